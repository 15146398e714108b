// Microbenchmark: how fast can a B200 stream strided pressure columns (nz planes of n2 floats)
// for a fragmented subset of cells?  Variants of the k_scan_quad access pattern.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstdint>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s:%d %s\n",__FILE__,__LINE__,cudaGetErrorString(e)); exit(1);} }while(0)

template <int G, int NT>
__global__ void __launch_bounds__(NT) scan_cpasync(const float* __restrict__ p, const int* __restrict__ quads, int nq, size_t n2, int nz, float tp, int* __restrict__ out)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* s_buf = reinterpret_cast<float4*>(smem_raw);
  const int tid = threadIdx.x; const int i = blockIdx.x * NT + tid; if (i >= nq) return;
  const int m = quads[i]; const size_t stride4 = n2 >> 2;
  const float4* pq = reinterpret_cast<const float4*>(p) + m;
  const int ngroups = (nz + G - 1) / G;
  auto issue = [&](int g) {
    float4* dst = s_buf + (size_t)((g & 1) * G) * NT + tid; const int kbase = g * G;
#pragma unroll
    for (int j = 0; j < G; ++j) if (kbase + j < nz) {
      const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + (size_t)j * NT);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(pq + (size_t)(kbase + j) * stride4) : "memory"); }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  issue(0);
  float b0=0,b1=0,b2=0,b3=0; int l0=0,l1=0,l2=0,l3=0;
  for (int g = 0; g < ngroups; ++g) {
    if (g + 1 < ngroups) { issue(g + 1); asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }
    else asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    const float4* src = s_buf + (size_t)((g & 1) * G) * NT + tid; const int kbase = g * G;
#pragma unroll
    for (int j = 0; j < G; ++j) { const int kk = kbase + j; if (kk < nz) {
        const float4 d = src[(size_t)j * NT];
        const float e0=fabsf(d.x-tp),e1=fabsf(d.y-tp),e2=fabsf(d.z-tp),e3=fabsf(d.w-tp);
        if (kk==0||e0<b0){b0=e0;l0=kk;} if (kk==0||e1<b1){b1=e1;l1=kk;} if (kk==0||e2<b2){b2=e2;l2=kk;} if (kk==0||e3<b3){b3=e3;l3=kk;} } }
  }
  out[i] = l0 + l1 + l2 + l3;
}

// register version: B float4 loads in flight, explicit two-phase via volatile asm loads
template <int B, int NT>
__global__ void __launch_bounds__(NT) scan_regs(const float* __restrict__ p, const int* __restrict__ quads, int nq, size_t n2, int nz, float tp, int* __restrict__ out)
{
  const int i = blockIdx.x * NT + threadIdx.x; if (i >= nq) return;
  const int m = quads[i]; const size_t stride4 = n2 >> 2;
  const float4* pq = reinterpret_cast<const float4*>(p) + m;
  float b0=3e38f,b1=3e38f,b2=3e38f,b3=3e38f; int l0=0,l1=0,l2=0,l3=0;
  for (int k0 = 0; k0 < nz; k0 += B) {
    float4 d[B];
#pragma unroll
    for (int j = 0; j < B; ++j) { const int kk = min(k0 + j, nz - 1);
      asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(d[j].x),"=f"(d[j].y),"=f"(d[j].z),"=f"(d[j].w) : "l"(pq + (size_t)kk * stride4)); }
#pragma unroll
    for (int j = 0; j < B; ++j) { const int kk = k0 + j;
        const float e0=fabsf(d[j].x-tp),e1=fabsf(d[j].y-tp),e2=fabsf(d[j].z-tp),e3=fabsf(d[j].w-tp);
        if (kk<nz&&e0<b0){b0=e0;l0=kk;} if (kk<nz&&e1<b1){b1=e1;l1=kk;} if (kk<nz&&e2<b2){b2=e2;l2=kk;} if (kk<nz&&e3<b3){b3=e3;l3=kk;} }
  }
  out[i] = l0 + l1 + l2 + l3;
}

// level-major: a CTA owns QPT*NT consecutive list entries; per level all its threads read QPT quads
template <int QPT, int NT>
__global__ void __launch_bounds__(NT) scan_levelmajor(const float* __restrict__ p, const int* __restrict__ quads, int nq, size_t n2, int nz, float tp, int* __restrict__ out)
{
  const int base = blockIdx.x * NT * QPT + threadIdx.x;
  const size_t stride4 = n2 >> 2;
  const float4* pq[QPT]; float b[QPT][4]; int l[QPT];
#pragma unroll
  for (int q = 0; q < QPT; ++q) { const int i = min(base + q * NT, nq - 1); pq[q] = reinterpret_cast<const float4*>(p) + quads[i]; l[q] = 0; for (int c=0;c<4;++c) b[q][c]=3e38f; }
#pragma unroll 2
  for (int k = 0; k < nz; ++k) {
    float4 d[QPT];
#pragma unroll
    for (int q = 0; q < QPT; ++q) d[q] = __ldcs(pq[q] + (size_t)k * stride4);
#pragma unroll
    for (int q = 0; q < QPT; ++q) { const float e0=fabsf(d[q].x-tp),e1=fabsf(d[q].y-tp),e2=fabsf(d[q].z-tp),e3=fabsf(d[q].w-tp);
      if (e0<b[q][0]){b[q][0]=e0;l[q]+=k;} if (e1<b[q][1]){b[q][1]=e1;l[q]+=k;} if (e2<b[q][2]){b[q][2]=e2;l[q]+=k;} if (e3<b[q][3]){b[q][3]=e3;l[q]+=k;} }
  }
#pragma unroll
  for (int q = 0; q < QPT; ++q) { const int i = base + q * NT; if (i < nq) out[i] = l[q]; }
}

// level split: S threads per quad, each scanning nz/S levels (B loads in flight each), combined by shuffle
template <int B, int S, int NT>
__global__ void __launch_bounds__(NT) scan_split(const float* __restrict__ p, const int* __restrict__ quads, int nq, size_t n2, int nz, float tp, int* __restrict__ out)
{
  // lanes: within a warp, lanes [0, 32/S) take part 0, the next 32/S part 1, ...: every part reads 32/S consecutive quads
  const int QW = 32 / S;
  const int warp = (blockIdx.x * NT + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int part = lane / QW, ql = lane % QW;
  const int i = min(warp * QW + ql, nq - 1);
  const int m = quads[i]; const size_t stride4 = n2 >> 2;
  const float4* pq = reinterpret_cast<const float4*>(p) + m;
  const int per = (nz + S - 1) / S, ka = part * per, kb = min(nz, ka + per);
  float b0=3e38f,b1=3e38f,b2=3e38f,b3=3e38f; int l0=0,l1=0,l2=0,l3=0;
  for (int k0 = ka; k0 < kb; k0 += B) {
    float4 d[B];
#pragma unroll
    for (int j = 0; j < B; ++j) { const int kk = min(k0 + j, kb - 1);
      asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(d[j].x),"=f"(d[j].y),"=f"(d[j].z),"=f"(d[j].w) : "l"(pq + (size_t)kk * stride4)); }
#pragma unroll
    for (int j = 0; j < B; ++j) { const int kk = k0 + j;
        const float e0=fabsf(d[j].x-tp),e1=fabsf(d[j].y-tp),e2=fabsf(d[j].z-tp),e3=fabsf(d[j].w-tp);
        if (kk<kb&&e0<b0){b0=e0;l0=kk;} if (kk<kb&&e1<b1){b1=e1;l1=kk;} if (kk<kb&&e2<b2){b2=e2;l2=kk;} if (kk<kb&&e3<b3){b3=e3;l3=kk;} }
  }
  // combine: strict < keeps the earlier part on ties (first minimum)
#pragma unroll
  for (int s = 1; s < S; s <<= 1) {
    const int src = lane + s * QW;
    float o; int ol;
    o = __shfl_sync(0xffffffffu, b0, src); ol = __shfl_sync(0xffffffffu, l0, src); if (o < b0) { b0 = o; l0 = ol; }
    o = __shfl_sync(0xffffffffu, b1, src); ol = __shfl_sync(0xffffffffu, l1, src); if (o < b1) { b1 = o; l1 = ol; }
    o = __shfl_sync(0xffffffffu, b2, src); ol = __shfl_sync(0xffffffffu, l2, src); if (o < b2) { b2 = o; l2 = ol; }
    o = __shfl_sync(0xffffffffu, b3, src); ol = __shfl_sync(0xffffffffu, l3, src); if (o < b3) { b3 = o; l3 = ol; }
  }
  if (part == 0 && warp * QW + ql < nq) out[i] = l0 + l1 + l2 + l3;
}

__global__ void copy_k(const float4* __restrict__ a, float4* __restrict__ b, size_t n) { size_t i = (size_t)blockIdx.x*blockDim.x+threadIdx.x; size_t st=(size_t)gridDim.x*blockDim.x; for(;i<n;i+=st) b[i]=a[i]; }
__global__ void read_k(const float4* __restrict__ a, float* __restrict__ o, size_t n) { size_t i = (size_t)blockIdx.x*blockDim.x+threadIdx.x; size_t st=(size_t)gridDim.x*blockDim.x; float s=0; for(;i<n;i+=st){float4 v=__ldcs(a+i); s+=v.x+v.y+v.z+v.w;} if (s==123.456f) o[0]=s; }

int main(int argc, char** argv)
{
  const int nx = 1024, ny = 769, nz = 70; const size_t n2 = (size_t)nx * ny;
  const int mean_run = argc > 1 ? atoi(argv[1]) : 16;     // mean active run length (cells)
  const double frac = argc > 2 ? atof(argv[2]) : 0.44;    // active fraction
  const int gran = argc > 3 ? atoi(argv[3]) : 0;
  if (gran) { CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran)); }
  { size_t g = 0; CK(cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity)); printf("L2 fetch granularity limit = %zu\n", g); }
  const int NSETS = 3;
  float* p[NSETS]; for (int s=0;s<NSETS;++s){ CK(cudaMalloc(&p[s], n2*nz*4)); CK(cudaMemset(p[s], 0, n2*nz*4)); }
  // fragmented active mask: alternating runs
  std::vector<uint8_t> act(n2, 0); srand(1);
  { size_t c = 0; bool on = false; while (c < n2) { double mr = on ? mean_run : mean_run * (1 - frac) / frac; int len = 1 + rand() % (int)(2 * mr); for (int j=0;j<len && c<n2;++j) act[c++] = on; on = !on; } }
  std::vector<int> quads; size_t nact=0; std::vector<uint8_t> g64(n2/16+1,0);
  for (size_t m = 0; m < n2/4; ++m) { bool any=false; for (int j=0;j<4;++j) { if (act[m*4+j]) {any=true; nact++; g64[(m*4+j)/16]=1;} } if (any) quads.push_back((int)m); }
  size_t n64=0; for (auto v: g64) n64+=v;
  std::vector<int> dense(n2/4); for (size_t m=0;m<n2/4;++m) dense[m]=(int)m;
  printf("active cells %zu (%.2f), quads %zu, 64B granules %zu -> useful %.1f MB, quad bytes %.1f MB, 64B-granule bytes %.1f MB\n", nact, (double)nact/n2, quads.size(), n64, nact*4.0*nz/1e6, quads.size()*16.0*nz/1e6, n64*64.0*nz/1e6);
  int *dq, *dd, *out; CK(cudaMalloc(&dq, quads.size()*4)); CK(cudaMalloc(&dd, dense.size()*4)); CK(cudaMalloc(&out, n2));
  CK(cudaMemcpy(dq, quads.data(), quads.size()*4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dd, dense.data(), dense.size()*4, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto timeit = [&](const char* name, auto launch, double bytes) {
    for (int w = 0; w < 3; ++w) launch(p[w % NSETS]);
    CK(cudaDeviceSynchronize());
    const int reps = 12; CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) launch(p[r % NSETS]);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaGetLastError());
    printf("%-44s %8.2f us  %7.1f GB/s (by bytes given)\n", name, ms/reps*1e3, bytes/(ms/reps*1e-3)/1e9);
  };
  { float4* tmp; CK(cudaMalloc(&tmp, n2*nz*4));
    timeit("copy (read+write) full p", [&](float* pp){ copy_k<<<148*16,256>>>((float4*)pp, tmp, n2*nz/4); }, 2.0*n2*nz*4);
    timeit("read full p", [&](float* pp){ read_k<<<148*16,256>>>((float4*)pp, (float*)out, n2*nz/4); }, 1.0*n2*nz*4);
    cudaFree(tmp); }
  const int nq = (int)quads.size(); const int nd = (int)dense.size(); const double bq = nq*16.0*nz, bd = nd*16.0*nz;
#define RUN_CP(G,NT,list,n,bytes,tag) { int sm = 2*G*NT*16; CK(cudaFuncSetAttribute(scan_cpasync<G,NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm)); \
    timeit(tag " cpasync G=" #G " NT=" #NT, [&](float* pp){ scan_cpasync<G,NT><<<(n+NT-1)/NT,NT,sm>>>(pp,list,n,n2,nz,70000.f,out); }, bytes); }
  RUN_CP(14,128,dd,nd,bd,"dense");
  RUN_CP(14,128,dq,nq,bq,"frag "); RUN_CP(3,256,dq,nq,bq,"frag ");
#define RUN_RG(B,NT,list,n,bytes,tag) timeit(tag " regs B=" #B " NT=" #NT, [&](float* pp){ scan_regs<B,NT><<<(n+NT-1)/NT,NT>>>(pp,list,n,n2,nz,70000.f,out); }, bytes);
  RUN_RG(10,128,dd,nd,bd,"dense"); RUN_RG(10,128,dq,nq,bq,"frag "); RUN_RG(14,128,dq,nq,bq,"frag ");
#define RUN_LM(Q,NT,list,n,bytes,tag) timeit(tag " levelmajor QPT=" #Q " NT=" #NT, [&](float* pp){ scan_levelmajor<Q,NT><<<(n+NT*Q-1)/(NT*Q),NT>>>(pp,list,n,n2,nz,70000.f,out); }, bytes);
  RUN_LM(1,256,dq,nq,bq,"frag ");
  RUN_RG(5,128,dq,nq,bq,"frag "); RUN_RG(7,128,dq,nq,bq,"frag "); RUN_RG(7,64,dq,nq,bq,"frag "); RUN_RG(7,256,dq,nq,bq,"frag "); RUN_RG(14,256,dq,nq,bq,"frag "); RUN_RG(18,128,dq,nq,bq,"frag ");
#define RUN_SP(B,S,NT,list,n,bytes,tag) timeit(tag " split B=" #B " S=" #S " NT=" #NT, [&](float* pp){ scan_split<B,S,NT><<<((n+(32/S)-1)/(32/S)*32+NT-1)/NT,NT>>>(pp,list,n,n2,nz,70000.f,out); }, bytes);
  RUN_SP(5,2,128,dq,nq,bq,"frag "); RUN_SP(9,2,128,dq,nq,bq,"frag "); RUN_SP(6,4,128,dq,nq,bq,"frag "); RUN_SP(9,4,128,dq,nq,bq,"frag "); RUN_SP(9,2,256,dq,nq,bq,"frag "); RUN_SP(9,8,128,dq,nq,bq,"frag ");
  // coarser fetch granules: every quad of each 64/128/256/512-byte block that holds an active cell
  for (int gcells : {16, 32, 64, 128}) {
    std::vector<int> ql; std::vector<uint8_t> blk(n2 / gcells + 1, 0);
    for (size_t c = 0; c < n2; ++c) if (act[c]) blk[c / gcells] = 1;
    for (size_t m = 0; m < n2 / 4; ++m) if (blk[(m * 4) / gcells]) ql.push_back((int)m);
    int* dl; CK(cudaMalloc(&dl, ql.size() * 4)); CK(cudaMemcpy(dl, ql.data(), ql.size() * 4, cudaMemcpyHostToDevice));
    const int nl = (int)ql.size(); const double bl = nl * 16.0 * nz;
    char tag[64]; snprintf(tag, sizeof tag, "gran %4d B (%.1f MB) regs B=10", gcells * 4, bl / 1e6);
    timeit(tag, [&](float* pp){ scan_regs<10,128><<<(nl+127)/128,128>>>(pp,dl,nl,n2,nz,70000.f,out); }, bl);
    snprintf(tag, sizeof tag, "gran %4d B (%.1f MB) split B=9 S=2", gcells * 4, bl / 1e6);
    timeit(tag, [&](float* pp){ scan_split<9,2,128><<<((nl+15)/16*32+127)/128,128>>>(pp,dl,nl,n2,nz,70000.f,out); }, bl);
    cudaFree(dl);
  }
  return 0;
}
