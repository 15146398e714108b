// Microbenchmark (B200): issue rates of what the window-sum kernels are made of.
//   FADD (independent chains), add.rn.f32x2 (FADD2), predicated FADD pairs, LDS.64 / LDS.128 feeding adds.
// Prints lane-adds per clock per SM for each variant; one CTA per SM, NW warps.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s:%d %s\n",__FILE__,__LINE__,cudaGetErrorString(e)); exit(1);} }while(0)

__device__ __forceinline__ unsigned long long add2(unsigned long long acc, unsigned long long v)
{
  asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(acc) : "l"(v));
  return acc;
}

template <int ILP>
__global__ void k_fadd(float* out, int iters, float x)
{
  float a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = (float)i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(x));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_fadd2(float* out, int iters, float x)
{
  unsigned long long a[ILP];
  unsigned long long v;
  asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x), "f"(0.0f));
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = 0ull;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = add2(a[i], v);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[i])); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// predicated pair: @p add Tl ; @!p add Ts  (one useful add per two issue slots)
template <int ILP>
__global__ void k_fadd_pred(float* out, int iters, float x, int pat)
{
  float l[ILP], s[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { l[i] = 0; s[i] = 0; }
  for (int it = 0; it < iters; ++it) {
    const int pr = (pat >> (it & 31)) & 1;
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t@q add.rn.f32 %0, %0, %2;\n\t@!q add.rn.f32 %1, %1, %2;\n\t}"
                   : "+f"(l[i]), "+f"(s[i]) : "f"(x), "r"(pr));
  }
  float t = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) t += l[i] + s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

// LDS.64 feeding CW FADD2 (strip of CW cells), conflict-free addresses (lane-consecutive pairs)
template <int CW, int VEC>
__global__ void k_lds_fadd2(float* out, int iters, int pitch)
{
  extern __shared__ __align__(16) unsigned long long sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 0ull;
  __syncthreads();
  unsigned long long a[CW];
#pragma unroll
  for (int i = 0; i < CW; ++i) a[i] = 0ull;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long* base = sm + (((warp * 37 + lane * pitch) * VEC) & 4095);
  for (int it = 0; it < iters; ++it) {
    const unsigned long long* rp = base + ((it * VEC) & 2047);
    if (VEC == 1) {
      const unsigned long long v = *reinterpret_cast<const volatile unsigned long long*>(rp);
#pragma unroll
      for (int i = 0; i < CW; ++i) a[i] = add2(a[i], v);
    } else {
      ulonglong2 v;
      asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"((unsigned)__cvta_generic_to_shared(rp)));
#pragma unroll
      for (int i = 0; i < CW; ++i) a[i] = add2(a[i], v.x);
#pragma unroll
      for (int i = 0; i < CW; ++i) a[i] = add2(a[i], v.y);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CW; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[i])); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
  int dev = 0; cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, dev));
  int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev));
  const int nsm = pr.multiProcessorCount;
  printf("%s, %d SMs, max clock %.0f MHz\n", pr.name, nsm, clk_khz / 1e3);
  float* out; CK(cudaMalloc(&out, 4 * 1024 * nsm * 2));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto timeit = [&](const char* name, auto launch, double lane_ops_per_thread, int threads) {
    launch(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); CK(cudaGetLastError());
    const double ops = lane_ops_per_thread * threads * nsm;
    printf("%-52s %8.3f ms  %7.1f Gop/s  = %6.1f lane-ops/clk/SM @%.0f MHz\n", name, ms, ops / ms / 1e6,
           ops / (ms * 1e-3) / nsm / (clk_khz * 1e3), clk_khz / 1e3);
  };
  const int IT = 20000;
  for (int nt : {128, 256, 512, 1024}) {
    char nm[96];
    snprintf(nm, sizeof nm, "FADD ILP8 NT=%d (adds)", nt);
    timeit(nm, [&] { k_fadd<8><<<nsm, nt>>>(out, IT, 1.0f); }, 8.0 * IT, nt);
    snprintf(nm, sizeof nm, "FADD2 ILP8 NT=%d (pair-adds; x2 for float adds)", nt);
    timeit(nm, [&] { k_fadd2<8><<<nsm, nt>>>(out, IT, 1.0f); }, 8.0 * IT, nt);
    snprintf(nm, sizeof nm, "pred FADD pair ILP8 NT=%d (useful adds)", nt);
    timeit(nm, [&] { k_fadd_pred<8><<<nsm, nt>>>(out, IT, 1.0f, 0x5a5a5a5a); }, 8.0 * IT, nt);
  }
  CK(cudaFuncSetAttribute(k_lds_fadd2<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(k_lds_fadd2<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(k_lds_fadd2<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(k_lds_fadd2<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  for (int nt : {256, 512, 1024}) {
    for (int pitch : {1, 2, 17, 16}) {
      char nm[96];
      snprintf(nm, sizeof nm, "LDS.64 + 4 FADD2, NT=%d lane pitch %d pairs", nt, pitch);
      timeit(nm, [&] { k_lds_fadd2<4, 1><<<nsm, nt, 65536>>>(out, IT, pitch); }, 4.0 * IT, nt);
      snprintf(nm, sizeof nm, "LDS.64 + 8 FADD2, NT=%d lane pitch %d pairs", nt, pitch);
      timeit(nm, [&] { k_lds_fadd2<8, 1><<<nsm, nt, 65536>>>(out, IT, pitch); }, 8.0 * IT, nt);
      snprintf(nm, sizeof nm, "LDS.128 + 2x8 FADD2, NT=%d lane pitch %d pairs", nt, pitch);
      timeit(nm, [&] { k_lds_fadd2<8, 2><<<nsm, nt, 65536>>>(out, IT, pitch); }, 16.0 * IT, nt);
      snprintf(nm, sizeof nm, "LDS.128 + 2x4 FADD2, NT=%d lane pitch %d pairs", nt, pitch);
      timeit(nm, [&] { k_lds_fadd2<4, 2><<<nsm, nt, 65536>>>(out, IT, pitch); }, 8.0 * IT, nt);
    }
  }
  return 0;
}
