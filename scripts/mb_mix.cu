// Microbenchmark (B200): does a packed add.rn.f32x2 (FADD2, half instruction rate on the fma pipe) leave issue slots
// for alu-pipe (FMNMX) and integer (IADD) instructions, where scalar FADD at full rate does not?
// Per loop iteration and thread: 8 pair-adds (16 FADD or 8 FADD2) + NM FMNMX + NI integer adds, all independent chains.
// Prints scheduler cycles per iteration per warp (one CTA per SM, NT/128 warps per scheduler).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s:%d %s\n",__FILE__,__LINE__,cudaGetErrorString(e)); exit(1);} }while(0)

template <bool PACKED, int NM, int NI>
__global__ void k_mix(float* out, int iters, float x)
{
  float al[8], as[8], m[NM > 0 ? NM : 1];
  int c[NI > 0 ? NI : 1];
#pragma unroll
  for (int i = 0; i < 8; ++i) { al[i] = (float)i; as[i] = -(float)i; }
#pragma unroll
  for (int i = 0; i < NM; ++i) m[i] = x * (float)(i + 1);
#pragma unroll
  for (int i = 0; i < NI; ++i) c[i] = threadIdx.x + i;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (PACKED)
        asm volatile("{\n\t.reg .b64 a, b;\n\tmov.b64 a, {%0, %1};\n\tmov.b64 b, {%2, %3};\n\tadd.rn.f32x2 a, a, b;\n\tmov.b64 {%0, %1}, a;\n\t}"
                     : "+f"(al[i]), "+f"(as[i]) : "f"(x), "f"(-x));
      else {
        asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(al[i]) : "f"(x));
        asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(as[i]) : "f"(-x));
      }
      if (i < NM) asm volatile("max.f32 %0, %0, %1;" : "+f"(m[i]) : "f"(al[(i + 3) & 7]));
      if (i < NI) asm volatile("add.s32 %0, %0, %1;" : "+r"(c[i]) : "r"(it));
    }
    if (NM > 8) {
#pragma unroll
      for (int i = 8; i < NM; ++i) asm volatile("max.f32 %0, %0, %1;" : "+f"(m[i]) : "f"(al[i & 7]));
    }
    if (NI > 8) {
#pragma unroll
      for (int i = 8; i < NI; ++i) asm volatile("add.s32 %0, %0, %1;" : "+r"(c[i]) : "r"(it));
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += al[i] + as[i];
#pragma unroll
  for (int i = 0; i < NM; ++i) s += m[i];
#pragma unroll
  for (int i = 0; i < NI; ++i) s += (float)c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  const int nsm = pr.multiProcessorCount;
  printf("%s, %d SMs, max clock %.0f MHz; per iteration: 8 pair-adds + NM FMNMX + NI IADD\n", pr.name, nsm, clk_khz / 1e3);
  float* out; CK(cudaMalloc(&out, 4 * 1024 * nsm));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int IT = 20000;
  auto timeit = [&](const char* name, auto launch, int nt, int ninst) {
    launch(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); CK(cudaGetLastError());
    const double cyc = ms * 1e-3 * clk_khz * 1e3 / IT / (nt / 128.0);
    printf("%-40s NT=%4d  %7.3f ms  %6.2f sched-cycles / warp-iteration  (%d instructions: %.2f inst/clk/scheduler)\n",
           name, nt, ms, cyc, ninst, ninst / cyc);
  };
#define RUN(P, NM, NI) for (int nt : {256, 512}) { char nm[64]; snprintf(nm, sizeof nm, "%s + %d FMNMX + %d IADD", P ? "8 FADD2 " : "16 FADD ", NM, NI); \
    timeit(nm, [&] { k_mix<P, NM, NI><<<nsm, nt>>>(out, IT, 1.0f); }, nt, (P ? 8 : 16) + NM + NI); }
  RUN(false, 0, 0) RUN(true, 0, 0)
  RUN(false, 4, 0) RUN(true, 4, 0)
  RUN(false, 8, 0) RUN(true, 8, 0)
  RUN(false, 0, 8) RUN(true, 0, 8)
  RUN(false, 4, 4) RUN(true, 4, 4)
  RUN(false, 8, 8) RUN(true, 8, 8)
  RUN(true, 4, 12) RUN(true, 16, 0)
  return 0;
}
