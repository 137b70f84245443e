// Throughput of the special-function (MUFU) and FMA pipes on one SM, in cycles per warp instruction per SM
// sub-partition.  Grounds the attention softmax design (tanh + exp per logit).  Build: nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 4096
template <int OP>
__global__ void k(float* out, long long* cyc, float seed) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = seed + 0.001f * (threadIdx.x + i);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[i]));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (OP == 3) { uint32_t u = __float_as_uint(x[i]); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u)); x[i] = __uint_as_float(u); }
      if (OP == 4) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[i]));
      if (OP == 5) { uint32_t u = __float_as_uint(x[i]); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u)); x[i] = __uint_as_float(u); }
      if (OP == 6) { asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[i])); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[i])); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[i])); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[i])); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[i])); }
      if (OP == 7) { uint32_t u = __float_as_uint(x[i]); asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %0; tanh.approx.f16 lo, lo; mov.b32 %0, {lo, hi};}" : "+r"(u)); x[i] = __uint_as_float(u); }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int warps, int per_elem) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  k<OP><<<148, warps * 32>>>(out, cyc, 0.3f);
  k<OP><<<148, warps * 32>>>(out, cyc, 0.3f);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  double winstr_per_smsp = (double)ITER * 8 * per_elem * warps / 4.0;
  printf("%-28s warps/SM=%2d  cycles=%9.0f  cycles per warp-instr per SMSP = %.2f\n", name, warps, c, c / winstr_per_smsp);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {4, 8, 16}) {
    run<0>("MUFU.TANH f32", w, 1);
    run<1>("MUFU.EX2 f32", w, 1);
    run<2>("MUFU.RCP f32", w, 1);
    run<3>("tanh f16x2 (2 MUFU)", w, 2);
    run<7>("tanh f16 (1 MUFU)", w, 1);
    run<5>("ex2 f16x2 (2 MUFU)", w, 2);
    run<4>("FFMA", w, 1);
    run<6>("TANH + 4 FFMA (per 5 instr)", w, 5);
  }
  return 0;
}
