// BatchRMSNorm with UPDATING statistics (reference alphagenome_pytorch/alphagenome.py, "AG": BatchRMSNorm.forward
// AG:324-361 with update_running_var — the reference's default, eval included — and get_maybe_dist_var AG:276-300).
//
// The frozen path folds the norm into the producer's GEMM epilogue.  When a norm re-estimates its variance the
// consumer's scale depends on ALL rows of the tensor being normalised, so that norm runs un-fused in three HBM-bound
// steps (all coalesced, 128-bit accesses):
//   1. bn_stats_kernel    per-channel sum and sum of squares of the fp32 tensor, accumulated in fp64
//                         (one pass; a sequence-parallel rank then adds the other ranks' [2][C] with ONE all-reduce,
//                         where the reference needs three, AG:289-298)
//   2. bn_update_kernel   var = E[x^2] - E[x]^2 (fp64), running_var.lerp_(var, momentum) IN the module's buffer,
//                         and the packed per-channel scale gamma / sqrt(running_var + eps) (+ optional folded
//                         shift bias * scale + beta for the sandwich post-norms) the fused kernels read
//   3. affine_act_kernel  y = act(x * scale[c] + shift[b, c]) (+ residual) -> fp32 and / or bf16
#include "common.cuh"
#include "kernels.h"

namespace agb {

constexpr int kBnTy = 8;          // row lanes per block
constexpr int kBnFlush = 8;       // fp32 partial sums are folded into fp64 every kBnFlush rows per thread

__global__ void __launch_bounds__(32 * kBnTy) bn_stats_kernel(const float* __restrict__ x, double* __restrict__ sums, long total_rows,
                                                               int rows_per_b, long bstride, int ld, int C, int rows_per_block) {
  // block = 128 channels (32 lanes x float4) x kBnTy row lanes; grid.x = C / 128, grid.y = row chunks
  const int c = (blockIdx.x * 32 + threadIdx.x) * 4;
  const long r0 = (long)blockIdx.y * rows_per_block;
  long r1 = r0 + rows_per_block;
  if (r1 > total_rows) r1 = total_rows;
  double ds[4] = {0, 0, 0, 0}, dq[4] = {0, 0, 0, 0};
  float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  int n = 0;
  for (long r = r0 + threadIdx.y; r < r1; r += kBnTy) {
    const long b = r / rows_per_b;
    const long rr = r - b * rows_per_b;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + b * bstride + rr * ld + c));
    s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
    q[0] = fmaf(v.x, v.x, q[0]); q[1] = fmaf(v.y, v.y, q[1]); q[2] = fmaf(v.z, v.z, q[2]); q[3] = fmaf(v.w, v.w, q[3]);
    if (++n == kBnFlush) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { ds[i] += (double)s[i]; dq[i] += (double)q[i]; s[i] = 0.f; q[i] = 0.f; }
      n = 0;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) { ds[i] += (double)s[i]; dq[i] += (double)q[i]; }
  __shared__ double sh[kBnTy][32][8];
#pragma unroll
  for (int i = 0; i < 4; ++i) { sh[threadIdx.y][threadIdx.x][i] = ds[i]; sh[threadIdx.y][threadIdx.x][4 + i] = dq[i]; }
  __syncthreads();
  if (threadIdx.y == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double a = 0;
      for (int y = 0; y < kBnTy; ++y) a += sh[y][threadIdx.x][i];
      atomicAdd(sums + (i < 4 ? 0 : C) + c + (i & 3), a);
    }
  }
}

__global__ void bn_update_kernel(const double* __restrict__ sums, double count, float* __restrict__ running_var,
                                 const float* __restrict__ gamma, float momentum, float eps, float* __restrict__ scale_out,
                                 const float* __restrict__ fold_bias, const float* __restrict__ fold_beta,
                                 float* __restrict__ fold_shift_out, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = sums[c] / count;
  double var = sums[C + c] / count - mean * mean;          // population variance (unbiased = False), AG:284 / 300
  if (var < 0) var = 0;
  const float rv0 = running_var[c];
  const float rv = rv0 + momentum * ((float)var - rv0);      // running_var.lerp_(batch_var, 1 - momentum_arg), AG:340
  running_var[c] = rv;
  const float sc = gamma[c] / sqrtf(rv + eps);               // x / sqrt(running_var + eps) * gamma, AG:344-359
  scale_out[c] = sc;
  if (fold_shift_out) fold_shift_out[c] = (fold_bias ? fold_bias[c] : 0.f) * sc + fold_beta[c];
}

__device__ __forceinline__ uint32_t bn_pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void affine_act_kernel(const float* __restrict__ x, long x_bs, int x_ld, const float* __restrict__ scale,
                                  const float* __restrict__ shift, long shift_bs, const float* __restrict__ res, long res_bs,
                                  int res_ld, float* __restrict__ out32, long o32_bs, int o32_ld, __nv_bfloat16* __restrict__ out16,
                                  long o16_bs, int o16_ld, int act, int rows_per_b, int C4, long total) {
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C4) * 4;
    const long row = idx / C4;
    const long b = row / rows_per_b;
    const long r = row - b * rows_per_b;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + b * x_bs + r * x_ld + c));
    const float4 s = scale ? __ldg(reinterpret_cast<const float4*>(scale + c)) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 t = shift ? __ldg(reinterpret_cast<const float4*>(shift + b * shift_bs + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 y;
    y.x = fmaf(v.x, s.x, t.x); y.y = fmaf(v.y, s.y, t.y); y.z = fmaf(v.z, s.z, t.z); y.w = fmaf(v.w, s.w, t.w);
    if (res) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(res + b * res_bs + r * res_ld + c));
      y.x += a.x; y.y += a.y; y.z += a.z; y.w += a.w;
    }
    if (out32) {
      float4 o = y;
      if (act != ACT_NONE) { o.x = act_apply(y.x, act); o.y = act_apply(y.y, act); o.z = act_apply(y.z, act); o.w = act_apply(y.w, act); }
      *reinterpret_cast<float4*>(out32 + b * o32_bs + r * o32_ld + c) = o;
    }
    if (out16) {
      float4 o = y;   // the same one-MUFU activation as the GEMM epilogue's bf16 outputs
      if (act != ACT_NONE) { o.x = act_apply_fast(y.x, act); o.y = act_apply_fast(y.y, act); o.z = act_apply_fast(y.z, act); o.w = act_apply_fast(y.w, act); }
      uint2 pk;
      pk.x = bn_pack2(o.x, o.y);
      pk.y = bn_pack2(o.z, o.w);
      *reinterpret_cast<uint2*>(out16 + b * o16_bs + r * o16_ld + c) = pk;
    }
  }
}

int launch_bn_stats(const float* x, double* sums, int B, int rows, long bstride, int ld, int C, int device, cudaStream_t st) {
  if (!x || !sums) return set_error("ag_bn_stats_fwd: null pointer");
  if (C < 128 || C % 128) return set_error("ag_bn_stats_fwd: C=%d must be a multiple of 128", C);
  if (B < 1 || rows < 1 || (ld & 3) || (bstride & 3) || (reinterpret_cast<uintptr_t>(x) & 15))
    return set_error("ag_bn_stats_fwd: bad shape / alignment (B=%d rows=%d ld=%d)", B, rows, ld);
  cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st);
  if (e != cudaSuccess) return set_error("ag_bn_stats_fwd: memset: %s", cudaGetErrorString(e));
  const long total = (long)B * rows;
  const int gx = C / 128;
  long chunks = (long)sm_count(device) * 8 / gx;
  if (chunks < 1) chunks = 1;
  long per = (total + chunks - 1) / chunks;
  if (per < kBnTy * 4) per = kBnTy * 4;
  per = (per + kBnTy - 1) / kBnTy * kBnTy;
  chunks = (total + per - 1) / per;
  if (chunks > 65535) return set_error("ag_bn_stats_fwd: too many row chunks");
  bn_stats_kernel<<<dim3(gx, (unsigned)chunks), dim3(32, kBnTy), 0, st>>>(x, sums, total, rows, bstride, ld, C, (int)per);
  return check_launch("ag_bn_stats_fwd");
}

int launch_bn_update(const double* sums, double count, float* running_var, const float* gamma, float momentum, float eps,
                     float* scale_out, const float* fold_bias, const float* fold_beta, float* fold_shift_out, int C, int device,
                     cudaStream_t st) {
  (void)device;
  if (!sums || !running_var || !gamma || !scale_out) return set_error("ag_bn_update_fwd: null pointer");
  if (!(count >= 1.0)) return set_error("ag_bn_update_fwd: count must be >= 1");
  if (fold_shift_out && !fold_beta) return set_error("ag_bn_update_fwd: the folded shift needs beta");
  bn_update_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, count, running_var, gamma, momentum, eps, scale_out, fold_bias, fold_beta,
                                                   fold_shift_out, C);
  return check_launch("ag_bn_update_fwd");
}

int launch_affine_act(const AgAffineArgs& a, int device, cudaStream_t st) {
  if (!a.x || (!a.out_f32 && !a.out_bf16)) return set_error("ag_affine_act_fwd: x and at least one output are required");
  if (a.C < 4 || (a.C & 3) || a.B < 1 || a.rows < 1) return set_error("ag_affine_act_fwd: bad sizes B=%d rows=%d C=%d", a.B, a.rows, a.C);
  if ((a.x_ld | a.res_ld | a.out_f32_ld | a.out_bf16_ld) & 3 || (a.x_bstride | a.res_bstride | a.out_f32_bstride | a.out_bf16_bstride | a.shift_bstride) & 3)
    return set_error("ag_affine_act_fwd: strides must be multiples of 4 elements");
  if ((reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.res) | reinterpret_cast<uintptr_t>(a.out_f32) |
       reinterpret_cast<uintptr_t>(a.scale) | reinterpret_cast<uintptr_t>(a.shift)) & 15 || (reinterpret_cast<uintptr_t>(a.out_bf16) & 7))
    return set_error("ag_affine_act_fwd: pointers must be 16-byte (bf16: 8-byte) aligned");
  if (a.act != AG_ACT_NONE && a.act != AG_ACT_GELU1702 && a.act != AG_ACT_GELU_ERF && a.act != AG_ACT_RELU)
    return set_error("ag_affine_act_fwd: unsupported activation %d", a.act);
  const long total = (long)a.B * a.rows * (a.C / 4);
  long blocks = (total + 255) / 256;
  const long cap = (long)sm_count(device) * 16;
  if (blocks > cap) blocks = cap;
  affine_act_kernel<<<(unsigned)blocks, 256, 0, st>>>(a.x, a.x_bstride, a.x_ld, a.scale, a.shift, a.shift_bstride, a.res, a.res_bstride,
                                                      a.res_ld, a.out_f32, a.out_f32_bstride, a.out_f32_ld,
                                                      reinterpret_cast<__nv_bfloat16*>(a.out_bf16), a.out_bf16_bstride, a.out_bf16_ld, a.act,
                                                      a.rows, a.C / 4, total);
  return check_launch("ag_affine_act_fwd");
}

}  // namespace agb
