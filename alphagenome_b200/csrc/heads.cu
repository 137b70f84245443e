// Small kernels of the prediction heads (reference alphagenome_pytorch/alphagenome.py "AG"):
//   * ContactMapHead's RMSNormWithBias + organism embedding + GELU1702 (AG:1292-1298) -> bf16 operand of its Linear;
//   * SpliceJunctionHead.tissue_scaled_rope (AG:1338-1361): per-context scale / offset and rotary embedding of the
//     gathered, projected splice-site rows -> bf16 operands of the donor x acceptor contraction.
// The heads' contractions themselves (track / splice Linear layers with softplus / sigmoid epilogues, the contact
// projection, the donor x acceptor einsum) run on ag_gemm_fwd.
#include "common.cuh"
#include "kernels.h"

namespace agb {

__device__ __forceinline__ uint32_t hpack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// out[row][c] = act(x[row][c] * rsqrt(mean_c x^2 + eps) * w[c] + b[c] + add[row / rows_per_batch][c]); one warp per row
__global__ void rmsnorm_act_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bvec,
                                   const float* __restrict__ add, __nv_bfloat16* __restrict__ out, long rows,
                                   long rows_per_batch, int C, int act) {
  const int lane = threadIdx.x & 31;
  const int nv = C / 128;
  for (long row = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5; row < rows; row += ((long)gridDim.x * blockDim.x) >> 5) {
    const float4* src = reinterpret_cast<const float4*>(x + row * C);
    const float* addrow = add ? add + (row / rows_per_batch) * C : nullptr;
    float ss = 0.f;
    for (int k = 0; k < nv; ++k) {
      const float4 v = __ldg(src + k * 32 + lane);
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    const float inv = rsqrtf(warp_sum(ss) / (float)C + 1e-5f);
    for (int k = 0; k < nv; ++k) {
      const int c4 = k * 32 + lane;
      const float4 v = __ldg(src + c4);
      const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + c4);
      float4 bb = __ldg(reinterpret_cast<const float4*>(bvec) + c4);
      if (addrow) {
        const float4 e = __ldg(reinterpret_cast<const float4*>(addrow) + c4);
        bb.x += e.x; bb.y += e.y; bb.z += e.z; bb.w += e.w;
      }
      uint2 pk;
      pk.x = hpack2(act_apply(fmaf(v.x * inv, ww.x, bb.x), act), act_apply(fmaf(v.y * inv, ww.y, bb.y), act));
      pk.y = hpack2(act_apply(fmaf(v.z * inv, ww.z, bb.z), act), act_apply(fmaf(v.w * inv, ww.w, bb.w), act));
      *reinterpret_cast<uint2*>(out + row * C + c4 * 4) = pk;
    }
  }
}

// out[b][t][p][:] = rope_t(x[b][p][:] * scale[t][:] + offset[t][:]), interleaved pairs (x0, x1) -> (x0 c - x1 s, x1 c + x0 s)
// with the context index t as the rotary position (AG:1352-1359).  One thread per 4 channels (two rotary pairs).
__global__ void splice_rope_kernel(const float* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ offset,
                                   const float* __restrict__ cos_tab, const float* __restrict__ sin_tab,
                                   __nv_bfloat16* __restrict__ out, int B, int P, int T, int Hd, int p_ld, int tab_per_site) {
  const int q4 = Hd / 4;
  const long total = (long)B * T * P * q4;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c4 = (int)(idx % q4);
    long r = idx / q4;
    const int p = (int)(r % P);
    r /= P;
    const int t = (int)(r % T);
    const long b = r / T;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + (b * P + p) * (long)Hd) + c4);
    const float4 s = __ldg(reinterpret_cast<const float4*>(scale + (long)t * Hd) + c4);
    const float4 o = __ldg(reinterpret_cast<const float4*>(offset + (long)t * Hd) + c4);
    // rotary position: the context index t (SpliceJunctionHead, AG:1356) or the site itself (SpliceSitesJunctionHead, AG:1495)
    const long trow = tab_per_site ? b * P + p : (long)t;
    const float2 c = __ldg(reinterpret_cast<const float2*>(cos_tab + trow * (Hd / 2)) + c4);
    const float2 sn = __ldg(reinterpret_cast<const float2*>(sin_tab + trow * (Hd / 2)) + c4);
    const float u0 = fmaf(v.x, s.x, o.x), u1 = fmaf(v.y, s.y, o.y), u2 = fmaf(v.z, s.z, o.z), u3 = fmaf(v.w, s.w, o.w);
    uint2 pk;
    pk.x = hpack2(u0 * c.x - u1 * sn.x, u1 * c.x + u0 * sn.x);
    pk.y = hpack2(u2 * c.y - u3 * sn.y, u3 * c.y + u2 * sn.y);
    *reinterpret_cast<uint2*>(out + ((b * T + t) * (long)p_ld + p) * Hd + c4 * 4) = pk;
  }
}

static inline unsigned head_grid(long work, int per_block, int device) {
  long blocks = (work + per_block - 1) / per_block;
  const long cap = (long)sm_count(device) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

int launch_rmsnorm_act(const float* x, const float* w, const float* b, const float* add, void* out, long rows,
                       long rows_per_batch, int C, int act, int device, cudaStream_t st) {
  if (!x || !w || !b || !out) return set_error("ag_rmsnorm_act_fwd: null pointer");
  if (C % 128) return set_error("ag_rmsnorm_act_fwd: C=%d must be a multiple of 128", C);
  if (rows < 1 || rows_per_batch < 1) return set_error("ag_rmsnorm_act_fwd: bad row counts");
  if (act != ACT_NONE && act != ACT_GELU1702 && act != ACT_GELU_ERF) return set_error("ag_rmsnorm_act_fwd: bad activation %d", act);
  rmsnorm_act_kernel<<<head_grid(rows * 32, 256, device), 256, 0, st>>>(x, w, b, add, (__nv_bfloat16*)out, rows, rows_per_batch, C, act);
  return check_launch("ag_rmsnorm_act_fwd");
}

int launch_splice_rope(const float* x, const float* scale, const float* offset, const float* cos_tab, const float* sin_tab,
                       void* out, int B, int P, int T, int Hd, int p_ld, int tab_per_site, int device, cudaStream_t st) {
  if (!x || !scale || !offset || !cos_tab || !sin_tab || !out) return set_error("ag_splice_rope_fwd: null pointer");
  if (Hd % 8 || B < 1 || P < 1 || T < 1 || p_ld < P) return set_error("ag_splice_rope_fwd: bad sizes (hidden %d must be a multiple of 8, p_ld >= P)", Hd);
  const long total = (long)B * T * P * (Hd / 4);
  splice_rope_kernel<<<head_grid(total, 256, device), 256, 0, st>>>(x, scale, offset, cos_tab, sin_tab, (__nv_bfloat16*)out, B, P, T, Hd, p_ld, tab_per_site);
  return check_launch("ag_splice_rope_fwd");
}

}  // namespace agb
