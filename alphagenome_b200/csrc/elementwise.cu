// K2/K7/K8 — vectorised, coalesced HBM-bound kernels of the trunk (no data reuse worth a tensor core):
// one-hot im2col, q/k/v LayerNorm + RoPE, bf16 transposes, pair->attention-bias projection,
// 16-token pooling + RMSNorm, the SingleToPairwise combine, row RMSNorm, organism-embedding add and the
// pair output embedding.  Reference lines are cited on each kernel (AG = alphagenome_pytorch/alphagenome.py).
#include "common.cuh"
#include "kernels.h"

namespace agb {

static inline unsigned grid_for(long work, int per_block, int device, int waves = 8) {
  long blocks = (work + per_block - 1) / per_block;
  long cap = (long)sm_count(device) * waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// ---------------------------------------------------------------------------------------------------
// DNAEmbed one-hot (AG:1171-1175) laid out as an im2col row per position so conv15 becomes one K=64 GEMM.
// One thread writes one 16-byte chunk (8 columns) of one row.
// ---------------------------------------------------------------------------------------------------
__global__ void onehot_im2col_kernel(const int64_t* __restrict__ tok, uint4* __restrict__ out, int B, int L, int width,
                                     int n_base, int ld) {
  const int chunks = ld / 8;
  const long total = (long)B * L * chunks;
  const int half = width / 2;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int ck = (int)(idx % chunks);
    const long pos = idx / chunks;
    const int l = (int)(pos % L);
    const long b = pos / L;
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int col = ck * 8 + e;
      const int t = col / n_base;
      const int base = col - t * n_base;
      float v = 0.0f;
      if (t < width) {
        const int src = l + t - half;
        if (src >= 0 && src < L) {
          const long tk = tok[b * L + src];
          v = (tk == base) ? 1.0f : 0.0f;   // negative tokens (N) match no base
        }
      }
      if (v != 0.0f) w[e >> 1] |= (e & 1) ? 0x3F800000u : 0x00003F80u;  // bf16(1.0) = 0x3F80
    }
    out[idx] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// The model's shape (4 bases, 64-column rows): a 16-byte chunk is exactly two taps, so a thread needs two tokens, shifts
// instead of the generic kernel's 64-bit divisions, and no per-column loop (the generic kernel is issue-bound: 144 us
// for the 134 MB it writes at 1 Mbp, 1 TB/s).
__global__ void onehot_im2col4_kernel(const int64_t* __restrict__ tok, uint4* __restrict__ out, long rows, int L, int width) {
  const int half = width / 2;
  const long total = rows * 8;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int ck = (int)(idx & 7);
    const long pos = idx >> 3;
    const int l = (int)(pos % L);
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int t = 2 * ck + h;
      const int src = l + t - half;
      if (t < width && src >= 0 && src < L) {
        const long tk = __ldg(tok + (pos - l) + src);
        const uint32_t one = (tk & 1) ? 0x3F800000u : 0x00003F80u;   // bf16(1.0) = 0x3F80 in the low / high half
        w[2 * h] = (tk >> 1) == 0 ? one : 0u;                        // bases 0, 1; negative tokens (N) match no base
        w[2 * h + 1] = (tk >> 1) == 1 ? one : 0u;                    // bases 2, 3
      }
    }
    out[idx] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// ---------------------------------------------------------------------------------------------------
// q/k/v LayerNorm + interleaved RoPE (AG:715-734, 559-566).  One warp per (token, unit); units 0..H-1 are
// the q heads, H is k, H+1 is v.  Each lane owns d/32 (or dv/32) consecutive channels, so RoPE pairs
// (2i, 2i+1) are lane-local.
// ---------------------------------------------------------------------------------------------------
template <int PER_LANE>
__device__ __forceinline__ void ln_rope_unit(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                             const float* __restrict__ w, const float* __restrict__ bvec,
                                             const float* __restrict__ cs, const float* __restrict__ sn, int lane,
                                             bool rope) {
  float x[PER_LANE];
  const int c0 = lane * PER_LANE;
#pragma unroll
  for (int i = 0; i < PER_LANE; i += 2) {
    const float2 t = *reinterpret_cast<const float2*>(src + c0 + i);
    x[i] = t.x;
    x[i + 1] = t.y;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) s += x[i];
  const float mean = warp_sum(s) / (32.0f * PER_LANE);
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) {
    x[i] -= mean;
    v += x[i] * x[i];
  }
  const float rstd = rsqrtf(warp_sum(v) / (32.0f * PER_LANE) + 1e-5f);
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) x[i] = x[i] * rstd * __ldg(w + c0 + i) + __ldg(bvec + c0 + i);
  if (rope) {
#pragma unroll
    for (int i = 0; i < PER_LANE; i += 2) {
      const int f = (c0 + i) >> 1;
      const float c = __ldg(cs + f), sgn = __ldg(sn + f);
      const float a = x[i], b2 = x[i + 1];
      x[i] = a * c - b2 * sgn;      // t*cos + rotate_half(t)*sin with rotate_half = (-x_odd, x_even)
      x[i + 1] = b2 * c + a * sgn;
    }
  }
#pragma unroll
  for (int i = 0; i < PER_LANE; i += 2) *reinterpret_cast<uint32_t*>(dst + c0 + i) = pack2(x[i], x[i + 1]);
}

__global__ void qkv_post_kernel(const float* __restrict__ qkv, __nv_bfloat16* __restrict__ Q, __nv_bfloat16* __restrict__ K,
                                __nv_bfloat16* __restrict__ V, const float* __restrict__ cos_tab,
                                const float* __restrict__ sin_tab, const float* q_w, const float* q_b, const float* k_w,
                                const float* k_b, const float* v_w, const float* v_b, long tokens, int n, int heads) {
  const int lane = threadIdx.x & 31;
  const int units = heads + 2;
  const long total = tokens * units;
  const int C = heads * 128 + 128 + 192;
  for (long wi = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5; wi < total; wi += ((long)gridDim.x * blockDim.x) >> 5) {
    const int u = (int)(wi % units);
    const long tkn = wi / units;
    const int pos = (int)(tkn % n);
    const float* src = qkv + tkn * C;
    const float* cs = cos_tab + (long)pos * 64;
    const float* sn = sin_tab + (long)pos * 64;
    if (u < heads) ln_rope_unit<4>(src + u * 128, Q + tkn * (heads * 128) + u * 128, q_w, q_b, cs, sn, lane, true);
    else if (u == heads) ln_rope_unit<4>(src + heads * 128, K + tkn * 128, k_w, k_b, cs, sn, lane, true);
    else ln_rope_unit<6>(src + heads * 128 + 128, V + tkn * 192, v_w, v_b, cs, sn, lane, false);
  }
}

// ---------------------------------------------------------------------------------------------------
// Sequence parallel: q/k/v LayerNorm + RoPE FUSED WITH THE K/V ALL-GATHER.  Every rank attends over the whole context
// (AG:748), so each rank's K and V rows are needed by all ranks.  Instead of writing K/V locally and calling an
// all-gather, the warp that finishes a K (V) row stores it straight into EVERY rank's gathered [B][n_all][d + dv]
// buffer (peer-mapped symmetric memory: remote stores travel over NVLink / NVSwitch while the other warps are still
// normalising).  Q stays local.  The consumers are ordered behind one cross-rank barrier by the caller.
// ---------------------------------------------------------------------------------------------------
struct PeerPtrs {
  __nv_bfloat16* p[8];
};

template <int PER_LANE>
__device__ __forceinline__ void ln_rope_unit_peers(const float* __restrict__ src, const PeerPtrs& peers, int world, long elem_off,
                                                   const float* __restrict__ w, const float* __restrict__ bvec,
                                                   const float* __restrict__ cs, const float* __restrict__ sn, int lane, bool rope) {
  float x[PER_LANE];
  const int c0 = lane * PER_LANE;
#pragma unroll
  for (int i = 0; i < PER_LANE; i += 2) {
    const float2 t = *reinterpret_cast<const float2*>(src + c0 + i);
    x[i] = t.x;
    x[i + 1] = t.y;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) s += x[i];
  const float mean = warp_sum(s) / (32.0f * PER_LANE);
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) {
    x[i] -= mean;
    v += x[i] * x[i];
  }
  const float rstd = rsqrtf(warp_sum(v) / (32.0f * PER_LANE) + 1e-5f);
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) x[i] = x[i] * rstd * __ldg(w + c0 + i) + __ldg(bvec + c0 + i);
  if (rope) {
#pragma unroll
    for (int i = 0; i < PER_LANE; i += 2) {
      const int f = (c0 + i) >> 1;
      const float c = __ldg(cs + f), sgn = __ldg(sn + f);
      const float a = x[i], b2 = x[i + 1];
      x[i] = a * c - b2 * sgn;
      x[i + 1] = b2 * c + a * sgn;
    }
  }
  uint32_t pk[PER_LANE / 2];
#pragma unroll
  for (int i = 0; i < PER_LANE; i += 2) pk[i / 2] = pack2(x[i], x[i + 1]);
  for (int r = 0; r < world; ++r) {
    __nv_bfloat16* row = peers.p[r] + elem_off + c0;
    if constexpr (PER_LANE == 4) {
      *reinterpret_cast<uint2*>(row) = make_uint2(pk[0], pk[1]);      // K: the warp writes one contiguous 256 B row
    } else {
      uint32_t* dst = reinterpret_cast<uint32_t*>(row);
#pragma unroll
      for (int i = 0; i < PER_LANE / 2; ++i) dst[i] = pk[i];
    }
  }
}

__global__ void qkv_post_gather_kernel(const float* __restrict__ qkv, __nv_bfloat16* __restrict__ Q, PeerPtrs kv, int world,
                                       long n_all, long row0, const float* __restrict__ cos_tab,
                                       const float* __restrict__ sin_tab, const float* q_w, const float* q_b, const float* k_w,
                                       const float* k_b, const float* v_w, const float* v_b, long tokens, int n, int heads) {
  const int lane = threadIdx.x & 31;
  const int units = heads + 2;
  const long total = tokens * units;
  const int C = heads * 128 + 128 + 192;
  for (long wi = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5; wi < total; wi += ((long)gridDim.x * blockDim.x) >> 5) {
    const int u = (int)(wi % units);
    const long tkn = wi / units;
    const int pos = (int)(tkn % n);
    const long bb = tkn / n;
    const float* src = qkv + tkn * C;
    const float* cs = cos_tab + (long)pos * 64;
    const float* sn = sin_tab + (long)pos * 64;
    const long grow = (bb * n_all + row0 + pos) * 320;          // row of [K | V] in a gathered buffer
    if (u < heads) ln_rope_unit<4>(src + u * 128, Q + tkn * (heads * 128) + u * 128, q_w, q_b, cs, sn, lane, true);
    else if (u == heads) ln_rope_unit_peers<4>(src + heads * 128, kv, world, grow, k_w, k_b, cs, sn, lane, true);
    else ln_rope_unit_peers<6>(src + heads * 128 + 128, kv, world, grow + 128, v_w, v_b, cs, sn, lane, false);
  }
}

// One-launch push all-gather over peer memory: `bytes` of src go to byte offset dst_off of every rank's buffer.
__global__ void peer_push_kernel(const uint4* __restrict__ src, PeerPtrs peers, int world, long n16, long dst_off16) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n16 * world; i += (long)gridDim.x * blockDim.x) {
    const int r = (int)(i / n16);
    const long j = i - r * n16;
    reinterpret_cast<uint4*>(peers.p[r])[dst_off16 + j] = __ldg(src + j);
  }
}

// ---------------------------------------------------------------------------------------------------
// bf16 transpose through a padded smem tile: out[b][c][r] = in[b][r][c]
// ---------------------------------------------------------------------------------------------------
__global__ void transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int R, int C,
                                      long in_bs, int in_ld, long out_bs, int out_ld) {
  __shared__ __nv_bfloat16 tile[64][66];
  const int b = blockIdx.z;
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 256 threads: 64 x 4
  for (int i = ty; i < 64; i += 4) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < R && c < C) ? in[b * in_bs + (long)r * in_ld + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int i = ty; i < 64; i += 4) {
    const int c = c0 + i, r = r0 + tx;
    if (c < C && r < R) out[b * out_bs + (long)c * out_ld + r] = tile[tx][i];
  }
}

// ---------------------------------------------------------------------------------------------------
// to_attn_bias (AG:688-693) for up to two attention layers that read the same pair tensor (the pair only changes
// on every second tower layer), in ONE pass over it.  A warp takes 8 consecutive cells: float4 per lane (C = 128),
// per cell sets*8 partial dot products, reduced over the 32 lanes with a transposing butterfly (16 + 8 + 4 + 2 + 2
// values exchanged instead of 5 shuffles per value); lane l ends up owning output plane l >> 1 and writes its 8
// cells as two float4 (full 32-byte sectors).  erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, two MUFU ops).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf_as(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = 1.0f - p * t * __expf(-z * z);     // erf(|x| / sqrt 2)
  return 0.5f * x * (1.0f + copysignf(e, x));
}

template <int SETS>
__global__ void __launch_bounds__(256) pair_bias_kernel(const float* __restrict__ pair, const float* __restrict__ scale,
                                                        const float* __restrict__ shift, const float* __restrict__ W,
                                                        float* __restrict__ bias, long groups, long per_b, int heads,
                                                        long set_stride) {
  constexpr int NV = SETS * 8;   // values per cell (heads == 8)
  const int lane = threadIdx.x & 31;
  float4 sc[SETS], sh[SETS];
#pragma unroll
  for (int s = 0; s < SETS; ++s) {
    sc[s] = __ldg(reinterpret_cast<const float4*>(scale + s * 128) + lane);
    sh[s] = __ldg(reinterpret_cast<const float4*>(shift + s * 128) + lane);
  }
  const int plane = lane >> (SETS == 2 ? 1 : 2);   // output plane (set * 8 + head) this lane ends up owning
  for (long grp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5; grp < groups; grp += ((long)gridDim.x * blockDim.x) >> 5) {
    const long cell0 = grp * 8;
    float4 x[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) x[c] = __ldg(reinterpret_cast<const float4*>(pair + (cell0 + c) * 128) + lane);
    float res[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float v[NV];
#pragma unroll
      for (int s = 0; s < SETS; ++s) {
        float4 y;
        y.x = gelu_erf_as(fmaf(x[c].x, sc[s].x, sh[s].x));
        y.y = gelu_erf_as(fmaf(x[c].y, sc[s].y, sh[s].y));
        y.z = gelu_erf_as(fmaf(x[c].z, sc[s].z, sh[s].z));
        y.w = gelu_erf_as(fmaf(x[c].w, sc[s].w, sh[s].w));
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(W + (s * 8 + g) * 128) + lane);
          v[s * 8 + g] = y.x * w.x + y.y * w.y + y.z * w.z + y.w * w.w;
        }
      }
      // transposing butterfly: after the step with offset o, a lane keeps the half of its values selected by (lane & o)
#pragma unroll
      for (int n = NV, o = 16; n > 1; n >>= 1, o >>= 1) {
        const bool hi = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
          const float send = hi ? v[i] : v[i + n / 2];
          const float keep = hi ? v[i + n / 2] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
      }
      // SETS == 2: 16 values used offsets 16..2, one more plain step; SETS == 1: 8 values used 16..4, two more
      float r = v[0];
#pragma unroll
      for (int o = (SETS == 2 ? 1 : 2); o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
      res[c] = r;
    }
    if ((lane & (SETS == 2 ? 1 : 3)) == 0) {
      const long b = cell0 / per_b;
      const long ij = cell0 - b * per_b;
      const int set = plane >> 3, g = plane & 7;
      float* dst = bias + set * set_stride + (b * heads + g) * per_b + ij;
      reinterpret_cast<float4*>(dst)[0] = make_float4(res[0], res[1], res[2], res[3]);
      reinterpret_cast<float4*>(dst)[1] = make_float4(res[4], res[5], res[6], res[7]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// to_attn_bias on the tensor core (the default path).  The projection is a [cells x 128] . [128 x 8] product per layer:
// as plain FMAs + a transposing butterfly it cost ~260 issue slots per cell and lane and ran at 0.16 of the HBM rate
// (ncu r1: ALU-bound).  Here a warp owns 16 consecutive cells and feeds them to mma.sync.m16n8k16 (bf16 in, fp32
// accumulate): N = 8 heads is exactly one n-tile, so the whole projection of a layer is 8 MMAs per 16 cells, and the
// result fragment (thread = 2 cells x 2 heads) stores as full 32-byte sectors.  (tcgen05 would need M = 128 x N >= 16
// tiles staged through shared memory and TMEM for a product whose tensor time is ~1 % of its HBM time; the legacy warp
// MMA keeps the operand in the registers the GELU just produced.)
//
// The MMA sums over k, so the k order is free: lane (g = lane >> 2, t = lane & 3) loads channels 16s + 4t .. + 3 of
// cells g and g + 8 with ONE 128-bit load each per k-step s and assigns them to k-slots {2t, 2t+1, 2t+8, 2t+9}; the W
// fragment uses the same map.  The operand is bf16 like every other tensor-core operand of the path, so erf-GELU is
// evaluated in the one-MUFU form 0.5 x (1 + tanh(x (c1 + c3 x^2 + c5 x^4))), coefficients fitted to erf-GELU
// (max |error| 2.5e-5 before tanh.approx's 2^-11, below the bf16 rounding of the operand).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf_tanh(float x) {
  const float t = fminf(x * x, 64.0f);            // beyond |x| = 8 the argument of tanh is >= 13.8: exactly +-1
  float q = fmaf(t, -3.51516786e-4f, 3.70056460e-2f);
  q = fmaf(t, q, 7.97507884e-1f);
  const float h = 0.5f * x;
  return fmaf(h, tanh_approx(x * q), h);
}

__device__ __forceinline__ void mma_m16n8k16_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int SETS>
__global__ void __launch_bounds__(256, 2) pair_bias_mma_kernel(const float* __restrict__ pair, const float* __restrict__ scale,
                                                               const float* __restrict__ shift, const float* __restrict__ W,
                                                               float* __restrict__ bias, long groups, long per_b, long set_stride) {
  __shared__ float4 s_sc[SETS][32], s_sh[SETS][32];
  if (threadIdx.x < SETS * 32) {
    const int s = threadIdx.x >> 5, i = threadIdx.x & 31;
    s_sc[s][i] = __ldg(reinterpret_cast<const float4*>(scale + s * 128) + i);
    s_sh[s][i] = __ldg(reinterpret_cast<const float4*>(shift + s * 128) + i);
  }
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  uint32_t bw[SETS][8][2];   // W fragments of head g: k-step s holds channels 16s + 4t .. + 3
#pragma unroll
  for (int s = 0; s < SETS; ++s)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(W + (s * 8 + g) * 128 + 16 * k) + t);
      bw[s][k][0] = pack2(w.x, w.y);
      bw[s][k][1] = pack2(w.z, w.w);
    }
  __syncthreads();
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  for (long grp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5; grp < groups; grp += nwarps) {
    const long cell0 = grp * 16;
    const float4* p0 = reinterpret_cast<const float4*>(pair + (cell0 + g) * 128) + t;   // cell g; cell g + 8 is 256 float4 on
    float acc[SETS][4];
#pragma unroll
    for (int s = 0; s < SETS; ++s) acc[s][0] = acc[s][1] = acc[s][2] = acc[s][3] = 0.f;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float4 xa[4], xb[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        xa[j] = __ldg(p0 + 4 * (half * 4 + j));
        xb[j] = __ldg(p0 + 256 + 4 * (half * 4 + j));
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = half * 4 + j;
#pragma unroll
        for (int s = 0; s < SETS; ++s) {
          const float4 sc = s_sc[s][4 * k + t], sh = s_sh[s][4 * k + t];
          const uint32_t a0 = pack2(gelu_erf_tanh(fmaf(xa[j].x, sc.x, sh.x)), gelu_erf_tanh(fmaf(xa[j].y, sc.y, sh.y)));
          const uint32_t a2 = pack2(gelu_erf_tanh(fmaf(xa[j].z, sc.z, sh.z)), gelu_erf_tanh(fmaf(xa[j].w, sc.w, sh.w)));
          const uint32_t a1 = pack2(gelu_erf_tanh(fmaf(xb[j].x, sc.x, sh.x)), gelu_erf_tanh(fmaf(xb[j].y, sc.y, sh.y)));
          const uint32_t a3 = pack2(gelu_erf_tanh(fmaf(xb[j].z, sc.z, sh.z)), gelu_erf_tanh(fmaf(xb[j].w, sc.w, sh.w)));
          mma_m16n8k16_bf16(acc[s], a0, a1, a2, a3, bw[s][k][0], bw[s][k][1]);
        }
      }
    }
    // result fragment: acc[.][0..1] = cell g, heads 2t, 2t+1; acc[.][2..3] = cell g + 8.  A store instruction writes, for
    // each of 4 heads, 8 consecutive cells = one full 32-byte sector.
    const long b = cell0 / per_b;
    const long ij = cell0 - b * per_b;
#pragma unroll
    for (int s = 0; s < SETS; ++s) {
      float* d0 = bias + s * set_stride + (b * 8 + 2 * t) * per_b + ij + g;
      d0[0] = acc[s][0];
      d0[per_b] = acc[s][1];
      d0[8] = acc[s][2];
      d0[per_b + 8] = acc[s][3];
    }
  }
}

// cells that do not fill a group of 8 (or unaligned shapes): one warp per cell, plain reduction
__global__ void pair_bias_tail_kernel(const float* __restrict__ pair, const float* __restrict__ scale,
                                      const float* __restrict__ shift, const float* __restrict__ W, float* __restrict__ bias,
                                      long cell_begin, long cells, long per_b, int heads, int sets, long set_stride) {
  const int lane = threadIdx.x & 31;
  for (long cell = cell_begin + ((blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5); cell < cells;
       cell += ((long)gridDim.x * blockDim.x) >> 5) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(pair + cell * 128) + lane);
    const long b = cell / per_b;
    const long ij = cell - b * per_b;
    for (int s = 0; s < sets; ++s) {
      const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + s * 128) + lane);
      const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + s * 128) + lane);
      float4 y;
      y.x = gelu_erf_as(fmaf(x.x, sc.x, sh.x));
      y.y = gelu_erf_as(fmaf(x.y, sc.y, sh.y));
      y.z = gelu_erf_as(fmaf(x.z, sc.z, sh.z));
      y.w = gelu_erf_as(fmaf(x.w, sc.w, sh.w));
      for (int g = 0; g < heads; ++g) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(W + (s * heads + g) * 128) + lane);
        const float d = warp_sum(y.x * w.x + y.y * w.y + y.z * w.z + y.w * w.w);
        if (lane == 0) bias[s * set_stride + (b * heads + g) * per_b + ij] = d;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// SingleToPairwise prologue (AG:828-829, 852).  One block per pooled row.
// ---------------------------------------------------------------------------------------------------
__global__ void pool_rmsnorm_kernel(const float* __restrict__ single, const float* __restrict__ w, const float* __restrict__ bvec,
                                    __nv_bfloat16* __restrict__ x_out, __nv_bfloat16* __restrict__ xg_out, int C, int pool) {
  extern __shared__ float sh_x[];  // [C]
  __shared__ float red[32];
  const long prow = blockIdx.x;
  const float* src = single + prow * pool * (long)C;
  float ss = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int t = 0; t < pool; ++t) acc += src[(long)t * C + c];
    acc /= (float)pool;
    sh_x[c] = acc;
    ss += acc * acc;
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) red[0] = t;
  }
  __syncthreads();
  const float inv = rsqrtf(red[0] / (float)C + 1e-5f);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float y = sh_x[c] * inv * __ldg(w + c) + __ldg(bvec + c);
    x_out[prow * C + c] = __float2bfloat16_rn(y);
    xg_out[prow * C + c] = __float2bfloat16_rn(act_apply(y, ACT_GELU_ERF));
  }
}

// C == 4 * blockDim (the model: 1536 channels, 384 threads), pool == 16: a thread owns 4 channels and has all 16 of its
// loads as 128-bit requests (the generic kernel above issued scalar loads from only 512 blocks of 256 threads:
// 1.4 TB/s).  The pooled means are bit-identical; the sum of squares is reduced in a different order (fp32 rounding).
__global__ void __launch_bounds__(384) pool16_rmsnorm4_kernel(const float* __restrict__ single, const float* __restrict__ w,
                                                              const float* __restrict__ bvec, __nv_bfloat16* __restrict__ x_out,
                                                              __nv_bfloat16* __restrict__ xg_out, int C) {
  __shared__ float red[32];
  const long prow = blockIdx.x;
  const float4* src = reinterpret_cast<const float4*>(single + prow * 16 * (long)C) + threadIdx.x;
  float4 v[16];
#pragma unroll
  for (int t = 0; t < 16; ++t) v[t] = __ldg(src + (long)t * (C / 4));
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int t = 0; t < 16; ++t) { acc.x += v[t].x; acc.y += v[t].y; acc.z += v[t].z; acc.w += v[t].w; }
  acc.x /= 16.0f; acc.y /= 16.0f; acc.z /= 16.0f; acc.w /= 16.0f;
  float ss = acc.x * acc.x + acc.y * acc.y + acc.z * acc.z + acc.w * acc.w;
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) red[0] = t;
  }
  __syncthreads();
  const float inv = rsqrtf(red[0] / (float)C + 1e-5f);
  const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + threadIdx.x);
  const float4 bb = __ldg(reinterpret_cast<const float4*>(bvec) + threadIdx.x);
  const float y0 = acc.x * inv * ww.x + bb.x, y1 = acc.y * inv * ww.y + bb.y, y2 = acc.z * inv * ww.z + bb.z, y3 = acc.w * inv * ww.w + bb.w;
  uint2 px, pg;
  px.x = pack2(y0, y1);
  px.y = pack2(y2, y3);
  pg.x = pack2(act_apply(y0, ACT_GELU_ERF), act_apply(y1, ACT_GELU_ERF));
  pg.y = pack2(act_apply(y2, ACT_GELU_ERF), act_apply(y3, ACT_GELU_ERF));
  reinterpret_cast<uint2*>(x_out + prow * C)[threadIdx.x] = px;
  reinterpret_cast<uint2*>(xg_out + prow * C)[threadIdx.x] = pg;
}

// ---------------------------------------------------------------------------------------------------
// SingleToPairwise combine (AG:835-856, relative_shift AG:523-530) + the pre-norm of the row attention that
// consumes the result (RMSNormWithBias, AG:400-405, via NormWrapper AG:974).
// Block = (b, 8 rows i, 32 columns j), 256 threads.
//   phase 1: sim[i][h][j] = qk[h][i][j] + 0.5 (relq[h][i][P+j-i] + relk[h][j][P+i-j]) into smem.  qk / relq are read
//            with j fastest (coalesced); relk with i fastest, so the 8 rows of the tile use one 32-byte sector per
//            (h, j) instead of one sector per element (the first version's 8x read amplification).
//   phase 2: warp = row i, lane = 4 consecutive channels; cells in groups of 4 columns: per head one LDS.128 of
//            Wp^T and one broadcast LDS.128 of sim feed 16 FMAs.  Then + bias + outer_q[i] + outer_k[j] + prev,
//            fp32 store (512 B per cell, coalesced), and the RMSNorm of the cell by a warp reduction -> bf16.
// ---------------------------------------------------------------------------------------------------
constexpr int kPcI = 8, kPcJ = 32, kPcH = 32, kPcD = 128;
constexpr int kPcSimI = kPcH * kPcJ + 4;                       // row stride of sim (floats); +4 skews the banks per i
constexpr int kPcSmem = (kPcH * kPcD + kPcI * kPcSimI) * 4;    // Wp^T [h][c] + sim [i][h][j]

__global__ void __launch_bounds__(256) pair_combine_kernel(const float* __restrict__ qk, const float* __restrict__ relq,
                                                           const float* __restrict__ relk, const float* __restrict__ Wp,
                                                           const float* __restrict__ bp, const float* __restrict__ outer,
                                                           const float* __restrict__ prev, float* __restrict__ pair,
                                                           const float* __restrict__ nw, const float* __restrict__ nb,
                                                           __nv_bfloat16* __restrict__ pn, int P, int qk_ld, int rel_ld,
                                                           int rows, int i_off) {
  extern __shared__ float pc_smem[];
  float* wt = pc_smem;                   // [32 h][128 c]
  float* sim = pc_smem + kPcH * kPcD;    // [8 i][32 h][32 j] (+4 per i)
  const int j0 = blockIdx.x * kPcJ, il0 = blockIdx.y * kPcI, b = blockIdx.z;
  const int tid = threadIdx.x;
  // Wp is [c][h]: transpose into smem once per block (16 KB, L2-resident)
  for (int idx = tid; idx < kPcH * kPcD; idx += 256) {
    const int c = idx >> 5, h = idx & 31;
    wt[h * kPcD + c] = __ldg(Wp + idx);
  }
  // pass A: j fastest.  The loads of 8 entries are issued before any is used (the loop is latency-bound otherwise).
  constexpr int kIters = kPcI * kPcH * kPcJ / 256;   // 32 entries per thread
#pragma unroll 1
  for (int it0 = 0; it0 < kIters; it0 += 8) {
    float va[8], vb[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = (it0 + u) * 256 + tid;
      const int jj = idx & 31, h = (idx >> 5) & 31, ii = idx >> 10;
      const int il = il0 + ii, j = j0 + jj;
      va[u] = vb[u] = 0.f;
      if (il < rows && j < P) {
        const long row = ((long)b * kPcH + h) * rows + il;
        va[u] = __ldg(qk + row * qk_ld + j);
        vb[u] = __ldg(relq + row * (long)rel_ld + (P + j - (il + i_off)));
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = (it0 + u) * 256 + tid;
      const int jj = idx & 31, h = (idx >> 5) & 31, ii = idx >> 10;
      sim[ii * kPcSimI + h * kPcJ + jj] = fmaf(0.5f, vb[u], va[u]);
    }
  }
  __syncthreads();
  // pass B: i fastest (8 consecutive floats of one relk row per (h, j))
#pragma unroll 1
  for (int it0 = 0; it0 < kIters; it0 += 8) {
    float vk[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = (it0 + u) * 256 + tid;
      const int ii = idx & 7, jj = (idx >> 3) & 31, h = idx >> 8;
      const int il = il0 + ii, j = j0 + jj;
      vk[u] = 0.f;
      if (il < rows && j < P) vk[u] = __ldg(relk + (((long)b * kPcH + h) * P + j) * (long)rel_ld + (P + (il + i_off) - j));
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = (it0 + u) * 256 + tid;
      const int ii = idx & 7, jj = (idx >> 3) & 31, h = idx >> 8;
      sim[ii * kPcSimI + h * kPcJ + jj] += 0.5f * vk[u];
    }
  }
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31;
  const int il = il0 + warp;
  if (il >= rows) return;
  const int i = il + i_off;
  const int c0 = lane * 4;
  const float4 bias4 = __ldg(reinterpret_cast<const float4*>(bp + c0));
  const float4 oq = __ldg(reinterpret_cast<const float4*>(outer + ((long)b * P + i) * (2 * kPcD) + c0));
  const float4 base = make_float4(bias4.x + oq.x, bias4.y + oq.y, bias4.z + oq.z, bias4.w + oq.w);
  float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = w4;
  if (pn) {
    w4 = __ldg(reinterpret_cast<const float4*>(nw + c0));
    b4 = __ldg(reinterpret_cast<const float4*>(nb + c0));
  }
  const float* simrow = sim + warp * kPcSimI;
  for (int jg = 0; jg < kPcJ / 4; ++jg) {
    const int jb = j0 + jg * 4;
    if (jb >= P) break;
    // outer_k / prev are requested here and only added after the head loop, so their latency hides behind it
    float4 acc[4], pv[4], okv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = jb + q < P ? jb + q : P - 1;   // clamped: columns past P are computed but never stored
      okv[q] = __ldg(reinterpret_cast<const float4*>(outer + ((long)b * P + j) * (2 * kPcD) + kPcD + c0));
      acc[q] = base;
      pv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (prev) pv[q] = __ldg(reinterpret_cast<const float4*>(prev + ((((long)b * rows + il) * P + j) * kPcD) + c0));
    }
#pragma unroll 8
    for (int h = 0; h < kPcH; ++h) {
      const float4 w = *reinterpret_cast<const float4*>(wt + h * kPcD + c0);
      const float4 sv = *reinterpret_cast<const float4*>(simrow + h * kPcJ + jg * 4);
      acc[0].x = fmaf(w.x, sv.x, acc[0].x); acc[0].y = fmaf(w.y, sv.x, acc[0].y); acc[0].z = fmaf(w.z, sv.x, acc[0].z); acc[0].w = fmaf(w.w, sv.x, acc[0].w);
      acc[1].x = fmaf(w.x, sv.y, acc[1].x); acc[1].y = fmaf(w.y, sv.y, acc[1].y); acc[1].z = fmaf(w.z, sv.y, acc[1].z); acc[1].w = fmaf(w.w, sv.y, acc[1].w);
      acc[2].x = fmaf(w.x, sv.z, acc[2].x); acc[2].y = fmaf(w.y, sv.z, acc[2].y); acc[2].z = fmaf(w.z, sv.z, acc[2].z); acc[2].w = fmaf(w.w, sv.z, acc[2].w);
      acc[3].x = fmaf(w.x, sv.w, acc[3].x); acc[3].y = fmaf(w.y, sv.w, acc[3].y); acc[3].z = fmaf(w.z, sv.w, acc[3].z); acc[3].w = fmaf(w.w, sv.w, acc[3].w);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = jb + q;
      float4 y = make_float4(acc[q].x + (pv[q].x + okv[q].x), acc[q].y + (pv[q].y + okv[q].y), acc[q].z + (pv[q].z + okv[q].z),
                             acc[q].w + (pv[q].w + okv[q].w));
      const float ss = warp_sum(y.x * y.x + y.y * y.y + y.z * y.z + y.w * y.w);   // all lanes take part
      if (j < P) {
        const long o = (((long)b * rows + il) * P + j) * kPcD + c0;
        *reinterpret_cast<float4*>(pair + o) = y;
        if (pn) {
          const float inv = rsqrtf(ss * (1.0f / kPcD) + 1e-5f);
          uint2 pk;
          pk.x = pack2(y.x * inv * w4.x + b4.x, y.y * inv * w4.y + b4.y);
          pk.y = pack2(y.z * inv * w4.z + b4.z, y.w * inv * w4.w + b4.w);
          *reinterpret_cast<uint2*>(pn + o) = pk;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// RMSNormWithBias (AG:400-405) fp32 -> bf16.  One warp per row, C = 128 * k.
// ---------------------------------------------------------------------------------------------------
__global__ void rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bvec,
                               __nv_bfloat16* __restrict__ out, long rows, int C) {
  const int lane = threadIdx.x & 31;
  const int nv = C / 128;  // float4 per lane
  for (long row = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5; row < rows;
       row += ((long)gridDim.x * blockDim.x) >> 5) {
    const float4* src = reinterpret_cast<const float4*>(x + row * C);
    float ss = 0.f;
    for (int k = 0; k < nv; ++k) {
      const float4 v = __ldg(src + k * 32 + lane);
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    const float inv = rsqrtf(warp_sum(ss) / (float)C + 1e-5f);
    for (int k = 0; k < nv; ++k) {
      const float4 v = __ldg(src + k * 32 + lane);
      const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + k * 32 + lane);
      const float4 bb = __ldg(reinterpret_cast<const float4*>(bvec) + k * 32 + lane);
      uint2 pk;
      pk.x = pack2(v.x * inv * ww.x + bb.x, v.y * inv * ww.y + bb.y);
      pk.y = pack2(v.z * inv * ww.z + bb.z, v.w * inv * ww.w + bb.w);
      *reinterpret_cast<uint2*>(out + row * C + (k * 32 + lane) * 4) = pk;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// organism embedding add + tower pre-norm (AG:1119-1120, 636)
// ---------------------------------------------------------------------------------------------------
__global__ void add_embed_norm_kernel(const float4* __restrict__ x, const float* __restrict__ emb,
                                      const float* __restrict__ scale, const float* __restrict__ shift,
                                      float4* __restrict__ single, uint2* __restrict__ a_out, long rows_per_b, int C4,
                                      long total) {
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c4 = (int)(idx % C4);
    const long b = idx / (rows_per_b * C4);
    float4 v = x[idx];
    const float4 e = __ldg(reinterpret_cast<const float4*>(emb + b * (C4 * 4L)) + c4);
    v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
    single[idx] = v;
    const float4 s = __ldg(reinterpret_cast<const float4*>(scale) + c4);
    const float4 t = __ldg(reinterpret_cast<const float4*>(shift) + c4);
    uint2 pk;
    pk.x = pack2(fmaf(v.x, s.x, t.x), fmaf(v.y, s.y, t.y));
    pk.y = pack2(fmaf(v.z, s.z, t.z), fmaf(v.w, s.w, t.w));
    a_out[idx] = pk;
  }
}

// ---------------------------------------------------------------------------------------------------
// OutputPairEmbedding (AG:1248-1259; symmetrize AG:121-122).  One warp per cell, C = 128.
// ---------------------------------------------------------------------------------------------------
__global__ void pair_out_embed_kernel(const float* __restrict__ pair, const float* __restrict__ pairT,
                                      const float* __restrict__ w, const float* __restrict__ bvec,
                                      const float* __restrict__ emb, float* __restrict__ out, int B, int rows, int P) {
  // pair: this rank's rows [B][rows][P][128].  pairT: the transposed blocks, [P/rows][B][rows (j)][rows (i)][128]
  // (block s holds pair[b][s*rows + jl][i_off + il]); with one rank rows == P and pairT == pair.
  const int lane = threadIdx.x & 31;
  const long cells = (long)B * rows * P;
  const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + lane);
  const float4 bb = __ldg(reinterpret_cast<const float4*>(bvec) + lane);
  for (long cell = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5; cell < cells;
       cell += ((long)gridDim.x * blockDim.x) >> 5) {
    const long b = cell / ((long)rows * P);
    const long ij = cell - b * (long)rows * P;
    const long il = ij / P, j = ij - il * P;
    const long sblk = j / rows, jl = j - sblk * rows;
    const float4 a = __ldg(reinterpret_cast<const float4*>(pair + cell * 128) + lane);
    const float4 t = __ldg(reinterpret_cast<const float4*>(pairT + (((sblk * B + b) * rows + jl) * rows + il) * 128) + lane);
    float4 v;
    v.x = 0.5f * (a.x + t.x); v.y = 0.5f * (a.y + t.y); v.z = 0.5f * (a.z + t.z); v.w = 0.5f * (a.w + t.w);
    const float inv = rsqrtf(warp_sum(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w) / 128.0f + 1e-5f);
    const float4 e = __ldg(reinterpret_cast<const float4*>(emb + b * 128) + lane);
    float4 y;
    y.x = act_apply(v.x * inv * ww.x + bb.x + e.x, ACT_GELU1702);
    y.y = act_apply(v.y * inv * ww.y + bb.y + e.y, ACT_GELU1702);
    y.z = act_apply(v.z * inv * ww.z + bb.z + e.z, ACT_GELU1702);
    y.w = act_apply(v.w * inv * ww.w + bb.w + e.w, ACT_GELU1702);
    reinterpret_cast<float4*>(out + cell * 128)[lane] = y;
  }
}

// ====================================================================================================
// host launchers
// ====================================================================================================
int launch_onehot_im2col(const int64_t* tokens, void* out, int B, int L, int width, int n_base, int ld, int device,
                         cudaStream_t st) {
  if (!tokens || !out) return set_error("ag_onehot_im2col_fwd: null pointer");
  if (ld % 8 || ld < width * n_base) return set_error("ag_onehot_im2col_fwd: ld=%d must be a multiple of 8 and >= width*n_base=%d", ld, width * n_base);
  if (!(width & 1)) return set_error("ag_onehot_im2col_fwd: width must be odd");
  const long total = (long)B * L * (ld / 8);
  if (n_base == 4 && ld == 64)
    onehot_im2col4_kernel<<<grid_for(total, 256, device), 256, 0, st>>>(tokens, reinterpret_cast<uint4*>(out), (long)B * L, L, width);
  else
    onehot_im2col_kernel<<<grid_for(total, 256, device), 256, 0, st>>>(tokens, reinterpret_cast<uint4*>(out), B, L, width, n_base, ld);
  return check_launch("ag_onehot_im2col_fwd");
}

int launch_qkv_post(const float* qkv, void* Q, void* K, void* V, const float* cos_tab, const float* sin_tab,
                    const float* q_w, const float* q_b, const float* k_w, const float* k_b, const float* v_w,
                    const float* v_b, int B, int n, int heads, int d, int dv, int device, cudaStream_t st) {
  if (d != 128 || dv != 192) return set_error("ag_qkv_post_fwd: built for d=128, dv=192 (got %d, %d)", d, dv);
  if (!qkv || !Q || !K || !V || !cos_tab || !sin_tab) return set_error("ag_qkv_post_fwd: null pointer");
  const long warps = (long)B * n * (heads + 2);
  qkv_post_kernel<<<grid_for(warps * 32, 256, device), 256, 0, st>>>(
      qkv, (__nv_bfloat16*)Q, (__nv_bfloat16*)K, (__nv_bfloat16*)V, cos_tab, sin_tab, q_w, q_b, k_w, k_b, v_w, v_b,
      (long)B * n, n, heads);
  return check_launch("ag_qkv_post_fwd");
}

int launch_qkv_post_gather(const float* qkv, void* Q, void* const* kv_peers, int world, long n_all, long row0, const float* cos_tab,
                           const float* sin_tab, const float* q_w, const float* q_b, const float* k_w, const float* k_b,
                           const float* v_w, const float* v_b, int B, int n, int heads, int d, int dv, int device, cudaStream_t st) {
  if (d != 128 || dv != 192) return set_error("ag_qkv_post_gather_fwd: built for d=128, dv=192 (got %d, %d)", d, dv);
  if (!qkv || !Q || !kv_peers || !cos_tab || !sin_tab) return set_error("ag_qkv_post_gather_fwd: null pointer");
  if (world < 1 || world > 8) return set_error("ag_qkv_post_gather_fwd: world=%d must be in [1, 8]", world);
  if (row0 < 0 || row0 + n > n_all) return set_error("ag_qkv_post_gather_fwd: rows [%ld, %ld) outside n_all=%ld", row0, row0 + n, n_all);
  PeerPtrs kv;
  for (int r = 0; r < 8; ++r) kv.p[r] = r < world ? static_cast<__nv_bfloat16*>(kv_peers[r]) : nullptr;
  for (int r = 0; r < world; ++r)
    if (!kv.p[r] || (reinterpret_cast<uintptr_t>(kv.p[r]) & 15)) return set_error("ag_qkv_post_gather_fwd: peer buffer %d is null or misaligned", r);
  const long warps = (long)B * n * (heads + 2);
  qkv_post_gather_kernel<<<grid_for(warps * 32, 256, device), 256, 0, st>>>(qkv, (__nv_bfloat16*)Q, kv, world, n_all, row0, cos_tab, sin_tab,
                                                                           q_w, q_b, k_w, k_b, v_w, v_b, (long)B * n, n, heads);
  return check_launch("ag_qkv_post_gather_fwd");
}

int launch_peer_push(const void* src, void* const* peers, int world, long bytes, long dst_offset, int device, cudaStream_t st) {
  if (!src || !peers) return set_error("ag_peer_push_fwd: null pointer");
  if (world < 1 || world > 8) return set_error("ag_peer_push_fwd: world=%d must be in [1, 8]", world);
  if ((bytes & 15) || (dst_offset & 15) || bytes <= 0 || (reinterpret_cast<uintptr_t>(src) & 15))
    return set_error("ag_peer_push_fwd: sizes, offsets and pointers must be multiples of 16 bytes");
  PeerPtrs pp;
  for (int r = 0; r < 8; ++r) pp.p[r] = r < world ? static_cast<__nv_bfloat16*>(peers[r]) : nullptr;
  for (int r = 0; r < world; ++r)
    if (!pp.p[r] || (reinterpret_cast<uintptr_t>(pp.p[r]) & 15)) return set_error("ag_peer_push_fwd: peer buffer %d is null or misaligned", r);
  const long n16 = bytes / 16;
  peer_push_kernel<<<grid_for(n16 * world, 256, device, 2), 256, 0, st>>>(static_cast<const uint4*>(src), pp, world, n16, dst_offset / 16);
  return check_launch("ag_peer_push_fwd");
}

int launch_transpose_bf16(const void* in, void* out, int nb, int R, int C, long in_bs, int in_ld, long out_bs,
                          int out_ld, int device, cudaStream_t st) {
  if (!in || !out) return set_error("ag_transpose_bf16: null pointer");
  if (nb > 65535) return set_error("ag_transpose_bf16: too many batches (%d)", nb);
  dim3 grid((C + 63) / 64, (R + 63) / 64, nb);
  transpose_bf16_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, R, C, in_bs, in_ld, out_bs, out_ld);
  return check_launch("ag_transpose_bf16");
}

int launch_pair_bias(const float* pair, const float* scale, const float* shift, const float* W, float* bias, int B,
                     int rows, int cols, int C, int heads, int sets, long set_stride, int device, cudaStream_t st) {
  if (C != 128) return set_error("ag_pair_bias_fwd: built for pair dim 128 (got %d)", C);
  if (sets != 1 && sets != 2) return set_error("ag_pair_bias_fwd: sets must be 1 or 2 (got %d)", sets);
  if (!pair || !scale || !shift || !W || !bias) return set_error("ag_pair_bias_fwd: null pointer");
  if ((reinterpret_cast<uintptr_t>(pair) | reinterpret_cast<uintptr_t>(scale) | reinterpret_cast<uintptr_t>(shift) | reinterpret_cast<uintptr_t>(W)) & 15)
    return set_error("ag_pair_bias_fwd: pair, scale, shift and W must be 16-byte aligned");
  const long per_b = (long)rows * cols;
  const long cells = (long)B * per_b;
  // tensor-core path: groups of 16 consecutive cells inside one batch element (AGB_PAIR_BIAS_MMA=0: the fp32 FMA kernels)
  static int use_mma = -1;
  if (use_mma < 0) {
    const char* e = getenv("AGB_PAIR_BIAS_MMA");
    use_mma = (e && e[0] == '0') ? 0 : 1;
  }
  if (use_mma && heads == 8 && per_b % 16 == 0) {
    const long groups16 = cells / 16;
    const unsigned grid = grid_for(groups16 * 32, 256, device, 2);
    if (sets == 2) pair_bias_mma_kernel<2><<<grid, 256, 0, st>>>(pair, scale, shift, W, bias, groups16, per_b, set_stride);
    else pair_bias_mma_kernel<1><<<grid, 256, 0, st>>>(pair, scale, shift, W, bias, groups16, per_b, set_stride);
    return check_launch("ag_pair_bias_fwd");
  }
  // fp32 path: groups of 8 consecutive cells inside one batch element, 16-byte aligned output rows
  const bool fast = heads == 8 && per_b % 8 == 0 && (reinterpret_cast<uintptr_t>(bias) & 15) == 0 && (set_stride % 4) == 0;
  const long groups = fast ? cells / 8 : 0;
  if (groups > 0) {
    const unsigned grid = grid_for(groups * 32, 256, device, 4);
    if (sets == 2) pair_bias_kernel<2><<<grid, 256, 0, st>>>(pair, scale, shift, W, bias, groups, per_b, heads, set_stride);
    else pair_bias_kernel<1><<<grid, 256, 0, st>>>(pair, scale, shift, W, bias, groups, per_b, heads, set_stride);
    if (int rc = check_launch("ag_pair_bias_fwd")) return rc;
  }
  if (groups * 8 < cells) {
    const long rest = cells - groups * 8;
    pair_bias_tail_kernel<<<grid_for(rest * 32, 256, device), 256, 0, st>>>(pair, scale, shift, W, bias, groups * 8, cells, per_b,
                                                                            heads, sets, set_stride);
    return check_launch("ag_pair_bias_fwd");
  }
  return 0;
}

int launch_pool_rmsnorm(const float* single, const float* w, const float* b, void* x, void* xg, int B, int n, int C,
                        int pool, int device, cudaStream_t st) {
  if (n % pool) return set_error("ag_pool_rmsnorm_fwd: n=%d not divisible by pool=%d", n, pool);
  const long rows = (long)B * (n / pool);
  if (pool == 16 && C == 1536 && ((reinterpret_cast<uintptr_t>(single) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(b)) & 15) == 0)
    pool16_rmsnorm4_kernel<<<(unsigned)rows, 384, 0, st>>>(single, w, b, (__nv_bfloat16*)x, (__nv_bfloat16*)xg, C);
  else
    pool_rmsnorm_kernel<<<(unsigned)rows, 256, C * sizeof(float), st>>>(single, w, b, (__nv_bfloat16*)x, (__nv_bfloat16*)xg, C, pool);
  return check_launch("ag_pool_rmsnorm_fwd");
}

int launch_pair_combine(const float* qk, const float* relq, const float* relk, const float* Wp, const float* bp,
                        const float* outer, const float* prev, float* pair, const float* norm_w, const float* norm_b,
                        void* pn, int B, int P, int H, int D, int qk_ld, int rel_ld, int rows, int i_off, int device,
                        cudaStream_t st) {
  if (H != kPcH || D != kPcD) return set_error("ag_pair_combine_fwd: built for 32 pair heads and pair dim 128 (got %d, %d)", H, D);
  if (qk_ld < P || rel_ld < 2 * P) return set_error("ag_pair_combine_fwd: qk_ld/rel_ld too small");
  if (rows < 1 || i_off < 0 || i_off + rows > P) return set_error("ag_pair_combine_fwd: row range [%d, %d) outside P=%d", i_off, i_off + rows, P);
  if (pn && (!norm_w || !norm_b)) return set_error("ag_pair_combine_fwd: the normalised output needs norm_w and norm_b");
  if (B > 65535 || (rows + kPcI - 1) / kPcI > 65535) return set_error("ag_pair_combine_fwd: grid too large");
  static thread_local int attr_for = -1;
  if (attr_for != device) {
    cudaError_t e = cudaFuncSetAttribute(pair_combine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPcSmem);
    if (e != cudaSuccess) return set_error("ag_pair_combine_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_for = device;
  }
  dim3 grid((P + kPcJ - 1) / kPcJ, (rows + kPcI - 1) / kPcI, B);
  pair_combine_kernel<<<grid, 256, kPcSmem, st>>>(qk, relq, relk, Wp, bp, outer, prev, pair, norm_w, norm_b,
                                                  (__nv_bfloat16*)pn, P, qk_ld, rel_ld, rows, i_off);
  return check_launch("ag_pair_combine_fwd");
}

int launch_rmsnorm(const float* x, const float* w, const float* b, void* out, long rows, int C, int device,
                   cudaStream_t st) {
  if (C % 128) return set_error("ag_rmsnorm_fwd: C=%d must be a multiple of 128", C);
  rmsnorm_kernel<<<grid_for(rows * 32, 256, device), 256, 0, st>>>(x, w, b, (__nv_bfloat16*)out, rows, C);
  return check_launch("ag_rmsnorm_fwd");
}

int launch_add_embed_norm(const float* x, const float* emb, const float* scale, const float* shift, float* single,
                          void* a, int B, int rows, int C, int device, cudaStream_t st) {
  if (C % 4) return set_error("ag_add_embed_norm_fwd: C must be a multiple of 4");
  const long total = (long)B * rows * (C / 4);
  add_embed_norm_kernel<<<grid_for(total, 256, device), 256, 0, st>>>((const float4*)x, emb, scale, shift, (float4*)single,
                                                                      (uint2*)a, rows, C / 4, total);
  return check_launch("ag_add_embed_norm_fwd");
}

int launch_pair_out_embed(const float* pair, const float* pairT, const float* w, const float* b, const float* emb,
                          float* out, int B, int rows, int P, int C, int device, cudaStream_t st) {
  if (C != 128) return set_error("ag_pair_out_embed_fwd: built for pair dim 128 (got %d)", C);
  if (rows < 1 || P % rows) return set_error("ag_pair_out_embed_fwd: rows=%d must divide P=%d", rows, P);
  const long cells = (long)B * rows * P;
  pair_out_embed_kernel<<<grid_for(cells * 32, 256, device), 256, 0, st>>>(pair, pairT, w, b, emb, out, B, rows, P);
  return check_launch("ag_pair_out_embed_fwd");
}

}  // namespace agb
