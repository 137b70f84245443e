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
// to_attn_bias (AG:688-693).  One warp per pair cell: float4 per lane (C = 128), 8 dot products, butterfly.
// ---------------------------------------------------------------------------------------------------
__global__ void pair_bias_kernel(const float* __restrict__ pair, const float* __restrict__ scale,
                                 const float* __restrict__ shift, const float* __restrict__ W, float* __restrict__ bias,
                                 int B, int rows, int cols, int heads) {
  const int lane = threadIdx.x & 31;
  const long per_b = (long)rows * cols;
  const long cells = (long)B * per_b;
  const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + lane);
  const float4 sh = __ldg(reinterpret_cast<const float4*>(shift) + lane);
  for (long cell = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5; cell < cells;
       cell += ((long)gridDim.x * blockDim.x) >> 5) {
    float4 x = __ldg(reinterpret_cast<const float4*>(pair + cell * 128) + lane);
    x.x = act_apply(fmaf(x.x, sc.x, sh.x), ACT_GELU_ERF);
    x.y = act_apply(fmaf(x.y, sc.y, sh.y), ACT_GELU_ERF);
    x.z = act_apply(fmaf(x.z, sc.z, sh.z), ACT_GELU_ERF);
    x.w = act_apply(fmaf(x.w, sc.w, sh.w), ACT_GELU_ERF);
    const long b = cell / per_b;
    const long ij = cell - b * per_b;
    for (int g = 0; g < heads; ++g) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(W + g * 128) + lane);
      float d = x.x * w.x + x.y * w.y + x.z * w.z + x.w * w.w;
      d = warp_sum(d);
      if (lane == 0) bias[(b * heads + g) * per_b + ij] = d;
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

// ---------------------------------------------------------------------------------------------------
// SingleToPairwise combine (AG:835-856, relative_shift AG:523-530).  Block = (b, i, 32 consecutive j).
// Phase 1: 256 threads compute sim[j][h] (32 x 32) into smem; phase 2: thread = output channel.
// ---------------------------------------------------------------------------------------------------
__global__ void pair_combine_kernel(const float* __restrict__ qk, const float* __restrict__ relq,
                                    const float* __restrict__ relk, const float* __restrict__ Wp,
                                    const float* __restrict__ bp, const float* __restrict__ outer,
                                    const float* __restrict__ prev, float* __restrict__ pair, int P, int H, int D, int qk_ld,
                                    int rel_ld, int rows, int i_off) {
  // il = local row (this rank owns rows [i_off, i_off + rows) of the P x P pair tensor); i = global row
  extern __shared__ float sim_s[];  // [32][H + 1]
  const int jt = blockIdx.x, il = blockIdx.y, b = blockIdx.z;
  const int i = il + i_off;
  const int j0 = jt * 32;
  const int HS = H + 1;
  for (int idx = threadIdx.x; idx < 32 * H; idx += blockDim.x) {
    const int jj = idx & 31, h = idx >> 5;
    const int j = j0 + jj;
    float s = 0.f;
    if (j < P) {
      const long bh = (long)b * H + h;
      s = qk[(bh * rows + il) * qk_ld + j] +
          0.5f * (relq[(bh * rows + il) * (long)rel_ld + (P + j - i)] + relk[(bh * P + j) * (long)rel_ld + (P + i - j)]);
    }
    sim_s[jj * HS + h] = s;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float wrow[32];
#pragma unroll
    for (int h = 0; h < 32; ++h) wrow[h] = h < H ? __ldg(Wp + c * H + h) : 0.f;
    const float base = __ldg(bp + c) + outer[((long)b * P + i) * (2 * D) + c];
    for (int jj = 0; jj < 32; ++jj) {
      const int j = j0 + jj;
      if (j >= P) break;
      float acc = base + outer[((long)b * P + j) * (2 * D) + D + c];
#pragma unroll
      for (int h = 0; h < 32; ++h) acc = fmaf(wrow[h], sim_s[jj * HS + h], acc);
      const long o = (((long)b * rows + il) * P + j) * D + c;
      if (prev) acc += prev[o];
      pair[o] = acc;
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

int launch_transpose_bf16(const void* in, void* out, int nb, int R, int C, long in_bs, int in_ld, long out_bs,
                          int out_ld, int device, cudaStream_t st) {
  if (!in || !out) return set_error("ag_transpose_bf16: null pointer");
  if (nb > 65535) return set_error("ag_transpose_bf16: too many batches (%d)", nb);
  dim3 grid((C + 63) / 64, (R + 63) / 64, nb);
  transpose_bf16_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, R, C, in_bs, in_ld, out_bs, out_ld);
  return check_launch("ag_transpose_bf16");
}

int launch_pair_bias(const float* pair, const float* scale, const float* shift, const float* W, float* bias, int B,
                     int rows, int cols, int C, int heads, int device, cudaStream_t st) {
  if (C != 128) return set_error("ag_pair_bias_fwd: built for pair dim 128 (got %d)", C);
  const long cells = (long)B * rows * cols;
  pair_bias_kernel<<<grid_for(cells * 32, 256, device), 256, 0, st>>>(pair, scale, shift, W, bias, B, rows, cols, heads);
  return check_launch("ag_pair_bias_fwd");
}

int launch_pool_rmsnorm(const float* single, const float* w, const float* b, void* x, void* xg, int B, int n, int C,
                        int pool, int device, cudaStream_t st) {
  if (n % pool) return set_error("ag_pool_rmsnorm_fwd: n=%d not divisible by pool=%d", n, pool);
  const long rows = (long)B * (n / pool);
  pool_rmsnorm_kernel<<<(unsigned)rows, 256, C * sizeof(float), st>>>(single, w, b, (__nv_bfloat16*)x, (__nv_bfloat16*)xg, C, pool);
  return check_launch("ag_pool_rmsnorm_fwd");
}

int launch_pair_combine(const float* qk, const float* relq, const float* relk, const float* Wp, const float* bp,
                        const float* outer, const float* prev, float* pair, int B, int P, int H, int D, int qk_ld,
                        int rel_ld, int rows, int i_off, int device, cudaStream_t st) {
  if (H != 32) return set_error("ag_pair_combine_fwd: built for 32 pair heads (got %d)", H);
  if (qk_ld < P || rel_ld < 2 * P) return set_error("ag_pair_combine_fwd: qk_ld/rel_ld too small");
  if (rows < 1 || i_off < 0 || i_off + rows > P) return set_error("ag_pair_combine_fwd: row range [%d, %d) outside P=%d", i_off, i_off + rows, P);
  dim3 grid((P + 31) / 32, rows, B);
  pair_combine_kernel<<<grid, 128, 32 * (H + 1) * sizeof(float), st>>>(qk, relq, relk, Wp, bp, outer, prev, pair, P, H, D, qk_ld, rel_ld, rows, i_off);
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
