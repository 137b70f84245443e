// K4/K6 — fused attention on tcgen05: S = Q K^T in TMEM, softmax in registers, P (bf16) staged in smem,
// O += P V accumulated in TMEM.  One kernel, two modes:
//
//   mode 0 (Attention.forward, AG:748-779): logits = 5*tanh((q.k/sqrt(d) + bias[b,h,i/16,j/16]) / 5).
//           The soft-cap bounds every logit to [-5, 5], so softmax needs no running max and O is never
//           rescaled: p = exp(logit), O = sum_j p_j v_j, out = O / sum_j p_j.
//   mode 1 (PairwiseRowAttention.forward, AG:871-889): plain softmax(q.k/sqrt(d)).  Two passes over the
//           keys: pass 1 computes only the exact row max (QK^T again is cheap: keys <= 512), pass 2 is the
//           same pipeline as mode 0 with p = exp(logit - max).  Epilogue adds to_v's bias (softmax rows
//           sum to 1, so the bias commutes with the weighted sum) and the residual, in fp32.
//
// One CTA per (batch, head, 128-query tile); 320 threads:
//   warp 0 (one lane)  TMA: Q once; per key block K [128 keys x 128] and V^T [dv x 128 keys], 2-stage rings
//   warp 1 (one lane)  tcgen05.mma issuer: S ping-pong (2 x 128 TMEM columns), O (dv columns)
//   warps 2-9          softmax: thread = query row, two warps per row quarter (one per 64-key half of a block);
//                      also the final O / rowsum epilogue
#include "common.cuh"
#include "kernels.h"

namespace agb {

constexpr int kAttD = 128;     // q/k head dim (both uses in the model)
constexpr int kAttBM = 128;    // queries per CTA
constexpr int kAttBN = 128;    // keys per block
constexpr int kAttMaxDv = 192;
constexpr int kAttQBytes = kAttBM * kAttD * 2;          // 32 KB (two 64-wide swizzle atoms)
constexpr int kAttKBytes = kAttBN * kAttD * 2;          // 32 KB
constexpr int kAttVBytesMax = kAttMaxDv * kAttBN * 2;   // 48 KB
constexpr int kAttPBytes = kAttBM * kAttBN * 2;         // 32 KB
constexpr int kAttThreads = 320;
constexpr int kAttSmem = 1024 + kAttQBytes + 2 * kAttKBytes + 2 * kAttVBytesMax + kAttPBytes + 256;

struct AttnKParams {
  CUtensorMap tmQ;   // (d, queries, heads, batch)
  CUtensorMap tmK;   // (d, keys, kv_heads, batch)
  CUtensorMap tmV;   // (keys, dv, kv_heads, batch)
  AgAttnArgs a;
  int q_tiles, nblk;
  unsigned int* err;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(kAttThreads, 1) attention_kernel(const __grid_constant__ AttnKParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const AgAttnArgs& a = p.a;

  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kAttQBytes;
  uint8_t* sV = sK + 2 * kAttKBytes;
  uint8_t* sP = sV + 2 * kAttVBytesMax;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + kAttPBytes);
  uint64_t* q_full = bars;         // [1]
  uint64_t* k_full = bars + 1;     // [2]
  uint64_t* k_empty = bars + 3;    // [2]
  uint64_t* v_full = bars + 5;     // [2]
  uint64_t* v_empty = bars + 7;    // [2]
  uint64_t* s_full = bars + 9;     // [2]
  uint64_t* s_empty = bars + 11;   // [2]
  uint64_t* p_full = bars + 13;    // [2] (key halves)
  uint64_t* p_empty = bars + 15;   // [2]
  uint64_t* o_full = bars + 17;    // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int qt = blockIdx.x % p.q_tiles;
  const int h = (blockIdx.x / p.q_tiles) % a.heads;
  const int b = blockIdx.x / (p.q_tiles * a.heads);
  const int hk = a.kv_heads == 1 ? 0 : h;
  const int q0 = qt * kAttBM;
  const int dv = a.dv;
  const int npre = a.mode == 1 ? p.nblk : 0;
  const int nsteps = npre + p.nblk;
  const uint32_t v_bytes = (uint32_t)dv * kAttBN * 2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 8);
      mbar_init(&p_full[i], 4);
      mbar_init(&p_empty[i], 1);
    }
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;         // 2 x 128 columns
  const uint32_t tmem_O = tmem_base + 256;   // dv columns

  if (warp == 0) {
    // ---------------------------------------------------------------------------- TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kAttQBytes);
      tma_load_4d(sQ, &p.tmQ, q_full, 0, q0, h, b);
      tma_load_4d(sQ + kAttQBytes / 2, &p.tmQ, q_full, 64, q0, h, b);
      for (int s = 0; s < nsteps; ++s) {
        const int kb = s < npre ? s : s - npre;
        const int st = s & 1;
        mbar_wait(&k_empty[st], ((s >> 1) & 1) ^ 1u, 0x110u + st, p.err);
        mbar_arrive_expect_tx(&k_full[st], kAttKBytes);
        tma_load_4d(sK + st * kAttKBytes, &p.tmK, &k_full[st], 0, kb * kAttBN, hk, b);
        tma_load_4d(sK + st * kAttKBytes + kAttKBytes / 2, &p.tmK, &k_full[st], 64, kb * kAttBN, hk, b);
        if (s >= npre) {
          const int f = s - npre;
          const int vs = f & 1;
          mbar_wait(&v_empty[vs], ((f >> 1) & 1) ^ 1u, 0x120u + vs, p.err);
          mbar_arrive_expect_tx(&v_full[vs], v_bytes);
          tma_load_4d(sV + vs * kAttVBytesMax, &p.tmV, &v_full[vs], kb * kAttBN, 0, hk, b);
          tma_load_4d(sV + vs * kAttVBytesMax + v_bytes / 2, &p.tmV, &v_full[vs], kb * kAttBN + 64, 0, hk, b);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(kAttBM, kAttBN);
      const uint32_t idesc_o = make_idesc_bf16(kAttBM, (uint32_t)dv);
      mbar_wait(q_full, 0, 0x100u, p.err);
      for (int s = 0; s <= nsteps; ++s) {
        if (s < nsteps) {
          const int st = s & 1;
          mbar_wait(&k_full[st], (s >> 1) & 1, 0x130u + st, p.err);
          mbar_wait(&s_empty[st], ((s >> 1) & 1) ^ 1u, 0x140u + st, p.err);
          tc_fence_after();
#pragma unroll
          for (int at = 0; at < 2; ++at) {
            const uint64_t dq = make_sw128_kmajor_desc(smem_u32(sQ + at * (kAttQBytes / 2)));
            const uint64_t dk = make_sw128_kmajor_desc(smem_u32(sK + st * kAttKBytes + at * (kAttKBytes / 2)));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_bf16_ss(tmem_S + (uint32_t)st * kAttBN, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_s,
                             (at > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit(&k_empty[st]);
          tc_commit(&s_full[st]);
        }
        if (s >= 1 && (s - 1) >= npre) {
          const int f = s - 1 - npre;
          const int vs = f & 1;
          mbar_wait(&v_full[vs], (f >> 1) & 1, 0x150u + vs, p.err);
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            mbar_wait(&p_full[hf], f & 1, 0x160u + hf, p.err);
            tc_fence_after();
            const uint64_t dp = make_sw128_kmajor_desc(smem_u32(sP + hf * (kAttPBytes / 2)));
            const uint64_t dvv = make_sw128_kmajor_desc(smem_u32(sV + vs * kAttVBytesMax + hf * (v_bytes / 2)));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_bf16_ss(tmem_O, dp + (uint64_t)(2 * k), dvv + (uint64_t)(2 * k), idesc_o,
                             (f > 0 || hf > 0 || k > 0) ? 1u : 0u);
            tc_commit(&p_empty[hf]);
          }
          tc_commit(&v_empty[vs]);
        }
      }
      tc_commit(o_full);
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------------------- softmax + epilogue
    // 8 warps: TMEM lane quarter qd = warp & 3 (thread = query row), key half hf = (warp - 2) >> 2: the warp
    // owns keys [64 hf, 64 hf + 64) of every 128-key block = one 128B-swizzled P atom column, so the two
    // halves never touch the same smem or barrier and each SM sub-partition has two softmax warps to overlap
    // MUFU latency.
    const int qd = warp & 3;
    const int hf = (warp - 2) >> 2;
    const int lrow = qd * 32 + lane;         // row within the tile
    const int row = q0 + lrow;               // query index
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const float LOG2E = 1.4426950408889634f;
    float rowsum = 0.0f;
    float rowmax = -INFINITY;
    const float sc_soft = a.scale / a.softcap;           // logits/softcap
    const float sc_exp = a.softcap * LOG2E;              // exp(softcap * tanh(.))
    const float sc_plain = a.scale * LOG2E;
    const float inv_cap = 1.0f / a.softcap;
    const int brow = (row < a.n_q ? row : a.n_q - 1) / 16;
    const int brows = a.bias_rows > 0 ? a.bias_rows : a.bias_p;
    const float* bias_row = a.mode == 0 ? a.bias + (((long)b * a.heads + h) * brows + brow) * a.bias_p : nullptr;
    uint8_t* prow = sP + hf * (kAttPBytes / 2) + lrow * 128;   // this row inside this half's swizzle atom column
    const int sw = lrow & 7;
    float* xch = reinterpret_cast<float*>(sP);   // [2][128] exchange between the halves (sP is idle at those points)
    float mneg = 0.0f;

    for (int s = 0; s < nsteps; ++s) {
      const int st = s & 1;
      const int kb = s < npre ? s : s - npre;
      const int key0 = kb * kAttBN + hf * 64;
      if (s == npre && npre > 0) {
        // mode 1: pass 1 done -> combine the two halves' row maxima before any exp (and before P is written)
        xch[hf * 128 + lrow] = rowmax;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        rowmax = fmaxf(rowmax, xch[(hf ^ 1) * 128 + lrow]);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        mneg = -rowmax * sc_plain;
      }
      mbar_wait(&s_full[st], (s >> 1) & 1, 0x170u + st, p.err);
      tc_fence_after();
      const uint32_t s_addr = tmem_S + lane_off + (uint32_t)st * kAttBN + (uint32_t)hf * 64;
      if (s < npre) {
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          uint32_t r[32];
          tmem_ld_32x32(s_addr + c2 * 32, r);
          tmem_ld_wait();
          if (key0 + c2 * 32 + 32 <= a.n_k) {
#pragma unroll
            for (int j = 0; j < 32; ++j) rowmax = fmaxf(rowmax, __uint_as_float(r[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (key0 + c2 * 32 + j < a.n_k) rowmax = fmaxf(rowmax, __uint_as_float(r[j]));
          }
        }
      } else {
        const int f = s - npre;
        // the four bias scalars (16 keys each) of this half-block, pre-divided by the soft-cap
        float bq[4] = {0.f, 0.f, 0.f, 0.f};
        if (a.mode == 0) {
          const int pc = kb * 8 + hf * 4;
#pragma unroll
          for (int t = 0; t < 4; ++t)
            if (pc + t < a.bias_p) bq[t] = __ldg(bias_row + pc + t) * inv_cap;
        }
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          uint32_t r[32];
          tmem_ld_32x32(s_addr + c2 * 32, r);
          tmem_ld_wait();
          if (c2 == 0) mbar_wait(&p_empty[hf], (f & 1) ^ 1u, 0x180u + hf, p.err);
          const bool full = key0 + c2 * 32 + 32 <= a.n_k;
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            float pv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int j = g8 * 8 + e;
              const float sv = __uint_as_float(r[j]);
              float pe;
              if (a.mode == 0) pe = ex2_approx(tanh_approx(fmaf(sv, sc_soft, bq[c2 * 2 + (j >> 4)])) * sc_exp);
              else pe = ex2_approx(fmaf(sv, sc_plain, mneg));
              if (!full && key0 + c2 * 32 + j >= a.n_k) pe = 0.0f;
              pv[e] = pe;
            }
            uint4 pk;
            pk.x = pack_bf16x2(pv[0], pv[1]);
            pk.y = pack_bf16x2(pv[2], pv[3]);
            pk.z = pack_bf16x2(pv[4], pv[5]);
            pk.w = pack_bf16x2(pv[6], pv[7]);
            // the normaliser sums exactly what the tensor core will multiply (bf16-rounded weights):
            // a bf16 is the top half of an fp32, so unpacking is a shift / mask
            rowsum += (__uint_as_float(pk.x << 16) + __uint_as_float(pk.x & 0xFFFF0000u)) +
                      (__uint_as_float(pk.y << 16) + __uint_as_float(pk.y & 0xFFFF0000u)) +
                      (__uint_as_float(pk.z << 16) + __uint_as_float(pk.z & 0xFFFF0000u)) +
                      (__uint_as_float(pk.w << 16) + __uint_as_float(pk.w & 0xFFFF0000u));
            const int chunk16 = c2 * 4 + g8;  // 16-byte chunk index inside the 64-key atom row
            *reinterpret_cast<uint4*>(prow + ((chunk16 ^ sw) << 4)) = pk;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[hf]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[st]);
    }

    // ---- epilogue: O / rowsum (+ bias + residual); each half writes dv/2 of the columns ----
    mbar_wait(o_full, 0, 0x190u, p.err);
    tc_fence_after();
    xch[hf * 128 + lrow] = rowsum;            // all MMAs are complete: sP is free
    asm volatile("bar.sync 1, 256;" ::: "memory");
    rowsum += xch[(hf ^ 1) * 128 + lrow];
    const float inv = 1.0f / rowsum;
    const long obase = (long)b * a.out_bstride + (long)h * a.out_hstride + (long)row * a.out_ld;
    const int cbeg = hf * (dv / 2), cend = cbeg + dv / 2;
#pragma unroll 1
    for (int c0 = cbeg; c0 < cend; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_O + lane_off + (uint32_t)c0, r);
      tmem_ld_wait();
      if (row < a.n_q) {
        if (a.out_dtype == 0) {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.out) + obase + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 pk;
            pk.x = pack_bf16x2(__uint_as_float(r[j]) * inv, __uint_as_float(r[j + 1]) * inv);
            pk.y = pack_bf16x2(__uint_as_float(r[j + 2]) * inv, __uint_as_float(r[j + 3]) * inv);
            pk.z = pack_bf16x2(__uint_as_float(r[j + 4]) * inv, __uint_as_float(r[j + 5]) * inv);
            pk.w = pack_bf16x2(__uint_as_float(r[j + 6]) * inv, __uint_as_float(r[j + 7]) * inv);
            *reinterpret_cast<uint4*>(o + j) = pk;
          }
        } else {
          float* o = reinterpret_cast<float*>(a.out) + obase + c0;
          const float* rs = a.res ? a.res + obase + c0 : nullptr;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 y;
            y.x = __uint_as_float(r[j]) * inv;
            y.y = __uint_as_float(r[j + 1]) * inv;
            y.z = __uint_as_float(r[j + 2]) * inv;
            y.w = __uint_as_float(r[j + 3]) * inv;
            if (a.out_bias) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(a.out_bias + c0 + j));
              y.x += bb.x; y.y += bb.y; y.z += bb.z; y.w += bb.w;
            }
            if (rs) {
              const float4 rr = *reinterpret_cast<const float4*>(rs + j);
              y.x += rr.x; y.y += rr.y; y.z += rr.z; y.w += rr.w;
            }
            *reinterpret_cast<float4*>(o + j) = y;
          }
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int launch_attention(const AgAttnArgs& a, int device, cudaStream_t stream) {
  if (!a.Q || !a.K || !a.Vt || !a.out) return set_error("ag_attention_fwd: Q, K, Vt and out must be non-null");
  if (a.d != kAttD) return set_error("ag_attention_fwd: head dim %d unsupported (kernel is built for d=128)", a.d);
  if (a.dv < 64 || a.dv > kAttMaxDv || a.dv % 64) return set_error("ag_attention_fwd: dv=%d must be a multiple of 64 in [64, 192]", a.dv);
  if (a.batch < 1 || a.heads < 1 || a.n_q < 1 || a.n_k < 1) return set_error("ag_attention_fwd: bad sizes");
  if (a.kv_heads != 1 && a.kv_heads != a.heads) return set_error("ag_attention_fwd: kv_heads must be 1 (MQA) or == heads");
  if (a.mode != 0 && a.mode != 1) return set_error("ag_attention_fwd: mode must be 0 (softcap+bias) or 1 (plain softmax)");
  if (a.mode == 0) {
    if (!a.bias || a.softcap <= 0.f) return set_error("ag_attention_fwd: mode 0 needs bias and softcap > 0");
    const int brows = a.bias_rows > 0 ? a.bias_rows : a.bias_p;
    if (a.bias_p * 16 < a.n_k || brows * 16 < a.n_q) return set_error("ag_attention_fwd: bias [%d x %d] must cover n_q/16 rows and n_k/16 cols", brows, a.bias_p);
  }
  if (a.out_dtype != 0 && a.out_dtype != 1) return set_error("ag_attention_fwd: bad out_dtype");
  if ((a.q_ld | a.k_ld | a.v_ld) % 8 || (a.q_hstride | a.k_hstride | a.v_hstride | a.q_bstride | a.k_bstride | a.v_bstride) % 8)
    return set_error("ag_attention_fwd: Q/K/Vt strides must be multiples of 8 elements");
  if ((a.out_ld | a.out_hstride | a.out_bstride) % 8) return set_error("ag_attention_fwd: out strides must be multiples of 8 elements");

  AttnKParams p;
  memset(&p, 0, sizeof(p));
  p.a = a;
  p.q_tiles = (a.n_q + kAttBM - 1) / kAttBM;
  p.nblk = (a.n_k + kAttBN - 1) / kAttBN;
  p.err = device_error_word(device);
  {
    uint64_t dims[4] = {(uint64_t)a.d, (uint64_t)a.n_q, (uint64_t)a.heads, (uint64_t)a.batch};
    uint64_t str[3] = {(uint64_t)a.q_ld * 2, (uint64_t)a.q_hstride * 2, (uint64_t)a.q_bstride * 2};
    uint32_t box[4] = {64, kAttBM, 1, 1};
    if (int rc = make_tensor_map_bf16(&p.tmQ, a.Q, 4, dims, str, box)) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)a.d, (uint64_t)a.n_k, (uint64_t)a.kv_heads, (uint64_t)a.batch};
    uint64_t str[3] = {(uint64_t)a.k_ld * 2, (uint64_t)a.k_hstride * 2, (uint64_t)a.k_bstride * 2};
    uint32_t box[4] = {64, kAttBN, 1, 1};
    if (int rc = make_tensor_map_bf16(&p.tmK, a.K, 4, dims, str, box)) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)a.n_k, (uint64_t)a.dv, (uint64_t)a.kv_heads, (uint64_t)a.batch};
    uint64_t str[3] = {(uint64_t)a.v_ld * 2, (uint64_t)a.v_hstride * 2, (uint64_t)a.v_bstride * 2};
    uint32_t box[4] = {64, (uint32_t)a.dv, 1, 1};
    if (int rc = make_tensor_map_bf16(&p.tmV, a.Vt, 4, dims, str, box)) return rc;
  }
  static thread_local int attr_set_for = -1;
  if (attr_set_for != device) {
    cudaError_t e = cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmem);
    if (e != cudaSuccess) return set_error("ag_attention_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set_for = device;
  }
  const long grid = (long)p.q_tiles * a.heads * a.batch;
  if (grid > 2147483647L) return set_error("ag_attention_fwd: grid too large");
  attention_kernel<<<(unsigned)grid, kAttThreads, kAttSmem, stream>>>(p);
  return check_launch("ag_attention_fwd");
}

}  // namespace agb
