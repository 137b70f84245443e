// K4/K6 — fused attention on tcgen05 with both A operands in tensor memory:
//   S = Q K^T   (Q staged once into TMEM, K tiles streamed through smem by TMA)        -> fp32 S in TMEM
//   softmax in registers (thread = query row), P written back to TMEM as bf16 over the S columns it was read from
//   O += P V    (A = P from TMEM, V^T tiles streamed through smem by TMA)              -> fp32 O in TMEM
// so shared memory only carries the K and V streams (160 KB of smem traffic per 128x128 block instead of 328 KB:
// the first version, with Q and P in smem, was bound by the 128 B/clk smem port, not by MUFU or the tensor pipe).
// One kernel, two modes:
//
//   mode 0 (Attention.forward, AG:748-779): logits = 5*tanh((q.k/sqrt(d) + bias[b,h,i/16,j/16]) / 5).
//           The soft-cap bounds every logit to [-5, 5], so softmax needs no running max and O is never
//           rescaled: p = exp(logit), O = sum_j p_j v_j, out = O / sum_j p_j.
//   mode 1 (PairwiseRowAttention.forward, AG:871-889): plain softmax(q.k/sqrt(d)).  Two passes over the
//           keys: pass 1 computes only the exact row max (QK^T again is cheap: keys <= 512), pass 2 is the
//           same pipeline as mode 0 with p = exp(logit - max).  Epilogue adds to_v's bias (softmax rows
//           sum to 1, so the bias commutes with the weighted sum) and the residual, in fp32.
//
// One CTA per (batch, head, 128-query tile); 576 threads:
//   warp 0 (one lane)  TMA: per key block K [128 keys x 128] and V^T [dv x 128 keys], 2-stage rings
//   warp 1 (one lane)  tcgen05.mma issuer
//   warps 2-17         Q rows global -> registers -> TMEM (once); softmax: four warps per TMEM lane quarter, one
//                      per 32-key column quarter of the block; kPolyOf8 of every 8 exponentials are evaluated on
//                      the FMA pipe (Cody-Waite split + degree-3 polynomial, 7.5e-5 relative error, below bf16
//                      rounding) so MUFU carries 1.5 instead of 2 operations per logit; bias scalars are
//                      requested one block ahead; finally the O / rowsum epilogue.
// TMEM (512 columns): S ping-pong 2 x 128 | O 192 | Q 64.  P(s) aliases S(s): warp cq overwrites the first 16 of
// its own 32 S columns, so no cross-warp hazard exists; S(s+2) cannot be overwritten before PV(s) has read P(s)
// because tcgen05.mma executes in issue order.
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace agb {

constexpr int kAttD = 128;     // q/k head dim (both uses in the model)
constexpr int kAttBM = 128;    // queries per CTA
constexpr int kAttBN = 128;    // keys per block
constexpr int kAttMaxDv = 192;
constexpr int kAttKBytes = kAttBN * kAttD * 2;          // 32 KB (two 64-wide swizzle atoms)
constexpr int kAttVBytesMax = kAttMaxDv * kAttBN * 2;   // 48 KB
constexpr int kAttXchBytes = 4 * kAttBM * 4;            // row max / row sum exchange between column quarters
constexpr int kAttSmxWarps = 16;
constexpr int kAttThreads = 64 + kAttSmxWarps * 32;
// K ring: 4 stages.  Mode 0 streams 64+ key blocks through it (two more loads in flight than the first version's
// 2-stage ring).  Mode 1 (<= 512 keys = 4 blocks) loads every K block ONCE and keeps it resident for both passes.
constexpr int kAttKStages = 4;
// the dynamic smem window of a CTA without static smem starts 1024-aligned (checked at kernel entry)
constexpr int kAttSmem = kAttKStages * kAttKBytes + 2 * kAttVBytesMax + kAttXchBytes + 256;   // 231,680 B <= 227 KB
constexpr uint32_t kTmemO = 256, kTmemQ = 448;

struct AttnKParams {
  CUtensorMap tmK;   // (d, keys, kv_heads, batch)
  CUtensorMap tmV;   // (keys, dv, kv_heads, batch)
  AgAttnArgs a;
  int q_tiles, nblk;
  unsigned long long* trace;   // AGB_TRACE builds only
  int tile_off;        // first tile of this launch (a call may be issued as a full-wave launch + a key-split tail launch)
  unsigned int* err;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^(t*c) for |t*c| < 120 on the FMA/ALU pipes: n = round(t*c) via the 1.5*2^23 trick, f = t*c - n in [-0.5, 0.5],
// 2^f by a degree-3 minimax polynomial (max relative error 7.5e-5), exponent patched in with one shift-add.
__device__ __forceinline__ float ex2_poly(float t, float c) {
  const float magic = 12582912.0f;
  const float j = fmaf(t, c, magic);
  const float f = fmaf(t, c, -(j - magic));
  float p = fmaf(0.05517177f, f, 0.24261113f);
  p = fmaf(p, f, 0.69326097f);
  p = fmaf(p, f, 0.99992807f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(j) << 23));
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// kSplit = 2 (mode 0 only): a cluster of two CTAs shares one (batch, head, query tile); each takes half of the key
// blocks.  Because soft-capped logits need no running max, the halves combine by plain addition: the second CTA
// ships its unnormalised O (fp32) and row sums into the first CTA's (by then idle) K/V shared memory over DSMEM.
// Used when the grid would otherwise leave SMs idle (a sequence-parallel rank has only n_q / 128 x heads tiles) or
// to cut the last partial wave.
template <int kPolyOf8, int kSplit>
__global__ void __launch_bounds__(kAttThreads, 1) attention_kernel(const __grid_constant__ AttnKParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const AgAttnArgs& a = p.a;

  uint8_t* sK = smem;
  uint8_t* sV = sK + kAttKStages * kAttKBytes;
  float* xch = reinterpret_cast<float*>(sV + 2 * kAttVBytesMax);   // [4][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * kAttVBytesMax + kAttXchBytes);
  uint64_t* q_full = bars;         // [1]  Q rows are in TMEM (16 softmax warps arrive)
  uint64_t* k_full = bars + 1;     // [kAttKStages]
  uint64_t* k_empty = bars + 5;    // [kAttKStages]
  uint64_t* v_full = bars + 9;     // [2]
  uint64_t* v_empty = bars + 11;   // [2]
  uint64_t* s_full = bars + 13;    // [2]
  uint64_t* s_empty = bars + 15;   // [2]
  uint64_t* p_full = bars + 17;    // [2 S stages][2 key halves]
  uint64_t* o_full = bars + 21;    // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tile = p.tile_off + (kSplit == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x);
  const uint32_t crank = kSplit == 2 ? cluster_ctarank() : 0u;
  const int qt = tile % p.q_tiles;
  const int h = (tile / p.q_tiles) % a.heads;
  const int b = tile / (p.q_tiles * a.heads);
  const int hk = a.kv_heads == 1 ? 0 : h;
  const int q0 = qt * kAttBM;
  const int dv = a.dv;
  // this CTA's key blocks [kb0, kb0 + nblk)
  const int nblk_lo = kSplit == 2 ? (p.nblk + 1) / 2 : p.nblk;
  const int kb0 = crank == 0 ? 0 : nblk_lo;
  const int nblk = crank == 0 ? nblk_lo : p.nblk - nblk_lo;
  const int npre = a.mode == 1 ? nblk : 0;
  const int nsteps = npre + nblk;
  // mode 1 with at most kAttKStages key blocks: pass 2 re-reads the K tiles pass 1 left in shared memory
  const bool k_res = a.mode == 1 && nblk <= kAttKStages;
  const uint32_t v_bytes = (uint32_t)dv * kAttBN * 2;

  if (warp == 0 && lane == 0) {
    if (smem_u32(smem) & 1023u) {   // the swizzled tiles need a 1024-aligned window
      if (p.err) { p.err[0] = 0xDEAD0A11u; __threadfence_system(); }
      __trap();
    }
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    mbar_init(q_full, kAttSmxWarps);
    for (int i = 0; i < kAttKStages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], kAttSmxWarps);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&p_full[i], kAttSmxWarps / 2);
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
  pdl_launch_dependents();
  pdl_wait();   // set-up above overlaps the previous kernel's tail; global memory is only read from here on
  const uint32_t tmem_S = tmem_base;              // 2 x 128 columns (P aliases the S stage it was computed from)
  const uint32_t tmem_O = tmem_base + kTmemO;     // dv columns
  const uint32_t tmem_Q = tmem_base + kTmemQ;     // 64 columns: 128 bf16 per row

  if (warp == 0) {
    // ---------------------------------------------------------------------------- TMA producer
    if (lane == 0) {
      for (int s = 0; s < nsteps; ++s) {
        const int kb = kb0 + (s < npre ? s : s - npre);
        if (!(k_res && s >= npre)) {
          const int st = s & (kAttKStages - 1);
          mbar_wait(&k_empty[st], ((s / kAttKStages) & 1) ^ 1u, 0x110u + st, p.err);
          mbar_arrive_expect_tx(&k_full[st], kAttKBytes);
          tma_load_4d(sK + st * kAttKBytes, &p.tmK, &k_full[st], 0, kb * kAttBN, hk, b);
          tma_load_4d(sK + st * kAttKBytes + kAttKBytes / 2, &p.tmK, &k_full[st], 64, kb * kAttBN, hk, b);
        }
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
      tc_fence_after();
      for (int s = 0; s <= nsteps; ++s) {
        if (s < nsteps) {
          const int st = s & 1;                                                   // S stage
          const bool reuse = k_res && s >= npre;                                  // K tile already resident (pass 2)
          const int kst = k_res ? (s < npre ? s : s - npre) : (s & (kAttKStages - 1));   // K stage
          if (!reuse) mbar_wait(&k_full[kst], (s / kAttKStages) & 1, 0x130u + kst, p.err);
          mbar_wait(&s_empty[st], ((s >> 1) & 1) ^ 1u, 0x140u + st, p.err);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {   // 16 channels of d per MMA: 8 TMEM columns of Q, 32 B inside a K atom row
            const uint64_t dk = make_sw128_kmajor_desc(smem_u32(sK + kst * kAttKBytes + (kk >> 2) * (kAttKBytes / 2)));
            tc_mma_bf16_ts(tmem_S + (uint32_t)st * kAttBN, tmem_Q + (uint32_t)(8 * kk), dk + (uint64_t)(2 * (kk & 3)), idesc_s,
                           kk > 0 ? 1u : 0u);
          }
          if (!k_res) tc_commit(&k_empty[kst]);   // resident tiles are never refilled
          tc_commit(&s_full[st]);
        }
        if (s >= 1 && (s - 1) >= npre) {
          const int f = s - 1 - npre;
          const int vs = f & 1;
          const int pst = (s - 1) & 1;       // the S stage P(s-1) lives in
          mbar_wait(&v_full[vs], (f >> 1) & 1, 0x150u + vs, p.err);
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            // stages alternate, so (f >> 1) P tiles were published on this barrier before this one
            mbar_wait(&p_full[pst * 2 + hf], (f >> 1) & 1, 0x160u + pst * 2 + hf, p.err);
            tc_fence_after();
            const uint64_t dvv = make_sw128_kmajor_desc(smem_u32(sV + vs * kAttVBytesMax + hf * (v_bytes / 2)));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // keys 64 hf + 16 k .. +16: written by column-quarter warp cq = 2 hf + (k >> 1) into columns 32 cq + 8 (k & 1)
              const uint32_t pa = tmem_S + (uint32_t)pst * kAttBN + (uint32_t)(32 * (2 * hf + (k >> 1)) + 8 * (k & 1));
              tc_mma_bf16_ts(tmem_O, pa, dvv + (uint64_t)(2 * k), idesc_o, (f > 0 || hf > 0 || k > 0) ? 1u : 0u);
            }
          }
          tc_commit(&v_empty[vs]);
        }
      }
      tc_commit(o_full);
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------------------- softmax + epilogue
    // 16 warps: TMEM lane quarter qd = warp & 3 (thread = query row), column quarter cq = (warp - 2) >> 2: the
    // warp owns keys [32 cq, 32 cq + 32) of every 128-key block.
    const int qd = warp & 3;
    const int cq = (warp - 2) >> 2;
    const int hf = cq >> 1;
    const int lrow = qd * 32 + lane;         // row within the tile
    const int row = q0 + lrow;               // query index
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;

    // ---- Q: this thread's row, channels [32 cq, 32 cq + 32) -> 16 packed TMEM columns ----
    {
      uint32_t qr[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) qr[j] = 0u;
      if (row < a.n_q) {
        const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.Q) + (long)b * a.q_bstride +
                                                          (long)h * a.q_hstride + (long)row * a.q_ld + 32 * cq);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 v = __ldg(src + j);
          qr[4 * j] = v.x; qr[4 * j + 1] = v.y; qr[4 * j + 2] = v.z; qr[4 * j + 3] = v.w;
        }
      }
      tmem_st_32x16(tmem_Q + lane_off + (uint32_t)(16 * cq), qr);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(q_full);
    }

    const float LOG2E = 1.4426950408889634f;
    float rowmax = -INFINITY;
    float rsum[4] = {0.f, 0.f, 0.f, 0.f};   // partial row sums of the (unrounded) weights; rounding to bf16 is unbiased
    const float sc_soft = a.scale / a.softcap;           // logits/softcap
    const float sc_exp = a.softcap * LOG2E;              // exp(softcap * tanh(.))
    const float sc_plain = a.scale * LOG2E;
    const float inv_cap = 1.0f / a.softcap;
    const int brow = (row < a.n_q ? row : a.n_q - 1) / 16;
    const int brows = a.bias_rows > 0 ? a.bias_rows : a.bias_p;
    const float* bias_row = a.mode == 0 ? a.bias + (((long)b * a.heads + h) * brows + brow) * a.bias_p : nullptr;
    float mneg = 0.0f;

    // bias scalars of the NEXT key block are requested one step ahead, so their global-load latency (the largest
    // stall of the first version, ncu: long_scoreboard on the first use) hides behind a whole block of math
    float nb0 = 0.f, nb1 = 0.f;
    if (a.mode == 0 && nsteps > 0) {
      const int pc = kb0 * 8 + cq * 2;
      if (pc < a.bias_p) nb0 = __ldg(bias_row + pc);
      if (pc + 1 < a.bias_p) nb1 = __ldg(bias_row + pc + 1);
    }

    for (int s = 0; s < nsteps; ++s) {
      const int st = s & 1;
      const int kb = kb0 + (s < npre ? s : s - npre);
      const int key0 = kb * kAttBN + cq * 32;
      if (s == npre && npre > 0) {
        // mode 1: pass 1 done -> combine the four column quarters' row maxima before any exp
        xch[cq * 128 + lrow] = rowmax;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        rowmax = fmaxf(fmaxf(xch[lrow], xch[128 + lrow]), fmaxf(xch[256 + lrow], xch[384 + lrow]));
        asm volatile("bar.sync 1, 512;" ::: "memory");
        mneg = -rowmax * sc_plain;
      }
      const float bq0 = nb0 * inv_cap, bq1 = nb1 * inv_cap;   // this block's bias / softcap
      if (a.mode == 0 && s + 1 < nsteps) {
        const int pc = (kb + 1) * 8 + cq * 2;
        nb0 = pc < a.bias_p ? __ldg(bias_row + pc) : 0.f;
        nb1 = pc + 1 < a.bias_p ? __ldg(bias_row + pc + 1) : 0.f;
      }
      mbar_wait(&s_full[st], (s >> 1) & 1, 0x170u + st, p.err);
      tc_fence_after();
      const uint32_t s_addr = tmem_S + lane_off + (uint32_t)st * kAttBN + (uint32_t)cq * 32;
      uint32_t r[32];
      tmem_ld_32x32(s_addr, r);
      tmem_ld_wait();
      const bool full = key0 + 32 <= a.n_k;
      if (s < npre) {
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; ++j) rowmax = fmaxf(rowmax, __uint_as_float(r[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (key0 + j < a.n_k) rowmax = fmaxf(rowmax, __uint_as_float(r[j]));
        }
      } else {
        float pv[32];
        if (a.mode == 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float t = tanh_approx(fmaf(__uint_as_float(r[j]), sc_soft, j < 16 ? bq0 : bq1));
            // interleave the two exponential flavours so the MUFU and FMA pipes both stay busy
            pv[j] = ((j & 7) * kPolyOf8 % 8 + kPolyOf8 > 7) ? ex2_poly(t, sc_exp) : ex2_approx(t * sc_exp);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) pv[j] = ex2_approx(fmaf(__uint_as_float(r[j]), sc_plain, mneg));
        }
        if (!full) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (key0 + j >= a.n_k) pv[j] = 0.0f;   // keys past the end contribute nothing (also to the row sum)
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) rsum[j & 3] += pv[j];
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(pv[2 * j], pv[2 * j + 1]);
        tmem_st_32x16(s_addr, pk);   // P over the first half of the S columns this thread just read
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[st * 2 + hf]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[st]);
    }

    // ---- epilogue: O / rowsum (+ bias + residual); each column quarter writes dv/4 of the columns ----
    // Row-attention epilogue (fp32 out + residual + RMSNorm): this thread's 32 residual values are requested NOW, so the
    // eight 16-byte loads (rows are 512 B apart: one sector each) are in flight while the last P.V MMAs drain, instead
    // of eight exposed round trips inside the store loop (ncu: 40 % of the kernel's stall samples sat on those adds).
    float4 rpre[8];
    const bool pre_res = a.norm_out != nullptr && a.res != nullptr && row < a.n_q;
    if (pre_res) {
      const float* rs = a.res + ((long)b * a.out_bstride + (long)h * a.out_hstride + (long)row * a.out_ld) + cq * (dv / 4);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(rpre[j].x), "=f"(rpre[j].y), "=f"(rpre[j].z), "=f"(rpre[j].w) : "l"(rs + 4 * j));
    }
    xch[cq * 128 + lrow] = (rsum[0] + rsum[1]) + (rsum[2] + rsum[3]);
    asm volatile("bar.sync 1, 512;" ::: "memory");
    float rtot = (xch[lrow] + xch[128 + lrow]) + (xch[256 + lrow] + xch[384 + lrow]);
    mbar_wait(o_full, 0, 0x190u, p.err);
    tc_fence_after();
    float4* land = reinterpret_cast<float4*>(smem);           // [dv/16 float4 slots][512 threads] + row sums
    const int tsoft = threadIdx.x - 64;
    if constexpr (kSplit == 2) {
      cluster_sync_all();                                      // #1: both CTAs' MMAs are complete, K/V smem is idle
      if (crank == 1) {
        const uint32_t land_remote = mapa_u32(smem_u32(land), 0u);
#pragma unroll 1
        for (int c16 = 0; c16 < dv / 64; ++c16) {
          uint32_t r[16];
          tmem_ld_32x16(tmem_O + lane_off + (uint32_t)(cq * (dv / 4) + 16 * c16), r);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q)
            st_cluster_v4(land_remote + (uint32_t)(((c16 * 4 + q) * 512 + tsoft) * 16), r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
        }
        if (cq == 0) st_cluster_f32(land_remote + (uint32_t)((dv / 16) * 512 * 16 + lrow * 4), rtot);
      }
      cluster_sync_all();                                      // #2: the partner's partial results have landed
      if (crank == 0) rtot += reinterpret_cast<const float*>(land + (dv / 16) * 512)[lrow];
    }
    if (crank == 0) {   // (kSplit == 1: always) this CTA writes the tile
    const float inv = 1.0f / rtot;
    const long obase = (long)b * a.out_bstride + (long)h * a.out_hstride + (long)row * a.out_ld;
    const int cbeg = cq * (dv / 4), cend = cbeg + dv / 4;
    if (a.norm_out != nullptr) {
      // fp32 out + bias + residual, and RMSNormWithBias of the finished row in bf16 (dv == 128: this thread holds
      // 32 of the row's 128 channels; the sum of squares is exchanged between the four column quarters)
      float y[32];
#pragma unroll
      for (int hc = 0; hc < 2; ++hc) {
        uint32_t r[16];
        tmem_ld_32x16(tmem_O + lane_off + (uint32_t)(cbeg + 16 * hc), r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) y[16 * hc + j] = __uint_as_float(r[j]) * inv;
      }
      float ss = 0.f;
      if (row < a.n_q) {
        float* o = reinterpret_cast<float*>(a.out) + obase + cbeg;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (a.out_bias) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(a.out_bias + cbeg + j));
            y[j] += bb.x; y[j + 1] += bb.y; y[j + 2] += bb.z; y[j + 3] += bb.w;
          }
          if (pre_res) {
            const float4 rr = rpre[j >> 2];
            y[j] += rr.x; y[j + 1] += rr.y; y[j + 2] += rr.z; y[j + 3] += rr.w;
          }
          *reinterpret_cast<float4*>(o + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
          ss += y[j] * y[j] + y[j + 1] * y[j + 1] + y[j + 2] * y[j + 2] + y[j + 3] * y[j + 3];
        }
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");   // every thread has read the row sums from xch
      xch[cq * 128 + lrow] = ss;
      asm volatile("bar.sync 1, 512;" ::: "memory");
      if (row < a.n_q) {
        const float rinv = rsqrtf(((xch[lrow] + xch[128 + lrow]) + (xch[256 + lrow] + xch[384 + lrow])) * (1.0f / 128.0f) + 1e-5f);
        __nv_bfloat16* no = reinterpret_cast<__nv_bfloat16*>(a.norm_out) + obase + cbeg;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(a.norm_w + cbeg + j));
          const float4 w1 = __ldg(reinterpret_cast<const float4*>(a.norm_w + cbeg + j + 4));
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.norm_b + cbeg + j));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.norm_b + cbeg + j + 4));
          uint4 pk;
          pk.x = pack_bf16x2(fmaf(y[j] * rinv, w0.x, b0.x), fmaf(y[j + 1] * rinv, w0.y, b0.y));
          pk.y = pack_bf16x2(fmaf(y[j + 2] * rinv, w0.z, b0.z), fmaf(y[j + 3] * rinv, w0.w, b0.w));
          pk.z = pack_bf16x2(fmaf(y[j + 4] * rinv, w1.x, b1.x), fmaf(y[j + 5] * rinv, w1.y, b1.y));
          pk.w = pack_bf16x2(fmaf(y[j + 6] * rinv, w1.z, b1.z), fmaf(y[j + 7] * rinv, w1.w, b1.w));
          *reinterpret_cast<uint4*>(no + j) = pk;
        }
      }
    } else {
#pragma unroll 1
    for (int c0 = cbeg; c0 < cend; c0 += 16) {
      uint32_t r[16];
      tmem_ld_32x16(tmem_O + lane_off + (uint32_t)c0, r);
      tmem_ld_wait();
      if constexpr (kSplit == 2) {
        const int c16 = (c0 - cbeg) >> 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 o2 = land[(c16 * 4 + q) * 512 + tsoft];
          r[4 * q] = __float_as_uint(__uint_as_float(r[4 * q]) + o2.x);
          r[4 * q + 1] = __float_as_uint(__uint_as_float(r[4 * q + 1]) + o2.y);
          r[4 * q + 2] = __float_as_uint(__uint_as_float(r[4 * q + 2]) + o2.z);
          r[4 * q + 3] = __float_as_uint(__uint_as_float(r[4 * q + 3]) + o2.w);
        }
      }
      if (row < a.n_q) {
        if (a.out_dtype == 0) {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.out) + obase + c0;
#pragma unroll
          for (int j = 0; j < 16; j += 8) {
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
          for (int j = 0; j < 16; j += 4) {
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
    }
    }
    tc_fence_before();
  }
  if constexpr (kSplit == 2) {
    if (warp < 2) {          // the TMA / MMA warps take part in the two cluster barriers of the combine
      cluster_sync_all();
      cluster_sync_all();
    }
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ----------------------------------------------------------------------------------------------------
// Persistent row attention (mode 1 as the trunk calls it: dv = 128, fp32 out + bias + residual + the pair feed-forward's
// RMSNorm).  The kernel above runs one 128-query tile per CTA with TWO passes over the keys; a pair block at 1 Mbp is 2048
// such tiles of 8 short steps each, i.e. 13.8 waves of CTAs whose launch, set-up, Q / K latencies and epilogue are all
// exposed: 0.29 ms per block alone (0.47 ms in the power-capped step) against a ~0.08 ms HBM floor.  What the r3
// timing experiments and the SM-clock timeline (-DAGB_TRACE build) showed, in the order it was fixed:
//   1. the epilogue's global traffic was half of the launch: TMEM hands a thread ONE ROW and rows are 512 B apart, so a
//      warp instruction touched 32 lines x 16 B.  Now the four warps that share 32 rows exchange the finished tile
//      through a swizzled staging buffer and every global access is 4 rows x 128 contiguous bytes (264 -> 189 us);
//   2. a persistent CTA per SM with every role running ahead across tile boundaries (Q double-buffered in TMEM: dv = 128
//      leaves 64 spare columns) hides set-up and load latency, but alone it bought only 290 -> 264 us, because
//   3. each step is paced by the TMEM READ port: 16 warps x tcgen05.ld.32x32b.x32 = 64 KB per 128-key block at
//      64 B/clk/SM = 1024 clk, and the two-pass scheme read S twice (row max, then exp) on top of computing it twice.
//      This version is SINGLE-PASS: an online softmax with a running row maximum, and O is rescaled lazily — only when
//      the maximum grows by more than 2^8 (the stale maximum just scales P and the row sum by the same factor, which
//      the final division removes; bf16 / fp32 carry 2^8 without loss).  With four key blocks of comparable logits the
//      rescale (a TMEM read-modify-write of O behind the previous P.V) practically never runs.  S is computed and read
//      once, K streams through a ring instead of staying resident (any number of keys), and the freed shared memory
//      holds the whole 128 x 128 fp32 tile for the epilogue exchange (one barrier instead of eight).
// Roles: warp 0 TMA (K ring 2, V^T ring 2), warp 1 MMA issuer, warps 2-17 softmax + epilogue (lane quarter qd = warp & 3,
// key quarter cq).  Barrier phases are carried in running counters.
// ----------------------------------------------------------------------------------------------------
constexpr int kRowVBytes = 128 * kAttBN * 2;   // dv = 128: 32 KB per V^T block
constexpr int kRowKStages = 2;
// Bring-up timeline (only in a -DAGB_TRACE build, python -m alphagenome_b200.csrc.build --trace): CTA 0's TMA thread, MMA
// thread and one softmax thread stamp (tag, SM clock) pairs into a global buffer that AGB_ROWATTN_TRACE=<file> dumps.
#ifdef AGB_TRACE
#define ROW_TRACE(role, tag)                                                                                      \
  do {                                                                                                            \
    if (p.trace && blockIdx.x == 0 && trc < 2040) p.trace[(role) * 2048 + trc++] = ((unsigned long long)(tag) << 48) | (unsigned long long)(clock64() & 0xFFFFFFFFFFFFull); \
  } while (0)
#else
#define ROW_TRACE(role, tag) do {} while (0)
#endif
constexpr int kRowStgBytes = kAttBM * 128 * 4;    // epilogue staging: the whole tile, 128 rows x 128 fp32 channels, swizzled
constexpr int kRowParBytes = 3 * 128 * 4;         // out_bias | norm_w | norm_b
constexpr int kRowSmem = kRowKStages * kAttKBytes + 2 * kRowVBytes + 3 * kAttXchBytes + kRowStgBytes + kRowParBytes + 256;   // 204,544 B
constexpr uint32_t kTmemQ1 = 384;   // second Q buffer (tile parity 1); parity 0 uses kTmemQ
constexpr float kRescaleLog2 = 8.0f;   // O is rescaled only when the running maximum grows by more than 2^8

__global__ void __launch_bounds__(kAttThreads, 1) row_attention_kernel(const __grid_constant__ AttnKParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const AgAttnArgs& a = p.a;
  constexpr int dv = 128;

  uint8_t* sK = smem;
  uint8_t* sV = sK + kRowKStages * kAttKBytes;
  uint8_t* stg_all = sV + 2 * kRowVBytes;
  float* xmax = reinterpret_cast<float*>(stg_all + kRowStgBytes);   // [2 (block parity)][4][128]
  float* xsum = xmax + 2 * 4 * kAttBM;                              // [4][128]
  float* spar = xsum + 4 * kAttBM;                                  // out_bias | norm_w | norm_b
  uint64_t* bars = reinterpret_cast<uint64_t*>(spar + 3 * 128);
  uint64_t* q_full = bars;         // [2]  Q rows of tile parity 0 / 1 are in TMEM (16 softmax warps arrive)
  uint64_t* k_full = bars + 2;     // [2]
  uint64_t* k_empty = bars + 4;    // [2]
  uint64_t* v_full = bars + 6;     // [2]
  uint64_t* v_empty = bars + 8;    // [2]
  uint64_t* s_full = bars + 10;    // [2]
  uint64_t* s_empty = bars + 12;   // [2]
  uint64_t* p_full = bars + 14;    // [2 S stages][2 key halves]
  uint64_t* o_full = bars + 18;    // [1]  the tile's last P.V has completed
  uint64_t* o_empty = bars + 19;   // [1]  the epilogue has O in registers (16 softmax warps arrive)
  uint64_t* pv_done = bars + 20;   // [1]  one phase per P.V product (a rescale of O waits for the product before it)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total = p.q_tiles * a.heads * a.batch;
  const int nblk = p.nblk;          // key blocks per tile, the same for every tile

  if (warp == 0 && lane == 0) {
    if (smem_u32(smem) & 1023u) {
      if (p.err) { p.err[0] = 0xDEAD0A12u; __threadfence_system(); }
      __trap();
    }
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], kAttSmxWarps);
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], kAttSmxWarps);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&p_full[i], kAttSmxWarps / 2);
    mbar_init(o_full, 1);
    mbar_init(o_empty, kAttSmxWarps);
    mbar_init(pv_done, 1);
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
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + kTmemO;

  if (warp == 0) {
    // ---------------------------------------------------------------------------- TMA producer
    if (lane == 0) {
      uint32_t gk = 0;   // key blocks loaded so far (K and V^T advance together)
      int trc = 0; (void)trc;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int h = (tile / p.q_tiles) % a.heads;
        const int b = tile / (p.q_tiles * a.heads);
        const int hk = a.kv_heads == 1 ? 0 : h;
        for (int j = 0; j < nblk; ++j, ++gk) {
          const int st = gk & 1;
          const uint32_t ph = ((gk >> 1) & 1u) ^ 1u;
          mbar_wait(&k_empty[st], ph, 0x210u + st, p.err);
          ROW_TRACE(0, 0x10 + j);
          mbar_arrive_expect_tx(&k_full[st], kAttKBytes);
          tma_load_4d(sK + st * kAttKBytes, &p.tmK, &k_full[st], 0, j * kAttBN, hk, b);
          tma_load_4d(sK + st * kAttKBytes + kAttKBytes / 2, &p.tmK, &k_full[st], 64, j * kAttBN, hk, b);
          mbar_wait(&v_empty[st], ph, 0x220u + st, p.err);
          ROW_TRACE(0, 0x20 + j);
          mbar_arrive_expect_tx(&v_full[st], kRowVBytes);
          tma_load_4d(sV + st * kRowVBytes, &p.tmV, &v_full[st], j * kAttBN, 0, hk, b);
          tma_load_4d(sV + st * kRowVBytes + kRowVBytes / 2, &p.tmV, &v_full[st], j * kAttBN + 64, 0, hk, b);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(kAttBM, kAttBN);
      const uint32_t idesc_o = make_idesc_bf16(kAttBM, (uint32_t)dv);
      uint32_t g = 0;    // QK^T products issued so far: S stage and K stage = g & 1
      uint32_t gv = 0;   // P.V products issued so far: V stage = gv & 1 (the S stage its P lives in is also gv & 1)
      int it = 0;
      int trc = 0; (void)trc;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
        const uint32_t tmem_Q = tmem_base + ((it & 1) ? kTmemQ1 : kTmemQ);
        ROW_TRACE(1, 0x01);
        mbar_wait(&q_full[it & 1], (uint32_t)(it >> 1) & 1u, 0x200u + (it & 1), p.err);
        tc_fence_after();
        ROW_TRACE(1, 0x02);
        for (int s = 0; s <= nblk; ++s) {
          if (s < nblk) {
            const int st = g & 1;
            mbar_wait(&k_full[st], (g >> 1) & 1u, 0x230u + st, p.err);
            mbar_wait(&s_empty[st], ((g >> 1) & 1u) ^ 1u, 0x240u + st, p.err);
            tc_fence_after();
            ROW_TRACE(1, 0x10 + s);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              const uint64_t dk = make_sw128_kmajor_desc(smem_u32(sK + st * kAttKBytes + (kk >> 2) * (kAttKBytes / 2)));
              tc_mma_bf16_ts(tmem_S + (uint32_t)st * kAttBN, tmem_Q + (uint32_t)(8 * kk), dk + (uint64_t)(2 * (kk & 3)), idesc_s,
                             kk > 0 ? 1u : 0u);
            }
            tc_commit(&k_empty[st]);
            tc_commit(&s_full[st]);
            ++g;
          }
          if (s >= 1) {
            const int f = s - 1;
            const int vs = gv & 1;
            mbar_wait(&v_full[vs], (gv >> 1) & 1u, 0x250u + vs, p.err);
            if (f == 0) {   // O still holds the previous tile until its epilogue has it in registers
              mbar_wait(o_empty, (uint32_t)(it & 1) ^ 1u, 0x280u, p.err);
            }
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              mbar_wait(&p_full[vs * 2 + hf], (gv >> 1) & 1u, 0x260u + vs * 2 + hf, p.err);
              tc_fence_after();
              const uint64_t dvv = make_sw128_kmajor_desc(smem_u32(sV + vs * kRowVBytes + hf * (kRowVBytes / 2)));
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t pa = tmem_S + (uint32_t)vs * kAttBN + (uint32_t)(32 * (2 * hf + (k >> 1)) + 8 * (k & 1));
                tc_mma_bf16_ts(tmem_O, pa, dvv + (uint64_t)(2 * k), idesc_o, (f > 0 || hf > 0 || k > 0) ? 1u : 0u);
              }
            }
            ROW_TRACE(1, 0x20 + f);
            tc_commit(&v_empty[vs]);
            tc_commit(pv_done);
            ++gv;
          }
        }
        tc_commit(o_full);
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------------------- softmax + epilogue
    const int qd = warp & 3;
    const int cq = (warp - 2) >> 2;
    const int hf = cq >> 1;
    const int lrow = qd * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const float LOG2E = 1.4426950408889634f;
    const float sc_plain = a.scale * LOG2E;
    const int rr = lane >> 3, cg = lane & 7;   // epilogue layout: rows 8 cq + rr (+ 4) of the group, channels 4 cg .. 4 cg + 3 of a 32-wide chunk
    uint8_t* stg = stg_all + qd * (32 * 512);   // this group's 32 rows x 512 B
    // the four warps of a 32-row group (same qd) only ever exchange data with each other: one named barrier per group
    auto group_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(2 + qd) : "memory"); };

    {   // per-channel epilogue parameters: once per kernel into shared memory
      const int t = threadIdx.x - 64;
      if (t < 128) {
        spar[t] = a.out_bias ? __ldg(a.out_bias + t) : 0.0f;
        spar[128 + t] = __ldg(a.norm_w + t);
        spar[256 + t] = __ldg(a.norm_b + t);
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");
    }

    // Q rows of a tile: this thread's row, channels [32 cq, 32 cq + 32) = 16 packed TMEM columns
    auto q_load = [&](int tile, uint32_t (&qr)[16]) {
      const int qt = tile % p.q_tiles;
      const int h = (tile / p.q_tiles) % a.heads;
      const int b = tile / (p.q_tiles * a.heads);
      const int row = qt * kAttBM + lrow;
#pragma unroll
      for (int j = 0; j < 16; ++j) qr[j] = 0u;
      if (row < a.n_q) {
        const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.Q) + (long)b * a.q_bstride +
                                                          (long)h * a.q_hstride + (long)row * a.q_ld + 32 * cq);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 v = __ldg(src + j);
          qr[4 * j] = v.x; qr[4 * j + 1] = v.y; qr[4 * j + 2] = v.z; qr[4 * j + 3] = v.w;
        }
      }
    };
    auto q_publish = [&](int parity, const uint32_t (&qr)[16]) {
      tmem_st_32x16(tmem_base + (parity ? kTmemQ1 : kTmemQ) + lane_off + (uint32_t)(16 * cq), qr);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&q_full[parity]);
    };
    if ((int)blockIdx.x < total) {
      uint32_t qr[16];
      q_load(blockIdx.x, qr);
      q_publish(0, qr);
    }

    uint32_t g = 0;    // key blocks processed so far (all tiles): S stage = g & 1, the P.V product of block g is product g
    int it = 0;
    int trc = (warp == 2 && lane == 0) ? 0 : 4096; (void)trc;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const int qt = tile % p.q_tiles;
      const int h = (tile / p.q_tiles) % a.heads;
      const int b = tile / (p.q_tiles * a.heads);
      const int next = tile + gridDim.x;
      ROW_TRACE(2, 0x01);
      uint32_t qn[16];
      if (next < total) q_load(next, qn);   // in flight during the first key block

      float m_run = 0.0f;                      // running row maximum (raw logits), set by the first block
      float rsum[4] = {0.f, 0.f, 0.f, 0.f};    // partial row sums of the weights, in units of exp(- m_run * scale)
      for (int s = 0; s < nblk; ++s, ++g) {
        const int st = g & 1;
        const int key0 = s * kAttBN + cq * 32;
        mbar_wait(&s_full[st], (g >> 1) & 1u, 0x270u + st, p.err);
        tc_fence_after();
        ROW_TRACE(2, 0x10 + s);
        const uint32_t s_addr = tmem_S + lane_off + (uint32_t)st * kAttBN + (uint32_t)cq * 32;
        uint32_t r[32];
        tmem_ld_32x32(s_addr, r);
        tmem_ld_wait();
        const bool full = key0 + 32 <= a.n_k;
        float bmax = -INFINITY;
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; ++j) bmax = fmaxf(bmax, __uint_as_float(r[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (key0 + j < a.n_k) bmax = fmaxf(bmax, __uint_as_float(r[j]));
        }
        // block maximum of the row: the four key quarters exchange theirs (buffers alternate, one barrier per block)
        float* xm = xmax + (s & 1) * 4 * kAttBM;
        xm[cq * kAttBM + lrow] = bmax;
        group_sync();
        bmax = fmaxf(fmaxf(xm[lrow], xm[kAttBM + lrow]), fmaxf(xm[2 * kAttBM + lrow], xm[3 * kAttBM + lrow]));
        if (s == 0) {
          m_run = bmax;
          if (next < total) q_publish((it + 1) & 1, qn);   // that Q buffer was last read by the tile before this one
        } else {
          // lazy rescale: all four warps of the group see the same (m_run, bmax), so they take the same branch per row
          const bool grow = (bmax - m_run) * sc_plain > kRescaleLog2;
          if (__any_sync(0xffffffffu, grow)) {
            const float fac = grow ? ex2_approx((m_run - bmax) * sc_plain) : 1.0f;
            if (grow) m_run = bmax;
            rsum[0] *= fac; rsum[1] *= fac; rsum[2] *= fac; rsum[3] *= fac;
            // O holds the products of blocks 0 .. s-1: product g-1 must have completed before it is scaled in place
            mbar_wait(pv_done, (g - 1) & 1u, 0x2A0u, p.err);
            tc_fence_after();
            uint32_t o[32];
            tmem_ld_32x32(tmem_O + lane_off + (uint32_t)(cq * 32), o);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * fac);
            tmem_st_32x16(tmem_O + lane_off + (uint32_t)(cq * 32), reinterpret_cast<uint32_t(&)[16]>(o[0]));
            tmem_st_32x16(tmem_O + lane_off + (uint32_t)(cq * 32 + 16), reinterpret_cast<uint32_t(&)[16]>(o[16]));
            tmem_st_wait();   // ordered before the P.V of this block by the fence + p_full arrive below
          }
        }
        const float mneg = -m_run * sc_plain;
        float pv[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) pv[j] = ex2_approx(fmaf(__uint_as_float(r[j]), sc_plain, mneg));
        if (!full) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (key0 + j >= a.n_k) pv[j] = 0.0f;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) rsum[j & 3] += pv[j];
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(pv[2 * j], pv[2 * j + 1]);
        tmem_st_32x16(s_addr, pk);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&p_full[st * 2 + hf]);
          mbar_arrive(&s_empty[st]);
        }
        ROW_TRACE(2, 0x30 + s);
      }

      // ---- epilogue: O / rowsum + to_v bias + residual (fp32), RMSNormWithBias of the finished row (bf16) ----
      // TMEM hands a thread ONE ROW and rows are 512 B apart in global memory, so the four warps that share 32 rows exchange
      // the tile through the swizzled staging buffer: warp cq then owns rows 8 cq .. 8 cq + 7 of the group with lanes across
      // the channels, and every global access is 4 rows x 128 contiguous bytes.
      const int r0 = qt * kAttBM + qd * 32 + cq * 8 + rr;       // this lane's rows: r0 and r0 + 4
      const long ob = (long)b * a.out_bstride + (long)h * a.out_hstride;
      const bool ok0 = r0 < a.n_q, ok1 = r0 + 4 < a.n_q;
      // y starts as the residual (requested now, in flight while the last P.V drains) and ends as the finished row
      float y[32];      // [chunk][row 0: 4 channels | row 1: 4 channels]
#pragma unroll
      for (int j = 0; j < 32; ++j) y[j] = 0.f;
      if (a.res != nullptr) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (ok0) asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(y[8 * c]), "=f"(y[8 * c + 1]), "=f"(y[8 * c + 2]), "=f"(y[8 * c + 3])
                                : "l"(a.res + ob + (long)r0 * a.out_ld + c * 32 + cg * 4));
          if (ok1) asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(y[8 * c + 4]), "=f"(y[8 * c + 5]), "=f"(y[8 * c + 6]), "=f"(y[8 * c + 7])
                                : "l"(a.res + ob + (long)(r0 + 4) * a.out_ld + c * 32 + cg * 4));
        }
      }
      xsum[cq * kAttBM + lrow] = (rsum[0] + rsum[1]) + (rsum[2] + rsum[3]);
      group_sync();
      const float inv = 1.0f / ((xsum[lrow] + xsum[kAttBM + lrow]) + (xsum[2 * kAttBM + lrow] + xsum[3 * kAttBM + lrow]));
      ROW_TRACE(2, 0x05);
      mbar_wait(o_full, (uint32_t)(it & 1), 0x290u, p.err);
      tc_fence_after();
      ROW_TRACE(2, 0x06);
      {
        uint32_t o[32];
        tmem_ld_32x32(tmem_O + lane_off + (uint32_t)(cq * 32), o);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_empty);   // O is in registers: the next tile's P.V may overwrite it
        // element (row, 16-byte unit u) of the group's tile lives at unit u ^ (row & 7): conflict-free both ways
#pragma unroll
        for (int u = 0; u < 8; ++u)
          *reinterpret_cast<float4*>(stg + lane * 512 + (((cq * 8 + u) ^ (lane & 7)) << 4)) =
              make_float4(__uint_as_float(o[4 * u]) * inv, __uint_as_float(o[4 * u + 1]) * inv, __uint_as_float(o[4 * u + 2]) * inv,
                          __uint_as_float(o[4 * u + 3]) * inv);
      }
      group_sync();
      float ss0 = 0.f, ss1 = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 bb = *reinterpret_cast<const float4*>(spar + c * 32 + cg * 4);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int gr = cq * 8 + rr + 4 * i;     // row inside the group
          float4 v = *reinterpret_cast<const float4*>(stg + gr * 512 + (((c * 8 + cg) ^ (gr & 7)) << 4));
          v.x = (v.x + bb.x) + y[8 * c + 4 * i]; v.y = (v.y + bb.y) + y[8 * c + 4 * i + 1];
          v.z = (v.z + bb.z) + y[8 * c + 4 * i + 2]; v.w = (v.w + bb.w) + y[8 * c + 4 * i + 3];
          if (i ? ok1 : ok0)
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(a.out) + ob + (long)(r0 + 4 * i) * a.out_ld + c * 32 + cg * 4) = v;
          const float sq = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
          if (i) ss1 += sq; else ss0 += sq;
          y[8 * c + 4 * i] = v.x; y[8 * c + 4 * i + 1] = v.y; y[8 * c + 4 * i + 2] = v.z; y[8 * c + 4 * i + 3] = v.w;
        }
      }
      // a row's 128 channels sit in the 8 lanes that share rr: three butterfly steps complete the sums of squares
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        ss0 += __shfl_xor_sync(0xffffffffu, ss0, o);
        ss1 += __shfl_xor_sync(0xffffffffu, ss1, o);
      }
      ROW_TRACE(2, 0x07);
      const float rinv0 = rsqrtf(ss0 * (1.0f / 128.0f) + 1e-5f), rinv1 = rsqrtf(ss1 * (1.0f / 128.0f) + 1e-5f);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 w = *reinterpret_cast<const float4*>(spar + 128 + c * 32 + cg * 4);
        const float4 nb4 = *reinterpret_cast<const float4*>(spar + 256 + c * 32 + cg * 4);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (!(i ? ok1 : ok0)) continue;
          const float ri = i ? rinv1 : rinv0;
          const float* yy = y + 8 * c + 4 * i;
          uint2 pk;
          pk.x = pack_bf16x2(fmaf(yy[0] * ri, w.x, nb4.x), fmaf(yy[1] * ri, w.y, nb4.y));
          pk.y = pack_bf16x2(fmaf(yy[2] * ri, w.z, nb4.z), fmaf(yy[3] * ri, w.w, nb4.w));
          *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(a.norm_out) + ob + (long)(r0 + 4 * i) * a.out_ld + c * 32 + cg * 4) = pk;
        }
      }
      ROW_TRACE(2, 0x08);
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// env AGB_ROWATTN=0 keeps the one-tile-per-CTA kernel for the row attention (A/B timing)
static bool rowattn_persistent() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("AGB_ROWATTN");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

// exponentials per 8 evaluated by polynomial in mode 0 (env AGB_ATTN_POLY = 0, 2, 4, 5 or 6; default 2, the fastest
// measured on B200: 0.369 ms per 8192-token layer against 0.393 (0), 0.375 (4), 0.407 (6))
static int attn_poly_of8() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("AGB_ATTN_POLY");
    v = e ? atoi(e) : 2;
    if (v != 0 && v != 2 && v != 4 && v != 5 && v != 6) v = 2;
  }
  return v;
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

  if (reinterpret_cast<uintptr_t>(a.Q) & 15) return set_error("ag_attention_fwd: Q must be 16-byte aligned");
  if (a.norm_out && (a.out_dtype != 1 || a.dv != 128 || !a.norm_w || !a.norm_b))
    return set_error("ag_attention_fwd: the normalised output needs fp32 out, dv == 128 and norm_w/norm_b");
  AttnKParams p;
  memset(&p, 0, sizeof(p));
  p.a = a;
  p.q_tiles = (a.n_q + kAttBM - 1) / kAttBM;
  p.nblk = (a.n_k + kAttBN - 1) / kAttBN;
  p.err = device_error_word(device);
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
  if (a.mode == 1 && a.norm_out != nullptr && a.dv == 128 && rowattn_persistent()) {
    const long tiles = (long)p.q_tiles * a.heads * a.batch;
    if (tiles > 1073741823L) return set_error("ag_attention_fwd: grid too large");
    static thread_local unsigned char attr_done[64];
    if (!attr_done[device]) {
      cudaError_t e = cudaFuncSetAttribute(row_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRowSmem);
      if (e != cudaSuccess) return set_error("ag_attention_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      attr_done[device] = 1;
    }
    const int sms = sm_count(device);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(tiles < sms ? tiles : sms), 1, 1);   // persistent: one CTA per SM; the q tiles of a row run side by side (L2)
    cfg.blockDim = dim3(kAttThreads, 1, 1);
    cfg.dynamicSmemBytes = kRowSmem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
#ifdef AGB_TRACE
    const char* trace_path = getenv("AGB_ROWATTN_TRACE");
    if (trace_path) {
      cudaMalloc(&p.trace, 3 * 2048 * sizeof(unsigned long long));
      cudaMemsetAsync(p.trace, 0, 3 * 2048 * sizeof(unsigned long long), stream);
    }
#endif
    cudaError_t le = cudaLaunchKernelEx(&cfg, row_attention_kernel, p);
    if (le != cudaSuccess) return set_error("ag_attention_fwd: launch failed: %s", cudaGetErrorString(le));
#ifdef AGB_TRACE
    if (trace_path) {
      static unsigned long long host[3 * 2048];
      cudaStreamSynchronize(stream);
      cudaMemcpy(host, p.trace, sizeof(host), cudaMemcpyDeviceToHost);
      cudaFree(p.trace);
      if (FILE* f = fopen(trace_path, "w")) {   // the last launch wins
        for (int r = 0; r < 3; ++r)
          for (int i = 0; i < 2048 && host[r * 2048 + i]; ++i)
            fprintf(f, "%d %llx %llu\n", r, host[r * 2048 + i] >> 48, host[r * 2048 + i] & 0xFFFFFFFFFFFFull);
        fclose(f);
      }
    }
#endif
    return check_launch("ag_attention_fwd");
  }
  // key split: two CTAs per tile when that shortens the schedule (waves x blocks per CTA; the DSMEM combine costs about
  // 6 blocks' worth: measured 0.371 -> 0.385 ms at 512 tiles x 64 blocks, 0.098 -> 0.063 ms at 64 tiles x 64 blocks).
  // With more tiles than SMs the full waves run unsplit and only a last partial wave of <= sms/2 tiles is split, so that
  // it costs about half a wave instead of a whole one.
  const long tiles = (long)p.q_tiles * a.heads * a.batch;
  if (tiles > 1073741823L) return set_error("ag_attention_fwd: grid too large");
  const int sms = sm_count(device);
  long n_plain = tiles, n_split = 0;
  if (a.mode == 0 && p.nblk >= 2 && a.norm_out == nullptr) {
    const long cost_split = (p.nblk + 1) / 2 + 6;
    if (tiles * 2 <= sms) {
      if (cost_split < p.nblk) { n_plain = 0; n_split = tiles; }
    } else if (tiles > sms) {
      const long tail = tiles % sms;
      if (tail > 0 && tail * 2 <= sms && cost_split < p.nblk) { n_plain = tiles - tail; n_split = tail; }
    }
    const char* e = getenv("AGB_ATTN_SPLIT");
    if (e && e[0] == '1') { n_plain = tiles; n_split = 0; }
    if (e && e[0] == '2') { n_plain = 0; n_split = tiles; }
  }
  for (int pass = 0; pass < 2; ++pass) {
    const int split = pass == 0 ? 1 : 2;
    const long ntile = pass == 0 ? n_plain : n_split;
    if (ntile == 0) continue;
    p.tile_off = pass == 0 ? 0 : (int)n_plain;
    void (*fn)(const AttnKParams) = nullptr;
    int slot = 0;
    switch (attn_poly_of8()) {
      case 0: fn = split == 2 ? attention_kernel<0, 2> : attention_kernel<0, 1>; slot = 0; break;
      case 4: fn = split == 2 ? attention_kernel<4, 2> : attention_kernel<4, 1>; slot = 2; break;
      case 5: fn = split == 2 ? attention_kernel<5, 2> : attention_kernel<5, 1>; slot = 3; break;
      case 6: fn = split == 2 ? attention_kernel<6, 2> : attention_kernel<6, 1>; slot = 4; break;
      default: fn = split == 2 ? attention_kernel<2, 2> : attention_kernel<2, 1>; slot = 1; break;
    }
    static thread_local unsigned char attr_done[64][2][5];
    if (!attr_done[device][split - 1][slot]) {
      cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmem);
      if (e != cudaSuccess) return set_error("ag_attention_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      attr_done[device][split - 1][slot] = 1;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(ntile * split), 1, 1);
    cfg.blockDim = dim3(kAttThreads, 1, 1);
    cfg.dynamicSmemBytes = kAttSmem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)split;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, fn, p);
    if (le != cudaSuccess) return set_error("ag_attention_fwd: launch failed: %s", cudaGetErrorString(le));
    if (pass == 0 && n_split > 0) {
      if (int rc = check_launch("ag_attention_fwd")) return rc;
    }
  }
  return check_launch("ag_attention_fwd");
}

}  // namespace agb
