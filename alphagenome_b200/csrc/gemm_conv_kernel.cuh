// K1/K3 — implicit-GEMM conv1d (taps 1/5) / linear / batched matmul on tcgen05 with a fused epilogue.
//
// Replaces (reference alphagenome_pytorch/alphagenome.py, "AG"): ConvBlock's Conv1d AG:431-452 as used by
// DNAEmbed.pointwise AG:1163, DownresBlock AG:470-481, UpresBlock AG:503-519, and the nn.Linear layers of
// Attention / FeedForward / SingleToPairwise / PairwiseRowAttention / OutputEmbedding.
//
// Design (one persistent CTA per SM, 576 threads; CG = 2: a cluster of two CTAs executes ONE cta_group::2 MMA of M = 256):
//   warp 0 (one lane)  TMA producer: per (tap, 64-wide K block) one A box [128 rows x 64] at row offset
//                      tap - taps/2 (out-of-range rows are zero-filled by TMA = the conv's zero padding)
//                      and one W box [n_tile / CG x 64], both 128B-swizzled, into a 5- or 6-stage smem ring
//                      (3 / 4 stages with single-CTA MMAs); an optional second operand pair (A2, W2) adds K blocks.
//   warp 1 (one lane)  MMA issuer: 4 x tcgen05.mma (M = 128 CG, N = n_tile, K = 16) per stage, fp32 accumulators in
//                      TMEM, double-buffered (2 x 256 columns) so the epilogue of tile i overlaps the
//                      mainloop of tile i+1.
//   warps 2-17         epilogue (4 warps per TMEM lane quarter, each takes every 4th EW-column chunk, EW = 32 or 16):
//                      tcgen05.ld 32x16 -> XOR-swizzled smem -> each lane owns 4 consecutive channels of EW / 4 rows ->
//                      fused affine / ReLU / residual / scale / activation / maxpool2 -> 128-bit global accesses that
//                      cover whole 128-byte lines (EW = 32, fp32).  The kernel is instantiated per set of enabled
//                      outputs (template FLAGS) so disabled paths cost no instructions; all pointer arithmetic is
//                      hoisted to one 64-bit base per tile.
//
// This header holds the kernel template; gemm_conv.cu (16-column epilogue chunks + the host side) and gemm_conv_wide.cu
// (32-column chunks) each instantiate one half of the <FLAGS, CG, EW> table, so the two halves compile in parallel.
#pragma once
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "kernels.h"

namespace agb {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;

constexpr int kMaxNTile = 256;
constexpr int kABytes = kBlockM * kBlockK * 2;      // 16 KB
constexpr int kBBytesMax = kMaxNTile * kBlockK * 2; // 32 KB
constexpr int kEpiWarps = 16;
// EW = columns of one staged epilogue chunk (16 or 32).  The epilogue-bound launches (width-1 convs, output embedding,
// K <= 128 linears) are bound by the L1TEX data pipe (ncu r2: 57-70 % of its wavefront peak, everything else < 45 %):
// one wavefront per distinct 128-byte line a shared- or global-memory instruction touches.  With 16-column chunks a
// row segment is 64 B of fp32 / 32 B of bf16, so every output line is written by 2 / 4 separate instructions; with
// 32-column chunks (a full 128-byte fp32 line per row) the global wavefronts of a tile halve.  The wider staging tile
// (4 KB per warp) costs one stage of the operand ring.
constexpr int kGemmThreads = 64 + kEpiWarps * 32;
// CG = CTAs per MMA (tcgen05 cta_group).  CG = 2: a cluster of two CTAs computes a 256 x n_tile tile; each CTA stages
// its own 128 rows of A and half of the W tile (<= 16 KB), which leaves room for a 6-deep ring.
template <int CG, int EW> struct GemmCfg {
  static constexpr int kBBytes = CG == 2 ? kBBytesMax / 2 : kBBytesMax;
  static constexpr int kStg = (CG == 2 ? 6 : 4) - (EW == 32 ? 1 : 0);
  static constexpr int kEpiBytesPerWarp = 32 * EW * 4;   // 2 / 4 KB, XOR-swizzled (no padding)
  static constexpr int kEpiBytes = kEpiWarps * kEpiBytesPerWarp;
};
template <int CG, int EW>
constexpr int gemm_smem_bytes() {
  return 1024 /*align slack*/ + GemmCfg<CG, EW>::kStg * (kABytes + GemmCfg<CG, EW>::kBBytes) + GemmCfg<CG, EW>::kEpiBytes + 256;
}

enum : int { F_RES = 1, F_OUT = 2, F_ACTA = 4, F_ACTB = 8, F_POOL = 16, F_COUNT = 32 };

struct GemmKParams {
  CUtensorMap tmA;
  CUtensorMap tmB;
  CUtensorMap tmB2;                                     // W box of the LAST N tile (n_last wide); == tmB when uniform
  CUtensorMap tmA2, tmW2, tmW2l;                        // optional second operand pair (AgGemmArgs.A2 / W2); tmW2l = last N tile
  AgGemmArgs g;
  int n_tile, m_tiles, n_tiles, total_tiles, kblocks;   // m_tiles counts (128 * CG)-row tiles
  int n_last;                                           // width of N tile n_tiles-1 (mixed-width tiling: k x n_tile + remainder)
  int kblocks2;                                         // 64-deep K blocks of the second operand pair (0 = none)
  int a_bmask;                                          // 0 when A is shared by every batch (a_bstride == 0), else -1
  unsigned int* err;
  unsigned long long* trace;   // -DAGB_TRACE builds only (AGB_GEMM_TRACE=<file>): CTA 0's per-role SM-clock timeline
};

#ifdef AGB_TRACE
#define GEMM_TRACE(role, tag)                                                                                     \
  do {                                                                                                            \
    if (p.trace && blockIdx.x == 0 && trc < 2040) p.trace[(role) * 2048 + trc++] = ((unsigned long long)(tag) << 48) | (unsigned long long)(clock64() & 0xFFFFFFFFFFFFull); \
  } while (0)
#else
#define GEMM_TRACE(role, tag) do {} while (0)
#endif

// GEMM epilogue activations (checked on the host): ACT_NONE / ACT_GELU1702 for bf16 and fp32 outputs; fp32 outputs
// also take the prediction heads' ACT_SOFTPLUS_MUL (softplus(v + shift) * scale, AG:1388-1393) and ACT_SIGMOID (AG:1321).
__device__ __forceinline__ float gelu1702_fast(float x) {   // 0.5x + 0.5x*tanh(0.851x): one MUFU op
  const float hx = 0.5f * x;
  return fmaf(hx, tanh_approx(0.851f * x), hx);
}
// x * sigmoid(1.702 x) for fp32 outputs: x * rcp(1 + 2^(-1.702 log2(e) x)), two MUFU ops and three FMA-pipe ops (the
// __expf / __fdividef intrinsics wrap the same two MUFU ops in range fix-ups — 5 more instructions per element — that
// cannot trigger here: 2^t overflowing to +inf gives rcp(inf) = 0, the correct limit)
__device__ __forceinline__ float gelu1702_acc(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -2.4554669595930157f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return x * r;
}
// F.softplus (beta 1, threshold 20): max(x, 0) + log1p(exp(-|x|)); the log1p argument is at most 1
__device__ __forceinline__ float softplus_acc(float x) {
  const float t = __expf(-fabsf(x));
  const float l = t < 1e-3f ? t * fmaf(-0.5f, t, 1.0f) : __logf(1.0f + t);
  return fmaxf(x, 0.0f) + l;
}
__device__ __forceinline__ float sigmoid_acc(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

__device__ __forceinline__ void store_bf16x4(char* dst, const float4& y) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(y.x, y.y);
  __nv_bfloat162 hi = __floats2bfloat162_rn(y.z, y.w);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&lo);
  pk.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(dst) = pk;
}

// One activated output for 4 consecutive channels of one row: dst = act(v*s + t), bf16 (fast GELU) or fp32.
__device__ __forceinline__ void epi_store4(char* base, uint32_t elem_off, bool is_bf16, int act, float4 v,
                                           const float4& s, const float4& t) {
  float4 y;
  if (is_bf16) {
    y.x = fmaf(v.x, s.x, t.x);
    y.y = fmaf(v.y, s.y, t.y);
    y.z = fmaf(v.z, s.z, t.z);
    y.w = fmaf(v.w, s.w, t.w);
    if (act != ACT_NONE) {
      y.x = gelu1702_fast(y.x);
      y.y = gelu1702_fast(y.y);
      y.z = gelu1702_fast(y.z);
      y.w = gelu1702_fast(y.w);
    }
    __nv_bfloat162 lo = __floats2bfloat162_rn(y.x, y.y);
    __nv_bfloat162 hi = __floats2bfloat162_rn(y.z, y.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(base + (size_t)elem_off * 2) = pk;
  } else {
    if (act == ACT_SOFTPLUS_MUL) {
      y.x = softplus_acc(v.x + t.x) * s.x;
      y.y = softplus_acc(v.y + t.y) * s.y;
      y.z = softplus_acc(v.z + t.z) * s.z;
      y.w = softplus_acc(v.w + t.w) * s.w;
    } else {
      y.x = fmaf(v.x, s.x, t.x);
      y.y = fmaf(v.y, s.y, t.y);
      y.z = fmaf(v.z, s.z, t.z);
      y.w = fmaf(v.w, s.w, t.w);
      if (act == ACT_GELU1702) {
        y.x = gelu1702_acc(y.x);
        y.y = gelu1702_acc(y.y);
        y.z = gelu1702_acc(y.z);
        y.w = gelu1702_acc(y.w);
      } else if (act == ACT_SIGMOID) {
        y.x = sigmoid_acc(y.x);
        y.y = sigmoid_acc(y.y);
        y.z = sigmoid_acc(y.z);
        y.w = sigmoid_acc(y.w);
      }
    }
    *reinterpret_cast<float4*>(base + (size_t)elem_off * 4) = y;
  }
}

// The same for the NG (row, 4-channel) groups a lane owns in one chunk, with the run-time format decided ONCE for all of
// them, so that the NG x 4 independent FMA -> MUFU -> store chains overlap.
template <int NG>
__device__ __forceinline__ void epi_store_groups(char* base, const uint32_t (&row)[NG], const bool (&ok)[NG], uint32_t ld, uint32_t col,
                                                 bool is_bf16, int act, const float4 (&v)[NG], const float4& s, const float4& t) {
  if (is_bf16) {
    if (act != ACT_NONE) {
#pragma unroll
      for (int k = 0; k < NG; ++k) {
        const float4 y = make_float4(gelu1702_fast(fmaf(v[k].x, s.x, t.x)), gelu1702_fast(fmaf(v[k].y, s.y, t.y)),
                                     gelu1702_fast(fmaf(v[k].z, s.z, t.z)), gelu1702_fast(fmaf(v[k].w, s.w, t.w)));
        if (ok[k]) store_bf16x4(base + (size_t)(row[k] * ld + col) * 2, y);
      }
    } else {
#pragma unroll
      for (int k = 0; k < NG; ++k) {
        const float4 y = make_float4(fmaf(v[k].x, s.x, t.x), fmaf(v[k].y, s.y, t.y), fmaf(v[k].z, s.z, t.z), fmaf(v[k].w, s.w, t.w));
        if (ok[k]) store_bf16x4(base + (size_t)(row[k] * ld + col) * 2, y);
      }
    }
  } else if (act == ACT_GELU1702) {
#pragma unroll
    for (int k = 0; k < NG; ++k) {
      const float4 y = make_float4(gelu1702_acc(fmaf(v[k].x, s.x, t.x)), gelu1702_acc(fmaf(v[k].y, s.y, t.y)),
                                   gelu1702_acc(fmaf(v[k].z, s.z, t.z)), gelu1702_acc(fmaf(v[k].w, s.w, t.w)));
      if (ok[k]) *reinterpret_cast<float4*>(base + (size_t)(row[k] * ld + col) * 4) = y;
    }
  } else {
#pragma unroll
    for (int k = 0; k < NG; ++k)
      if (ok[k]) epi_store4(base, row[k] * ld + col, false, act, v[k], s, t);
  }
}
__device__ __forceinline__ void epi_store4x4(char* base, const uint32_t (&row)[4], const bool (&ok)[4], uint32_t ld, uint32_t col, bool is_bf16,
                                             int act, const float4 (&v)[4], const float4& s, const float4& t) {
  epi_store_groups<4>(base, row, ok, ld, col, is_bf16, act, v, s, t);
}
__device__ __forceinline__ void epi_store2x4(char* base, const uint32_t (&row)[2], const bool (&ok)[2], uint32_t ld, uint32_t col, bool is_bf16,
                                             int act, const float4 (&v)[2], const float4& s, const float4& t) {
  epi_store_groups<2>(base, row, ok, ld, col, is_bf16, act, v, s, t);
}

// Per-channel epilogue parameters are re-read by every tile: keep them in what little L1 the 226 KB of shared memory
// leaves (evict_last).  ncu (profiles/ncu_r02_summary.md): 39 % of the HBM-bound w1 GEMM's stall samples sit on the first
// FFMA after the staging read (long scoreboard).  Letting the residual rows bypass L1 (no_allocate, ldg4_stream) was
// tried as the remedy and measured 1 % SLOWER (A/B on one box, AGB_EPI_STREAM=1): it stays off.
__device__ __forceinline__ float4 ldg4_or(const float* p, float dflt) {
  if (p) {
    float4 v;
    asm volatile("ld.global.nc.L1::evict_last.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
  }
  return make_float4(dflt, dflt, dflt, dflt);
}
// One float4 of a per-channel vector held by lane `src` of this warp (see the per-tile parameter registers below).
__device__ __forceinline__ float4 shfl4(const float4& v, int src) {
  float4 r;
  r.x = __shfl_sync(0xffffffffu, v.x, src);
  r.y = __shfl_sync(0xffffffffu, v.y, src);
  r.z = __shfl_sync(0xffffffffu, v.z, src);
  r.w = __shfl_sync(0xffffffffu, v.w, src);
  return r;
}

template <int FLAGS, int CG, int EW>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_conv_kernel(const __grid_constant__ GemmKParams p) {
  constexpr int kStages = GemmCfg<CG, EW>::kStg;
  constexpr int kBStage = GemmCfg<CG, EW>::kBBytes;
  constexpr int kEpiCols = EW;
  constexpr int kEpiBytesPerWarp = GemmCfg<CG, EW>::kEpiBytesPerWarp;
  constexpr int kEpiBytes = GemmCfg<CG, EW>::kEpiBytes;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);   // same offset in both CTAs of a pair

  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * kABytes;
  uint8_t* sEpi = smem + kStages * (kABytes + kBStage);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * (kABytes + kBStage) + kEpiBytes);
  uint64_t* full_bar = bars;                 // [kStages]   (CG = 2: only the leader CTA's are used)
  uint64_t* empty_bar = bars + kStages;      // [kStages]
  uint64_t* tfull_bar = bars + 2 * kStages;  // [2]
  uint64_t* tempty_bar = bars + 2 * kStages + 2;  // [2]   (CG = 2: the leader's, arrived on by both CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const AgGemmArgs& g = p.g;
  const uint32_t crank = CG == 2 ? cluster_ctarank() : 0u;   // 0 = leader (issues the MMAs)
  const int tile0 = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmB2);
    if (p.kblocks2) {
      tma_prefetch_desc(&p.tmA2);
      tma_prefetch_desc(&p.tmW2);
      tma_prefetch_desc(&p.tmW2l);
    }
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kEpiWarps * CG);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (CG == 2) {
      tmem_alloc_cg2(tmem_slot, 512);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(tmem_slot, 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();   // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();   // the next kernel may be scheduled as our CTAs retire ...
  pdl_wait();                // ... and we read nothing from global memory before the previous kernel has completed

  const int iters = g.taps * p.kblocks + p.kblocks2;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int trc = 0; (void)trc;
      for (int tile = tile0; tile < p.total_tiles; tile += tile_step) {
        GEMM_TRACE(0, 0x01);
        const int nt = tile % p.n_tiles;
        const int rest = tile / p.n_tiles;
        const int mt = rest % p.m_tiles;
        const int b = rest / p.m_tiles;
        const int row0 = g.a_row0 + (mt * CG + (int)crank) * kBlockM - g.taps / 2;
        const bool last_n = nt == p.n_tiles - 1;
        const int nw = last_n ? p.n_last : p.n_tile;                         // this tile's width
        const CUtensorMap* tmB = last_n ? &p.tmB2 : &p.tmB;
        const uint32_t b_bytes = (uint32_t)(nw / CG) * kBlockK * 2;          // this CTA's share of the W tile
        const int wrow0 = nt * p.n_tile + (int)crank * (nw / CG);
        for (int tap = 0; tap < g.taps; ++tap) {
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1u, 0x10u + stage, p.err);
            if constexpr (CG == 2) {
              // both CTAs' boxes complete on the leader's barrier, which expects the pair's bytes
              if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * (kABytes + b_bytes));
              tma_load_3d_cg2(sA + stage * kABytes, &p.tmA, &full_bar[stage], kb * kBlockK, row0 + tap, b & p.a_bmask);
              tma_load_3d_cg2(sB + stage * kBStage, tmB, &full_bar[stage], kb * kBlockK, wrow0, g.w_batched ? b : tap);
            } else {
              mbar_arrive_expect_tx(&full_bar[stage], kABytes + b_bytes);
              tma_load_3d(sA + stage * kABytes, &p.tmA, &full_bar[stage], kb * kBlockK, row0 + tap, b & p.a_bmask);
              tma_load_3d(sB + stage * kBStage, tmB, &full_bar[stage], kb * kBlockK, wrow0, g.w_batched ? b : tap);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
        // second operand pair: same boxes from its own tensor maps, accumulated into the same tile
        const CUtensorMap* tmW2 = last_n ? &p.tmW2l : &p.tmW2;
        const int row2 = g.a2_row0 + (mt * CG + (int)crank) * kBlockM;
        for (int kb = 0; kb < p.kblocks2; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u, 0x10u + stage, p.err);
          if constexpr (CG == 2) {
            if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * (kABytes + b_bytes));
            tma_load_3d_cg2(sA + stage * kABytes, &p.tmA2, &full_bar[stage], kb * kBlockK, row2, b);
            tma_load_3d_cg2(sB + stage * kBStage, tmW2, &full_bar[stage], kb * kBlockK, wrow0, 0);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], kABytes + b_bytes);
            tma_load_3d(sA + stage * kABytes, &p.tmA2, &full_bar[stage], kb * kBlockK, row2, b);
            tma_load_3d(sB + stage * kBStage, tmW2, &full_bar[stage], kb * kBlockK, wrow0, 0);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && crank == 0) {
      const uint32_t idesc_full = make_idesc_bf16(kBlockM * CG, (uint32_t)p.n_tile);
      const uint32_t idesc_last = make_idesc_bf16(kBlockM * CG, (uint32_t)p.n_last);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      int trc = 0; (void)trc;
      for (int tile = tile0; tile < p.total_tiles; tile += tile_step) {
        GEMM_TRACE(1, 0x01);
        mbar_wait(&tempty_bar[as], aphase ^ 1u, 0x30u + as, p.err);
        tc_fence_after();
        GEMM_TRACE(1, 0x02);
        const uint32_t idesc = (tile % p.n_tiles == p.n_tiles - 1) ? idesc_last : idesc_full;
        const uint32_t d_tmem = tmem_base + (uint32_t)as * kMaxNTile;
        for (int it = 0; it < iters; ++it) {
          mbar_wait(&full_bar[stage], phase, 0x20u + stage, p.err);
          tc_fence_after();
          const uint64_t da = make_sw128_kmajor_desc(smem_u32(sA + stage * kABytes));
          const uint64_t db = make_sw128_kmajor_desc(smem_u32(sB + stage * kBStage));
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // +32 bytes per K=16 step inside the 128B swizzle atom (encoded >>4 -> +2)
            if constexpr (CG == 2)
              tc_mma_bf16_ss_cg2(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (it > 0 || k > 0) ? 1u : 0u);
            else
              tc_mma_bf16_ss(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (it > 0 || k > 0) ? 1u : 0u);
          }
          if constexpr (CG == 2) tc_commit_cg2(&empty_bar[stage]); else tc_commit(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        GEMM_TRACE(1, 0x03);
        if constexpr (CG == 2) tc_commit_cg2(&tfull_bar[as]); else tc_commit(&tfull_bar[as]);
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..17)
    // Four warps per TMEM lane quarter (q = warp & 3); warp slot = (warp-2)>>2 takes chunks ch = slot, slot+4, ...
    // Per chunk: residual rows requested from global first, then LDTM (thread = row) -> swizzled smem -> each
    // lane re-reads 4 consecutive channels of 4 rows (two adjacent-row pairs, so maxpool2 is thread-local).
    constexpr bool kRes = (FLAGS & F_RES) != 0, kOut = (FLAGS & F_OUT) != 0, kActA = (FLAGS & F_ACTA) != 0,
                   kActB = (FLAGS & F_ACTB) != 0, kPool = (FLAGS & F_POOL) != 0;
    const int q = warp & 3;
    const int slot = (warp - 2) >> 2;
    uint8_t* stg = sEpi + (warp - 2) * kEpiBytesPerWarp;
    // Read-back layout: kLPR lanes span a row's EW columns (4 channels each), 32 / kLPR rows per instruction; a lane owns
    // kSteps rows as adjacent pairs (2m, 2m + 1), so maxpool2 is thread-local.
    constexpr int kLPR = EW / 4;            // 4 or 8 lanes per row
    constexpr int kSteps = kLPR;            // row steps per chunk: 32 rows / (32 / kLPR rows per step)
    const int cg = lane & (kLPR - 1);       // channel group: columns 4cg..4cg+3 of the chunk
    const int rsel = lane / kLPR;           // row selector
    // row (inside this warp's 32) of step k.  EW = 16: rows 2 rsel + {0,1} (+16), odd row-selectors take their odd row
    // first so that the two rows a quarter-warp reads in one LDS.128 hit disjoint bank groups; EW = 32: a quarter-warp
    // reads ONE row (8 lanes x 16 B), rows 2 rsel + {0,1} + 8 (k >> 1).
    auto row_of = [&](int k) -> int {
      if constexpr (EW == 16) return 2 * rsel + ((k & 1) ^ (rsel & 1)) + 16 * (k >> 1);
      else return 2 * rsel + (k & 1) + 8 * (k >> 1);
    };
    // staging tile.  EW = 16: 64 B rows, two rows per 128 B line; EW = 32: one row per line.  The 16 B unit index is
    // XORed with (line & 7) both ways, which makes the thread = row writes and the read-back above conflict-free.
    const uint32_t st_line = EW == 16 ? (uint32_t)(lane >> 1) : (uint32_t)lane, st_u0 = EW == 16 ? (uint32_t)(lane & 1) * 4 : 0u;
    uint8_t* st_wr = stg + st_line * 128;
    const uint32_t st_x = st_line & 7;
    auto st_rd = [&](int wr) -> const float4* {
      if constexpr (EW == 16) {
        const uint32_t line = (uint32_t)(wr >> 1);
        return reinterpret_cast<const float4*>(stg + line * 128 + ((((uint32_t)(wr & 1) * 4 + cg) ^ (line & 7)) << 4));
      } else {
        return reinterpret_cast<const float4*>(stg + (uint32_t)wr * 128 + (((uint32_t)cg ^ ((uint32_t)wr & 7)) << 4));
      }
    };
    int as = 0;
    uint32_t aphase = 0;
    int trc = (warp == 2 && lane == 0) ? 0 : 4096; (void)trc;
    const float post = g.post_scale ? __ldg(g.post_scale) : 1.0f;
    const bool relu = g.relu != 0;
    const bool a16 = g.actA.dtype == 0, b16 = g.actB.dtype == 0, p16 = g.actP.dtype == 0;
    const int aG = g.actA.act, bG = g.actB.act, pG = g.actP.act;
    const bool has_poolf = kPool && g.pool != nullptr, has_actP = kPool && g.actP.ptr != nullptr;
    for (int tile = tile0; tile < p.total_tiles; tile += tile_step) {
      const int nt = tile % p.n_tiles;
      const int nchunks = (nt == p.n_tiles - 1 ? p.n_last : p.n_tile) / kEpiCols;
      const int rest = tile / p.n_tiles;
      const int mt = rest % p.m_tiles;
      const int b = rest / p.m_tiles;
      const int n0 = nt * p.n_tile;
      const int row_t = (mt * CG + (int)crank) * kBlockM;   // first row of this CTA's 128-row tile
      const int lr_base = q * 32;                  // this warp's rows inside the tile
      // one 64-bit base per tensor per tile; everything below is 32-bit in-tile offsets
      const float* res_t = kRes ? g.res + (long)b * g.res_bstride + (long)(row_t >> g.res_shift) * g.res_ld + n0 : nullptr;
      char* out_t = kOut ? reinterpret_cast<char*>(g.out + (long)b * g.out_bstride + (long)row_t * g.out_ld + n0) : nullptr;
      char* actA_t = kActA ? reinterpret_cast<char*>(g.actA.ptr) + ((long)b * g.actA.bstride + (long)row_t * g.actA.ld + n0) * (a16 ? 2 : 4) : nullptr;
      char* actB_t = kActB ? reinterpret_cast<char*>(g.actB.ptr) + ((long)b * g.actB.bstride + (long)row_t * g.actB.ld + n0) * (b16 ? 2 : 4) : nullptr;
      char* pool_t = has_poolf ? reinterpret_cast<char*>(g.pool + (long)b * g.pool_bstride + (long)(row_t >> 1) * g.pool_ld + n0) : nullptr;
      char* actP_t = has_actP ? reinterpret_cast<char*>(g.actP.ptr) + ((long)b * g.actP.bstride + (long)(row_t >> 1) * g.actP.ld + n0) * (p16 ? 2 : 4) : nullptr;
      const float* ps_t = g.pre_scale ? g.pre_scale + n0 : nullptr;
      const float* pt_t = g.pre_shift ? g.pre_shift + n0 : nullptr;
      const float* sA_t = (kActA && g.actA.scale) ? g.actA.scale + n0 : nullptr;
      const float* tA_t = (kActA && g.actA.shift) ? g.actA.shift + (long)b * g.actA.shift_bstride + n0 : nullptr;
      const float* sB_t = (kActB && g.actB.scale) ? g.actB.scale + n0 : nullptr;
      const float* tB_t = (kActB && g.actB.shift) ? g.actB.shift + (long)b * g.actB.shift_bstride + n0 : nullptr;
      const float* sP_t = (has_actP && g.actP.scale) ? g.actP.scale + n0 : nullptr;
      const float* tP_t = (has_actP && g.actP.shift) ? g.actP.shift + (long)b * g.actP.shift_bstride + n0 : nullptr;
      // Per-channel parameters of the whole tile, ONE float4 register per (scale, shift) pair and lane: this warp's
      // chunks ch = slot + 4k cover 16 float4 positions idx = kLPR k + cg; lanes 0-15 hold the scale vector at
      // position idx = lane, lanes 16-31 the shift vector at idx = lane - 16.  Loaded here, BEFORE the wait for the
      // accumulator, and handed to the chunk loop by warp shuffles — the first version loaded them inside the loop
      // right before their first use, where 39 % of the HBM-bound w1 GEMM's stall samples sat (ncu, long scoreboard).
      float4 regPre, regA, regB, regP;
      {
        const int idx = lane & 15;
        const int pcol = ((slot + 4 * (idx / kLPR)) * kEpiCols) + 4 * (idx & (kLPR - 1));   // in-tile column of this lane's float4
        const bool inb = pcol < (nt == p.n_tiles - 1 ? p.n_last : p.n_tile);
        const bool hi = lane >= 16;
        regPre = ldg4_or(inb ? (hi ? (pt_t ? pt_t + pcol : nullptr) : (ps_t ? ps_t + pcol : nullptr)) : nullptr, hi ? 0.0f : 1.0f);
        if constexpr (kActA) regA = ldg4_or(inb ? (hi ? (tA_t ? tA_t + pcol : nullptr) : (sA_t ? sA_t + pcol : nullptr)) : nullptr, hi ? 0.0f : 1.0f);
        if constexpr (kActB) regB = ldg4_or(inb ? (hi ? (tB_t ? tB_t + pcol : nullptr) : (sB_t ? sB_t + pcol : nullptr)) : nullptr, hi ? 0.0f : 1.0f);
        if (has_actP) regP = ldg4_or(inb ? (hi ? (tP_t ? tP_t + pcol : nullptr) : (sP_t ? sP_t + pcol : nullptr)) : nullptr, hi ? 0.0f : 1.0f);
      }
      const int res_cols = kRes ? g.res_ch - n0 : 0;   // residual applies to in-tile columns < res_cols
      const int rows_left = g.M - row_t;               // in-tile rows >= rows_left do not exist

      GEMM_TRACE(2, 0x01);
      mbar_wait(&tfull_bar[as], aphase, 0x40u + as, p.err);
      tc_fence_after();
      GEMM_TRACE(2, 0x02);
      const uint32_t t_addr = tmem_base + ((uint32_t)lr_base << 16) + (uint32_t)as * kMaxNTile;

      // kG = false: every row of the tile exists (all tiles but the last M tile): no per-row predicates, so the
      // compiler keeps one straight-line body and overlaps the 16 elements' chains instead of branching per row group
      auto run_chunks = [&](auto guard) {
      constexpr bool kG = decltype(guard)::value;
      for (int ch = slot; ch < nchunks; ch += 4) {
        const int col = ch * kEpiCols + 4 * cg;  // in-tile column of this lane's 4 channels
        // ---- residual rows of the first four row steps: in flight while the accumulators travel TMEM -> registers -> smem ----
        float4 rr[4];
        auto load_res = [&](int k0) {
          if constexpr (kRes) {
            const bool on = col < res_cols;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int lr = lr_base + row_of(k0 + k);
              rr[k] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (on && (!kG || lr < rows_left))
                rr[k] = __ldg(reinterpret_cast<const float4*>(res_t + (uint32_t)((lr >> g.res_shift) * g.res_ld + col)));
            }
          }
        };
        load_res(0);
        GEMM_TRACE(2, 0x10);
#pragma unroll
        for (int part = 0; part < EW / 16; ++part) {
          uint32_t r[16];
          tmem_ld_32x16(t_addr + (uint32_t)(ch * kEpiCols + 16 * part), r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(st_wr + (((st_u0 + 4 * part + j) ^ st_x) << 4)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        }
        GEMM_TRACE(2, 0x11);
        // per-channel parameters of this chunk: from the per-tile registers of lanes kLPR k + cg (scale) / 16 + kLPR k + cg (shift)
        const int psrc = ((ch - slot) >> 2) * kLPR + cg;
        const float4 ps = shfl4(regPre, psrc), pt = shfl4(regPre, psrc + 16);
        float4 sA, tA, sB, tB, sP, tP;
        if constexpr (kActA) { sA = shfl4(regA, psrc); tA = shfl4(regA, psrc + 16); }
        if constexpr (kActB) { sB = shfl4(regB, psrc); tB = shfl4(regB, psrc + 16); }
        if (has_actP) { sP = shfl4(regP, psrc); tP = shfl4(regP, psrc + 16); }
        __syncwarp();
        GEMM_TRACE(2, 0x12);

        // Four row steps (two adjacent-row pairs) of this lane at a time, processed TOGETHER stage by stage: the output
        // formats are run-time values, and one call per row group would put every group's FMA -> MUFU -> store chain
        // behind its own reconvergence point.
#pragma unroll
        for (int k0 = 0; k0 < kSteps; k0 += 4) {
          float4 x[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float4 v = *st_rd(row_of(k0 + k));
            v.x = fmaf(v.x, ps.x, pt.x);
            v.y = fmaf(v.y, ps.y, pt.y);
            v.z = fmaf(v.z, ps.z, pt.z);
            v.w = fmaf(v.w, ps.w, pt.w);
            if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            if constexpr (kRes) {
              const float4 ra = rr[k];
              v.x += ra.x; v.y += ra.y; v.z += ra.z; v.w += ra.w;
            }
            v.x *= post; v.y *= post; v.z *= post; v.w *= post;
            x[k] = v;
          }
          if (k0 + 4 < kSteps) load_res(k0 + 4);   // the next four rows' residual: requested before this group's stores
          bool rowok[4];
          uint32_t lrk[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            lrk[k] = (uint32_t)(lr_base + row_of(k0 + k));   // row inside the tile
            rowok[k] = !kG || (int)lrk[k] < rows_left;
          }
          if constexpr (kOut) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (rowok[k]) *reinterpret_cast<float4*>(out_t + (size_t)(lrk[k] * (uint32_t)g.out_ld + (uint32_t)col) * 4) = x[k];
          }
          if constexpr (kActA) epi_store4x4(actA_t, lrk, rowok, (uint32_t)g.actA.ld, (uint32_t)col, a16, aG, x, sA, tA);
          if constexpr (kActB) epi_store4x4(actB_t, lrk, rowok, (uint32_t)g.actB.ld, (uint32_t)col, b16, bG, x, sB, tB);
          if constexpr (kPool) {
            float4 pv[2];
            bool pok[2];
            uint32_t prk[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              pv[i].x = fmaxf(x[2 * i].x, x[2 * i + 1].x);
              pv[i].y = fmaxf(x[2 * i].y, x[2 * i + 1].y);
              pv[i].z = fmaxf(x[2 * i].z, x[2 * i + 1].z);
              pv[i].w = fmaxf(x[2 * i].w, x[2 * i + 1].w);
              const int lr = (lr_base + row_of(k0 + 2 * i)) & ~1;   // even row of the pair
              pok[i] = !kG || lr + 1 < rows_left;
              prk[i] = (uint32_t)(lr >> 1);
            }
            if (has_poolf) {
#pragma unroll
              for (int i = 0; i < 2; ++i)
                if (pok[i]) *reinterpret_cast<float4*>(pool_t + (size_t)(prk[i] * (uint32_t)g.pool_ld + (uint32_t)col) * 4) = pv[i];
            }
            if (has_actP) epi_store2x4(actP_t, prk, pok, (uint32_t)g.actP.ld, (uint32_t)col, p16, pG, pv, sP, tP);
          }
        }
        __syncwarp();
        GEMM_TRACE(2, 0x13);
      }
      };
      if (rows_left >= kBlockM) run_chunks(std::false_type{});
      else run_chunks(std::true_type{});
      tc_fence_before();
      __syncwarp();
      GEMM_TRACE(2, 0x03);
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_cluster(&tempty_bar[as], 0u);   // the leader's MMA thread owns the accumulators
        else mbar_arrive(&tempty_bar[as]);
      }
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();   // neither CTA leaves while the pair's MMAs / arrives may still touch it
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_cg2(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

typedef void (*GemmKernelFn)(const GemmKParams);

// one instantiation per set of enabled outputs (res / out / actA / actB / pool) and per CTA-group size, for one chunk width
template <int EW, int... I>
struct GemmTable {
  GemmKernelFn fn[2][sizeof...(I)] = {{gemm_conv_kernel<I, 1, EW>...}, {gemm_conv_kernel<I, 2, EW>...}};   // [cg - 1][flags]
};
template <int EW, int N, int... I>
struct MakeGemmTable : MakeGemmTable<EW, N - 1, N - 1, I...> {};
template <int EW, int... I>
struct MakeGemmTable<EW, 0, I...> {
  typedef GemmTable<EW, I...> type;
};
// the kernel for (cg, flags) with 32-column epilogue chunks (gemm_conv_wide.cu)
GemmKernelFn gemm_wide_kernel(int cg, int flags);

}  // namespace agb
