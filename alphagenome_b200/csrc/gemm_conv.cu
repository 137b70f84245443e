// K1/K3 — implicit-GEMM conv1d (taps 1/5) / linear / batched matmul on tcgen05 with a fused epilogue.
//
// Replaces (reference alphagenome_pytorch/alphagenome.py, "AG"): ConvBlock's Conv1d AG:431-452 as used by
// DNAEmbed.pointwise AG:1163, DownresBlock AG:470-481, UpresBlock AG:503-519, and the nn.Linear layers of
// Attention / FeedForward / SingleToPairwise / PairwiseRowAttention / OutputEmbedding.
//
// Design (one persistent CTA per SM, 320 threads):
//   warp 0 (one lane)  TMA producer: per (tap, 64-wide K block) one A box [128 rows x 64] at row offset
//                      tap - taps/2 (out-of-range rows are zero-filled by TMA = the conv's zero padding)
//                      and one W box [n_tile x 64], both 128B-swizzled, into a 4-stage smem ring.
//   warp 1 (one lane)  MMA issuer: 4 x tcgen05.mma (M=128, N=n_tile, K=16) per stage, fp32 accumulators in
//                      TMEM, double-buffered (2 x 256 columns) so the epilogue of tile i overlaps the
//                      mainloop of tile i+1.
//   warps 2-9          epilogue: tcgen05.ld 32x16 blocks -> padded smem -> each lane owns 4 consecutive channels
//                      of 4 rows -> fused affine / ReLU / residual / scale / activation / maxpool2 ->
//                      128-bit global loads and stores.
#include "common.cuh"
#include "kernels.h"

namespace agb {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kStages = 4;
constexpr int kMaxNTile = 256;
constexpr int kABytes = kBlockM * kBlockK * 2;      // 16 KB
constexpr int kBBytesMax = kMaxNTile * kBlockK * 2; // 32 KB
constexpr int kEpiWarps = 8;
constexpr int kEpiCols = 16;                          // columns per TMEM load / staged chunk
constexpr int kEpiStride = 20;                        // floats per staged row (16 + 4 pad; 5 x 16 B, odd)
constexpr int kEpiBytes = kEpiWarps * 32 * kEpiStride * 4;
constexpr int kGemmThreads = 64 + kEpiWarps * 32;
constexpr int kGemmSmem = 1024 /*align slack*/ + kStages * (kABytes + kBBytesMax) + kEpiBytes + 256;

struct GemmKParams {
  CUtensorMap tmA;
  CUtensorMap tmB;
  AgGemmArgs g;
  int n_tile, m_tiles, n_tiles, total_tiles, kblocks;
  unsigned int* err;
};

// One activated output for 4 consecutive channels of one row (scale/shift pre-loaded per chunk).
// Only ACT_NONE / ACT_GELU1702 occur in GEMM epilogues (checked on the host): one uniform branch per 4 elements.
__device__ __forceinline__ float gelu1702_fast(float x) {   // 0.5x + 0.5x*tanh(0.851x): one MUFU op
  const float hx = 0.5f * x;
  return fmaf(hx, tanh_approx(0.851f * x), hx);
}
__device__ __forceinline__ float gelu1702_acc(float x) { return __fdividef(x, 1.0f + __expf(-1.702f * x)); }

__device__ __forceinline__ void epi_store4(const AgEpiOut& o, long off, float4 v, const float4& s, const float4& t) {
  float4 y;
  y.x = fmaf(v.x, s.x, t.x);
  y.y = fmaf(v.y, s.y, t.y);
  y.z = fmaf(v.z, s.z, t.z);
  y.w = fmaf(v.w, s.w, t.w);
  if (o.dtype == 0) {
    if (o.act != ACT_NONE) {
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
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(o.ptr) + off) = pk;
  } else {
    if (o.act != ACT_NONE) {
      y.x = gelu1702_acc(y.x);
      y.y = gelu1702_acc(y.y);
      y.z = gelu1702_acc(y.z);
      y.w = gelu1702_acc(y.w);
    }
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(o.ptr) + off) = y;
  }
}

__device__ __forceinline__ float4 ldg4_or(const float* p, long off, float dflt) {
  if (p) return __ldg(reinterpret_cast<const float4*>(p + off));
  return make_float4(dflt, dflt, dflt, dflt);
}

__global__ void __launch_bounds__(kGemmThreads, 1) gemm_conv_kernel(const __grid_constant__ GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);

  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * kABytes;
  float* sEpi = reinterpret_cast<float*>(smem + kStages * (kABytes + kBBytesMax));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * (kABytes + kBBytesMax) + kEpiBytes);
  uint64_t* full_bar = bars;                 // [kStages]
  uint64_t* empty_bar = bars + kStages;      // [kStages]
  uint64_t* tfull_bar = bars + 2 * kStages;  // [2]
  uint64_t* tempty_bar = bars + 2 * kStages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const AgGemmArgs& g = p.g;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kEpiWarps);
    }
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

  const int iters = g.taps * p.kblocks;
  const uint32_t b_bytes = (uint32_t)p.n_tile * kBlockK * 2;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles;
        const int rest = tile / p.n_tiles;
        const int mt = rest % p.m_tiles;
        const int b = rest / p.m_tiles;
        const int row0 = g.a_row0 + mt * kBlockM - g.taps / 2;
        for (int tap = 0; tap < g.taps; ++tap) {
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1u, 0x10u + stage, p.err);
            mbar_arrive_expect_tx(&full_bar[stage], kABytes + b_bytes);
            tma_load_3d(sA + stage * kABytes, &p.tmA, &full_bar[stage], kb * kBlockK, row0 + tap, b);
            tma_load_3d(sB + stage * kBBytesMax, &p.tmB, &full_bar[stage], kb * kBlockK, nt * p.n_tile,
                        g.w_batched ? b : tap);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(kBlockM, (uint32_t)p.n_tile);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[as], aphase ^ 1u, 0x30u + as, p.err);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)as * kMaxNTile;
        for (int it = 0; it < iters; ++it) {
          mbar_wait(&full_bar[stage], phase, 0x20u + stage, p.err);
          tc_fence_after();
          const uint64_t da = make_sw128_kmajor_desc(smem_u32(sA + stage * kABytes));
          const uint64_t db = make_sw128_kmajor_desc(smem_u32(sB + stage * kBBytesMax));
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // +32 bytes per K=16 step inside the 128B swizzle atom (encoded >>4 -> +2)
            tc_mma_bf16_ss(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (it > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        tc_commit(&tfull_bar[as]);
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    // Two warps per TMEM lane quarter; each takes every other 16-column chunk.  Per chunk: LDTM (thread =
    // row) -> padded smem -> each lane re-reads 4 consecutive channels of 4 rows (two adjacent-row pairs, so
    // maxpool2 is thread-local) -> 128-bit global loads/stores.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    float* stg = sEpi + (warp - 2) * 32 * kEpiStride;
    const int cg = lane & 3;     // channel group: columns 4cg..4cg+3 of the chunk
    const int rsel = lane >> 2;  // rows 2*rsel + {0,1} (+16)
    int as = 0;
    uint32_t aphase = 0;
    const float post = g.post_scale ? __ldg(g.post_scale) : 1.0f;
    const bool do_pool = (g.pool != nullptr) || (g.actP.ptr != nullptr);
    const int nchunks = p.n_tile / kEpiCols;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles;
      const int rest = tile / p.n_tiles;
      const int mt = rest % p.m_tiles;
      const int b = rest / p.m_tiles;
      const int n0 = nt * p.n_tile;
      const int r_base = mt * kBlockM + q * 32;

      mbar_wait(&tfull_bar[as], aphase, 0x40u + as, p.err);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)as * kMaxNTile;

      for (int ch = half; ch < nchunks; ch += 2) {
        const int c = n0 + ch * kEpiCols + 4 * cg;
        // ---- global loads first: residual rows and per-channel parameters are in flight while the
        //      accumulators travel TMEM -> registers -> smem ----
        const bool has_res = (g.res != nullptr) && (c < g.res_ch);
        float4 rr[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int row = r_base + 2 * rsel + (k & 1) + 16 * (k >> 1);
          rr[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (has_res && row < g.M)
            rr[k] = __ldg(reinterpret_cast<const float4*>(g.res + (long)b * g.res_bstride + (long)(row >> g.res_shift) * g.res_ld + c));
        }
        const float4 ps = ldg4_or(g.pre_scale, c, 1.0f);
        const float4 pt = ldg4_or(g.pre_shift, c, 0.0f);
        float4 sA, tA, sB, tB, sP, tP;
        if (g.actA.ptr) { sA = ldg4_or(g.actA.scale, c, 1.0f); tA = ldg4_or(g.actA.shift, (long)b * g.actA.shift_bstride + c, 0.0f); }
        if (g.actB.ptr) { sB = ldg4_or(g.actB.scale, c, 1.0f); tB = ldg4_or(g.actB.shift, (long)b * g.actB.shift_bstride + c, 0.0f); }
        if (g.actP.ptr) { sP = ldg4_or(g.actP.scale, c, 1.0f); tP = ldg4_or(g.actP.shift, (long)b * g.actP.shift_bstride + c, 0.0f); }

        uint32_t r[16];
        tmem_ld_32x16(t_addr + (uint32_t)(ch * kEpiCols), r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(stg + lane * kEpiStride + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        __syncwarp();

        const long out_base = (long)b * g.out_bstride + c;
        const long a_base = (long)b * g.actA.bstride + c;
        const long b_base = (long)b * g.actB.bstride + c;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          float4 v[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int lr = 2 * rsel + h + 16 * i;
            const int row = r_base + lr;
            float4 x = *reinterpret_cast<const float4*>(stg + lr * kEpiStride + 4 * cg);
            const float4 ra = rr[i * 2 + h];
            x.x = fmaf(x.x, ps.x, pt.x);
            x.y = fmaf(x.y, ps.y, pt.y);
            x.z = fmaf(x.z, ps.z, pt.z);
            x.w = fmaf(x.w, ps.w, pt.w);
            if (g.relu) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
            x.x = (x.x + ra.x) * post;
            x.y = (x.y + ra.y) * post;
            x.z = (x.z + ra.z) * post;
            x.w = (x.w + ra.w) * post;
            if (row < g.M) {
              if (g.out) *reinterpret_cast<float4*>(g.out + out_base + (long)row * g.out_ld) = x;
              if (g.actA.ptr) epi_store4(g.actA, a_base + (long)row * g.actA.ld, x, sA, tA);
              if (g.actB.ptr) epi_store4(g.actB, b_base + (long)row * g.actB.ld, x, sB, tB);
            }
            v[h] = x;
          }
          if (do_pool) {
            const int row = r_base + 2 * rsel + 16 * i;
            if (row + 1 < g.M) {
              float4 pv;
              pv.x = fmaxf(v[0].x, v[1].x);
              pv.y = fmaxf(v[0].y, v[1].y);
              pv.z = fmaxf(v[0].z, v[1].z);
              pv.w = fmaxf(v[0].w, v[1].w);
              const long prow = row >> 1;
              if (g.pool) *reinterpret_cast<float4*>(g.pool + (long)b * g.pool_bstride + prow * g.pool_ld + c) = pv;
              if (g.actP.ptr) epi_store4(g.actP, (long)b * g.actP.bstride + prow * g.actP.ld + c, pv, sP, tP);
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ----------------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------------
static int pick_n_tile(int N) {
  for (int t = 256; t >= 16; t -= 16)
    if (N % t == 0) return t;
  return 0;
}

int launch_gemm(const AgGemmArgs& a, int device, cudaStream_t stream) {
  if (!a.A || !a.W) return set_error("ag_gemm_fwd: A and W must be non-null");
  if (a.batch < 1 || a.M < 1 || a.N < 16 || a.K < 8) return set_error("ag_gemm_fwd: bad sizes batch=%d M=%d N=%d K=%d", a.batch, a.M, a.N, a.K);
  if (a.N % 16) return set_error("ag_gemm_fwd: N=%d must be a multiple of 16", a.N);
  if (a.K % 8 || a.a_ld % 8 || a.w_ld % 8 || a.a_bstride % 8 || a.w_zstride % 8)
    return set_error("ag_gemm_fwd: K, a_ld, w_ld and strides must be multiples of 8 (16-byte TMA alignment)");
  if ((a.taps & 1) == 0 || a.taps < 1 || a.taps > 15) return set_error("ag_gemm_fwd: taps=%d must be odd and <= 15", a.taps);
  if (a.w_batched && a.taps != 1) return set_error("ag_gemm_fwd: batched W requires taps == 1");
  if (a.w_rows < 0 || a.w_rows > a.N) return set_error("ag_gemm_fwd: w_rows=%d out of range", a.w_rows);
  if ((a.pool || a.actP.ptr) && (a.M & 1)) return set_error("ag_gemm_fwd: maxpool2 epilogue needs even M (got %d)", a.M);
  if ((reinterpret_cast<uintptr_t>(a.A) | reinterpret_cast<uintptr_t>(a.W)) & 15) return set_error("ag_gemm_fwd: A/W must be 16-byte aligned");
  // the epilogue moves 4 channels per lane with 128-bit (fp32) / 64-bit (bf16) accesses
  if (a.res && ((a.res_ch & 3) || (a.res_ld & 3) || (a.res_bstride & 3) || (reinterpret_cast<uintptr_t>(a.res) & 15)))
    return set_error("ag_gemm_fwd: res needs res_ch, res_ld, res_bstride multiples of 4 and a 16-byte aligned pointer");
  if (a.out && ((a.out_ld & 3) || (a.out_bstride & 3) || (reinterpret_cast<uintptr_t>(a.out) & 15)))
    return set_error("ag_gemm_fwd: out needs out_ld/out_bstride multiples of 4 and a 16-byte aligned pointer");
  if (a.pool && ((a.pool_ld & 3) || (a.pool_bstride & 3) || (reinterpret_cast<uintptr_t>(a.pool) & 15)))
    return set_error("ag_gemm_fwd: pool needs pool_ld/pool_bstride multiples of 4 and a 16-byte aligned pointer");
  for (const AgEpiOut* o : {&a.actA, &a.actB, &a.actP}) {
    if (!o->ptr) continue;
    const uintptr_t al = o->dtype == 0 ? 7 : 15;
    if ((o->ld & 3) || (o->bstride & 3) || (reinterpret_cast<uintptr_t>(o->ptr) & al) || (o->shift_bstride & 3) ||
        (reinterpret_cast<uintptr_t>(o->scale) & 15) || (reinterpret_cast<uintptr_t>(o->shift) & 15))
      return set_error("ag_gemm_fwd: activated outputs need ld/bstride multiples of 4 and aligned pointers");
    if (o->dtype != 0 && o->dtype != 1) return set_error("ag_gemm_fwd: bad output dtype %d", o->dtype);
    if (o->act != AG_ACT_NONE && o->act != AG_ACT_GELU1702) return set_error("ag_gemm_fwd: epilogue activations are AG_ACT_NONE or AG_ACT_GELU1702 (got %d)", o->act);
  }
  if ((reinterpret_cast<uintptr_t>(a.pre_scale) | reinterpret_cast<uintptr_t>(a.pre_shift)) & 15)
    return set_error("ag_gemm_fwd: pre_scale/pre_shift must be 16-byte aligned");

  GemmKParams p;
  memset(&p, 0, sizeof(p));
  p.g = a;
  p.n_tile = pick_n_tile(a.N);
  if (!p.n_tile) return set_error("ag_gemm_fwd: no N tile for N=%d", a.N);
  p.m_tiles = (a.M + kBlockM - 1) / kBlockM;
  p.n_tiles = a.N / p.n_tile;
  p.total_tiles = p.m_tiles * p.n_tiles * a.batch;
  p.kblocks = (a.K + kBlockK - 1) / kBlockK;
  p.err = device_error_word(device);

  {
    uint64_t dims[3] = {(uint64_t)a.K, (uint64_t)a.a_rows, (uint64_t)a.batch};
    uint64_t strides[2] = {(uint64_t)a.a_ld * 2, (uint64_t)(a.batch > 1 ? a.a_bstride : (int64_t)a.a_ld * a.a_rows) * 2};
    uint32_t box[3] = {kBlockK, kBlockM, 1};
    if (int rc = make_tensor_map_bf16(&p.tmA, a.A, 3, dims, strides, box)) return rc;
  }
  {
    const int nz = a.w_batched ? a.batch : a.taps;
    uint64_t dims[3] = {(uint64_t)a.K, (uint64_t)(a.w_rows > 0 ? a.w_rows : a.N), (uint64_t)nz};
    uint64_t strides[2] = {(uint64_t)a.w_ld * 2, (uint64_t)(nz > 1 ? a.w_zstride : (int64_t)a.w_ld * (a.w_rows > 0 ? a.w_rows : a.N)) * 2};
    uint32_t box[3] = {kBlockK, (uint32_t)p.n_tile, 1};
    if (int rc = make_tensor_map_bf16(&p.tmB, a.W, 3, dims, strides, box)) return rc;
  }

  static thread_local int attr_set_for = -1;
  if (attr_set_for != device) {
    cudaError_t e = cudaFuncSetAttribute(gemm_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem);
    if (e != cudaSuccess) return set_error("ag_gemm_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set_for = device;
  }
  int grid = p.total_tiles < sm_count(device) ? p.total_tiles : sm_count(device);
  gemm_conv_kernel<<<grid, kGemmThreads, kGemmSmem, stream>>>(p);
  return check_launch("ag_gemm_fwd");
}

}  // namespace agb
