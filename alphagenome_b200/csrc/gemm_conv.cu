// K1/K3 — implicit-GEMM conv1d (taps 1/5) / linear / batched matmul on tcgen05 with a fused epilogue.
//
// Replaces (reference alphagenome_pytorch/alphagenome.py, "AG"): ConvBlock's Conv1d AG:431-452 as used by
// DNAEmbed.pointwise AG:1163, DownresBlock AG:470-481, UpresBlock AG:503-519, and the nn.Linear layers of
// Attention / FeedForward / SingleToPairwise / PairwiseRowAttention / OutputEmbedding.
//
// Design (one persistent CTA per SM, 192 threads):
//   warp 0 (one lane)  TMA producer: per (tap, 64-wide K block) one A box [128 rows x 64] at row offset
//                      tap - taps/2 (out-of-range rows are zero-filled by TMA = the conv's zero padding)
//                      and one W box [n_tile x 64], both 128B-swizzled, into a 4-stage smem ring.
//   warp 1 (one lane)  MMA issuer: 4 x tcgen05.mma (M=128, N=n_tile, K=16) per stage, fp32 accumulators in
//                      TMEM, double-buffered (2 x 256 columns) so the epilogue of tile i overlaps the
//                      mainloop of tile i+1.
//   warps 2-5          epilogue: tcgen05.ld 32x32 blocks -> smem transpose (lane = channel) -> fused
//                      affine / ReLU / residual / scale / activation / maxpool2 -> coalesced global stores.
#include "common.cuh"
#include "kernels.h"

namespace agb {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kStages = 4;
constexpr int kMaxNTile = 256;
constexpr int kABytes = kBlockM * kBlockK * 2;      // 16 KB
constexpr int kBBytesMax = kMaxNTile * kBlockK * 2; // 32 KB
constexpr int kEpiWarps = 4;
constexpr int kEpiPad = 33;
constexpr int kEpiBytes = kEpiWarps * 32 * kEpiPad * 4;
constexpr int kGemmThreads = 192;
constexpr int kGemmSmem = 1024 /*align slack*/ + kStages * (kABytes + kBBytesMax) + kEpiBytes + 256;

struct GemmKParams {
  CUtensorMap tmA;
  CUtensorMap tmB;
  AgGemmArgs g;
  int n_tile, m_tiles, n_tiles, total_tiles, kblocks;
  unsigned int* err;
};

__device__ __forceinline__ void epi_store(const AgEpiOut& o, int b, long row, int c, float v) {
  float s = o.scale ? __ldg(o.scale + c) : 1.0f;
  float t = o.shift ? __ldg(o.shift + (long)b * o.shift_bstride + c) : 0.0f;
  float y = act_apply(fmaf(v, s, t), o.act);
  long idx = (long)b * o.bstride + row * o.ld + c;
  if (o.dtype == 0) reinterpret_cast<__nv_bfloat16*>(o.ptr)[idx] = __float2bfloat16_rn(y);
  else reinterpret_cast<float*>(o.ptr)[idx] = y;
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
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    float* stg = sEpi + (warp - 2) * 32 * kEpiPad;
    int as = 0;
    uint32_t aphase = 0;
    const float post = g.post_scale ? __ldg(g.post_scale) : 1.0f;
    const bool do_pool = (g.pool != nullptr) || (g.actP.ptr != nullptr);
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

      for (int c0 = 0; c0 < p.n_tile; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(t_addr + (uint32_t)c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) stg[lane * kEpiPad + j] = __uint_as_float(r[j]);
        __syncwarp();

        const int c = n0 + c0 + lane;
        const bool c_ok = (c0 + lane) < p.n_tile;
        if (c_ok) {
          const float ps = g.pre_scale ? __ldg(g.pre_scale + c) : 1.0f;
          const float pt = g.pre_shift ? __ldg(g.pre_shift + c) : 0.0f;
          const bool has_res = (g.res != nullptr) && (c < g.res_ch);
          const float* resb = g.res + (long)b * g.res_bstride + c;
#pragma unroll 4
          for (int rp = 0; rp < 16; ++rp) {
            float v[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int row = r_base + 2 * rp + h;
              float x = fmaf(stg[(2 * rp + h) * kEpiPad + lane], ps, pt);
              if (g.relu) x = fmaxf(x, 0.0f);
              if (row < g.M) {
                if (has_res) x += __ldg(resb + (long)(row >> g.res_shift) * g.res_ld);
                x *= post;
                if (g.out) g.out[(long)b * g.out_bstride + (long)row * g.out_ld + c] = x;
                if (g.actA.ptr) epi_store(g.actA, b, row, c, x);
                if (g.actB.ptr) epi_store(g.actB, b, row, c, x);
              }
              v[h] = x;
            }
            if (do_pool) {
              const int row = r_base + 2 * rp;
              if (row + 1 < g.M) {
                const float pv = fmaxf(v[0], v[1]);
                const long prow = row >> 1;
                if (g.pool) g.pool[(long)b * g.pool_bstride + prow * g.pool_ld + c] = pv;
                if (g.actP.ptr) epi_store(g.actP, b, prow, c, pv);
              }
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
  if ((a.pool || a.actP.ptr) && (a.M & 1)) return set_error("ag_gemm_fwd: maxpool2 epilogue needs even M (got %d)", a.M);
  if ((reinterpret_cast<uintptr_t>(a.A) | reinterpret_cast<uintptr_t>(a.W)) & 15) return set_error("ag_gemm_fwd: A/W must be 16-byte aligned");

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
    uint64_t dims[3] = {(uint64_t)a.K, (uint64_t)a.N, (uint64_t)nz};
    uint64_t strides[2] = {(uint64_t)a.w_ld * 2, (uint64_t)(nz > 1 ? a.w_zstride : (int64_t)a.w_ld * a.N) * 2};
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
