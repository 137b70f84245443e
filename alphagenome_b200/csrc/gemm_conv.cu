// Host side of ag_gemm_fwd and the 16-column-chunk half of the kernel table (see gemm_conv_kernel.cuh).
#include "gemm_conv_kernel.cuh"

namespace agb {

// ----------------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------------
static const MakeGemmTable<16, F_COUNT>::type kGemmTable;

// env AGB_GEMM_CG=1|2 seeds ag_gemm_set_cta_group() (A/B timing, bring-up); default 0 = pairs whenever there are
// at least two 128-row tiles
static int forced_cg() {
  static bool seeded = false;
  if (!seeded) {
    seeded = true;
    const char* e = getenv("AGB_GEMM_CG");
    if (e && (e[0] == '1' || e[0] == '2') && g_gemm_cta_group.load() == 0) g_gemm_cta_group.store(e[0] - '0');
  }
  return g_gemm_cta_group.load();
}

// N tiling.  The tensor pipe spends the same time on a 176/192/224-wide tile as on a 256-wide one (measured: ~510 clk
// per 64-deep K block for every width in (128, 256]), so N is cut into 256-wide tiles plus ONE narrower remainder tile
// (a multiple of 16 per CTA) instead of the largest divisor of N: 1152 = 4 x 256 + 128 (was 6 x 192), 1408 = 5 x 256 +
// 128 (was 8 x 176), 896 = 3 x 256 + 128 (was 4 x 224).  env AGB_GEMM_UNIFORM=1 restores the divisor tiling (A/B timing).
static bool uniform_tiles() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("AGB_GEMM_UNIFORM");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v != 0;
}

static void pick_n_tiles(int N, int cg, int* n_tile, int* n_tiles, int* n_last) {
  if (N % kMaxNTile == 0 || N < kMaxNTile || uniform_tiles() || (N % kMaxNTile) % (16 * cg) != 0) {
    int t = kMaxNTile;
    for (; t >= 16; t -= 16)
      if (N % t == 0) break;
    *n_tile = t < 16 ? 0 : t;
    *n_tiles = t < 16 ? 0 : N / t;
    *n_last = *n_tile;
    return;
  }
  *n_tile = kMaxNTile;
  *n_tiles = N / kMaxNTile + 1;
  *n_last = N % kMaxNTile;
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
    const bool head_act = o->act == AG_ACT_SOFTPLUS_MUL || o->act == AG_ACT_SIGMOID;
    if (o->act != AG_ACT_NONE && o->act != AG_ACT_GELU1702 && !(head_act && o->dtype == 1))
      return set_error("ag_gemm_fwd: epilogue activations are AG_ACT_NONE / AG_ACT_GELU1702, or AG_ACT_SOFTPLUS_MUL / AG_ACT_SIGMOID on fp32 outputs (got act %d, dtype %d)", o->act, o->dtype);
  }
  if ((reinterpret_cast<uintptr_t>(a.pre_scale) | reinterpret_cast<uintptr_t>(a.pre_shift)) & 15)
    return set_error("ag_gemm_fwd: pre_scale/pre_shift must be 16-byte aligned");
  if ((a.A2 == nullptr) != (a.W2 == nullptr)) return set_error("ag_gemm_fwd: A2 and W2 must be given together");
  if (a.A2) {
    if (a.w_batched || a.w_rows) return set_error("ag_gemm_fwd: the second operand pair needs a shared, unpadded W");
    if (a.K2 < 8 || a.K2 % 8 || a.a2_ld % 8 || a.w2_ld % 8 || a.a2_bstride % 8 || a.a2_rows < 1)
      return set_error("ag_gemm_fwd: K2, a2_ld, w2_ld, a2_bstride must be multiples of 8 and a2_rows >= 1");
    if ((reinterpret_cast<uintptr_t>(a.A2) | reinterpret_cast<uintptr_t>(a.W2)) & 15) return set_error("ag_gemm_fwd: A2/W2 must be 16-byte aligned");
  }

  GemmKParams p;
  memset(&p, 0, sizeof(p));
  p.g = a;
  if (a.cta_group < 0 || a.cta_group > 2) return set_error("ag_gemm_fwd: cta_group=%d must be 0, 1 or 2", a.cta_group);
  int cg = a.cta_group ? a.cta_group : forced_cg();
  if (cg == 0) cg = a.M > kBlockM ? 2 : 1;
  pick_n_tiles(a.N, cg, &p.n_tile, &p.n_tiles, &p.n_last);
  if (!p.n_tile) return set_error("ag_gemm_fwd: no N tile for N=%d", a.N);
  p.m_tiles = (a.M + kBlockM * cg - 1) / (kBlockM * cg);
  p.total_tiles = p.m_tiles * p.n_tiles * a.batch;
  p.kblocks = (a.K + kBlockK - 1) / kBlockK;
  p.err = device_error_word(device);

  {
    const bool a_shared = a.batch > 1 && a.a_bstride == 0;   // one A for every batch (e.g. a weight as the row operand)
    p.a_bmask = a_shared ? 0 : -1;
    uint64_t dims[3] = {(uint64_t)a.K, (uint64_t)a.a_rows, (uint64_t)(a_shared ? 1 : a.batch)};
    uint64_t strides[2] = {(uint64_t)a.a_ld * 2, (uint64_t)(a.batch > 1 && !a_shared ? a.a_bstride : (int64_t)a.a_ld * a.a_rows) * 2};
    uint32_t box[3] = {kBlockK, kBlockM, 1};
    if (int rc = make_tensor_map_bf16(&p.tmA, a.A, 3, dims, strides, box)) return rc;
  }
  {
    const int nz = a.w_batched ? a.batch : a.taps;
    uint64_t dims[3] = {(uint64_t)a.K, (uint64_t)(a.w_rows > 0 ? a.w_rows : a.N), (uint64_t)nz};
    uint64_t strides[2] = {(uint64_t)a.w_ld * 2, (uint64_t)(nz > 1 ? a.w_zstride : (int64_t)a.w_ld * (a.w_rows > 0 ? a.w_rows : a.N)) * 2};
    uint32_t box[3] = {kBlockK, (uint32_t)(p.n_tile / cg), 1};   // a CTA of a pair stages half of the W tile
    if (int rc = make_tensor_map_bf16(&p.tmB, a.W, 3, dims, strides, box)) return rc;
    box[1] = (uint32_t)(p.n_last / cg);
    if (int rc = make_tensor_map_bf16(&p.tmB2, a.W, 3, dims, strides, box)) return rc;
  }
  if (a.A2) {
    p.kblocks2 = (a.K2 + kBlockK - 1) / kBlockK;
    {
      uint64_t dims[3] = {(uint64_t)a.K2, (uint64_t)a.a2_rows, (uint64_t)a.batch};
      uint64_t strides[2] = {(uint64_t)a.a2_ld * 2, (uint64_t)(a.batch > 1 ? a.a2_bstride : (int64_t)a.a2_ld * a.a2_rows) * 2};
      uint32_t box[3] = {kBlockK, kBlockM, 1};
      if (int rc = make_tensor_map_bf16(&p.tmA2, a.A2, 3, dims, strides, box)) return rc;
    }
    {
      uint64_t dims[3] = {(uint64_t)a.K2, (uint64_t)a.N, 1};
      uint64_t strides[2] = {(uint64_t)a.w2_ld * 2, (uint64_t)a.w2_ld * a.N * 2};
      uint32_t box[3] = {kBlockK, (uint32_t)(p.n_tile / cg), 1};
      if (int rc = make_tensor_map_bf16(&p.tmW2, a.W2, 3, dims, strides, box)) return rc;
      box[1] = (uint32_t)(p.n_last / cg);
      if (int rc = make_tensor_map_bf16(&p.tmW2l, a.W2, 3, dims, strides, box)) return rc;
    }
  }

  const int flags = (a.res ? F_RES : 0) | (a.out ? F_OUT : 0) | (a.actA.ptr ? F_ACTA : 0) | (a.actB.ptr ? F_ACTB : 0) |
                    ((a.pool || a.actP.ptr) ? F_POOL : 0);
  if (flags == 0) return set_error("ag_gemm_fwd: no output requested");
  // 32-column epilogue chunks whenever every N tile is a multiple of 32 wide (env AGB_GEMM_EW=16 / 32w1: A/B timing)
  static int ew_mode = -1;   // 0 = wide whenever possible, 1 = never, 2 = only for width-1 launches
  if (ew_mode < 0) {
    const char* e = getenv("AGB_GEMM_EW");
    ew_mode = !e ? 0 : (e[0] == '1' ? 1 : (e[0] == '3' && e[2] == 'w' ? 2 : 0));
  }
  const int wide = (p.n_tile % 32 == 0 && p.n_last % 32 == 0 && ew_mode != 1 && (ew_mode != 2 || a.taps == 1)) ? 1 : 0;
  GemmKernelFn fn = wide ? gemm_wide_kernel(cg, flags) : kGemmTable.fn[cg - 1][flags];
  const int smem = cg == 2 ? (wide ? gemm_smem_bytes<2, 32>() : gemm_smem_bytes<2, 16>()) : (wide ? gemm_smem_bytes<1, 32>() : gemm_smem_bytes<1, 16>());
  static thread_local unsigned char attr_done[64][2][2][F_COUNT];
  if (!attr_done[device][cg - 1][wide][flags]) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error("ag_gemm_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_done[device][cg - 1][wide][flags] = 1;
  }
  const int units = sm_count(device) / cg;   // persistent: one CTA (or CTA pair) per SM (or TPC)
  const int grid = (p.total_tiles < units ? p.total_tiles : units) * cg;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(kGemmThreads, 1, 1);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cg;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
#ifdef AGB_TRACE
  const char* trace_path = getenv("AGB_GEMM_TRACE");
  if (trace_path) {
    cudaMalloc(&p.trace, 3 * 2048 * sizeof(unsigned long long));
    cudaMemsetAsync(p.trace, 0, 3 * 2048 * sizeof(unsigned long long), stream);
  }
#endif
  cudaError_t le = cudaLaunchKernelEx(&cfg, fn, p);
  if (le != cudaSuccess) return set_error("ag_gemm_fwd: launch failed: %s", cudaGetErrorString(le));
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
  return check_launch("ag_gemm_fwd");
}

}  // namespace agb
