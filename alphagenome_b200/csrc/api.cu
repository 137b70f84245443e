// C-ABI surface of libagb200.so (declared in include/agb200.h) and the small host runtime behind it:
// thread-local error text, launch accounting, per-device watchdog word, TMA descriptor encoding.
#include <atomic>
#include <mutex>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "kernels.h"

namespace agb {

static thread_local char tl_error[1024] = "";
static std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_gemm_cta_group{0};   // see ag_gemm_set_cta_group
static std::mutex g_mu;
static unsigned int* g_err_host[64] = {nullptr};
static unsigned int* g_err_dev[64] = {nullptr};
static int g_sm_count[64] = {0};
static int g_cc[64] = {0};

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("AGB_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

int set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tl_error, sizeof(tl_error), fmt, ap);
  va_end(ap);
  return 1;
}

static int ensure_device(int device) {
  if (device < 0 || device >= 64) return set_error("bad device index %d", device);
  int cur = -1;
  cudaError_t e = cudaGetDevice(&cur);
  if (e != cudaSuccess) return set_error("cudaGetDevice: %s", cudaGetErrorString(e));
  if (cur != device) {
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return set_error("cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
  }
  if (g_sm_count[device] == 0) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_sm_count[device] == 0) {
      cudaDeviceProp prop;
      e = cudaGetDeviceProperties(&prop, device);
      if (e != cudaSuccess) return set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e));
      g_cc[device] = prop.major * 10 + prop.minor;
      void* h = nullptr;
      e = cudaHostAlloc(&h, 64, cudaHostAllocMapped | cudaHostAllocPortable);
      if (e != cudaSuccess) return set_error("cudaHostAlloc(watchdog): %s", cudaGetErrorString(e));
      memset(h, 0, 64);
      void* d = nullptr;
      e = cudaHostGetDevicePointer(&d, h, 0);
      if (e != cudaSuccess) return set_error("cudaHostGetDevicePointer: %s", cudaGetErrorString(e));
      g_err_host[device] = static_cast<unsigned int*>(h);
      g_err_dev[device] = static_cast<unsigned int*>(d);
      g_sm_count[device] = prop.multiProcessorCount;
    }
  }
  if (g_cc[device] != 100) return set_error("device %d is sm_%d; libagb200 is built for sm_100a (B200) only", device, g_cc[device]);
  return 0;
}

// Makes `device` current for the duration of one entry point and restores the caller's device afterwards (the library
// must not flip the current device under a host framework that tracks it).
struct DeviceGuard {
  int prev = -1;
  int rc = 0;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    rc = ensure_device(device);
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

unsigned int* device_error_word(int device) { return g_err_dev[device]; }
int sm_count(int device) { return g_sm_count[device]; }

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
  return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

int make_tensor_map_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box) {
  if (!g_encode) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_encode) {
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult qres;
      cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
      if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
        return set_error("cuTensorMapEncodeTiled entry point unavailable: %s", cudaGetErrorString(e));
      g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
  }
  cuuint64_t gdims[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    bx[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstr,
                        bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return set_error("cuTensorMapEncodeTiled failed (%d): rank=%d dims=[%llu,%llu,%llu] strides=[%llu,%llu] box=[%u,%u,%u] base=%p",
                     (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                     (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)strides_bytes[0],
                     (unsigned long long)(rank > 2 ? strides_bytes[1] : 0), box[0], rank > 1 ? box[1] : 0,
                     rank > 2 ? box[2] : 0, base);
  }
  return 0;
}

}  // namespace agb

using namespace agb;

extern "C" {

int ag_version(void) { return AGB200_VERSION; }
const char* ag_last_error(void) { return tl_error; }
uint64_t ag_launch_count(void) { return g_launches.load(); }

int ag_check_device(int device) {
  DeviceGuard guard(device);
  return guard.rc;
}

int ag_watchdog_status(int device, uint32_t out[4]) {
  if (device < 0 || device >= 64 || !g_err_host[device]) {
    out[0] = out[1] = out[2] = out[3] = 0;
    return 0;
  }
  for (int i = 0; i < 4; ++i) out[i] = g_err_host[device][i];
  return out[0] != 0;
}

int ag_gemm_set_cta_group(int cta_group) {
  if (cta_group < 0 || cta_group > 2) return -1;
  return agb::g_gemm_cta_group.exchange(cta_group);
}

int ag_gemm_fwd(const AgGemmArgs* args, int device, void* stream) {
  if (!args) return set_error("ag_gemm_fwd: null args");
  DeviceGuard guard(device);
  if (guard.rc) return guard.rc;
  return launch_gemm(*args, device, static_cast<cudaStream_t>(stream));
}

int ag_attention_fwd(const AgAttnArgs* args, int device, void* stream) {
  if (!args) return set_error("ag_attention_fwd: null args");
  DeviceGuard guard(device);
  if (guard.rc) return guard.rc;
  return launch_attention(*args, device, static_cast<cudaStream_t>(stream));
}

int ag_onehot_im2col_fwd(const int64_t* tokens, void* out_bf16, int32_t B, int32_t L, int32_t width, int32_t n_base,
                         int32_t ld, int device, void* stream) {
  DeviceGuard guard(device);
  if (guard.rc) return guard.rc;
  return launch_onehot_im2col(tokens, out_bf16, B, L, width, n_base, ld, device, static_cast<cudaStream_t>(stream));
}

int ag_qkv_post_fwd(const float* qkv, void* Q, void* K, void* V, const float* cos_tab, const float* sin_tab,
                    const float* q_w, const float* q_b, const float* k_w, const float* k_b, const float* v_w,
                    const float* v_b, int32_t B, int32_t n, int32_t heads, int32_t d, int32_t dv, int device,
                    void* stream) {
  DeviceGuard guard(device);
  if (guard.rc) return guard.rc;
  return launch_qkv_post(qkv, Q, K, V, cos_tab, sin_tab, q_w, q_b, k_w, k_b, v_w, v_b, B, n, heads, d, dv, device,
                         static_cast<cudaStream_t>(stream));
}

int ag_transpose_bf16(const void* in, void* out, int32_t nb, int32_t R, int32_t C, int64_t in_bstride, int32_t in_ld,
                      int64_t out_bstride, int32_t out_ld, int device, void* stream) {
  DeviceGuard guard(device);
  if (guard.rc) return guard.rc;
  return launch_transpose_bf16(in, out, nb, R, C, in_bstride, in_ld, out_bstride, out_ld, device,
                               static_cast<cudaStream_t>(stream));
}

int ag_pair_bias_fwd(const float* pair, const float* scale, const float* shift, const float* W, float* bias, int32_t B,
                     int32_t rows, int32_t cols, int32_t C, int32_t heads, int32_t sets, int64_t set_stride, int device,
                     void* stream) {
  DeviceGuard guard(device);
  if (guard.rc) return guard.rc;
  return launch_pair_bias(pair, scale, shift, W, bias, B, rows, cols, C, heads, sets, set_stride, device,
                          static_cast<cudaStream_t>(stream));
}

int ag_pool_rmsnorm_fwd(const float* single, const float* w, const float* b, void* x_bf16, void* xg_bf16, int32_t B,
                        int32_t n, int32_t C, int32_t pool, int device, void* stream) {
  DeviceGuard guard(device);
  if (guard.rc) return guard.rc;
  return launch_pool_rmsnorm(single, w, b, x_bf16, xg_bf16, B, n, C, pool, device, static_cast<cudaStream_t>(stream));
}

int ag_pair_combine_fwd(const float* qk, const float* relq, const float* relk, const float* Wp, const float* bp,
                        const float* outer, const float* prev, float* pair, const float* norm_w, const float* norm_b,
                        void* pair_norm_bf16, int32_t B, int32_t P, int32_t H, int32_t D, int32_t qk_ld, int32_t rel_ld,
                        int32_t rows, int32_t i_off, int device, void* stream) {
  DeviceGuard guard(device);
  if (guard.rc) return guard.rc;
  return launch_pair_combine(qk, relq, relk, Wp, bp, outer, prev, pair, norm_w, norm_b, pair_norm_bf16, B, P, H, D, qk_ld,
                             rel_ld, rows, i_off, device, static_cast<cudaStream_t>(stream));
}

int ag_rmsnorm_fwd(const float* x, const float* w, const float* b, void* out_bf16, int64_t rows, int32_t C, int device,
                   void* stream) {
  DeviceGuard guard(device);
  if (guard.rc) return guard.rc;
  return launch_rmsnorm(x, w, b, out_bf16, rows, C, device, static_cast<cudaStream_t>(stream));
}

int ag_add_embed_norm_fwd(const float* x, const float* emb, const float* scale, const float* shift, float* single,
                          void* a_bf16, int32_t B, int32_t rows, int32_t C, int device, void* stream) {
  DeviceGuard guard(device);
  if (guard.rc) return guard.rc;
  return launch_add_embed_norm(x, emb, scale, shift, single, a_bf16, B, rows, C, device,
                               static_cast<cudaStream_t>(stream));
}

int ag_pair_out_embed_fwd(const float* pair, const float* pairT, const float* w, const float* b, const float* emb,
                          float* out, int32_t B, int32_t rows, int32_t P, int32_t C, int device, void* stream) {
  DeviceGuard guard(device);
  if (guard.rc) return guard.rc;
  return launch_pair_out_embed(pair, pairT, w, b, emb, out, B, rows, P, C, device, static_cast<cudaStream_t>(stream));
}

int ag_rmsnorm_act_fwd(const float* x, const float* w, const float* b, const float* add, void* out_bf16, int64_t rows,
                       int64_t rows_per_batch, int32_t C, int32_t act, int device, void* stream) {
  DeviceGuard guard(device);
  if (guard.rc) return guard.rc;
  return launch_rmsnorm_act(x, w, b, add, out_bf16, rows, rows_per_batch, C, act, device, static_cast<cudaStream_t>(stream));
}

int ag_splice_rope_fwd(const float* x, const float* scale, const float* offset, const float* cos_tab, const float* sin_tab,
                       void* out_bf16, int32_t B, int32_t P, int32_t T, int32_t hidden, int32_t p_ld, int32_t tab_per_site,
                       int device, void* stream) {
  DeviceGuard guard(device);
  if (guard.rc) return guard.rc;
  return launch_splice_rope(x, scale, offset, cos_tab, sin_tab, out_bf16, B, P, T, hidden, p_ld, tab_per_site, device,
                            static_cast<cudaStream_t>(stream));
}

int ag_qkv_post_gather_fwd(const float* qkv, void* Q, void* const* kv_peers, int32_t world, int64_t n_all, int64_t row0,
                           const float* cos_tab, const float* sin_tab, const float* q_w, const float* q_b, const float* k_w,
                           const float* k_b, const float* v_w, const float* v_b, int32_t B, int32_t n, int32_t heads, int32_t d,
                           int32_t dv, int device, void* stream) {
  DeviceGuard guard(device);
  if (guard.rc) return guard.rc;
  return launch_qkv_post_gather(qkv, Q, kv_peers, world, n_all, row0, cos_tab, sin_tab, q_w, q_b, k_w, k_b, v_w, v_b, B, n, heads, d, dv,
                                device, static_cast<cudaStream_t>(stream));
}

int ag_peer_push_fwd(const void* src, void* const* peers, int32_t world, int64_t bytes, int64_t dst_offset, int device, void* stream) {
  DeviceGuard guard(device);
  if (guard.rc) return guard.rc;
  return launch_peer_push(src, peers, world, bytes, dst_offset, device, static_cast<cudaStream_t>(stream));
}

int ag_bn_stats_fwd(const float* x, double* sums, int32_t B, int32_t rows, int64_t bstride, int32_t ld, int32_t C, int device,
                    void* stream) {
  DeviceGuard guard(device);
  if (guard.rc) return guard.rc;
  return launch_bn_stats(x, sums, B, rows, bstride, ld, C, device, static_cast<cudaStream_t>(stream));
}

int ag_bn_update_fwd(const double* sums, double count, float* running_var, const float* gamma, float momentum, float eps,
                     float* scale_out, const float* fold_bias, const float* fold_beta, float* fold_shift_out, int32_t C,
                     int device, void* stream) {
  DeviceGuard guard(device);
  if (guard.rc) return guard.rc;
  return launch_bn_update(sums, count, running_var, gamma, momentum, eps, scale_out, fold_bias, fold_beta, fold_shift_out, C,
                          device, static_cast<cudaStream_t>(stream));
}

int ag_affine_act_fwd(const AgAffineArgs* args, int device, void* stream) {
  if (!args) return set_error("ag_affine_act_fwd: null args");
  DeviceGuard guard(device);
  if (guard.rc) return guard.rc;
  return launch_affine_act(*args, device, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
