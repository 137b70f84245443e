// Internal host-side declarations shared by the .cu files of libagb200.so (not part of the C ABI).
#pragma once
#include <atomic>

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/agb200.h"

namespace agb {

// thread-local error text; always returns non-zero so callers can `return set_error(...)`
int set_error(const char* fmt, ...);
// cudaGetLastError() after a launch; bumps the launch counter; formats watchdog info if present
int check_launch(const char* what);
// host-mapped pinned word block [4] the device watchdog writes before trapping (device pointer)
unsigned int* device_error_word(int device);
int sm_count(int device);
// tiled, 128B-swizzled bf16 tensor map (rank 2..4); dims[0] is the contiguous dimension
int make_tensor_map_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box);

// programmatic dependent launch of the GEMM / attention kernels (env AGB_PDL=0 disables)
bool pdl_enabled();
extern std::atomic<int> g_gemm_cta_group;   // 0 auto, 1, 2 (ag_gemm_set_cta_group)
int launch_gemm(const AgGemmArgs& a, int device, cudaStream_t stream);
int launch_attention(const AgAttnArgs& a, int device, cudaStream_t stream);
int launch_onehot_im2col(const int64_t* tokens, void* out, int B, int L, int width, int n_base, int ld, int device,
                         cudaStream_t st);
int launch_qkv_post(const float* qkv, void* Q, void* K, void* V, const float* cos_tab, const float* sin_tab,
                    const float* q_w, const float* q_b, const float* k_w, const float* k_b, const float* v_w,
                    const float* v_b, int B, int n, int heads, int d, int dv, int device, cudaStream_t st);
int launch_transpose_bf16(const void* in, void* out, int nb, int R, int C, long in_bs, int in_ld, long out_bs,
                          int out_ld, int device, cudaStream_t st);
int launch_pair_bias(const float* pair, const float* scale, const float* shift, const float* W, float* bias, int B,
                     int rows, int cols, int C, int heads, int sets, long set_stride, int device, cudaStream_t st);
int launch_pool_rmsnorm(const float* single, const float* w, const float* b, void* x, void* xg, int B, int n, int C,
                        int pool, int device, cudaStream_t st);
int launch_pair_combine(const float* qk, const float* relq, const float* relk, const float* Wp, const float* bp,
                        const float* outer, const float* prev, float* pair, const float* norm_w, const float* norm_b,
                        void* pn, int B, int P, int H, int D, int qk_ld, int rel_ld, int rows, int i_off, int device,
                        cudaStream_t st);
int launch_rmsnorm(const float* x, const float* w, const float* b, void* out, long rows, int C, int device,
                   cudaStream_t st);
int launch_add_embed_norm(const float* x, const float* emb, const float* scale, const float* shift, float* single,
                          void* a, int B, int rows, int C, int device, cudaStream_t st);
int launch_pair_out_embed(const float* pair, const float* pairT, const float* w, const float* b, const float* emb,
                          float* out, int B, int rows, int P, int C, int device, cudaStream_t st);

int launch_rmsnorm_act(const float* x, const float* w, const float* b, const float* add, void* out, long rows,
                       long rows_per_batch, int C, int act, int device, cudaStream_t st);
int launch_splice_rope(const float* x, const float* scale, const float* offset, const float* cos_tab, const float* sin_tab,
                       void* out, int B, int P, int T, int Hd, int p_ld, int tab_per_site, int device, cudaStream_t st);

int launch_qkv_post_gather(const float* qkv, void* Q, void* const* kv_peers, int world, long n_all, long row0, const float* cos_tab,
                           const float* sin_tab, const float* q_w, const float* q_b, const float* k_w, const float* k_b,
                           const float* v_w, const float* v_b, int B, int n, int heads, int d, int dv, int device, cudaStream_t st);
int launch_peer_push(const void* src, void* const* peers, int world, long bytes, long dst_offset, int device, cudaStream_t st);
int launch_bn_stats(const float* x, double* sums, int B, int rows, long bstride, int ld, int C, int device, cudaStream_t st);
int launch_bn_update(const double* sums, double count, float* running_var, const float* gamma, float momentum, float eps,
                     float* scale_out, const float* fold_bias, const float* fold_beta, float* fold_shift_out, int C, int device,
                     cudaStream_t st);
int launch_affine_act(const AgAffineArgs& a, int device, cudaStream_t st);

}  // namespace agb
