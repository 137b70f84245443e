/*
 * agb200.h — C ABI of libagb200.so: the sm_100a (B200) kernels behind the AlphaGenome trunk forward.
 *
 * The reference (lucidrains/alphagenome, alphagenome_pytorch/alphagenome.py, "AG" below) has no FFI or
 * plugin registry: its operator interface for this path is the nn.Module classes instantiated by
 * TransformerUnet (AG:1054-1093) and called from TransformerUnet.forward (AG:1095-1136).  Each entry
 * point below replaces the arithmetic of one group of those module forwards; the Python host in
 * alphagenome_b200/ keeps the modules' names, constructor signatures, forward signatures and
 * state_dict keys, and calls these functions with raw device pointers (tensor.data_ptr()) and the
 * current CUDA stream.  INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on device `device` unless stated otherwise;
 *   - the library never allocates, frees or retains memory past a call: outputs and workspaces are
 *     caller-owned (SURVEY.md 8b "Ownership");
 *   - calls are asynchronous and ordered on `stream` (a cudaStream_t passed as void*); no host sync;
 *   - return 0 on success; non-zero => ag_last_error() (thread-local text) says why.  Unsupported
 *     shapes/dtypes/architectures are errors, never a silent fallback;
 *   - activations are channels-last: a "row" is one sequence position (or one pair cell), channels
 *     are contiguous.  bf16 buffers feed the tensor cores, fp32 buffers carry the residual stream.
 */
#ifndef AGB200_H
#define AGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGB200_VERSION 1

/* activation selector used by every fused epilogue */
enum AgAct {
  AG_ACT_NONE = 0,
  AG_ACT_GELU1702 = 1, /* x * sigmoid(1.702 x)            AG:130-137 */
  AG_ACT_GELU_ERF = 2, /* nn.GELU() (exact erf)           AG:690, AG:810 */
  AG_ACT_RELU = 3      /* nn.ReLU()                       AG:903 */
};

/* One "activated" epilogue output:  dst[b, row, c] = act( v * scale[c] + shift[b?, c] )
 * (a BatchRMSNorm in eval mode is exactly such a per-channel affine: AG:344-361). */
typedef struct AgEpiOut {
  void* ptr;            /* NULL = output disabled */
  const float* scale;   /* [N] or NULL (=1) */
  const float* shift;   /* [N] (+ b*shift_bstride) or NULL (=0) */
  int64_t bstride;      /* elements between batches of dst */
  int64_t shift_bstride;/* elements between batches of shift (0 = shared) */
  int32_t ld;           /* elements between rows of dst */
  int32_t act;          /* enum AgAct */
  int32_t dtype;        /* 0 = bf16, 1 = fp32 */
  int32_t reserved;
} AgEpiOut;

/* ag_gemm_fwd: implicit-GEMM conv1d (taps 1 or 5, "same" zero padding) / linear / batched matmul on
 * tcgen05 tensor cores with a fused epilogue.  Replaces ConvBlock's Conv1d (AG:431-452) inside
 * DNAEmbed.pointwise, DownresBlock and UpresBlock (AG:454-519, 1163), and every nn.Linear on the
 * path (Attention.to_qkv/to_out AG:676-677, FeedForward AG:893-906, SingleToPairwise.to_qk /
 * to_outer_sum / the q.k and relative-position einsums AG:832-846, PairwiseRowAttention.to_qk/to_v
 * AG:868-869, OutputEmbedding.double_features/skip_proj AG:1207-1208).
 *
 *   acc[b, r, n] = sum_{t<taps} sum_{k<K} A[b, r + t - taps/2, k] * W[z(b,t), n, k]        (fp32 accumulate)
 *   v = acc * pre_scale[n] + pre_shift[n];  if relu: v = max(v, 0)
 *   if res && n < res_ch: v += res[b, r >> res_shift, n]
 *   if post_scale: v *= *post_scale
 *   out[b, r, n] = v;  actA / actB[b, r, n] = act(v * scale + shift)
 *   pool[b, r/2, n] = max(v[r], v[r^1]);  actP[b, r/2, n] = act(pool * scale + shift)      (maxpool2, AG:468)
 *
 * A rows outside [0, a_rows) read as zero (the conv's zero padding, AG:448); rows are addressed as
 * buffer row r + a_row0 so a caller can keep halo rows in front of the data (sequence-parallel shards).
 */
typedef struct AgGemmArgs {
  const void* A;      /* bf16 */
  int64_t a_bstride;  /* elements between batches */
  int32_t a_ld;       /* elements between rows (multiple of 8) */
  int32_t a_rows;     /* rows per batch present in the buffer (incl. halo rows) */
  int32_t a_row0;     /* buffer row of logical row 0 */
  int32_t w_batched;  /* 0: W is [taps][N][K] shared by all batches; 1: W is [batch][N][K] (taps must be 1) */
  const void* W;      /* bf16 */
  int64_t w_zstride;  /* elements between taps (or between batches if w_batched) */
  int32_t w_ld;       /* elements between the N rows of W (multiple of 8) */
  int32_t batch;
  int32_t M;          /* output rows per batch */
  int32_t N;          /* output channels (multiple of 16) */
  int32_t K;          /* input channels (multiple of 8) */
  int32_t taps;       /* 1 or 5 (any odd value <= 15 is accepted) */
  const float* pre_scale; /* [N] or NULL */
  const float* pre_shift; /* [N] or NULL (bias, possibly folded with a post-norm) */
  int32_t relu;
  int32_t res_ch;     /* residual is added to channels n < res_ch (zero-padded residual, AG:472) */
  const float* res;   /* fp32 or NULL */
  int64_t res_bstride;
  int32_t res_ld;
  int32_t res_shift;  /* residual row = r >> res_shift (nearest upsample, AG:513 / AG:1227) */
  const float* post_scale; /* device scalar or NULL (UpresBlock.residual_scale, AG:501) */
  float* out;         /* fp32 or NULL */
  int64_t out_bstride;
  int32_t out_ld;
  int32_t pool_ld;
  float* pool;        /* fp32 [batch][M/2][N] or NULL */
  int64_t pool_bstride;
  AgEpiOut actA, actB, actP;
} AgGemmArgs;

int ag_gemm_fwd(const AgGemmArgs* args, int device, void* stream);

/* library / error plumbing */
int ag_version(void);
const char* ag_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches claim) */
uint64_t ag_launch_count(void);
/* 0 if `device` is an sm_100 part and the kernels can be loaded, else non-zero with ag_last_error() */
int ag_check_device(int device);

#ifdef __cplusplus
}
#endif
#endif /* AGB200_H */
