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
  AG_ACT_RELU = 3,     /* nn.ReLU()                       AG:903 */
  /* prediction heads (ag_gemm_fwd fp32 outputs only): */
  AG_ACT_SOFTPLUS_MUL = 4, /* softplus(v + shift[c]) * scale[c]   TracksScaledPrediction AG:1388-1393 (scale = softplus(param)) */
  AG_ACT_SIGMOID = 5       /* sigmoid(v * scale[c] + shift[c])    SpliceSiteUsage AG:1321 */
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
 *   (+ an optional second operand pair A2 . W2^T accumulated into acc, see the end of the struct)
 *   v = acc * pre_scale[n] + pre_shift[n];  if relu: v = max(v, 0)
 *   if res && n < res_ch: v += res[b, r >> res_shift, n]
 *   if post_scale: v *= *post_scale
 *   out[b, r, n] = v;  actA / actB[b, r, n] = act(v * scale + shift)        (act: AG_ACT_NONE or AG_ACT_GELU1702;
 *                                                                           fp32 outputs also AG_ACT_SOFTPLUS_MUL, AG_ACT_SIGMOID)
 *   pool[b, r/2, n] = max(v[r], v[r^1]);  actP[b, r/2, n] = act(pool * scale + shift)      (maxpool2, AG:468)
 *
 * A rows outside [0, a_rows) read as zero (the conv's zero padding, AG:448); rows are addressed as
 * buffer row r + a_row0 so a caller can keep halo rows in front of the data (sequence-parallel shards).
 */
typedef struct AgGemmArgs {
  const void* A;      /* bf16 */
  int64_t a_bstride;  /* elements between batches; 0 with batch > 1 = ONE A shared by every batch */
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
  int32_t w_rows;     /* rows of W present in memory (0 = N); rows in [w_rows, N) read as zero, so N can be padded to 16 */
  int32_t cta_group;     /* CTAs per tcgen05.mma for THIS call: 0 = the process default (automatic unless ag_gemm_set_cta_group
                          * changed it), 1 = single-CTA MMAs, 2 = CTA pairs (cta_group::2).  Same results either way. */
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
  /* Optional SECOND operand pair accumulated into the same tile before the epilogue (A2 == NULL: none):
   *   acc[b, r, n] += sum_{k<K2} A2[b, r, k] * W2[n, k]          (one tap; rows outside [0, a2_rows) read as zero)
   * DNAEmbed (AG:1177-1178) uses it to add the width-15 convolution  x0 = conv(one_hot)  — a K2 = 64 im2col product —
   * to pointwise(x0)'s accumulator on the tensor core, so the 1 bp fp32 x0 is neither written nor re-read as a residual. */
  const void* A2;     /* bf16 [batch][a2_rows][K2], addressed like A (buffer row r + a2_row0) */
  const void* W2;     /* bf16 [N][K2] shared by every batch */
  int64_t a2_bstride;
  int32_t a2_ld, a2_rows, a2_row0;
  int32_t w2_ld;      /* elements between the N rows of W2 (multiple of 8) */
  int32_t K2;         /* multiple of 8 */
  int32_t reserved2;
} AgGemmArgs;

int ag_gemm_fwd(const AgGemmArgs* args, int device, void* stream);

/* ag_attention_fwd: fused attention (QK^T -> softmax -> .V) with tcgen05; replaces the materialised
 * einsum/softmax/einsum of Attention.forward (AG:748-779; mode 0: + pair bias upsampled x16, tanh soft-cap)
 * and of PairwiseRowAttention.forward (AG:880-889; mode 1: plain softmax, fp32 out = result + out_bias + res).
 *   Q  bf16 [batch][heads][n_q][d]      addressed as b*q_bstride + h*q_hstride + i*q_ld + c
 *   K  bf16 [batch][kv_heads][n_k][d]   (kv_heads == 1: multi-query, shared by all heads)
 *   Vt bf16 [batch][kv_heads][dv][n_k]  (values TRANSPOSED: keys contiguous; v_ld = elements between dv rows)
 *   bias fp32 [batch][heads][bias_rows][bias_p]; rows index queries/16, columns keys/16     (mode 0 only)
 *   out [batch][heads][n_q][dv] addressed as b*out_bstride + h*out_hstride + i*out_ld + c; bf16 or fp32
 */
typedef struct AgAttnArgs {
  const void* Q;
  const void* K;
  const void* Vt;
  int64_t q_bstride, q_hstride;
  int64_t k_bstride, k_hstride;
  int64_t v_bstride, v_hstride;
  int32_t q_ld, k_ld, v_ld;
  int32_t batch, heads, kv_heads, n_q, n_k, d, dv;
  int32_t mode;       /* 0 = soft-capped + bias (AG:752-765), 1 = plain softmax (AG:880-885) */
  int32_t bias_p;     /* bias columns (>= n_k/16) */
  float scale;        /* d^-0.5 */
  float softcap;      /* 5.0 (AG:653) */
  const float* bias;
  void* out;
  int64_t out_bstride, out_hstride;
  int32_t out_ld;
  int32_t out_dtype;  /* 0 = bf16, 1 = fp32 */
  const float* out_bias; /* [dv] fp32 or NULL (fp32 out only) */
  const float* res;      /* fp32, same addressing as out, or NULL (fp32 out only; may alias out) */
  int32_t bias_rows;     /* bias rows (>= n_q/16); 0 = bias_p.  bias is [batch][heads][bias_rows][bias_p] */
  int32_t reserved1;
  /* optional (fp32 out, dv == 128): RMSNormWithBias (AG:400-405) of each finished output row, bf16, addressed like out —
   * the pre-norm of the pair feed-forward that follows the row attention (NormWrapper, AG:972-975) */
  const float* norm_w;   /* [dv] or NULL */
  const float* norm_b;   /* [dv] */
  void* norm_out;        /* bf16 or NULL */
} AgAttnArgs;

/* mode 1 with norm_out (the trunk's call) runs a persistent single-pass kernel: online softmax with a running row maximum
 * and a lazy rescale of the accumulator — the exact softmax, any n_k; without norm_out the two-pass kernel (n_k <= 512
 * keys stay resident between its passes, more are re-streamed). */
int ag_attention_fwd(const AgAttnArgs* args, int device, void* stream);

/* --- vectorised HBM-bound kernels ------------------------------------------------------------------- */

/* tokens int64 [B][L] (A,C,G,T = 0..3; negative = N -> all-zero row, AG:1171-1174) -> bf16 one-hot im2col
 * [B][L][ld] with column t*n_base + base = 1 iff token[l + t - width/2] == base; columns >= width*n_base are 0.
 * Feeding this to ag_gemm_fwd with the packed [C][ld] weight is DNAEmbed.conv (AG:1162, 1177). */
int ag_onehot_im2col_fwd(const int64_t* tokens, void* out_bf16, int32_t B, int32_t L, int32_t width,
                         int32_t n_base, int32_t ld, int device, void* stream);

/* qkv fp32 [B][n][heads*d + d + dv] -> LayerNorm(eps 1e-5) per head slice, interleaved RoPE on q and k
 * (AG:715-734, 559-566) -> Q bf16 [B][n][heads*d], K bf16 [B][n][d], V bf16 [B][n][dv].
 * cos/sin: fp32 [n][d/2] tables of the angles position*inv_freq (AG:545-557). norm_wb: q_w,q_b,k_w,k_b,v_w,v_b. */
int ag_qkv_post_fwd(const float* qkv, void* Q, void* K, void* V, const float* cos_tab, const float* sin_tab,
                    const float* q_w, const float* q_b, const float* k_w, const float* k_b, const float* v_w,
                    const float* v_b, int32_t B, int32_t n, int32_t heads, int32_t d, int32_t dv, int device,
                    void* stream);

/* bf16 [nb][R][C] (row stride in_ld, batch stride in_bstride) -> bf16 [nb][C][out_ld] with out[b][c][r] = in[b][r][c] */
int ag_transpose_bf16(const void* in, void* out, int32_t nb, int32_t R, int32_t C, int64_t in_bstride,
                      int32_t in_ld, int64_t out_bstride, int32_t out_ld, int device, void* stream);

/* to_attn_bias (AG:688-693) of `sets` (1 or 2) attention layers that read the same pair tensor, in one pass over it:
 * bias[s][b][g][i][j] = sum_c W[s][g][c] * gelu_erf(pair[b][i][j][c]*scale[s][c] + shift[s][c]);
 * pair is [B][rows][cols][C] (rows < cols for a sequence-parallel row shard); scale/shift [sets][C]; W [sets][heads][C];
 * bias holds `sets` tensors [B][heads][rows][cols] that are `set_stride` floats apart.  When rows*cols is a multiple of 16
 * the projection runs on the tensor core with bf16 operands (like every other contraction of the path); other shapes
 * take fp32 FMA kernels.  pair, scale, shift and W must be 16-byte aligned. */
int ag_pair_bias_fwd(const float* pair, const float* scale, const float* shift, const float* W, float* bias,
                     int32_t B, int32_t rows, int32_t cols, int32_t C, int32_t heads, int32_t sets, int64_t set_stride,
                     int device, void* stream);

/* SingleToPairwise prologue (AG:828-829, 852): mean over `pool` tokens -> RMSNormWithBias ->
 * x_bf16 [B][P][C] and gelu_erf(x) bf16 [B][P][C] */
int ag_pool_rmsnorm_fwd(const float* single, const float* w, const float* b, void* x_bf16, void* xg_bf16,
                        int32_t B, int32_t n, int32_t C, int32_t pool, int device, void* stream);

/* SingleToPairwise combine (AG:835-856): pair[b][i][j][:] = Wp . sim[b][i][j][:] + bp + outer[b][i][:D] +
 * outer[b][j][D:] (+ prev), with sim[..h] = qk[b][h][i][j] + 0.5*(relq[b][h][i][P+j-i] + relk[b][h][j][P+i-j]);
 * qk rows are qk_ld floats apart, relq/relk rows rel_ld floats apart (padded to the GEMM's N multiple of 16).
 * Sequence-parallel row shard: this call produces global rows i in [i_off, i_off+rows); qk, relq, prev and pair hold
 * only those rows ([B][H][rows][..] / [B][rows][P][D]); relk and outer hold all P rows.
 * If pair_norm_bf16 is non-null it also receives RMSNormWithBias(pair) (AG:400-405 with norm_w/norm_b [D]) in bf16 —
 * the pre-norm of the PairwiseRowAttention that consumes the result (NormWrapper, AG:974). */
int ag_pair_combine_fwd(const float* qk, const float* relq, const float* relk, const float* Wp, const float* bp,
                        const float* outer, const float* prev, float* pair, const float* norm_w, const float* norm_b,
                        void* pair_norm_bf16, int32_t B, int32_t P, int32_t H, int32_t D, int32_t qk_ld, int32_t rel_ld,
                        int32_t rows, int32_t i_off, int device, void* stream);

/* RMSNormWithBias over the last dim (AG:400-405): fp32 [rows][C] -> bf16 [rows][C] */
int ag_rmsnorm_fwd(const float* x, const float* w, const float* b, void* out_bf16, int64_t rows, int32_t C,
                   int device, void* stream);

/* single[b][r][c] = x[b][r][c] + emb[b][c]  (AG:1119-1120);  a_bf16 = single*scale[c] + shift[c] (pre-norm) */
int ag_add_embed_norm_fwd(const float* x, const float* emb, const float* scale, const float* shift,
                          float* single, void* a_bf16, int32_t B, int32_t rows, int32_t C, int device,
                          void* stream);

/* OutputPairEmbedding (AG:1248-1259): out = gelu1702(RMSNormWithBias(0.5*(pair[b][i][j] + pair[b][j][i])) + emb[b]).
 * pair/out are this rank's rows [B][rows][P][C]; pairT holds the transposed blocks [P/rows][B][rows][rows][C] with
 * pairT[s][b][jl][il] = pair_global[b][s*rows + jl][i_off + il] (one rank: rows == P and pairT == pair). */
int ag_pair_out_embed_fwd(const float* pair, const float* pairT, const float* w, const float* b, const float* emb,
                          float* out, int32_t B, int32_t rows, int32_t P, int32_t C, int device, void* stream);

/* --- sequence parallel over NVLink peer memory (no reference counterpart: the reference has no sequence parallelism) --- */

/* ag_qkv_post_fwd FUSED WITH THE K/V ALL-GATHER of a sequence-parallel rank (AG:748 attends over the whole context):
 * Q [B][n][heads*d] stays local; the K and V rows of this rank's n tokens are stored, as [K | V] rows of width d + dv, at
 * rows [row0, row0 + n) of EVERY rank's gathered buffer kv_peers[r] (bf16 [B][n_all][d + dv]; peer-mapped device
 * pointers of a symmetric allocation, r < world <= 8; kv_peers itself is a HOST array) — remote stores over NVLink /
 * NVSwitch issued by the warp that just finished the row.  The caller orders consumers behind a cross-rank barrier. */
int ag_qkv_post_gather_fwd(const float* qkv, void* Q, void* const* kv_peers, int32_t world, int64_t n_all, int64_t row0,
                           const float* cos_tab, const float* sin_tab, const float* q_w, const float* q_b, const float* k_w,
                           const float* k_b, const float* v_w, const float* v_b, int32_t B, int32_t n, int32_t heads, int32_t d,
                           int32_t dv, int device, void* stream);

/* One-launch push all-gather: `bytes` at src are stored at byte offset dst_offset of every rank's buffer peers[r]
 * (HOST array of peer-mapped device pointers); sizes / offsets / pointers multiples of 16 bytes. */
int ag_peer_push_fwd(const void* src, void* const* peers, int32_t world, int64_t bytes, int64_t dst_offset, int device, void* stream);

/* --- BatchRMSNorm with updating statistics (AG:324-361 with update_running_var, the reference's default; ---------
 * --- get_maybe_dist_var AG:276-300).  The frozen path is folded into ag_gemm_fwd's epilogue instead. ---------- */

/* sums[0..C) = sum over (b, row) of x[b][row][c];  sums[C..2C) = sum of squares; fp64, overwritten.  x fp32
 * [B][rows][C] with row stride ld and batch stride bstride (elements); C a multiple of 128.  A sequence-parallel rank
 * adds the other ranks' sums with ONE all-reduce (the reference uses three, AG:289-298). */
int ag_bn_stats_fwd(const float* x, double* sums, int32_t B, int32_t rows, int64_t bstride, int32_t ld, int32_t C,
                    int device, void* stream);

/* var[c] = sums[C+c]/count - (sums[c]/count)^2 (population variance, AG:284);  running_var.lerp_(var, momentum) IN
 * PLACE (AG:340; momentum = 1 - the constructor's momentum argument = 0.1);  scale_out[c] = gamma[c] /
 * sqrt(running_var[c] + eps) (AG:344-359).  If fold_shift_out is non-null it also receives fold_bias[c] * scale + fold_beta[c]
 * (the sandwich post-norm folded behind a Linear's bias, as ag_gemm_fwd's pre_scale / pre_shift use it). */
int ag_bn_update_fwd(const double* sums, double count, float* running_var, const float* gamma, float momentum, float eps,
                     float* scale_out, const float* fold_bias, const float* fold_beta, float* fold_shift_out, int32_t C,
                     int device, void* stream);

/* y = x * scale[c] + shift[b][c] (+ res);  out_f32 = act(y) and / or out_bf16 = act(y).  All tensors [B][rows][C]
 * with their own row / batch strides (elements, multiples of 4).  act: NONE / GELU1702 / GELU_ERF / RELU. */
typedef struct AgAffineArgs {
  const float* x;
  const float* scale;      /* [C] or NULL (= 1) */
  const float* shift;      /* [C] (+ b*shift_bstride) or NULL (= 0) */
  const float* res;        /* NULL = none */
  float* out_f32;          /* NULL = disabled */
  void* out_bf16;          /* NULL = disabled */
  int64_t x_bstride, shift_bstride, res_bstride, out_f32_bstride, out_bf16_bstride;
  int32_t x_ld, res_ld, out_f32_ld, out_bf16_ld;
  int32_t B, rows, C, act;
} AgAffineArgs;
int ag_affine_act_fwd(const AgAffineArgs* args, int device, void* stream);

/* --- prediction heads; their contractions run on ag_gemm_fwd ---------------------------- */

/* ContactMapHead AG:1292-1298 after symmetrize (embeds_pair is already symmetric): RMSNormWithBias over the last dim,
 * + add[row / rows_per_batch][:] (the organism embedding; NULL = none), activation -> bf16 [rows][C] */
int ag_rmsnorm_act_fwd(const float* x, const float* w, const float* b, const float* add, void* out_bf16, int64_t rows,
                       int64_t rows_per_batch, int32_t C, int32_t act, int device, void* stream);

/* SpliceJunctionHead.tissue_scaled_rope AG:1338-1361 / SpliceSitesJunctionHead._apply_rope AG:1478-1499:
 * x fp32 [B][P][hidden] (projected splice-site rows), scale/offset fp32 [T][hidden] ->
 * bf16 [B][T][p_ld][hidden] with out[b][t][p] = rope(x[b][p] * scale[t] + offset[t]) (rows p >= P untouched).
 * Rotary angle tables cos/sin fp32: tab_per_site == 0: [T][hidden/2], position = context index t (AG:1356);
 * tab_per_site == 1: [B][P][hidden/2], position = the site's genomic index (AG:1495). */
int ag_splice_rope_fwd(const float* x, const float* scale, const float* offset, const float* cos_tab, const float* sin_tab,
                       void* out_bf16, int32_t B, int32_t P, int32_t T, int32_t hidden, int32_t p_ld, int32_t tab_per_site,
                       int device, void* stream);

/* library / error plumbing */
int ag_version(void);
const char* ag_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches claim) */
uint64_t ag_launch_count(void);
/* 0 if `device` is an sm_100 part and the kernels can be loaded, else non-zero with ag_last_error() */
int ag_check_device(int device);
/* device watchdog: a kernel whose mbarrier wait exceeds ~4 s records {0xDEAD0000|tag, blockIdx, threadIdx, parity}
 * in host-mapped memory and traps (never hangs the GPU); returns non-zero if such a record exists. */
int ag_watchdog_status(int device, uint32_t out[4]);
/* Process-wide tuning / bring-up state (no reference counterpart).  Everything else in this library is re-entrant and
 * keeps no state between calls; these switches are the exception, are meant for A/B timing, and must not be changed
 * while other threads are inside the library:
 *   - ag_gemm_set_cta_group: default CTAs per tcgen05.mma in ag_gemm_fwd for calls that leave AgGemmArgs.cta_group = 0.
 *     0 = automatic (CTA pairs whenever the problem has at least two 128-row tiles), 1, 2.  Results are bit-identical
 *     either way (same products, same fp32 accumulation order); returns the previous setting.
 *   - environment variables read ONCE per process: AGB_GEMM_CG (seeds the default above), AGB_GEMM_UNIFORM=1 (divisor
 *     N tiling instead of 256-wide tiles + remainder), AGB_GEMM_EW=16 (16- instead of 32-column epilogue chunks),
 *     AGB_ATTN_POLY (exponentials per 8 on the FMA pipe), AGB_ATTN_SPLIT (force the attention key split off / on),
 *     AGB_ROWATTN=0 (mode-1 attention always on the one-tile-per-CTA two-pass kernel), AGB_PAIR_BIAS_MMA=0 (fp32 FMA
 *     kernels for ag_pair_bias_fwd), AGB_PDL=0 (no programmatic dependent launch). */
int ag_gemm_set_cta_group(int cta_group);

#ifdef __cplusplus
}
#endif
#endif /* AGB200_H */
