/*
 * vipant_b200 -- C-ABI of the B200-native InfoNCE / retrieval-scoring hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every
 * entry point replaces a span of the reference's PyTorch code in
 *   /root/reference/cvap/module/decoder/loss_head.py        (cited per function)
 * and is what a binding on the reference side calls (ctypes stub: INTEGRATION.md;
 * the in-tree Python mirror of the reference's loss-head API is vipant_b200/loss_head.py).
 *
 * Conventions
 *   - all data pointers are DEVICE pointers unless the name ends in _host;
 *   - matrices are row-major, leading dimension in ELEMENTS;
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work on it:
 *     no allocation, no host synchronisation, no global state besides the error string;
 *   - scratch memory is caller-owned, sized by the matching *_workspace_bytes query;
 *   - return value: 0 ok, <0 invalid argument (VPA_E_*), >0 a cudaError_t;
 *     vpa_last_error_string() describes the last non-zero return of the calling thread;
 *   - there is NO CPU fallback: on a machine without an sm_100 device every compute
 *     entry point returns an error.
 */
#ifndef VIPANT_B200_H_
#define VIPANT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VPA_VERSION 100 /* 0.1.0 */

/* element types of user-facing feature matrices */
#define VPA_F32 0
#define VPA_BF16 1
#define VPA_F16 2
#define VPA_U8 3 /* label matrices only */

/* arithmetic of the similarity contraction */
#define VPA_PREC_BF16_TC 0   /* bf16 operands, fp32 accumulate, tcgen05 tensor cores (TMA + TMEM)   */
#define VPA_PREC_FP32_SIMT 1 /* fp32 FFMA; "fp32 mode" parity bars, any D % 4 == 0                    */

#define VPA_E_INVALID (-1)     /* bad argument (null pointer, negative size, unsupported D ...) */
#define VPA_E_WORKSPACE (-2)   /* workspace too small */
#define VPA_E_UNSUPPORTED (-3) /* shape/precision combination not supported by this build */
#define VPA_E_NO_DEVICE (-4)   /* no sm_100 device / driver entry point missing */
#define VPA_E_COMM (-5)        /* NCCL could not be loaded / a collective failed */

int vpa_version(void);
const char* vpa_last_error_string(void);

/* Opt-in launch timing for measurement (bench.py): when enabled, the library brackets its dominant kernels
 * with CUDA events on the launch stream.  kind: VPA_PROF_* below.  vpa_profile_read synchronises on the recorded events, returns the summed kernel time and the
 * number of launches since the last read, and resets the counter.  Off by default; not thread-safe. */
#define VPA_PROF_NORMALIZE 0
#define VPA_PROF_FWD_SWEEP 1
#define VPA_PROF_BWD_SWEEP 2
#define VPA_PROF_SIM 3
#define VPA_PROF_RANK 4
#define VPA_PROF_FWD_GENERAL 5 /* the exact two-sweep forward when the single-pass one is also enqueued */
#define VPA_PROF_FINALIZE 6
#define VPA_PROF_PUSH 7 /* peer-memory transport: the stand-alone operand relay kernel (shapes without the fused forward) */
int vpa_profile_enable(int on);
int vpa_profile_hold(int hold); /* pause (1) / resume (0) the bracketing without discarding what was recorded */
/* Kernel launches of this library since it was loaded (every launch site counts itself): bench.py's `gpu_launches`. */
unsigned long long vpa_launch_count(void);
/* Work decomposition chosen for a shape (diagnostics / tests; host only, no device needed): out10 = n_tiles, single-pass
 * forward {chunks, tiles per equal chunk, tiles of the short tail chunk}, backward {same three}, forward row blocks,
 * backward row blocks, impl (1 = CTA-pair kernels).  peer_memory != 0: the plan of the peer-memory transport, whose relay
 * CTAs take SMs from the sweeps. */
int vpa_plan_query(int64_t rows_local, int64_t rows_global, int D, int precision, int peer_memory, int* out10);
int vpa_profile_read(int kind, float* total_ms, int* launches);

/* ------------------------------------------------------------------------------------------
 * L2 normalisation + cast.  Replaces  x / x.norm(dim=-1, keepdim=True)
 *   loss_head.py:271-273 (CELossHead.forward), :38-40 (LossHead.infer); no epsilon, a zero
 *   row yields NaN exactly as the reference does.
 * vpa_normalize_cast: one matrix.  Any of y_bf16 / y_f32 / inv_norm may be NULL.
 *   already_normalized != 0 reproduces the `normalized=True` entry condition (:271): the
 *   values are only cast, inv_norm is written as 1.
 * vpa_normalize_pair: the fused training-path kernel: both modalities in one launch plus
 *   diag_cos[i] = <a_i, t_i> computed from the bf16-rounded (precision 0) or fp32 rows.
 * Requires D % 4 == 0, ld >= D, 16-byte aligned rows.
 * ------------------------------------------------------------------------------------------ */
int vpa_normalize_cast(const void* x, int in_dtype, int64_t rows, int D, int64_t ld,
                       int already_normalized, void* y_bf16, float* y_f32, float* inv_norm,
                       void* stream);

int vpa_normalize_pair(const void* x1, const void* x2, int in_dtype, int64_t rows, int D,
                       int64_t ld1, int64_t ld2, int already_normalized,
                       void* a_bf16, void* t_bf16, float* a_f32, float* t_f32,
                       float* inv_norm1, float* inv_norm2, float* diag_cos, int diag_from_bf16,
                       void* stream);

/* ------------------------------------------------------------------------------------------
 * InfoNCE forward.  Replaces loss_head.py:276-283 for the rows this process owns.
 *   s        = min(exp(*logit_scale), scale_max)   (:276; scale_max <= 0 or +inf: no clamp, :254)
 *   S        = s * A_loc . T_all^T   and   S' = s * T_loc . A_all^T   are formed tile by tile on
 *              the tensor cores and never written to memory;
 *   row_lse[i] = logsumexp_j S[i, j],  col_lse[i] = logsumexp_j S'[i, j],  diag[i] = s * diag_cos[i]
 *   scale_out[0] = s, scale_out[1] = 1 if the clamp passes the gradient (exp(l) <= scale_max) else 0.
 * a_loc/t_loc: this rank's rows_local normalised rows (rows [row_offset, row_offset+rows_local) of
 * the global batch); a_all/t_all: all rows_global rows (== a_loc/t_loc on one GPU).  Element type:
 * bf16 for VPA_PREC_BF16_TC (requires D % 64 == 0, D <= 512, contiguous rows ld == D),
 * fp32 for VPA_PREC_FP32_SIMT.
 * ------------------------------------------------------------------------------------------ */
size_t vpa_infonce_workspace_bytes(int64_t rows_local, int64_t rows_global, int D, int precision);

int vpa_infonce_fwd(const void* a_loc, const void* t_loc, const void* a_all, const void* t_all,
                    int precision, int64_t rows_local, int64_t rows_global, int D,
                    int64_t row_offset, const float* logit_scale, float scale_max,
                    const float* diag_cos, void* workspace, size_t workspace_bytes,
                    float* row_lse, float* col_lse, float* diag, float* scale_out, void* stream);

/* The same forward split around the one exchange step of a row-sharded batch (north star: "a small all-reduce
 * combines the column statistics").  While s * log2(e) <= 62 (s <= 43; every exp2(S*s2 - s2) is a normal fp32 number)
 * the tensor-core path sweeps S ONCE: row sums for the local rows and, for every column, the sum over the LOCAL rows,
 * reduced in fixed order to col_sum[VPA_COLSUM_SPLIT][rows_global].  The caller all-reduces (SUM) col_sum over the
 * ranks and calls _finish, which turns the sums into row_lse / col_lse / diag of the local rows.  For larger s (decided
 * on the device, no host sync) the exact two-sweep kernel runs instead and col_sum is zero / ignored.
 * `parts`: 3 = everything; 1 = only the single-pass kernel, which reads a_loc and t_all (so the all-gather of a_all can
 * still be in flight); 2 = the rest (exact kernel + column-sum reduction), which reads all four operands.  Call 1 then 2.
 * vpa_infonce_fwd == _sweep + _finish with an internal col_sum; when rows_local < rows_global it always takes the
 * exact two-sweep route (no exchange needed). */
#define VPA_COLSUM_SPLIT 8
size_t vpa_infonce_colsum_floats(int64_t rows_global);
int vpa_infonce_fwd_sweep(const void* a_loc, const void* t_loc, const void* a_all, const void* t_all,
                          int precision, int64_t rows_local, int64_t rows_global, int D, int64_t row_offset,
                          const float* logit_scale, float scale_max, void* workspace, size_t workspace_bytes,
                          float* col_sum, int parts, void* stream);
int vpa_infonce_fwd_finish(int precision, int64_t rows_local, int64_t rows_global, int D, int64_t row_offset,
                           const float* logit_scale, float scale_max, const float* diag_cos, void* workspace,
                           size_t workspace_bytes, const float* col_sum, float* row_lse, float* col_lse,
                           float* diag, float* scale_out, void* stream);

/* loss = mean_i(row_lse - diag) + mean_i(col_lse - diag) over the GLOBAL batch (:280-283: two
 * mean-reduced cross entropies, summed).  Deterministic fixed-order reduction, one block. */
int vpa_infonce_loss(const float* row_lse, const float* col_lse, const float* diag,
                     int64_t rows_global, float* loss_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * InfoNCE backward (the autograd of loss_head.py:271-283, SURVEY.md section 8 closed form).
 * Recomputes the logit tiles, forms G = (softmax_rows + softmax_cols - 2I)/B in registers and
 * contracts it with the other modality on the tensor cores:
 *   dx1[i] = J1_i ( g * s * sum_j G[i,j] t_j ),   dx2[i] = J2_i ( g * s * sum_j G[j,i] a_j )
 * for the local rows, J = (I - a a^T)/||x|| the normalisation Jacobian (identity when
 * already_normalized), g = *grad_out (the incoming dL, e.g. the AMP loss scale).
 *   dlogit_scale[0] = g * flows * sum_{i local, j} G[i,j] S[i,j]   (this rank's share; all-reduce it)
 * row_lse_all / col_lse_all: the GLOBAL (rows_global,) statistic vectors from the forward.
 * x1/x2: the original (un-normalised) local rows, dtype in_dtype; dx1/dx2 same dtype and ld.
 * ------------------------------------------------------------------------------------------ */
int vpa_infonce_bwd(const void* a_loc, const void* t_loc, const void* a_all, const void* t_all,
                    int precision, int64_t rows_local, int64_t rows_global, int D,
                    int64_t row_offset, const float* scale /* scale_out of the forward */,
                    const float* row_lse_all, const float* col_lse_all,
                    const float* grad_out, const void* x1, const void* x2, int in_dtype,
                    int64_t ld1, int64_t ld2, const float* inv_norm1, const float* inv_norm2,
                    int already_normalized, void* workspace, size_t workspace_bytes,
                    void* dx1, void* dx2, float* dlogit_scale, void* stream);

/* ------------------------------------------------------------------------------------------
 * Similarity + rank / top-k for the monitors' scoring.  Replaces, per query row,
 *   S = Q @ K.t(); ind = S.argsort(descending=True); torch.where(ind == gt)[1]; ind[:, :k]
 *   loss_head.py:115-117,128-130 (N==M), :139-142,156-158 (1-vs-5), :81-91,95-103
 *   (retrieval_eval), :381-385 (zero-shot argmax).
 * Q (N,D), K (M,D) fp32.  ranks[i,c] = #{m : S[i,m] > S[i,gt[i,c]]} + #{m < gt : S[i,m] == S[i,gt]}
 * (position in a stable descending sort); topk_idx/topk_val: the k best keys per query,
 * descending, lower index first on ties (k <= 32).  gt_idx may be NULL (g = 0); topk_* may be NULL
 * (k = 0).  Similarities are fp32 FFMA dot products; S (N x M fp32) lives in the workspace.
 * ------------------------------------------------------------------------------------------ */
size_t vpa_sim_workspace_bytes(int64_t N, int64_t M);

int vpa_sim_rank_topk(const float* Q, const float* K, int64_t N, int64_t M, int D,
                      int64_t ldq, int64_t ldk, const int32_t* gt_idx, int g, int k,
                      int64_t* topk_idx, float* topk_val, int32_t* ranks,
                      void* workspace, size_t workspace_bytes, void* stream);

/* The same scoring with the similarity tile consumed in registers -- S is never written -- and BOTH directions from one pass:
 * ranks_q[i,c] = position of key gt_q[i,c] in row i of S = Q.K^T (as above), ranks_k[j,c] = position of query gt_k[j,c] in
 * COLUMN j of S, i.e. in row j of the other direction's similarity K.Q^T (fp32 fmaf is commutative in its factors: the same
 * values bit for bit).  One call therefore serves both halves of LossHead.report / retrieval_eval (loss_head.py:115-117 and
 * :128-130; :139-142 and :156-158; :81-91 and :95-103) for 2*N*M*D flops.  top1_*: the best key per query / best query per
 * key (index int64 + value), "larger value, then lower index" -- the k = 1 of `ind[:, :1]` (:182, :381-385); NaN similarities
 * sort FIRST, as torch.argsort(descending=True) places them.  Any output group may be omitted (g = 0 / NULL pointers).
 * Ground-truth indices outside [0, M) / [0, N) yield rank 0 (never counted); g_q, g_k <= 8; k > 1 needs vpa_sim_rank_topk. */
size_t vpa_sim_fused_workspace_bytes(int64_t N, int64_t M, int g_q, int g_k);
int vpa_sim_rank_fused(const float* Q, const float* K, int64_t N, int64_t M, int D, int64_t ldq, int64_t ldk,
                       const int32_t* gt_q, int g_q, const int32_t* gt_k, int g_k, int32_t* ranks_q, int32_t* ranks_k,
                       int64_t* top1_q, float* top1_val_q, int64_t* top1_k, float* top1_val_k, void* workspace,
                       size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Encoder tail fused up to the InfoNCE operand.  Replaces the last ops of the reference's towers
 *     x = ln_post(x[:, 0, :]);  x = x @ proj;  x = x / x.norm(dim=-1, keepdim=True)
 *   /root/reference/cvap/module/val.py:288-290 (ViTPostEncoder), :143-146 (GPTPostEncoder); the normalisation in
 *   cvap/module/encoder/clip_head.py:117-118, audio_head.py:209-210 -- and the cast to the bf16 operand rows the sweeps read.
 * x: (rows, width) CLS rows, row stride ld (the view hidden[:, 0, :] is read in place), dtype in_dtype; ln_gamma / ln_beta:
 * fp32 (width,); proj_t_bf16: the projection TRANSPOSED, (N, width) bf16 row-major.  LayerNorm in fp32 (biased variance, eps),
 * projection on the tensor cores (bf16 operands, fp32 accumulate), one CTA pair per 256 rows holding all N columns in tensor
 * memory so that the row norm and the cast happen in the GEMM's epilogue.
 * Outputs: a_bf16 (rows, N) normalised operand rows; inv_norm (rows,) = 1 / ||y||; optional y_f32 (rows, N) = the
 * un-normalised projected features (what the backward's normalisation Jacobian needs), optional mean / rstd (rows,) of the
 * LayerNorm.  ln_scratch_bf16: (rows, width) bf16 scratch (the LayerNorm output, also needed by the backward of proj).
 * width % 64 == 0, 64 <= width <= 1024; N in {256, 512}.
 * ------------------------------------------------------------------------------------------ */
int vpa_encoder_tail(const void* x, int in_dtype, int64_t rows, int width, int64_t ld, const float* ln_gamma,
                     const float* ln_beta, float eps, const void* proj_t_bf16, int N, void* ln_scratch_bf16, float* mean,
                     float* rstd, void* a_bf16, float* y_f32, float* inv_norm, void* stream);

/* ------------------------------------------------------------------------------------------
 * Multi-label ranking metrics of the AudioSet zero-shot / tagging evaluation.  Replaces the per-class scikit-learn calls of
 *   /root/reference/cvap/module/decoder/loss_more.py:92-123 (BCELossHead.report, reached from BCELossHead.zero_shot :77-84 and
 *   cvap/monitor/audioset_clf.py:377-404): average_precision_score, roc_auc_score and the middle point of
 *   precision_recall_curve for every class, plus the micro-averaged AP over all (sample, class) pairs.
 * S (N, C) fp32 scores, Y (N, C) labels (VPA_F32 or VPA_U8; positive == 1), row-major with leading dimensions.
 * per_class[c] = { AP, ROC-AUC, precision_mid, recall_mid } (double; AP / AUC are NaN where scikit-learn has none);
 * flags[c] bit 0: class without a positive, bit 1: without a negative; support[c] = positives; micro_ap (may be NULL).
 * truncate_pr != 0: scikit-learn 1.0.1's precision_recall_curve (the release the reference pins), which stops the curve
 * at full recall; 0: the un-truncated curve of releases >= 1.1.  N <= 32768 samples per call (one CTA sorts a class in
 * shared memory).  All sums are fixed-order fp64: results are deterministic.
 * ------------------------------------------------------------------------------------------ */
size_t vpa_multilabel_workspace_bytes(int64_t N, int C);
int vpa_multilabel_scores(const float* S, int64_t ld_s, const void* Y, int y_dtype, int64_t ld_y, int64_t N, int C,
                          int truncate_pr, double* per_class, int32_t* flags, int32_t* support, double* micro_ap,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Several InfoNCE pairs over shared modalities in ONE step: the composite heads of the reference run up to five
 * CELossHead pairs per training step on two to five feature matrices (VALCELossHead va / lv / al, loss_head.py:421-495;
 * VACELossHead vp / ap / va / vv / aa, :497-598).  Here every modality is normalised once, the single-pass forwards of all
 * pairs are one launch, the backward sweeps of all pairs are one launch, and the gradient of a modality that takes part in
 * several pairs is summed before its normalisation Jacobian is applied once: 8 launches instead of 7 per pair.
 *   x[m], ld[m]: n_mod <= 5 feature matrices (rows, D) of dtype in_dtype; pair p = (x[pair_x[p]], x[pair_y[p]]), n_pairs <= 5,
 *   with its own temperature logit_scale[p] (device pointers) and clamp scale_max[p] (host floats; <= 0: none).
 *   loss_out[p]: the pair's loss (loss_head.py:280-283).  grad_out[p] (device): dL/d loss_p -- the pair weights of
 *   VACELossHead and the AMP loss scale go here.  dx[m]: gradient of modality m summed over its pairs; dlogit_scale[p].
 * Tensor-core path only (precision VPA_PREC_BF16_TC, D in {256, 512}), rows <= 8192, one GPU; other shapes return
 * VPA_E_UNSUPPORTED and the caller runs the pairs one by one.
 * ------------------------------------------------------------------------------------------ */
#define VPA_MAX_PAIRS 5
size_t vpa_infonce_multi_state_bytes(int64_t rows, int D, int n_mod, int n_pairs, int precision);
int vpa_infonce_multi_fwd(const void* const* x, const int64_t* ld, int in_dtype, int64_t rows, int D, int n_mod,
                          int already_normalized, const int32_t* pair_x, const int32_t* pair_y, int n_pairs,
                          const float* const* logit_scale, const float* scale_max, int precision, void* state,
                          size_t state_bytes, float* loss_out, void* stream);
int vpa_infonce_multi_bwd(const void* const* x, const int64_t* ld, int in_dtype, int64_t rows, int D, int n_mod,
                          int already_normalized, const int32_t* pair_x, const int32_t* pair_y, int n_pairs, int precision,
                          const float* grad_out, void* state, size_t state_bytes, void* const* dx, float* dlogit_scale,
                          void* stream);

/* ------------------------------------------------------------------------------------------
 * Row-sharded training step, orchestrated inside the library: TWO calls per step (forward, backward) instead of a
 * dozen host-side launches and five framework collectives -- at 8 GPUs the step is ~0.6 ms of kernels, so host
 * enqueue time decides the scaling.  One process per GPU; rank r owns rows [r*b, (r+1)*b) of the global batch
 * (B = b * world).  Semantics == loss_head.py:271-283 on the concatenated batch (the reference's `dp` mode).
 *
 * Communicator: NCCL, the copy the host framework ships, bound with dlopen (vpa_comm_load(path or NULL)).  Rank 0
 * obtains a 128-byte unique id (vpa_comm_unique_id), the host broadcasts it (e.g. torch.distributed), every rank
 * calls vpa_comm_init.  world == 1 needs no communicator (comm = NULL): the same two calls are the single-GPU step.
 *
 * forward:  normalise+cast the local rows straight into the gathered operand buffers -> all-gather of the x2 operands,
 *   then of the x1 operands on a side stream (overlapping the single-pass forward, which does not read them) ->
 *   sweeps -> per-rank message [column sums (B) | row_lse (b) | col_lse (b) | diag (b)] -> ONE all-gather ->
 *   statistics of all rows + the global loss (identical on every rank).  Everything the backward needs stays in `state`.
 * backward: recompute sweeps for the local rows, d(global loss)/d(local rows) into dx1 / dx2 (dtype / ld of x1 / x2),
 *   d logit_scale: dls_reduce != 0 -> summed over the ranks (the full derivative, identical on every rank); dls_reduce == 0
 *   -> this rank's partial (the sum over its own rows), for trainers that reduce parameter gradients themselves: under
 *   DistributedDataParallel (which AVERAGES) the feature gradients are per-rank partials too, so `loss * world` together with
 *   the partial d logit_scale reproduces the single-process update exactly.
 * ------------------------------------------------------------------------------------------ */
int vpa_comm_load(const char* libnccl_path);
int vpa_comm_unique_id(void* out128);
int vpa_comm_init(const void* id128, int rank, int world, void** comm_out);
int vpa_comm_destroy(void* comm);

size_t vpa_sharded_state_bytes(int64_t rows_local, int world, int D, int precision);

int vpa_infonce_fwd_sharded(void* comm, const void* x1, const void* x2, int in_dtype, int64_t rows_local, int world,
                            int rank, int D, int64_t ld1, int64_t ld2, int already_normalized,
                            const float* logit_scale, float scale_max, int precision, void* state,
                            size_t state_bytes, float* loss_out, void* stream);

int vpa_infonce_bwd_sharded(void* comm, const void* x1, const void* x2, int in_dtype, int64_t rows_local, int world,
                            int rank, int D, int64_t ld1, int64_t ld2, int already_normalized, int precision,
                            const float* grad_out, void* state, size_t state_bytes, void* dx1, void* dx2,
                            float* dlogit_scale, int dls_reduce, void* stream);

/* ------------------------------------------------------------------------------------------
 * The same row-sharded step over NVLink PEER MEMORY instead of NCCL.  The all-gather of the normalised operands is FUSED INTO
 * the sweep kernels: the first CTAs of a grid are relays that pull the peers' rows with TMA bulk copies through a shared-
 * memory ring (256-row chunks) and raise one arrival flag per chunk; the sweep CTAs of the same grid start on the local block
 * and consume the peers' tiles as their flags flip.  The forward kernel gathers the x2 operands (all it reads), the backward
 * kernel the x1 operands (which only its second problem reads) while its first problem already runs.  The statistics exchange
 * is one kernel (message into every peer, wait for R messages, merge, loss) and the d logit_scale exchange rides in the
 * backward's finalize kernel: six launches per step, no collective library call on the data path.  Loss and d logit_scale are
 * bitwise identical on every rank and to the NCCL transport; the feature gradients agree with it to fp32 rounding (the
 * backward visits its tiles in another order).
 *
 * Setup (once per (rows_local, world, D, precision) and call site): every rank calls vpa_p2p_create -- the library
 * cudaMalloc's its symmetric segment (gathered operands x 2 steps, messages, flags, workspace) and returns a 64-byte CUDA IPC
 * handle -- the host exchanges the handles (e.g. torch.distributed all_gather, any backend), every rank calls
 * vpa_p2p_connect with the world x 64 bytes in rank order.  Ranks may share a device (tests) or own one each (peer access is
 * enabled lazily).  world <= 8 (one NVSwitch node).  vpa_p2p_destroy synchronises the device and frees everything.
 *
 * vpa_infonce_fwd_p2p returns the step number in *epoch_out; pass it to vpa_infonce_bwd_p2p.  The segment keeps the two
 * most recent steps: a backward for an older step returns VPA_E_INVALID (a module that runs several InfoNCE pairs per step
 * uses one segment per pair).  A peer that never arrives trips a device-side timeout (8 s) and the kernel traps -- an error
 * on the stream, not a hang.
 * ------------------------------------------------------------------------------------------ */
int vpa_p2p_create(int64_t rows_local, int world, int rank, int D, int precision, void** p2p_out,
                   void* ipc_handle_out64);
int vpa_p2p_connect(void* p2p, const void* all_ipc_handles /* world x 64 bytes, rank order */);
int vpa_p2p_destroy(void* p2p);

/* Diagnostics / tests (host only): the relay CTAs' work-item map for matrices [m0, 2).  out5 = matrix (0: x2 operands, 1: x1),
 * source rank, chunk index, first row within the source's block, row count.  Items 0 .. (2 - m0) * chunks_per_rank *
 * (world-1) - 1; relay CTA k of n takes items k, k + n, ...  source_major: 0 = chunk k of every peer before chunk k+1 (the
 * forward's order), 1 = peer after peer starting at me+1 (the backward's order). */
int vpa_debug_relay_item(int item, int m0, int source_major, int world, int me, int chunks_per_rank, int64_t rows_local, int* out5);

int vpa_infonce_fwd_p2p(void* p2p, const void* x1, const void* x2, int in_dtype, int64_t rows_local, int world, int rank,
                        int D, int64_t ld1, int64_t ld2, int already_normalized, const float* logit_scale,
                        float scale_max, int precision, float* loss_out, uint32_t* epoch_out, void* stream);

int vpa_infonce_bwd_p2p(void* p2p, uint32_t epoch, const void* x1, const void* x2, int in_dtype, int64_t rows_local,
                        int world, int rank, int D, int64_t ld1, int64_t ld2, int already_normalized, int precision,
                        const float* grad_out, void* dx1, void* dx2, float* dlogit_scale, int dls_reduce, void* stream);

/* ------------------------------------------------------------------------------------------
 * Host-buffer convenience entry (the end-to-end call timed as `e2e` in bench.py): pageable or
 * pinned HOST x1/x2 (fp32, contiguous), copies them to the device, runs normalise -> forward ->
 * loss -> backward on `stream`, copies loss / dlogit_scale (and dx1/dx2 when non-NULL) back and
 * synchronises the stream.  dev_scratch must hold vpa_infonce_host_scratch_bytes() bytes.
 * ------------------------------------------------------------------------------------------ */
size_t vpa_infonce_host_scratch_bytes(int64_t rows, int D, int precision);

int vpa_infonce_step_host(const float* x1_host, const float* x2_host, int64_t rows, int D,
                          float logit_scale, float scale_max, float grad_out, int precision,
                          void* dev_scratch, size_t dev_scratch_bytes,
                          float* loss_host, float* dlogit_scale_host,
                          float* dx1_host, float* dx2_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VIPANT_B200_H_ */
