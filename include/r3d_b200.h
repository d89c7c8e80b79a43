/*
 * r3d_b200 -- C ABI of the B200-native R3D token-fuser / effective-rank path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every
 * pointer is a DEVICE pointer unless the name ends in `_host`.  Every call is
 * asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant, and
 * keeps no global mutable state except the thread-local last-error string.
 * Return value: 0 on success, non-zero on error (r3d_last_error() explains).
 *
 * Reference citations are relative to the reference repository root
 * (olivesgatech/R3D): each entry point names the reference lines it replaces.
 *
 * dtype codes: R3D_F32 = 0 (float), R3D_BF16 = 1 (__nv_bfloat16).
 * blend codes: R3D_BLEND_SWAP = 0, R3D_BLEND_SCALE = 1, R3D_BLEND_CONVEX = 2.
 */
#ifndef R3D_B200_H
#define R3D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define R3D_F32 0
#define R3D_BF16 1

#define R3D_BLEND_SWAP 0   /* model/futr_safuser_tokenfusion.py:59-60          */
#define R3D_BLEND_SCALE 1  /* model/futr_safuser_tokenfusion_vary.py:51-56     */
#define R3D_BLEND_CONVEX 2 /* model/futr_safuser_batchnormalization.py:65-74   */

/* ---- library ------------------------------------------------------------ */
const char* r3d_last_error(void);
int r3d_abi_version(void);
/* Number of kernels this library has launched from the calling thread since
 * the counter was last reset (bench.py's `gpu_launches`). */
int64_t r3d_launch_count(int reset);
/* Per-stage timing with CUDA events recorded on the launching stream (off by
 * default).  r3d_profile_read synchronises the recorded events and returns, per
 * stage, accumulated milliseconds, bracketed calls and kernel launches. */
/* Thread-local tuning knobs (unknown keys are an error).  Solver:
 *   "erank_passes"        2 (default): after the first Jacobi pass a second one runs on G2 = Y Y^T, which is graded
 *                         and nearly diagonal, and updates U and Y -- gradients within 1e-4 of float64 on square
 *                         samples; 1 = single pass (erank unchanged, gradients 1e-3 .. 5e-2 on square samples)
 *   "jacobi_tol"          relative rotation threshold |s_pq| > tol sqrt(s_pp s_qq), default 1e-6
 *   "jacobi_tol_pass1"    the same for the first pass of the two-pass solver, default 1e-6
 *   "jacobi_nu_pass1"     first-pass absolute significance floor in units of 2^-23 max|diag|, default 8192 (the first
 *                         pass stops early, the second finishes; the single-pass solver uses 4)
 *   "jacobi_nu_pass2"     second pass: < 0 (default -1e-10) = scale-free test, a rotation is significant when it passes
 *                         the relative test and one of its two diagonal entries exceeds |value| max|diag|; > 0 = absolute
 *                         floor in the units of "jacobi_nu_pass1"
 *   "jacobi_max_sweeps"   sweep cap, default 16, at most 32; "erank_pass1_sweeps" (12) / "erank_pass2_sweeps" (-1 = 6 up to n = 512, 8 beyond)
 *                         are the caps of the two passes
 * Kernel selection (defaults are the fast paths; the alternatives exist for A/B measurements and as fallbacks):
 *   "jacobi_update_tc"    1 = tcgen05 3xTF32 panel update, 0 = SIMT fp32 tile update
 *   "jacobi_inner_regs"   1 = register-resident inner solver for the cross rounds, 0 = shared-memory solver
 *   "gemm_tc"             1 = refinement / backward / fp32-Gram GEMMs on tcgen05 through bf16 planes, 0 = SIMT
 *   "jacobi_overlap_v"    1 = eigenvector update on a library-owned side stream; "jacobi_v_after_g" 0 = it starts
 *                         right after the inner solve (default), 1 = after the G passes of the round
 *   "jacobi_chunks"       2 (default; up to 4): the batch is cut into chunks whose Jacobi iterations run on separate
 *                         library-owned streams, so that one chunk's issue-bound inner solve shares the SMs with another
 *                         chunk's latency-bound panel passes (only while every chunk still fills the GPU); 1 = one sequence.
 *                         Results do not depend on the chunk count.  "jacobi_own_streams" 1 (default) = all chunks on
 *                         library streams of high priority (V streams: low), 0 = chunk 0 on the caller's stream
 *   "panel_sym"           1 (default) = G <- Q^T G Q as one in-place pass over the upper block triangle (jacobi_sym.cu);
 *                         0 = two passes through the scratch matrix H
 *   "panel_merged"        0 (default); 1 = both G passes in one launch with H in an L2-resident ring, sized by
 *                         "panel_group_mb" (8) and "panel_ring" (6) (less DRAM traffic, measured slower)
 *   "jacobi_schedule"     0 = circle-method round robin, one panel update of G and V per round;
 *                         2 = XOR matchings grouped three at a time into super-rounds confined to 128-column groups
 *                         (where the block count is a power of two >= 8, i.e. n in (192,256], (448,512], (960,1024], ...):
 *                         G is updated every round, the three V updates of a super-round run as ONE chained pass over V;
 *                         1 = spread schedule on the same grouping: group-local problems + one K = 128 pass over G and V
 *                         per super-round (measured slower)
 *   "row_chunk_mult"      row chunks per SM of the column-reduction kernels, default 4
 *   "panel_debug", "panel_grid_cap"   test / timing hooks of the panel kernel */
int r3d_set_option(const char* key, double value);
/* Test hook: one tensor-core panel-update round (G <- Q^T G Q via H, V <- V Q) on caller buffers. */
int r3d_debug_panel_round(float* G, float* H, float* V, const float* Qb, int64_t B, int np, int round,
                          int* scratch, void* stream);
/* Test hook: the chained V update  V <- V Q1 Q2 Q3  of three XOR rounds (masks ga, gb, ga ^ gb); Q3 = the three rounds'
 * Q^T buffers back to back, scratch = B * 32 + 3 * B * np/64 ints. */
int r3d_debug_vchain(float* V, float* Q3, int64_t B, int np, int ga, int gb, int* scratch, void* stream);
/* Host-only test hooks of the eigensolver's schedule (no GPU work).  r3d_debug_round_plan: the rounds of one sweep for nb
 * blocks as (a, b) entries -- b == 0: a single XOR round with mask a; else the super-round {a, b, a ^ b} whose three V
 * updates run as one chained pass; returns the entry count (0: this block count uses the circle method).
 * r3d_debug_chain_plan: tile bookkeeping of that pass for group g: out22[0..3] global blocks of the coset,
 * out22[4 + 4k + pos] local block at accumulator position pos in round k, out22[16 + 2k + p] task of pair p in round k. */
int r3d_debug_round_plan(int nb, int32_t* out_pairs, int cap);
int r3d_debug_chain_plan(int ga, int gb, int g, int32_t* out22);
/* Measurement hook: panel tiles (128 rows x 64 output columns: 32 KB read + 32 KB written) the Jacobi panel kernel
 * has actually processed on the current device since the last reset -- out3[0] G passes, out3[1] V passes, out3[2]
 * tiles of the group-local 128 x 128 problems of the spread schedule (an L2-resident working set, not HBM traffic).
 * The launch sequence is fixed and launches after convergence exit at once, so the bytes a timed region really
 * moved are tiles x 65536, not launches x (one full pass).  Synchronises the device. */
int r3d_panel_tiles(uint64_t* out3, int reset);
/* Side-stream sets (8 streams: per batch chunk one high-priority stream and one low-priority V stream; 41 events each) created
 * in this process so far.  They are pooled per device and
 * lent to host threads, so the count stays flat when short-lived threads (nn.DataParallel replicas,
 * main_utkinects.py:129) call the effective-rank entry points. */
int r3d_stream_sets_created(void);
int r3d_profile_enable(int on);
int r3d_profile_num_stages(void);
const char* r3d_profile_stage_name(int stage);
int r3d_profile_read(double* ms_out, int64_t* calls_out, int64_t* launches_out, int reset);

/* ---- a1: channel score ---------------------------------------------------
 * Replaces  x.abs().mean(dim=(0,1))  for both modalities in one launch:
 * model/futr_safuser_tokenfusion.py:49-50, ..._vary.py:41-42.
 * rgb, depth: (rows, C) row-major, rows = B*T.  partial: workspace of
 * r3d_score_workspace_floats(rows, C) floats.  Stage 1 only: per-CTA column
 * sums of |x| in a fixed order; r3d_bottomk() (or r3d_score_finalize) reduces
 * them, again in a fixed order, so results are run-to-run bit-identical. */
size_t r3d_score_workspace_floats(int64_t rows, int64_t C);
int r3d_channel_score_partial(const void* rgb, const void* depth, int64_t rows, int64_t C, int dtype,
                              float* partial, void* stream);
/* sums_out (2, C): column sums of |x| (NOT divided) -- this is the buffer that is
 * all-reduced across ranks in global-score mode; score_out (2, C) = sums / rows. */
int r3d_score_finalize(const float* partial, int64_t rows, int64_t C, float* sums_out, float* score_out,
                       void* stream);
/* The packed statistic of the batch-sharded path in one launch, no host involvement:
 * packed_out (2C + 2) = [sum|rgb| (C) | sum|depth| (C) | sum of er[0..n_er) | rows].  `er` may be NULL (the sum is
 * then 0).  This buffer is what global-score mode all-reduces; r3d_bottomk_scaled consumes it. */
int r3d_score_finalize_packed(const float* partial, int64_t rows, int64_t C, const float* er, int64_t n_er,
                              float* packed_out, void* stream);
/* The same packed statistic from column-sum partials that the producers of rgb / depth emitted as a by-product
 * (r3d_gemm's colsum_partial of the RGB embedding, tokenfusion.py:179-183; r3d_ln_relu_fwd's of the depth projection,
 * :194-197): part_r (parts_r, C), part_d (parts_d, C) -> packed_out (2C + 2).  With it the separate score pass (a1)
 * disappears from the launch list. */
int r3d_score_pack(const float* part_r, int64_t parts_r, const float* part_d, int64_t parts_d, int64_t rows, int64_t C,
                   const float* er, int64_t n_er, float* packed_out, void* stream);

/* ---- a4: bottom-k -----------------------------------------------------------
 * Replaces  torch.topk(score, k, dim=-1, largest=False)[1]  for `nvec` score
 * vectors at once: model/futr_safuser_tokenfusion.py:52-54, ..._vary.py:44-46,
 * ..._batchnormalization.py:58-60.  score: (nvec, C) float; idx_out: (nvec, k)
 * int64, ascending by score, ties -> lower index, NaN last.  C <= 8192. */
int r3d_bottomk(const float* score, int nvec, int64_t C, int64_t k, int64_t* idx_out, void* stream);
/* Same selection on  sums[i] / *denom  (IEEE fp32 division of the column sums by the -- possibly all-reduced -- row
 * count, i.e. exactly the mean of tokenfusion.py:49-50) without a separate elementwise kernel.  score_out: optional
 * (nvec, C) quotient. */
int r3d_bottomk_scaled(const float* sums, int nvec, int64_t C, int64_t k, const float* denom, int64_t* idx_out,
                       float* score_out, void* stream);

/* ---- a3: BatchNorm1d front end of the BN variant ------------------------------
 * Replaces  bn(x.permute(0,2,1)).permute(0,2,1)  batch statistics:
 * model/futr_safuser_batchnormalization.py:45-46.  Two launches: per-CTA Welford
 * partials, then a fixed-order Chan merge.  stats_out (2 modalities, 3, C):
 * [mean, biased var, unbiased var].  */
size_t r3d_bn_workspace_floats(int64_t rows, int64_t C);
int r3d_bn_stats(const void* rgb, const void* depth, int64_t rows, int64_t C, int dtype, float* workspace,
                 float* stats_out, void* stream);

/* ---- a5/a6/a7: exchange + stack ---------------------------------------------
 * Replaces clone x2 + index_put x2 + stack:
 *   swap   model/futr_safuser_tokenfusion.py:56-62
 *   scale  model/futr_safuser_tokenfusion_vary.py:48-57
 *   convex model/futr_safuser_batchnormalization.py:62-75
 * out: (rows, 2, C), same dtype.  idx_r / idx_d: k int64 channel indices each.
 * alpha: (C) float or NULL (swap).  affine: NULL, or (2 modalities, 2, C) float
 * [scale, shift] applied to the inputs first (x*scale + shift) -- the fused
 * BatchNorm normalise of the BN variant. */
int r3d_exchange_fwd(const void* rgb, const void* depth, const int64_t* idx_r, const int64_t* idx_d, int64_t k,
                     const float* alpha, const float* affine, int blend, void* out, int64_t rows, int64_t C,
                     int dtype, void* stream);

/* ---- a8: backward of the exchange ---------------------------------------------
 * Replaces autograd through the lines above.  g: (rows, 2, C).  d_rgb, d_depth:
 * (rows, C) gradients w.r.t. the tensors that entered the blend (for the BN
 * variant: w.r.t. the normalised tensors).  For blend != swap, rgb/depth (and
 * `affine` if the forward used it) must be given and colsum_partial receives
 * per-CTA column sums, reduced by r3d_exchange_bwd_finalize into
 * colsums_out (5, C) = [d_alpha, sum dRhat, sum dRhat*Rhat_n, sum dDhat, sum dDhat*Dhat_n]
 * (the last four only when `bn_norm` != NULL: the BatchNorm backward sums).
 * bn_norm: (2 modalities, 2, C) float [rstd, -mean*rstd], so xn = x*rstd - mean*rstd. */
size_t r3d_exchange_bwd_workspace_floats(int64_t rows, int64_t C);
int r3d_exchange_bwd(const void* g, const void* rgb, const void* depth, const int64_t* idx_r, const int64_t* idx_d,
                     int64_t k, const float* alpha, const float* affine, const float* bn_norm, int blend,
                     void* d_rgb, void* d_depth, float* colsum_partial, int64_t rows, int64_t C, int dtype,
                     void* stream);
int r3d_exchange_bwd_finalize(const float* colsum_partial, int64_t rows, int64_t C, float* colsums_out, void* stream);
/* BatchNorm backward, in place on d_rgb / d_depth (which hold dRhat / dDhat):
 * dx = gamma*rstd * (dy - sum(dy)/N - xn * sum(dy*xn)/N). */
int r3d_bn_bwd_apply(const void* rgb, const void* depth, const float* bn_norm, const float* gamma_r,
                     const float* gamma_d, const float* colsums, void* d_rgb, void* d_depth, int64_t rows, int64_t C,
                     int dtype, void* stream);

/* ---- a12: effective rank -------------------------------------------------------
 * No reference symbol; follows SURVEY.md appendix B.
 * x: (B, T, C).  n = min(T, C), m = max(T, C).
 * Workspace layout is private; query the size, hand in one device buffer. */
size_t r3d_erank_workspace_bytes(int64_t B, int64_t T, int64_t C, int dtype);
/* Forward: erank_out (B) float, sigma_out (B, n) float (solver order, not sorted),
 * U_out (B, n, n) float eigenvectors of the Gram (column j <-> sigma j),
 * Y_out (B, n, m) float = U^T A with A the (n, m) short-side-major view of x.
 * sweeps_out: optional (B) int32 Jacobi sweeps used, both passes of the two-pass solver added up; NEGATIVE when the
 * final pass ran into its sweep cap while still rotating (not converged: sigma / the gradient are less accurate than
 * the bounds DESIGN.md states; raise "erank_pass2_sweeps" / "jacobi_max_sweeps").  The same holds for the sweeps
 * r3d_jacobi_eigh reports.  gram_impl: 0 = tcgen05 tensor-core path, 1 = SIMT fp32 path (kept for A/B accuracy
 * checks). */
int r3d_erank_fwd(const void* x, int64_t B, int64_t T, int64_t C, int dtype, float rtol, int gram_impl,
                  void* workspace, float* erank_out, float* sigma_out, float* U_out, float* Y_out,
                  int32_t* sweeps_out, void* stream);
/* Backward: dx (B, T, C) same dtype as x = d(sum_b g[b]*erank[b]) / dx;
 * accumulate != 0 adds into dx instead of overwriting. */
int r3d_erank_bwd(const float* g, const float* erank, const float* sigma, const float* U, const float* Y,
                  int64_t B, int64_t T, int64_t C, int dtype, float rtol, void* workspace, void* dx,
                  int accumulate, void* stream);
/* Stage-level entry points (benchmarks and parity tests address stages directly). */
int r3d_gram(const void* x, int64_t B, int64_t T, int64_t C, int dtype, int gram_impl, void* workspace,
             float* G_out, void* stream);
int r3d_jacobi_eigh(const float* G, int64_t B, int64_t n, void* workspace, float* lambda_out, float* U_out,
                    int32_t* sweeps_out, int max_sweeps, void* stream);
size_t r3d_jacobi_workspace_bytes(int64_t B, int64_t n);
/* a13: per-token (short-side) informativeness  s_t = sum_j p_j U[t,j]^2. */
int r3d_token_informativeness(const float* sigma, const float* U, int64_t B, int64_t n, float rtol,
                              float* score_out, void* stream);

/* ---- f1 / f2 / f3: linear-layer GEMM on tcgen05 with fused epilogues ---------------------------------------
 * Replaces nn.Linear + the elementwise ops around it in the fuser Block (model/extras/transformerblock.py:79-93,
 * 118-135 as called from model/futr_safuser_tokenfusion.py:86-95) and in the input projections
 * (tokenfusion.py:111,143,179-197), forward and backward:
 *     D (M, N) = epilogue( sum_k A[m, k] * B[n, k] ),   row-major D with pitch N.
 * a_kmajor: A is stored (M, K) [1] or (K, M) [0];  b_kmajor: B is stored (N, K) [1, an nn.Linear weight] or (K, N) [0].
 *     forward            Y  = X W^T      A = X (1),  B = W (1)
 *     input gradient     dX = dY W       A = dY (1), B = W (0)
 *     weight gradient    dW = dY^T X     A = dY (0), B = X (0)       (split along K internally, fixed-order reduce)
 * dtype R3D_BF16: bf16 operands / outputs, fp32 accumulate.  R3D_F32: fp32 tensors, computed from three bf16 planes
 * per operand (six plane products, ~2^-24 relative) -- needs the workspace.  The contiguous dimension of A and B must
 * be a multiple of 8 elements, bases 16-byte aligned.
 * Epilogue (NULL = none; every member optional; applied in this order to the fp32 accumulator x):
 *     x += bias[n];  aux_out[m, n] = x;  x = act(x);  x *= gelu'(aux_in[m, n]);  x += residual[m, n];  D[m, n] = x;
 *     colsum_partial[m / 128][n] = sum over the 128-row tile of D[m, n] (|D[m, n]| with colsum_abs), rounded values:
 *     ceil(M / 128) partial rows, finalised by r3d_colsum_finalize -- the bias gradient of the NEXT layer down, or the
 *     channel-score sums of tokenfusion.py:49-50 for free.
 * bias (N), residual / aux_* (M, N) have the output dtype.  Weight-gradient calls (split-K) take no epilogue. */
typedef struct r3d_epilogue {
  const void* bias;
  const void* residual;
  void* aux_out;
  const void* aux_in;
  float* colsum_partial;
  int act;         /* 0 none, 1 GELU (erf form, nn.GELU default), 2 ReLU */
  int colsum_abs;
} r3d_epilogue;
size_t r3d_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int dtype);
int r3d_gemm(const void* A, const void* B, void* D, int64_t M, int64_t N, int64_t K, int a_kmajor, int b_kmajor,
             int dtype, const r3d_epilogue* epi, void* workspace, void* stream);
/* out (N) of dtype `dtype` = sum over `parts` partial rows (fixed order). */
int r3d_colsum_finalize(const float* partial, int64_t parts, int64_t N, int dtype, void* out, void* stream);
/* out (C) = column sums of x (rows, C) -- the bias gradient of a Linear whose output gradient is x; two fixed-order
 * stages.  workspace: r3d_colsum_workspace_floats(rows, C) floats. */
size_t r3d_colsum_workspace_floats(int64_t rows, int64_t C);
int r3d_colsum(const void* x, int64_t rows, int64_t C, int dtype, float* workspace, void* out, void* stream);
/* ReLU backward of the input projections (F.relu at tokenfusion.py:183,197): dpre (rows, C) = y > 0 ? dy : 0 and
 * dbias (C) = column sums of dpre in one pass (same workspace size as r3d_colsum). */
int r3d_relu_bwd(const void* dy, const void* y, int64_t rows, int64_t C, int dtype, float* workspace, void* dpre,
                 void* dbias, void* stream);

/* ---- N2: M-modality fuser (BASELINE.json configs[4]: RGB + depth + gaze) --------------------------------------
 * The reference reads M = len(modal_feats) (model/futr_safuser_tokenfusion.py:76) but hard-codes two modalities
 * (:77,:79).  Semantics fixed here -- for M = 2 exactly the reference, beyond that PARITY UNPINNED (oracle:
 * oracle/torch_port.py:PortCMFuserM): modality m swaps its k lowest-score channels for those of modality (m+1) mod M;
 * the Block attends over the M modality tokens with the -inf diagonal mask of tokenfusion.py:68-72 generalised to
 * M x M; the final LayerNorm is followed by the mean over the M tokens (:93-95).
 * r3d_exchange_one_fwd: one stream of the stack -- out (row pitch out_pitch, already offset to slot m) =
 *   c in idx ? other[row, c] : own[row, c].
 * r3d_exchange_one_bwd: dx of modality j = g_j outside idx_j + g_prev inside idx_prev (prev = (j-1) mod M); g_* point
 *   at their slots of the stacked gradient, row pitch g_pitch.
 * r3d_mtoken_attn_fwd / _bwd: qkv (rows, M, 3C) as nn.Linear(C, 3C) lays it out (transformerblock.py:22-23) ->
 *   out (rows, M, C), softmax over the OTHER tokens per head, scale head_dim^-0.5; 2 <= M <= 4, head_dim a multiple
 *   of 32 and <= 256.  The backward recomputes the weights.
 * r3d_token_mean: (rows, M, C) -> (rows, C) mean over the tokens; backward != 0: (rows, C) -> (rows, M, C) / M. */
int r3d_exchange_one_fwd(const void* own, const void* other, const int64_t* idx, int64_t k, void* out,
                         int64_t out_pitch, int64_t rows, int64_t C, int dtype, void* stream);
int r3d_exchange_one_bwd(const void* g_own, const void* g_prev, int64_t g_pitch, const int64_t* idx_own,
                         const int64_t* idx_prev, int64_t k, void* dx, int64_t rows, int64_t C, int dtype, void* stream);
int r3d_mtoken_attn_fwd(const void* qkv, void* out, int64_t rows, int M, int64_t C, int heads, int dtype, void* stream);
int r3d_mtoken_attn_bwd(const void* qkv, const void* dout, void* dqkv, int64_t rows, int M, int64_t C, int heads,
                        int dtype, void* stream);
int r3d_token_mean(const void* x, void* out, int64_t rows, int M, int64_t C, int dtype, int backward, void* stream);

/* ---- N1: token-axis selection (north_star kernels 3-6) -----------------------------------------------------
 * No reference symbol: the reference ships only the channel exchange (model/futr_safuser_tokenfusion.py:33-66,
 * SURVEY.md F2) and describes the token form in prose (README.md:13).  PARITY UNPINNED; oracle:
 * oracle/fuser_oracle.py:token_fusion_tokens.
 * r3d_token_scores: score_out (B, T) = sum_j p_j u_{tj}^2 over the LEFT singular vectors of every (T, C) sample, from
 * the sigma / U / Y that r3d_erank_fwd saved (U when T < C, the rows of Y when T >= C), kept singular values only.
 * r3d_token_mask: per-sample index lists idx_r, idx_d (B, k) int64 (from r3d_bottomk on the scores) -> mask_out (B, T)
 * bytes, bit 0 = token in S_rgb(b), bit 1 = token in S_depth(b).  T <= 8192.
 * r3d_token_exchange_fwd: out (rows, 2, C): out[r, 0] = bit0 ? depth[r] : rgb[r], out[r, 1] = bit1 ? rgb[r] : depth[r]
 * (rows = B*T; the token analogue of tokenfusion.py:56-62).  2 N s read + 2 N s written.
 * r3d_token_exchange_bwd: g (rows, 2, C) -> d_rgb, d_depth (rows, C): masked select per token; the index sets hold no
 * duplicates, so no atomics / scatter-add are needed. */
int r3d_token_scores(const float* sigma, const float* U, const float* Y, int64_t B, int64_t T, int64_t C, float rtol,
                     float* score_out, void* stream);
int r3d_token_mask(const int64_t* idx_r, const int64_t* idx_d, int64_t B, int64_t T, int64_t k, uint8_t* mask_out,
                   void* stream);
int r3d_token_exchange_fwd(const void* rgb, const void* depth, const uint8_t* mask, void* out, int64_t rows, int64_t C,
                           int dtype, void* stream);
int r3d_token_exchange_bwd(const void* g, const uint8_t* mask, void* d_rgb, void* d_depth, int64_t rows, int64_t C,
                           int dtype, void* stream);

/* ---- f1 (next row, first step): row LayerNorm of the fuser Block ------------------
 * Replaces the three nn.LayerNorm of a fuser call and their autograd backward:
 *   Block.norm1 / Block.norm2  model/extras/transformerblock.py:122,127 (used :132,:134)
 *   CMFuser.norm               model/futr_safuser_tokenfusion.py:25 (used :93), same lines in the
 *                              _vary / _batchnormalization / futr_safuser_depth variants.
 * x, y, dy, dx: (rows, C) row-major in `dtype`; gamma, beta in the same dtype (nn.LayerNorm keeps its
 * parameters in the module dtype); mean, rstd: (rows) fp32 saved by the forward for the backward.
 * C must be a multiple of 8 (bf16) / 4 (fp32) and at most 2048 (bf16) / 1024 (fp32); tensors 16-byte aligned.
 * r3d_ln_bwd writes dx and dgamma_dbeta = (2, C) fp32 [dgamma | dbeta] (deterministic two-stage reduction).
 * pair_mean = 1 fuses the mean over the fuser's two modality tokens that follows the final norm
 * (tokenfusion.py:93-95  x = self.norm(x); x = x.mean(dim=1)): rows come in pairs, y and dy have rows/2 rows,
 * y[r] = (LN(x[2r]) + LN(x[2r+1])) / 2. */
size_t r3d_ln_bwd_workspace_floats(int64_t rows, int64_t C);
int r3d_ln_fwd(const void* x, const void* gamma, const void* beta, int64_t rows, int64_t C, int dtype, float eps,
               int pair_mean, void* y, float* mean, float* rstd, void* stream);
int r3d_ln_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const void* gamma, int64_t rows,
               int64_t C, int dtype, int pair_mean, void* dx, float* workspace, float* dgamma_dbeta, void* stream);
/* The same two kernels with the Block's row swap and residual gradient fused in.  flags: bit 0 = pair_mean, bit 1 =
 * swap the two rows of every pair (forward: row r is written to row r ^ 1; backward: dy is read from row r ^ 1) --
 * the closed-form 2-token attention hands token m the V of token 1 - m, and swapping inside norm1 keeps
 * every GEMM of the Block plain.  addend (backward, may be NULL): dx += addend, the gradient arriving over the
 * residual connection. */
int r3d_ln_fwd2(const void* x, const void* gamma, const void* beta, int64_t rows, int64_t C, int dtype, float eps,
                int flags, void* y, float* mean, float* rstd, void* stream);
int r3d_ln_bwd2(const void* dy, const void* x, const float* mean, const float* rstd, const void* gamma, int64_t rows,
                int64_t C, int dtype, int flags, const void* addend, void* dx, float* workspace, float* dgamma_dbeta,
                void* stream);
/* f2 tail -- depth projection (model/futr_safuser_tokenfusion.py:195-197): y = relu(LayerNorm(x)) in one pass, with the
 * channel-score partial sums of |y| (tokenfusion.py:49-50) as a by-product: colsum_partial gets r3d_ln_relu_parts(rows)
 * rows of C floats, finalised by r3d_colsum_finalize.  The backward masks dy by the ReLU (recomputed, y is not needed)
 * and applies the LayerNorm backward; dgamma_dbeta (2, C) float. */
int64_t r3d_ln_relu_parts(int64_t rows);
int r3d_ln_relu_fwd(const void* x, const void* gamma, const void* beta, int64_t rows, int64_t C, int dtype, float eps,
                    void* y, float* mean, float* rstd, float* colsum_partial, void* stream);
int r3d_ln_relu_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const void* gamma,
                    const void* beta, int64_t rows, int64_t C, int dtype, void* dx, float* workspace,
                    float* dgamma_dbeta, void* stream);
/* Residual add of the closed-form 2-token attention: out[row] = a[row] + b[row ^ 1] (token m receives the projected
 * V of token 1-m: transformerblock.py:19-36 with the -inf diagonal mask of tokenfusion.py:68-72), which
 * replaces  x + proj(v.flip(1)).  a == NULL gives the pair-swapped copy (its backward).  rows even. */
int r3d_swap_add(const void* a, const void* b, int64_t rows, int64_t C, int dtype, void* out, void* stream);

/* ---- host-buffer convenience (what a non-Python caller binds; used for `e2e`) ----
 * All pointers are HOST pointers (pinned for full speed); the call copies in,
 * runs score -> bottom-k -> exchange on the given stream, copies the stacked
 * result and the indices out, and synchronises the stream. */
int r3d_token_fusion_host(const void* rgb_host, const void* depth_host, int64_t B, int64_t T, int64_t C,
                          int dtype, int64_t k, void* out_host, int64_t* idx_r_host, int64_t* idx_d_host,
                          void* stream);
/* The whole hot path from HOST buffers (what a non-Python caller binds; `e2e_c_abi` in bench.py):
 *   H2D -> erank(rgb), erank(depth) -> channel score -> bottom-k -> exchange/stack -> [exchange backward of the
 *   upstream gradient gst + d(erank_weight * mean erank)/dX accumulated] -> D2H.
 * rgb_host, depth_host (B, T, C); gst_host (B, T, 2, C) or NULL (forward only; then d_*_host must be NULL too).
 * Outputs: stacked_out_host (B, T, 2, C); erank_out_host (2B) float (rgb samples, then depth); d_rgb_host, d_depth_host
 * (B, T, C); idx_host (2, k) int64 or NULL.  One stream-ordered device arena per call; returns after the stream has
 * been synchronised. */
int r3d_fuser_step_host(const void* rgb_host, const void* depth_host, const void* gst_host, int64_t B, int64_t T,
                        int64_t C, int dtype, int64_t k, float rtol, float erank_weight, void* stacked_out_host,
                        float* erank_out_host, void* d_rgb_host, void* d_depth_host, int64_t* idx_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* R3D_B200_H */
