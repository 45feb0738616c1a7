/*
 * tq100.h -- C ABI of libtq100.so: the B200 (sm_100a) ternary-GPTQ hot path.
 *
 * The reference (shuhan-wang1/SNLP---Tenary-Post-train-Quantization) has no FFI: its hot path
 * is three Python modules calling ATen.  This header is the boundary a binding for that path
 * would target; every entry point names the reference call site it replaces (file:line under
 * /root/reference).  The Python mirror of the reference API (gptq.GPTQ, quantizer.
 * AsymmetricTernaryQuantizer, reorder.select_next_block_ssr, utils.pack_ternary) calls these
 * through ctypes; see INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless its name ends in
 *     _host; the library never allocates persistent memory -- callers own all buffers;
 *   - matrices are row-major, leading dimensions are in ELEMENTS;
 *   - column index arrays are int32;
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it;
 *   - return value: 0 = ok, >0 = cudaError_t, <0 = TQ_E_* domain status;
 *     tq_last_error_string() describes the last failure on the calling thread.
 */
#ifndef TQ100_H
#define TQ100_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TQ100_ABI_VERSION 1

/* domain statuses */
#define TQ_E_BADARG       (-1)   /* shape / alignment / enum outside what the kernels accept */
#define TQ_E_UNSUPPORTED  (-2)   /* valid request the selected kernel path cannot serve        */
#define TQ_E_NOTPD        (-3)   /* reported through *info, see tq_chol_inverse               */
#define TQ_E_WORKSPACE    (-4)   /* workspace too small                                       */

/* activation dtypes accepted by tq_hessian_accum */
#define TQ_F32  0
#define TQ_F16  1
#define TQ_BF16 2

/* Hessian kernel selection */
#define TQ_HESS_AUTO     0       /* tcgen05 for f16/bf16 with m % 8 == 0, FFMA otherwise */
#define TQ_HESS_FFMA     1       /* fp32 CUDA-core SYRK (any dtype, any shape)           */
#define TQ_HESS_TCGEN05  2       /* TMA-fed tcgen05/TMEM SYRK (f16/bf16 only)            */

/* AGA input (quantizer.py:177-248 is fed differently by its two callers) */
#define TQ_AGA_NONE         0    /* quantizer.py:274 with X=None                                   */
#define TQ_AGA_HESSIAN      1    /* gptq.py:147-150: the damped, normalised H[blk,blk] is "X"      */
#define TQ_AGA_ACTIVATIONS  2    /* main.py:177-180: raw X[:,blk]  =>  Gram = (X'X)[blk,blk]       */

/* column order of the sweep */
#define TQ_ORDER_SEQUENTIAL 0    /* gptq.py:136-137 (use_ssr=False)                                */
#define TQ_ORDER_SSR        1    /* gptq.py:130, reorder.py:107-143 (use_ssr=True)                 */
#define TQ_ORDER_STATIC     2    /* caller-supplied permutation (act-order extension)  */

int         tq_abi_version(void);
const char* tq_last_error_string(void);
/* number of kernels this library has launched on the calling process since load (bench.py's
 * gpu_launches claim is read from here) */
long long   tq_launch_count(void);

/* ---- A2  GPTQ.add_batch, gptq.py:59-76 (main.py:128):  H += X' X -------------------------------
 * H   f32 [m, ldh]  accumulated in place.  The tcgen05 path writes only tiles that touch the
 *     upper triangle; call tq_symmetrize before reading H as a full matrix.
 * X   [Nt, ldx] of `dtype`, row = token.  */
int tq_hessian_accum(float* H, int64_t ldh, const void* X, int64_t Nt, int64_t m, int64_t ldx,
                     int dtype, int path, void* stream);
/* development aid: device buffer (>= 32 long long) that receives clock64 stamps of the first panel's phases inside
 * the diagonal-block kernel of tq_chol_inverse (scripts/diag_prof.py); NULL switches it off */
void tq_debug_set_diag_prof(long long* dev_buf);

/* lower triangle := upper triangle */
int tq_symmetrize(float* H, int64_t ldh, int64_t m, void* stream);

/* ---- A3  GPTQ.quantize prologue, gptq.py:94-106 (main.py:129-139) -----------------------------
 * Hd = Hraw / nsamples;  Hd.diag += percdamp * mean(diag Hd).   scratch: >= 1 float. */
int tq_hessian_finalize(float* Hd, const float* Hraw, int64_t m, double nsamples, double percdamp,
                        float* scratch, void* stream);
/* Hinv = potri(potrf(Hd)) as a full symmetric matrix (torch.linalg.cholesky + torch.
 * cholesky_inverse, gptq.py:102-103).  work: tq_chol_workspace_floats(m) floats.
 * *info_dev (device int) = 0, or 1-based index of the first non-positive pivot -- the caller
 * then takes the reference's pinv route (gptq.py:104-106). */
int64_t tq_chol_workspace_floats(int64_t m);
int tq_chol_inverse(float* Hinv, const float* Hd, int64_t m, float* work, int* info_dev, void* stream);

/* ---- A4  select_next_block_ssr, reorder.py:107-143 + :36-61 -----------------------------------
 * Stage 1: row means of the remaining columns and per-row-chunk partial column statistics.
 *   rowmean  f32 [n]
 *   partials f32 [tq_ssr_num_chunks(n), 2, rem]  (dot with the mean vector; squared norm)
 * Stage 2 (one CTA): similarities, top-`block` (descending, ties -> lower position), the new
 *   remaining list (order kept).  Between the stages a row-sharded caller all-reduces
 *   `partials` and the scalar sum of rowmean^2 (pass it in wbar_sq_dev, else NULL).
 *   sims f32 [2*rem + 2]: first rem entries = compute_column_similarity_to_mean's return value, the
 *   rest is scratch (selection keys, ||wbar||^2). */
int64_t tq_ssr_num_chunks(int64_t n);
int tq_ssr_stats(const float* W, int64_t ldw, int64_t n, const int32_t* rem_idx, int64_t rem,
                 float* rowmean, float* partials, void* stream);
int tq_ssr_select(const float* partials, int64_t num_chunks, const float* rowmean, int64_t n,
                  const float* wbar_sq_dev, const int32_t* rem_idx, int64_t rem, int64_t block,
                  int32_t* blk_idx, int32_t* new_rem_idx, float* sims, void* stream);

/* ---- A7 input  s1 = S 1, d = 1' S 1 for the block (quantizer.py:207-218) -----------------------
 * mode TQ_AGA_HESSIAN: S = Hb' Hb with Hb = Hsrc[blk,blk];  TQ_AGA_ACTIVATIONS: S = Hsrc[blk,blk].
 * s1d f32 [b + 1]: s1 then d.  blk_idx NULL => columns col0 .. col0+b-1. */
int tq_aga_vector(const float* Hsrc, int64_t ldh, const int32_t* blk_idx, int64_t col0, int64_t b,
                  int mode, float* s1d, void* stream);

/* ---- A5-A7 + error  AsymmetricTernaryQuantizer.quantize, quantizer.py:250-277; gptq.py:158-159 --
 * One warp per row: init (:32-69), ITF with per-row exit (:136-175), AGA from s1d (:215-246),
 * E = W_b - (alpha T + mu).  blk_idx NULL => contiguous columns from col0.
 *   T      int8 [n, ldt]   codes of this block (position p of the block at T[r*ldt + p])
 *   alpha, mu  f32, element r at alpha[r*ld_am]
 *   E      f32 [n, lde] or NULL;  iters int32 [n] or NULL (per-row ITF iterations)
 *   s1d    NULL => no AGA */
int tq_atq_block(const float* W, int64_t ldw, int64_t n, const int32_t* blk_idx, int64_t col0,
                 int64_t b, const float* s1d, int max_iter, int8_t* T, int64_t ldt,
                 float* alpha, float* mu, int64_t ld_am, float* E, int64_t lde, int32_t* iters,
                 void* stream);

/* The individual stages of the quantizer API (quantizer.py:32-248), one op per call; not on the
 * sweep's path.  W [n, ldw] with the block's b columns contiguous from column 0; T_in/T_out int8
 * [n, b]; alpha/mu f32 [n].
 *   INIT  -> (alpha, mu, T) of ternary_init             GRID  (T_in)            -> (alpha, mu)
 *   ROUND (alpha_in, mu_in) -> T                        ITF   (T_in, alpha_in, mu_in) -> (alpha, mu, T)
 *   AGA   (T_in, s1d)       -> (alpha, mu) */
#define TQ_ATQ_OP_INIT  0
#define TQ_ATQ_OP_GRID  1
#define TQ_ATQ_OP_ROUND 2
#define TQ_ATQ_OP_ITF   3
#define TQ_ATQ_OP_AGA   4
int tq_atq_stage(int op, const float* W, int64_t ldw, int64_t n, int64_t b, const int8_t* T_in,
                 const float* alpha_in, const float* mu_in, const float* s1d, int max_iter,
                 int8_t* T_out, float* alpha_out, float* mu_out, void* stream);

/* ---- A8  error feedback, gptq.py:173-186 (main.py:201-214) ------------------------------------
 * W[:, rem] -= E @ (Hinv[blk, rem] / clamp(diag(Hinv)[blk], 1e-8)[:, None])
 * blk_idx/rem_idx NULL => contiguous ranges starting at blk0 / rem0. */
int tq_err_feedback(float* W, int64_t ldw, int64_t n, const float* E, int64_t lde,
                    const float* Hinv, int64_t ldh, const int32_t* blk_idx, int64_t blk0, int64_t b,
                    const int32_t* rem_idx, int64_t rem0, int64_t rem, void* stream);

/* Tensor-core form of the same update: E C as a split-TF32 tcgen05 GEMM (lo*lo' + lo*hi' + hi*lo' + hi*hi',
 * fp32 accumulate in TMEM), ~2^-22 relative per product.  workspace:
 * tq_err_feedback_tc_workspace_floats(n, b, rem) floats, 16-byte aligned.  tq_split_tf32 writes the
 * (hi, lo) split of a row-major matrix (hi = rn_tf32(x), lo = rn_tf32(x - hi)). */
int64_t tq_err_feedback_tc_workspace_floats(int64_t n, int64_t b, int64_t rem);
int tq_err_feedback_tc(float* W, int64_t ldw, int64_t n, const float* E, int64_t lde,
                       const float* Hinv, int64_t ldh, const int32_t* blk_idx, int64_t blk0, int64_t b,
                       const int32_t* rem_idx, int64_t rem0, int64_t rem, float* workspace, void* stream);
int tq_split_tf32(const float* x, int64_t ld, int64_t rows, int64_t cols, float* hi, float* lo,
                  int64_t ld_out, void* stream);

/* ---- A9-A11  epilogue, dequant, 2-bit codec ---------------------------------------------------
 * Tperm holds block k's codes at columns [k*block, ...) in sweep order; Torig[:, perm[p]] = Tperm[:, p]
 * (gptq.py:155 stores T in ORIGINAL positions).  Tf32 optional float copy (gptq.py:109). */
int tq_unpermute_codes(const int8_t* Tperm, int64_t n, int64_t m, const int32_t* perm,
                       int8_t* Torig, float* Tf32, void* stream);
/* gptq.py:201-230: Wq[:, perm[blk_k]] = alpha_k * T[:, perm[blk_k]] + mu_k (T in original positions) */
int tq_dequant(const float* alpha, const float* mu, int64_t nb, const int8_t* Torig, int64_t n,
               int64_t m, const int32_t* perm, int64_t block, float* Wq, void* stream);
/* utils.py:189-219 / :222-248: code = T+1, 4 codes per byte, flat row-major, zero padded */
int tq_pack2b(const int8_t* T, int64_t count, uint8_t* packed, void* stream);
int tq_pack2b_f32(const float* T, int64_t count, uint8_t* packed, void* stream);
int tq_unpack2b(const uint8_t* packed, int64_t count, int8_t* T, void* stream);

/* ---- whole-layer column sweep, gptq.py:108-199 (main.py:143-223) -------------------------------
 * Runs every block step (select -> AGA vector -> ATQ -> feedback) back to back on `stream` with
 * no host synchronisation.  W is the working copy and is destroyed.
 *   Hd    damped normalised H (AGA 'hessian');  Hraw raw accumulated H (AGA 'activations');
 *   static_perm int32 [m] for TQ_ORDER_STATIC, else NULL;
 *   Torig int8 [n,m], alpha/mu f32 [n, nb], perm int32 [m] (nb = ceil(m/block));
 *   workspace: tq_sweep_workspace_bytes(n, m, block) bytes;
 *   flags: TQ_SWEEP_ROW_SHARD = W holds one contiguous row slab of the layer and the other slabs are
 *          swept by the other ranks of the communicator (tq_comm_init): with SSR the per-block column
 *          statistics are all-reduced (2*rem+1 floats, NCCL, on `stream`) so every rank selects the
 *          same block; sequential / static orders need no exchange. */
#define TQ_SWEEP_ROW_SHARD 1
#define TQ_SWEEP_FFMA_FEEDBACK 2   /* error feedback on the fp32 CUDA cores instead of the split-TF32 tcgen05 GEMM */
#define TQ_SWEEP_UNFUSED_STATS 4   /* SSR statistics by two passes over W per block instead of from the feedback epilogue */
int64_t tq_sweep_workspace_bytes(int64_t n, int64_t m, int64_t block);
int tq_sweep_layer(float* W, int64_t ldw, int64_t n, int64_t m, const float* Hd, const float* Hraw,
                   const float* Hinv, int64_t block, int order, int aga, int max_iter,
                   const int32_t* static_perm, int8_t* Torig, float* alpha, float* mu,
                   int32_t* perm, void* workspace, int64_t workspace_bytes, int flags, void* stream);

/* ---- SURVEY 8f N2  TernaryLinear on 2-bit codes, model.py:17-127 -----------------------------------
 * Layer format: codes u32 [n, wpr] (wpr >= tq_tl_words_per_row(m) = ceil(m/16)): word w of row r holds SWEEP
 * positions 16w..16w+15 as two bit planes -- bit j = 1 iff T[r, perm[16w+j]] = +1, bit 16+j = 1 iff it is -1,
 * positions >= m are 0 in both -- so block k of alpha/mu (gptq.py:153-155) covers positions [k*block, (k+1)*block).
 * wtab f32 [n, nb, 4] = (fl(mu-alpha), mu, fl(alpha+mu), 0): `alpha * T + mu` of model.py:108 evaluated and
 * rounded in the layer dtype `wdtype` (TQ_F32/F16/BF16), for T = -1, 0, +1.  perm NULL = identity.
 * block must be a multiple of 16 (else TQ_E_UNSUPPORTED).
 *   tq_tl_pack     T int8 [n, m] in ORIGINAL positions (gptq.py:155) + perm -> codes
 *   tq_tl_wtab     alpha, mu f32 [n, nb] -> wtab
 *   tq_tl_gemv     y[t, r] = sum_p wtab[r, p/block][T(r,p)] * x[t, perm[p]] (+ bias[r]) for M tokens, i.e.
 *                  model.py:75-95 with the dequantised weight of gptq.py:201-230; x [M, ldx] of xdtype, y f32
 *                  [M, ldy]; meant for decode-sized M (4 tokens per launch; subset sums of x looked up by nibble)
 *   tq_tl_gemm_tc  the same product for MANY tokens as one tcgen05 GEMM that expands the codes to the layer's 16-bit
 *                  dtype in shared memory (no dense weight in HBM): x, y [M, ld] of xdtype (TQ_F16 / TQ_BF16 = the layer
 *                  dtype wtab was built for), fp32 accumulation; needs m % 8 == 0 and 16-byte aligned x; xperm_work
 *                  [M, m] of xdtype is the gather buffer for x[:, perm] (model.py:84), unused when perm is NULL
 *   tq_tl_dequant  dense Wq [n, ldw] of wdtype in ORIGINAL column positions (model.py:97-110 / gptq.py:201-230)
 *   tq_tl_unpack   T int8 [n, m] in original positions
 * tq_tl_dequant / tq_tl_unpack: inv_perm (int32 [m], inv_perm[perm[p]] = p) is optional; with it the kernels gather
 * codes and write coalesced rows, without it a permuted layer is scattered column by column. */
int64_t tq_tl_words_per_row(int64_t m);
int tq_tl_pack(const int8_t* Torig, int64_t n, int64_t m, const int32_t* perm, uint32_t* codes, int64_t wpr,
               void* stream);
int tq_tl_wtab(const float* alpha, const float* mu, int64_t n, int64_t nb, int wdtype, float* wtab, void* stream);
int tq_tl_gemv(const uint32_t* codes, int64_t wpr, const float* wtab, int64_t n, int64_t m, int64_t block,
               const void* x, int xdtype, int64_t ldx, int64_t M, const int32_t* perm, const float* bias,
               float* y, int64_t ldy, void* stream);
int tq_tl_gemm_tc(const uint32_t* codes, int64_t wpr, const float* wtab, int64_t n, int64_t m, int64_t block,
                  const void* x, int xdtype, int64_t ldx, int64_t M, const int32_t* perm, void* xperm_work,
                  const float* bias, void* y, int64_t ldy, void* stream);
int tq_tl_dequant(const uint32_t* codes, int64_t wpr, const float* wtab, int64_t n, int64_t m, int64_t block,
                  const int32_t* perm, const int32_t* inv_perm, void* W, int wdtype, int64_t ldw, void* stream);
int tq_tl_unpack(const uint32_t* codes, int64_t wpr, int64_t n, int64_t m, const int32_t* perm,
                 const int32_t* inv_perm, int8_t* Torig, void* stream);

/* ---- communicator for the row-sharded sweep ------------------------------------------
 * One process per GPU.  Rank 0 calls tq_comm_unique_id (128 bytes, host), ships the bytes to the other
 * ranks (torch.distributed broadcast), every rank calls tq_comm_init on its device.  NCCL is taken from
 * the libnccl.so.2 already loaded in the process; the library itself does not link against it.
 * tq_comm_ready returns the communicator size (0 = none). */
int tq_comm_unique_id(void* out128_host);
int tq_comm_init(const void* id128_host, int rank, int nranks);
int tq_comm_ready(void);
int tq_comm_destroy(void);
int tq_comm_allreduce_f32(float* buf, int64_t count, void* stream);

/* ---- peer-memory mailboxes for the same exchange (one-shot all-reduce over NVLink, csrc/comm.cu) ---------------------
 * The per-block statistics of the row-sharded SSR sweep (2*rem+1 floats, reorder.py:36-61 summed over all rows) are
 * stored by each rank directly into every rank's mailbox and summed locally in rank order; with the mailboxes open
 * tq_sweep_layer uses them instead of NCCL.  Setup, once per process: every rank calls tq_comm_p2p_alloc (cudaMalloc of
 * its own buffer, cap_floats >= 2 * widest layer + 1; returns the 64-byte CUDA IPC handle), the handles are all-gathered
 * by the host (torch.distributed), every rank calls tq_comm_p2p_open with the nranks x 64 bytes in rank order.
 * tq_comm_p2p_ready returns the number of ranks (0 = not open).  All ranks must run the same sequence of sharded sweeps,
 * one at a time per process (the mailboxes and the sequence counter are per process, not per stream). */
int tq_comm_p2p_alloc(int64_t cap_floats, int rank, int nranks, void* handle64_out);
int tq_comm_p2p_open(const void* handles);
int tq_comm_p2p_ready(void);

#ifdef __cplusplus
}
#endif
#endif /* TQ100_H */
