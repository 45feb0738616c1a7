// The whole-layer block column sweep: gptq.py:108-199 (main.py:143-223) as one host-side loop that
// enqueues every block step on the caller's stream with no host synchronisation -- the reference's
// per-block syncs (`.tolist()` gptq.py:133, `torch.equal` quantizer.py:164, `.item()` quantizer.py:218)
// are gone: the permutation, the remaining-column list and the AGA vector all stay on the device.
#include "tc_common.cuh"

namespace tq {

int launch_atq_block(const float*, int64_t, int64_t, const int32_t*, int64_t, int64_t, const float*, int, int8_t*,
                     int64_t, float*, float*, int64_t, float*, float*, int64_t, int32_t*, const float*, int64_t, const double*,
                     int64_t, float*, cudaStream_t);
int launch_aga_vector(const float*, int64_t, const int32_t*, int64_t, int64_t, int, float*, const float*, int64_t, double*,
                      cudaStream_t);
int launch_ssr_stats(const float*, int64_t, int64_t, const int32_t*, int64_t, float*, float*, float*, cudaStream_t);
int launch_ssr_select(const float*, int64_t, const float*, int64_t, const float*, const int32_t*, int64_t, int64_t,
                      int32_t*, int32_t*, float*, uint32_t*, cudaStream_t);
int launch_err_feedback(float*, int64_t, int64_t, const float*, int64_t, const float*, int64_t, const int32_t*,
                        int64_t, int64_t, const int32_t*, int64_t, int64_t, cudaStream_t);
int launch_unpermute(const int8_t*, int64_t, int64_t, const int32_t*, int8_t*, float*, int32_t*, cudaStream_t);
int launch_ssr_fold(const float*, int64_t, const float*, int64_t, int64_t, float*, cudaStream_t);
bool comm_active();
int comm_allreduce_sum_f32(float*, int64_t, cudaStream_t);
bool p2p_active();
int p2p_fold_allreduce(const float*, int64_t, const float*, int64_t, int64_t, float*, cudaStream_t);

__global__ void iota_kernel(int32_t* __restrict__ a, int m) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) a[i] = i;
}

struct SweepWs {
    float* E;
    float* E_lo;       // low halves of the tf32 split of E (tensor-core feedback)
    float* coef;       // [2][m][ldb] hi/lo coefficient operand of the feedback GEMM
    float* rowmean;
    float* partials;
    float* sims;       // [2*m + 2]: similarities, selection keys, ||wbar||^2
    float* s1d;
    float* folded;     // [2*m + 1] all-reduced SSR statistics (row-sharded sweep)
    float* rowsum_part[2]; // [2*ceil(m/128)][n] row-sum partials over the remaining columns (read by ATQ / written by the
                           // feedback epilogue)
    double* csum;      // [block] row sums of the coefficient matrix C 1
    float* csum_part;  // [ceil(m/32)][block] per-CTA partials of C 1
    int32_t* rem[2];
    int8_t* Tperm;
    int64_t bytes;
};

static inline int64_t align256(int64_t x) { return (x + 255) & ~(int64_t)255; }

static SweepWs carve(void* base, int64_t n, int64_t m, int64_t block) {
    SweepWs w;
    char* p = static_cast<char*>(base);
    int64_t off = 0;
    auto take = [&](int64_t bytes) { char* q = p ? p + off : nullptr; off += align256(bytes); return q; };
    const int64_t chunks = tq_ssr_num_chunks(n);
    const int64_t ldb = (block + 3) & ~(int64_t)3;
    w.E = reinterpret_cast<float*>(take(sizeof(float) * n * ldb));
    w.E_lo = reinterpret_cast<float*>(take(sizeof(float) * n * ldb));
    w.coef = reinterpret_cast<float*>(take(sizeof(float) * 2 * m * ldb));
    w.rowmean = reinterpret_cast<float*>(take(sizeof(float) * n));
    w.partials = reinterpret_cast<float*>(take(sizeof(float) * chunks * 2 * m));
    w.sims = reinterpret_cast<float*>(take(sizeof(float) * (2 * m + 2)));
    w.s1d = reinterpret_cast<float*>(take(sizeof(float) * (block + 1)));
    w.folded = reinterpret_cast<float*>(take(sizeof(float) * (2 * m + 1)));
    w.rowsum_part[0] = reinterpret_cast<float*>(take(sizeof(float) * 2 * ((m + 127) / 128) * n));
    w.rowsum_part[1] = reinterpret_cast<float*>(take(sizeof(float) * 2 * ((m + 127) / 128) * n));
    w.csum = reinterpret_cast<double*>(take(sizeof(double) * block));
    w.csum_part = reinterpret_cast<float*>(take(sizeof(float) * ((m + 31) / 32) * block));
    w.rem[0] = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * m));
    w.rem[1] = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * m));
    w.Tperm = reinterpret_cast<int8_t*>(take(n * m));
    w.bytes = off;
    return w;
}

}  // namespace tq

extern "C" int64_t tq_sweep_workspace_bytes(int64_t n, int64_t m, int64_t block) {
    return tq::carve(nullptr, n, m, block).bytes;
}

extern "C" int tq_sweep_layer(float* W, int64_t ldw, int64_t n, int64_t m, const float* Hd, const float* Hraw,
                              const float* Hinv, int64_t block, int order, int aga, int max_iter,
                              const int32_t* static_perm, int8_t* Torig, float* alpha, float* mu, int32_t* perm,
                              void* workspace, int64_t workspace_bytes, int flags, void* stream) {
    using namespace tq;
    TQ_CHECK_ARG(W && Hinv && Torig && alpha && mu && perm && workspace, "tq_sweep_layer: null pointer");
    TQ_CHECK_ARG(n > 0 && m > 0 && ldw >= m && block >= 1 && block <= 512, "tq_sweep_layer: bad shape");
    TQ_CHECK_ARG(order == TQ_ORDER_SEQUENTIAL || order == TQ_ORDER_SSR || order == TQ_ORDER_STATIC,
                 "tq_sweep_layer: unknown order %d", order);
    TQ_CHECK_ARG(order != TQ_ORDER_STATIC || static_perm != nullptr, "tq_sweep_layer: TQ_ORDER_STATIC needs static_perm");
    TQ_CHECK_ARG(aga == TQ_AGA_NONE || (aga == TQ_AGA_HESSIAN && Hd) || (aga == TQ_AGA_ACTIVATIONS && Hraw),
                 "tq_sweep_layer: AGA mode %d needs its Hessian (Hd for HESSIAN, Hraw for ACTIVATIONS)", aga);
    TQ_CHECK_ARG(max_iter >= 0, "tq_sweep_layer: max_iter < 0");
    const bool sharded = (flags & TQ_SWEEP_ROW_SHARD) != 0;
    TQ_CHECK_ARG(!sharded || order != TQ_ORDER_SSR || comm_active() || p2p_active(),
                 "tq_sweep_layer: TQ_SWEEP_ROW_SHARD with SSR needs the communicator (tq_comm_init) or peer mailboxes");
    cudaStream_t st = (cudaStream_t)stream;
    SweepWs ws = carve(workspace, n, m, block);
    if (ws.bytes > workspace_bytes) {
        set_error("tq_sweep_layer: workspace %lld bytes < required %lld", (long long)workspace_bytes, (long long)ws.bytes);
        return TQ_E_WORKSPACE;
    }
    const int64_t nb = ceil_div(m, block);
    const int64_t chunks = tq_ssr_num_chunks(n);
    const int64_t ldb = (block + 3) & ~(int64_t)3;
    const bool tc_feedback = (flags & TQ_SWEEP_FFMA_FEEDBACK) == 0;
    const float* Haga = (aga == TQ_AGA_HESSIAN) ? Hd : Hraw;
    int rc;

    if (order == TQ_ORDER_SSR) {
        iota_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, st>>>(ws.rem[0], (int)m);
        TQ_LAUNCH_CHECK("iota_kernel");
    } else if (order == TQ_ORDER_SEQUENTIAL) {
        iota_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, st>>>(perm, (int)m);
        TQ_LAUNCH_CHECK("iota_kernel");
    } else {
        TQ_CUDA(cudaMemcpyAsync(perm, static_perm, sizeof(int32_t) * m, cudaMemcpyDeviceToDevice, st));
    }

    // TMA descriptors of the feedback GEMM's operands: E (hi, lo) [n, b] and the coefficient operand (hi, lo)
    // [remaining, b].  Encoded once for the full-width blocks (and once more if the last block is ragged); the
    // coefficient operand is declared with m rows -- rows past the current `rem` only reach masked outputs.
    float* coef_hi = ws.coef;
    float* coef_lo = ws.coef + m * ldb;
    GemmOperands ops_full, ops_tail;
    GemmC c_W;                  // W as the output matrix of the TMA epilogue (contiguous remaining columns only)
    c_W.valid = false;
    if (tc_feedback && order == TQ_ORDER_SEQUENTIAL && (rc = gemm_cmap_encode(&c_W, W, n, m, ldw))) return rc;
    const int64_t tail = m % block;
    if (tc_feedback) {
        if (m > block && (rc = gemm_operands_encode(&ops_full, ws.E, ws.E_lo, ldb, n, coef_hi, coef_lo, ldb, m, block))) return rc;
        if (tail != 0 && m > tail && (rc = gemm_operands_encode(&ops_tail, ws.E, ws.E_lo, ldb, n, coef_hi, coef_lo, ldb, m, tail)))
            return rc;
    }

    // SSR statistics: computed by two passes over W[:, rem] for the first block; afterwards (tensor-core feedback) the
    // feedback GEMM's epilogue emits them for the columns it has just updated, from row means predicted exactly by the
    // ATQ kernel (TQ_SWEEP_UNFUSED_STATS switches back to the two passes per block).
    const bool fused_stats = tc_feedback && (flags & TQ_SWEEP_UNFUSED_STATS) == 0;
    bool have_stats = false;              // ws.partials / ws.rowmean / ws.rowsum_part[rs] describe the current remaining set
    int rs = 0;
    int64_t rs_parts = 1;                 // number of row-sum partials in ws.rowsum_part[rs]

    int cur = 0;
    int64_t done = 0, rem = m;
    for (int64_t k = 0; done < m; ++k) {
        const int64_t b = (m - done < block) ? (m - done) : block;
        const int32_t* blk_idx;
        const int32_t* rem_idx;
        if (order == TQ_ORDER_SSR) {
            if (rem <= block) {                                   // reorder.py:125-126
                TQ_CUDA(cudaMemcpyAsync(perm + done, ws.rem[cur], sizeof(int32_t) * rem, cudaMemcpyDeviceToDevice, st));
            } else {
                if (!have_stats) {
                    if ((rc = launch_ssr_stats(W, ldw, n, ws.rem[cur], rem, ws.rowmean, ws.partials, ws.rowsum_part[rs], st))) return rc;
                    rs_parts = 1;
                }
                if (sharded) {
                    // rows are one shard of the layer: column statistics must cover every shard's rows
                    if (p2p_active()) {
                        // fold + exchange through peer memory (one-shot all-reduce, comm.cu)
                        if ((rc = p2p_fold_allreduce(ws.partials, chunks, ws.rowmean, n, rem, ws.folded, st))) return rc;
                    } else {
                        if ((rc = launch_ssr_fold(ws.partials, chunks, ws.rowmean, n, rem, ws.folded, st))) return rc;
                        if ((rc = comm_allreduce_sum_f32(ws.folded, 2 * rem + 1, st))) return rc;
                    }
                    if ((rc = launch_ssr_select(ws.folded, 1, nullptr, n, ws.folded + 2 * rem, ws.rem[cur], rem, block,
                                                perm + done, ws.rem[cur ^ 1], ws.sims,
                                                reinterpret_cast<uint32_t*>(ws.sims + m), st)))
                        return rc;
                } else if ((rc = launch_ssr_select(ws.partials, chunks, ws.rowmean, n, nullptr, ws.rem[cur], rem, block,
                                                   perm + done, ws.rem[cur ^ 1], ws.sims,
                                                   reinterpret_cast<uint32_t*>(ws.sims + m), st)))
                    return rc;
                cur ^= 1;
            }
            rem -= b;
            blk_idx = perm + done;
            rem_idx = ws.rem[cur];
        } else {
            rem = m - done - b;
            const bool contiguous = (order == TQ_ORDER_SEQUENTIAL);
            blk_idx = contiguous ? nullptr : perm + done;
            rem_idx = contiguous ? nullptr : perm + done + b;
        }
        // will the NEXT block be chosen by similarity?  then this step must leave statistics of the updated W behind
        const bool emit_stats = fused_stats && order == TQ_ORDER_SSR && rem > block;
        if (tc_feedback && rem > 0) {
            if ((rc = launch_feedback_coef(Hinv, m, blk_idx, done, b, rem_idx, done + b, rem, coef_hi, coef_lo, ldb,
                                           emit_stats ? ws.csum_part : nullptr, st)))
                return rc;
        }
        const float* s1d = (aga != TQ_AGA_NONE) ? ws.s1d : nullptr;
        if ((rc = launch_aga_vector(Haga, m, blk_idx, done, b, aga, ws.s1d, emit_stats ? ws.csum_part : nullptr,
                                    ceil_div(rem, 32), emit_stats ? ws.csum : nullptr, st)))
            return rc;
        if ((rc = launch_atq_block(W, ldw, n, blk_idx, done, b, s1d, max_iter, ws.Tperm + done, m, alpha + k, mu + k,
                                   nb, ws.E, tc_feedback ? ws.E_lo : nullptr, ldb, nullptr,
                                   emit_stats ? ws.rowsum_part[rs] : nullptr, rs_parts, emit_stats ? ws.csum : nullptr, rem,
                                   emit_stats ? ws.rowmean : nullptr, st)))
            return rc;
        have_stats = false;
        if (rem > 0) {                                            // gptq.py:170 (and SURVEY Q3)
            if (tc_feedback) {
                const GemmOperands* ops = (b == block) ? &ops_full : &ops_tail;
                if (emit_stats) {
                    rc = launch_gemm_feedback_stats(W, ldw, n, rem, ops, rem_idx, done + b, ws.rowmean, ws.partials,
                                                    ws.rowsum_part[rs ^ 1], st);
                    rs ^= 1;
                    rs_parts = 2 * ceil_div(rem, 128);
                    have_stats = true;
                } else if (c_W.valid && rem_idx == nullptr) {
                    rc = launch_gemm_tf32x3_at(GX_FEEDBACK, &c_W, 0, done + b, n, rem, ops, 0, 0, st);
                } else {
                    rc = launch_gemm_tf32x3_ops(GX_FEEDBACK, W, ldw, n, rem, ops, rem_idx, done + b, st);
                }
            } else {
                rc = launch_err_feedback(W, ldw, n, ws.E, ldb, Hinv, m, blk_idx, done, b, rem_idx, done + b, rem, st);
            }
            if (rc) return rc;
        }
        done += b;
    }
    // (the remaining-column lists are dead by now: ws.rem[0] serves as the inverse-permutation scratch)
    return launch_unpermute(ws.Tperm, n, m, perm, Torig, nullptr, ws.rem[0], st);
}
