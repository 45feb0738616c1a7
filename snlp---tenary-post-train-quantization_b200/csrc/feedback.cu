// A8: lazy-batch error feedback  W[:, rem] -= E @ (Hinv[blk, rem] / clamp(diag(Hinv)[blk], 1e-8)).
// Reference: gptq.py:173-186 (main.py:201-214).  fp32 CUDA-core tile; the coefficient matrix is formed
// on the fly while its slab is staged (gather of Hinv[blk_i, rem_j], one rounded division by the
// clamped diagonal -- the reference's `coefficients = H_inv_block_rem / H_inv_diag`), so it never
// exists in HBM.  Bound: K = block (128) gives 32 flop per byte of W read-modify-write, i.e. HBM-bound
// on tensor cores but FFMA-bound here; the tcgen05 3xTF32 version is the round-2 target.
#include "gemm_simt.cuh"

namespace tq {

__global__ void __launch_bounds__(GT_THREADS, 2)
err_feedback_kernel(float* __restrict__ W, int64_t ldw, int n, const float* __restrict__ E, int64_t lde,
                    const float* __restrict__ Hinv, int64_t ldh, const int32_t* __restrict__ blk_idx, int blk0,
                    int b, const int32_t* __restrict__ rem_idx, int rem0, int rem) {
    __shared__ GemmSmem sm;
    __shared__ int s_cols[GT_N];          // absolute column of each of this tile's remaining positions
    __shared__ float s_rdiag[512];        // clamped Hinv diagonal of the block's columns
    __shared__ int s_bcols[512];
    const int i0 = blockIdx.y * GT_M;     // row tile
    const int j0 = blockIdx.x * GT_N;     // tile of remaining positions
    for (int j = threadIdx.x; j < GT_N; j += GT_THREADS)
        s_cols[j] = (j0 + j < rem) ? (rem_idx ? rem_idx[j0 + j] : rem0 + j0 + j) : -1;
    for (int k = threadIdx.x; k < b; k += GT_THREADS) {
        const int c = blk_idx ? blk_idx[k] : blk0 + k;
        s_bcols[k] = c;
        s_rdiag[k] = fmaxf(Hinv[(int64_t)c * ldh + c], kTiny);
    }
    __syncthreads();
    float acc[8][8];
    gemm_tile<WALK_K, WALK_MN>(
        sm, 0, b,
        [&](int k, int i) { return (i0 + i < n) ? E[(int64_t)(i0 + i) * lde + k] : 0.f; },
        [&](int k, int j) {
            const int c = s_cols[j];
            return (c >= 0) ? __fdiv_rn(Hinv[(int64_t)s_bcols[k] * ldh + c], s_rdiag[k]) : 0.f;
        },
        acc);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = i0 + gt_row(ty, i);
        if (r >= n) continue;
        float* wrow = W + (int64_t)r * ldw;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = s_cols[gt_col(tx, j)];
            if (c >= 0) wrow[c] = __fsub_rn(wrow[c], acc[i][j]);
        }
    }
}

int launch_err_feedback(float* W, int64_t ldw, int64_t n, const float* E, int64_t lde, const float* Hinv,
                        int64_t ldh, const int32_t* blk_idx, int64_t blk0, int64_t b, const int32_t* rem_idx,
                        int64_t rem0, int64_t rem, cudaStream_t st) {
    if (rem <= 0) return 0;
    dim3 grid((unsigned)ceil_div(rem, GT_N), (unsigned)ceil_div(n, GT_M));
    err_feedback_kernel<<<grid, GT_THREADS, 0, st>>>(W, ldw, (int)n, E, lde, Hinv, ldh, blk_idx, (int)blk0, (int)b,
                                                     rem_idx, (int)rem0, (int)rem);
    TQ_LAUNCH_CHECK("err_feedback_kernel");
    return 0;
}

}  // namespace tq

extern "C" int tq_err_feedback(float* W, int64_t ldw, int64_t n, const float* E, int64_t lde, const float* Hinv,
                               int64_t ldh, const int32_t* blk_idx, int64_t blk0, int64_t b,
                               const int32_t* rem_idx, int64_t rem0, int64_t rem, void* stream) {
    using namespace tq;
    TQ_CHECK_ARG(W && E && Hinv, "tq_err_feedback: null pointer");
    TQ_CHECK_ARG(n > 0 && b > 0 && b <= 512 && rem >= 0 && lde >= b, "tq_err_feedback: bad shape");
    return launch_err_feedback(W, ldw, n, E, lde, Hinv, ldh, blk_idx, blk0, b, rem_idx, rem0, rem,
                               (cudaStream_t)stream);
}
