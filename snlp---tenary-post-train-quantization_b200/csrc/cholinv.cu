// A3: Hinv = potri(potrf(Hd)) as a full symmetric fp32 matrix.
// Reference: gptq.py:101-103 (torch.linalg.cholesky + torch.cholesky_inverse; main.py:137-139).
//
// Three right-looking blocked phases, every bulk update a K = 128 fp32 GEMM on the shared CUDA-core
// tile (gemm_simt.cuh), every 128x128 diagonal block handled by one CTA entirely in shared memory:
//   1. potrf : for each panel k   L_kk = chol(A_kk), D_k = L_kk^-1            (chol_diag_kernel)
//                                 L_ik = A_ik D_k'            for i > k       (chol_trsm_kernel)
//                                 A_ij -= L_ik L_jk'          for i >= j > k  (chol_syrk_kernel)
//   2. trtri : X = L^-1 by block forward substitution on R = I:
//                                 X_k,: = D_k R_k,:                           (trtri_row_kernel)
//                                 R_i,: -= L_ik X_k,:         for i > k       (trtri_update_kernel)
//   3. lauum : Hinv = X'X (upper tiles, K range starts at the tile's column), then mirrored.
// By default the three bulk updates (chol_syrk, trtri_update, lauum: all of the O(m^3) work) run on the
// tensor cores through the 3xTF32 GEMM of gemm_tc.cu, fed with (hi, lo) splits of the panels; the fp32
// CUDA-core kernels below remain as the reference path (environment TQ_CHOL_FFMA=1).
// A non-positive pivot is reported through *info (1-based index, LAPACK convention); the caller takes
// the reference's pinv route (gptq.py:104-106).
#include "gemm_simt.cuh"

#include <cuda.h>
#include <stdlib.h>

namespace tq {

constexpr int CB = 128;            // panel width
constexpr int CB_LD = CB + 1;      // shared-memory row stride
constexpr int DIAG_THREADS = 512;

// One CTA factors a 128x128 diagonal block and inverts its Cholesky factor, all in shared memory, in four
// 32-column steps:  warp 0 factors the 32x32 diagonal sub-block with its rows in registers (pivots and
// multipliers travel by shuffle), the rows below are solved against it one thread per row, the trailing
// sub-blocks take the rank-32 update.  The inverse is built the same way: four warps invert the 32x32
// diagonal sub-blocks, then the off-diagonal sub-blocks follow level by level as 32x32 products.
constexpr int SB = 32;                 // sub-block width
constexpr int NSB = CB / SB;           // 4

__device__ __forceinline__ void diag_factor_32(float* S, int o, int nb, int k0, int* info, int lane) {
    float a[SB];
#pragma unroll
    for (int c = 0; c < SB; ++c) a[c] = S[(o + lane) * CB_LD + o + c];
#pragma unroll
    for (int j = 0; j < SB; ++j) {
        float pj = __shfl_sync(0xffffffffu, a[j], j);
        if (!(pj > 0.f)) {                                   // also catches NaN
            if (lane == 0 && o + j < nb) atomicCAS(info, 0, k0 + o + j + 1);
            pj = 1.f;                                        // keep going so nothing downstream divides by zero
        }
        const float dj = sqrtf(pj);
        const float l = (lane == j) ? dj : __fdiv_rn(a[j], dj);
        a[j] = l;
#pragma unroll
        for (int c = j + 1; c < SB; ++c) {
            const float lc = __shfl_sync(0xffffffffu, l, c);
            if (lane >= c) a[c] = fmaf(-l, lc, a[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < SB; ++c) S[(o + lane) * CB_LD + o + c] = (c <= lane) ? a[c] : 0.f;
}

__global__ void __launch_bounds__(DIAG_THREADS)
chol_diag_kernel(float* __restrict__ A, int64_t ld, int m, int k0, float* __restrict__ Dk, int* __restrict__ info) {
    extern __shared__ float sh[];
    float* S = sh;                       // [CB][CB_LD]  A_kk -> L_kk
    float* V = sh + CB * CB_LD;          // [CB][CB_LD]  L_kk^-1
    float* Tm = sh + 2 * CB * CB_LD;     // [3][SB][SB+1] scratch for the off-diagonal inverse blocks
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nb = min(CB, m - k0);
    for (int e = tid; e < CB * CB; e += DIAG_THREADS) {
        const int i = e >> 7, j = e & (CB - 1);
        float v = (i == j) ? 1.f : 0.f;                          // identity padding for a ragged last panel
        if (i < nb && j < nb) v = (j <= i) ? A[(int64_t)(k0 + i) * ld + k0 + j] : 0.f;
        S[i * CB_LD + j] = v;
        V[i * CB_LD + j] = 0.f;
    }
    __syncthreads();

    // ---- L = chol(S), right-looking over 32-column steps
    for (int d = 0; d < NSB; ++d) {
        const int o = d * SB;
        if (warp == 0) diag_factor_32(S, o, nb, k0, info, lane);
        __syncthreads();
        const int below = CB - o - SB;                           // rows under the diagonal sub-block
        if (tid < below) {                                       // x L_dd' = a : one row per thread
            float* row = S + (o + SB + tid) * CB_LD + o;
            float x[SB];
#pragma unroll
            for (int j = 0; j < SB; ++j) {
                float s = row[j];
#pragma unroll
                for (int k = 0; k < j; ++k) s = fmaf(-x[k], S[(o + j) * CB_LD + o + k], s);
                x[j] = __fdiv_rn(s, S[(o + j) * CB_LD + o + j]);
            }
#pragma unroll
            for (int j = 0; j < SB; ++j) row[j] = x[j];
        }
        __syncthreads();
        if (below > 0) {                                         // trailing -= P P', 4 threads per row
            const int r = o + SB + (tid >> 2);
            if (r < CB) {
                float pr[SB];
#pragma unroll
                for (int k = 0; k < SB; ++k) pr[k] = S[r * CB_LD + o + k];
                for (int c = o + SB + (tid & 3); c <= r; c += 4) {
                    float s = 0.f;
#pragma unroll
                    for (int k = 0; k < SB; ++k) s = fmaf(pr[k], S[c * CB_LD + o + k], s);
                    S[r * CB_LD + c] -= s;
                }
            }
        }
        __syncthreads();
    }

    // ---- V = L^-1: diagonal sub-blocks, one warp each, lane = column of the inverse
    if (warp < NSB) {
        const int o = warp * SB;
        float x[SB];
#pragma unroll
        for (int i = 0; i < SB; ++i) {
            float s = (i == lane) ? 1.f : 0.f;
#pragma unroll
            for (int k = 0; k < i; ++k) s = fmaf(-S[(o + i) * CB_LD + o + k], x[k], s);
            x[i] = (i >= lane) ? __fdiv_rn(s, S[(o + i) * CB_LD + o + i]) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < SB; ++i) V[(o + i) * CB_LD + o + lane] = x[i];
    }
    __syncthreads();
    // off-diagonal sub-blocks, level by level:  V_ij = -V_ii * sum_{k=j}^{i-1} L_ik V_kj
    for (int lev = 1; lev < NSB; ++lev) {
        const int nblk = NSB - lev;                              // blocks (i, j) = (j + lev, j)
        const int g = tid >> 7, t = tid & 127;                   // one 128-thread group per block
        const int bj = g, bi = g + lev;
        const int er = t >> 2, ec0 = (t & 3) * 8;                // each thread: row er, 8 columns
        if (g < nblk) {
            float acc[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] = 0.f;
            for (int k = bj * SB; k < bi * SB; ++k) {
                const float l = S[(bi * SB + er) * CB_LD + k];
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[q] = fmaf(l, V[k * CB_LD + bj * SB + ec0 + q], acc[q]);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) Tm[(g * SB + er) * (SB + 1) + ec0 + q] = acc[q];
        }
        __syncthreads();
        if (g < nblk) {
            float acc[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] = 0.f;
            for (int k = 0; k <= er; ++k) {                      // V_ii is lower triangular
                const float v = V[(bi * SB + er) * CB_LD + bi * SB + k];
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[q] = fmaf(v, Tm[(g * SB + k) * (SB + 1) + ec0 + q], acc[q]);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) V[(bi * SB + er) * CB_LD + bj * SB + ec0 + q] = -acc[q];
        }
        __syncthreads();
    }
    for (int e = tid; e < CB * CB; e += DIAG_THREADS) {
        const int i = e >> 7, j = e & (CB - 1);
        Dk[e] = V[i * CB_LD + j];
        if (i < nb && j < nb && j <= i) A[(int64_t)(k0 + i) * ld + k0 + j] = S[i * CB_LD + j];
    }
}

// L_ik = A_ik D_k'   (rows below the panel), in place
__global__ void __launch_bounds__(GT_THREADS, 2)
chol_trsm_kernel(float* __restrict__ A, int64_t ld, int m, int k0, const float* __restrict__ Dk) {
    __shared__ GemmSmem sm;
    const int nb = min(CB, m - k0);
    const int i0 = k0 + CB + blockIdx.x * GT_M;
    float acc[8][8];
    gemm_tile<WALK_K, WALK_K>(
        sm, 0, nb,
        [&](int c, int i) { return (i0 + i < m) ? A[(int64_t)(i0 + i) * ld + k0 + c] : 0.f; },
        [&](int c, int j) { return Dk[j * CB + c]; }, acc);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = i0 + gt_row(ty, i);
        if (r >= m) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = gt_col(tx, j);
            if (c < nb) A[(int64_t)r * ld + k0 + c] = acc[i][j];
        }
    }
}

// A_ij -= L_ik L_jk'   for tiles i >= j below/right of the panel
__global__ void __launch_bounds__(GT_THREADS, 2)
chol_syrk_kernel(float* __restrict__ A, int64_t ld, int m, int k0) {
    if (blockIdx.y < blockIdx.x) return;
    __shared__ GemmSmem sm;
    const int i0 = k0 + CB + blockIdx.y * GT_M, j0 = k0 + CB + blockIdx.x * GT_N;
    float acc[8][8];
    gemm_tile<WALK_K, WALK_K>(
        sm, 0, CB,
        [&](int c, int i) { return (i0 + i < m) ? A[(int64_t)(i0 + i) * ld + k0 + c] : 0.f; },
        [&](int c, int j) { return (j0 + j < m) ? A[(int64_t)(j0 + j) * ld + k0 + c] : 0.f; }, acc);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = i0 + gt_row(ty, i);
        if (r >= m) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = j0 + gt_col(tx, j);
            if (c <= r) A[(int64_t)r * ld + c] -= acc[i][j];
        }
    }
}

__global__ void set_identity_kernel(float* __restrict__ X, int64_t ld, int m) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) X[(int64_t)i * ld + i] = 1.f;
}

// X_k,: = D_k R_k,:   over columns [0, k0+nb), in place on the panel's rows
__global__ void __launch_bounds__(GT_THREADS, 2)
trtri_row_kernel(float* __restrict__ X, int64_t ld, int m, int k0, const float* __restrict__ Dk) {
    __shared__ GemmSmem sm;
    const int nb = min(CB, m - k0);
    const int c0 = blockIdx.x * GT_N;
    const int cend = k0 + nb;
    float acc[8][8];
    gemm_tile<WALK_K, WALK_MN>(
        sm, 0, nb,
        [&](int q, int i) { return Dk[i * CB + q]; },
        [&](int q, int j) { return (c0 + j < cend) ? X[(int64_t)(k0 + q) * ld + c0 + j] : 0.f; }, acc);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int ri = gt_row(ty, i);
        if (ri >= nb) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = c0 + gt_col(tx, j);
            if (c < cend) X[(int64_t)(k0 + ri) * ld + c] = acc[i][j];
        }
    }
}

// R_i,: -= L_ik X_k,:   for row tiles below the panel, columns [0, k0+nb)
__global__ void __launch_bounds__(GT_THREADS, 2)
trtri_update_kernel(float* __restrict__ X, int64_t ld, const float* __restrict__ L, int64_t ldl, int m, int k0) {
    __shared__ GemmSmem sm;
    const int nb = min(CB, m - k0);
    const int i0 = k0 + CB + blockIdx.y * GT_M;
    const int c0 = blockIdx.x * GT_N;
    const int cend = k0 + nb;
    float acc[8][8];
    gemm_tile<WALK_K, WALK_MN>(
        sm, 0, nb,
        [&](int q, int i) { return (i0 + i < m) ? L[(int64_t)(i0 + i) * ldl + k0 + q] : 0.f; },
        [&](int q, int j) { return (c0 + j < cend) ? X[(int64_t)(k0 + q) * ld + c0 + j] : 0.f; }, acc);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = i0 + gt_row(ty, i);
        if (r >= m) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = c0 + gt_col(tx, j);
            if (c < cend) X[(int64_t)r * ld + c] -= acc[i][j];
        }
    }
}

// Hinv (upper tiles) = X'X with X lower triangular: sum over rows q >= the tile's first column
__global__ void __launch_bounds__(GT_THREADS, 2)
lauum_kernel(float* __restrict__ Hinv, int64_t ldh, const float* __restrict__ X, int64_t ld, int m) {
    if (blockIdx.y > blockIdx.x) return;
    __shared__ GemmSmem sm;
    const int i0 = blockIdx.y * GT_M, j0 = blockIdx.x * GT_N;
    float acc[8][8];
    gemm_tile<WALK_MN, WALK_MN>(
        sm, j0, m,
        [&](int q, int i) { return (i0 + i < m) ? X[(int64_t)q * ld + i0 + i] : 0.f; },
        [&](int q, int j) { return (j0 + j < m) ? X[(int64_t)q * ld + j0 + j] : 0.f; }, acc);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = i0 + gt_row(ty, i);
        if (r >= m) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = j0 + gt_col(tx, j);
            if (c < m) Hinv[(int64_t)r * ldh + c] = acc[i][j];
        }
    }
}

int launch_split(const float*, int64_t, int64_t, int64_t, float*, float*, int64_t, int, cudaStream_t);     // gemm_tc.cu
int launch_gemm_tf32x3(int, float*, int64_t, int64_t, int64_t, int64_t, const float*, const float*, int64_t, const float*,
                       const float*, int64_t, const int32_t*, int64_t, cudaStream_t);
struct GemmOperands {
    CUtensorMap ah, al, bh, bl;
    int64_t K;
};
int gemm_operands_encode(GemmOperands*, const float*, const float*, int64_t, int64_t, const float*, const float*, int64_t,
                         int64_t, int64_t);
int launch_gemm_tf32x3_ops(int, float*, int64_t, int64_t, int64_t, const GemmOperands*, const int32_t*, int64_t, cudaStream_t);
enum { GXM_SUB_LOWER = 1, GXM_SUB_RECT = 2, GXM_STORE_UPPER = 3 };   // GxMode of gemm_tc.cu

static inline int64_t chol_split_floats(int64_t m) {
    const int64_t mp = ceil_div(m, CB) * CB;
    const int64_t a = 2 * m * m, b = 4 * mp * CB;
    return a > b ? a : b;
}

}  // namespace tq

extern "C" int64_t tq_chol_workspace_floats(int64_t m) {
    const int64_t panels = tq::ceil_div(m, tq::CB);
    return m * m + panels * tq::CB * tq::CB + tq::chol_split_floats(m);
}

extern "C" int tq_chol_inverse(float* Hinv, const float* Hd, int64_t m, float* work, int* info_dev, void* stream) {
    using namespace tq;
    TQ_CHECK_ARG(Hinv && Hd && work && info_dev && m > 0 && m < (1 << 30), "tq_chol_inverse: bad arguments");
    TQ_CHECK_ARG(Hinv != Hd, "tq_chol_inverse: Hinv must not alias Hd");
    cudaStream_t st = (cudaStream_t)stream;
    const int M = (int)m;
    const int64_t ld = m;
    float* L = Hinv;                       // potrf runs in the output buffer
    float* X = work;                       // L^-1
    float* D = work + m * m;               // per-panel inverses of the diagonal blocks
    const int panels = (int)ceil_div(m, CB);
    float* S = D + (int64_t)panels * CB * CB;          // (hi, lo) operand splits for the tensor-core updates
    const int64_t mp = (int64_t)panels * CB;
    float *Ph = S, *Pl = S + mp * CB, *Bh = S + 2 * mp * CB, *Bl = S + 3 * mp * CB;   // panel / strip operands, ld = CB
    // TQ_CHOL_FFMA bit mask selects the fp32 CUDA-core kernel per phase: 1 = potrf update, 2 = trtri update, 4 = lauum
    static const int ffma_mask = []() { const char* e = getenv("TQ_CHOL_FFMA"); return e ? atoi(e) : 0; }();
    const bool tc_potrf = !(ffma_mask & 1), tc_trtri = !(ffma_mask & 2), tc_lauum = !(ffma_mask & 4);
    int rc;
    // operand descriptors, encoded once per inversion with the largest extents (see gemm_operands_encode)
    GemmOperands ops_potrf, ops_trtri;
    if (panels > 1) {
        if (tc_potrf && (rc = gemm_operands_encode(&ops_potrf, Ph, Pl, CB, mp, Ph, Pl, CB, mp, CB))) return rc;
        if (tc_trtri && (rc = gemm_operands_encode(&ops_trtri, Ph, Pl, CB, mp, Bh, Bl, CB, mp, CB))) return rc;
    }

    static bool attr_set = false;
    const int diag_smem = (2 * CB * CB_LD + 3 * SB * (SB + 1)) * (int)sizeof(float);
    if (!attr_set) {
        TQ_CUDA(cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, diag_smem));
        attr_set = true;
    }
    TQ_CUDA(cudaMemcpyAsync(L, Hd, sizeof(float) * m * m, cudaMemcpyDeviceToDevice, st));
    TQ_CUDA(cudaMemsetAsync(info_dev, 0, sizeof(int), st));
    TQ_CUDA(cudaMemsetAsync(X, 0, sizeof(float) * m * m, st));
    set_identity_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, st>>>(X, ld, M);
    TQ_LAUNCH_CHECK("set_identity_kernel");

    // phase 1: potrf
    for (int k = 0; k < panels; ++k) {
        const int k0 = k * CB;
        float* Dk = D + (int64_t)k * CB * CB;
        chol_diag_kernel<<<1, DIAG_THREADS, diag_smem, st>>>(L, ld, M, k0, Dk, info_dev);
        TQ_LAUNCH_CHECK("chol_diag_kernel");
        const int below = M - k0 - CB;
        if (below > 0) {
            const int tiles = (int)ceil_div(below, GT_M);
            chol_trsm_kernel<<<tiles, GT_THREADS, 0, st>>>(L, ld, M, k0, Dk);
            TQ_LAUNCH_CHECK("chol_trsm_kernel");
            if (tc_potrf) {
                // A_ij -= L_ik L_jk' on the tensor cores: both operands are the freshly solved panel
                if ((rc = launch_split(L + (int64_t)(k0 + CB) * ld + k0, ld, below, CB, Ph, Pl, CB, 0, st))) return rc;
                if ((rc = launch_gemm_tf32x3_ops(GXM_SUB_LOWER, L + (int64_t)(k0 + CB) * ld + (k0 + CB), ld, below, below,
                                                 &ops_potrf, nullptr, 0, st)))
                    return rc;
            } else {
                chol_syrk_kernel<<<dim3(tiles, tiles), GT_THREADS, 0, st>>>(L, ld, M, k0);
                TQ_LAUNCH_CHECK("chol_syrk_kernel");
            }
        }
    }
    // phase 2: X = L^-1
    for (int k = 0; k < panels; ++k) {
        const int k0 = k * CB;
        const int nb = (M - k0 < CB) ? (M - k0) : CB;
        const float* Dk = D + (int64_t)k * CB * CB;
        const int ctiles = (int)ceil_div(k0 + nb, GT_N);
        trtri_row_kernel<<<ctiles, GT_THREADS, 0, st>>>(X, ld, M, k0, Dk);
        TQ_LAUNCH_CHECK("trtri_row_kernel");
        const int below = M - k0 - CB;
        if (below > 0) {
            if (tc_trtri) {
                // R_i,: -= L_ik X_k,: : A = L panel rows, B = the just-finished strip X_k,: transposed to K-major
                const int64_t cend = k0 + nb;
                if ((rc = launch_split(L + (int64_t)(k0 + CB) * ld + k0, ld, below, CB, Ph, Pl, CB, 0, st))) return rc;
                if ((rc = launch_split(X + (int64_t)k0 * ld, ld, nb, cend, Bh, Bl, CB, 1, st))) return rc;
                if ((rc = launch_gemm_tf32x3_ops(GXM_SUB_RECT, X + (int64_t)(k0 + CB) * ld, ld, below, cend, &ops_trtri, nullptr,
                                                 0, st)))
                    return rc;
            } else {
                trtri_update_kernel<<<dim3(ctiles, (unsigned)ceil_div(below, GT_M)), GT_THREADS, 0, st>>>(X, ld, L, ld, M, k0);
                TQ_LAUNCH_CHECK("trtri_update_kernel");
            }
        }
    }
    // phase 3: Hinv = X'X (upper), then mirror
    if (tc_lauum && (m % 4) == 0) {
        // Y = X' (upper triangular), Hinv = Y Y': rows of Y are K-major operands; the K range of tile (i, j <= ... )
        // starts at the tile's first column because Y[i][q] = 0 for q < i
        float *Yh = S, *Yl = S + m * m;
        if ((rc = launch_split(X, ld, m, m, Yh, Yl, m, 1, st))) return rc;
        if ((rc = launch_gemm_tf32x3(GXM_STORE_UPPER, Hinv, ld, m, m, m, Yh, Yl, m, Yh, Yl, m, nullptr, 0, st))) return rc;
    } else {
        const int nt = (int)ceil_div(m, GT_M);
        lauum_kernel<<<dim3(nt, nt), GT_THREADS, 0, st>>>(Hinv, ld, X, ld, M);
        TQ_LAUNCH_CHECK("lauum_kernel");
    }
    return tq_symmetrize(Hinv, ld, m, stream);
}
