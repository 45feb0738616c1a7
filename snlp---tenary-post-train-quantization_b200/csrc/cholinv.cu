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
#include "tc_common.cuh"

#include <cuda.h>
#include <stdlib.h>
#include <unordered_map>
#include <vector>

namespace tq {

constexpr int CB = 128;            // panel width
constexpr int CB_LD = CB + 1;      // shared-memory row stride
constexpr int DIAG_THREADS = 512;

constexpr int SB = 32;                 // sub-block width
constexpr int NSB = CB / SB;           // 4

// ---- diagonal-block kernel, second form --------------------------------------------------------------------------
// One CTA factors a 128x128 diagonal block A = L L' and inverts L, in four 32-column steps.  The only serial part is the
// 32x32 diagonal sub-block: warp 0 factors it with its rows in registers and inverts the factor right away (lane = column
// of the inverse).  Everything else is a small shared-memory matrix product spread over all warps:
//   rows below       L21 = A21 V11'                      (V11 = L11^-1: no substitution chain)
//   trailing update  A22 -= L21 L21'
//   inverse, row d   V[d, j<d] = -V_dd (sum_k L[d,k] V[k,j]); the inner sum runs on warps 1.. WHILE warp 0 factors
//                    block d (it needs only rows of L and V that are already final)
constexpr int DG_P_LD = 3 * SB + 1;      // scratch P: 32 rows x (up to 96 columns)

// 1 / sqrt(p): hardware approximation + one Newton step (relative error ~1 ulp)
__device__ __forceinline__ float fast_rsqrt(float p) {
    float r = rsqrtf(p);
    return r * fmaf(-0.5f * p, r * r, 1.5f);
}

// warp 0: S[o:o+32, o:o+32] -> L_dd (in place, upper part zeroed) and V[o:o+32, o:o+32] = L_dd^-1.
// This code runs once per launch on a single warp, so it has to be SHORT as well as register-resident: fully unrolled over
// the 32 columns (several thousand straight-line instructions) it is bound by instruction fetch (measured 38k cycles per
// sub-block), and with the block left in shared memory by load latency (47k).  Hence the ROTATING register file: lane i
// keeps row i in a[0..31] and after every column the row is shifted left by one, so the current column is always a[0] and
// the loop body does not depend on the column index -- one rolled loop of ~60 instructions.  The multipliers of the other
// rows travel through a 32-float buffer indexed by (row - column), read back as eight 128-bit broadcast loads.
__device__ __forceinline__ void diag_factor_invert_32(float* S, float* V, float* colbuf, float* rdiag, int o, int nb, int k0,
                                                      int* info, int lane) {
    float a[SB];
    float* myrow = S + (o + lane) * CB_LD + o;
#pragma unroll
    for (int c = 0; c < SB; ++c) a[c] = myrow[c];
#pragma unroll 1
    for (int j = 0; j < SB; ++j) {
        const int rel = lane - j;                            // row index relative to the pivot row
        float pj = __shfl_sync(0xffffffffu, a[0], j);
        if (!(pj > 0.f)) {                                   // also catches NaN
            if (lane == 0 && o + j < nb) atomicCAS(info, 0, k0 + o + j + 1);
            pj = 1.f;                                        // keep going so nothing downstream divides by zero
        }
        const float rinv = fast_rsqrt(pj);
        const float l = (rel == 0) ? pj * rinv : ((rel > 0) ? a[0] * rinv : 0.f);
        myrow[j] = l;                                        // L[i][j]; zero above the diagonal
        if (rel == 0) rdiag[j] = rinv;
        float* cb = colbuf + (j & 1) * SB;                   // double-buffered: one warp barrier per column
        cb[rel & (SB - 1)] = l;
        __syncwarp();
        float lc[SB];                                        // lc[t] = multiplier of row j + t
#pragma unroll
        for (int v = 0; v < SB / 4; ++v) {
            const float4 t = *reinterpret_cast<const float4*>(cb + 4 * v);
            lc[4 * v] = t.x; lc[4 * v + 1] = t.y; lc[4 * v + 2] = t.z; lc[4 * v + 3] = t.w;
        }
        // a[t] holds column j + t of this row: rank-1 update and shift left.  No predicate: for t > rel the entry lies above
        // the diagonal of its row, is never read (the stored column is forced to 0 there, rows above the pivot have l = 0)
#pragma unroll
        for (int t = 1; t < SB; ++t) a[t - 1] = fmaf(-l, lc[t], a[t]);
        a[SB - 1] = 0.f;
    }
    __syncwarp();
    // inverse: lane q owns column q of V_dd.  acc[t] accumulates row i + t of  delta - L x ; when x_i = acc[0] / L[i][i] is
    // known every later row takes its contribution (column i of L: warp-uniform addresses, broadcast loads) and the
    // accumulators shift -- again a column-independent loop body
    float acc[SB];
#pragma unroll
    for (int t = 0; t < SB; ++t) acc[t] = (t == lane) ? 1.f : 0.f;
    float* vcol = V + o * CB_LD + o + lane;
#pragma unroll 1
    for (int i = 0; i < SB; ++i) {
        const float x = acc[0] * rdiag[i];                   // 0 for i < q: nothing has reached the accumulator yet
        vcol[i * CB_LD] = x;
        // L[i + t][i] at lcol[t * CB_LD]: constant offsets, no index arithmetic.  Rows past the sub-block (i + t >= 32) read
        // whatever follows in shared memory (S is followed by V and P: in bounds) into accumulators that are never used.
        const float* lcol = S + (o + i) * CB_LD + o + i;
#pragma unroll
        for (int t = 1; t < SB; ++t) acc[t - 1] = fmaf(-lcol[t * CB_LD], x, acc[t]);
        acc[SB - 1] = 0.f;
    }
}

__global__ void __launch_bounds__(DIAG_THREADS)
chol_diag_kernel(float* __restrict__ A, int64_t ld, int m, int k0, float* __restrict__ Dk, int* __restrict__ info,
                 long long* __restrict__ prof) {
    extern __shared__ float sh[];
    int pi = 0;
#define TQ_PROF() do { if (prof && threadIdx.x == 0) prof[pi++] = clock64(); } while (0)
    TQ_PROF();
    float* S = sh;                       // [CB][CB_LD]  A_kk -> L_kk
    float* V = sh + CB * CB_LD;          // [CB][CB_LD]  L_kk^-1
    float* P = sh + 2 * CB * CB_LD;      // [SB][DG_P_LD] partial products of the inverse's current block row
    __shared__ float rdiag[SB];
    __shared__ __align__(16) float colbuf[2 * SB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nb = min(CB, m - k0);
    // tid -> (row = tid / 4, 32 consecutive columns): eight independent 128-bit loads when the block is interior and aligned
    {
        const int i = tid >> 2, j0 = (tid & 3) * 32;
        const bool fast = (nb == CB) && ((ld & 3) == 0) && ((k0 & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
        float vals[32];
        if (fast) {
            const float4* src = reinterpret_cast<const float4*>(A + (int64_t)(k0 + i) * ld + k0 + j0);
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                const float4 t = src[v];
                vals[4 * v] = t.x; vals[4 * v + 1] = t.y; vals[4 * v + 2] = t.z; vals[4 * v + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const int j = j0 + c;
                vals[c] = (i < nb && j < nb) ? A[(int64_t)(k0 + i) * ld + k0 + j] : ((i == j) ? 1.f : 0.f);   // identity padding
            }
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            const int j = j0 + c;
            S[i * CB_LD + j] = (j <= i) ? vals[c] : 0.f;
            V[i * CB_LD + j] = 0.f;
        }
    }
    __syncthreads();
    TQ_PROF();

    for (int d = 0; d < NSB; ++d) {
        const int o = d * SB;
        const int below = CB - o - SB;                           // rows under the diagonal sub-block
        if (warp == 0) {
            diag_factor_invert_32(S, V, colbuf, rdiag, o, nb, k0, info, lane);
        } else if (d > 0) {
            // P[i][c] = sum_{k<o} L[o+i][k] V[k][c]  for c < o (V[k][c] = 0 for k < c).  Thread (i, g) owns the eight columns
            // g + q * (o / 8): neighbouring threads read neighbouring columns of the same row of V (no bank conflicts), and
            // every thread walks the same k (warp-uniform trip count)
            const int t = tid - 32;
            if (t < 4 * o) {
                const int ng = o >> 3, i = t / ng, g = t - i * ng;
                float acc[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[q] = 0.f;
                const float* lrow = S + (o + i) * CB_LD;
#pragma unroll 2
                for (int k = 0; k < o; ++k) {
                    const float l = lrow[k];
                    const float* vrow = V + k * CB_LD + g;
#pragma unroll
                    for (int q = 0; q < 8; ++q) acc[q] = fmaf(l, vrow[q * ng], acc[q]);
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) P[i * DG_P_LD + g + q * ng] = acc[q];
            }
        }
        __syncthreads();
        TQ_PROF();
        // L21 = A21 V_dd' : row r = o + 32 + tid / 4, eight columns per thread; the four threads of a row share a warp
        {
            const int rr = tid >> 2, j0 = (tid & 3) * 8;
            float acc[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] = 0.f;
            float* arow = S + (o + SB + rr) * CB_LD + o;
            if (rr < below) {
                for (int k = 0; k < j0 + 8; ++k) {               // V_dd[j][k] = 0 for k > j
                    const float a = arow[k];
#pragma unroll
                    for (int q = 0; q < 8; ++q) acc[q] = fmaf(a, V[(o + j0 + q) * CB_LD + o + k], acc[q]);
                }
            }
            __syncwarp();
            if (rr < below) {
#pragma unroll
                for (int q = 0; q < 8; ++q) arow[j0 + q] = acc[q];
            }
        }
        // inverse, block row d:  V[o+i][c] = -sum_{k<=i} V_dd[i][k] P[k][c]   (same thread -> column mapping as P)
        if (d > 0 && tid < 4 * o) {
            const int ng = o >> 3, i = tid / ng, g = tid - i * ng;
            float acc[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] = 0.f;
            const float* vdd = V + (o + i) * CB_LD + o;
#pragma unroll 2
            for (int k = 0; k < SB; ++k) {                       // V_dd[i][k] = 0 for k > i
                const float v = vdd[k];
                const float* prow = P + k * DG_P_LD + g;
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[q] = fmaf(v, prow[q * ng], acc[q]);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) V[(o + i) * CB_LD + g + q * ng] = -acc[q];
        }
        __syncthreads();
        TQ_PROF();
        // trailing update A22 -= L21 L21' (entries with column <= row): thread (ti, tj) owns rows ti + T i', columns
        // tj + T j' (i', j' < 4) so that neighbouring threads touch neighbouring rows (pitch 129: distinct banks)
        if (below > 0) {
            const int T = below >> 2;                            // 24, 16, 8
            for (int idx = tid; idx < T * T; idx += DIAG_THREADS) {
                const int ti = idx / T, tj = idx - ti * T;
                if (tj > ti + 3 * T) continue;
                const float* pr = S + (o + SB + ti) * CB_LD + o;
                const float* pc = S + (o + SB + tj) * CB_LD + o;
                float acc[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
                for (int k = 0; k < SB; ++k) {
                    float av[4], bv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) av[i] = pr[i * T * CB_LD + k];
#pragma unroll
                    for (int j = 0; j < 4; ++j) bv[j] = pc[j * T * CB_LD + k];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int r = o + SB + ti + i * T, c = o + SB + tj + j * T;
                        if (c <= r) S[r * CB_LD + c] -= acc[i][j];
                    }
            }
        }
        __syncthreads();
        TQ_PROF();
    }
    for (int e = tid; e < CB * CB; e += DIAG_THREADS) {
        const int i = e >> 7, j = e & (CB - 1);
        Dk[e] = V[i * CB_LD + j];
        if (i < nb && j < nb && j <= i) A[(int64_t)(k0 + i) * ld + k0 + j] = S[i * CB_LD + j];
    }
    __syncthreads();
    TQ_PROF();
#undef TQ_PROF
}

// L_ik = A_ik D_k'   (rows below the panel), in place; when split_hi / split_lo are given the solved panel is also written
// as the (hi, lo) tf32 operand pair of the tensor-core updates (row r - (k0 + CB) of the panel's slot in the stacked split
// arrays, pitch CB) -- one launch less on the per-panel critical path than a separate split pass
__device__ __forceinline__ float trsm_rn_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__global__ void __launch_bounds__(GT_THREADS, 2)
chol_trsm_kernel(float* __restrict__ A, int64_t ld, int m, int k0, const float* __restrict__ Dk,
                 float* __restrict__ split_hi, float* __restrict__ split_lo) {
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
            if (split_hi != nullptr && c < CB) {
                // columns past a ragged panel (c >= nb) are zero operands, exactly what the separate split pass left there
                const float v = (c < nb) ? acc[i][j] : 0.f;
                const float h = trsm_rn_tf32(v);
                const int64_t o = (int64_t)(r - (k0 + CB)) * CB + c;
                split_hi[o] = h;
                split_lo[o] = trsm_rn_tf32(__fsub_rn(v, h));
            }
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

long long* g_diag_prof = nullptr;

// rows of the stacked L-panel splits: panel k contributes the rows below it
static inline int64_t chol_stack_rows(int64_t m) {
    const int64_t panels = ceil_div(m, CB);
    int64_t rows = 0;
    for (int64_t k = 0; k + 1 < panels; ++k) rows += m - (k + 1) * CB;
    return rows;
}

// operand region: phases 1/2 keep the stacked (hi, lo) splits of every L panel plus the transposed strip of phase 2;
// phase 3 reuses the region for Y = (L^-1)' (hi, lo)
static inline int64_t chol_split_floats(int64_t m) {
    const int64_t mp = ceil_div(m, CB) * CB;
    const int64_t a = 2 * m * m, b = 2 * chol_stack_rows(m) * CB + 2 * mp * CB + 2 * CB * CB;
    return a > b ? a : b;
}

// helper streams per caller stream: phase 2 (L^-1) chases phase 1 (potrf) panel by panel on `aux`; the bulk of every
// trailing update runs on `bulk` (lowest priority) behind the look-ahead part
struct ChainStreams {
    cudaStream_t aux = nullptr, bulk = nullptr;
};
static ChainStreams helper_streams_for(cudaStream_t st) {
    static std::unordered_map<cudaStream_t, ChainStreams> pool;
    auto it = pool.find(st);
    if (it != pool.end()) return it->second;
    ChainStreams cs;
    int lo = 0, hi = 0, mine = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);                 // lo = least priority (numerically greatest)
    // the helpers follow the caller's priority: a chain that was placed above the Hessian stream must not have half of
    // its inverse queue behind the Hessian's CTAs; `bulk` sits one level below its chain (but not below other chains' aux)
    if (cudaStreamGetPriority(st, &mine) != cudaSuccess) mine = lo;
    const int bulk_prio = (mine + 1 <= lo) ? mine + 1 : lo;
    if (cudaStreamCreateWithPriority(&cs.aux, cudaStreamNonBlocking, mine) != cudaSuccess) cs.aux = nullptr;
    if (cudaStreamCreateWithPriority(&cs.bulk, cudaStreamNonBlocking, bulk_prio) != cudaSuccess) cs.bulk = nullptr;
    pool[st] = cs;
    return cs;
}
static cudaEvent_t chol_event(size_t i) {
    static std::vector<cudaEvent_t> pool;
    while (pool.size() <= i) {
        cudaEvent_t e = nullptr;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        pool.push_back(e);
    }
    return pool[i];
}

}  // namespace tq

// development aid: device buffer (>= 32 long long) that receives clock64 stamps of the first panel's phases
extern "C" void tq_debug_set_diag_prof(long long* dev_buf) { tq::g_diag_prof = dev_buf; }

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
    float* S = D + (int64_t)panels * CB * CB;          // (hi, lo) operand region, see chol_split_floats
    const int64_t mp = (int64_t)panels * CB;
    const int64_t srows = chol_stack_rows(m);
    float *Sh = S, *Sl = S + srows * CB;                                   // stacked L-panel splits, ld = CB
    float *Bh = S + 2 * srows * CB, *Bl = Bh + mp * CB;                    // transposed strip X_k,: of phase 2
    // TQ_CHOL_FFMA bit mask selects the fp32 CUDA-core kernel per phase: 1 = potrf update, 2 = trtri update, 4 = lauum
    static const int ffma_mask = []() { const char* e = getenv("TQ_CHOL_FFMA"); return e ? atoi(e) : 0; }();
    const bool tc_potrf = !(ffma_mask & 1), tc_trtri = !(ffma_mask & 2), tc_lauum = !(ffma_mask & 4);
    static const bool one_stream = []() { const char* e = getenv("TQ_CHOL_ONE_STREAM"); return e && atoi(e) != 0; }();
    int rc;
    // operand descriptors, encoded once per inversion over the whole stacked / strip arrays
    GemmOperands ops_potrf, ops_trtri;
    GemmC c_L, c_X;                            // output matrices of the tensor-core updates (TMA reduce-add epilogue)
    if ((rc = gemm_cmap_encode(&c_L, L, m, m, ld))) return rc;
    if ((rc = gemm_cmap_encode(&c_X, X, m, m, ld))) return rc;
    if (panels > 1) {
        if (tc_potrf && (rc = gemm_operands_encode(&ops_potrf, Sh, Sl, CB, srows, Sh, Sl, CB, srows, CB))) return rc;
        if (tc_trtri && (rc = gemm_operands_encode(&ops_trtri, Sh, Sl, CB, srows, Bh, Bl, CB, mp, CB))) return rc;
    }
    const ChainStreams helpers = one_stream ? ChainStreams() : helper_streams_for(st);
    cudaStream_t aux = helpers.aux;
    cudaStream_t s2 = aux ? aux : st;                 // stream of phase 2
    // look-ahead: of update k only the column panel k+1 (what the next diagonal block and trsm read) stays on `st`;
    // the rest runs on `bulk` while `st` factors panel k+1 (TQ_CHOL_NO_LOOKAHEAD=1 keeps it all on `st`)
    static const bool no_lookahead = []() { const char* e = getenv("TQ_CHOL_NO_LOOKAHEAD"); return e && atoi(e) != 0; }();
    cudaStream_t bulk = (tc_potrf && !no_lookahead) ? helpers.bulk : nullptr;
    bool bulk_pending = false;                        // an update on `bulk` has not been joined by `st` yet

    static bool attr_set = false;
    const int diag_smem = (2 * CB * CB_LD + SB * DG_P_LD) * (int)sizeof(float);
    if (!attr_set) {
        TQ_CUDA(cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, diag_smem));
        attr_set = true;
    }
    TQ_CUDA(cudaMemcpyAsync(L, Hd, sizeof(float) * m * m, cudaMemcpyDeviceToDevice, st));
    TQ_CUDA(cudaMemsetAsync(info_dev, 0, sizeof(int), st));
    TQ_CUDA(cudaMemsetAsync(X, 0, sizeof(float) * m * m, st));
    set_identity_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, st>>>(X, ld, M);
    TQ_LAUNCH_CHECK("set_identity_kernel");

    // Phase 1 (potrf, on `st`) and phase 2 (X = L^-1, on the helper stream) run panel-interleaved: step k of phase 2
    // needs only D_k and the solved panel k of L, both final once phase 1 has passed its trsm of panel k.
    int64_t off = 0;                                  // first row of panel k inside the stacked splits
    for (int k = 0; k < panels; ++k) {
        const int k0 = k * CB;
        const int nb = (M - k0 < CB) ? (M - k0) : CB;
        float* Dk = D + (int64_t)k * CB * CB;
        const int below = M - k0 - CB;
        // ---- phase 1, panel k
        chol_diag_kernel<<<1, DIAG_THREADS, diag_smem, st>>>(L, ld, M, k0, Dk, info_dev, k == 0 ? g_diag_prof : nullptr);
        TQ_LAUNCH_CHECK("chol_diag_kernel");
        if (below > 0) {
            const int tiles = (int)ceil_div(below, GT_M);
            const bool want_split = tc_potrf || tc_trtri;
            chol_trsm_kernel<<<tiles, GT_THREADS, 0, st>>>(L, ld, M, k0, Dk, want_split ? Sh + off * CB : nullptr,
                                                           want_split ? Sl + off * CB : nullptr);
            TQ_LAUNCH_CHECK("chol_trsm_kernel");
        }
        if (aux) {
            TQ_CUDA(cudaEventRecord(chol_event(3 * k), st));
            TQ_CUDA(cudaStreamWaitEvent(aux, chol_event(3 * k), 0));
        }
        if (below > 0) {
            if (tc_potrf && bulk) {
                // A_ij -= L_ik L_jk' on the tensor cores: both operands are the freshly solved panel.
                // (a) column panel k+1 on `st` -- it also receives the bulk of update k-1, so join that first
                if (bulk_pending) TQ_CUDA(cudaStreamWaitEvent(st, chol_event(3 * (k - 1) + 2), 0));
                const int64_t la = below < CB ? below : CB;
                if ((rc = launch_gemm_tf32x3_at(GX_SUB_LOWER, &c_L, k0 + CB, k0 + CB, below, la, &ops_potrf, off, off, st)))
                    return rc;
                bulk_pending = false;
                // (b) the remaining columns on `bulk`, released only after (a) so the short look-ahead GEMM gets the SMs
                if (below > CB) {
                    TQ_CUDA(cudaEventRecord(chol_event(3 * k + 1), st));
                    TQ_CUDA(cudaStreamWaitEvent(bulk, chol_event(3 * k + 1), 0));
                    if ((rc = launch_gemm_tf32x3_at(GX_SUB_LOWER, &c_L, k0 + 2 * CB, k0 + 2 * CB, below - CB, below - CB,
                                                    &ops_potrf, off + CB, off + CB, bulk)))
                        return rc;
                    TQ_CUDA(cudaEventRecord(chol_event(3 * k + 2), bulk));
                    bulk_pending = true;
                }
            } else if (tc_potrf) {
                if ((rc = launch_gemm_tf32x3_at(GX_SUB_LOWER, &c_L, k0 + CB, k0 + CB, below, below, &ops_potrf, off, off, st)))
                    return rc;
            } else {
                const int tiles = (int)ceil_div(below, GT_M);
                chol_syrk_kernel<<<dim3(tiles, tiles), GT_THREADS, 0, st>>>(L, ld, M, k0);
                TQ_LAUNCH_CHECK("chol_syrk_kernel");
            }
        }
        // ---- phase 2, panel k
        const int ctiles = (int)ceil_div(k0 + nb, GT_N);
        trtri_row_kernel<<<ctiles, GT_THREADS, 0, s2>>>(X, ld, M, k0, Dk);
        TQ_LAUNCH_CHECK("trtri_row_kernel");
        if (below > 0) {
            if (tc_trtri) {
                // R_i,: -= L_ik X_k,: : A = stacked split of the L panel, B = the just-finished strip X_k,: transposed
                const int64_t cend = k0 + nb;
                if ((rc = launch_split(X + (int64_t)k0 * ld, ld, nb, cend, Bh, Bl, CB, 1, s2))) return rc;
                if ((rc = launch_gemm_tf32x3_at(GX_SUB_RECT, &c_X, k0 + CB, 0, below, cend, &ops_trtri, off, 0, s2)))
                    return rc;
            } else {
                trtri_update_kernel<<<dim3(ctiles, (unsigned)ceil_div(below, GT_M)), GT_THREADS, 0, s2>>>(X, ld, L, ld, M, k0);
                TQ_LAUNCH_CHECK("trtri_update_kernel");
            }
            off += below;
        }
    }
    // (every update on `bulk` has been joined by the look-ahead GEMM of the following panel)
    if (aux) {
        TQ_CUDA(cudaEventRecord(chol_event(3 * panels), aux));
        TQ_CUDA(cudaStreamWaitEvent(st, chol_event(3 * panels), 0));
    }

    // phase 3: Hinv = X'X (upper), then mirror
    if (tc_lauum && (m % 4) == 0) {
        // Y = X' (upper triangular), Hinv = Y Y': rows of Y are K-major operands; the K range of tile (i, j <= ... )
        // starts at the tile's first column because Y[i][q] = 0 for q < i
        float *Yh = S, *Yl = S + m * m;
        if ((rc = launch_split(X, ld, m, m, Yh, Yl, m, 1, st))) return rc;
        GemmOperands ops_lauum;
        if ((rc = gemm_operands_encode(&ops_lauum, Yh, Yl, m, m, Yh, Yl, m, m, m))) return rc;
        if ((rc = launch_gemm_tf32x3_at(GX_STORE_UPPER, &c_L, 0, 0, m, m, &ops_lauum, 0, 0, st))) return rc;
    } else {
        const int nt = (int)ceil_div(m, GT_M);
        lauum_kernel<<<dim3(nt, nt), GT_THREADS, 0, st>>>(Hinv, ld, X, ld, M);
        TQ_LAUNCH_CHECK("lauum_kernel");
    }
    return tq_symmetrize(Hinv, ld, m, stream);
}
