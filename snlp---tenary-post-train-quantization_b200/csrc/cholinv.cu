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
// A non-positive pivot is reported through *info (1-based index, LAPACK convention); the caller takes
// the reference's pinv route (gptq.py:104-106).
#include "gemm_simt.cuh"

namespace tq {

constexpr int CB = 128;            // panel width
constexpr int CB_LD = CB + 1;      // shared-memory row stride
constexpr int DIAG_THREADS = 512;

__global__ void __launch_bounds__(DIAG_THREADS)
chol_diag_kernel(float* __restrict__ A, int64_t ld, int m, int k0, float* __restrict__ Dk, int* __restrict__ info) {
    extern __shared__ float sh[];
    float* S = sh;                       // [CB][CB_LD]  A_kk -> L_kk
    float* V = sh + CB * CB_LD;          // [CB][CB_LD]  L_kk^-1
    const int tid = threadIdx.x;
    const int nb = min(CB, m - k0);
    for (int e = tid; e < CB * CB; e += DIAG_THREADS) {
        const int i = e / CB, j = e % CB;
        float v = (i == j) ? 1.f : 0.f;                          // identity padding for a ragged last panel
        if (i < nb && j < nb) v = (j <= i) ? A[(int64_t)(k0 + i) * ld + k0 + j] : 0.f;
        S[i * CB_LD + j] = v;
        V[i * CB_LD + j] = 0.f;
    }
    __syncthreads();
    // right-looking Cholesky of the block
    for (int j = 0; j < CB; ++j) {
        if (tid == 0) {
            float p = S[j * CB_LD + j];
            if (!(p > 0.f)) {                                    // also catches NaN
                if (j < nb) atomicCAS(info, 0, k0 + j + 1);
                p = 1.f;                                         // keep going so nothing downstream divides by zero
            }
            S[j * CB_LD + j] = sqrtf(p);
        }
        __syncthreads();
        const float d = S[j * CB_LD + j];
        for (int i = j + 1 + tid; i < CB; i += DIAG_THREADS) S[i * CB_LD + j] = __fdiv_rn(S[i * CB_LD + j], d);
        __syncthreads();
        const int t = CB - 1 - j;                                // trailing size
        for (int e = tid; e < t * t; e += DIAG_THREADS) {
            const int i = j + 1 + e / t, c = j + 1 + e % t;
            if (c <= i) S[i * CB_LD + c] = fmaf(-S[i * CB_LD + j], S[c * CB_LD + j], S[i * CB_LD + c]);
        }
        __syncthreads();
    }
    // V = L^-1 : column c by forward substitution, 4 lanes per column share each dot product
    {
        const int c = tid >> 2, sub = tid & 3;
        if (sub == 0) V[c * CB_LD + c] = __fdiv_rn(1.f, S[c * CB_LD + c]);
        __syncwarp();
        for (int i = 1; i < CB; ++i) {
            float part = 0.f;
            if (i > c)
                for (int q = c + sub; q < i; q += 4) part = fmaf(S[i * CB_LD + q], V[q * CB_LD + c], part);
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            if (i > c && sub == 0) V[i * CB_LD + c] = __fdiv_rn(-part, S[i * CB_LD + i]);
            __syncwarp();
        }
    }
    __syncthreads();
    for (int e = tid; e < CB * CB; e += DIAG_THREADS) {
        const int i = e / CB, j = e % CB;
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

}  // namespace tq

extern "C" int64_t tq_chol_workspace_floats(int64_t m) {
    const int64_t panels = tq::ceil_div(m, tq::CB);
    return m * m + panels * tq::CB * tq::CB;
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

    static bool attr_set = false;
    const int diag_smem = 2 * CB * CB_LD * (int)sizeof(float);
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
            chol_syrk_kernel<<<dim3(tiles, tiles), GT_THREADS, 0, st>>>(L, ld, M, k0);
            TQ_LAUNCH_CHECK("chol_syrk_kernel");
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
            trtri_update_kernel<<<dim3(ctiles, (unsigned)ceil_div(below, GT_M)), GT_THREADS, 0, st>>>(X, ld, L, ld, M, k0);
            TQ_LAUNCH_CHECK("trtri_update_kernel");
        }
    }
    // phase 3: Hinv = X'X (upper), then mirror
    {
        const int nt = (int)ceil_div(m, GT_M);
        lauum_kernel<<<dim3(nt, nt), GT_THREADS, 0, st>>>(Hinv, ld, X, ld, M);
        TQ_LAUNCH_CHECK("lauum_kernel");
    }
    return tq_symmetrize(Hinv, ld, m, stream);
}
