// fp32 CUDA-core GEMM tile shared by the fp32-exact stages (Cholesky / inverse updates, the fp32
// Hessian path, the error-feedback update).  One CTA = 256 threads = one 128x128 output tile,
// K streamed in slabs of 16 through double-buffered shared memory; each thread owns an 8x8
// register block (two 4-wide groups per dimension so every shared-memory read is a conflict-free
// 128-bit load).  Operands come from functors so each caller can gather / convert / scale on the
// way in (e.g. Hinv[blk, rem] / diag for gptq.py:173-181) without a staging pass through HBM.
#pragma once

#include "common.cuh"

namespace tq {

constexpr int GT_M = 128;
constexpr int GT_N = 128;
constexpr int GT_K = 16;
constexpr int GT_THREADS = 256;
constexpr int GT_LD = 132;   // padded row (floats): keeps float4 alignment, breaks K-major store conflicts

// how consecutive threads walk an operand slab when loading it from global memory
enum OperandWalk { WALK_MN = 0,   // consecutive threads -> consecutive m (or n): operand is MN-contiguous
                   WALK_K = 1 };  // consecutive threads -> consecutive k: operand is K-contiguous

struct GemmSmem {
    float a[2][GT_K][GT_LD];
    float b[2][GT_K][GT_LD];
};

template <int WALK>
__device__ __forceinline__ void slab_coords(int q, int tid, int& k, int& x) {
    const int e = tid + GT_THREADS * q;          // 0 .. 2047
    if (WALK == WALK_MN) { x = e & (GT_M - 1); k = e >> 7; }
    else                 { k = e & (GT_K - 1); x = e >> 4; }
}

// acc[i][j]: i -> rows {ty*4+0..3, 64+ty*4+0..3}, j -> cols {tx*4+0..3, 64+tx*4+0..3}
__device__ __forceinline__ int gt_row(int ty, int i) { return (i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4); }
__device__ __forceinline__ int gt_col(int tx, int j) { return (j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4); }

// LoadA(k, i) -> A[i][k]  (i tile-local in [0,128), k absolute in [k_begin, k_end)), zero when out of range
// LoadB(k, j) -> B[k][j]
template <int WALK_A, int WALK_B, class LoadA, class LoadB>
__device__ __forceinline__ void gemm_tile(GemmSmem& sm, int k_begin, int k_end, LoadA la, LoadB lb,
                                          float (&acc)[8][8]) {
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int nk = (k_end - k_begin + GT_K - 1) / GT_K;
    if (nk <= 0) return;

    float ra[8], rb[8];
    auto fetch = [&](int kt) {
        const int k0 = k_begin + kt * GT_K;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            int k, x;
            slab_coords<WALK_A>(q, tid, k, x);
            ra[q] = (k0 + k < k_end) ? la(k0 + k, x) : 0.f;
            slab_coords<WALK_B>(q, tid, k, x);
            rb[q] = (k0 + k < k_end) ? lb(k0 + k, x) : 0.f;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            int k, x;
            slab_coords<WALK_A>(q, tid, k, x);
            sm.a[buf][k][x] = ra[q];
            slab_coords<WALK_B>(q, tid, k, x);
            sm.b[buf][k][x] = rb[q];
        }
    };

    fetch(0);
    stash(0);
    __syncthreads();
    int cur = 0;
    for (int kt = 0; kt < nk; ++kt) {
        if (kt + 1 < nk) fetch(kt + 1);
#pragma unroll
        for (int k = 0; k < GT_K; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&sm.a[cur][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&sm.a[cur][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&sm.b[cur][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&sm.b[cur][k][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < nk) stash(cur ^ 1);
        __syncthreads();
        cur ^= 1;
    }
}

}  // namespace tq
