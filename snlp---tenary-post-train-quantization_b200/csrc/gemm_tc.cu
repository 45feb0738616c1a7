// fp32-faithful tensor-core GEMM for the fp32 stages:  C (op)= A B'  with A [M, K] and B [N, K] both K-major fp32.
//
// tcgen05 has no fp32 MMA; kind::tf32 keeps 11 significant bits of each operand.  Every operand is therefore
// supplied as a (hi, lo) pair prepared by the producing kernel:  hi = rn_tf32(x), lo = rn_tf32(x - hi), both
// exactly representable in tf32 (|x - hi - lo| <= 2^-23 |x|), and four MMAs accumulate
//     lo*lo' + lo*hi' + hi*lo' + hi*hi'     in fp32 in TMEM
// so each product carries ~2^-22 relative error (fp32 FFMA: 2^-24); measured against the fp32 CUDA-core
// kernels in tests/.  The kernel is bound by the read-modify-write of C, not by the MMAs.  Used by
//   error feedback (gptq.py:173-186)     W[:, rem] -= E C,  A = E, B = C' (built by feedback_coef_kernel)
//   potrf trailing update                A_ij -= L_ik L_jk'
//   trtri update                         R_i,: -= L_ik X_k,:
//   lauum                                Hinv = Y Y' with Y = (L^-1)'
//
//   tile 128 x 128, K slabs of 32 floats (one 128-byte swizzled row), 3 smem stages x (A_hi, A_lo, B_hi, B_lo) x 16 KB
//   TMEM: 2 accumulators x 128 columns; persistent CTAs; roles as in hessian_tc.cu
//   epilogue: TMEM -> registers -> global read-modify-write (the RMW of C is the dominant, compulsory traffic:
//   8 bytes per element per call at K = 128, i.e. HBM-bound -- SURVEY 8d)
#include "tc_common.cuh"

#include <stdlib.h>

namespace tq {

constexpr int GX_BM = 128, GX_BN = 128, GX_BK = 32, GX_UMMA_K = 8;
constexpr int GX_STAGES_RMW = 2;                     // epilogue = read-modify-write through a shared-memory C tile
constexpr int GX_STAGES_TMA = 3;                     // epilogue = TMA store / reduce-add from two 16 KB staging slabs
constexpr int GX_SLAB = GX_BM * GX_BK * 4;            // 16384
constexpr int GX_STAGE_BYTES = 4 * GX_SLAB;           // 65536
constexpr int GX_THREADS = 384;                 // 4 control warps + 8 epilogue warps
constexpr int GX_EPI_WARPS = 8;
constexpr int GX_CT_LD = GX_BN + 1;                // pitch of the C tile in shared memory (bank-conflict-free both ways)
constexpr int GX_COLPART = 2 * GX_EPI_WARPS * 64 * 2 * 4;   // double-buffered per-warp column partials (dot, sq)
constexpr int GX_SMEM = GX_STAGES_RMW * GX_STAGE_BYTES + 256 + GX_BM * GX_CT_LD * 4 + GX_COLPART + 1024;   // stages, barriers, C tile, slack
constexpr int GX_EPI_COLS = 32;                      // one TMA box of C: 128 rows x 32 columns (128-byte swizzled rows)
constexpr int GX_EPI_BYTES = GX_BM * GX_EPI_COLS * 4;                                             // 16384
constexpr int GX_SMEM_TMA = GX_STAGES_TMA * GX_STAGE_BYTES + 2 * GX_EPI_BYTES + 256 + 1024;       // stages, staging, barriers, slack
constexpr int GX_TMEM_COLS = 256;

struct GxProblem {
    int M, N, K;            // C is M x N, contraction length K
    int mt, nt, tiles;      // tile grid and number of scheduled tiles
    int mode;
    float* C;
    int64_t ldc;
    const int32_t* col_idx; // FEEDBACK: absolute column of remaining position j (NULL: col0 + j)
    int col0;
    int a_row0, b_row0;     // first row of A / B inside the (larger) arrays the descriptors were encoded over
    // FEEDBACK only, optional (SSR): statistics of the UPDATED C for the next block selection (reorder.py:36-61),
    // emitted from the epilogue so W is not read again: per 128-row tile partial sums over rows of
    // C[r,j] * wbar[r] and C[r,j]^2 for every remaining position j, and exact row sums of the updated values
    const float* wbar;      // [M] predicted row means of the updated remaining columns
    float* stat_partials;   // [mt][2][N]
    float* rowsum_part;     // [2*nt][M]: row sums of the updated values per (column tile, column half)
    int k_chunk;            // > 0: TMEM accumulates at most k_chunk of K at a time; chunks are summed in the
                            // smem C tile with round-to-nearest fp32 adds (the tensor core's accumulate truncates)
    int debug;              // development switch (env TQ_GX_DEBUG): 1 = drain accumulators only, 2 = no prefetch loads
    int c_row0, c_col0;     // TMA epilogue: coordinates of C[0, 0] inside the matrix the C descriptor was encoded over
    int skip_lolo;          // 1: three products (lo*hi' + hi*lo' + hi*hi'); the lo*lo' term is ~2^-22 of the product
};

__device__ __forceinline__ void gx_decode(const GxProblem& p, int t, int& bi, int& bj) {
    if (p.mode == GX_SUB_LOWER) {
        int j = 0;
        for (;; ++j) {
            const int cnt = p.mt - j;
            if (t < cnt) break;
            t -= cnt;
        }
        bj = j;
        bi = j + t;
    } else if (p.mode == GX_STORE_UPPER) {
        int j = 0;
        for (;; ++j) {
            const int cnt = min(p.mt, j + 1);
            if (t < cnt) break;
            t -= cnt;
        }
        bj = j;
        bi = t;
    } else {
        bj = t / p.mt;
        bi = t - bj * p.mt;
    }
}

// K-major, 128B-swizzled slab [128 rows x 32 floats]: 8 rows x 128 B per swizzle atom, next 8 rows at +1024 B
__device__ __forceinline__ uint64_t gx_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// TMA_EPI = false: the epilogue reads C into shared memory, subtracts, writes back (needed when the columns of C are
//   gathered through col_idx, or when the updated values feed the SSR statistics).
// TMA_EPI = true : C is a plain tile of a matrix described by map_c.  The SMs never read it: the negated accumulator
//   goes TMEM -> registers -> swizzled staging slab -> TMA reduce-add, i.e. the fp32 add `C + (-acc)` is performed by
//   the L2 (round-to-nearest: bit-identical to the subtraction the other path does); GX_STORE_UPPER stores the first K
//   chunk and reduce-adds the later ones.  The shared memory the C tile no longer needs holds a third pipeline stage.
template <bool TMA_EPI>
__global__ void __launch_bounds__(GX_THREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                   const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl,
                   const __grid_constant__ CUtensorMap map_c, const GxProblem p) {
    constexpr int GX_STAGES = TMA_EPI ? GX_STAGES_TMA : GX_STAGES_RMW;
    constexpr int GX_BAR_OFF = GX_STAGES * GX_STAGE_BYTES + (TMA_EPI ? 2 * GX_EPI_BYTES : 0);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_bar = s_base + GX_BAR_OFF;
    auto full_bar = [&](int s) { return s_bar + 8 * s; };
    auto empty_bar = [&](int s) { return s_bar + 8 * (GX_STAGES + s); };
    auto tfull_bar = [&](int a) { return s_bar + 8 * (2 * GX_STAGES + a); };
    auto tempty_bar = [&](int a) { return s_bar + 8 * (2 * GX_STAGES + 2 + a); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + GX_BAR_OFF + 8 * (2 * GX_STAGES + 4));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ah) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_al) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_bh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_bl) : "memory");
        if (TMA_EPI) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < GX_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), GX_EPI_WARPS * 32); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_slot), GX_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // tcgen05 instruction descriptor (kind::tf32): D = f32, A/B = TF32, both K-major, N = 128, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(GX_BN >> 3) << 17) | ((uint32_t)(GX_BM >> 4) << 24);

    auto k_begin_of = [&](int bj) { return (p.mode == GX_STORE_UPPER) ? bj * GX_BN : 0; };

    if (warp == 0 && lane == 0) {
        // ===== TMA producer =====
        int stage = 0, phase = 0;
        for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
            int bi, bj;
            gx_decode(p, t, bi, bj);
            for (int k = k_begin_of(bj); k < p.K; k += GX_BK) {
                mbar_wait(empty_bar(stage), phase ^ 1);
                const uint32_t s0 = s_base + stage * GX_STAGE_BYTES;
                mbar_expect_tx(full_bar(stage), GX_STAGE_BYTES);
                tma_load_2d(s0 + 0 * GX_SLAB, &map_ah, full_bar(stage), k, p.a_row0 + bi * GX_BM);
                tma_load_2d(s0 + 1 * GX_SLAB, &map_al, full_bar(stage), k, p.a_row0 + bi * GX_BM);
                tma_load_2d(s0 + 2 * GX_SLAB, &map_bh, full_bar(stage), k, p.b_row0 + bj * GX_BN);
                tma_load_2d(s0 + 3 * GX_SLAB, &map_bl, full_bar(stage), k, p.b_row0 + bj * GX_BN);
                if (++stage == GX_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer: lo*lo' + lo*hi' + hi*lo' + hi*hi' =====
        int stage = 0, phase = 0, n_item = 0;
        for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
            int bi, bj;
            gx_decode(p, t, bi, bj);
            const int kb0 = k_begin_of(bj);
            const int kc = (p.k_chunk > 0) ? p.k_chunk : p.K;
            for (int c0 = kb0; c0 < p.K; c0 += kc, ++n_item) {
                const int c1 = min(p.K, c0 + kc);
                const int acc = n_item & 1;
                mbar_wait(tempty_bar(acc), ((n_item >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * GX_BN;
                bool first = true;
                for (int k = c0; k < c1; k += GX_BK) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t s0 = s_base + stage * GX_STAGE_BYTES;
#pragma unroll
                    for (int kk = 0; kk < GX_BK / GX_UMMA_K; ++kk) {
                        const uint32_t off = kk * GX_UMMA_K * 4;             // 32 bytes along K inside the swizzled row
                        const uint64_t ah = gx_desc(s0 + 0 * GX_SLAB + off), al = gx_desc(s0 + 1 * GX_SLAB + off);
                        const uint64_t bh = gx_desc(s0 + 2 * GX_SLAB + off), bl = gx_desc(s0 + 3 * GX_SLAB + off);
                        if (!p.skip_lolo) {
                            umma_tf32(tmem_d, al, bl, idesc, first ? 0u : 1u);     // smallest term first
                            first = false;
                        }
                        umma_tf32(tmem_d, al, bh, idesc, first ? 0u : 1u);
                        first = false;
                        umma_tf32(tmem_d, ah, bl, idesc, 1u);
                        umma_tf32(tmem_d, ah, bh, idesc, 1u);
                    }
                    umma_commit(empty_bar(stage));
                    if (k + GX_BK >= c1) umma_commit(tfull_bar(acc));
                    if (++stage == GX_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (TMA_EPI && warp >= 4) {
        // ===== epilogue through TMA.  Eight warps: warp % 4 is the TMEM lane quarter (32 rows), (warp - 4) / 4 the half of
        // the tile's columns; each half owns one 16 KB staging slab (128 rows x 32 columns, 128B-swizzled like the box of
        // map_c) and walks its two 32-column groups: tcgen05.ld -> negate / mask -> st.shared -> one elected thread issues
        // the bulk tensor reduce-add.  Nothing of C is read by the SM.
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        const bool leader = (q == 0 && lane == 0);
        const uint32_t s_stage = s_base + GX_STAGES * GX_STAGE_BYTES + half * GX_EPI_BYTES;
        const bool store_mode = (p.mode == GX_STORE_UPPER);
        const bool lower = (p.mode == GX_SUB_LOWER);
        int n_item = 0;
        for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
            int bi, bj;
            gx_decode(p, t, bi, bj);
            const int r_loc = bi * GX_BM + row;
            const bool row_ok = r_loc < p.M;
            const int kb0 = k_begin_of(bj);
            const int kc = (p.k_chunk > 0) ? p.k_chunk : p.K;
            bool first_chunk = true;
            for (int c0 = kb0; c0 < p.K; c0 += kc, ++n_item) {
                const int acc = n_item & 1;
                mbar_wait(tfull_bar(acc), (n_item >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int cgi = 0; cgi < 2; ++cgi) {
                    const int cg = half * 2 + cgi;
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * GX_BN + cg * GX_EPI_COLS, v);
                    tmem_ld_wait();
                    if (cgi == 1) {
                        tc_fence_before();
                        mbar_arrive(tempty_bar(acc));
                    }
                    const int j0 = bj * GX_BN + cg * GX_EPI_COLS;
                    const bool diag_tile = lower && bi == bj;
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const int j = j0 + c;
                        const bool ok = row_ok && j < p.N && (!diag_tile || j <= r_loc);
                        const float a = __uint_as_float(v[c]);
                        v[c] = __float_as_uint(ok ? (store_mode ? a : -a) : 0.f);
                    }
                    // this half's previous bulk operation must have read the slab (and, when chunks of one tile are
                    // stored then added, must have completed) before the slab is rewritten
                    if (leader) {
                        if (store_mode) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                        else bulk_wait_read<0>();
                    }
                    named_bar_sync(1 + half, 128);
                    const uint32_t sdst = s_stage + row * 128;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const uint32_t a = sdst + ((c ^ (row & 7)) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v[4 * c]), "r"(v[4 * c + 1]),
                                     "r"(v[4 * c + 2]), "r"(v[4 * c + 3])
                                     : "memory");
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(1 + half, 128);
                    if (leader) {
                        if (store_mode && first_chunk) tma_store_2d(&map_c, s_stage, p.c_col0 + j0, p.c_row0 + bi * GX_BM);
                        else tma_reduce_add_2d(&map_c, s_stage, p.c_col0 + j0, p.c_row0 + bi * GX_BM);
                        bulk_commit();
                    }
                }
                first_chunk = false;
            }
        }
        if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else if (!TMA_EPI && warp >= 4 && p.mode == GX_FEEDBACK && p.k_chunk == 0 && p.debug == 0) {
        // ===== epilogue of the error feedback with gathered columns and / or SSR statistics: read-modify-write through
        // a shared-memory C tile, software-pipelined ACROSS tiles.  Each warp owns a private 32 x 64 block of the tile
        // (warp % 4 = TMEM lane quarter, (warp - 4) / 4 = column half), handled as two 32-column groups:
        //    wait for the group's old values (cp.async, issued one tile earlier)  ->  tcgen05.ld (thread = row): subtract
        //    in the smem block, row sums on the fly  ->  walk the block row by row with lanes along the (gathered)
        //    columns: coalesced store + column statistics  ->  cp.async the SAME group of the warp's NEXT tile into the
        //    slot just drained.
        // so the global->shared latency of a group is covered by the other group's work and the wait for the next
        // accumulator, instead of being paid at the top of every tile (6.7 -> ~4 us per tile at 4096 x 10880).
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        float* ctile = reinterpret_cast<float*>(smem + GX_BAR_OFF + 256);
        float* cw = ctile + (q * 32) * GX_CT_LD + half * 64;    // this warp's 32 x 64 block
        float* myrow = cw + lane * GX_CT_LD;
        const bool stats = (p.stat_partials != nullptr);
        float* colpart = ctile + GX_BM * GX_CT_LD;               // [2][8 warps][64 cols][2]
        auto col_of = [&](int bj_, int cg) {
            const int j = bj_ * GX_BN + half * 64 + cg * 32 + lane;
            return (j < p.N) ? (p.col_idx ? p.col_idx[j] : p.col0 + j) : -1;
        };
        auto prefetch = [&](int r0_, int rows_, int colv, int cg) {
            if (colv >= 0) {
                const float* src = p.C + (int64_t)r0_ * p.ldc + colv;
                const uint32_t dst = smem_u32(cw + cg * 32 + lane);
#pragma unroll 8
                for (int rr = 0; rr < rows_; ++rr)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + rr * GX_CT_LD * 4),
                                 "l"(src + (int64_t)rr * p.ldc)
                                 : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        int t = blockIdx.x, n_item = 0, n_tile = 0;
        int bi = 0, bj = 0, col[2] = {-1, -1};
        if (t < p.tiles) {
            gx_decode(p, t, bi, bj);
            col[0] = col_of(bj, 0);
            col[1] = col_of(bj, 1);
            const int r0 = bi * GX_BM + q * 32;
            prefetch(r0, min(32, p.M - r0), col[0], 0);
            prefetch(r0, min(32, p.M - r0), col[1], 1);
        }
        for (; t < p.tiles; ++n_item, ++n_tile) {
            const int r0 = bi * GX_BM + q * 32;                  // first row of this warp's block
            const int rows = min(32, p.M - r0);                  // <= 0 on a ragged last row tile
            // the next tile of this CTA: its gathered column indices are fetched now, used at the end of each group
            const int tn = t + gridDim.x;
            int bin = 0, bjn = 0, coln[2] = {-1, -1};
            if (tn < p.tiles) {
                gx_decode(p, tn, bin, bjn);
                coln[0] = col_of(bjn, 0);
                coln[1] = col_of(bjn, 1);
            }
            const int r0n = bin * GX_BM + q * 32;
            const int rowsn = (tn < p.tiles) ? min(32, p.M - r0n) : 0;
            const int acc = n_item & 1;
            mbar_wait(tfull_bar(acc), (n_item >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * GX_BN + half * 64;
            const float wb = (stats && r0 + lane < p.M) ? p.wbar[r0 + lane] : 0.f;
            float rs = 0.f;
#pragma unroll
            for (int cg = 0; cg < 2; ++cg) {
                // outstanding cp.async groups here: (this tile, cg), then one younger group -> wait for all but one
                asm volatile("cp.async.wait_group 1;" ::: "memory");
                __syncwarp();
                uint32_t v[32];
                tmem_ld_32x32(taddr + cg * 32, v);
                tmem_ld_wait();
                if (cg == 1) {
                    tc_fence_before();
                    mbar_arrive(tempty_bar(acc));
                }
                const int jbase = bj * GX_BN + half * 64 + cg * 32;
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const float nv = __fsub_rn(myrow[cg * 32 + c], __uint_as_float(v[c]));
                    myrow[cg * 32 + c] = nv;
                    if (jbase + c < p.N) rs += nv;               // exact row sum of the updated values (statistics)
                }
                __syncwarp();
                float dot = 0.f, sq = 0.f;
                const bool okc = col[cg] >= 0;
                float* dstp = p.C + (int64_t)r0 * p.ldc + (okc ? col[cg] : 0);
#pragma unroll 8
                for (int rr = 0; rr < rows; ++rr) {
                    const float wbr = stats ? __shfl_sync(0xffffffffu, wb, rr) : 0.f;    // warp-uniform control flow
                    if (okc) {
                        const float x = cw[rr * GX_CT_LD + cg * 32 + lane];
                        dstp[(int64_t)rr * p.ldc] = x;
                        dot = fmaf(x, wbr, dot);
                        sq = fmaf(x, x, sq);
                    }
                }
                if (stats) {
                    float* cp = colpart + (((n_tile & 1) * GX_EPI_WARPS + (warp - 4)) * 64 + cg * 32 + lane) * 2;
                    cp[0] = dot;
                    cp[1] = sq;
                }
                __syncwarp();
                prefetch(r0n, rowsn, coln[cg], cg);              // (commits an empty group when there is no next tile)
            }
            if (stats) {
                if (r0 + lane < p.M) p.rowsum_part[(int64_t)(bj * 2 + half) * p.M + r0 + lane] = rs;
                // the four row-quarter warps of each column half are summed in a fixed order by the q == 0 warp
                named_bar_sync(2, GX_EPI_WARPS * 32);
                if (q == 0) {
#pragma unroll
                    for (int cg = 0; cg < 2; ++cg) {
                        const int j = bj * GX_BN + half * 64 + cg * 32 + lane;
                        if (j < p.N) {
                            float dot = 0.f, sq = 0.f;
#pragma unroll
                            for (int qq = 0; qq < 4; ++qq) {
                                const float* cp = colpart + (((n_tile & 1) * GX_EPI_WARPS + (half * 4 + qq)) * 64 + cg * 32 + lane) * 2;
                                dot += cp[0];
                                sq += cp[1];
                            }
                            float* out = p.stat_partials + (int64_t)bi * 2 * p.N;
                            out[j] = dot;
                            out[p.N + j] = sq;
                        }
                    }
                }
            }
            t = tn; bi = bin; bj = bjn; col[0] = coln[0]; col[1] = coln[1];
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else if (!TMA_EPI && warp >= 4) {
        // ===== epilogue: read-modify-write of the C tile through shared memory (generic form: K chunks, masks).
        // Eight warps; warp % 4 is the TMEM lane quarter (32 rows), (warp - 4) / 4 the half of the tile's columns.
        //  1. BEFORE the accumulator is awaited, the old values of the warp's 32 x 64 block of C are fetched with
        //     cp.async straight into a padded smem tile: no registers, 64 independent 128-byte row segments in
        //     flight per warp (64 KB per SM) while this tile's MMAs run.  Lanes run along the columns, so every
        //     request is one row segment of 32 consecutive (with SSR: ascending, near-consecutive) columns.
        //  2. tcgen05.ld hands each thread one ROW of accumulators; it subtracts them from its row of the smem
        //     tile (pitch 129: conflict-free).
        //  3. The warp walks its block row by row, lanes along the columns again, and stores coalesced.
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        constexpr int NCG = GX_BN / 64;                          // column groups of 32 per warp
        float* ctile = reinterpret_cast<float*>(smem + GX_BAR_OFF + 256);
        float* cw = ctile + (q * 32) * GX_CT_LD + half * 64;    // this warp's 32 x 64 block
        const bool rmw = (p.mode != GX_STORE_UPPER) && p.debug != 2;
        const bool lower = (p.mode == GX_SUB_LOWER);
        const bool stats = (p.stat_partials != nullptr);
        float* colpart = ctile + GX_BM * GX_CT_LD;               // [2][8 warps][64 cols][2]
        int n_item = 0, n_tile = 0;
        for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++n_tile) {
            int bi, bj;
            gx_decode(p, t, bi, bj);
            const int r0 = bi * GX_BM + q * 32;                  // first row of this warp's block
            const int rows = min(32, p.M - r0);                  // <= 0 on a ragged last row tile
            int col[NCG];
#pragma unroll
            for (int cg = 0; cg < NCG; ++cg) {
                const int j = bj * GX_BN + half * 64 + cg * 32 + lane;      // this lane's position in the N range
                col[cg] = -1;
                if (j < p.N) col[cg] = (p.mode == GX_FEEDBACK) ? (p.col_idx ? p.col_idx[j] : p.col0 + j) : j;
            }
            if (rmw) {
#pragma unroll
                for (int cg = 0; cg < NCG; ++cg) {
                    if (col[cg] >= 0) {
                        const float* src = p.C + (int64_t)r0 * p.ldc + col[cg];
                        const uint32_t dst = smem_u32(cw + cg * 32 + lane);
#pragma unroll 8
                        for (int rr = 0; rr < rows; ++rr)
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + rr * GX_CT_LD * 4),
                                         "l"(src + (int64_t)rr * p.ldc)
                                         : "memory");
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            float* myrow = cw + lane * GX_CT_LD;
            const int kb0 = k_begin_of(bj);
            const int kc = (p.k_chunk > 0) ? p.k_chunk : p.K;
            bool first_chunk = true;
            for (int c0 = kb0; c0 < p.K; c0 += kc, ++n_item) {
                const int acc = n_item & 1;
                mbar_wait(tfull_bar(acc), (n_item >> 1) & 1);
                tc_fence_after();
                if (first_chunk) {
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                    __syncwarp();
                }
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * GX_BN + half * 64;
#pragma unroll
                for (int cg = 0; cg < NCG; ++cg) {
                    uint32_t v[32];
                    tmem_ld_32x32(taddr + cg * 32, v);
                    tmem_ld_wait();
                    if (cg == NCG - 1) {
                        tc_fence_before();
                        mbar_arrive(tempty_bar(acc));
                    }
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const float a = __uint_as_float(v[c]);
                        const float cur = myrow[cg * 32 + c];
                        // RMW modes subtract every chunk from the fetched C; STORE mode sums the chunks
                        myrow[cg * 32 + c] = rmw ? __fsub_rn(cur, a) : (first_chunk ? a : __fadd_rn(cur, a));
                    }
                }
                first_chunk = false;
            }
            if (stats) {
                // exact row sum of the updated values of this thread's row (its 64 columns of this tile)
                float rs = 0.f;
                const int jbase = bj * GX_BN + half * 64;
#pragma unroll 8
                for (int c = 0; c < 64; ++c)
                    if (jbase + c < p.N) rs += myrow[c];
                if (r0 + lane < p.M) p.rowsum_part[(int64_t)(bj * 2 + half) * p.M + r0 + lane] = rs;
            }
            __syncwarp();
            if (p.debug != 1) {
                const float wb = (stats && r0 + lane < p.M) ? p.wbar[r0 + lane] : 0.f;
#pragma unroll
                for (int cg = 0; cg < NCG; ++cg) {
                    float dot = 0.f, sq = 0.f;
                    const bool okc = col[cg] >= 0;
                    float* dstp = p.C + (int64_t)r0 * p.ldc + (okc ? col[cg] : 0);
#pragma unroll 8
                    for (int rr = 0; rr < rows; ++rr) {
                        const float wbr = stats ? __shfl_sync(0xffffffffu, wb, rr) : 0.f;    // warp-uniform control flow
                        if (okc) {
                            const float v = cw[rr * GX_CT_LD + cg * 32 + lane];
                            if (!lower || col[cg] <= r0 + rr) dstp[(int64_t)rr * p.ldc] = v;
                            dot = fmaf(v, wbr, dot);
                            sq = fmaf(v, v, sq);
                        }
                    }
                    if (stats) {
                        float* cp = colpart + (((n_tile & 1) * GX_EPI_WARPS + (warp - 4)) * 64 + cg * 32 + lane) * 2;
                        cp[0] = dot;
                        cp[1] = sq;
                    }
                }
            }
            if (stats) {
                // the four row-quarter warps of each column half are summed in a fixed order by the q == 0 warp
                named_bar_sync(2, GX_EPI_WARPS * 32);
                if (q == 0) {
#pragma unroll
                    for (int cg = 0; cg < NCG; ++cg) {
                        const int j = bj * GX_BN + half * 64 + cg * 32 + lane;
                        if (j < p.N) {
                            float dot = 0.f, sq = 0.f;
#pragma unroll
                            for (int qq = 0; qq < 4; ++qq) {
                                const float* cp = colpart + (((n_tile & 1) * GX_EPI_WARPS + (half * 4 + qq)) * 64 + cg * 32 + lane) * 2;
                                dot += cp[0];
                                sq += cp[1];
                            }
                            float* out = p.stat_partials + (int64_t)bi * 2 * p.N;
                            out[j] = dot;
                            out[p.N + j] = sq;
                        }
                    }
                }
            }
            __syncwarp();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, GX_TMEM_COLS);
}

// ---- operand preparation -------------------------------------------------------------------------
// x = hi + lo' + delta with hi = rn_tf32(x) (11 significant bits), lo' = rn_tf32(x - hi); both are exactly
// representable in tf32, so the tensor core's own operand conversion is exact, and |delta| <= 2^-23 |x|.
__device__ __forceinline__ float rn_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = rn_tf32(x);
    lo = rn_tf32(__fsub_rn(x, hi));
}

// hi/lo copies of a row-major matrix; optionally transposed on the way (out[c][r] = in[r][c])
__global__ void __launch_bounds__(256)
split_kernel(const float* __restrict__ in, int64_t ld_in, int rows, int cols, float* __restrict__ hi,
             float* __restrict__ lo, int64_t ld_out, int transpose) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    if (!transpose) {
        for (int i = ty; i < 32; i += 8) {
            const int r = r0 + i, c = c0 + tx;
            if (r < rows && c < cols) {
                float h, l;
                split_tf32(in[(int64_t)r * ld_in + c], h, l);
                hi[(int64_t)r * ld_out + c] = h;
                lo[(int64_t)r * ld_out + c] = l;
            }
        }
        return;
    }
    for (int i = ty; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + tx;
        tile[i][tx] = (r < rows && c < cols) ? in[(int64_t)r * ld_in + c] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + tx;                           // output row = input column
        if (r < rows && c < cols) {
            float h, l;
            split_tf32(tile[tx][i], h, l);
            hi[(int64_t)c * ld_out + r] = h;
            lo[(int64_t)c * ld_out + r] = l;
        }
    }
}

// B operand of the error feedback, K-major:  coef[j][i] = Hinv[blk_i, rem_j] / clamp(Hinv[blk_i, blk_i], 1e-8)
// (gptq.py:173-181), written as hi/lo.  Reads walk rem_j (ascending, near-contiguous) along rows blk_i of Hinv.
constexpr int COEF_SPAN = 1;          // 32-position sub-tiles per CTA (4 was tried: fewer partials of C 1 for the fold in
                                      // aga_vector_kernel, 23 -> 22 us, but this kernel went 6 -> 15 us)
__global__ void __launch_bounds__(256)
feedback_coef_kernel(const float* __restrict__ Hinv, int64_t ldh, const int32_t* __restrict__ blk_idx, int blk0, int b,
                     const int32_t* __restrict__ rem_idx, int rem0, int rem, float* __restrict__ hi,
                     float* __restrict__ lo, int64_t ldb, float* __restrict__ csum_part) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i0 = blockIdx.y * 32;
    float csum = 0.f;                                     // ty == 0: sum over this CTA's positions of C[i0 + tx, j]
    for (int sp = 0; sp < COEF_SPAN; ++sp) {
        const int j0 = (blockIdx.x * COEF_SPAN + sp) * 32;
        if (j0 >= rem) break;                             // CTA-uniform
        for (int ii = ty; ii < 32; ii += 8) {
            const int i = i0 + ii, j = j0 + tx;
            float v = 0.f;
            if (i < b && j < rem) {
                const int bc = blk_idx ? blk_idx[i] : blk0 + i;
                const int rc = rem_idx ? rem_idx[j] : rem0 + j;
                const float* hrow = Hinv + (int64_t)bc * ldh;
                v = __fdiv_rn(hrow[rc], fmaxf(hrow[bc], kTiny));
            }
            tile[ii][tx] = v;
        }
        __syncthreads();
        if (csum_part != nullptr && ty == 0) {
            // fixed order: position by position inside the sub-tile, sub-tile by sub-tile
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) csum += tile[tx][jj];
        }
        for (int jj = ty; jj < 32; jj += 8) {
            const int j = j0 + jj, i = i0 + tx;
            if (j < rem && i < b) {
                float h, l;
                split_tf32(tile[tx][jj], h, l);
                hi[(int64_t)j * ldb + i] = h;
                lo[(int64_t)j * ldb + i] = l;
            }
        }
        __syncthreads();
    }
    if (csum_part != nullptr && ty == 0 && i0 + tx < b) csum_part[(int64_t)blockIdx.x * b + i0 + tx] = csum;
}

int launch_split(const float* in, int64_t ld_in, int64_t rows, int64_t cols, float* hi, float* lo, int64_t ld_out,
                 int transpose, cudaStream_t st) {
    dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32));
    split_kernel<<<grid, 256, 0, st>>>(in, ld_in, (int)rows, (int)cols, hi, lo, ld_out, transpose);
    TQ_LAUNCH_CHECK("split_kernel");
    return 0;
}

// TMA descriptors of the four operand arrays.  Encoding one costs microseconds of host time, so callers that
// launch many GEMMs over the same buffers (the sweep: one per block; the inverse: two per panel) encode once
// with the LARGEST extents they will use: rows beyond a launch's M / N only feed accumulators the epilogue
// never stores.
int gemm_operands_encode(GemmOperands* ops, const float* Ah, const float* Al, int64_t lda, int64_t Mmax, const float* Bh,
                         const float* Bl, int64_t ldb, int64_t Nmax, int64_t K) {
    TQ_CHECK_ARG((lda % 4) == 0 && (ldb % 4) == 0, "gemm_tf32x3: operand leading dimensions must be multiples of 4");
    int rc;
    if ((rc = make_tmap_2d(&ops->ah, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, Ah, Mmax, K, lda, GX_BM, GX_BK, "gemm_tf32x3(A hi)"))) return rc;
    if ((rc = make_tmap_2d(&ops->al, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, Al, Mmax, K, lda, GX_BM, GX_BK, "gemm_tf32x3(A lo)"))) return rc;
    if ((rc = make_tmap_2d(&ops->bh, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, Bh, Nmax, K, ldb, GX_BN, GX_BK, "gemm_tf32x3(B hi)"))) return rc;
    if ((rc = make_tmap_2d(&ops->bl, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, Bl, Nmax, K, ldb, GX_BN, GX_BK, "gemm_tf32x3(B lo)"))) return rc;
    ops->K = K;
    return 0;
}

// TMA descriptor of a whole output matrix (box 128 rows x 32 columns) for the TMA epilogue; `valid` stays false when
// the matrix cannot be described (base not 16-byte aligned, row pitch not a multiple of 16 bytes): callers then get the
// read-modify-write epilogue.
int gemm_cmap_encode(GemmC* c, float* base, int64_t rows, int64_t cols, int64_t ld) {
    c->valid = false;
    c->base = base;
    c->ld = ld;
    static const bool off = []() { const char* e = getenv("TQ_GX_NO_TMA_EPI"); return e && atoi(e) != 0; }();
    if (off || (ld % 4) != 0 || (reinterpret_cast<uintptr_t>(base) & 15) != 0) return 0;
    int rc = make_tmap_2d(&c->map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, rows, cols, ld, GX_BM, GX_EPI_COLS, "gemm_tf32x3(C)");
    if (rc) return rc;
    c->valid = true;
    return 0;
}

// C (op)= A B' with pre-encoded operands; M, N <= the extents the operands were encoded with
int launch_gemm_tf32x3_ops(int mode, float* C, int64_t ldc, int64_t M, int64_t N, const GemmOperands* ops,
                           const int32_t* col_idx, int64_t col0, cudaStream_t st) {
    return launch_gemm_tf32x3_rows(mode, C, ldc, M, N, ops, 0, 0, col_idx, col0, st);
}

// as above, with A and B starting at rows a_row0 / b_row0 of the arrays the descriptors describe
static int launch_gemm_impl(int mode, float* C, int64_t ldc, int64_t M, int64_t N, const GemmOperands* ops, int64_t a_row0,
                            int64_t b_row0, const int32_t* col_idx, int64_t col0, const float* wbar, float* stat_partials,
                            float* rowsum_part, const GemmC* cmap, int64_t c_row0, int64_t c_col0, cudaStream_t st);

int launch_gemm_tf32x3_rows(int mode, float* C, int64_t ldc, int64_t M, int64_t N, const GemmOperands* ops,
                            int64_t a_row0, int64_t b_row0, const int32_t* col_idx, int64_t col0, cudaStream_t st) {
    return launch_gemm_impl(mode, C, ldc, M, N, ops, a_row0, b_row0, col_idx, col0, nullptr, nullptr, nullptr, nullptr, 0, 0, st);
}

// C = the sub-matrix of `cmap`'s matrix whose first element sits at (c_row0, c_col0); plain (ungathered) columns.
// Takes the TMA epilogue when the descriptor is valid, else the read-modify-write epilogue on the same addresses.
int launch_gemm_tf32x3_at(int mode, const GemmC* cmap, int64_t c_row0, int64_t c_col0, int64_t M, int64_t N,
                          const GemmOperands* ops, int64_t a_row0, int64_t b_row0, cudaStream_t st) {
    float* C = cmap->base + c_row0 * cmap->ld + c_col0;
    return launch_gemm_impl(mode, C, cmap->ld, M, N, ops, a_row0, b_row0, nullptr, 0, nullptr, nullptr, nullptr, cmap, c_row0,
                            c_col0, st);
}

// error feedback with the next block's SSR statistics emitted from the epilogue
int launch_gemm_feedback_stats(float* C, int64_t ldc, int64_t M, int64_t N, const GemmOperands* ops, const int32_t* col_idx,
                               int64_t col0, const float* wbar, float* stat_partials, float* rowsum_part, cudaStream_t st) {
    return launch_gemm_impl(GX_FEEDBACK, C, ldc, M, N, ops, 0, 0, col_idx, col0, wbar, stat_partials, rowsum_part, nullptr, 0, 0, st);
}

static int launch_gemm_impl(int mode, float* C, int64_t ldc, int64_t M, int64_t N, const GemmOperands* ops, int64_t a_row0,
                            int64_t b_row0, const int32_t* col_idx, int64_t col0, const float* wbar, float* stat_partials,
                            float* rowsum_part, const GemmC* cmap, int64_t c_row0, int64_t c_col0, cudaStream_t st) {
    const int64_t K = ops->K;
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    GxProblem p;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.mt = (int)ceil_div(M, GX_BM);
    p.nt = (int)ceil_div(N, GX_BN);
    p.mode = mode;
    p.C = C; p.ldc = ldc; p.col_idx = col_idx; p.col0 = (int)col0;
    p.k_chunk = (mode == GX_STORE_UPPER) ? 256 : 0;
    p.a_row0 = (int)a_row0; p.b_row0 = (int)b_row0;
    p.wbar = wbar; p.stat_partials = stat_partials; p.rowsum_part = rowsum_part;
    static const int dbg = []() { const char* e = getenv("TQ_GX_DEBUG"); return e ? atoi(e) : 0; }();
    p.debug = dbg;
    p.c_row0 = (int)c_row0; p.c_col0 = (int)c_col0;
    static const int products = []() { const char* e = getenv("TQ_GX_PRODUCTS"); return e ? atoi(e) : 4; }();
    p.skip_lolo = (products == 3) ? 1 : 0;
    const bool tma_epi = cmap != nullptr && cmap->valid && col_idx == nullptr && stat_partials == nullptr && dbg == 0;
    if (mode == GX_SUB_LOWER) {
        p.tiles = 0;
        for (int j = 0; j < p.nt; ++j) p.tiles += (p.mt - j > 0) ? p.mt - j : 0;
    } else if (mode == GX_STORE_UPPER) {
        p.tiles = 0;
        for (int j = 0; j < p.nt; ++j) p.tiles += (p.mt < j + 1) ? p.mt : j + 1;
    } else {
        p.tiles = p.mt * p.nt;
    }
    if (p.tiles <= 0) return 0;
    static std::atomic<unsigned long long> attr_mask{0};
    int dev;
    if (dyn_smem_pending(attr_mask, dev)) {
        TQ_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GX_SMEM));
        TQ_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GX_SMEM_TMA));
        dyn_smem_done(attr_mask, dev);
    }
    // Persistent CTAs of this kernel fill an SM's shared memory.  The chains of other linears on other streams are made of
    // short one-CTA kernels (diagonal-block factorisation, block selection, ...); a few SMs are left to them so they do
    // not queue behind every GEMM (TQ_GX_SPARE_SMS, default 8).
    static const int spare = []() { const char* e = getenv("TQ_GX_SPARE_SMS"); return e ? atoi(e) : 8; }();
    int sms = sm_count() - spare;
    if (sms < 1) sms = 1;
    const int grid = p.tiles < sms ? p.tiles : sms;
    if (tma_epi)
        gemm_tf32x3_kernel<true><<<grid, GX_THREADS, GX_SMEM_TMA, st>>>(ops->ah, ops->al, ops->bh, ops->bl, cmap->map, p);
    else
        gemm_tf32x3_kernel<false><<<grid, GX_THREADS, GX_SMEM, st>>>(ops->ah, ops->al, ops->bh, ops->bl, ops->ah, p);
    TQ_LAUNCH_CHECK("gemm_tf32x3_kernel");
    return 0;
}

int launch_gemm_tf32x3(int mode, float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, const float* Ah, const float* Al,
                       int64_t lda, const float* Bh, const float* Bl, int64_t ldb, const int32_t* col_idx, int64_t col0,
                       cudaStream_t st) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    GemmOperands ops;
    int rc;
    if ((rc = gemm_operands_encode(&ops, Ah, Al, lda, M, Bh, Bl, ldb, N, K))) return rc;
    return launch_gemm_tf32x3_ops(mode, C, ldc, M, N, &ops, col_idx, col0, st);
}

int launch_feedback_coef(const float* Hinv, int64_t ldh, const int32_t* blk_idx, int64_t blk0, int64_t b,
                         const int32_t* rem_idx, int64_t rem0, int64_t rem, float* ch, float* cl, int64_t ldb,
                         float* csum_part, cudaStream_t st) {
    dim3 grid((unsigned)ceil_div(rem, 32 * COEF_SPAN), (unsigned)ceil_div(b, 32));
    feedback_coef_kernel<<<grid, 256, 0, st>>>(Hinv, ldh, blk_idx, (int)blk0, (int)b, rem_idx, (int)rem0, (int)rem, ch, cl, ldb,
                                               csum_part);
    TQ_LAUNCH_CHECK("feedback_coef_kernel");
    return 0;
}

// W[:, rem] -= E C on the tensor cores.  E is given as hi/lo [n, lde]; coef_ws holds 2 * rem * ldb floats.
int launch_err_feedback_tc(float* W, int64_t ldw, int64_t n, const float* Eh, const float* El, int64_t lde,
                           const float* Hinv, int64_t ldh, const int32_t* blk_idx, int64_t blk0, int64_t b,
                           const int32_t* rem_idx, int64_t rem0, int64_t rem, float* coef_ws, int64_t ldb,
                           cudaStream_t st) {
    if (rem <= 0) return 0;
    float* ch = coef_ws;
    float* cl = coef_ws + rem * ldb;
    int rc;
    if ((rc = launch_feedback_coef(Hinv, ldh, blk_idx, blk0, b, rem_idx, rem0, rem, ch, cl, ldb, nullptr, st))) return rc;
    return launch_gemm_tf32x3(GX_FEEDBACK, W, ldw, n, rem, b, Eh, El, lde, ch, cl, ldb, rem_idx, rem0, st);
}

}  // namespace tq

extern "C" int tq_split_tf32(const float* x, int64_t ld, int64_t rows, int64_t cols, float* hi, float* lo,
                             int64_t ld_out, void* stream) {
    using namespace tq;
    TQ_CHECK_ARG(x && hi && lo && rows > 0 && cols > 0 && ld >= cols && ld_out >= cols, "tq_split_tf32: bad arguments");
    return launch_split(x, ld, rows, cols, hi, lo, ld_out, 0, (cudaStream_t)stream);
}

extern "C" int64_t tq_err_feedback_tc_workspace_floats(int64_t n, int64_t b, int64_t rem) {
    const int64_t ldb = (b + 3) & ~(int64_t)3;
    return 2 * n * ldb + 2 * rem * ldb;
}

extern "C" int tq_err_feedback_tc(float* W, int64_t ldw, int64_t n, const float* E, int64_t lde, const float* Hinv,
                                  int64_t ldh, const int32_t* blk_idx, int64_t blk0, int64_t b, const int32_t* rem_idx,
                                  int64_t rem0, int64_t rem, float* workspace, void* stream) {
    using namespace tq;
    TQ_CHECK_ARG(W && E && Hinv && workspace, "tq_err_feedback_tc: null pointer");
    TQ_CHECK_ARG(n > 0 && b > 0 && b <= 512 && rem >= 0 && lde >= b, "tq_err_feedback_tc: bad shape");
    TQ_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "tq_err_feedback_tc: workspace must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t ldb = (b + 3) & ~(int64_t)3;
    float* eh = workspace;
    float* el = eh + n * ldb;
    float* coef = el + n * ldb;
    int rc;
    if ((rc = launch_split(E, lde, n, b, eh, el, ldb, 0, st))) return rc;
    return launch_err_feedback_tc(W, ldw, n, eh, el, ldb, Hinv, ldh, blk_idx, blk0, b, rem_idx, rem0, rem, coef, ldb, st);
}
