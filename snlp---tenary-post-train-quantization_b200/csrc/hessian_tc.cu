// A2 (tensor-core path): H += X'X as a TMA-fed tcgen05 SYRK with fp32 accumulators in TMEM.
// Reference: gptq.py:59-76 (`self.H += inp @ inp.t()`, inp = X'), main.py:128.
//
// X is token-major [Nt, m] fp16/bf16, so with K = tokens BOTH operands of H = X'X are "MN-major"
// slabs of the same tensor: one 2-D tensor map serves A (128 columns) and B (256 columns).
//
//   tile      : 128 (rows of H) x 256 (cols of H) per CTA, UMMA 128x256x16, cta_group::1
//   pipeline  : 4 smem stages of 64 tokens (A 16 KB + B 32 KB, 128B-swizzled TMA boxes of 64 cols x 64 tokens)
//   TMEM      : 2 accumulators x 256 columns (all 512) so the epilogue of item i overlaps the MMAs of i+1
//   epilogue  : tcgen05.ld 32 columns -> swizzled smem -> TMA reduce-add (fp32 add in L2) into H;
//               this is the reference's `H +=`, and it lets several token chunks of one tile be
//               in flight on different SMs with no ordering between them
//   schedule  : persistent, one CTA per SM; work item = (group of tiles, token chunk, tile): see TileSched;
//               only tiles that touch the upper triangle are computed (tq_symmetrize mirrors later)
//   roles     : warp 0 TMA producer, warp 1 MMA issuer (one thread), warp 2 TMEM allocator,
//               warps 4-7 epilogue (TMEM lane quarter = warp % 4)
#include "tc_common.cuh"

#include <stdlib.h>

namespace tq {

constexpr int HT_BM = 128, HT_BN = 256, HT_BK = 64, HT_STAGES = 4, HT_UMMA_K = 16;
constexpr int HT_A_BYTES = HT_BM * HT_BK * 2;              // 16384
constexpr int HT_B_BYTES = HT_BN * HT_BK * 2;              // 32768
constexpr int HT_STAGE_BYTES = HT_A_BYTES + HT_B_BYTES;    // 49152
constexpr int HT_BOX_BYTES = 64 * HT_BK * 2;               // one TMA box: 64 columns x 64 tokens = 8192
constexpr int HT_EPI_COLS = 32;
constexpr int HT_EPI_BYTES = HT_BM * HT_EPI_COLS * 4;      // 16384
constexpr int HT_EPI_STAGES = 2;
constexpr int HT_THREADS = 256;
constexpr int HT_SMEM = HT_STAGES * HT_STAGE_BYTES + HT_EPI_STAGES * HT_EPI_BYTES + 256 + 1024;

// smem matrix descriptor for an MN-major, 128B-swizzled operand slab made of [64 cols x 64 tokens] TMA boxes:
//   8 token rows x 128 B = one 1024 B swizzle atom; next 8 tokens at +1024 B (SBO); next 64 columns at +8192 B (LBO)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(HT_BOX_BYTES >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
           (1ull << 46) | (2ull << 61);
}

// Work schedule.  Tiles (128 rows x 256 cols of H) that touch the upper triangle are bundled into square GROUPS of
// 28 x 14 tiles (3584 x 3584 of H, <= 49 MB): a group's tiles stay L2-resident while ALL token chunks are streamed
// through it, so the per-chunk TMA reduce-adds hit L2 instead of DRAM (for m > ~4600 the whole H does not fit and a
// plain chunk-major sweep turns every reduce into a DRAM read-modify-write).  Inside a group: chunk-major, tile-minor,
// so the chunk's rows of X for the group's columns (<= 29 MB) are fetched from DRAM once per group.
constexpr int HT_GI = 28, HT_GJ = 14, HT_MAX_GROUPS = 136;

struct TileSched {
    int nbi, nbj, gdim, ngroups, num_chunks;
    int gj_tiles;                         // group width in 256-column tiles: HT_GJ, or the whole matrix when H (upper
                                          // triangle) plus a token chunk fits L2 (m <= 4608): X is then read from DRAM once
    int pair;                             // 1: a row index counts PAIRS of 128-row tiles (256 rows): the two CTAs of a cluster
                                          // take the two halves and share the 256-column operand by TMA multicast
    int round_ctas, items_per_cta;        // CTA c works on items (c / round_ctas) * round_ctas * items_per_cta + c % round_ctas
                                          // + j * round_ctas, j < items_per_cta: inside a round the CTAs interleave over
                                          // consecutive items exactly like a persistent grid would (same L2 locality), but a
                                          // CTA retires after ~0.15 ms, so the block scheduler can hand its SM to a
                                          // higher-priority stream (the dependent-kernel chains of linears whose Hessian is
                                          // already complete) instead of holding every SM for the whole launch
    int tile_prefix[HT_MAX_GROUPS + 1];   // tiles in groups [0, g)

    __host__ __device__ int tiles_in_col(int bj, int bi0, int bi1) const {   // valid bi in [bi0, bi1): bi <= 2*bj + 1 (pairs: bi <= bj)
        const int lim = pair ? bj + 1 : 2 * bj + 2;
        const int hi = (lim < bi1) ? lim : bi1;
        return hi > bi0 ? hi - bi0 : 0;
    }
    __host__ __device__ void group_rect(int g, int& bi0, int& bi1, int& bj0, int& bj1) const {
        // groups enumerated over the upper triangle of the gdim x gdim group grid: (gi <= gj), gj-major
        int gj = 0, rem = g;
        while (rem > gj) { rem -= gj + 1; ++gj; }
        const int gi = rem;
        const int GI = pair ? gj_tiles : 2 * gj_tiles;
        bi0 = gi * GI; bi1 = (bi0 + GI < nbi) ? bi0 + GI : nbi;
        bj0 = gj * gj_tiles; bj1 = (bj0 + gj_tiles < nbj) ? bj0 + gj_tiles : nbj;
    }
    __host__ __device__ int group_tiles(int g) const {
        int bi0, bi1, bj0, bj1, n = 0;
        group_rect(g, bi0, bi1, bj0, bj1);
        for (int bj = bj0; bj < bj1; ++bj) n += tiles_in_col(bj, bi0, bi1);
        return n;
    }
    __device__ __forceinline__ long long total_items() const { return (long long)tile_prefix[ngroups] * num_chunks; }
    // item -> (chunk, bi, bj)
    __device__ __forceinline__ void decode(long long it, int& chunk, int& bi, int& bj) const {
        int g = 0;
        while ((long long)tile_prefix[g + 1] * num_chunks <= it) ++g;
        const int tg = tile_prefix[g + 1] - tile_prefix[g];
        const long long local = it - (long long)tile_prefix[g] * num_chunks;
        chunk = (int)(local / tg);
        int t = (int)(local - (long long)chunk * tg);
        int bi0, bi1, bj0, bj1;
        group_rect(g, bi0, bi1, bj0, bj1);
        bj = bj0;
        for (;; ++bj) {
            const int cnt = tiles_in_col(bj, bi0, bi1);
            if (t < cnt) break;
            t -= cnt;
        }
        bi = bi0 + t;
    }
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load delivered to the same shared-memory offset of every CTA in `mask`; each destination's mbarrier (same offset)
// receives the complete_tx
__device__ __forceinline__ void tma_load_2d_multicast(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                                      uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
// MMA completion -> arrive on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_multicast(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask)
                 : "memory");
}

// CLUSTER = 2: the two CTAs of a cluster work on the two 128-row halves of a 256 x 256 tile of H.  Both need the same
// 256-column operand: each loads HALF of it and multicasts the boxes into both shared memories, so a CTA pulls 32 KB
// per 64-token step through L2 -> SM instead of 48 KB (the kernel's measured limiter: 13.8 TB/s of L2 reads at 70 %
// tensor-pipe activity).  A stage may be refilled only when BOTH CTAs have consumed it: the MMA thread's commit arrives
// on the `empty` barrier of both CTAs (count 2).  The MMAs themselves stay cta_group::1 (one accumulator per CTA).
template <int CLUSTER>
__global__ void __launch_bounds__(HT_THREADS, 1)
hessian_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_h,
                  int Nt, int kc, const __grid_constant__ TileSched sched, uint32_t idesc) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_epi = s_base + HT_STAGES * HT_STAGE_BYTES;
    const uint32_t s_bar = s_epi + HT_EPI_STAGES * HT_EPI_BYTES;
    auto full_bar = [&](int s) { return s_bar + 8 * s; };
    auto empty_bar = [&](int s) { return s_bar + 8 * (HT_STAGES + s); };
    auto tfull_bar = [&](int a) { return s_bar + 8 * (2 * HT_STAGES + a); };
    auto tempty_bar = [&](int a) { return s_bar + 8 * (2 * HT_STAGES + 2 + a); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + HT_STAGES * HT_STAGE_BYTES + HT_EPI_STAGES * HT_EPI_BYTES + 8 * (2 * HT_STAGES + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_h) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < HT_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), CLUSTER); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 128); }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc_512(smem_u32(tmem_slot));
    }
    tc_fence_before();
    __syncthreads();
    if (CLUSTER > 1) cluster_sync_all();             // the peer's barriers exist before anything is multicast to them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int rank = (CLUSTER > 1) ? (int)cluster_ctarank() : 0;
    const int unit = blockIdx.x / CLUSTER;           // the scheduling unit: a CTA, or a cluster of two

    const long long total_items = sched.total_items();
    const long long item0 = (long long)(unit / sched.round_ctas) * sched.round_ctas * sched.items_per_cta +
                            unit % sched.round_ctas;
    const int item_stride = sched.round_ctas, my_items = sched.items_per_cta;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer =====
        int stage = 0, phase = 0;
        for (int ji = 0; ji < my_items; ++ji) {
            const long long it = item0 + (long long)ji * item_stride;
            if (it >= total_items) break;
            int chunk, bi, bj;
            sched.decode(it, chunk, bi, bj);
            const int t0 = chunk * kc;
            const int t1 = min(Nt, t0 + kc);
            const int nkb = (t1 - t0 + HT_BK - 1) / HT_BK;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(empty_bar(stage), phase ^ 1);
                const uint32_t sa = s_base + stage * HT_STAGE_BYTES;
                const uint32_t sb = sa + HT_A_BYTES;
                mbar_expect_tx(full_bar(stage), HT_STAGE_BYTES);     // own A + both halves of B (one arrives from the peer)
                const int tok = t0 + kb * HT_BK;
                const int bia = (CLUSTER > 1) ? 2 * bi + rank : bi;
#pragma unroll
                for (int c = 0; c < HT_BM / 64; ++c) tma_load_2d(sa + c * HT_BOX_BYTES, &map_x, full_bar(stage), bia * HT_BM + c * 64, tok);
                if (CLUSTER > 1) {
#pragma unroll
                    for (int c = 0; c < HT_BN / 128; ++c) {
                        const int box = rank * (HT_BN / 128) + c;
                        tma_load_2d_multicast(sb + box * HT_BOX_BYTES, &map_x, full_bar(stage), bj * HT_BN + box * 64, tok, 0x3);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < HT_BN / 64; ++c) tma_load_2d(sb + c * HT_BOX_BYTES, &map_x, full_bar(stage), bj * HT_BN + c * 64, tok);
                }
                if (++stage == HT_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer =====
        int stage = 0, phase = 0, n_item = 0;
        for (int ji = 0; ji < my_items; ++ji, ++n_item) {
            const long long it = item0 + (long long)ji * item_stride;
            if (it >= total_items) break;
            int chunk, bi, bj;
            sched.decode(it, chunk, bi, bj);
            const int t0 = chunk * kc;
            const int t1 = min(Nt, t0 + kc);
            const int nkb = (t1 - t0 + HT_BK - 1) / HT_BK;
            const int acc = n_item & 1;
            mbar_wait(tempty_bar(acc), ((n_item >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * HT_BN;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                const uint32_t sa = s_base + stage * HT_STAGE_BYTES;
                const uint32_t sb = sa + HT_A_BYTES;
#pragma unroll
                for (int k = 0; k < HT_BK / HT_UMMA_K; ++k) {
                    const uint64_t ad = make_desc(sa + k * HT_UMMA_K * 128);
                    const uint64_t bd = make_desc(sb + k * HT_UMMA_K * 128);
                    umma_f16(tmem_d, ad, bd, idesc, (kb | k) != 0);
                }
                if (CLUSTER > 1) umma_commit_multicast(empty_bar(stage), 0x3);   // frees the slot in BOTH CTAs' books
                else umma_commit(empty_bar(stage));     // frees the smem slot once these MMAs have read it
                if (kb == nkb - 1) umma_commit(tfull_bar(acc));
                if (++stage == HT_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> swizzled smem -> TMA reduce-add into H =====
        const int q = warp & 3;                          // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;                   // row of the 128-row tile
        const bool leader = (warp == 4 && lane == 0);
        int n_item = 0, estage = 0;
        for (int ji = 0; ji < my_items; ++ji, ++n_item) {
            const long long it = item0 + (long long)ji * item_stride;
            if (it >= total_items) break;
            int chunk, bi, bj;
            sched.decode(it, chunk, bi, bj);
            const int acc = n_item & 1;
            mbar_wait(tfull_bar(acc), (n_item >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * HT_BN;
            for (int cg = 0; cg < HT_BN / HT_EPI_COLS; ++cg) {
                uint32_t v[32];
                tmem_ld_32x32(taddr + cg * HT_EPI_COLS, v);
                tmem_ld_wait();
                if (cg == HT_BN / HT_EPI_COLS - 1) {     // accumulator fully read: hand it back to the MMA warp
                    tc_fence_before();
                    mbar_arrive(tempty_bar(acc));
                }
                // staging buffer `estage` must have been read by the TMA store issued two groups ago
                if (leader) bulk_wait_read<HT_EPI_STAGES - 1>();
                named_bar_sync(1, 128);
                const uint32_t sdst = s_epi + estage * HT_EPI_BYTES + row * 128;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint32_t a = sdst + ((c ^ (row & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v[4 * c]), "r"(v[4 * c + 1]),
                                 "r"(v[4 * c + 2]), "r"(v[4 * c + 3])
                                 : "memory");
                }
                fence_proxy_async_smem();
                named_bar_sync(1, 128);
                if (leader) {
                    const int bia = (CLUSTER > 1) ? 2 * bi + rank : bi;
                    tma_reduce_add_2d(&map_h, s_epi + estage * HT_EPI_BYTES, bj * HT_BN + cg * HT_EPI_COLS, bia * HT_BM);
                    bulk_commit();
                }
                estage ^= 1;
            }
        }
        if (leader) bulk_wait_read<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tmem_dealloc_512(tmem_base);
    }
    if (CLUSTER > 1) cluster_sync_all();             // neither CTA leaves while the other may still signal its barriers
}


// ---- 2-SM MMA form (tcgen05 cta_group::2) ---------------------------------------------------------------------------
// The two CTAs of a cluster issue ONE 256 x 256 x 16 MMA per step between them: CTA r supplies rows [128 r, 128 r + 128)
// of the A operand and columns [128 r, 128 r + 128) of the B operand from its own shared memory, and receives rows
// [128 r, ...) of the accumulator in its own TMEM.  Per 64-token step a CTA stages 16 KB of A + 16 KB of B (was 16 + 32):
// six pipeline stages instead of four in the same shared memory, half the shared-memory operand reads per flop.
//   producer (warp 0, both CTAs)  TMA `.cta_group::2` loads into the CTA's own stage; the bytes are counted on the LEADER's
//                                 `full` barrier (peer bit cleared), which the leader arms with both CTAs' 64 KB
//   MMA (warp 1, leader only)     tcgen05.mma.cta_group::2; tcgen05.commit.cta_group::2 ... multicast frees the stage in
//                                 both CTAs (`empty`) and publishes the accumulator to both epilogues (`tfull`)
//   epilogue (warps 4-7, both)    as the single-CTA kernel on the CTA's own 128 rows; every thread arrives on the
//                                 LEADER's `tempty` (count 256) when it has read its part of the accumulator
constexpr int HP_STAGES = 6;
constexpr int HP_STAGE_BYTES = HT_A_BYTES + HT_A_BYTES;            // A 128 x 64 + B half 128 x 64 (bf16 / fp16)
constexpr int HP_SMEM = HP_STAGES * HP_STAGE_BYTES + HT_EPI_STAGES * HT_EPI_BYTES + 256 + 1024;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;                          // shared::cluster address of the same offset in CTA 0

__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma2_commit_multicast(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}

__global__ void __launch_bounds__(HT_THREADS, 1)
hessian_tc_pair_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_h,
                       int Nt, int kc, const __grid_constant__ TileSched sched, uint32_t idesc) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_epi = s_base + HP_STAGES * HP_STAGE_BYTES;
    const uint32_t s_bar = s_epi + HT_EPI_STAGES * HT_EPI_BYTES;
    auto full_bar = [&](int s) { return s_bar + 8 * s; };
    auto empty_bar = [&](int s) { return s_bar + 8 * (HP_STAGES + s); };
    auto tfull_bar = [&](int a) { return s_bar + 8 * (2 * HP_STAGES + a); };
    auto tempty_bar = [&](int a) { return s_bar + 8 * (2 * HP_STAGES + 2 + a); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + HP_STAGES * HP_STAGE_BYTES + HT_EPI_STAGES * HT_EPI_BYTES + 8 * (2 * HP_STAGES + 4));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_h) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < HP_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 2 * 128); }
        fence_barrier_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int unit = blockIdx.x >> 1;

    const long long total_items = sched.total_items();
    const long long item0 = (long long)(unit / sched.round_ctas) * sched.round_ctas * sched.items_per_cta +
                            unit % sched.round_ctas;
    const int item_stride = sched.round_ctas, my_items = sched.items_per_cta;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer (both CTAs) =====
        int stage = 0, phase = 0;
        for (int ji = 0; ji < my_items; ++ji) {
            const long long it = item0 + (long long)ji * item_stride;
            if (it >= total_items) break;
            int chunk, bi, bj;
            sched.decode(it, chunk, bi, bj);
            const int t0 = chunk * kc;
            const int t1 = min(Nt, t0 + kc);
            const int nkb = (t1 - t0 + HT_BK - 1) / HT_BK;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(empty_bar(stage), phase ^ 1);
                const uint32_t sa = s_base + stage * HP_STAGE_BYTES;
                const uint32_t sb = sa + HT_A_BYTES;
                const uint32_t lbar = full_bar(stage) & PEER_MASK;
                if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * HP_STAGE_BYTES);   // this CTA's and the peer's bytes
                const int tok = t0 + kb * HT_BK;
#pragma unroll
                for (int c = 0; c < 2; ++c) tma_load_2d_2sm(sa + c * HT_BOX_BYTES, &map_x, lbar, (2 * bi + rank) * HT_BM + c * 64, tok);
#pragma unroll
                for (int c = 0; c < 2; ++c) tma_load_2d_2sm(sb + c * HT_BOX_BYTES, &map_x, lbar, bj * HT_BN + rank * 128 + c * 64, tok);
                if (++stage == HP_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && lane == 0 && rank == 0) {
        // ===== MMA issuer (leader CTA only) =====
        int stage = 0, phase = 0, n_item = 0;
        for (int ji = 0; ji < my_items; ++ji, ++n_item) {
            const long long it = item0 + (long long)ji * item_stride;
            if (it >= total_items) break;
            int chunk, bi, bj;
            sched.decode(it, chunk, bi, bj);
            const int t0 = chunk * kc;
            const int t1 = min(Nt, t0 + kc);
            const int nkb = (t1 - t0 + HT_BK - 1) / HT_BK;
            const int acc = n_item & 1;
            mbar_wait(tempty_bar(acc), ((n_item >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * HT_BN;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                const uint32_t sa = s_base + stage * HP_STAGE_BYTES;
                const uint32_t sb = sa + HT_A_BYTES;
#pragma unroll
                for (int k = 0; k < HT_BK / HT_UMMA_K; ++k) {
                    const uint64_t ad = make_desc(sa + k * HT_UMMA_K * 128);
                    const uint64_t bd = make_desc(sb + k * HT_UMMA_K * 128);
                    umma2_f16(tmem_d, ad, bd, idesc, (kb | k) != 0);
                }
                umma2_commit_multicast(empty_bar(stage), 0x3);
                if (kb == nkb - 1) umma2_commit_multicast(tfull_bar(acc), 0x3);
                if (++stage == HP_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue (both CTAs): TMEM -> registers -> swizzled smem -> TMA reduce-add into H =====
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const bool leader = (warp == 4 && lane == 0);
        int n_item = 0, estage = 0;
        for (int ji = 0; ji < my_items; ++ji, ++n_item) {
            const long long it = item0 + (long long)ji * item_stride;
            if (it >= total_items) break;
            int chunk, bi, bj;
            sched.decode(it, chunk, bi, bj);
            const int acc = n_item & 1;
            mbar_wait(tfull_bar(acc), (n_item >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * HT_BN;
            for (int cg = 0; cg < HT_BN / HT_EPI_COLS; ++cg) {
                uint32_t v[32];
                tmem_ld_32x32(taddr + cg * HT_EPI_COLS, v);
                tmem_ld_wait();
                if (cg == HT_BN / HT_EPI_COLS - 1) {     // accumulator fully read: tell the leader's MMA thread
                    tc_fence_before();
                    mbar_arrive_cluster(tempty_bar(acc) & PEER_MASK);
                }
                if (leader) bulk_wait_read<HT_EPI_STAGES - 1>();
                named_bar_sync(1, 128);
                const uint32_t sdst = s_epi + estage * HT_EPI_BYTES + row * 128;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint32_t a = sdst + ((c ^ (row & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v[4 * c]), "r"(v[4 * c + 1]),
                                 "r"(v[4 * c + 2]), "r"(v[4 * c + 3])
                                 : "memory");
                }
                fence_proxy_async_smem();
                named_bar_sync(1, 128);
                if (leader) {
                    tma_reduce_add_2d(&map_h, s_epi + estage * HT_EPI_BYTES, bj * HT_BN + cg * HT_EPI_COLS, (2 * bi + rank) * HT_BM);
                    bulk_commit();
                }
                estage ^= 1;
            }
        }
        if (leader) bulk_wait_read<0>();
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                              // nobody frees TMEM or leaves while the peer may still use it
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// ---- host side ---------------------------------------------------------------------------------
PFN_tmapEncodeTiled tmap_encode_fn() {
    static PFN_tmapEncodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
    }
    return fn;
}

int make_tmap_2d(CUtensorMap* map, CUtensorMapDataType dt, int elem_bytes, const void* base, int64_t rows, int64_t cols,
                 int64_t ld, int box_rows, int box_cols, const char* what) {
    PFN_tmapEncodeTiled enc = tmap_encode_fn();
    if (!enc) {
        set_error("%s: cuTensorMapEncodeTiled not available from the driver", what);
        return TQ_E_UNSUPPORTED;
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * (cuuint64_t)elem_bytes};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("%s: cuTensorMapEncodeTiled failed with CUresult %d (base %p rows %lld cols %lld ld %lld)", what, (int)r,
                  base, (long long)rows, (long long)cols, (long long)ld);
        return TQ_E_BADARG;
    }
    return 0;
}

int hessian_accum_tcgen05(float* H, int64_t ldh, const void* X, int64_t Nt, int64_t m, int64_t ldx, int dtype,
                          cudaStream_t st) {
    TQ_CHECK_ARG(Nt < (1ll << 31) && m < (1ll << 31), "tq_hessian_accum: sizes exceed the TMA coordinate range");
    CUtensorMap map_x, map_h;
    int rc;
    if ((rc = make_tmap_2d(&map_x, dtype == TQ_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, X,
                           Nt, m, ldx, HT_BK, 64, "tq_hessian_accum(X)")))
        return rc;
    if ((rc = make_tmap_2d(&map_h, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, H, m, m, ldh, HT_BM, HT_EPI_COLS, "tq_hessian_accum(H)")))
        return rc;

    // TQ_HESS_CLUSTER=1 selects the single-CTA kernel (48 KB of operands per step and CTA instead of 32 KB)
    // TQ_HESS_CLUSTER: 1 = single-CTA kernel, 2 = cluster of two with TMA multicast of the shared operand,
    // 3 = cluster of two issuing one 256 x 256 MMA between them (tcgen05 cta_group::2; default).  Measured at Nt = 262 144,
    // m = 4096 / 11008, timed alone: 1 094 / 1 107 -> 1 191 / 1 151 -> 1 266 / 1 201 TFLOP/s useful; inside the 7B bench
    // 1 162 -> 1 259 TFLOP/s (0.83 -> 0.89 of the sustained cuBLAS peak) for variants 2 -> 3.
    static const int variant = []() { const char* e = getenv("TQ_HESS_CLUSTER"); const int v = e ? atoi(e) : 3; return (v >= 1 && v <= 3) ? v : 3; }();
    const int cluster = (variant == 1) ? 1 : 2;
    TileSched sched;
    sched.pair = (cluster == 2) ? 1 : 0;
    sched.nbi = (int)ceil_div(ceil_div(m, HT_BM), cluster);
    sched.nbj = (int)ceil_div(m, HT_BN);
    static const int one_group_max = []() { const char* e = getenv("TQ_HESS_ONE_GROUP_MAX_M"); return e ? atoi(e) : 4608; }();
    sched.gj_tiles = (m <= one_group_max) ? sched.nbj : HT_GJ;
    sched.gdim = (int)ceil_div(sched.nbj, sched.gj_tiles);
    sched.ngroups = sched.gdim * (sched.gdim + 1) / 2;
    TQ_CHECK_ARG(sched.ngroups <= HT_MAX_GROUPS, "tq_hessian_accum: m = %lld is beyond the scheduler's group table", (long long)m);
    sched.tile_prefix[0] = 0;
    for (int g = 0; g < sched.ngroups; ++g) sched.tile_prefix[g + 1] = sched.tile_prefix[g] + sched.group_tiles(g);
    const int64_t tiles = sched.tile_prefix[sched.ngroups];
    const int64_t tiles_per_group = tiles / sched.ngroups > 0 ? tiles / sched.ngroups : 1;

    // token chunk = one TMEM accumulation: small enough that the group's slab of X (chunk x <= 7168 cols x 2 B) plus the
    // group's H tiles stay L2-resident, capped at 2048 tokens (see below), and giving >= ~6 work items per SM
    // TQ_HESS_SPARE_SMS: SMs this persistent kernel leaves to the chains of other linears (inverse / sweep kernels on other
    // streams) that run while the remaining Hessians of the layer are still being accumulated
    static const int spare = []() { const char* e = getenv("TQ_HESS_SPARE_SMS"); return e ? atoi(e) : 0; }();
    const int sms = ((sm_count() - spare > 16) ? sm_count() - spare : sm_count()) / cluster;   // scheduling units per round
    int64_t kc_l2 = (40ll << 20) / (2 * (m < 7168 ? m : 7168));
    kc_l2 = (kc_l2 / HT_BK) * HT_BK;
    if (kc_l2 < 256) kc_l2 = 256;
    static const int kc_cap = []() { const char* e = getenv("TQ_HESS_KC_MAX"); return e ? atoi(e) : 2048; }();
    if (kc_l2 > kc_cap) kc_l2 = kc_cap;   // also bounds the length of one TMEM accumulation: the tensor core's fp32 add truncates, and
                                      // a sum of squares over L tokens picks up a relative bias of ~L * 2^-25 (measured 2.8e-5 at
                                      // L = 4864, 9e-6 at 1792); 2048 is the reference's own per-add_batch granularity
    const int64_t want_chunks = ceil_div((int64_t)6 * sms, tiles_per_group);
    int64_t kc = ceil_div(ceil_div(Nt, want_chunks), HT_BK) * HT_BK;
    if (kc < 256) kc = 256;
    if (kc > kc_l2) kc = kc_l2;
    const int64_t num_chunks = ceil_div(Nt, kc);
    sched.num_chunks = (int)num_chunks;
    const int64_t items = num_chunks * tiles;

    const uint32_t fmt = (dtype == TQ_F16) ? 0u : 1u;
    // tcgen05 instruction descriptor (kind::f16): D=f32, A/B format, A and B MN-major, N=256, M=128
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) |
                           ((uint32_t)(HT_BN >> 3) << 17) | ((uint32_t)(HT_BM >> 4) << 24);

    static std::atomic<unsigned long long> attr_mask{0};
    int dev;
    if (dyn_smem_pending(attr_mask, dev)) {
        TQ_CUDA(cudaFuncSetAttribute(hessian_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, HT_SMEM));
        TQ_CUDA(cudaFuncSetAttribute(hessian_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, HT_SMEM));
        TQ_CUDA(cudaFuncSetAttribute(hessian_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HP_SMEM));
        dyn_smem_done(attr_mask, dev);
    }
    // TQ_HESS_ITEMS_PER_CTA: work items (one <= 2048-token chunk of one 128 x 256 tile, ~9 us) a CTA processes before it
    // retires; 0 = persistent (one CTA per SM for the whole launch)
    static const int ipc_env = []() { const char* e = getenv("TQ_HESS_ITEMS_PER_CTA"); return e ? atoi(e) : 16; }();
    sched.round_ctas = (int)((items < sms) ? items : sms);
    const int64_t per_cta_persistent = ceil_div(items, sched.round_ctas);
    sched.items_per_cta = (int)((ipc_env > 0 && ipc_env < per_cta_persistent) ? ipc_env : per_cta_persistent);
    const int grid = (int)(ceil_div(items, (int64_t)sched.round_ctas * sched.items_per_cta) * sched.round_ctas) * cluster;
    if (cluster == 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(HT_THREADS);
        cfg.dynamicSmemBytes = (variant == 3) ? HP_SMEM : HT_SMEM;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (variant == 3) {
            // instruction descriptor of the pair MMA: as above with M = 256 (the two CTAs' 128 rows)
            const uint32_t idesc2 = (idesc & ~(0x1Fu << 24)) | ((uint32_t)(256 >> 4) << 24);
            TQ_CUDA(cudaLaunchKernelEx(&cfg, hessian_tc_pair_kernel, map_x, map_h, (int)Nt, (int)kc, sched, idesc2));
        } else {
            TQ_CUDA(cudaLaunchKernelEx(&cfg, hessian_tc_kernel<2>, map_x, map_h, (int)Nt, (int)kc, sched, idesc));
        }
    } else {
        hessian_tc_kernel<1><<<grid, HT_THREADS, HT_SMEM, st>>>(map_x, map_h, (int)Nt, (int)kc, sched, idesc);
    }
    TQ_LAUNCH_CHECK("hessian_tc_kernel");
    return 0;
}

}  // namespace tq
