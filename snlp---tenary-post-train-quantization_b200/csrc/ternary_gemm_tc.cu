// SURVEY 8f N2, many-token path: y = x Wq' (+ bias) for a TernaryLinear (model.py:75-110) as ONE tcgen05 GEMM that
// dequantises the 2-bit codes on the fly -- the dense weight never exists in HBM (the reference materialises it on
// every forward, model.py:87, :97-110).
//
//   D[128 weight rows, 256 tokens] (fp32, TMEM) += A[128 x 64] * B[256 x 64]',  kind::f16, UMMA 128x256x16
//   (tile widths of 128 / 256 / 512 tokens; a 512-token tile issues two 256-wide MMAs per K step on one A slab)
//   A = dequantised weights, K-major, 128B-swizzled, WRITTEN BY SIXTEEN DEQUANT WARPS from the TL2 code words
//       (ternary_linear.cu): value = wtab[row, p/block][code], already rounded to the layer's 16-bit dtype, so the
//       tensor core multiplies exactly the reference's fp16/bf16 weight `alpha * T + mu`;
//   B = activations [tokens, m] in sweep order (gathered by perm beforehand when perm is not the identity,
//       model.py:84), K-major, fetched by TMA (one 256 x 64 box per stage, out-of-range rows/columns zero-filled);
//   pipeline: 4 smem stages x (A 16 KB + B 32 KB), or 2 x (16 + 64 KB) for 512-token tiles; per stage two "full" barriers (TMA bytes; one arrival per
//       dequant warp after every lane's fence.proxy.async) and one "empty" barrier (tcgen05.commit) that both producers wait on;
//   TMEM: 2 accumulators x 256 columns, so the epilogue of one tile overlaps the MMAs of the next (512-token tiles
//       use all 512 columns for one accumulator);
//   epilogue: tcgen05.ld -> (+ bias) -> 16-bit stores y[token, row] with lanes along the rows (contiguous in y);
//   schedule: persistent CTAs; tiles ordered token-tile-major so concurrently running CTAs share the same x slab in L2.
//   roles: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 epilogue, warps 8-23 dequant
//   (sixteen warps: the expansion is ~5 integer instructions per weight and needs the issue slots of all four schedulers).
#include "tc_common.cuh"

#include <stdlib.h>

namespace tq {

constexpr int TG_BM = 128, TG_BK = 64, TG_UMMA_K = 16;
constexpr int TG_A_BYTES = TG_BM * TG_BK * 2;                 // 16384
constexpr int TG_MAX_STAGES = 4;
// token-tile width -> pipeline geometry: <= 256 tokens: 4 stages x (A 16 KB + B 32 KB) = 192 KB, two TMEM accumulators;
// 512 tokens: 2 stages x (A 16 KB + B 64 KB) = 160 KB, ONE 512-column accumulator (two 256-wide MMAs per K step share
// the dequantised A slab, halving the expansion work per MMA; the epilogue is then not overlapped with the next tile)
__host__ __device__ constexpr int tg_stages(int bn) { return bn == 512 ? 2 : 4; }
__host__ __device__ constexpr int tg_stage_bytes(int bn) { return TG_A_BYTES + (bn == 512 ? 512 : 256) * TG_BK * 2; }
__host__ __device__ constexpr int tg_acc_bufs(int bn) { return bn == 512 ? 1 : 2; }
constexpr int TG_THREADS = 768;
constexpr int TG_DEQ_THREADS = 512;
constexpr int TG_DEQ_WARPS = TG_DEQ_THREADS / 32;
constexpr int TG_SMEM = 4 * tg_stage_bytes(256) + 256 + 1024;          // >= 2 * tg_stage_bytes(512)

struct TgProblem {
    const uint32_t* codes;
    int64_t wpr;
    const float4* wtab;
    int n, m, nb, block, M;
    int wpb_shift;                   // log2 of the code words per scale block (block / 16) when a power of two, else -1
    uint32_t wpb_magic;              // ceil(2^32 / (block / 16)): multiply-high division otherwise
    int row_tiles, tok_tiles, ksteps;
    const float* bias;
    void* y;
    int64_t ldy;
};

// K-major, 128B-swizzled slab [rows x 64 halfs]: 8 rows x 128 B per swizzle atom, next 8 rows at +1024 B
__device__ __forceinline__ uint64_t tg_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

template <typename HT> __device__ __forceinline__ uint32_t tg_bits(float v);
template <> __device__ __forceinline__ uint32_t tg_bits<__half>(float v) { return (uint32_t)__half_as_ushort(__float2half_rn(v)); }
template <> __device__ __forceinline__ uint32_t tg_bits<__nv_bfloat16>(float v) { return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v)); }

template <typename HT> __device__ __forceinline__ HT tg_out(float v);
template <> __device__ __forceinline__ __half tg_out<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 tg_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// The 16 positions of one code word (bit j: +1, bit 16+j: -1) -> 8 packed pairs of 16-bit weights (position 2i in the
// low half of q[i]); h0/h1/h2 = bit patterns of w-, w0, w+.  Byte-permute lookup instead of per-position selects: the
// six bytes of (h0, h1, h2) sit in a register pair, and each output byte is picked by a selector nibble -- position j
// needs the selector byte 0x10 + 0x22 u_j with u_j = 1 + P_j - N_j.  Four positions at a time: a nibble of a plane is
// spread to one bit per byte by a multiply (copies at shifts 0/7/14/21 never overlap) and a mask, then
// S = 0x32323232 + 0x22 spread(P) - 0x22 spread(N) holds the four selector bytes (no borrow: a position is never in
// both planes) and two PRMTs emit the two pairs.  ~3.3 integer instructions per weight instead of ~4.5.
__device__ __forceinline__ void tg_expand_word(uint32_t word, uint32_t h0, uint32_t h1, uint32_t h2, uint32_t (&q)[8]) {
    const uint32_t A = h0 | (h1 << 16), B = h2;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const uint32_t pb = (((word >> (4 * g)) & 0xFu) * 0x00204081u) & 0x01010101u;
        const uint32_t nb = (((word >> (16 + 4 * g)) & 0xFu) * 0x00204081u) & 0x01010101u;
        const uint32_t S = 0x32323232u + 0x22u * pb - 0x22u * nb;
        q[2 * g] = __byte_perm(A, B, S);
        q[2 * g + 1] = __byte_perm(A, B, S >> 16);
    }
}

// TG_BN = tokens per tile (128 / 256 / 512), chosen by the host to minimise the number of tile waves
template <typename HT, int TG_BN>
__global__ void __launch_bounds__(TG_THREADS, 1)
tl_gemm_tc_kernel(const __grid_constant__ CUtensorMap map_x, const TgProblem p, uint32_t idesc) {
    constexpr int TG_STAGES = tg_stages(TG_BN);
    constexpr int TG_STAGE_BYTES = tg_stage_bytes(TG_BN);
    constexpr int NBUF = tg_acc_bufs(TG_BN);                  // accumulator buffers in TMEM
    constexpr int BOX = TG_BN < 256 ? TG_BN : 256;            // tokens per TMA box and per MMA
    constexpr int NBOX = TG_BN / BOX;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_bar = s_base + TG_STAGES * TG_STAGE_BYTES;
    auto fullx_bar = [&](int s) { return s_bar + 8 * s; };
    auto fullw_bar = [&](int s) { return s_bar + 8 * (TG_MAX_STAGES + s); };
    auto empty_bar = [&](int s) { return s_bar + 8 * (2 * TG_MAX_STAGES + s); };
    auto tfull_bar = [&](int a) { return s_bar + 8 * (3 * TG_MAX_STAGES + a); };
    auto tempty_bar = [&](int a) { return s_bar + 8 * (3 * TG_MAX_STAGES + 2 + a); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + TG_STAGES * TG_STAGE_BYTES + 8 * (3 * TG_MAX_STAGES + 4));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < TG_STAGES; ++s) {
            mbar_init(fullx_bar(s), 1);
            mbar_init(fullw_bar(s), TG_DEQ_WARPS);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < NBUF; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 128); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_512(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int tiles = p.row_tiles * p.tok_tiles;
    constexpr int TG_B_BYTES = TG_BN * TG_BK * 2;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer: activations =====
        int stage = 0, phase = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
            const int bj = t / p.row_tiles;
            for (int ks = 0; ks < p.ksteps; ++ks) {
                mbar_wait(empty_bar(stage), phase ^ 1);
                mbar_expect_tx(fullx_bar(stage), TG_B_BYTES);
#pragma unroll
                for (int b = 0; b < NBOX; ++b)
                    tma_load_2d(s_base + stage * TG_STAGE_BYTES + TG_A_BYTES + b * BOX * TG_BK * 2, &map_x, fullx_bar(stage),
                                ks * TG_BK, bj * TG_BN + b * BOX);
                if (++stage == TG_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer =====
        int stage = 0, phase = 0, n_item = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++n_item) {
            const int acc = n_item % NBUF;
            mbar_wait(tempty_bar(acc), ((n_item / NBUF) & 1) ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * TG_BN;
            for (int ks = 0; ks < p.ksteps; ++ks) {
                mbar_wait(fullw_bar(stage), phase);
                mbar_wait(fullx_bar(stage), phase);
                tc_fence_after();
                const uint32_t sa = s_base + stage * TG_STAGE_BYTES;
                const uint32_t sb = sa + TG_A_BYTES;
#pragma unroll
                for (int k = 0; k < TG_BK / TG_UMMA_K; ++k)
#pragma unroll
                    for (int b = 0; b < NBOX; ++b)
                        umma_f16(tmem_d + b * BOX, tg_desc(sa + k * TG_UMMA_K * 2),
                                 tg_desc(sb + b * BOX * TG_BK * 2 + k * TG_UMMA_K * 2), idesc, (ks | k) != 0);
                umma_commit(empty_bar(stage));
                if (ks == p.ksteps - 1) umma_commit(tfull_bar(acc));
                if (++stage == TG_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ===== epilogue: each thread owns one weight row of the tile; lanes along rows = contiguous in y =====
        const int q = warp & 3;
        HT* y = reinterpret_cast<HT*>(p.y);
        int n_item = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++n_item) {
            const int bj = t / p.row_tiles, bi = t - bj * p.row_tiles;
            const int acc = n_item % NBUF;
            const int r = bi * TG_BM + q * 32 + lane;
            const float bv = (p.bias && r < p.n) ? p.bias[r] : 0.f;
            mbar_wait(tfull_bar(acc), (n_item / NBUF) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * TG_BN;
#pragma unroll 1
            for (int cg = 0; cg < TG_BN / 32; ++cg) {
                uint32_t v[32];
                tmem_ld_32x32(taddr + cg * 32, v);
                tmem_ld_wait();
                if (cg == TG_BN / 32 - 1) {
                    tc_fence_before();
                    mbar_arrive(tempty_bar(acc));
                }
                const int tok0 = bj * TG_BN + cg * 32;
                if (r < p.n) {
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        if (tok0 + c < p.M) y[(int64_t)(tok0 + c) * p.ldy + r] = tg_out<HT>(__fadd_rn(__uint_as_float(v[c]), bv));
                }
            }
        }
    } else if (warp >= 8) {
        // ===== dequant producers: thread = (row of the tile, quarter of the 64-wide K slab = 1 code word) =====
        // The code word and its weight-table entry are fetched TWO slabs ahead of the one being expanded: with one
        // CTA per SM nothing else hides the L2 latency of these loads.
        const int dt = threadIdx.x - 256;
        const int row = dt >> 2, h = dt & 3;
        const int words = (p.m + 15) >> 4;
        int stage = 0, phase = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
            const int bj = t / p.row_tiles, bi = t - bj * p.row_tiles;
            const int r = bi * TG_BM + row;
            const bool live = r < p.n;
            const uint32_t* crow = p.codes + (int64_t)(live ? r : 0) * p.wpr;
            const float4* trow = p.wtab + (int64_t)(live ? r : 0) * p.nb;
            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
            auto fetch = [&](int ks, uint32_t& a, float4& ta) {
                const int w = ks * 4 + h;
                a = 0u; ta = zero4;                             // out of range: empty planes AND a zero table
                if (live && w < words && ks < p.ksteps) {
                    a = __ldg(crow + w);
                    const int k = p.wpb_shift >= 0 ? (w >> p.wpb_shift) : (int)__umulhi((uint32_t)w, p.wpb_magic);
                    ta = __ldg(trow + min(k, p.nb - 1));
                }
            };
            uint32_t w0, n0;
            float4 ta, na;
            fetch(0, w0, ta);
            fetch(1, n0, na);
            for (int ks = 0; ks < p.ksteps; ++ks) {
                uint32_t f0;
                float4 fa;
                fetch(ks + 2, f0, fa);
                uint32_t qa[8];
                tg_expand_word(w0, tg_bits<HT>(ta.x), tg_bits<HT>(ta.y), tg_bits<HT>(ta.z), qa);
                mbar_wait(empty_bar(stage), phase ^ 1);
                const uint32_t srow = s_base + stage * TG_STAGE_BYTES + row * 128;
                const int c0 = 2 * h;                        // this thread's two 16-byte chunks of the 128-byte row
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + (((c0 + 0) ^ (row & 7)) << 4)),
                             "r"(qa[0]), "r"(qa[1]), "r"(qa[2]), "r"(qa[3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + (((c0 + 1) ^ (row & 7)) << 4)),
                             "r"(qa[4]), "r"(qa[5]), "r"(qa[6]), "r"(qa[7]) : "memory");
                fence_proxy_async_smem();                    // generic-proxy writes -> visible to the tensor core's async proxy
                __syncwarp();
                if (lane == 0) mbar_arrive(fullw_bar(stage)); // one arrival per warp: 512 arrivals on one barrier would serialise
                w0 = n0; ta = na;
                n0 = f0; na = fa;
                if (++stage == TG_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc_512(tmem_base);
}

// xp[t, p] = x[t, perm[p]] (model.py:84), 16-bit elements
__global__ void __launch_bounds__(256)
tl_gather_x_kernel(const uint16_t* __restrict__ x, int64_t ldx, int64_t M, int64_t m, const int32_t* __restrict__ perm,
                   uint16_t* __restrict__ xp) {
    const int64_t total = M * m;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = q / m, pcol = q - t * m;
        xp[q] = x[t * ldx + perm[pcol]];
    }
}

}  // namespace tq

extern "C" int tq_tl_gemm_tc(const uint32_t* codes, int64_t wpr, const float* wtab, int64_t n, int64_t m, int64_t block,
                             const void* x, int xdtype, int64_t ldx, int64_t M, const int32_t* perm, void* xperm_work,
                             const float* bias, void* y, int64_t ldy, void* stream) {
    using namespace tq;
    TQ_CHECK_ARG(codes && wtab && x && y && n > 0 && m > 0 && M >= 0 && wpr >= ceil_div(m, 16) && ldx >= m && ldy >= n,
                 "tq_tl_gemm_tc: bad arguments");
    TQ_CHECK_ARG(n < (1ll << 31) && m < (1ll << 31) && M < (1ll << 31), "tq_tl_gemm_tc: shape too large");
    if (xdtype != TQ_F16 && xdtype != TQ_BF16) {
        set_error("tq_tl_gemm_tc: activations must be f16 or bf16 (the layer dtype); fp32 layers use tq_tl_dequant + a dense GEMM");
        return TQ_E_UNSUPPORTED;
    }
    if (block <= 0 || block % 16 != 0) {
        set_error("tq_tl_gemm_tc: block size %lld is not a positive multiple of 16", (long long)block);
        return TQ_E_UNSUPPORTED;
    }
    TQ_CHECK_ARG((reinterpret_cast<uintptr_t>(wtab) & 15) == 0, "tq_tl_gemm_tc: wtab must be 16-byte aligned");
    TQ_CHECK_ARG(perm == nullptr || xperm_work != nullptr, "tq_tl_gemm_tc: a permuted layer needs the [M, m] gather workspace");
    if (M == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const void* xs = x;
    int64_t lds = ldx;
    if (perm) {
        int64_t g = ceil_div(M * m, 256);
        const int64_t cap = (int64_t)sm_count() * 16;
        if (g > cap) g = cap;
        tl_gather_x_kernel<<<(unsigned)g, 256, 0, st>>>((const uint16_t*)x, ldx, M, m, perm, (uint16_t*)xperm_work);
        TQ_LAUNCH_CHECK("tl_gather_x_kernel");
        xs = xperm_work;
        lds = m;
    }
    if ((reinterpret_cast<uintptr_t>(xs) & 15) != 0 || (lds * 2) % 16 != 0) {
        set_error("tq_tl_gemm_tc: activations need a 16-byte aligned base and row pitch (m %% 8 == 0) for TMA");
        return TQ_E_UNSUPPORTED;
    }
    const int sms = sm_count();
    // The kernel is bound by the expansion of the weight slab, which costs about the same per K step whatever the tile
    // width: time ~ (waves of tiles) x K steps x t(width), with t measured on B200 at 0.93 / 1.0 / 1.5 us for 128 / 256 /
    // 512 tokens (profiles/r01d_tl_bench_prefill.json; the 512-wide tile has a 2-stage pipeline and no spare
    // accumulator).  Take the cheapest width, but never pad a call to more than twice its tokens.
    int bn = 128;
    {
        const double t_step[3] = {0.93, 1.0, 1.5};
        double best = (double)ceil_div(ceil_div(n, TG_BM) * ceil_div(M, 128), sms) * t_step[0];
        int idx = 1;
        for (int cand = 256; cand <= 512; cand *= 2, ++idx) {
            if (M <= cand / 2) break;
            const double cost = (double)ceil_div(ceil_div(n, TG_BM) * ceil_div(M, cand), sms) * t_step[idx];
            if (cost < best) { best = cost; bn = cand; }
        }
    }
    if (const char* e = getenv("TQ_TL_GEMM_BN")) {      // development / test override of the tile width
        const int v = atoi(e);
        if (v == 128 || v == 256 || v == 512) bn = v;
    }
    const int box = bn < 256 ? bn : 256;
    CUtensorMap map_x;
    int rc = make_tmap_2d(&map_x, xdtype == TQ_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, xs,
                          M, m, lds, box, TG_BK, "tq_tl_gemm_tc(x)");
    if (rc) return rc;

    TgProblem p;
    p.codes = codes;
    p.wpr = wpr;
    p.wtab = reinterpret_cast<const float4*>(wtab);
    p.n = (int)n; p.m = (int)m; p.nb = (int)ceil_div(m, block); p.block = (int)block; p.M = (int)M;
    const uint32_t wpb = (uint32_t)(block / 16);
    p.wpb_shift = -1;
    for (int sft = 0; sft < 28; ++sft)
        if ((1u << sft) == wpb) p.wpb_shift = sft;
    p.wpb_magic = wpb > 1 ? (uint32_t)(((1ull << 32) + wpb - 1) / wpb) : 0u;
    p.row_tiles = (int)ceil_div(n, TG_BM);
    p.tok_tiles = (int)ceil_div(M, bn);
    p.ksteps = (int)ceil_div(m, TG_BK);
    p.bias = bias;
    p.y = y;
    p.ldy = ldy;
    const uint32_t fmt = (xdtype == TQ_F16) ? 0u : 1u;
    // tcgen05 instruction descriptor (kind::f16): D = f32, A/B format, both K-major, N = tokens per tile, M = 128
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(box >> 3) << 17) | ((uint32_t)(TG_BM >> 4) << 24);
    const int64_t tiles = (int64_t)p.row_tiles * p.tok_tiles;
    const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
    static std::atomic<unsigned long long> attr_devices{0};
    int dev;
    if (dyn_smem_pending(attr_devices, dev)) {
        TQ_CUDA(cudaFuncSetAttribute(tl_gemm_tc_kernel<__half, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM));
        TQ_CUDA(cudaFuncSetAttribute(tl_gemm_tc_kernel<__half, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM));
        TQ_CUDA(cudaFuncSetAttribute(tl_gemm_tc_kernel<__half, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM));
        TQ_CUDA(cudaFuncSetAttribute(tl_gemm_tc_kernel<__nv_bfloat16, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM));
        TQ_CUDA(cudaFuncSetAttribute(tl_gemm_tc_kernel<__nv_bfloat16, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM));
        TQ_CUDA(cudaFuncSetAttribute(tl_gemm_tc_kernel<__nv_bfloat16, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM));
        dyn_smem_done(attr_devices, dev);
    }
    if (xdtype == TQ_F16) {
        if (bn == 512) tl_gemm_tc_kernel<__half, 512><<<grid, TG_THREADS, TG_SMEM, st>>>(map_x, p, idesc);
        else if (bn == 256) tl_gemm_tc_kernel<__half, 256><<<grid, TG_THREADS, TG_SMEM, st>>>(map_x, p, idesc);
        else tl_gemm_tc_kernel<__half, 128><<<grid, TG_THREADS, TG_SMEM, st>>>(map_x, p, idesc);
    } else {
        if (bn == 512) tl_gemm_tc_kernel<__nv_bfloat16, 512><<<grid, TG_THREADS, TG_SMEM, st>>>(map_x, p, idesc);
        else if (bn == 256) tl_gemm_tc_kernel<__nv_bfloat16, 256><<<grid, TG_THREADS, TG_SMEM, st>>>(map_x, p, idesc);
        else tl_gemm_tc_kernel<__nv_bfloat16, 128><<<grid, TG_THREADS, TG_SMEM, st>>>(map_x, p, idesc);
    }
    TQ_LAUNCH_CHECK("tl_gemm_tc_kernel");
    return 0;
}
