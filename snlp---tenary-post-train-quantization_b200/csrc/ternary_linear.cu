// the consumer of the packed codes -- the reference's TernaryLinear (model.py:17-127) on 2-bit codes.
//
// Layer format ("TL2"), chosen so that one 32-bit word never straddles a scale block:
//   codes  u32 [n, wpr], wpr >= ceil(m/16): word w of row r holds sweep positions 16w .. 16w+15, position p at
//          bits 2(p%16) .. +1, code = T[r, perm[p]] + 1 (utils.py:203 coding); positions >= m hold code 1 (T = 0).
//          Sweep order, not original order: block k of alpha/mu (gptq.py:153-155) then covers the contiguous
//          positions [k*block, (k+1)*block) and the input is gathered once by perm (model.py:84's x[..., perm]).
//   wtab   f32 [n, nb, 4]: the three values a weight of (row, block) can take, rounded like the reference's
//          dequantisation `alpha * T + mu` evaluated in the layer dtype (model.py:106-108):
//          (fl(mu - alpha), mu, fl(alpha + mu), 0) indexed by code.
// With them  y[t, r] = sum_p wtab[r, p/block][code(r, p)] * x[t, perm[p]]  is exactly F.linear(x, Wq) with
// Wq = get_quantized_weight (gptq.py:201-230) up to summation order.  (model.py:84-90 permutes the input AND
// un-permutes W while its T is stored in original positions -- SURVEY Q11 -- so the reference's forward equals this
// only for the identity permutation; tests pin both facts.)
//
// Kernels (CUDA cores; the codes are 0.25 B per weight, so a decode-sized call is bound by reading them once):
//   tl_pack_kernel     int8 T in original positions + perm -> codes                       (HBM: n*m in, n*m/4 out)
//   tl_wtab_kernel     alpha, mu -> wtab
//   tl_gemv_kernel     <= 4 tokens per launch: gathered x staged in shared memory (transposed + padded so that a
//                      warp's 32 words read 32 different banks), one warp per row, 16 codes per lane per load,
//                      the weight picked from wtab by code and multiplied into every token's accumulator
//   tl_expand_kernel   codes -> dense Wq (f32/f16/bf16) or T (int8) in ORIGINAL column positions, for the
//                      many-token path (library GEMM on the dense weight) and for reading T back
#include "common.cuh"

namespace tq {

constexpr int TL_KW = 128;            // words of one row staged per chunk (2048 positions)
constexpr int TL_ROWS_PER_WARP = 2;
constexpr int TL_WARPS = 8;

__global__ void __launch_bounds__(256)
tl_pack_kernel(const int8_t* __restrict__ Torig, int64_t n, int64_t m, const int32_t* __restrict__ perm,
               uint32_t* __restrict__ codes, int64_t wpr) {
    const int64_t total = n * wpr;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = q / wpr, w = q - r * wpr;
        const int8_t* row = Torig + r * m;
        uint32_t word = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int64_t p = w * 16 + j;
            uint32_t c = 1u;
            if (p < m) c = (uint32_t)((int)row[perm ? perm[p] : p] + 1) & 3u;
            word |= c << (2 * j);
        }
        codes[q] = word;
    }
}

template <int WD>
__device__ __forceinline__ float tl_add_rounded(float a, float u) {      // fl_dtype(a + u), a and u exact in dtype
    if constexpr (WD == TQ_F16) return __half2float(__hadd(__float2half_rn(a), __float2half_rn(u)));
    else if constexpr (WD == TQ_BF16) return __bfloat162float(__hadd(__float2bfloat16_rn(a), __float2bfloat16_rn(u)));
    else return __fadd_rn(a, u);
}

template <int WD>
__global__ void __launch_bounds__(256)
tl_wtab_kernel(const float* __restrict__ alpha, const float* __restrict__ mu, int64_t count, float4* __restrict__ wtab) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < count; q += (int64_t)gridDim.x * blockDim.x) {
        float a = alpha[q], u = mu[q];
        if constexpr (WD == TQ_F16) { a = __half2float(__float2half_rn(a)); u = __half2float(__float2half_rn(u)); }
        if constexpr (WD == TQ_BF16) { a = __bfloat162float(__float2bfloat16_rn(a)); u = __bfloat162float(__float2bfloat16_rn(u)); }
        // model.py:108: alpha * T + mu with T in {-1, 0, +1}: the product is exact, the sum rounds once
        wtab[q] = make_float4(tl_add_rounded<WD>(-a, u), tl_add_rounded<WD>(0.f, u), tl_add_rounded<WD>(a, u), 0.f);
    }
}

__device__ __forceinline__ float tl_pick(const float4& wt, uint32_t c) {
    return c == 0u ? wt.x : (c == 1u ? wt.y : wt.z);
}

// xs layout: [16][TL_KW + 1][MT] floats; position (j, wl) = chunk position 16*wl + j.  A warp reads (j, lane + 32 i)
// for fixed j: consecutive lanes are MT floats apart -> conflict-free vector loads.
template <typename XT, int MT>
__global__ void __launch_bounds__(TL_WARPS * 32)
tl_gemv_kernel(const uint32_t* __restrict__ codes, int64_t wpr, const float4* __restrict__ wtab, int n, int m, int nb,
               int block, const XT* __restrict__ x, int64_t ldx, int M, const int32_t* __restrict__ perm,
               const float* __restrict__ bias, float* __restrict__ y, int64_t ldy) {
    __shared__ __align__(16) float xs[16 * (TL_KW + 1) * MT];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row0 = (blockIdx.x * TL_WARPS + warp) * TL_ROWS_PER_WARP;
    const int words = (m + 15) >> 4;                         // words of a row that hold real positions

    float acc[TL_ROWS_PER_WARP][MT];
#pragma unroll
    for (int i = 0; i < TL_ROWS_PER_WARP; ++i)
#pragma unroll
        for (int t = 0; t < MT; ++t) acc[i][t] = 0.f;

    for (int c0 = 0; c0 < m; c0 += TL_KW * 16) {
        const int w0 = c0 >> 4;
        // the chunk's code words first (independent global loads in flight while x is staged)
        uint32_t wd[TL_ROWS_PER_WARP][TL_KW / 32];
#pragma unroll
        for (int i = 0; i < TL_ROWS_PER_WARP; ++i)
#pragma unroll
            for (int s = 0; s < TL_KW / 32; ++s) {
                const int w = w0 + lane + 32 * s;
                const int r = row0 + i;
                wd[i][s] = (r < n && w < words) ? __ldg(codes + (int64_t)r * wpr + w) : 0x55555555u;
            }
        __syncthreads();                                     // previous chunk fully consumed
        for (int idx = threadIdx.x; idx < TL_KW * 16; idx += TL_WARPS * 32) {
            const int p = c0 + idx;
            const int col = p < m ? (perm ? perm[p] : p) : -1;
            float* dst = xs + ((idx & 15) * (TL_KW + 1) + (idx >> 4)) * MT;
#pragma unroll
            for (int t = 0; t < MT; ++t) dst[t] = (col >= 0 && t < M) ? to_f32<XT>(x[(int64_t)t * ldx + col]) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < TL_KW / 32; ++s) {
            const int wl = lane + 32 * s;
            const int w = w0 + wl;
            if (w >= words) continue;
            const int k = min((w * 16) / block, nb - 1);
            float4 wt[TL_ROWS_PER_WARP];
#pragma unroll
            for (int i = 0; i < TL_ROWS_PER_WARP; ++i)
                wt[i] = (row0 + i < n) ? __ldg(wtab + (int64_t)(row0 + i) * nb + k) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float xv[MT];
                const float* src = xs + (j * (TL_KW + 1) + wl) * MT;
                if constexpr (MT == 4) {
                    const float4 v = *reinterpret_cast<const float4*>(src);
                    xv[0] = v.x; xv[1] = v.y; xv[2] = v.z; xv[3] = v.w;
                } else if constexpr (MT == 2) {
                    const float2 v = *reinterpret_cast<const float2*>(src);
                    xv[0] = v.x; xv[1] = v.y;
                } else {
                    xv[0] = src[0];
                }
#pragma unroll
                for (int i = 0; i < TL_ROWS_PER_WARP; ++i) {
                    const float wv = tl_pick(wt[i], (wd[i][s] >> (2 * j)) & 3u);
#pragma unroll
                    for (int t = 0; t < MT; ++t) acc[i][t] = fmaf(wv, xv[t], acc[i][t]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < TL_ROWS_PER_WARP; ++i)
#pragma unroll
        for (int t = 0; t < MT; ++t) {
            const float v = warp_sum(acc[i][t]);
            const int r = row0 + i;
            if (lane == 0 && r < n && t < M) y[(int64_t)t * ldy + r] = bias ? __fadd_rn(v, bias[r]) : v;
        }
}

template <typename OT> __device__ __forceinline__ OT tl_cast(float v);
template <> __device__ __forceinline__ float tl_cast<float>(float v) { return v; }
template <> __device__ __forceinline__ __half tl_cast<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 tl_cast<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ int8_t tl_cast<int8_t>(float v) { return (int8_t)v; }

// one thread per code word; wtab == nullptr expands to T = code - 1
template <typename OT>
__global__ void __launch_bounds__(256)
tl_expand_kernel(const uint32_t* __restrict__ codes, int64_t wpr, const float4* __restrict__ wtab, int64_t n, int64_t m,
                 int64_t nb, int64_t block, const int32_t* __restrict__ perm, OT* __restrict__ out, int64_t ldo) {
    const int64_t words = (m + 15) >> 4;
    const int64_t total = n * words;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = q / words, w = q - r * words;
        const uint32_t word = codes[r * wpr + w];
        int64_t k = (w * 16) / block;
        if (k > nb - 1) k = nb - 1;
        const float4 wt = wtab ? wtab[r * nb + k] : make_float4(-1.f, 0.f, 1.f, 0.f);
        OT* row = out + r * ldo;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int64_t p = w * 16 + j;
            if (p < m) row[perm ? perm[p] : p] = tl_cast<OT>(tl_pick(wt, (word >> (2 * j)) & 3u));
        }
    }
}

static inline unsigned tl_grid(int64_t work, int threads) {
    int64_t g = ceil_div(work, threads);
    const int64_t cap = (int64_t)sm_count() * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

template <typename XT>
static int launch_gemv(const uint32_t* codes, int64_t wpr, const float* wtab, int64_t n, int64_t m, int64_t nb,
                       int64_t block, const XT* x, int64_t ldx, int64_t M, const int32_t* perm, const float* bias,
                       float* y, int64_t ldy, cudaStream_t st) {
    const unsigned grid = (unsigned)ceil_div(n, TL_WARPS * TL_ROWS_PER_WARP);
    const float4* wt = reinterpret_cast<const float4*>(wtab);
    for (int64_t t0 = 0; t0 < M; t0 += 4) {
        const int mt = (int)((M - t0) < 4 ? (M - t0) : 4);
        const XT* xt = x + t0 * ldx;
        float* yt = y + t0 * ldy;
        if (mt == 1)
            tl_gemv_kernel<XT, 1><<<grid, TL_WARPS * 32, 0, st>>>(codes, wpr, wt, (int)n, (int)m, (int)nb, (int)block, xt,
                                                                  ldx, mt, perm, bias, yt, ldy);
        else if (mt == 2)
            tl_gemv_kernel<XT, 2><<<grid, TL_WARPS * 32, 0, st>>>(codes, wpr, wt, (int)n, (int)m, (int)nb, (int)block, xt,
                                                                  ldx, mt, perm, bias, yt, ldy);
        else
            tl_gemv_kernel<XT, 4><<<grid, TL_WARPS * 32, 0, st>>>(codes, wpr, wt, (int)n, (int)m, (int)nb, (int)block, xt,
                                                                  ldx, mt, perm, bias, yt, ldy);
        TQ_LAUNCH_CHECK("tl_gemv_kernel");
    }
    return 0;
}

}  // namespace tq

extern "C" int64_t tq_tl_words_per_row(int64_t m) { return m > 0 ? (m + 15) / 16 : 0; }

extern "C" int tq_tl_pack(const int8_t* Torig, int64_t n, int64_t m, const int32_t* perm, uint32_t* codes, int64_t wpr,
                          void* stream) {
    using namespace tq;
    TQ_CHECK_ARG(Torig && codes && n > 0 && m > 0 && wpr >= ceil_div(m, 16), "tq_tl_pack: bad arguments");
    tl_pack_kernel<<<tl_grid(n * wpr, 256), 256, 0, (cudaStream_t)stream>>>(Torig, n, m, perm, codes, wpr);
    TQ_LAUNCH_CHECK("tl_pack_kernel");
    return 0;
}

extern "C" int tq_tl_wtab(const float* alpha, const float* mu, int64_t n, int64_t nb, int wdtype, float* wtab,
                          void* stream) {
    using namespace tq;
    TQ_CHECK_ARG(alpha && mu && wtab && n > 0 && nb > 0, "tq_tl_wtab: bad arguments");
    TQ_CHECK_ARG((reinterpret_cast<uintptr_t>(wtab) & 15) == 0, "tq_tl_wtab: wtab must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    float4* out = reinterpret_cast<float4*>(wtab);
    const int64_t count = n * nb;
    const unsigned grid = tl_grid(count, 256);
    if (wdtype == TQ_F32) tl_wtab_kernel<TQ_F32><<<grid, 256, 0, st>>>(alpha, mu, count, out);
    else if (wdtype == TQ_F16) tl_wtab_kernel<TQ_F16><<<grid, 256, 0, st>>>(alpha, mu, count, out);
    else if (wdtype == TQ_BF16) tl_wtab_kernel<TQ_BF16><<<grid, 256, 0, st>>>(alpha, mu, count, out);
    else { set_error("tq_tl_wtab: unknown dtype %d", wdtype); return TQ_E_BADARG; }
    TQ_LAUNCH_CHECK("tl_wtab_kernel");
    return 0;
}

extern "C" int tq_tl_gemv(const uint32_t* codes, int64_t wpr, const float* wtab, int64_t n, int64_t m, int64_t block,
                          const void* x, int xdtype, int64_t ldx, int64_t M, const int32_t* perm, const float* bias,
                          float* y, int64_t ldy, void* stream) {
    using namespace tq;
    TQ_CHECK_ARG(codes && wtab && x && y && n > 0 && m > 0 && M >= 0 && wpr >= ceil_div(m, 16) && ldx >= m && ldy >= n,
                 "tq_tl_gemv: bad arguments");
    TQ_CHECK_ARG(n < (1ll << 31) && m < (1ll << 31), "tq_tl_gemv: shape too large");
    if (block <= 0 || block % 16 != 0) {
        set_error("tq_tl_gemv: block size %lld is not a positive multiple of 16 (a code word must not straddle blocks)",
                  (long long)block);
        return TQ_E_UNSUPPORTED;
    }
    TQ_CHECK_ARG((reinterpret_cast<uintptr_t>(wtab) & 15) == 0, "tq_tl_gemv: wtab must be 16-byte aligned");
    if (M == 0) return 0;
    const int64_t nb = ceil_div(m, block);
    cudaStream_t st = (cudaStream_t)stream;
    switch (xdtype) {
        case TQ_F32: return launch_gemv(codes, wpr, wtab, n, m, nb, block, (const float*)x, ldx, M, perm, bias, y, ldy, st);
        case TQ_F16: return launch_gemv(codes, wpr, wtab, n, m, nb, block, (const __half*)x, ldx, M, perm, bias, y, ldy, st);
        case TQ_BF16: return launch_gemv(codes, wpr, wtab, n, m, nb, block, (const __nv_bfloat16*)x, ldx, M, perm, bias, y, ldy, st);
        default: set_error("tq_tl_gemv: unknown dtype %d", xdtype); return TQ_E_BADARG;
    }
}

extern "C" int tq_tl_dequant(const uint32_t* codes, int64_t wpr, const float* wtab, int64_t n, int64_t m, int64_t block,
                             const int32_t* perm, void* W, int wdtype, int64_t ldw, void* stream) {
    using namespace tq;
    TQ_CHECK_ARG(codes && wtab && W && n > 0 && m > 0 && wpr >= ceil_div(m, 16) && ldw >= m, "tq_tl_dequant: bad arguments");
    if (block <= 0 || block % 16 != 0) {
        set_error("tq_tl_dequant: block size %lld is not a positive multiple of 16", (long long)block);
        return TQ_E_UNSUPPORTED;
    }
    TQ_CHECK_ARG((reinterpret_cast<uintptr_t>(wtab) & 15) == 0, "tq_tl_dequant: wtab must be 16-byte aligned");
    const int64_t nb = ceil_div(m, block);
    cudaStream_t st = (cudaStream_t)stream;
    const float4* wt = reinterpret_cast<const float4*>(wtab);
    const unsigned grid = tl_grid(n * ceil_div(m, 16), 256);
    if (wdtype == TQ_F32) tl_expand_kernel<float><<<grid, 256, 0, st>>>(codes, wpr, wt, n, m, nb, block, perm, (float*)W, ldw);
    else if (wdtype == TQ_F16) tl_expand_kernel<__half><<<grid, 256, 0, st>>>(codes, wpr, wt, n, m, nb, block, perm, (__half*)W, ldw);
    else if (wdtype == TQ_BF16) tl_expand_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(codes, wpr, wt, n, m, nb, block, perm, (__nv_bfloat16*)W, ldw);
    else { set_error("tq_tl_dequant: unknown dtype %d", wdtype); return TQ_E_BADARG; }
    TQ_LAUNCH_CHECK("tl_expand_kernel");
    return 0;
}

extern "C" int tq_tl_unpack(const uint32_t* codes, int64_t wpr, int64_t n, int64_t m, const int32_t* perm, int8_t* Torig,
                            void* stream) {
    using namespace tq;
    TQ_CHECK_ARG(codes && Torig && n > 0 && m > 0 && wpr >= ceil_div(m, 16), "tq_tl_unpack: bad arguments");
    tl_expand_kernel<int8_t><<<tl_grid(n * ceil_div(m, 16), 256), 256, 0, (cudaStream_t)stream>>>(
        codes, wpr, nullptr, n, m, 1, m, perm, Torig, m);
    TQ_LAUNCH_CHECK("tl_expand_kernel<int8>");
    return 0;
}
