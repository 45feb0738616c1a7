// the consumer of the packed codes -- the reference's TernaryLinear (model.py:17-127) on 2 bits per weight.
//
// Layer format ("TL2"), chosen so that one 32-bit word never straddles a scale block and so that subset sums of the
// input can be looked up by nibble:
//   codes  u32 [n, wpr], wpr >= ceil(m/16): word w of row r holds sweep positions 16w .. 16w+15 as two bit planes:
//          bit j = 1 iff T[r, perm[16w+j]] = +1, bit 16+j = 1 iff it is -1; positions >= m hold 0 in both planes.
//          Sweep order, not original order: block k of alpha/mu (gptq.py:153-155) then covers the contiguous
//          positions [k*block, (k+1)*block) and the input is gathered once by perm (model.py:84's x[..., perm]).
//   wtab   f32 [n, nb, 4]: the three values a weight of (row, block) can take, rounded like the reference's
//          dequantisation `alpha * T + mu` evaluated in the layer dtype (model.py:106-108):
//          (w- = fl(mu - alpha), w0 = mu, w+ = fl(alpha + mu), 0).
// With them  y[t, r] = sum_p w[r, p/block][T(r,p)] * x[t, perm[p]]  is F.linear(x, Wq) with
// Wq = get_quantized_weight (gptq.py:201-230).  (model.py:84-90 permutes the input AND un-permutes W while its T is
// stored in original positions -- SURVEY Q11 -- so the reference's forward equals this only for the identity
// permutation; tests pin both facts.)
//
// Kernels (CUDA cores; the codes are 0.25 B per weight, so a decode-sized call is bound by reading them once):
//   tl_pack_kernel     int8 T in original positions + perm -> codes
//   tl_wtab_kernel     alpha, mu -> wtab
//   tl_gemv_kernel     <= 4 tokens per launch, by table lookup: per chunk of positions the CTA first builds, in shared
//                      memory, the 16 subset sums of every group of 4 gathered inputs; a weight row then costs two
//                      lookups per group (the +1 nibble and the -1 nibble) instead of 4 multiply-adds:
//                          y += w0 * sum(x) + (w+ - w0) * sum_{T=+1} x + (w- - w0) * sum_{T=-1} x      per code word
//                      (the differences are formed in fp32 from the rounded weights).  One warp per row, one code word
//                      per lane per step; the table columns are rotated so the 32 lanes of a lookup hit 32 banks.
//   tl_expand_kernel   codes -> dense Wq (f32/f16/bf16) or T (int8) in ORIGINAL column positions, for fp32 layers'
//                      many-token path (library GEMM on the dense weight) and for reading T back
#include "common.cuh"

namespace tq {

constexpr int TL_THREADS = 512;           // gemv CTA: 16 warps = 16 rows; two or three CTAs per SM overlap their phases
constexpr int TL_ROWS = TL_THREADS / 32;
constexpr int TL_ENTRIES = 4096;          // positions per chunk * tokens (64 KB of table per CTA)

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
            if (p < m) {
                const int t = row[perm ? perm[p] : p];
                word |= (t > 0 ? 1u : 0u) << j;
                word |= (t < 0 ? 1u : 0u) << (16 + j);
            }
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

// weight of position j of a code word
__device__ __forceinline__ float tl_pick(const float4& wt, uint32_t word, int j) {
    return ((word >> j) & 1u) ? wt.z : (((word >> (16 + j)) & 1u) ? wt.x : wt.y);
}

// Lookup table of one chunk: KC = TL_ENTRIES / MT positions = KC/4 groups of 4 positions; 128 groups (512 positions,
// the 32 code words one warp step reads) form a segment.  Entry (token t, segment s, nibble v, group gl of the segment)
// is the float at  tab[((t * SEGS + s) * 16 + v) * 128 + col(gl)],  col(gl) = (gl & ~31) | ((gl + (gl >> 5)) & 31):
// lane l looks up groups 4l + i, i.e. columns 32a + ((4b + i + a) & 31) with l = 8a + b -- 32 distinct banks for every
// nibble value, so a lookup is one conflict-free 4-byte shared load per token.
__device__ __forceinline__ int tl_col(int gl) { return (gl & ~31) | ((gl + (gl >> 5)) & 31); }

template <typename XT, int MT>
__global__ void __launch_bounds__(TL_THREADS, 2)
tl_gemv_kernel(const uint32_t* __restrict__ codes, int64_t wpr, const float4* __restrict__ wtab, int n, int m, int nb,
               int wpb_shift, uint32_t wpb_magic, const XT* __restrict__ x, int64_t ldx, int M,
               const int32_t* __restrict__ perm, const float* __restrict__ bias, float* __restrict__ y, int64_t ldy) {
    constexpr int KC = TL_ENTRIES / MT;          // positions per chunk
    constexpr int GPC = KC / 4;                  // groups per chunk
    constexpr int SEGS = KC / 512;               // warp steps per chunk
    constexpr int TOK_BYTES = SEGS * 16 * 128 * 4;   // one token's table
    extern __shared__ __align__(16) float tl_smem[];
    float* tab = tl_smem;                        // [MT][SEGS][16][128]
    float* wsum = tl_smem + MT * SEGS * 16 * 128;   // [MT][KC/16] sum of the 16 inputs of each code word

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = blockIdx.x * TL_ROWS + warp;
    const bool live = r < n;
    const int words = (m + 15) >> 4;
    const uint32_t* crow = codes + (int64_t)(live ? r : 0) * wpr;
    const float4* trow = wtab + (int64_t)(live ? r : 0) * nb;

    uint32_t cb[4];                              // byte offset of this lane's i-th group column within a nibble row
#pragma unroll
    for (int i = 0; i < 4; ++i) cb[i] = (uint32_t)tl_col(4 * lane + i) * 4u;

    float acc[MT];
#pragma unroll
    for (int t = 0; t < MT; ++t) acc[t] = 0.f;

    for (int c0 = 0; c0 < m; c0 += KC) {
        // this chunk's code words first: independent global loads in flight while the table is built
        uint32_t wd[SEGS];
#pragma unroll
        for (int s = 0; s < SEGS; ++s) {
            const int w = (c0 >> 4) + 32 * s + lane;
            wd[s] = (live && w < words) ? __ldg(crow + w) : 0u;
        }
        __syncthreads();                         // previous chunk's table fully consumed
        // build: item = (token bt, group bg); the 16 subset sums of inputs 4 bg .. 4 bg + 3 (entry v = sum over the bits of v)
#pragma unroll
        for (int item = threadIdx.x; item < GPC * MT; item += TL_THREADS) {
            const int bt = item / GPC, bg = item - bt * GPC;
            float xv[4];
            const int p0 = c0 + 4 * bg;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int p = p0 + j;
                float v = 0.f;
                if (p < m && bt < M) v = to_f32<XT>(x[(int64_t)bt * ldx + (perm ? perm[p] : p)]);
                xv[j] = v;
            }
            float e[16];
            e[0] = 0.f;
#pragma unroll
            for (int v = 1; v < 16; ++v) {
                const int low = v & (-v);                                  // lowest set bit
                const int j = (low == 1) ? 0 : (low == 2) ? 1 : (low == 4) ? 2 : 3;
                e[v] = e[v ^ low] + xv[j];
            }
            float* dst = tab + ((bt * SEGS + (bg >> 7)) * 16) * 128 + tl_col(bg & 127);
#pragma unroll
            for (int v = 0; v < 16; ++v) dst[v * 128] = e[v];
            // word totals: groups 4q .. 4q+3 are adjacent lanes
            float tot = e[15];
            tot += __shfl_xor_sync(0xffffffffu, tot, 1);
            tot += __shfl_xor_sync(0xffffffffu, tot, 2);
            if ((bg & 3) == 0) wsum[bt * (KC / 16) + (bg >> 2)] = tot;
        }
        __syncthreads();
        if (live) {
            const char* tabc = reinterpret_cast<const char*>(tab);
#pragma unroll
            for (int s = 0; s < SEGS; ++s) {
                const int wl = 32 * s + lane;                              // word within the chunk
                const int w = (c0 >> 4) + wl;
                if (w >= words) continue;
                const uint32_t word = wd[s];
                const char* seg = tabc + s * (16 * 128 * 4);
                // byte offset of (nibble value, column): value * 512 | column * 4
                uint32_t op[4], on[4];
                op[0] = ((word << 9) & 0x1E00u) | cb[0];
                op[1] = ((word << 5) & 0x1E00u) | cb[1];
                op[2] = ((word << 1) & 0x1E00u) | cb[2];
                op[3] = ((word >> 3) & 0x1E00u) | cb[3];
                on[0] = ((word >> 7) & 0x1E00u) | cb[0];
                on[1] = ((word >> 11) & 0x1E00u) | cb[1];
                on[2] = ((word >> 15) & 0x1E00u) | cb[2];
                on[3] = ((word >> 19) & 0x1E00u) | cb[3];
                const uint32_t k = wpb_shift >= 0 ? ((uint32_t)w >> wpb_shift) : __umulhi((uint32_t)w, wpb_magic);
                const float4 wt = __ldg(trow + min((int)k, nb - 1));
                const float dp = __fsub_rn(wt.z, wt.y), dn = __fsub_rn(wt.x, wt.y);
#pragma unroll
                for (int t = 0; t < MT; ++t) {
                    const char* tk = seg + t * TOK_BYTES;
                    float sp = 0.f, sn = 0.f;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        sp += *reinterpret_cast<const float*>(tk + op[i]);
                        sn += *reinterpret_cast<const float*>(tk + on[i]);
                    }
                    const float tot = wsum[t * (KC / 16) + wl];
                    acc[t] = fmaf(wt.y, tot, fmaf(dp, sp, fmaf(dn, sn, acc[t])));
                }
            }
        }
    }
#pragma unroll
    for (int t = 0; t < MT; ++t) {
        const float v = warp_sum(acc[t]);
        if (lane == 0 && live && t < M) y[(int64_t)t * ldy + r] = bias ? __fadd_rn(v, bias[r]) : v;
    }
}

template <typename OT> __device__ __forceinline__ OT tl_cast(float v);
template <> __device__ __forceinline__ float tl_cast<float>(float v) { return v; }
template <> __device__ __forceinline__ __half tl_cast<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 tl_cast<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ int8_t tl_cast<int8_t>(float v) { return (int8_t)v; }

// One thread per PAIR of output columns, consecutive threads = consecutive pairs of one row (grid.y strides over rows), so
// a warp writes 64 consecutive elements.  Identity order: column = position.  Permuted layer with inv_perm: the thread
// GATHERS the code of position inv_perm[column] (the row's code words are 1-3 KB, L1-resident) and the writes stay
// coalesced; with perm only it falls back to scattering position p to column perm[p].  wtab == nullptr expands to T.
template <typename OT>
__global__ void __launch_bounds__(256)
tl_expand_kernel(const uint32_t* __restrict__ codes, int64_t wpr, const float4* __restrict__ wtab, int n, int m, int nb,
                 int wpb_shift, uint32_t wpb_magic, const int32_t* __restrict__ perm, const int32_t* __restrict__ inv_perm,
                 OT* __restrict__ out, int64_t ldo) {
    const int pairs = (m + 1) >> 1;
    const float4 unit = make_float4(-1.f, 0.f, 1.f, 0.f);
    for (int r = blockIdx.y; r < n; r += gridDim.y) {
        const uint32_t* crow = codes + (int64_t)r * wpr;
        const float4* trow = wtab ? wtab + (int64_t)r * nb : nullptr;
        OT* row = out + (int64_t)r * ldo;
        auto value = [&](int p) {
            const int w = p >> 4;
            const uint32_t k = wpb_shift >= 0 ? ((uint32_t)w >> wpb_shift) : __umulhi((uint32_t)w, wpb_magic);
            const float4 wt = trow ? __ldg(trow + min((int)k, nb - 1)) : unit;
            return tl_cast<OT>(tl_pick(wt, __ldg(crow + w), p & 15));
        };
        for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < pairs; q += gridDim.x * blockDim.x) {
            const int c0 = q * 2;
            const bool two = c0 + 1 < m;
            if (perm && !inv_perm) {                       // scatter
                row[perm[c0]] = value(c0);
                if (two) row[perm[c0 + 1]] = value(c0 + 1);
            } else {
                row[c0] = value(inv_perm ? inv_perm[c0] : c0);
                if (two) row[c0 + 1] = value(inv_perm ? inv_perm[c0 + 1] : c0 + 1);
            }
        }
    }
}

template <typename OT>
static int launch_expand(const uint32_t* codes, int64_t wpr, const float4* wt, int64_t n, int64_t m, int64_t nb, int64_t block,
                         const int32_t* perm, const int32_t* inv_perm, OT* out, int64_t ldo, cudaStream_t st) {
    const uint32_t wpb = block / 16 > 0 ? (uint32_t)(block / 16) : 1u;
    int wpb_shift = -1;
    for (int sft = 0; sft < 28; ++sft)
        if ((1u << sft) == wpb) wpb_shift = sft;
    const uint32_t magic = wpb > 1 ? (uint32_t)(((1ull << 32) + wpb - 1) / wpb) : 0u;
    const int64_t pairs = (m + 1) / 2;
    dim3 grid((unsigned)(ceil_div(pairs, 256) < 32 ? ceil_div(pairs, 256) : 32), (unsigned)(n < 65535 ? n : 65535));
    tl_expand_kernel<OT><<<grid, 256, 0, st>>>(codes, wpr, wt, (int)n, (int)m, (int)nb, wpb_shift, magic, perm, inv_perm, out, ldo);
    TQ_LAUNCH_CHECK("tl_expand_kernel");
    return 0;
}

static inline unsigned tl_grid(int64_t work, int threads) {
    int64_t g = ceil_div(work, threads);
    const int64_t cap = (int64_t)sm_count() * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

template <int MT> constexpr int tl_gemv_smem() { return TL_ENTRIES * 4 * 4 + (TL_ENTRIES / 16) * 4; }

template <typename XT, int MT>
static int launch_gemv_mt(const uint32_t* codes, int64_t wpr, const float4* wt, int64_t n, int64_t m, int64_t nb,
                          int64_t block, const XT* x, int64_t ldx, int mt, const int32_t* perm, const float* bias, float* y,
                          int64_t ldy, cudaStream_t st) {
    constexpr int smem = tl_gemv_smem<MT>();
    static std::atomic<unsigned long long> attr_devices{0};
    int dev;
    if (dyn_smem_pending(attr_devices, dev)) {
        TQ_CUDA(cudaFuncSetAttribute(tl_gemv_kernel<XT, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        dyn_smem_done(attr_devices, dev);
    }
    // block index of code word w = w / (block / 16): a shift, or a multiply-high by ceil(2^32 / wpb) (exact for
    // w * wpb < 2^32, i.e. any m < 2^31)
    const uint32_t wpb = (uint32_t)(block / 16);
    int wpb_shift = -1;
    for (int sft = 0; sft < 28; ++sft)
        if ((1u << sft) == wpb) wpb_shift = sft;
    const uint32_t magic = wpb > 1 ? (uint32_t)(((1ull << 32) + wpb - 1) / wpb) : 0u;
    tl_gemv_kernel<XT, MT><<<(unsigned)ceil_div(n, TL_ROWS), TL_THREADS, smem, st>>>(
        codes, wpr, wt, (int)n, (int)m, (int)nb, wpb_shift, magic, x, ldx, mt, perm, bias, y, ldy);
    TQ_LAUNCH_CHECK("tl_gemv_kernel");
    return 0;
}

template <typename XT>
static int launch_gemv(const uint32_t* codes, int64_t wpr, const float* wtab, int64_t n, int64_t m, int64_t nb,
                       int64_t block, const XT* x, int64_t ldx, int64_t M, const int32_t* perm, const float* bias,
                       float* y, int64_t ldy, cudaStream_t st) {
    const float4* wt = reinterpret_cast<const float4*>(wtab);
    for (int64_t t0 = 0; t0 < M; t0 += 4) {
        const int mt = (int)((M - t0) < 4 ? (M - t0) : 4);
        const XT* xt = x + t0 * ldx;
        float* yt = y + t0 * ldy;
        int rc;
        if (mt == 1) rc = launch_gemv_mt<XT, 1>(codes, wpr, wt, n, m, nb, block, xt, ldx, mt, perm, bias, yt, ldy, st);
        else if (mt == 2) rc = launch_gemv_mt<XT, 2>(codes, wpr, wt, n, m, nb, block, xt, ldx, mt, perm, bias, yt, ldy, st);
        else rc = launch_gemv_mt<XT, 4>(codes, wpr, wt, n, m, nb, block, xt, ldx, mt, perm, bias, yt, ldy, st);
        if (rc) return rc;
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
    TQ_CHECK_ARG(n < (1ll << 31) - 64 && m < (1ll << 31) - 8192, "tq_tl_gemv: shape too large");
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
                             const int32_t* perm, const int32_t* inv_perm, void* W, int wdtype, int64_t ldw, void* stream) {
    using namespace tq;
    TQ_CHECK_ARG(codes && wtab && W && n > 0 && m > 0 && wpr >= ceil_div(m, 16) && ldw >= m, "tq_tl_dequant: bad arguments");
    if (block <= 0 || block % 16 != 0) {
        set_error("tq_tl_dequant: block size %lld is not a positive multiple of 16", (long long)block);
        return TQ_E_UNSUPPORTED;
    }
    TQ_CHECK_ARG((reinterpret_cast<uintptr_t>(wtab) & 15) == 0, "tq_tl_dequant: wtab must be 16-byte aligned");
    TQ_CHECK_ARG(n < (1ll << 31) && m < (1ll << 31) - 2, "tq_tl_dequant: shape too large");
    const int64_t nb = ceil_div(m, block);
    cudaStream_t st = (cudaStream_t)stream;
    const float4* wt = reinterpret_cast<const float4*>(wtab);
    if (wdtype == TQ_F32) return launch_expand<float>(codes, wpr, wt, n, m, nb, block, perm, inv_perm, (float*)W, ldw, st);
    if (wdtype == TQ_F16) return launch_expand<__half>(codes, wpr, wt, n, m, nb, block, perm, inv_perm, (__half*)W, ldw, st);
    if (wdtype == TQ_BF16) return launch_expand<__nv_bfloat16>(codes, wpr, wt, n, m, nb, block, perm, inv_perm, (__nv_bfloat16*)W, ldw, st);
    set_error("tq_tl_dequant: unknown dtype %d", wdtype);
    return TQ_E_BADARG;
}

extern "C" int tq_tl_unpack(const uint32_t* codes, int64_t wpr, int64_t n, int64_t m, const int32_t* perm,
                            const int32_t* inv_perm, int8_t* Torig, void* stream) {
    using namespace tq;
    TQ_CHECK_ARG(codes && Torig && n > 0 && m > 0 && wpr >= ceil_div(m, 16), "tq_tl_unpack: bad arguments");
    TQ_CHECK_ARG(n < (1ll << 31) && m < (1ll << 31) - 2, "tq_tl_unpack: shape too large");
    return launch_expand<int8_t>(codes, wpr, nullptr, n, m, 1, ((m + 15) / 16) * 16, perm, inv_perm, Torig, m,
                                 (cudaStream_t)stream);
}
