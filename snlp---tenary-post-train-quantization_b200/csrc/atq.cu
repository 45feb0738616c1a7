// A5-A7 + block error: the asymmetric ternary quantizer of one column block, one warp per row.
// Reference: quantizer.py:32-69 (init), :71-108 (grid), :110-134 (round), :136-175 (ITF),
// :177-248 (AGA), gptq.py:158-159 (E = W_b - (alpha T + mu)).
//
// Each lane keeps EPL consecutive block positions of its row in registers (EPL = 4 for the default
// 128-column block: one 128-bit load).  Float row sums are butterfly shuffles; the integer sums
// (sum T, sum T^2) are one REDUX each.  The ITF stop test is per row: a converged row is a fixed
// point of the reference's global loop (quantizer.py:164), an oscillating row runs to max_iter
// either way, and -- like the global loop whenever any other row is still moving -- a row whose
// initial T is all zero is NOT stopped at iteration 0.
//
// All closed-form expressions use explicitly rounded mul/sub/div so the compiler cannot contract
// them into FMAs the reference (separate ATen ops) does not perform.
#include "common.cuh"

namespace tq {

template <int EPL>
struct Row {
    float w[EPL];
    int t[EPL];
    bool valid[EPL];
    float ws;        // sum W over the block (constant while T changes)
    float fb;        // block width as float
    int p0;

    __device__ __forceinline__ void load(const float* __restrict__ wrow, const int32_t* __restrict__ blk_idx, int col0,
                                         int b, int lane, bool vec_ok) {
        p0 = lane * EPL;
        fb = (float)b;
#pragma unroll
        for (int e = 0; e < EPL; ++e) valid[e] = (p0 + e) < b;
        if (blk_idx == nullptr) {
            bool vec = false;
            if (EPL == 4) {
                vec = vec_ok && (p0 + 3 < b);
                if (vec) {
                    const float4 v = *reinterpret_cast<const float4*>(wrow + col0 + p0);
                    w[0] = v.x; w[1 % EPL] = v.y; w[2 % EPL] = v.z; w[3 % EPL] = v.w;
                }
            }
            if (!vec) {
#pragma unroll
                for (int e = 0; e < EPL; ++e) w[e] = valid[e] ? wrow[col0 + p0 + e] : 0.f;
            }
        } else {
#pragma unroll
            for (int e = 0; e < EPL; ++e) w[e] = valid[e] ? wrow[blk_idx[p0 + e]] : 0.f;
        }
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < EPL; ++e) s += w[e];
        ws = warp_sum(s);
    }

    // quantizer.py:49-67
    __device__ __forceinline__ void init(float& alpha, float& mu) {
        mu = __fdiv_rn(ws, fb);
        float sabs = 0.f;
#pragma unroll
        for (int e = 0; e < EPL; ++e) sabs += valid[e] ? fabsf(__fsub_rn(w[e], mu)) : 0.f;
        sabs = warp_sum(sabs);
        const float delta = __fmul_rn(0.75f, __fdiv_rn(sabs, fb));
        float num = 0.f;
        int cnt = 0;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            const float wc = __fsub_rn(w[e], mu);
            t[e] = valid[e] ? ((wc > delta) ? 1 : ((wc < -delta) ? -1 : 0)) : 0;
            num += (float)t[e] * wc;
            cnt += t[e] * t[e];
        }
        num = warp_sum(num);
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        alpha = __fdiv_rn(num, fmaxf((float)cnt, kTiny));
    }

    // quantizer.py:93-106
    __device__ __forceinline__ void grid(float& alpha, float& mu) const {
        float wt = 0.f;
        int ts = 0, t2 = 0;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            wt += (float)t[e] * w[e];
            ts += t[e];
            t2 += t[e] * t[e];
        }
        wt = warp_sum(wt);
        ts = __reduce_add_sync(0xffffffffu, ts);
        t2 = __reduce_add_sync(0xffffffffu, t2);
        const float fts = (float)ts, ft2 = (float)t2;
        const float den = fmaxf(__fsub_rn(__fmul_rn(fb, ft2), __fmul_rn(fts, fts)), kTiny);
        alpha = __fdiv_rn(__fsub_rn(__fmul_rn(fb, wt), __fmul_rn(fts, ws)), den);
        mu = __fdiv_rn(__fsub_rn(__fmul_rn(ft2, ws), __fmul_rn(fts, wt)), den);
    }

    // quantizer.py:125-132; returns whether any code of the row changed
    __device__ __forceinline__ bool round(float alpha, float mu) {
        const float a_safe = fmaxf(alpha, kTiny);
        bool changed = false;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            const float z = __fdiv_rn(__fsub_rn(w[e], mu), a_safe);
            const int tn = valid[e] ? ((z > 0.5f) ? 1 : ((z < -0.5f) ? -1 : 0)) : 0;
            changed |= (tn != t[e]);
            t[e] = tn;
        }
        return __any_sync(0xffffffffu, changed);
    }

    // quantizer.py:162-175 from the current T; returns the number of grid+round rounds performed
    __device__ __forceinline__ int itf(int max_iter, float& alpha, float& mu) {
        int it = 0;
        while (it < max_iter) {
            grid(alpha, mu);
            ++it;
            if (!round(alpha, mu)) break;
        }
        // (alpha, mu) are the grid of the T that was rounded last, i.e. of the returned T at convergence and of
        // its predecessor when max_iter ran out -- exactly what the reference returns.
        return it;
    }

    // quantizer.py:215-246 from s1 = S 1 and d = 1'S1
    __device__ __forceinline__ void aga(const float* __restrict__ s1d, int b, float& alpha, float& mu) const {
        float v = 0.f, wsd = 0.f, wts = 0.f, t2s = 0.f;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            const float s = valid[e] ? s1d[p0 + e] : 0.f;
            const float ft = (float)t[e];
            v += ft * s;
            wsd += w[e] * s;
            wts += __fmul_rn(w[e], ft) * s;
            t2s += __fmul_rn(ft, ft) * s;
        }
        v = warp_sum(v); wsd = warp_sum(wsd); wts = warp_sum(wts); t2s = warp_sum(t2s);
        const float d = s1d[b];
        const float den = fmaxf(__fsub_rn(__fmul_rn(d, t2s), __fmul_rn(v, v)), kTiny);
        alpha = __fdiv_rn(__fsub_rn(__fmul_rn(d, wts), __fmul_rn(v, wsd)), den);
        mu = __fdiv_rn(__fsub_rn(__fmul_rn(t2s, wsd), __fmul_rn(v, wts)), den);
    }

    __device__ __forceinline__ void load_codes(const int8_t* __restrict__ trow) {
#pragma unroll
        for (int e = 0; e < EPL; ++e) t[e] = valid[e] ? (int)trow[p0 + e] : 0;
    }

    __device__ __forceinline__ void store_codes(int8_t* __restrict__ trow, int b, bool vec_ok) const {
        if (EPL == 4 && vec_ok && (p0 + 3 < b)) {
            char4 c;
            c.x = (signed char)t[0]; c.y = (signed char)t[1 % EPL]; c.z = (signed char)t[2 % EPL]; c.w = (signed char)t[3 % EPL];
            *reinterpret_cast<char4*>(trow + p0) = c;
        } else {
#pragma unroll
            for (int e = 0; e < EPL; ++e)
                if (valid[e]) trow[p0 + e] = (int8_t)t[e];
        }
    }

    // gptq.py:158-159.  With erow_lo the error is written as the (hi, lo) tf32 split the tensor-core feedback
    // GEMM consumes (hi = rn_tf32(e), lo = rn_tf32(e - hi)), else as plain fp32.
    __device__ __forceinline__ void store_error(float* __restrict__ erow, float* __restrict__ erow_lo, int b, float alpha,
                                                float mu, bool vec_ok) const {
        float ev[EPL], el[EPL];
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            ev[e] = __fsub_rn(w[e], __fadd_rn(__fmul_rn(alpha, (float)t[e]), mu));
            el[e] = 0.f;
            if (erow_lo != nullptr) {
                uint32_t hb, lb;                         // hi = rn_tf32(e), lo = rn_tf32(e - hi)  (gemm_tc.cu)
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(ev[e]));
                const float hi = __uint_as_float(hb);
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(__fsub_rn(ev[e], hi)));
                el[e] = __uint_as_float(lb);
                ev[e] = hi;
            }
        }
        if (EPL == 4 && vec_ok && (p0 + 3 < b)) {
            *reinterpret_cast<float4*>(erow + p0) = make_float4(ev[0], ev[1 % EPL], ev[2 % EPL], ev[3 % EPL]);
            if (erow_lo != nullptr)
                *reinterpret_cast<float4*>(erow_lo + p0) = make_float4(el[0], el[1 % EPL], el[2 % EPL], el[3 % EPL]);
        } else {
#pragma unroll
            for (int e = 0; e < EPL; ++e)
                if (valid[e]) {
                    erow[p0 + e] = ev[e];
                    if (erow_lo != nullptr) erow_lo[p0 + e] = el[e];
                }
        }
    }
};

template <int EPL>
__global__ void __launch_bounds__(256)
atq_block_kernel(const float* __restrict__ W, int64_t ldw, int n, const int32_t* __restrict__ blk_idx, int col0,
                 int b, const float* __restrict__ s1d, int max_iter, int8_t* __restrict__ T, int64_t ldt,
                 float* __restrict__ alpha_out, float* __restrict__ mu_out, int64_t ld_am,
                 float* __restrict__ E, float* __restrict__ E_lo, int64_t lde, int32_t* __restrict__ iters_out,
                 const float* __restrict__ rowsum_part, int rowsum_parts, const double* __restrict__ csum, int rem_next,
                 float* __restrict__ wbar_next) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    Row<EPL> r;
    const bool w_vec = ((ldw & 3) == 0) && ((col0 & 3) == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
    r.load(W + (int64_t)row * ldw, blk_idx, col0, b, lane, w_vec);
    float alpha, mu;
    r.init(alpha, mu);
    const int it = r.itf(max_iter, alpha, mu);
    if (s1d != nullptr) r.aga(s1d, b, alpha, mu);
    if (lane == 0) {
        alpha_out[(int64_t)row * ld_am] = alpha;
        mu_out[(int64_t)row * ld_am] = mu;
        if (iters_out) iters_out[row] = it;
    }
    r.store_codes(T + (int64_t)row * ldt, b, ((ldt & 3) == 0) && ((reinterpret_cast<uintptr_t>(T) & 3) == 0));
    if (wbar_next != nullptr) {
        // SSR bookkeeping for the NEXT selection: the row sum over the next remaining set, predicted exactly from
        // this block's removal and feedback:  sum_new = sum_cur - sum(W_b) - E . (C 1)   (gptq.py:186 summed over j)
        double bs = 0.0, ec = 0.0, cursum = 0.0;
        // current exact row sum = sum of the partials the previous feedback epilogue (or the first statistics pass) left
        for (int t = lane; t < rowsum_parts; t += 32) cursum += (double)rowsum_part[(int64_t)t * n + row];
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            if (r.valid[e]) {
                const float ev = __fsub_rn(r.w[e], __fadd_rn(__fmul_rn(alpha, (float)r.t[e]), mu));
                bs += (double)r.w[e];
                ec += (double)ev * csum[r.p0 + e];
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            bs += __shfl_xor_sync(0xffffffffu, bs, o);
            ec += __shfl_xor_sync(0xffffffffu, ec, o);
            cursum += __shfl_xor_sync(0xffffffffu, cursum, o);
        }
        if (lane == 0) wbar_next[row] = (float)((cursum - bs - ec) / (double)rem_next);
    }
    if (E != nullptr)
        r.store_error(E + (int64_t)row * lde, E_lo ? E_lo + (int64_t)row * lde : nullptr, b, alpha, mu,
                      ((lde & 3) == 0) && ((reinterpret_cast<uintptr_t>(E) & 15) == 0) &&
                          ((reinterpret_cast<uintptr_t>(E_lo) & 15) == 0));
}

// The individual stages of the quantizer API (not on the sweep's path): one op per launch.
enum AtqOp { ATQ_OP_INIT = 0, ATQ_OP_GRID = 1, ATQ_OP_ROUND = 2, ATQ_OP_ITF = 3, ATQ_OP_AGA = 4 };

template <int EPL>
__global__ void __launch_bounds__(256)
atq_stage_kernel(int op, const float* __restrict__ W, int64_t ldw, int n, int b, const int8_t* __restrict__ T_in,
                 const float* __restrict__ alpha_in, const float* __restrict__ mu_in, const float* __restrict__ s1d,
                 int max_iter, int8_t* __restrict__ T_out, float* __restrict__ alpha_out, float* __restrict__ mu_out) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    Row<EPL> r;
    r.load(W + (int64_t)row * ldw, nullptr, 0, b, lane, ((ldw & 3) == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0));
    float alpha = alpha_in ? alpha_in[row] : 0.f, mu = mu_in ? mu_in[row] : 0.f;
    if (T_in) r.load_codes(T_in + (int64_t)row * b);
    else {
#pragma unroll
        for (int e = 0; e < EPL; ++e) r.t[e] = 0;
    }
    switch (op) {
        case ATQ_OP_INIT:  r.init(alpha, mu); break;
        case ATQ_OP_GRID:  r.grid(alpha, mu); break;
        case ATQ_OP_ROUND: r.round(alpha, mu); break;
        case ATQ_OP_ITF:   r.itf(max_iter, alpha, mu); break;
        case ATQ_OP_AGA:   r.aga(s1d, b, alpha, mu); break;
    }
    if (lane == 0) {
        if (alpha_out) alpha_out[row] = alpha;
        if (mu_out) mu_out[row] = mu;
    }
    if (T_out) r.store_codes(T_out + (int64_t)row * b, b, false);
}

// s1 = S 1, d = 1'S1 for one block.  HESSIAN: S = Hb'Hb (Hb = Hsrc[blk,blk], what gptq.py:147-150 feeds
// quantizer.py:207); ACTIVATIONS: S = Hsrc[blk,blk] = (X'X)[blk,blk] (main.py:177-180).  One CTA of 32
// warps; every output is a row-times-vector product of the gathered sub-block (warp per row, lanes along
// the row).  Hsrc is exactly symmetric (tq_hessian_finalize / tq_symmetrize mirror the upper triangle), so
// Hb'(Hb 1) is computed as Hb (Hb 1) with the same row-wise access.
__global__ void __launch_bounds__(1024)
aga_vector_kernel(const float* __restrict__ Hsrc, int64_t ldh, const int32_t* __restrict__ blk_idx, int col0,
                  int b, int mode, float* __restrict__ s1d, const float* __restrict__ csum_part, int csum_parts,
                  double* __restrict__ csum) {
    // One CTA, run once per block step: the code is kept SMALL on purpose (rolled loops, no per-row unrolling) -- the first
    // version compiled to 5 800 straight-line instructions and spent its 27 us fetching them (ncu: 47 k cycles for 40 k
    // warp instructions, profiles/r02a notes in DESIGN section 3).
    extern __shared__ float sh[];           // cols[b] (as int), u[b], s1[b]
    int* cols = reinterpret_cast<int*>(sh);
    float* u = sh + b;
    float* s1 = sh + 2 * b;
    __shared__ float red[32];
    __shared__ double fold[1024];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int p = threadIdx.x; p < b; p += blockDim.x) cols[p] = blk_idx ? blk_idx[p] : col0 + p;
    // (independent of the AGA vector) fold the coefficient kernel's per-CTA partials of C 1 in a fixed order:
    // blockDim / b threads per column take interleaved partials (eight loads in flight), one thread adds their sums
    if (csum != nullptr) {
        const int subs = max(1, (int)blockDim.x / b);
        const int sub = threadIdx.x / b, i = threadIdx.x - sub * b;
        if (sub < subs) {
            double s = 0.0;
            int t = sub;
#pragma unroll 1
            for (; t + 7 * subs < csum_parts; t += 8 * subs) {
                float v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = csum_part[(int64_t)(t + q * subs) * b + i];
#pragma unroll
                for (int q = 0; q < 8; ++q) s += (double)v[q];
            }
#pragma unroll 1
            for (; t < csum_parts; t += subs) s += (double)csum_part[(int64_t)t * b + i];
            fold[sub * b + i] = s;
        }
    }
    __syncthreads();
    if (csum != nullptr && threadIdx.x < b) {
        const int subs = max(1, (int)blockDim.x / b);
        double s = 0.0;
#pragma unroll 1
        for (int q = 0; q < subs; ++q) s += fold[q * b + threadIdx.x];
        csum[threadIdx.x] = s;
    }
    if (mode == TQ_AGA_NONE) return;
    // row-times-vector products of the gathered sub-block: warp w owns rows w, w + nwarp, ...; lanes walk the row.
    // Phase 0: u = Hb 1.  Phase 1 (HESSIAN: S = Hb'Hb, Hb symmetric): s1 = Hb u.
#pragma unroll 1
    for (int phase = 0; phase < 2; ++phase) {
        if (phase == 1 && mode != TQ_AGA_HESSIAN) break;
        float* out = (phase == 0) ? u : s1;
#pragma unroll 1
        for (int i = warp; i < b; i += nwarp) {
            const float* hrow = Hsrc + (int64_t)cols[i] * ldh;
            float acc = 0.f;
#pragma unroll 4
            for (int j = lane; j < b; j += 32) {
                const float h = hrow[cols[j]];
                acc = (phase == 0) ? acc + h : fmaf(h, u[j], acc);
            }
            acc = warp_sum(acc);
            if (lane == 0) out[i] = acc;
        }
        __syncthreads();
    }
    const float* res = (mode == TQ_AGA_HESSIAN) ? s1 : u;
    float part = 0.f;
    for (int j = threadIdx.x; j < b; j += blockDim.x) {
        const float v = res[j];
        s1d[j] = v;
        part += v;
    }
    part = warp_sum(part);
    if (lane == 0) red[warp] = part;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = (threadIdx.x < nwarp) ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) s1d[b] = v;
    }
}

int launch_atq_block(const float* W, int64_t ldw, int64_t n, const int32_t* blk_idx, int64_t col0, int64_t b,
                     const float* s1d, int max_iter, int8_t* T, int64_t ldt, float* alpha, float* mu,
                     int64_t ld_am, float* E, float* E_lo, int64_t lde, int32_t* iters, const float* rowsum_part,
                     int64_t rowsum_parts, const double* csum, int64_t rem_next, float* wbar_next, cudaStream_t st) {
    const int warps = 8;
    dim3 grid((unsigned)ceil_div(n, warps)), block(warps * 32);
#define TQ_ATQ_CASE(EPL)                                                                                      \
    atq_block_kernel<EPL><<<grid, block, 0, st>>>(W, ldw, (int)n, blk_idx, (int)col0, (int)b, s1d, max_iter, \
                                                  T, ldt, alpha, mu, ld_am, E, E_lo, lde, iters, rowsum_part,          \
                                                  (int)rowsum_parts, csum, (int)rem_next, wbar_next)
    if (b <= 32) TQ_ATQ_CASE(1);
    else if (b <= 64) TQ_ATQ_CASE(2);
    else if (b <= 128) TQ_ATQ_CASE(4);
    else if (b <= 256) TQ_ATQ_CASE(8);
    else TQ_ATQ_CASE(16);
#undef TQ_ATQ_CASE
    TQ_LAUNCH_CHECK("atq_block_kernel");
    return 0;
}

// AGA vector of the block (mode HESSIAN / ACTIVATIONS) and/or the fold of the coefficient row sums (csum != NULL)
int launch_aga_vector(const float* Hsrc, int64_t ldh, const int32_t* blk_idx, int64_t col0, int64_t b, int mode,
                      float* s1d, const float* csum_part, int64_t csum_parts, double* csum, cudaStream_t st) {
    if (mode == TQ_AGA_NONE && csum == nullptr) return 0;
    aga_vector_kernel<<<1, 1024, 3 * b * sizeof(float), st>>>(Hsrc, ldh, blk_idx, (int)col0, (int)b, mode, s1d, csum_part,
                                                              (int)csum_parts, csum);
    TQ_LAUNCH_CHECK("aga_vector_kernel");
    return 0;
}

}  // namespace tq

extern "C" int tq_atq_block(const float* W, int64_t ldw, int64_t n, const int32_t* blk_idx, int64_t col0,
                            int64_t b, const float* s1d, int max_iter, int8_t* T, int64_t ldt, float* alpha,
                            float* mu, int64_t ld_am, float* E, int64_t lde, int32_t* iters, void* stream) {
    using namespace tq;
    TQ_CHECK_ARG(W && T && alpha && mu, "tq_atq_block: null pointer");
    TQ_CHECK_ARG(n > 0 && b > 0 && b <= 512, "tq_atq_block: need n > 0 and 1 <= block <= 512 (got n=%lld b=%lld)",
                 (long long)n, (long long)b);
    TQ_CHECK_ARG(ldt >= b && ld_am >= 1 && (E == nullptr || lde >= b) && max_iter >= 0, "tq_atq_block: bad strides");
    return launch_atq_block(W, ldw, n, blk_idx, col0, b, s1d, max_iter, T, ldt, alpha, mu, ld_am, E, nullptr, lde, iters,
                            nullptr, 0, nullptr, 0, nullptr, (cudaStream_t)stream);
}

extern "C" int tq_atq_stage(int op, const float* W, int64_t ldw, int64_t n, int64_t b, const int8_t* T_in,
                            const float* alpha_in, const float* mu_in, const float* s1d, int max_iter,
                            int8_t* T_out, float* alpha_out, float* mu_out, void* stream) {
    using namespace tq;
    TQ_CHECK_ARG(W && n > 0 && b > 0 && b <= 512, "tq_atq_stage: bad shape");
    TQ_CHECK_ARG(op >= ATQ_OP_INIT && op <= ATQ_OP_AGA, "tq_atq_stage: unknown op %d", op);
    TQ_CHECK_ARG(op == ATQ_OP_INIT || op == ATQ_OP_ROUND || T_in, "tq_atq_stage: this op needs T_in");
    TQ_CHECK_ARG(op != ATQ_OP_ROUND || (alpha_in && mu_in), "tq_atq_stage: ROUND needs alpha_in and mu_in");
    TQ_CHECK_ARG(op != ATQ_OP_AGA || s1d, "tq_atq_stage: AGA needs s1d");
    cudaStream_t st = (cudaStream_t)stream;
    const int warps = 8;
    dim3 grid((unsigned)ceil_div(n, warps)), block(warps * 32);
#define TQ_STAGE_CASE(EPL)                                                                                       \
    atq_stage_kernel<EPL><<<grid, block, 0, st>>>(op, W, ldw, (int)n, (int)b, T_in, alpha_in, mu_in, s1d, max_iter, \
                                                  T_out, alpha_out, mu_out)
    if (b <= 32) TQ_STAGE_CASE(1);
    else if (b <= 64) TQ_STAGE_CASE(2);
    else if (b <= 128) TQ_STAGE_CASE(4);
    else if (b <= 256) TQ_STAGE_CASE(8);
    else TQ_STAGE_CASE(16);
#undef TQ_STAGE_CASE
    TQ_LAUNCH_CHECK("atq_stage_kernel");
    return 0;
}

extern "C" int tq_aga_vector(const float* Hsrc, int64_t ldh, const int32_t* blk_idx, int64_t col0, int64_t b,
                             int mode, float* s1d, void* stream) {
    using namespace tq;
    TQ_CHECK_ARG(Hsrc && s1d && b > 0 && b <= 512, "tq_aga_vector: bad arguments");
    TQ_CHECK_ARG(mode == TQ_AGA_HESSIAN || mode == TQ_AGA_ACTIVATIONS, "tq_aga_vector: mode must be HESSIAN or ACTIVATIONS");
    return launch_aga_vector(Hsrc, ldh, blk_idx, col0, b, mode, s1d, nullptr, 0, nullptr, (cudaStream_t)stream);
}
