// A2 (fp32 path) + A3 prologue: H += X'X on the CUDA cores, symmetrise, scale + damp.
// Reference: gptq.py:59-76 (add_batch), gptq.py:94-98 (normalise + damp).
#include "gemm_simt.cuh"

namespace tq {

// One CTA per (upper-triangular 128x128 tile of H) x (token split).  Both operands are slabs of the
// same row-major X (token-major), i.e. MN-contiguous: consecutive threads read consecutive columns.
template <typename T>
__global__ void __launch_bounds__(GT_THREADS, 2)
hessian_ffma_kernel(float* __restrict__ H, int64_t ldh, const T* __restrict__ X, int64_t Nt, int m,
                    int64_t ldx, int64_t tokens_per_split, int use_atomics) {
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bi > bj) return;
    __shared__ GemmSmem sm;
    const int i0 = bi * GT_M, j0 = bj * GT_N;
    const int64_t t0 = (int64_t)blockIdx.z * tokens_per_split;
    const int64_t t1 = min(Nt, t0 + tokens_per_split);
    const T* Xs = X + t0 * ldx;
    const int kspan = (int)(t1 - t0);
    float acc[8][8];
    gemm_tile<WALK_MN, WALK_MN>(
        sm, 0, kspan,
        [&](int k, int i) { return (i0 + i < m) ? to_f32(Xs[(int64_t)k * ldx + i0 + i]) : 0.f; },
        [&](int k, int j) { return (j0 + j < m) ? to_f32(Xs[(int64_t)k * ldx + j0 + j]) : 0.f; }, acc);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = i0 + gt_row(ty, i);
        if (r >= m) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = j0 + gt_col(tx, j);
            if (c >= m) continue;
            float* p = H + (int64_t)r * ldh + c;
            if (use_atomics) atomicAdd(p, acc[i][j]);
            else *p += acc[i][j];
        }
    }
}

// lower := upper, 32x32 tiles through shared memory so both sides stay coalesced
__global__ void symmetrize_kernel(float* __restrict__ H, int64_t ldh, int m) {
    const int by = blockIdx.y, bx = blockIdx.x;
    if (by < bx) return;                    // only tiles on or below the diagonal are written
    __shared__ float s[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
    // source tile = (bx, by): rows bx*32.., cols by*32..
    for (int r = ty; r < 32; r += 8) {
        const int gr = bx * 32 + r, gc = by * 32 + tx;
        s[r][tx] = (gr < m && gc < m) ? H[(int64_t)gr * ldh + gc] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int gr = by * 32 + r, gc = bx * 32 + tx;      // destination element (gr, gc), gr >= gc region
        if (gr < m && gc < m && gr > gc) H[(int64_t)gr * ldh + gc] = s[tx][r];
    }
}

__global__ void diag_damp_kernel(const float* __restrict__ Hraw, int m, float ns, float percdamp,
                                 float* __restrict__ out) {
    __shared__ float red[32];
    float s = 0.f;
    for (int i = threadIdx.x; i < m; i += blockDim.x) s += __fdiv_rn(Hraw[(int64_t)i * m + i], ns);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) out[0] = __fmul_rn(percdamp, __fdiv_rn(v, (float)m));
    }
}

// Hd[i][j] = Hraw[min][max] / ns (+ damp on the diagonal); reads only the upper triangle of Hraw
__global__ void finalize_kernel(float* __restrict__ Hd, const float* __restrict__ Hraw, int m, float ns,
                                const float* __restrict__ damp) {
    const int by = blockIdx.y, bx = blockIdx.x;
    __shared__ float s[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int sy = min(by, bx), sx = max(by, bx);
    for (int r = ty; r < 32; r += 8) {
        const int gr = sy * 32 + r, gc = sx * 32 + tx;
        s[r][tx] = (gr < m && gc < m) ? Hraw[(int64_t)gr * m + gc] : 0.f;
    }
    __syncthreads();
    const float d = damp[0];
    for (int r = ty; r < 32; r += 8) {
        const int gr = by * 32 + r, gc = bx * 32 + tx;
        if (gr >= m || gc >= m) continue;
        float v;
        if (by < bx) v = s[r][tx];
        else if (by > bx) v = s[tx][r];
        else v = (r <= tx) ? s[r][tx] : s[tx][r];
        v = __fdiv_rn(v, ns);
        if (gr == gc) v = __fadd_rn(v, d);
        Hd[(int64_t)gr * m + gc] = v;
    }
}

template <typename T>
static int launch_hessian_ffma(float* H, int64_t ldh, const void* X, int64_t Nt, int64_t m, int64_t ldx,
                               cudaStream_t st) {
    const int nt = (int)ceil_div(m, GT_M);
    const int64_t tiles = (int64_t)nt * (nt + 1) / 2;
    int64_t splits = ceil_div((int64_t)sm_count() * 4, tiles);
    const int64_t max_splits = ceil_div(Nt, 512);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    int64_t per = ceil_div(Nt, splits);
    per = ceil_div(per, GT_K) * GT_K;
    splits = ceil_div(Nt, per);
    dim3 grid(nt, nt, (unsigned)splits);
    hessian_ffma_kernel<T><<<grid, GT_THREADS, 0, st>>>(H, ldh, (const T*)X, Nt, (int)m, ldx, per,
                                                        splits > 1 ? 1 : 0);
    TQ_LAUNCH_CHECK("hessian_ffma_kernel");
    return 0;
}

int hessian_accum_ffma(float* H, int64_t ldh, const void* X, int64_t Nt, int64_t m, int64_t ldx, int dtype,
                       cudaStream_t st) {
    switch (dtype) {
        case TQ_F32:  return launch_hessian_ffma<float>(H, ldh, X, Nt, m, ldx, st);
        case TQ_F16:  return launch_hessian_ffma<__half>(H, ldh, X, Nt, m, ldx, st);
        case TQ_BF16: return launch_hessian_ffma<__nv_bfloat16>(H, ldh, X, Nt, m, ldx, st);
    }
    set_error("tq_hessian_accum: unknown dtype %d", dtype);
    return TQ_E_BADARG;
}

int hessian_accum_tcgen05(float* H, int64_t ldh, const void* X, int64_t Nt, int64_t m, int64_t ldx, int dtype,
                          cudaStream_t st);   // hessian_tc.cu

}  // namespace tq

extern "C" int tq_hessian_accum(float* H, int64_t ldh, const void* X, int64_t Nt, int64_t m, int64_t ldx,
                                int dtype, int path, void* stream) {
    using namespace tq;
    TQ_CHECK_ARG(H && X, "tq_hessian_accum: null pointer");
    TQ_CHECK_ARG(m > 0 && m < (1 << 30) && ldh >= m && ldx >= m, "tq_hessian_accum: bad shape m=%lld ldh=%lld ldx=%lld",
                 (long long)m, (long long)ldh, (long long)ldx);
    TQ_CHECK_ARG(Nt >= 0, "tq_hessian_accum: negative token count");
    if (Nt == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const bool tc_ok = (dtype == TQ_F16 || dtype == TQ_BF16) && (m % 8 == 0) && (ldx % 8 == 0) && (ldh % 4 == 0) &&
                       ((reinterpret_cast<uintptr_t>(X) & 15) == 0) && ((reinterpret_cast<uintptr_t>(H) & 15) == 0);
    if (path == TQ_HESS_TCGEN05 && !tc_ok) {
        set_error("tq_hessian_accum: tcgen05 path needs f16/bf16, m%%8==0, ldx%%8==0, ldh%%4==0, 16B-aligned pointers");
        return TQ_E_UNSUPPORTED;
    }
    if (path == TQ_HESS_TCGEN05 || (path == TQ_HESS_AUTO && tc_ok))
        return hessian_accum_tcgen05(H, ldh, X, Nt, m, ldx, dtype, st);
    TQ_CHECK_ARG(path == TQ_HESS_AUTO || path == TQ_HESS_FFMA, "tq_hessian_accum: unknown path %d", path);
    return hessian_accum_ffma(H, ldh, X, Nt, m, ldx, dtype, st);
}

extern "C" int tq_symmetrize(float* H, int64_t ldh, int64_t m, void* stream) {
    using namespace tq;
    TQ_CHECK_ARG(H && m > 0 && ldh >= m, "tq_symmetrize: bad arguments");
    const int nt = (int)ceil_div(m, 32);
    symmetrize_kernel<<<dim3(nt, nt), dim3(32, 8), 0, (cudaStream_t)stream>>>(H, ldh, (int)m);
    TQ_LAUNCH_CHECK("symmetrize_kernel");
    return 0;
}

extern "C" int tq_hessian_finalize(float* Hd, const float* Hraw, int64_t m, double nsamples, double percdamp,
                                   float* scratch, void* stream) {
    using namespace tq;
    TQ_CHECK_ARG(Hd && Hraw && scratch && m > 0, "tq_hessian_finalize: bad arguments");
    TQ_CHECK_ARG(nsamples > 0, "tq_hessian_finalize: nsamples must be positive (add_batch never called?)");
    TQ_CHECK_ARG(Hd != Hraw, "tq_hessian_finalize: Hd must not alias Hraw");
    cudaStream_t st = (cudaStream_t)stream;
    diag_damp_kernel<<<1, 1024, 0, st>>>(Hraw, (int)m, (float)nsamples, (float)percdamp, scratch);
    TQ_LAUNCH_CHECK("diag_damp_kernel");
    const int nt = (int)ceil_div(m, 32);
    finalize_kernel<<<dim3(nt, nt), dim3(32, 8), 0, st>>>(Hd, Hraw, (int)m, (float)nsamples, scratch);
    TQ_LAUNCH_CHECK("finalize_kernel");
    return 0;
}
