// A4: SSR dynamic block selection.  Reference: reorder.py:36-61 (similarity of every remaining column
// to the mean of the remaining columns) and reorder.py:107-143 (top-k, remaining list).
//
// Stage 1 (HBM-bound, two passes over W[:, rem]):
//   ssr_rowmean_kernel : wbar[r] = mean_j W[r, rem_j]                       (reorder.py:52)
//   ssr_colstats_kernel: per row-chunk partials of  dot_j = sum_r W[r,j] wbar[r]  and
//                        sq_j = sum_r W[r,j]^2                               (reorder.py:55-59)
//   Partials are reduced in a fixed order in stage 2, so the selection is run-to-run deterministic;
//   a row-sharded caller all-reduces them between the stages.
// Stage 2 (one CTA): similarities, radix select of the `block` largest with ties going to the lower
//   position, bitonic sort of the winners into descending-similarity order (torch.topk's output
//   order, reorder.py:133-136), order-preserving compaction of the rest (reorder.py:139-141).
#include "common.cuh"

namespace tq {

constexpr int SSR_ROWS_PER_CHUNK = 128;
constexpr int SSR_COLS_PER_CTA = 128;
constexpr int SEL_THREADS = 1024;

__global__ void __launch_bounds__(256)
ssr_rowmean_kernel(const float* __restrict__ W, int64_t ldw, int n, const int32_t* __restrict__ rem_idx, int rem,
                   float* __restrict__ rowmean, float* __restrict__ rowsum) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const float* wrow = W + (int64_t)row * ldw;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int j = lane;
    for (; j + 96 < rem; j += 128) {
        s0 += wrow[rem_idx[j]];
        s1 += wrow[rem_idx[j + 32]];
        s2 += wrow[rem_idx[j + 64]];
        s3 += wrow[rem_idx[j + 96]];
    }
    for (; j < rem; j += 32) s0 += wrow[rem_idx[j]];
    const float s = warp_sum((s0 + s1) + (s2 + s3));
    if (lane == 0) {
        rowmean[row] = __fdiv_rn(s, (float)rem);
        if (rowsum) rowsum[row] = s;            // the single row-sum partial the next ATQ kernel starts from
    }
}

// grid: (ceil(rem/128), num_chunks); 128 threads, thread = one remaining column, loops over the chunk's rows
__global__ void __launch_bounds__(SSR_COLS_PER_CTA)
ssr_colstats_kernel(const float* __restrict__ W, int64_t ldw, int n, const int32_t* __restrict__ rem_idx, int rem,
                    const float* __restrict__ rowmean, float* __restrict__ partials) {
    __shared__ float wb[SSR_ROWS_PER_CHUNK];
    const int r0 = blockIdx.y * SSR_ROWS_PER_CHUNK;
    const int nr = min(SSR_ROWS_PER_CHUNK, n - r0);
    for (int i = threadIdx.x; i < nr; i += blockDim.x) wb[i] = rowmean[r0 + i];
    __syncthreads();
    const int j = blockIdx.x * SSR_COLS_PER_CTA + threadIdx.x;
    if (j >= rem) return;
    const float* wc = W + (int64_t)r0 * ldw + rem_idx[j];
    float dot = 0.f, sq = 0.f;
#pragma unroll 8
    for (int i = 0; i < nr; ++i) {
        const float v = wc[(int64_t)i * ldw];
        dot = fmaf(v, wb[i], dot);
        sq = fmaf(v, v, sq);
    }
    float* out = partials + (int64_t)blockIdx.y * 2 * rem;
    out[j] = dot;
    out[rem + j] = sq;
}

__device__ __forceinline__ uint32_t order_key(float f) {   // larger float -> larger key
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ float block_sum_1024(float v, float* red) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = (threadIdx.x < 32) ? red[threadIdx.x] : 0.f;
    if (threadIdx.x < 32) t = warp_sum(t);
    if (threadIdx.x == 0) red[0] = t;
    __syncthreads();
    t = red[0];
    __syncthreads();
    return t;
}

// exclusive scan of one int per thread over the 1024-thread CTA; returns the exclusive prefix, total in *total
__device__ int block_excl_scan_1024(int v, int* wsum, int* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = wsum[lane];
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += y;
        }
        wsum[lane] = winc - w;               // exclusive warp offsets
        if (lane == 31) wsum[32] = winc;     // grand total
    }
    __syncthreads();
    const int res = wsum[wid] + inc - v;
    *total = wsum[32];
    __syncthreads();
    return res;
}

// similarities (reorder.py:56-59) and order-preserving keys, one thread per remaining column; the per-chunk
// partials are folded in chunk order, so the result does not depend on the launch geometry
__global__ void __launch_bounds__(256)
ssr_sims_kernel(const float* __restrict__ partials, int num_chunks, const float* __restrict__ wbar_sq,
                const float* __restrict__ rowmean, int n, int rem, float* __restrict__ sims, uint32_t* __restrict__ keys) {
    // ||wbar||^2 (reorder.py:55): taken from the caller (row-sharded sweep: all-reduced) or summed here, by every CTA
    // in the same fixed order so all CTAs agree bit for bit
    __shared__ float red[8];
    __shared__ float s_msq;
    if (wbar_sq == nullptr) {
        float s = 0.f;
        for (int i = threadIdx.x; i < n; i += blockDim.x) s = fmaf(rowmean[i], rowmean[i], s);
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x < 32) {
            float v = (threadIdx.x < 8) ? red[threadIdx.x] : 0.f;
            v = warp_sum(v);
            if (threadIdx.x == 0) s_msq = v;
        }
    } else if (threadIdx.x == 0) {
        s_msq = wbar_sq[0];
    }
    __syncthreads();
    // 64 columns per CTA, four threads per column: thread (g, jj) folds chunks g, g + 4, ... (independent loads), thread
    // (0, jj) adds the four sums in a fixed order -- the result does not depend on the launch geometry, and the dependent
    // load chain is a quarter of what one thread per column had (grid 4x larger: the kernel was latency-bound on 28 CTAs)
    __shared__ float part[4][64][2];
    const int jj = threadIdx.x & 63, g = threadIdx.x >> 6;
    const int j = blockIdx.x * 64 + jj;
    float dot = 0.f, sq = 0.f;
    if (j < rem) {
#pragma unroll 4
        for (int c = g; c < num_chunks; c += 4) {
            dot += partials[(int64_t)c * 2 * rem + j];
            sq += partials[(int64_t)c * 2 * rem + rem + j];
        }
    }
    part[g][jj][0] = dot;
    part[g][jj][1] = sq;
    __syncthreads();
    if (g != 0 || j >= rem) return;
    dot = ((part[0][jj][0] + part[1][jj][0]) + part[2][jj][0]) + part[3][jj][0];
    sq = ((part[0][jj][1] + part[1][jj][1]) + part[2][jj][1]) + part[3][jj][1];
    const float mnorm = fmaxf(sqrtf(s_msq), kTiny);
    const float cn = fmaxf(sqrtf(sq), kTiny);
    const float sim = __fdiv_rn(__fdiv_rn(dot, cn), mnorm);
    sims[j] = sim;
    keys[j] = order_key(sim);
}

__global__ void __launch_bounds__(SEL_THREADS)
ssr_select_kernel(const uint32_t* __restrict__ keys, const int32_t* __restrict__ rem_idx, int rem, int block,
                  int32_t* __restrict__ blk_idx, int32_t* __restrict__ new_rem_idx) {
    __shared__ int wsum[33];
    __shared__ int whist[SEL_THREADS / 32][256];   // one histogram per warp: similarities of one layer share their leading
                                                   // key bytes, so a single shared histogram serialises ~rem atomics on ONE
                                                   // address per pass (measured 26 us per selection at rem = 11008)
    __shared__ uint32_t s_prefix;
    __shared__ int s_need;
    __shared__ unsigned long long win[1024];   // selected (key, position), sorted at the end
    const int tid = threadIdx.x;

    // radix select: find the key of the block-th largest element, 8 bits at a time from the top
    uint32_t prefix = 0, pmask = 0;
    int need = block;                         // how many still to take among keys matching the prefix
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = tid; i < (SEL_THREADS / 32) * 256; i += SEL_THREADS) (&whist[0][0])[i] = 0;
        __syncthreads();
        {
            int* mine = whist[tid >> 5];
            for (int j = tid; j < rem; j += SEL_THREADS) {
                const uint32_t k = keys[j];
                const bool in = (k & pmask) == prefix;
                const int bin = (int)((k >> shift) & 255);
                // integer counts: order-free, deterministic.  When every matching lane of the warp hits the same bin
                // (the leading bytes of one layer's similarities) one lane adds the population count; otherwise the
                // lanes add to the warp's private histogram (conflicts are at most 32-way and rare in the low bytes)
                const unsigned act = __activemask();
                const unsigned inm = __ballot_sync(act, in);
                if (inm == 0) continue;
                const int lead = __ffs(inm) - 1;
                const int bin0 = __shfl_sync(act, bin, lead);
                if (__all_sync(act, !in || bin == bin0)) {
                    if ((int)(tid & 31) == lead) atomicAdd(&mine[bin0], __popc(inm));
                } else if (in) {
                    atomicAdd(&mine[bin], 1);
                }
            }
        }
        __syncthreads();
        // thread t < 256 owns bin 255 - t: the exclusive prefix of the reversed histogram is the number of keys in HIGHER
        // bins, so the bin holding the need-th largest key is found by all bins at once (no serial walk by one thread)
        int mycount = 0;
        if (tid < 256) {
#pragma unroll 8
            for (int w = 0; w < SEL_THREADS / 32; ++w) mycount += whist[w][255 - tid];
        }
        int tot_unused;
        const int above = block_excl_scan_1024(mycount, wsum, &tot_unused);
        if (tid < 256 && above < need && need <= above + mycount) {
            s_prefix = prefix | ((uint32_t)(255 - tid) << shift);
            s_need = need - above;
        }
        __syncthreads();
        prefix = s_prefix;
        need = s_need;
        pmask |= (255u << shift);
        __syncthreads();
    }
    const uint32_t kth = prefix;              // key of the block-th largest; take all > kth and the first `need` == kth

    // each thread owns a contiguous span of positions so scans preserve order
    const int per = (rem + SEL_THREADS - 1) / SEL_THREADS;
    const int j0 = min(rem, tid * per), j1 = min(rem, j0 + per);
    int tot;
    int eq = 0;
    for (int j = j0; j < j1; ++j) eq += (keys[j] == kth);
    const int eq_excl = block_excl_scan_1024(eq, wsum, &tot);      // ties at lower positions
    int nsel = 0;
    {
        int e = eq_excl;
        for (int j = j0; j < j1; ++j) {
            const uint32_t k = keys[j];
            nsel += (k > kth) || (k == kth && (e++ < need));
        }
    }
    const int sel_excl = block_excl_scan_1024(nsel, wsum, &tot);   // winners at lower positions
    {
        int e = eq_excl, s = sel_excl;
        for (int j = j0; j < j1; ++j) {
            const uint32_t k = keys[j];
            const bool take = (k > kth) || (k == kth && (e++ < need));
            if (take) {
                // ascending on (~key, position)  ==  descending similarity, ties -> lower position
                win[s++] = ((unsigned long long)(~k) << 32) | (uint32_t)j;
            } else {
                new_rem_idx[j - s] = rem_idx[j];                   // order-preserving compaction
            }
        }
    }
    // pad to a power of two and bitonic-sort the winners
    int npow = 1;
    while (npow < block) npow <<= 1;
    for (int i = block + tid; i < npow; i += SEL_THREADS) win[i] = ~0ull;
    __syncthreads();
    for (int k = 2; k <= npow; k <<= 1) {
        for (int jj = k >> 1; jj > 0; jj >>= 1) {
            for (int i = tid; i < npow; i += SEL_THREADS) {
                const int ixj = i ^ jj;
                if (ixj > i) {
                    const unsigned long long a = win[i], bb = win[ixj];
                    const bool up = ((i & k) == 0);
                    if ((a > bb) == up) { win[i] = bb; win[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < block; i += SEL_THREADS) blk_idx[i] = rem_idx[(uint32_t)(win[i] & 0xffffffffu)];
}

// row-sharded sweep: fold this rank's per-chunk partials into one [2*rem + 1] vector (dot, sq, sum wbar^2),
// in a fixed order, ready for the all-reduce across row shards
__global__ void __launch_bounds__(256)
ssr_fold_kernel(const float* __restrict__ partials, int num_chunks, const float* __restrict__ rowmean, int n, int rem,
                float* __restrict__ folded) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < 2 * rem) {
        float s = 0.f;
        for (int c = 0; c < num_chunks; ++c) s += partials[(int64_t)c * 2 * rem + j];
        folded[j] = s;
    }
    if (blockIdx.x == 0) {
        __shared__ float red[8];
        float s = 0.f;
        for (int i = threadIdx.x; i < n; i += blockDim.x) s = fmaf(rowmean[i], rowmean[i], s);
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x < 32) {
            float v = (threadIdx.x < 8) ? red[threadIdx.x] : 0.f;
            v = warp_sum(v);
            if (threadIdx.x == 0) folded[2 * rem] = v;
        }
    }
}

int launch_ssr_fold(const float* partials, int64_t num_chunks, const float* rowmean, int64_t n, int64_t rem,
                    float* folded, cudaStream_t st) {
    ssr_fold_kernel<<<(unsigned)ceil_div(2 * rem, 256), 256, 0, st>>>(partials, (int)num_chunks, rowmean, (int)n,
                                                                      (int)rem, folded);
    TQ_LAUNCH_CHECK("ssr_fold_kernel");
    return 0;
}

int launch_ssr_stats(const float* W, int64_t ldw, int64_t n, const int32_t* rem_idx, int64_t rem, float* rowmean,
                     float* partials, float* rowsum, cudaStream_t st) {
    ssr_rowmean_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, st>>>(W, ldw, (int)n, rem_idx, (int)rem, rowmean, rowsum);
    TQ_LAUNCH_CHECK("ssr_rowmean_kernel");
    dim3 grid((unsigned)ceil_div(rem, SSR_COLS_PER_CTA), (unsigned)ceil_div(n, SSR_ROWS_PER_CHUNK));
    ssr_colstats_kernel<<<grid, SSR_COLS_PER_CTA, 0, st>>>(W, ldw, (int)n, rem_idx, (int)rem, rowmean, partials);
    TQ_LAUNCH_CHECK("ssr_colstats_kernel");
    return 0;
}

int launch_ssr_select(const float* partials, int64_t num_chunks, const float* rowmean, int64_t n,
                      const float* wbar_sq_dev, const int32_t* rem_idx, int64_t rem, int64_t block,
                      int32_t* blk_idx, int32_t* new_rem_idx, float* sims, uint32_t* keys, cudaStream_t st) {
    ssr_sims_kernel<<<(unsigned)ceil_div(rem, 64), 256, 0, st>>>(partials, (int)num_chunks, wbar_sq_dev, rowmean, (int)n,
                                                                   (int)rem, sims, keys);
    TQ_LAUNCH_CHECK("ssr_sims_kernel");
    ssr_select_kernel<<<1, SEL_THREADS, 0, st>>>(keys, rem_idx, (int)rem, (int)block, blk_idx, new_rem_idx);
    TQ_LAUNCH_CHECK("ssr_select_kernel");
    return 0;
}

}  // namespace tq

extern "C" int64_t tq_ssr_num_chunks(int64_t n) { return tq::ceil_div(n, tq::SSR_ROWS_PER_CHUNK); }

extern "C" int tq_ssr_stats(const float* W, int64_t ldw, int64_t n, const int32_t* rem_idx, int64_t rem,
                            float* rowmean, float* partials, void* stream) {
    using namespace tq;
    TQ_CHECK_ARG(W && rem_idx && rowmean && partials && n > 0 && rem > 0, "tq_ssr_stats: bad arguments");
    return launch_ssr_stats(W, ldw, n, rem_idx, rem, rowmean, partials, nullptr, (cudaStream_t)stream);
}

extern "C" int tq_ssr_select(const float* partials, int64_t num_chunks, const float* rowmean, int64_t n,
                             const float* wbar_sq_dev, const int32_t* rem_idx, int64_t rem, int64_t block,
                             int32_t* blk_idx, int32_t* new_rem_idx, float* sims, void* stream) {
    using namespace tq;
    TQ_CHECK_ARG(partials && rem_idx && blk_idx && new_rem_idx && sims, "tq_ssr_select: null pointer (sims doubles as key scratch)");
    TQ_CHECK_ARG(rem > block && block >= 1 && block <= 1024,
                 "tq_ssr_select: needs rem > block and block <= 1024; with rem <= block the block is the whole "
                 "remaining list (reorder.py:125-126)");
    TQ_CHECK_ARG(rowmean != nullptr || wbar_sq_dev != nullptr, "tq_ssr_select: need rowmean or wbar_sq_dev");
    // the key scratch (and 2 floats for ||wbar||^2) live behind the sims array: callers give sims room for 2*rem + 2 floats
    return launch_ssr_select(partials, num_chunks, rowmean, n, wbar_sq_dev, rem_idx, rem, block, blk_idx,
                             new_rem_idx, sims, reinterpret_cast<uint32_t*>(sims + rem), (cudaStream_t)stream);
}
