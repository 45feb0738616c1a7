// A9-A11: un-permute the sweep-order codes into original column positions (gptq.py:155), dequantise
// (gptq.py:201-230), and the 2-bit codec (utils.py:189-248).  Byte/integer work, HBM-bound: one thread
// per output byte group, 128-bit accesses on the contiguous side.
#include "common.cuh"

namespace tq {

// inverse of the sweep order: inv[perm[p]] = p
__global__ void __launch_bounds__(256)
invert_perm_kernel(const int32_t* __restrict__ perm, int m, int32_t* __restrict__ inv) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < m) inv[perm[p]] = p;
}

// Codes from sweep order back to ORIGINAL column positions (gptq.py:155), gather form: a CTA stages rows of the
// sweep-ordered codes in shared memory (16-byte loads), every thread assembles FOUR consecutive original columns from it and
// stores one 32-bit word -- coalesced on both sides of HBM.  (The scatter form it replaces stored single bytes at permuted
// addresses: 45 MB of partial-sector writes, 0.2-1.0 ms per linear.)
constexpr int UNP_ROWS = 4;
__global__ void __launch_bounds__(256)
unpermute_gather_kernel(const int8_t* __restrict__ Tperm, int n, int m, const int32_t* __restrict__ inv,
                        int8_t* __restrict__ Torig, float* __restrict__ Tf32) {
    extern __shared__ int8_t rows[];                       // [UNP_ROWS][mp], mp = m rounded up to 16
    const int mp = (m + 15) & ~15;
    const int r0 = blockIdx.x * UNP_ROWS;
    const int nr = min(UNP_ROWS, n - r0);
    const bool vec = ((m & 15) == 0) && ((reinterpret_cast<uintptr_t>(Tperm) & 15) == 0);
    for (int rr = 0; rr < nr; ++rr) {
        const int8_t* src = Tperm + (int64_t)(r0 + rr) * m;
        if (vec) {
            const int4* s4 = reinterpret_cast<const int4*>(src);
            int4* d4 = reinterpret_cast<int4*>(rows + rr * mp);
            for (int i = threadIdx.x; i < m / 16; i += blockDim.x) d4[i] = s4[i];
        } else {
            for (int i = threadIdx.x; i < m; i += blockDim.x) rows[rr * mp + i] = src[i];
        }
    }
    __syncthreads();
    const bool wvec = ((m & 3) == 0) && ((reinterpret_cast<uintptr_t>(Torig) & 3) == 0);
    for (int c0 = threadIdx.x * 4; c0 < m; c0 += blockDim.x * 4) {
        int src_pos[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) src_pos[k] = (c0 + k < m) ? inv[c0 + k] : 0;
        for (int rr = 0; rr < nr; ++rr) {
            const int8_t* row = rows + rr * mp;
            int8_t v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = row[src_pos[k]];
            const int64_t o = (int64_t)(r0 + rr) * m + c0;
            if (wvec) {
                const uint32_t w = (uint32_t)(uint8_t)v[0] | ((uint32_t)(uint8_t)v[1] << 8) | ((uint32_t)(uint8_t)v[2] << 16) |
                                   ((uint32_t)(uint8_t)v[3] << 24);
                *reinterpret_cast<uint32_t*>(Torig + o) = w;
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (c0 + k < m) Torig[o + k] = v[k];
            }
            if (Tf32) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (c0 + k < m) Tf32[o + k] = (float)v[k];
            }
        }
    }
}

__global__ void __launch_bounds__(256)
dequant_kernel(const float* __restrict__ alpha, const float* __restrict__ mu, int nb, const int8_t* __restrict__ Torig,
               int n, int m, const int32_t* __restrict__ perm, int block, float* __restrict__ Wq) {
    const int r = blockIdx.y;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < m; p += gridDim.x * blockDim.x) {
        const int k = p / block;
        const int c = perm[p];
        const float a = alpha[(int64_t)r * nb + k], u = mu[(int64_t)r * nb + k];
        Wq[(int64_t)r * m + c] = __fadd_rn(__fmul_rn(a, (float)Torig[(int64_t)r * m + c]), u);
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
pack2b_kernel(const T* __restrict__ in, int64_t count, uint8_t* __restrict__ out) {
    const int64_t nbytes = (count + 3) / 4;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nbytes; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t base = q * 4;
        unsigned byte = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t i = base + k;
            const unsigned code = (i < count) ? (unsigned)((int)in[i] + 1) & 3u : 0u;   // zero pad (utils.py:207-209)
            byte |= code << (2 * k);
        }
        out[q] = (uint8_t)byte;
    }
}

// vector path: 16 int8 codes -> one 32-bit word of packed output per thread
__global__ void __launch_bounds__(256)
pack2b_vec_kernel(const int4* __restrict__ in, int64_t groups, uint32_t* __restrict__ out) {
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
        const int4 v = in[g];
        const uint32_t wv[4] = {(uint32_t)v.x, (uint32_t)v.y, (uint32_t)v.z, (uint32_t)v.w};
        uint32_t word = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            // four int8 codes in wv[q]: keep 2 bits (-1 -> 3), +1 per byte (max 4, no carry), keep 2 bits again
            const uint32_t c = ((wv[q] & 0x03030303u) + 0x01010101u) & 0x03030303u;
            const uint32_t byte = (c & 3u) | ((c >> 6) & 0xCu) | ((c >> 12) & 0x30u) | ((c >> 18) & 0xC0u);
            word |= byte << (8 * q);
        }
        out[g] = word;
    }
}

__global__ void __launch_bounds__(256)
unpack2b_kernel(const uint8_t* __restrict__ in, int64_t count, int8_t* __restrict__ out) {
    const int64_t nbytes = (count + 3) / 4;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nbytes; q += (int64_t)gridDim.x * blockDim.x) {
        const unsigned byte = in[q];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t i = q * 4 + k;
            if (i < count) out[i] = (int8_t)((int)((byte >> (2 * k)) & 3u) - 1);
        }
    }
}

// vector path: one 32-bit word of packed input -> 16 int8 codes (one 128-bit store) per thread
__global__ void __launch_bounds__(256)
unpack2b_vec_kernel(const uint32_t* __restrict__ in, int64_t groups, int4* __restrict__ out) {
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t word = in[g];
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t byte = (word >> (8 * q)) & 0xFFu;
            // spread the four 2-bit codes of this byte over four bytes, then subtract 1 per byte (codes are 0,1,2: no borrow
            // crosses a byte because every byte is first OR-ed with 0x80 and the marker is removed afterwards)
            const uint32_t c = (byte & 3u) | ((byte & 0xCu) << 6) | ((byte & 0x30u) << 12) | ((byte & 0xC0u) << 18);
            o[q] = ((c | 0x80808080u) - 0x01010101u) ^ 0x80808080u;
        }
        out[g] = make_int4((int)o[0], (int)o[1], (int)o[2], (int)o[3]);
    }
}

static inline unsigned grid_for(int64_t work, int threads) {
    int64_t g = ceil_div(work, threads);
    const int64_t cap = (int64_t)sm_count() * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

// `inv_scratch`: m ints of device scratch for the inverse permutation
int launch_unpermute(const int8_t* Tperm, int64_t n, int64_t m, const int32_t* perm, int8_t* Torig, float* Tf32,
                     int32_t* inv_scratch, cudaStream_t st) {
    invert_perm_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, st>>>(perm, (int)m, inv_scratch);
    TQ_LAUNCH_CHECK("invert_perm_kernel");
    const int mp = (int)((m + 15) & ~(int64_t)15);
    const int smem = UNP_ROWS * mp;
    static std::atomic<unsigned long long> attr_mask{0};
    int dev;
    if (smem > 48 * 1024 && dyn_smem_pending(attr_mask, dev)) {
        TQ_CUDA(cudaFuncSetAttribute(unpermute_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        dyn_smem_done(attr_mask, dev);
    }
    TQ_CHECK_ARG(smem <= 200 * 1024, "unpermute: m = %lld is beyond the staged-row kernel's shared memory", (long long)m);
    unpermute_gather_kernel<<<(unsigned)ceil_div(n, UNP_ROWS), 256, smem, st>>>(Tperm, (int)n, (int)m, inv_scratch, Torig, Tf32);
    TQ_LAUNCH_CHECK("unpermute_gather_kernel");
    return 0;
}

}  // namespace tq

extern "C" int tq_unpermute_codes(const int8_t* Tperm, int64_t n, int64_t m, const int32_t* perm, int8_t* Torig,
                                  float* Tf32, void* stream) {
    using namespace tq;
    TQ_CHECK_ARG(Tperm && perm && Torig && n > 0 && m > 0 && n <= 65535 * 1024, "tq_unpermute_codes: bad arguments");
    TQ_CHECK_ARG(Tperm != Torig, "tq_unpermute_codes: in-place not supported");
    // the inverse permutation needs m ints of scratch: a stream-ordered temporary (the sweep passes its own workspace)
    cudaStream_t st = (cudaStream_t)stream;
    int32_t* inv = nullptr;
    TQ_CUDA(cudaMallocAsync(&inv, sizeof(int32_t) * m, st));
    const int rc = launch_unpermute(Tperm, n, m, perm, Torig, Tf32, inv, st);
    TQ_CUDA(cudaFreeAsync(inv, st));
    return rc;
}

extern "C" int tq_dequant(const float* alpha, const float* mu, int64_t nb, const int8_t* Torig, int64_t n, int64_t m,
                          const int32_t* perm, int64_t block, float* Wq, void* stream) {
    using namespace tq;
    TQ_CHECK_ARG(alpha && mu && Torig && perm && Wq && n > 0 && m > 0 && block > 0, "tq_dequant: bad arguments");
    TQ_CHECK_ARG(nb == ceil_div(m, block), "tq_dequant: nb must equal ceil(m / block)");
    dim3 grid((unsigned)(ceil_div(m, 256) < 64 ? ceil_div(m, 256) : 64), (unsigned)n);
    dequant_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(alpha, mu, (int)nb, Torig, (int)n, (int)m, perm, (int)block, Wq);
    TQ_LAUNCH_CHECK("dequant_kernel");
    return 0;
}

extern "C" int tq_pack2b(const int8_t* T, int64_t count, uint8_t* packed, void* stream) {
    using namespace tq;
    TQ_CHECK_ARG(T && packed && count >= 0, "tq_pack2b: bad arguments");
    if (count == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const bool aligned = ((reinterpret_cast<uintptr_t>(T) & 15) == 0) && ((reinterpret_cast<uintptr_t>(packed) & 3) == 0);
    const int64_t groups = aligned ? count / 16 : 0;
    if (groups > 0) {
        pack2b_vec_kernel<<<grid_for(groups, 256), 256, 0, st>>>(reinterpret_cast<const int4*>(T), groups,
                                                                 reinterpret_cast<uint32_t*>(packed));
        TQ_LAUNCH_CHECK("pack2b_vec_kernel");
    }
    const int64_t done = groups * 16;
    if (done < count) {
        pack2b_kernel<int8_t><<<grid_for((count - done + 3) / 4, 256), 256, 0, st>>>(T + done, count - done, packed + done / 4);
        TQ_LAUNCH_CHECK("pack2b_kernel");
    }
    return 0;
}

extern "C" int tq_pack2b_f32(const float* T, int64_t count, uint8_t* packed, void* stream) {
    using namespace tq;
    TQ_CHECK_ARG(T && packed && count >= 0, "tq_pack2b_f32: bad arguments");
    if (count == 0) return 0;
    pack2b_kernel<float><<<grid_for((count + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(T, count, packed);
    TQ_LAUNCH_CHECK("pack2b_kernel<float>");
    return 0;
}

extern "C" int tq_unpack2b(const uint8_t* packed, int64_t count, int8_t* T, void* stream) {
    using namespace tq;
    TQ_CHECK_ARG(T && packed && count >= 0, "tq_unpack2b: bad arguments");
    if (count == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const bool aligned = ((reinterpret_cast<uintptr_t>(T) & 15) == 0) && ((reinterpret_cast<uintptr_t>(packed) & 3) == 0);
    const int64_t groups = aligned ? count / 16 : 0;
    if (groups > 0) {
        unpack2b_vec_kernel<<<grid_for(groups, 256), 256, 0, st>>>(reinterpret_cast<const uint32_t*>(packed), groups,
                                                                   reinterpret_cast<int4*>(T));
        TQ_LAUNCH_CHECK("unpack2b_vec_kernel");
    }
    const int64_t done = groups * 16;
    if (done < count) {
        unpack2b_kernel<<<grid_for((count - done + 3) / 4, 256), 256, 0, st>>>(packed + done / 4, count - done, T + done);
        TQ_LAUNCH_CHECK("unpack2b_kernel");
    }
    return 0;
}
