// Inline-PTX wrappers shared by the tcgen05 kernels: mbarrier, TMA (bulk tensor copies and reduce), tcgen05
// MMA / commit / TMEM load, proxy fences.  sm_100a only.
#pragma once

#include "common.cuh"

#include <cuda.h>

namespace tq {

constexpr unsigned long long TC_SPIN_LIMIT = 4000000000ull;   // ~2 s of SM clocks: trap instead of hanging the box

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const unsigned long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > TC_SPIN_LIMIT) {
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// whole-warp calls; ncols is a power of two in [32, 512]
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem_addr, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem_addr), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t tmem_base, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_alloc_512(uint32_t slot_smem_addr) { tmem_alloc(slot_smem_addr, 512); }
__device__ __forceinline__ void tmem_dealloc_512(uint32_t tmem_base) { tmem_dealloc(tmem_base, 512); }

// ---- host: cuTensorMapEncodeTiled through the runtime's driver entry point (no libcuda link dependency)
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_tmapEncodeTiled tmap_encode_fn();
// 2-D row-major tensor [rows, cols] (cols contiguous, row stride ld elements), box [box_rows, box_cols], 128B swizzle
int make_tmap_2d(CUtensorMap* map, CUtensorMapDataType dt, int elem_bytes, const void* base, int64_t rows, int64_t cols,
                 int64_t ld, int box_rows, int box_cols, const char* what);


// ---- split-TF32 GEMM (gemm_tc.cu): C (op)= A B' with A [M, K], B [N, K] K-major fp32 given as (hi, lo) pairs
struct GemmOperands {
    CUtensorMap ah, al, bh, bl;     // TMA descriptors of the four operand arrays
    int64_t K;
};
struct GemmC {                      // output matrix of the TMA epilogue (gemm_cmap_encode)
    CUtensorMap map;
    float* base;
    int64_t ld;
    bool valid;
};
enum GxMode {
    GX_FEEDBACK = 0,      // all tiles; C[r, colmap(j)] -= acc
    GX_SUB_LOWER = 1,     // tiles bi >= bj; C[r, c] -= acc for c <= r  (potrf trailing update)
    GX_SUB_RECT = 2,      // all tiles; C[r, c] -= acc                  (trtri update)
    GX_STORE_UPPER = 3    // tiles bi <= bj, K range starts at the tile's first column; C[r, c] = acc (lauum)
};
int gemm_operands_encode(GemmOperands* ops, const float* Ah, const float* Al, int64_t lda, int64_t Mmax, const float* Bh,
                         const float* Bl, int64_t ldb, int64_t Nmax, int64_t K);
int launch_gemm_tf32x3_rows(int mode, float* C, int64_t ldc, int64_t M, int64_t N, const GemmOperands* ops, int64_t a_row0,
                            int64_t b_row0, const int32_t* col_idx, int64_t col0, cudaStream_t st);
int launch_gemm_tf32x3_ops(int mode, float* C, int64_t ldc, int64_t M, int64_t N, const GemmOperands* ops,
                           const int32_t* col_idx, int64_t col0, cudaStream_t st);
int gemm_cmap_encode(GemmC* c, float* base, int64_t rows, int64_t cols, int64_t ld);
int launch_gemm_tf32x3_at(int mode, const GemmC* cmap, int64_t c_row0, int64_t c_col0, int64_t M, int64_t N,
                          const GemmOperands* ops, int64_t a_row0, int64_t b_row0, cudaStream_t st);
int launch_gemm_tf32x3(int mode, float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, const float* Ah, const float* Al,
                       int64_t lda, const float* Bh, const float* Bl, int64_t ldb, const int32_t* col_idx, int64_t col0,
                       cudaStream_t st);
int launch_split(const float* in, int64_t ld_in, int64_t rows, int64_t cols, float* hi, float* lo, int64_t ld_out,
                 int transpose, cudaStream_t st);
int launch_feedback_coef(const float* Hinv, int64_t ldh, const int32_t* blk_idx, int64_t blk0, int64_t b,
                         const int32_t* rem_idx, int64_t rem0, int64_t rem, float* ch, float* cl, int64_t ldb,
                         float* csum_part, cudaStream_t st);
int launch_gemm_feedback_stats(float* C, int64_t ldc, int64_t M, int64_t N, const GemmOperands* ops, const int32_t* col_idx,
                               int64_t col0, const float* wbar, float* stat_partials, float* rowsum_part, cudaStream_t st);

}  // namespace tq
