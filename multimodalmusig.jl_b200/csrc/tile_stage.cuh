// tile_stage.cuh -- dense count tiles and their asynchronous staging (sm_100a: cp.async.bulk + mbarrier).
//
// The tile kernels (theta_tile.cuh, lda_tile.cuh) work on 32 samples x V terms at a time.  With CSR records
// every tile starts by clearing a dense 32 x V tile in shared memory and scattering the tile's records into
// it: two passes over shared memory and a dependent chain rowptr -> records -> scatter of global loads at the
// head of every tile (the `long_scoreboard` stall of profiles/r02e_summary.md).  Mutation-count matrices are
// dense (79 % of the cells at BASELINE's shapes), so a modality whose density is above kDenseFrac also keeps
// its counts as DENSE TILES in HBM -- int32 [ceil(D / 32)][32][V], 4 bytes per cell against 8 per record -- and
// a tile's counts are ONE contiguous block: a single bulk copy (TMA engine, SASS UBLKCP) brings the next tile's
// counts into a staging buffer while the current tile computes, and completion is signalled on an mbarrier.
// No clear, no scatter, no row pointers; the arithmetic on the counts is unchanged (the same n, the same order).
#pragma once
#include <cstdint>

namespace mmsig {

constexpr double kDenseFrac = 0.40;     // nnz / (D V) above which a modality also keeps dense tiles

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// one arrival that also announces `bytes` of asynchronous copies to come
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// global -> shared bulk copy (16-byte aligned, size a multiple of 16), completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// generic-proxy reads of the staging buffer (by every thread, ordered by the preceding __syncthreads) before the
// asynchronous proxy overwrites it
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// thread 0 of the block: start the copy of tile `tile`'s counts (32 x V int32, contiguous) into `nt`
__device__ __forceinline__ void stage_tile(const int *cnt, long long tile, int V, int *nt, uint64_t *bar) {
    const unsigned bytes = 32u * (unsigned)V * 4u;
    fence_proxy_async();
    mbar_expect_tx(bar, bytes);
    bulk_g2s(nt, cnt + (size_t)tile * 32 * V, bytes, bar);
}

// CSR records -> dense tiles, warp per sample (the target range has been zeroed).  rowptr is the view of the first
// sample of the range, cnt its row.
__global__ void __launch_bounds__(256) k_densify(const long long *__restrict__ rowptr, const int2 *__restrict__ rec, long long D,
                                                 int V, int *__restrict__ cnt) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long nw = (long long)gridDim.x * 8;
    for (long long d = (long long)blockIdx.x * 8 + warp; d < D; d += nw) {
        const long long beg = rowptr[d], end = rowptr[d + 1];
        for (long long w = beg + lane; w < end; w += 32) {
            const int2 r = rec[w];
            cnt[d * V + (r.x & 0xffff)] = r.y;
        }
    }
}

}  // namespace mmsig
