// ingest_kernels.cuh -- format_counts_mmctm / _ctm / _lda on the device (reference src/utils.jl:1-36).
//
// make_count_matrix (src/utils.jl:1-7) turns one sample's dense count vector into the nnz x 2
// matrix [term, count] of its entries > 0 (entries <= 0 are dropped: `findall(counts .> 0)`).
// Here the whole dense V x D matrix of one modality becomes the CSR the E-step streams
// (rowptr int64[D+1], rec int2[nnz] = (term 0-based, count)) in two passes over the dense data:
//   k_dense_panel<.., false>   per sample: number of entries > 0 (-> rowptr[d+1]) and their sum (-> N_dm)
//   k_scan_*                   rowptr = exclusive prefix sum of the row sizes (three small kernels)
//   k_dense_panel<.., true>    per sample: the (term, count) records in ascending term order
// (k_dense_count / k_dense_fill are the row-at-a-time versions, kept for V too large to stage)
// Pure integer / byte work, HBM-bound: algorithmic bytes = 2 * elem * V * D (dense read twice)
// + 8 * nnz (records) + 24 * D (row sizes written, scanned, read).
// Layouts: TERM_MAJOR dense[v * D + d] (the TSV files: one line per term; a C-order (V, D) array):
// thread <-> sample, so a warp reads 32 consecutive samples of one term per load;
// SAMPLE_MAJOR dense[d * V + v] (Julia's column-major V x D Matrix{Int}, one DataFrame column per
// sample): warp <-> sample, lanes over terms, ballot / popc compaction.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mmsig {

constexpr int SCAN_ITEMS = 8;                  // per thread
constexpr int SCAN_BLOCK = 256 * SCAN_ITEMS;   // row sizes per block of the local scan

// flags[0] bit 3: a count does not fit int32.  total: Σ of the counts kept.
template <typename T>
__global__ void __launch_bounds__(256) k_dense_count(const T *__restrict__ dense, long long D, int V, int layout,
                                                     long long *rowptr, double *N, int n_stride, int n_off, int *flags,
                                                     unsigned long long *total) {
    const int lane = threadIdx.x & 31;
    unsigned long long tot_all = 0;
    int bad = 0;
    if (layout == 0) {
        const long long nthr = (long long)gridDim.x * blockDim.x;
        for (long long d = (long long)blockIdx.x * blockDim.x + threadIdx.x; d < D; d += nthr) {
            const T *col = dense + d;
            long long nnz = 0;
            unsigned long long tot = 0;
#pragma unroll 8
            for (int v = 0; v < V; ++v) {
                const long long x = (long long)col[(size_t)v * D];
                if (x > 0) { ++nnz; tot += (unsigned long long)x; }
                if (x > 2147483647LL) bad = 8;
            }
            rowptr[d + 1] = nnz;
            if (N) N[d * n_stride + n_off] = (double)tot;
            tot_all += tot;
        }
    } else {
        const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
        for (long long d = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); d < D; d += nw) {
            const T *row = dense + (size_t)d * V;
            int nnz = 0;
            unsigned long long tot = 0;
            for (int v = lane; v < V; v += 32) {
                const long long x = (long long)row[v];
                if (x > 0) { ++nnz; tot += (unsigned long long)x; }
                if (x > 2147483647LL) bad = 8;
            }
            for (int off = 16; off >= 1; off >>= 1) {
                nnz += __shfl_xor_sync(0xffffffffu, nnz, off);
                tot += __shfl_xor_sync(0xffffffffu, tot, off);
            }
            if (lane == 0) {
                rowptr[d + 1] = nnz;
                if (N) N[d * n_stride + n_off] = (double)tot;
                tot_all += tot;
            }
        }
    }
    for (int off = 16; off >= 1; off >>= 1) {
        tot_all += __shfl_xor_sync(0xffffffffu, tot_all, off);
        bad |= __shfl_xor_sync(0xffffffffu, bad, off);
    }
    if (lane == 0) {
        if (tot_all) atomicAdd(total, tot_all);
        if (bad) atomicOr(flags, bad);
    }
}

// inclusive scan of x[1..n] (x[0] stays 0), block-local part: SCAN_BLOCK values per block
__global__ void __launch_bounds__(256) k_scan_local(long long *x, long long n, long long *block_sum) {
    __shared__ long long wsum[8];
    const long long base = 1 + (long long)blockIdx.x * SCAN_BLOCK + (long long)threadIdx.x * SCAN_ITEMS;
    long long v[SCAN_ITEMS];
    long long s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i <= n) ? x[base + i] : 0;
        s += v[i];
        v[i] = s;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long inc = s;
    for (int off = 1; off < 32; off <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    long long woff = 0;
    for (int w = 0; w < warp; ++w) woff += wsum[w];
    const long long excl = woff + inc - s;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (base + i <= n) x[base + i] = v[i] + excl;
    if (threadIdx.x == 255) block_sum[blockIdx.x] = woff + inc;
}
// exclusive scan of the block sums, one block
__global__ void __launch_bounds__(1024) k_scan_sums(long long *block_sum, int nb) {
    __shared__ long long wsum[32];
    __shared__ long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int b0 = 0; b0 < nb; b0 += 1024) {
        const int i = b0 + threadIdx.x;
        const long long v = i < nb ? block_sum[i] : 0;
        long long inc = v;
        for (int off = 1; off < 32; off <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += t;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        long long woff = carry_s;
        for (int w = 0; w < warp; ++w) woff += wsum[w];
        if (i < nb) block_sum[i] = woff + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = woff + inc;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256) k_scan_add(long long *x, long long n, const long long *block_off) {
    const long long off = block_off[blockIdx.x];
    const long long base = 1 + (long long)blockIdx.x * SCAN_BLOCK + (long long)threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (base + i <= n) x[base + i] += off;
}

template <typename T>
__global__ void __launch_bounds__(256) k_dense_fill(const T *__restrict__ dense, long long D, int V, int layout,
                                                    const long long *__restrict__ rowptr, int2 *rec, int tag) {
    if (layout == 0) {
        const long long nthr = (long long)gridDim.x * blockDim.x;
        for (long long d = (long long)blockIdx.x * blockDim.x + threadIdx.x; d < D; d += nthr) {
            const T *col = dense + d;
            long long w = rowptr[d];
#pragma unroll 8
            for (int v = 0; v < V; ++v) {
                const long long x = (long long)col[(size_t)v * D];
                if (x > 0) rec[w++] = make_int2(tag ? (v | (int)((d & 31) << 16)) : v, (int)x);
            }
        }
    } else {
        const int lane = threadIdx.x & 31;
        const unsigned lt = (1u << lane) - 1u;
        const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
        for (long long d = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); d < D; d += nw) {
            const T *row = dense + (size_t)d * V;
            long long w = rowptr[d];
            for (int v0 = 0; v0 < V; v0 += 32) {
                const int v = v0 + lane;
                const long long x = v < V ? (long long)row[v] : 0;
                const unsigned m = __ballot_sync(0xffffffffu, x > 0);
                if (x > 0) rec[w + __popc(m & lt)] = make_int2(tag ? (v | (int)((d & 31) << 16)) : v, (int)x);
                w += __popc(m);
            }
        }
    }
}

// Panel kernels (both layouts, both passes): a block stages PS consecutive samples x V terms in
// shared memory as int32, sample-major with an odd row stride -- every thread issues its share of
// the panel's loads back to back (PS*V/256 independent, fully coalesced requests: the row-at-a-time
// kernels above are bound by one memory latency per row, not by bytes) -- and the warps then
// take one staged row each: ballot / popc for the row size and the record positions, redux.sync on
// 16-bit limbs for the row total.  FILL = false: rowptr[d+1] = row size, N, Σ counts, overflow flag;
// FILL = true: the (term, count) records at rowptr[d] (a row's records go out as <= 256-byte runs).
template <typename T, bool FILL>
__global__ void __launch_bounds__(256) k_dense_panel(const T *__restrict__ dense, long long D, int V, int layout, int PS,
                                                     long long *rowptr, double *N, int n_stride, int n_off, int *flags,
                                                     unsigned long long *total, int2 *rec, int tag) {
    extern __shared__ int sm[];                    // PS * VP staged counts, then PS + 1 row pointers (FILL)
    const int VP = V | 1, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long *rp = reinterpret_cast<long long *>(sm + ((PS * VP + 1) & ~1));
    const unsigned magic = (unsigned)((0x100000000ULL + V - 1) / V);       // i / V, exact for i < 2^16 (PS * V <= 49152)
    const unsigned lt = (1u << lane) - 1u;
    const int psh = 31 - __clz(PS);                // PS is a power of two
    const long long npanels = (D + PS - 1) / PS;
    unsigned long long tot_all = 0;
    int bad = 0;
    for (long long panel = blockIdx.x; panel < npanels; panel += gridDim.x) {
        const long long d0 = panel * PS;
        const int ns = (int)min((long long)PS, D - d0);
        if (layout == 1) {
            const T *src = dense + (size_t)d0 * V;
            const int n = ns * V;
#pragma unroll 8
            for (int i = threadIdx.x; i < n; i += 256) {
                const long long x = (long long)src[i];
                const int sidx = (int)(((unsigned long long)i * magic) >> 32), v = i - sidx * V;
                if (x > 2147483647LL) bad = 8;
                sm[sidx * VP + v] = x > 0 ? (int)x : 0;
            }
        } else {
            const int n = V << psh;
#pragma unroll 8
            for (int i = threadIdx.x; i < n; i += 256) {
                const int v = i >> psh, sidx = i & (PS - 1);
                const long long x = sidx < ns ? (long long)dense[(size_t)v * D + d0 + sidx] : 0;
                if (x > 2147483647LL) bad = 8;
                sm[sidx * VP + v] = x > 0 ? (int)x : 0;
            }
        }
        if (FILL)
            for (int i = threadIdx.x; i <= ns; i += 256) rp[i] = rowptr[d0 + i];
        __syncthreads();
        for (int sidx = warp; sidx < ns; sidx += 8) {
            const int *row = sm + sidx * VP;
            if (FILL) {
                long long w = rp[sidx];
                for (int v0 = 0; v0 < V; v0 += 32) {
                    const int v = v0 + lane;
                    const int x = v < V ? row[v] : 0;
                    const unsigned m = __ballot_sync(0xffffffffu, x > 0);
                    if (x > 0) rec[w + __popc(m & lt)] = make_int2(tag ? (v | (int)(((d0 + sidx) & 31) << 16)) : v, x);
                    w += __popc(m);
                }
            } else {
                unsigned nnz = 0;
                unsigned long long tot = 0;
                for (int v0 = 0; v0 < V; v0 += 32) {
                    const int v = v0 + lane;
                    const int x = v < V ? row[v] : 0;
                    nnz += __popc(__ballot_sync(0xffffffffu, x > 0));
                    tot += (unsigned)x;
                }
                const unsigned long long a = __reduce_add_sync(0xffffffffu, (unsigned)(tot & 0xffffu));
                const unsigned long long b = __reduce_add_sync(0xffffffffu, (unsigned)((tot >> 16) & 0xffffu));
                const unsigned long long c = __reduce_add_sync(0xffffffffu, (unsigned)(tot >> 32));
                const unsigned long long t = a + (b << 16) + (c << 32);
                if (lane == 0) {
                    rowptr[d0 + sidx + 1] = nnz;
                    if (N) N[(d0 + sidx) * n_stride + n_off] = (double)t;
                    tot_all += t;
                }
            }
        }
        __syncthreads();
    }
    if (!FILL) {
        for (int off = 16; off >= 1; off >>= 1) {
            tot_all += __shfl_xor_sync(0xffffffffu, tot_all, off);
            bad |= __shfl_xor_sync(0xffffffffu, bad, off);
        }
        if (lane == 0) {
            if (tot_all) atomicAdd(total, tot_all);
            if (bad) atomicOr(flags, bad);
        }
    }
}

// records -> separate term / count arrays (what mmsig_*_set_data takes)
__global__ void __launch_bounds__(256) k_unpack_rec(const int2 *__restrict__ rec, long long n, int *term, int *count) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int2 r = rec[i];
        term[i] = r.x & 0xffff;
        count[i] = r.y;
    }
}

}  // namespace mmsig
