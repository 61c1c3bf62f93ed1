// fp64_peak.cu -- measures the FP64 pipe's sustained rate on this GPU (warp-level DFMA / DADD / DMUL
// per clock per SM), the second roof of the exact-MMA E-step (k_solve is bound by
// FP64 issue, not by HBM).  Not part of the product; round-2 measurement aid.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int OP>
__global__ void __launch_bounds__(256) k_fp64(double *out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = a + threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (OP == 0) x[i] = fma(x[i], a, b);
                else if (OP == 1) x[i] = x[i] + b;
                else x[i] = x[i] * a;
            }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;      // keeps the chains alive
}

template <int ILP, int OP>
static void run(const char *name, int sms, int blocks_per_sm, double clock_ghz) {
    double *out;
    cudaMalloc(&out, 8);
    const int iters = 4096, grid = sms * blocks_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_fp64<ILP, OP><<<grid, 256>>>(out, 64, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_fp64<ILP, OP><<<grid, 256>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_inst = (double)grid * 8 * iters * 16 * ILP;
    const double per_clk_sm = warp_inst / (ms * 1e-3 * clock_ghz * 1e9 * sms);
    printf("{\"op\": \"%s\", \"ilp\": %d, \"warps_per_sm\": %d, \"ms\": %.3f, \"warp_inst_per_clk_per_sm\": %.3f, \"lane_ops_per_s\": %.4g}\n",
           name, ILP, blocks_per_sm * 8, ms, per_clk_sm, warp_inst * 32 / (ms * 1e-3));
    cudaFree(out);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_ghz_nominal\": %.3f}\n", p.name, p.multiProcessorCount, ghz);
    for (int bps : {1, 3, 4}) {
        run<1, 0>("dfma", p.multiProcessorCount, bps, ghz);
        run<4, 0>("dfma", p.multiProcessorCount, bps, ghz);
        run<4, 1>("dadd", p.multiProcessorCount, bps, ghz);
        run<4, 2>("dmul", p.multiProcessorCount, bps, ghz);
    }
    return 0;
}
