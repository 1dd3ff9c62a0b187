// Microbenchmark: non-FMA FP64 issue rate and dependent latency on this GPU (the Gaussian IIR's roofline).
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -fmad=false tools/fp64_peak.cu -o tools/_bin/fp64_peak
#include <cuda_runtime.h>
#include <cstdio>
template <int CHAINS> __global__ void dp_kernel(double *out, double a, double b, int iters) {
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 1e-3 + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) { x[c] = __dmul_rn(x[c], a); x[c] = __dadd_rn(x[c], b); }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CHAINS> void run(int blocks_per_sm, int threads, int sms, double clock_ghz) {
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks_per_sm * sms * threads);
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    dp_kernel<CHAINS><<<blocks_per_sm * sms, threads>>>(out, 0.999999, 1e-7, 100);
    cudaEventRecord(e0);
    dp_kernel<CHAINS><<<blocks_per_sm * sms, threads>>>(out, 0.999999, 1e-7, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instrs = 2.0 * CHAINS * iters * (threads / 32.0) * blocks_per_sm;  // per SM
    const double cycles = ms * 1e-3 * clock_ghz * 1e9;
    printf("chains %d  warps/SM %3d: %.3f ms  %.3f DP warp-instr/clk/SM  (%.1f lanes/clk/SM); per-warp dependent issue interval %.1f clk\n", CHAINS,
           blocks_per_sm * threads / 32, ms, warp_instrs / cycles, 32 * warp_instrs / cycles, cycles / (2.0 * iters));
    cudaFree(out);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double ghz = clk * 1e-6;
    printf("%s, %d SMs, %.3f GHz nominal\n", p.name, p.multiProcessorCount, ghz);
    const int sms = p.multiProcessorCount;
    run<1>(1, 32, sms, ghz);    // one warp per SM, one chain: dependent latency
    run<1>(1, 128, sms, ghz);   // 1 warp per scheduler
    run<1>(2, 128, sms, ghz);
    run<1>(4, 128, sms, ghz);   // 4 warps per scheduler
    run<1>(8, 128, sms, ghz);
    run<2>(4, 128, sms, ghz);
    run<4>(4, 128, sms, ghz);
    run<8>(4, 128, sms, ghz);
    run<8>(8, 128, sms, ghz);
    run<2>(1, 128, sms, ghz);
    run<4>(1, 128, sms, ghz);
    return 0;
}
