// kernels_microbench.cu -- FP64-pipe issue-rate microbenchmarks (the roofline denominator).
//
// MEASURED_PEAKS.json has HBM and bf16 tensor peaks but no FP64 (non-tensor) figure, and the DP is bound by
// the FP64 pipe (one DADD and one DSETP per cell-update).  These kernels measure, on the box
// the benchmark runs on and under a sustained full-chip load:
//   mode 0: independent DADD chains only            -> FP64 instruction issue rate (lane-ops/s)
//   mode 1: DADD + DSETP.GT + predicated selects    -> the relaxation's own instruction mix, operands in registers
#include "bb200_internal.cuh"
#include "kernels.cuh"

namespace bb200 {

template <int MODE>
__global__ void __launch_bounds__(256) fp64_rate_kernel(double *out, int iters, double seed)
{
    constexpr int C = 16;  // independent chains per thread
    __shared__ double xs[256];
    xs[threadIdx.x] = seed * (double)(threadIdx.x % 17) - 3.0;
    __syncthreads();
    double acc[C], y[C];
    int arg[C];
#pragma unroll
    for (int k = 0; k < C; ++k) {
        acc[k] = (MODE == 0) ? seed * (k + 1 + threadIdx.x) : d_inf();
        y[k] = seed * (double)(k + 1);
        arg[k] = 0;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep) {
            const double x = xs[(it * 8 + rep) & 255];  // one broadcast LDS per 16 relaxations
#pragma unroll
            for (int k = 0; k < C; ++k) {
                if (MODE == 0) {
                    acc[k] = __dadd_rn(acc[k], x);
                } else {
                    const double v = __dadd_rn(y[k], x);                        // candidate: one DADD
                    if (acc[k] > v) { acc[k] = v; arg[k] = it * 8 + rep; }      // DSETP.GT + selects
                }
            }
        }
    }
    double s = 0.;
    int a = 0;
#pragma unroll
    for (int k = 0; k < C; ++k) { s += acc[k]; a += arg[k]; }
    if (s == 123.456 || a == -1) out[0] = s;  // keep the chains alive
}

// Returns lane-operations per second of FP64-pipe instructions (mode 0: DADD; mode 1: DADD+DSETP pairs
// counted as 2 ops) measured with CUDA events over roughly `target_ms`.
cudaError_t measure_fp64_rate(int mode, int num_sms, double target_ms, double *ops_per_s, double *elapsed_ms,
                              cudaStream_t st)
{
    double *d_out = nullptr;
    cudaError_t e = cudaMalloc(&d_out, sizeof(double));
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = num_sms * 4, threads = 256;
    int iters = 2000;
    float ms = 0.f;
    for (int round = 0; round < 6; ++round) {
        cudaEventRecord(e0, st);
        if (mode == 0) fp64_rate_kernel<0><<<blocks, threads, 0, st>>>(d_out, iters, 1.000001);
        else fp64_rate_kernel<1><<<blocks, threads, 0, st>>>(d_out, iters, 1.000001);
        cudaEventRecord(e1, st);
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) break;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms >= 0.6 * target_ms) break;
        const double scale = ms > 0.01 ? target_ms / ms : 50.;
        iters = (int)(iters * (scale > 50. ? 50. : scale)) + 1;
    }
    const double per_thread = (double)iters * 8 * 16 * (mode == 0 ? 1. : 2.);
    *ops_per_s = per_thread * (double)blocks * threads / (ms * 1e-3);
    *elapsed_ms = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    return e;
}

}  // namespace bb200
