// pipe_probe.cu -- issue-slot / pipe-throughput probes for the relaxation's instruction mix on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_probe.bin tools/pipe_probe.cu
// Each kernel runs `iters` iterations of an unrolled body over C independent chains per thread, 8 warps per
// SMSP resident (4 CTAs x 256 threads per SM... grid = SMs*4), and reports SM cycles per warp-level body item.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define C 16
__device__ __forceinline__ double inf_d() { return __longlong_as_double(0x7ff0000000000000LL); }

template <int MODE>
__global__ void __launch_bounds__(256) probe(double *out, long long *cyc, int iters, double seed)
{
    __shared__ double xs[256];
    xs[threadIdx.x] = seed * (double)(threadIdx.x % 17) - 3.0;
    __syncthreads();
    double acc[C], y[C];
    int arg[C];
    float facc[C];
#pragma unroll
    for (int k = 0; k < C; ++k) {
        acc[k] = (MODE == 0) ? seed * (k + 1 + threadIdx.x) : inf_d();
        y[k] = seed * (double)(k + 1);
        arg[k] = 0;
        facc[k] = 1e30f;
    }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep) {
            const double x = xs[(it * 8 + rep) & 255];
#pragma unroll
            for (int k = 0; k < C; ++k) {
                if (MODE == 0) {                         // DADD only
                    acc[k] = __dadd_rn(acc[k], x);
                } else if (MODE == 1) {                  // full relaxation: DADD, DSETP, 2 FSEL, SEL
                    const double v = __dadd_rn(y[k], x);
                    if (acc[k] > v) { acc[k] = v; arg[k] = it * 8 + rep; }
                } else if (MODE == 2) {                  // value only: DADD, DSETP, 2 FSEL
                    const double v = __dadd_rn(y[k], x);
                    if (acc[k] > v) { acc[k] = v; }
                } else if (MODE == 3) {                  // DADD + DSETP + predicated int add (1 ALU)
                    const double v = __dadd_rn(y[k], x);
                    if (acc[k] > v) arg[k] += 1;
                } else if (MODE == 4) {                  // fp32 relaxation: FADD, FSETP, FSEL, SEL
                    const float v = __fadd_rn((float)y[k], (float)x);
                    if (facc[k] > v) { facc[k] = v; arg[k] = it * 8 + rep; }
                } else if (MODE == 5) {                  // 3 selects only, predicate from a cheap integer test
                    const bool p = ((it + k) & 7) == 0;
                    const double v = y[k];
                    acc[k] = p ? v : acc[k];
                    arg[k] = p ? rep : arg[k];
                    y[k] = acc[k];
                } else if (MODE == 6) {                  // DSETP only (compare against a moving scalar), int add
                    if (acc[k] > x) arg[k] += 1;
                } else if (MODE == 7) {                  // 64-bit integer min relaxation: DADD + s64 min (+ arg via compare)
                    const double v = __dadd_rn(y[k], x);
                    long long a = __double_as_longlong(acc[k]), b = __double_as_longlong(v);
                    acc[k] = __longlong_as_double(a < b ? a : b);
                } else if (MODE == 8) {                  // DADD + hi-word integer compare + 3 selects
                    const double v = __dadd_rn(y[k], x);
                    if (__double2hiint(acc[k]) > __double2hiint(v)) { acc[k] = v; arg[k] = it * 8 + rep; }
                }
            }
        }
    }
    long long t1 = clock64();
    double s = 0.;
    int a = 0;
    float f = 0.f;
#pragma unroll
    for (int k = 0; k < C; ++k) { s += acc[k] + y[k]; a += arg[k]; f += facc[k]; }
    if (s == 123.456 || a == -1 || f == 77.f) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE>
void run(const char *name, int sms, int ctas_per_sm, double items_per_body)
{
    double *d_out; long long *d_cyc;
    cudaMalloc(&d_out, 8); cudaMalloc(&d_cyc, 8);
    const int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<sms * ctas_per_sm, 256>>>(d_out, d_cyc, 200, 1.000001);
    cudaEventRecord(e0);
    probe<MODE><<<sms * ctas_per_sm, 256>>>(d_out, d_cyc, iters, 1.000001);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long cyc; cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
    // per SMSP: warps = ctas_per_sm*8/4; warp-level items executed = warps * iters * 8 * C
    const double warps_per_smsp = ctas_per_sm * 8 / 4.0;
    const double items = warps_per_smsp * iters * 8.0 * C * items_per_body;
    printf("%-44s ctas/SM=%d  %8.3f ms  %10lld cyc  cycles per warp-item per SMSP = %6.3f   (%.2f T lane-items/s)\n", name,
           ctas_per_sm, ms, cyc, cyc / items, (double)sms * ctas_per_sm * 256 * iters * 8.0 * C * items_per_body / (ms * 1e-3) / 1e12);
    cudaFree(d_out); cudaFree(d_cyc);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("%s  SMs=%d\n", p.name, sms);
    for (int c : {1, 2, 4}) {
        run<0>("0 DADD", sms, c, 1);
        run<1>("1 relax: DADD DSETP FSEL FSEL SEL", sms, c, 1);
        run<2>("2 value only: DADD DSETP FSEL FSEL", sms, c, 1);
        run<3>("3 DADD DSETP @P IADD", sms, c, 1);
        run<4>("4 fp32 relax: FADD FSETP FSEL SEL", sms, c, 1);
        run<5>("5 selects only (3 per item)", sms, c, 1);
        run<6>("6 DSETP @P IADD", sms, c, 1);
        run<7>("7 DADD + s64 min", sms, c, 1);
        run<8>("8 DADD + ISETP(hi) + 3 sel", sms, c, 1);
    }
    return 0;
}
