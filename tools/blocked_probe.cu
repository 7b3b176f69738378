// blocked_probe.cu -- phase-B loop variants in isolation: per-candidate argmin tracking (production) against
// value-only relaxation with the argmin tracked per block of 8 candidates.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -std=c++17 -o tools/blocked_probe.bin tools/blocked_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double d_inf() { return __longlong_as_double(0x7ff0000000000000LL); }

template <int TB, int TL>
__device__ __forceinline__ void scan_percand(const double *__restrict__ Prow, const double *__restrict__ crow,
                                             const double *__restrict__ srow, double *__restrict__ pv,
                                             unsigned char *__restrict__ pa, int jb, int je, int Kp)
{
    const double inf = d_inf();
    double best[TB][TL];
    int arg[TB][TL];
#pragma unroll
    for (int a = 0; a < TB; ++a)
#pragma unroll
        for (int q = 0; q < TL; ++q) { best[a][q] = inf; arg[a][q] = 255; }
    double s[TL];
#pragma unroll
    for (int q = 0; q < TL; ++q) s[q] = srow[q];
#pragma unroll 4
    for (int j = jb; j < je; j += 2) {
        double p0[TB], p1[TB];
#pragma unroll
        for (int r = 0; r < TB; ++r) {
            const double2 x = *reinterpret_cast<const double2 *>(Prow + (size_t)r * Kp + j);
            p0[r] = x.x;
            p1[r] = x.y;
        }
        double a0[TL], a1[TL];
#pragma unroll
        for (int k = 0; k < TL / 2; ++k) {
            const double2 x = *reinterpret_cast<const double2 *>(crow + (size_t)j * Kp + 2 * k);
            const double2 y = *reinterpret_cast<const double2 *>(crow + (size_t)(j + 1) * Kp + 2 * k);
            a0[2 * k] = __dadd_rn(s[2 * k], x.x);
            a0[2 * k + 1] = __dadd_rn(s[2 * k + 1], x.y);
            a1[2 * k] = __dadd_rn(s[2 * k], y.x);
            a1[2 * k + 1] = __dadd_rn(s[2 * k + 1], y.y);
        }
#pragma unroll
        for (int q = 0; q < TL; ++q)
#pragma unroll
            for (int r = 0; r < TB; ++r) {
                const double v = __dadd_rn(a0[q], p0[r]);
                if (best[r][q] > v) { best[r][q] = v; arg[r][q] = j; }
            }
#pragma unroll
        for (int q = 0; q < TL; ++q)
#pragma unroll
            for (int r = 0; r < TB; ++r) {
                const double v = __dadd_rn(a1[q], p1[r]);
                if (best[r][q] > v) { best[r][q] = v; arg[r][q] = j + 1; }
            }
    }
#pragma unroll
    for (int r = 0; r < TB; ++r)
#pragma unroll
        for (int q = 0; q < TL; ++q) {
            pv[(size_t)r * Kp + q] = best[r][q];
            pa[(size_t)r * Kp + q] = (unsigned char)arg[r][q];
        }
}

// value-only inside a block of 8 candidates, (min, block index) against the running best once per block
template <int TB, int TL>
__device__ __forceinline__ void scan_blocked(const double *__restrict__ Prow, const double *__restrict__ crow,
                                             const double *__restrict__ srow, double *__restrict__ pv,
                                             unsigned char *__restrict__ pa, int jb, int je, int Kp)
{
    const double inf = d_inf();
    double best[TB][TL];
    int blk[TB][TL];
#pragma unroll
    for (int a = 0; a < TB; ++a)
#pragma unroll
        for (int q = 0; q < TL; ++q) { best[a][q] = inf; blk[a][q] = 255; }
    double s[TL];
#pragma unroll
    for (int q = 0; q < TL; ++q) s[q] = srow[q];
#pragma unroll 1
    for (int j8 = jb; j8 < je; j8 += 8) {
        double m[TB][TL];
#pragma unroll
        for (int a = 0; a < TB; ++a)
#pragma unroll
            for (int q = 0; q < TL; ++q) m[a][q] = inf;
#pragma unroll
        for (int jj = 0; jj < 8; jj += 2) {
            const int j = j8 + jj;
            double p0[TB], p1[TB];
#pragma unroll
            for (int r = 0; r < TB; ++r) {
                const double2 x = *reinterpret_cast<const double2 *>(Prow + (size_t)r * Kp + j);
                p0[r] = x.x;
                p1[r] = x.y;
            }
            double a0[TL], a1[TL];
#pragma unroll
            for (int k = 0; k < TL / 2; ++k) {
                const double2 x = *reinterpret_cast<const double2 *>(crow + (size_t)j * Kp + 2 * k);
                const double2 y = *reinterpret_cast<const double2 *>(crow + (size_t)(j + 1) * Kp + 2 * k);
                a0[2 * k] = __dadd_rn(s[2 * k], x.x);
                a0[2 * k + 1] = __dadd_rn(s[2 * k + 1], x.y);
                a1[2 * k] = __dadd_rn(s[2 * k], y.x);
                a1[2 * k + 1] = __dadd_rn(s[2 * k + 1], y.y);
            }
#pragma unroll
            for (int q = 0; q < TL; ++q)
#pragma unroll
                for (int r = 0; r < TB; ++r) {
                    const double v = __dadd_rn(a0[q], p0[r]);
                    if (m[r][q] > v) m[r][q] = v;
                }
#pragma unroll
            for (int q = 0; q < TL; ++q)
#pragma unroll
                for (int r = 0; r < TB; ++r) {
                    const double v = __dadd_rn(a1[q], p1[r]);
                    if (m[r][q] > v) m[r][q] = v;
                }
        }
        const int b = j8 >> 3;
#pragma unroll
        for (int q = 0; q < TL; ++q)
#pragma unroll
            for (int r = 0; r < TB; ++r)
                if (best[r][q] > m[r][q]) { best[r][q] = m[r][q]; blk[r][q] = b; }
    }
#pragma unroll
    for (int r = 0; r < TB; ++r)
#pragma unroll
        for (int q = 0; q < TL; ++q) {
            pv[(size_t)r * Kp + q] = best[r][q];
            pa[(size_t)r * Kp + q] = (unsigned char)blk[r][q];
        }
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(long long *out, int stages)
{
    extern __shared__ double sh[];
    constexpr int Kp = 128, R = 7;
    double *P = sh;                  // [R][Kp]
    double *cs = P + R * Kp;         // [Kp][Kp]
    double *ss = cs + Kp * Kp;       // [Kp]
    double *pv = ss + Kp;            // [4][R][Kp]
    unsigned char *pa = reinterpret_cast<unsigned char *>(pv + 4 * R * Kp);
    for (int x = threadIdx.x; x < R * Kp; x += blockDim.x) P[x] = (x * 37 % 101) * 0.25;
    for (int x = threadIdx.x; x < Kp * Kp; x += blockDim.x) cs[x] = (x * 13 % 89) * 0.5;
    for (int x = threadIdx.x; x < Kp; x += blockDim.x) ss[x] = x * 0.125;
    __syncthreads();
    const int tid = threadIdx.x;
    __shared__ volatile int done_flag;
    __shared__ unsigned long long bar;
    if (tid == 0) {
        done_flag = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar)), "r"(1));
    }
    __syncthreads();
    if (tid >= 256) {
        // MODE 4/6: helper warps that poll shared memory between short sleeps (like the comm / publisher warps);
        // MODE 5: helper warps parked on an mbarrier (like the scatter warps waiting for the scan)
        if (MODE == 4 || MODE == 6) {
            while (!done_flag) __nanosleep(MODE == 4 ? 64 : 1000);
        } else if (MODE == 5) {
            unsigned ok = 0;
            while (!ok && !done_flag)
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                             : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(&bar)), "r"(0) : "memory");
        }
        return;
    }
    const int jg = tid / 64, lg = tid % 64;
    const int jb = jg * 32, je = jb + 32;
    const long long t0 = clock64();
    for (int sidx = 0; sidx < stages; ++sidx) {
        if (MODE == 0 || MODE >= 4) {
            scan_percand<4, 2>(P, cs + lg * 2, ss + lg * 2, pv + (size_t)jg * R * Kp + lg * 2, pa + (size_t)jg * R * Kp + lg * 2, jb, je, Kp);
            scan_percand<3, 2>(P + 4 * Kp, cs + lg * 2, ss + lg * 2, pv + ((size_t)jg * R + 4) * Kp + lg * 2, pa + ((size_t)jg * R + 4) * Kp + lg * 2, jb, je, Kp);
        } else if (MODE == 1) {
            scan_blocked<4, 2>(P, cs + lg * 2, ss + lg * 2, pv + (size_t)jg * R * Kp + lg * 2, pa + (size_t)jg * R * Kp + lg * 2, jb, je, Kp);
            scan_blocked<3, 2>(P + 4 * Kp, cs + lg * 2, ss + lg * 2, pv + ((size_t)jg * R + 4) * Kp + lg * 2, pa + ((size_t)jg * R + 4) * Kp + lg * 2, jb, je, Kp);
        } else if (MODE == 2) {
            scan_percand<7, 2>(P, cs + lg * 2, ss + lg * 2, pv + (size_t)jg * R * Kp + lg * 2, pa + (size_t)jg * R * Kp + lg * 2, jb, je, Kp);
        } else {
            scan_blocked<7, 2>(P, cs + lg * 2, ss + lg * 2, pv + (size_t)jg * R * Kp + lg * 2, pa + (size_t)jg * R * Kp + lg * 2, jb, je, Kp);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid < 7 * 16) P[tid * 8] = pv[tid * 8] * 0.5;  // keep the stages dependent
        asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    const long long t1 = clock64();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (tid == 0) {
        out[blockIdx.x] = (t1 - t0) / stages;
        done_flag = 1;
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(&bar)) : "memory");
    }
}

template <int MODE>
void run(const char *name)
{
    long long *d;
    cudaMalloc(&d, 148 * sizeof(long long));
    const int smem = (7 * 128 + 128 * 128 + 128 + 4 * 7 * 128) * 8 + 4 * 7 * 128 + 1024;
    cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<MODE><<<148, 512, smem>>>(d, 2000);
    cudaDeviceSynchronize();
    probe<MODE><<<148, 512, smem>>>(d, 2000);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mn = h[0], mx = h[0];
    for (int i = 0; i < 148; ++i) { mn = h[i] < mn ? h[i] : mn; mx = h[i] > mx ? h[i] : mx; }
    printf("%-44s cycles per stage (7 rows x 128 levels x 128 successors, 8 warps): min %lld max %lld  [%s]\n", name, mn, mx,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}

int main()
{
    run<0>("per-candidate argmin, tiles 4x2 + 3x2");
    run<1>("blocked (8) argmin, tiles 4x2 + 3x2");
    run<2>("per-candidate argmin, tile 7x2");
    run<3>("blocked (8) argmin, tile 7x2");
    run<4>("per-candidate 4x2+3x2, + 8 warps polling smem / nanosleep(64)");
    run<6>("per-candidate 4x2+3x2, + 8 warps polling smem / nanosleep(1000)");
    run<5>("per-candidate 4x2+3x2, + 8 warps parked on an mbarrier");
    return 0;
}
