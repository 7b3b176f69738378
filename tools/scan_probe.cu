// scan_probe.cu -- the phase-B scan of kernel_wavefront.cu in isolation, with the relaxation's update
// (`if (best > v) { best = v; arg = j; }`, HelpFunctions.jl:73-76) written as different instruction mixes.
// Question: the three selects (FSEL, FSEL, SEL) run on the half-rate ALU pipe and bind the
// scan; can the conditional moves be issued to the idle FMA pipe (predicated IMAD / IMAD.WIDE) instead?
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -std=c++17 -lineinfo -o tools/scan_probe.bin tools/scan_probe.cu
// Every variant computes the same (min, argmin) table; a checksum over the partial tables proves bit-equality.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ double d_inf() { return __longlong_as_double(0x7ff0000000000000LL); }

// U = 0  CUDA C: the front end makes selects of it (ptxas: DSETP + FSEL + FSEL + SEL, three ALU-pipe instructions)
// U = 1  value: selp.f64 (FSEL + FSEL, ALU pipe);  arg: @P IMAD arg = z * hi(v) + j  (FMA pipe; z == 0 at run time,
//        the multiplier differs per candidate so that ptxas cannot hoist or fold it)
// U = 2  an empty volatile asm in the taken branch stops the front end's select conversion; ptxas if-converts the
//        short branch into three predicated moves and spreads them over both pipes (@P MOV / @P IMAD.MOV.U32)
// (predicated mad.wide / 64-bit moves written in PTX come back as SEL/FSEL; value halves as two predicated IMADs
//  work but ptxas then allocates the halves in unpaired registers and copies them before every DSETP.)
template <int U>
__device__ __forceinline__ void relax(double &best, int &arg, double v, int j, unsigned z)
{
    if constexpr (U == 0) {
        if (best > v) { best = v; arg = j; }
    } else if constexpr (U == 1) {
        asm volatile(
            "{\n.reg .pred p;\n.reg .b32 vl, vh;\nsetp.gt.f64 p, %0, %2;\nselp.f64 %0, %2, %0, p;\nmov.b64 {vl, vh}, %2;\n"
            "@p mad.lo.u32 %1, %3, vh, %4;\n}\n"
            : "+d"(best), "+r"(arg) : "d"(v), "r"(z), "r"(j));
    } else {
        if (best > v) { asm volatile(""); best = v; arg = j; }
    }
}

template <int U, int TB, int TL, int UNR>
__device__ __forceinline__ void scan(const double *__restrict__ Prow, const double *__restrict__ crow,
                                     const double *__restrict__ srow, double *__restrict__ pv,
                                     unsigned char *__restrict__ pa, int jb, int je, int Kp, unsigned z)
{
    double best[TB][TL];
    int arg[TB][TL];
#pragma unroll
    for (int a = 0; a < TB; ++a)
#pragma unroll
        for (int q = 0; q < TL; ++q) { best[a][q] = d_inf(); arg[a][q] = 255; }
    double s[TL];
#pragma unroll
    for (int q = 0; q < TL; ++q) s[q] = srow[q];
#pragma unroll UNR
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
            for (int r = 0; r < TB; ++r) relax<U>(best[r][q], arg[r][q], __dadd_rn(a0[q], p0[r]), j, z);
#pragma unroll
        for (int q = 0; q < TL; ++q)
#pragma unroll
            for (int r = 0; r < TB; ++r) relax<U>(best[r][q], arg[r][q], __dadd_rn(a1[q], p1[r]), j + 1, z);
    }
#pragma unroll
    for (int r = 0; r < TB; ++r)
#pragma unroll
        for (int q = 0; q < TL; ++q) {
            pv[(size_t)r * Kp + q] = best[r][q];
            pa[(size_t)r * Kp + q] = (unsigned char)arg[r][q];
        }
}

// GEO 0: tiles 4x2 then 3x2 on the same 8 warps (production: two sub-slices one after the other)
// GEO 1: one 7x2 tile, 8 warps
// GEO 2: tiles 4x2 and 3x2 on DIFFERENT warps at the same time (16 warps: four per scheduler instead of two)
// JS = 4 successor groups of 32 successors; 64 lanes per group cover the 128 levels in pairs.
template <int U, int GEO, int UNR>
__global__ void __launch_bounds__(GEO == 2 ? 512 : 256, 1) probe(long long *out, unsigned long long *chk, int stages, unsigned z)
{
    extern __shared__ double sh[];
    constexpr int Kp = 128, R = 7;
    double *P = sh;                  // [R][Kp]
    double *cs = P + R * Kp;         // [Kp][Kp]
    double *ss = cs + Kp * Kp;       // [Kp]
    double *pv = ss + Kp;            // [4][R][Kp]
    unsigned char *pa = reinterpret_cast<unsigned char *>(pv + 4 * R * Kp);
    for (int x = threadIdx.x; x < R * Kp; x += blockDim.x) P[x] = (x * 37 % 101) * 0.25 - 3.0;
    for (int x = threadIdx.x; x < Kp * Kp; x += blockDim.x) cs[x] = (x * 13 % 89) * 0.5;
    for (int x = threadIdx.x; x < Kp; x += blockDim.x) ss[x] = x * 0.125 - 4.0;
    __syncthreads();
    const int tid = threadIdx.x;
    const int half = tid / 256;  // GEO 2: 0 = sub-slice A, 1 = sub-slice B
    const int jg = (tid % 256) / 64, lg = tid % 64;
    const int jb = jg * 32, je = jb + 32;
    unsigned long long sum = 0;
    const long long t0 = clock64();
    for (int sidx = 0; sidx < stages; ++sidx) {
        if (GEO == 0 || (GEO == 2 && half == 0))
            scan<U, 4, 2, UNR>(P, cs + lg * 2, ss + lg * 2, pv + (size_t)jg * R * Kp + lg * 2, pa + (size_t)jg * R * Kp + lg * 2, jb, je, Kp, z);
        if (GEO == 0 || (GEO == 2 && half == 1))
            scan<U, 3, 2, UNR>(P + 4 * Kp, cs + lg * 2, ss + lg * 2, pv + ((size_t)jg * R + 4) * Kp + lg * 2, pa + ((size_t)jg * R + 4) * Kp + lg * 2, jb, je, Kp, z);
        if (GEO == 1)
            scan<U, 7, 2, UNR>(P, cs + lg * 2, ss + lg * 2, pv + (size_t)jg * R * Kp + lg * 2, pa + (size_t)jg * R * Kp + lg * 2, jb, je, Kp, z);
        __syncthreads();
        if (sidx + 1 == stages || sidx == 0)
            for (int x = tid; x < 4 * R * Kp; x += blockDim.x)
                sum += (unsigned long long)__double_as_longlong(pv[x]) * 1315423911ULL + pa[x] * (unsigned long long)(x + 1);
        if (tid < 7 * 16) P[tid * 8] = pv[tid * 8] * 0.5;  // keep the stages dependent (and make ties / negatives appear)
        __syncthreads();
    }
    const long long t1 = clock64();
    atomicAdd(&chk[blockIdx.x], sum);
    if (tid == 0) out[blockIdx.x] = (t1 - t0) / stages;
}

template <int U, int GEO, int UNR>
void run(const char *name)
{
    long long *d;
    unsigned long long *c;
    cudaMalloc(&d, 148 * sizeof(long long));
    cudaMalloc(&c, 148 * sizeof(unsigned long long));
    const int smem = (7 * 128 + 128 * 128 + 128 + 4 * 7 * 128) * 8 + 4 * 7 * 128 + 1024;
    const int threads = GEO == 2 ? 512 : 256;
    cudaFuncSetAttribute(probe<U, GEO, UNR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<U, GEO, UNR><<<148, threads, smem>>>(d, c, 500, 0u);
    cudaDeviceSynchronize();
    cudaMemset(c, 0, 148 * sizeof(unsigned long long));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<U, GEO, UNR><<<148, threads, smem>>>(d, c, 2000, 0u);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148];
    unsigned long long hc[148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaMemcpy(hc, c, sizeof(hc), cudaMemcpyDeviceToHost);
    long long mn = h[0], mx = h[0];
    for (int i = 0; i < 148; ++i) { mn = h[i] < mn ? h[i] : mn; mx = h[i] > mx ? h[i] : mx; }
    // 7 rows x 128 levels x 128 successors per stage and SM; FP64 floor: 2 ops per candidate (+ the s + c adds) on 64 lanes
    const double cand = 7.0 * 128 * 128;
    printf("%-52s cyc/stage min %6lld max %6lld | %5.2f cyc per warp-candidate per scheduler | %6.3f T upd/s | chk %016llx [%s]\n", name, mn, mx,
           mn / (cand / 32 / 4), 148.0 * cand * 2000 / (ms * 1e-3) / 1e12, hc[0], cudaGetErrorString(cudaGetLastError()));
    cudaFree(d); cudaFree(c);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s  SMs=%d  (FP64 floor: 4.16 cyc per warp-candidate for DADD + DSETP, +14%% for the s + c adds of the 4x2 + 3x2 tiles)\n", p.name, p.multiProcessorCount);
    run<0, 0, 4>("U0 selects (FSEL FSEL SEL)      4x2+3x2  unroll 4");
    run<1, 0, 4>("U1 FSEL FSEL + @P IMAD          4x2+3x2  unroll 4");
    run<2, 0, 4>("U2 predicated moves (ALU+FMA)   4x2+3x2  unroll 4");
    run<0, 1, 4>("U0 selects                      7x2      unroll 4");
    run<1, 1, 4>("U1 FSEL FSEL + @P IMAD          7x2      unroll 4");
    run<2, 1, 4>("U2 predicated moves             7x2      unroll 4");
    run<0, 2, 4>("U0 selects                      4x2|3x2 16 warps unroll 4");
    run<1, 2, 4>("U1 FSEL FSEL + @P IMAD          4x2|3x2 16 warps unroll 4");
    run<2, 2, 4>("U2 predicated moves             4x2|3x2 16 warps unroll 4");
    run<2, 0, 2>("U2 predicated moves             4x2+3x2  unroll 2");
    run<2, 0, 8>("U2 predicated moves             4x2+3x2  unroll 8");
    run<1, 0, 2>("U1 FSEL FSEL + @P IMAD          4x2+3x2  unroll 2");
    run<2, 2, 2>("U2 predicated moves             4x2|3x2 16 warps unroll 2");
    return 0;
}
