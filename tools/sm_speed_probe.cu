// sm_speed_probe.cu -- does every SM run the relaxation's instruction mix at the same rate?
// One CTA per SM (forced by the shared-memory request), 8 warps, operands in registers; prints cycles per SM id.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o tools/sm_speed_probe.bin tools/sm_speed_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>

// LDS-fed variant: operands come from shared memory like in phase B (warp-broadcast 16-byte loads of the value
// rows, per-lane 16-byte loads of the cost rows), 8 relaxations per 3 loads.
__global__ void __launch_bounds__(256, 1) probe_lds(long long *out, int iters, double seed)
{
    extern __shared__ double sh[];
    for (int x = threadIdx.x; x < 16384; x += blockDim.x) sh[x] = seed + (x % 97);
    __syncthreads();
    double best[8];
    int arg[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { best[k] = 1e300; arg[k] = 0; }
    const int lane2 = (threadIdx.x & 63) * 2;
    const long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < iters; ++it) {
        const int j = (it * 2) & 127;
        const double2 p0 = *reinterpret_cast<const double2 *>(sh + j);          // broadcast
        const double2 p1 = *reinterpret_cast<const double2 *>(sh + 128 + j);    // broadcast
        const double2 c = *reinterpret_cast<const double2 *>(sh + 1024 + j * 64 + lane2);  // per lane
        const double a0 = __dadd_rn(c.x, seed), a1 = __dadd_rn(c.y, seed);
        const double v[8] = {__dadd_rn(a0, p0.x), __dadd_rn(a0, p1.x), __dadd_rn(a1, p0.x), __dadd_rn(a1, p1.x),
                             __dadd_rn(a0, p0.y), __dadd_rn(a0, p1.y), __dadd_rn(a1, p0.y), __dadd_rn(a1, p1.y)};
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (best[k] > v[k]) { best[k] = v[k]; arg[k] = it; }
    }
    const long long t1 = clock64();
    double acc = 0;
    int a = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { acc += best[k]; a += arg[k]; }
    if (acc == 12345.678 && a == 77) sh[0] = acc;
    if (threadIdx.x == 0) {
        unsigned int smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        out[2 * blockIdx.x] = smid;
        out[2 * blockIdx.x + 1] = t1 - t0;
    }
}

__global__ void __launch_bounds__(256, 1) probe(long long *out, int iters, double seed)
{
    extern __shared__ double sh[];
    double best[8], v[8];
    int arg[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { best[k] = 1e300; v[k] = seed + threadIdx.x + k; arg[k] = 0; }
    sh[threadIdx.x] = seed;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const double c = __dadd_rn(v[k], -1.0);
            v[k] = c;
            if (best[k] > c) { best[k] = c; arg[k] = it; }
        }
    }
    const long long t1 = clock64();
    double acc = 0;
    int a = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { acc += best[k]; a += arg[k]; }
    if (acc == 12345.678 && a == 77) sh[0] = acc;
    if (threadIdx.x == 0) {
        unsigned int smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        out[2 * blockIdx.x] = smid;
        out[2 * blockIdx.x + 1] = t1 - t0;
    }
}

int main(int argc, char **argv)
{
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long *d;
    cudaMalloc(&d, sizeof(long long) * 2 * sms);
    const int smem = 150 * 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const bool lds = argc > 1;
    cudaFuncSetAttribute(probe_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rep = 0; rep < 3; ++rep) {
        if (lds) probe_lds<<<sms, 256, smem>>>(d, 200000, 1.0);
        else probe<<<sms, 256, smem>>>(d, 200000, 1.0);
        cudaDeviceSynchronize();
    }
    std::vector<long long> h(2 * sms);
    cudaMemcpy(h.data(), d, sizeof(long long) * 2 * sms, cudaMemcpyDeviceToHost);
    std::vector<std::pair<long long, long long>> v;
    for (int i = 0; i < sms; ++i) v.push_back({h[2 * i], h[2 * i + 1]});
    std::sort(v.begin(), v.end());
    long long mn = v[0].second, mx = v[0].second;
    for (auto &p : v) { mn = std::min(mn, p.second); mx = std::max(mx, p.second); }
    printf("per-SM cycles for 200000 x 8 relaxations x 8 warps: min %lld max %lld (%.2f%% spread), err=%s\n", mn, mx,
           100.0 * (mx - mn) / mn, cudaGetErrorString(cudaGetLastError()));
    for (int i = 0; i < sms; ++i)
        if ((double)v[i].second / mn > 1.002) printf("slow SM %lld: %.4f\n", v[i].first, (double)v[i].second / mn);
    return 0;
}
