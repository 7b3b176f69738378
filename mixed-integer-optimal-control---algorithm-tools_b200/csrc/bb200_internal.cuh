// bb200_internal.cuh -- shared definitions of the B200 trust-region DP library (sm_100a only).
//
// Device data layout (all per plan unless "per slot"):
//   lvd   double[K][M]    level values nu_k[m] as Float64 (HelpFunctions.jl:54-56 converts Int64->Float64)
//   Kp = K rounded up to a multiple of 32 (one warp-wide row segment)
//   cost  double[K][Kp]   cost[j*Kp + l] jump cost successor j (stage i+1) <- level l (stage i); pad = +Inf
//   goff  int64[K]        column-major grid offset of admissible tuple k
//   per slot:
//   df, u_old, u  double[nPad][M]   reference layout (Julia M x n), rows padded to the TMA chunk
//   ss_all double[n][Kp], bt_all int[n][Kp]  per-stage level costs / budget uses (S3), written once per DP by prep
//   phi   double[2][B1][Kp]  exit state (S7): [0] = stage-1 values, [1] = stage-2 values, level fastest
//   arg   ArgT[n-1][B1][Kp]  packed argmin, indexed by SOURCE budget row b' = b - b~_l(i):
//                            arg[i-1][b'][l] = winner j of target cell (b' + b~_l(i), l) at stage i
//                            (MARK where no candidate won, i.e. the reference wrote nothing)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bb200 {

constexpr int kChunk = 64;        // stages per TMA chunk of df / u_old
constexpr int kWaveThreadsSmall = 512;  // wavefront CTA cap (up to 16 warps: compute + comm + publisher + scatter): 128 regs/thread
constexpr int kMaxM = 8;          // controls supported by the kernels
constexpr int kFlagStride = 16;   // u64 words between per-CTA progress flags (128 B apart)
constexpr int kHaloRing = 32;     // stages of halo kept in flight between neighbouring CTAs (deep: a push can span many slices)

struct SlotDev {
    const double *df;     // [nPad][M]
    const double *u_old;  // [nPad][M]
    double *u;            // [nPad][M]
    double *phi;          // [2][B1][Kp]
    void *arg;            // ArgT[n-1][B1][Kp]
    double *ss_all;       // [n][Kp] stage cost s_l(i) of stage i in row i-1 (filled by the prep kernel)
    int *bt_all;          // [n][Kp] budget use b~_l(i), clamped to B1 = unreachable; pad levels hold B1
    unsigned long long *n_updates;  // exact relaxation count of the last DP
    double *rec;          // [kMaxRadii][4] phi_star, b_star, k_star, stale flag of selection r; then [4] slot status:
                          // rec[4 * kMaxRadii] != 0: u_old of this slot is not integer valued (InexactError)
};

constexpr int kMaxRadii = 16;                    // selections per subproblem in one batched sweep
constexpr int kRecDoubles = 4 * (kMaxRadii + 1);  // per-slot record block

// Batched radius sweep (multi-trust.jl:109-110 for many subproblems at once): CTA x of the selection / backtrack
// launch serves slot x / n_radii with radius x % n_radii and writes its trajectory to u_sweep[x].
struct SweepDev {
    const SlotDev *slots;  // nullptr: single-slot launch
    const int *radii;      // [n_radii] trial budgets B'
    int n_radii;
    double *u_sweep;       // [slots * n_radii][n][M]
    size_t u_stride;       // n * M
};

struct Tables {
    int n, M, K, Kp, B1;  // B1 = B + 1
    double dt;
    const double *lvd;
    const double *cost;
    const long long *goff;
};

__device__ __forceinline__ double d_inf() { return __longlong_as_double(0x7ff0000000000000LL); }

// Stage cost and budget use of level l at 0-based stage row `row` (S3):
//   s = 0. + (dt*df[0])*nu[0] + (dt*df[1])*nu[1] + ..   every * and + rounded separately
//   bt = sum_m Int64(|nu[m] - u_old[m]|), clamped to B1 (anything > B is unreachable)
__device__ __forceinline__ void stage_cost(const Tables &t, const double *__restrict__ lv,
                                           const double *__restrict__ dfrow,
                                           const double *__restrict__ uorow, double &s, int &bt)
{
    double acc = 0.;
    double b = 0.;
#pragma unroll 1
    for (int m = 0; m < t.M; ++m) {
        const double nu = lv[m];
        acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(t.dt, dfrow[m]), nu));
        b += fabs(nu - uorow[m]);
    }
    s = acc;
    bt = (b < (double)t.B1) ? (int)b : t.B1;  // NaN compares false -> unreachable
}

// Julia findmin order (S8): isgreater(fm, fx) == "fx strictly precedes fm".
__device__ __forceinline__ bool julia_isless(double a, double b)
{
    if (a != a) return false;
    if (b != b) return true;
    if (a < b) return true;
    if (a == b) return (__double_as_longlong(a) < 0) && !(__double_as_longlong(b) < 0);
    return false;
}
__device__ __forceinline__ bool julia_isgreater(double fm, double fx)
{
    return (fm != fm || fx != fx) ? julia_isless(fm, fx) : julia_isless(fx, fm);
}

// ---- mbarrier / 1-D bulk TMA helpers (sm_90+/sm_100a PTX) -----------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait parks the warp in hardware until the phase completes or the time hint expires; without a hint it returns
// almost at once and a waiting warp burns the issue slots of the warps that share its scheduler (the waiting loops
// executed half as many instructions as the scan before the hint was added).
constexpr uint32_t kMbarSuspendNs = 1000000u;
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity), "r"(kMbarSuspendNs)
            : "memory");
    }
}
// Non-blocking probe: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
// 1-D bulk TMA global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

}  // namespace bb200
