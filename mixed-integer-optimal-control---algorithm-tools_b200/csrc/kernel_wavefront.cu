// kernel_wavefront.cu -- persistent, pipelined min-plus DP over all stages (the hot path).
//
// One launch walks every stage i = n-1 .. 1 of bellman_TRM! (HelpFunctions.jl:45-82) for one or more
// subproblems.  Work decomposition:
//
//   * The budget axis is cut into G slices of R consecutive SOURCE rows b' (one persistent CTA each,
//     one CTA per SM).  A target cell (b, l) of stage i has exactly one source row b' = b - b~_l(i)
//     (HelpFunctions.jl:69-71), so slicing by source row partitions the cells, every CTA reads only the
//     value rows it owns (resident in shared memory) and PUSHES results whose target row b' + b~_l
//     belongs to a higher slice into a global ring of value rows (row-major, the same layout as the rows in
//     shared memory), so the consumer takes its R rows with ONE bulk TMA straight into the rows the next stage
//     reads, and its own results simply overwrite the cells it produces itself.  Budget only flows upwards, therefore slice g depends on slices
//     g-1 .. g-D only (D = ceil(max b~ / R)): the slices form a pipeline and low-budget CTAs run ahead in
//     time.  Neighbours synchronise through per-CTA progress counters in global memory (release/acquire),
//     never through a grid-wide barrier.
//   * Warp specialisation inside a CTA.  COMPUTE warps do the arithmetic; one COMM warp runs one stage
//     ahead: it brings the stage's level costs s_l(i) and budget uses b~_l(i) (evaluated from df[:, i] and
//     u_old[:, i] by the prep kernel, S3) and the halo rows the predecessors pushed into shared memory by 1-D
//     bulk TMA (cp.async.bulk + mbarrier), polls the neighbours' progress counters, merges the halo cells
//     (and the +Inf of unreachable cells) into the next value rows, and publishes this CTA's own progress.  Compute and comm hand over through two shared-memory mbarriers
//     (`full`: rows ready, `done`: stage finished), so no global-memory latency is exposed in steady state.
//   * Phase B (compute): the stage is a small min-plus matrix product C[b', l] = min_j (s_l + c_jl) + P[b', j].
//     A thread owns a TB x TL register tile of cells, thread groups split the successor range j (JS groups).
//     Every candidate is two separately rounded FP64 adds (:67, :71) and a strict '>' (:73) that keeps the
//     earliest successor on ties and never lets +Inf/NaN win -- exactly the reference's arithmetic.  On sm_100a
//     an FP64 instruction occupies two issue slots, so a candidate costs DADD(2)+DSETP(2)+2 FSEL+SEL = 7 slots
//     (profiles/pipe_probe_r01.txt); the kernel is issue bound, not FP64-pipe or HBM bound.
//   * Phase C (compute): the JS partial (min, argmin) pairs of a cell are combined in ascending-j order with the
//     same strict '>' (the earliest group holding the minimum wins).  The value goes to the next stage's rows (own slice: shared memory; higher slice: global halo ring), the
//     argmin to HBM once per cell as uint8/uint16 indexed by source row (coalesced rows).
#include "bb200_internal.cuh"
#include "kernels.cuh"

namespace bb200 {

// ---- PTX helpers (mbarrier / bulk TMA helpers live in bb200_internal.cuh) -----------------------------
__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void compute_barrier(int nthreads)
{
    asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

// Spin until *flag >= want: relaxed polls, then ONE acquire load of the satisfied counter (an acquire load is far
// cheaper than a full fence.acq_rel.gpu, which costs 1-3 thousand cycles on a busy SM).  A bounded watchdog turns a
// lost dependency into an error code instead of a hung GPU: after ~2^23 polls the CTA raises the abort flag and
// every poller gives up.
__device__ __forceinline__ void wait_flag(const unsigned long long *flag, long long want, int *err)
{
    if (want <= 0) return;
    unsigned int spins = 0;
    while ((long long)ld_relaxed(flag) < want) {
        __nanosleep(100);  // the comm warp shares a scheduler with compute warps: do not burn their issue slots
        if ((++spins & 0x3ffu) == 0) {
            if (*(volatile int *)&err[3]) return;
            if (spins > (1u << 23)) {
                atomicOr(&err[2], 1);
                atomicOr(&err[3], 1);
                return;
            }
        }
    }
    (void)ld_acquire(flag);
}

struct Smem {
    uint64_t *mbar;   // [0..2] cost rows landed (buffer i%3), [3] halo landed, [4] full, [5] done, [6] relay counter
    double *ss;       // [3][Kp]   stage cost of stage i in ss[i%3]         (TMA destination)
    int *bts;         // [3][Kp]   budget use of stage i in bts[i%3]        (TMA destination)
    double *Ps;       // [2][R][Kp] value rows (row-major) read by stage i in Ps[i&1]  (halo rows: TMA destination)
    double *cs;       // [K*Kp]    jump costs
    double *pv;       // [JS*R*Kp] partial minima of the j-groups
    unsigned char *pa;  // ArgT[JS*R*Kp] partial argmins
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline size_t carve(const Tables &t, const WaveCfg &c, int argw, unsigned char *base, Smem *s)
{
    size_t off = 0;
    size_t o[7];
    const size_t sizes[7] = {8 * sizeof(uint64_t),
                             3 * (size_t)t.Kp * sizeof(double),
                             3 * (size_t)t.Kp * sizeof(int),
                             2 * (size_t)t.Kp * c.R * sizeof(double),
                             (size_t)t.K * t.Kp * sizeof(double),
                             (size_t)c.JS * c.R * t.Kp * sizeof(double),
                             (size_t)c.JS * c.R * t.Kp * (size_t)argw};
    for (int k = 0; k < 7; ++k) {
        o[k] = off;
        off = align_up(off + sizes[k], 128);
    }
    if (s) {
        s->mbar = reinterpret_cast<uint64_t *>(base + o[0]);
        s->ss = reinterpret_cast<double *>(base + o[1]);
        s->bts = reinterpret_cast<int *>(base + o[2]);
        s->Ps = reinterpret_cast<double *>(base + o[3]);
        s->cs = reinterpret_cast<double *>(base + o[4]);
        s->pv = reinterpret_cast<double *>(base + o[5]);
        s->pa = base + o[6];
    }
    return off;
}

// Phase B of one thread: TB x TL cells, successors [jb, je): the reference's innermost loop (HelpFunctions.jl:71-76).
//   Prow: value rows of this thread's row group, row-major [r][Kp];  crow: jump costs of its levels, [j][Kp]
//   srow: stage costs of its levels;  pv/pa: partial (min, argmin) out, [r][Kp]
// Two successors per trip: P[r][j], P[r][j+1] are neighbours in the row-major value rows, so one 16-byte
// warp-broadcast load serves both (jb is even by construction).
// Cost per candidate and cell on sm_100a: DADD + DSETP (two issue slots each) + 2 FSEL + SEL = 7 slots.
#ifndef BB_BVAR
#define BB_BVAR 0
#endif
template <int TB, int TL, typename ArgT>
__device__ __forceinline__ void phase_b(const double *__restrict__ Prow, const double *__restrict__ crow,
                                        const double *__restrict__ srow, double *__restrict__ pv,
                                        ArgT *__restrict__ pa, int jb, int je, int Kp)
{
    constexpr int MARKI = (int)(ArgT) ~(ArgT)0;
    const double inf = d_inf();
    double best[TB][TL];
    int arg[TB][TL];
#pragma unroll
    for (int a = 0; a < TB; ++a)
#pragma unroll
        for (int q = 0; q < TL; ++q) { best[a][q] = inf; arg[a][q] = MARKI; }
    double s[TL];
#pragma unroll
    for (int q = 0; q < TL; ++q) s[q] = srow[q];
    // main loop: full pairs, one straight-line block so that the two candidates' chains interleave
    const int je2 = jb + ((je - jb) & ~1);
#if BB_BVAR == 1
#pragma unroll 2
#elif BB_BVAR == 3
#pragma unroll 8
#elif BB_BVAR == 4
#pragma unroll 3
#elif BB_BVAR == 5
#pragma unroll 1
#else
#pragma unroll 4
#endif
    for (int j = jb; j < je2; j += 2) {
        double p0[TB], p1[TB];
#pragma unroll
        for (int r = 0; r < TB; ++r) {
            const double2 x = *reinterpret_cast<const double2 *>(Prow + (size_t)r * Kp + j);
            p0[r] = x.x;
            p1[r] = x.y;
        }
        double a0[TL], a1[TL];
        if constexpr (TL % 2 == 0) {
#pragma unroll
            for (int k = 0; k < TL / 2; ++k) {
                const double2 x = *reinterpret_cast<const double2 *>(crow + (size_t)j * Kp + 2 * k);
                const double2 y = *reinterpret_cast<const double2 *>(crow + (size_t)(j + 1) * Kp + 2 * k);
                a0[2 * k] = __dadd_rn(s[2 * k], x.x);  // HelpFunctions.jl:67
                a0[2 * k + 1] = __dadd_rn(s[2 * k + 1], x.y);
                a1[2 * k] = __dadd_rn(s[2 * k], y.x);
                a1[2 * k + 1] = __dadd_rn(s[2 * k + 1], y.y);
            }
        } else {
#pragma unroll
            for (int q = 0; q < TL; ++q) {
                a0[q] = __dadd_rn(s[q], crow[(size_t)j * Kp + q]);
                a1[q] = __dadd_rn(s[q], crow[(size_t)(j + 1) * Kp + q]);
            }
        }
#pragma unroll
        for (int q = 0; q < TL; ++q) {
#pragma unroll
            for (int r = 0; r < TB; ++r) {
                const double v = __dadd_rn(a0[q], p0[r]);               // :71
                if (best[r][q] > v) { best[r][q] = v; arg[r][q] = j; }  // :73-76, strict: the earliest j wins
            }
        }
#pragma unroll
        for (int q = 0; q < TL; ++q) {
#pragma unroll
            for (int r = 0; r < TB; ++r) {
                const double v = __dadd_rn(a1[q], p1[r]);
                if (best[r][q] > v) { best[r][q] = v; arg[r][q] = j + 1; }
            }
        }
    }
    if (je2 < je) {  // odd tail: one last successor
        const int j = je2;
#pragma unroll
        for (int q = 0; q < TL; ++q) {
            const double a = __dadd_rn(s[q], crow[(size_t)j * Kp + q]);
#pragma unroll
            for (int r = 0; r < TB; ++r) {
                const double v = __dadd_rn(a, Prow[(size_t)r * Kp + j]);
                if (best[r][q] > v) { best[r][q] = v; arg[r][q] = j; }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < TB; ++r)
#pragma unroll
        for (int q = 0; q < TL; ++q) {
            pv[(size_t)r * Kp + q] = best[r][q];
            pa[(size_t)r * Kp + q] = (ArgT)arg[r][q];
        }
}

// ======================================= COMM warp ===============================================
template <int TB>
__device__ __forceinline__ void comm_warp(const Tables &t, const WaveCfg &c, const Smem &sm, int lane)
{
    const int g = blockIdx.x;
    const int r0 = g * c.R;
    const int K = t.K, Kp = t.Kp, B1 = t.B1, R = c.R, n = t.n;
    const double inf = d_inf();
    uint64_t *mb_cost = &sm.mbar[0], *mb_halo = &sm.mbar[3], *mb_full = &sm.mbar[4], *mb_done = &sm.mbar[5];
    const int btm = min(*c.btmax, B1 - 1);
    const int D = (btm + R - 1) / R;  // slices a push can span
    const int my_rows = min(R, B1 - r0);  // rows of this slice that exist in the table (>= 1)
    const bool halo_on = (g > 0 && D > 0);  // lower slices push into this one
    const int lblocks = Kp >> 5;
    uint32_t cost_phase = 0;  // parity bit per cost buffer
    uint32_t done_phase = 0;
    bool prepolled = false;  // the predecessors' counters for the stage about to be prepared were already seen
    long long tick0 = 0;  // tick of the terminal stage of the current subproblem; stage i has tick0 + n - i
    long long pc[7] = {0, 0, 0, 0, 0, 0, 0};  // profile: cost wait, flag wait, merge, done wait, publish, stages, halo TMA
    long long tp = clock64();
#define PROF_LAP(k) do { if (c.prof) { const long long tq = clock64(); pc[k] += tq - tp; tp = tq; } } while (0)
    // level costs / budget uses of stage i (row i-1 of the prep kernel's tables) into buffer i%3.  Three buffers:
    // the stage being computed, the stage being prepared, and the one in flight for the stage after that.
    auto load_costs = [&](const SlotDev &sl, int i) {
        const int b = i % 3;
        mbar_expect_tx(&mb_cost[b], (uint32_t)(Kp * (sizeof(double) + sizeof(int))));
        tma_load_1d(sm.ss + (size_t)b * Kp, sl.ss_all + (size_t)(i - 1) * Kp, (uint32_t)(Kp * sizeof(double)), &mb_cost[b]);
        tma_load_1d(sm.bts + (size_t)b * Kp, sl.bt_all + (size_t)(i - 1) * Kp, (uint32_t)(Kp * sizeof(int)), &mb_cost[b]);
    };
    auto wait_costs = [&](int i) {
        mbar_wait(&mb_cost[i % 3], (cost_phase >> (i % 3)) & 1u);
        cost_phase ^= 1u << (i % 3);
    };

    for (int sub = 0; sub < c.nsub; ++sub, tick0 += n) {
        const SlotDev sl = c.slots[sub];
        if (lane == 0) {
            load_costs(sl, n);                  // terminal stage
            if (n >= 2) load_costs(sl, n - 1);  // first computed stage
        }
        wait_costs(n);
        if (n == 1) {
            // only the terminal stage exists: it is the exit state (slot 1 of the reference)
            for (int x = lane; x < my_rows * Kp; x += 32) {
                const int row = x / Kp, l = x % Kp;
                if (l < K) sl.phi[(size_t)(r0 + row) * Kp + l] = (r0 + row == sm.bts[Kp + l]) ? sm.ss[Kp + l] : inf;  // buffer 1 % 3
            }
            __syncwarp();
            continue;
        }

        for (int s = n - 1; s >= 0; --s) {
            const long long tick_s = tick0 + (n - s);  // tick of stage s
            if (s >= 1) {
                double *Pw = sm.Ps + (size_t)(s & 1) * R * Kp;
                if (s == n - 1) {
                    // ---- terminal stage n (HelpFunctions.jl:27-43) = the rows stage n-1 reads ----------------
                    // P[b][l] = (b == b~_l(n)) ? s_l(n) : Inf
                    const double *sn = sm.ss + (size_t)(n % 3) * Kp;
                    const int *bn = sm.bts + (size_t)(n % 3) * Kp;
                    for (int row = 0; row < my_rows; ++row)
                        for (int blk = 0; blk < lblocks; ++blk) {
                            const int l = (blk << 5) + lane, b = r0 + row;
                            if (l < K) {
                                const double v = (b == bn[l]) ? sn[l] : inf;
                                Pw[row * Kp + l] = v;
                                if (n == 2) sl.phi[((size_t)B1 + b) * Kp + l] = v;  // stage 2 is exit slot 2
                            }
                        }
                    __syncwarp();
                }
                // ---- my value rows as the lower slices pushed them during stage s+1: one bulk TMA straight into the
                // rows stage s reads (same row-major layout), issued first thing: the predecessors' counters were
                // already checked at the end of the previous iteration, and these rows were released by done(s+2).
                // The compute warps wait on mb_halo before they scatter their own stage-(s+1) results over the block.
                if (s <= n - 2 && halo_on) {
                    if (!prepolled) {  // not seen early (deep halo or short lag): wait for the predecessors here
                        for (int idx = lane; idx < D; idx += 32)
                            if (idx + 1 <= g) wait_flag(c.flags + (size_t)(g - idx - 1) * kFlagStride, tick_s - 1, c.err);
                        __syncwarp();
                    }
                    if (lane == 0) {
                        asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy acquire -> async-proxy read
                        const uint32_t bytes = (uint32_t)((size_t)my_rows * Kp * sizeof(double));
                        mbar_expect_tx(mb_halo, bytes);
                        tma_load_1d(Pw, c.halo + ((size_t)((tick_s - 1) % kHaloRing) * B1 + r0) * Kp, bytes, mb_halo);
                    }
                }
                PROF_LAP(6);
                // the buffer of stage s+2 is free (its `done` was observed in the previous iteration): prefetch s-1
                if (lane == 0 && s - 1 >= 1) load_costs(sl, s - 1);
                wait_costs(s);
                PROF_LAP(0);
                // ---- back-pressure: successors consumed the ring slot stage s will overwrite (one counter per lane) ----
                for (int idx = lane; idx < D; idx += 32)
                    if (g + idx + 1 < c.G)
                        wait_flag(c.flags + (size_t)(g + idx + 1) * kFlagStride, tick_s - kHaloRing + 1, c.err);
                // ---- look ahead: have the predecessors finished stage s already?  Then the next iteration's TMA can go
                // out at once.  Only the comm warp's idle time is spent on this: the probing stops as soon as the compute
                // warps finish stage s+1, so it never delays the hand-over (and cannot deadlock against back-pressure).
                prepolled = false;
                if (s - 1 >= 1 && halo_on) {
                    const bool have_done_next = (s + 1 <= n - 1);
                    for (;;) {
                        bool ok = true;
                        for (int idx = lane; idx < D; idx += 32)
                            if (idx + 1 <= g && (long long)ld_relaxed(c.flags + (size_t)(g - idx - 1) * kFlagStride) < tick_s) ok = false;
                        if (__all_sync(0xffffffffu, ok)) { prepolled = true; break; }
                        if (!have_done_next || mbar_test(mb_done, done_phase)) break;
                        __nanosleep(100);
                    }
                    if (prepolled)
                        for (int idx = lane; idx < D; idx += 32)
                            if (idx + 1 <= g) (void)ld_acquire(c.flags + (size_t)(g - idx - 1) * kFlagStride);
                }
                __syncwarp();  // every polling lane finished with an acquire load; lane 0 inherits the order through the warp sync
                PROF_LAP(1);
                PROF_LAP(2);
            }
            // ---- hand-over: wait for the compute warps to finish stage s+1, then release stage s -----------
            const bool have_done = (s + 1 <= n - 1);
            if (have_done) {
                mbar_wait(mb_done, done_phase);
                done_phase ^= 1u;
            }
            if (s >= 1) mbar_arrive(mb_full);
            PROF_LAP(3);
            // stage s+1 is complete in this CTA: hand its tick to the publisher warp, which makes the pushes visible
            // GPU-wide, moves the progress counter and prefetches the next cost rows off this warp's critical path
            if (have_done && lane == 0) {
                if (c.pub) {
                    asm volatile("st.release.cta.shared::cta.u64 [%0], %1;" ::"r"(smem_u32(&sm.mbar[6])), "l"((unsigned long long)(tick_s - 1)) : "memory");
                } else {
                    // long stages hide the fence: publish from here and save the extra warp
                    fence_gpu();
                    st_relaxed(c.flags + (size_t)g * kFlagStride, (unsigned long long)(tick_s - 1));
                }
            }
            PROF_LAP(4);
            pc[5] += 1;
        }
    }
    if (c.prof && lane == 0)
        for (int k = 0; k < 7; ++k) c.prof[(size_t)g * 16 + 8 + k] = pc[k];
#undef PROF_LAP
}

// ===================================== PUBLISHER warp ============================================
// Waits for the comm warp's relay (stage finished in this CTA), then: fence.acq_rel.gpu so that the pushes the
// compute threads made during that stage are visible GPU-wide, then the relaxed store of the progress counter --
// about 1 200 cycles that no longer sit on the comm warp's per-stage critical path.
__device__ __forceinline__ void publisher_warp(const Tables &t, const WaveCfg &c, const Smem &sm, int lane)
{
    if (lane != 0) return;
    const int g = blockIdx.x;
    const int n = t.n;
    unsigned long long *myflag = c.flags + (size_t)g * kFlagStride;
    long long tick0 = 0;
    for (int sub = 0; sub < c.nsub; ++sub, tick0 += n) {
        for (int s = n - 1; s >= 1; --s) {
            const unsigned long long tick = (unsigned long long)(tick0 + (n - s));
            unsigned long long seen;
            unsigned int spins = 0;
            do {
                asm volatile("ld.acquire.cta.shared::cta.u64 %0, [%1];" : "=l"(seen) : "r"(smem_u32(&sm.mbar[6])) : "memory");
                if (seen < tick) {
                    __nanosleep(250);  // a quiet poll: this warp shares a scheduler with compute warps
                    if ((++spins & 0xffffu) == 0 && *(volatile int *)&c.err[3]) return;  // watchdog fired elsewhere
                }
            } while (seen < tick);
            fence_gpu();
            st_relaxed(myflag, tick);
        }
    }
}

// MAXT is a multiple of 128: the register file is split evenly over the four schedulers, so the per-thread
// budget is set by the scheduler that hosts the most warps.
template <int TB, int TL, typename ArgT, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) wavefront_kernel(Tables t, WaveCfg c)
{
    constexpr ArgT MARK = (ArgT)~(ArgT)0;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem sm;
    carve(t, c, (int)sizeof(ArgT), smem_raw, &sm);

    const int tid = threadIdx.x;
    const int NC = c.JS * c.tpg;  // compute threads; comm warp = threads [NC, NC+32), publisher warp = [NC+32, NC+64)
    const int g = blockIdx.x;
    const int r0 = g * c.R;
    const int K = t.K, Kp = t.Kp, B1 = t.B1, R = c.R, n = t.n;
    const double inf = d_inf();
    uint64_t *mb_full = &sm.mbar[4], *mb_done = &sm.mbar[5];

    // one-time: jump costs into shared memory, value rows and halo staging to +Inf, barriers
    for (int x = tid; x < K * Kp; x += blockDim.x) sm.cs[x] = t.cost[x];
    for (int x = tid; x < 2 * R * Kp; x += blockDim.x) sm.Ps[x] = inf;
    if (tid == 0) {
        mbar_init(&sm.mbar[0], 1);
        mbar_init(&sm.mbar[1], 1);
        mbar_init(&sm.mbar[2], 1);
        mbar_init(&sm.mbar[3], 1);
        mbar_init(mb_full, 32);
        mbar_init(mb_done, NC);
        sm.mbar[6] = 0;  // relay counter
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= NC + 32) {  // only launched when c.pub
        publisher_warp(t, c, sm, tid - NC - 32);
        return;
    }
    if (tid >= NC) {
        comm_warp<TB>(t, c, sm, tid - NC);
        return;
    }

    // ========================================= COMPUTE warps ============================================
    const int jg = tid / c.tpg, tig = tid % c.tpg;
    const bool active = tig < c.RG * c.nLG;
    const int rg = active ? tig / c.nLG : 0;
    const int lg = active ? tig % c.nLG : 0;
    const int jb = jg * c.jper;
    const int je = min(K, jb + c.jper);
    uint32_t full_phase = 0, halo_phase = 0;
    long long tick0 = 0;
    const int btm_c = min(*c.btmax, B1 - 1);
    const bool halo_on = (g > 0 && btm_c > 0);  // same condition as the comm warp's (D > 0)
    uint64_t *mb_halo = &sm.mbar[3];
    long long pc[5] = {0, 0, 0, 0, 0};  // profile: full wait, phase B, barrier, phase C, stages
    long long pcc[3] = {0, 0, 0};       // phase C split: halo wait, combine, scatter
    long long tpc = 0;
    long long tp = clock64();
#define PROF_LAP(k) do { if (c.prof) { const long long tq = clock64(); pc[k] += tq - tp; tp = tq; } } while (0)

    for (int sub = 0; sub < c.nsub; ++sub, tick0 += n) {
        const SlotDev sl = c.slots[sub];
        ArgT *argtab = reinterpret_cast<ArgT *>(sl.arg);
        for (int i = n - 1; i >= 1; --i) {
            const long long tick = tick0 + (n - i);
            const double *Pc = sm.Ps + (size_t)(i & 1) * R * Kp;
            double *Pn = sm.Ps + (size_t)((i - 1) & 1) * R * Kp;
            const double *ssc = sm.ss + (size_t)(i % 3) * Kp;
            const int *bt_cur = sm.bts + (size_t)(i % 3) * Kp;

            mbar_wait(mb_full, full_phase);  // rows, stage costs and back-pressure for stage i are ready
            full_phase ^= 1u;
            PROF_LAP(0);

            // ---- phase B: register-tiled min-plus scan over this group's successors (values only) --------
            if (active)
                phase_b<TB, TL, ArgT>(Pc + (size_t)rg * TB * Kp, sm.cs + lg * TL, ssc + lg * TL,
                                      sm.pv + ((size_t)jg * R + rg * TB) * Kp + lg * TL,
                                      reinterpret_cast<ArgT *>(sm.pa) + ((size_t)jg * R + rg * TB) * Kp + lg * TL, jb, je,
                                      Kp);
            PROF_LAP(1);
            compute_barrier(NC);
            PROF_LAP(2);

            // ---- phase C: combine the j-groups in ascending order, scatter the value, store the argmin --------
            // Work unit = 32 consecutive levels of one source row (a warp-wide, coalesced row segment); every warp
            // handles up to CU units at once so that the dependent compare chains of different cells overlap.
            {
                constexpr int CU = 4;
                // ring slot of this stage's pushes: value rows [B1][Kp], row-major like the rows in shared memory
                double *hring = c.halo + (size_t)(tick % kHaloRing) * B1 * Kp;
                // my next rows receive the lower slices' block by TMA (issued by the comm warp): it must have landed
                // before my own results overwrite the cells I produce myself
                if (c.prof) tpc = clock64();
                if (halo_on && i - 1 >= 1 && i - 1 <= n - 2) {
                    mbar_wait(mb_halo, halo_phase);
                    halo_phase ^= 1u;
                }
                if (c.prof) { const long long tq = clock64(); pcc[0] += tq - tpc; tpc = tq; }
                const ArgT *pa_all = reinterpret_cast<const ArgT *>(sm.pa);
                const int lblocks = Kp >> 5;
                const int units = R * lblocks;
                const int warp = tid >> 5, lane = tid & 31, nwarps = NC >> 5;
                for (int u0 = warp; u0 < units; u0 += nwarps * CU) {
                    double val[CU];
                    int row_[CU], l_[CU], tgt_[CU], arg_[CU];
                    bool ok[CU], live_[CU];
#pragma unroll
                    for (int u = 0; u < CU; ++u) {
                        const int unit = u0 + u * nwarps;
                        const bool live = unit < units;
                        const int row = live ? unit / lblocks : 0;
                        const int l = live ? ((unit - row * lblocks) << 5) + lane : lane;
                        const int bsrc = r0 + row;
                        const int tgt = bsrc + bt_cur[l];
                        // cells outside `for b = 0:B-b~` (:69) are not computed by the reference either
                        ok[u] = live && l < K && bsrc < B1 && tgt < B1;
                        live_[u] = live;
                        row_[u] = row; l_[u] = l; tgt_[u] = tgt;
                        val[u] = inf; arg_[u] = (int)MARK;
                    }
                    for (int q = 0; q < c.JS; ++q) {
#pragma unroll
                        for (int u = 0; u < CU; ++u) {
                            const size_t x = ((size_t)q * R + row_[u]) * Kp + l_[u];
                            const double v = sm.pv[x];
                            if (val[u] > v) { val[u] = v; arg_[u] = (int)pa_all[x]; }  // strict: earliest group wins ties
                        }
                    }
                    if (c.prof) { const long long tq = clock64(); pcc[1] += tq - tpc; tpc = tq; }
#pragma unroll
                    for (int u = 0; u < CU; ++u) {
                        // target cell (bsrc, l) of my rows has no source row when bsrc < b~_l: +Inf (:47).  This also
                        // covers levels that are unreachable at this stage (b~ clamped to B1).
                        if (live_[u] && l_[u] < K && r0 + row_[u] < B1 && r0 + row_[u] < tgt_[u] - (r0 + row_[u]))
                            Pn[row_[u] * Kp + l_[u]] = inf;
                        if (!ok[u]) continue;
                        const int bsrc = r0 + row_[u];
                        argtab[((size_t)(i - 1) * B1 + bsrc) * Kp + l_[u]] = (ArgT)arg_[u];
                        if (tgt_[u] < r0 + R) Pn[(tgt_[u] - r0) * Kp + l_[u]] = val[u];
                        else hring[(size_t)tgt_[u] * Kp + l_[u]] = val[u];
                        if (i <= 2) sl.phi[((size_t)((i + 1) & 1) * B1 + tgt_[u]) * Kp + l_[u]] = val[u];
                    }
                }
            }
            if (c.prof) { const long long tq = clock64(); pcc[2] += tq - tpc; tpc = tq; }
            mbar_arrive(mb_done);
            PROF_LAP(3);
            pc[4] += 1;
        }
    }
    if (c.prof && tid == 0)
        for (int k = 0; k < 5; ++k) c.prof[(size_t)g * 16 + k] = pc[k];
    if (c.prof && tid == 0) {  // phase C split of warp 0: halo wait, combine, scatter
        c.prof[(size_t)g * 16 + 5] = pcc[0];
        c.prof[(size_t)g * 16 + 6] = pcc[1];
        c.prof[(size_t)g * 16 + 7] = pcc[2];
    }
#undef PROF_LAP
}

// ---- host side -----------------------------------------------------------------------------------
struct Variant { int TB, TL, maxt; };
// Large register tiles run with 8 compute warps (224 registers per thread); small tiles with up to 16 compute
// warps (120 registers per thread), which hides the FP64 compare->select latency with more warps in flight.
static const Variant kVariants[] = {{7, 4, kWaveThreadsBig},   {8, 4, kWaveThreadsBig},   {4, 4, kWaveThreadsSmall},
                                    {7, 2, kWaveThreadsSmall}, {8, 2, kWaveThreadsSmall}, {8, 1, kWaveThreadsSmall},
                                    {4, 1, kWaveThreadsSmall}};
static const int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);

static void fill_geometry(const Tables &t, int argw, int G, int JS, int v, int pub, WaveCfg &c)
{
    c.pub = pub;
    c.variant = v;
    c.TB = kVariants[v].TB;
    c.TL = kVariants[v].TL;
    const int rows_per_cta = (t.B1 + G - 1) / G;
    c.RG = (rows_per_cta + c.TB - 1) / c.TB;
    c.R = c.RG * c.TB;
    c.G = (t.B1 + c.R - 1) / c.R;  // drop CTAs that would own no row
    c.nLG = (t.K + c.TL - 1) / c.TL;
    c.JS = JS;
    c.jper = ((t.K + JS - 1) / JS + 1) & ~1;  // even: successors are taken in aligned pairs
    c.tpg = ((c.RG * c.nLG + 31) / 32) * 32;
    c.RP = c.R;
    c.threads = c.JS * c.tpg + 32 + 32 * c.pub;  // + the comm warp (+ the publisher warp)
    c.smem = carve(t, c, argw, nullptr, nullptr);
}

bool wave_configure(const Tables &t, int argw, int num_sms, size_t smem_max, int want_ctas, int want_js,
                    int want_variant, WaveCfg &cfg)
{
    if (t.M > kMaxM || t.K > 4096) return false;
    double best_score = -1.;
    bool found = false;
    for (int v = 0; v < kNumVariants; ++v) {
        if (want_variant > 0 && v != want_variant - 1) continue;
        if (((t.K + kVariants[v].TL - 1) / kVariants[v].TL) * kVariants[v].TL > t.Kp) continue;
        for (int js = 1; js <= 16; ++js) {
            if (want_js > 0 && js != want_js) continue;
            if (js > t.K) break;
            int gmax = want_ctas > 0 ? want_ctas : num_sms;
            if (gmax > num_sms) gmax = num_sms;
            WaveCfg c = cfg;
            fill_geometry(t, argw, gmax, js, v, 0, c);
            if (c.threads > kVariants[v].maxt || c.threads < 64) continue;
            if (c.smem > smem_max) continue;
            // Cost model per stage and CTA in scheduler cycles, fitted to the in-kernel profile on B200
            // (profiles/phase_profile_r01.txt).  Phase B is issue bound: an FP64 instruction takes two issue slots,
            // so a candidate costs DADD(2)+DSETP(2)+2 FSEL+SEL = 7 slots, plus the per-successor loads and the
            // s_l + c_jl adds; the busiest scheduler hosts ceil(warps/4) warps; ~25% of the slots are lost to
            // dependency stalls.  Phase C (combine + scatter) is latency bound and grows with the j-split.
            const double warps = (double)(c.threads - 32) / 32.0;
            const double tile = (double)(c.TB * c.TL);
            const double per_j = 7.0 * tile + 2.0 * c.TL + ((c.TB + 1) / 2) + ((c.TL + 1) / 2) + 3.0;
            const double b_cycles = per_j * c.jper * (double)(((int)warps + 3) / 4) * 1.25;
            const double units = (double)c.R * (t.Kp / 32);
            const double batches = (double)(((int)units + (int)warps * 4 - 1) / ((int)warps * 4));
            const double c_cycles = batches * (1500.0 + 300.0 * js);
            const double stage = b_cycles + c_cycles + 600.0;
            const double score = 1.0 / stage;  // every CTA does the same work per stage: smaller is better
            if (score > best_score) {
                best_score = score;
                cfg = c;
                found = true;
                // short stages cannot hide the ~1 200-cycle publish fence behind compute: give it its own warp
                if (stage < 9000.0 && c.threads + 32 <= kVariants[v].maxt) {
                    cfg.pub = 1;
                    cfg.threads += 32;
                }
            }
        }
    }
    return found;
}

template <int TB, int TL, int MAXT>
static cudaError_t launch_variant(const Tables &t, const WaveCfg &cfg, int argw, cudaStream_t st)
{
    void *args[] = {(void *)&t, (void *)&cfg};
    const void *fn = (argw == 1) ? (const void *)wavefront_kernel<TB, TL, uint8_t, MAXT>
                                 : (const void *)wavefront_kernel<TB, TL, uint16_t, MAXT>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem);
    if (e != cudaSuccess) return e;
    // cooperative launch: the CTAs wait on one another, so all of them must be co-resident
    return cudaLaunchCooperativeKernel(fn, dim3(cfg.G), dim3(cfg.threads), args, cfg.smem, st);
}

// Small tiles are compiled for two CTA sizes: up to 384 threads (168 registers per thread) and up to 512 (128).
template <int TB, int TL>
static cudaError_t launch_small(const Tables &t, const WaveCfg &cfg, int argw, cudaStream_t st)
{
    if (cfg.threads <= kWaveThreadsMid) return launch_variant<TB, TL, kWaveThreadsMid>(t, cfg, argw, st);
    return launch_variant<TB, TL, kWaveThreadsSmall>(t, cfg, argw, st);
}

cudaError_t launch_wavefront(const Tables &t, const WaveCfg &cfg, int argw, cudaStream_t st)
{
    switch (cfg.variant) {
        case 0: return launch_variant<7, 4, kWaveThreadsBig>(t, cfg, argw, st);
        case 1: return launch_variant<8, 4, kWaveThreadsBig>(t, cfg, argw, st);
        case 2: return launch_small<4, 4>(t, cfg, argw, st);
        case 3: return launch_small<7, 2>(t, cfg, argw, st);
        case 4: return launch_small<8, 2>(t, cfg, argw, st);
        case 5: return launch_small<8, 1>(t, cfg, argw, st);
        case 6: return launch_small<4, 1>(t, cfg, argw, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace bb200
