// kernel_wavefront.cu -- persistent, pipelined min-plus DP over all stages (the hot path).
//
// One launch walks every stage i = n-1 .. 1 of bellman_TRM! (HelpFunctions.jl:45-82) for one or more
// subproblems.  Work decomposition:
//
//   * The budget axis is cut into G slices of R consecutive SOURCE rows b' (one persistent CTA each, one CTA per
//     SM).  A target cell (b, l) of stage i has exactly one source row b' = b - b~_l(i) (HelpFunctions.jl:69-71),
//     so slicing by source row partitions the cells: every CTA reads only the value rows it owns (resident in
//     shared memory, row-major) and PUSHES results whose target row b' + b~_l belongs to a higher slice into a
//     global ring of value rows with the same layout.  The consumer takes its R rows with ONE bulk TMA straight
//     into the rows the next stage reads; its own results then overwrite the cells it produces itself.  Budget
//     only flows upwards, so slice g depends on slices g-1 .. g-D only (D = ceil(max b~ / R)): the slices form a
//     pipeline and low-budget CTAs run ahead in time.  Neighbours synchronise through per-CTA progress counters
//     in global memory (release/acquire), never through a grid-wide barrier.
//   * The same "budget only flows upwards" argument is used once more INSIDE a CTA: its rows are split into a
//     lower sub-slice A and an upper sub-slice B.  The compute warps alternate scan(A, i), scan(B, i),
//     scan(A, i-1), ...; SCATTER warps trail one sub-step behind and finish sub-slice A of stage i (combine the
//     partial minima, store the argmin, scatter the values) while the compute warps already scan sub-slice B.
//     scan(A, i-1) needs only finish(A, i) -- rows of A never receive values from B -- so the latency-bound
//     finishing work is hidden behind the issue-bound scan instead of alternating with it.
//   * Warp roles:  COMPUTE warps (phase B below);  SCATTER warps (phase C, the terminal stage);  one COMM warp
//     (cost rows and halo rows by 1-D bulk TMA, neighbour counters, back-pressure on the ring);  one PUBLISHER
//     warp (fence + progress counter).  Compute and scatter hand over through shared-memory mbarriers
//     (`scanned[v]`, `finished[v]`), the comm and publisher warps follow monotone counters in shared memory, so
//     no role ever blocks another one that could make progress.
//   * Phase B (compute): the stage is a small min-plus matrix product C[b', l] = min_j (s_l + c_jl) + P[b', j].
//     A thread owns a TB x TL register tile of cells, thread groups split the successor range j (JS groups).
//     Every candidate is two separately rounded FP64 adds (:67, :71) and a strict '>' (:73) that keeps the
//     earliest successor on ties and never lets +Inf/NaN win -- exactly the reference's arithmetic.  On sm_100a
//     an FP64 instruction occupies two issue slots, so a candidate costs DADD(2)+DSETP(2)+2 FSEL+SEL = 7 slots
//     (profiles/pipe_probe_r01.txt); the kernel is issue bound, not FP64-pipe or HBM bound.
//   * Phase C (scatter): the JS partial (min, argmin) pairs of a cell are combined in ascending-j order with the
//     same strict '>' (the earliest group holding the minimum wins).  The value goes to the next stage's rows
//     (own slice: shared memory; higher slice: global ring), the argmin to HBM once per cell as uint8 indexed
//     by source row (coalesced rows).  Tiles with one j-group skip phase C: the thread scatters its finished
//     cells from registers ("direct" mode).
//   * Pruned scan (tiles with PR > 0; the default for wide level sets): the scan above is exhaustive like the
//     reference's loops, but most candidates cannot win.  For a block of PR consecutive successors,
//         LB = (s_l + min_{j in block} c_jl) + min_{j in block} P[b', j]
//     is a lower bound of every candidate of the block in the SAME floating-point arithmetic (rounded addition is
//     monotone in both operands), so if the running minimum is not greater than LB no candidate of the block can win
//     the strict '>' (HelpFunctions.jl:73) and the block is skipped -- for the whole warp, by a vote, because only a
//     warp-uniform skip saves issue slots.  Successors are still visited in ascending order, so the results are
//     bit-identical (value, argmin and ties).  A lane is a level, a warp is 32 levels x a group of rows, the two row
//     groups of a CTA run on different warps, the finished cells are scattered from registers (no phase C, no
//     scatter warps).  The block minima of the jump costs are a per-launch table, those of the value rows are
//     recomputed per stage by every warp for its own rows.
//   * Code size is a first-order concern: five roles run different code on one SM and the code executed every
//     stage has to stay below the 32 KB instruction-cache tier (it was 36 KB: 95 % hit rate, 6 % slower).  Hence
//     one out-of-line wait loop, one phase-C unit in flight per warp, +Inf pad rows in the jump-cost table (every
//     j-group runs whole trips of the unrolled scan; remainder code stays cold), 32-bit ticks, and the cycle
//     counters as a separate instantiation (template parameter PROF).
#include "bb200_internal.cuh"
#include "kernels.cuh"
#include "pruned_scan.cuh"

namespace bb200 {

// ---- PTX helpers (mbarrier / bulk TMA helpers live in bb200_internal.cuh) -----------------------------
__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// monotone counters in shared memory (CTA scope)
__device__ __forceinline__ unsigned long long lds_acquire(const uint64_t *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.cta.shared::cta.u64 %0, [%1];" : "=l"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void sts_release_u32(uint64_t *p, unsigned int v)
{
    asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_release(uint64_t *p, unsigned long long v)
{
    asm volatile("st.release.cta.shared::cta.u64 [%0], %1;" ::"r"(smem_u32(p)), "l"(v) : "memory");
}
// 32-bit event counters: native shared-memory RED; compared wrap-safe (difference as signed)
// One hand-over = ONE release fence, then relaxed signals: an arrive with release semantics and a release RED are two
// fences, and a CTA-scope fence waits for the warp's outstanding global stores (~400 cycles each after the argmin /
// ring stores of a stage).  fence + relaxed write is a release pattern of the PTX memory model.
__device__ __forceinline__ void fence_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_relaxed(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void reds_relaxed_inc(uint64_t *p)
{
    asm volatile("red.relaxed.cta.shared::cta.add.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(1u) : "memory");
}
__device__ __forceinline__ void reds_release_inc(uint64_t *p)
{
    asm volatile("red.release.cta.shared::cta.add.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(1u) : "memory");
}
__device__ __forceinline__ unsigned int lds_acquire_u32(const uint64_t *p)
{
    unsigned int v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
// has the counter reached `want` events?  (events are counted modulo 2^32; the lag is always far below 2^31)
__device__ __forceinline__ bool reached(unsigned int cnt, unsigned int want) { return (int)(cnt - want) >= 0; }

// mbarrier wait with a watchdog: a lost hand-over becomes an error code (err[2] watchdog, err[3] abort) instead of
// a hung GPU.  Once the abort flag is up every wait gives up quickly so that the launch drains.  The waiting loop is
// ONE out-of-line copy: the kernel's hot code has to stay below the 32 KB instruction-cache tier.
__device__ __noinline__ void mbar_wait_slow(uint32_t bar_addr, uint32_t parity, int *err, long long wd_cycles)
{
    unsigned int spins = 0;
    long long t0 = 0;
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar_addr), "r"(parity), "r"(kMbarSuspendNs)
            : "memory");
        if (done) return;
        if ((++spins & 0x3fu) == 0) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            if (*(volatile int *)&err[3]) {
                if (spins >= 256u) return;
            } else if (wd_cycles > 0 && now - t0 > wd_cycles) {  // default ~3 s; 0 disables (debugger, sanitizer, time slicing)
                atomicOr(&err[2], 2);
                atomicOr(&err[3], 1);
                return;
            }
        }
    }
}
__device__ __forceinline__ void mbar_wait_wd(uint64_t *bar, uint32_t parity, int *err, long long wd_cycles = 6000000000LL)
{
    if (mbar_test(bar, parity)) return;
    mbar_wait_slow(smem_u32(bar), parity, err, wd_cycles);
}

#ifndef BB_COMM_SLEEP
#define BB_COMM_SLEEP 64
#endif
#ifndef BB_PUB_SLEEP
#define BB_PUB_SLEEP 200
#endif
#ifndef BB_UNROLL
#define BB_UNROLL 4
#endif
// Value-row buffers in shared memory.  With THREE buffers the halo rows of stage i-1 may land while stage i is still being
// scanned (the buffer they go to was last read by stage i+2); with two, the TMA is issued after the scan of stage i+1.
// (Measured at config 4: no difference -- the slices are paced by ring back-pressure, not by this round trip.)
#ifndef BB_PS_BUFS
#define BB_PS_BUFS 3
#endif
#ifndef BB_KPC
#define BB_KPC 1   // 0: never launch the instantiations with a compile-time level count (A/B builds)
#endif
constexpr int kMaxRowGroups = 4;   // row groups of a pruned tile (wave_configure rejects more)
constexpr int kPsBufs = BB_PS_BUFS;
constexpr int kUnrollB = BB_UNROLL;  // successor pairs per trip of the phase-B loop
#ifndef BB_TWO_ZONE
#define BB_TWO_ZONE 1
#endif
constexpr bool kTwoZoneSlices = BB_TWO_ZONE != 0;  // pruned tiles: upper slices with one row group less (slice_r0 / slice_rows)
constexpr bool kAutoPruned = true;  // may the geometry model pick pruned tiles by itself (wide level sets only)?
constexpr int kNever = 0x7fffffff;  // "no bound": ticks are 32-bit, wave_configure refuses launches with >= 2^30 steps

// shared-memory synchronisation words
enum {
    MB_COST = 0,      // [0..2]  cost rows of global step T landed in buffer T % 3           (TMA, tx count)
    MB_HALO = 3,      // [3..5]  halo rows pushed during stage i landed, barrier i % 3        (TMA, tx count)
    MB_SCANNED = 6,   // [6..7]  compute warps finished phase B of sub-slice v                (one arrival per warp)
    MB_FINISHED = 8,  // [8..9]  scatter warps finished phase C of sub-slice v                (one arrival per warp)
    CNT_SCANNED = 10,  // counter: compute-warp arrivals after the LAST sub-slice of a stage
    CNT_FINISHED = 11,  // counter: scatter-warp arrivals after the LAST sub-slice of a step (terminal stage included)
    RING_OK = 12,     // counter: highest global step whose pushes may overwrite their ring slot
    SYNC_WORDS = 16
};

struct Smem {
    uint64_t *mbar;   // SYNC_WORDS synchronisation words, see the enum above
    double *ss;       // [3][Kp]   stage cost of global step T in ss[T%3]        (TMA destination)
    int *bts;         // [3][Kp]   budget use of global step T in bts[T%3]       (TMA destination)
    double *Ps;       // [kPsBufs][R][Kp] value rows (row-major) read by stage i in Ps[i % kPsBufs]  (halo rows: TMA destination)
    double *cs;       // [K*Kp]    jump costs
    double *pv;       // [JS*R*Kp] partial minima of the j-groups
    unsigned char *pa;  // ArgT[JS*R*Kp] partial argmins
    int *umap;        // [R*ceil(Kp/64)] phase-C work unit -> (row << 16) | first level
    float *cminf;     // pruned scan: [Kr/PR][Kp] block minima of the jump costs ROUNDED DOWN to float,
                      //              cminf[q][l] <= min_{j in block q} c_jl
    float *pminf;     // pruned scan: block minima of the CTA's value rows rounded down to float [8][Kr/PR], then the rows' seed
                      //              successors int[16], then two per-launch tables: cw[Kp/32][32] (level block, block q: the
                      //              smallest cminf[q][l] over the live levels of the level block) and cmx[Kp] (the largest
                      //              finite |jump cost| into level l, rounded up)
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Two-zone slices: the CTAs 0 .. GA-1 own R consecutive source rows each, the CTAs above them Rtop (< R) rows.  The bound
// tests of the pruned scan drop fewer blocks at high budgets, so the upper slices are the slow ones; giving them one row
// group less evens the stage time out along the pipeline and lets the slices cover all SMs (config 4: 56 x 8 + 92 x 6 rows
// on 148 SMs instead of 143 x 7).  Uniform slices are GA = G.
__host__ __device__ inline int slice_r0(const WaveCfg &c, int g) { return g < c.GA ? g * c.R : c.GA * c.R + (g - c.GA) * c.Rtop; }
__host__ __device__ inline int slice_rows(const WaveCfg &c, int g) { return g < c.GA ? c.R : c.Rtop; }

__host__ __device__ inline size_t carve(const Tables &t, const WaveCfg &c, int argw, unsigned char *base, Smem *s)
{
    size_t off = 0;
    size_t o[10];
    const int nblk = c.PR ? c.Kr / c.PR : 0;
    const int js_smem = c.PR ? 0 : c.JS;  // pruned tiles keep no partial minima in shared memory
    const size_t sizes[10] = {SYNC_WORDS * sizeof(uint64_t),
                             3 * (size_t)t.Kp * sizeof(double),
                             3 * (size_t)t.Kp * sizeof(int),
                             kPsBufs * (size_t)t.Kp * c.R * sizeof(double),
                             (size_t)c.Kr * t.Kp * sizeof(double),
                             (size_t)js_smem * c.R * t.Kp * sizeof(double),
                             (size_t)js_smem * c.R * t.Kp * (size_t)argw,
                             (size_t)c.R * (t.Kp / 32) * sizeof(int),
                             (size_t)nblk * t.Kp * sizeof(float),
                             c.PR ? (size_t)(8 * nblk + 16 + 2 * t.Kp) * sizeof(float) : 0};
    for (int k = 0; k < 10; ++k) {
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
        s->umap = reinterpret_cast<int *>(base + o[7]);
        s->cminf = reinterpret_cast<float *>(base + o[8]);
        s->pminf = reinterpret_cast<float *>(base + o[9]);
    }
    return off;
}

// Phase B of one thread: TB x TL cells, successors [jb, je): the reference's innermost loop (HelpFunctions.jl:71-76).
//   Prow: value rows of this thread's row group, row-major [r][Kp];  crow: jump costs of its levels, [j][Kp]
//   srow: stage costs of its levels;  pv/pa: partial (min, argmin) out, [r][Kp]
// Two successors per trip: P[r][j], P[r][j+1] are neighbours in the row-major value rows, so one 16-byte
// warp-broadcast load serves both (jb is even by construction).
// Cost per candidate and cell on sm_100a: DADD + DSETP (two issue slots each) + 2 FSEL + SEL = 7 slots.
template <int TB, int TL, typename ArgT>
__device__ __forceinline__ void scan_tile(const double *__restrict__ Prow, const double *__restrict__ crow,
                                          const double *__restrict__ srow, int jb, int je, int Kp,
                                          double (&best)[TB][TL], int (&arg)[TB][TL])
{
    constexpr int MARKI = (int)(ArgT) ~(ArgT)0;
    const double inf = d_inf();
#pragma unroll
    for (int a = 0; a < TB; ++a)
#pragma unroll
        for (int q = 0; q < TL; ++q) { best[a][q] = inf; arg[a][q] = MARKI; }
    double s[TL];
#pragma unroll
    for (int q = 0; q < TL; ++q) s[q] = srow[q];
    // main loop: full pairs, one straight-line block so that the two candidates' chains interleave
    const int je2 = jb + ((je - jb) & ~1);
#pragma unroll kUnrollB
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
                if (best[r][q] > v) { BB_KEEP_BRANCH; best[r][q] = v; arg[r][q] = j; }  // :73-76, strict: the earliest j wins
            }
        }
#pragma unroll
        for (int q = 0; q < TL; ++q) {
#pragma unroll
            for (int r = 0; r < TB; ++r) {
                const double v = __dadd_rn(a1[q], p1[r]);
                if (best[r][q] > v) { BB_KEEP_BRANCH; best[r][q] = v; arg[r][q] = j + 1; }
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
                if (best[r][q] > v) { BB_KEEP_BRANCH; best[r][q] = v; arg[r][q] = j; }
            }
        }
    }
}

// ---- pruned scan (pruned_scan.cuh) ---------------------------------------------------------------------------
// One tile of the pipelined kernel: bounds, then the one segment of at most 32 blocks its jump-cost table allows.
template <int TB, int BK, typename ArgT, bool PROF>
__device__ __forceinline__ void scan_pruned(const double *__restrict__ Prow, const double *__restrict__ cs_l,
                                            const float *__restrict__ cmf_l, const float *__restrict__ pmf, int row0,
                                            const int *__restrict__ qseed, double s, float cw, float cmx, int nblk,
                                            int Kp, bool live, int rows_live, int lane, const double (&vself)[TB], double (&best)[TB][1],
                                            int (&arg)[TB][1], unsigned int &executed, long long (&ph)[4])
{
    constexpr int MARKI = (int)(ArgT) ~(ArgT)0;
    long long tq = 0;
    if constexpr (PROF) tq = clock64();
#define PH_LAP(k) do { if constexpr (PROF) { const long long tn = clock64(); ph[k] += tn - tq; tq = tn; } } while (0)
    PrunedBounds<TB> pb;
    pruned_bounds<TB>(Prow, cs_l, qseed, s, cmx, Kp, live, rows_live, vself, pb);
    PH_LAP(1);
#pragma unroll
    for (int r = 0; r < TB; ++r) { best[r][0] = d_inf(); arg[r][0] = MARKI; }
    executed += pruned_segment<TB, BK, ArgT>(Prow, cs_l, cmf_l, pmf, row0, pb, s, cw, nblk, Kp, live, lane, 0, best, arg);
    PH_LAP(3);
#undef PH_LAP
}

// scan + partial (min, argmin) of this thread's j-group into shared memory for phase C
template <int TB, int TL, typename ArgT>
__device__ __forceinline__ void phase_b(const double *__restrict__ Prow, const double *__restrict__ crow,
                                        const double *__restrict__ srow, double *__restrict__ pv,
                                        ArgT *__restrict__ pa, int jb, int je, int Kp)
{
    double best[TB][TL];
    int arg[TB][TL];
    scan_tile<TB, TL, ArgT>(Prow, crow, srow, jb, je, Kp, best, arg);
#pragma unroll
    for (int r = 0; r < TB; ++r)
#pragma unroll
        for (int q = 0; q < TL; ++q) {
            pv[(size_t)r * Kp + q] = best[r][q];
            pa[(size_t)r * Kp + q] = (ArgT)arg[r][q];
        }
}

// ===================================== SCATTER warps (phase C) ====================================
struct FinishArgs {
    const double *pv;
    const unsigned char *pa;  // ArgT[]
    const int *bt;            // budget use of this stage's levels
    const int *umap;          // work unit (64 levels of one row) -> (row << 16) | first level
    double *Pn;               // rows of the next stage, [R][Kp]
    double *hring;            // ring slot of this step at row r0, [B1 - r0][Kp]
    double *phi;              // exit slot that receives this stage's values (i <= 2) at row r0, or nullptr
    unsigned char *argrow;    // ArgT*: argmin table of this stage at source row r0
    int JS, R, Kp, K, B1, r0;
    int Rown;                 // rows this CTA owns (<= R, the row stride of the shared-memory arrays)
};

// Finishes the work units [ub, ue) of a sub-slice; warp sw of NS takes every NS-th one.  A unit is 64 consecutive
// levels of one source row: a lane owns EC = 2 neighbouring cells and fetches their partial minima with one 16-byte load
// per j-group and both argmins with one 2-byte load.  (Four cells per lane -- a whole 128-level row per unit -- shortened
// the finishing pass but its larger code slowed the scan by more, profiles/README.md; the variant was removed.)
template <int JSC, int EC, typename ArgT, bool PROF>
__device__ __forceinline__ void finish_rows(const FinishArgs &a, int ub, int ue, int sw, int NS, int lane, long long *pcc)
{
    static_assert(sizeof(ArgT) == 1, "the packed argmin load assumes one byte per cell");
    static_assert(EC == 2, "two cells per lane");
    long long tq0 = 0;
    if constexpr (PROF) tq0 = pcc ? clock64() : 0;
    const double inf = d_inf();
    const double *__restrict__ pv = a.pv;
    const unsigned char *__restrict__ pa = a.pa;
    const int *__restrict__ btp = a.bt;
    const int *__restrict__ umap = a.umap;
    unsigned char *__restrict__ argrow = a.argrow;
    double *__restrict__ Pn = a.Pn;
    double *__restrict__ hring = a.hring;
    const int RK = a.R * a.Kp;
    const int rows_left = a.B1 - a.r0;  // rows of this slice that exist in the table
#pragma unroll 1
    for (int unit = ub + sw; unit < ue; unit += NS) {
        const int m = umap[unit];
        const int row = m >> 16;
        const int l0 = min((m & 0xffff) + EC * lane, a.Kp - EC);  // a partly filled last unit re-reads the last cells
        const bool lane_live = (m & 0xffff) + EC * lane < a.Kp;
        const int x = row * a.Kp + l0;
        const int2 b2 = *reinterpret_cast<const int2 *>(btp + l0);
        const int bt[EC] = {b2.x, b2.y};
        double val[EC];
        int arg[EC], y[EC];
        bool no_src[EC], ok[EC];
#pragma unroll
        for (int e = 0; e < EC; ++e) {
            const bool in_tab = lane_live && l0 + e < a.K && row < rows_left;
            y[e] = x + e + bt[e] * a.Kp;  // (target row - r0) * Kp + l
            // target cell (bsrc, l) of my rows has no source row when bsrc < b~_l: +Inf (:47).  This also covers
            // levels that are unreachable at this stage (b~ clamped to B1).
            no_src[e] = in_tab && a.r0 + row < bt[e];
            // cells outside `for b = 0:B-b~` (:69) are not computed by the reference either
            ok[e] = in_tab && row + bt[e] < rows_left;
        }
        auto load = [&](int q, double (&v)[EC], unsigned int &g) {
#pragma unroll
            for (int h = 0; h < EC / 2; ++h) {
                const double2 w = *reinterpret_cast<const double2 *>(pv + q * RK + x + 2 * h);
                v[2 * h] = w.x;
                v[2 * h + 1] = w.y;
            }
            g = (unsigned int)*reinterpret_cast<const unsigned short *>(pa + q * RK + x);
        };
        unsigned int gsel[EC];  // the packed argmins of the group that currently wins cell e
        if constexpr (JSC > 0) {
            // all partial (min, argmin) pairs are loaded before the first compare; tournament in ascending group order,
            // on ties the earlier group stays (strict '>', HelpFunctions.jl:73).  Partial minima are never NaN and
            // carry MARK with +Inf, so this equals the sequential scan.
            double v[JSC][EC];
            unsigned int g[JSC];
#pragma unroll
            for (int q = 0; q < JSC; ++q) load(q, v[q], g[q]);
            unsigned int gw[JSC][EC];
#pragma unroll
            for (int q = 0; q < JSC; ++q)
#pragma unroll
                for (int e = 0; e < EC; ++e) gw[q][e] = g[q];
#pragma unroll
            for (int w = 1; w < JSC; w *= 2)
#pragma unroll
                for (int q = 0; q + w < JSC; q += 2 * w)
#pragma unroll
                    for (int e = 0; e < EC; ++e)
                        if (v[q][e] > v[q + w][e]) { v[q][e] = v[q + w][e]; gw[q][e] = gw[q + w][e]; }
#pragma unroll
            for (int e = 0; e < EC; ++e) { val[e] = v[0][e]; gsel[e] = gw[0][e]; }
        } else {
            // any other split: sequential scan over the groups
#pragma unroll
            for (int e = 0; e < EC; ++e) { val[e] = inf; gsel[e] = 0xffffffffu; }
#pragma unroll 1
            for (int q = 0; q < a.JS; ++q) {
                double v[EC];
                unsigned int g;
                load(q, v, g);
#pragma unroll
                for (int e = 0; e < EC; ++e)
                    if (val[e] > v[e]) { val[e] = v[e]; gsel[e] = g; }  // strict: earliest group wins ties
            }
        }
#pragma unroll
        for (int e = 0; e < EC; ++e) arg[e] = (int)((gsel[e] >> (8 * e)) & 0xffu);
        if constexpr (PROF) { if (pcc) { const long long tq = clock64(); pcc[0] += tq - tq0; tq0 = tq; } }
#pragma unroll
        for (int e = 0; e < EC; ++e) {
            if (no_src[e]) Pn[x + e] = inf;
            // written once, streamed (do not displace the ring in L2).  EVERY byte of an owned row is written -- pad levels
            // and cells outside `for b = 0:B-b~` get MARK -- so that no 32-byte sector of the table is written partially
            // (a partial sector costs a DRAM read-modify-write: 3.6 GB of reads per launch at config 4)
            if (lane_live && row < rows_left) __stcs(argrow + x + e, ok[e] ? (unsigned char)arg[e] : (unsigned char)0xff);
            if (ok[e] && y[e] < RK) Pn[y[e]] = val[e];
            if (ok[e] && y[e] >= RK) hring[y[e]] = val[e];
        }
        if (a.phi) {  // stages 2 and 1 are the exit state (S7)
#pragma unroll
            for (int e = 0; e < EC; ++e)
                if (ok[e]) a.phi[y[e]] = val[e];
        }
        if constexpr (PROF) { if (pcc) { const long long tq = clock64(); pcc[1] += tq - tq0; tq0 = tq; } }
    }
}

// With a single j-group a thread's tile already holds the final (min, argmin) of its cells: it stores the argmin
// and scatters the values straight from registers -- no partials in shared memory, no second hand-over.
template <int TB, int TL, typename ArgT>
__device__ __forceinline__ void scatter_tile(const FinishArgs &a, int row0, int l0, const double (&best)[TB][TL],
                                             const int (&arg)[TB][TL])
{
    const double inf = d_inf();
    ArgT *argrow = reinterpret_cast<ArgT *>(a.argrow);
    const int RK = a.Rown * a.Kp;   // targets below stay in shared memory, the others belong to higher slices
    const int rows_left = a.B1 - a.r0;
#pragma unroll
    for (int q = 0; q < TL; ++q) {
        const int l = l0 + q;
        const int bt = a.bt[min(l, a.Kp - 1)];
#pragma unroll
        for (int r = 0; r < TB; ++r) {
            const int row = row0 + r;
            const bool in_tab = l < a.K && row < rows_left && row < a.Rown;  // (a ragged last row group reaches past the CTA's rows)
            const int x = row * a.Kp + l;
            const int y = x + bt * a.Kp;
            if (in_tab && a.r0 + row < bt) a.Pn[x] = inf;  // no source row: +Inf (:47)
            const bool okc = in_tab && row + bt < rows_left;  // inside `for b = 0:B-b~` (:69)
            // argmin: written once, streamed; every byte of an owned row is written (MARK in pad levels and out-of-range
            // cells) so that no 32-byte sector is written partially (see finish_rows)
            if (row < rows_left && row < a.Rown && l < a.Kp) __stcs(argrow + x, okc ? (ArgT)arg[r][q] : (ArgT) ~(ArgT)0);
            if (okc) {
                if (y < RK) a.Pn[y] = best[r][q];
                else a.hring[y] = best[r][q];
                if (a.phi) a.phi[y] = best[r][q];
            }
        }
    }
}

// Phase C and the terminal stage for the rows of one CTA, executed by NF "finisher" warps: the scatter warps, or --
// for tiles without a second sub-slice -- the compute warps themselves after their scan.
template <typename ArgT, bool PROF>
struct Finisher {
    const Tables &t;
    const WaveCfg &c;
    const Smem &sm;
    const int fw, NF, lane;  // this warp's index among the NF finisher warps
    const int r0, myR, lblocks;  // first row and number of rows of this CTA's slice; 32-level blocks per row (terminal stage)
    const int ublocks;      // work units (32 * EC levels) per row (phase C)
    bool halo_on, pushes;
    uint32_t halo_phase = 0;
    long long pcc[2] = {0, 0};  // profile: loads + combine, stores
    long long ring_wait = 0;    // profile: cycles spent waiting for ring space

    __device__ __forceinline__ Finisher(const Tables &t_, const WaveCfg &c_, const Smem &sm_, int fw_, int NF_, int lane_)
        : t(t_), c(c_), sm(sm_), fw(fw_), NF(NF_), lane(lane_), r0(slice_r0(c_, blockIdx.x)), myR(slice_rows(c_, blockIdx.x)), lblocks(t_.Kp >> 5), ublocks((t_.Kp + 32 * c_.EC - 1) / (32 * c_.EC))
    {
        const int btm = min(*c.btmax, t.B1 - 1);
        halo_on = (blockIdx.x > 0 && btm > 0);  // lower slices push into this one
        pushes = (blockIdx.x + 1 < (unsigned)c.G);
    }

    // terminal stage n (HelpFunctions.jl:27-43) = the rows stage n-1 reads:  P[b][l] = (b == b~_l(n)) ? s_l(n) : Inf
    __device__ __forceinline__ void terminal(const SlotDev &sl, int T)
    {
        const int K = t.K, Kp = t.Kp, B1 = t.B1, R = c.R, n = t.n;
        const double inf = d_inf();
        const double *sn = sm.ss + (size_t)(T % 3) * Kp;
        const int *bn = sm.bts + (size_t)(T % 3) * Kp;
        double *Pw = sm.Ps + (size_t)((n - 1) % kPsBufs) * R * Kp;
        // n == 1: only the terminal stage exists, it is the exit state (slot 1 of the reference); n == 2: it is slot 2
        double *phi = (n == 1) ? sl.phi : (n == 2) ? sl.phi + (size_t)B1 * Kp : nullptr;
        for (int unit = fw; unit < R * lblocks; unit += NF) {
            const int row = unit / lblocks, l = ((unit - row * lblocks) << 5) + lane, b = r0 + row;
            if (l < K && b < B1 && row < myR) {
                const double v = (b == bn[l]) ? sn[l] : inf;
                Pw[row * Kp + l] = v;
                if (phi) phi[(size_t)b * Kp + l] = v;
            }
        }
    }

    // before the first store of stage i: halo rows landed, ring slot free
    __device__ __forceinline__ void wait_inputs(int i, int T)
    {
        // my next rows receive the lower slices' block by TMA (issued by the comm warp): it must have landed before
        // my own results overwrite the cells I produce myself
        if (halo_on && i >= 2) {
            mbar_wait_wd(&sm.mbar[MB_HALO + (i % 3)], (halo_phase >> (i % 3)) & 1u, c.err, c.wd_cycles);
            halo_phase ^= 1u << (i % 3);
        }
        // back-pressure: the successors consumed the ring slot this step overwrites
        long long tr0 = 0;
        if constexpr (PROF) tr0 = clock64();
        if (pushes) {
            unsigned int spins = 0;
            while ((int)lds_acquire_u32(&sm.mbar[RING_OK]) < T) {
                __nanosleep(64);
                if ((++spins & 0xfffffu) == 0 && *(volatile int *)&c.err[3]) break;  // the comm warp gave up
            }
        }
        if constexpr (PROF) ring_wait += clock64() - tr0;
    }

    __device__ __forceinline__ FinishArgs stage_args(const SlotDev &sl, int i, int T) const
    {
        const int Kp = t.Kp, B1 = t.B1, R = c.R;
        FinishArgs a;
        a.pv = sm.pv;
        a.pa = sm.pa;
        a.bt = sm.bts + (size_t)(T % 3) * Kp;
        a.umap = sm.umap;
        a.Pn = sm.Ps + (size_t)((i - 1) % kPsBufs) * R * Kp;
        a.hring = c.halo + ((size_t)(T % kHaloRing) * B1 + r0) * Kp;
        a.phi = (i <= 2) ? sl.phi + ((size_t)((i + 1) & 1) * B1 + r0) * Kp : nullptr;
        a.argrow = reinterpret_cast<unsigned char *>(sl.arg) + ((size_t)(i - 1) * B1 + r0) * Kp * sizeof(ArgT);
        a.JS = c.JS; a.R = R; a.Kp = Kp; a.K = t.K; a.B1 = B1; a.r0 = r0; a.Rown = myR;
        return a;
    }

    // phase C of work units [ub, ue) of the stage described by `a` (stage_args)
    __device__ __forceinline__ void rows(const FinishArgs &a, int ub, int ue)
    {
        long long *pccp = (PROF && c.prof) ? pcc : nullptr;
        switch (c.JS) {
            case 2: finish_rows<2, 2, ArgT, PROF>(a, ub, ue, fw, NF, lane, pccp); break;
            case 4: finish_rows<4, 2, ArgT, PROF>(a, ub, ue, fw, NF, lane, pccp); break;
            default: finish_rows<0, 2, ArgT, PROF>(a, ub, ue, fw, NF, lane, pccp); break;
        }
    }
};

template <typename ArgT, bool PROF>
__device__ __forceinline__ void scatter_warp(const Tables &t, const WaveCfg &c, const Smem &sm, int sw, int lane)
{
    const int g = blockIdx.x;
    const int R = c.R, n = t.n;
    const int NV = c.RB > 0 ? 2 : 1;
    Finisher<ArgT, PROF> fin(t, c, sm, sw, c.NS, lane);
    const int ublocks = fin.ublocks;
    uint32_t cost_phase = 0, scanned_phase = 0;
    long long pc[3] = {0, 0, 0};  // profile (warp 0): wait for the scan, wait for halo / ring, work
    long long tp = clock64();
#define PROF_LAP(k) do { if constexpr (PROF) { if (c.prof) { const long long tq = clock64(); pc[k] += tq - tp; tp = tq; } } } while (0)
    auto wait_costs = [&](int T) {
        const int b = (int)(T % 3);
        mbar_wait_wd(&sm.mbar[MB_COST + b], (cost_phase >> b) & 1u, c.err, c.wd_cycles);
        cost_phase ^= 1u << b;
    };
    auto finished = [&](int v) {
        __syncwarp();
        if (lane == 0) {
            fence_cta();
            mbar_arrive_relaxed(&sm.mbar[MB_FINISHED + v]);
            if (v == NV - 1) reds_relaxed_inc(&sm.mbar[CNT_FINISHED]);
        }
    };

    int T = 0;  // global step: subproblem * n + (n - stage)
    for (int sub = 0; sub < c.nsub; ++sub) {
        const SlotDev sl = c.slots[sub];
        wait_costs(T);
        fin.terminal(sl, T);
        for (int v = 0; v < NV; ++v) finished(v);
        ++T;
        for (int i = n - 1; i >= 1; --i, ++T) {
            wait_costs(T);
            const FinishArgs fa = fin.stage_args(sl, i, T);  // once per stage, shared by both sub-slices
            // everything that does not need the scan happens while the compute warps are still scanning sub-slice A:
            // the halo rows of this stage have landed, the ring slot is free
            fin.wait_inputs(i, T);
            PROF_LAP(1);
            for (int v = 0; v < NV; ++v) {
                mbar_wait_wd(&sm.mbar[MB_SCANNED + v], (scanned_phase >> v) & 1u, c.err, c.wd_cycles);
                scanned_phase ^= 1u << v;
                PROF_LAP(0);
                const int ub = v == 0 ? 0 : c.RA * ublocks, ue = v == 0 ? c.RA * ublocks : R * ublocks;
                if (!PROF || c.decouple < 2) fin.rows(fa, ub, ue);
                finished(v);
                PROF_LAP(2);
            }
        }
    }
    if (PROF && c.prof && sw == 0 && lane == 0) {
        for (int k = 0; k < 3; ++k) c.prof[(size_t)g * 16 + 5 + k] = pc[k];
        c.prof[(size_t)g * 16 + 14] = fin.pcc[0];
        c.prof[(size_t)g * 16 + 15] = fin.pcc[1];
        c.prof[(size_t)g * 16 + 13] = fin.ring_wait;
    }
#undef PROF_LAP
}

// ======================================= COMM warp ===============================================
// Event loop over three independent cursors, none of which ever blocks the others:
//   cost:  level costs / budget uses of global step T (rows of the prep kernel's tables) into buffer T % 3 as soon
//          as step T-3 has been finished by the scatter warps;
//   halo:  my value rows as the lower slices pushed them during stage i, one bulk TMA straight into the rows stage
//          i-1 reads, as soon as the predecessors have published stage i and the compute warps no longer read
//          that buffer (they finished scanning stage i+1);
//   ring:  how far this CTA's pushes may run ahead of the slowest successor (ring of kHaloRing slots).
// A bounded watchdog turns a lost dependency into an error code instead of a hung GPU.
__device__ __forceinline__ int warp_min(int v) { return __reduce_min_sync(0xffffffffu, v); }

template <bool PROF>
__device__ __forceinline__ void comm_warp(const Tables &t, const WaveCfg &c, const Smem &sm, int lane)
{
    const int g = blockIdx.x;
    const int r0 = slice_r0(c, g);
    const int Kp = t.Kp, B1 = t.B1, R = c.R, n = t.n, NF = c.NF;
    const int ncw = (c.JS * c.tpg) >> 5;
    const int btm = min(*c.btmax, B1 - 1);
    const int Rmin = min(R, c.Rtop);
    const int D = (btm + Rmin - 1) / Rmin;  // slices a push can span (an upper bound where the slices have R rows)
    const int my_rows = min(slice_rows(c, g), B1 - r0);  // rows of this slice that exist in the table (>= 1)
    const int npred = min(D, g), nsucc = min(D, c.G - 1 - g);
    const int Ttot = c.nsub * n;
    // cost cursor
    int cT = 0;
    int csub = 0, ck = 0;
    // halo cursor: stage i = n - hk, hk in [1, n-2]
    bool h_active = (npred > 0 && n >= 3);
    int hsub = 0, hk = 1;
    int pred_seen = 0;
    // ring cursor
    int ring_val = (c.G - 1 - g > 0) ? (kHaloRing - 1) : kNever;  // value set at kernel start
    if (nsucc == 0 && ring_val != kNever) {  // successors exist but never receive pushes (b~ = 0 everywhere)
        ring_val = kNever;
        if (lane == 0) sts_release_u32(&sm.mbar[RING_OK], (unsigned int)ring_val);
    }
    if constexpr (PROF) {
        if (c.decouple) {  // timing experiment of the profiling build only (results are wrong): ignore the neighbours
            pred_seen = kNever;
            ring_val = kNever;
            if (lane == 0) sts_release_u32(&sm.mbar[RING_OK], (unsigned int)ring_val);
        }
    }
    unsigned int idle = 0;
    long long pc[6] = {0, 0, 0, 0, 0, 0};  // profile: loop trips, idle trips, pred polls, succ polls, SM id, cost rows loaded

    for (;;) {
        bool progress = false;
        const unsigned int fin_cnt = lds_acquire_u32(&sm.mbar[CNT_FINISHED]);  // NF arrivals per finished step
        const unsigned int scn_cnt = lds_acquire_u32(&sm.mbar[CNT_SCANNED]);   // ncw arrivals per scanned stage
        pc[0] += 1;
        // ---- cost cursor --------------------------------------------------------------------------------
        while (cT < Ttot && (cT < 3 || reached(fin_cnt, (unsigned int)(cT - 2) * (unsigned int)NF))) {
            if (lane == 0) {
                const SlotDev &sl = c.slots[csub];
                const int b = (int)(cT % 3);
                const int row = n - ck - 1;  // stage i = n - ck uses row i-1 of the prep tables
                uint64_t *bar = &sm.mbar[MB_COST + b];
                mbar_expect_tx(bar, (uint32_t)(Kp * (sizeof(double) + sizeof(int))));
                tma_load_1d(sm.ss + (size_t)b * Kp, sl.ss_all + (size_t)row * Kp, (uint32_t)(Kp * sizeof(double)), bar);
                tma_load_1d(sm.bts + (size_t)b * Kp, sl.bt_all + (size_t)row * Kp, (uint32_t)(Kp * sizeof(int)), bar);
            }
            ++cT;
            if (++ck == n) { ck = 0; ++csub; }
            progress = true;
            pc[5] += 1;
        }
        // ---- halo cursor --------------------------------------------------------------------------------
        if (h_active) {
            const int hT = hsub * n + hk;
            if (pred_seen < hT) {
                int m = kNever;
#pragma unroll 1
                for (int idx = lane; idx < npred; idx += 32)
                    m = min(m, (int)ld_acquire(c.flags + (size_t)(g - idx - 1) * kFlagStride));
                pred_seen = max(pred_seen, warp_min(m));
                __syncwarp();  // every polling lane finished its acquire load; lane 0 inherits the order
                pc[2] += 1;
            }
            // the buffer the block lands in was last read by the scan of stage i + kPsBufs - 1 (or by the previous subproblem)
            const int need_scanned = hsub * (n - 1) + max(hk - (kPsBufs - 1), 0);
            // Pruned tiles with three buffers send no "scanned" signal: the block of stage i lands in the buffer stage i+2
            // read, and a step that every warp has FINISHED has also been scanned -- step hT - 2 (or the whole previous
            // subproblem) is complete when (hT - 1) NF arrivals have been counted.  One fence and one hand-over less per
            // warp and stage; the block still has a whole stage to land.
            const bool buffer_free = (c.PR > 0 && kPsBufs == 3) ? reached(fin_cnt, (unsigned int)(hT - 1) * (unsigned int)NF)
                                                                : reached(scn_cnt, (unsigned int)need_scanned * (unsigned int)ncw);
            if (pred_seen >= hT && buffer_free) {
                if (lane == 0) {
                    const int i = n - hk;
                    asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy acquire -> async-proxy read
                    const uint32_t bytes = (uint32_t)((size_t)my_rows * Kp * sizeof(double));
                    uint64_t *bar = &sm.mbar[MB_HALO + (i % 3)];
                    mbar_expect_tx(bar, bytes);
                    tma_load_1d(sm.Ps + (size_t)((i - 1) % kPsBufs) * R * Kp,
                                c.halo + ((size_t)(hT % kHaloRing) * B1 + r0) * Kp, bytes, bar);
                }
                if (++hk > n - 2) { hk = 1; if (++hsub >= c.nsub) h_active = false; }
                progress = true;
            }
        }
        // ---- ring cursor --------------------------------------------------------------------------------
        // scatter warps work on step >= cT - 3 at most cT: keep the ring bound a few steps ahead of them
        if (ring_val < Ttot && ring_val < cT + 2) {
            int m = kNever;
#pragma unroll 1
            for (int idx = lane; idx < nsucc; idx += 32)
                m = min(m, (int)ld_acquire(c.flags + (size_t)(g + idx + 1) * kFlagStride));
            m = warp_min(m);
            __syncwarp();
            pc[3] += 1;
            const int nv = (m >= kNever) ? kNever : m + kHaloRing - 1;
            if (nv > ring_val) {
                ring_val = nv;
                if (lane == 0) sts_release_u32(&sm.mbar[RING_OK], (unsigned int)ring_val);
                progress = true;
            }
        }
        if (cT >= Ttot && !h_active && ring_val >= Ttot) break;
        if (progress) { idle = 0; continue; }
        __nanosleep(BB_COMM_SLEEP);
        pc[1] += 1;
        if ((++idle & 0x3ffu) == 0) {
            bool abort_now = *(volatile int *)&c.err[3] != 0;
            if (!abort_now && c.wd_cycles > 0 && idle > (1u << 22)) {
                atomicOr(&c.err[2], 1);
                atomicOr(&c.err[3], 1);
                abort_now = true;
            }
            if (abort_now) {  // give up on the neighbours so that this CTA drains and the launch ends
                pred_seen = kNever;
                ring_val = kNever;
                if (lane == 0) sts_release_u32(&sm.mbar[RING_OK], (unsigned int)ring_val);
            }
        }
    }
    if (c.prof && lane == 0) {
        unsigned int smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        pc[4] = smid;  // which SM hosted this slice
        for (int k = 0; k < 5; ++k) c.prof[(size_t)g * 16 + 8 + k] = pc[k];  // slot 13: the scatter warps' ring wait
    }
}

// ===================================== PUBLISHER warp ============================================
// Follows the scatter warps' counter; for every finished stage: fence.acq_rel.gpu so that the pushes made during
// that stage are visible GPU-wide, then the relaxed store of the progress counter -- about 1 200 cycles that sit on
// nobody's critical path.
__device__ __forceinline__ void publisher_warp(const Tables &t, const WaveCfg &c, const Smem &sm, int lane)
{
    if (lane != 0) return;
    const int g = blockIdx.x;
    const int n = t.n;
    unsigned long long *myflag = c.flags + (size_t)g * kFlagStride;
    int T = 0;
    for (int sub = 0; sub < c.nsub; ++sub) {
        ++T;  // the terminal stage pushes nothing
        for (int i = n - 1; i >= 1; --i, ++T) {
            const unsigned int want = (unsigned int)(T + 1) * (unsigned int)c.NF;
            unsigned int spins = 0;
            while (!reached(lds_acquire_u32(&sm.mbar[CNT_FINISHED]), want)) {
                __nanosleep(BB_PUB_SLEEP);  // a quiet poll: this warp shares a scheduler with compute warps
                if ((++spins & 0xffffu) == 0 && *(volatile int *)&c.err[3] && spins > (1u << 22)) return;  // stuck after an abort
            }
            fence_gpu();
            st_relaxed(myflag, (unsigned long long)T);
        }
    }
}

// MAXT is a multiple of 128: the register file is split evenly over the four schedulers, so the per-thread
// budget is set by the scheduler that hosts the most warps.
// KPC > 0: the padded level count Kp (= Kr for pruned tiles) is the compile-time constant KPC -- every row stride, block
// count and table offset of the stage loop folds into immediates (the production shape K = 125 runs with KPC = 128).
template <int TBA, int TBB, int TL, typename ArgT, bool PROF, int PR>
__device__ __forceinline__ void wavefront_body(const Tables &t, const WaveCfg &c)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem sm;
    carve(t, c, (int)sizeof(ArgT), smem_raw, &sm);

    const int tid = threadIdx.x;
    const int NC = c.JS * c.tpg;  // compute threads, then: comm warp, publisher warp, NS scatter warps
    const int g = blockIdx.x;
    const int K = t.K, Kp = t.Kp, R = c.R, n = t.n;
    constexpr int NV = (TBB > 0 && PR == 0) ? 2 : 1;  // the pruned scan runs its two row groups side by side: one hand-over
    const double inf = d_inf();

    // one-time: jump costs into shared memory, value rows to +Inf, barriers and counters
    for (int x = tid; x < c.Kr * Kp; x += blockDim.x) sm.cs[x] = x < K * Kp ? t.cost[x] : inf;  // pad rows: never win
    for (int x = tid; x < kPsBufs * R * Kp; x += blockDim.x) sm.Ps[x] = inf;
    {
        const int ul = 32 * c.EC, ubl = (Kp + ul - 1) / ul;  // levels per phase-C work unit, units per row
        for (int x = tid; x < R * ubl; x += blockDim.x) sm.umap[x] = ((x / ubl) << 16) | ((x % ubl) * ul);
    }
    if (tid == 0) {
        for (int k = 0; k < 6; ++k) mbar_init(&sm.mbar[k], 1);  // cost[3], halo[3]: one arming arrival + tx bytes
        for (int v = 0; v < 2; ++v) {
            mbar_init(&sm.mbar[MB_SCANNED + v], NC >> 5);
            mbar_init(&sm.mbar[MB_FINISHED + v], c.NF);
        }
        sm.mbar[CNT_SCANNED] = 0;
        sm.mbar[CNT_FINISHED] = 0;
        sm.mbar[RING_OK] = (g + 1 < c.G) ? (uint64_t)(kHaloRing - 1) : (uint64_t)kNever;  // read as a 32-bit tick
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if constexpr (PR > 0) {  // block minima of the jump costs, rounded down to float (pad rows are +Inf and never lower a minimum)
        const int nblk = c.Kr / PR;
        for (int x = tid; x < nblk * Kp; x += blockDim.x) {
            const int q = x / Kp, l = x - q * Kp;
            double m = inf;
#pragma unroll
            for (int jj = 0; jj < PR; ++jj) m = fmin(m, sm.cs[(size_t)(q * PR + jj) * Kp + l]);
            sm.cminf[x] = __double2float_rd(m);
        }
        for (int x = tid; x < 8 * nblk; x += blockDim.x) sm.pminf[x] = __int_as_float(0x7f800000);  // rows the CTA does not own bound nothing
        if (tid < 16) reinterpret_cast<int *>(sm.pminf + 8 * nblk)[tid] = 0;
        __syncthreads();
    }

    if (tid >= NC + 64) {
        scatter_warp<ArgT, PROF>(t, c, sm, (tid - NC - 64) >> 5, tid & 31);
        return;
    }
    if (tid >= NC + 32) {
        publisher_warp(t, c, sm, tid - NC - 32);
        return;
    }
    if (tid >= NC) {
        comm_warp<PROF>(t, c, sm, tid - NC);
        return;
    }

    // ========================================= COMPUTE warps ============================================
    // Pruned scan (PR > 0): a lane is a level, a warp is a block of 32 levels x a group of TBA consecutive rows (the
    // last group may be ragged: the CTA owns TBB rows); the warps of a row group are consecutive, so that with four
    // level blocks every scheduler hosts one warp of each group.
    const int nLB = Kp >> 5;
    const int pr_grp = PR > 0 ? (tid >> 5) / nLB : 0;
    const int pr_l = PR > 0 ? (((tid >> 5) % nLB) << 5) + (tid & 31) : 0;
    const int jg = PR > 0 ? 0 : tid / c.tpg, tig = tid % c.tpg;
    const bool active = PR > 0 ? pr_l < K : tig < c.RG * c.nLG;
    const int rg = (active && PR == 0) ? tig / c.nLG : 0;
    const int lg = PR > 0 ? min(pr_l, Kp - 1) : (active ? tig % c.nLG : 0);
    const int jb = jg * c.jper;
    const int je = min(c.Kr, jb + c.jper);  // rows K .. Kr-1 of the cost table are +Inf: (s + Inf) + P never wins
    const int lane = tid & 31;
    uint32_t fin_phase = 0;
    int T = 0;  // global step: subproblem * n + (n - stage)
    long long pc[5] = {0, 0, 0, 0, 0};  // profile: wait A (+ costs), phase B, hand-over, wait B, stages
    long long tp = clock64();
#define PROF_LAP(k) do { if constexpr (PROF) { if (c.prof) { const long long tq = clock64(); pc[k] += tq - tp; tp = tq; } } } while (0)
    auto wait_costs = [&](int T) {
        const int b = (int)(T % 3);   // step T is completion number T / 3 of barrier T % 3: no phase word to keep in a register
        mbar_wait_wd(&sm.mbar[MB_COST + b], (uint32_t)(T / 3) & 1u, c.err, c.wd_cycles);
    };
    auto wait_finished = [&](int v) {
        if constexpr (PR > 0) {  // one hand-over per step: the wait before step T is completion number T - 1
            mbar_wait_wd(&sm.mbar[MB_FINISHED], (uint32_t)(T - 1) & 1u, c.err, c.wd_cycles);
        } else {
            mbar_wait_wd(&sm.mbar[MB_FINISHED + v], (fin_phase >> v) & 1u, c.err, c.wd_cycles);
            fin_phase ^= 1u << v;
        }
    };
    auto scanned = [&](int v) {
        __syncwarp();
        if (lane == 0) {
            fence_cta();
            mbar_arrive_relaxed(&sm.mbar[MB_SCANNED + v]);
            if (v == NV - 1) reds_relaxed_inc(&sm.mbar[CNT_SCANNED]);
        }
    };
    const int rowA = PR > 0 ? pr_grp * TBA : rg * TBA, rowB = PR > 0 ? 0 : c.RA + rg * TBB;
    // pruned scan, once per launch: (lane = block q) the smallest block minimum of the jump costs over the live levels of this
    // warp's level block; (lane = level) the largest finite jump cost into my level, rounded up.  Kept in shared memory and
    // re-read every stage: two registers less in a kernel that is capped at 96.
    if constexpr (PR > 0) {
        const int nblk = c.Kr / PR, l0 = ((tid >> 5) % nLB) << 5;
        float *cwv = sm.pminf + 8 * nblk + 16, *cmxv = cwv + Kp;
        if (pr_grp == 0) {
            float cw = __int_as_float(0x7f800000), cmx = 0.f;
            if (lane < nblk)
#pragma unroll 1
                for (int l = l0; l < min(l0 + 32, K); ++l) cw = fminf(cw, sm.cminf[(size_t)lane * Kp + l]);
#pragma unroll 1
            for (int j = 0; j < K; ++j) {
                const double cj = sm.cs[(size_t)j * Kp + lg];
                if (fabs(cj) < inf) cmx = fmaxf(cmx, __double2float_ru(fabs(cj)));
            }
            cwv[l0 + lane] = cw;
            cmxv[lg] = cmx;
        }
        asm volatile("bar.sync %0, %1;" ::"r"(kMaxRowGroups + 1), "r"(NC) : "memory");  // compute warps only
    }
    unsigned int executed = 0;  // pruned scan: blocks this warp really scanned
    long long ph[4] = {0, 0, 0, 0};  // profile of the pruned scan: block minima + seed, upper bounds, masks, scan
    ArgT *pa = reinterpret_cast<ArgT *>(sm.pa);

    // tiles without a second sub-slice have no scatter warps: the compute warps finish their own stage
    const bool self_finish = (TBB == 0 || PR > 0) && (c.NS == 0);  // two sub-slices one after the other always have scatter warps
    const bool direct = self_finish && c.JS == 1;         // final values stay in registers until they are scattered
    const int cwarp = tid >> 5;
    Finisher<ArgT, PROF> fin(t, c, sm, cwarp, NC >> 5, lane);
    uint32_t scanned_phase = 0;
    const int pr_rows_live = min(fin.myR, t.B1 - fin.r0) - rowA;  // pruned tiles: rows of my row group that exist (<= 0: none)
    auto finished = [&]() {
        __syncwarp();
        if (lane == 0) {
            fence_cta();
            mbar_arrive_relaxed(&sm.mbar[MB_FINISHED]);
            reds_relaxed_inc(&sm.mbar[CNT_FINISHED]);
        }
    };

    for (int sub = 0; sub < c.nsub; ++sub) {
        const SlotDev sl = c.slots[sub];
        wait_costs(T);  // terminal stage: nothing to scan, but every role follows every cost phase
        if (self_finish) {
            fin.terminal(sl, T);
            finished();
        }
        ++T;
        for (int i = n - 1; i >= 1; --i, ++T) {
            const double *Pc = sm.Ps + (size_t)(i % kPsBufs) * R * Kp;
            const double *ssc = sm.ss + (size_t)(T % 3) * Kp;
            wait_costs(T);
            // ---- phase B: register-tiled min-plus scan over this group's successors, sub-slice A then B ------
            wait_finished(0);  // rows of A for stage i are complete (finish(A, i+1) or the terminal stage)
            PROF_LAP(0);
            if constexpr (PR > 0) {
                // ---- pruned scan: block minima of my rows, branch-and-bound scan, scatter from registers ----------
                static_assert(TL == 1, "pruned scan: a lane is a level");
                if (pr_rows_live <= 0) {
                    // a row group without rows (two-zone slices: this CTA owns one group less; or rows beyond the table):
                    // nothing to scan or scatter, but the warp keeps every hand-over of the stage
                    PROF_LAP(1);
                    if constexpr (kPsBufs != 3) scanned(0);
                    fin.wait_inputs(i, T);
                    PROF_LAP(2);
                    finished();
                    PROF_LAP(3);
                    pc[4] += 1;
                    continue;
                }
                const int nblk = c.Kr / PR;
                float *pmR = sm.pminf;
                int *qseed = reinterpret_cast<int *>(sm.pminf + 8 * nblk);
                // block minima and seed of the rows of my row group, once per stage: the group's warps (one per level block)
                // share its rows, then meet at the group's own named barrier (not a CTA-wide one)
                for (int r = rowA + (tid >> 5) % nLB; r < min(rowA + TBA, fin.myR); r += nLB)
                    row_minima<PR>(Pc + (size_t)r * Kp, pmR, r, qseed + r, nblk, lane);
                // the no-jump candidates need neither minima nor seeds: their loads and adds run while the group gathers
                const double s_l = ssc[lg];
                double vself[TBA];
                pruned_self_candidates<TBA>(Pc + (size_t)rowA * Kp, sm.cs + lg, s_l, Kp, min(lg, K - 1), vself);
                asm volatile("bar.sync %0, %1;" ::"r"(1 + pr_grp), "r"(32 * nLB) : "memory");
                const FinishArgs fa = fin.stage_args(sl, i, T);
                double best[TBA][1];
                int arg[TBA][1];
                scan_pruned<TBA, PR, ArgT, PROF>(Pc + (size_t)rowA * Kp, sm.cs + lg, sm.cminf + lg, pmR, rowA, qseed + rowA,
                                                 s_l, sm.pminf[8 * nblk + 16 + lg], sm.pminf[8 * nblk + 16 + Kp + lg], nblk, Kp, active, pr_rows_live, lane,
                                                 vself, best, arg, executed, ph);
                PROF_LAP(1);
                if constexpr (kPsBufs != 3) scanned(0);  // (three buffers: the comm warp follows the "finished" count instead)
                fin.wait_inputs(i, T);
                PROF_LAP(2);
                scatter_tile<TBA, 1, ArgT>(fa, rowA, lg, best, arg);  // pad levels only write their MARK bytes
                finished();
                PROF_LAP(3);
                pc[4] += 1;
                continue;
            } else if (direct) {
                // ---- one j-group, no scatter warps: scan, then scatter the finished tile from registers ------------
                double best[TBA][TL];
                int arg[TBA][TL];
                if (active) scan_tile<TBA, TL, ArgT>(Pc + (size_t)rowA * Kp, sm.cs + lg * TL, ssc + lg * TL, 0, c.Kr, Kp, best, arg);
                PROF_LAP(1);
                scanned(0);  // the comm warp may refill the rows this stage read
                fin.wait_inputs(i, T);
                PROF_LAP(2);
                if (active) scatter_tile<TBA, TL, ArgT>(fin.stage_args(sl, i, T), rowA, lg * TL, best, arg);
                finished();
                PROF_LAP(3);
                pc[4] += 1;
                continue;
            }
            if (active)
                phase_b<TBA, TL, ArgT>(Pc + (size_t)rowA * Kp, sm.cs + lg * TL, ssc + lg * TL,
                                       sm.pv + ((size_t)jg * R + rowA) * Kp + lg * TL,
                                       pa + ((size_t)jg * R + rowA) * Kp + lg * TL, jb, je, Kp);
            PROF_LAP(1);
            scanned(0);
            PROF_LAP(2);
            if constexpr (TBB > 0) {
                wait_finished(1);
                PROF_LAP(3);
                if (active)
                    phase_b<TBB, TL, ArgT>(Pc + (size_t)rowB * Kp, sm.cs + lg * TL, ssc + lg * TL,
                                           sm.pv + ((size_t)jg * R + rowB) * Kp + lg * TL,
                                           pa + ((size_t)jg * R + rowB) * Kp + lg * TL, jb, je, Kp);
                PROF_LAP(1);
                scanned(1);
                PROF_LAP(2);
            } else if (self_finish) {
                // ---- phase C by the compute warps: all partial minima are in shared memory once every warp scanned
                mbar_wait_wd(&sm.mbar[MB_SCANNED], scanned_phase, c.err, c.wd_cycles);
                scanned_phase ^= 1u;
                fin.wait_inputs(i, T);
                PROF_LAP(2);
                fin.rows(fin.stage_args(sl, i, T), 0, R * fin.ublocks);
                finished();
                PROF_LAP(3);
            }
            pc[4] += 1;
        }
        // drain: stage 1 (or the terminal stage when n == 1) is finished; keeps the barrier phases aligned
        for (int v = 0; v < NV; ++v) wait_finished(v);
    }
    if constexpr (PR > 0) {  // candidates really evaluated: blocks x successors x rows x live lanes of this warp
        if (c.exec && lane == 0) {
            const int live = min(32, max(0, K - (((tid >> 5) % nLB) << 5)));
            atomicAdd(c.exec, (unsigned long long)executed * (unsigned long long)(PR * max(0, min(TBA, fin.myR - rowA)) * live));
        }
    }
    if (PROF && c.prof && tid == 0 && self_finish) {
        c.prof[(size_t)g * 16 + 13] = fin.ring_wait;  // the part of the hand-over spent waiting for ring space
        c.prof[(size_t)g * 16 + 14] = fin.pcc[0];
        c.prof[(size_t)g * 16 + 15] = fin.pcc[1];
    }
    if constexpr (PROF && PR > 0) {
        if (c.prof && tid == 0) {  // the scatter warps' slots are free with pruned tiles: sub-phases of the scan (warp 0)
            c.prof[(size_t)g * 16 + 5] = ph[0];
            c.prof[(size_t)g * 16 + 6] = ph[1];
            c.prof[(size_t)g * 16 + 7] = ph[2];
            c.prof[(size_t)g * 16 + 14] = ph[3];
            c.prof[(size_t)g * 16 + 15] = executed;
        }
    }
    if (PROF && c.prof && tid == 0) {
        c.prof[(size_t)g * 16 + 0] = pc[0];
        c.prof[(size_t)g * 16 + 1] = pc[1];
        c.prof[(size_t)g * 16 + 2] = pc[2];
        c.prof[(size_t)g * 16 + 3] = pc[3];
        c.prof[(size_t)g * 16 + 4] = pc[4];
    }
#undef PROF_LAP
}

template <int TBA, int TBB, int TL, typename ArgT, int MAXT, bool PROF, int PR, int KPC = 0, int KRC = KPC>
__global__ void __launch_bounds__(MAXT, 1) wavefront_kernel(const __grid_constant__ Tables t_in, const __grid_constant__ WaveCfg c_in)
{
    if constexpr (KPC > 0) {
        Tables t = t_in;
        WaveCfg c = c_in;
        t.Kp = KPC;
        c.Kr = KRC;   // rows of the (padded) jump-cost table: Kp for pruned tiles, the j-groups' whole trips for exhaustive ones
        if constexpr (PR > 0) { c.R = TBB; c.RA = TBB; c.JS = 1; c.NS = 0; c.RB = 0; }  // what fill_geometry sets for pruned tiles anyway
        if constexpr (PR == 0 && TBA == 4 && TBB == 3 && TL == 2 && KPC == 128) {
            // the config-4 geometry of the exhaustive tile (launch_variant checks that the plan's geometry is this one)
            c.JS = 4; c.jper = 32; c.NS = 6; c.NF = 6; c.RG = 1; c.R = 7; c.RA = 4; c.RB = 3; c.tpg = 64; c.nLG = 63; c.EC = 2;
        }
        wavefront_body<TBA, TBB, TL, ArgT, PROF, PR>(t, c);
    } else {
        wavefront_body<TBA, TBB, TL, ArgT, PROF, PR>(t_in, c_in);
    }
}

// ---- host side -----------------------------------------------------------------------------------
struct Variant { int TBA, TBB, TL, PR; };
// (rows of sub-slice A, rows of sub-slice B, levels) per thread tile.  TBB = 0: one sub-slice; phase C then runs
// after the scan, either on the compute warps themselves (NS = 0) or on scatter warps.  Every variant is built for
// CTAs of up to 512 threads (128 registers per thread; the largest tiles spill a little and are penalised by the model).
// The last column is PR, the block size of the pruned scan (0 = exhaustive scan).  Pruned tiles: lane = level, a warp is 32
// levels x TBA consecutive rows, and the second column is the number of rows of the CTA (the last row group may be ragged:
// 7 rows = 4 + 3 or 2 + 2 + 2 + 1); the row groups run side by side on different warps and the compute warps finish their
// own stage.  Built for 4 * ceil(rows / TBA) compute warps + comm + publisher.
#ifdef BB_FAST_BUILD   // register / SASS checks of the production tile only (nvcc -cubin -DBB_FAST_BUILD): not a usable library
#define BB200_VARIANTS(X) X(27, 2, 8, 1, 4)
#else
#define BB200_VARIANTS(X)                                                                                      \
    X(0, 7, 0, 2, 0) X(1, 8, 0, 2, 0) X(2, 6, 0, 2, 0) X(3, 5, 0, 2, 0) X(4, 4, 0, 2, 0) X(5, 3, 0, 2, 0) X(6, 2, 0, 2, 0) X(7, 1, 0, 2, 0) \
    X(8, 8, 0, 1, 0) X(9, 4, 0, 1, 0) X(10, 2, 0, 1, 0) X(11, 1, 0, 1, 0)                                                    \
    X(12, 4, 3, 2, 0) X(13, 4, 4, 2, 0) X(14, 3, 3, 2, 0) X(15, 3, 2, 2, 0) X(16, 2, 2, 2, 0) X(17, 2, 1, 2, 0) X(18, 1, 1, 2, 0)     \
    X(19, 4, 4, 1, 0) X(20, 2, 2, 1, 0) X(21, 1, 1, 1, 0) X(22, 4, 3, 1, 0) X(23, 3, 3, 1, 0)                               \
    X(24, 4, 7, 1, 4) X(25, 2, 7, 1, 4) X(26, 4, 8, 1, 4) X(27, 2, 8, 1, 4) X(28, 2, 4, 1, 4) X(29, 1, 2, 1, 4) X(30, 4, 4, 1, 4) X(31, 3, 6, 1, 4)
#endif
static const Variant kVariants[] = {
#define X(idx, a, b, l, pr) {a, b, l, pr},
    BB200_VARIANTS(X)
#undef X
};
static const int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);
constexpr int kWaveThreads = kWaveThreadsSmall;
// pruned tiles: 4 level blocks x ceil(rows / TBA) row groups of compute warps + comm + publisher, rounded to 128 threads
constexpr int pruned_max_threads(int tba, int rows) { return (32 * (4 * ((rows + tba - 1) / tba) + 2) + 127) / 128 * 128; }

static void fill_geometry(const Tables &t, int argw, int G, int JS, int v, int NS, WaveCfg &c)
{
    c.variant = v;
    c.TB = kVariants[v].TBA;
    c.TBB = kVariants[v].TBB;
    c.TL = kVariants[v].TL;
    c.PR = kVariants[v].PR;
    const int tb = c.TB + c.TBB;
    const int rows_per_cta = (t.B1 + G - 1) / G;
    if (c.PR > 0) {
        // pruned scan: the CTA owns TBB rows in ceil(TBB / TB) row groups of TB rows, one lane per level, one warp per 32
        // levels and row group; the caller rejects the variant when the slices do not cover the table (G > SMs)
        const int ng = (c.TBB + c.TB - 1) / c.TB;
        c.RG = ng;
        c.R = c.TBB;
        c.RA = c.TBB;
        c.RB = 0;
        c.G = (t.B1 + c.R - 1) / c.R;
        c.GA = c.G;
        c.Rtop = c.R;
        if (kTwoZoneSlices && ng >= 2 && c.TBB % c.TB == 0 && G > 0) {
            // whole row groups only: the upper zone runs one group less.  With G CTAs, GA of them must own R rows so that
            // GA R + (G - GA) Rtop >= B1; when Rtop rows per CTA already cover the table, every slice gets Rtop rows.
            const int rtop = c.R - c.TB;
            const int ga = (t.B1 - rtop * G + (c.R - rtop) - 1) / (c.R - rtop);
            if (ga <= 0) {
                c.Rtop = rtop; c.GA = 0; c.G = (t.B1 + rtop - 1) / rtop;
            } else if (ga < G) {
                c.Rtop = rtop; c.GA = ga; c.G = G;
            }  // ga >= G: uniform slices of R rows (the caller rejects the variant if they need more than G CTAs)
        }
        c.nLG = t.K;
        c.JS = 1;
        c.jper = c.Kr = (t.K + 8 * c.PR - 1) / (8 * c.PR) * (8 * c.PR);  // whole groups of eight blocks; rows K .. Kr-1 of the cost table are +Inf
        c.tpg = t.Kp * ng;                               // 32 lanes per level block and row group
        c.NS = 0;
        c.EC = 2;
        c.NF = c.tpg / 32;
        c.threads = c.tpg + 64;
        c.smem = carve(t, c, argw, nullptr, nullptr);
        return;
    }
    c.RG = (rows_per_cta + tb - 1) / tb;
    c.R = c.RG * tb;
    c.RA = c.RG * c.TB;
    c.RB = c.RG * c.TBB;
    c.G = (t.B1 + c.R - 1) / c.R;  // drop CTAs that would own no row
    c.GA = c.G;
    c.Rtop = c.R;
    c.nLG = (t.K + c.TL - 1) / c.TL;
    c.JS = JS;
    c.jper = ((t.K + JS - 1) / JS + 1) & ~1;  // even: successors are taken in aligned pairs
    // Pad the successor axis with +Inf cost rows up to whole trips of the unrolled scan loop (8 successors) when the
    // value rows are wide enough (Kp columns exist) -- every j-group then runs the same straight-line loop and the
    // remainder / odd-tail code stays cold (instruction cache).  Otherwise exactly K rows.
    {
        const int jper8 = (c.jper + 7) & ~7;
        const int need = JS * jper8;
        if (need <= t.Kp) { c.jper = jper8; c.Kr = need; }
        else c.Kr = t.K;
    }
    c.tpg = ((c.RG * c.nLG + 31) / 32) * 32;
    c.NS = NS;
    c.EC = 2;  // cells per lane in phase C (the kernel dispatches EC = 2 only, see Finisher::rows)
    c.NF = c.NS > 0 ? c.NS : c.JS * c.tpg / 32;  // warps that finish a stage: scatter warps, else the compute warps
    c.threads = c.JS * c.tpg + 64 + 32 * c.NS;  // + comm warp + publisher warp + scatter warps
    c.smem = carve(t, c, argw, nullptr, nullptr);
}

bool wave_configure(const Tables &t, int argw, int num_sms, size_t smem_max, int want_ctas, int want_js,
                    int want_variant, WaveCfg &cfg)
{
    if (t.M > kMaxM || argw != 1) return false;  // K > 255: the jump-cost table would not fit in shared memory anyway
    if ((long long)t.n >= ((long long)1 << 30)) return false;  // ticks are 32-bit; the host splits long batches into launches
    const bool no_prune = want_variant < 0;  // -1: automatic choice among the exhaustive tiles only (the plan saw that pruning does not pay)
    if (no_prune) want_variant = 0;
    const int want_ns = want_variant / 100;  // 0: automatic, 1..8: that many scatter warps, 10: none (NS = 0)
    want_variant %= 100;
    static const int kNsChoices[] = {0, 1, 2, 3, 4, 6, 8};
    double best_stage = -1.;
    bool found = false;
    for (int v = 0; v < kNumVariants; ++v) {
        if (want_variant > 0 && v != want_variant - 1) continue;
        if (((t.K + kVariants[v].TL - 1) / kVariants[v].TL) * kVariants[v].TL > t.Kp) continue;
        const int pr = kVariants[v].PR;
        if (pr > 0) {
            // the pruned scan pays when a stage has many successors to skip; narrow level sets keep the exhaustive tiles
            if (want_variant == 0 && (!kAutoPruned || no_prune || t.K < 64)) continue;
            if ((t.K + 8 * pr - 1) / (8 * pr) * (8 * pr) > t.Kp || (t.K + 8 * pr - 1) / (8 * pr) * 8 > 32) continue;  // whole groups of eight blocks, one mask word
        }
        for (int js = 1; js <= 16; ++js) {
            if (want_js > 0 && js != want_js) continue;
            if (js > t.K) break;
            if (pr > 0 && js > 1) break;  // one thread scans all successors of its cells
            for (int nsi = 0; nsi < 7; ++nsi) {
                const int ns = kNsChoices[nsi];
                if (kVariants[v].TBB > 0 && ns == 0 && pr == 0) continue;  // two sub-slices one after the other need scatter warps
                if (pr > 0 && ns != 0) continue;                            // pruned scan: the compute warps finish the stage
                if (want_ns > 0 && ns != want_ns % 10) continue;  // 100 * NS forces NS scatter warps, 1000: none
                int gmax = want_ctas > 0 ? want_ctas : num_sms;
                if (gmax > num_sms) gmax = num_sms;
                WaveCfg c = cfg;
                fill_geometry(t, argw, gmax, js, v, ns, c);
                if (pr > 0 && (c.G > gmax || c.RG > kMaxRowGroups || c.threads > pruned_max_threads(kVariants[v].TBA, kVariants[v].TBB))) continue;  // rows x CTAs must cover the table
                if (pr == 0 && c.threads > kWaveThreads) continue;
                if (c.smem > smem_max) continue;
                if ((size_t)c.JS * c.R * t.Kp >= ((size_t)1 << 30) || (size_t)t.B1 * t.Kp >= ((size_t)1 << 31) ||
                    c.R >= (1 << 15) || t.Kp >= (1 << 16))
                    continue;  // 32-bit cell indices in phase C
                // Cost model per stage and CTA in scheduler cycles, fitted to the in-kernel profile on B200
                // (profiles/phase_profile_r01.txt).  Phase B is issue bound: an FP64 instruction takes two issue
                // slots, so a candidate costs DADD(2)+DSETP(2)+2 FSEL+SEL = 7 slots, plus the per-successor loads and
                // the s_l + c_jl adds; the busiest scheduler hosts ceil(warps/4) warps; ~20% of the slots are lost to
                // dependency stalls.  Phase C is latency bound; it only shows when it outlasts the scan of the other
                // sub-slice that hides it.
                const int cwarps = c.JS * c.tpg / 32;
                const int sched = (cwarps + 3) / 4;
                // one warp per scheduler cannot hide the compare -> select latency; three or more run with 128
                // registers and a longer schedule (measured: 1.65 / 1.3 / 1.4)
                const double stall = sched == 1 ? 1.65 : sched == 2 ? 1.3 : 1.4;
                auto scan = [&](int tb) {
                    if (tb == 0) return 0.0;
                    const double per_j = 7.0 * tb * c.TL + 2.0 * c.TL + ((tb + 1) / 2) + ((c.TL + 1) / 2) + 3.0;
                    return per_j * c.jper * sched * stall + 150.0;
                };
                // phase C is latency bound: time per unit (64 cells) and finisher warp; warps that share the SM with the
                // scan only get the issue slots it leaves (constants fitted to tools/tune_sweep.py runs of the
                // config-4 and the heat-shaped instance, profiles/tune_sweep*_r01.txt)
                auto finish = [&](int rows, bool hidden) {
                    const int units = rows * ((t.Kp + 32 * c.EC - 1) / (32 * c.EC));
                    const int per_warp = (units + c.NF - 1) / c.NF;
                    const double generic = (js == 2 || js == 4) ? 1.0 : 1.5;  // other splits: rolled combine loop
                    return per_warp * generic * (hidden ? 1200.0 + 50.0 * js : 100.0 + 310.0 * js) + 300.0;
                };
                const double sa = scan(c.TB), sb = scan(c.TBB);
                double stage;
                if (pr > 0) {
                    // per scheduler: the warps it hosts scan (TBA + TBB) rows x Kr successors between them; about half
                    // of the blocks are skipped on the BASELINE shapes (the skip rate is data dependent: it is measured,
                    // bb200_stats[17], not modelled); every block costs a bound check
                    // fitted to profiles/phase_profile_r02_pruned.txt (config 4: 7 400 cycles per stage with two rows per warp)
                    const int nlb = t.Kp / 32, per_sched = (nlb + 3) / 4;
                    const double full = 7.0 * (c.GA < c.G ? c.Rtop : c.R) * c.Kr * per_sched;  // the upper slices set the pace
                    stage = 0.35 * full + 5000.0;
                    if (c.TB > 2) stage *= 1.25;  // four rows per warp: half the warps, the latency of the bound phases shows (measured +22 %)
                } else if (c.TBB > 0) {
                    const double fa = finish(c.RA, true), fb = finish(c.RB, true);
                    // a shorter finish also shortens the lag a successor slice needs behind this one
                    stage = sa + sb + (fa > sb ? fa - sb : 0.) + (fb > sa ? fb - sa : 0.) + 0.15 * (fa + fb) + 800.0;
                } else {
                    if (js == 1 && ns == 0)  // direct: the finished tile is scattered from registers, one hand-over
                        stage = sa + 14.0 * c.TB * c.TL * sched + 450.0;
                    else
                        stage = sa + finish(c.RA, false) + 600.0;  // + the hand-over after the scan
                }
                stage += 10.0 * c.NS;  // scatter warps take issue slots from the scan
                // tiles ptxas cannot keep in 128 registers (-Xptxas -v: 8x2 596 B, 6x2 104 B, 8x1 40 B, 7x2 24 B of spills)
                // pay for their local-memory traffic inside the scan loop
                if (pr == 0) {
                    const int cells = (c.TB > c.TBB ? c.TB : c.TBB) * c.TL;
                    if (cells >= 16) stage *= 1.5;
                    else if (cells >= 12 || (c.TB == 8 && c.TL == 1)) stage *= 1.15;
                }
                if (!found || stage < best_stage) {
                    best_stage = stage;
                    cfg = c;
                    found = true;
                }
            }
        }
    }
    return found;
}

template <int TBA, int TBB, int TL, int PR>
static cudaError_t launch_variant(const Tables &t, const WaveCfg &cfg, cudaStream_t st)
{
    void *args[] = {(void *)&t, (void *)&cfg};
    constexpr int MAXT = PR > 0 ? pruned_max_threads(TBA, TBB) : kWaveThreads;
    // the cycle counters are a separate instantiation: their code would cost the production kernel ~1.5 %
    const void *fn = cfg.prof ? (const void *)wavefront_kernel<TBA, TBB, TL, uint8_t, MAXT, true, PR>
                              : (const void *)wavefront_kernel<TBA, TBB, TL, uint8_t, MAXT, false, PR>;
    // the tiles of the production shape (pruned, and the exhaustive tile it falls back to): Kp as a compile-time constant
    if constexpr ((PR > 0 && TBA == 2 && TBB >= 7) || (PR == 0 && TBA == 4 && TBB == 3 && TL == 2)) {
        const bool geom_ok = PR > 0 || (cfg.JS == 4 && cfg.jper == 32 && cfg.NS == 6 && cfg.NF == 6 && cfg.RG == 1 && cfg.R == 7 &&
                                        cfg.RA == 4 && cfg.RB == 3 && cfg.tpg == 64 && cfg.nLG == 63 && cfg.EC == 2);
        if (BB_KPC != 0 && t.Kp == 128 && cfg.Kr == 128 && geom_ok)
            fn = cfg.prof ? (const void *)wavefront_kernel<TBA, TBB, TL, uint8_t, MAXT, true, PR, 128>
                          : (const void *)wavefront_kernel<TBA, TBB, TL, uint8_t, MAXT, false, PR, 128>;
    }
    if constexpr (PR == 0 && (TBA == 1 || TBA == 2 || TBA == 4) && TBB == 0 && TL == 1) {  // the reference's heat example: 6 x 6 levels, one j-group (direct tiles)
        if (BB_KPC != 0 && !cfg.prof && t.Kp == 64 && cfg.Kr == 40 && cfg.JS == 1 && cfg.NS == 0)
            fn = (const void *)wavefront_kernel<TBA, TBB, TL, uint8_t, MAXT, false, PR, 64, 40>;
    }
    if constexpr (PR > 0 && TBA == 2 && TBB == 8) {  // the other level counts the pruned tiles take: 64 < K <= 96, K = 64
        if (BB_KPC != 0 && !cfg.prof && t.Kp == 96 && cfg.Kr == 96) fn = (const void *)wavefront_kernel<TBA, TBB, TL, uint8_t, MAXT, false, PR, 96>;
        if (BB_KPC != 0 && !cfg.prof && t.Kp == 64 && cfg.Kr == 64) fn = (const void *)wavefront_kernel<TBA, TBB, TL, uint8_t, MAXT, false, PR, 64>;
    }
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem);
    if (e != cudaSuccess) return e;
    // cooperative launch: the CTAs wait on one another, so all of them must be co-resident
    return cudaLaunchCooperativeKernel(fn, dim3(cfg.G), dim3(cfg.threads), args, cfg.smem, st);
}

cudaError_t launch_wavefront(const Tables &t, const WaveCfg &cfg, int argw, cudaStream_t st)
{
    if (argw != 1) return cudaErrorInvalidValue;
    switch (cfg.variant) {
#define X(idx, a, b, l, pr) case idx: return launch_variant<a, b, l, pr>(t, cfg, st);
        BB200_VARIANTS(X)
#undef X
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace bb200
