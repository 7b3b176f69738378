// kernel_wavefront.cu -- persistent, pipelined min-plus DP over all stages (the hot path).
//
// One launch walks every stage i = n-1 .. 1 of bellman_TRM! (HelpFunctions.jl:45-82) for one or more
// subproblems.  Work decomposition:
//
//   * The budget axis is cut into G slices of R consecutive SOURCE rows b' (one persistent CTA each,
//     one CTA per SM).  A target cell (b, l) of stage i has exactly one source row b' = b - b~_l(i)
//     (HelpFunctions.jl:69-71), so slicing by source row partitions the cells, every CTA reads only the
//     value rows it owns (resident in shared memory) and PUSHES results whose target row b' + b~_l
//     belongs to a higher slice.  Budget only flows upwards, therefore slice g depends on slices
//     g-1 .. g-D only (D = ceil(max b~ / R)): the slices form a pipeline and low-budget CTAs run ahead in
//     time.  Neighbours synchronise through per-CTA progress counters in global memory (release/acquire),
//     never through a grid-wide barrier.
//   * Inside a CTA the stage is a small min-plus matrix product C[b', l] = min_j (s_l + c_jl) + P[b', j].
//     A thread owns a TB x TL register tile of cells (TB source rows, TL levels), thread groups split the
//     successor range j (JS groups); every candidate is two separately rounded FP64 adds and a strict '>'
//     (earliest successor wins ties, +Inf/NaN never win) -- exactly the reference's arithmetic.  Partial
//     (min, argmin) pairs of the JS groups are combined in ascending-j order through shared memory.
//   * df[:, i] and u_old[:, i] arrive in kChunk-stage chunks by 1-D bulk TMA (cp.async.bulk + mbarrier).
//   * The argmin goes to HBM once per cell as uint8/uint16, indexed by source row (coalesced rows).
#include "bb200_internal.cuh"
#include "kernels.cuh"

namespace bb200 {

// ---- PTX helpers ---------------------------------------------------------------------------------
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
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
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

// Spin until *flag >= want.  A bounded watchdog turns a lost dependency into an error code instead of a
// hung GPU: after ~2^24 polls the CTA raises the abort flag and every poller gives up.
__device__ __forceinline__ void wait_flag(const unsigned long long *flag, long long want, int *err)
{
    if (want <= 0) return;
    unsigned int spins = 0;
    while ((long long)ld_acquire(flag) < want) {
        if ((++spins & 0x3ffu) == 0) {
            if (*(volatile int *)&err[3]) return;
            if (spins > (1u << 24)) {
                atomicOr(&err[2], 1);
                atomicOr(&err[3], 1);
                return;
            }
        }
    }
}

struct Smem {
    uint64_t *mbar;   // [2]
    double *dfb;      // [2][kChunk*M]
    double *uob;      // [2][kChunk*M]
    double *lvs;      // [K*M]
    double *ss;       // [Kp]
    int *bts;         // [2][Kp]
    double *Ps;       // [2][Kp*RP]
    double *cs;       // [K*Kp]
    double *pv;       // [JS*R*Kp]
    unsigned char *pa;  // ArgT[JS*R*Kp]
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Carves the dynamic shared memory; argw = bytes per partial argmin entry (1 or 2).
__host__ __device__ inline size_t carve(const Tables &t, const WaveCfg &c, int argw, unsigned char *base,
                                        Smem *s)
{
    size_t off = 0;
    size_t o[10];
    const size_t chunk = (size_t)kChunk * t.M * sizeof(double);
    const size_t sizes[10] = {2 * sizeof(uint64_t),
                              2 * chunk,
                              2 * chunk,
                              (size_t)t.K * t.M * sizeof(double),
                              (size_t)t.Kp * sizeof(double),
                              2 * (size_t)t.Kp * sizeof(int),
                              2 * (size_t)t.Kp * c.RP * sizeof(double),
                              (size_t)t.K * t.Kp * sizeof(double),
                              (size_t)c.JS * c.R * t.Kp * sizeof(double),
                              (size_t)c.JS * c.R * t.Kp * (size_t)argw};
    for (int k = 0; k < 10; ++k) {
        o[k] = off;
        off = align_up(off + sizes[k], 128);
    }
    if (s) {
        s->mbar = reinterpret_cast<uint64_t *>(base + o[0]);
        s->dfb = reinterpret_cast<double *>(base + o[1]);
        s->uob = reinterpret_cast<double *>(base + o[2]);
        s->lvs = reinterpret_cast<double *>(base + o[3]);
        s->ss = reinterpret_cast<double *>(base + o[4]);
        s->bts = reinterpret_cast<int *>(base + o[5]);
        s->Ps = reinterpret_cast<double *>(base + o[6]);
        s->cs = reinterpret_cast<double *>(base + o[7]);
        s->pv = reinterpret_cast<double *>(base + o[8]);
        s->pa = base + o[9];
    }
    return off;
}

template <int TB, int TL, typename ArgT>
__global__ void __launch_bounds__(kMaxWaveThreads, 1) wavefront_kernel(Tables t, WaveCfg c)
{
    constexpr int TBP = (TB + 1) & ~1;  // row positions per row group (even: 16-byte aligned loads)
    constexpr ArgT MARK = (ArgT)~(ArgT)0;
    constexpr int MARKI = (int)MARK;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem sm;
    carve(t, c, (int)sizeof(ArgT), smem_raw, &sm);

    const int tid = threadIdx.x, NT = blockDim.x;
    const int g = blockIdx.x;
    const int r0 = g * c.R;
    const int K = t.K, Kp = t.Kp, M = t.M, B1 = t.B1, R = c.R, RP = c.RP;
    const int jg = tid / c.tpg, tig = tid % c.tpg;
    const bool active = tig < c.RG * c.nLG;
    const int rg = active ? tig / c.nLG : 0;
    const int lg = active ? tig % c.nLG : 0;
    const int jb = jg * c.jper;
    const int je = min(K, jb + c.jper);
    const double inf = d_inf();
    const uint32_t chunk_bytes = (uint32_t)(kChunk * M * sizeof(double));

    // one-time: constant tables into shared memory
    for (int x = tid; x < K * Kp; x += NT) sm.cs[x] = t.cost[x];
    for (int x = tid; x < K * M; x += NT) sm.lvs[x] = t.lvd[x];
    if (tid == 0) {
        mbar_init(&sm.mbar[0], 1);
        mbar_init(&sm.mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int btm = min(*c.btmax, B1 - 1);
    const int D = (btm + R - 1) / R;  // predecessor / successor slices a push can span
    unsigned long long *myflag = c.flags + (size_t)g * kFlagStride;
    uint32_t phase_bits = 0;  // mbarrier parity per chunk buffer
    long long tick = 0;       // stages completed by this CTA in this launch (monotone across subproblems)

    auto rowpos = [&](int row) { return (row / TB) * TBP + (row % TB); };

    for (int sub = 0; sub < c.nsub; ++sub) {
        const SlotDev sl = c.slots[sub];
        ArgT *argtab = reinterpret_cast<ArgT *>(sl.arg);
        const int n = t.n;
        int chunk_cur = (n - 1) / kChunk;  // chunk holding the terminal stage row n-1
        __syncthreads();                   // previous subproblem fully done with the chunk buffers
        if (tid == 0) {
            const int b0 = chunk_cur & 1;
            mbar_expect_tx(&sm.mbar[b0], 2 * chunk_bytes);
            tma_load_1d(sm.dfb + (size_t)b0 * kChunk * M, sl.df + (size_t)chunk_cur * kChunk * M, chunk_bytes,
                        &sm.mbar[b0]);
            tma_load_1d(sm.uob + (size_t)b0 * kChunk * M, sl.u_old + (size_t)chunk_cur * kChunk * M,
                        chunk_bytes, &sm.mbar[b0]);
            if (chunk_cur >= 1) {
                const int b1 = (chunk_cur - 1) & 1;
                mbar_expect_tx(&sm.mbar[b1], 2 * chunk_bytes);
                tma_load_1d(sm.dfb + (size_t)b1 * kChunk * M, sl.df + (size_t)(chunk_cur - 1) * kChunk * M,
                            chunk_bytes, &sm.mbar[b1]);
                tma_load_1d(sm.uob + (size_t)b1 * kChunk * M, sl.u_old + (size_t)(chunk_cur - 1) * kChunk * M,
                            chunk_bytes, &sm.mbar[b1]);
            }
        }
        mbar_wait(&sm.mbar[chunk_cur & 1], (phase_bits >> (chunk_cur & 1)) & 1u);
        phase_bits ^= 1u << (chunk_cur & 1);

        // ---- terminal stage n (HelpFunctions.jl:27-43): P[b][l] = (b == b~_l(n)) ? s_l(n) : Inf ----
        int cur = 0;
        {
            const int ri = n - 1;
            const double *dfr = sm.dfb + ((size_t)(chunk_cur & 1) * kChunk + (ri % kChunk)) * M;
            const double *uor = sm.uob + ((size_t)(chunk_cur & 1) * kChunk + (ri % kChunk)) * M;
            for (int l = tid; l < Kp; l += NT) {
                double s = 0.;
                int bt = B1;
                if (l < K) stage_cost(t, sm.lvs + l * M, dfr, uor, s, bt);
                sm.ss[l] = s;
                sm.bts[(tick & 1) * Kp + l] = bt;
            }
            __syncthreads();
            double *Pc = sm.Ps + (size_t)cur * Kp * RP;
            for (int x = tid; x < Kp * RP; x += NT) Pc[x] = inf;
            __syncthreads();
            for (int x = tid; x < R * Kp; x += NT) {
                const int row = x / Kp, l = x % Kp;
                const int b = r0 + row;
                if (l < K && b < B1) {
                    const bool hit = (b == sm.bts[(tick & 1) * Kp + l]);
                    const double v = hit ? sm.ss[l] : inf;
                    if (hit) Pc[l * RP + rowpos(row)] = v;
                    if (n <= 2) sl.phi[((size_t)((n + 1) & 1) * B1 + b) * Kp + l] = v;
                }
            }
            // (tick numbering: the terminal stage of subproblem `sub` is tick sub*n)
            __syncthreads();
            if (tid == 0) st_release(myflag, (unsigned long long)tick);
        }

        // ---- stages i = n-1 .. 1 --------------------------------------------------------------
        for (int i = n - 1; i >= 1; --i) {
            ++tick;
            const int ri = i - 1;
            const int ch = ri / kChunk;
            if (ch != chunk_cur) {
                // first stage of a new chunk: everybody finished reading chunk ch+1 one barrier ago
                chunk_cur = ch;
                if (tid == 0 && ch >= 1) {
                    const int b1 = (ch - 1) & 1;
                    mbar_expect_tx(&sm.mbar[b1], 2 * chunk_bytes);
                    tma_load_1d(sm.dfb + (size_t)b1 * kChunk * M, sl.df + (size_t)(ch - 1) * kChunk * M,
                                chunk_bytes, &sm.mbar[b1]);
                    tma_load_1d(sm.uob + (size_t)b1 * kChunk * M, sl.u_old + (size_t)(ch - 1) * kChunk * M,
                                chunk_bytes, &sm.mbar[b1]);
                }
                mbar_wait(&sm.mbar[ch & 1], (phase_bits >> (ch & 1)) & 1u);
                phase_bits ^= 1u << (ch & 1);
            }
            const int nxt = cur ^ 1;
            double *Pc = sm.Ps + (size_t)cur * Kp * RP;
            double *Pn = sm.Ps + (size_t)nxt * Kp * RP;
            int *bt_cur = sm.bts + (tick & 1) * Kp;
            const int *bt_prev = sm.bts + ((tick - 1) & 1) * Kp;
            const bool need_halo = (i < n - 1) && (g > 0);

            // ---- phase A: stage costs, neighbour waits, halo gather, +Inf fill of the next rows -----
            {
                const double *dfr = sm.dfb + ((size_t)(ch & 1) * kChunk + (ri % kChunk)) * M;
                const double *uor = sm.uob + ((size_t)(ch & 1) * kChunk + (ri % kChunk)) * M;
                for (int l = tid; l < Kp; l += NT) {
                    double s = 0.;
                    int bt = B1;
                    if (l < K) stage_cost(t, sm.lvs + l * M, dfr, uor, s, bt);
                    sm.ss[l] = s;
                    bt_cur[l] = bt;
                }
                if (tid == NT - 1) {
                    // data: predecessors finished the previous tick (their pushes are visible)
                    if (need_halo)
                        for (int d = 1; d <= D && d <= g; ++d)
                            wait_flag(c.flags + (size_t)(g - d) * kFlagStride, tick - 1, c.err);
                    // back-pressure: successors consumed the ring slot this tick overwrites
                    for (int d = 1; d <= D && g + d < c.G; ++d)
                        wait_flag(c.flags + (size_t)(g + d) * kFlagStride, tick - kHaloRing + 1, c.err);
                }
                for (int x = tid; x < Kp * RP; x += NT) Pn[x] = inf;
                __syncthreads();
                if (need_halo) {
                    const double *hsrc = c.halo + (size_t)((tick - 1) % kHaloRing) * B1 * Kp;
                    for (int x = tid; x < R * Kp; x += NT) {
                        const int row = x / Kp, l = x % Kp;
                        const int b = r0 + row;
                        if (l < K && b < B1) {
                            const int src = b - bt_prev[l];
                            if (src >= 0 && src < r0) Pc[l * RP + rowpos(row)] = __ldcg(hsrc + (size_t)b * Kp + l);
                        }
                    }
                    __syncthreads();
                }
            }

            // ---- phase B: register-tiled min-plus scan over this group's successors ---------------
            if (active) {
                double best[TB][TL];
                int arg[TB][TL];
#pragma unroll
                for (int a = 0; a < TB; ++a)
#pragma unroll
                    for (int q = 0; q < TL; ++q) { best[a][q] = inf; arg[a][q] = MARKI; }
                double s[TL];
#pragma unroll
                for (int q = 0; q < TL; ++q) s[q] = sm.ss[lg * TL + q];
                const double *Prow = Pc + rg * TBP;
                const double *crow = sm.cs + lg * TL;
#pragma unroll 2
                for (int j = jb; j < je; ++j) {
                    double p[TBP];
#pragma unroll
                    for (int k = 0; k < TBP / 2; ++k) {
                        const double2 x = *reinterpret_cast<const double2 *>(Prow + (size_t)j * RP + 2 * k);
                        p[2 * k] = x.x;
                        p[2 * k + 1] = x.y;
                    }
                    double a[TL];
                    if constexpr (TL % 2 == 0) {
#pragma unroll
                        for (int k = 0; k < TL / 2; ++k) {
                            const double2 x = *reinterpret_cast<const double2 *>(crow + (size_t)j * Kp + 2 * k);
                            a[2 * k] = __dadd_rn(s[2 * k], x.x);          // HelpFunctions.jl:67
                            a[2 * k + 1] = __dadd_rn(s[2 * k + 1], x.y);
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < TL; ++q) a[q] = __dadd_rn(s[q], crow[(size_t)j * Kp + q]);
                    }
#pragma unroll
                    for (int r = 0; r < TB; ++r)
#pragma unroll
                        for (int q = 0; q < TL; ++q) {
                            const double v = __dadd_rn(a[q], p[r]);       // :71
                            if (best[r][q] > v) { best[r][q] = v; arg[r][q] = j; }  // :73-76
                        }
                }
                // partial results of this j-group
                double *pv = sm.pv + ((size_t)jg * R + rg * TB) * Kp + lg * TL;
                ArgT *pa = reinterpret_cast<ArgT *>(sm.pa) + ((size_t)jg * R + rg * TB) * Kp + lg * TL;
#pragma unroll
                for (int r = 0; r < TB; ++r)
#pragma unroll
                    for (int q = 0; q < TL; ++q) {
                        pv[(size_t)r * Kp + q] = best[r][q];
                        pa[(size_t)r * Kp + q] = (ArgT)arg[r][q];
                    }
            }
            __syncthreads();

            // ---- phase C: combine the j-groups in ascending order, scatter values, store the argmin --
            {
                double *hdst = c.halo + (size_t)(tick % kHaloRing) * B1 * Kp;
                for (int x = tid; x < R * Kp; x += NT) {
                    const int row = x / Kp, l = x % Kp;
                    const int bsrc = r0 + row;
                    if (l >= K || bsrc >= B1) continue;
                    const int tgt = bsrc + bt_cur[l];
                    if (tgt >= B1) continue;  // outside `for b = 0:B-b~` (:69): the reference computes nothing
                    double val = inf;
                    ArgT a = MARK;
                    const ArgT *pa_all = reinterpret_cast<const ArgT *>(sm.pa);
                    for (int q = 0; q < c.JS; ++q) {
                        const double v = sm.pv[((size_t)q * R + row) * Kp + l];
                        if (val > v) { val = v; a = pa_all[((size_t)q * R + row) * Kp + l]; }
                    }
                    argtab[((size_t)(i - 1) * B1 + bsrc) * Kp + l] = a;
                    if (tgt < r0 + R) Pn[l * RP + rowpos(tgt - r0)] = val;
                    else hdst[(size_t)tgt * Kp + l] = val;
                    if (i <= 2) sl.phi[((size_t)((i + 1) & 1) * B1 + tgt) * Kp + l] = val;
                }
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) st_release(myflag, (unsigned long long)tick);
            cur = nxt;
        }
        ++tick;  // the next subproblem's terminal stage
    }
}

// ---- host side -----------------------------------------------------------------------------------
struct Variant { int TB, TL; };
static const Variant kVariants[] = {{7, 4}, {8, 4}, {4, 4}, {8, 2}, {8, 1}, {4, 1}};
static const int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);

static void fill_geometry(const Tables &t, int argw, int G, int JS, int v, WaveCfg &c)
{
    c.variant = v;
    c.TB = kVariants[v].TB;
    c.TL = kVariants[v].TL;
    const int TBP = (c.TB + 1) & ~1;
    c.G = G;
    const int rows_per_cta = (t.B1 + G - 1) / G;
    c.RG = (rows_per_cta + c.TB - 1) / c.TB;
    c.R = c.RG * c.TB;
    c.G = (t.B1 + c.R - 1) / c.R;  // drop CTAs that would own no row
    c.nLG = (t.K + c.TL - 1) / c.TL;
    c.JS = JS;
    c.jper = (t.K + JS - 1) / JS;
    c.tpg = ((c.RG * c.nLG + 31) / 32) * 32;
    c.RP = c.RG * TBP;
    c.threads = c.JS * c.tpg;
    c.smem = carve(t, c, argw, nullptr, nullptr);
}

bool wave_configure(const Tables &t, int argw, int num_sms, size_t smem_max, int want_ctas, int want_js,
                    int want_variant, WaveCfg &cfg)
{
    if (t.M > kMaxM || t.K > 4096) return false;
    // candidate search: prefer many CTAs, full lanes, and enough rows per CTA to amortise a stage
    double best_score = -1.;
    bool found = false;
    for (int v = 0; v < kNumVariants; ++v) {
        if (want_variant > 0 && v != want_variant - 1) continue;
        // level groups must cover Kp-aligned vector loads: TL*nLG <= Kp is guaranteed by Kp = roundup16(K)
        if (((t.K + kVariants[v].TL - 1) / kVariants[v].TL) * kVariants[v].TL > t.Kp) continue;
        for (int js = 1; js <= 16; ++js) {
            if (want_js > 0 && js != want_js) continue;
            if (js > t.K) break;
            int gmax = want_ctas > 0 ? want_ctas : num_sms;
            if (gmax > num_sms) gmax = num_sms;
            // tiny stages do not pay for inter-CTA pipelining: at least ~2 row tiles of work per CTA
            WaveCfg c = cfg;
            fill_geometry(t, argw, gmax, js, v, c);
            if (c.threads > kMaxWaveThreads || c.threads < 32) continue;
            if (c.smem > smem_max) continue;
            // score: relaxations per issue slot.  lanes used x tile efficiency, minus combine overhead.
            const double lanes = (double)(c.RG * c.nLG) / c.tpg;
            const double rows_eff = (double)t.B1 / ((double)c.G * c.R);
            const double lev_eff = (double)t.K / (c.nLG * c.TL);
            const double tile = (double)(c.TB * c.TL);
            const double per_j = 5.0 * tile + (c.TB + 1) / 2 + (c.TL + 1) / 2 + c.TL;  // issue slots per j
            const double combine = 12.0 * js / (double)((t.K + js - 1) / js);          // per cell, amortised
            const double eff = lanes * rows_eff * lev_eff * (5.0 * tile) / (per_j + combine * tile / 5.0);
            const double par = (double)c.G * c.threads;  // parallel lanes in flight
            double score = eff * (par < 148.0 * 128 ? par / (148.0 * 128) : 1.0);
            if (c.threads >= 256) score *= 1.02;  // two warps per scheduler hide barrier bubbles
            if (score > best_score) { best_score = score; cfg = c; found = true; }
        }
    }
    return found;
}

template <int TB, int TL>
static cudaError_t launch_variant(const Tables &t, const WaveCfg &cfg, int argw, cudaStream_t st)
{
    void *args[] = {(void *)&t, (void *)&cfg};
    const void *fn = (argw == 1) ? (const void *)wavefront_kernel<TB, TL, uint8_t>
                                 : (const void *)wavefront_kernel<TB, TL, uint16_t>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem);
    if (e != cudaSuccess) return e;
    // cooperative launch: the CTAs wait on one another, so all of them must be co-resident
    return cudaLaunchCooperativeKernel(fn, dim3(cfg.G), dim3(cfg.threads), args, cfg.smem, st);
}

cudaError_t launch_wavefront(const Tables &t, const WaveCfg &cfg, int argw, cudaStream_t st)
{
    switch (cfg.variant) {
        case 0: return launch_variant<7, 4>(t, cfg, argw, st);
        case 1: return launch_variant<8, 4>(t, cfg, argw, st);
        case 2: return launch_variant<4, 4>(t, cfg, argw, st);
        case 3: return launch_variant<8, 2>(t, cfg, argw, st);
        case 4: return launch_variant<8, 1>(t, cfg, argw, st);
        case 5: return launch_variant<4, 1>(t, cfg, argw, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace bb200
