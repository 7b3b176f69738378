// kernels_common.cu -- prep / per-stage (validation path) / selection / backtrack kernels.
//
// Reference semantics ("Normative semantics" S1-S10) of
//   bellman_TRM!  HelpFunctions.jl:20-83      eval_u_TRM!  HelpFunctions.jl:98-124
#include "bb200_internal.cuh"
#include "kernels.cuh"

namespace bb200 {

// ------------------------------------------------------------------------------------------------
// prep: validates u_old (Julia's InexactError, HelpFunctions.jl:37,57), counts the exact number of
// innermost-loop executions N, finds the largest budget use <= B (pipeline halo depth)
// and resets the exit-state rows to +Inf (HelpFunctions.jl:27,47).
// ------------------------------------------------------------------------------------------------
__global__ void prep_kernel(Tables t, SlotDev slot, int *err, int *btmax)
{
    __shared__ unsigned long long s_upd;
    __shared__ int s_bad, s_max;
    if (threadIdx.x == 0) { s_upd = 0ull; s_bad = 0; s_max = 0; }
    __syncthreads();
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nthr = (long long)gridDim.x * blockDim.x;
    const double inf = d_inf();
    for (long long x = gtid; x < 2ll * t.B1 * t.Kp; x += nthr) slot.phi[x] = inf;
    unsigned long long upd = 0ull;
    int bad = 0, mx = 0;
    // one thread per (stage row, level): stage cost and budget use of every stage (S3), kept for both kernel paths
    for (long long x = gtid; x < (long long)t.n * t.Kp; x += nthr) {
        const long long row = x / t.Kp;
        const int l = (int)(x - row * t.Kp);
        const double *uo = slot.u_old + row * t.M;
        double s = 0.;
        int bt = t.B1;
        if (l < t.K) {
            stage_cost(t, t.lvd + l * t.M, slot.df + row * t.M, uo, s, bt);
            if (bt < t.B1) {
                if (row < t.n - 1) upd += (unsigned long long)(t.B1 - bt) * (unsigned long long)t.K;
                mx = max(mx, bt);
            }
        }
        slot.ss_all[x] = s;
        slot.bt_all[x] = bt;
        if (l == 0)
            for (int m = 0; m < t.M; ++m) {
                const double v = uo[m];
                if (!(fabs(v) <= 1073741824.0) || v != floor(v)) bad = 1;
            }
    }
    atomicAdd(&s_upd, upd);
    if (bad) atomicOr(&s_bad, 1);
    atomicMax(&s_max, mx);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_upd) atomicAdd(slot.n_updates, s_upd);
        if (s_bad) {
            atomicOr(&err[0], 1);
            slot.rec[4 * kMaxRadii] = 1.;  // per-slot flag: one bad subproblem must not condemn its whole batch
        }
        atomicMax(btmax, s_max);
    }
}

// ------------------------------------------------------------------------------------------------
// Validation path: one launch per stage, value rows ping-pong in HBM.  Thread = (source row b', level l).
//   terminal stage (HelpFunctions.jl:27-43): phi[t][l] = (t == b~_l(n)) ? s_l(n) : Inf
//   stage i (HelpFunctions.jl:45-82): target (b'+b~_l, l) = min_j (s_l + c_jl) + phi_next[b'][j],
//   strict '>' scan in iterator order, so the earliest j wins ties and Inf/NaN never win.
// ------------------------------------------------------------------------------------------------
__global__ void terminal_kernel(Tables t, SlotDev slot, int pslot)
{
    const int l = threadIdx.x;
    const int b = blockIdx.x * blockDim.y + threadIdx.y;
    if (l >= t.K || b >= t.B1) return;
    const long long row = t.n - 1;
    const double s = slot.ss_all[row * t.Kp + l];
    const int bt = slot.bt_all[row * t.Kp + l];
    slot.phi[((long long)pslot * t.B1 + b) * t.Kp + l] = (b == bt) ? s : d_inf();
}

template <typename ArgT>
__global__ void stage_kernel(Tables t, SlotDev slot, int i /* 1-based stage */)
{
    constexpr ArgT MARK = (ArgT)~(ArgT)0;
    const int l = threadIdx.x;
    const int bsrc = blockIdx.x * blockDim.y + threadIdx.y;
    if (l >= t.K || bsrc >= t.B1) return;
    const int cur = (i + 1) & 1, nxt = i & 1;  // slot(i) = (i+1)%2 0-based, slot(i+1) = i%2
    const long long row = i - 1;
    const double s = slot.ss_all[row * t.Kp + l];
    const int bt = slot.bt_all[row * t.Kp + l];
    double *pc = slot.phi + (long long)cur * t.B1 * t.Kp;
    const double *pn = slot.phi + ((long long)nxt * t.B1 + bsrc) * t.Kp;
    if (bsrc < bt) pc[(long long)bsrc * t.Kp + l] = d_inf();  // target rows nobody reaches (:47)
    const int tgt = bsrc + bt;
    if (tgt >= t.B1) return;  // outside `for b = 0:B-b~` (:69)
    double best = d_inf();
    int arg = (int)MARK;
    for (int j = 0; j < t.K; ++j) {
        const double a = __dadd_rn(s, t.cost[j * t.Kp + l]);  // :67
        const double v = __dadd_rn(a, pn[j]);                 // :71
        if (best > v) { best = v; arg = j; }                  // :73-76
    }
    pc[(long long)tgt * t.Kp + l] = best;
    reinterpret_cast<ArgT *>(slot.arg)[((long long)(i - 1) * t.B1 + bsrc) * t.Kp + l] = (ArgT)arg;
}

// ------------------------------------------------------------------------------------------------
// Small-problem path: one CTA walks all stages of one subproblem with both value rows in shared memory
// (the reference's rolling 2-slot Phi, HelpFunctions.jl:27,47,71) and one __syncthreads per stage.  Used when a
// stage is too little work to spread over the GPU (K = 3..5 example shapes: ~1.5e3 candidates per stage), where
// any inter-CTA hand-over would cost more than the stage itself.  blockIdx.x = subproblem slot, so a batch of
// small subproblems runs on as many SMs concurrently.  Same arithmetic as stage_kernel (thread = source cell,
// straight scan in iterator order); the stage's level costs arrive in 32-stage chunks by 1-D bulk TMA.
// ------------------------------------------------------------------------------------------------
constexpr int kMiniChunk = 32;

template <typename ArgT>
__global__ void __launch_bounds__(1024, 1) mini_kernel(Tables t, const SlotDev *slots)
{
    constexpr ArgT MARK = (ArgT)~(ArgT)0;
    extern __shared__ __align__(128) unsigned char smem_mini[];
    const SlotDev slot = slots[blockIdx.x];
    const int K = t.K, Kp = t.Kp, B1 = t.B1, n = t.n;
    const int cells = B1 * K;
    // layout: mbar[2] | ss[2][chunk][Kp] | bt[2][chunk][Kp] | cost[K][K] | P[2][B1][K]
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_mini);
    double *ssb = reinterpret_cast<double *>(smem_mini + 128);
    int *btb = reinterpret_cast<int *>(ssb + 2 * kMiniChunk * Kp);
    double *cs = reinterpret_cast<double *>(btb + 2 * kMiniChunk * Kp);
    double *P = cs + (size_t)K * K;
    const int tid = threadIdx.x, NT = blockDim.x;
    const double inf = d_inf();
    ArgT *arg = reinterpret_cast<ArgT *>(slot.arg);
    for (int x = tid; x < K * K; x += NT) cs[x] = t.cost[(x / K) * Kp + (x % K)];
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // chunk c holds stage rows [c*chunk, (c+1)*chunk); rows are consumed downwards from n-1
    auto load_chunk = [&](int c) {
        const int b = c & 1;
        const int row0 = c * kMiniChunk;
        const int rows = min(kMiniChunk, n - row0);
        mbar_expect_tx(&mbar[b], (uint32_t)(rows * Kp * (sizeof(double) + sizeof(int))));
        tma_load_1d(ssb + (size_t)b * kMiniChunk * Kp, slot.ss_all + (size_t)row0 * Kp, (uint32_t)(rows * Kp * sizeof(double)), &mbar[b]);
        tma_load_1d(btb + (size_t)b * kMiniChunk * Kp, slot.bt_all + (size_t)row0 * Kp, (uint32_t)(rows * Kp * sizeof(int)), &mbar[b]);
    };
    int chunk_cur = (n - 1) / kMiniChunk;
    uint32_t phase_bits = 0;
    if (tid == 0) {
        load_chunk(chunk_cur);
        if (chunk_cur >= 1) load_chunk(chunk_cur - 1);
    }
    mbar_wait(&mbar[chunk_cur & 1], (phase_bits >> (chunk_cur & 1)) & 1u);
    phase_bits ^= 1u << (chunk_cur & 1);
    // terminal stage n (HelpFunctions.jl:27-43) into the slot of stage n
    {
        const double *sn = ssb + ((size_t)(chunk_cur & 1) * kMiniChunk + ((n - 1) % kMiniChunk)) * Kp;
        const int *bn = btb + ((size_t)(chunk_cur & 1) * kMiniChunk + ((n - 1) % kMiniChunk)) * Kp;
        double *Pt = P + (size_t)((n + 1) & 1) * cells;
        for (int x = tid; x < cells; x += NT) {
            const int b = x / K, l = x - b * K;
            Pt[x] = (b == bn[l]) ? sn[l] : inf;
        }
    }
    __syncthreads();
    for (int i = n - 1; i >= 1; --i) {
        const int ri = i - 1, ch = ri / kMiniChunk;
        if (ch != chunk_cur) {
            // every thread finished reading chunk ch+1 before the barrier that ended the previous stage
            chunk_cur = ch;
            if (tid == 0 && ch >= 1) load_chunk(ch - 1);
            mbar_wait(&mbar[ch & 1], (phase_bits >> (ch & 1)) & 1u);
            phase_bits ^= 1u << (ch & 1);
        }
        const double *ss = ssb + ((size_t)(ch & 1) * kMiniChunk + (ri % kMiniChunk)) * Kp;
        const int *bt = btb + ((size_t)(ch & 1) * kMiniChunk + (ri % kMiniChunk)) * Kp;
        double *Pc = P + (size_t)((i + 1) & 1) * cells;       // slot(i)   written
        const double *Pn = P + (size_t)(i & 1) * cells;       // slot(i+1) read
        for (int x = tid; x < cells; x += NT) {
            const int bsrc = x / K, l = x - bsrc * K;
            const int b = bt[l];
            if (bsrc < b) Pc[x] = inf;                        // target rows nobody reaches (:47)
            const int tgt = bsrc + b;
            if (tgt >= B1) continue;                          // outside `for b = 0:B-b~` (:69)
            const double s = ss[l];
            const double *pn = Pn + (size_t)bsrc * K;
            double best = inf;
            int a = (int)MARK;
            for (int j = 0; j < K; ++j) {
                const double v = __dadd_rn(__dadd_rn(s, cs[j * K + l]), pn[j]);  // :67, :71
                if (best > v) { best = v; a = j; }                                // :73-76
            }
            Pc[(size_t)tgt * K + l] = best;
            arg[((size_t)(i - 1) * B1 + bsrc) * Kp + l] = (ArgT)a;
        }
        __syncthreads();
    }
    // exit state (S7): stage-1 values in slot 0, stage-2 values in slot 1 (the prep kernel pre-filled +Inf)
    for (int sl = 0; sl < 2; ++sl) {
        if (n == 1 && sl == 1) break;
        const double *Ps = P + (size_t)sl * cells;
        for (int x = tid; x < cells; x += NT) {
            const int b = x / K, l = x - b * K;
            slot.phi[((size_t)sl * B1 + b) * Kp + l] = Ps[x];
        }
    }
}

size_t mini_smem_bytes(const Tables &t)
{
    return 128 + (size_t)2 * kMiniChunk * t.Kp * (sizeof(double) + sizeof(int)) + (size_t)t.K * t.K * sizeof(double) +
           (size_t)2 * t.B1 * t.K * sizeof(double) + 64;
}

// The small-problem kernel is chosen when a stage has little work and everything fits in shared memory.
bool mini_applicable(const Tables &t, size_t smem_max)
{
    const double relax_per_stage = (double)t.B1 * t.K * t.K;
    return relax_per_stage <= 60000. && mini_smem_bytes(t) <= smem_max;
}

cudaError_t launch_mini(const Tables &t, const SlotDev *d_slots, int count, int argw, cudaStream_t st)
{
    const size_t smem = mini_smem_bytes(t);
    int threads = ((t.B1 * t.K + 31) / 32) * 32;
    if (threads > 1024) threads = 1024;
    if (threads < 64) threads = 64;
    cudaError_t e;
    if (argw == 1) {
        e = cudaFuncSetAttribute(mini_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        mini_kernel<uint8_t><<<count, threads, smem, st>>>(t, d_slots);
    } else {
        e = cudaFuncSetAttribute(mini_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        mini_kernel<uint16_t><<<count, threads, smem, st>>>(t, d_slots);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Selection (HelpFunctions.jl:102-112): argmin over phi[0][0..Bnew][*] in the reference's column-major
// order (budget fastest, then grid offset) with Julia 1.10 findmin semantics (S8).  One CTA; per-thread
// scan, warp-shuffle reduction, then across warps through shared memory.
// ------------------------------------------------------------------------------------------------
struct Cand { double v; long long pos; int k; int b; };

__device__ __forceinline__ bool cand_precedes(const Cand &x, const Cand &y)
{
    if (julia_isgreater(y.v, x.v)) return true;
    if (julia_isgreater(x.v, y.v)) return false;
    return x.pos < y.pos;
}

// slot / radius / output of this CTA in a batched sweep launch
__device__ __forceinline__ void sweep_slot(const SweepDev &sw, SlotDev &slot, int &Bnew)
{
    if (!sw.slots) return;
    const int s = blockIdx.x / sw.n_radii, r = blockIdx.x - s * sw.n_radii;
    slot = sw.slots[s];
    slot.rec += 4 * r;
    slot.u = sw.u_sweep + (size_t)blockIdx.x * sw.u_stride;
    Bnew = sw.radii[r];
}

__global__ void select_kernel(Tables t, SlotDev slot, int Bnew_arg, const int *bnew_ptr, int *err, SweepDev sw)
{
    // the graph-replayed path passes the trial budget through device memory so that one captured graph serves every B'
    int Bnew = bnew_ptr ? *bnew_ptr : Bnew_arg;
    sweep_slot(sw, slot, Bnew);
    if (Bnew < 0) Bnew = 0;
    if (Bnew > t.B1 - 1) Bnew = t.B1 - 1;
    __shared__ Cand s_c[32];
    Cand best;
    best.v = d_inf(); best.pos = 0x7fffffffffffffffLL; best.k = -1; best.b = -1;
    const long long cells = (long long)(Bnew + 1) * t.K;
    for (long long x = threadIdx.x; x < cells; x += blockDim.x) {
        const int b = (int)(x / t.K), k = (int)(x % t.K);
        Cand c;
        c.v = slot.phi[(long long)b * t.Kp + k];
        c.pos = t.goff[k] * (long long)t.B1 + b;
        c.k = k; c.b = b;
        if (cand_precedes(c, best)) best = c;
    }
    for (int off = 16; off > 0; off >>= 1) {
        Cand o;
        o.v = __shfl_down_sync(0xffffffffu, best.v, off);
        o.pos = __shfl_down_sync(0xffffffffu, best.pos, off);
        o.k = __shfl_down_sync(0xffffffffu, best.k, off);
        o.b = __shfl_down_sync(0xffffffffu, best.b, off);
        if (cand_precedes(o, best)) best = o;
    }
    if ((threadIdx.x & 31) == 0) s_c[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x < 32) {
        const int nw = (blockDim.x + 31) >> 5;
        Cand c = s_c[threadIdx.x < nw ? threadIdx.x : 0];
        for (int off = 16; off > 0; off >>= 1) {
            Cand o;
            o.v = __shfl_down_sync(0xffffffffu, c.v, off);
            o.pos = __shfl_down_sync(0xffffffffu, c.pos, off);
            o.k = __shfl_down_sync(0xffffffffu, c.k, off);
            o.b = __shfl_down_sync(0xffffffffu, c.b, off);
            if (cand_precedes(o, c)) c = o;
        }
        if (threadIdx.x == 0) {
            slot.rec[0] = c.v;
            slot.rec[1] = (double)c.b;
            slot.rec[2] = (double)c.k;
            // +Inf selected: no feasible trajectory; the reference would go on to read stale U.
            const bool stale = (c.k < 0) || (c.v == d_inf());
            slot.rec[3] = stale ? 1. : 0.;
            if (stale) atomicOr(&err[1], 1);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Backtrack (HelpFunctions.jl:108-122): a dependent chase of n-1 argmin entries.  With the source-row
// indexed table one step is  b' = b - b~_l(i);  l <- arg[i-1][b'][l];  b <- b'.
// The budget only shrinks along the trajectory, so the entries the chase can touch in the next S stages lie
// in a window of W budget rows below the current b.  The CTA brings that window (S x W x Kp entries, one
// contiguous block per stage) and the stages' budget uses into shared memory by 1-D bulk TMA, one thread
// chases through shared memory (two dependent LDS per stage instead of two dependent HBM reads), then all
// threads write the S controls.  Two buffers: while chunk c is chased, the window of chunk c+1 is already in
// flight, placed below the budget known at the START of chunk c (the budget moves little within a chunk).  A
// prefetch that turns out misplaced (the chunk ended early, or the budget fell too far) is dropped and the
// window is reloaded around the actual budget.
// ------------------------------------------------------------------------------------------------
template <typename ArgT>
__global__ void __launch_bounds__(256, 1) backtrack_kernel(Tables t, SlotDev slot, int *err, int S, int W, SweepDev sw)
{
    { int unused = 0; sweep_slot(sw, slot, unused); }
    extern __shared__ __align__(128) unsigned char smem_bt[];
    const int Kp = t.Kp;
    const uint32_t win_bytes = (uint32_t)(((size_t)S * W * Kp * sizeof(ArgT) + 127) / 128 * 128);
    const uint32_t buf_bytes = win_bytes + (uint32_t)((size_t)S * Kp * sizeof(int));
    // buffer k: window ArgT[S][W][Kp] at k * buf_bytes, budget uses int[S][Kp] behind it
    int *lseq = reinterpret_cast<int *>(smem_bt + 2 * (size_t)buf_bytes);  // [S]
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ int s_b, s_l, s_done, s_fail;
    if (slot.rec[3] != 0.) return;
    const int tid = threadIdx.x, NT = blockDim.x;
    const ArgT *arg = reinterpret_cast<const ArgT *>(slot.arg);
    if (tid == 0) {
        s_b = (int)slot.rec[1];
        s_l = (int)slot.rec[2];
        s_fail = 0;
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    for (int m = tid; m < t.M; m += NT) slot.u[m] = t.lvd[s_l * t.M + m];  // u[:,1] (:108-112)

    // bulk-TMA the window of stages i0 .. i0+cnt-1 below budget row `top` and their budget uses into buffer k;
    // returns the window's first budget row
    auto issue = [&](int k, long long i0, int top) {
        const int cnt = (int)min((long long)S, (long long)t.n - i0);
        int wb = top - W + 1;
        if (wb < 0) wb = 0;
        const int wrows = min(W, t.B1 - wb);
        if (tid >= 32 && tid < 64) {  // warp 1 issues: warp 0 hosts the chase thread and must not be delayed
            const int ln = tid - 32;
            unsigned char *base = smem_bt + (size_t)k * buf_bytes;
            const uint32_t wbytes = (uint32_t)((size_t)wrows * Kp * sizeof(ArgT));
            if (ln == 0)
                mbar_expect_tx(&s_bar[k], (uint32_t)cnt * wbytes + (uint32_t)((size_t)cnt * Kp * sizeof(int)));
            __syncwarp();
            for (int st = ln; st < cnt; st += 32)
                tma_load_1d(base + (size_t)st * W * Kp * sizeof(ArgT), arg + ((size_t)(i0 + st - 1) * t.B1 + wb) * Kp, wbytes,
                            &s_bar[k]);
            if (ln == 0)
                tma_load_1d(base + win_bytes, slot.bt_all + (size_t)(i0 - 1) * Kp, (uint32_t)((size_t)cnt * Kp * sizeof(int)),
                            &s_bar[k]);
        }
        return wb;
    };

    // what the current / the other buffer holds or has in flight (identical in every thread; scalars, not arrays,
    // so that nothing lives in local memory)
    int cur = 0;
    long long c_i0 = 0, o_i0 = 0;
    int c_wb = 0, o_wb = 0;
    bool c_live = false, o_live = false;
    uint32_t c_phase = 0, o_phase = 0;
    for (long long i0 = 1; i0 <= t.n - 1;) {
        const int cnt = (int)min((long long)S, (long long)t.n - i0);  // stages i0 .. i0+cnt-1
        const int b0 = s_b;
        // is the block in flight for this buffer the one we need, with room for the budget to move?
        const bool usable = c_live && c_i0 == i0 && b0 - c_wb >= min(W / 2, b0);
        if (!usable) {
            if (c_live) {  // let the misplaced block land before its buffer is overwritten
                mbar_wait(&s_bar[cur], c_phase);
                c_phase ^= 1u;
            }
            asm volatile("fence.proxy.async;" ::: "memory");
            c_wb = issue(cur, i0, b0);
            c_i0 = i0;
        }
        mbar_wait(&s_bar[cur], c_phase);
        c_phase ^= 1u;
        c_live = false;
        // the next chunk's window goes out now (the other buffer was consumed in the previous iteration)
        if (i0 + cnt <= t.n - 1) {
            if (o_live) {
                mbar_wait(&s_bar[cur ^ 1], o_phase);
                o_phase ^= 1u;
            }
            o_wb = issue(cur ^ 1, i0 + cnt, b0);
            o_i0 = i0 + cnt;
            o_live = true;
        }
        // ---- chase through shared memory -----------------------------------------------------------------
        if (tid == 0) {
            const ArgT *wn = reinterpret_cast<const ArgT *>(smem_bt + (size_t)cur * buf_bytes);
            const int *bt = reinterpret_cast<const int *>(smem_bt + (size_t)cur * buf_bytes + win_bytes);
            const int wb = c_wb, WK = W * Kp;
            int b = b0, l = s_l, st = 0, fail = 0;
            // Fast path: almost every chunk stays inside its window and steps only on written cells.  The chase is one
            // dependent chain (budget use of the current level -> source row -> argmin entry -> next level); walk it without
            // a branch on the chain -- indices are clamped so that every load stays inside the window, the two conditions
            // are collected in a flag -- and fall back to the careful loop below, from the start of the chunk, if the flag
            // says the walk left the window or met an unwritten cell.  (232 -> ~100 cycles per stage.)
            {
                int bb = b0, ll = s_l;
                unsigned int ok = 1u;
                const ArgT *w = wn;
                const int *btp = bt;
                const int Km1 = t.K - 1;
#pragma unroll 4
                for (int k = 0; k < cnt; ++k, w += WK, btp += Kp) {
                    const int bsrc = bb - btp[ll];
                    const int rel = bsrc - wb;                      // <= W - 1: the budget only shrinks below the window's top
                    ok &= (unsigned int)(rel >= 0);
                    const int a = (int)w[max(rel, 0) * Kp + ll];
                    ok &= (unsigned int)(a <= Km1);                 // MARK (all ones) or garbage: no candidate won this cell
                    ll = min(a, Km1);
                    bb = bsrc;
                    lseq[k] = ll;
                }
                if (ok) { b = bb; l = ll; st = cnt; }
            }
            for (; st < cnt; ++st, wn += WK, bt += Kp) {
                const int bsrc = b - bt[l];
                ArgT a;
                if (bsrc >= wb) {                      // the common case: one branch, two dependent shared-memory loads
                    a = wn[(bsrc - wb) * Kp + l];
                } else {
                    if (bsrc < 0) { fail = 1; break; }  // unreachable level: the reference would read stale U
                    if (st > 0) break;                  // left the window: re-centre
                    // a first step that jumps below the window reads its one entry straight from HBM
                    a = arg[((size_t)(i0 + st - 1) * t.B1 + bsrc) * Kp + l];
                }
                if ((int)a >= t.K) { fail = 1; break; }  // MARK (all ones) or garbage: no candidate won this cell
                l = (int)a;
                b = bsrc;
                lseq[st] = l;
            }
            s_b = b; s_l = l; s_done = st; s_fail = fail;
        }
        __syncthreads();
        const int done = s_done;
        // ---- u[:, i+1] = nu_l for the stages walked (:117-119) ---------------------------------------------
        for (int x = tid; x < done * t.M; x += NT) {
            const int st = x / t.M, m = x - st * t.M;
            slot.u[(size_t)(i0 + st) * t.M + m] = t.lvd[lseq[st] * t.M + m];
        }
        if (s_fail) {
            if (tid == 0) { atomicOr(&err[1], 1); slot.rec[3] = 1.; }
            break;
        }
        i0 += done;  // done >= 1: the first step of a chunk always completes
        // swap the roles of the two buffers
        cur ^= 1;
        { const long long x = c_i0; c_i0 = o_i0; o_i0 = x; }
        { const int x = c_wb; c_wb = o_wb; o_wb = x; }
        { const bool x = c_live; c_live = o_live; o_live = x; }
        { const uint32_t x = c_phase; c_phase = o_phase; o_phase = x; }
        __syncthreads();  // everyone is finished with the window and lseq before they are overwritten
    }
    // never leave with a bulk copy in flight into this CTA's shared memory
    if (c_live) mbar_wait(&s_bar[cur], c_phase);
    if (o_live) mbar_wait(&s_bar[cur ^ 1], o_phase);
}

// ------------------------------------------------------------------------------------------------
// "Next" rows: sequential-order reductions on the resident arrays.
//   N1  int_val = dt * sum_j df[:,j]'(u_old[:,j]-u[:,j])   multi-trust.jl:117-121
//   N3  TV_p(u,p), p in {Inf, 1, 2}                           HelpFunctions.jl:251-268
// The floating-point sums are order dependent, so one thread accumulates in the reference's order while
// the CTA stages the operands through shared memory in coalesced chunks.
// ------------------------------------------------------------------------------------------------
__global__ void pred_integral_kernel(Tables t, SlotDev slot, double *out)
{
    extern __shared__ double sh[];
    double *term = sh;  // [chunk]
    const int chunk = blockDim.x;
    double acc = 0.;
    for (long long j0 = 0; j0 < t.n; j0 += chunk) {
        const long long j = j0 + threadIdx.x;
        if (j < t.n) {
            double dot = 0.;
            for (int m = 0; m < t.M; ++m) {
                const double x = __dmul_rn(slot.df[j * t.M + m],
                                           __dsub_rn(slot.u_old[j * t.M + m], slot.u[j * t.M + m]));
                dot = (m == 0) ? x : __dadd_rn(dot, x);
            }
            term[threadIdx.x] = dot;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const int cnt = (int)min((long long)chunk, t.n - j0);
            for (int k = 0; k < cnt; ++k) acc = __dadd_rn(acc, term[k]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = __dmul_rn(acc, t.dt);
}

__global__ void tv_kernel(Tables t, SlotDev slot, int mode /*0 = Inf, 1, 2*/, double *out)
{
    extern __shared__ double sh[];
    double *term = sh;
    const int chunk = blockDim.x;
    double acc = 0.;
    for (long long i0 = 1; i0 < t.n; i0 += chunk) {
        const long long i = i0 + threadIdx.x;
        if (i < t.n) {
            double v = 0.;
            if (mode == 0) {
                v = -d_inf();
                for (int m = 0; m < t.M; ++m) {
                    const double d = fabs(__dsub_rn(slot.u[i * t.M + m], slot.u[(i - 1) * t.M + m]));
                    if (d > v || d != d) v = d;
                }
            } else {
                for (int m = 0; m < t.M; ++m) {
                    const double d = fabs(__dsub_rn(slot.u[i * t.M + m], slot.u[(i - 1) * t.M + m]));
                    const double pw = (mode == 1) ? d : __dmul_rn(d, d);
                    v = (m == 0) ? pw : __dadd_rn(v, pw);
                }
                if (mode == 2) v = sqrt(v);  // integer-valued u: exact when the sum is a perfect square
            }
            term[threadIdx.x] = v;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const int cnt = (int)min((long long)chunk, t.n - i0);
            for (int k = 0; k < cnt; ++k) acc = __dadd_rn(acc, term[k]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = acc;
}

// ---- launchers ---------------------------------------------------------------------------------
void launch_prep(const Tables &t, const SlotDev &slot, int *err, int *btmax, cudaStream_t st)
{
    const int threads = 256;
    long long work = (long long)t.n * t.Kp > 2ll * t.B1 * t.Kp ? (long long)t.n * t.Kp : 2ll * t.B1 * t.Kp;
    int blocks = (int)((work / 4 + threads - 1) / threads);
    if (blocks > 1184) blocks = 1184;
    if (blocks < 1) blocks = 1;
    prep_kernel<<<blocks, threads, 0, st>>>(t, slot, err, btmax);
}

int launch_terminal_stage(const Tables &t, const SlotDev &slot, cudaStream_t st)
{
    int bx = ((t.K + 31) / 32) * 32;
    if (bx > 1024) return -1;
    int by = 1024 / bx;
    if (by > 8) by = 8;
    terminal_kernel<<<dim3((t.B1 + by - 1) / by), dim3(bx, by), 0, st>>>(t, slot, (t.n + 1) & 1);
    return 1;
}

int launch_stage_path(const Tables &t, const SlotDev &slot, int argw, cudaStream_t st)
{
    // blockDim.x = levels rounded to a warp, blockDim.y = source rows
    int bx = ((t.K + 31) / 32) * 32;
    if (bx > 1024) return -1;
    int by = 1024 / bx;
    if (by > 8) by = 8;
    dim3 block(bx, by);
    dim3 grid((t.B1 + by - 1) / by);
    int launches = launch_terminal_stage(t, slot, st);
    for (int i = t.n - 1; i >= 1; --i) {
        if (argw == 1) stage_kernel<uint8_t><<<grid, block, 0, st>>>(t, slot, i);
        else stage_kernel<uint16_t><<<grid, block, 0, st>>>(t, slot, i);
        ++launches;
    }
    return launches;
}

void launch_select(const Tables &t, const SlotDev &slot, int Bnew, const int *bnew_ptr, int *err, cudaStream_t st,
                   const SweepDev *sweep, int ctas)
{
    const SweepDev none{nullptr, nullptr, 1, nullptr, 0};
    select_kernel<<<sweep ? ctas : 1, 1024, 0, st>>>(t, slot, Bnew, bnew_ptr, err, sweep ? *sweep : none);
}

void launch_backtrack(const Tables &t, const SlotDev &slot, int argw, int *err, cudaStream_t st, const SweepDev *sweep,
                      int ctas)
{
    const SweepDev none{nullptr, nullptr, 1, nullptr, 0};
    const SweepDev sw = sweep ? *sweep : none;
    const int grid = sweep ? ctas : 1;
    // window rows W: at most 16 (a step that uses more budget than the window holds falls back to one HBM read);
    // stages per chunk S: at most 32, two buffers within ~200 KB of shared memory
    int W = t.B1 < 16 ? t.B1 : 16;
    auto per_stage = [&](int w) { return (size_t)w * t.Kp * argw + (size_t)t.Kp * sizeof(int); };
    long long S = (long long)(100 * 1024) / (long long)per_stage(W);
    if (S > 32) S = 32;
    if (S < 1) S = 1;
    while (W > 1 && 2 * (size_t)S * per_stage(W) > 200 * 1024) --W;
    const size_t win_bytes = ((size_t)S * W * t.Kp * argw + 127) / 128 * 128;
    const size_t smem = 2 * (win_bytes + (size_t)S * t.Kp * sizeof(int)) + (size_t)S * sizeof(int) + 128;
    if (argw == 1) {
        cudaFuncSetAttribute(backtrack_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        backtrack_kernel<uint8_t><<<grid, 256, smem, st>>>(t, slot, err, (int)S, W, sw);
    } else {
        cudaFuncSetAttribute(backtrack_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        backtrack_kernel<uint16_t><<<grid, 256, smem, st>>>(t, slot, err, (int)S, W, sw);
    }
}

void launch_pred_integral(const Tables &t, const SlotDev &slot, double *out, cudaStream_t st)
{
    pred_integral_kernel<<<1, 1024, 1024 * sizeof(double), st>>>(t, slot, out);
}

void launch_tv(const Tables &t, const SlotDev &slot, int mode, double *out, cudaStream_t st)
{
    tv_kernel<<<1, 1024, 1024 * sizeof(double), st>>>(t, slot, mode, out);
}

}  // namespace bb200
