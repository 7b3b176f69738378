// kernel_stage_pruned.cu -- per-stage DP kernels with the branch-and-bound scan, for level sets the persistent pipelined
// kernel does not take: a jump-cost table larger than shared memory (K > ~150) or a uint16 argmin table (K > 255).
//
// Same I/O as stage_kernel (kernels_common.cu): one launch per stage i = n-1 .. 1, value rows ping-pong in the exit-state
// buffers of the slot, no inter-CTA synchronisation at all.  A CTA owns TB = 4 consecutive SOURCE budget rows and all
// levels (thread = level, warp = 32 levels); the rows are staged in shared memory, the jump costs c[j][l] stay in global
// memory (a few hundred KB, L2 resident; lanes read consecutive levels: coalesced) in a copy padded with +Inf rows up to
// whole blocks, and the successor axis is walked in segments of 32 blocks with pruned_segment (pruned_scan.cuh): the
// bound tests drop most blocks, the surviving ones are scanned in ascending order with the reference's strict '>'
// (HelpFunctions.jl:69-78) -- value, argmin and ties are identical to the exhaustive scan.
#include "bb200_internal.cuh"
#include "kernels.cuh"
#include "pruned_scan.cuh"

namespace bb200 {

constexpr int kSpTB = 4;   // source rows per CTA
constexpr int kSpBK = 4;   // successors per block

// Per-plan tables (built once per plan from the jump-cost table):
//   cpad[Kp][Kp]    jump costs with +Inf rows K .. Kp-1 (a scanned block may reach past the last successor)
//   cminf[nblk][Kp] block minima of the jump costs rounded DOWN to float
//   cwv[Kp/32][nblk] per level block: the smallest cminf[q][l] over its live levels (row test)
//   cmx[Kp]         largest finite |jump cost| into level l, rounded up (slack of the row test)
__global__ void stage_pruned_setup_kernel(Tables t, StagePrunedTabs pt)
{
    const int Kp = t.Kp, K = t.K, nblk = pt.nblk;
    const double inf = d_inf();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int x = tid; x < Kp * Kp; x += nth) {
        const int j = x / Kp, l = x - j * Kp;
        pt.cpad[x] = (j < K && l < K) ? t.cost[(size_t)j * Kp + l] : inf;
    }
    for (int x = tid; x < nblk * Kp; x += nth) {
        const int q = x / Kp, l = x - q * Kp;
        double m = inf;
        for (int jj = 0; jj < kSpBK; ++jj) {
            const int j = q * kSpBK + jj;
            if (j < K && l < K) m = fmin(m, t.cost[(size_t)j * Kp + l]);
        }
        pt.cminf[x] = __double2float_rd(m);
    }
    for (int l = tid; l < Kp; l += nth) {
        float cmx = 0.f;
        if (l < K)
            for (int j = 0; j < K; ++j) {
                const double cj = t.cost[(size_t)j * Kp + l];
                if (fabs(cj) < inf) cmx = fmaxf(cmx, __double2float_ru(fabs(cj)));
            }
        pt.cmx[l] = cmx;
    }
}
// (second pass: needs cminf)
__global__ void stage_pruned_setup2_kernel(Tables t, StagePrunedTabs pt)
{
    const int Kp = t.Kp, K = t.K, nblk = pt.nblk, nLB = Kp >> 5;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int x = tid; x < nLB * nblk; x += nth) {
        const int w = x / nblk, q = x - w * nblk;
        float cw = __int_as_float(0x7f800000);
        for (int l = w * 32; l < min(w * 32 + 32, K); ++l) cw = fminf(cw, pt.cminf[(size_t)q * Kp + l]);
        pt.cwv[x] = cw;
    }
}

template <typename ArgT>
__global__ void __launch_bounds__(512) stage_pruned_kernel(Tables t, SlotDev slot, int i /* 1-based stage */, StagePrunedTabs pt,
                                                            unsigned long long *exec)
{
    constexpr int TB = kSpTB, BK = kSpBK;
    constexpr int MARKI = (int)(ArgT) ~(ArgT)0;
    extern __shared__ __align__(16) unsigned char sp_smem[];
    const int Kp = t.Kp, K = t.K, B1 = t.B1, nblk = pt.nblk, nseg = (nblk + 31) >> 5;
    double *Ps = reinterpret_cast<double *>(sp_smem);                        // [TB][Kp] value rows of my source rows
    float *pmf = reinterpret_cast<float *>(Ps + (size_t)TB * Kp);            // [nseg][TB * 32] block minima (pm_idx per segment)
    unsigned int *keymin = reinterpret_cast<unsigned int *>(pmf + (size_t)nseg * TB * 32);  // [TB]
    int *jseed = reinterpret_cast<int *>(keymin + TB);                      // [TB]
    const int l = threadIdx.x, lane = l & 31, wid = l >> 5;
    const int row0 = blockIdx.x * TB;
    const int rows_live = min(TB, B1 - row0);
    const double inf = d_inf();
    const int cur = (i + 1) & 1, nxt = i & 1;  // slot(i) = (i+1)%2 0-based, slot(i+1) = i%2  (as in stage_kernel)
    double *pc = slot.phi + (size_t)cur * B1 * Kp;
    const double *pn = slot.phi + (size_t)nxt * B1 * Kp;

    // 1. my source rows of the next stage's values; pad levels and rows beyond the table are +Inf (bound nothing, never win)
#pragma unroll
    for (int r = 0; r < TB; ++r) Ps[(size_t)r * Kp + l] = (r < rows_live && l < K) ? pn[(size_t)(row0 + r) * Kp + l] : inf;
    if (l < TB) keymin[l] = 0xffffffffu;
    __syncthreads();
    // 2. block minima (thread = block) and one seed successor per row (a smallest value: 25-bit key + block index)
    int jmin[TB];
    unsigned int key[TB];
    if (l < nblk) {
        const int seg = l >> 5, ql = l & 31, nb = min(32, nblk - 32 * seg);
#pragma unroll
        for (int r = 0; r < TB; ++r) {
            const double *p = Ps + (size_t)r * Kp + l * BK;
            const double m01 = fmin(p[0], p[1]), m23 = fmin(p[2], p[3]), m = fmin(m01, m23);
            jmin[r] = (m == m01) ? (m == p[0] ? 0 : 1) : (m == p[2] ? 2 : 3);
            const float mf = __double2float_rd(m);
            pmf[(size_t)seg * TB * 32 + pm_idx(r, ql, nb)] = mf;
            key[r] = (f2key(mf) & ~127u) | (unsigned int)l;
            atomicMin(&keymin[r], key[r]);
        }
    }
    __syncthreads();
    if (l < nblk) {
#pragma unroll
        for (int r = 0; r < TB; ++r)
            if (keymin[r] == key[r]) jseed[r] = l * BK + jmin[r];
    }
    __syncthreads();
    // 3. bounds, then the segments of 32 blocks in ascending order
    const bool live = l < K;
    const int lc = min(l, K - 1);
    const size_t srow = (size_t)(i - 1) * Kp;
    const double s = slot.ss_all[srow + lc];
    const int bt = slot.bt_all[srow + lc];
    const double *cs_l = pt.cpad + lc;
    const float *cmf_l = pt.cminf + lc;
    PrunedBounds<TB> pb;
    double vself[TB];
    pruned_self_candidates<TB>(Ps, cs_l, s, Kp, lc, vself);
    pruned_bounds<TB>(Ps, cs_l, jseed, s, pt.cmx[lc], Kp, live, rows_live, vself, pb);
    double best[TB][1];
    int arg[TB][1];
#pragma unroll
    for (int r = 0; r < TB; ++r) { best[r][0] = inf; arg[r][0] = MARKI; }
    unsigned int executed = 0;
    for (int seg = 0; seg < nseg; ++seg) {
        const int nb = min(32, nblk - 32 * seg);
        const float cw = pt.cwv[(size_t)wid * nblk + 32 * seg + min(lane, nb - 1)];
        executed += pruned_segment<TB, BK, ArgT>(Ps + seg * 32 * BK, cs_l + (size_t)seg * 32 * BK * Kp, cmf_l + (size_t)seg * 32 * Kp,
                                                 pmf + (size_t)seg * TB * 32, 0, pb, s, cw, nb, Kp, live, lane, seg * 32 * BK, best, arg);
    }
    // 4. results: exactly what stage_kernel writes
    if (live) {
#pragma unroll
        for (int r = 0; r < TB; ++r) {
            const int bsrc = row0 + r;
            if (bsrc >= B1) break;
            if (bsrc < bt) pc[(size_t)bsrc * Kp + l] = inf;  // target rows nobody reaches (:47)
            const int tgt = bsrc + bt;
            if (tgt < B1) {                                  // inside `for b = 0:B-b~` (:69)
                pc[(size_t)tgt * Kp + l] = best[r][0];
                reinterpret_cast<ArgT *>(slot.arg)[((size_t)(i - 1) * B1 + bsrc) * Kp + l] = (ArgT)arg[r][0];
            }
        }
    }
    if (exec && lane == 0 && executed) {
        const int lv = min(32, max(0, K - (wid << 5)));
        atomicAdd(exec, (unsigned long long)executed * (unsigned long long)(BK * rows_live * lv));
    }
}

bool stage_pruned_applicable(const Tables &t)
{
    // wide level sets only (narrow ones scan faster than they can be bounded); one thread per level, at most 128 blocks
    return t.K >= 64 && t.Kp <= 512 && t.n >= 2;
}

size_t stage_pruned_table_bytes(const Tables &t)
{
    const size_t Kp = (size_t)t.Kp, nblk = Kp / kSpBK;
    return Kp * Kp * sizeof(double) + nblk * Kp * sizeof(float) + (Kp / 32) * nblk * sizeof(float) + Kp * sizeof(float);
}

// `base` is one device allocation of stage_pruned_table_bytes(t); fills pt and builds the tables on `st`.
cudaError_t stage_pruned_setup(const Tables &t, void *base, StagePrunedTabs &pt, cudaStream_t st)
{
    const size_t Kp = (size_t)t.Kp, nblk = Kp / kSpBK;
    unsigned char *b = static_cast<unsigned char *>(base);
    pt.nblk = (int)nblk;
    pt.cpad = reinterpret_cast<double *>(b);
    pt.cminf = reinterpret_cast<float *>(b + Kp * Kp * sizeof(double));
    pt.cwv = pt.cminf + nblk * Kp;
    pt.cmx = pt.cwv + (Kp / 32) * nblk;
    stage_pruned_setup_kernel<<<64, 256, 0, st>>>(t, pt);
    stage_pruned_setup2_kernel<<<8, 256, 0, st>>>(t, pt);
    return cudaGetLastError();
}

int launch_stage_pruned_path(const Tables &t, const SlotDev &slot, int argw, const StagePrunedTabs &pt, unsigned long long *exec,
                             cudaStream_t st)
{
    const int nseg = (pt.nblk + 31) / 32;
    const size_t smem = (size_t)kSpTB * t.Kp * sizeof(double) + (size_t)nseg * kSpTB * 32 * sizeof(float) + 2 * kSpTB * sizeof(int) + 16;
    dim3 block(t.Kp);
    dim3 grid((t.B1 + kSpTB - 1) / kSpTB);
    int launches = launch_terminal_stage(t, slot, st);
    if (launches < 0) return -1;
    for (int i = t.n - 1; i >= 1; --i) {
        if (argw == 1) stage_pruned_kernel<uint8_t><<<grid, block, smem, st>>>(t, slot, i, pt, exec);
        else stage_pruned_kernel<uint16_t><<<grid, block, smem, st>>>(t, slot, i, pt, exec);
        ++launches;
    }
    return launches;
}

}  // namespace bb200
