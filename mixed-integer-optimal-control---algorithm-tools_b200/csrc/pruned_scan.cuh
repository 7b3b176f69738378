// pruned_scan.cuh -- the branch-and-bound ("pruned") scan of one warp tile: TB rows x 32 levels, successors in blocks of 4.
// Shared by the persistent pipelined kernel (kernel_wavefront.cu: value rows and jump costs in shared memory, at most 32
// blocks) and the per-stage kernels for wide level sets (kernel_stage_pruned.cu: jump costs in global memory / L2, the
// successor axis walked in segments of 32 blocks).
#pragma once
#include "bb200_internal.cuh"

namespace bb200 {

// An empty volatile asm inside the update keeps the front end from turning `if (best > v) { best = v; arg = j; }`
// into selects (DSETP + FSEL + FSEL + SEL: three instructions on the half-rate ALU pipe, which then binds the scan).
// ptxas if-converts the short branch itself into three PREDICATED MOVES and spreads them over the ALU and the FMA
// pipe (@P MOV / @P IMAD.MOV.U32), profiles/scan_probe_r02.txt.
#ifndef BB_SELECT_UPDATE
#define BB_KEEP_BRANCH asm volatile("")
#else
#define BB_KEEP_BRANCH
#endif

// ---- pruned scan ------------------------------------------------------------------------------------------
// The bound tests run in FP32 with DIRECTED rounding; only the surviving candidates are evaluated in FP64.  Why that is
// rigorous: a candidate is v = fl64(fl64(s + c_jl) + P[j]) (HelpFunctions.jl:67, :71).  With sf <= s, cf <= c_jl, pf <= P[j]
// (floats, rounded down) the float chain rd32(rd32(sf + cf) + pf) is a double that is <= the exact sum of its operands
// at every step, and fl64 is monotone and leaves doubles unchanged -- hence LB32 <= v, for every candidate of the block
// whose minima cf, pf are.  An upper bound UB of the cell's minimum is rounded UP to float.  LB32 > UBf therefore proves
// v > UB: the candidate can neither be the minimum nor tie with it.  No error analysis is involved in this test.
//
// order-preserving map float -> uint32 (for REDUX.MIN/MAX, which only take integers) and back
__device__ __forceinline__ unsigned int f2key(float x)
{
    const unsigned int u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned int k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

// Block minima of ONE value row, computed once per stage and row by one warp for all the warps that scan the row
// (lane = block of BK successors; nblk is a multiple of 8, at most 32):
//   pmr[q] = rd32(min_{j in block q} P[j])  (NaN ignored: a NaN candidate never wins; pad columns may hold anything finite
//   or +Inf, they only lower a bound),  *jseed = a successor with a (nearly) smallest value -- any successor gives a valid
//   upper bound, so a 27-bit key of the block minimum with the lane in the low bits and one REDUX.MIN are enough.
// Layout of the block minima of the CTA's value rows: the two rows of a pair are neighbours, pm[((r >> 1) * nblk + q) * 2 +
// (r & 1)], so that a warp that owns an even-aligned pair of rows takes both minima of a block with one 8-byte load.
__device__ __forceinline__ int pm_idx(int r, int q, int nblk) { return (((r >> 1) * nblk + q) << 1) + (r & 1); }

template <int BK>
__device__ __forceinline__ void row_minima(const double *__restrict__ Prow, float *__restrict__ pm, int r, int *__restrict__ jseed,
                                           int nblk, int lane)
{
    static_assert(BK == 4, "two successor pairs per block");
    unsigned int key = 0xffffffffu;
    int jmin = 0;
    if (lane < nblk) {
        const double2 w0 = *reinterpret_cast<const double2 *>(Prow + lane * BK);
        const double2 w1 = *reinterpret_cast<const double2 *>(Prow + lane * BK + 2);
        const double m01 = fmin(w0.x, w0.y), m23 = fmin(w1.x, w1.y), m = fmin(m01, m23);
        jmin = (m == m01) ? (m == w0.x ? 0 : 1) : (m == w1.x ? 2 : 3);   // (all NaN: any successor will do)
        const float mf = __double2float_rd(m);
        pm[pm_idx(r, lane, nblk)] = mf;
        key = (f2key(mf) & ~31u) | (unsigned int)lane;
    }
    const unsigned int kmin = __reduce_min_sync(0xffffffffu, key);
    if (lane == (int)(kmin & 31u) && lane < nblk) *jseed = lane * BK + jmin;   // the lane that holds the smallest key
}

// Pruned scan of one thread: TB rows x ONE level (lane = level), successors in blocks of BK.
//   Prow: the warp's value rows [r][Kp];  cs_l = cs + l: jump costs c[j][l] at stride Kp;  cmf_l: their block minima (float,
//   rounded down) at stride Kp;  pmf / qseed: the rows' block minima (float, rounded down) and seeds (row_minima), row
//   strides nblk / 1;  s: stage cost of this level;  cw: (lane = block q) the smallest cmf[q][l] over the live levels of
//   this warp;  cmx: the largest finite jump cost into this level (for the slack below).
//   1. UB[r]: an upper bound of every cell's minimum -- the candidate of its row's seed successor (a smallest value of the
//      row) and the no-jump candidate j = l, evaluated exactly (FP64).  (The other candidates of the seed's block tighten
//      the bound by < 5 % of the surviving blocks, tools/prune_stats.py, and cost more than that.);
//   2. row test, lane = block: block q of row r can matter to SOME level of this warp only if
//          cw[q] + pm[q][r]  <=  max_l (UB[r][l] - s_l)
//      -- one compare per (block, row) for the whole warp instead of one per (block, row, level).  The merge over the
//      levels is an argument in real arithmetic.  With pq = pmf[q][r] (a float, hence a double, <= every P[j] of the
//      block) a candidate is v = fl64(fl64(s + c) + P[j]) >= fl64(fl64(s + c) + pq) by monotonicity, and that is
//      >= s + c + pq - 2^-52 (|s| + |c|) - 2^-53 |pq| (two roundings to nearest).  The test passes only if
//      cw + pq - 2^-30 |pq|  >  UB - s + 2^-30 (|s| + cmx)  with cw <= c and cmx >= |c|, every float operation rounded
//      in the safe direction -- the slack covers the rounding terms a million times over, so v > UB.  Infinite and NaN
//      operands make the comparison false or the bound infinite: nothing is pruned;
//   3. level test, lane = level, on the blocks that survive 2.:  LB32[q][r] = rd32(rd32(sf + cmf[q][l]) + pmf[q][r]) > UBf[r]
//      (the monotone chain of the header comment).  The masks are OR-reduced over the warp: only a warp-uniform skip
//      saves issue slots;
//   4. the surviving blocks are scanned in ascending order with the reference's strict '>' from +Inf: the same minimum,
//      the same (earliest) argmin, bit for bit.
// FP64 adds / compares have ~40 cycles of latency on sm_100a: phases 1 and 4 are written as independent chains, and in the
// scan the next block's candidates are loaded and added while the compare -> move chain of the current block runs.
// Upper bounds and the per-row thresholds of the row test, computed once per stage and tile (all segments share them).
template <int TB>
struct PrunedBounds {
    double ub[TB];   // exact upper bound of every cell's minimum (-Inf: the cell asks for nothing)
    float ubf[TB];   // the same rounded up to float
    float U[TB];     // max over the warp's levels of (UB - s + slack), rounded up: the right-hand side of the row test
    float sf;        // stage cost rounded down to float
};

// The no-jump candidates j = l of a tile: they need neither the block minima nor the seeds, so a caller that has to wait for
// those (a barrier) can evaluate them first.
template <int TB>
__device__ __forceinline__ void pruned_self_candidates(const double *__restrict__ Prow, const double *__restrict__ cs_l, double s, int Kp,
                                                       int l_self, double (&vself)[TB])
{
    const double a = __dadd_rn(s, cs_l[(size_t)l_self * Kp]);
#pragma unroll
    for (int r = 0; r < TB; ++r) vself[r] = __dadd_rn(a, Prow[(size_t)r * Kp + l_self]);
}

template <int TB>
__device__ __forceinline__ void pruned_bounds(const double *__restrict__ Prow, const double *__restrict__ cs_l,
                                              const int *__restrict__ qseed, double s, float cmx, int Kp, bool live,
                                              int rows_live, const double (&vself_in)[TB], PrunedBounds<TB> &pb)
{
    const double inf = d_inf();
    const float finf = __int_as_float(0x7f800000);
    // ---- 1. upper bounds ---------------------------------------------------------------------------------
#pragma unroll
    for (int r = 0; r < TB; ++r) {
        const int js = qseed[r];
        const double vseed = __dadd_rn(__dadd_rn(s, cs_l[(size_t)js * Kp]), Prow[(size_t)r * Kp + js]);
        const double vself = vself_in[r];
        double u = inf;               // from +Inf with '>', so that a NaN candidate is ignored
        if (u > vseed) u = vseed;
        if (u > vself) u = vself;
        pb.ub[r] = u;
        if (!live || r >= rows_live) pb.ub[r] = -inf;  // pad levels and rows beyond the table never ask for a block
    }
    // ---- 2a. thresholds of the row test ---------------------------------------------------------------------
    pb.sf = __double2float_rd(s);
    // slack of the level side: 2^-30 (|s| + largest finite jump cost), rounded up
    const float sl_l = __fmul_ru(__fadd_ru(__double2float_ru(fabs(s)), cmx), 0x1p-30f);
#pragma unroll
    for (int r = 0; r < TB; ++r) {
        pb.ubf[r] = __double2float_ru(pb.ub[r]);
        float e = __fadd_ru(__double2float_ru(__dadd_ru(pb.ub[r], -s)), sl_l);   // >= UB - s + slack in real arithmetic
        if (!(e == e)) e = finf;                                                // NaN stage cost: prune nothing
        pb.U[r] = key2f(__reduce_max_sync(0xffffffffu, f2key(e)));
    }
}

// Row test, level test and scan of ONE segment of at most 32 blocks.  All pointers are those of the segment's first block
// (Prow: its first successor's column; cs_l / cmf_l: its first jump-cost row / block-minimum row; pmf: its block minima in
// the pm_idx layout with `nblk` blocks per row); jbase is the successor index of that block (added to the argmin);
// best / arg continue from the segments before (the caller starts them at +Inf / MARK).  Returns blocks scanned.
template <int TB, int BK, typename ArgT>
__device__ __forceinline__ unsigned int pruned_segment(const double *__restrict__ Prow, const double *__restrict__ cs_l,
                                                       const float *__restrict__ cmf_l, const float *__restrict__ pmf, int row0,
                                                       const PrunedBounds<TB> &pb, double s, float cw, int nblk, int Kp,
                                                       bool live, int lane, int jbase, double (&best)[TB][1], int (&arg)[TB][1])
{
    static_assert(TB <= 4 && BK == 4, "the scan takes two successor pairs per block");
    const float fmax = __int_as_float(0x7f7fffff);
    unsigned int executed = 0;
    // ---- 2b. row test (lane = block) -----------------------------------------------------------------------
    unsigned int cand = 0;
    const int qlane = min(lane, nblk - 1);
#pragma unroll
    for (int r = 0; r < TB; ++r) {
        const float pq = pmf[pm_idx(row0 + r, qlane, nblk)];
        const float slq = fminf(__fmul_ru(fabsf(pq), 0x1p-30f), fmax);        // slack of the block side
        const float t = __fadd_rd(__fadd_rd(cw, pq), -slq);                   // <= cw + pm - slack in real arithmetic
        cand |= __ballot_sync(0xffffffffu, lane < nblk && !(t > pb.U[r]));
    }
    // ---- 3. level test (lane = level) on the candidate blocks -----------------------------------------------
    unsigned int pneed = 0;
    for (unsigned int ms = cand; ms;) {
        // four candidate blocks per trip; an exhausted mask repeats the first one (its bit is 0: no effect)
        unsigned int bit[4];
        int q[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            bit[k] = ms & (0u - ms);
            ms ^= bit[k];
            q[k] = (k == 0 || bit[k]) ? 31 - __clz((int)bit[k]) : q[0];
        }
        float a[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) a[k] = __fadd_rd(pb.sf, cmf_l[(size_t)q[k] * Kp]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            // LB32 > UBf for every row of the warp: the block cannot matter to this level.  Straight-line code (no short
            // circuit: a branch per row costs more than the add and the compare it would skip)
            unsigned int above = 1u;
            if constexpr (TB % 2 == 0) {
#pragma unroll
                for (int r = 0; r < TB; r += 2) {   // row0 is a multiple of TB: the pair (r, r + 1) is one 8-byte load
                    const float2 pp = *reinterpret_cast<const float2 *>(pmf + pm_idx(row0 + r, q[k], nblk));
                    above &= (unsigned int)(__fadd_rd(a[k], pp.x) > pb.ubf[r]) & (unsigned int)(__fadd_rd(a[k], pp.y) > pb.ubf[r + 1]);
                }
            } else {
#pragma unroll
                for (int r = 0; r < TB; ++r) above &= (unsigned int)(__fadd_rd(a[k], pmf[pm_idx(row0 + r, q[k], nblk)]) > pb.ubf[r]);
            }
            pneed |= bit[k] & (above - 1u);   // above == 0: keep the block
        }
    }
    if (!live) pneed = 0;
    unsigned int m2 = __reduce_or_sync(0xffffffffu, pneed);
    // ---- 4. exhaustive scan of the surviving blocks, ascending, strict '>' -----------------------------------
    auto candidates = [&](int q, double (&v)[TB][BK]) {
        double a[BK];
#pragma unroll
        for (int jj = 0; jj < BK; ++jj) a[jj] = __dadd_rn(s, cs_l[(size_t)(q * BK + jj) * Kp]);   // HelpFunctions.jl:67
#pragma unroll
        for (int r = 0; r < TB; ++r) {
            const double2 w0 = *reinterpret_cast<const double2 *>(Prow + (size_t)r * Kp + q * BK);
            const double2 w1 = *reinterpret_cast<const double2 *>(Prow + (size_t)r * Kp + q * BK + 2);
            v[r][0] = __dadd_rn(a[0], w0.x);                                                        // :71
            v[r][1] = __dadd_rn(a[1], w0.y);
            v[r][2] = __dadd_rn(a[2], w1.x);
            v[r][3] = __dadd_rn(a[3], w1.y);
        }
    };
    auto relax = [&](const double (&v)[TB][BK], int q) {
#pragma unroll
        for (int jj = 0; jj < BK; ++jj)
#pragma unroll
            for (int r = 0; r < TB; ++r)
                if (best[r][0] > v[r][jj]) { BB_KEEP_BRANCH; best[r][0] = v[r][jj]; arg[r][0] = jbase + q * BK + jj; }  // :73-76
    };
    // (a per-block tournament that keeps only one compare -> move step on the chain through `best` was measured: 50 % more
    // moves, and with four warps per scheduler the chain latency is hidden anyway -- slower)
    if (m2) {
        double va[TB][BK], vb[TB][BK];   // two blocks in flight, alternating roles: no register copies between trips
        int qa = __ffs(m2) - 1, qb = 0;
        m2 &= m2 - 1;
        candidates(qa, va);
        for (;;) {
            if (m2) { qb = __ffs(m2) - 1; candidates(qb, vb); }
            relax(va, qa);
            executed += 1;
            if (!m2) break;
            m2 &= m2 - 1;
            if (m2) { qa = __ffs(m2) - 1; candidates(qa, va); }
            relax(vb, qb);
            executed += 1;
            if (!m2) break;
            m2 &= m2 - 1;
        }
    }
    return executed;
}

}  // namespace bb200
