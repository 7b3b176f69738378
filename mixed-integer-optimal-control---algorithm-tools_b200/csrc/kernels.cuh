// kernels.cuh -- launcher prototypes shared between the kernel translation units and the C ABI.
#pragma once
#include "bb200_internal.cuh"

namespace bb200 {

// kernels_common.cu
void launch_prep(const Tables &t, const SlotDev &slot, int *err, int *btmax, cudaStream_t st);
int launch_stage_path(const Tables &t, const SlotDev &slot, int argw, cudaStream_t st);
int launch_terminal_stage(const Tables &t, const SlotDev &slot, cudaStream_t st);  // the terminal stage alone (1 launch; -1: K > 1024)

// kernel_stage_pruned.cu -- per-stage kernels with the branch-and-bound scan, for level sets the pipelined kernel does not take
struct StagePrunedTabs {
    double *cpad;   // [Kp][Kp] jump costs, +Inf rows K .. Kp-1
    float *cminf;   // [nblk][Kp] block minima of the jump costs, rounded down
    float *cwv;     // [Kp/32][nblk] per level block: smallest cminf over its live levels
    float *cmx;     // [Kp] largest finite |jump cost| into a level, rounded up
    int nblk;       // blocks of 4 successors = Kp / 4
};
bool stage_pruned_applicable(const Tables &t);
size_t stage_pruned_table_bytes(const Tables &t);
cudaError_t stage_pruned_setup(const Tables &t, void *base, StagePrunedTabs &pt, cudaStream_t st);
int launch_stage_pruned_path(const Tables &t, const SlotDev &slot, int argw, const StagePrunedTabs &pt, unsigned long long *exec,
                             cudaStream_t st);
bool mini_applicable(const Tables &t, size_t smem_max);
cudaError_t launch_mini(const Tables &t, const SlotDev *d_slots, int count, int argw, cudaStream_t st);
// sweep == nullptr: one CTA for `slot`; otherwise `ctas` = slots * radii CTAs, one per (slot, radius) of the sweep
void launch_select(const Tables &t, const SlotDev &slot, int Bnew, const int *bnew_ptr, int *err, cudaStream_t st,
                   const SweepDev *sweep = nullptr, int ctas = 1);
void launch_backtrack(const Tables &t, const SlotDev &slot, int argw, int *err, cudaStream_t st,
                      const SweepDev *sweep = nullptr, int ctas = 1);
void launch_pred_integral(const Tables &t, const SlotDev &slot, double *out, cudaStream_t st);
void launch_tv(const Tables &t, const SlotDev &slot, int mode, double *out, cudaStream_t st);

// kernel_wavefront.cu -- the persistent pipelined DP
struct WaveCfg {
    int variant;  // index into the tile instantiation table
    int TB, TBB;  // source rows per thread tile in sub-slice A / sub-slice B (TBB = 0: one sub-slice)
    int TL;       // levels per thread tile
    int G;        // CTAs (each owns R consecutive source budget rows)
    int RG;       // row groups per sub-slice
    int R;        // RA + RB
    int GA, Rtop; // two-zone slices (pruned tiles): CTAs 0 .. GA-1 own R source rows, the CTAs above them Rtop; GA = G: uniform
    int RA, RB;   // rows of sub-slice A (lower) and B (upper): RG * TB, RG * TBB
    int nLG;      // level groups = ceil(K / TL)
    int JS;       // j-split: thread groups scanning disjoint successor ranges
    int PR;       // pruned scan: successors per block of the branch-and-bound scan (0 = off), see kernel_wavefront.cu
    int jper;     // successors per group = ceil(K / JS), rounded up to even (to 8 when the pad rows fit)
    int Kr;       // rows of the jump-cost table in shared memory: K, or JS * jper with +Inf pad rows
    int tpg;      // threads per group (multiple of 32)
    int NS;       // scatter warps (phase C); 0 when the tile has one sub-slice and the compute warps finish it
    int NF;       // warps that finish a stage: NS, or the compute warps
    int EC;       // cells per lane in phase C (2 or 4): a work unit is 32 * EC levels of one row
    int threads;  // JS * tpg compute threads + comm warp + publisher warp + NS scatter warps
    size_t smem;  // dynamic shared memory bytes
    int nsub;     // subproblems walked by this launch
    const SlotDev *slots;  // device array [nsub]
    double *halo;                // [kHaloRing][B1][Kp]
    unsigned long long *flags;   // [G * kFlagStride]
    int *err;                    // [4]: inexact, stale, watchdog, abort
    const int *btmax;
    long long *prof;             // optional [G][16] cycle counters (NULL = off)
    unsigned long long *exec;    // pruned scan: number of candidates actually evaluated by this launch (all CTAs)
    int decouple;                // timing experiments only: CTAs ignore their neighbours (wrong results)
    long long wd_cycles;         // watchdog of the inter-CTA waits in SM cycles (0 = off); $BELLMAN_B200_WATCHDOG_S, default 3 s
};

// Fills the geometry fields of cfg for the given tables; returns false when the shape cannot run on the
// wavefront kernel (e.g. the jump-cost table does not fit in shared memory).  want_variant = 0 picks the tile by
// the cost model, v > 0 forces tile v - 1; adding 100 * NS forces the number of scatter warps.
bool wave_configure(const Tables &t, int argw, int num_sms, size_t smem_max, int want_ctas,
                    int want_js, int want_variant, WaveCfg &cfg);
cudaError_t launch_wavefront(const Tables &t, const WaveCfg &cfg, int argw, cudaStream_t st);

// kernels_microbench.cu
cudaError_t measure_fp64_rate(int mode, int num_sms, double target_ms, double *ops_per_s, double *elapsed_ms,
                              cudaStream_t st);

}  // namespace bb200
