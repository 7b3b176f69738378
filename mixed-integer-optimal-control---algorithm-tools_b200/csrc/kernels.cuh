// kernels.cuh -- launcher prototypes shared between the kernel translation units and the C ABI.
#pragma once
#include "bb200_internal.cuh"

namespace bb200 {

// kernels_common.cu
void launch_prep(const Tables &t, const SlotDev &slot, int *err, int *btmax, cudaStream_t st);
int launch_stage_path(const Tables &t, const SlotDev &slot, int argw, cudaStream_t st);
bool mini_applicable(const Tables &t, size_t smem_max);
cudaError_t launch_mini(const Tables &t, const SlotDev *d_slots, int count, int argw, cudaStream_t st);
void launch_select(const Tables &t, const SlotDev &slot, int Bnew, const int *bnew_ptr, int *err, cudaStream_t st);
void launch_backtrack(const Tables &t, const SlotDev &slot, int argw, int *err, cudaStream_t st);
void launch_pred_integral(const Tables &t, const SlotDev &slot, double *out, cudaStream_t st);
void launch_tv(const Tables &t, const SlotDev &slot, int mode, double *out, cudaStream_t st);

// kernel_wavefront.cu -- the persistent pipelined DP
struct WaveCfg {
    int variant;  // index into the (TB, TL) instantiation table
    int TB, TL;   // source rows / levels per thread tile
    int G;        // CTAs (each owns R consecutive source budget rows)
    int RG;       // row groups per CTA
    int R;        // RG * TB
    int nLG;      // level groups = ceil(K / TL)
    int JS;       // j-split: thread groups scanning disjoint successor ranges
    int jper;     // successors per group = ceil(K / JS)
    int tpg;      // threads per group (multiple of 32)
    int RP;       // padded row positions per level in the shared value rows
    int pub;      // 1: a separate publisher warp moves the progress counter (stages too short to hide the fence)
    int threads;  // JS * tpg
    size_t smem;  // dynamic shared memory bytes
    int nsub;     // subproblems walked by this launch
    const SlotDev *slots;  // device array [nsub]
    double *halo;                // [kHaloRing][B1][Kp]
    unsigned long long *flags;   // [G * kFlagStride]
    int *err;                    // [4]: inexact, stale, watchdog, abort
    const int *btmax;
    long long *prof;             // optional [G][16] cycle counters (NULL = off)
};

// Fills the geometry fields of cfg for the given tables; returns false when the shape cannot run on the
// wavefront kernel (e.g. the jump-cost table does not fit in shared memory).
bool wave_configure(const Tables &t, int argw, int num_sms, size_t smem_max, int want_ctas,
                    int want_js, int want_variant, WaveCfg &cfg);
cudaError_t launch_wavefront(const Tables &t, const WaveCfg &cfg, int argw, cudaStream_t st);

// kernels_microbench.cu
cudaError_t measure_fp64_rate(int mode, int num_sms, double target_ms, double *ops_per_s, double *elapsed_ms,
                              cudaStream_t st);

}  // namespace bb200
