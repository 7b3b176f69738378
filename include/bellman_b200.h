/*
 * bellman_b200.h -- C ABI of the B200-native trust-region subproblem solver.
 *
 * Drop-in boundary for the reference's two hot-path functions (paths relative to the
 * reference repository Jonas477/mixed-integer-optimal-control---algorithm-tools):
 *
 *     bellman_TRM!(∇f, u_old, B, β, p, Δt, nu, U, Φ, iterator)     HelpFunctions.jl:20-83
 *     eval_u_TRM!(u, u_old, U, Φ, B, nu)                            HelpFunctions.jl:98-124
 *
 * both called only from TRM (multi-trust.jl:108-114).  The reference has no FFI of its own; a
 * Julia file loaded after `include("multi-trust.jl")` re-defines the two methods and forwards to
 * this library through `ccall` (see INTEGRATION.md and julia/BellmanB200.jl).
 *
 * Conventions
 *   - every entry point returns int: 0 = BB200_OK, otherwise an error code; the message of the
 *     last failure on the calling thread is available from bb200_last_error().  Nothing throws,
 *     aborts or exits across this boundary.
 *   - all pointers are HOST pointers unless the parameter name starts with `d_`.
 *   - host arrays use the reference's (Julia, column-major) memory layout:
 *         df, u_old, u : double[n][M]      element (m,i) at (i-1)*M + (m-1)        (Julia M x n)
 *         Phi          : double[2][G][B+1]  Julia Float64[B+1, L1..LM, 2]           (multi-trust.jl:77)
 *         U            : int64 [n-1][G][B+1][M]  Julia Int64[M, B+1, L1..LM, n-1]   (multi-trust.jl:71-76)
 *     with G = prod(grid_dims) and grid cells addressed by their column-major offset.
 *   - the caller owns all host memory; the plan owns all device memory.
 *   - every entry point selects the plan's device itself and holds a per-plan mutex, so a plan may
 *     be driven from any OS thread (Julia tasks migrate); distinct plans may run concurrently.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef BELLMAN_B200_H
#define BELLMAN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BB200_OK 0
#define BB200_ERR_ARG 1      /* invalid argument / shape                                           */
#define BB200_ERR_CUDA 2     /* CUDA runtime failure (message has the CUDA error string)           */
#define BB200_ERR_INEXACT 3  /* u_old not integer valued / not finite: Julia's InexactError,       */
                             /* HelpFunctions.jl:37,57 (convert(Int64, abs(numl - u_old[m,i])));   */
                             /* control levels are integers by type (the reference's nu is         */
                             /* Vector{Vector{Int64}}, multi-trust.jl:64): level_values is int32   */
#define BB200_ERR_STALE 4    /* selection/backtrack reached a cell the DP never wrote (the          */
                             /* reference would read stale U there)                     */
#define BB200_ERR_STATE 5    /* call order violated (e.g. backtrack before any DP)                 */
#define BB200_ERR_NOMEM 6    /* device or host allocation failed                                   */

/* plan flags */
#define BB200_FLAG_STAGE_KERNELS 1u /* force the one-launch-per-stage kernels (validation path)    */
#define BB200_FLAG_NO_GRAPH 2u      /* do not capture TR iterations into a CUDA graph              */
#define BB200_FLAG_FORCE_WAVEFRONT 4u /* never use the single-CTA small-problem kernel (tests)       */

typedef struct bb200_plan bb200_plan;
typedef struct bb200_comm bb200_comm;   /* one rank of an NCCL communicator (one process per GPU)  */
typedef struct bb200_multi bb200_multi; /* all GPUs of one process: one plan + host thread per GPU */

const char *bb200_last_error(void);
/* Library/ABI version (major*1000+minor). */
int bb200_version(void);
/* Number of visible CUDA devices (0 when there is none; never fails). */
int bb200_device_count(void);

/*
 * One plan per TRM run (replaces the table allocation at multi-trust.jl:69-77).
 *   n, M            size(u_old) = (M, n)                                  HelpFunctions.jl:23
 *   K               number of admissible level tuples, in iterator order  AdmissibleIterators.jl:9-34
 *   B               budget floor(Δ⁰/Δt)                                   multi-trust.jl:69
 *   grid_dims[M]    L_m = length(nu[m])
 *   level_values    int32[K][M], nu[m][l_k[m]]
 *   grid_offset     int64[K], 0-based column-major offset of tuple k in the L1 x .. x LM grid
 *   jump_cost       double[K][K], jump_cost[j*K + l] = β * (Σ_m |ν_j[m]-ν_l[m]|^p)^(1/p), evaluated by
 *                   the CALLER with the reference's own expression (HelpFunctions.jl:63-67); j is the
 *                   level at stage i+1, l the level at stage i
 *   dt              Δt
 *   batch           number of resident subproblem slots (>= 1); each slot has its own df, u_old, u,
 *                   value rows and argmin table
 */
int bb200_plan_create(int device, int64_t n, int32_t M, int32_t K, int64_t B,
                      const int64_t *grid_dims, const int32_t *level_values,
                      const int64_t *grid_offset, const double *jump_cost, double dt,
                      int32_t batch, uint32_t flags, bb200_plan **out);
int bb200_plan_destroy(bb200_plan *plan);

/* Run subsequent work of this plan on the given cudaStream_t (e.g. torch's current stream)
 * instead of the plan's own stream.  NULL restores the plan's stream. */
int bb200_plan_set_stream(bb200_plan *plan, void *cuda_stream);

/*
 * bellman_TRM!  (HelpFunctions.jl:20-83) for slot 0: copies df and u_old in, runs the DP for budget
 * B and returns when the value rows and the packed argmin table are resident and consumable.
 */
int bb200_bellman(bb200_plan *plan, const double *df, const double *u_old);

/*
 * eval_u_TRM!  (HelpFunctions.jl:98-124) for slot 0 against the resident table; B_new <= B; may be
 * called repeatedly with shrinking B_new without re-running the DP (multi-trust.jl:105-114).
 *   u_out           double[n][M]
 *   phi_star        value of the selected stage-1 cell           (may be NULL)
 *   b_star          its budget row                               (may be NULL)
 *   k_star          its admissible index, 0-based                (may be NULL)
 */
int bb200_select_and_backtrack(bb200_plan *plan, int64_t B_new, double *u_out, double *phi_star,
                               int64_t *b_star, int64_t *k_star);

/* One whole TR inner iteration for slot 0: H2D(df,u_old) -> DP -> selection -> backtrack -> D2H(u),
 * replayed from a captured CUDA graph unless BB200_FLAG_NO_GRAPH.  Same results as bb200_bellman
 * followed by bb200_select_and_backtrack(B_new). */
int bb200_solve(bb200_plan *plan, const double *df, const double *u_old, int64_t B_new,
                double *u_out, double *phi_star, int64_t *b_star, int64_t *k_star);

/* Batched variants: S independent subproblems (own df, u_old) sharing the plan's tables, processed in
 * waves of `batch` slots.  df_all/u_old_all: double[S][n][M].  For every subproblem, n_radii (<= 16)
 * selections/backtracks are taken from its one table (a radius sweep costs one DP, multi-trust.jl:109-110):
 *   u_out_all       double[S][n_radii][n][M]   (may be NULL: only the optima are returned)
 *   phi_star/b_star/k_star   [S][n_radii]       (each may be NULL)
 *   status          int32[S][n_radii]: BB200_OK, BB200_ERR_INEXACT (that subproblem's u_old) or BB200_ERR_STALE
 *                   (that selection), per entry; may be NULL.  The return value is the worst of them (the
 *                   other entries are valid) or a hard error.
 * Pipeline: the DP of a wave is one persistent launch; its selections and backtracks are ONE launch each with
 * a CTA per (slot, radius); inputs of wave w+1 and outputs of wave w-1 move on a second stream / through
 * pinned staging while wave w computes; the host waits once per wave.
 */
int bb200_solve_batched(bb200_plan *plan, int64_t S, const double *df_all, const double *u_old_all,
                        int32_t n_radii, const int64_t *B_new, double *u_out_all, double *phi_star,
                        int64_t *b_star, int64_t *k_star, int32_t *status);
/* The same for the shard {first, first+stride, ...} < S of the caller's arrays (indices stay global): what one
 * GPU of several does (subproblem s -> rank s mod G). */
int bb200_solve_batched_shard(bb200_plan *plan, int64_t S, int64_t first, int64_t stride, const double *df_all,
                              const double *u_old_all, int32_t n_radii, const int64_t *B_new, double *u_out_all,
                              double *phi_star, int64_t *b_star, int64_t *k_star, int32_t *status);

/* Deterministic best-candidate reduction over gathered (value, global index) records in Julia's findmin
 * order, like the selection (S8): the smallest value wins, NaN precedes every number, -0.0 precedes +0.0, equal
 * values go to the smallest index.  Used after the per-rank records were exchanged (ncclAllGather). */
int bb200_best_candidate(const double *values, const int64_t *indices, int64_t count,
                         double *best_value, int64_t *best_index);

/* ---- multi-GPU ------------------------------------------------------------------------
 * One DP is sequential in time, so one subproblem lives on one GPU; S independent subproblems (multi-start
 * x0, each with its radius sweep) are sharded s -> rank s mod G with no data-path collective.  The one
 * exchange is the best-candidate reduction: a 16-byte (value, global index) record per rank through
 * ncclAllGather over NVLink/NVSwitch + bb200_best_candidate on every rank (NCCL has no MINLOC).  NCCL is
 * loaded at run time (libnccl.so.2, or $BELLMAN_B200_NCCL); without it these entry points return
 * BB200_ERR_STATE.  bb200_nccl_version() is 0 then.
 *
 * (a) all GPUs of ONE process -- what a Julia session uses: one ccall drives every GPU of the box. */
int bb200_nccl_version(void);
int bb200_multi_create(const int32_t *devices, int32_t n_dev, int64_t n, int32_t M, int32_t K, int64_t B,
                       const int64_t *grid_dims, const int32_t *level_values, const int64_t *grid_offset,
                       const double *jump_cost, double dt, int32_t batch_per_device, uint32_t flags,
                       bb200_multi **out);
int bb200_multi_destroy(bb200_multi *m);
/* Arguments as bb200_solve_batched (all arrays global, [S]...), plus the reduction's result:
 *   best_value, best_subproblem, best_radius   smallest selected value over all entries with status OK (Julia
 *                   findmin order; ties -> smallest (subproblem, radius)); -1 indices if there is none
 *   u_best          double[n][M], the winner's trajectory; taken from u_out_all or, if that is NULL, recomputed
 *                   by one more DP of the winning subproblem on its device.  Each may be NULL. */
int bb200_multi_solve_batched(bb200_multi *m, int64_t S, const double *df_all, const double *u_old_all,
                              int32_t n_radii, const int64_t *B_new, double *u_out_all, double *phi_star,
                              int64_t *b_star, int64_t *k_star, int32_t *status, double *best_value,
                              int64_t *best_subproblem, int32_t *best_radius, double *u_best);
/* out[0] = devices, out[1] = host wall time [ms] of the last call, out[2+d] = device time [ms] of device d's shard */
int bb200_multi_stats(bb200_multi *m, double *out, int32_t count);

/* (b) one process per GPU (torchrun, MPI, Julia Distributed): rank 0 draws the 128-byte NCCL unique id, the host
 * program ships it to the other ranks by whatever channel it has, every rank creates its communicator. */
int bb200_comm_unique_id(void *id128);
int bb200_comm_create(int device, int32_t nranks, int32_t rank, const void *id128, bb200_comm **out);
int bb200_comm_destroy(bb200_comm *comm);
/* Collective: every rank passes its local best (value, global index < 2^53); every rank receives the global
 * best and the rank that contributed it. */
int bb200_comm_best_candidate(bb200_comm *comm, double value, int64_t index, double *best_value,
                              int64_t *best_index, int32_t *owner_rank);
/* Collective: `count` doubles from rank `root`'s host_buf to every rank's host_buf (the winner's u, n*M*8
 * bytes) by ncclBroadcast. */
int bb200_comm_broadcast(bb200_comm *comm, int32_t root, double *host_buf, int64_t count);

/* ---- resident (device-side) interface: what bench.py times with inputs already in HBM -------- */
/* H2D copy of one subproblem's inputs into slot `slot` (asynchronous on the plan's stream). */
int bb200_upload(bb200_plan *plan, int32_t slot, const double *df, const double *u_old);
/* Same, from DEVICE pointers (e.g. torch tensors): device-to-device on the plan's stream. */
int bb200_upload_device(bb200_plan *plan, int32_t slot, const double *d_df, const double *d_u_old);
/* Launch the DP for slots [slot0, slot0+count) on resident inputs; asynchronous. */
int bb200_bellman_resident(bb200_plan *plan, int32_t slot0, int32_t count);
/* Selection + backtrack for one slot on the device; u stays resident; asynchronous. */
int bb200_backtrack_resident(bb200_plan *plan, int32_t slot, int64_t B_new);
/* D2H of slot's u and optimum record; synchronises the plan's stream. */
int bb200_download(bb200_plan *plan, int32_t slot, double *u_out, double *phi_star, int64_t *b_star,
                   int64_t *k_star);
/* Block until everything queued on the plan's stream is done; reports deferred device-side errors
 * (BB200_ERR_INEXACT / BB200_ERR_STALE). */
int bb200_sync(bb200_plan *plan);

/* ---- parity / inspection --------------------------------------------------------------------- */
/* Reference-shaped value table (S7): Phi_out double[2][G][B+1]; +Inf in inadmissible grid cells. */
int bb200_export_phi(bb200_plan *plan, int32_t slot, double *Phi_out);
/* Reference-shaped argmin table for stages [i0, i1) (1-based, i1 <= n): U_out int64[i1-i0][G][B+1][M]
 * holding 1-based index tuples; `fill` in cells the DP did not write (the reference leaves those
 * stale). */
int bb200_export_argmin(bb200_plan *plan, int32_t slot, int64_t i0, int64_t i1, int64_t *U_out,
                        int64_t fill);
/* Exact number of executions of the innermost loop body (HelpFunctions.jl:71-76) of the last DP of
 * `slot`: N = Σ_i K·Σ_l max(0, B+1-b~_l(i)). */
int bb200_count_updates(bb200_plan *plan, int32_t slot, int64_t *n_updates);

/* Device-side "next" rows: evaluated on the resident u / u_old / df of `slot`.
 *   pred integral Δt·Σ_j ∇f[:,j]'(u_old[:,j]-u[:,j])  (multi-trust.jl:117-121); the sum runs left to right over j and,
 *       inside a column, over m -- the reference's `∇f[:,j]' * (..)` is a BLAS dot whose internal order is not
 *       specified, so for M > 2 the last bit of `pred` is only pinned against this order
 *   TV_p(u, p)  with p = +Inf, 1 or 2                    (HelpFunctions.jl:251-268)
 *       p = 1 and p = Inf are bit-exact (integer-valued u: every term is exact).  p = 2 is NOT bit-pinned: the
 *       reference evaluates (Σ|Δ|^2)^(1/p) with Julia's Float64 `^`, this kernel uses sqrt -- identical whenever the
 *       sum is a perfect square, otherwise possibly one ulp apart.  Callers who need the reference's bits for p = 2
 *       keep TV_p in Julia (it is O(M·n)). */
int bb200_pred_integral(bb200_plan *plan, int32_t slot, double *int_val);
int bb200_tv(bb200_plan *plan, int32_t slot, double p, double *tv);

/* Statistics of the plan.  Index meaning (values beyond `count` are not written):
 *   0 last DP device time [ms] (CUDA events on the launching stream, all slots of the launch)
 *   1 last selection+backtrack device time [ms]
 *   2 kernels launched by this plan so far
 *   3 kernel path of the last DP: 0 = per-stage kernels, 1 = persistent wavefront kernel,
 *     2 = single-CTA small-problem kernel (value rows in shared memory, one CTA per subproblem)
 *   4 CTAs used by the last DP launch          5 source rows per CTA
 *   6 argmin bytes per cell (1 or 2)           7 device bytes owned by the plan
 *   8 threads per CTA of the last DP launch    9 j-split of the last DP launch
 *  10 device time [ms] of the last persistent wavefront kernel alone (events around that launch)
 *  11 number of CUDA-graph replays performed by bb200_solve
 *  12 tile variant of the wavefront kernel (1-based, as accepted by bb200_plan_tune)
 *  13 scatter warps per CTA of the wavefront kernel
 *  14 device time [ms] of the last bb200_solve_batched call (first prep kernel to last D2H)
 *  15 waves of that call            16 host waits (event synchronisations) of all batched calls so far
 *  17 candidates the last DP really evaluated when it ran the pruned scan (else 0): the exhaustive count is
 *     bb200_count_updates; the difference was skipped by the bound test, results are bit-identical
 *  18 block size of the pruned scan of the current geometry (0: exhaustive scan); also 4 while a wide level set runs the
 *     per-stage kernels with the pruned scan
 *  19 1 if the plan went back to the exhaustive tiles because its DP evaluated more candidates than the measured
 *     break-even of the pruned scan (25 %): a data / horizon dependent choice, results are identical either way
 *  20 CTAs that own (5) source rows; the CTAs above them own (21) rows (two-zone slices; 20 == 4: uniform slices)
 */
int bb200_stats(bb200_plan *plan, double *out, int32_t count);

/* FP64-pipe issue-rate microbenchmark on `device` (the roofline denominator; MEASURED_PEAKS.json has no
 * FP64 figure).  mode 0: independent DADD chains; mode 1: DADD + DSETP + selects (the relaxation's mix,
 * counted as 2 FP64 ops).  Runs a full-chip kernel for about target_ms; returns lane-operations per second. */
int bb200_fp64_peak(int device, int32_t mode, double target_ms, double *ops_per_s, double *elapsed_ms);

/* In-kernel cycle profile of the wavefront kernel.  enable != 0 switches the counters on for later launches;
 * if out != NULL the counters of the last launch are copied out, 16 int64 per CTA for min(max_ctas, SMs) CTAs:
 *   [0..4]  compute warp 0: cycles waiting for finished rows / cost rows, in phase B, handing over; [3] unused; stages
 *   [5..7]  scatter warp 0: cycles waiting for the scan, waiting for halo rows / ring space, in phase C
 *   [8..13] comm warp: event-loop trips, idle trips, predecessor polls, successor polls, unused, cost rows loaded */
int bb200_profile(bb200_plan *plan, int32_t enable, int64_t *out, int32_t max_ctas);

/* Tuning knobs for experiments (0 = automatic): number of CTAs, j-split, tile variant (1-based index into the
 * wavefront kernel's tile table; add 100 * NS to force NS scatter warps per CTA). */
int bb200_plan_tune(bb200_plan *plan, int32_t ctas, int32_t jsplit, int32_t variant);

/* The geometry the persistent wavefront kernel would pick for a table shape on a GPU with `num_sms` SMs and
 * `smem_max` bytes of opt-in shared memory per CTA.  Pure host code (no CUDA call): it lets CPU-only tests pin the
 * geometry model.  ctas / jsplit / variant as in bb200_plan_tune.  out (values beyond `count` are not written):
 *   0 1 if the shape runs on the wavefront kernel, else 0      1 tile variant (1-based)
 *   2 rows per thread tile in sub-slice A   3 in sub-slice B (0: one sub-slice)   4 levels per thread tile
 *   5 CTAs   6 source rows per CTA   7 j-split   8 successors per j-group   9 rows of the padded jump-cost table
 *  10 scatter warps   11 threads per CTA   12 dynamic shared memory per CTA [bytes]
 *  13 block size of the pruned scan (0: exhaustive scan)
 *  14 CTAs that own (6) rows; the CTAs above them own (15) rows (two-zone slices; 14 == 5: uniform slices) */
int bb200_wave_geometry(int64_t n, int32_t M, int32_t K, int64_t B, int32_t num_sms, int64_t smem_max, int32_t ctas,
                        int32_t jsplit, int32_t variant, int64_t *out, int32_t count);

#ifdef __cplusplus
}
#endif
#endif /* BELLMAN_B200_H */
