// bb200_api.cu -- the C ABI (include/bellman_b200.h): plans, resident buffers, launch sequencing.
//
// Replaces, behind plain C entry points, the reference's
//   table allocation        multi-trust.jl:69-77        -> bb200_plan_create
//   bellman_TRM!            HelpFunctions.jl:20-83      -> bb200_bellman / bb200_bellman_resident
//   eval_u_TRM!             HelpFunctions.jl:98-124     -> bb200_select_and_backtrack
// There is no CPU implementation in this library: every compute entry point needs a CUDA device.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/bellman_b200.h"
#include "bb200_internal.cuh"
#include "kernels.cuh"
#include <cstdlib>

using namespace bb200;

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                               \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(e_ == cudaErrorMemoryAllocation ? BB200_ERR_NOMEM : BB200_ERR_CUDA,    \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct SlotHost {
    double *df = nullptr, *u_old = nullptr, *u = nullptr, *phi = nullptr, *rec = nullptr, *ss_all = nullptr;
    int *bt_all = nullptr;
    void *arg = nullptr;
    unsigned long long *n_updates = nullptr;
    bool has_dp = false;
};

}  // namespace

struct bb200_plan {
    int device = 0;
    int64_t n = 0, B = 0, G = 1;
    int M = 0, K = 0, Kp = 0, B1 = 0, batch = 1, argw = 1;
    int64_t nPad = 0;
    double dt = 0.;
    uint32_t flags = 0;
    std::vector<int64_t> grid_dims, grid_offset;
    std::vector<int32_t> level_values;
    Tables tab{};
    // device
    double *d_lvd = nullptr, *d_cost = nullptr, *d_halo = nullptr, *d_scalar = nullptr;
    long long *d_goff = nullptr;
    size_t halo_elems = 0;
    unsigned long long *d_flags = nullptr;
    int *d_err = nullptr, *d_btmax = nullptr;
    unsigned long long *d_exec = nullptr;  // pruned scan: candidates really evaluated by the last DP
    unsigned long long *h_exec = nullptr;  // pinned copy, fetched with the error words
    int last_dp_slots = 1;                 // slots of the last DP launch (the counter above sums over them)
    bool prune_off = false;                // the pruned scan did not pay on this plan's data: exhaustive tiles from now on
    // per-stage kernels with the pruned scan (level sets the pipelined kernel does not take)
    void *d_sp = nullptr;
    StagePrunedTabs sp{};
    bool sp_ok = false, sp_off = false;
    double prune_switches = 0.;
    long long *d_prof = nullptr;
    bool prof_on = false;
    SlotDev *d_slots = nullptr;
    std::vector<SlotHost> slots;
    std::vector<SlotDev> slots_dev;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double *d_rec_all = nullptr;  // [batch][kRecDoubles]: the slots' record blocks, contiguous (one D2H per wave)
    // batched radius sweep (bb200_solve_batched): trial budgets, per-(slot, radius) trajectories, second stream,
    // double-buffered pinned staging so that the H2D of wave w+1 and the D2H of wave w-1 overlap the DP of wave w
    int *d_radii = nullptr;
    double *d_usweep = nullptr;
    size_t usweep_elems = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_prep = nullptr, ev_in[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr}, ev_batch[2] = {nullptr, nullptr};
    double *h_in[2] = {nullptr, nullptr}, *h_out[2] = {nullptr, nullptr}, *h_recs[2] = {nullptr, nullptr};
    int *h_errs[2] = {nullptr, nullptr};
    unsigned long long *h_execs[2] = {nullptr, nullptr};
    size_t h_in_elems = 0, h_out_elems = 0;
    double last_batch_ms = 0., batch_syncs = 0., batch_waves = 0.;
    // pinned staging
    double *h_rec = nullptr;
    int *h_err = nullptr;
    // CUDA-graph replay of one TR iteration (bb200_solve): pinned input/output staging + the captured graph
    double *h_df = nullptr, *h_uold = nullptr, *h_u = nullptr;
    int *h_bnew = nullptr, *d_bnew = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    bool graph_failed = false;
    double graph_replays = 0.;
    // geometry
    int num_sms = 0;
    size_t smem_max = 0;
    bool wave_ok = false, mini_ok = false;
    WaveCfg cfg{};
    int tune_ctas = 0, tune_js = 0, tune_variant = 0;
    // stats
    double last_dp_ms = 0., last_bt_ms = 0., last_wave_ms = 0.;
    double launches = 0.;
    int last_path = -1;
    size_t dev_bytes = 0;
    bool dp_timed = false, bt_timed = false;
    std::mutex mu;
};

namespace {

// Watchdog of the persistent kernel's inter-CTA waits: seconds from $BELLMAN_B200_WATCHDOG_S (default 3; 0 disables it
// for runs under a debugger / compute-sanitizer / on a time-sliced GPU), converted with the nominal 2 GHz clock.
long long watchdog_cycles()
{
    static const long long v = [] {
        const char *e = getenv("BELLMAN_B200_WATCHDOG_S");
        const double sec = e ? atof(e) : 3.0;
        return sec <= 0. ? 0LL : (long long)(sec * 2.0e9);
    }();
    return v;
}

int decouple_env()
{
    static const int v = [] { const char *e = getenv("BELLMAN_B200_DECOUPLE"); return e ? atoi(e) : 0; }();
    return v;
}

template <typename T>
int dev_alloc(bb200_plan *p, T **ptr, size_t count)
{
    size_t bytes = count * sizeof(T);
    if (bytes == 0) bytes = sizeof(T);
    cudaError_t e = cudaMalloc((void **)ptr, bytes);
    if (e != cudaSuccess)
        return fail(BB200_ERR_NOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    p->dev_bytes += bytes;
    return BB200_OK;
}

// The halo ring (a few tens of MB) is re-read every stage while the argmin table streams through L2 once: ask the
// driver to keep the ring's lines (persisting access-policy window on the plan's stream; best effort).
void pin_ring_in_l2(bb200_plan *p)
{
    if (!p->d_halo || !p->stream) return;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, p->device) != cudaSuccess) { cudaGetLastError(); return; }
    const size_t bytes = p->halo_elems * sizeof(double);
    const size_t win = bytes < (size_t)prop.accessPolicyMaxWindowSize ? bytes : (size_t)prop.accessPolicyMaxWindowSize;
    if (win == 0 || prop.persistingL2CacheMaxSize == 0) return;
    const size_t want = win < (size_t)prop.persistingL2CacheMaxSize ? win : (size_t)prop.persistingL2CacheMaxSize;
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
    cudaStreamAttrValue attr{};
    attr.accessPolicyWindow.base_ptr = p->d_halo;
    attr.accessPolicyWindow.num_bytes = win;
    attr.accessPolicyWindow.hitRatio = 1.0f;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cudaStreamSetAttribute(p->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
    cudaGetLastError();  // best effort: a refusal changes traffic, not results
}

int reconfigure(bb200_plan *p)
{
    p->wave_ok = false;
    p->mini_ok = false;
    p->sp_ok = false;
    if (p->flags & BB200_FLAG_STAGE_KERNELS) return BB200_OK;
    // small stages: one CTA per subproblem, rows in shared memory (unless a wavefront geometry was requested)
    if (!(p->tune_ctas || p->tune_js || p->tune_variant) && !(p->flags & BB200_FLAG_FORCE_WAVEFRONT) &&
        mini_applicable(p->tab, p->smem_max)) {
        p->mini_ok = true;
        return BB200_OK;
    }
    WaveCfg c{};
    const int want_variant = (p->prune_off && p->tune_variant == 0) ? -1 : p->tune_variant;
    if (wave_configure(p->tab, p->argw, p->num_sms, p->smem_max, p->tune_ctas, p->tune_js, want_variant, c)) {
        // the halo ring holds kHaloRing stages of value rows [B1][Kp], row-major like the rows in shared memory
        const size_t need = (size_t)kHaloRing * p->B1 * p->Kp;
        if (need > p->halo_elems) {
            cudaStreamSynchronize(p->stream);
            if (p->d_halo) { cudaFree(p->d_halo); p->dev_bytes -= p->halo_elems * sizeof(double); }
            p->d_halo = nullptr;
            p->halo_elems = 0;
            if (cudaMalloc((void **)&p->d_halo, need * sizeof(double)) != cudaSuccess) {
                cudaGetLastError();
                return fail(BB200_ERR_NOMEM, "cudaMalloc of the halo ring (%zu bytes) failed", need * sizeof(double));
            }
            p->halo_elems = need;
            p->dev_bytes += need * sizeof(double);
            // pad columns of the ring rows are never written by a kernel but travel with the rows (bulk TMA) and are read
            // by the pruned scan's block minima: give them a huge finite value (0x7f7f.. = 1.4e306) instead of whatever the
            // allocation held
            if (cudaMemset(p->d_halo, 0x7f, need * sizeof(double)) != cudaSuccess) {
                cudaGetLastError();
                return fail(BB200_ERR_CUDA, "cudaMemset of the halo ring failed");
            }
        }
        p->cfg = c;
        p->wave_ok = true;
        pin_ring_in_l2(p);
    } else if (!(p->tune_ctas || p->tune_js || p->tune_variant) && stage_pruned_applicable(p->tab)) {
        // one launch per stage, but with the branch-and-bound scan (kernel_stage_pruned.cu); its tables are built once
        if (!p->d_sp) {
            const size_t bytes = stage_pruned_table_bytes(p->tab);
            if (cudaMalloc(&p->d_sp, bytes) != cudaSuccess) {
                cudaGetLastError();
                p->d_sp = nullptr;
                return BB200_OK;  // no tables: the plain per-stage kernels still run
            }
            p->dev_bytes += bytes;
            if (stage_pruned_setup(p->tab, p->d_sp, p->sp, p->stream) != cudaSuccess) {
                cudaGetLastError();
                return BB200_OK;
            }
        }
        p->sp_ok = true;
    }
    return BB200_OK;
}

void destroy_plan(bb200_plan *p)
{
    if (!p) return;
    cudaSetDevice(p->device);
    if (p->own_stream) cudaStreamSynchronize(p->own_stream);
    if (p->copy_stream) cudaStreamSynchronize(p->copy_stream);
    for (auto &s : p->slots) {
        cudaFree(s.df); cudaFree(s.u_old); cudaFree(s.u); cudaFree(s.phi);
        cudaFree(s.arg); cudaFree(s.n_updates); cudaFree(s.ss_all); cudaFree(s.bt_all);
    }
    cudaFree(p->d_rec_all); cudaFree(p->d_radii); cudaFree(p->d_usweep);
    for (int k = 0; k < 2; ++k) {
        if (p->h_in[k]) cudaFreeHost(p->h_in[k]);
        if (p->h_out[k]) cudaFreeHost(p->h_out[k]);
        if (p->h_recs[k]) cudaFreeHost(p->h_recs[k]);
        if (p->h_errs[k]) cudaFreeHost(p->h_errs[k]);
        if (p->h_execs[k]) cudaFreeHost(p->h_execs[k]);
        if (p->ev_in[k]) cudaEventDestroy(p->ev_in[k]);
        if (p->ev_out[k]) cudaEventDestroy(p->ev_out[k]);
        if (p->ev_batch[k]) cudaEventDestroy(p->ev_batch[k]);
    }
    if (p->ev_prep) cudaEventDestroy(p->ev_prep);
    if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
    cudaFree(p->d_lvd); cudaFree(p->d_cost); cudaFree(p->d_halo); cudaFree(p->d_scalar);
    if (p->d_sp) cudaFree(p->d_sp);
    cudaFree(p->d_goff); cudaFree(p->d_flags); cudaFree(p->d_err); cudaFree(p->d_btmax); cudaFree(p->d_exec);
    cudaFree(p->d_slots);
    cudaFree(p->d_prof);
    if (p->h_rec) cudaFreeHost(p->h_rec);
    if (p->h_err) cudaFreeHost(p->h_err);
    if (p->h_exec) cudaFreeHost(p->h_exec);
    if (p->h_df) cudaFreeHost(p->h_df);
    if (p->h_uold) cudaFreeHost(p->h_uold);
    if (p->h_u) cudaFreeHost(p->h_u);
    if (p->h_bnew) cudaFreeHost(p->h_bnew);
    cudaFree(p->d_bnew);
    if (p->graph_exec) cudaGraphExecDestroy(p->graph_exec);
    for (auto &e : p->ev) if (e) cudaEventDestroy(e);
    if (p->own_stream) cudaStreamDestroy(p->own_stream);
    delete p;
}

// Queue the DP for slots [slot0, slot0+count) on the plan's stream.
int queue_dp(bb200_plan *p, int slot0, int count, bool capturing = false, cudaEvent_t after_prep = nullptr)
{
    cudaStream_t st = p->stream;
    if (!capturing) CU(cudaEventRecord(p->ev[0], st));
    CU(cudaMemsetAsync(p->d_err, 0, 4 * sizeof(int), st));
    CU(cudaMemsetAsync(p->d_btmax, 0, sizeof(int), st));
    CU(cudaMemsetAsync(p->d_exec, 0, sizeof(unsigned long long), st));
    CU(cudaMemsetAsync(p->d_rec_all + (size_t)slot0 * kRecDoubles, 0, (size_t)count * kRecDoubles * sizeof(double), st));
    p->last_dp_slots = count;
    for (int s = slot0; s < slot0 + count; ++s) {
        CU(cudaMemsetAsync(p->slots[s].n_updates, 0, sizeof(unsigned long long), st));
        launch_prep(p->tab, p->slots_dev[s], p->d_err, p->d_btmax, st);
        p->launches += 1;
        p->slots[s].has_dp = true;
    }
    if (after_prep) CU(cudaEventRecord(after_prep, st));  // df / u_old of these slots may be overwritten from here on
    if (p->mini_ok) {
        if (!capturing) CU(cudaEventRecord(p->ev[4], st));
        CU(launch_mini(p->tab, p->d_slots + slot0, count, p->argw, st));
        if (!capturing) CU(cudaEventRecord(p->ev[5], st));
        p->launches += 1;
        p->last_path = 2;
    } else if (p->wave_ok) {
        // the kernel counts global steps (slots * n) in 32 bits: very long batches go out in several launches
        const long long per_launch = ((((long long)1 << 30) - 1) / p->n) > 0 ? ((((long long)1 << 30) - 1) / p->n) : 1;
        if (!capturing) CU(cudaEventRecord(p->ev[4], st));
        for (int done = 0; done < count;) {
            const int chunk = (int)((long long)(count - done) < per_launch ? (count - done) : per_launch);
            CU(cudaMemsetAsync(p->d_flags, 0, (size_t)p->cfg.G * kFlagStride * sizeof(unsigned long long), st));
            WaveCfg c = p->cfg;
            c.nsub = chunk;
            c.slots = p->d_slots + slot0 + done;
            c.halo = p->d_halo;
            c.flags = p->d_flags;
            c.err = p->d_err;
            c.btmax = p->d_btmax;
            c.prof = p->prof_on ? p->d_prof : nullptr;
            c.exec = p->d_exec;
            c.wd_cycles = watchdog_cycles();
            c.decouple = p->prof_on ? decouple_env() : 0;  // profiling experiment (PROF instantiation only), never a result
            CU(launch_wavefront(p->tab, c, p->argw, st));
            done += chunk;
            if (done < count) p->launches += 1;
        }
        if (!capturing) CU(cudaEventRecord(p->ev[5], st));
        p->launches += 1;
        p->last_path = 1;
    } else {
        for (int s = slot0; s < slot0 + count; ++s) {
            int l = (p->sp_ok && !p->sp_off) ? launch_stage_pruned_path(p->tab, p->slots_dev[s], p->argw, p->sp, p->d_exec, st)
                                             : launch_stage_path(p->tab, p->slots_dev[s], p->argw, st);
            if (l < 0) return fail(BB200_ERR_ARG, "shape not supported by the per-stage kernels (K=%d)", p->K);
            p->launches += l;
        }
        p->last_path = 0;
    }
    if (capturing) return BB200_OK;
    CU(cudaGetLastError());
    CU(cudaEventRecord(p->ev[1], st));
    p->dp_timed = true;
    return BB200_OK;
}

int queue_backtrack(bb200_plan *p, int slot, int64_t B_new, int rec_idx, bool capturing = false)
{
    if (!p->slots[slot].has_dp) return fail(BB200_ERR_STATE, "slot %d has no DP result yet", slot);
    if (B_new < 0 || B_new > p->B) return fail(BB200_ERR_ARG, "B_new=%lld outside [0, %lld]", (long long)B_new, (long long)p->B);
    cudaStream_t st = p->stream;
    SlotDev sd = p->slots_dev[slot];
    sd.rec = p->slots[slot].rec + 4 * rec_idx;
    if (!capturing) CU(cudaEventRecord(p->ev[2], st));
    CU(cudaMemsetAsync(p->d_err + 1, 0, sizeof(int), st));  // the stale flag is per selection
    launch_select(p->tab, sd, (int)B_new, capturing ? p->d_bnew : nullptr, p->d_err, st);
    launch_backtrack(p->tab, sd, p->argw, p->d_err, st);
    p->launches += 2;
    if (capturing) return BB200_OK;
    CU(cudaGetLastError());
    CU(cudaEventRecord(p->ev[3], st));
    p->bt_timed = true;
    return BB200_OK;
}

// Captures one TR inner iteration of slot 0 -- H2D(df, u_old, B') -> prep -> DP -> selection -> backtrack ->
// D2H(u, optimum, error word) -- into a CUDA graph (multi-trust.jl:112-113 as one replayable unit).  All shapes
// are fixed for a TRM run, the trial budget travels through device memory.  Returns false when the
// capture is not possible (then bb200_solve launches directly).
bool ensure_graph(bb200_plan *p)
{
    if (p->graph_exec) return true;
    if (p->graph_failed || (p->flags & BB200_FLAG_NO_GRAPH)) return false;
    const size_t io = (size_t)p->n * p->M * sizeof(double);
    if (!p->h_df) {
        if (cudaMallocHost((void **)&p->h_df, io) != cudaSuccess || cudaMallocHost((void **)&p->h_uold, io) != cudaSuccess ||
            cudaMallocHost((void **)&p->h_u, io) != cudaSuccess || cudaMallocHost((void **)&p->h_bnew, sizeof(int)) != cudaSuccess ||
            cudaMalloc((void **)&p->d_bnew, sizeof(int)) != cudaSuccess) {
            cudaGetLastError();
            p->graph_failed = true;
            return false;
        }
    }
    cudaStream_t st = p->stream;
    cudaGraph_t graph = nullptr;
    const double launches_before = p->launches;
    const bool had_dp = p->slots[0].has_dp;
    bool ok = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
        ok = cudaMemcpyAsync(p->slots[0].df, p->h_df, io, cudaMemcpyHostToDevice, st) == cudaSuccess &&
             cudaMemcpyAsync(p->slots[0].u_old, p->h_uold, io, cudaMemcpyHostToDevice, st) == cudaSuccess &&
             cudaMemcpyAsync(p->d_bnew, p->h_bnew, sizeof(int), cudaMemcpyHostToDevice, st) == cudaSuccess &&
             queue_dp(p, 0, 1, true) == BB200_OK && queue_backtrack(p, 0, p->B, 0, true) == BB200_OK &&
             cudaMemcpyAsync(p->h_u, p->slots[0].u, io, cudaMemcpyDeviceToHost, st) == cudaSuccess &&
             cudaMemcpyAsync(p->h_rec, p->slots[0].rec, 4 * sizeof(double), cudaMemcpyDeviceToHost, st) == cudaSuccess &&
             cudaMemcpyAsync(p->h_err, p->d_err, 4 * sizeof(int), cudaMemcpyDeviceToHost, st) == cudaSuccess &&
             cudaMemcpyAsync(p->h_exec, p->d_exec, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st) == cudaSuccess;
        cudaError_t e = cudaStreamEndCapture(st, &graph);
        ok = ok && e == cudaSuccess && graph != nullptr;
    }
    p->launches = launches_before;  // capture launches nothing
    p->slots[0].has_dp = had_dp;
    if (ok) ok = cudaGraphInstantiate(&p->graph_exec, graph, 0) == cudaSuccess;
    if (graph) cudaGraphDestroy(graph);
    if (!ok) {
        cudaGetLastError();
        p->graph_exec = nullptr;
        p->graph_failed = true;
    }
    return ok;
}

// The pruned scan only pays when its bound test drops most blocks; that depends on the data (jump costs against the
// spread of the value rows, and the horizon: the first stages after the terminal one bound little).  Measured on B200
// against the exhaustive tiles (config 4 shape, horizon sweep, profiles/prune_break_even_r02.txt and the final builds): the
// pruned stage costs 1.3 us + 12 us x (fraction of the candidates evaluated), the exhaustive one 4.3 us -- break-even at
// 25 % (n = 100 000 evaluates 7.7 %: 2.2 us; n = 10 000 16 %; n = 2 500 27 %).  After a synchronised DP that evaluated more,
// the plan goes back to the exhaustive tiles for its following DPs (TRM / a batch call the DP again on similar data).
constexpr double kPruneBreakEven = 0.25;
// (The per-stage pruned kernels pay for their bound tests with fewer candidates too: past half of them the plain
// per-stage kernels take over.)
constexpr double kStagePruneBreakEven = 0.5;
void adapt_pruning(bb200_plan *p, int slots, unsigned long long executed)
{
    if (p->sp_ok && !p->sp_off && !p->wave_ok && !p->mini_ok) {
        const double full = (double)(p->n > 1 ? p->n - 1 : 1) * p->B1 * (double)p->Kp * p->K * slots;
        if ((double)executed > kStagePruneBreakEven * full) {
            p->sp_off = true;
            p->prune_switches += 1;
            if (p->graph_exec) { cudaGraphExecDestroy(p->graph_exec); p->graph_exec = nullptr; }  // the launches are baked in
        }
        return;
    }
    if (!p->wave_ok || p->cfg.PR == 0 || p->tune_variant != 0 || p->prune_off) return;
    const double full = (double)(p->n > 1 ? p->n - 1 : 1) * p->B1 * (double)p->cfg.Kr * p->K * slots;
    if ((double)executed > kPruneBreakEven * full) {
        p->prune_off = true;
        p->prune_switches += 1;
        if (p->graph_exec) { cudaGraphExecDestroy(p->graph_exec); p->graph_exec = nullptr; }  // geometry is baked in
        reconfigure(p);
    }
}

// Synchronise and translate the deferred device-side error word.
int sync_and_check(bb200_plan *p)
{
    CU(cudaMemcpyAsync(p->h_err, p->d_err, 4 * sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    CU(cudaMemcpyAsync(p->h_exec, p->d_exec, sizeof(unsigned long long), cudaMemcpyDeviceToHost, p->stream));
    CU(cudaStreamSynchronize(p->stream));
    if (p->dp_timed) adapt_pruning(p, p->last_dp_slots, *p->h_exec);
    if (p->dp_timed) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p->ev[0], p->ev[1]) == cudaSuccess) p->last_dp_ms = ms;
        if (p->last_path >= 1 && cudaEventElapsedTime(&ms, p->ev[4], p->ev[5]) == cudaSuccess) p->last_wave_ms = ms;
        p->dp_timed = false;
    }
    if (p->bt_timed) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p->ev[2], p->ev[3]) == cudaSuccess) p->last_bt_ms = ms;
        p->bt_timed = false;
    }
    if (p->h_err[2]) return fail(BB200_ERR_CUDA, "wavefront kernel watchdog fired: a pipeline dependency was never satisfied");
    if (p->h_err[0]) return fail(BB200_ERR_INEXACT, "InexactError: u_old is not integer valued / finite (HelpFunctions.jl:37,57)");
    if (p->h_err[1]) return fail(BB200_ERR_STALE, "selection/backtrack reached a cell the DP never wrote (no feasible trajectory for this u_old / budget)");
    return BB200_OK;
}

int check_slot(bb200_plan *p, int slot)
{
    if (!p) return fail(BB200_ERR_ARG, "plan is NULL");
    if (slot < 0 || slot >= p->batch) return fail(BB200_ERR_ARG, "slot %d outside [0, %d)", slot, p->batch);
    return BB200_OK;
}

struct Guard {
    std::lock_guard<std::mutex> lk;
    explicit Guard(bb200_plan *p) : lk(p->mu) { cudaSetDevice(p->device); }
};

}  // namespace

namespace bb200 {
// the calling thread's last-error message, for the other host translation units (bb200_multi.cu)
int set_error(int code, const char *msg)
{
    g_err = msg;
    return code;
}
const char *get_error() { return g_err.c_str(); }
}  // namespace bb200

extern "C" {

const char *bb200_last_error(void) { return g_err.c_str(); }
int bb200_version(void) { return 1000; }

int bb200_device_count(void)
{
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); return 0; }
    return c;
}

int bb200_plan_create(int device, int64_t n, int32_t M, int32_t K, int64_t B, const int64_t *grid_dims,
                      const int32_t *level_values, const int64_t *grid_offset, const double *jump_cost,
                      double dt, int32_t batch, uint32_t flags, bb200_plan **out)
{
    if (!out) return fail(BB200_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (!grid_dims || !level_values || !grid_offset || !jump_cost) return fail(BB200_ERR_ARG, "NULL table pointer");
    if (n < 1 || n > 0x7fffff00ll) return fail(BB200_ERR_ARG, "n=%lld out of range", (long long)n);
    if (M < 1 || M > kMaxM) return fail(BB200_ERR_ARG, "M=%d out of range [1, %d]", M, kMaxM);
    if (K < 1 || K > 65534) return fail(BB200_ERR_ARG, "K=%d out of range [1, 65534]", K);
    if (B < 0 || B > 0x3fffffffll) return fail(BB200_ERR_ARG, "B=%lld out of range", (long long)B);
    if (batch < 1) return fail(BB200_ERR_ARG, "batch=%d must be >= 1", batch);
    int64_t G = 1;
    for (int m = 0; m < M; ++m) {
        if (grid_dims[m] < 1) return fail(BB200_ERR_ARG, "grid_dims[%d]=%lld", m, (long long)grid_dims[m]);
        G *= grid_dims[m];
    }
    for (int k = 0; k < K; ++k)
        if (grid_offset[k] < 0 || grid_offset[k] >= G) return fail(BB200_ERR_ARG, "grid_offset[%d] outside the grid", k);
    int ndev = bb200_device_count();
    if (ndev == 0) return fail(BB200_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(BB200_ERR_ARG, "device %d outside [0, %d)", device, ndev);
    CU(cudaSetDevice(device));

    bb200_plan *p = new bb200_plan();
    p->device = device; p->n = n; p->M = M; p->K = K; p->B = B; p->B1 = (int)(B + 1);
    p->Kp = (K + 31) / 32 * 32; p->dt = dt; p->batch = batch; p->flags = flags; p->G = G;
    p->argw = (K <= 255) ? 1 : 2;
    p->nPad = (n + kChunk - 1) / kChunk * kChunk;
    p->grid_dims.assign(grid_dims, grid_dims + M);
    p->grid_offset.assign(grid_offset, grid_offset + K);
    p->level_values.assign(level_values, level_values + (size_t)K * M);
    int rc = BB200_OK;
    auto bail = [&](int code) { destroy_plan(p); return code; };

    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(fail(BB200_ERR_CUDA, "cudaGetDeviceProperties failed"));
    p->num_sms = prop.multiProcessorCount;
    p->smem_max = prop.sharedMemPerBlockOptin;

    // constant tables
    std::vector<double> lvd((size_t)K * M), cost((size_t)K * p->Kp, std::numeric_limits<double>::infinity());
    for (size_t x = 0; x < lvd.size(); ++x) lvd[x] = (double)level_values[x];
    for (int j = 0; j < K; ++j)
        for (int l = 0; l < K; ++l) cost[(size_t)j * p->Kp + l] = jump_cost[(size_t)j * K + l];
    std::vector<long long> goff(grid_offset, grid_offset + K);
    if ((rc = dev_alloc(p, &p->d_lvd, lvd.size()))) return bail(rc);
    if ((rc = dev_alloc(p, &p->d_cost, cost.size()))) return bail(rc);
    if ((rc = dev_alloc(p, &p->d_goff, goff.size()))) return bail(rc);
    if ((rc = dev_alloc(p, &p->d_err, 4))) return bail(rc);
    if ((rc = dev_alloc(p, &p->d_btmax, 1))) return bail(rc);
    if ((rc = dev_alloc(p, &p->d_exec, 1))) return bail(rc);
    if ((rc = dev_alloc(p, &p->d_scalar, 8))) return bail(rc);
    if ((rc = dev_alloc(p, &p->d_flags, (size_t)prop.multiProcessorCount * kFlagStride))) return bail(rc);
    if ((rc = dev_alloc(p, &p->d_slots, (size_t)batch))) return bail(rc);
    if ((rc = dev_alloc(p, &p->d_prof, (size_t)prop.multiProcessorCount * 16))) return bail(rc);
#define CUB(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return bail(fail(BB200_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)));      \
    } while (0)
    CUB(cudaMemcpy(p->d_lvd, lvd.data(), lvd.size() * sizeof(double), cudaMemcpyHostToDevice));
    CUB(cudaMemcpy(p->d_cost, cost.data(), cost.size() * sizeof(double), cudaMemcpyHostToDevice));
    CUB(cudaMemcpy(p->d_goff, goff.data(), goff.size() * sizeof(long long), cudaMemcpyHostToDevice));
    CUB(cudaMemset(p->d_err, 0, 4 * sizeof(int)));

    p->tab.n = (int)n; p->tab.M = M; p->tab.K = K; p->tab.Kp = p->Kp; p->tab.B1 = p->B1; p->tab.dt = dt;
    p->tab.lvd = p->d_lvd; p->tab.cost = p->d_cost; p->tab.goff = p->d_goff;

    if ((rc = dev_alloc(p, &p->d_rec_all, (size_t)batch * kRecDoubles))) return bail(rc);
    if ((rc = dev_alloc(p, &p->d_radii, (size_t)kMaxRadii))) return bail(rc);
    p->slots.resize(batch);
    p->slots_dev.resize(batch);
    const size_t io = (size_t)p->nPad * M;
    const size_t argcells = (size_t)(n > 1 ? n - 1 : 1) * p->B1 * p->Kp;
    for (int s = 0; s < batch; ++s) {
        SlotHost &h = p->slots[s];
        if ((rc = dev_alloc(p, &h.df, io))) return bail(rc);
        if ((rc = dev_alloc(p, &h.u_old, io))) return bail(rc);
        if ((rc = dev_alloc(p, &h.u, io))) return bail(rc);
        if ((rc = dev_alloc(p, &h.phi, (size_t)2 * p->B1 * p->Kp))) return bail(rc);
        h.rec = p->d_rec_all + (size_t)s * kRecDoubles;
        if ((rc = dev_alloc(p, &h.n_updates, 1))) return bail(rc);
        if ((rc = dev_alloc(p, &h.ss_all, (size_t)n * p->Kp))) return bail(rc);
        if ((rc = dev_alloc(p, &h.bt_all, (size_t)n * p->Kp))) return bail(rc);
        unsigned char *a = nullptr;
        if ((rc = dev_alloc(p, &a, argcells * p->argw))) return bail(rc);
        h.arg = a;
        CUB(cudaMemset(h.df, 0, io * sizeof(double)));
        CUB(cudaMemset(h.u_old, 0, io * sizeof(double)));
        CUB(cudaMemset(h.u, 0, io * sizeof(double)));
        SlotDev &d = p->slots_dev[s];
        d.df = h.df; d.u_old = h.u_old; d.u = h.u; d.phi = h.phi; d.arg = h.arg;
        d.n_updates = h.n_updates; d.rec = h.rec; d.ss_all = h.ss_all; d.bt_all = h.bt_all;
    }
    CUB(cudaMemcpy(p->d_slots, p->slots_dev.data(), sizeof(SlotDev) * batch, cudaMemcpyHostToDevice));
    CUB(cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking));
    p->stream = p->own_stream;
    for (auto &e : p->ev) CUB(cudaEventCreate(&e));
    CUB(cudaMallocHost((void **)&p->h_rec, kRecDoubles * sizeof(double)));
    CUB(cudaMemset(p->d_rec_all, 0, (size_t)batch * kRecDoubles * sizeof(double)));
    CUB(cudaMallocHost((void **)&p->h_err, 4 * sizeof(int)));
    CUB(cudaMallocHost((void **)&p->h_exec, sizeof(unsigned long long)));
    *p->h_exec = 0ull;
#undef CUB
    if ((rc = reconfigure(p))) return bail(rc);
    *out = p;
    return BB200_OK;
}

int bb200_plan_destroy(bb200_plan *plan)
{
    if (!plan) return BB200_OK;
    destroy_plan(plan);
    return BB200_OK;
}

int bb200_plan_set_stream(bb200_plan *plan, void *cuda_stream)
{
    if (!plan) return fail(BB200_ERR_ARG, "plan is NULL");
    Guard g(plan);
    CU(cudaStreamSynchronize(plan->stream));
    plan->stream = cuda_stream ? (cudaStream_t)cuda_stream : plan->own_stream;
    if (plan->graph_exec) { cudaGraphExecDestroy(plan->graph_exec); plan->graph_exec = nullptr; }  // captured on the old stream
    if (plan->wave_ok) pin_ring_in_l2(plan);
    return BB200_OK;
}

int bb200_plan_tune(bb200_plan *plan, int32_t ctas, int32_t jsplit, int32_t variant)
{
    if (!plan) return fail(BB200_ERR_ARG, "plan is NULL");
    Guard g(plan);
    plan->tune_ctas = ctas; plan->tune_js = jsplit; plan->tune_variant = variant;
    if (plan->graph_exec) { cudaGraphExecDestroy(plan->graph_exec); plan->graph_exec = nullptr; }  // geometry is baked in
    int rc = reconfigure(plan);
    if (rc) return rc;
    if (!plan->wave_ok && !plan->mini_ok && !(plan->flags & BB200_FLAG_STAGE_KERNELS) && (ctas || jsplit || variant))
        return fail(BB200_ERR_ARG, "requested tuning (ctas=%d, jsplit=%d, variant=%d) is not launchable", ctas, jsplit, variant);
    return BB200_OK;
}

int bb200_wave_geometry(int64_t n, int32_t M, int32_t K, int64_t B, int32_t num_sms, int64_t smem_max, int32_t ctas,
                        int32_t jsplit, int32_t variant, int64_t *out, int32_t count)
{
    if (!out || n < 1 || M < 1 || K < 1 || B < 0 || num_sms < 1 || smem_max < 0) return fail(BB200_ERR_ARG, "bad arguments");
    Tables t{};
    t.n = (int)n; t.M = M; t.K = K; t.Kp = (K + 31) / 32 * 32; t.B1 = (int)B + 1; t.dt = 1.0;
    WaveCfg c{};
    const bool ok = wave_configure(t, K <= 255 ? 1 : 2, num_sms, (size_t)smem_max, ctas, jsplit, variant, c);
    const int64_t v[16] = {ok ? 1 : 0, ok ? c.variant + 1 : 0, c.TB, c.TBB, c.TL, c.G, c.R, c.JS, c.jper, c.Kr, c.NS,
                           c.threads, (int64_t)c.smem, c.PR, c.GA, c.Rtop};
    for (int k = 0; k < count && k < 16; ++k) out[k] = ok || k == 0 ? v[k] : 0;
    return BB200_OK;
}

int bb200_upload(bb200_plan *plan, int32_t slot, const double *df, const double *u_old)
{
    int rc = check_slot(plan, slot);
    if (rc) return rc;
    if (!df || !u_old) return fail(BB200_ERR_ARG, "df/u_old is NULL");
    Guard g(plan);
    const size_t bytes = (size_t)plan->n * plan->M * sizeof(double);
    CU(cudaMemcpyAsync(plan->slots[slot].df, df, bytes, cudaMemcpyHostToDevice, plan->stream));
    CU(cudaMemcpyAsync(plan->slots[slot].u_old, u_old, bytes, cudaMemcpyHostToDevice, plan->stream));
    return BB200_OK;
}

int bb200_upload_device(bb200_plan *plan, int32_t slot, const double *d_df, const double *d_u_old)
{
    int rc = check_slot(plan, slot);
    if (rc) return rc;
    if (!d_df || !d_u_old) return fail(BB200_ERR_ARG, "d_df/d_u_old is NULL");
    Guard g(plan);
    const size_t bytes = (size_t)plan->n * plan->M * sizeof(double);
    CU(cudaMemcpyAsync(plan->slots[slot].df, d_df, bytes, cudaMemcpyDeviceToDevice, plan->stream));
    CU(cudaMemcpyAsync(plan->slots[slot].u_old, d_u_old, bytes, cudaMemcpyDeviceToDevice, plan->stream));
    return BB200_OK;
}

int bb200_bellman_resident(bb200_plan *plan, int32_t slot0, int32_t count)
{
    int rc = check_slot(plan, slot0);
    if (rc) return rc;
    if (count < 1 || slot0 + count > plan->batch) return fail(BB200_ERR_ARG, "slot range [%d, %d) outside the plan", slot0, slot0 + count);
    Guard g(plan);
    return queue_dp(plan, slot0, count);
}

int bb200_backtrack_resident(bb200_plan *plan, int32_t slot, int64_t B_new)
{
    int rc = check_slot(plan, slot);
    if (rc) return rc;
    Guard g(plan);
    return queue_backtrack(plan, slot, B_new, 0);
}

int bb200_sync(bb200_plan *plan)
{
    if (!plan) return fail(BB200_ERR_ARG, "plan is NULL");
    Guard g(plan);
    return sync_and_check(plan);
}

static int download_locked(bb200_plan *plan, int slot, int rec_idx, double *u_out, double *phi_star,
                           int64_t *b_star, int64_t *k_star)
{
    if (u_out)
        CU(cudaMemcpyAsync(u_out, plan->slots[slot].u, (size_t)plan->n * plan->M * sizeof(double),
                           cudaMemcpyDeviceToHost, plan->stream));
    CU(cudaMemcpyAsync(plan->h_rec, plan->slots[slot].rec + 4 * rec_idx, 4 * sizeof(double),
                       cudaMemcpyDeviceToHost, plan->stream));
    int rc = sync_and_check(plan);
    if (phi_star) *phi_star = plan->h_rec[0];
    if (b_star) *b_star = (int64_t)plan->h_rec[1];
    if (k_star) *k_star = (int64_t)plan->h_rec[2];
    return rc;
}

int bb200_download(bb200_plan *plan, int32_t slot, double *u_out, double *phi_star, int64_t *b_star,
                   int64_t *k_star)
{
    int rc = check_slot(plan, slot);
    if (rc) return rc;
    Guard g(plan);
    return download_locked(plan, slot, 0, u_out, phi_star, b_star, k_star);
}

int bb200_bellman(bb200_plan *plan, const double *df, const double *u_old)
{
    int rc = bb200_upload(plan, 0, df, u_old);
    if (rc) return rc;
    Guard g(plan);
    rc = queue_dp(plan, 0, 1);
    if (rc) return rc;
    return sync_and_check(plan);
}

int bb200_select_and_backtrack(bb200_plan *plan, int64_t B_new, double *u_out, double *phi_star,
                               int64_t *b_star, int64_t *k_star)
{
    if (!plan) return fail(BB200_ERR_ARG, "plan is NULL");
    if (!u_out) return fail(BB200_ERR_ARG, "u_out is NULL");
    Guard g(plan);
    int rc = queue_backtrack(plan, 0, B_new, 0);
    if (rc) return rc;
    return download_locked(plan, 0, 0, u_out, phi_star, b_star, k_star);
}

int bb200_solve(bb200_plan *plan, const double *df, const double *u_old, int64_t B_new, double *u_out,
                double *phi_star, int64_t *b_star, int64_t *k_star)
{
    if (!plan) return fail(BB200_ERR_ARG, "plan is NULL");
    if (!df || !u_old || !u_out) return fail(BB200_ERR_ARG, "df/u_old/u_out is NULL");
    if (B_new < 0 || B_new > plan->B) return fail(BB200_ERR_ARG, "B_new=%lld outside [0, %lld]", (long long)B_new, (long long)plan->B);
    Guard g(plan);
    const size_t io = (size_t)plan->n * plan->M * sizeof(double);
    if (ensure_graph(plan)) {
        // replay: stage the inputs in pinned memory, one graph launch, one synchronisation
        std::memcpy(plan->h_df, df, io);
        std::memcpy(plan->h_uold, u_old, io);
        *plan->h_bnew = (int)B_new;
        CU(cudaEventRecord(plan->ev[0], plan->stream));
        CU(cudaGraphLaunch(plan->graph_exec, plan->stream));
        CU(cudaEventRecord(plan->ev[1], plan->stream));
        CU(cudaStreamSynchronize(plan->stream));
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, plan->ev[0], plan->ev[1]) == cudaSuccess) plan->last_dp_ms = ms;
        plan->slots[0].has_dp = true;
        const bool one_launch = plan->wave_ok || plan->mini_ok;
        plan->launches += 4 + (one_launch ? 0 : (double)plan->n - 1);  // prep, DP, selection, backtrack
        plan->last_path = plan->mini_ok ? 2 : (plan->wave_ok ? 1 : 0);
        plan->graph_replays += 1;
        adapt_pruning(plan, 1, *plan->h_exec);
        std::memcpy(u_out, plan->h_u, io);
        if (phi_star) *phi_star = plan->h_rec[0];
        if (b_star) *b_star = (int64_t)plan->h_rec[1];
        if (k_star) *k_star = (int64_t)plan->h_rec[2];
        if (plan->h_err[2]) return fail(BB200_ERR_CUDA, "wavefront kernel watchdog fired: a pipeline dependency was never satisfied");
        if (plan->h_err[0]) return fail(BB200_ERR_INEXACT, "InexactError: u_old is not integer valued / finite (HelpFunctions.jl:37,57)");
        if (plan->h_err[1]) return fail(BB200_ERR_STALE, "selection/backtrack reached a cell the DP never wrote (no feasible trajectory for this u_old / budget)");
        return BB200_OK;
    }
    CU(cudaMemcpyAsync(plan->slots[0].df, df, io, cudaMemcpyHostToDevice, plan->stream));
    CU(cudaMemcpyAsync(plan->slots[0].u_old, u_old, io, cudaMemcpyHostToDevice, plan->stream));
    int rc = queue_dp(plan, 0, 1);
    if (rc) return rc;
    rc = queue_backtrack(plan, 0, B_new, 0);
    if (rc) return rc;
    return download_locked(plan, 0, 0, u_out, phi_star, b_star, k_star);
}

// ---- batched radius sweep as a pipeline -----------------------------------------------------------------------
// Lazily creates what only the batched path needs: the copy stream, the events and the pinned staging buffers
// (inputs of one wave: [batch][2][n*M]; outputs: [batch][n_radii][n*M] + the record blocks), two of each.
static int ensure_batch_pipe(bb200_plan *p, int n_radii, bool want_u)
{
    const size_t io = (size_t)p->n * p->M;
    if (!p->copy_stream) {
        CU(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&p->ev_prep, cudaEventDisableTiming));
        for (int k = 0; k < 2; ++k) {
            CU(cudaEventCreateWithFlags(&p->ev_in[k], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&p->ev_out[k], cudaEventDisableTiming));
            CU(cudaEventCreate(&p->ev_batch[k]));
            CU(cudaMallocHost((void **)&p->h_recs[k], (size_t)p->batch * kRecDoubles * sizeof(double)));
            CU(cudaMallocHost((void **)&p->h_errs[k], 4 * sizeof(int)));
            CU(cudaMallocHost((void **)&p->h_execs[k], sizeof(unsigned long long)));
        }
    }
    const size_t in_need = (size_t)p->batch * 2 * io;
    if (in_need > p->h_in_elems) {
        for (int k = 0; k < 2; ++k) {
            if (p->h_in[k]) cudaFreeHost(p->h_in[k]);
            p->h_in[k] = nullptr;
            CU(cudaMallocHost((void **)&p->h_in[k], in_need * sizeof(double)));
        }
        p->h_in_elems = in_need;
    }
    const size_t out_need = want_u ? (size_t)p->batch * n_radii * io : 0;
    if (out_need > p->h_out_elems) {
        for (int k = 0; k < 2; ++k) {
            if (p->h_out[k]) cudaFreeHost(p->h_out[k]);
            p->h_out[k] = nullptr;
            CU(cudaMallocHost((void **)&p->h_out[k], out_need * sizeof(double)));
        }
        p->h_out_elems = out_need;
    }
    const size_t sweep_need = (size_t)p->batch * n_radii * io;
    if (sweep_need > p->usweep_elems) {
        CU(cudaStreamSynchronize(p->stream));
        if (p->d_usweep) { cudaFree(p->d_usweep); p->dev_bytes -= p->usweep_elems * sizeof(double); }
        p->d_usweep = nullptr;
        p->usweep_elems = 0;
        int rc = dev_alloc(p, &p->d_usweep, sweep_need);
        if (rc) return rc;
        p->usweep_elems = sweep_need;
    }
    return BB200_OK;
}

// The subproblems first, first + stride, ... < S of the caller's arrays, in waves of `batch` resident slots:
//   copy stream:     H2D(w+1) as soon as the prep kernels of wave w have consumed df / u_old
//   compute stream:  prep(w) -> DP(w) (one persistent launch for all slots of the wave) -> ONE selection and ONE
//                    backtrack launch with a CTA per (slot, radius) -> D2H(w) into pinned staging
//   host:            stages wave w+1 while wave w runs; one event wait per wave (for the outputs of wave w-1)
int bb200_solve_batched_shard(bb200_plan *plan, int64_t S, int64_t first, int64_t stride, const double *df_all,
                              const double *u_old_all, int32_t n_radii, const int64_t *B_new, double *u_out_all,
                              double *phi_star, int64_t *b_star, int64_t *k_star, int32_t *status)
{
    if (!plan) return fail(BB200_ERR_ARG, "plan is NULL");
    if (S < 1 || !df_all || !u_old_all) return fail(BB200_ERR_ARG, "bad batch arguments");
    if (first < 0 || stride < 1) return fail(BB200_ERR_ARG, "bad shard (first=%lld, stride=%lld)", (long long)first, (long long)stride);
    if (n_radii < 1 || n_radii > kMaxRadii || !B_new) return fail(BB200_ERR_ARG, "n_radii=%d outside [1, %d]", n_radii, kMaxRadii);
    for (int r = 0; r < n_radii; ++r)
        if (B_new[r] < 0 || B_new[r] > plan->B) return fail(BB200_ERR_ARG, "B_new[%d]=%lld outside [0, %lld]", r, (long long)B_new[r], (long long)plan->B);
    Guard g(plan);
    const int64_t mine = first < S ? (S - first + stride - 1) / stride : 0;  // subproblems of this shard
    if (mine == 0) return BB200_OK;
    const bool want_u = u_out_all != nullptr;
    int rc = ensure_batch_pipe(plan, n_radii, want_u);
    if (rc) return rc;
    const size_t io = (size_t)plan->n * plan->M;
    const int batch = plan->batch;
    const int64_t waves = (mine + batch - 1) / batch;
    cudaStream_t st = plan->stream, cs = plan->copy_stream;
    int radii[kMaxRadii];
    for (int r = 0; r < n_radii; ++r) radii[r] = (int)B_new[r];
    CU(cudaMemcpyAsync(plan->d_radii, radii, n_radii * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(plan->ev_batch[0], st));

    auto sub = [&](int64_t w, int s) { return first + (w * batch + s) * stride; };  // global index of slot s in wave w
    auto count = [&](int64_t w) { return (int)std::min<int64_t>(batch, mine - w * batch); };
    auto stage_in = [&](int64_t w) -> int {  // host -> pinned -> device, on the copy stream
        double *h = plan->h_in[w & 1];
        const int cnt = count(w);
        if (w >= 2) CU(cudaEventSynchronize(plan->ev_in[w & 1]));  // the H2D of wave w-2 has left this buffer (long ago)
        for (int s = 0; s < cnt; ++s) {
            std::memcpy(h + (size_t)(2 * s) * io, df_all + (size_t)sub(w, s) * io, io * sizeof(double));
            std::memcpy(h + (size_t)(2 * s + 1) * io, u_old_all + (size_t)sub(w, s) * io, io * sizeof(double));
        }
        if (w > 0) CU(cudaStreamWaitEvent(cs, plan->ev_prep, 0));  // wave w-1's prep kernels have read their inputs
        for (int s = 0; s < cnt; ++s) {
            CU(cudaMemcpyAsync(plan->slots[s].df, h + (size_t)(2 * s) * io, io * sizeof(double), cudaMemcpyHostToDevice, cs));
            CU(cudaMemcpyAsync(plan->slots[s].u_old, h + (size_t)(2 * s + 1) * io, io * sizeof(double), cudaMemcpyHostToDevice, cs));
        }
        CU(cudaEventRecord(plan->ev_in[w & 1], cs));
        return BB200_OK;
    };
    int worst = BB200_OK;
    std::string worst_msg;
    auto drain = [&](int64_t w) -> int {  // outputs of wave w: pinned -> the caller's arrays
        CU(cudaEventSynchronize(plan->ev_out[w & 1]));
        plan->batch_syncs += 1;
        const int cnt = count(w);
        const double *hr = plan->h_recs[w & 1];
        const int *he = plan->h_errs[w & 1];
        if (he[2]) return fail(BB200_ERR_CUDA, "wavefront kernel watchdog fired: a pipeline dependency was never satisfied");
        adapt_pruning(plan, cnt, *plan->h_execs[w & 1]);  // later waves run on the exhaustive tiles if pruning did not pay
        for (int s = 0; s < cnt; ++s) {
            const int64_t gs = sub(w, s);
            const bool inexact = hr[(size_t)s * kRecDoubles + 4 * kMaxRadii] != 0.;
            for (int r = 0; r < n_radii; ++r) {
                const double *rec = hr + (size_t)s * kRecDoubles + 4 * r;
                const size_t o = (size_t)gs * n_radii + r;
                const int code = inexact ? BB200_ERR_INEXACT : (rec[3] != 0. ? BB200_ERR_STALE : BB200_OK);
                if (phi_star) phi_star[o] = rec[0];
                if (b_star) b_star[o] = (int64_t)rec[1];
                if (k_star) k_star[o] = (int64_t)rec[2];
                if (status) status[o] = code;
                if (want_u) std::memcpy(u_out_all + o * io, plan->h_out[w & 1] + ((size_t)s * n_radii + r) * io, io * sizeof(double));
                if (code && (!worst || (code == BB200_ERR_INEXACT && worst == BB200_ERR_STALE))) {
                    worst = code;
                    char buf[256];
                    snprintf(buf, sizeof buf, code == BB200_ERR_INEXACT
                                 ? "subproblem %lld: InexactError: u_old is not integer valued / finite (HelpFunctions.jl:37,57)"
                                 : "subproblem %lld, radius %d: selection/backtrack reached a cell the DP never wrote",
                             (long long)gs, r);
                    worst_msg = buf;
                }
            }
        }
        return BB200_OK;
    };

    if ((rc = stage_in(0))) return rc;
    for (int64_t w = 0; w < waves; ++w) {
        const int cnt = count(w);
        CU(cudaStreamWaitEvent(st, plan->ev_in[w & 1], 0));
        rc = queue_dp(plan, 0, cnt, false, plan->ev_prep);
        if (rc) return rc;
        if (w + 1 < waves && (rc = stage_in(w + 1))) return rc;  // overlaps the DP of wave w
        SweepDev sw{plan->d_slots, plan->d_radii, n_radii, plan->d_usweep, io};
        CU(cudaMemsetAsync(plan->d_err + 1, 0, sizeof(int), st));
        launch_select(plan->tab, plan->slots_dev[0], 0, nullptr, plan->d_err, st, &sw, cnt * n_radii);
        launch_backtrack(plan->tab, plan->slots_dev[0], plan->argw, plan->d_err, st, &sw, cnt * n_radii);
        plan->launches += 2;
        CU(cudaGetLastError());
        if (want_u)
            CU(cudaMemcpyAsync(plan->h_out[w & 1], plan->d_usweep, (size_t)cnt * n_radii * io * sizeof(double), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(plan->h_recs[w & 1], plan->d_rec_all, (size_t)cnt * kRecDoubles * sizeof(double), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(plan->h_errs[w & 1], plan->d_err, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(plan->h_execs[w & 1], plan->d_exec, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CU(cudaEventRecord(plan->ev_out[w & 1], st));
        if (w + 1 == waves) CU(cudaEventRecord(plan->ev_batch[1], st));
        if (w > 0 && (rc = drain(w - 1))) return rc;
    }
    if ((rc = drain(waves - 1))) return rc;
    plan->dp_timed = false;
    plan->batch_waves = (double)waves;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, plan->ev_batch[0], plan->ev_batch[1]) == cudaSuccess) plan->last_batch_ms = ms;
    if (worst) g_err = worst_msg;
    return worst;
}

int bb200_solve_batched(bb200_plan *plan, int64_t S, const double *df_all, const double *u_old_all,
                        int32_t n_radii, const int64_t *B_new, double *u_out_all, double *phi_star,
                        int64_t *b_star, int64_t *k_star, int32_t *status)
{
    return bb200_solve_batched_shard(plan, S, 0, 1, df_all, u_old_all, n_radii, B_new, u_out_all, phi_star, b_star,
                                     k_star, status);
}

// Julia findmin order over (value, index) records, like the selection kernel (S8): NaN precedes every number,
// -0.0 precedes +0.0, equal values go to the smaller index.
int bb200_best_candidate(const double *values, const int64_t *indices, int64_t count, double *best_value,
                         int64_t *best_index)
{
    if (!values || !indices || count < 1 || !best_value || !best_index) return fail(BB200_ERR_ARG, "bad arguments");
    auto isless = [](double a, double b) {
        if (a != a) return false;
        if (b != b) return true;
        if (a < b) return true;
        if (a == b) return std::signbit(a) && !std::signbit(b);
        return false;
    };
    auto isgreater = [&](double fm, double fx) { return (fm != fm || fx != fx) ? isless(fm, fx) : isless(fx, fm); };
    double bv = values[0];
    int64_t bi = indices[0];
    for (int64_t x = 1; x < count; ++x) {
        const double v = values[x];
        const bool better = isgreater(bv, v) || (!isgreater(v, bv) && indices[x] < bi);
        if (better) { bv = v; bi = indices[x]; }
    }
    *best_value = bv;
    *best_index = bi;
    return BB200_OK;
}

int bb200_export_phi(bb200_plan *plan, int32_t slot, double *Phi_out)
{
    int rc = check_slot(plan, slot);
    if (rc) return rc;
    if (!Phi_out) return fail(BB200_ERR_ARG, "Phi_out is NULL");
    Guard g(plan);
    if (!plan->slots[slot].has_dp) return fail(BB200_ERR_STATE, "slot %d has no DP result yet", slot);
    const size_t cells = (size_t)2 * plan->B1 * plan->Kp;
    std::vector<double> h(cells);
    CU(cudaStreamSynchronize(plan->stream));
    CU(cudaMemcpy(h.data(), plan->slots[slot].phi, cells * sizeof(double), cudaMemcpyDeviceToHost));
    const size_t B1 = plan->B1;
    const double inf = std::numeric_limits<double>::infinity();
    for (size_t x = 0; x < 2 * (size_t)plan->G * B1; ++x) Phi_out[x] = inf;
    for (int s = 0; s < 2; ++s)
        for (int k = 0; k < plan->K; ++k)
            for (size_t b = 0; b < B1; ++b)
                Phi_out[b + B1 * ((size_t)plan->grid_offset[k] + (size_t)plan->G * s)] = h[((size_t)s * B1 + b) * plan->Kp + k];
    // n == 1: the reference never touches slot 2 (it keeps the caller's contents); we report +Inf there.
    return BB200_OK;
}

int bb200_export_argmin(bb200_plan *plan, int32_t slot, int64_t i0, int64_t i1, int64_t *U_out, int64_t fill)
{
    int rc = check_slot(plan, slot);
    if (rc) return rc;
    if (!U_out) return fail(BB200_ERR_ARG, "U_out is NULL");
    if (i0 < 1 || i1 > plan->n || i0 > i1) return fail(BB200_ERR_ARG, "stage range [%lld, %lld) invalid", (long long)i0, (long long)i1);
    Guard g(plan);
    if (!plan->slots[slot].has_dp) return fail(BB200_ERR_STATE, "slot %d has no DP result yet", slot);
    CU(cudaStreamSynchronize(plan->stream));
    const size_t B1 = plan->B1, Kp = plan->Kp;
    const int M = plan->M, K = plan->K;
    const size_t rows = (size_t)(i1 - i0);
    if (rows == 0) return BB200_OK;
    std::vector<unsigned char> a(rows * B1 * Kp * plan->argw);
    std::vector<double> uo(rows * M);
    CU(cudaMemcpy(a.data(), (unsigned char *)plan->slots[slot].arg + (size_t)(i0 - 1) * B1 * Kp * plan->argw,
                  a.size(), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(uo.data(), plan->slots[slot].u_old + (size_t)(i0 - 1) * M, uo.size() * sizeof(double), cudaMemcpyDeviceToHost));
    // 1-based tuple of every admissible level
    std::vector<int64_t> tup((size_t)K * M);
    for (int k = 0; k < K; ++k) {
        int64_t rem = plan->grid_offset[k];
        for (int m = 0; m < M; ++m) { tup[(size_t)k * M + m] = rem % plan->grid_dims[m] + 1; rem /= plan->grid_dims[m]; }
    }
    const size_t G = (size_t)plan->G;
    for (size_t x = 0; x < rows * G * B1 * M; ++x) U_out[x] = fill;
    const unsigned mark = plan->argw == 1 ? 0xffu : 0xffffu;
    for (size_t r = 0; r < rows; ++r)
        for (int k = 0; k < K; ++k) {
            double bd = 0.;
            for (int m = 0; m < M; ++m) bd += std::fabs((double)plan->level_values[(size_t)k * M + m] - uo[r * M + m]);
            if (!(bd < (double)B1)) continue;
            const size_t bt = (size_t)bd;
            for (size_t bsrc = 0; bsrc + bt < B1; ++bsrc) {
                const size_t cell = (r * B1 + bsrc) * Kp + k;
                const unsigned v = plan->argw == 1 ? a[cell] : ((const uint16_t *)a.data())[cell];
                if (v == mark || v >= (unsigned)K) continue;
                int64_t *dst = U_out + M * ((bsrc + bt) + B1 * ((size_t)plan->grid_offset[k] + G * r));
                for (int m = 0; m < M; ++m) dst[m] = tup[(size_t)v * M + m];
            }
        }
    return BB200_OK;
}

int bb200_count_updates(bb200_plan *plan, int32_t slot, int64_t *n_updates)
{
    int rc = check_slot(plan, slot);
    if (rc) return rc;
    if (!n_updates) return fail(BB200_ERR_ARG, "n_updates is NULL");
    Guard g(plan);
    if (!plan->slots[slot].has_dp) return fail(BB200_ERR_STATE, "slot %d has no DP result yet", slot);
    CU(cudaStreamSynchronize(plan->stream));
    unsigned long long v = 0;
    CU(cudaMemcpy(&v, plan->slots[slot].n_updates, sizeof v, cudaMemcpyDeviceToHost));
    *n_updates = (int64_t)v;
    return BB200_OK;
}

int bb200_pred_integral(bb200_plan *plan, int32_t slot, double *int_val)
{
    int rc = check_slot(plan, slot);
    if (rc) return rc;
    if (!int_val) return fail(BB200_ERR_ARG, "int_val is NULL");
    Guard g(plan);
    launch_pred_integral(plan->tab, plan->slots_dev[slot], plan->d_scalar, plan->stream);
    plan->launches += 1;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(plan->h_rec, plan->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, plan->stream));
    CU(cudaStreamSynchronize(plan->stream));
    *int_val = plan->h_rec[0];
    return BB200_OK;
}

int bb200_tv(bb200_plan *plan, int32_t slot, double p, double *tv)
{
    int rc = check_slot(plan, slot);
    if (rc) return rc;
    if (!tv) return fail(BB200_ERR_ARG, "tv is NULL");
    int mode;
    if (std::isinf(p) && p > 0) mode = 0;
    else if (p == 1.) mode = 1;
    else if (p == 2.) mode = 2;
    else return fail(BB200_ERR_ARG, "TV_p on the device supports p in {1, 2, Inf}; got %g", p);
    Guard g(plan);
    launch_tv(plan->tab, plan->slots_dev[slot], mode, plan->d_scalar, plan->stream);
    plan->launches += 1;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(plan->h_rec, plan->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, plan->stream));
    CU(cudaStreamSynchronize(plan->stream));
    *tv = plan->h_rec[0];
    return BB200_OK;
}

int bb200_profile(bb200_plan *plan, int32_t enable, int64_t *out, int32_t max_ctas)
{
    if (!plan) return fail(BB200_ERR_ARG, "plan is NULL");
    Guard g(plan);
    plan->prof_on = enable != 0;
    if (out && max_ctas > 0) {
        CU(cudaStreamSynchronize(plan->stream));
        const int nct = std::min<int>(max_ctas, plan->num_sms);
        CU(cudaMemcpy(out, plan->d_prof, (size_t)nct * 16 * sizeof(long long), cudaMemcpyDeviceToHost));
    }
    return BB200_OK;
}

int bb200_fp64_peak(int device, int32_t mode, double target_ms, double *ops_per_s, double *elapsed_ms)
{
    if (!ops_per_s || !elapsed_ms || mode < 0 || mode > 1) return fail(BB200_ERR_ARG, "bad arguments");
    if (bb200_device_count() == 0) return fail(BB200_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    CU(measure_fp64_rate(mode, prop.multiProcessorCount, target_ms, ops_per_s, elapsed_ms, 0));
    return BB200_OK;
}

int bb200_stats(bb200_plan *plan, double *out, int32_t count)
{
    if (!plan || !out) return fail(BB200_ERR_ARG, "bad arguments");
    Guard g(plan);
    unsigned long long exec = 0;
    const bool sp_on = plan->sp_ok && !plan->sp_off && !plan->wave_ok && !plan->mini_ok;
    if (count > 17 && ((plan->cfg.PR > 0 && plan->wave_ok) || sp_on)) {
        CU(cudaStreamSynchronize(plan->stream));
        CU(cudaMemcpy(&exec, plan->d_exec, sizeof exec, cudaMemcpyDeviceToHost));
    }
    const double v[22] = {plan->last_dp_ms, plan->last_bt_ms, plan->launches, (double)plan->last_path,
                          plan->wave_ok ? (double)plan->cfg.G : 0., plan->wave_ok ? (double)plan->cfg.R : 0.,
                          (double)plan->argw, (double)plan->dev_bytes,
                          plan->wave_ok ? (double)plan->cfg.threads : 0., plan->wave_ok ? (double)plan->cfg.JS : 0.,
                          plan->last_wave_ms, plan->graph_replays,
                          plan->wave_ok ? (double)(plan->cfg.variant + 1) : 0., plan->wave_ok ? (double)plan->cfg.NS : 0.,
                          plan->last_batch_ms, plan->batch_waves, plan->batch_syncs, (double)exec,
                          plan->wave_ok ? (double)plan->cfg.PR : (sp_on ? 4. : 0.), plan->prune_switches,
                          plan->wave_ok ? (double)plan->cfg.GA : 0., plan->wave_ok ? (double)plan->cfg.Rtop : 0.};
    for (int k = 0; k < count && k < 22; ++k) out[k] = v[k];
    return BB200_OK;
}

}  // extern "C"
