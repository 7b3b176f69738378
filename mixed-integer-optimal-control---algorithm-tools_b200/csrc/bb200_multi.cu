// bb200_multi.cu -- multi-GPU behind the C ABI: independent subproblems sharded s -> device s mod G, no
// data-path collective, ONE exchange at the end: every rank contributes a 16-byte (value, global index) record,
// ncclAllGather over NVLink/NVSwitch, then the deterministic local reduction bb200_best_candidate (NCCL has no MINLOC).
//
//   bb200_comm_*    one rank of a communicator (multi-process use: one process per GPU, the unique id travels
//                   through whatever channel the host program has -- torch.distributed, MPI, Julia Distributed)
//   bb200_multi_*   all GPUs of one process: one plan and one host thread per device, ncclCommInitAll
//
// NCCL is loaded at run time (dlopen "libnccl.so.2"): the library has no link-time dependency on it, so it still
// loads where NCCL is absent, and inside a torch process it binds to the NCCL torch has already loaded instead of a
// second copy.  Without NCCL the comm / multi entry points fail with BB200_ERR_STATE; nothing falls back.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/bellman_b200.h"
#include "bb200_internal.cuh"

namespace bb200 {
int set_error(int code, const char *msg);
const char *get_error();
}  // namespace bb200

namespace {

int failf(int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    return bb200::set_error(code, buf);
}

struct Nccl {
    void *h = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    std::string why;
};

Nccl &nccl()
{
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {getenv("BELLMAN_B200_NCCL"), "libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            if (!nm || !*nm) continue;
            n.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (n.h) break;
        }
        if (!n.h) { n.why = "libnccl.so.2 not found (set BELLMAN_B200_NCCL to its path)"; return; }
#define SYM(field, name)                                                      \
    n.field = reinterpret_cast<decltype(n.field)>(dlsym(n.h, name));          \
    if (!n.field) { n.why = std::string("symbol missing in NCCL: ") + name; n.h = nullptr; return; }
        SYM(GetUniqueId, "ncclGetUniqueId")
        SYM(CommInitRank, "ncclCommInitRank")
        SYM(CommInitAll, "ncclCommInitAll")
        SYM(CommDestroy, "ncclCommDestroy")
        SYM(AllGather, "ncclAllGather")
        SYM(Broadcast, "ncclBroadcast")
        SYM(GetErrorString, "ncclGetErrorString")
        SYM(GetVersion, "ncclGetVersion")
#undef SYM
    });
    return n;
}

#define NC(call)                                                                                        \
    do {                                                                                                \
        ncclResult_t r_ = (call);                                                                       \
        if (r_ != ncclSuccess) return failf(BB200_ERR_CUDA, "%s failed: %s", #call, nccl().GetErrorString(r_)); \
    } while (0)
#define CUM(call)                                                                                       \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) return failf(BB200_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

}  // namespace

struct bb200_comm {
    int device = 0, rank = 0, nranks = 1;
    ncclComm_t comm = nullptr;
    cudaStream_t stream = nullptr;
    double *d_send = nullptr, *d_recv = nullptr, *h_buf = nullptr;  // h_buf pinned: [2 + 2 * nranks]
    double *d_bcast = nullptr;
    size_t bcast_elems = 0;
    double collectives = 0.;
};

static int comm_finish_init(bb200_comm *c)
{
    CUM(cudaSetDevice(c->device));
    CUM(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUM(cudaMalloc((void **)&c->d_send, 2 * sizeof(double)));
    CUM(cudaMalloc((void **)&c->d_recv, (size_t)2 * c->nranks * sizeof(double)));
    CUM(cudaMallocHost((void **)&c->h_buf, (size_t)(2 + 2 * c->nranks) * sizeof(double)));
    return BB200_OK;
}

extern "C" {

int bb200_nccl_version(void)
{
    if (!nccl().h) return 0;
    int v = 0;
    return nccl().GetVersion(&v) == ncclSuccess ? v : 0;
}

int bb200_comm_unique_id(void *id128)
{
    if (!id128) return failf(BB200_ERR_ARG, "id128 is NULL");
    if (!nccl().h) return failf(BB200_ERR_STATE, "NCCL is not available: %s", nccl().why.c_str());
    static_assert(sizeof(ncclUniqueId) == 128, "the ABI passes the NCCL unique id as 128 opaque bytes");
    ncclUniqueId id;
    NC(nccl().GetUniqueId(&id));
    std::memcpy(id128, &id, sizeof id);
    return BB200_OK;
}

int bb200_comm_create(int device, int32_t nranks, int32_t rank, const void *id128, bb200_comm **out)
{
    if (!out) return failf(BB200_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) return failf(BB200_ERR_ARG, "bad communicator arguments");
    if (!nccl().h) return failf(BB200_ERR_STATE, "NCCL is not available: %s", nccl().why.c_str());
    if (bb200_device_count() == 0) return failf(BB200_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
    bb200_comm *c = new bb200_comm();
    c->device = device; c->rank = rank; c->nranks = nranks;
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof id);
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { delete c; return failf(BB200_ERR_CUDA, "cudaSetDevice(%d) failed: %s", device, cudaGetErrorString(e)); }
    ncclResult_t r = nccl().CommInitRank(&c->comm, nranks, id, rank);
    if (r != ncclSuccess) { delete c; return failf(BB200_ERR_CUDA, "ncclCommInitRank failed: %s", nccl().GetErrorString(r)); }
    int rc = comm_finish_init(c);
    if (rc) { bb200_comm_destroy(c); return rc; }
    *out = c;
    return BB200_OK;
}

int bb200_comm_destroy(bb200_comm *c)
{
    if (!c) return BB200_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm && nccl().h) nccl().CommDestroy(c->comm);
    cudaFree(c->d_send); cudaFree(c->d_recv); cudaFree(c->d_bcast);
    if (c->h_buf) cudaFreeHost(c->h_buf);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return BB200_OK;
}

// Every rank calls this with its local best; all ranks return the same global (value, index) and the rank that
// contributed it.  16 bytes per rank through ncclAllGather, then the deterministic host reduction.
int bb200_comm_best_candidate(bb200_comm *c, double value, int64_t index, double *best_value, int64_t *best_index,
                              int32_t *owner_rank)
{
    if (!c || !best_value || !best_index) return failf(BB200_ERR_ARG, "bad arguments");
    if (index < 0 || index >= ((int64_t)1 << 53)) return failf(BB200_ERR_ARG, "index %lld does not fit a double exactly", (long long)index);
    CUM(cudaSetDevice(c->device));
    c->h_buf[0] = value;
    c->h_buf[1] = (double)index;
    CUM(cudaMemcpyAsync(c->d_send, c->h_buf, 2 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    NC(nccl().AllGather(c->d_send, c->d_recv, 2, ncclDouble, c->comm, c->stream));
    CUM(cudaMemcpyAsync(c->h_buf + 2, c->d_recv, (size_t)2 * c->nranks * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUM(cudaStreamSynchronize(c->stream));
    c->collectives += 1;
    std::vector<double> v(c->nranks);
    std::vector<int64_t> ix(c->nranks);
    for (int r = 0; r < c->nranks; ++r) { v[r] = c->h_buf[2 + 2 * r]; ix[r] = (int64_t)c->h_buf[3 + 2 * r]; }
    int rc = bb200_best_candidate(v.data(), ix.data(), c->nranks, best_value, best_index);
    if (rc) return rc;
    if (owner_rank) {
        *owner_rank = 0;
        for (int r = 0; r < c->nranks; ++r)
            if (ix[r] == *best_index && std::memcmp(&v[r], best_value, sizeof(double)) == 0) { *owner_rank = r; break; }
    }
    return BB200_OK;
}

// The winner's trajectory (or any host array of doubles) from rank `root` to every rank: root -> device -> ncclBroadcast
// over NVLink -> host of the other ranks.
int bb200_comm_broadcast(bb200_comm *c, int32_t root, double *host_buf, int64_t count)
{
    if (!c || !host_buf || count < 1 || root < 0 || root >= c->nranks) return failf(BB200_ERR_ARG, "bad arguments");
    CUM(cudaSetDevice(c->device));
    if ((size_t)count > c->bcast_elems) {
        cudaFree(c->d_bcast);
        c->d_bcast = nullptr;
        c->bcast_elems = 0;
        CUM(cudaMalloc((void **)&c->d_bcast, (size_t)count * sizeof(double)));
        c->bcast_elems = (size_t)count;
    }
    if (c->rank == root) CUM(cudaMemcpyAsync(c->d_bcast, host_buf, (size_t)count * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    NC(nccl().Broadcast(c->d_bcast, c->d_bcast, (size_t)count, ncclDouble, root, c->comm, c->stream));
    if (c->rank != root) CUM(cudaMemcpyAsync(host_buf, c->d_bcast, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUM(cudaStreamSynchronize(c->stream));
    c->collectives += 1;
    return BB200_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// All GPUs of one process
// ---------------------------------------------------------------------------------------------------------------
struct bb200_multi {
    std::vector<int> devices;
    std::vector<bb200_plan *> plans;
    std::vector<bb200_comm *> comms;
    int64_t n = 0;
    int M = 0;
    std::vector<double> last_ms;  // per device: device time of its shard in the last call
    double last_wall_ms = 0.;
    std::mutex mu;
};

extern "C" {

int bb200_multi_destroy(bb200_multi *m)
{
    if (!m) return BB200_OK;
    for (auto *c : m->comms) bb200_comm_destroy(c);
    for (auto *p : m->plans) bb200_plan_destroy(p);
    delete m;
    return BB200_OK;
}

int bb200_multi_create(const int32_t *devices, int32_t n_dev, int64_t n, int32_t M, int32_t K, int64_t B,
                       const int64_t *grid_dims, const int32_t *level_values, const int64_t *grid_offset,
                       const double *jump_cost, double dt, int32_t batch_per_device, uint32_t flags, bb200_multi **out)
{
    if (!out) return failf(BB200_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (!devices || n_dev < 1) return failf(BB200_ERR_ARG, "no devices given");
    for (int a = 0; a < n_dev; ++a)
        for (int b = a + 1; b < n_dev; ++b)
            if (devices[a] == devices[b]) return failf(BB200_ERR_ARG, "device %d listed twice", devices[a]);
    bb200_multi *m = new bb200_multi();
    m->n = n; m->M = M;
    m->devices.assign(devices, devices + n_dev);
    m->last_ms.assign(n_dev, 0.);
    for (int d = 0; d < n_dev; ++d) {
        bb200_plan *p = nullptr;
        int rc = bb200_plan_create(devices[d], n, M, K, B, grid_dims, level_values, grid_offset, jump_cost, dt,
                                   batch_per_device, flags, &p);
        if (rc) { bb200_multi_destroy(m); return rc; }
        m->plans.push_back(p);
    }
    if (n_dev > 1) {
        if (!nccl().h) { bb200_multi_destroy(m); return failf(BB200_ERR_STATE, "NCCL is not available: %s", nccl().why.c_str()); }
        std::vector<ncclComm_t> cs(n_dev);
        std::vector<int> devs(devices, devices + n_dev);
        ncclResult_t r = nccl().CommInitAll(cs.data(), n_dev, devs.data());
        if (r != ncclSuccess) { bb200_multi_destroy(m); return failf(BB200_ERR_CUDA, "ncclCommInitAll failed: %s", nccl().GetErrorString(r)); }
        for (int d = 0; d < n_dev; ++d) {
            bb200_comm *c = new bb200_comm();
            c->device = devices[d]; c->rank = d; c->nranks = n_dev; c->comm = cs[d];
            m->comms.push_back(c);
            int rc = comm_finish_init(c);
            if (rc) { bb200_multi_destroy(m); return rc; }
        }
    }
    *out = m;
    return BB200_OK;
}

// S independent subproblems with a radius sweep each over all devices of `m`, subproblem s on device s mod G
// (equal cost -> static partition), then the best-candidate reduction.  Outputs as bb200_solve_batched; additionally
//   best_value / best_subproblem / best_radius   the smallest selected value over all entries with status OK, in
//                   Julia findmin order; ties go to the smallest (subproblem, radius)      (each may be NULL)
//   u_best          double[n][M], the winner's trajectory (may be NULL); taken from u_out_all, or -- when that is
//                   NULL -- recomputed by one more DP of the winning subproblem on its device
int bb200_multi_solve_batched(bb200_multi *m, int64_t S, const double *df_all, const double *u_old_all, int32_t n_radii,
                              const int64_t *B_new, double *u_out_all, double *phi_star, int64_t *b_star, int64_t *k_star,
                              int32_t *status, double *best_value, int64_t *best_subproblem, int32_t *best_radius,
                              double *u_best)
{
    if (!m) return failf(BB200_ERR_ARG, "multi plan is NULL");
    if (S < 1 || !df_all || !u_old_all || !B_new || n_radii < 1) return failf(BB200_ERR_ARG, "bad batch arguments");
    std::lock_guard<std::mutex> lk(m->mu);
    const int G = (int)m->plans.size();
    const size_t io = (size_t)m->n * m->M;
    std::vector<double> phi_tmp;
    std::vector<int32_t> st_tmp;
    if (!phi_star) { phi_tmp.assign((size_t)S * n_radii, std::numeric_limits<double>::quiet_NaN()); phi_star = phi_tmp.data(); }
    if (!status) { st_tmp.assign((size_t)S * n_radii, 0); status = st_tmp.data(); }
    std::vector<int> rcs(G, BB200_OK);
    std::vector<std::string> msgs(G);
    std::vector<double> gv(G, 0.);
    std::vector<int64_t> gi(G, 0);
    std::vector<int32_t> owner(G, 0);
    const double inf = std::numeric_limits<double>::infinity();
    auto t_start = std::chrono::steady_clock::now();
    auto work = [&](int d) {
        int rc = bb200_solve_batched_shard(m->plans[d], S, d, G, df_all, u_old_all, n_radii, B_new, u_out_all, phi_star,
                                           b_star, k_star, status);
        if (rc) msgs[d] = bb200::get_error();
        const bool hard = rc != BB200_OK && rc != BB200_ERR_INEXACT && rc != BB200_ERR_STALE;
        rcs[d] = rc;
        double st[17];
        if (bb200_stats(m->plans[d], st, 17) == BB200_OK) m->last_ms[d] = st[14];
        // this rank's best over the entries it owns (global index = s * n_radii + r)
        std::vector<double> v;
        std::vector<int64_t> ix;
        if (!hard)
            for (int64_t s = d; s < S; s += G)
                for (int r = 0; r < n_radii; ++r)
                    if (status[(size_t)s * n_radii + r] == BB200_OK) { v.push_back(phi_star[(size_t)s * n_radii + r]); ix.push_back(s * n_radii + r); }
        double bv = inf;
        int64_t bi = ((int64_t)1 << 53) - 1;  // "no candidate": loses every tie
        if (!v.empty()) bb200_best_candidate(v.data(), ix.data(), (int64_t)v.size(), &bv, &bi);
        if (G > 1) {
            int rc2 = bb200_comm_best_candidate(m->comms[d], bv, bi, &gv[d], &gi[d], &owner[d]);
            if (rc2 && !hard) { rcs[d] = rc2; msgs[d] = bb200::get_error(); }
        } else {
            gv[d] = bv; gi[d] = bi; owner[d] = 0;
        }
    };
    std::vector<std::thread> th;
    for (int d = 1; d < G; ++d) th.emplace_back(work, d);
    work(0);
    for (auto &t : th) t.join();
    m->last_wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count();
    int worst = BB200_OK;
    for (int d = 0; d < G; ++d) {
        const bool hard = rcs[d] != BB200_OK && rcs[d] != BB200_ERR_INEXACT && rcs[d] != BB200_ERR_STALE;
        if (hard) return bb200::set_error(rcs[d], msgs[d].c_str());
        if (rcs[d] && (!worst || (rcs[d] == BB200_ERR_INEXACT && worst == BB200_ERR_STALE))) { worst = rcs[d]; bb200::set_error(worst, msgs[d].c_str()); }
    }
    for (int d = 1; d < G; ++d)
        if (gi[d] != gi[0] || std::memcmp(&gv[d], &gv[0], sizeof(double)) != 0)
            return failf(BB200_ERR_CUDA, "best-candidate reduction disagrees between ranks 0 and %d", d);
    const bool have = gi[0] < ((int64_t)1 << 53) - 1;
    const int64_t bs = have ? gi[0] / n_radii : -1;
    const int br = have ? (int)(gi[0] % n_radii) : -1;
    if (best_value) *best_value = gv[0];
    if (best_subproblem) *best_subproblem = bs;
    if (best_radius) *best_radius = br;
    if (u_best) {
        if (!have) return failf(BB200_ERR_STALE, "no subproblem produced a feasible trajectory");
        if (u_out_all) {
            std::memcpy(u_best, u_out_all + ((size_t)bs * n_radii + br) * io, io * sizeof(double));
        } else {
            int rc = bb200_solve(m->plans[bs % G], df_all + (size_t)bs * io, u_old_all + (size_t)bs * io, B_new[br], u_best,
                                 nullptr, nullptr, nullptr);
            if (rc) return rc;
        }
    }
    return worst;
}

/* out[0] = devices, out[1] = host wall time [ms] of the last call, out[2 + d] = device time [ms] of device d's shard */
int bb200_multi_stats(bb200_multi *m, double *out, int32_t count)
{
    if (!m || !out) return failf(BB200_ERR_ARG, "bad arguments");
    std::lock_guard<std::mutex> lk(m->mu);
    std::vector<double> v{(double)m->plans.size(), m->last_wall_ms};
    v.insert(v.end(), m->last_ms.begin(), m->last_ms.end());
    for (int k = 0; k < count && k < (int)v.size(); ++k) out[k] = v[k];
    return BB200_OK;
}

}  // extern "C"
