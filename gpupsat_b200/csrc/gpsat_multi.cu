// Multi-GPU host in ONE process (include/gpsat.h: gpsat_multi_*): the N-GPU form of what SATSolver/main.cu:197-310
// does for one GPU.  One gpsat handle and one host thread per GPU; formula and cube list are replicated,
// and the handles are joined in a mesh (gpsat_mesh_attach_local) so that the GPUs behave as one work pool over
// NVLink peer memory.  A solve is ONE persistent launch per GPU; afterwards the per-cube outcome flags / open-descendant
// counts / records of the ranks are reduced with ncclAllReduce (MAX, SUM, SUM).  libnccl is loaded at run time with
// dlopen, so that libgpsat.so itself carries no NCCL dependency and a process that also loads torch's NCCL sees one
// copy; when it cannot be loaded the same reduction runs on the host (reduce_backend = 0).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gpsat.h"
#include "host_formula.h"

namespace {

using gpsat_host::set_error;

// ---- the few NCCL entry points used, resolved at run time (nccl.h: ncclResult_t 0 = success) -----------------------
typedef struct ncclComm *ncclComm_t;
enum { kNcclInt32 = 2, kNcclInt64 = 4 };   // ncclDataType_t: ncclInt32 = 2, ncclInt64 = 4
enum { kNcclSum = 0, kNcclMax = 2 };       // ncclRedOp_t: ncclSum = 0, ncclProd = 1, ncclMax = 2
struct Nccl {
    void *lib = nullptr;
    int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};

Nccl &nccl()
{
    static Nccl n;
    static bool tried = false;
    if (tried) return n;
    tried = true;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
        n.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
        if (n.lib) break;
    }
    if (!n.lib) return n;
    n.CommInitAll = (int (*)(ncclComm_t *, int, const int *))dlsym(n.lib, "ncclCommInitAll");
    n.CommDestroy = (int (*)(ncclComm_t))dlsym(n.lib, "ncclCommDestroy");
    n.GroupStart = (int (*)())dlsym(n.lib, "ncclGroupStart");
    n.GroupEnd = (int (*)())dlsym(n.lib, "ncclGroupEnd");
    n.AllReduce = (int (*)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(n.lib, "ncclAllReduce");
    n.GetErrorString = (const char *(*)(int))dlsym(n.lib, "ncclGetErrorString");
    n.ok = n.CommInitAll && n.CommDestroy && n.GroupStart && n.GroupEnd && n.AllReduce;
    return n;
}

const double kStepBudgetMs = 2000.0;

}  // namespace

struct gpsat_multi {
    int n = 0;
    std::vector<int> devices;
    std::vector<gpsat_t *> h;
    int32_t n_vars = 0;
    int32_t n_cubes = 0;
    bool cubes_set = false;
    gpsat_opts opts{};
    // result reduction
    std::vector<int32_t *> block;       // per GPU: device result block
    std::vector<cudaStream_t> stream;
    int64_t block_words = 0;
    std::vector<ncclComm_t> comm;
    bool nccl_ready = false;
    std::vector<gpsat_job_record> records;
    double time_limit_ms = 0;           // 0 = until done
};

namespace {

void free_blocks(gpsat_multi *m)
{
    for (int r = 0; r < m->n; r++) {
        if ((size_t)r < m->block.size() && m->block[r]) {
            cudaSetDevice(m->devices[r]);
            cudaFree(m->block[r]);
            m->block[r] = nullptr;
        }
    }
    m->block_words = 0;
}

int ensure_blocks(gpsat_multi *m, int64_t words)
{
    if (m->block_words >= words) return GPSAT_OK;
    free_blocks(m);
    m->block.assign((size_t)m->n, nullptr);
    for (int r = 0; r < m->n; r++) {
        cudaSetDevice(m->devices[r]);
        if (cudaMalloc((void **)&m->block[r], (size_t)words * 4) != cudaSuccess) {
            set_error("cudaMalloc of the result block failed");
            return GPSAT_E_CUDA;
        }
    }
    m->block_words = words;
    return GPSAT_OK;
}

}  // namespace

extern "C" {

int gpsat_multi_create(gpsat_multi_t **out, int32_t n_gpus, const int32_t *devices, int32_t n_vars, int64_t n_clauses,
                       const int64_t *offsets, const int32_t *lits, const gpsat_opts *opts)
{
    if (!out) {
        set_error("null handle pointer");
        return GPSAT_E_ARG;
    }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error(std::string("no usable CUDA device (") + cudaGetErrorString(e) + "); gpupsat_b200 has no CPU fallback");
        return GPSAT_E_NO_DEVICE;
    }
    if (n_gpus <= 0) n_gpus = std::min(count, GPSAT_MESH_MAX_GPUS);
    if (n_gpus > GPSAT_MESH_MAX_GPUS || (!devices && n_gpus > count)) {
        set_error("gpsat_multi_create: " + std::to_string(n_gpus) + " GPUs requested, " + std::to_string(count) + " visible");
        return GPSAT_E_ARG;
    }
    gpsat_multi *m = new gpsat_multi();
    m->n = n_gpus;
    m->n_vars = n_vars;
    if (opts) m->opts = *opts;
    else gpsat_opts_default(&m->opts);
    m->opts.dynamic_split = 1;   // a mesh is a pool of split-off cubes
    m->h.assign((size_t)n_gpus, nullptr);
    m->stream.assign((size_t)n_gpus, nullptr);
    int prev = 0;
    cudaGetDevice(&prev);
    for (int r = 0; r < n_gpus; r++) {
        m->devices.push_back(devices ? devices[r] : r);
        gpsat_opts o = m->opts;
        o.device = m->devices[r];
        int rc = gpsat_create(&m->h[r], n_vars, n_clauses, offsets, lits, &o);
        if (rc == GPSAT_OK && cudaStreamCreateWithFlags(&m->stream[r], cudaStreamNonBlocking) != cudaSuccess) {
            set_error("cudaStreamCreate failed");
            rc = GPSAT_E_CUDA;
        }
        if (rc != GPSAT_OK) {
            cudaSetDevice(prev);
            gpsat_multi_destroy(m);
            return rc;
        }
    }
    cudaSetDevice(prev);
    *out = m;
    return GPSAT_OK;
}

int gpsat_multi_n_gpus(gpsat_multi_t *m) { return m ? m->n : 0; }

int gpsat_multi_set_time_limit(gpsat_multi_t *m, double ms)
{
    if (!m || ms < 0) {
        set_error("bad arguments");
        return GPSAT_E_ARG;
    }
    m->time_limit_ms = ms;
    return GPSAT_OK;
}

int gpsat_multi_set_cubes(gpsat_multi_t *m, int32_t n_cubes, const int64_t *cube_offsets, const int32_t *cube_lits)
{
    if (!m || n_cubes < 0 || (n_cubes > 0 && !cube_offsets)) {
        set_error("bad arguments");
        return GPSAT_E_ARG;
    }
    // every GPU holds the complete cube list; the mesh hands the cubes out through ONE cursor.  n_cubes = 0: the single
    // empty cube (the reference's sequential mode) — one GPU takes it, the others live off the cubes it splits off
    const int32_t total = std::max(n_cubes, 1);
    for (int r = 0; r < m->n; r++) {
        int rc = gpsat_set_cubes(m->h[r], n_cubes, cube_offsets, cube_lits);
        if (rc != GPSAT_OK) return rc;
    }
    int rc = gpsat_mesh_attach_local(m->h.data(), m->n);
    if (rc != GPSAT_OK) return rc;
    m->n_cubes = total;
    m->cubes_set = true;
    return GPSAT_OK;
}

int gpsat_multi_solve(gpsat_multi_t *m, int32_t *verdict, uint8_t *model, gpsat_stats *stats, int32_t *reduce_backend)
{
    if (!m) {
        set_error("null handle");
        return GPSAT_E_ARG;
    }
    if (!m->cubes_set) {
        int rc = gpsat_multi_set_cubes(m, 0, nullptr, nullptr);
        if (rc != GPSAT_OK) return rc;
    }
    const int N = m->n;
    std::vector<int> rcs((size_t)N, GPSAT_OK);
    std::vector<std::string> errs((size_t)N);
    std::vector<int32_t> local_verdict((size_t)N, GPSAT_UNDEF);
    std::vector<gpsat_stats> local_stats((size_t)N);
    std::vector<std::vector<uint8_t>> local_model((size_t)N, std::vector<uint8_t>((size_t)std::max(m->n_vars, 1), 0));
    std::atomic<int> arrived{0}, failed{0};
    const int64_t words = gpsat_mesh_result_words(m->h[0]);
    int rc0 = ensure_blocks(m, words);
    if (rc0 != GPSAT_OK) return rc0;

    // one host thread per GPU: begin -> barrier (no rank may steal from a ring its owner has not reset) -> ONE launch
    auto worker = [&](int r) {
        int rc = gpsat_solve_begin(m->h[r]);
        if (rc != GPSAT_OK) failed.fetch_add(1);
        arrived.fetch_add(1);
        while (arrived.load() < N) std::this_thread::yield();
        if (failed.load() == 0) {
            // time-bounded launches (unfinished cubes park in place and resume): a GPU whose peer never shows up ends
            // its step instead of spinning for ever; a run shorter than the budget is a single launch
            int32_t done = 0, v = GPSAT_UNDEF;
            const auto t_start = std::chrono::steady_clock::now();
            double left = m->time_limit_ms;
            do {
                const double budget = m->time_limit_ms > 0 ? std::min(kStepBudgetMs, std::max(left, 1.0)) : kStepBudgetMs;
                rc = gpsat_solve_step(m->h[r], budget, &done, &v);
                left = m->time_limit_ms - std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count();
            } while (rc == GPSAT_OK && !done && (m->time_limit_ms <= 0 || left > 0));
            if (rc == GPSAT_OK) rc = gpsat_solve_end(m->h[r], &local_verdict[(size_t)r], local_model[(size_t)r].data(), &local_stats[(size_t)r]);
            if (rc == GPSAT_OK) rc = gpsat_mesh_results_pack(m->h[r], m->block[r], words);
        }
        if (rc != GPSAT_OK) errs[(size_t)r] = gpsat_last_error();
        rcs[(size_t)r] = rc;
    };
    std::vector<std::thread> threads;
    for (int r = 1; r < N; r++) threads.emplace_back(worker, r);
    worker(0);
    for (auto &t : threads) t.join();
    for (int r = 0; r < N; r++)
        if (rcs[(size_t)r] != GPSAT_OK) {
            set_error("GPU " + std::to_string(m->devices[r]) + ": " + errs[(size_t)r]);
            return rcs[(size_t)r];
        }

    // ---- reduce the per-rank result blocks: [flags: MAX int32][open descendants: SUM int32][records: SUM int64]
    const int64_t nr = m->n_cubes;
    int backend = 0;
    int prev = 0;
    cudaGetDevice(&prev);
    if (N > 1) {
        Nccl &nc = nccl();
        if (nc.ok && !m->nccl_ready) {
            m->comm.assign((size_t)N, nullptr);
            if (nc.CommInitAll(m->comm.data(), N, m->devices.data()) == 0) m->nccl_ready = true;
            else m->comm.clear();
        }
        if (m->nccl_ready) {
            bool ok = nc.GroupStart() == 0;
            for (int r = 0; r < N && ok; r++) {
                int32_t *b = m->block[r];
                ok = ok && nc.AllReduce(b, b, (size_t)nr, kNcclInt32, kNcclMax, m->comm[(size_t)r], m->stream[(size_t)r]) == 0;
                ok = ok && nc.AllReduce(b + nr, b + nr, (size_t)nr, kNcclInt32, kNcclSum, m->comm[(size_t)r], m->stream[(size_t)r]) == 0;
                ok = ok && nc.AllReduce(b + 2 * nr, b + 2 * nr, (size_t)nr * (sizeof(gpsat_job_record) / 8), kNcclInt64, kNcclSum,
                                        m->comm[(size_t)r], m->stream[(size_t)r]) == 0;
            }
            ok = (nc.GroupEnd() == 0) && ok;
            for (int r = 0; r < N; r++) {
                cudaSetDevice(m->devices[r]);
                ok = (cudaStreamSynchronize(m->stream[(size_t)r]) == cudaSuccess) && ok;
            }
            if (!ok) {
                cudaSetDevice(prev);
                set_error("ncclAllReduce of the result blocks failed");
                return GPSAT_E_CUDA;
            }
            backend = 1;
        } else {
            // host reduction (libnccl not loadable): same arithmetic on the host, result written back to rank 0's block
            std::vector<int32_t> acc((size_t)words, 0), tmp((size_t)words);
            for (int r = 0; r < N; r++) {
                cudaSetDevice(m->devices[r]);
                if (cudaMemcpy(tmp.data(), m->block[r], (size_t)words * 4, cudaMemcpyDeviceToHost) != cudaSuccess) {
                    cudaSetDevice(prev);
                    set_error("copy of a result block failed");
                    return GPSAT_E_CUDA;
                }
                for (int64_t i = 0; i < nr; i++) acc[(size_t)i] = r == 0 ? tmp[(size_t)i] : std::max(acc[(size_t)i], tmp[(size_t)i]);
                for (int64_t i = nr; i < 2 * nr; i++) acc[(size_t)i] += tmp[(size_t)i];
                int64_t *a64 = (int64_t *)(acc.data() + 2 * nr);
                const int64_t *t64 = (const int64_t *)(tmp.data() + 2 * nr);
                for (int64_t i = 0; i < nr * (int64_t)(sizeof(gpsat_job_record) / 8); i++) a64[i] += t64[i];
            }
            cudaSetDevice(m->devices[0]);
            cudaMemcpy(m->block[0], acc.data(), (size_t)words * 4, cudaMemcpyHostToDevice);
        }
    }
    cudaSetDevice(prev);
    int32_t v = GPSAT_UNDEF;
    gpsat_stats total;
    int rc = gpsat_mesh_results_unpack(m->h[0], m->block[0], words, &v, &total);
    if (rc != GPSAT_OK) return rc;
    m->records.resize((size_t)nr);
    rc = gpsat_job_records(m->h[0], m->records.data(), (int32_t)nr);
    if (rc != GPSAT_OK) return rc;
    // whole-run statistics: the counters come from the reduced records; launch-level figures from the ranks
    total.kernel_ms = 0;
    total.warp_busy_frac = 0;
    total.steals = 0;
    total.foreign_clauses = 0;
    total.pool_clauses = 0;
    for (int r = 0; r < N; r++) {
        total.kernel_ms = std::max(total.kernel_ms, local_stats[(size_t)r].kernel_ms);
        total.warp_busy_frac += local_stats[(size_t)r].warp_busy_frac / N;
        total.steals += local_stats[(size_t)r].steals;
        total.foreign_clauses += local_stats[(size_t)r].foreign_clauses;
        total.pool_clauses += local_stats[(size_t)r].pool_clauses;
    }
    total.blocks = local_stats[0].blocks * N;
    if (v == GPSAT_SAT && model) {
        bool found = false;
        for (int r = 0; r < N && !found; r++)
            if (local_verdict[(size_t)r] == GPSAT_SAT) {
                std::memcpy(model, local_model[(size_t)r].data(), (size_t)m->n_vars);
                found = true;
            }
        if (!found) {
            set_error("SAT verdict without a model on any GPU");
            return GPSAT_E_STATE;
        }
    }
    if (verdict) *verdict = v;
    if (stats) *stats = total;
    if (reduce_backend) *reduce_backend = backend;
    return GPSAT_OK;
}

int gpsat_multi_job_records(gpsat_multi_t *m, gpsat_job_record *records, int32_t cap)
{
    if (!m || !records || cap < (int32_t)m->records.size() || m->records.empty()) {
        set_error("record buffer too small or no run yet");
        return GPSAT_E_CAPACITY;
    }
    std::memcpy(records, m->records.data(), m->records.size() * sizeof(gpsat_job_record));
    return GPSAT_OK;
}

void gpsat_multi_destroy(gpsat_multi_t *m)
{
    if (!m) return;
    int prev = 0;
    cudaGetDevice(&prev);
    if (m->nccl_ready)
        for (auto c : m->comm)
            if (c) nccl().CommDestroy(c);
    free_blocks(m);
    for (int r = 0; r < (int)m->h.size(); r++) {
        if ((size_t)r < m->stream.size() && m->stream[(size_t)r]) {
            cudaSetDevice(m->devices[(size_t)r]);
            cudaStreamDestroy(m->stream[(size_t)r]);
        }
        gpsat_destroy(m->h[(size_t)r]);
    }
    cudaSetDevice(prev);
    delete m;
}

}  // extern "C"
