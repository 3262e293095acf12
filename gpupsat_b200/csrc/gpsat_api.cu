// C ABI of the device side (include/gpsat.h): handle management, uploads, launches, result gathering.
// One handle = one GPU = one host thread.  No CPU fallback: without a usable device every entry point fails with
// GPSAT_E_NO_DEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/gpsat.h"
#include "gpsat_device.h"
#include "host_formula.h"
#include "kernels.h"

namespace {

using gpsat_host::set_error;

#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                           \
            return (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver) ? GPSAT_E_NO_DEVICE \
                                                                                  : GPSAT_E_CUDA;    \
        }                                                                                            \
    } while (0)

// Device buffers are recycled through a process-wide cache, so that create/solve/destroy cycles — one per solve in
// the end-to-end path — do not pay cudaMalloc/cudaFree every time: the big ones (per-warp arenas, hand-off blocks,
// pools: hundreds of MB) because allocating them costs milliseconds, the small ones (a few dozen per handle) because
// cudaFree synchronises the device and, once NCCL has enabled peer access between the GPUs of a box, every
// cudaMalloc / cudaFree also maps / unmaps the block for the peers (bench.py at N = 8: 7 ms of a 32 ms end-to-end
// solve went into these calls).
struct CachedBlock {
    int device;
    void *p;
    size_t bytes;
};
std::vector<CachedBlock> g_block_cache;
std::mutex g_block_cache_mutex;
const size_t kCacheMaxBlocks = 160;

cudaError_t cached_alloc(void **p, size_t bytes, size_t *got, int *device)
{
    int dev = 0;
    cudaGetDevice(&dev);
    *device = dev;
    bytes = (bytes + 255) / 256 * 256;
    {
        std::lock_guard<std::mutex> lock(g_block_cache_mutex);
        size_t best = g_block_cache.size();
        for (size_t i = 0; i < g_block_cache.size(); i++) {
            const CachedBlock &b = g_block_cache[i];
            if (b.device == dev && b.bytes >= bytes && b.bytes <= 2 * bytes &&
                (best == g_block_cache.size() || b.bytes < g_block_cache[best].bytes))
                best = i;
        }
        if (best < g_block_cache.size()) {
            *p = g_block_cache[best].p;
            *got = g_block_cache[best].bytes;
            g_block_cache.erase(g_block_cache.begin() + (long)best);
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {   // out of memory: drop the cache and retry once
        std::lock_guard<std::mutex> lock(g_block_cache_mutex);
        for (auto &b : g_block_cache) cudaFree(b.p);
        g_block_cache.clear();
        cudaGetLastError();
        e = cudaMalloc(p, bytes);
    }
    *got = bytes;
    return e;
}

void cached_free(void *p, size_t bytes, int dev)
{
    if (!p) return;
    std::lock_guard<std::mutex> lock(g_block_cache_mutex);
    if (g_block_cache.size() >= kCacheMaxBlocks) {   // evict the block that has been unused for the longest
        cudaFree(g_block_cache.front().p);           // cudaFree works on a pointer of any device
        g_block_cache.erase(g_block_cache.begin());
    }
    g_block_cache.push_back({dev, p, bytes});        // the device the block was allocated on, not the current one
}

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    size_t bytes = 0;
    int device = 0;
    ~DevBuf() { release(); }
    void release()
    {
        if (p) cached_free(p, bytes, device);
        p = nullptr;
        n = 0;
        bytes = 0;
    }
    cudaError_t ensure(size_t count)
    {
        if (count <= n && p) return cudaSuccess;
        release();
        void *q = nullptr;
        size_t got = 0;
        cudaError_t e = cached_alloc(&q, std::max<size_t>(count, 1) * sizeof(T), &got, &device);
        if (e == cudaSuccess) {
            p = (T *)q;
            n = count;
            bytes = got;
        }
        return e;
    }
    cudaError_t upload(const T *src, size_t count, cudaStream_t s)
    {
        cudaError_t e = ensure(count);
        if (e != cudaSuccess || count == 0) return e;
        return cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, s);
    }
};

}  // namespace

struct gpsat {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaDeviceProp prop{};
    gpsat_opts opts{};
    gpsat_host::DeviceFormula D;
    std::vector<int32_t> coffsets_h;   // compact CSR offsets (int32) for the evaluation kernel / index lookups
    // device formula
    DevBuf<int32_t> cstart, cl2, ostart, occ2, vsids0, coffsets, clits;
    DevBuf<uint32_t> wbits0;
    DevBuf<uint8_t> val0;
    // occurrence-mode BCP (opts.bcp == GPSAT_BCP_OCCURRENCE)
    DevBuf<int32_t> occ_clause, occ_pair, orange;
    DevBuf<uint32_t> valbits_cta, occ_bucket;
    DevBuf<int32_t> cube_lits_sorted, cube_short;   // ternary sweep kernel: every cube's literals ordered by occurrence-count class
    int32_t tern_state_bytes = 0;
    DevBuf<int64_t> sweep_counters;
    int uniform3 = 0;
    // cubes
    int32_t n_cubes = 0;
    int32_t cube_base = 0;             // gpsat_propagate: index of the one cube the narrowed job list starts at
    bool cubes_set = false;
    std::vector<int64_t> cube_offsets_h;
    DevBuf<int64_t> cube_offsets;
    DevBuf<int32_t> cube_lits;
    // run buffers
    DevBuf<int32_t> ctrl;              // [0] next_job [1] stop_flag [2] sat_job
    DevBuf<unsigned long long> t0;
    DevBuf<uint8_t> model;
    DevBuf<gpsat_job_record> records;
    DevBuf<int32_t> implied, n_implied;
    DevBuf<int64_t> conflict_clause;
    DevBuf<int32_t> gstate;
    DevBuf<int32_t> pool, pool_cursor;
    // queue region: control block, ring of split-off cubes with their hand-off blocks, foreign pool, facts — ONE
    // allocation with the same layout on every rank, so that the other GPUs of a mesh can map it (gpsat_mesh_*)
    DevBuf<char> region;
    gpsat_mesh_layout ML{};
    bool region_full = false;          // false: control block only (propagate-only handles)
    int32_t hand_words = 0;
    DevBuf<int32_t> root_pending, root_flag, park, stage;
    // mesh (several GPUs as one work pool)
    int32_t mesh_ranks = 1, mesh_rank = 0;
    char *mesh_base[GPSAT_MESH_MAX_RANKS] = {nullptr};
    std::vector<int32_t> root_pending_h, root_flag_h;
    int32_t dq_ctrl_h[GPSAT_DQC_WORDS] = {0};   // control block after the last launch
    DevBuf<int32_t> arena;
    // geometry
    gpsat_state_layout Ly{};
    int blocks = 0, warps_per_block = 0, state_in_smem = 0, formula_in_smem = 0, formula_smem_words = 0;
    size_t smem_bytes = 0;
    int64_t arena_words = 0;
    // last run
    std::vector<gpsat_job_record> records_h;
    bool solving = false;
    int32_t run_mode = GPSAT_MODE_SOLVE;
    double kernel_ms = 0;
    int kernel_launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

namespace {

const int64_t kPoolWords = 1 << 22;   // 16 MB shared learnt pool per GPU (and as much for clauses from other GPUs)
const int64_t kPoolSlots = kPoolWords / GPSAT_POOL_SLOT_WORDS;

// every entry point runs on the handle's device (several handles on different GPUs may live in one process)
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard()
    {
        if (switched) cudaSetDevice(prev);
    }
};

// learnt-clause words a split-off cube inherits (plus the level-0 facts, which share the space)
int32_t hand_clause_words(const gpsat *h)
{
    const int32_t w = h->opts.split_hand_words > 0 ? h->opts.split_hand_words : 2048;   // measured best on C2
    return std::max(w, 2 * h->D.n_vars + 64);
}

int32_t *region_i32(gpsat *h, int64_t off) { return (int32_t *)(h->region.p + off); }
int32_t *dq_ctrl(gpsat *h) { return region_i32(h, h->ML.ctrl); }
int32_t *xpool(gpsat *h) { return h->region_full ? region_i32(h, h->ML.xpool) : nullptr; }
int32_t *xpool_cursor(gpsat *h) { return h->region_full ? region_i32(h, h->ML.xcur) : nullptr; }
uint8_t *facts(gpsat *h) { return h->region_full ? (uint8_t *)(h->region.p + h->ML.facts) : nullptr; }

// The queue region.  full = ring + hand-off blocks + foreign pool + facts (solve runs, exchange); otherwise only the
// control block.  A full region is never shrunk back, and a handle attached to a mesh keeps its region for life.
int ensure_region(gpsat *h, bool full)
{
    if (h->region.p && (h->region_full || !full)) return GPSAT_OK;
    if (h->mesh_ranks > 1) {
        set_error("the queue region of a handle attached to a mesh cannot change");
        return GPSAT_E_STATE;
    }
    gpsat_mesh_layout m;
    h->hand_words = (1 + 2 * h->D.n_vars + hand_clause_words(h) + 3) / 4 * 4;
    if (full) gpsat_make_mesh_layout(h->D.n_vars, h->hand_words, GPSAT_DQ_CAP, kPoolWords, &m);
    else gpsat_make_mesh_layout(0, 0, 0, 0, &m);
    h->region.release();
    CU(h->region.ensure((size_t)m.total));
    h->ML = m;
    h->region_full = full;
    CU(cudaMemsetAsync(h->region.p + m.ctrl, 0, GPSAT_DQC_WORDS * 4, h->stream));
    if (full) {
        CU(cudaMemsetAsync(h->region.p + m.xcur, 0, 64, h->stream));
        CU(cudaMemsetAsync(h->region.p + m.xpool, 0, (size_t)kPoolWords * 4, h->stream));
        CU(cudaMemsetAsync(h->region.p + m.facts, 0, (size_t)std::max(h->D.n_vars, 1), h->stream));
    }
    return GPSAT_OK;
}

int ensure_pools(gpsat *h, bool foreign)
{
    if (!h->pool.p) {
        CU(h->pool.ensure((size_t)kPoolWords));
        CU(h->pool_cursor.ensure(4));
        CU(cudaMemsetAsync(h->pool_cursor.p, 0, 4 * sizeof(int32_t), h->stream));
        CU(cudaMemsetAsync(h->pool.p, 0, (size_t)kPoolWords * sizeof(int32_t), h->stream));
    }
    if (foreign || !h->region_full) return ensure_region(h, true);
    return GPSAT_OK;
}

gpsat_formula_view make_view(gpsat *h)
{
    gpsat_formula_view F;
    F.n_vars = h->D.n_vars;
    F.n_clauses = (int32_t)h->D.n_clauses;
    F.n_lits = (int32_t)h->D.n_lits;
    F.wbits_words = (int32_t)h->D.wbits0.size();
    F.cstart = h->cstart.p;
    F.cl2 = h->cl2.p;
    F.ostart = h->ostart.p;
    F.occ2 = h->occ2.p;
    F.wbits0 = h->wbits0.p;
    F.vsids0 = h->vsids0.p;
    F.val0 = h->val0.p;
    return F;
}

int32_t default_max_learnts(int64_t n_clauses, int32_t refs_cap, int32_t n_vars)
{
    int64_t v = std::max<int64_t>(n_clauses / 3, 300);
    v = std::min<int64_t>(v, (int64_t)refs_cap - n_vars - 2);
    return (int32_t)std::max<int64_t>(v, 1);
}

gpsat_solve_params make_params(gpsat *h, int mode, int64_t implied_stride)
{
    gpsat_solve_params P;
    std::memset(&P, 0, sizeof(P));
    P.mode = mode;
    P.decision = h->opts.decision;
    P.bcp = h->opts.bcp;
    P.restart_first = h->opts.restart_first;
    P.restart_factor = h->opts.restart_factor;
    P.max_iterations = h->opts.max_iterations;
    P.stop_on_sat = h->opts.stop_on_sat;
    P.share_learnts = h->opts.share_learnts;
    P.share_max_len = std::min(h->opts.share_max_len, GPSAT_POOL_SLOT_WORDS - 1);
    P.learnt_refs_cap = 16384;
    if (P.learnt_refs_cap < h->D.n_vars + 64) P.learnt_refs_cap = h->D.n_vars + 64;
    P.max_learnts_first = default_max_learnts(h->D.n_clauses, P.learnt_refs_cap, h->D.n_vars);
    P.max_conflicts = h->opts.max_conflicts;
    P.arena_words = mode == GPSAT_MODE_SOLVE ? h->arena_words : 0;
    P.implied_stride = implied_stride;
    P.dynamic_split = (mode == GPSAT_MODE_SOLVE && h->opts.dynamic_split) ? 1 : 0;
    P.split_force = 0;
    P.split_gap = h->opts.split_gap > 0 ? h->opts.split_gap : 8;
    P.split_burst = h->opts.split_burst > 0 ? h->opts.split_burst : 4;
    P.share_import_max = h->opts.share_import_max > 0 ? h->opts.share_import_max : 256;
    P.split_gap_hot = h->opts.split_gap_hot > 0 ? h->opts.split_gap_hot : std::max(1, P.split_gap / 2);   // measured: DESIGN.md
    if (P.split_gap_hot > P.split_gap) P.split_gap_hot = P.split_gap;
    P.split_hot_demand = std::max(1, h->blocks * h->warps_per_block / 8);
    P.split_at_start = h->opts.split_at_start > 0 ? 1 : 0;
    P.mesh_flags = h->opts.mesh_flags;
    P.split_reserve = h->opts.split_reserve > 0 ? h->opts.split_reserve : (h->opts.split_reserve < 0 ? 0 : 32);   // measured: DESIGN.md section 3
    P.split_mode = h->opts.split_mode;
    P.phase_stats = (mode == GPSAT_MODE_SOLVE && h->opts.phase_stats) ? 1 : 0;
    P.split_min = h->opts.split_min > 0 ? h->opts.split_min : 0;
    P.split_hard = h->opts.split_hard > 0 ? h->opts.split_hard : 0x7fffffff;
    if (h->opts.max_learnts > 0)
        P.max_learnts_first = std::max(1, std::min(h->opts.max_learnts, P.learnt_refs_cap - h->D.n_vars - 2));
    return P;
}

// launch geometry: as many resident warps per SM as shared memory (per-job state) and registers allow
int plan_geometry(gpsat *h, int mode)
{
    const size_t smem_block_max = h->prop.sharedMemPerBlockOptin;                 // 227 KB on B200
    const size_t smem_sm = h->prop.sharedMemPerMultiprocessor;                    // 228 KB
    gpsat_geometry G;
    gpsat_plan_warps(h->D.n_vars, h->D.n_lits, h->D.n_clauses, h->opts.phase_stats, mode == GPSAT_MODE_SOLVE ? 1 : 0,
                     h->opts.warps_per_block, gpsat_kernels::cdcl_max_warps_per_block(), 24,
                     (int64_t)std::min(smem_block_max, smem_sm - 1024), &G);
    h->Ly = G.ly;
    h->state_in_smem = G.state_in_smem;
    h->formula_in_smem = G.formula_in_smem;
    h->formula_smem_words = G.formula_smem_words;
    h->warps_per_block = G.warps;
    h->smem_bytes = (size_t)G.smem_bytes;
    const int w = G.warps;
    int occ = 0;
    CU(gpsat_kernels::cdcl_occupancy(w, h->smem_bytes, h->state_in_smem != 0, h->formula_in_smem != 0, &occ));
    if (occ < 1) {
        set_error("kernel configuration does not fit on an SM");
        return GPSAT_E_CUDA;
    }
    int blocks = h->opts.blocks > 0 ? h->opts.blocks : h->prop.multiProcessorCount * occ;
    // no more warps than can find work (each warp owns an arena / state block): one per cube, or — when running cubes
    // hand sub-cubes to idle warps (dynamic_split) — up to eight per cube, so that a GPU that received fewer cubes than
    // it has warps (a shard of a multi-GPU run) still fills all its SMs
    const int64_t takers = (mode == GPSAT_MODE_SOLVE && h->opts.dynamic_split) ? 8 : 1;
    const int64_t need = (takers * (int64_t)std::max(h->n_cubes, 1) + w - 1) / w;
    // a GPU of a mesh also takes children of the other GPUs: it always launches every warp it can hold
    if (blocks > need && !(h->mesh_ranks > 1 && mode == GPSAT_MODE_SOLVE)) blocks = (int)need;
    h->blocks = std::max(blocks, 1);
    return GPSAT_OK;
}

int32_t roots_of(const gpsat *h) { return std::max(h->n_cubes, 1); }   // every rank of a mesh holds ALL cubes

int ensure_run_buffers(gpsat *h, int mode)
{
    const size_t n_warps = (size_t)h->blocks * h->warps_per_block;
    const size_t nr = (size_t)roots_of(h);
    const bool split = mode == GPSAT_MODE_SOLVE && h->opts.dynamic_split;
    CU(h->ctrl.ensure(4));
    CU(h->t0.ensure(2));   // [0] globaltimer stamp of the launch, [1] summed busy time of the warps
    CU(h->model.ensure((size_t)std::max(h->D.n_vars, 1)));
    CU(h->records.ensure(nr));
    // solve runs always get the full region: the exchange / pool calls of an epoch may come at any time afterwards
    int rc = ensure_region(h, mode == GPSAT_MODE_SOLVE || h->opts.share_learnts || h->mesh_ranks > 1);
    if (rc != GPSAT_OK) return rc;
    if (h->opts.share_learnts) {   // the pools exist only when clause sharing is on
        rc = ensure_pools(h, false);
        if (rc != GPSAT_OK) return rc;
    }
    CU(h->root_pending.ensure(nr));
    CU(h->root_flag.ensure(nr));
    if (split) CU(h->park.ensure(n_warps * (size_t)gpsat_park_words(h->D.n_vars)));
    if (split && h->mesh_ranks > 1) CU(h->stage.ensure(n_warps * (size_t)(h->hand_words + GPSAT_DQ_MAXK)));
    if (!h->state_in_smem) CU(h->gstate.ensure(n_warps * (size_t)h->Ly.total_words));
    if (mode == GPSAT_MODE_SOLVE) CU(h->arena.ensure(n_warps * (size_t)h->arena_words));
    return GPSAT_OK;
}

gpsat_run_buffers make_buffers(gpsat *h, int mode, double budget_ms)
{
    gpsat_run_buffers B;
    std::memset(&B, 0, sizeof(B));
    const bool split = mode == GPSAT_MODE_SOLVE && h->opts.dynamic_split && h->region_full;
    B.cube_offsets = h->cube_offsets.p;
    B.cube_lits = h->cube_lits.p;
    B.n_cubes = h->n_cubes;
    // the root cursor: this handle's control block, or — mesh — rank 0's, which every GPU advances over NVLink
    B.next_job = (mode == GPSAT_MODE_SOLVE && h->mesh_ranks > 1 ? (int32_t *)(h->mesh_base[0] + h->ML.ctrl) : dq_ctrl(h)) + GPSAT_DQC_CURSOR;
    B.stop_flag = dq_ctrl(h) + GPSAT_DQC_STOP;
    B.sat_job = h->ctrl.p + 2;
    B.model = h->model.p;
    B.records = h->records.p;
    B.arena = mode == GPSAT_MODE_SOLVE ? h->arena.p : nullptr;
    B.gstate = h->gstate.p;
    B.pool = h->pool.p;
    B.pool_cursor = h->pool_cursor.p;
    B.pool_cap_words = (int32_t)kPoolWords;
    B.xpool = xpool(h);
    B.xpool_cursor = xpool_cursor(h);
    B.facts = (h->opts.share_learnts || h->mesh_ranks > 1) ? facts(h) : nullptr;
    B.state_in_smem = h->state_in_smem;
    B.formula_in_smem = h->formula_in_smem;
    B.formula_smem_words = h->formula_smem_words;
    B.dq_ctrl = dq_ctrl(h);
    B.dq_lits = split ? region_i32(h, h->ML.lits) : nullptr;
    B.dq_meta = split ? region_i32(h, h->ML.meta) : nullptr;
    B.dq_hand = split ? region_i32(h, h->ML.hand) : nullptr;
    B.root_pending = h->root_pending.p;
    B.hand_words = h->hand_words;
    B.dq_cap = GPSAT_DQ_CAP;
    B.root_flag = h->root_flag.p;
    B.t0 = h->t0.p;
    B.busy_ns = (long long *)(h->t0.p + 1);
    B.park = split ? h->park.p : nullptr;
    B.park_words = gpsat_park_words(h->D.n_vars);
    B.budget_ns = budget_ms > 0 ? (unsigned long long)(budget_ms * 1e6) : 0ull;
    B.mesh_ranks = mode == GPSAT_MODE_SOLVE ? h->mesh_ranks : 1;
    B.mesh_rank = h->mesh_rank;
    for (int r = 0; r < GPSAT_MESH_MAX_RANKS; r++) B.mesh_base[r] = r < h->mesh_ranks ? h->mesh_base[r] : nullptr;
    B.mesh_off_ctrl = h->ML.ctrl;
    B.mesh_off_meta = h->ML.meta;
    B.mesh_off_lits = h->ML.lits;
    B.mesh_off_hand = h->ML.hand;
    B.mesh_off_xcur = h->ML.xcur;
    B.mesh_off_xpool = h->ML.xpool;
    B.mesh_off_facts = h->ML.facts;
    B.stage = (split && h->mesh_ranks > 1) ? h->stage.p : nullptr;
    B.xpool_cap_slots = (int32_t)kPoolSlots;
    B.mesh_n_vars = h->D.n_vars;
    return B;
}

// queue state, root arrays and run counters of a new run: one small kernel + memsets, nothing from pageable memory
int reset_ctrl(gpsat *h)
{
    const size_t nr = (size_t)roots_of(h);
    CU(cudaMemsetAsync(h->records.p, 0, nr * sizeof(gpsat_job_record), h->stream));
    if (h->park.p) CU(cudaMemsetAsync(h->park.p, 0, h->park.n * sizeof(int32_t), h->stream));
    CU(gpsat_kernels::launch_queue_init(dq_ctrl(h), h->region_full ? region_i32(h, h->ML.meta) : nullptr, GPSAT_DQ_CAP,
                                        h->root_pending.p, h->root_flag.p, (int)nr,
                                        (h->mesh_ranks > 1 && h->mesh_rank != 0) ? 0 : 1, xpool_cursor(h), facts(h),
                                        h->D.n_vars, h->ctrl.p, h->t0.p, h->stream));
    if (h->region_full && (h->opts.share_learnts || h->mesh_ranks > 1))   // foreign slots must read "empty" (length 0)
        CU(cudaMemsetAsync(xpool(h), 0, (size_t)kPoolWords * sizeof(int32_t), h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return GPSAT_OK;
}

int launch_timed(gpsat *h, const gpsat_solve_params &P, const gpsat_run_buffers &B)
{
    CU(gpsat_kernels::launch_stamp(h->t0.p, h->stream));
    CU(cudaEventRecord(h->ev0, h->stream));
    CU(gpsat_kernels::launch_cdcl(make_view(h), P, h->Ly, B, h->blocks, h->warps_per_block, h->smem_bytes, h->stream));
    CU(cudaEventRecord(h->ev1, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->kernel_ms += ms;
    h->kernel_launches += 1;
    return GPSAT_OK;
}

int fetch_records(gpsat *h)
{
    const size_t nc = (size_t)roots_of(h);
    h->records_h.resize(nc);
    h->root_pending_h.resize(nc);
    h->root_flag_h.resize(nc);
    CU(cudaMemcpyAsync(h->records_h.data(), h->records.p, nc * sizeof(gpsat_job_record), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(h->root_pending_h.data(), h->root_pending.p, nc * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(h->root_flag_h.data(), h->root_flag.p, nc * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(h->dq_ctrl_h, dq_ctrl(h), sizeof(h->dq_ctrl_h), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    // the status of an original cube is decided by all jobs that descend from it (dynamic splitting); on a mesh rank
    // these are this GPU's contributions only, merged over ranks by gpsat_mesh_results_*
    for (size_t j = 0; j < nc; j++) h->records_h[j].status = gpsat_root_status(h->root_flag_h[j], h->root_pending_h[j]);
    return GPSAT_OK;
}

void fill_stats(gpsat *h, gpsat_stats *s)
{
    if (!s) return;
    std::memset(s, 0, sizeof(*s));
    const size_t nc = std::min(h->records_h.size(), (size_t)roots_of(h));
    s->jobs_total = (int64_t)nc;
    for (size_t j = 0; j < nc; j++) {
        const gpsat_job_record &r = h->records_h[j];
        // counters include cubes that are still open (parked between steps, or cut short by the stop flag)
        if (r.status == GPSAT_SAT) s->jobs_sat++;
        else if (r.status == GPSAT_UNSAT) s->jobs_unsat++;
        else if (r.status != GPSAT_JOB_NOT_RUN) s->jobs_undef++;
        if (r.status != GPSAT_JOB_ABORTED && r.status != GPSAT_JOB_NOT_RUN) s->jobs_done++;
        s->decisions += r.decisions;
        s->implications += r.implications;
        s->conflicts += r.conflicts;
        s->learnt_clauses += r.learnt_clauses;
        s->learnt_literals += r.learnt_literals;
        s->restarts += r.restarts;
        s->watchers_visited += r.watchers_visited;
        s->clause_words_read += r.clause_words_read;
        s->splits += r.reserved;
    }
    int32_t cur[2] = {0, 0}, xcur[2] = {0, 0};
    if (h->pool_cursor.p) cudaMemcpy(cur, h->pool_cursor.p, sizeof(cur), cudaMemcpyDeviceToHost);
    if (xpool_cursor(h)) cudaMemcpy(xcur, xpool_cursor(h), sizeof(xcur), cudaMemcpyDeviceToHost);
    s->pool_clauses = cur[1];
    s->foreign_clauses = std::min<int64_t>(xcur[0], kPoolSlots);
    unsigned long long busy = 0;
    if (h->t0.p && h->run_mode == GPSAT_MODE_SOLVE) cudaMemcpy(&busy, h->t0.p + 1, sizeof(busy), cudaMemcpyDeviceToHost);
    const double warp_ms = (double)h->blocks * h->warps_per_block * h->kernel_ms;
    s->warp_busy_frac = warp_ms > 0 ? (double)busy * 1e-6 / warp_ms : 0.0;
    s->kernel_ms = h->kernel_ms;
    s->kernel_launches = h->kernel_launches;
    s->blocks = h->blocks;
    s->warps_per_block = h->warps_per_block;
    s->smem_bytes_per_block = (int32_t)h->smem_bytes;
    s->state_in_smem = h->state_in_smem;
    s->steals = h->dq_ctrl_h[GPSAT_DQC_STEALS];
}

// verdict of the whole run from the per-job records (≙ Results::get_status after parallel_kernel_retrieve_results)
int32_t run_verdict(gpsat *h, bool *all_done)
{
    bool any_sat = false, any_open = false, any_undef = false;
    const size_t nc = std::min(h->records_h.size(), (size_t)roots_of(h));
    for (size_t j = 0; j < nc; j++) {
        const int st = h->records_h[j].status;
        if (st == GPSAT_SAT) any_sat = true;
        else if (st == GPSAT_UNSAT) continue;
        else if (st == GPSAT_JOB_NOT_RUN || st == GPSAT_JOB_ABORTED) any_open = true;
        else any_undef = true;   // UNDEF (cap) or OOM
    }
    if (all_done) *all_done = !any_open;
    if (any_sat) return GPSAT_SAT;
    if (any_open || any_undef) return GPSAT_UNDEF;
    return GPSAT_UNSAT;
}

// occurrence-list BCP for large clause databases (GPSAT_BCP_OCCURRENCE): picks and launches one of the sweep kernels
int propagate_all_occurrence(gpsat *h, int32_t *status, int32_t *n_implied, int32_t *implied, int64_t implied_stride,
                             int64_t *conflict_clause, gpsat_job_record *records)
{
    const size_t nc = (size_t)h->n_cubes;
    if (implied_stride <= 0) implied_stride = h->D.n_vars;       // the implied block doubles as the trail
    const int32_t val_words = (h->D.n_vars + 15) / 16;
    // Kernel choice (measured on C4):
    //   1. ternary kernel — pure 3-SAT whose base-3 state fits one SM (gpsat_create built the bucket index);
    //   2. one CTA per job, assigned-bit filter in shared memory + 2-bit values in an L2-resident global block — any
    //      clause lengths, any variable count (the filter is exact up to 2^20 variables, aliased beyond).
    // opts.sweep_flags (test hook): 1 = general kernel even for pure 3-SAT, bits 8.. = log2 of the filter size.
    const bool use_tern = h->tern_state_bytes > 0 && !(h->opts.sweep_flags & 1);
    int cluster = -1, slice_log2 = 10, cthreads = 1024;
    {
        int lg = 10;
        while (((int64_t)1 << lg) < h->D.n_vars) lg++;
        while (lg > 10 && ((size_t)1 << (lg - 3)) + 1024 > h->prop.sharedMemPerBlockOptin) lg--;   // aliased filter
        const int forced = (h->opts.sweep_flags >> 8) & 31;
        if (forced >= 10 && forced <= lg) lg = forced;
        slice_log2 = lg;
    }
    int wpb = cthreads / 32;
    int blocks = h->prop.multiProcessorCount;
    int32_t cta_val_words = 0;
    if (!use_tern) {
        int per_sm = 0;
        CU(gpsat_kernels::sweep_cta_capacity(slice_log2, cthreads, 1, &per_sm));
        if (per_sm < 1) {
            set_error("sweep kernel configuration does not fit on an SM");
            return GPSAT_E_CUDA;
        }
        blocks = h->prop.multiProcessorCount * per_sm;
        if (h->opts.blocks > 0) blocks = std::min(blocks, h->opts.blocks);
        if ((int64_t)blocks > (int64_t)nc) blocks = (int)std::max<size_t>(nc, 1);
        cta_val_words = (val_words + 3) / 4 * 4;
        CU(h->valbits_cta.ensure((size_t)blocks * (size_t)cta_val_words));   // zeroed by the kernel per job
    }
    CU(h->ctrl.ensure(4));
    CU(h->implied.ensure(nc * (size_t)implied_stride));
    CU(h->n_implied.ensure(nc));
    CU(h->conflict_clause.ensure(nc));
    CU(h->records.ensure(nc));           // reused as int32 status scratch below
    CU(h->sweep_counters.ensure(2 * nc));
    DevBuf<int32_t> d_status;
    CU(d_status.ensure(nc));
    CU(cudaMemsetAsync(h->ctrl.p, 0, 4 * sizeof(int32_t), h->stream));
    gpsat_kernels::SweepLaunch L;
    L.n_vars = h->D.n_vars;
    L.n_clauses = (int32_t)h->D.n_clauses;
    L.n_cubes = h->n_cubes;
    L.uniform3 = h->uniform3;
    L.ostart = h->orange.p;
    L.occ_clause = h->occ_clause.p;
    L.occ_pair = h->occ_pair.p;
    L.coffsets = h->coffsets.p;
    L.clits = h->clits.p;
    L.cube_offsets = h->cube_offsets.p;
    L.cube_lits = h->cube_lits.p;
    L.valbits = h->valbits_cta.p;
    L.val_words = cta_val_words;
    L.implied = h->implied.p;
    L.stride = implied_stride;
    L.n_implied = h->n_implied.p;
    L.status = d_status.p;
    L.conflict_clause = h->conflict_clause.p;
    L.counters = h->sweep_counters.p;
    L.next_job = h->ctrl.p;
    L.blocks = blocks;
    L.warps_per_block = wpb;
    L.cluster_size = cluster;
    // measured (C4, 1184 jobs): CTA-filter kernel — evict-first index loads 12.9 -> 11.2 ms, L2 persistence of the value
    // blocks 11.0 ms; ternary kernel — 2 x 32-byte read-only bucket loads
    L.stream_index = use_tern ? 2 : 1;
    if (!use_tern) {
        // keep the per-CTA value blocks resident in L2 while the occurrence index streams through it
        const size_t bytes = (size_t)blocks * (size_t)cta_val_words * sizeof(uint32_t);
        int max_win = 0, max_persist = 0;
        cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, h->device);
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, h->device);
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min<size_t>(bytes, (size_t)max_persist));
        cudaStreamAttrValue av;
        std::memset(&av, 0, sizeof(av));
        av.accessPolicyWindow.base_ptr = h->valbits_cta.p;
        av.accessPolicyWindow.num_bytes = std::min<size_t>(bytes, (size_t)max_win);
        av.accessPolicyWindow.hitRatio = 1.0f;
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &av);
        cudaGetLastError();
    }
    L.slice_log2 = slice_log2;
    if (use_tern) {   // one CTA per SM, whole job state in shared memory: nothing else to size
        L.bucket = h->occ_bucket.p;
        L.tern_state_bytes = h->tern_state_bytes;
        L.l2_prefetch = (h->opts.sweep_flags & 4) ? 0 : 1;       // prefetch.global.L2 of the next batch's buckets: 4.12 -> 4.08 ms at L = 1e5, 31.9 -> 30.7 ms at 2e5
        if (h->cube_lits_sorted.p)
            L.cube_lits = h->cube_lits_sorted.p, L.cube_short = h->cube_short.p + h->cube_base;   // per-cube info follows the narrowed job list
        L.tern_prefetch = (h->opts.sweep_flags & 2) ? 1 : 0;     // measured: 5.59 ms without, 5.67 ms with the bucket fetched one batch ahead in registers
        L.tern_first_hit = (h->opts.sweep_flags & 64) ? 0 : 1;    // first hit of a bucket kept during the scan: 4.07 -> 3.91 ms at L = 1e5 (64 switches it off)
        // test hook 32: lane-private code table (no bank conflicts on the second lookup) when it fits beside the state
        L.tern_private_lut = ((h->opts.sweep_flags & 32) &&
                              gpsat_kernels::tern_smem_bytes(h->tern_state_bytes, true) + 256 <= h->prop.sharedMemPerBlockOptin) ? 1 : 0;
        blocks = h->prop.multiProcessorCount;
        if (h->opts.blocks > 0) blocks = std::min(blocks, h->opts.blocks);
        if ((int64_t)blocks > (int64_t)nc) blocks = (int)std::max<size_t>(nc, 1);
        wpb = 32;
        L.blocks = blocks;
        L.warps_per_block = wpb;
    }
    L.cluster_size = cluster;
    CU(cudaMemsetAsync(h->sweep_counters.p, 0, 2 * nc * sizeof(int64_t), h->stream));
    h->kernel_ms = 0;
    h->kernel_launches = 0;
    h->blocks = blocks;
    h->warps_per_block = wpb;
    h->smem_bytes = use_tern ? gpsat_kernels::tern_smem_bytes(h->tern_state_bytes, L.tern_private_lut != 0) : ((size_t)1 << (slice_log2 - 3));
    h->state_in_smem = 1;
    CU(cudaEventRecord(h->ev0, h->stream));
    CU(gpsat_kernels::launch_bcp_sweep(L, h->stream));
    CU(cudaEventRecord(h->ev1, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->kernel_ms = ms;
    h->kernel_launches = 1;
    std::vector<int32_t> st(nc), ni(nc);
    std::vector<int64_t> cnt(2 * nc);
    CU(cudaMemcpy(st.data(), d_status.p, nc * sizeof(int32_t), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(ni.data(), h->n_implied.p, nc * sizeof(int32_t), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(cnt.data(), h->sweep_counters.p, 2 * nc * sizeof(int64_t), cudaMemcpyDeviceToHost));
    h->records_h.assign(nc, gpsat_job_record());
    for (size_t j = 0; j < nc; j++) {
        gpsat_job_record &r = h->records_h[j];
        std::memset(&r, 0, sizeof(r));
        r.status = st[j];
        r.implications = ni[j];
        r.conflicts = st[j] == GPSAT_UNSAT ? 1 : 0;
        r.watchers_visited = cnt[2 * j];
        r.clause_words_read = cnt[2 * j + 1];
    }
    h->run_mode = GPSAT_MODE_PROPAGATE;
    if (status) std::memcpy(status, st.data(), nc * sizeof(int32_t));
    if (n_implied) std::memcpy(n_implied, ni.data(), nc * sizeof(int32_t));
    if (records) std::memcpy(records, h->records_h.data(), nc * sizeof(gpsat_job_record));
    if (conflict_clause)
        CU(cudaMemcpy(conflict_clause, h->conflict_clause.p, nc * sizeof(int64_t), cudaMemcpyDeviceToHost));
    if (implied)
        CU(cudaMemcpy(implied, h->implied.p, nc * (size_t)implied_stride * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return GPSAT_OK;
}

}  // namespace

extern "C" {

void gpsat_opts_default(gpsat_opts *o)
{
    if (!o) return;
    std::memset(o, 0, sizeof(*o));
    o->struct_size = (int32_t)sizeof(gpsat_opts);
    o->device = -1;
    o->decision = GPSAT_DECIDE_VSIDS;
    o->bcp = GPSAT_BCP_WATCHED;
    o->restart_first = 100;
    o->restart_factor = 1.3f;
    o->max_iterations = 0;
    o->stop_on_sat = 1;
    o->max_conflicts = 0;
    o->share_learnts = 0;
    o->share_max_len = 8;
    o->warps_per_block = 0;
    o->blocks = 0;
    o->arena_words = 0;
    o->dynamic_split = 1;
}

int gpsat_create(gpsat_t **out, int32_t n_vars, int64_t n_clauses, const int64_t *offsets, const int32_t *lits,
                 const gpsat_opts *opts)
{
    if (!out) {
        set_error("null handle pointer");
        return GPSAT_E_ARG;
    }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error(std::string("no usable CUDA device (") + cudaGetErrorString(e) +
                  "); gpupsat_b200 has no CPU fallback");
        return GPSAT_E_NO_DEVICE;
    }
    gpsat *h = new gpsat();
    if (opts) {
        if (opts->struct_size != (int32_t)sizeof(gpsat_opts)) {
            delete h;
            set_error("gpsat_opts.struct_size mismatch: initialise with gpsat_opts_default");
            return GPSAT_E_ARG;
        }
        h->opts = *opts;
    } else {
        gpsat_opts_default(&h->opts);
    }
    int rc = gpsat_host::build_device_formula(n_vars, n_clauses, offsets, lits, h->D);
    if (rc != GPSAT_OK) {
        delete h;
        return rc;
    }
    auto fail = [&](int code) {
        gpsat_destroy(h);
        return code;
    };
#define CUH(call)                                                              \
    do {                                                                       \
        cudaError_t e_ = (call);                                               \
        if (e_ != cudaSuccess) {                                               \
            set_error(std::string(#call) + ": " + cudaGetErrorString(e_));     \
            return fail(GPSAT_E_CUDA);                                         \
        }                                                                      \
    } while (0)
    if (h->opts.device >= 0) {
        CUH(cudaSetDevice(h->opts.device));
        h->device = h->opts.device;
    } else {
        CUH(cudaGetDevice(&h->device));
    }
    {   // cudaGetDeviceProperties costs milliseconds (and varies): query the three attributes used, once per device
        static std::mutex attr_mutex;
        static cudaDeviceProp cached[64];
        static bool have[64] = {false};
        std::lock_guard<std::mutex> lock(attr_mutex);
        const int slot = h->device >= 0 && h->device < 64 ? h->device : 63;
        if (!have[slot] || slot == 63) {
            int v = 0;
            CUH(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, h->device));
            cached[slot].multiProcessorCount = v;
            CUH(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
            cached[slot].sharedMemPerBlockOptin = (size_t)v;
            CUH(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerMultiprocessor, h->device));
            cached[slot].sharedMemPerMultiprocessor = (size_t)v;
            have[slot] = true;
        }
        h->prop = cached[slot];
    }
    CUH(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CUH(cudaEventCreate(&h->ev0));
    CUH(cudaEventCreate(&h->ev1));

    // compact CSR with int32 offsets (clause evaluation kernel)
    h->coffsets_h.resize((size_t)n_clauses + 1);
    const int64_t base = n_clauses ? offsets[0] : 0;
    for (int64_t c = 0; c <= n_clauses; c++) h->coffsets_h[(size_t)c] = (int32_t)((n_clauses ? offsets[c] : 0) - base);

    CUH(h->cstart.upload(h->D.cstart.data(), h->D.cstart.size(), h->stream));
    CUH(h->cl2.upload(h->D.cl2.data(), h->D.cl2.size(), h->stream));
    CUH(h->ostart.upload(h->D.ostart.data(), h->D.ostart.size(), h->stream));
    CUH(h->occ2.upload(h->D.occ2.data(), h->D.occ2.size(), h->stream));
    CUH(h->wbits0.upload(h->D.wbits0.data(), h->D.wbits0.size(), h->stream));
    CUH(h->vsids0.upload(h->D.vsids0.data(), h->D.vsids0.size(), h->stream));
    CUH(h->val0.upload(h->D.val0.data(), h->D.val0.size(), h->stream));
    CUH(h->coffsets.upload(h->coffsets_h.data(), h->coffsets_h.size(), h->stream));
    CUH(h->clits.upload(lits + base, (size_t)h->D.n_lits, h->stream));
    if (h->opts.bcp == GPSAT_BCP_OCCURRENCE) {
        // Occurrence index of the sweep kernels (host_formula.cpp: build_sweep_index).  The bucket index and with it the
        // ternary kernel (gpsat_bcp_sweep_tern_kernel) are used for pure 3-SAT whose literal ids fit 21 bits and whose
        // base-3 state (five variables per byte) fits one SM's shared memory beside the lookup table.
        const int32_t state_bytes = (int32_t)((((int64_t)n_vars + 1 + 4) / 5 + 15) / 16 * 16);
        const bool want_buckets = !(h->opts.sweep_flags & 1) &&
                                  gpsat_kernels::tern_smem_bytes(state_bytes) + 256 <= h->prop.sharedMemPerBlockOptin;
        gpsat_host::SweepIndex X;
        gpsat_host::build_sweep_index(h->D, want_buckets, X);
        h->uniform3 = X.uniform3;
        h->tern_state_bytes = 0;
        CUH(h->orange.upload(X.orange.data(), X.orange.size(), h->stream));
        CUH(h->occ_clause.upload(X.occ_clause.data(), X.occ_clause.size(), h->stream));
        if (X.uniform3) CUH(h->occ_pair.upload(X.occ_pair.data(), X.occ_pair.size(), h->stream));
        if (!X.bucket.empty()) {
            CUH(h->occ_bucket.upload(X.bucket.data(), X.bucket.size(), h->stream));
            h->tern_state_bytes = state_bytes;
        }
        CUH(cudaStreamSynchronize(h->stream));
    }
    CUH(cudaStreamSynchronize(h->stream));
#undef CUH

    // per-warp learnt arena: header (watch-vector heads, histogram, clause list) + room for clauses and watches
    const int64_t header = 6 * (int64_t)n_vars + 64 + std::max<int64_t>(16384, n_vars + 64);
    h->arena_words = h->opts.arena_words > 0 ? h->opts.arena_words : std::max<int64_t>((int64_t)1 << 19, 4 * header);
    if (h->arena_words < header + 1024) h->arena_words = header + 1024;
    if (h->arena_words >= ((int64_t)1 << 31)) h->arena_words = ((int64_t)1 << 31) - 1;

    // default: the single empty cube (sequential mode)
    h->n_cubes = 1;
    h->cube_offsets_h.assign(2, 0);
    h->cubes_set = false;
    *out = h;
    return GPSAT_OK;
}

void gpsat_destroy(gpsat_t *h)
{
    if (!h) return;
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int gpsat_set_cubes(gpsat_t *h, int32_t n_cubes, const int64_t *cube_offsets, const int32_t *cube_lits)
{
    if (!h || n_cubes < 0 || (n_cubes > 0 && !cube_offsets)) {
        set_error("bad arguments");
        return GPSAT_E_ARG;
    }
    DeviceGuard guard(h->device);
    h->cube_lits_sorted.release();
    h->cube_short.release();
    if (n_cubes == 0) {
        h->n_cubes = 1;
        h->cube_offsets_h.assign(2, 0);
    } else {
        h->n_cubes = n_cubes;
        h->cube_offsets_h.assign(cube_offsets, cube_offsets + n_cubes + 1);
        const int64_t base = cube_offsets[0];
        for (auto &o : h->cube_offsets_h) o -= base;
        const int64_t total = h->cube_offsets_h.back();
        for (int32_t j = 0; j < n_cubes; j++)
            if (h->cube_offsets_h[(size_t)j + 1] < h->cube_offsets_h[(size_t)j]) {
                set_error("cube offsets not monotone");
                return GPSAT_E_ARG;
            }
        if (total > 0 && !cube_lits) {
            set_error("null cube literals");
            return GPSAT_E_ARG;
        }
        for (int64_t i = 0; i < total; i++) {
            const int32_t x = cube_lits[base + i];
            if (x < 0 || (x >> 1) >= h->D.n_vars) {
                set_error("cube literal out of range");
                return GPSAT_E_ARG;
            }
        }
        CU(h->cube_lits.upload(cube_lits + base, (size_t)total, h->stream));
        if (h->tern_state_bytes > 0) {
            // Ternary sweep kernel: every cube's literals ordered by the length of the list they make the kernel visit
            // (host_formula.cpp: order_cubes_for_sweep).  BCP is confluent: status and implied set do not depend on the
            // order in which a cube's literals are visited.
            std::vector<int32_t> sorted, n_short;
            gpsat_host::order_cubes_for_sweep(h->D, n_cubes, cube_offsets, cube_lits, sorted, n_short);
            CU(h->cube_lits_sorted.upload(sorted.data(), sorted.size(), h->stream));
            CU(h->cube_short.upload(n_short.data(), n_short.size(), h->stream));
            CU(cudaStreamSynchronize(h->stream));
        }
    }
    CU(h->cube_offsets.upload(h->cube_offsets_h.data(), h->cube_offsets_h.size(), h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->cubes_set = true;
    return GPSAT_OK;
}

static int ensure_cubes(gpsat_t *h)
{
    if (h->cubes_set) return GPSAT_OK;
    return gpsat_set_cubes(h, 0, nullptr, nullptr);
}

int gpsat_propagate_all(gpsat_t *h, int32_t *status, int32_t *n_implied, int32_t *implied, int64_t implied_stride,
                        int64_t *conflict_clause, gpsat_job_record *records)
{
    if (!h || implied_stride < 0) {
        set_error("bad arguments");
        return GPSAT_E_ARG;
    }
    DeviceGuard guard(h->device);
    int rc = ensure_cubes(h);
    if (rc != GPSAT_OK) return rc;
    if (h->opts.bcp == GPSAT_BCP_OCCURRENCE)
        return propagate_all_occurrence(h, status, n_implied, implied, implied_stride, conflict_clause, records);
    rc = plan_geometry(h, GPSAT_MODE_PROPAGATE);
    if (rc != GPSAT_OK) return rc;
    rc = ensure_run_buffers(h, GPSAT_MODE_PROPAGATE);
    if (rc != GPSAT_OK) return rc;
    const size_t nc = (size_t)h->n_cubes;
    if (implied) {
        CU(h->implied.ensure(nc * (size_t)implied_stride));
        CU(cudaMemsetAsync(h->implied.p, 0xFF, nc * (size_t)implied_stride * sizeof(int32_t), h->stream));   // -1 padding
    }
    CU(h->n_implied.ensure(nc));
    CU(h->conflict_clause.ensure(nc));
    rc = reset_ctrl(h);
    if (rc != GPSAT_OK) return rc;
    gpsat_solve_params P = make_params(h, GPSAT_MODE_PROPAGATE, implied ? implied_stride : 0);
    gpsat_run_buffers B = make_buffers(h, GPSAT_MODE_PROPAGATE, 0);
    B.implied = implied ? h->implied.p : nullptr;
    B.n_implied = h->n_implied.p;
    B.conflict_clause = h->conflict_clause.p;
    h->kernel_ms = 0;
    h->kernel_launches = 0;
    h->run_mode = GPSAT_MODE_PROPAGATE;
    rc = launch_timed(h, P, B);
    if (rc != GPSAT_OK) return rc;
    rc = fetch_records(h);
    if (rc != GPSAT_OK) return rc;
    if (status)
        for (size_t j = 0; j < nc; j++) status[j] = h->records_h[j].status;
    if (records) std::memcpy(records, h->records_h.data(), nc * sizeof(gpsat_job_record));
    if (n_implied) CU(cudaMemcpy(n_implied, h->n_implied.p, nc * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (conflict_clause)
        CU(cudaMemcpy(conflict_clause, h->conflict_clause.p, nc * sizeof(int64_t), cudaMemcpyDeviceToHost));
    if (implied && implied_stride > 0)
        CU(cudaMemcpy(implied, h->implied.p, nc * (size_t)implied_stride * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return GPSAT_OK;
}

int gpsat_propagate(gpsat_t *h, int32_t cube, int32_t *status, int32_t *implied, int32_t *n_implied,
                    int64_t *conflict_clause)
{
    if (!h) {
        set_error("null handle");
        return GPSAT_E_ARG;
    }
    int rc = ensure_cubes(h);
    if (rc != GPSAT_OK) return rc;
    if (cube < 0 || cube >= h->n_cubes) {
        set_error("cube index out of range");
        return GPSAT_E_ARG;
    }
    // run just this cube: temporarily narrow the job list to [cube, cube+1)
    const int32_t saved_n = h->n_cubes;
    int64_t *saved_off = h->cube_offsets.p;
    h->cube_offsets.p = saved_off + cube;
    h->n_cubes = 1;
    h->cube_base = cube;
    std::vector<int32_t> imp((size_t)std::max(h->D.n_vars, 1));
    int32_t st = GPSAT_UNDEF, n = 0;
    int64_t cc = -1;
    rc = gpsat_propagate_all(h, &st, &n, implied ? imp.data() : nullptr, h->D.n_vars, &cc, nullptr);
    h->cube_offsets.p = saved_off;
    h->n_cubes = saved_n;
    h->cube_base = 0;
    if (rc != GPSAT_OK) return rc;
    if (status) *status = st;
    if (n_implied) *n_implied = n;
    if (conflict_clause) *conflict_clause = cc;
    if (implied) std::memcpy(implied, imp.data(), (size_t)std::min(n, h->D.n_vars) * sizeof(int32_t));
    return GPSAT_OK;
}

int gpsat_eval_clauses(gpsat_t *h, int32_t n_assignments, const uint8_t *assignment, int32_t *status_per_clause,
                       int32_t *unit_lit)
{
    if (!h || n_assignments < 0 || (n_assignments > 0 && (!assignment || !status_per_clause))) {
        set_error("bad arguments");
        return GPSAT_E_ARG;
    }
    DeviceGuard guard(h->device);
    const size_t na = (size_t)n_assignments, nv = (size_t)h->D.n_vars, nc = (size_t)h->D.n_clauses;
    if (na == 0 || nc == 0) return GPSAT_OK;
    DevBuf<uint8_t> d_as;
    DevBuf<int32_t> d_st, d_un;
    CU(d_as.upload(assignment, na * nv, h->stream));
    CU(d_st.ensure(na * nc));
    if (unit_lit) CU(d_un.ensure(na * nc));
    CU(gpsat_kernels::launch_eval_clauses(h->D.n_vars, (int32_t)h->D.n_clauses, h->coffsets.p, h->clits.p, n_assignments,
                                          d_as.p, d_st.p, unit_lit ? d_un.p : nullptr, h->stream));
    CU(cudaMemcpyAsync(status_per_clause, d_st.p, na * nc * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (unit_lit) CU(cudaMemcpyAsync(unit_lit, d_un.p, na * nc * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return GPSAT_OK;
}

int gpsat_solve_begin(gpsat_t *h)
{
    if (!h) {
        set_error("null handle");
        return GPSAT_E_ARG;
    }
    DeviceGuard guard(h->device);
    int rc = ensure_cubes(h);
    if (rc != GPSAT_OK) return rc;
    rc = plan_geometry(h, GPSAT_MODE_SOLVE);
    if (rc != GPSAT_OK) return rc;
    rc = ensure_run_buffers(h, GPSAT_MODE_SOLVE);
    if (rc != GPSAT_OK) return rc;
    // every solve starts without shared knowledge: clauses of an earlier run on this handle would be cached work
    // (reset_ctrl clears the foreign pool, its cursor and the facts)
    if (h->pool.p) {
        CU(cudaMemsetAsync(h->pool_cursor.p, 0, 4 * sizeof(int32_t), h->stream));
        CU(cudaMemsetAsync(h->pool.p, 0, (size_t)kPoolWords * sizeof(int32_t), h->stream));
    }
    rc = reset_ctrl(h);
    if (rc != GPSAT_OK) return rc;
    h->kernel_ms = 0;
    h->kernel_launches = 0;
    h->run_mode = GPSAT_MODE_SOLVE;
    h->solving = true;
    return GPSAT_OK;
}

int gpsat_solve_step(gpsat_t *h, double budget_ms, int32_t *done, int32_t *verdict)
{
    if (!h || !h->solving) {
        set_error("gpsat_solve_step without gpsat_solve_begin");
        return GPSAT_E_STATE;
    }
    DeviceGuard guard(h->device);
    gpsat_solve_params P = make_params(h, GPSAT_MODE_SOLVE, 0);
    gpsat_run_buffers B = make_buffers(h, GPSAT_MODE_SOLVE, budget_ms);
    int rc = launch_timed(h, P, B);
    if (rc != GPSAT_OK) return rc;
    rc = fetch_records(h);
    if (rc != GPSAT_OK) return rc;
    bool all_done = false;
    int32_t v = run_verdict(h, &all_done);
    const bool stopped = h->dq_ctrl_h[GPSAT_DQC_STOP] != 0;
    if (h->mesh_ranks > 1) {
        // this rank sees only its own contributions: "done" is the mesh-wide termination flag of the communication warp
        // (or the stop flag), the verdict is decided by gpsat_mesh_results_* / gpsat_multi_solve over all ranks
        all_done = h->dq_ctrl_h[GPSAT_DQC_DONE] != 0;
        if (v != GPSAT_SAT) v = GPSAT_UNDEF;
    }
    if (done) *done = (all_done || stopped || (v == GPSAT_SAT && h->opts.stop_on_sat)) ? 1 : 0;
    if (verdict) *verdict = v;
    return GPSAT_OK;
}

int gpsat_solve_end(gpsat_t *h, int32_t *verdict, uint8_t *model, gpsat_stats *stats)
{
    if (!h || !h->solving) {
        set_error("gpsat_solve_end without gpsat_solve_begin");
        return GPSAT_E_STATE;
    }
    DeviceGuard guard(h->device);
    h->solving = false;
    if (h->records_h.size() < (size_t)roots_of(h)) {
        int rc = fetch_records(h);
        if (rc != GPSAT_OK) return rc;
    }
    const int32_t v = run_verdict(h, nullptr);
    if (verdict) *verdict = v;
    if (v == GPSAT_SAT && model)
        CU(cudaMemcpy(model, h->model.p, (size_t)h->D.n_vars, cudaMemcpyDeviceToHost));
    fill_stats(h, stats);
    return GPSAT_OK;
}

int gpsat_solve(gpsat_t *h, int32_t *verdict, uint8_t *model, gpsat_stats *stats)
{
    int rc = gpsat_solve_begin(h);
    if (rc != GPSAT_OK) return rc;
    int32_t done = 0, v = GPSAT_UNDEF;
    rc = gpsat_solve_step(h, 0.0, &done, &v);
    if (rc != GPSAT_OK) {
        h->solving = false;
        return rc;
    }
    return gpsat_solve_end(h, verdict, model, stats);
}

int gpsat_request_stop(gpsat_t *h)
{
    if (!h || !h->region.p) {
        set_error("no run in progress");
        return GPSAT_E_STATE;
    }
    DeviceGuard guard(h->device);
    const int32_t two = 2;
    CU(cudaMemcpy(dq_ctrl(h) + GPSAT_DQC_STOP, &two, sizeof(two), cudaMemcpyHostToDevice));
    return GPSAT_OK;
}

double gpsat_last_kernel_ms(gpsat_t *h) { return h ? h->kernel_ms : 0.0; }

int gpsat_job_records(gpsat_t *h, gpsat_job_record *records, int32_t cap)
{
    if (!h || !records) {
        set_error("bad arguments");
        return GPSAT_E_ARG;
    }
    const int32_t nr = roots_of(h);
    if (cap < nr || h->records_h.size() < (size_t)nr) {
        set_error("record buffer too small or no run yet");
        return GPSAT_E_CAPACITY;
    }
    std::memcpy(records, h->records_h.data(), (size_t)nr * sizeof(gpsat_job_record));
    return GPSAT_OK;
}

int gpsat_pool_export(gpsat_t *h, int32_t *buf, int64_t cap_words, int64_t *n_words)
{
    if (!h || !n_words || (cap_words > 0 && !buf)) {
        set_error("bad arguments");
        return GPSAT_E_ARG;
    }
    DeviceGuard guard(h->device);
    *n_words = 0;
    if (!h->pool.p) return GPSAT_OK;
    int32_t cur[4] = {0, 0, 0, 0};
    CU(cudaMemcpy(cur, h->pool_cursor.p, sizeof(cur), cudaMemcpyDeviceToHost));
    const int64_t used = std::min<int64_t>(cur[0], kPoolSlots);
    int64_t mark = cur[2];
    if (used <= mark) return GPSAT_OK;
    std::vector<int32_t> tmp((size_t)(used - mark) * GPSAT_POOL_SLOT_WORDS);
    CU(cudaMemcpy(tmp.data(), h->pool.p + mark * GPSAT_POOL_SLOT_WORDS, tmp.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    int64_t at = 0;
    const int64_t n_fresh = used - mark;
    for (int64_t sl = 0; sl < n_fresh; sl++) {   // fixed slots -> packed records, whole records only
        const int32_t *rec = tmp.data() + sl * GPSAT_POOL_SLOT_WORDS;
        const int32_t len = rec[0];
        if (len > 0 && len < GPSAT_POOL_SLOT_WORDS) {
            if (at + 1 + len > cap_words) break;
            std::memcpy(buf + at, rec, (size_t)(1 + len) * sizeof(int32_t));
            at += 1 + len;
        }
        mark++;
    }
    *n_words = at;
    const int32_t m32 = (int32_t)mark;
    CU(cudaMemcpy(h->pool_cursor.p + 2, &m32, sizeof(m32), cudaMemcpyHostToDevice));
    return GPSAT_OK;
}

int gpsat_pool_import(gpsat_t *h, const int32_t *buf, int64_t n_words)
{
    if (!h || n_words < 0 || (n_words > 0 && !buf)) {
        set_error("bad arguments");
        return GPSAT_E_ARG;
    }
    DeviceGuard guard(h->device);
    if (n_words == 0) return GPSAT_OK;
    int rc = ensure_pools(h, true);
    if (rc != GPSAT_OK) return rc;
    std::vector<int32_t> slots;
    int64_t at = 0;
    while (at < n_words) {
        const int32_t len = buf[at];
        if (len <= 0 || at + 1 + len > n_words) {
            set_error("malformed pool records");
            return GPSAT_E_ARG;
        }
        for (int32_t i = 0; i < len; i++)
            if (buf[at + 1 + i] < 0 || (buf[at + 1 + i] >> 1) >= h->D.n_vars) {
                set_error("pool literal out of range");
                return GPSAT_E_ARG;
            }
        if (len == 1) {
            const uint8_t f = (uint8_t)(1 + (buf[at + 1] & 1));
            CU(cudaMemcpy(facts(h) + (buf[at + 1] >> 1), &f, 1, cudaMemcpyHostToDevice));
        }
        if (len < GPSAT_POOL_SLOT_WORDS) {   // longer clauses do not fit a slot: optional knowledge, dropped
            const size_t o = slots.size();
            slots.resize(o + GPSAT_POOL_SLOT_WORDS, 0);
            std::memcpy(slots.data() + o, buf + at, (size_t)(1 + len) * sizeof(int32_t));
        }
        at += 1 + len;
    }
    int32_t cur[4] = {0, 0, 0, 0};
    CU(cudaMemcpy(cur, xpool_cursor(h), sizeof(cur), cudaMemcpyDeviceToHost));
    const int64_t n_slots = (int64_t)slots.size() / GPSAT_POOL_SLOT_WORDS;
    if (n_slots == 0 || cur[0] + n_slots > kPoolSlots) return GPSAT_OK;   // pool full: foreign clauses are optional
    CU(cudaMemcpy(xpool(h) + (int64_t)cur[0] * GPSAT_POOL_SLOT_WORDS, slots.data(), slots.size() * sizeof(int32_t),
                  cudaMemcpyHostToDevice));
    cur[0] += (int32_t)n_slots;
    cur[1] += (int32_t)n_slots;
    CU(cudaMemcpy(xpool_cursor(h), cur, sizeof(cur), cudaMemcpyHostToDevice));
    return GPSAT_OK;
}

int64_t gpsat_exchange_block_words(int32_t max_clauses)
{
    return GPSAT_XCHG_HEADER_WORDS + (int64_t)std::max(max_clauses, 0) * GPSAT_POOL_SLOT_WORDS;
}

int gpsat_exchange_pack(gpsat_t *h, void *dev_block, int64_t block_words, int32_t rank, int32_t done, int32_t verdict)
{
    if (!h || !dev_block || block_words < GPSAT_XCHG_HEADER_WORDS || block_words > INT32_MAX) {
        set_error("bad arguments");
        return GPSAT_E_ARG;
    }
    DeviceGuard guard(h->device);
    int rc = ensure_pools(h, true);
    if (rc != GPSAT_OK) return rc;
    int64_t jobs_done = 0;
    for (size_t j = 0; j < h->records_h.size() && j < (size_t)roots_of(h); j++) {
        const int st = h->records_h[j].status;
        jobs_done += (st == GPSAT_SAT || st == GPSAT_UNSAT || st == GPSAT_UNDEF) ? 1 : 0;
    }
    CU(gpsat_kernels::launch_xchg_pack(h->pool.p, h->pool_cursor.p, (int)kPoolSlots, (int *)dev_block, (int)block_words,
                                       rank, done, verdict, (int)jobs_done, h->stream));
    CU(cudaStreamSynchronize(h->stream));   // the caller's collective runs on another stream
    return GPSAT_OK;
}

int gpsat_exchange_unpack(gpsat_t *h, const void *dev_blocks, int32_t n_ranks, int32_t my_rank, int64_t block_words,
                          int32_t *sat_rank, int32_t *all_done, int32_t *any_undef, int64_t *imported_clauses,
                          int64_t *jobs_done_total)
{
    if (!h || !dev_blocks || n_ranks <= 0 || my_rank < 0 || my_rank >= n_ranks ||
        block_words < GPSAT_XCHG_HEADER_WORDS || block_words > INT32_MAX) {
        set_error("bad arguments");
        return GPSAT_E_ARG;
    }
    DeviceGuard guard(h->device);
    int rc = ensure_pools(h, true);
    if (rc != GPSAT_OK) return rc;
    CU(gpsat_kernels::launch_xchg_unpack((const int *)dev_blocks, n_ranks, my_rank, (int)block_words, xpool(h),
                                         xpool_cursor(h), (int)kPoolSlots, facts(h), h->D.n_vars, h->stream));
    std::vector<int32_t> hdr((size_t)n_ranks * GPSAT_XCHG_HEADER_WORDS);
    CU(cudaMemcpy2DAsync(hdr.data(), GPSAT_XCHG_HEADER_WORDS * sizeof(int32_t), dev_blocks,
                         (size_t)block_words * sizeof(int32_t), GPSAT_XCHG_HEADER_WORDS * sizeof(int32_t), (size_t)n_ranks,
                         cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    int32_t sat = -1, done = 1, undef = 0;
    int64_t imported = 0, jobs = 0;
    for (int r = 0; r < n_ranks; r++) {
        const int32_t *b = hdr.data() + (size_t)r * GPSAT_XCHG_HEADER_WORDS;
        if (b[0] != GPSAT_XCHG_MAGIC) {
            set_error("exchange block without header (rank " + std::to_string(r) + ")");
            return GPSAT_E_ARG;
        }
        if (b[1] == GPSAT_SAT && sat < 0) sat = r;
        if (!b[2]) done = 0;
        else if (b[1] == GPSAT_UNDEF) undef = 1;
        if (r != my_rank) imported += b[4];
        jobs += b[6];
    }
    if (sat_rank) *sat_rank = sat;
    if (all_done) *all_done = done;
    if (any_undef) *any_undef = undef;
    if (imported_clauses) *imported_clauses = imported;
    if (jobs_done_total) *jobs_done_total = jobs;
    return GPSAT_OK;
}

int gpsat_debug_ctrl(gpsat_t *h, int32_t *out16)
{
    if (!h || !out16 || !h->ctrl.p || !h->region.p) {
        set_error("no run buffers");
        return GPSAT_E_STATE;
    }
    DeviceGuard guard(h->device);
    // cudaMemcpy on the legacy stream does not wait for the handle's non-blocking stream: usable while a kernel runs
    CU(cudaMemcpy(out16, h->ctrl.p, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost));
    {   // [4] tail [5] head [6] open jobs [7] idle warps [8] splits in flight [9] done [10] steals [11] stop flag
        int32_t c[GPSAT_DQC_WORDS];
        CU(cudaMemcpy(c, dq_ctrl(h), sizeof(c), cudaMemcpyDeviceToHost));
        out16[4] = c[GPSAT_DQC_TAIL];
        out16[5] = c[GPSAT_DQC_HEAD];
        out16[6] = c[GPSAT_DQC_CREATED] - c[GPSAT_DQC_CLOSED];
        out16[7] = c[GPSAT_DQC_IDLE];
        out16[8] = c[GPSAT_DQC_INFLIGHT];
        out16[9] = c[GPSAT_DQC_DONE];
        out16[10] = c[GPSAT_DQC_STEALS];
        out16[11] = c[GPSAT_DQC_STOP];
        out16[1] = c[GPSAT_DQC_STOP];
        out16[0] = c[GPSAT_DQC_CURSOR];
    }
    unsigned long long t[2] = {0, 0};
    CU(cudaMemcpy(t, h->t0.p, sizeof(t), cudaMemcpyDeviceToHost));
    out16[12] = (int32_t)(t[0] & 0x7fffffff);
    out16[13] = (int32_t)(t[1] / 1000);
    out16[14] = h->blocks;
    out16[15] = h->warps_per_block;
    return GPSAT_OK;
}

int gpsat_device_ptrs(gpsat_t *h, void **pool_words, void **pool_cursor, void **stop_flag, void **stream)
{
    if (!h) {
        set_error("null handle");
        return GPSAT_E_ARG;
    }
    if (pool_words) *pool_words = h->pool.p;
    if (pool_cursor) *pool_cursor = h->pool_cursor.p;
    if (stop_flag) *stop_flag = h->region.p ? (void *)(dq_ctrl(h) + GPSAT_DQC_STOP) : nullptr;
    if (stream) *stream = (void *)h->stream;
    return GPSAT_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// Mesh: the GPUs of one box as ONE work pool over NVLink peer memory (no reference equivalent: the reference is
// single-GPU).  Every rank maps the queue regions of the others; the communication warp inside the
// solve kernel (kernels.cu: gpsat_comm_loop) and the stealing path of the warp loop do the rest.
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct IpcMapping {
    unsigned char handle[64];
    void *p;
};
std::vector<IpcMapping> g_ipc_open;   // peer regions stay mapped for the life of the process (regions are recycled
std::mutex g_ipc_mutex;               // through the block cache, so the same few handles come back every solve)

int mesh_prepare(gpsat *h)
{
    if (h->opts.bcp != GPSAT_BCP_WATCHED || !h->opts.dynamic_split) {
        set_error("a mesh needs the watched-literal solver with dynamic_split = 1");
        return GPSAT_E_ARG;
    }
    if (h->mesh_ranks > 1) return GPSAT_OK;
    return ensure_region(h, true);
}

int mesh_set(gpsat *h, int32_t n_ranks, int32_t rank, char *const *bases)
{
    if (n_ranks < 1 || n_ranks > GPSAT_MESH_MAX_RANKS || rank < 0 || rank >= n_ranks) {
        set_error("bad mesh arguments");
        return GPSAT_E_ARG;
    }
    int rc = h->cubes_set ? GPSAT_OK : gpsat_set_cubes(h, 0, nullptr, nullptr);
    if (rc != GPSAT_OK) return rc;
    h->mesh_ranks = n_ranks;
    h->mesh_rank = rank;
    for (int r = 0; r < GPSAT_MESH_MAX_RANKS; r++) h->mesh_base[r] = r < n_ranks ? bases[r] : nullptr;
    h->mesh_base[rank] = h->region.p;
    h->records_h.clear();
    return GPSAT_OK;
}
}  // namespace

int gpsat_mesh_export(gpsat_t *h, void *ipc_handle)
{
    if (!h || !ipc_handle) {
        set_error("bad arguments");
        return GPSAT_E_ARG;
    }
    DeviceGuard guard(h->device);
    int rc = mesh_prepare(h);
    if (rc != GPSAT_OK) return rc;
    static_assert(sizeof(cudaIpcMemHandle_t) == GPSAT_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
    cudaIpcMemHandle_t mh;
    CU(cudaIpcGetMemHandle(&mh, h->region.p));
    std::memcpy(ipc_handle, &mh, sizeof(mh));
    return GPSAT_OK;
}

int gpsat_mesh_attach_ipc(gpsat_t *h, int32_t n_ranks, int32_t rank, const void *ipc_handles)
{
    if (!h || !ipc_handles || n_ranks < 1 || n_ranks > GPSAT_MESH_MAX_RANKS || rank < 0 || rank >= n_ranks) {
        set_error("bad arguments");
        return GPSAT_E_ARG;
    }
    DeviceGuard guard(h->device);
    int rc = mesh_prepare(h);
    if (rc != GPSAT_OK) return rc;
    char *bases[GPSAT_MESH_MAX_RANKS] = {nullptr};
    for (int r = 0; r < n_ranks; r++) {
        if (r == rank) continue;
        const unsigned char *hb = (const unsigned char *)ipc_handles + (size_t)r * GPSAT_IPC_HANDLE_BYTES;
        std::lock_guard<std::mutex> lock(g_ipc_mutex);
        void *p = nullptr;
        for (const auto &m : g_ipc_open)
            if (std::memcmp(m.handle, hb, GPSAT_IPC_HANDLE_BYTES) == 0) p = m.p;
        if (!p) {
            cudaIpcMemHandle_t mh;
            std::memcpy(&mh, hb, sizeof(mh));
            CU(cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
            IpcMapping m;
            std::memcpy(m.handle, hb, GPSAT_IPC_HANDLE_BYTES);
            m.p = p;
            g_ipc_open.push_back(m);
        }
        bases[r] = (char *)p;
    }
    return mesh_set(h, n_ranks, rank, bases);
}

int gpsat_mesh_attach_local(gpsat_t *const *handles, int32_t n_ranks)
{
    if (!handles || n_ranks < 1 || n_ranks > GPSAT_MESH_MAX_RANKS) {
        set_error("bad arguments");
        return GPSAT_E_ARG;
    }
    char *bases[GPSAT_MESH_MAX_RANKS] = {nullptr};
    for (int r = 0; r < n_ranks; r++) {
        if (!handles[r]) {
            set_error("null handle");
            return GPSAT_E_ARG;
        }
        DeviceGuard guard(handles[r]->device);
        int rc = mesh_prepare(handles[r]);
        if (rc != GPSAT_OK) return rc;
        if (handles[r]->ML.total != handles[0]->ML.total) {
            set_error("mesh: the handles were created for different formulas or options");
            return GPSAT_E_ARG;
        }
        bases[r] = handles[r]->region.p;
    }
    for (int r = 0; r < n_ranks; r++) {
        DeviceGuard guard(handles[r]->device);
        for (int q = 0; q < n_ranks; q++) {
            if (handles[q]->device == handles[r]->device) continue;
            int can = 0;
            CU(cudaDeviceCanAccessPeer(&can, handles[r]->device, handles[q]->device));
            if (!can) {
                set_error("mesh: GPU " + std::to_string(handles[r]->device) + " cannot access GPU " +
                          std::to_string(handles[q]->device) + " as a peer");
                return GPSAT_E_CUDA;
            }
            cudaError_t e = cudaDeviceEnablePeerAccess(handles[q]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(e);
            cudaGetLastError();
        }
        if (handles[r]->n_cubes != handles[0]->n_cubes) {
            set_error("mesh: every handle must hold the same (complete) cube list");
            return GPSAT_E_ARG;
        }
        int rc = mesh_set(handles[r], n_ranks, r, bases);
        if (rc != GPSAT_OK) return rc;
    }
    return GPSAT_OK;
}

int gpsat_mesh_detach(gpsat_t *h)
{
    if (!h) {
        set_error("null handle");
        return GPSAT_E_ARG;
    }
    h->mesh_ranks = 1;
    h->mesh_rank = 0;
    h->records_h.clear();
    return GPSAT_OK;
}

// result block of a mesh rank: [root_flag n_roots][root_pending n_roots][records n_roots x 20 words]
int64_t gpsat_mesh_result_words(gpsat_t *h) { return h ? (int64_t)roots_of(h) * (2 + (int64_t)(sizeof(gpsat_job_record) / 4)) : 0; }

int gpsat_mesh_results_pack(gpsat_t *h, void *dev_block, int64_t words)
{
    if (!h || !dev_block || words < gpsat_mesh_result_words(h) || !h->records.p) {
        set_error("bad arguments");
        return GPSAT_E_ARG;
    }
    DeviceGuard guard(h->device);
    const size_t nr = (size_t)roots_of(h);
    int32_t *b = (int32_t *)dev_block;
    CU(cudaMemcpyAsync(b, h->root_flag.p, nr * 4, cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaMemcpyAsync(b + nr, h->root_pending.p, nr * 4, cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaMemcpyAsync(b + 2 * nr, h->records.p, nr * sizeof(gpsat_job_record), cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return GPSAT_OK;
}

int gpsat_mesh_results_unpack(gpsat_t *h, const void *dev_block, int64_t words, int32_t *verdict, gpsat_stats *stats)
{
    if (!h || !dev_block || words < gpsat_mesh_result_words(h)) {
        set_error("bad arguments");
        return GPSAT_E_ARG;
    }
    DeviceGuard guard(h->device);
    const size_t nr = (size_t)roots_of(h);
    const int32_t *b = (const int32_t *)dev_block;
    h->records_h.resize(nr);
    h->root_pending_h.resize(nr);
    h->root_flag_h.resize(nr);
    CU(cudaMemcpyAsync(h->root_flag_h.data(), b, nr * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(h->root_pending_h.data(), b + nr, nr * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(h->records_h.data(), b + 2 * nr, nr * sizeof(gpsat_job_record), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (size_t j = 0; j < nr; j++) h->records_h[j].status = gpsat_root_status(h->root_flag_h[j], h->root_pending_h[j]);
    if (verdict) *verdict = run_verdict(h, nullptr);
    fill_stats(h, stats);
    return GPSAT_OK;
}

int gpsat_handle_device(gpsat_t *h) { return h ? h->device : -1; }

int gpsat_get_phase_stats(gpsat_t *h, gpsat_phase_stats *out)
{
    if (!h || !out) {
        set_error("bad arguments");
        return GPSAT_E_ARG;
    }
    std::memset(out, 0, sizeof(*out));
    if (!h->opts.phase_stats) {
        set_error("the handle was created without opts.phase_stats");
        return GPSAT_E_STATE;
    }
    const int64_t *w = (const int64_t *)(h->dq_ctrl_h + GPSAT_DQC_PHASE);
    for (int i = 0; i < GPSAT_N_PHASES; i++) {
        out->ns[i] = w[i];
        out->count[i] = w[GPSAT_N_PHASES + i];
    }
    out->backtracked_levels = w[2 * GPSAT_N_PHASES];
    out->jobs = h->dq_ctrl_h[GPSAT_DQC_CLOSED];
    DeviceGuard guard(h->device);
    unsigned long long busy = 0;
    if (h->t0.p) CU(cudaMemcpy(&busy, h->t0.p + 1, sizeof(busy), cudaMemcpyDeviceToHost));
    out->job_ns = (int64_t)busy;
    const double warp_ns = (double)h->blocks * h->warps_per_block * h->kernel_ms * 1e6;
    out->idle_ns = (int64_t)std::max(0.0, warp_ns - (double)busy);
    return GPSAT_OK;
}

int gpsat_debug_words(gpsat_t *h, int32_t *out, int32_t n)
{
    if (!h || !out || n < 0 || n > GPSAT_DQC_WORDS) {
        set_error("bad arguments");
        return GPSAT_E_ARG;
    }
    std::memcpy(out, h->dq_ctrl_h, (size_t)n * sizeof(int32_t));
    return GPSAT_OK;
}

}  // extern "C"
