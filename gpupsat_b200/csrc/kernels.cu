// sm_100a kernels of gpupsat_b200.
//
//   gpsat_cdcl_kernel   persistent, one warp per cube (≙ parallel_kernel, SATSolver/Parallelizer.cu:180-228, and
//                       run_sequential :230-278): each warp pulls the next cube from an atomic cursor
//                       (≙ JobsQueue::next_job, SATSolver/JobsQueue.cu:10-32), runs the whole CDCL search for it
//                       (cdcl_warp.inl), writes its record, and polls the early-termination flag inside the loop
//                       rather than at kernel boundaries (SATSolver/main.cu:259-269).
//   gpsat_eval_kernel   clause evaluation (≙ VariablesStateHandler::clause_status, SATSolver/VariablesStateHandler.cu:180-206)
//
// Tensor cores are not used anywhere: this path is integer gather / scan work (BASELINE.json north_star).
#include "kernels.h"
#include "cdcl_warp.inl"

#define GPSAT_MAX_THREADS 896       // widest block: 25 to 28 warps, 7 on some scheduler — 72 registers per thread
#define GPSAT_MID_THREADS 768       // 21 to 24 warps, 6 on some scheduler — 80 registers per thread
#define GPSAT_WIDE_THREADS 640      // blocks of up to 20 warps get up to 96 registers per thread

namespace {

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void gpsat_stamp_kernel(unsigned long long *t0) { *t0 = globaltimer_ns(); }

// ---------------------------------------------------------------------------------------------------------------
// Communication warp (mesh runs; no reference equivalent — the reference is single-GPU).
// One warp per GPU lives inside the persistent solve kernel and does, over NVLink peer memory, what would otherwise
// need the kernel to end and the host to run a collective:
//   * advertises this GPU's queued children (PEER_QUEUE) and its unmet demand (PEER_IDLE) in the control block of
//     every other GPU — idle warps there steal the children, busy cubes there split for the demand;
//   * pushes the learnt clauses this GPU's jobs published since the last round into the foreign pool of every other
//     GPU (the in-kernel form of the per-epoch all-gather: a 64-byte slot per clause, length written last) and learnt
//     units into their `facts`;
//   * forwards the early-termination flag (≙ managed *state, SATSolver/main.cu:185,266);
//   * detects global termination: sum over ranks of `closed`, read BEFORE the sum of `created`; both counters only
//     grow and closed <= created at every instant, so equal sums mean that no job was open at the moment the last
//     `closed` was read — and then none can be created any more.
// ---------------------------------------------------------------------------------------------------------------
// what the communication warp needs of the kernel parameters, passed BY VALUE: handing the parameter structs to a
// non-inlined function by reference would make the compiler keep an addressable copy of them in local memory for
// every warp of the kernel
struct CommArgs {
    int mesh_ranks, mesh_rank, share_learnts, pool_cap_words, xpool_cap_slots, mesh_n_vars;
    int *dq_ctrl;
    const int *pool;
    const int *pool_cursor;
    const unsigned long long *t0;
    unsigned long long budget_ns;
    char *mesh_base[GPSAT_MESH_MAX_RANKS];
    long long mesh_off_ctrl, mesh_off_xcur, mesh_off_xpool, mesh_off_facts;
};

__device__ __noinline__ void gpsat_comm_loop(const CommArgs B)
{
    const struct { int share_learnts; } P = {B.share_learnts};
    const int lane = (int)(threadIdx.x & 31);
    const int R = B.mesh_ranks, me = B.mesh_rank;
    int *ctrl = B.dq_ctrl;
    const bool peer_lane = lane < R && lane != me;
    int *pctrl = peer_lane ? (int *)(B.mesh_base[lane] + B.mesh_off_ctrl) : ctrl;
    const int n_workers = (int)(gridDim.x * (blockDim.x >> 5)) - 1;
    const int pool_slots = B.pool_cap_words / GPSAT_POOL_SLOT_WORDS;
    int pushed = *(volatile int *)(ctrl + GPSAT_DQC_PUSHED);
    unsigned spins = 0;
    while (true) {
        const int stop = *(volatile int *)(ctrl + GPSAT_DQC_STOP);
        if (stop) {   // a model was found here, or the host / another GPU asked to stop: tell everybody, leave
            if (peer_lane && *(volatile int *)(pctrl + GPSAT_DQC_STOP) == 0) *(volatile int *)(pctrl + GPSAT_DQC_STOP) = 2;
            break;
        }
        if (B.budget_ns && globaltimer_ns() > *B.t0 + B.budget_ns) break;
        const int idle = *(volatile int *)(ctrl + GPSAT_DQC_IDLE);
        const int queued = *(volatile int *)(ctrl + GPSAT_DQC_TAIL) - *(volatile int *)(ctrl + GPSAT_DQC_HEAD);
        const int inflight = *(volatile int *)(ctrl + GPSAT_DQC_INFLIGHT);
        if (peer_lane) {
            int want = idle - queued - inflight;   // warps here that would take a child right now
            want = want > 0 ? (want + R - 2) / (R - 1) : 0;
            const int avail = *(volatile int *)(ctrl + GPSAT_DQC_AVAIL);   // published children nobody has claimed yet
            *(volatile int *)(pctrl + GPSAT_DQC_PEER_QUEUE + me) = avail > 0 ? avail : 0;
            *(volatile int *)(pctrl + GPSAT_DQC_PEER_IDLE + me) = want;
        }
        // ---- learnt clauses published on this GPU since the last round -> every peer's foreign pool
        if (B.pool != nullptr && P.share_learnts) {
            int used = *(volatile int *)B.pool_cursor;
            if (used > pool_slots) used = pool_slots;
            while (pushed < used) {
                const int slot = pushed + lane;
                const int len = slot < used ? __ldcg(B.pool + (size_t)slot * GPSAT_POOL_SLOT_WORDS) : 0;
                const unsigned ready = __ballot_sync(0xffffffffu, len > 0);
                const int cnt = __ffs(~ready) - 1 < 0 ? 32 : __ffs(~ready) - 1;   // complete slots form a prefix
                if (cnt == 0) break;
                int4 w0 = make_int4(0, 0, 0, 0), w1 = w0, w2 = w0, w3 = w0;
                if (lane < cnt) {
                    const int4 *src = reinterpret_cast<const int4 *>(B.pool + (size_t)slot * GPSAT_POOL_SLOT_WORDS);
                    w0 = __ldcg(src);
                    w1 = __ldcg(src + 1);
                    w2 = __ldcg(src + 2);
                    w3 = __ldcg(src + 3);
                }
                int base = 0;   // lane r reserves cnt slots of rank r's foreign pool
                if (peer_lane) base = atomicAdd_system((int *)(B.mesh_base[lane] + B.mesh_off_xcur), cnt);
                for (int r = 0; r < R; ++r) {
                    const int rb = __shfl_sync(0xffffffffu, base, r);
                    if (r == me) continue;
                    const int dst = rb + lane;
                    if (lane < cnt && dst < B.xpool_cap_slots) {
                        int *x = (int *)(B.mesh_base[r] + B.mesh_off_xpool) + (size_t)dst * GPSAT_POOL_SLOT_WORDS;
                        int4 *x4 = reinterpret_cast<int4 *>(x);
                        x4[1] = w1;
                        x4[2] = w2;
                        x4[3] = w3;
                        x[1] = w0.y;
                        x[2] = w0.z;
                        x[3] = w0.w;
                        if (w0.x == 1 && (w0.y >> 1) < B.mesh_n_vars)
                            ((volatile unsigned char *)(B.mesh_base[r] + B.mesh_off_facts))[w0.y >> 1] = (unsigned char)(1 + (w0.y & 1));
                        __threadfence_system();
                        *(volatile int *)x = w0.x;   // length last: the slot is complete
                    }
                }
                pushed += cnt;
                if (cnt < 32) break;
            }
        }
        // ---- termination: only worth the two NVLink round trips when this GPU has nothing to do
        if ((idle >= n_workers && queued <= 0) || (spins & 15u) == 15u) {
            const unsigned long long a = *(const volatile unsigned long long *)(pctrl + GPSAT_DQC_CREATED);
            int closed = (lane < R) ? (int)(a >> 32) : 0;
            __threadfence_system();
            const unsigned long long b = *(const volatile unsigned long long *)(pctrl + GPSAT_DQC_CREATED);
            int created = (lane < R) ? (int)(b & 0xffffffffull) : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                closed += __shfl_xor_sync(0xffffffffu, closed, o);
                created += __shfl_xor_sync(0xffffffffu, created, o);
            }
            if (closed == created) {
                if (lane == 0) *(volatile int *)(ctrl + GPSAT_DQC_DONE) = 1;
                break;
            }
        }
        spins++;
        __nanosleep(1500);
    }
    if (lane == 0) *(volatile int *)(ctrl + GPSAT_DQC_PUSHED) = pushed;
}

// Queue state of a solve, initialised on the device (one small launch instead of several pageable H2D copies):
// ring slots empty (sequence number = slot index), control block, one open job per root cube this rank owns.
__global__ void __launch_bounds__(256)
gpsat_queue_init_kernel(int *ctrl, int *meta, int dq_cap, int *root_pending, int *root_flag, int n_roots, int owner,
                        int *xcur, unsigned char *facts, int n_vars, int *run_ctrl, unsigned long long *t0)
{
    // owner: this rank's control block holds the root cursor of the run (single GPU, or rank 0 of a mesh): the root
    // cubes count as created here, and their "one open job each" lives in this rank's root_pending
    const int tid = (int)(blockIdx.x * blockDim.x + threadIdx.x), nt = (int)(gridDim.x * blockDim.x);
    for (int i = tid; i < GPSAT_DQC_WORDS; i += nt) ctrl[i] = (i == GPSAT_DQC_CREATED && owner) ? n_roots : 0;
    if (meta)
        for (int i = tid; i < dq_cap; i += nt) {
            meta[4 * i] = 0;
            meta[4 * i + 1] = 0;
            meta[4 * i + 2] = i;
            meta[4 * i + 3] = 0;
        }
    for (int i = tid; i < n_roots; i += nt) {
        root_pending[i] = owner ? 1 : 0;
        root_flag[i] = 0;
    }
    if (xcur)
        for (int i = tid; i < 16; i += nt) xcur[i] = 0;
    if (facts)
        for (int i = tid; i < n_vars; i += nt) facts[i] = 0;
    if (tid == 0) {
        run_ctrl[0] = 0;
        run_ctrl[1] = 0;
        run_ctrl[2] = -1;   // sat_job
        run_ctrl[3] = 0;
        t0[0] = 0;
        t0[1] = 0;
    }
}

// kSmemState:   the per-job state blocks live in dynamic shared memory.
// kSmemFormula: the read-only formula index (cl2, occ2, ostart) is staged once per block in shared memory, in front
//               of the state blocks, and every warp of the block reads it from there; the (x, y) pairs of cl2 / occ2
//               are packed into one word each on the way (a formula that fits has fewer than 65 536 slots), which
//               halves the staged copy: C2 31 KB instead of 61 KB, room for 23 warps of state instead of 20.
// kThreads:     launch bound — the register budget follows from it.
// Both are template parameters (not run-time selects) so that the pointers provably derive from the shared window
// and compile to LDS/STS/ATOMS instead of generic loads.
template <bool kSmemState, bool kSmemFormula, int kThreads>
__global__ void __launch_bounds__(kThreads, 1)
gpsat_cdcl_kernel(const gpsat_formula_view F, const gpsat_solve_params P, const gpsat_state_layout Ly,
                  const gpsat_run_buffers B)
{
    extern __shared__ __align__(16) int gpsat_smem[];
    // %tid.x through a volatile asm, read ONCE: the compiler otherwise rematerialises the per-warp state pointers from
    // S2R %tid.x wherever it runs short of registers — 11 % of the instructions and 7 % of the stall samples of this
    // kernel were such recomputations (profiles/r02_cdcl_lines_b.txt: kernels.cu:196/219, cdcl_warp.inl:1465)
    unsigned tid_x;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid_x));
    // through a warp reduction: its result lives in a UNIFORM register, so the compiler knows that the state pointers
    // derived from it are the same for all lanes (uniform datapath) instead of spending vector registers on them
    const int warp_in_block = (int)__reduce_min_sync(0xffffffffu, tid_x >> 5);
    const int warps_per_block = (int)(blockDim.x >> 5);
    const long long gwarp = (long long)blockIdx.x * warps_per_block + warp_in_block;

    gpsat_formula_view Fv = F;
    int *state_base = gpsat_smem;
    if (kSmemFormula) {
        // layout: cl2 | occ2 | ostart, each rounded up to 4 words; one word per (x, y) pair (gpsat_formula_smem_words)
        const int n_cl2 = F.n_lits + F.n_clauses, n_occ2 = F.n_lits, n_os = 2 * F.n_vars + 1;
        int *s_cl2 = gpsat_smem;
        int *s_occ2 = s_cl2 + ((n_cl2 + 3) & ~3);
        int *s_os = s_occ2 + ((n_occ2 + 3) & ~3);
        const int2 *g_cl2 = (const int2 *)F.cl2, *g_occ2 = (const int2 *)F.occ2;
        for (int i = (int)tid_x; i < n_cl2; i += (int)blockDim.x) {
            const int2 q = g_cl2[i];
            s_cl2[i] = (int)((uint32_t)q.x | ((uint32_t)q.y << 16));
        }
        for (int i = (int)tid_x; i < n_occ2; i += (int)blockDim.x) {
            const int2 q = g_occ2[i];
            s_occ2[i] = (int)((uint32_t)q.x | ((uint32_t)q.y << 16));
        }
        for (int i = (int)tid_x; i < n_os; i += (int)blockDim.x) s_os[i] = F.ostart[i];
        __syncthreads();
        Fv.cl2 = s_cl2;
        Fv.occ2 = s_occ2;
        Fv.ostart = s_os;
        state_base = gpsat_smem + B.formula_smem_words;
    }
    int *state;
    if (kSmemState) state = state_base + (size_t)warp_in_block * Ly.total_words;
    else state = B.gstate + (size_t)gwarp * Ly.total_words;
    int *arena = B.arena ? B.arena + (size_t)gwarp * (size_t)P.arena_words : nullptr;

    int *park = B.park ? B.park + (size_t)gwarp * (size_t)B.park_words : nullptr;
    int *stage = B.stage ? B.stage + (size_t)gwarp * (size_t)(B.hand_words + GPSAT_DQ_MAXK) : nullptr;

    // mesh: warp 0 of this GPU does not solve; it is the GPU's link to the other GPUs of the box
    if (B.mesh_ranks > 1 && gwarp == 0) {
        CommArgs C;
        C.mesh_ranks = B.mesh_ranks;
        C.mesh_rank = B.mesh_rank;
        C.share_learnts = (P.mesh_flags & 2) ? 0 : P.share_learnts;
        C.pool_cap_words = B.pool_cap_words;
        C.xpool_cap_slots = B.xpool_cap_slots;
        C.mesh_n_vars = B.mesh_n_vars;
        C.dq_ctrl = B.dq_ctrl;
        C.pool = B.pool;
        C.pool_cursor = B.pool_cursor;
        C.t0 = B.t0;
        C.budget_ns = B.budget_ns;
#pragma unroll
        for (int r = 0; r < GPSAT_MESH_MAX_RANKS; ++r) C.mesh_base[r] = B.mesh_base[r];
        C.mesh_off_ctrl = B.mesh_off_ctrl;
        C.mesh_off_xcur = B.mesh_off_xcur;
        C.mesh_off_xpool = B.mesh_off_xpool;
        C.mesh_off_facts = B.mesh_off_facts;
        gpsat_comm_loop(C);
        return;
    }
    WarpSolverT<kSmemFormula> S;
    gpsat_bind(S, Fv, P, Ly, state, arena, park, B);
    gpsat_warp_loop(S, P, B, stage);
}

typedef void (*cdcl_kernel_t)(const gpsat_formula_view, const gpsat_solve_params, const gpsat_state_layout,
                              const gpsat_run_buffers);
cdcl_kernel_t pick_kernel(bool smem_state, bool smem_formula, int threads)
{
    const int bound = threads <= GPSAT_WIDE_THREADS ? 0 : threads <= GPSAT_MID_THREADS ? 1 : 2;
    if (smem_state && smem_formula)
        return bound == 0 ? gpsat_cdcl_kernel<true, true, GPSAT_WIDE_THREADS>
                          : bound == 1 ? gpsat_cdcl_kernel<true, true, GPSAT_MID_THREADS> : gpsat_cdcl_kernel<true, true, GPSAT_MAX_THREADS>;
    if (smem_state)
        return bound == 0 ? gpsat_cdcl_kernel<true, false, GPSAT_WIDE_THREADS>
                          : bound == 1 ? gpsat_cdcl_kernel<true, false, GPSAT_MID_THREADS> : gpsat_cdcl_kernel<true, false, GPSAT_MAX_THREADS>;
    return gpsat_cdcl_kernel<false, false, GPSAT_WIDE_THREADS>;
}

// One thread per (assignment, clause).  Clause literals are read from the compact CSR (4 B per literal, coalesced
// across the warp for fixed-width clauses); variable values are byte gathers served by L1/L2.
__global__ void __launch_bounds__(256)
gpsat_eval_kernel(int32_t n_vars, int32_t n_clauses, const int32_t *__restrict__ coffsets,
                  const int32_t *__restrict__ clits, const uint8_t *__restrict__ assignment,
                  int32_t *__restrict__ status, int32_t *__restrict__ unit)
{
    const int a = (int)blockIdx.y;
    const uint8_t *__restrict__ as = assignment + (size_t)a * (size_t)n_vars;
    for (int c = (int)(blockIdx.x * blockDim.x + threadIdx.x); c < n_clauses; c += (int)(gridDim.x * blockDim.x)) {
        const int b = __ldg(coffsets + c), e = __ldg(coffsets + c + 1);
        int n_false = 0, last_undef = -1, st = -1;
        for (int i = b; i < e; ++i) {
            const int x = __ldg(clits + i);
            const int sv = as[x >> 1];   // reference sat_status: 0 true, 1 false, 2 unassigned
            if (sv == 2) {
                last_undef = x;
            } else if ((sv == 0) == ((x & 1) == 1)) {
                st = GPSAT_SAT;
                break;
            } else {
                n_false++;
            }
        }
        int u = -1;
        if (st != GPSAT_SAT) {
            st = (n_false == e - b) ? GPSAT_UNSAT : GPSAT_UNDEF;
            if (n_false == e - b - 1) u = last_undef;
        }
        const size_t o = (size_t)a * (size_t)n_clauses + (size_t)c;
        status[o] = st;
        if (unit) unit[o] = u;
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Epoch exchange (multi-GPU; no reference equivalent — the reference is single-GPU).
// One exchange block per GPU: [magic, verdict, done, payload words, clauses, rank, jobs done, -] + pool slots.  The blocks
// of all ranks are all-gathered by ONE NCCL collective per epoch; unpack appends the other ranks' slots to the
// foreign pool that jobs import when they start.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gpsat_xchg_pack_kernel(const int *__restrict__ pool, int *pool_cursor, int pool_cap_slots, int *__restrict__ block,
                       int block_words, int rank, int done, int verdict, int jobs_done)
{
    __shared__ int s_from, s_n;
    if (threadIdx.x == 0) {
        int used = pool_cursor[0];
        if (used > pool_cap_slots) used = pool_cap_slots;
        const int mark = pool_cursor[2];
        const int room = (block_words - GPSAT_XCHG_HEADER_WORDS) / GPSAT_POOL_SLOT_WORDS;
        int n = used - mark;
        if (n > room) n = room;
        if (n < 0) n = 0;
        s_from = mark;
        s_n = n;
    }
    __syncthreads();
    const int from = s_from, n = s_n;
    const int4 *src = reinterpret_cast<const int4 *>(pool + (size_t)from * GPSAT_POOL_SLOT_WORDS);
    int4 *dst = reinterpret_cast<int4 *>(block + GPSAT_XCHG_HEADER_WORDS);
    for (int i = (int)(blockIdx.x * blockDim.x + threadIdx.x); i < n * (GPSAT_POOL_SLOT_WORDS / 4);
         i += (int)(gridDim.x * blockDim.x))
        dst[i] = src[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        block[0] = GPSAT_XCHG_MAGIC;
        block[1] = verdict;
        block[2] = done;
        block[3] = n * GPSAT_POOL_SLOT_WORDS;
        block[4] = n;
        block[5] = rank;
        block[6] = jobs_done;
        block[7] = 0;
        pool_cursor[2] = from + n;
    }
}

__global__ void __launch_bounds__(256)
gpsat_xchg_unpack_kernel(const int *__restrict__ blocks, int n_ranks, int my_rank, int block_words,
                         int *__restrict__ xpool, int *xpool_cursor, int xpool_cap_slots, unsigned char *facts,
                         int n_vars)
{
    // every thread recomputes the (tiny) prefix over ranks; thread 0 publishes the new cursor at the end
    int base = xpool_cursor[0];
    for (int r = 0; r < n_ranks; ++r) {
        const int *b = blocks + (size_t)r * block_words;
        if (r == my_rank || b[0] != GPSAT_XCHG_MAGIC) continue;
        int n = b[4];
        if (n > xpool_cap_slots - base) n = xpool_cap_slots - base;
        if (n <= 0) continue;
        const int4 *src = reinterpret_cast<const int4 *>(b + GPSAT_XCHG_HEADER_WORDS);
        int4 *dst = reinterpret_cast<int4 *>(xpool + (size_t)base * GPSAT_POOL_SLOT_WORDS);
        for (int i = (int)threadIdx.x; i < n * (GPSAT_POOL_SLOT_WORDS / 4); i += (int)blockDim.x) dst[i] = src[i];
        // unit clauses of the other GPU become facts of this one
        for (int i = (int)threadIdx.x; i < n; i += (int)blockDim.x) {
            const int *rec = b + GPSAT_XCHG_HEADER_WORDS + (size_t)i * GPSAT_POOL_SLOT_WORDS;
            if (rec[0] == 1 && facts && rec[1] >= 0 && (rec[1] >> 1) < n_vars) facts[rec[1] >> 1] = (unsigned char)(1 + (rec[1] & 1));
        }
        base += n;
    }
    __syncthreads();
    if (threadIdx.x == 0) xpool_cursor[0] = base;
}

}  // namespace

namespace gpsat_kernels {

cudaError_t launch_cdcl(const gpsat_formula_view &F, const gpsat_solve_params &P, const gpsat_state_layout &Ly,
                        const gpsat_run_buffers &B, int blocks, int warps_per_block, size_t smem_bytes,
                        cudaStream_t stream)
{
    if (warps_per_block * 32 > GPSAT_MAX_THREADS) return cudaErrorInvalidConfiguration;
    cdcl_kernel_t k = pick_kernel(B.state_in_smem != 0, B.formula_in_smem != 0, warps_per_block * 32);
    if (smem_bytes > 0) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return e;
    }
    k<<<blocks, warps_per_block * 32, smem_bytes, stream>>>(F, P, Ly, B);
    return cudaGetLastError();
}

cudaError_t cdcl_occupancy(int warps_per_block, size_t smem_bytes, bool smem_state, bool smem_formula,
                           int *blocks_per_sm)
{
    cdcl_kernel_t k = pick_kernel(smem_state, smem_formula, warps_per_block * 32);
    if (smem_bytes > 0) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return e;
    }
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k, warps_per_block * 32, smem_bytes);
}

int cdcl_max_warps_per_block() { return GPSAT_MAX_THREADS / 32; }

cudaError_t cdcl_attributes(int *regs_per_thread, size_t *local_bytes)
{
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, gpsat_cdcl_kernel<true, true, GPSAT_MAX_THREADS>);
    if (e != cudaSuccess) return e;
    *regs_per_thread = a.numRegs;
    *local_bytes = a.localSizeBytes;
    return cudaSuccess;
}

cudaError_t launch_xchg_pack(const int *pool, int *pool_cursor, int pool_cap_slots, int *block, int block_words,
                             int rank, int done, int verdict, int jobs_done, cudaStream_t stream)
{
    gpsat_xchg_pack_kernel<<<1, 256, 0, stream>>>(pool, pool_cursor, pool_cap_slots, block, block_words, rank, done,
                                                  verdict, jobs_done);
    return cudaGetLastError();
}

cudaError_t launch_xchg_unpack(const int *blocks, int n_ranks, int my_rank, int block_words, int *xpool,
                               int *xpool_cursor, int xpool_cap_slots, unsigned char *facts, int n_vars,
                               cudaStream_t stream)
{
    gpsat_xchg_unpack_kernel<<<1, 256, 0, stream>>>(blocks, n_ranks, my_rank, block_words, xpool, xpool_cursor,
                                                    xpool_cap_slots, facts, n_vars);
    return cudaGetLastError();
}

cudaError_t launch_queue_init(int *ctrl, int *meta, int dq_cap, int *root_pending, int *root_flag, int n_roots, int owner,
                              int *xcur, unsigned char *facts, int n_vars, int *run_ctrl, unsigned long long *t0,
                              cudaStream_t stream)
{
    gpsat_queue_init_kernel<<<32, 256, 0, stream>>>(ctrl, meta, dq_cap, root_pending, root_flag, n_roots, owner, xcur,
                                                    facts, n_vars, run_ctrl, t0);
    return cudaGetLastError();
}

cudaError_t launch_stamp(unsigned long long *t0, cudaStream_t stream)
{
    gpsat_stamp_kernel<<<1, 1, 0, stream>>>(t0);
    return cudaGetLastError();
}

cudaError_t launch_eval_clauses(int32_t n_vars, int32_t n_clauses, const int32_t *coffsets, const int32_t *clits,
                                int32_t n_assignments, const uint8_t *assignment, int32_t *status, int32_t *unit,
                                cudaStream_t stream)
{
    if (n_clauses <= 0 || n_assignments <= 0) return cudaSuccess;
    int bx = (n_clauses + 255) / 256;
    const int cap = 148 * 16;
    if (bx > cap) bx = cap;
    dim3 grid((unsigned)bx, (unsigned)n_assignments);
    gpsat_eval_kernel<<<grid, 256, 0, stream>>>(n_vars, n_clauses, coffsets, clits, assignment, status, unit);
    return cudaGetLastError();
}

}  // namespace gpsat_kernels

// ---------------------------------------------------------------------------------------------------------------
// BCP by clause evaluation over occurrence lists for LARGE clause databases (config 4: n = 1e6, m = 4e6), where neither
// watch state nor assignments of a job fit beside the formula in shared memory.
//   ≙ ConflictAnalyzer::propagate_all_clauses + VariablesStateHandler::clause_status
//     (ConflictAnalysis/ConflictAnalyzer.cu:173-245, SATSolver/VariablesStateHandler.cu:180-206) made incremental:
//     only clauses containing a literal that just became false are evaluated.
// One CTA per job (cube); the trail is the cube followed by the implied literals, written straight into the caller's
// `implied` block; all cube literals are assigned before propagation starts, as the reference does
// (VariablesStateHandler::set_assumptions, SATSolver.cu:231-246).  Unit propagation is confluent: status and, without
// conflict, the implied SET do not depend on the visiting order.  Two kernels:
//   gpsat_bcp_sweep_tern_kernel  pure 3-SAT, n <= ~1.04 M: whole assignment on chip (base-3 digits), bucket index
//   gpsat_bcp_sweep_cta_kernel   every other database: assigned-bit filter on chip, values in an L2-resident block
// (round 1 also shipped a warp-per-job kernel with the bitmap in HBM and a thread-block-cluster kernel with the bitmap in
// distributed shared memory; both were 3-20x slower — profiles/r01_sweep_ncu_e.json, r01_sweepc_ncu_f.json — and are gone.)
// ---------------------------------------------------------------------------------------------------------------
namespace {

struct SweepArgs {
    int32_t n_vars, n_clauses, n_cubes;
    int32_t uniform3;                 // every clause has exactly 3 literals: use the (other, other) pair entries
    int32_t stream_index;             // read the occurrence index with evict-first loads (keeps the value blocks in L2)
    const int2 *orange;               // 2n : (begin, end) of a literal's PADDED occurrence list: begin and end are even,
                                      //      padding entries hold -1
    const int32_t *occ_clause;        // per entry: clause index
    const int2 *occ_pair;             // per entry: the two other literals of that clause (uniform3 only)
    const uint4 *bucket;              // ternary kernel: one 64-byte bucket per literal id (2n + 2 of them)
    int32_t tern_state_bytes;         // ternary kernel: ceil((n + 1) / 5) rounded up to 16
    int32_t l2_prefetch;              // ternary kernel: prefetch the next batch's buckets into L2
    const int32_t *cube_short;        // ternary kernel: per cube, how many leading literals have at most 5 occurrences (or null)
    const int32_t *coffsets;          // n_clauses+1 (general path)
    const int32_t *clits;             // compact literals
    const int64_t *cube_offsets;
    const int32_t *cube_lits;
    uint32_t *valbits;                // n_warps * val_words, zero on entry and on exit
    int32_t val_words;                // ceil(n/16)
    int32_t *implied;                 // n_cubes * stride
    int64_t stride;
    int32_t *n_implied, *status;
    int64_t *conflict_clause;
    int64_t *counters;                // per cube: [0] occurrence entries visited [1] clause literals read
    int32_t *next_job;
};

// ---------------------------------------------------------------------------------------------------------------
// gpsat_bcp_sweep_cta_kernel — occurrence-list BCP with ONE CTA per job and a two-level assignment:
//   A  "assigned" bit per variable in SHARED memory (n/8 bytes: 125 KB at n = 1e6, fits one SM);
//   V  2-bit (valid, value) field per variable in global memory (this CTA's 250 KB block, L2 resident).
// A literal is assigned by ONE atomicOr on V — its return value says whether the variable was still free (the
// authority for "append to the trail exactly once") or already carried the same / the opposite value — followed by
// setting the A bit.  A lookup first tests A in shared memory: 87 % of the lookups of config 4 hit an unassigned
// variable and end there; only when A is set is V read (V was written before A, so it is valid by then).
// The filter may be SMALLER than the variable count (2^k bits indexed by var & mask, for databases whose exact filter
// does not fit one SM): a set bit may then belong to an aliased variable and V decides — correctness does not depend on
// the filter, only the share of lookups that end in shared memory does.
// A stale A bit can only read "unassigned" for a variable assigned during the current round; the literal that made
// it assigned is processed in a later round (after a barrier) and re-examines the clause, so no unit is lost.
// ---------------------------------------------------------------------------------------------------------------
#define GPSAT_SWEEP_CHUNK 8
struct CtaBits {
    uint32_t *a;          // shared: assigned bits, a FILTER of 2^k bits indexed by var & mask (exact when 2^k >= n;
                          // when smaller, a set bit may belong to an aliased variable and V decides)
    uint32_t *v;          // global: 2-bit fields, 16 variables per word
    uint32_t mask;
    __device__ __forceinline__ int value(int x) const   // 1 true, 0 false, 2 unassigned
    {
        const int var = x >> 1;
        const uint32_t fb = (uint32_t)var & mask;
        if (!((a[fb >> 5] >> (fb & 31)) & 1u)) return 2;
        const uint32_t f = (__ldcg(v + (var >> 4)) >> ((var & 15) * 2)) & 3u;
        return (f & 2u) ? (int)((f & 1u) == (uint32_t)(x & 1)) : 2;
    }
    __device__ __forceinline__ uint32_t assign(int x) const   // previous 2-bit field (0 = was unassigned)
    {
        const int var = x >> 1, sh = (var & 15) * 2;
        const uint32_t prev = (atomicOr(v + (var >> 4), (2u | (uint32_t)(x & 1)) << sh) >> sh) & 3u;
        if (prev == 0) {
            const uint32_t fb = (uint32_t)var & mask;
            atomicOr(a + (fb >> 5), 1u << (fb & 31));
        }
        return prev;
    }
};

template <int kThreads, int kBlocksPerSm>
__global__ void __launch_bounds__(kThreads, kBlocksPerSm) gpsat_bcp_sweep_cta_kernel(const SweepArgs A, const int filter_log2)
{
    extern __shared__ __align__(16) uint32_t s_abits[];
    __shared__ int s_count, s_conflict, s_clause, s_job, s_total, s_stop;
    const int tid = (int)threadIdx.x, nthreads = (int)blockDim.x;
    const int a_words = 1 << (filter_log2 - 5);
    CtaBits bits;
    bits.a = s_abits;
    bits.mask = (1u << filter_log2) - 1u;
    bits.v = A.valbits + (size_t)blockIdx.x * (size_t)A.val_words;

    while (true) {
        if (tid == 0) {
            s_job = atomicAdd(A.next_job, 1);
            s_count = 0;
            s_conflict = 0;
            s_clause = -1;
        }
        for (int i = tid; i < a_words; i += nthreads) s_abits[i] = 0u;
        {   // this CTA's V block back to all-unassigned (coalesced 16-byte stores; val_words is a multiple of 4)
            uint4 *z = reinterpret_cast<uint4 *>(bits.v);
            for (int i = tid; i < A.val_words / 4; i += nthreads) z[i] = make_uint4(0, 0, 0, 0);
        }
        __threadfence();
        __syncthreads();
        const int job = s_job;
        if (job >= A.n_cubes) break;
        const long long c0 = A.cube_offsets[job], c1 = A.cube_offsets[job + 1];
        const int k = (int)(c1 - c0);
        const int32_t *cube = A.cube_lits + c0;
        int32_t *imp = A.implied + (long long)job * A.stride;

        for (int i0 = tid; i0 < k; i0 += 4 * nthreads) {   // phase 0: the whole cube is assigned up front,
            int x[4];                                       // four atomics in flight per thread
            uint32_t prev[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) x[u] = i0 + u * nthreads < k ? __ldg(cube + i0 + u * nthreads) : -1;
#pragma unroll
            for (int u = 0; u < 4; ++u) prev[u] = x[u] >= 0 ? bits.assign(x[u]) : 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (x[u] >= 0 && (prev[u] & 2u) && (prev[u] & 1u) != (uint32_t)(x[u] & 1)) s_conflict = 1;   // x and ~x
        }
        __syncthreads();
        if (tid == 0) {
            s_total = k;
            s_stop = s_conflict;
        }
        __syncthreads();

        long long visited = 0, words = 0;
        int qhead = 0;
        while (true) {
            const int total = s_total;
            if (s_stop || qhead >= total) break;
            for (int t = qhead + tid; t < total; t += nthreads) {
                const int p = t < k ? __ldg(cube + t) : __ldcg(imp + (t - k));
                const int f = p ^ 1;
                const int2 rg = A.stream_index ? __ldcs(A.orange + f) : __ldg(A.orange + f);
                const int os = rg.x, oe = rg.y;   // both even: the list is read as 16-byte loads of two entries
                for (int e0 = os; e0 < oe; e0 += GPSAT_SWEEP_CHUNK) {
                    const int cnt = min(GPSAT_SWEEP_CHUNK, oe - e0);
                    if (A.uniform3) {
                        // all loads of a chunk are issued before anything depends on them: pairs, then the assigned
                        // bits (shared), then the value fields (global) — three round trips per chunk, and a chunk of
                        // 8 entries covers most literals of a 3-SAT formula (6 occurrences on average)
                        int2 pr[GPSAT_SWEEP_CHUNK];
                        int va[GPSAT_SWEEP_CHUNK], vb[GPSAT_SWEEP_CHUNK];
#pragma unroll
                        for (int j = 0; j < GPSAT_SWEEP_CHUNK; j += 2) {
                            pr[j] = pr[j + 1] = make_int2(-1, -1);
                            if (j < cnt) {
                                const int4 *qp = reinterpret_cast<const int4 *>(A.occ_pair + e0 + j);
                                const int4 q = A.stream_index ? __ldcs(qp) : __ldg(qp);
                                pr[j] = make_int2(q.x, q.y);
                                pr[j + 1] = make_int2(q.z, q.w);
                            }
                        }
                        int real = 0;
#pragma unroll
                        for (int j = 0; j < GPSAT_SWEEP_CHUNK; ++j)
                            if (pr[j].x >= 0) {
                                va[j] = bits.value(pr[j].x);
                                vb[j] = bits.value(pr[j].y);
                                real++;
                            }
#pragma unroll
                        for (int j = 0; j < GPSAT_SWEEP_CHUNK; ++j)
                            if (pr[j].x >= 0) {
                                if (va[j] == 1 || vb[j] == 1) continue;
                                const int n_undef = (va[j] == 2) + (vb[j] == 2);
                                if (n_undef > 1) continue;
                                if (n_undef == 0) {   // every other literal false: conflict
                                    s_clause = __ldg(A.occ_clause + e0 + j);
                                    s_conflict = 1;
                                    continue;
                                }
                                const int unit = (va[j] == 2) ? pr[j].x : pr[j].y;
                                const uint32_t prev = bits.assign(unit);
                                if (prev == 0) {
                                    const int pos = atomicAdd(&s_count, 1);
                                    if (pos < A.stride) imp[pos] = unit;
                                } else if ((prev & 1u) != (uint32_t)(unit & 1)) {   // lost a race against ~unit
                                    s_clause = __ldg(A.occ_clause + e0 + j);
                                    s_conflict = 1;
                                }
                            }
                        visited += real;
                        words += 2 * real;
                    } else {
                        for (int j = 0; j < cnt; ++j) {
                            const int c = __ldg(A.occ_clause + e0 + j);
                            if (c < 0) continue;   // padding
                            const int lb = __ldg(A.coffsets + c), le = __ldg(A.coffsets + c + 1);
                            int unit = -1, n_undef = 0;
                            bool sat = false;
                            for (int i = lb; i < le && !sat; ++i) {
                                const int x = __ldg(A.clits + i);
                                words++;
                                if (x == f) continue;
                                const int v = bits.value(x);
                                if (v == 1) sat = true;
                                else if (v == 2) { n_undef++; unit = x; }
                            }
                            visited++;
                            if (sat || n_undef > 1) continue;
                            if (n_undef == 0) {
                                s_clause = c;
                                s_conflict = 1;
                                continue;
                            }
                            const uint32_t prev = bits.assign(unit);
                            if (prev == 0) {
                                const int pos = atomicAdd(&s_count, 1);
                                if (pos < A.stride) imp[pos] = unit;
                            } else if ((prev & 1u) != (uint32_t)(unit & 1)) {
                                s_clause = c;
                                s_conflict = 1;
                            }
                        }
                    }
                }
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                s_total = k + (int)min((long long)s_count, (long long)A.stride);
                s_stop = s_conflict;
            }
            __syncthreads();
            qhead = total;
        }
        for (int o = 16; o > 0; o >>= 1) {
            visited += __shfl_xor_sync(0xffffffffu, visited, o);
            words += __shfl_xor_sync(0xffffffffu, words, o);
        }
        if ((tid & 31) == 0 && A.counters) {
            atomicAdd((unsigned long long *)(A.counters + 2 * job), (unsigned long long)visited);
            atomicAdd((unsigned long long *)(A.counters + 2 * job + 1), (unsigned long long)words);
        }
        if (tid == 0) {
            const int n_imp = s_count, conflict = s_conflict;
            A.status[job] = conflict ? GPSAT_UNSAT : (n_imp > A.stride ? GPSAT_JOB_OOM : GPSAT_UNDEF);
            A.n_implied[job] = n_imp;
            A.conflict_clause[job] = conflict ? (long long)s_clause : -1;
        }
        __syncthreads();
    }
}


// ---------------------------------------------------------------------------------------------------------------
// gpsat_bcp_sweep_tern_kernel — occurrence-list BCP, one CTA per job, for pure 3-SAT databases of up to ~1.04 M
// variables: the WHOLE job state stays in the SM's shared memory and a literal costs one 32- or 64-byte memory access.
//
// State: base-3 digits, five variables per byte (3^5 = 243 <= 256): 0 unassigned, 1 false, 2 true — n = 1e6 needs
//   200 KB, which fits one SM where the 2-bit encoding (250 KB) does not.  A lookup is branch-free:
//   q = x / 10 (one IMAD.HI), byte = S[q], code = LUT[byte * 20 + (x - 10 q)] — the 5 KB table folds digit
//   extraction and the literal's sign into one load and returns 0 (false), 1 (unassigned) or 4 (true), so that a
//   clause with other literals (a, b) needs attention iff code(a) + code(b) <= 1.  Assigning is a compare-and-swap
//   on the 32-bit word that holds the byte; its outcome is the authority for "append to the trail exactly once"
//   (a cube without repeated variables is written with plain adds instead).
//   No global-memory state at all (the CTA kernel above keeps 2-bit fields in L2 behind a shared-memory filter: ncu
//   showed 38 % of its stall samples on those reads / atomics and 12 of 32 lanes active on average).
// Index: one 64-byte bucket per literal, the occurrence list packed INTO its head (host_formula.cpp:
//   build_sweep_index):
//   word 0            number of occurrences of the literal (8 bits, saturating) | half the index of its list in the
//                     plain pair list (24 bits): where entries 11.. are, without asking orange first
//   bits  32 .. 241   entries 0..4   } 42 bits per entry: the two OTHER literals of the clause, 21 bits each;
//   bits 256 .. 507   entries 5..10  } unused entries hold the always-true literal of a sentinel variable n
//   so a literal is ONE round trip of one or two adjacent sectors instead of head -> list (two dependent round trips,
//   a 32-byte sector for the 8-byte head plus the sectors an unaligned list straddles).  The 2 % of literals with
//   more than 11 occurrences (Poisson, mean 6) read entries 11.. from the plain pair list.
// Warp-uniform control flow: a warp takes 32 trail literals; every lane decodes and looks its bucket's entries up
//   with no branch (padding evaluates as "satisfied"), as many entries as the longest list of the batch has (the
//   host orders a cube's literals by list length, so the lists of a batch are about equally long), and marks the
//   clauses that became unit or conflicting (one literal in four has one) in a bit mask; the lanes with marks then
//   resolve them and rejoin the warp at the top of the next batch (__syncwarp).  The trail literal is fetched two
//   batches ahead.
// A lookup that misses an assignment made during the current round reads "unassigned"; the literal that was
// assigned is processed in a later round (after a barrier) and re-examines the clause, so no unit is lost.
// ---------------------------------------------------------------------------------------------------------------
#define GPSAT_TERN_LIT_BITS 21
#define GPSAT_TERN_ENTRIES 11
#define GPSAT_TERN_LUT_ROW 20          // bytes per table row: 5 words, odd, so that the few byte values that occur spread over the banks
#define GPSAT_TERN_LUT_BYTES 5120

// kPriv: the second lookup goes to a LANE-PRIVATE copy of the table (2-bit codes, 16 per word, word w of lane l at
// [w * 32 + l]: every lane reads its own bank, no conflict whatever the byte values are) — 20 KB instead of 5 KB
#define GPSAT_TERN_LUT2_WORDS 160      // 2560 codes (byte * 10 + rs), 16 per word
#define GPSAT_TERN_LUT2_BYTES (GPSAT_TERN_LUT2_WORDS * 32 * 4)

template <bool kPriv> struct TernJob {
    // dynamic shared memory: [LUT | state bytes]
    static constexpr uint32_t kLutBytes = kPriv ? GPSAT_TERN_LUT2_BYTES : GPSAT_TERN_LUT_BYTES;
    static constexpr uint32_t kTrue = kPriv ? 3u : 4u;
    uint8_t *smem;
    uint32_t lane;
    int32_t *imp;
    long long stride;
    int *s_count, *s_conflict, *s_conf_f, *s_conf_j;
    // code of the literal with in-byte index rs (2 * digit position + sign) given the state byte
    __device__ __forceinline__ uint32_t lut(uint32_t byte, uint32_t rs) const
    {
        if constexpr (kPriv) {
            const uint32_t e = byte * 10u + rs;
            return (reinterpret_cast<const uint32_t *>(smem)[(e >> 4) * 32u + lane] >> ((e & 15u) * 2u)) & 3u;
        } else {
            return smem[byte * GPSAT_TERN_LUT_ROW + rs];
        }
    }
    __device__ __forceinline__ int code(uint32_t x) const   // 0 false, 1 unassigned, kTrue true
    {
        const uint32_t q = __umulhi(x, 0x1999999Au);   // x / 10 (exact below 2^30)
        return (int)lut(smem[kLutBytes + q], x - 10u * q);
    }
    // previous digit of the variable: 0 it was unassigned (and now carries x), 1 it was false, 2 it was true
    __device__ __forceinline__ uint32_t assign(uint32_t x) const
    {
        const uint32_t var = x >> 1, q = var / 5u, r = var - 5u * q, sh = (q & 3u) * 8u;
        const uint32_t p3 = r == 4u ? 81u : (0x1B090301u >> (8u * r)) & 255u;   // 1, 3, 9, 27, 81
        uint32_t *w = reinterpret_cast<uint32_t *>(smem + kLutBytes) + (q >> 2);
        uint32_t old = *reinterpret_cast<volatile uint32_t *>(w);
        while (true) {
            const uint32_t c = lut((old >> sh) & 255u, 2u * r);   // code of the NEGATIVE literal of var
            if (c != 1u) return c == kTrue ? 1u : 2u;
            const uint32_t seen = atomicCAS(w, old, old + (((1u + (x & 1u)) * p3) << sh));
            if (seen == old) return 0u;
            old = seen;
        }
    }
    __device__ __forceinline__ void conflict_at(uint32_t f, int j) const
    {
        if (atomicCAS(s_conflict, 0, 1) == 0) {
            *s_conf_f = (int)f;
            *s_conf_j = j;
        }
    }
    // entry j of literal f's list, other literals (a, b): evaluate it now and act on it
    __device__ __forceinline__ void resolve(uint32_t a, uint32_t b, uint32_t f, int j) const
    {
        const int ca = code(a), cb = code(b);
        if (ca + cb > 1) return;
        if (ca + cb == 0) {   // every other literal false
            conflict_at(f, j);
            return;
        }
        resolve_unit(ca ? a : b, 1, f, j);
    }
    // a clause found unit (s == 1, on `unit`) or all-false (s == 0) when it was looked at; `unit` may have been assigned
    // by another thread since: assign() tells
    __device__ __forceinline__ void resolve_unit(uint32_t unit, int s, uint32_t f, int j) const
    {
        if (s == 0) {
            conflict_at(f, j);
            return;
        }
        const uint32_t prev = assign(unit);
        if (prev == 0) {
            const int pos = atomicAdd(s_count, 1);
            if (pos < stride) imp[pos] = (int32_t)unit;
        } else if (prev - 1u != (unit & 1u)) {   // lost a race against ~unit
            conflict_at(f, j);
        }
    }
};

template <int kOff> __device__ __forceinline__ uint32_t tern_field(const uint32_t (&w)[16])
{
    constexpr int wi = kOff >> 5, sh = kOff & 31;
    uint32_t v;
    if constexpr (sh + GPSAT_TERN_LIT_BITS <= 32) v = w[wi] >> sh;
    else v = __funnelshift_r(w[wi], w[wi + 1 < 16 ? wi + 1 : 15], sh);
    return v & ((1u << GPSAT_TERN_LIT_BITS) - 1u);
}
template <int kJ> struct TernEntry {
    static constexpr int bit = kJ < 5 ? 32 + 42 * kJ : 256 + 42 * (kJ - 5);
};

// entries kJ .. kEnd-1 of a bucket, branch-free: bit j of the result is set when entry j needs attention (unit /
// conflict)
// kFirst: the FIRST such entry is remembered as it goes by (two predicated moves per entry) — fu its literal that is
// not false (the unit), fs the code sum (0: conflict) — so that the usual case, one hit in the bucket, is resolved
// without the 11-way select of tern_pick and without a second evaluation
template <int kEnd, bool kFirst, int kJ = 0, class JobT>
__device__ __forceinline__ uint32_t tern_scan(const JobT &J, const uint32_t (&w)[16], uint32_t &fu, int &fs, uint32_t seen = 0u)
{
    if constexpr (kJ < kEnd) {
        const uint32_t a = tern_field<TernEntry<kJ>::bit>(w), b = tern_field<TernEntry<kJ>::bit + 21>(w);
        const int ca = J.code(a), s = ca + J.code(b);
        if constexpr (kFirst) {
            if (s <= 1 && seen == 0u) {
                fu = ca ? a : b;
                fs = s;
            }
        }
        const uint32_t now = seen | (s <= 1 ? 1u << kJ : 0u);
        return tern_scan<kEnd, kFirst, kJ + 1, JobT>(J, w, fu, fs, now);
    } else {
        return seen;
    }
}
// the two other literals of entry j (run-time j: a select over the 11 compile-time positions)
template <int kJ = 0>
__device__ __forceinline__ void tern_pick(const uint32_t (&w)[16], int j, uint32_t &a, uint32_t &b)
{
    if constexpr (kJ < GPSAT_TERN_ENTRIES) {
        if (j == kJ) {
            a = tern_field<TernEntry<kJ>::bit>(w);
            b = tern_field<TernEntry<kJ>::bit + 21>(w);
        }
        tern_pick<kJ + 1>(w, j, a, b);
    }
}

// kind 0: four 16-byte evict-first loads; 1: four 16-byte read-only loads; 2: two 32-byte read-only loads
// first_only (warp-uniform): every list of the batch has at most 5 entries — the second sector is not fetched
__device__ __forceinline__ void tern_load_bucket(const uint4 *bucket, uint32_t f, uint32_t (&w)[16], int kind, bool first_only)
{
    const uint4 *bp = bucket + 4 * (size_t)f;
    if (kind == 2) {
        asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(bp));
        if (first_only) return;
        asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]) : "l"(bp + 2));
        return;
    }
    uint4 q0, q1, q2, q3;
    if (kind == 0) { q0 = __ldcs(bp); q1 = __ldcs(bp + 1); q2 = __ldcs(bp + 2); q3 = __ldcs(bp + 3); }
    else { q0 = __ldg(bp); q1 = __ldg(bp + 1); q2 = __ldg(bp + 2); q3 = __ldg(bp + 3); }
    w[0] = q0.x; w[1] = q0.y; w[2] = q0.z; w[3] = q0.w;
    w[4] = q1.x; w[5] = q1.y; w[6] = q1.z; w[7] = q1.w;
    w[8] = q2.x; w[9] = q2.y; w[10] = q2.z; w[11] = q2.w;
    w[12] = q3.x; w[13] = q3.y; w[14] = q3.z; w[15] = q3.w;
}

template <bool kPrefetch, bool kPriv, bool kFirst>
__global__ void __launch_bounds__(1024, 1) gpsat_bcp_sweep_tern_kernel(const SweepArgs A)
{
    extern __shared__ __align__(16) uint8_t s_dyn[];   // [LUT | state bytes]
    __shared__ int s_count, s_conflict, s_conf_f, s_conf_j, s_job, s_total, s_stop;
    const int tid = (int)threadIdx.x, nthreads = (int)blockDim.x;
    const uint32_t lane = (uint32_t)tid & 31u, warp = (uint32_t)tid >> 5;
    const int state_bytes = A.tern_state_bytes;        // multiple of 16
    const uint32_t sentinel_true = 2u * (uint32_t)A.n_vars + 1u;   // literal of the sentinel variable n, always true
    typedef TernJob<kPriv> Job;
    constexpr uint32_t kLutBytes = Job::kLutBytes;
    Job J;
    J.smem = s_dyn;
    J.lane = lane;
    J.stride = A.stride;
    J.s_count = &s_count;
    J.s_conflict = &s_conflict;
    J.s_conf_f = &s_conf_f;
    J.s_conf_j = &s_conf_j;

    auto code_of = [](uint32_t byte, uint32_t rs) -> uint32_t {
        const uint32_t r = rs >> 1, sgn = rs & 1u;
        const uint32_t p3 = r == 0 ? 1u : r == 1 ? 3u : r == 2 ? 9u : r == 3 ? 27u : 81u;
        const uint32_t digit = r < 5 ? (byte / p3) % 3u : 0u;
        return digit == 0 ? 1u : (digit - 1u == sgn ? Job::kTrue : 0u);
    };
    if constexpr (kPriv) {
        for (int i = tid; i < GPSAT_TERN_LUT2_WORDS * 32; i += nthreads) {   // the 32 lanes' copies are identical
            const uint32_t e0 = ((uint32_t)i >> 5) * 16u;
            uint32_t word = 0;
            for (uint32_t c = 0; c < 16u; ++c) word |= code_of((e0 + c) / 10u, (e0 + c) % 10u) << (2u * c);
            reinterpret_cast<uint32_t *>(s_dyn)[i] = word;
        }
    } else {
        for (int i = tid; i < GPSAT_TERN_LUT_BYTES; i += nthreads)
            s_dyn[i] = (uint8_t)code_of((uint32_t)i / GPSAT_TERN_LUT_ROW, (uint32_t)i % GPSAT_TERN_LUT_ROW);
    }

    while (true) {
        if (tid == 0) {
            s_job = atomicAdd(A.next_job, 1);
            s_count = 0;
            s_conflict = 0;
        }
        {
            uint4 *z = reinterpret_cast<uint4 *>(s_dyn + kLutBytes);
            for (int i = tid; i < state_bytes / 16; i += nthreads) z[i] = make_uint4(0, 0, 0, 0);
        }
        __syncthreads();
        if (tid == 0) J.assign(sentinel_true);
        const int job = s_job;
        if (job >= A.n_cubes) break;
        const long long c0 = A.cube_offsets[job], c1 = A.cube_offsets[job + 1];
        const int k = (int)(c1 - c0);
        const int32_t *cube = A.cube_lits + c0;
        const int cube_info = A.cube_short ? __ldg(A.cube_short + job) : 0;
        const int n_short = cube_info & 0x3FFFFFFF;
        const bool distinct = (cube_info >> 30) & 1;   // the host checked: no variable occurs twice in this cube
        int32_t *imp = A.implied + (long long)job * A.stride;
        J.imp = imp;

        // phase 0: the whole cube is assigned up front
        if (distinct) {
            // every variable at most once: its digit goes from 0 to 1 or 2 with ONE fire-and-forget shared-memory add
            // (no carry can reach a neighbour's digit), four loads in flight per thread
            uint32_t *sw = reinterpret_cast<uint32_t *>(s_dyn + kLutBytes);
            for (int i0 = tid; i0 < k; i0 += 4 * nthreads) {
                uint32_t x[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) x[u] = i0 + u * nthreads < k ? (uint32_t)__ldg(cube + i0 + u * nthreads) : 0xFFFFFFFFu;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (x[u] != 0xFFFFFFFFu) {
                        const uint32_t var = x[u] >> 1, q = var / 5u, r = var - 5u * q;
                        const uint32_t p3 = r == 4u ? 81u : (0x1B090301u >> (8u * r)) & 255u;
                        atomicAdd(sw + (q >> 2), ((1u + (x[u] & 1u)) * p3) << ((q & 3u) * 8u));
                    }
            }
        } else {
            for (int i = tid; i < k; i += nthreads) {
                const uint32_t x = (uint32_t)__ldg(cube + i);
                const uint32_t prev = J.assign(x);
                if (prev && prev - 1u != (x & 1u)) s_conflict = 2;   // x and ~x in the cube: no clause to blame
            }
        }
        __syncthreads();
        if (tid == 0) {
            s_total = k;
            s_stop = s_conflict;
        }
        __syncthreads();

        long long visited = 0;
        int qhead = 0;
        while (true) {
            const int total = s_total;
            if (s_stop || qhead >= total) break;
            // the negation of trail literal t, or (past the end) the sentinel's false literal: an all-padding bucket
            auto trail_f = [&](int t) -> uint32_t {
                if (t >= total) return sentinel_true ^ 1u;
                return (uint32_t)(t < k ? __ldg(cube + t) : __ldcg(imp + (t - k))) ^ 1u;
            };
            int base = qhead + (int)warp * 32;
            // the trail literal is fetched two batches ahead and (kPrefetch: into registers; otherwise: into L2 with a
            // prefetch, which costs no register) its bucket one batch ahead
            uint32_t f = trail_f(base + (int)lane), f1 = trail_f(base + nthreads + (int)lane);
            uint32_t w[16], wn[16];
            if (kPrefetch) tern_load_bucket(A.bucket, f, w, A.stream_index, base + 32 <= n_short);
            for (; base < total; base += nthreads) {   // warp-uniform trip count
                __syncwarp();   // the lanes that resolved a hit rejoin here, not at the end of the round
                const uint32_t f2 = trail_f(base + 2 * nthreads + (int)lane);
                if (kPrefetch) {
                    tern_load_bucket(A.bucket, f1, wn, A.stream_index, base + nthreads + 32 <= n_short);
                } else {
                    if (A.l2_prefetch && base + nthreads < total)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(A.bucket + 4 * (size_t)f1));
                    tern_load_bucket(A.bucket, f, w, A.stream_index, base + 32 <= n_short);
                }
                // word 0: occurrences (8 bits, saturating) | half the index of the list's first entry in the plain pair
                // list (24 bits, all ones: too large, ask orange) — the tail of a long list is ONE further round trip,
                // started here so that it runs under the scan
                const int cnt = (int)(w[0] & 255u);
                const uint32_t tail_at = w[0] >> 8;
                if (cnt > GPSAT_TERN_ENTRIES && tail_at != 0xFFFFFFu)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(A.occ_pair + 2 * (size_t)tail_at + GPSAT_TERN_ENTRIES));
                visited += cnt;
                // the host sorts a cube's literals by occurrence count (gpsat_set_cubes), so the 32 lists of a batch
                // have nearly the same length and the scan stops at the longest of them instead of visiting padding
                const int cmax = __reduce_max_sync(0xffffffffu, cnt);
                uint32_t hits, fu = 0;
                int fs = 0;
                switch (cmax) {
                case 0: hits = 0u; break;
                case 1: case 2: hits = tern_scan<2, kFirst>(J, w, fu, fs); break;
                case 3: hits = tern_scan<3, kFirst>(J, w, fu, fs); break;
                case 4: hits = tern_scan<4, kFirst>(J, w, fu, fs); break;
                case 5: hits = tern_scan<5, kFirst>(J, w, fu, fs); break;
                case 6: hits = tern_scan<6, kFirst>(J, w, fu, fs); break;
                case 7: hits = tern_scan<7, kFirst>(J, w, fu, fs); break;
                case 8: hits = tern_scan<8, kFirst>(J, w, fu, fs); break;
                case 9: hits = tern_scan<9, kFirst>(J, w, fu, fs); break;
                default: hits = tern_scan<GPSAT_TERN_ENTRIES, kFirst>(J, w, fu, fs); break;
                }
                if (kFirst && hits) {   // one in four literals has an entry that needs attention, nearly always just one
                    const int j = __ffs((int)hits) - 1;
                    hits &= hits - 1u;
                    J.resolve_unit(fu, fs, f, j);
                }
                while (hits) {
                    const int j = __ffs((int)hits) - 1;
                    hits &= hits - 1u;
                    uint32_t a = 0, b = 0;
                    tern_pick<0>(w, j, a, b);
                    J.resolve(a, b, f, j);
                }
                if (cnt > GPSAT_TERN_ENTRIES) {   // 2 % of the literals: the tail of a long list comes from the plain pair index
                    long long os = 2 * (long long)tail_at;
                    int end = cnt;
                    if (cnt == 255 || tail_at == 0xFFFFFFu) {   // saturated fields: (begin, end) of the padded list
                        const int2 oc = __ldg(A.orange + f);
                        os = oc.x;
                        end = oc.y - oc.x;
                        if (__ldg(A.occ_pair + oc.y - 1).x < 0) end--;   // padded to an even length with -1
                        visited += end - cnt;
                    }
                    for (int j = GPSAT_TERN_ENTRIES; j < end; ++j) {
                        const int2 q = __ldg(A.occ_pair + os + j);
                        J.resolve((uint32_t)q.x, (uint32_t)q.y, f, j);
                    }
                }
                if (kPrefetch) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) w[i] = wn[i];
                }
                f = f1;
                f1 = f2;
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                s_total = k + (int)min((long long)s_count, (long long)A.stride);
                s_stop = s_conflict;
            }
            __syncthreads();
            qhead = total;
        }
        for (int o = 16; o > 0; o >>= 1) visited += __shfl_xor_sync(0xffffffffu, visited, o);
        if (lane == 0 && A.counters) {
            atomicAdd((unsigned long long *)(A.counters + 2 * job), (unsigned long long)visited);
            atomicAdd((unsigned long long *)(A.counters + 2 * job + 1), (unsigned long long)(2 * visited));
        }
        if (tid == 0) {
            const int n_imp = s_count, conflict = s_conflict;
            A.status[job] = conflict ? GPSAT_UNSAT : (n_imp > A.stride ? GPSAT_JOB_OOM : GPSAT_UNDEF);
            A.n_implied[job] = n_imp;
            long long cc = -1;
            if (conflict == 1) cc = __ldg(A.occ_clause + __ldg(A.orange + s_conf_f).x + s_conf_j);
            A.conflict_clause[job] = cc;
        }
        __syncthreads();
    }
}

typedef void (*cta_sweep_kernel_t)(const SweepArgs, const int);
// register budgets: 1 CTA x 1024 threads (64 registers), 2 x 768 (42), 2 x 1024 (32)
cta_sweep_kernel_t pick_cta_sweep(int threads, int per_sm)
{
    if (per_sm >= 2 && threads > 768) return gpsat_bcp_sweep_cta_kernel<1024, 2>;
    if (per_sm >= 2) return gpsat_bcp_sweep_cta_kernel<768, 2>;
    return gpsat_bcp_sweep_cta_kernel<1024, 1>;
}

}  // namespace

namespace gpsat_kernels {

size_t tern_smem_bytes(int32_t state_bytes, bool private_lut)
{
    return (size_t)state_bytes + (private_lut ? GPSAT_TERN_LUT2_BYTES : GPSAT_TERN_LUT_BYTES);
}

cudaError_t launch_bcp_sweep(const SweepLaunch &L, cudaStream_t stream)
{
    SweepArgs A;
    A.n_vars = L.n_vars;
    A.n_clauses = L.n_clauses;
    A.n_cubes = L.n_cubes;
    A.uniform3 = L.uniform3;
    A.stream_index = L.stream_index;
    A.orange = (const int2 *)L.ostart;
    A.occ_clause = L.occ_clause;
    A.occ_pair = (const int2 *)L.occ_pair;
    A.bucket = (const uint4 *)L.bucket;
    A.tern_state_bytes = L.tern_state_bytes;
    A.cube_short = L.cube_short;
    A.l2_prefetch = L.l2_prefetch;
    A.coffsets = L.coffsets;
    A.clits = L.clits;
    A.cube_offsets = L.cube_offsets;
    A.cube_lits = L.cube_lits;
    A.valbits = L.valbits;
    A.val_words = L.val_words;
    A.implied = L.implied;
    A.stride = L.stride;
    A.n_implied = L.n_implied;
    A.status = L.status;
    A.conflict_clause = L.conflict_clause;
    A.counters = L.counters;
    A.next_job = L.next_job;
    if (L.bucket) {   // ternary kernel: whole job state in shared memory, bucket index
        const size_t smem = tern_smem_bytes(L.tern_state_bytes, L.tern_private_lut != 0);
        typedef void (*tern_kernel_t)(const SweepArgs);
        static const tern_kernel_t table[8] = {
            gpsat_bcp_sweep_tern_kernel<false, false, false>, gpsat_bcp_sweep_tern_kernel<true, false, false>,
            gpsat_bcp_sweep_tern_kernel<false, true, false>,  gpsat_bcp_sweep_tern_kernel<true, true, false>,
            gpsat_bcp_sweep_tern_kernel<false, false, true>,  gpsat_bcp_sweep_tern_kernel<true, false, true>,
            gpsat_bcp_sweep_tern_kernel<false, true, true>,   gpsat_bcp_sweep_tern_kernel<true, true, true>};
        tern_kernel_t kfn = table[(L.tern_prefetch ? 1 : 0) | (L.tern_private_lut ? 2 : 0) | (L.tern_first_hit ? 4 : 0)];
        cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kfn<<<L.blocks, 1024, smem, stream>>>(A);
        return cudaGetLastError();
    }
    if (L.cluster_size < 0) {   // one CTA per job, assigned-bit filter in shared memory, values in this CTA's global block
        const size_t smem = ((size_t)1 << (L.slice_log2 - 3)) + 16;
        cta_sweep_kernel_t kfn = pick_cta_sweep(L.warps_per_block * 32, -L.cluster_size);
        cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kfn<<<L.blocks, L.warps_per_block * 32, smem, stream>>>(A, L.slice_log2);
        return cudaGetLastError();
    }
    return cudaErrorInvalidConfiguration;
}

cudaError_t sweep_cta_capacity(int filter_log2, int threads, int want_per_sm, int *blocks_per_sm)
{
    const size_t smem = ((size_t)1 << (filter_log2 - 3)) + 16;
    cta_sweep_kernel_t kfn = pick_cta_sweep(threads, want_per_sm);
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kfn, threads, smem);
}

}  // namespace gpsat_kernels
