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

#define GPSAT_MAX_THREADS 768   // 24 warps per block: up to 85 registers per thread

namespace {

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void gpsat_stamp_kernel(unsigned long long *t0) { *t0 = globaltimer_ns(); }

// kSmemState:   the per-job state blocks live in dynamic shared memory.
// kSmemFormula: the read-only formula index (cl2, occ2, ostart) is staged once per block in shared memory, in front
//               of the state blocks, and every warp of the block reads it from there.
// Both are template parameters (not run-time selects) so that the pointers provably derive from the shared window
// and compile to LDS/STS/ATOMS instead of generic loads.
template <bool kSmemState, bool kSmemFormula>
__global__ void __launch_bounds__(GPSAT_MAX_THREADS, 1)
gpsat_cdcl_kernel(const gpsat_formula_view F, const gpsat_solve_params P, const gpsat_state_layout Ly,
                  const gpsat_run_buffers B)
{
    extern __shared__ __align__(16) int gpsat_smem[];
    const int warp_in_block = (int)(threadIdx.x >> 5);
    const int warps_per_block = (int)(blockDim.x >> 5);
    const long long gwarp = (long long)blockIdx.x * warps_per_block + warp_in_block;

    gpsat_formula_view Fv = F;
    int *state_base = gpsat_smem;
    if (kSmemFormula) {
        // layout: cl2 | occ2 | ostart, each rounded up to 4 words
        const int n_cl2 = 2 * (F.n_lits + F.n_clauses), n_occ2 = 2 * F.n_lits, n_os = 2 * F.n_vars + 1;
        int *s_cl2 = gpsat_smem;
        int *s_occ2 = s_cl2 + ((n_cl2 + 3) & ~3);
        int *s_os = s_occ2 + ((n_occ2 + 3) & ~3);
        const int *g_cl2 = (const int *)F.cl2, *g_occ2 = (const int *)F.occ2;
        for (int i = (int)threadIdx.x; i < n_cl2; i += (int)blockDim.x) s_cl2[i] = g_cl2[i];
        for (int i = (int)threadIdx.x; i < n_occ2; i += (int)blockDim.x) s_occ2[i] = g_occ2[i];
        for (int i = (int)threadIdx.x; i < n_os; i += (int)blockDim.x) s_os[i] = F.ostart[i];
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

    WarpSolver S;
    gpsat_bind(S, Fv, P, Ly, state, arena, B);
    gpsat_warp_loop(S, P, B);
}

typedef void (*cdcl_kernel_t)(const gpsat_formula_view, const gpsat_solve_params, const gpsat_state_layout,
                              const gpsat_run_buffers);
cdcl_kernel_t pick_kernel(bool smem_state, bool smem_formula)
{
    if (smem_state && smem_formula) return gpsat_cdcl_kernel<true, true>;
    if (smem_state) return gpsat_cdcl_kernel<true, false>;
    return gpsat_cdcl_kernel<false, false>;
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

}  // namespace

namespace gpsat_kernels {

cudaError_t launch_cdcl(const gpsat_formula_view &F, const gpsat_solve_params &P, const gpsat_state_layout &Ly,
                        const gpsat_run_buffers &B, int blocks, int warps_per_block, size_t smem_bytes,
                        cudaStream_t stream)
{
    if (warps_per_block * 32 > GPSAT_MAX_THREADS) return cudaErrorInvalidConfiguration;
    cdcl_kernel_t k = pick_kernel(B.state_in_smem != 0, B.formula_in_smem != 0);
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
    cdcl_kernel_t k = pick_kernel(smem_state, smem_formula);
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
    cudaError_t e = cudaFuncGetAttributes(&a, gpsat_cdcl_kernel<true, true>);
    if (e != cudaSuccess) return e;
    *regs_per_thread = a.numRegs;
    *local_bytes = a.localSizeBytes;
    return cudaSuccess;
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
