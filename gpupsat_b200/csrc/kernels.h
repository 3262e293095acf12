// Launchers of the sm_100a kernels (kernels.cu), called by the C-ABI layer (gpsat_api.cu).
#pragma once
#include <cuda_runtime.h>
#include "gpsat_device.h"

namespace gpsat_kernels {

// Persistent warp-per-cube CDCL / BCP kernel.  grid = blocks, block = warps_per_block*32 threads,
// smem_bytes = warps_per_block * layout.total_words * 4 when the per-job state lives in shared memory, else 0.
cudaError_t launch_cdcl(const gpsat_formula_view &F, const gpsat_solve_params &P, const gpsat_state_layout &Ly,
                        const gpsat_run_buffers &B, int blocks, int warps_per_block, size_t smem_bytes,
                        cudaStream_t stream);
// resident blocks per SM for that configuration (cudaOccupancyMaxActiveBlocksPerMultiprocessor)
cudaError_t cdcl_occupancy(int warps_per_block, size_t smem_bytes, bool smem_state, bool smem_formula,
                           int *blocks_per_sm);
cudaError_t cdcl_attributes(int *regs_per_thread, size_t *local_bytes);
int cdcl_max_warps_per_block();

// queue state of a solve initialised on the device: empty ring, control block, root arrays, foreign-pool cursor, facts,
// run_ctrl = [-, -, sat_job, -], t0 = [launch stamp, busy time].  owner = this rank's control block holds the root
// cursor (single GPU, or rank 0 of a mesh): created = n_roots and one open job per root are booked here.
cudaError_t launch_queue_init(int *ctrl, int *meta, int dq_cap, int *root_pending, int *root_flag, int n_roots, int owner,
                              int *xcur, unsigned char *facts, int n_vars, int *run_ctrl, unsigned long long *t0,
                              cudaStream_t stream);
// stamps *t0 with the GPU's globaltimer (deadline base for budgeted steps)
cudaError_t launch_stamp(unsigned long long *t0, cudaStream_t stream);

// epoch exchange between GPUs: pack this GPU's fresh pool slots + status header into an exchange block; append the
// other ranks' slots (from the all-gathered blocks) to the foreign pool
cudaError_t launch_xchg_pack(const int *pool, int *pool_cursor, int pool_cap_slots, int *block, int block_words,
                             int rank, int done, int verdict, int jobs_done, cudaStream_t stream);
cudaError_t launch_xchg_unpack(const int *blocks, int n_ranks, int my_rank, int block_words, int *xpool,
                               int *xpool_cursor, int xpool_cap_slots, unsigned char *facts, int n_vars,
                               cudaStream_t stream);

// clause evaluation: one thread per (assignment, clause)
cudaError_t launch_eval_clauses(int32_t n_vars, int32_t n_clauses, const int32_t *coffsets, const int32_t *clits,
                                int32_t n_assignments, const uint8_t *assignment, int32_t *status, int32_t *unit,
                                cudaStream_t stream);

// BCP by clause evaluation over occurrence lists for large clause databases (one CTA per job)
struct SweepLaunch {
    int32_t n_vars, n_clauses, n_cubes, uniform3;
    const int32_t *ostart;   // (begin, end) pairs of the padded occurrence lists (int2 per literal)
    const int32_t *occ_clause, *occ_pair, *coffsets, *clits;
    // ternary kernel (pure 3-SAT, whole job state in one SM's shared memory): 64-byte bucket per literal id, 2n + 2 of
    // them (16 words each), and the bytes of base-3 state; nullptr selects the other kernels
    const uint32_t *bucket = nullptr;
    int32_t tern_state_bytes = 0;
    const int32_t *cube_short = nullptr;   // per cube: leading literals (sorted order) whose lists have at most 5 entries
    int tern_prefetch = 0;
    int tern_first_hit = 0;     // the first hit of a bucket is resolved from literals kept during the scan
    int tern_private_lut = 0;   // second lookup in a lane-private copy of the code table (20 KB, no bank conflicts)
    int l2_prefetch = 0;                // bucket of the next batch prefetched into L2 (no register cost; no gain measured)              // bucket fetched one batch ahead (registers), trail literal two
    const int64_t *cube_offsets;
    const int32_t *cube_lits;
    uint32_t *valbits;
    int32_t val_words;
    int32_t *implied;
    int64_t stride;
    int32_t *n_implied, *status;
    int64_t *conflict_clause;
    int64_t *counters;
    int32_t *next_job;
    int blocks, warps_per_block;
    int stream_index;              // index loads: 0 evict-first 16-byte, 1 read-only 16-byte, 2 read-only 32-byte
    int cluster_size;              // -(CTAs per SM the CTA-filter kernel variant is compiled for): -1 or -2
    int slice_log2;                // CTA-filter kernel: log2 of the filter bits
};
cudaError_t launch_bcp_sweep(const SweepLaunch &L, cudaStream_t stream);
size_t tern_smem_bytes(int32_t state_bytes, bool private_lut = false);   // dynamic shared memory of the ternary kernel
cudaError_t sweep_cta_capacity(int filter_log2, int threads, int want_per_sm, int *blocks_per_sm);

}  // namespace gpsat_kernels
