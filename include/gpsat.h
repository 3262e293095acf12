/*
 * gpsat.h — C ABI of libgpsat (gpupsat_b200): the B200-native drop-in for the data-parallel hot path of
 * nvzoll/gpupsat (two-watched-literal BCP + clause evaluation + CDCL over independent cubes).
 *
 * The reference has no FFI; its seams are C++ classes and four __global__ kernels.  Each entry
 * point below names the reference interface it replaces (paths relative to the reference's src/).
 *
 * Conventions
 *   - literals: int32, reference encoding  x = 2*var + (positive ? 1 : 0), vars 0-based      (SATSolver/SolverTypes.cu:6-32)
 *   - status / verdict / per-variable value: the reference's sat_status                       (SATSolver/SolverTypes.cuh:125)
 *         GPSAT_SAT = 0 (true), GPSAT_UNSAT = 1 (false), GPSAT_UNDEF = 2 (unassigned / no verdict)
 *   - every buffer is caller-owned; functions return GPSAT_OK (0) or a negative gpsat_error and never exit()
 *     (the reference's check() prints, cudaDeviceReset()s and exit(1)s: ErrorHandler/CudaMemoryErrorHandler.cu:3-10)
 *   - one handle is used from one host thread and owns one GPU; several GPUs = several handles joined in a mesh
 *     (gpsat_mesh_*: one process per GPU, or gpsat_multi_*: one process, one host thread per GPU)
 *   - there is NO CPU fallback: device entry points fail with GPSAT_E_NO_DEVICE when no CUDA device is usable
 */
#ifndef GPSAT_H
#define GPSAT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPSAT_SAT   0
#define GPSAT_UNSAT 1
#define GPSAT_UNDEF 2

typedef enum gpsat_error {
    GPSAT_OK            =  0,
    GPSAT_E_ARG         = -1,   /* bad argument / malformed formula (unit or empty clause, duplicate literal, ...) */
    GPSAT_E_NO_DEVICE   = -2,   /* no usable CUDA device: the product path has no CPU fallback */
    GPSAT_E_CUDA        = -3,   /* a CUDA call failed; see gpsat_last_error() */
    GPSAT_E_IO          = -4,   /* file missing / unreadable */
    GPSAT_E_PARSE       = -5,   /* DIMACS syntax the reference's grammar rejects */
    GPSAT_E_CAPACITY    = -6,   /* caller buffer too small / per-job arena exhausted */
    GPSAT_E_STATE       = -7    /* call order (e.g. solve before set_cubes) */
} gpsat_error;

const char *gpsat_last_error(void);          /* thread-local text of the last failure */
const char *gpsat_version(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Host side: formula container, DIMACS reader, host preprocessing, cube generation.
 * Pure host code (no GPU needed); mirrors the reference's L4/L3 layers because they decide what the hot path sees.
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct gpsat_cnf gpsat_cnf;

/* ≙ CnfManager::read_cnf + cnf_reader::parse_cnf (FileManager/CnfReader.cpp:30-50,86-135): header is read but not
 * trusted, n_vars = highest variable seen, comments only before the "p cnf" line, one clause per line ending in 0. */
int gpsat_cnf_read(const char *path, gpsat_cnf **out);
/* same container from arrays (offsets has n_clauses+1 entries) */
int gpsat_cnf_from_arrays(int64_t n_clauses, const int64_t *offsets, const int32_t *lits, gpsat_cnf **out);
void gpsat_cnf_free(gpsat_cnf *f);

/* ≙ FormulaData::copy_host_clauses_to_dev (FileManager/FormulaData.cu:82-105) = RepeatedLiteralsRemover::process
 * (Preprocessing/RepeatedLiteralsRemover.cu:25-62) then UnaryClausesRemover::process (Preprocessing/
 * UnaryClausesRemover.cu:13-173).  *out is a new container holding the reduced formula (same clause order), the
 * solved literals in discovery order, and the status after preprocessing (GPSAT_UNDEF = still to solve). */
int gpsat_cnf_preprocess(const gpsat_cnf *in, gpsat_cnf **out);

int32_t        gpsat_cnf_n_vars(const gpsat_cnf *f);              /* highest var + 1 (FormulaData::get_n_vars) */
int64_t        gpsat_cnf_n_clauses(const gpsat_cnf *f);
int64_t        gpsat_cnf_n_lits(const gpsat_cnf *f);
const int64_t *gpsat_cnf_offsets(const gpsat_cnf *f);
const int32_t *gpsat_cnf_lits(const gpsat_cnf *f);
int32_t        gpsat_cnf_status(const gpsat_cnf *f);              /* FormulaData::get_status_after_preprocessing */
int32_t        gpsat_cnf_n_solved(const gpsat_cnf *f);            /* FormulaData::get_solved_literals().size() */
const int32_t *gpsat_cnf_solved(const gpsat_cnf *f);
int32_t        gpsat_cnf_header_vars(const gpsat_cnf *f);
int64_t        gpsat_cnf_header_clauses(const gpsat_cnf *f);
int32_t        gpsat_cnf_largest_clause(const gpsat_cnf *f);      /* FormulaData::get_largest_clause_size (pre-preprocessing) */
int32_t        gpsat_cnf_most_common_var(const gpsat_cnf *f);     /* FormulaData::get_most_common_var */
int32_t        gpsat_cnf_most_common_freq(const gpsat_cnf *f);
int32_t        gpsat_cnf_n_lines(const gpsat_cnf *f);             /* n_lines() of the file (FileManager/FileUtils.cu:4-27) */

#define GPSAT_STRATEGY_DISTRIBUTED 0   /* ChoosingStrategy::DISTRIBUTE_JOBS_PER_THREAD */
#define GPSAT_STRATEGY_UNIFORM     1   /* ChoosingStrategy::UNIFORM */
#define GPSAT_STRATEGY_SIMPLE      2   /* SimpleJobChooser (JobsManager/SimpleJobChooser.cu:22-75; USE_SIMPLE_JOBS_GENERATION, off as shipped) */
/* ≙ MaxClauseJobChooser::evaluate/getJobs + VariableChooser::evaluate (JobsManager/JobChooser.cu:52-133,
 * JobsManager/VariableChooser.cu:23-40).  `pre` is a preprocessed formula.  Writes vars-per-job and the job count
 * (2^k); when cube_lits != NULL fills n_cubes*k literals (cube j, position i positive iff bit (k-1-i) of j is 0). */
int gpsat_choose_cubes(const gpsat_cnf *pre, int32_t blocks, int32_t threads, int32_t strategy,
                       int32_t *vars_per_job, int32_t *n_cubes, int32_t *cube_lits, int64_t cube_lits_cap);

/* ------------------------------------------------------------------------------------------------------------------
 * Device side: the hot path.
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct gpsat gpsat_t;

#define GPSAT_DECIDE_REFERENCE 0   /* shipped rule: a free variable from the top, positive (SATSolver/DecisionMaker.cu:45-55) */
#define GPSAT_DECIDE_VSIDS     1   /* per-literal integer counters, halved every 50 learnt clauses (DecisionStrategy/VSIDS.cu:77-124) */

#define GPSAT_BCP_WATCHED    0     /* two watched literals (BCPStrategy/WatchedClausesList.cu:46-221) */
#define GPSAT_BCP_OCCURRENCE 1     /* clause evaluation over occurrence lists (ConflictAnalysis/ConflictAnalyzer.cu:173-245 made incremental) */

typedef struct gpsat_opts {
    int32_t struct_size;          /* = sizeof(gpsat_opts), set by gpsat_opts_default */
    int32_t device;               /* CUDA ordinal; -1 = current device */
    int32_t decision;             /* GPSAT_DECIDE_* (default VSIDS) */
    int32_t bcp;                  /* GPSAT_BCP_* (default WATCHED) */
    int32_t restart_first;        /* conflicts before the first restart, 0 = off (Configs.cuh:103-107: 100) */
    float   restart_factor;       /* geometric factor, int-truncated (Restarts/GeometricRestartsManager.cu:16-20: 1.3) */
    int32_t max_iterations;       /* decisions per job before GPSAT_UNDEF; 0 = none (Configs.cuh:23 ships 1000) */
    int32_t stop_on_sat;          /* 1: first satisfied cube ends the run (Parallelizer.cu:213-227) */
    int64_t max_conflicts;        /* per job, 0 = none */
    int32_t share_learnts;        /* 1: short 1-UIP clauses go to the per-GPU pool and are imported by later jobs */
    int32_t share_max_len;        /* longest clause exported to the pool (at most 15: pool slots are 16 words) */
    int32_t warps_per_block;      /* 0 = auto (as many as the state blocks allow, at most 24); up to 28 on request */
    int32_t blocks;               /* 0 = auto (SM count x resident blocks) */
    int64_t arena_words;          /* per-warp learnt-clause arena (int32 words); 0 = auto */
    int32_t dynamic_split;        /* 1 (default): at a restart a long-running cube hands half of its remaining search
                                     space to an idle warp (the reference's wave-synchronous loop waits for the slowest
                                     job instead: SATSolver/main.cu:259-269).  Verdicts are unaffected; per-cube counters
                                     then depend on timing, so parity runs set 0. */
    int32_t split_gap;            /* conflicts a cube runs between two rounds of splitting; 0 = default (8) */
    int32_t split_burst;          /* children handed out per round while warps are idle; 0 = default (4) */
    int32_t share_import_max;     /* non-unit shared clauses a cube imports per pool when it starts; 0 = default (256) */
    int32_t split_hand_words;     /* learnt-clause words a split-off cube inherits from its parent; 0 = default */
    int32_t split_gap_hot;        /* gap used instead of split_gap while more than 1/8 of the GPU's warps are idle; 0 = split_gap / 2 */
    int32_t split_at_start;       /* 1: a cube may split before its first conflict while more than 1/8 of the warps are idle
                                     (default 0: measured slower on C2) */
    int32_t mesh_flags;           /* test hooks: 1 = never take children of other GPUs, 2 = never push clauses to them */
    int32_t sweep_flags;          /* test hooks (GPSAT_BCP_OCCURRENCE): 1 = general kernel even for pure 3-SAT; 2 = bucket one batch ahead in
                                     registers; 4 = no L2 prefetch of the next buckets; 32 = lane-private code table; 64 = the first hit of a
                                     bucket is NOT kept during the scan; bits 8..12 = log2 of the assigned-bit filter (smaller than the
                                     variable count = aliased filter) */
    int32_t split_mode;           /* 0 (default): back to the cube, branch on the VSIDS-best literal p, keep p, hand out ~p;
                                     1: guiding path — hand out the untried side of the OLDEST open decision and keep searching
                                     where the cube is; 2: as 0 but keep ~p (measured on C2: DESIGN.md section 3) */
    int32_t max_learnts;          /* learnt clauses a job keeps before its first database reduction; 0 = default */
    int32_t split_min;            /* a cube splits only once it has proved hard: conflicts (its own + half of what its parent
                                     had when it was split off) before its first split; 0 = default */
    int32_t split_reserve;        /* split-off cubes kept queued AHEAD of demand, so that a warp that runs out of work finds one at once
                                     instead of waiting for a busy cube's next conflict; 0 = default, -1 = none */
    int32_t phase_stats;          /* 1: time and count per solver phase (gpsat_phase_stats), the counterpart of the reference's
                                     RuntimeStatistics timers (Statistics/RuntimeStatistics.cuh:17-66); default 0 */
    int32_t split_hard;           /* hardness from which a cube splits after EVERY conflict (and right when it starts) while warps
                                     are idle: the few very hard cubes that decide the tail of a run; 0 = default, -1 = off */
} gpsat_opts;

void gpsat_opts_default(gpsat_opts *o);

/* per-job record (one per cube), filled by gpsat_solve / gpsat_propagate_all; all counters are exact and, with
 * share_learnts = 0, a pure function of (formula, cube, opts) — the parity tests compare them with the oracle. */
typedef struct gpsat_job_record {
    int32_t status;               /* GPSAT_SAT / UNSAT / UNDEF; -1 = not run (run ended early) */
    int32_t reserved;             /* number of times this cube was split (dynamic_split) */
    int64_t decisions;
    int64_t implications;         /* literals assigned by unit propagation (≙ VariablesStateHandler::new_implication) */
    int64_t conflicts;
    int64_t learnt_clauses;
    int64_t learnt_literals;
    int64_t restarts;
    int64_t watchers_visited;     /* watched (or occurrence) entries examined */
    int64_t clause_words_read;    /* clause literals read while examining them */
    int64_t learnt_hash;          /* order-sensitive checksum of every learnt clause, for bit-exact parity */
} gpsat_job_record;

typedef struct gpsat_stats {
    int64_t jobs_total, jobs_done, jobs_sat, jobs_unsat, jobs_undef;
    int64_t decisions, implications, conflicts, learnt_clauses, learnt_literals, restarts;
    int64_t watchers_visited, clause_words_read;
    int64_t pool_clauses;         /* clauses in the per-GPU shared pool at the end of the run */
    double  kernel_ms;            /* CUDA-event time of the solve kernel(s) on the handle's stream */
    int32_t kernel_launches;
    int32_t blocks, warps_per_block, smem_bytes_per_block;
    int32_t state_in_smem;        /* 1: per-job assignment/trail/watch bitmap live in shared memory */
    int32_t reserved;
    int64_t splits;               /* children created by dynamic splitting */
    double  warp_busy_frac;       /* sum of the time warps spent inside jobs / (warps x kernel time) */
    int64_t foreign_clauses;      /* clauses received from other GPUs (exchange, pool import, or pushed over NVLink by a mesh peer) */
    int64_t steals;               /* mesh: children of cubes split on another GPU that this GPU took over NVLink peer memory */
} gpsat_stats;

/* ≙ DataToDevice ctor + CUDAClauseVec::alloc_and_copy_to_dev (SATSolver/DataToDevice.cu:11-56, Utils/CUDAClauseVec.cu:85-118):
 * uploads the (preprocessed: no unit/empty clauses, no repeated variables inside a clause) formula as packed int32 CSR plus
 * the occurrence index.  dead variables (solved in preprocessing) simply do not occur. */
int gpsat_create(gpsat_t **h, int32_t n_vars, int64_t n_clauses, const int64_t *offsets, const int32_t *lits,
                 const gpsat_opts *opts);
void gpsat_destroy(gpsat_t *h);

/* ≙ DataToDevice::prepare_parallel + JobsQueue::add/close (SATSolver/DataToDevice.cu:60-95, SATSolver/JobsQueue.cu:34-63).
 * n_cubes = 0 with NULL arrays = one empty cube (the reference's sequential mode, Parallelizer.cu:230-278). */
int gpsat_set_cubes(gpsat_t *h, int32_t n_cubes, const int64_t *cube_offsets, const int32_t *cube_lits);

/* ≙ ConflictAnalyzerWithWatchedLits::set_assumptions as driven by SATSolver::preprocess (ConflictAnalysis/
 * ConflictAnalyzerWithWatchedLits.cu:46-111, SATSolver/SATSolver.cu:231-246): BCP of every cube from an empty trail to
 * fixpoint or first conflict.  status[j] = GPSAT_UNDEF (fixpoint) or GPSAT_UNSAT (conflict); implied literals of cube j
 * (trail order, cube variables excluded) are written at implied[implied_offsets[j] ...], implied_offsets has n_cubes+1
 * entries and is computed by the library (stride = min(n_vars, implied_cap/n_cubes)); conflict_clause[j] = index of the
 * falsified clause or -1.  Any output pointer may be NULL. */
int gpsat_propagate_all(gpsat_t *h, int32_t *status, int32_t *n_implied, int32_t *implied, int64_t implied_stride,
                        int64_t *conflict_clause, gpsat_job_record *records);
/* single-cube convenience with the exact shape of SURVEY.md §8b */
int gpsat_propagate(gpsat_t *h, int32_t cube, int32_t *status, int32_t *implied, int32_t *n_implied,
                    int64_t *conflict_clause);

/* ≙ VariablesStateHandler::clause_status over the whole formula (SATSolver/VariablesStateHandler.cu:180-206), the
 * evaluation step of ConflictAnalyzer::propagate_all_clauses (ConflictAnalysis/ConflictAnalyzer.cu:173-245).
 * assignment[v] in {GPSAT_SAT(true), GPSAT_UNSAT(false), GPSAT_UNDEF}; n_assignments independent assignments are
 * evaluated in one launch (assignment is [n_assignments][n_vars], outputs [n_assignments][n_clauses]).
 * unit_lit (optional) gets the single unassigned literal of a unit clause, else -1. */
int gpsat_eval_clauses(gpsat_t *h, int32_t n_assignments, const uint8_t *assignment, int32_t *status_per_clause,
                       int32_t *unit_lit);

/* ≙ parallel_kernel_init + the parallel_kernel relaunch loop + parallel_kernel_retrieve_results (SATSolver/Parallelizer.cu:
 * 133-228, SATSolver/main.cu:256-285), or run_sequential for the empty cube.  verdict: GPSAT_SAT if some cube is
 * satisfiable (model[v] = 1/0, a total assignment that satisfies the uploaded formula), GPSAT_UNSAT if every cube is
 * refuted, GPSAT_UNDEF if a cap (max_iterations / max_conflicts) stopped some job and none was SAT. */
int gpsat_solve(gpsat_t *h, int32_t *verdict, uint8_t *model, gpsat_stats *stats);

/* ≙ RuntimeStatistics (Statistics/RuntimeStatistics.cuh:17-66, printed by print_function_time_statistics,
 * Statistics/RuntimeStatistics.cu:299-360): time (ns, summed over warps, GPU globaltimer) and number of runs per phase
 * of the last solve on a handle created with opts.phase_stats = 1.  The reference keeps one clock64 total per thread and
 * phase; here one total per GPU and phase.  Mapping of the reference's phases: "job" = job_ns / jobs, "pre-processing"
 * (cube placement, SATSolver::preprocess) = RESET + IMPORT, "decision" = DECIDE, "conflict analyzing" (its name for
 * propagate-until-no-conflict, ConflictAnalyzer::propagate) = PROPAGATE + ANALYZE, "backtracking" = BACKTRACK, "structures
 * reset" = RESET, "next job" = idle_ns (warps waiting for a cube); the reference's "creating structures", "add jobs to
 * assumptions", "processing results" and the three pre-processing sub-phases have no counterpart (no per-thread object
 * construction, cubes are read in place, records are a few atomics): reported as not run. */
#define GPSAT_N_PHASES 8
#define GPSAT_PHASE_RESET 0        /* per-job state reset (≙ reset_structures) */
#define GPSAT_PHASE_IMPORT 1       /* hand-off block / shared clauses attached at job start */
#define GPSAT_PHASE_PROPAGATE 2    /* two-watched-literal BCP */
#define GPSAT_PHASE_ANALYZE 3      /* first-UIP analysis */
#define GPSAT_PHASE_SPLIT 4        /* handing half of the cube to another warp */
#define GPSAT_PHASE_REDUCE 5       /* learnt-database reduction */
#define GPSAT_PHASE_DECIDE 6       /* VSIDS / reference decision */
#define GPSAT_PHASE_BACKTRACK 7    /* cancel_until after a conflict or restart */
typedef struct gpsat_phase_stats {
    int64_t ns[GPSAT_N_PHASES];
    int64_t count[GPSAT_N_PHASES];
    int64_t backtracked_levels;   /* sum over backjumps of the levels undone (≙ total_backtracked_levels) */
    int64_t jobs;                 /* jobs run (cubes + split-off cubes) */
    int64_t job_ns;               /* warp time inside jobs */
    int64_t idle_ns;              /* warp time waiting for a job */
} gpsat_phase_stats;
int gpsat_get_phase_stats(gpsat_t *h, gpsat_phase_stats *out);

/* CUDA-event time (ms) of the kernels of the last gpsat_solve / gpsat_propagate_all on this handle */
double gpsat_last_kernel_ms(gpsat_t *h);

/* per-job records of the last gpsat_solve / gpsat_propagate_all (n_cubes entries) */
int gpsat_job_records(gpsat_t *h, gpsat_job_record *records, int32_t cap);

/* --- epoch API for one-process-per-GPU runs (cube-and-conquer over several GPUs; no reference equivalent, the
 * reference is single-GPU: SURVEY.md §8e).  gpsat_solve == begin + step(until done) + end. ------------------------- */
int gpsat_solve_begin(gpsat_t *h);
/* runs the persistent kernel until all local cubes are closed, a cube is SAT, or ~budget_ms elapsed (0 = no limit).
 * *done = 1 when nothing is left to do on this GPU. *verdict as gpsat_solve (UNDEF while cubes remain). */
int gpsat_solve_step(gpsat_t *h, double budget_ms, int32_t *done, int32_t *verdict);
int gpsat_solve_end(gpsat_t *h, int32_t *verdict, uint8_t *model, gpsat_stats *stats);
/* raises the early-termination flag (another GPU found a model): running cubes abort, no new cube starts */
int gpsat_request_stop(gpsat_t *h);
/* learnt-clause pool exchange through HOST buffers: records are [len, lit0, ..., lit(len-1)] packed back to back.
 * export returns clauses appended to this GPU's pool since the previous export / pack; import appends foreign clauses
 * (clauses of 16 or more literals do not fit a pool slot and are dropped). */
int gpsat_pool_export(gpsat_t *h, int32_t *buf, int64_t cap_words, int64_t *n_words);
int gpsat_pool_import(gpsat_t *h, const int32_t *buf, int64_t n_words);
/* Device-side exchange for ONE collective per epoch: gpsat_exchange_pack writes this GPU's exchange
 * block — header [magic, verdict, done, payload words, clauses, rank, jobs done, 0] followed by the pool slots published
 * since the previous pack (16-word slots [len, lit0 ...]) — into DEVICE memory owned by the caller (e.g. a torch
 * tensor); the caller all-gathers the blocks of all ranks (ncclAllGather / torch.distributed.all_gather_into_tensor)
 * and hands the gathered DEVICE buffer [n_ranks][block_words] to gpsat_exchange_unpack, which appends the other ranks'
 * clauses to this GPU's foreign pool (imported by every job that starts afterwards) and reduces the headers:
 * sat_rank (lowest rank that holds a model, -1 if none), all_done (every rank has closed all its cubes), any_undef (a finished rank ended
 * UNDEF).  Both calls run on the handle's stream and return after it is idle. */
int64_t gpsat_exchange_block_words(int32_t max_clauses);
int gpsat_exchange_pack(gpsat_t *h, void *dev_block, int64_t block_words, int32_t rank, int32_t done, int32_t verdict);
int gpsat_exchange_unpack(gpsat_t *h, const void *dev_blocks, int32_t n_ranks, int32_t my_rank, int64_t block_words,
                          int32_t *sat_rank, int32_t *all_done, int32_t *any_undef, int64_t *imported_clauses,
                          int64_t *jobs_done_total);
/* diagnostics: [next_job, stop_flag, sat_job, -, queue tail, head, outstanding jobs, idle warps, splits in flight, -,-,-,
 * launch stamp (low bits), busy us, blocks, warps per block]; may be called from another host thread while a step runs */
int gpsat_debug_ctrl(gpsat_t *h, int32_t *out16);
/* raw device pointers so a caller can run NCCL collectives on the pool / flag without staging through the host */
int gpsat_device_ptrs(gpsat_t *h, void **pool_words, void **pool_cursor, void **stop_flag, void **stream);

/* --- mesh: the GPUs of one box as ONE work pool over NVLink peer memory (no reference equivalent). ---
 * Every rank holds the formula AND the complete cube list, and owns a queue region (control block, ring of split-off cubes
 * with their hand-off blocks, foreign learnt-clause pool) that the other ranks map.  Inside ONE persistent launch per GPU:
 * root cubes are handed out by ONE cursor (rank 0's, advanced by every GPU with atomics over NVLink), idle warps take
 * split-off cubes from the rings of other GPUs, busy cubes split for the demand other GPUs advertise, short learnt clauses
 * are stored straight into the peers' foreign pools, the early-termination flag and global termination travel the same way
 * — no kernel boundary, no host collective on the data path.  Per-cube records / outcome flags are per-rank contributions
 * over ALL cubes: gpsat_mesh_results_pack copies them into a caller-owned DEVICE block
 * [flags n | open descendants n | records n x 20 words] (n = number of cubes); the caller reduces the blocks of all ranks
 * (MAX over the first n int32 words, SUM over the rest: int32 for the next n words, int64 for the records —
 * ncclAllReduce / torch.distributed.all_reduce) and gpsat_mesh_results_unpack turns the reduced block into the global
 * verdict, records (gpsat_job_records) and statistics.
 * One process per GPU: gpsat_mesh_export + all-gather of the 64-byte handles + gpsat_mesh_attach_ipc (CUDA IPC).
 * One process, several GPUs: gpsat_mesh_attach_local (peer access), or simply the gpsat_multi_* host below.
 * Call order per solve: [all ranks] gpsat_solve_begin -> barrier -> gpsat_solve_step(budget) until done -> reduce results.
 * The barrier matters: a peer must not touch a cursor or ring that its owner has not reset yet. */
#define GPSAT_IPC_HANDLE_BYTES 64
#define GPSAT_MESH_MAX_GPUS 8
int gpsat_mesh_export(gpsat_t *h, void *ipc_handle /* GPSAT_IPC_HANDLE_BYTES */);
int gpsat_mesh_attach_ipc(gpsat_t *h, int32_t n_ranks, int32_t rank, const void *ipc_handles /* n_ranks x 64 bytes */);
int gpsat_mesh_attach_local(gpsat_t *const *handles, int32_t n_ranks);   /* rank r = handles[r] */
int gpsat_mesh_detach(gpsat_t *h);
int64_t gpsat_mesh_result_words(gpsat_t *h);
int gpsat_mesh_results_pack(gpsat_t *h, void *dev_block, int64_t words);
int gpsat_mesh_results_unpack(gpsat_t *h, const void *dev_block, int64_t words, int32_t *verdict, gpsat_stats *stats);
int gpsat_handle_device(gpsat_t *h);
/* diagnostics: the first n words of the queue control block as it was after the last launch (csrc/gpsat_device.h GPSAT_DQC_*) */
int gpsat_debug_words(gpsat_t *h, int32_t *out, int32_t n);

/* --- multi-GPU host in one process (≙ the host side of SATSolver/main.cu:197-310 for N GPUs; the reference drives one) ---
 * One host thread per GPU, formula and cube list replicated, the GPUs meshed as above (one root cursor); per-cube results are
 * reduced with ncclAllReduce (libnccl is loaded at run time; the reduction runs on the host when it cannot be loaded).
 * n_gpus = 0: every visible GPU (at most GPSAT_MESH_MAX_GPUS); devices = NULL: ordinals 0 .. n_gpus-1. */
typedef struct gpsat_multi gpsat_multi_t;
int gpsat_multi_create(gpsat_multi_t **m, int32_t n_gpus, const int32_t *devices, int32_t n_vars, int64_t n_clauses,
                       const int64_t *offsets, const int32_t *lits, const gpsat_opts *opts);
int gpsat_multi_n_gpus(gpsat_multi_t *m);
/* time-bounded solves (0 = until done): when the limit is reached the open cubes stay parked, the verdict is GPSAT_UNDEF
 * unless a model was found, and the records say how many cubes were closed */
int gpsat_multi_set_time_limit(gpsat_multi_t *m, double ms);
int gpsat_multi_set_cubes(gpsat_multi_t *m, int32_t n_cubes, const int64_t *cube_offsets, const int32_t *cube_lits);
/* stats: counters summed over GPUs; kernel_ms = slowest GPU; warp_busy_frac = mean; reduce_backend (optional): 1 NCCL, 0 host */
int gpsat_multi_solve(gpsat_multi_t *m, int32_t *verdict, uint8_t *model, gpsat_stats *stats, int32_t *reduce_backend);
int gpsat_multi_job_records(gpsat_multi_t *m, gpsat_job_record *records, int32_t cap);
void gpsat_multi_destroy(gpsat_multi_t *m);

#ifdef __cplusplus
}
#endif
#endif /* GPSAT_H */
