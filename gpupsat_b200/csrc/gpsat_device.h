// Plain-data structures shared by the host API (gpsat_api.cu), the kernels (kernels.cu) and the per-warp solver
// (cdcl_warp.inl).  No CUDA types here so the test-only lockstep emulator can include it too.
#pragma once
#include <stdint.h>
#include "../../include/gpsat.h"

#define GPSAT_NO_CONFLICT ((int)0x80000000)
#define GPSAT_REASON_NONE (-1)
// clause references: >= 0  original clause = slot of its first literal in cl2 (header at slot-1);
//                    <= -2 learnt clause at arena word r = -2 - cref
#define GPSAT_LEARNT_CREF(r) (-2 - (r))
#define GPSAT_LEARNT_OFF(cref) (-2 - (cref))

#define GPSAT_MODE_SOLVE 0
#define GPSAT_MODE_PROPAGATE 1

#define GPSAT_VAL_FALSE 0
#define GPSAT_VAL_TRUE 1
#define GPSAT_VAL_UNDEF 2
#define GPSAT_VAL_ABSENT 4   // variable does not occur in the formula: assignable by a cube, never decided

// job status beyond the public verdicts
#define GPSAT_JOB_NOT_RUN (-1)
#define GPSAT_JOB_ABORTED (-2)   // stopped by the early-termination flag
#define GPSAT_JOB_OOM (-3)       // learnt arena exhausted even after reduction
#define GPSAT_JOB_SUSPENDED (-4) // parked in the ring at the end of a budgeted step; a later launch resumes it

// dynamic cube queue (children of split cubes)
#define GPSAT_DQ_MAXK 64         // literals per queued cube
#define GPSAT_DQ_CAP 16384       // slots of the ring of queued cubes (power of two, > 2 x resident warps)
#define GPSAT_HAND_CLAUSE_WORDS 8192   // learnt-clause words a queued cube inherits from the job that queued it
// shared learnt-clause pools: fixed slots [len, lit0 .. ] (len written last); exchange block = header + slots
#define GPSAT_POOL_SLOT_WORDS 16
#define GPSAT_XCHG_HEADER_WORDS 8
#define GPSAT_XCHG_MAGIC 0x47505358
// Queue control block (dq_ctrl): word indices.  Every group sits on its own 128-byte line so that the ring tickets,
// the job accounting and the words other GPUs write do not share a sector.
#define GPSAT_DQC_TAIL 0          // push tickets (fetch-add)
#define GPSAT_DQC_HEAD 8          // pop tickets (fetch-add; a ticket is drawn only after a claim on AVAIL succeeded)
#define GPSAT_DQC_AVAIL 16        // children published and not yet claimed (the semaphore consumers claim from)
#define GPSAT_DQC_CURSOR 24       // next root cube (≙ JobsQueue::next_job_index, SATSolver/JobsQueue.cu:16); mesh: rank 0's is THE cursor of all GPUs
#define GPSAT_DQC_ROOTS_DONE 25   // this GPU has seen the cursor run past the last root cube (no more remote fetches)
#define GPSAT_DQC_CREATED 32      // jobs created on this GPU: the root cubes it owns + the children it queued (monotonic)
#define GPSAT_DQC_CLOSED 33       // jobs closed on this GPU, wherever they were created (monotonic)
#define GPSAT_DQC_IDLE 64         // warps of this GPU with nothing to do
#define GPSAT_DQC_INFLIGHT 65     // splits being prepared
#define GPSAT_DQC_DONE 66         // set once every rank's closed == created (mesh: by the communication warp)
#define GPSAT_DQC_STEALS 67       // children this GPU took from the rings of other GPUs
#define GPSAT_DQC_STOP 68         // early-termination flag: 0 run, 1 SAT found, 2 external stop (≙ managed *state, main.cu:185)
#define GPSAT_DQC_PUSHED 69       // pool slots already pushed to the peers (communication warp)
#define GPSAT_DQC_REMOTE_TRIES 70 // remote pop attempts (diagnostics)
#define GPSAT_DQC_HIST 72         // [16] children by floor(log2(conflicts + 1)) (diagnostics: how big are split-off cubes?)

#define GPSAT_DQC_PEER_QUEUE 96   // [GPSAT_MESH_MAX_RANKS] children queued on rank r, written by r's communication warp
#define GPSAT_DQC_PEER_IDLE 128   // [GPSAT_MESH_MAX_RANKS] unmet demand of rank r (idle warps - queued children), its share for us
#define GPSAT_DQC_PHASE 160       // [GPSAT_N_PHASES] int64 ns per phase, then [GPSAT_N_PHASES] int64 counts, then int64 backtracked levels (opts.phase_stats)
#define GPSAT_DQC_WORDS 200
#define GPSAT_MESH_MAX_RANKS 8

// per-root outcome flags, combined with atomicMax (higher wins)
#define GPSAT_FLAG_UNSAT 1
#define GPSAT_FLAG_ABORTED 2
#define GPSAT_FLAG_UNDEF 3
#define GPSAT_FLAG_OOM 4
#define GPSAT_FLAG_SAT 5

struct gpsat_formula_view {
    int32_t n_vars;
    int32_t n_clauses;
    int32_t n_lits;
    int32_t wbits_words;
    const int32_t *cstart;    // n_clauses+1 : clause index -> header slot in cl2 (literals follow)
    const void *cl2;          // int2[n_lits + n_clauses] : header (len, clause index) then (literal, occurrence slot)
    const int32_t *ostart;    // 2*n_vars+1
    const void *occ2;         // int2[n_lits] : (first literal slot, len) of the clause owning occurrence slot k
    const uint32_t *wbits0;   // initial watch bitmap over occurrence slots
    const int32_t *vsids0;    // 2*n_vars initial VSIDS counters
    const uint8_t *val0;      // n_vars initial values: UNDEF or ABSENT
};

struct gpsat_solve_params {
    int32_t mode;             // GPSAT_MODE_*
    int32_t decision;         // GPSAT_DECIDE_*
    int32_t bcp;              // GPSAT_BCP_*
    int32_t restart_first;
    float restart_factor;
    int32_t max_iterations;
    int32_t stop_on_sat;
    int32_t share_learnts;
    int32_t share_max_len;
    int32_t max_learnts_first;   // learnt clauses kept before the first reduction
    int32_t learnt_refs_cap;     // capacity of the per-warp learnt clause list
    int64_t max_conflicts;
    int64_t arena_words;         // per warp
    int64_t implied_stride;      // propagate mode: words reserved per cube in `implied`
    int32_t dynamic_split;       // 1: a long-running cube hands half of its search space to an idle warp at restarts
    int32_t split_force;         // test hook: split at every restart even when no warp is idle
    int32_t split_gap;           // conflicts a job runs between two rounds of splitting
    int32_t split_burst;         // children handed out per round while warps are idle
    int32_t share_import_max;    // non-unit clauses a job takes from each shared pool when it starts (newest first)
    int32_t split_gap_hot;       // gap while demand >= split_hot_demand (many idle warps: fill them fast)
    int32_t split_hot_demand;    // demand from which the hot gap applies (1/8 of this GPU's warps)
    int32_t split_at_start;      // 1: a cube may split before its first conflict while warps are idle
    int32_t mesh_flags;          // test hooks: 1 no stealing, 2 no clause push
    int32_t split_reserve;       // children kept queued ahead of demand
    int32_t phase_stats;         // 1: per-phase time / count accumulators (≙ RuntimeStatistics, Statistics/RuntimeStatistics.cuh:17-66)
    int32_t split_mode;          // 0 back to the cube + VSIDS-best, 1 guiding path (oldest open decision), 2 as 0 with sides swapped
    int32_t split_min;           // hardness (own conflicts + inherited) a job needs before its first split
    int32_t split_hard;          // hardness from which a job splits after every conflict and at its start (0x7fffffff = never)
};

// word offsets (int32 units) of the per-warp state arrays inside one warp's state block
struct gpsat_state_layout {
    int32_t val, seen, level, reason, trail, trail_lim, wbits, vs, lbuf, cube, lwbits, ph;
    int32_t total_words;
    int32_t lbuf_words;
    int32_t idx16;
};

struct gpsat_run_buffers {
    const int64_t *cube_offsets;   // n_cubes+1
    const int32_t *cube_lits;
    int32_t n_cubes;
    int32_t *next_job;             // atomic cursor (≙ JobsQueue::next_job_index)
    int32_t *stop_flag;            // 0 run, 1 SAT found, 2 external stop
    int32_t *sat_job;              // first SAT job index (-1)
    uint8_t *model;                // n_vars, written by the SAT job that wins sat_job
    gpsat_job_record *records;     // n_cubes
    int32_t *implied;              // propagate mode (may be null)
    int32_t *n_implied;            // propagate mode (may be null)
    int64_t *conflict_clause;      // propagate mode (may be null)
    int32_t *arena;                // n_warps * arena_words
    int32_t *gstate;               // n_warps * layout.total_words when state is not in shared memory
    int32_t *pool;                 // shared learnt pool words
    int32_t *pool_cursor;          // [0] slots reserved, [1] clauses published, [2] export mark (slots)
    int32_t pool_cap_words;
    const int32_t *xpool;          // clauses received from other GPUs (same slot format, may be null)
    const int32_t *xpool_cursor;   // [0] slots used
    uint8_t *facts;                // n_vars: 0 none, 1 false, 2 true — learnt unit clauses of all GPUs (may be null)
    int32_t state_in_smem;
    int32_t formula_in_smem;       // cl2 / occ2 / ostart staged once per block in front of the warps' state blocks
    int32_t formula_smem_words;    // size of that staging area (multiple of 4 words)
    // dynamic splitting: children of split cubes are queued here and popped by idle warps
    int32_t *dq_lits;              // dq_cap * GPSAT_DQ_MAXK
    int32_t *dq_meta;              // dq_cap * 4 : (root cube, length, slot sequence number, inherited hardness); sequence starts at the slot index
    int32_t *dq_ctrl;              // GPSAT_DQC_* words
    int32_t *dq_hand;              // dq_cap * hand_words : per queued cube [n][vs 2n][records] (non-null when dynamic_split)
    int32_t hand_words;
    int32_t dq_cap;                // ring slots (power of two)
    int32_t *root_pending;         // n_cubes: open jobs descending from each original cube
    int32_t *root_flag;            // n_cubes: GPSAT_FLAG_* (atomicMax)
    const unsigned long long *t0;     // globaltimer stamp taken right before the launch
    unsigned long long budget_ns;     // warps stop pulling new cubes once now > *t0 + budget_ns (0 = no limit)
    long long *busy_ns;               // sum over warps of the time spent inside jobs (may be null)
    int32_t *park;                    // n_warps * park_words: per-warp parking blocks of budgeted steps (may be null)
    int32_t park_words;
    // ---- mesh: the GPUs of one box as one work pool over NVLink peer memory (no reference equivalent).
    // Every rank's queue region (control block, ring, hand-off blocks, foreign pool, facts) has the same layout; rank
    // r's region starts at mesh_base[r] (own rank: the local pointers above point into it).  mesh_ranks <= 1: off.
    int32_t mesh_ranks, mesh_rank;
    char *mesh_base[GPSAT_MESH_MAX_RANKS];
    int64_t mesh_off_ctrl, mesh_off_meta, mesh_off_lits, mesh_off_hand, mesh_off_xcur, mesh_off_xpool, mesh_off_facts;
    int32_t *stage;                   // n_warps * hand_words: local copy of a hand-off block popped from another GPU
    int32_t xpool_cap_slots;
    int32_t mesh_n_vars;
};

// Layout of one rank's mesh region (bytes, every part 256-byte aligned): identical on every rank because it depends
// only on (n_vars, hand_words, ring capacity, pool size).
struct gpsat_mesh_layout {
    int64_t ctrl, meta, lits, hand, xcur, xpool, facts, total;
};
static inline void gpsat_make_mesh_layout(int32_t n_vars, int32_t hand_words, int32_t dq_cap, int64_t pool_words,
                                          gpsat_mesh_layout *m)
{
    int64_t at = 0;
#define GPSAT_MTAKE(field, bytes)                 \
    do {                                          \
        m->field = at;                            \
        at += (((int64_t)(bytes)) + 255) / 256 * 256; \
    } while (0)
    GPSAT_MTAKE(ctrl, GPSAT_DQC_WORDS * 4);
    GPSAT_MTAKE(meta, (int64_t)dq_cap * 4 * 4);
    GPSAT_MTAKE(lits, (int64_t)dq_cap * GPSAT_DQ_MAXK * 4);
    GPSAT_MTAKE(hand, (int64_t)dq_cap * hand_words * 4);
    GPSAT_MTAKE(xcur, 64);
    GPSAT_MTAKE(xpool, pool_words * 4);
    GPSAT_MTAKE(facts, n_vars > 0 ? n_vars : 1);
#undef GPSAT_MTAKE
    m->total = at;
}

// per-warp state block: word offsets of each array (every array starts on a 16-byte boundary)
// idx16: level / trail / trail_lim hold 16-bit elements (the kernel variant with the staged formula: 2n < 65 536)
static inline void gpsat_make_layout(int32_t n_vars, int64_t n_lits, int32_t phase_stats, int32_t idx16, gpsat_state_layout *ly)
{
    int32_t at = 0;
    const int32_t n = n_vars > 0 ? n_vars : 1;
#define GPSAT_TAKE(field, words)          \
    do {                                  \
        ly->field = at;                   \
        at += (((words) + 3) / 4) * 4;    \
    } while (0)
    GPSAT_TAKE(val, (n + 3) / 4);
    GPSAT_TAKE(seen, (n + 3) / 4);
    GPSAT_TAKE(level, idx16 ? (n + 1) / 2 : n);
    GPSAT_TAKE(reason, n);
    GPSAT_TAKE(trail, idx16 ? (n + 1) / 2 : n);
    GPSAT_TAKE(trail_lim, idx16 ? (n + 2) / 2 : n + 1);
    GPSAT_TAKE(wbits, (int32_t)((n_lits + 31) / 32));
    GPSAT_TAKE(vs, 2 * n);
    ly->lbuf_words = (n + 1) > 64 ? (n + 1) : 64;
    GPSAT_TAKE(lbuf, ly->lbuf_words);
    GPSAT_TAKE(cube, GPSAT_DQ_MAXK);
    GPSAT_TAKE(lwbits, (2 * n + 31) / 32);
    // int64 ns[8] | int32 count[8] | int64 backtracked levels — only when the run keeps phase statistics
    GPSAT_TAKE(ph, phase_stats ? 2 * GPSAT_N_PHASES + GPSAT_N_PHASES + 2 : 0);
#undef GPSAT_TAKE
    ly->total_words = at;
    ly->idx16 = idx16;
}

// Launch geometry of the CDCL kernel from the formula's size and the shared memory of an SM (pure arithmetic: the CPU
// tests pin it, a silent change here costs 10 % of the throughput).
//   * the formula index is staged in shared memory (one word per cl2 / occ2 pair, every field below 2^16) when that
//     leaves room for at least 8 warps of 16-bit-layout state (and for the requested number of warps, if one is given);
//   * as many warps as fit, at most w_auto_max (24: measured on C2 — 20 warps 34.0 ms, 24 30.3, 28 30.3) unless the
//     caller asks for a number (up to w_max);
//   * state in global memory (16 warps) when fewer than 4 warps of state fit, or the requested number does not.
struct gpsat_geometry {
    gpsat_state_layout ly;
    int32_t warps, state_in_smem, formula_in_smem, formula_smem_words;
    int64_t smem_bytes;
};
static inline void gpsat_plan_warps(int32_t n_vars, int64_t n_lits, int64_t n_clauses, int32_t phase_stats, int32_t solve_mode,
                                    int32_t w_request, int32_t w_max, int32_t w_auto_max, int64_t smem_max, gpsat_geometry *g)
{
    gpsat_state_layout ly32, ly16;
    gpsat_make_layout(n_vars, n_lits, phase_stats, 0, &ly32);
    gpsat_make_layout(n_vars, n_lits, phase_stats, 1, &ly16);
    const int64_t f_words = ((n_lits + n_clauses + 3) & ~(int64_t)3) + ((n_lits + 3) & ~(int64_t)3) +
                            ((2 * (int64_t)n_vars + 1 + 3) & ~(int64_t)3);
    const int packs = n_lits + n_clauses < 65536 && 2 * (int64_t)n_vars < 65536;
    int32_t w = w_request > w_max ? w_max : w_request;
    g->state_in_smem = 1;
    g->formula_in_smem = 0;
    g->formula_smem_words = 0;
    if (solve_mode && packs && f_words * 4 + 8 * (int64_t)ly16.total_words * 4 <= smem_max &&
        (w <= 0 || f_words * 4 + (int64_t)w * ly16.total_words * 4 <= smem_max)) {
        g->formula_in_smem = 1;
        g->formula_smem_words = (int32_t)f_words;
    }
    g->ly = g->formula_in_smem ? ly16 : ly32;
    const int64_t bytes_per_warp = (int64_t)g->ly.total_words * 4;
    const int64_t room = smem_max - (int64_t)g->formula_smem_words * 4;
    if (w <= 0) {
        const int64_t cap = w_auto_max < w_max ? w_auto_max : w_max;
        const int64_t fit = room / (bytes_per_warp > 0 ? bytes_per_warp : 1);
        if (fit >= 4) {
            w = (int32_t)(fit < cap ? fit : cap);
        } else {
            g->state_in_smem = 0;
            w = 16;
        }
    } else if ((int64_t)w * bytes_per_warp > room) {
        g->state_in_smem = 0;
    }
    if (!g->state_in_smem) {
        g->formula_in_smem = 0;
        g->formula_smem_words = 0;
        g->ly = ly32;
    }
    g->warps = w;
    g->smem_bytes = g->state_in_smem ? (int64_t)w * g->ly.total_words * 4 + (int64_t)g->formula_smem_words * 4 : 0;
}

static inline int32_t gpsat_park_words(int32_t n_vars) { return ((16 + GPSAT_DQ_MAXK + 3 * (n_vars > 0 ? n_vars : 1)) + 3) / 4 * 4; }

// outcome of an original cube from the flags of all jobs that descend from it
static inline int32_t gpsat_root_status(int32_t flag, int32_t pending)
{
    if (flag == GPSAT_FLAG_SAT) return GPSAT_SAT;
    if (flag == GPSAT_FLAG_OOM) return GPSAT_JOB_OOM;
    if (flag == GPSAT_FLAG_UNDEF) return GPSAT_UNDEF;
    if (flag == GPSAT_FLAG_ABORTED) return GPSAT_JOB_ABORTED;
    if (flag == GPSAT_FLAG_UNSAT && pending == 0) return GPSAT_UNSAT;
    return GPSAT_JOB_NOT_RUN;   // not started, or descendants still open
}
