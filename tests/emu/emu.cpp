// TEST-ONLY lockstep emulator of the kernel source (gpupsat_b200/csrc/cdcl_warp.inl compiled with
// GPSAT_WARP_EMU).  It lets the CPU test-suite step the exact warp program the GPU runs and compare it with
// the oracle when no GPU is present.  Never linked into libgpsat.so; not a fallback of any kind.
#define GPSAT_WARP_EMU 1
#include <vector>
#include <cstring>
#include "../../gpupsat_b200/csrc/cdcl_warp.inl"
#include "../../gpupsat_b200/csrc/host_formula.h"

// kPacked: the variant the GPU runs beside a formula staged in shared memory — cl2 / occ2 pairs packed into one word,
// 16-bit level / trail / trail_lim (WarpSolverT<true>, kernels.cu: gpsat_cdcl_kernel<.., kSmemFormula = true, ..>)
static int g_emu_packed = 0;
extern "C" void gpsat_emu_set_packed(int on) { g_emu_packed = on; }

template <bool kPacked>
static int emu_run(int32_t n_vars, int64_t n_clauses, const int64_t *offsets, const int32_t *lits,
                   const gpsat_solve_params *params, int32_t n_cubes, const int64_t *cube_offsets,
                   const int32_t *cube_lits, gpsat_job_record *records, uint8_t *model, int32_t *sat_job,
                   int32_t *implied, int32_t *n_implied, int64_t *conflict_clause, int32_t *pool,
                   int32_t *pool_cursor, int32_t pool_cap_words, uint64_t budget_ticks, int32_t *n_launches)
{
    gpsat_host::DeviceFormula D;
    int rc = gpsat_host::build_device_formula(n_vars, n_clauses, offsets, lits, D);
    if (rc != 0) return rc;
    gpsat_formula_view F;
    F.n_vars = D.n_vars;
    F.n_clauses = (int32_t)D.n_clauses;
    F.n_lits = (int32_t)D.n_lits;
    F.wbits_words = (int32_t)D.wbits0.size();
    F.cstart = D.cstart.data();
    F.cl2 = D.cl2.data();
    F.ostart = D.ostart.data();
    F.occ2 = D.occ2.data();
    F.wbits0 = D.wbits0.data();
    F.vsids0 = D.vsids0.data();
    F.val0 = D.val0.data();
    std::vector<uint32_t> cl2p, occ2p;
    if (kPacked) {   // what the kernel's staging loop does (kernels.cu)
        if (D.n_lits + D.n_clauses >= 65536 || 2 * (int64_t)D.n_vars >= 65536) return -100;
        cl2p.resize((size_t)(D.n_lits + D.n_clauses));
        occ2p.resize((size_t)D.n_lits);
        for (size_t i = 0; i < cl2p.size(); i++) cl2p[i] = (uint32_t)D.cl2[2 * i] | ((uint32_t)D.cl2[2 * i + 1] << 16);
        for (size_t i = 0; i < occ2p.size(); i++) occ2p[i] = (uint32_t)D.occ2[2 * i] | ((uint32_t)D.occ2[2 * i + 1] << 16);
        F.cl2 = cl2p.data();
        F.occ2 = occ2p.data();
    }
    gpsat_solve_params P = *params;
    gpsat_state_layout Ly;
    gpsat_make_layout(n_vars, D.n_lits, P.phase_stats, kPacked ? 1 : 0, &Ly);
    std::vector<int32_t> state((size_t)Ly.total_words, 0);
    std::vector<int32_t> arena((size_t)P.arena_words, 0);
    *sat_job = -1;
    gpsat_run_buffers B;
    std::memset(&B, 0, sizeof(B));
    B.cube_offsets = cube_offsets;
    B.cube_lits = cube_lits;
    B.n_cubes = n_cubes;
    B.sat_job = sat_job;
    B.model = model;
    B.records = records;
    B.implied = implied;
    B.n_implied = n_implied;
    B.conflict_clause = conflict_clause;
    B.arena = arena.data();
    B.pool = pool;
    B.pool_cursor = pool_cursor;
    B.pool_cap_words = pool_cap_words;
    std::vector<uint8_t> facts((size_t)(n_vars > 0 ? n_vars : 1), 0);
    B.facts = facts.data();
    std::memset(records, 0, sizeof(gpsat_job_record) * (size_t)n_cubes);
    std::vector<int32_t> dq_lits((size_t)GPSAT_DQ_CAP * GPSAT_DQ_MAXK, 0), dq_meta((size_t)GPSAT_DQ_CAP * 4, 0);
    for (int i = 0; i < GPSAT_DQ_CAP; i++) dq_meta[4 * (size_t)i + 2] = i;
    std::vector<int32_t> root_pending((size_t)n_cubes, 1), root_flag((size_t)n_cubes, 0);
    int32_t dq_ctrl[GPSAT_DQC_WORDS] = {0};
    dq_ctrl[GPSAT_DQC_CREATED] = n_cubes;
    B.stop_flag = dq_ctrl + GPSAT_DQC_STOP;
    B.next_job = dq_ctrl + GPSAT_DQC_CURSOR;
    B.dq_lits = dq_lits.data();
    B.dq_meta = dq_meta.data();
    B.dq_ctrl = dq_ctrl;
    B.dq_cap = GPSAT_DQ_CAP;
    B.root_pending = root_pending.data();
    B.hand_words = 1 + 2 * n_vars + GPSAT_HAND_CLAUSE_WORDS;
    std::vector<int32_t> dq_hand(P.dynamic_split ? (size_t)GPSAT_DQ_CAP * B.hand_words : 1, 0);
    B.dq_hand = P.dynamic_split ? dq_hand.data() : nullptr;
    B.root_flag = root_flag.data();
    WarpSolverT<kPacked> S;
    std::memset(&S, 0, sizeof(S));
    // budgeted steps: relaunch the warp program until nothing is outstanding (≙ gpsat_solve_step in a loop)
    std::vector<int32_t> park((size_t)gpsat_park_words(n_vars), 0);
    B.park = park.data();
    B.park_words = (int32_t)park.size();
    unsigned long long t0 = 0;
    B.t0 = &t0;
    B.budget_ns = budget_ticks;
    int launches = 0;
    do {
        t0 = gpsat_now_ns();
        gpsat_bind(S, F, P, Ly, state.data(), arena.data(), B.park, B);
        gpsat_warp_loop(S, P, B, nullptr);
        launches++;
    } while (budget_ticks && dq_ctrl[GPSAT_DQC_CREATED] != dq_ctrl[GPSAT_DQC_CLOSED] && !dq_ctrl[GPSAT_DQC_STOP] && launches < 100000);
    if (n_launches) *n_launches = launches;
    for (int j = 0; j < n_cubes; j++) records[j].status = gpsat_root_status(root_flag[j], root_pending[j]);
    return 0;
}

extern "C" int gpsat_emu_run(int32_t n_vars, int64_t n_clauses, const int64_t *offsets, const int32_t *lits,
                             const gpsat_solve_params *params, int32_t n_cubes, const int64_t *cube_offsets,
                             const int32_t *cube_lits, gpsat_job_record *records, uint8_t *model, int32_t *sat_job,
                             int32_t *implied, int32_t *n_implied, int64_t *conflict_clause, int32_t *pool,
                             int32_t *pool_cursor, int32_t pool_cap_words, uint64_t budget_ticks, int32_t *n_launches)
{
    if (g_emu_packed)
        return emu_run<true>(n_vars, n_clauses, offsets, lits, params, n_cubes, cube_offsets, cube_lits, records, model,
                             sat_job, implied, n_implied, conflict_clause, pool, pool_cursor, pool_cap_words, budget_ticks,
                             n_launches);
    return emu_run<false>(n_vars, n_clauses, offsets, lits, params, n_cubes, cube_offsets, cube_lits, records, model,
                          sat_job, implied, n_implied, conflict_clause, pool, pool_cursor, pool_cap_words, budget_ticks,
                          n_launches);
}

// ---- launch geometry arithmetic of the CDCL kernel (gpsat_device.h: gpsat_plan_warps), exposed for the CPU tests -----
extern "C" void gpsat_emu_plan(int32_t n_vars, int64_t n_lits, int64_t n_clauses, int32_t phase_stats, int32_t solve_mode,
                               int32_t w_request, int32_t w_max, int32_t w_auto_max, int64_t smem_max, int64_t *out8)
{
    gpsat_geometry G;
    gpsat_plan_warps(n_vars, n_lits, n_clauses, phase_stats, solve_mode, w_request, w_max, w_auto_max, smem_max, &G);
    out8[0] = G.warps;
    out8[1] = G.state_in_smem;
    out8[2] = G.formula_in_smem;
    out8[3] = G.formula_smem_words;
    out8[4] = G.smem_bytes;
    out8[5] = G.ly.total_words;
    out8[6] = G.ly.idx16;
    out8[7] = G.ly.lbuf_words;
}

// ---- host-side index of the large-database sweep kernels (host_formula.cpp), exposed for the CPU tests --------------
extern "C" int gpsat_emu_bucket_index(int32_t n_vars, int64_t n_clauses, const int64_t *offsets, const int32_t *lits,
                                      uint32_t *bucket /* 16 * (2n + 2) */, int32_t *orange /* 2 * 2n */,
                                      int64_t *n_entries)
{
    gpsat_host::DeviceFormula D;
    int rc = gpsat_host::build_device_formula(n_vars, n_clauses, offsets, lits, D);
    if (rc != 0) return rc;
    gpsat_host::SweepIndex X;
    gpsat_host::build_sweep_index(D, true, X);
    if (X.bucket.empty()) return 1;   // database does not qualify for the bucket index
    std::copy(X.bucket.begin(), X.bucket.end(), bucket);
    std::copy(X.orange.begin(), X.orange.end(), orange);
    *n_entries = (int64_t)X.occ_clause.size();
    return 0;
}

extern "C" int gpsat_emu_order_cubes(int32_t n_vars, int64_t n_clauses, const int64_t *offsets, const int32_t *lits,
                                     int32_t n_cubes, const int64_t *cube_offsets, const int32_t *cube_lits,
                                     int32_t *sorted, int32_t *info)
{
    gpsat_host::DeviceFormula D;
    int rc = gpsat_host::build_device_formula(n_vars, n_clauses, offsets, lits, D);
    if (rc != 0) return rc;
    std::vector<int32_t> s, i;
    gpsat_host::order_cubes_for_sweep(D, n_cubes, cube_offsets, cube_lits, s, i);
    std::copy(s.begin(), s.end(), sorted);
    std::copy(i.begin(), i.end(), info);
    return 0;
}
