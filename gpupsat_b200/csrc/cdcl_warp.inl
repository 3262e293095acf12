// One cube = one warp.  Conflict-driven clause learning with two-watched-literal BCP, written in the lockstep
// vocabulary of warp_lockstep.h.  Included by kernels.cu (the product, sm_100a) and by tests/emu (test-only).
//
// What replaces what (reference file:line):
//   propagate()      WatchedClausesList::new_decision / process_clause / handle_implication
//                    (BCPStrategy/WatchedClausesList.cu:46-101,103-221,253-282) — per-literal watches over a
//                    bitmap of the static occurrence index instead of per-variable linked lists
//   analyze()        analyze_graph / get_conflicting_assignment (ConflictAnalysis/GraphAnalyzer.cu:31-138) — first
//                    UIP over trail + reason array instead of the dense back-edge matrix
//   learn()          LearntClausesManager::learn_clause (ClauseLearning/LearntClausesManager.cu:16-41)
//   cancel_until()   Backtracker::handle_backtrack + VariablesStateHandler::backtrack_to
//                    (SATSolver/Backtracker.cu:38-75, SATSolver/VariablesStateHandler.cu:367-399)
//   pick_branch()    DecisionMaker::decide / VSIDS::next_higher_literal
//                    (SATSolver/DecisionMaker.cu:45-87, DecisionStrategy/VSIDS.cu:97-124)
//   run_job()        SATSolver::solve + preprocess (SATSolver/SATSolver.cu:67-218,231-272) with the cube as
//                    pseudo-decision levels 1..k; restarts per Restarts/GeometricRestartsManager.cu:16-31
//
// Canonical order (what makes every job a pure function of (formula, cube, params); the oracle follows it):
//   * trail FIFO; for the falsified literal f = ~trail[qhead]: first its occurrence slots in ascending order, in
//     chunks of 32 examined against the assignment at chunk start, watch moves applied, then units / conflicts
//     committed in slot order; then the learnt clauses watching f one by one in watch-list order.
//   * learnt clause = first-UIP clause, literals in discovery order, highest other level swapped to position 1.
#pragma once
#include "gpsat_device.h"
#include "warp_lockstep.h"

// kPacked: cl2 / occ2 are the block's staged copy in shared memory, one 32-bit word per pair (x in the low half, y in
// the high half — every field of a formula small enough to be staged fits 16 bits); otherwise int2 in global memory
template <bool kPacked> struct gpsat_idx { typedef int type; };
template <> struct gpsat_idx<true> { typedef uint16_t type; };

template <bool kPacked> struct WarpSolverT {
    // element of level / trail / trail_lim: 16 bits beside a staged formula (2n < 65 536), so that more warps' state fits
    typedef typename gpsat_idx<kPacked>::type idx_t;
    int lane_id;          // this thread's lane, read once (see GPSAT_LANE_DECL)
    // ---- read-only formula index (global, L1/L2 resident, or staged in shared memory)
    int n_vars, n_clauses, n_lits, wbits_words;
    const gint2 *cl2;
    const int *ostart;
    const gint2 *occ2;
    GPSAT_DEV gint2 ld_pair(const gint2 *base, int i) const
    {
        if (kPacked) {
            const uint32_t w = reinterpret_cast<const uint32_t *>(base)[i];
            gint2 r;
            r.x = (int)(w & 0xFFFFu);
            r.y = (int)(w >> 16);
            return r;
        }
        return gpsat_ld2(base + i);
    }
    GPSAT_DEV gint2 ld_cl2(int i) const { return ld_pair(cl2, i); }
    GPSAT_DEV gint2 ld_occ2(int i) const { return ld_pair(occ2, i); }
    const uint32_t *wbits0;
    const int *vsids0;
    const uint8_t *val0;
    // ---- per-job state (shared memory when it fits, else this warp's global block)
    uint8_t *val;
    uint8_t *seen;
    idx_t *level;
    int *reason;
    idx_t *trail;
    idx_t *trail_lim;
    uint32_t *wbits;
    int *vs;
    int *lbuf;
    int lbuf_words;
    int *cube_buf;        // this job's cube when it may grow by splitting (GPSAT_DQ_MAXK literals)
    // ---- this warp's learnt arena (global): [lw_head: (ptr, size, cap) per literal | hist | refs | clauses ->   <- watch vectors]
    // the three words of a watch-vector head share a sector, so one round trip fetches them
    int *arena;
    int arena_words;
    int *lw_head, *hist, *refs;
    uint32_t *lwbits;     // per literal: does any learnt clause watch it?  (shared memory; saves the round trip when not)
    int refs_cap;
    int clause_base;
    // ---- warp-uniform scalars
    int trail_size, qhead, dlevel;
    int arena_top, watch_bot;
    int n_learnts, max_learnts;
    int conflicts_since_restart, restart_limit;
    int vs_clauses;
    int oom;
    int use_learnts;   // 0 in propagate-only runs: no learnt arena, no VSIDS state
    // params
    int decision_mode, restart_first, max_iterations, share_learnts, share_max_len, share_import_max;
    float restart_factor;
    long long max_conflicts;
    // dynamic splitting
    int dynamic_split, split_force, split_gap, split_burst, split_gap_hot, split_hot_demand, split_at_start, split_mode, split_min, split_hard, split_reserve, inherited, root;
    int *dq_lits, *dq_meta, *dq_ctrl, *root_pending, *dq_hand;
    int hand_words, dq_cap;
    int rel_slot, rel_seq;                 // queue slot this job was popped from (released once its content is consumed)
    int *rel_meta;                         // ... and the ring it belongs to
    // per-phase statistics (opts.phase_stats; ≙ RuntimeStatistics, Statistics/RuntimeStatistics.cuh:17-66): lane 0 keeps
    // ns / count per phase in this warp's state block and folds them into the control block when a job ends
    int phase_stats;
    long long *phs;   // [GPSAT_N_PHASES] ns, then [1] backtracked levels behind the counts
    int *phn;         // [GPSAT_N_PHASES] counts
#define GPSAT_PH(i, stmt)                                              \
    {                                                                  \
        unsigned long long t_ph = 0;                                   \
        if (phase_stats) t_ph = gpsat_now_ns();                        \
        stmt;                                                          \
        if (phase_stats) {                                             \
            LANE0                                                      \
            {                                                          \
                phs[i] += (long long)(gpsat_now_ns() - t_ph);          \
                phn[i] += 1;                                           \
            }                                                          \
        }                                                              \
    }
    int mesh_ranks;                        // > 1: other GPUs pop from this GPU's ring (system-scope fences / atomics)
    const unsigned long long *t0;          // budgeted steps: jobs park themselves once now > *t0 + budget_ns
    unsigned long long budget_ns;
    long long c_splits;
    // shared pool
    int *pool;            // this GPU's pool: fixed slots of GPSAT_POOL_SLOT_WORDS words, [len, lit0 ...], len written last
    int *pool_cursor;     // [0] slots reserved [1] clauses published [2] export mark (host / exchange kernels)
    int pool_cap_words;
    const int *xpool;     // clauses received from the other GPUs (same slot format, filled between launches)
    const int *xpool_cursor;
    uint8_t *facts;              // per variable: 0 none, 1 false, 2 true — unit clauses learnt by any job on any GPU
    int pool_mark, xpool_mark;   // slots of the two pools this job has already imported
    int *park;                   // this warp's parking block (budgeted steps), see park_job()
    // counters (uniform) + per-lane counters
    long long c_decisions, c_implications, c_conflicts, c_learnt_clauses, c_learnt_literals, c_restarts;
    unsigned long long c_hash;
    long long c_lwatchers, c_lwords;   // learnt-clause part of the counters (warp-uniform)
    LANEVAR(unsigned, l_watchers);
    LANEVAR(unsigned, l_words);

    // -----------------------------------------------------------------------------------------------------------
    GPSAT_DEV int lit_value(int x) const
    {
        int v = val[x >> 1];
        return v >= 2 ? 2 : (v ^ (x & 1) ^ 1);
    }
    GPSAT_DEV bool wbit(int k) const { return (wbits[k >> 5] >> (k & 31)) & 1u; }

    GPSAT_DEV void enqueue(int x, int why)
    {
        GPSAT_LANE_DECL
        SYNCWARP();   // every lane has finished reading val[] (lit_value of the same literal) before lane 0 overwrites it
        LANE0
        {
            int v = x >> 1;
            val[v] = (uint8_t)(x & 1);
            level[v] = (idx_t)dlevel;
            reason[v] = why;
            trail[trail_size] = (idx_t)x;
        }
        trail_size++;
        SYNCWARP();
    }

    GPSAT_DEV void new_level()
    {
        GPSAT_LANE_DECL
        LANE0 { trail_lim[dlevel] = (idx_t)trail_size; }
        dlevel++;
        SYNCWARP();
    }

    GPSAT_DEV void cancel_until(int lv)
    {
        GPSAT_LANE_DECL
        if (dlevel <= lv) return;
        const int start = trail_lim[lv];
        LANES
        {
            GPSAT_NOUNROLL
            for (int i = start + lane; i < trail_size; i += 32) {
                int v = trail[i] >> 1;
                val[v] = val0[v];
            }
        }
        trail_size = start;
        qhead = start;
        dlevel = lv;
        SYNCWARP();
    }

    // ---- learnt watch vectors --------------------------------------------------------------------------------
    GPSAT_DEV void lw_append(int x, int cref, int blocker)
    {
        GPSAT_LANE_DECL
        int n = lw_head[3 * x + 1], cap = lw_head[3 * x + 2], ptr = lw_head[3 * x];
        if (n == cap) {
            const int ncap = cap ? 2 * cap : 4;
            const int nptr = watch_bot - 2 * ncap;
            if (nptr < arena_top) {
                oom = 1;
                return;
            }
            watch_bot = nptr;
            LANES
            {
                GPSAT_NOUNROLL
                for (int i = lane; i < 2 * n; i += 32) arena[nptr + i] = arena[ptr + i];
            }
            SYNCWARP();
            LANE0
            {
                lw_head[3 * x] = nptr;
                lw_head[3 * x + 2] = ncap;
            }
            ptr = nptr;
        }
        LANE0
        {
            arena[ptr + 2 * n] = cref;
            arena[ptr + 2 * n + 1] = blocker;
            lw_head[3 * x + 1] = n + 1;
            lwbits[x >> 5] |= 1u << (x & 31);
        }
        SYNCWARP();
    }

    // ---- BCP -------------------------------------------------------------------------------------------------
    // returns GPSAT_NO_CONFLICT or the falsified clause's cref
    // Grouped propagation: a 3-SAT literal has ~6 occurrence slots, so examining one literal per step leaves 26 lanes
    // idle.  The warp takes a GROUP of consecutive trail literals whose occurrence lists fit 32 slots together (always at
    // least one; a literal with more than 32 occurrences is a group of its own, examined 32 slots at a time) and examines
    // all slots of the group at once against the assignment at group start.  A clause two of whose watched literals are
    // falsified by two literals of the same group would be examined twice against the same state (and both lanes would
    // move their watch onto the same replacement): the group is CUT before the literal that owns the second occurrence,
    // which then opens the next group.  Then the learnt clauses watching the group's literals, literal after literal.
    GPSAT_DEV int propagate()
    {
        GPSAT_LANE_DECL
        while (qhead < trail_size) {
            const int avail = (trail_size - qhead) < 32 ? (trail_size - qhead) : 32;
            LANEVAR(int, gf);     // lane i: falsified literal i of the candidate group, its occurrence range,
            LANEVAR(int, gos);    // and the head of its learnt watch vector (fetched now: one global round trip that
            LANEVAR(int, goe);    // overlaps the scan of the original clauses; nothing in part (1) changes it)
            LANEVAR(int, gwp);
            LANEVAR(int, gwn);
            LANES
            {
                LV(gf) = 0;
                LV(gos) = 0;
                LV(goe) = 0;
                LV(gwp) = 0;
                LV(gwn) = 0;
                if (lane < avail) {
                    const int f = trail[qhead + lane] ^ 1;
                    LV(gf) = f;
                    LV(gos) = gpsat_ld(ostart + f);
                    LV(goe) = gpsat_ld(ostart + f + 1);
                    if (use_learnts && ((lwbits[f >> 5] >> (f & 31)) & 1u)) {
                        LV(gwp) = lw_head[3 * f];
                        LV(gwn) = lw_head[3 * f + 1];
                    }
                }
            }
            // group = literals 0 .. g-1; lane L of the (single) chunk examines slot gk of literal gli
            LANEVAR(int, gli);
            LANEVAR(int, gk);
            LANES
            {
                LV(gli) = -1;
                LV(gk) = 0;
            }
            int g = 0, total = 0;
            const int os0 = SHFL(gos, 0);
            const int cnt0 = SHFL(goe, 0) - os0;
            if (cnt0 > 32) {
                g = 1;
                total = cnt0;
            } else {
                GPSAT_NOUNROLL
                for (int t = 0; t < avail; ++t) {
                    const int ost = SHFL(gos, t);
                    const int c = SHFL(goe, t) - ost;
                    if (t > 0 && total + c > 32) break;
                    LANES
                    {
                        if (lane >= total && lane < total + c) {
                            LV(gli) = t;
                            LV(gk) = ost + lane - total;
                        }
                    }
                    total += c;
                    g++;
                }
            }
            int cut = g;

            // (1) original clauses: the occurrence slots of the group whose watch bit is set
            GPSAT_NOUNROLL
            for (int base = 0; base < total; base += 32) {
                LANEVAR(int, act);    // 0 none, 1 move, 2 unit, 3 conflict
                LANEVAR(int, alit);   // unit literal, or new occurrence slot for a move
                LANEVAR(int, acl);    // clause cref (first literal slot); -(lane + 1) when the lane examined nothing
                LANEVAR(int, nrd);    // clause words read by this lane
                LANES
                {
                    if (cnt0 > 32) {   // one long list, 32 slots at a time
                        LV(gli) = base + lane < total ? 0 : -1;
                        LV(gk) = os0 + base + lane;
                    }
                    const int k = LV(gk);
                    LV(act) = 0;
                    LV(alit) = 0;
                    LV(acl) = -(lane + 1);
                    LV(nrd) = -1;
                    if (LV(gli) >= 0 && wbit(k)) {
                        const gint2 e = ld_occ2(k);
                        const int s = e.x, len = e.y;
                        int other = -1, other_val = 2, repl = -1, nread = 0;
                        GPSAT_HOTLOOP
                        for (int i = 0; i < len; ++i) {
                            const gint2 q = ld_cl2(s + i);
                            nread++;
                            if (q.y == k) continue;
                            const int v = lit_value(q.x);
                            if (wbit(q.y)) {
                                other = q.x;
                                other_val = v;
                                if (v == 1) break;
                            } else if (v != 0 && repl < 0) {
                                repl = q.y;
                            }
                            if (other >= 0 && repl >= 0) break;
                        }
                        LV(nrd) = nread;
                        LV(acl) = s;
                        if (other_val == 1) {
                            LV(act) = 0;
                        } else if (repl >= 0) {
                            LV(act) = 1;
                            LV(alit) = repl;
                        } else if (other_val == 2) {
                            LV(act) = 2;
                            LV(alit) = other;
                        } else {
                            LV(act) = 3;
                        }
                    }
                }
                SYNCWARP();
                if (g > 1) {   // a clause examined twice: cut the group before the literal of its second occurrence
                    LANEVAR(unsigned, same);
                    LANES { LV(same) = MATCH_ANY(acl); }
                    const unsigned second = BALLOT(LV(nrd) >= 0 && (LV(same) & GPSAT_LANEMASK_LT) != 0u);
                    if (second) cut = SHFL(gli, gpsat_ffs(second) - 1);   // lanes are in literal order: the lowest one has the smallest literal
                }
                LANES
                {
                    if (LV(gli) >= cut) {
                        LV(act) = 0;
                        LV(nrd) = -1;
                    }
                    if (LV(nrd) >= 0) {
                        LV(l_watchers) += 1u;
                        LV(l_words) += (unsigned)LV(nrd);
                    }
                    if (LV(act) == 1) {
                        const int k = LV(gk), r = LV(alit);
                        gpsat_atomic_and(wbits + (k >> 5), ~(1u << (k & 31)));
                        gpsat_atomic_or(wbits + (r >> 5), 1u << (r & 31));
                    }
                }
                SYNCWARP();
                unsigned todo = BALLOT(LV(act) >= 2);
                while (todo) {
                    const int src = gpsat_ffs(todo) - 1;
                    todo &= todo - 1;
                    const int a = SHFL(act, src);
                    const int s = SHFL(acl, src);
                    const int u = SHFL(alit, src);
                    if (a == 3) return s;
                    const int v = lit_value(u);
                    if (v == 2) {
                        enqueue(u, s);
                        c_implications++;
                    } else if (v == 0) {
                        return s;
                    }
                }
            }
            qhead += cut;

            // (2) learnt clauses watching the literals of the group, literal after literal
            GPSAT_NOUNROLL
            for (int t = 0; t < cut; ++t) {
            const int f = SHFL(gf, t);
            const int wp = SHFL(gwp, t);
            const int wn = SHFL(gwn, t);
            // MiniSat order, the warp cooperating on one clause at a time
            if (wn == 0) continue;
            c_lwatchers += wn;
            int confl = GPSAT_NO_CONFLICT;
            int j = 0;
            GPSAT_NOUNROLL
            for (int base = 0; base < wn; base += 32) {
                LANEVAR(int, ecr);
                LANEVAR(int, ebl);
                LANEVAR(int, keep);   // 0 dropped, 1 kept, 2 to examine
                LANEVAR(int, hlen);   // header of the lane's clause (length, first two literals), fetched by all lanes at once:
                LANEVAR(int, hl0);    // the clauses are then examined one after the other, and a clause is only ever
                LANEVAR(int, hl1);    // modified while it is being examined, so these stay valid
                LANES
                {
                    const int i = base + lane;
                    LV(keep) = 0;
                    LV(ecr) = 0;
                    LV(ebl) = 0;
                    LV(hlen) = 0;
                    LV(hl0) = 0;
                    LV(hl1) = 0;
                    if (i < wn) {
                        LV(ecr) = arena[wp + 2 * i];
                        LV(ebl) = arena[wp + 2 * i + 1];
                        LV(keep) = (confl != GPSAT_NO_CONFLICT || lit_value(LV(ebl)) == 1) ? 1 : 2;
                        if (LV(keep) == 2) {
                            const int *hc = arena + GPSAT_LEARNT_OFF(LV(ecr));
                            LV(hlen) = hc[0];
                            LV(hl0) = hc[1];
                            LV(hl1) = hc[2];
                        }
                    }
                }
                SYNCWARP();
                unsigned todo = BALLOT(LV(keep) == 2);
                while (todo) {
                    const int src = gpsat_ffs(todo) - 1;
                    todo &= todo - 1;
                    const int cref = SHFL(ecr, src);
                    const int bl = SHFL(ebl, src);
                    if (confl != GPSAT_NO_CONFLICT || lit_value(bl) == 1) {
                        SETLANE(keep, src, 1);
                        continue;
                    }
                    int *cl = arena + GPSAT_LEARNT_OFF(cref) + 1;
                    const int len = SHFL(hlen, src);
                    int l0 = SHFL(hl0, src), l1 = SHFL(hl1, src);
                    if (l0 == f) {
                        LANE0
                        {
                            cl[0] = l1;
                            cl[1] = l0;
                        }
                        l0 = l1;
                        l1 = f;
                        SYNCWARP();
                    }
                    const int first = l0;
                    c_lwords += 2;
                    if (first != bl && lit_value(first) == 1) {
                        SETLANE(keep, src, 1);
                        SETLANE(ebl, src, first);
                        continue;
                    }
                    int found = -1;
                    GPSAT_NOUNROLL
                    for (int b2 = 2; b2 < len && found < 0; b2 += 32) {
                        const unsigned mm = BALLOT((b2 + lane < len) && lit_value(cl[b2 + lane]) != 0);
                        c_lwords += (len - b2) < 32 ? (len - b2) : 32;
                        if (mm) found = b2 + gpsat_ffs(mm) - 1;
                    }
                    if (found >= 0) {
                        const int nl = cl[found];
                        LANE0
                        {
                            cl[1] = nl;
                            cl[found] = f;
                        }
                        SYNCWARP();
                        lw_append(nl, cref, first);
                        SETLANE(keep, src, 0);
                        if (oom) {   // could not move the watch: keep the structure consistent and bail out
                            LANE0
                            {
                                cl[found] = nl;
                                cl[1] = f;
                            }
                            SYNCWARP();
                            SETLANE(keep, src, 1);
                            confl = cref;   // unwinds the loop; run_job() reports OOM
                        }
                        continue;
                    }
                    SETLANE(keep, src, 1);
                    SETLANE(ebl, src, first);
                    if (lit_value(first) == 0) {
                        confl = cref;
                    } else {
                        enqueue(first, cref);
                        c_implications++;
                    }
                }
                const unsigned km = BALLOT(LV(keep) == 1);
                LANES
                {
                    if (LV(keep) == 1) {
                        const int dst = j + gpsat_popc(km & GPSAT_LANEMASK_LT);
                        arena[wp + 2 * dst] = LV(ecr);
                        arena[wp + 2 * dst + 1] = LV(ebl);
                    }
                }
                j += gpsat_popc(km);
                SYNCWARP();
            }
            LANE0
            {
                lw_head[3 * (f) + 1] = j;
                if (j == 0) lwbits[f >> 5] &= ~(1u << (f & 31));
            }
            SYNCWARP();
            if (confl != GPSAT_NO_CONFLICT) return confl;
            }   // literals of the group
        }
        return GPSAT_NO_CONFLICT;
    }

    // ---- first-UIP analysis ----------------------------------------------------------------------------------
    // fills lbuf[0..n_out) (lbuf[0] = asserting literal, lbuf[1] = a literal of the backjump level), returns n_out
    GPSAT_DEV int analyze(int confl, int &bt_level)
    {
        GPSAT_LANE_DECL
        int pathC = 0;
        int p = -1;
        int n_out = 1;
        int index = trail_size - 1;
        do {
            const bool orig = confl >= 0;
            const int s = orig ? confl : GPSAT_LEARNT_OFF(confl) + 1;
            const int len = orig ? ld_cl2(s - 1).x : arena[s - 1];
            GPSAT_NOUNROLL
            for (int b = 0; b < len; b += 32) {
                LANEVAR(int, q);
                LANEVAR(int, kind);   // 1: current level (path), 2: lower level (goes to the clause)
                LANES
                {
                    LV(kind) = 0;
                    LV(q) = 0;
                    const int i = b + lane;
                    if (i < len) {
                        const int x = orig ? ld_cl2(s + i).x : arena[s + i];
                        const int v = x >> 1;
                        LV(q) = x;
                        if (x != p && !seen[v] && level[v] > 0) {
                            seen[v] = 1;
                            LV(kind) = (level[v] >= dlevel) ? 1 : 2;
                        }
                    }
                }
                SYNCWARP();
                const unsigned m1 = BALLOT(LV(kind) == 1), m2 = BALLOT(LV(kind) == 2);
                pathC += gpsat_popc(m1);
                LANES
                {
                    if (LV(kind) == 2) lbuf[n_out + gpsat_popc(m2 & GPSAT_LANEMASK_LT)] = LV(q);
                }
                n_out += gpsat_popc(m2);
                SYNCWARP();
            }
            // most recent seen literal on the trail
            while (true) {
                const unsigned m = BALLOT((index - lane >= 0) && seen[trail[index - lane] >> 1]);
                if (m) {
                    index -= gpsat_ffs(m) - 1;
                    break;
                }
                index -= 32;
            }
            p = trail[index];
            index--;
            confl = reason[p >> 1];
            SYNCWARP();   // every lane has read seen[] (the ballot above) before lane 0 clears the entry
            LANE0 { seen[p >> 1] = 0; }
            SYNCWARP();
            pathC--;
        } while (pathC > 0);
        LANE0 { lbuf[0] = p ^ 1; }
        SYNCWARP();

        // backjump level = highest level among lbuf[1..); first such literal goes to position 1
        bt_level = 0;
        if (n_out > 1) {
            LANEVAR(long long, best);
            LANES
            {
                LV(best) = -1;
                GPSAT_NOUNROLL
                for (int i = 1 + lane; i < n_out; i += 32) {
                    const long long key = ((long long)level[lbuf[i] >> 1] << 32) | (long long)(0x7fffffff - i);
                    if (key > LV(best)) LV(best) = key;
                }
            }
            const long long top = LANE_MAX_I64(best);
            const int pos = 0x7fffffff - (int)(top & 0xffffffffll);
            bt_level = (int)(top >> 32);
            if (pos != 1) {
                const int a = lbuf[1], b = lbuf[pos];
                SYNCWARP();
                LANE0
                {
                    lbuf[1] = b;
                    lbuf[pos] = a;
                }
                SYNCWARP();
            }
        }
        LANES
        {
            GPSAT_NOUNROLL
            for (int i = 1 + lane; i < n_out; i += 32) seen[lbuf[i] >> 1] = 0;
        }
        SYNCWARP();
        return n_out;
    }

    // order-sensitive checksum of a learnt clause, folded into c_hash
    GPSAT_DEV void hash_learnt(int n_out)
    {
        GPSAT_LANE_DECL
        LANEVAR(unsigned long long, part);
        LANES
        {
            LV(part) = 0;
            GPSAT_NOUNROLL
            for (int i = lane; i < n_out; i += 32)
                LV(part) += (unsigned long long)(lbuf[i] + 1) * (unsigned long long)(i + 1) * 0x9E3779B97F4A7C15ull;
        }
        const unsigned long long hsum = LANE_SUM_U64(part);
        c_hash = c_hash * 0x100000001B3ull + hsum + (unsigned long long)n_out;
    }

    // store lbuf[0..n_out) as a learnt clause, watch positions 0 and 1; returns its cref (or NO_CONFLICT on OOM)
    GPSAT_DEV int learn(int n_out)
    {
        GPSAT_LANE_DECL
        if (n_learnts >= refs_cap || arena_top + n_out + 1 > watch_bot) {
            oom = 1;
            return GPSAT_NO_CONFLICT;
        }
        const int r = arena_top;
        LANES
        {
            GPSAT_NOUNROLL
            for (int i = lane; i < n_out; i += 32) arena[r + 1 + i] = lbuf[i];
        }
        LANE0
        {
            arena[r] = n_out;
            refs[n_learnts] = r;
        }
        arena_top += n_out + 1;
        n_learnts++;
        SYNCWARP();
        const int cref = GPSAT_LEARNT_CREF(r);
        lw_append(lbuf[0], cref, lbuf[1]);
        lw_append(lbuf[1], cref, lbuf[0]);
        return oom ? GPSAT_NO_CONFLICT : cref;
    }

    // VSIDS::handle_clause on the learnt clause: +1 per literal, halve everything every 50 clauses
    GPSAT_DEV void vsids_learnt(int n_out)
    {
        GPSAT_LANE_DECL
        LANES
        {
            GPSAT_NOUNROLL
            for (int i = lane; i < n_out; i += 32) vs[lbuf[i]] += 1;
        }
        SYNCWARP();
        vs_clauses++;
        if (vs_clauses % 50 == 0) {
            LANES
            {
                GPSAT_NOUNROLL
                for (int x = lane; x < 2 * n_vars; x += 32) vs[x] /= 2;
            }
            SYNCWARP();
        }
    }

    // append a short learnt clause to the per-GPU pool: one atomic on the slot cursor, length written last
    GPSAT_DEV void pool_publish(int n_out)
    {
        GPSAT_LANE_DECL
        if (!share_learnts || n_out > share_max_len || n_out >= GPSAT_POOL_SLOT_WORDS || pool == nullptr) return;
        if (n_out == 1 && facts != nullptr) {
            LANE0 { ((volatile uint8_t *)facts)[lbuf[0] >> 1] = (uint8_t)(1 + (lbuf[0] & 1)); }
        }
        LANEVAR(int, off);
        LANES { LV(off) = 0; }
        LANE0 { LV(off) = gpsat_atomic_add(pool_cursor, 1); }
        const int slot = SHFL(off, 0);
        if (slot >= pool_cap_words / GPSAT_POOL_SLOT_WORDS) return;   // pool full: clause stays private
        const int at = slot * GPSAT_POOL_SLOT_WORDS;
        LANES
        {
            GPSAT_NOUNROLL
            for (int i = lane; i < n_out; i += 32) pool[at + 1 + i] = lbuf[i];
        }
        SYNCWARP();
        gpsat_threadfence();
        LANE0
        {
            ((volatile int *)pool)[at] = n_out;
            gpsat_atomic_add(pool_cursor + 1, 1);
        }
        SYNCWARP();
    }

    // ---- learnt database reduction ---------------------------------------------------------------------------
    GPSAT_DEV bool locked(int r) const
    {
        const int x0 = arena[r + 1];
        return lit_value(x0) == 1 && reason[x0 >> 1] == GPSAT_LEARNT_CREF(r);
    }

    // drop the longer half of the unlocked learnt clauses (len > 2), compact the arena, rebuild the watch vectors
    GPSAT_DEV void reduce_db()
    {
        GPSAT_LANE_DECL
        // 1. histogram of candidate lengths
        LANES
        {
            GPSAT_NOUNROLL
            for (int i = lane; i < 64; i += 32) hist[i] = 0;
        }
        SYNCWARP();
        LANES
        {
            GPSAT_NOUNROLL
            for (int i = lane; i < n_learnts; i += 32) {
                const int r = refs[i];
                const int len = arena[r];
                if (len > 2 && !locked(r)) gpsat_atomic_add(hist + (len < 63 ? len : 63), 1);
            }
        }
        SYNCWARP();
        int n_cand = 0;
        GPSAT_NOUNROLL
        for (int b = 0; b < 64; ++b) n_cand += hist[b];
        const int target = n_cand / 2;
        int thr = 64, partial = 0;
        {
            int acc = 0;
            GPSAT_NOUNROLL
            for (int b = 63; b >= 3 && acc < target; --b) {
                const int h = hist[b];
                if (acc + h >= target) {
                    thr = b;
                    partial = target - acc;
                    acc = target;
                } else {
                    acc += h;
                }
            }
        }
        // 2. mark (negative header) in age order, 3. compact, relocating reasons of locked clauses
        int kept = 0;
        int top = clause_base;
        GPSAT_NOUNROLL
        for (int base = 0; base < n_learnts; base += 32) {
            LANEVAR(int, rr);
            LANEVAR(int, ll);
            LANEVAR(int, cls);   // 0 keep, 1 remove, 2 remove if partial budget left
            LANES
            {
                const int i = base + lane;
                LV(rr) = 0;
                LV(ll) = 0;
                LV(cls) = 0;
                if (i < n_learnts) {
                    const int r = refs[i];
                    const int len = arena[r];
                    LV(rr) = r;
                    LV(ll) = len;
                    if (len > 2 && !locked(r)) {
                        const int b = len < 63 ? len : 63;
                        LV(cls) = b > thr ? 1 : (b == thr ? 2 : 0);
                    }
                }
            }
            SYNCWARP();
            const int cnt = (n_learnts - base) < 32 ? (n_learnts - base) : 32;
            GPSAT_NOUNROLL
            for (int t = 0; t < cnt; ++t) {
                const int r = SHFL(rr, t);
                const int len = SHFL(ll, t);
                int c = SHFL(cls, t);
                if (c == 2) {
                    if (partial > 0) {
                        partial--;
                        c = 1;
                    } else {
                        c = 0;
                    }
                }
                if (c == 1) continue;
                if (top != r) {
                    const int x0 = arena[r + 1];
                    const bool lk = lit_value(x0) == 1 && reason[x0 >> 1] == GPSAT_LEARNT_CREF(r);
                    SYNCWARP();
                    // move down (top < r, ranges may overlap: ascending order, 32 words at a time)
                    GPSAT_NOUNROLL
                    for (int w = 0; w < len + 1; w += 32) {
                        LANEVAR(int, tmp);
                        LANES
                        {
                            LV(tmp) = (w + lane < len + 1) ? arena[r + w + lane] : 0;
                        }
                        SYNCWARP();
                        LANES
                        {
                            if (w + lane < len + 1) arena[top + w + lane] = LV(tmp);
                        }
                        SYNCWARP();
                    }
                    if (lk) {
                        LANE0 { reason[x0 >> 1] = GPSAT_LEARNT_CREF(top); }
                    }
                }
                LANE0 { refs[kept] = top; }
                SYNCWARP();
                kept++;
                top += len + 1;
            }
        }
        n_learnts = kept;
        arena_top = top;
        // 4. rebuild watch vectors: count, carve exact vectors from the top of the arena, fill in age order
        LANES
        {
            GPSAT_NOUNROLL
            for (int x = lane; x < 2 * n_vars; x += 32) lw_head[3 * (x) + 1] = 0;
            GPSAT_NOUNROLL
            for (int w = lane; w < (2 * n_vars + 31) / 32; w += 32) lwbits[w] = 0u;
        }
        SYNCWARP();
        LANES
        {
            GPSAT_NOUNROLL
            for (int i = lane; i < n_learnts; i += 32) {
                const int r = refs[i];
                gpsat_atomic_add(lw_head + 3 * arena[r + 1] + 1, 1);
                gpsat_atomic_add(lw_head + 3 * arena[r + 2] + 1, 1);
            }
        }
        SYNCWARP();
        watch_bot = arena_words;
        GPSAT_NOUNROLL
        for (int base = 0; base < 2 * n_vars; base += 32) {
            LANEVAR(int, cntv);
            LANEVAR(int, ptrv);
            LANES
            {
                const int x = base + lane;
                LV(cntv) = x < 2 * n_vars ? lw_head[3 * (x) + 1] : 0;
                LV(ptrv) = 0;
            }
            GPSAT_NOUNROLL
            for (int t = 0; t < 32; ++t) {
                const int c = SHFL(cntv, t);
                if (c > 0) {
                    watch_bot -= 2 * (2 * c + 4);
                    SETLANE(ptrv, t, watch_bot);
                }
            }
            LANES
            {
                const int x = base + lane;
                if (x < 2 * n_vars) {
                    lw_head[3 * (x)] = LV(ptrv);
                    lw_head[3 * (x) + 2] = LV(cntv) > 0 ? 2 * LV(cntv) + 4 : 0;
                    lw_head[3 * (x) + 1] = 0;
                }
            }
            SYNCWARP();
        }
        if (watch_bot < arena_top) {
            oom = 1;
            return;
        }
        GPSAT_NOUNROLL
        for (int i = 0; i < n_learnts; ++i) {
            const int r = refs[i];
            const int x0 = arena[r + 1], x1 = arena[r + 2];
            lw_append(x0, GPSAT_LEARNT_CREF(r), x1);
            lw_append(x1, GPSAT_LEARNT_CREF(r), x0);
        }
        max_learnts = max_learnts + max_learnts / 10 + 1;
        if (max_learnts > refs_cap - n_vars - 2) max_learnts = refs_cap - n_vars - 2;
    }

    // ---- decisions -------------------------------------------------------------------------------------------
    GPSAT_DEV int pick_branch()
    {
        GPSAT_LANE_DECL
        if (decision_mode == GPSAT_DECIDE_VSIDS) {
            LANEVAR(long long, best);
            LANES
            {
                LV(best) = -1;
                GPSAT_NOUNROLL
                for (int v = lane; v < n_vars; v += 32) {
                    if (val[v] != GPSAT_VAL_UNDEF) continue;
                    const long long kp = ((long long)vs[2 * v + 1] << 32) | (long long)(0x7fffffff - 2 * v);
                    const long long kn = ((long long)vs[2 * v] << 32) | (long long)(0x7fffffff - (2 * v + 1));
                    const long long k = kp > kn ? kp : kn;
                    if (k > LV(best)) LV(best) = k;
                }
            }
            const long long top = LANE_MAX_I64(best);
            if (top < 0) return -1;
            const int o = 0x7fffffff - (int)(top & 0xffffffffll);   // order index: 2v = positive, 2v+1 = negative
            return o ^ 1;
        }
        // shipped rule: topmost free variable, positive polarity
        GPSAT_NOUNROLL
        for (int base = n_vars - 1; base >= 0; base -= 32) {
            const unsigned m = BALLOT((base - lane >= 0) && val[base - lane] == GPSAT_VAL_UNDEF);
            if (m) return 2 * (base - (gpsat_ffs(m) - 1)) + 1;
        }
        return -1;
    }

    // ---- job -------------------------------------------------------------------------------------------------
    // keep_learnts: a parked job resumes in this warp — its learnt clauses, their watch vectors and the VSIDS counters
    // are still in place (every watch is legal on an empty trail); only the per-job assignment state starts over
    GPSAT_DEV void reset_job(bool keep_learnts = false)
    {
        GPSAT_LANE_DECL
        LANES
        {
            GPSAT_NOUNROLL
            for (int v = lane; v < n_vars; v += 32) {
                val[v] = val0[v];
                seen[v] = 0;
            }
            GPSAT_NOUNROLL
            for (int w = lane; w < wbits_words; w += 32) wbits[w] = wbits0[w];
            if (use_learnts && !keep_learnts) {
                GPSAT_NOUNROLL
                for (int x = lane; x < 2 * n_vars; x += 32) {
                    vs[x] = vsids0[x];
                    lw_head[3 * (x) + 1] = 0;
                    lw_head[3 * (x) + 2] = 0;
                    lw_head[3 * (x)] = 0;
                }
            }
            // learnt-watch bitmap: empty for a fresh job, rebuilt from the heads when a parked job resumes
            GPSAT_NOUNROLL
            for (int w = lane; w < (2 * n_vars + 31) / 32; w += 32) {
                uint32_t bits = 0u;
                if (use_learnts && keep_learnts) {
                    GPSAT_NOUNROLL
                    for (int b = 0; b < 32; ++b) {
                        const int x = 32 * w + b;
                        if (x < 2 * n_vars && lw_head[3 * x + 1] > 0) bits |= 1u << b;
                    }
                }
                lwbits[w] = bits;
            }
            LV(l_watchers) = 0;
            LV(l_words) = 0;
        }
        trail_size = qhead = dlevel = 0;
        if (!keep_learnts) {
            arena_top = clause_base;
            watch_bot = arena_words;
            n_learnts = 0;
            conflicts_since_restart = 0;
            restart_limit = restart_first;
            vs_clauses = n_clauses;
            pool_mark = xpool_mark = 0;
        }
        oom = 0;
        c_decisions = c_implications = c_conflicts = c_learnt_clauses = c_learnt_literals = c_restarts = 0;
        c_hash = 0;
        c_splits = 0;
        c_lwatchers = c_lwords = 0;
        SYNCWARP();
    }

    // Attach one foreign clause rec[0..len) (2 <= len <= 32) to this job's private database at the current (root)
    // level: satisfied clauses are skipped, a clause with one non-false literal becomes a fact, none refutes the job.
    // returns GPSAT_UNDEF to go on, GPSAT_UNSAT, or GPSAT_JOB_OOM
    GPSAT_DEV int attach_record(const int *rec, int len)
    {
        GPSAT_LANE_DECL
        LANEVAR(int, x);
        LANEVAR(int, v);
        LANES
        {
            LV(x) = lane < len ? gpsat_ld_cg(rec + lane) : 0;
            LV(v) = lane < len ? lit_value(LV(x)) : 0;
        }
        const unsigned in = len == 32 ? 0xffffffffu : ((1u << len) - 1u);
        const unsigned mt = BALLOT(LV(v) == 1) & in;
        const unsigned mnf = BALLOT(LV(v) != 0) & in;
        if (mt) return GPSAT_UNDEF;
        const int nf = gpsat_popc(mnf);
        if (nf == 0) return GPSAT_UNSAT;
        LANES
        {
            if (lane < len) {
                const bool isnf = (mnf >> lane) & 1u;
                const int dst = isnf ? gpsat_popc(mnf & GPSAT_LANEMASK_LT) : nf + gpsat_popc(~mnf & in & GPSAT_LANEMASK_LT);
                lbuf[dst] = LV(x);
            }
        }
        SYNCWARP();
        if (nf == 1) {
            enqueue(lbuf[0], GPSAT_REASON_NONE);
        } else if (learn(len) == GPSAT_NO_CONFLICT) {
            return GPSAT_JOB_OOM;
        }
        return GPSAT_UNDEF;
    }

    // room left for foreign clauses: they may take at most half of the arena and a quarter of the clause list
    GPSAT_DEV bool import_room() const
    {
        return (watch_bot - arena_top) >= (arena_words - clause_base) / 2 && n_learnts < refs_cap / 4;
    }

    // Unit clauses learnt anywhere (this GPU's jobs publish them, the exchange kernels add those of other GPUs) live
    // in one byte per variable; a job takes them all as level-0 facts when it starts or resumes.
    GPSAT_DEV int import_facts()
    {
        GPSAT_LANE_DECL
        if (facts == nullptr) return GPSAT_UNDEF;
        GPSAT_NOUNROLL
        for (int v0 = 0; v0 < n_vars; v0 += 32) {
            LANEVAR(int, f);
            LANES
            {
                const int v = v0 + lane;
                LV(f) = v < n_vars ? (int)((volatile const uint8_t *)facts)[v] : 0;
            }
            unsigned m = BALLOT(LV(f) != 0);
            while (m) {
                const int src = gpsat_ffs(m) - 1;
                m &= m - 1;
                const int lit = 2 * (v0 + src) + (SHFL(f, src) - 1);
                const int v = lit_value(lit);
                if (v == 0) return GPSAT_UNSAT;
                if (v == 2) enqueue(lit, GPSAT_REASON_NONE);
            }
        }
        return GPSAT_UNDEF;
    }

    // One importer for both record layouts (one attach site keeps the kernel's code small; nothing here propagates —
    // the caller's next propagate() does, at the root level):
    //   stride > 0   pool slots [from, used) of `stride` words, length written last (a slot whose length is still 0 is
    //                being written by another warp and is skipped): the NEWEST clauses of 2..stride-1 literals, at
    //                most share_import_max of them, looking at no more than 2048 slots, 32 slot headers per step;
    //                unit clauses are not taken here, they travel through `facts`
    //   stride == 0  hand-off block of a split cube: records [len, lits...] packed back to back in words [0, used):
    //                unit records (the parent's level-0 facts) become facts, every clause is attached
    GPSAT_DEV int import_stream(const int *base, int from, int used, int stride)
    {
        GPSAT_LANE_DECL
        if (stride && used - from > 2048) from = used - 2048;
        int taken = 0, at = 0, s0 = used, chunk_top = used;
        unsigned m = 0;
        LANEVAR(int, len_v);
        LANES { LV(len_v) = 0; }
        while (true) {
            const int *rec;
            int len;
            if (stride) {
                if (taken >= share_import_max) break;
                if (!m) {
                    if (s0 <= from) break;
                    const int lo = s0 - 32 > from ? s0 - 32 : from;
                    LANES
                    {
                        const int sl = s0 - 1 - lane;
                        LV(len_v) = sl >= lo ? gpsat_ld_cg(base + sl * stride) : 0;
                    }
                    m = BALLOT(LV(len_v) >= 2 && LV(len_v) < stride);
                    chunk_top = s0;
                    s0 -= 32;
                    if (!m) continue;
                }
                const int src = gpsat_ffs(m) - 1;
                m &= m - 1;
                rec = base + (chunk_top - 1 - src) * stride + 1;
                len = SHFL(len_v, src);
            } else {
                if (at >= used) break;
                len = gpsat_ld_cg(base + at);
                if (len <= 0 || at + 1 + len > used) break;
                rec = base + at + 1;
                at += len + 1;
                if (len == 1) {
                    const int u = gpsat_ld_cg(rec);
                    const int v = lit_value(u);
                    if (v == 0) return GPSAT_UNSAT;
                    if (v == 2) enqueue(u, GPSAT_REASON_NONE);
                    continue;
                }
                if (len > 32 || len > lbuf_words) continue;
            }
            if (!import_room()) break;
            const int st = attach_record(rec, len);
            if (st != GPSAT_UNDEF) return st;
            taken++;
        }
        // imported clauses do not count towards the job's own learnt-clause budget: otherwise a job that starts with a
        // few hundred foreign clauses reduces its database at once and keeps doing so (measured on uf300 over 2 GPUs:
        // twice the conflicts with sharing on)
        max_learnts += taken;
        if (max_learnts > refs_cap - n_vars - 2) max_learnts = refs_cap - n_vars - 2;
        return GPSAT_UNDEF;
    }

    // Import what the two shared pools gained since this job last looked (all of it for a fresh job).
    GPSAT_DEV int pool_import()
    {
        if (!share_learnts) return GPSAT_UNDEF;
        const int cap_slots = pool_cap_words / GPSAT_POOL_SLOT_WORDS;
        {
            const int st = import_facts();
            if (st != GPSAT_UNDEF) return st;
        }
        if (pool != nullptr) {
            int used = gpsat_ld_volatile(pool_cursor);
            if (used > cap_slots) used = cap_slots;
            const int from = pool_mark < used ? pool_mark : used;
            pool_mark = used;
            const int st = import_stream(pool, from, used, GPSAT_POOL_SLOT_WORDS);
            if (st != GPSAT_UNDEF) return st;
        }
        if (xpool != nullptr) {
            int used = gpsat_ld_volatile(xpool_cursor);
            if (used > cap_slots) used = cap_slots;
            const int from = xpool_mark < used ? xpool_mark : used;
            xpool_mark = used;
            const int st = import_stream(xpool, from, used, GPSAT_POOL_SLOT_WORDS);
            if (st != GPSAT_UNDEF) return st;
        }
        return GPSAT_UNDEF;
    }

    // Runs one cube.  mode SOLVE: full CDCL; mode PROPAGATE: stop once every cube literal is placed.
    // returns GPSAT_SAT / GPSAT_UNSAT / GPSAT_UNDEF / GPSAT_JOB_ABORTED / GPSAT_JOB_OOM; conflict_out = falsified
    // clause of the last conflict (cref) or NO_CONFLICT
    // What a child inherits from the cube it was split off: the VSIDS counters, the level-0 facts and the newest
    // learnt clauses that fit (all of them are implied by the formula alone, cube literals being decisions).
    // Block layout: [n_record_words][vs: 2*n_vars][records ...]
    GPSAT_DEV void write_handoff(int slot)
    {
        GPSAT_LANE_DECL
        if (dq_hand == nullptr) return;
        int *h = dq_hand + (long long)slot * hand_words;
        int *rec = h + 1 + 2 * n_vars;
        const int cap = hand_words - 1 - 2 * n_vars;
        LANES
        {
            GPSAT_NOUNROLL
            for (int x = lane; x < 2 * n_vars; x += 32) h[1 + x] = vs[x];
        }
        // level-0 facts as unit records
        const int n0 = dlevel > 0 ? trail_lim[0] : trail_size;
        int used = 0;
        const int n_units = (2 * n0 <= cap) ? n0 : cap / 2;
        LANES
        {
            GPSAT_NOUNROLL
            for (int i = lane; i < n_units; i += 32) {
                rec[2 * i] = 1;
                rec[2 * i + 1] = trail[i];
            }
        }
        used = 2 * n_units;
        // newest learnt clauses: the arena tail [refs[first] .. arena_top) that fits
        int first = n_learnts;
        GPSAT_NOUNROLL
        for (int base = 0; base < n_learnts && first == n_learnts; base += 32) {
            const unsigned m = BALLOT((base + lane < n_learnts) && (arena_top - refs[base + lane] <= cap - used));
            if (m) first = base + gpsat_ffs(m) - 1;
        }
        if (first < n_learnts) {
            const int from = refs[first];
            const int nw = arena_top - from;
            LANES
            {
                GPSAT_NOUNROLL
                for (int i = lane; i < nw; i += 32) rec[used + i] = arena[from + i];
            }
            used += nw;
        }
        LANE0 { h[0] = used; }
        SYNCWARP();
    }

    // ---- dynamic queue: a bounded multi-producer multi-consumer ring (per-slot sequence numbers, dq_meta[4*slot+2]).
    // dq_ctrl words: GPSAT_DQC_* (gpsat_device.h).
    // demand = idle warps - queued children - splits in flight (+ what the other GPUs of the mesh advertise as their
    // unmet demand): how many more children would find a taker right now.
    GPSAT_DEV int demand_hint() const
    {
        int d = gpsat_ld_volatile(dq_ctrl + GPSAT_DQC_IDLE) -
                (gpsat_ld_volatile(dq_ctrl + GPSAT_DQC_TAIL) - gpsat_ld_volatile(dq_ctrl + GPSAT_DQC_HEAD)) -
                gpsat_ld_volatile(dq_ctrl + GPSAT_DQC_INFLIGHT) + split_reserve;
        // this GPU's idle warps take a new child before any other GPU can: what the peers advertise counts only once
        // the local warps are served (d <= 0: -d children are queued beyond local demand)
        if (d <= 0) {
            GPSAT_NOUNROLL
            for (int r = 0; r < mesh_ranks; ++r) d += gpsat_ld_volatile(dq_ctrl + GPSAT_DQC_PEER_IDLE + r);
        }
        return d;
    }

    // reserves a slot for writing; returns it (ticket in `ticket`) or -1 when the ring is (nearly) full.
    // Push tickets are a fetch-add too; the margin covers every warp of the GPU drawing a ticket at the same moment.
    GPSAT_DEV int dq_acquire(int &ticket)
    {
        GPSAT_LANE_DECL
        LANEVAR(int, slot_v);
        LANEVAR(int, pos_v);
        LANES
        {
            LV(slot_v) = -1;
            LV(pos_v) = 0;
        }
        LANE0
        {
            if (gpsat_ld_volatile(dq_ctrl + GPSAT_DQC_TAIL) - gpsat_ld_volatile(dq_ctrl + GPSAT_DQC_HEAD) < dq_cap - 4096) {
                const int pos = gpsat_atomic_add(dq_ctrl + GPSAT_DQC_TAIL, 1);
                const int slot = pos & (dq_cap - 1);
                // freed long ago by the consumer of ticket pos - dq_cap (it releases the slot as soon as it has copied it)
                while (gpsat_ld_volatile(dq_meta + 4 * slot + 2) != pos) gpsat_nanosleep(200);
                LV(slot_v) = slot;
                LV(pos_v) = pos;
            }
        }
        ticket = SHFL(pos_v, 0);
        return SHFL(slot_v, 0);
    }

    // writes cube[0..k) (+ extra literal when >= 0) and the hand-off block into `slot` and publishes it
    GPSAT_DEV void dq_fill_and_publish(int slot, int ticket, int k, int extra)
    {
        GPSAT_LANE_DECL
        LANES
        {
            GPSAT_NOUNROLL
            for (int i = lane; i < k; i += 32) dq_lits[(long long)slot * GPSAT_DQ_MAXK + i] = cube_buf[i];
        }
        LANE0
        {
            if (extra >= 0) dq_lits[(long long)slot * GPSAT_DQ_MAXK + k] = extra;
            dq_meta[4 * slot] = root;
            dq_meta[4 * slot + 1] = extra >= 0 ? k + 1 : k;
            dq_meta[4 * slot + 3] = (int)((inherited + c_conflicts) / 2);   // the child is as hard as its parent was (a guess)
            gpsat_atomic_add(root_pending + root, 1);
            gpsat_atomic_add(dq_ctrl + GPSAT_DQC_CREATED, 1);
        }
        SYNCWARP();
        write_handoff(slot);
        if (mesh_ranks > 1) gpsat_threadfence_sys();   // the taker may be a warp of another GPU
        else gpsat_threadfence();
        LANE0
        {
            ((volatile int *)dq_meta)[4 * slot + 2] = ticket + 1;   // publish
            if (mesh_ranks > 1) gpsat_atomic_add_sys(dq_ctrl + GPSAT_DQC_AVAIL, 1);
            else gpsat_atomic_add(dq_ctrl + GPSAT_DQC_AVAIL, 1);
        }
        SYNCWARP();
    }

    // a job popped from the ring frees its slot as soon as it has copied the cube and imported the hand-off
    GPSAT_DEV void release_slot()
    {
        GPSAT_LANE_DECL
        if (rel_slot < 0) return;
        gpsat_threadfence();
        LANE0 { ((volatile int *)rel_meta)[4 * rel_slot + 2] = rel_seq; }
        SYNCWARP();
        rel_slot = -1;
    }

    // Hand half of the remaining search space to another warp: queue cube + ~p, keep cube + p.
    // Guiding path (split_mode 0, called at any point of the search with dlevel > k): p is the OLDEST open decision of
    // this job (the first literal of level k+1) — the untried side of that decision is the largest unexplored subtree,
    // the job simply adopts p as one more cube literal and keeps its trail, its learnt clauses and its place in the
    // search.  Otherwise (called with exactly the k cube literals decided and propagated): p = the VSIDS-best variable.
    // Returns the new k.
    GPSAT_DEV int try_split(int k)
    {
        GPSAT_LANE_DECL
        if (k + 1 >= GPSAT_DQ_MAXK) return k;
        LANEVAR(int, claim_v);
        LANES { LV(claim_v) = 0; }
        LANE0
        {
            gpsat_atomic_add(dq_ctrl + GPSAT_DQC_INFLIGHT, 1);   // splits in flight, this one included
            LV(claim_v) = (split_force || demand_hint() >= 0) ? 1 : 0;
        }
        int slot = -1, ticket = 0, p = -1;
        if (SHFL(claim_v, 0)) {
            p = dlevel > k ? trail[trail_lim[k]] : pick_branch();
            if (p >= 0 && split_mode == 2 && dlevel == k) p ^= 1;   // keep the side VSIDS would NOT try first
            if (p >= 0) slot = dq_acquire(ticket);
        }
        LANE0 { gpsat_atomic_add(dq_ctrl + GPSAT_DQC_INFLIGHT, -1); }   // from here on the child is counted by tail - head
        if (slot < 0) return k;                        // no demand, nothing to branch on, or the ring is full
        LANE0 { cube_buf[k] = p; }
        SYNCWARP();
        dq_fill_and_publish(slot, ticket, k, p ^ 1);
        c_splits++;
        return k + 1;
    }

    // Budgeted steps (gpsat_solve_step): when the step's budget is spent the job parks itself so that the kernel can
    // end and the host can run the epoch's exchange.  Everything expensive stays where it is — the learnt clauses and
    // their watch vectors in this warp's arena — and the parking block only records the cube, the level-0 facts, the
    // VSIDS counters and a few scalars; the same warp of the next launch resumes from there (a restart that loses
    // nothing).  Block layout: [valid, root, k, n_facts, n_learnts, arena_top, watch_bot, max_learnts, vs_clauses,
    // restart_limit, conflicts_since_restart, pool_mark, xpool_mark, -, -, -][cube 64][facts n][vs 2n]
    GPSAT_DEV void park_job(int k)
    {
        GPSAT_LANE_DECL
        const int n0 = dlevel > 0 ? trail_lim[0] : trail_size;
        LANES
        {
            GPSAT_NOUNROLL
            for (int i = lane; i < k; i += 32) park[16 + i] = cube_buf[i];
            GPSAT_NOUNROLL
            for (int i = lane; i < n0; i += 32) park[16 + GPSAT_DQ_MAXK + i] = trail[i];
            GPSAT_NOUNROLL
            for (int x = lane; x < 2 * n_vars; x += 32) park[16 + GPSAT_DQ_MAXK + n_vars + x] = vs[x];
        }
        LANE0
        {
            park[1] = root;
            park[2] = k;
            park[3] = n0;
            park[4] = n_learnts;
            park[5] = arena_top;
            park[6] = watch_bot;
            park[7] = max_learnts;
            park[8] = vs_clauses;
            park[9] = restart_limit;
            park[10] = conflicts_since_restart;
            park[11] = pool_mark;
            park[12] = xpool_mark;
            park[13] = inherited;
        }
        SYNCWARP();
        gpsat_threadfence();
        LANE0 { ((volatile int *)park)[0] = 1; }
        SYNCWARP();
    }

    GPSAT_DEV int run_job(const int *cube, int k, int mode, volatile const int *stop_flag, int &conflict_out,
                          const int *hand, bool resume)
    {
        GPSAT_LANE_DECL
        conflict_out = GPSAT_NO_CONFLICT;
        GPSAT_PH(0, reset_job(resume))
        if (resume) {   // parked by this warp at the end of the previous step (see park_job)
            n_learnts = park[4];
            arena_top = park[5];
            watch_bot = park[6];
            max_learnts = park[7];
            vs_clauses = park[8];
            restart_limit = park[9];
            conflicts_since_restart = park[10];
            pool_mark = park[11];
            xpool_mark = park[12];
            inherited = park[13];
            cube = park + 16;
            const int n0 = park[3];
            LANES
            {
                GPSAT_NOUNROLL
                for (int x = lane; x < 2 * n_vars; x += 32) vs[x] = park[16 + GPSAT_DQ_MAXK + n_vars + x];
            }
            SYNCWARP();
            GPSAT_NOUNROLL
            for (int i = 0; i < n0; ++i) {
                const int u = park[16 + GPSAT_DQ_MAXK + i];
                const int v = lit_value(u);
                if (v == 0) return GPSAT_UNSAT;
                if (v == 2) enqueue(u, GPSAT_REASON_NONE);
            }
        }
        // a cube that may be split or parked works on a private copy (GPSAT_DQ_MAXK words of the state block, and as
        // many in the parking block); a longer cube runs in place and can neither split nor park
        const bool queued_ok = mode == GPSAT_MODE_SOLVE && dynamic_split && k < GPSAT_DQ_MAXK;
        const bool may_split = queued_ok && k + 1 < GPSAT_DQ_MAXK;
        int want_split = 0, burst = 0;
        int at_start = (may_split && split_at_start && !resume) ? 1 : 0;
        int hard_start = (may_split && !resume && inherited >= split_hard) ? 1 : 0;
        long long last_split_at = 0;
        if (queued_ok) {   // the cube may grow (splits) or be parked (budgeted steps): work on a private copy
            LANES
            {
                GPSAT_NOUNROLL
                for (int i = lane; i < k; i += 32) cube_buf[i] = gpsat_ld_cg(cube + i);
            }
            SYNCWARP();
            cube = cube_buf;
        }
        if (resume) {
            LANE0 { ((volatile int *)park)[0] = 0; }
            SYNCWARP();
        }
        if (hand != nullptr) {   // popped from the ring: inherit the parent's counters, facts and newest clauses
            LANES
            {
                GPSAT_NOUNROLL
                for (int x = lane; x < 2 * n_vars; x += 32) vs[x] = gpsat_ld_cg(hand + 1 + x);
            }
            SYNCWARP();
            int st;
            GPSAT_PH(1, st = import_stream(hand + 1 + 2 * n_vars, 0, gpsat_ld_cg(hand), 0))
            release_slot();
            if (st != GPSAT_UNDEF) return st;
        }
        if (mode == GPSAT_MODE_SOLVE) {
            const int st = pool_import();
            if (st != GPSAT_UNDEF) return st;
        }
        while (true) {
            int confl;
            GPSAT_PH(2, confl = propagate())
            if (oom) return GPSAT_JOB_OOM;
            if (confl != GPSAT_NO_CONFLICT) {
                conflict_out = confl;
                c_conflicts++;
                conflicts_since_restart++;
                if (dlevel == 0 || mode == GPSAT_MODE_PROPAGATE) return GPSAT_UNSAT;
                int bt;
                int n_out;
                GPSAT_PH(3, n_out = analyze(confl, bt))
                const int dl_before = dlevel;
                hash_learnt(n_out);
                c_learnt_clauses++;
                c_learnt_literals += n_out;
                GPSAT_PH(GPSAT_PHASE_BACKTRACK, cancel_until(bt))
                if (phase_stats) {
                    LANE0 { phs[GPSAT_N_PHASES + GPSAT_N_PHASES / 2] += (long long)(dl_before - bt); }
                }
                if (n_out == 1) {
                    enqueue(lbuf[0], GPSAT_REASON_NONE);
                } else {
                    const int cref = learn(n_out);
                    if (cref == GPSAT_NO_CONFLICT) return GPSAT_JOB_OOM;
                    enqueue(lbuf[0], cref);
                }
                c_implications++;
                if (decision_mode == GPSAT_DECIDE_VSIDS) vsids_learnt(n_out);
                pool_publish(n_out);
                if (max_conflicts && c_conflicts >= max_conflicts) return GPSAT_UNDEF;
                if ((c_conflicts & 7) == 0) {   // every 8 conflicts (< 1 ms): stop flag and step budget
                    if (stop_flag && *stop_flag) return GPSAT_JOB_ABORTED;
                    if (budget_ns && queued_ok) {
                        LANEVAR(int, late_v);
                        LANES { LV(late_v) = 0; }
                        LANE0 { LV(late_v) = gpsat_now_ns() > *t0 + budget_ns ? 1 : 0; }
                        if (SHFL(late_v, 0) && park != nullptr) {
                            park_job(k);
                            return GPSAT_JOB_SUSPENDED;
                        }
                    }
                }
                if (may_split && !want_split && inherited + c_conflicts >= split_hard) {
                    if (demand_hint() > 0) {   // a very hard cube: every conflict is a chance to hand work out
                        want_split = 1;
                        burst = 0;
                    }
                } else if (may_split && !want_split && c_conflicts - last_split_at >= split_gap_hot &&
                    inherited + c_conflicts >= split_min) {
                    // a cube splits once it has proved hard (split_min), then every split_gap conflicts while warps are
                    // idle — every split_gap_hot conflicts while more than 1/8 of the GPU's warps are idle (start and tail)
                    const int d = demand_hint();
                    if (d > 0 && (c_conflicts - last_split_at >= split_gap || d >= split_hot_demand)) {
                        want_split = 1;
                        burst = 0;
                    }
                }
                continue;
            }
            if (mode == GPSAT_MODE_PROPAGATE && dlevel >= k) return GPSAT_UNDEF;
            if (restart_first > 0 && conflicts_since_restart >= restart_limit) {
                conflicts_since_restart = 0;
                restart_limit = (int)((float)restart_limit * restart_factor);
                c_restarts++;
                cancel_until(k < dlevel ? k : dlevel);
                if (split_force) want_split = 1;
            }
            if (at_start && dlevel >= k) {   // the cube is placed and nothing has been searched yet: idle warps get their share now
                at_start = 0;
                if (demand_hint() >= split_hot_demand) {
                    want_split = 1;
                    burst = 0;
                }
            } else if (hard_start && dlevel >= k) {   // split off a very hard cube: pass work on before searching
                hard_start = 0;
                if (demand_hint() > 0) {
                    want_split = 1;
                    burst = 0;
                }
            }
            if (want_split && dlevel >= k) {   // an idle warp is waiting: give it half of what is left
                want_split = 0;
                if (split_mode != 1 || dlevel == k) cancel_until(k);
                int k2;
                GPSAT_PH(4, k2 = try_split(k))
                if (k2 != k) {
                    // while warps are still idle keep peeling children off (cube+~p1, cube+p1+~p2, ...): the
                    // number of busy warps then grows by split_burst per gap instead of doubling
                    last_split_at = c_conflicts;
                    if (++burst < split_burst && demand_hint() > 0) want_split = 1;
                }
                k = k2;
            }
            if (use_learnts && (n_learnts >= max_learnts || (watch_bot - arena_top) < (arena_words - clause_base) / 4)) {
                GPSAT_PH(5, reduce_db())
                if (oom) return GPSAT_JOB_OOM;
            }
            int next = -1;
            while (dlevel < k) {
                const int x = cube[dlevel];
                const int v = lit_value(x);
                if (v == 1) {
                    new_level();
                } else if (v == 0) {
                    // the cube contradicts what propagation already fixed: the clause that implied ~x is falsified
                    // once the whole cube is assigned (the reference assigns it up front: SATSolver.cu:231-246)
                    const int why = reason[x >> 1];
                    if (why != GPSAT_REASON_NONE) conflict_out = why;
                    return GPSAT_UNSAT;
                } else {
                    next = x;
                    break;
                }
            }
            if (next < 0) {
                if (mode == GPSAT_MODE_PROPAGATE) return GPSAT_UNDEF;
                GPSAT_PH(GPSAT_PHASE_DECIDE, next = pick_branch())
                if (next < 0) return GPSAT_SAT;
                c_decisions++;
                if (max_iterations && c_decisions > max_iterations) return GPSAT_UNDEF;
            }
            new_level();
            enqueue(next, GPSAT_REASON_NONE);
        }
    }
};
typedef WarpSolverT<false> WarpSolver;

// ---------------------------------------------------------------------------------------------------------------
// glue shared by the kernel and the test-only emulator: point a WarpSolver at its memory, run one job, record it
// ---------------------------------------------------------------------------------------------------------------
template <class WS>
GPSAT_DEV void gpsat_bind(WS &S, const gpsat_formula_view &F, const gpsat_solve_params &P,
                          const gpsat_state_layout &Ly, int *state, int *arena, int *park, const gpsat_run_buffers &B)
{
    S.park = park;
    S.pool_mark = S.xpool_mark = 0;
#if !defined(GPSAT_WARP_EMU)
    S.lane_id = gpsat_read_lane();
#endif
    S.n_vars = F.n_vars;
    S.n_clauses = F.n_clauses;
    S.n_lits = F.n_lits;
    S.wbits_words = F.wbits_words;
    S.cl2 = (const gint2 *)F.cl2;
    S.ostart = F.ostart;
    S.occ2 = (const gint2 *)F.occ2;
    S.wbits0 = F.wbits0;
    S.vsids0 = F.vsids0;
    S.val0 = F.val0;
    S.val = (uint8_t *)(state + Ly.val);
    S.seen = (uint8_t *)(state + Ly.seen);
    S.level = (typename WS::idx_t *)(state + Ly.level);
    S.reason = state + Ly.reason;
    S.trail = (typename WS::idx_t *)(state + Ly.trail);
    S.trail_lim = (typename WS::idx_t *)(state + Ly.trail_lim);
    S.wbits = (uint32_t *)(state + Ly.wbits);
    S.vs = state + Ly.vs;
    S.lbuf = state + Ly.lbuf;
    S.lbuf_words = Ly.lbuf_words;
    S.cube_buf = state + Ly.cube;
    S.dynamic_split = P.dynamic_split;
    S.split_force = P.split_force;
    S.split_gap = P.split_gap > 0 ? P.split_gap : 1;
    S.split_burst = P.split_burst > 0 ? P.split_burst : 1;
    S.split_gap_hot = P.split_gap_hot > 0 ? P.split_gap_hot : S.split_gap;
    S.split_hot_demand = P.split_hot_demand > 0 ? P.split_hot_demand : 0x7fffffff;
    S.split_at_start = P.split_at_start;
    S.split_mode = P.split_mode;
    S.split_min = P.split_min;
    S.split_reserve = P.split_reserve;
    S.split_hard = P.split_hard > 0 ? P.split_hard : 0x7fffffff;
    S.inherited = 0;
    S.root = 0;
    S.dq_lits = B.dq_lits;
    S.dq_meta = B.dq_meta;
    S.dq_ctrl = B.dq_ctrl;
    S.root_pending = B.root_pending;
    S.dq_hand = B.dq_hand;
    S.hand_words = B.hand_words;
    S.dq_cap = B.dq_cap;
    S.mesh_ranks = B.mesh_ranks > 1 ? B.mesh_ranks : 0;
    S.rel_slot = -1;
    S.rel_seq = 0;
    S.rel_meta = B.dq_meta;
    S.t0 = B.t0;
    S.budget_ns = B.budget_ns;
    S.use_learnts = (P.mode == GPSAT_MODE_SOLVE) ? 1 : 0;
    S.arena = arena;
    S.arena_words = (int)P.arena_words;
    S.lw_head = arena;
    S.lwbits = (uint32_t *)(state + Ly.lwbits);
    S.phase_stats = P.phase_stats;
    S.phs = (long long *)(state + Ly.ph);
    S.phn = (int *)(state + Ly.ph + 2 * GPSAT_N_PHASES);
    S.hist = arena + 6 * F.n_vars;
    S.refs = arena + 6 * F.n_vars + 64;
    S.refs_cap = P.learnt_refs_cap;
    S.clause_base = 6 * F.n_vars + 64 + P.learnt_refs_cap;
    S.max_learnts = P.max_learnts_first;
    S.decision_mode = P.decision;
    S.restart_first = P.restart_first;
    S.restart_factor = P.restart_factor;
    S.max_iterations = P.max_iterations;
    S.max_conflicts = P.max_conflicts;
    S.share_learnts = P.share_learnts;
    S.share_max_len = P.share_max_len;
    S.share_import_max = P.share_import_max > 0 ? P.share_import_max : 256;
    S.pool = B.pool;
    S.pool_cursor = B.pool_cursor;
    S.pool_cap_words = B.pool_cap_words;
    S.xpool = B.xpool;
    S.xpool_cursor = B.xpool_cursor;
    S.facts = B.facts;
}

// Runs one job (an original cube, or a child produced by a split) and folds its outcome into the record of the
// original cube it descends from.
template <class WS>
GPSAT_DEV void gpsat_run_and_record(WS &S, int root, const int *cube, int k, const int *hand,
                                    const gpsat_solve_params &P, const gpsat_run_buffers &B, bool resume = false)
{
    GPSAT_LANE_DECL_S
    int confl;
    S.max_learnts = P.max_learnts_first;
    S.root = root;
    const int job = root;
    const int status = S.run_job(cube, k, P.mode, B.stop_flag, confl, hand, resume);

    const long long watchers = LANE_SUM_I64(S.l_watchers) + S.c_lwatchers;
    const long long words = LANE_SUM_I64(S.l_words) + S.c_lwords;
    LANE0
    {
        gpsat_job_record *r = B.records + job;
        gpsat_atomic_add_ll((long long *)&r->decisions, S.c_decisions);
        gpsat_atomic_add_ll((long long *)&r->implications, S.c_implications);
        gpsat_atomic_add_ll((long long *)&r->conflicts, S.c_conflicts);
        gpsat_atomic_add_ll((long long *)&r->learnt_clauses, S.c_learnt_clauses);
        gpsat_atomic_add_ll((long long *)&r->learnt_literals, S.c_learnt_literals);
        gpsat_atomic_add_ll((long long *)&r->restarts, S.c_restarts);
        gpsat_atomic_add_ll((long long *)&r->watchers_visited, watchers);
        gpsat_atomic_add_ll((long long *)&r->clause_words_read, words);
        gpsat_atomic_add_ll((long long *)&r->learnt_hash, (long long)S.c_hash);
        gpsat_atomic_add(&r->reserved, (int)S.c_splits);
        const int flag = status == GPSAT_SAT ? GPSAT_FLAG_SAT
                       : status == GPSAT_UNSAT ? GPSAT_FLAG_UNSAT
                       : status == GPSAT_UNDEF ? GPSAT_FLAG_UNDEF
                       : status == GPSAT_JOB_OOM ? GPSAT_FLAG_OOM
                       : status == GPSAT_JOB_SUSPENDED ? 0 : GPSAT_FLAG_ABORTED;
        gpsat_atomic_max(B.root_flag + job, flag);
    }
    if (P.mode == GPSAT_MODE_PROPAGATE) {
        if (B.conflict_clause) {
            long long cidx = -1;
            if (status == GPSAT_UNSAT && confl != GPSAT_NO_CONFLICT && confl >= 0) cidx = S.ld_cl2(confl - 1).y;
            LANE0 { B.conflict_clause[job] = cidx; }
        }
        // implied literals = trail entries with a reason, cube variables excluded, trail order
        LANES
        {
            GPSAT_NOUNROLL
            for (int i = lane; i < k; i += 32) S.seen[cube[i] >> 1] = 1;
        }
        SYNCWARP();
        int n_imp = 0;
        GPSAT_NOUNROLL
        for (int base = 0; base < S.trail_size; base += 32) {
            LANEVAR(int, x);
            LANEVAR(int, take);
            LANES
            {
                const int i = base + lane;
                LV(take) = 0;
                LV(x) = 0;
                if (i < S.trail_size) {
                    LV(x) = S.trail[i];
                    LV(take) = (S.reason[LV(x) >> 1] != GPSAT_REASON_NONE && !S.seen[LV(x) >> 1]) ? 1 : 0;
                }
            }
            const unsigned m = BALLOT(LV(take) != 0);
            if (B.implied) {
                LANES
                {
                    if (LV(take)) {
                        const long long dst = n_imp + gpsat_popc(m & GPSAT_LANEMASK_LT);
                        if (dst < P.implied_stride) B.implied[(long long)job * P.implied_stride + dst] = LV(x);
                    }
                }
            }
            n_imp += gpsat_popc(m);
        }
        SYNCWARP();
        LANES
        {
            GPSAT_NOUNROLL
            for (int i = lane; i < k; i += 32) S.seen[cube[i] >> 1] = 0;
        }
        if (B.n_implied) {
            LANE0 { B.n_implied[job] = n_imp; }
        }
        SYNCWARP();
    } else if (status == GPSAT_SAT) {
        LANEVAR(int, won);
        LANES { LV(won) = 0; }
        LANE0 { LV(won) = (gpsat_atomic_cas(B.sat_job, -1, job) == -1) ? 1 : 0; }
        if (SHFL(won, 0)) {
            LANES
            {
                GPSAT_NOUNROLL
                for (int v = lane; v < S.n_vars; v += 32) B.model[v] = (S.val[v] == GPSAT_VAL_FALSE) ? 0 : 1;
            }
            SYNCWARP();
            gpsat_threadfence();
            if (P.stop_on_sat) {
                LANE0 { gpsat_atomic_cas(B.stop_flag, 0, 1); }
            }
        }
    }
    // this job is closed: one fewer open descendant of the original cube, one fewer outstanding job
    // (a parked job stays open: the next launch resumes it)
    gpsat_threadfence();
    LANE0
    {
        if (status != GPSAT_JOB_SUSPENDED) {
            gpsat_atomic_add(B.root_pending + job, -1);
            gpsat_atomic_add(B.dq_ctrl + GPSAT_DQC_CLOSED, 1);
            if (P.phase_stats) {
                GPSAT_NOUNROLL
                for (int i = 0; i < GPSAT_N_PHASES; ++i) {
                    gpsat_atomic_add_ll((long long *)(B.dq_ctrl + GPSAT_DQC_PHASE) + i, S.phs[i]);
                    gpsat_atomic_add_ll((long long *)(B.dq_ctrl + GPSAT_DQC_PHASE) + GPSAT_N_PHASES + i, (long long)S.phn[i]);
                    S.phs[i] = 0;
                    S.phn[i] = 0;
                }
                gpsat_atomic_add_ll((long long *)(B.dq_ctrl + GPSAT_DQC_PHASE) + 2 * GPSAT_N_PHASES, S.phs[GPSAT_N_PHASES + GPSAT_N_PHASES / 2]);
                S.phs[GPSAT_N_PHASES + GPSAT_N_PHASES / 2] = 0;
            }
            if (hand != nullptr) {   // size of a split-off cube, in conflicts (diagnostics)
                int b = 0;
                GPSAT_NOUNROLL
                for (long long c = S.c_conflicts + 1; c > 1 && b < 15; c >>= 1) ++b;
                gpsat_atomic_add(B.dq_ctrl + GPSAT_DQC_HIST + b, 1);
            }
        }
    }
    SYNCWARP();
}

// Lane-0 code: takes the oldest child off a ring (this GPU's, or — `sys` — one that warps of several GPUs pop from).
// Returns the pop ticket, or -1 when the ring holds no published child.
// One atomic per step and no retry loop: a consumer first claims one of the published children on the AVAIL semaphore,
// only then draws a pop ticket with a fetch-add, and waits for exactly that slot if its writer is still a few
// microseconds from publishing (tickets are handed out in ring order, publications complete in any order).  The
// compare-and-swap ring this replaces collapsed under ~2000 idle warps: 6 million lost CAS for 23 thousand pops in one
// C2 shard, 1400 children queued in front of 1400 idle warps (profiles/r02_ring_*.txt).
GPSAT_DEV int gpsat_ring_pop(int *ctrl, int *meta, int cap, bool sys)
{
    if (gpsat_ld_volatile(ctrl + GPSAT_DQC_AVAIL) <= 0) return -1;
    const int had = sys ? gpsat_atomic_add_sys(ctrl + GPSAT_DQC_AVAIL, -1) : gpsat_atomic_add(ctrl + GPSAT_DQC_AVAIL, -1);
    if (had <= 0) {   // somebody else took the last one: give the claim back
        if (sys) gpsat_atomic_add_sys(ctrl + GPSAT_DQC_AVAIL, 1);
        else gpsat_atomic_add(ctrl + GPSAT_DQC_AVAIL, 1);
        return -1;
    }
    const int pos = sys ? gpsat_atomic_add_sys(ctrl + GPSAT_DQC_HEAD, 1) : gpsat_atomic_add(ctrl + GPSAT_DQC_HEAD, 1);
    const int slot = pos & (cap - 1);
    while (gpsat_ld_volatile(meta + 4 * slot + 2) != pos + 1) gpsat_nanosleep(200);
    return pos;
}

// The warp's main loop: original cubes from the atomic cursor (≙ JobsQueue::next_job, SATSolver/JobsQueue.cu:10-32),
// then children of split cubes from this GPU's ring, then — mesh — children queued on the other GPUs of the box, read
// over NVLink peer memory; a warp with nothing to do advertises itself as idle (so that long-running cubes split) and
// leaves when no job is open anywhere, the stop flag is up, or the step budget is spent.
template <class WS>
GPSAT_DEV void gpsat_warp_loop(WS &S, const gpsat_solve_params &P, const gpsat_run_buffers &B, int *stage)
{
    GPSAT_LANE_DECL_S
    int is_idle = 0, idle_spins = 0, rot = 0;
    unsigned long long busy_ns = 0;
    const bool mesh = B.mesh_ranks > 1;
    if (P.phase_stats) {   // the state block (shared memory) starts uninitialised
        LANE0
        {
            GPSAT_NOUNROLL
            for (int i = 0; i < GPSAT_N_PHASES + GPSAT_N_PHASES / 2 + 1; ++i) S.phs[i] = 0;
        }
        SYNCWARP();
    }
    while (true) {
        LANEVAR(int, kind_v);   // 0 exit, 1 original cube, 2 queued child, 3 wait, 4 resume the job this warp parked, 5 child of another GPU
        LANEVAR(int, idx_v);
        LANEVAR(int, peer_v);
        LANES
        {
            LV(kind_v) = 0;
            LV(idx_v) = 0;
            LV(peer_v) = 0;
        }
        LANE0
        {
            int kind = 3, idx = 0, peer = 0;
            if (gpsat_ld_volatile(B.stop_flag)) {
                kind = 0;
            } else if (B.budget_ns && gpsat_now_ns() > *B.t0 + B.budget_ns) {
                kind = 0;
            } else if (S.park != nullptr && gpsat_ld_volatile(S.park) == 1) {
                kind = 4;   // this warp parked a job at the end of the previous step
            } else {
                // root cubes: ONE cursor for all GPUs of a mesh (rank 0's, over NVLink) — whichever warp of whichever GPU is
                // free takes the next cube, so no GPU sits on unstarted cubes while another splits running ones
                if (gpsat_ld_volatile(B.dq_ctrl + GPSAT_DQC_ROOTS_DONE) == 0) {
                    idx = mesh ? gpsat_atomic_add_sys(B.next_job, 1) : gpsat_atomic_add(B.next_job, 1);
                    if (idx < B.n_cubes) kind = 1;
                    else ((volatile int *)B.dq_ctrl)[GPSAT_DQC_ROOTS_DONE] = 1;
                }
                if (kind == 3 && P.mode == GPSAT_MODE_SOLVE && P.dynamic_split) {
                    idx = gpsat_ring_pop(B.dq_ctrl, B.dq_meta, B.dq_cap, mesh);
                    if (idx >= 0) kind = 2;
                    // ... but only once this GPU is running dry (more than 1/8 of its warps idle): a warp that merely found
                    // its own ring empty for a moment gets a local child within microseconds
                    if (kind == 3 && mesh && stage != nullptr && !(P.mesh_flags & 1) &&
                        gpsat_ld_volatile(B.dq_ctrl + GPSAT_DQC_IDLE) >= P.split_hot_demand) {
                        // children advertised by the other GPUs (their communication warps refresh PEER_QUEUE): claim one
                        // locally first, so that at most as many warps go out over NVLink as there are children to take
                        GPSAT_NOUNROLL
                        for (int t = 0; t < B.mesh_ranks - 1 && kind == 3; ++t) {
                            const int r = (B.mesh_rank + 1 + (rot + t) % (B.mesh_ranks - 1)) % B.mesh_ranks;
                            if (gpsat_ld_volatile(B.dq_ctrl + GPSAT_DQC_PEER_QUEUE + r) <= 0) continue;
                            if (gpsat_atomic_add(B.dq_ctrl + GPSAT_DQC_PEER_QUEUE + r, -1) <= 0) continue;
                            int *rctrl = (int *)(B.mesh_base[r] + B.mesh_off_ctrl);
                            int *rmeta = (int *)(B.mesh_base[r] + B.mesh_off_meta);
                            gpsat_atomic_add(B.dq_ctrl + GPSAT_DQC_REMOTE_TRIES, 1);
                            idx = gpsat_ring_pop(rctrl, rmeta, B.dq_cap, true);
                            if (idx >= 0) {
                                kind = 5;
                                peer = r;
                            }
                        }
                        rot++;
                    }
                }
                if (kind == 3) {   // nothing to take: is anything still open?
                    if (gpsat_ld_volatile(B.dq_ctrl + GPSAT_DQC_DONE)) {
                        kind = 0;
                    } else if (!mesh) {
                        // one 8-byte snapshot of (created, closed): equal = no open job, and then none can appear
                        const unsigned long long cc = gpsat_ld_volatile64(B.dq_ctrl + GPSAT_DQC_CREATED);
                        if ((unsigned)(cc & 0xffffffffull) == (unsigned)(cc >> 32)) kind = 0;
                    }
                }
            }
            LV(kind_v) = kind;
            LV(idx_v) = idx;
            LV(peer_v) = peer;
        }
        const int kind = SHFL(kind_v, 0);
        const int idx = SHFL(idx_v, 0);
        if (kind == 0) break;
        if (kind == 3) {
            if (!is_idle && P.dynamic_split) {
                is_idle = 1;
                LANE0 { gpsat_atomic_add(B.dq_ctrl + GPSAT_DQC_IDLE, 1); }   // one more idle warp
            }
            // idle warps must not steal issue slots from the busy ones, nor hammer the queue counters in L2 (in the
            // tail of a run thousands of them poll the same sector that the splitting warps update): back off to 32 us
#if !defined(GPSAT_IDLE_NS)
#define GPSAT_IDLE_NS 4000u
#endif
            gpsat_nanosleep(GPSAT_IDLE_NS << (idle_spins < 3 ? idle_spins : 3));
            idle_spins++;
            continue;
        }
        idle_spins = 0;
        if (is_idle) {
            is_idle = 0;
            LANE0 { gpsat_atomic_add(B.dq_ctrl + GPSAT_DQC_IDLE, -1); }
        }
        const unsigned long long t_job = gpsat_now_ns();
        // ONE call site of the solver for every kind of job: each extra site would be another inlined copy of the whole
        // program, and the instruction cache is this kernel's first bottleneck (profiles/r02_cdcl_ncu_a.json)
        int job_root, job_k;
        const int *job_cube, *job_hand = nullptr;
        bool job_resume = false;
        if (kind == 4) {
            job_root = S.park[1];
            job_cube = S.park + 16;
            job_k = S.park[2];
            job_resume = true;
        } else if (kind == 1) {
            const long long c0 = B.cube_offsets[idx], c1 = B.cube_offsets[idx + 1];
            S.inherited = 0;
            job_root = idx;
            job_cube = B.cube_lits + c0;
            job_k = (int)(c1 - c0);
        } else if (kind == 2) {
            const int slot = idx & (B.dq_cap - 1);
            gpsat_threadfence();
            job_root = gpsat_ld_cg(B.dq_meta + 4 * slot);
            job_k = gpsat_ld_cg(B.dq_meta + 4 * slot + 1);
            S.inherited = gpsat_ld_cg(B.dq_meta + 4 * slot + 3);
            job_hand = B.dq_hand + (long long)slot * B.hand_words;
            job_cube = B.dq_lits + (long long)slot * GPSAT_DQ_MAXK;
            S.rel_slot = slot;
            S.rel_seq = idx + B.dq_cap;   // the ticket that may write this slot next
            S.rel_meta = B.dq_meta;
        } else {
            // a child queued on GPU `peer`: copy its cube and hand-off block over NVLink into this warp's staging block
            // (one bulk copy instead of a dependent remote load per imported clause), free the remote slot, run it here.
            // The job counts as closed on THIS GPU; its record goes into this GPU's arrays (merged at the end).
            const int peer = SHFL(peer_v, 0);
            const int slot = idx & (B.dq_cap - 1);
            int *rmeta = (int *)(B.mesh_base[peer] + B.mesh_off_meta);
            const int *rlits = (const int *)(B.mesh_base[peer] + B.mesh_off_lits) + (long long)slot * GPSAT_DQ_MAXK;
            const int *rhand = (const int *)(B.mesh_base[peer] + B.mesh_off_hand) + (long long)slot * B.hand_words;
            gpsat_threadfence_sys();
            job_root = gpsat_ld_cg(rmeta + 4 * slot);
            job_k = gpsat_ld_cg(rmeta + 4 * slot + 1);
            S.inherited = gpsat_ld_cg(rmeta + 4 * slot + 3);
            int used = gpsat_ld_cg(rhand);
            if (used < 0 || used > B.hand_words - 1 - 2 * S.n_vars) used = 0;
            LANES
            {
                gpsat_copy_cg(stage, rhand, 1 + 2 * S.n_vars + used, lane);
                gpsat_copy_cg(stage + B.hand_words, rlits, GPSAT_DQ_MAXK, lane);
            }
            SYNCWARP();
            gpsat_threadfence_sys();
            LANE0
            {
                ((volatile int *)rmeta)[4 * slot + 2] = idx + B.dq_cap;   // the remote slot is free again
                gpsat_atomic_add(B.dq_ctrl + GPSAT_DQC_STEALS, 1);
            }
            SYNCWARP();
            job_cube = stage + B.hand_words;
            job_hand = stage;
        }
        gpsat_run_and_record(S, job_root, job_cube, job_k, job_hand, P, B, job_resume);
        busy_ns += gpsat_now_ns() - t_job;
    }
    LANE0
    {
        if (is_idle) gpsat_atomic_add(B.dq_ctrl + GPSAT_DQC_IDLE, -1);
        if (B.busy_ns) gpsat_atomic_add_ll(B.busy_ns, (long long)busy_ns);   // utilisation = busy / (warps x kernel time)
    }
}
