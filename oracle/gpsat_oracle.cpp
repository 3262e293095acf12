/* TEST INFRASTRUCTURE ONLY — the CPU oracle of gpupsat_b200.
 *
 * Plain sequential C++ restatement of the reference's hot path (nvzoll/gpupsat), one job at a time, used by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg as the CHECKER of the CUDA path.  The product never
 * includes, links or calls this file (libgpsat.so has no CPU fallback).
 *
 * What is restated (reference file:line), and how the oracle is pinned:
 *   clause evaluation    VariablesStateHandler::clause_status (SATSolver/VariablesStateHandler.cu:180-206)
 *   BCP to fixpoint      WatchedClausesList::new_decision/process_clause/handle_implication
 *                        (BCPStrategy/WatchedClausesList.cu:46-101,103-221,253-282) driven like
 *                        ConflictAnalyzerWithWatchedLits::set_assumptions (ConflictAnalysis/ConflictAnalyzerWithWatchedLits.cu:46-111)
 *   conflict analysis    ConflictAnalyzer::handle_conflict_with_clause_learning (ConflictAnalysis/ConflictAnalyzer.cu:57-98);
 *                        the clause is the first-UIP clause BASELINE.json:north_star asks for instead of the reference's
 *                        decision cut (GraphAnalyzer.cu:92-138) — SURVEY.md §8a row 9 defines parity for that row
 *   learnt DB            LearntClausesManager::learn_clause (ClauseLearning/LearntClausesManager.cu:16-41), unbounded + reduction
 *   decisions            DecisionMaker::new_literal / VSIDS (SATSolver/DecisionMaker.cu:45-55, DecisionStrategy/VSIDS.cu:77-124)
 *   restarts             GeometricRestartsManager (Restarts/GeometricRestartsManager.cu:16-31), used as SATSolver.cu:170-178
 *   job driver           SATSolver::solve / preprocess (SATSolver/SATSolver.cu:67-218,231-272)
 * Pinning: tests/test_oracle.py and tests/test_host_layers.py checks this file against oracle/_ref (the reference's own sources built
 * for the host) — verdicts on tests/cnf + uf20/uf50/PHP, BCP implication SETS and conflict status on every cube of
 * random instances — and tests/golden/ holds those reference outputs as committed fixtures for machines without
 * /root/reference.  The reference ships no vectors for implication lists or learnt clauses, so the
 * ORDER below (which clause reports a conflict first, the literal order of a learnt clause) is this repo's canonical
 * order, documented in DESIGN.md §"canonical order"; the CUDA kernel must reproduce it bit-exactly.
 *
 * The data structures here are deliberately different from the kernel's (clause-major watch flags and std::vector
 * watch lists instead of an occurrence bitmap and a bump arena): agreement is evidence, not tautology.
 */
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

enum { SAT = 0, UNSAT = 1, UNDEF = 2 };
enum { V_FALSE = 0, V_TRUE = 1, V_UNDEF = 2, V_ABSENT = 4 };
enum { DECIDE_REFERENCE = 0, DECIDE_VSIDS = 1 };
enum { MODE_SOLVE = 0, MODE_PROPAGATE = 1, MODE_PROPAGATE_BATCH = 2 };
const int JOB_OOM = -3;

struct Params {
    int32_t mode, decision, restart_first;
    float restart_factor;
    int32_t max_iterations;
    int64_t max_conflicts;
    int32_t max_learnts_first, learnt_refs_cap;
    int64_t arena_words;
};

struct Record {   /* same fields as gpsat_job_record */
    int32_t status, reserved;
    int64_t decisions, implications, conflicts, learnt_clauses, learnt_literals, restarts, watchers_visited,
        clause_words_read, learnt_hash;
};

struct Occ { int32_t clause, pos; };
struct LWatch { int32_t id, blocker; };

/* reason encoding: -1 none, >= 0 original clause index, <= -2 learnt clause id = -2 - reason */
struct Oracle {
    int n_vars = 0, n_clauses = 0;
    std::vector<int64_t> off;
    std::vector<int32_t> lits;
    std::vector<std::vector<Occ>> occ;          /* per literal, ascending clause index */
    std::vector<int32_t> vsids0;
    std::vector<uint8_t> val0;

    /* per job */
    std::vector<uint8_t> val, seen;
    std::vector<uint8_t> watched;               /* per literal slot of the original formula */
    std::vector<int32_t> level, reason, trail, trail_lim, vs;
    std::vector<std::vector<int32_t>> learnt;   /* learnt[id] = literals, watched at [0],[1]; empty = deleted */
    std::vector<int32_t> learnt_order;          /* live ids in age order */
    std::vector<std::vector<LWatch>> lw;
    size_t qhead = 0;
    int dlevel = 0;
    int conflicts_since_restart = 0, restart_limit = 0, vs_clauses = 0, max_learnts = 0;
    Params P{};
    Record R{};
    /* arena accounting that mirrors the kernel's out-of-memory behaviour */
    int64_t arena_top = 0, watch_bot = 0, clause_base = 0;
    std::vector<int32_t> lw_cap;
    bool oom = false;

    int lit_value(int x) const
    {
        int v = val[x >> 1];
        return v >= 2 ? 2 : (v ^ (x & 1) ^ 1);
    }
    void enqueue(int x, int why)
    {
        int v = x >> 1;
        val[v] = (uint8_t)(x & 1);
        level[v] = dlevel;
        reason[v] = why;
        trail.push_back(x);
    }
    void new_level()
    {
        trail_lim.push_back((int32_t)trail.size());
        dlevel++;
    }
    void cancel_until(int lv)
    {
        if (dlevel <= lv) return;
        size_t start = (size_t)trail_lim[lv];
        for (size_t i = start; i < trail.size(); i++) val[trail[i] >> 1] = val0[trail[i] >> 1];
        trail.resize(start);
        trail_lim.resize(lv);
        qhead = start;
        dlevel = lv;
    }

    /* bump-arena bookkeeping of a learnt watch vector growing by one entry (kernel: lw_append) */
    void account_append(int x)
    {
        int n = (int)lw[x].size();   /* size before the append */
        if (n == lw_cap[x]) {
            int ncap = lw_cap[x] ? 2 * lw_cap[x] : 4;
            int64_t nptr = watch_bot - 2 * (int64_t)ncap;
            if (nptr < arena_top) { oom = true; return; }
            watch_bot = nptr;
            lw_cap[x] = ncap;
        }
    }
    void lw_append(int x, int id, int blocker)
    {
        account_append(x);
        if (oom) return;
        lw[x].push_back(LWatch{id, blocker});
    }

    /* ---- BCP.  Returns reason-encoded conflict or INT32_MIN ------------------------------------------------ */
    /* Canonical order: GROUPS of consecutive trail literals whose occurrence lists fit 32 slots
     * together (at least one literal; a literal with more than 32 occurrences is a group of its own, examined 32 slots
     * at a time).  All slots of a group are examined against the state at group start; if a clause is examined twice
     * (two of its watched literals falsified by two literals of the group) the group is cut before the literal that owns
     * the second occurrence.  Watch moves of the surviving slots are applied, then units / conflicts committed in slot
     * order; then the learnt clauses watching the group's literals, literal after literal. */
    int propagate()
    {
        const int NOC = INT32_MIN;
        while (qhead < trail.size()) {
            const size_t avail = std::min<size_t>(trail.size() - qhead, 32);
            std::vector<int> gf(avail);
            for (size_t i = 0; i < avail; i++) gf[i] = trail[qhead + i] ^ 1;
            struct Slot { int li; size_t t; };
            std::vector<Slot> slots;
            size_t g = 0;
            if (occ[gf[0]].size() > 32) {
                g = 1;
                for (size_t t = 0; t < occ[gf[0]].size(); t++) slots.push_back(Slot{0, t});
            } else {
                for (size_t i = 0; i < avail; i++) {
                    const size_t c = occ[gf[i]].size();
                    if (i > 0 && slots.size() + c > 32) break;
                    for (size_t t = 0; t < c; t++) slots.push_back(Slot{(int)i, t});
                    g++;
                }
            }
            size_t cut = g;
            for (size_t base = 0; base < slots.size(); base += 32) {
                const size_t end = std::min(slots.size(), base + 32);
                struct Act { int kind, lit, clause, newpos, li, nread; };
                std::vector<Act> acts;   /* one per EXAMINED slot (kind 0 = nothing to do) */
                acts.reserve(32);
                for (size_t q = base; q < end; q++) {                       /* examine against the group-start state */
                    const Occ &o = occ[gf[slots[q].li]][slots[q].t];
                    const int c = o.clause, mypos = o.pos;
                    const int64_t b = off[c];
                    const int len = (int)(off[c + 1] - b);
                    if (!watched[b + mypos]) continue;
                    int other = -1, other_val = 2, repl = -1, nread = 0;
                    for (int i = 0; i < len; i++) {
                        nread++;
                        if (i == mypos) continue;
                        const int x = lits[b + i];
                        const int v = lit_value(x);
                        if (watched[b + i]) {
                            other = x; other_val = v;
                            if (v == 1) break;
                        } else if (v != 0 && repl < 0) {
                            repl = i;
                        }
                        if (other >= 0 && repl >= 0) break;
                    }
                    Act a{0, 0, c, 0, slots[q].li, nread};
                    if (other_val == 1) a.kind = 0;
                    else if (repl >= 0) { a.kind = 1; a.lit = mypos; a.newpos = repl; }
                    else if (other_val == 2) { a.kind = 2; a.lit = other; }
                    else a.kind = 3;
                    acts.push_back(a);
                }
                if (g > 1)                                                   /* a clause examined twice: cut the group */
                    for (size_t x = 0; x < acts.size() && cut == g; x++)
                        for (size_t y = 0; y < x; y++)
                            if (acts[y].clause == acts[x].clause) { cut = (size_t)acts[x].li; break; }
                for (const Act &a : acts)
                    if ((size_t)a.li < cut) {
                        R.watchers_visited++;
                        R.clause_words_read += a.nread;
                        if (a.kind == 1) {                                   /* all watch moves of the chunk */
                            watched[off[a.clause] + a.lit] = 0;
                            watched[off[a.clause] + a.newpos] = 1;
                        }
                    }
                for (const Act &a : acts) {                                  /* units / conflicts in slot order */
                    if ((size_t)a.li >= cut) continue;
                    if (a.kind == 3) return a.clause;
                    if (a.kind == 2) {
                        const int v = lit_value(a.lit);
                        if (v == 2) { enqueue(a.lit, a.clause); R.implications++; }
                        else if (v == 0) return a.clause;
                    }
                }
            }
            qhead += cut;
            for (size_t gi = 0; gi < cut; gi++) {
            const int f = gf[gi];
            /* learnt clauses watching f: sequential two-watched-literal scheme with blockers */
            if (P.mode != MODE_SOLVE) continue;
            std::vector<LWatch> &ws = lw[f];
            if (ws.empty()) continue;
            R.watchers_visited += (int64_t)ws.size();
            int confl = NOC;
            size_t j = 0;
            for (size_t i = 0; i < ws.size(); i++) {
                LWatch w = ws[i];
                if (confl != NOC || lit_value(w.blocker) == 1) { ws[j++] = w; continue; }
                std::vector<int32_t> &cl = learnt[w.id];
                if (cl[0] == f) std::swap(cl[0], cl[1]);
                const int first = cl[0];
                R.clause_words_read += 2;
                if (first != w.blocker && lit_value(first) == 1) { ws[j++] = LWatch{w.id, first}; continue; }
                int found = -1;
                const int len = (int)cl.size();
                for (int q = 2; q < len; q++)
                    if (lit_value(cl[q]) != 0) { found = q; break; }
                if (len > 2) {   /* the kernel reads clause words 32 at a time */
                    int scanned = found >= 0 ? ((found - 2) / 32 + 1) * 32 : len - 2;
                    R.clause_words_read += std::min(scanned, len - 2);
                }
                if (found >= 0) {
                    const int nl = cl[found];
                    cl[1] = nl; cl[found] = f;
                    lw_append(nl, w.id, first);
                    if (oom) { cl[found] = nl; cl[1] = f; ws[j++] = w; confl = -2 - w.id; }
                    continue;
                }
                ws[j++] = LWatch{w.id, first};
                if (lit_value(first) == 0) confl = -2 - w.id;
                else { enqueue(first, -2 - w.id); R.implications++; }
            }
            ws.resize(j);
            if (confl != NOC) return confl;
            }   /* literals of the group */
        }
        return NOC;
    }

    void clause_lits(int why, const int32_t *&p, int &len) const
    {
        if (why >= 0) { p = &lits[off[why]]; len = (int)(off[why + 1] - off[why]); }
        else { const std::vector<int32_t> &c = learnt[-2 - why]; p = c.data(); len = (int)c.size(); }
    }

    /* first UIP; literals in discovery order, highest remaining level moved to slot 1 (first such literal) */
    std::vector<int32_t> analyze(int confl, int &bt)
    {
        std::vector<int32_t> out(1, -1);
        int pathC = 0, p = -1;
        int index = (int)trail.size() - 1;
        do {
            const int32_t *cl; int len;
            clause_lits(confl, cl, len);
            for (int i = 0; i < len; i++) {
                const int x = cl[i], v = x >> 1;
                if (x == p || seen[v] || level[v] <= 0) continue;
                seen[v] = 1;
                if (level[v] >= dlevel) pathC++; else out.push_back(x);
            }
            while (!seen[trail[index] >> 1]) index--;
            p = trail[index--];
            confl = reason[p >> 1];
            seen[p >> 1] = 0;
            pathC--;
        } while (pathC > 0);
        out[0] = p ^ 1;
        bt = 0;
        if (out.size() > 1) {
            size_t best = 1;
            for (size_t i = 2; i < out.size(); i++)
                if (level[out[i] >> 1] > level[out[best] >> 1]) best = i;
            bt = level[out[best] >> 1];
            std::swap(out[1], out[best]);
        }
        for (size_t i = 1; i < out.size(); i++) seen[out[i] >> 1] = 0;
        return out;
    }

    void hash_learnt(const std::vector<int32_t> &c)
    {
        uint64_t s = 0;
        for (size_t i = 0; i < c.size(); i++) s += (uint64_t)(c[i] + 1) * (uint64_t)(i + 1) * 0x9E3779B97F4A7C15ull;
        uint64_t h = (uint64_t)R.learnt_hash;
        h = h * 0x100000001B3ull + s + (uint64_t)c.size();
        R.learnt_hash = (int64_t)h;
    }

    int learn(const std::vector<int32_t> &c)
    {
        if ((int)learnt_order.size() >= P.learnt_refs_cap || arena_top + (int64_t)c.size() + 1 > watch_bot) { oom = true; return 0; }
        arena_top += (int64_t)c.size() + 1;
        const int id = (int)learnt.size();
        learnt.push_back(c);
        learnt_order.push_back(id);
        lw_append(c[0], id, c[1]);
        if (!oom) lw_append(c[1], id, c[0]);
        return id;
    }

    bool locked(int id) const
    {
        const int x0 = learnt[id][0];
        return lit_value(x0) == 1 && reason[x0 >> 1] == -2 - id;
    }

    /* drop the longer half of the unlocked learnt clauses of length > 2 (oldest first inside the threshold length),
     * then rebuild the learnt watch lists in age order */
    void reduce_db()
    {
        int hist[64] = {0};
        for (int id : learnt_order) {
            const int len = (int)learnt[id].size();
            if (len > 2 && !locked(id)) hist[std::min(len, 63)]++;
        }
        int n_cand = 0;
        for (int b = 0; b < 64; b++) n_cand += hist[b];
        const int target = n_cand / 2;
        int thr = 64, partial = 0, acc = 0;
        for (int b = 63; b >= 3 && acc < target; --b) {
            if (acc + hist[b] >= target) { thr = b; partial = target - acc; acc = target; }
            else acc += hist[b];
        }
        std::vector<int32_t> keep;
        int64_t top = clause_base;
        for (int id : learnt_order) {
            const int len = (int)learnt[id].size();
            bool remove = false;
            if (len > 2 && !locked(id)) {
                const int b = std::min(len, 63);
                if (b > thr) remove = true;
                else if (b == thr && partial > 0) { partial--; remove = true; }
            }
            if (remove) { learnt[id].clear(); learnt[id].shrink_to_fit(); }
            else { keep.push_back(id); top += len + 1; }
        }
        learnt_order.swap(keep);
        arena_top = top;
        std::vector<int> cnt((size_t)2 * n_vars, 0);
        for (int id : learnt_order) { cnt[learnt[id][0]]++; cnt[learnt[id][1]]++; }
        watch_bot = P.arena_words;
        for (int x = 0; x < 2 * n_vars; x++) {
            lw[x].clear();
            lw_cap[x] = cnt[x] > 0 ? 2 * cnt[x] + 4 : 0;
            if (cnt[x] > 0) watch_bot -= 2 * (int64_t)lw_cap[x];
        }
        if (watch_bot < arena_top) { oom = true; return; }
        for (int id : learnt_order) {
            lw_append(learnt[id][0], id, learnt[id][1]);
            lw_append(learnt[id][1], id, learnt[id][0]);
        }
        max_learnts = max_learnts + max_learnts / 10 + 1;
        if (max_learnts > P.learnt_refs_cap - n_vars - 2) max_learnts = P.learnt_refs_cap - n_vars - 2;
    }

    void vsids_learnt(const std::vector<int32_t> &c)
    {
        for (int x : c) vs[x]++;
        vs_clauses++;
        if (vs_clauses % 50 == 0)
            for (auto &s : vs) s /= 2;
    }

    int pick_branch() const
    {
        if (P.decision == DECIDE_VSIDS) {
            int best = -1, best_score = -1;
            for (int v = 0; v < n_vars; v++) {
                if (val[v] != V_UNDEF) continue;
                if (vs[2 * v + 1] > best_score) { best_score = vs[2 * v + 1]; best = 2 * v + 1; }
                if (vs[2 * v] > best_score) { best_score = vs[2 * v]; best = 2 * v; }
            }
            return best;
        }
        for (int v = n_vars - 1; v >= 0; v--)
            if (val[v] == V_UNDEF) return 2 * v + 1;
        return -1;
    }

    void reset_job()
    {
        val = val0;
        seen.assign((size_t)n_vars, 0);
        level.assign((size_t)n_vars, 0);
        reason.assign((size_t)n_vars, -1);
        trail.clear();
        trail_lim.clear();
        watched.assign(lits.size(), 0);
        for (int c = 0; c < n_clauses; c++) { watched[off[c]] = 1; watched[off[c] + 1] = 1; }
        vs = vsids0;
        learnt.clear();
        learnt_order.clear();
        lw.assign((size_t)2 * n_vars, std::vector<LWatch>());
        lw_cap.assign((size_t)2 * n_vars, 0);
        qhead = 0; dlevel = 0;
        conflicts_since_restart = 0;
        restart_limit = P.restart_first;
        vs_clauses = n_clauses;
        max_learnts = P.max_learnts_first;
        clause_base = 6 * (int64_t)n_vars + 64 + P.learnt_refs_cap;
        arena_top = clause_base;
        watch_bot = P.arena_words;
        oom = false;
        std::memset(&R, 0, sizeof(R));
    }

    int run_job(const int32_t *cube, int k, int &conflict_out)
    {
        const int NOC = INT32_MIN;
        conflict_out = NOC;
        reset_job();
        if (P.mode == MODE_PROPAGATE_BATCH) {
            /* the reference's own order: the whole cube is assigned first (VariablesStateHandler::set_assumptions,
             * SATSolver.cu:231-246), then BCP runs to fixpoint or first conflict */
            new_level();
            for (int i = 0; i < k; i++) {
                const int v = lit_value(cube[i]);
                if (v == 0) return UNSAT;
                if (v == 2) enqueue(cube[i], -1);
            }
            const int confl = propagate();
            if (confl != NOC) { conflict_out = confl; R.conflicts++; return UNSAT; }
            return UNDEF;
        }
        while (true) {
            const int confl = propagate();
            if (oom) return JOB_OOM;
            if (confl != NOC) {
                conflict_out = confl;
                R.conflicts++;
                conflicts_since_restart++;
                if (dlevel == 0 || P.mode == MODE_PROPAGATE) return UNSAT;
                int bt;
                std::vector<int32_t> c = analyze(confl, bt);
                hash_learnt(c);
                R.learnt_clauses++;
                R.learnt_literals += (int64_t)c.size();
                cancel_until(bt);
                if (c.size() == 1) enqueue(c[0], -1);
                else {
                    const int id = learn(c);
                    if (oom) return JOB_OOM;
                    enqueue(c[0], -2 - id);
                }
                R.implications++;
                if (P.decision == DECIDE_VSIDS) vsids_learnt(c);
                if (P.max_conflicts && R.conflicts >= P.max_conflicts) return UNDEF;
                continue;
            }
            if (P.mode == MODE_PROPAGATE && dlevel >= k) return UNDEF;
            if (P.restart_first > 0 && conflicts_since_restart >= restart_limit) {
                conflicts_since_restart = 0;
                restart_limit = (int)((float)restart_limit * P.restart_factor);
                R.restarts++;
                cancel_until(std::min(k, dlevel));
            }
            if (P.mode == MODE_SOLVE &&
                ((int)learnt_order.size() >= max_learnts || (watch_bot - arena_top) < (P.arena_words - clause_base) / 4)) {
                reduce_db();
                if (oom) return JOB_OOM;
            }
            int next = -1;
            while (dlevel < k) {
                const int x = cube[dlevel];
                const int v = lit_value(x);
                if (v == 1) new_level();
                else if (v == 0) {   /* cube literal already false: its reason clause is falsified under the full cube */
                    if (reason[x >> 1] != -1) conflict_out = reason[x >> 1];
                    return UNSAT;
                }
                else { next = x; break; }
            }
            if (next < 0) {
                if (P.mode == MODE_PROPAGATE) return UNDEF;
                next = pick_branch();
                if (next < 0) return SAT;
                R.decisions++;
                if (P.max_iterations && R.decisions > P.max_iterations) return UNDEF;
            }
            new_level();
            enqueue(next, -1);
        }
    }
};

}  // namespace

extern "C" {

void *oracle_open(int32_t n_vars, int64_t n_clauses, const int64_t *offsets, const int32_t *lits)
{
    Oracle *o = new Oracle();
    o->n_vars = n_vars;
    o->n_clauses = (int)n_clauses;
    o->off.assign(offsets, offsets + n_clauses + 1);
    const int64_t base = n_clauses ? offsets[0] : 0;
    for (auto &x : o->off) x -= base;
    o->lits.assign(lits + base, lits + base + o->off[n_clauses]);
    o->occ.assign((size_t)2 * n_vars, std::vector<Occ>());
    o->val0.assign((size_t)n_vars, V_ABSENT);
    o->vsids0.assign((size_t)2 * n_vars, 0);
    /* VSIDS::handle_clause over the formula (DecisionMaker.cu:3-16) with VSIDS::decay every 50 clauses (VSIDS.cu:84-89).
     * Small formulas: literally that.  Large ones: each counter is halved lazily for the decays it missed, which
     * yields the same integers (repeated floor-halving == right shift). */
    const bool literal_decay = n_clauses <= 20000;
    std::vector<int32_t> seen_epoch((size_t)2 * n_vars, 0);
    int32_t epoch = 0;
    for (int c = 0; c < (int)n_clauses; c++) {
        for (int64_t i = o->off[c]; i < o->off[c + 1]; i++) {
            const int x = o->lits[i];
            o->occ[x].push_back(Occ{c, (int32_t)(i - o->off[c])});
            o->val0[x >> 1] = V_UNDEF;
            if (!literal_decay && epoch > seen_epoch[x]) {
                const int d = epoch - seen_epoch[x];
                o->vsids0[x] = d >= 31 ? 0 : (o->vsids0[x] >> d);
                seen_epoch[x] = epoch;
            }
            o->vsids0[x]++;
        }
        if ((c + 1) % 50 == 0) {
            if (literal_decay) for (auto &s : o->vsids0) s /= 2;
            else epoch++;
        }
    }
    if (!literal_decay)
        for (size_t x = 0; x < o->vsids0.size(); x++) {
            const int d = epoch - seen_epoch[x];
            if (d > 0) o->vsids0[x] = d >= 31 ? 0 : (o->vsids0[x] >> d);
        }
    return o;
}
void oracle_close(void *h) { delete (Oracle *)h; }

/* params: {mode, decision, restart_first, max_iterations, max_learnts_first, learnt_refs_cap} + restart_factor,
 * max_conflicts, arena_words.  Outputs per cube as the C ABI's gpsat_job_record; for PROPAGATE mode also the implied
 * literals (trail order, cube variables excluded) and the falsified clause index (-1 none / learnt). */
int oracle_run(void *h, const int32_t *iparams, float restart_factor, int64_t max_conflicts, int64_t arena_words,
               int32_t n_cubes, const int64_t *cube_offsets, const int32_t *cube_lits, void *records,
               uint8_t *model, int32_t *sat_job, int32_t stop_on_sat, int32_t *implied, int64_t implied_stride,
               int32_t *n_implied, int64_t *conflict_clause)
{
    Oracle *o = (Oracle *)h;
    o->P.mode = iparams[0];
    o->P.decision = iparams[1];
    o->P.restart_first = iparams[2];
    o->P.max_iterations = iparams[3];
    o->P.max_learnts_first = iparams[4];
    o->P.learnt_refs_cap = iparams[5];
    o->P.restart_factor = restart_factor;
    o->P.max_conflicts = max_conflicts;
    o->P.arena_words = arena_words;
    Record *rec = (Record *)records;
    *sat_job = -1;
    for (int j = 0; j < n_cubes; j++) rec[j].status = -1;
    for (int j = 0; j < n_cubes; j++) {
        const int32_t *cube = cube_lits + cube_offsets[j];
        const int k = (int)(cube_offsets[j + 1] - cube_offsets[j]);
        int confl;
        const int st = o->run_job(cube, k, confl);
        o->R.status = st;
        rec[j] = o->R;
        if (o->P.mode != MODE_SOLVE) {
            if (conflict_clause) conflict_clause[j] = (st == UNSAT && confl != INT32_MIN && confl >= 0) ? confl : -1;
            std::vector<uint8_t> in_cube((size_t)o->n_vars, 0);
            for (int i = 0; i < k; i++) in_cube[cube[i] >> 1] = 1;
            int n = 0;
            for (int x : o->trail) {
                if (o->reason[x >> 1] == -1 || in_cube[x >> 1]) continue;
                if (implied && n < implied_stride) implied[(int64_t)j * implied_stride + n] = x;
                n++;
            }
            if (n_implied) n_implied[j] = n;
        } else if (st == SAT && *sat_job < 0) {
            *sat_job = j;
            if (model)
                for (int v = 0; v < o->n_vars; v++) model[v] = o->val[v] == V_FALSE ? 0 : 1;
            if (stop_on_sat) break;
        }
    }
    return 0;
}

/* VariablesStateHandler::clause_status for every clause: status SAT(0)/UNSAT(1)/UNDEF(2); unit = the LAST unassigned
 * literal when exactly len-1 literals are false, else -1.  assignment[v] uses the reference's sat_status encoding
 * (0 = true, 1 = false, 2 = unassigned). */
int oracle_eval_clauses(void *h, int32_t n_assignments, const uint8_t *assignment, int32_t *status, int32_t *unit)
{
    Oracle *o = (Oracle *)h;
    for (int a = 0; a < n_assignments; a++) {
        const uint8_t *as = assignment + (int64_t)a * o->n_vars;
        for (int c = 0; c < o->n_clauses; c++) {
            int n_false = 0, last_undef = -1, st = -1;
            const int len = (int)(o->off[c + 1] - o->off[c]);
            for (int64_t i = o->off[c]; i < o->off[c + 1]; i++) {
                const int x = o->lits[i];
                const int sv = as[x >> 1];                       /* sat_status of the variable */
                int ls = sv == 2 ? UNDEF : ((sv == 0) == ((x & 1) == 1) ? SAT : UNSAT);
                if (ls == SAT) { st = SAT; break; }
                if (ls == UNSAT) n_false++;
                else last_undef = x;
            }
            int u = -1;
            if (st != SAT) {
                st = n_false == len ? UNSAT : UNDEF;
                if (n_false == len - 1) u = last_undef;
            }
            status[(int64_t)a * o->n_clauses + c] = st;
            if (unit) unit[(int64_t)a * o->n_clauses + c] = u;
        }
    }
    return 0;
}

}  /* extern "C" */
