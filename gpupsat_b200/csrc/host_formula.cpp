// Host-side layers that decide what the hot path sees: DIMACS in, host preprocessing, cube generation, and the
// static index (CSR + occurrence lists) the kernels read.  Behaviour follows the reference (file:line cited per
// function); the code is new.
#include "host_formula.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>

#include "../../include/gpsat.h"

namespace gpsat_host {

static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
const char *last_error() { return g_err.c_str(); }

// ---------------------------------------------------------------------------------------------------------------
// DIMACS.  The reference parses with a Boost.Spirit grammar (FileManager/CnfReader.cpp:30-50):
//     *( "c" ... eol | eol )  >>  "p" "cnf" uint uint (eol|eoi)  >>  *( *(int - 0) 0 eol )  >>  (eol|eoi)
// with blanks skipped.  Consequences kept here: comments and blank lines only before the header; the header is
// mandatory; one clause per line, closed by 0 on that line; an empty clause line "0" is a clause with no literals;
// the clause block ends at EOF or at an empty line (anything after it is ignored); any other line is a parse error.
// n_lines follows FileManager/FileUtils.cu:4-27.
// ---------------------------------------------------------------------------------------------------------------
static int count_lines(const std::string &text)
{
    int n = 0;
    for (char ch : text)
        if (ch == '\n') n++;
    // the reference adds one for a last line without '\n' only when at least one '\n' was seen; its loop ends on
    // EOF, so the "last char" it tests is always EOF != '\n' -> it always adds one when n != 0
    if (n != 0) n++;
    return n;
}

static bool is_blank_line(const std::string &s)
{
    for (char ch : s)
        if (ch != ' ' && ch != '\t' && ch != '\r') return false;
    return true;
}

int read_dimacs(const char *path, gpsat_cnf &out)
{
    std::ifstream in(path, std::ios::binary);
    if (!in) {
        set_error(std::string("cannot open ") + (path ? path : "(null)"));
        return GPSAT_E_IO;
    }
    std::stringstream ss;
    ss << in.rdbuf();
    const std::string text = ss.str();
    out = gpsat_cnf();
    out.n_lines = count_lines(text);

    size_t pos = 0;
    auto next_line = [&](std::string &line) -> bool {
        if (pos >= text.size()) return false;
        size_t e = text.find('\n', pos);
        if (e == std::string::npos) e = text.size();
        line.assign(text, pos, e - pos);
        if (!line.empty() && line.back() == '\r') line.pop_back();
        pos = e + 1;
        return true;
    };

    std::string line;
    bool have_header = false;
    while (next_line(line)) {
        size_t i = line.find_first_not_of(" \t");
        if (i == std::string::npos) continue;                 // blank line
        if (line[i] == 'c') continue;                         // comment
        if (line[i] == 'p') {
            char tag[16] = {0};
            long long v = -1, c = -1;
            if (std::sscanf(line.c_str() + i, "p %15s %lld %lld", tag, &v, &c) != 3 || std::strcmp(tag, "cnf") != 0 ||
                v < 0 || c < 0) {
                set_error("malformed DIMACS header: " + line);
                return GPSAT_E_PARSE;
            }
            out.header_vars = (int32_t)v;
            out.header_clauses = c;
            have_header = true;
            break;
        }
        set_error("text before the 'p cnf' header: " + line);
        return GPSAT_E_PARSE;
    }
    if (!have_header) {
        set_error("missing 'p cnf' header");
        return GPSAT_E_PARSE;
    }

    while (next_line(line)) {
        if (is_blank_line(line)) break;                        // grammar's trailing eol: rest of file ignored
        const char *p = line.c_str();
        char *end = nullptr;
        bool closed = false;
        size_t first = out.lits.size();
        while (true) {
            while (*p == ' ' || *p == '\t') p++;
            if (*p == 0) break;
            long k = std::strtol(p, &end, 10);
            if (end == p) {
                set_error("unexpected text in clause line: " + line);
                return GPSAT_E_PARSE;
            }
            p = end;
            if (k == 0) {
                closed = true;
                while (*p == ' ' || *p == '\t') p++;
                if (*p != 0) {
                    set_error("text after the closing 0: " + line);
                    return GPSAT_E_PARSE;
                }
                break;
            }
            long v = std::labs(k) - 1;
            out.lits.push_back((int32_t)(2 * v + (k > 0 ? 1 : 0)));
        }
        if (!closed) {
            set_error("clause line not closed by 0: " + line);
            return GPSAT_E_PARSE;
        }
        if (out.lits.size() - first > 10000) {                 // MAX_CLAUSE_SIZE, SATSolver/Configs.cuh:72
            set_error("clause longer than MAX_CLAUSE_SIZE (10000)");
            return GPSAT_E_PARSE;
        }
        out.offsets.push_back((int64_t)out.lits.size());
    }
    finish_raw(out);
    return GPSAT_OK;
}

// CnfManager::add_lit bookkeeping (FileManager/CnfReader.cpp:57-84): n_vars = highest var + 1; most frequent
// variable = first one to reach the running maximum count; largest clause as FormulaData::add_clause sees it
// (FileManager/FormulaData.cu:31-33).
void finish_raw(gpsat_cnf &f)
{
    f.n_vars = -1;
    f.largest_clause = 0;
    f.most_common_var = -1;
    f.most_common_freq = -1;
    std::map<int32_t, int32_t> occurrences;
    for (int64_t c = 0; c < f.n_clauses(); c++) {
        int64_t b = f.offsets[c], e = f.offsets[c + 1];
        f.largest_clause = std::max<int32_t>(f.largest_clause, (int32_t)(e - b));
        for (int64_t i = b; i < e; i++) {
            int32_t v = f.lits[i] >> 1;
            if (v + 1 > f.n_vars) f.n_vars = v + 1;
            int32_t cnt = ++occurrences[v];
            if (cnt > f.most_common_freq) {
                f.most_common_freq = cnt;
                f.most_common_var = v;
            }
        }
    }
    f.status = GPSAT_UNDEF;
}

// ---------------------------------------------------------------------------------------------------------------
// Preprocessing = RepeatedLiteralsRemover::process then UnaryClausesRemover::process, as FormulaData::
// copy_host_clauses_to_dev chains them (FileManager/FormulaData.cu:82-105).
// ---------------------------------------------------------------------------------------------------------------
namespace {
using ClauseList = std::vector<std::vector<int32_t>>;

// Preprocessing/RepeatedLiteralsRemover.cu:25-62: inside a clause keep the first copy of a repeated literal, drop
// the whole clause when a variable occurs with both signs; formula empty afterwards -> SAT.
int drop_repeats(ClauseList &f, int32_t n_vars)
{
    std::vector<int8_t> seen((size_t)std::max(n_vars, 1), -1);
    ClauseList kept;
    kept.reserve(f.size());
    for (auto &c : f) {
        bool tautology = false;
        std::vector<int32_t> r;
        r.reserve(c.size());
        for (int32_t x : c) {
            int8_t &s = seen[(size_t)(x >> 1)];
            if (s < 0) {
                s = (int8_t)(x & 1);
                r.push_back(x);
            } else if (s != (int8_t)(x & 1)) {
                tautology = true;
                break;
            }
        }
        for (int32_t x : c) seen[(size_t)(x >> 1)] = -1;
        if (!tautology) kept.push_back(std::move(r));
    }
    f.swap(kept);
    return f.empty() ? GPSAT_SAT : GPSAT_UNDEF;
}

struct UnitFixpoint {
    ClauseList &f;
    std::vector<int32_t> solved;
    int status = GPSAT_UNDEF;
    explicit UnitFixpoint(ClauseList &formula) : f(formula) {}

    // UnaryClausesRemover::add (Preprocessing/UnaryClausesRemover.cu:46-64)
    bool add(int32_t x)
    {
        for (int32_t s : solved) {
            if (s == x) return false;
            if (s == (x ^ 1)) {
                status = GPSAT_UNSAT;
                return false;
            }
        }
        solved.push_back(x);
        return true;
    }

    // process_unary_clauses (:13-44): a unit clause leaves the formula only when its literal is NEW
    void collect_units()
    {
        ClauseList kept;
        kept.reserve(f.size());
        for (auto &c : f) {
            if (c.size() == 1 && add(c[0])) continue;
            kept.push_back(std::move(c));
        }
        f.swap(kept);
    }

    // process_clause (:66-100): SAT as soon as a solved literal is met; `unit` = last literal that is not false,
    // valid only when exactly size-1 literals are false
    int classify(const std::vector<int32_t> &c, int32_t &unit) const
    {
        size_t n_false = 0;
        unit = -1;
        for (int32_t x : c) {
            bool is_false = false;
            for (int32_t s : solved) {
                if (s == x) {
                    unit = -1;
                    return GPSAT_SAT;
                }
                if (s == (x ^ 1)) {
                    n_false++;
                    is_false = true;
                    break;
                }
            }
            if (!is_false) unit = x;
        }
        if (c.empty() || n_false != c.size() - 1) unit = -1;
        return n_false == c.size() ? GPSAT_UNSAT : GPSAT_UNDEF;
    }

    // propagate_literals (:113-150): one in-order sweep; `solved` grows while sweeping
    void sweep()
    {
        ClauseList kept;
        kept.reserve(f.size());
        size_t i = 0;
        for (; i < f.size(); i++) {
            auto &c = f[i];
            int32_t unit;
            int st = classify(c, unit);
            if (unit != -1) {
                add(unit);
                if (status == GPSAT_UNSAT) break;
            }
            if (st == GPSAT_SAT || unit != -1) continue;      // clause erased
            if (st == GPSAT_UNSAT) {
                status = GPSAT_UNSAT;
                break;
            }
            // clean_clause (:102-111): strip literals that are false under `solved`
            std::vector<int32_t> r;
            r.reserve(c.size());
            for (int32_t x : c) {
                bool is_false = false;
                for (int32_t s : solved)
                    if (s == (x ^ 1)) {
                        is_false = true;
                        break;
                    }
                if (!is_false) r.push_back(x);
            }
            kept.push_back(std::move(r));
        }
        if (status == GPSAT_UNSAT) {
            // the reference returns mid-sweep leaving the tail untouched; keep the same (unused) remainder
            for (; i < f.size(); i++) kept.push_back(std::move(f[i]));
            f.swap(kept);
            return;
        }
        f.swap(kept);
        if (f.empty()) status = GPSAT_SAT;
    }

    // process (:152-173)
    void run()
    {
        collect_units();
        if (status != GPSAT_UNDEF) return;
        size_t last = 0, cur = solved.size();
        while (cur != last) {
            sweep();
            if (status != GPSAT_UNDEF) return;
            last = cur;
            cur = solved.size();
        }
    }
};
}  // namespace

int preprocess(const gpsat_cnf &in, gpsat_cnf &out)
{
    ClauseList f((size_t)in.n_clauses());
    for (int64_t c = 0; c < in.n_clauses(); c++)
        f[(size_t)c].assign(in.lits.begin() + in.offsets[c], in.lits.begin() + in.offsets[c + 1]);

    out = gpsat_cnf();
    out.n_vars = in.n_vars;
    out.header_vars = in.header_vars;
    out.header_clauses = in.header_clauses;
    out.largest_clause = in.largest_clause;
    out.most_common_var = in.most_common_var;
    out.most_common_freq = in.most_common_freq;
    out.n_lines = in.n_lines;

    int status = drop_repeats(f, in.n_vars);
    if (status == GPSAT_UNDEF) {
        UnitFixpoint u(f);
        u.run();
        status = u.status;
        out.solved = u.solved;
    }
    out.status = status;
    for (auto &c : f) {
        out.lits.insert(out.lits.end(), c.begin(), c.end());
        out.offsets.push_back((int64_t)out.lits.size());
    }
    return GPSAT_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Cubes.  MaxClauseJobChooser::evalVarPerJobsDistribute / Uniform (JobsManager/JobChooser.cu:115-133) with
// JOBS_PER_THREAD 10, MIN_FREE_VARS 2, MAX_VARS 15, UNIFORM_NUMBER_OF_VARS 7 (SATSolver/Configs.cuh:38-41).
// The reference computes n_working_vars - MIN_FREE_VARS in size_t, so fewer than 2 live vars wraps to "huge".
// ---------------------------------------------------------------------------------------------------------------
int vars_per_job(int64_t n_working_vars, int64_t blocks, int64_t threads, int strategy)
{
    const uint64_t live_minus = (uint64_t)n_working_vars - 2u;   // wraps like the reference's size_t
    if (strategy == GPSAT_STRATEGY_UNIFORM) {
        uint64_t mx = std::max<uint64_t>(live_minus, 1);
        return (int)std::min<uint64_t>(std::min<uint64_t>(7, mx), 15);
    }
    double expected = (double)((uint64_t)threads * (uint64_t)blocks) * 10.0;
    double mx = std::pow(2.0, (double)std::min<uint64_t>(std::max<uint64_t>(live_minus, 1), 62));
    expected = std::min(expected, mx);
    uint64_t k = (uint64_t)std::log2(expected) + 1;
    return (int)std::min<uint64_t>(k, 15);
}

// VariableChooser::evaluate (JobsManager/VariableChooser.cu:23-40): score(v) = sum over occurrences of v of the
// variable INDEX (sic), sorted descending with std::sort (ties as libstdc++ leaves them: same call, same input order,
// same result as the reference build on this platform).
int choose_cube_vars(const gpsat_cnf &pre, int k, std::vector<int32_t> &vars)
{
    struct Eval {
        int32_t var;
        int32_t score;
    };
    std::vector<Eval> ev((size_t)std::max(pre.n_vars, 0));
    for (int32_t v = 0; v < pre.n_vars; v++) ev[(size_t)v] = Eval{v, 0};
    for (int32_t x : pre.lits) ev[(size_t)(x >> 1)].score += (x >> 1);
    std::sort(ev.begin(), ev.end(), [](Eval const &l, Eval const &r) { return l.score > r.score; });
    if ((size_t)k > ev.size()) {
        set_error("more cube variables requested than the formula has");
        return GPSAT_E_ARG;
    }
    vars.resize((size_t)k);
    for (int i = 0; i < k; i++) vars[(size_t)i] = ev[(size_t)i].var;
    return GPSAT_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Static device index.  Layout (all int32, DESIGN.md "data layout"):
//   cl2     int2 per slot.  Clause c owns slots cstart[c] .. cstart[c]+len: a header (len, c) followed by one
//           (literal, occurrence slot) pair per literal.  A clause reference (cref) is the slot of its first literal.
//   ostart  literal x -> first occurrence slot; occurrence slots of x are sorted by clause index
//   occ2    int2 per occurrence slot k: (first literal slot, len) of the clause in which x occurs
//   wbits0  bit k set <=> occurrence slot k is watched initially (clause positions 0 and 1)
//   vsids0  VSIDS::handle_clause over the formula in order, halving every 50 clauses
//           (SATSolver/DecisionMaker.cu:3-16, DecisionStrategy/VSIDS.cu:77-90)
//   val0    initial value per variable: UNDEF (2) if it occurs, ABSENT (4) otherwise
// ---------------------------------------------------------------------------------------------------------------
int build_device_formula(int32_t n_vars, int64_t n_clauses, const int64_t *offsets, const int32_t *lits,
                         DeviceFormula &out)
{
    if (n_vars < 0 || n_clauses < 0 || (n_clauses > 0 && (!offsets || !lits))) {
        set_error("bad formula arguments");
        return GPSAT_E_ARG;
    }
    const int64_t L = n_clauses ? offsets[n_clauses] - offsets[0] : 0;
    if (L + n_clauses >= (int64_t)1 << 30 || n_vars >= (1 << 29)) {
        set_error("formula too large for int32 slots");
        return GPSAT_E_ARG;
    }
    out = DeviceFormula();
    out.n_vars = n_vars;
    out.n_clauses = n_clauses;
    out.n_lits = L;
    out.cstart.resize((size_t)n_clauses + 1);
    out.cl2.resize((size_t)(2 * (L + n_clauses)));
    out.ostart.assign((size_t)(2 * (int64_t)n_vars + 1), 0);
    out.occ2.resize((size_t)(2 * L));
    out.wbits0.assign((size_t)((L + 31) / 32), 0u);
    out.vsids0.assign((size_t)(2 * (int64_t)n_vars), 0);
    out.val0.assign((size_t)n_vars, 4);

    const int64_t base = n_clauses ? offsets[0] : 0;
    std::vector<int8_t> seen((size_t)std::max(n_vars, 1), 0);
    for (int64_t c = 0; c < n_clauses; c++) {
        const int64_t b = offsets[c] - base, e = offsets[c + 1] - base;
        const int64_t len = e - b;
        if (len < 2) {
            set_error("clause " + std::to_string(c) + " has fewer than 2 literals: run gpsat_cnf_preprocess first");
            return GPSAT_E_ARG;
        }
        out.max_clause_len = std::max<int32_t>(out.max_clause_len, (int32_t)len);
        out.cstart[(size_t)c] = (int32_t)(b + c);
        for (int64_t i = b; i < e; i++) {
            const int32_t x = lits[base + i];
            if (x < 0 || (x >> 1) >= n_vars) {
                set_error("literal out of range in clause " + std::to_string(c));
                return GPSAT_E_ARG;
            }
            if (seen[(size_t)(x >> 1)]) {
                set_error("variable repeated inside clause " + std::to_string(c) + ": run gpsat_cnf_preprocess first");
                return GPSAT_E_ARG;
            }
            seen[(size_t)(x >> 1)] = 1;
            out.val0[(size_t)(x >> 1)] = 2;
            out.ostart[(size_t)x + 1]++;
        }
        for (int64_t i = b; i < e; i++) seen[(size_t)(lits[base + i] >> 1)] = 0;
    }
    out.cstart[(size_t)n_clauses] = (int32_t)(L + n_clauses);
    for (size_t x = 0; x < (size_t)(2 * (int64_t)n_vars); x++) out.ostart[x + 1] += out.ostart[x];

    std::vector<int32_t> fill(out.ostart.begin(), out.ostart.end() - 1);
    for (int64_t c = 0; c < n_clauses; c++) {
        const int64_t b = offsets[c] - base, e = offsets[c + 1] - base;
        const int64_t h = b + c;   // header slot
        out.cl2[(size_t)(2 * h)] = (int32_t)(e - b);
        out.cl2[(size_t)(2 * h + 1)] = (int32_t)c;
        for (int64_t i = b; i < e; i++) {
            const int32_t x = lits[base + i];
            const int32_t k = fill[(size_t)x]++;
            const int64_t slot = h + 1 + (i - b);
            out.cl2[(size_t)(2 * slot)] = x;
            out.cl2[(size_t)(2 * slot + 1)] = k;
            out.occ2[(size_t)(2 * (int64_t)k)] = (int32_t)(h + 1);
            out.occ2[(size_t)(2 * (int64_t)k + 1)] = (int32_t)(e - b);
            if (i - b < 2) out.wbits0[(size_t)(k >> 5)] |= 1u << (k & 31);
        }
    }

    // "+1 per occurrence, halve every counter every 50 clauses", with the halvings applied lazily per literal
    // (c >>= missed epochs): identical integers, O(n_lits) instead of O(n_clauses/50 * n_vars)
    {
        std::vector<int32_t> epoch_of((size_t)(2 * (int64_t)n_vars), 0);
        int32_t epoch = 0;
        auto settle = [&](size_t x) {
            const int32_t d = epoch - epoch_of[x];
            if (d > 0) {
                out.vsids0[x] = d >= 31 ? 0 : (out.vsids0[x] >> d);
                epoch_of[x] = epoch;
            }
        };
        for (int64_t c = 0; c < n_clauses; c++) {
            for (int64_t i = offsets[c]; i < offsets[c + 1]; i++) {
                settle((size_t)lits[i]);
                out.vsids0[(size_t)lits[i]]++;
            }
            if ((c + 1) % 50 == 0) epoch++;
        }
        for (size_t x = 0; x < out.vsids0.size(); x++) settle(x);
    }
    return GPSAT_OK;
}

void build_sweep_index(const DeviceFormula &D, bool want_buckets, SweepIndex &out)
{
    const size_t L = (size_t)D.n_lits;
    const size_t n_lit_ids = 2 * (size_t)std::max(D.n_vars, 0);
    out.uniform3 = (D.n_clauses > 0 && D.n_lits == 3 * D.n_clauses && D.max_clause_len == 3) ? 1 : 0;
    out.orange.assign(2 * n_lit_ids, 0);
    out.occ_clause.clear();
    out.occ_pair.clear();
    out.bucket.clear();
    out.occ_clause.reserve(L + n_lit_ids);
    if (out.uniform3) out.occ_pair.reserve(2 * (L + n_lit_ids));
    for (size_t f = 0; f < n_lit_ids; f++) {
        out.orange[2 * f] = (int32_t)out.occ_clause.size();
        for (int32_t k = D.ostart[f]; k < D.ostart[f + 1]; k++) {
            const int32_t s0 = D.occ2[2 * (size_t)k], len = D.occ2[2 * (size_t)k + 1];
            out.occ_clause.push_back(D.cl2[2 * (size_t)(s0 - 1) + 1]);
            if (out.uniform3)
                for (int i = 0; i < len; i++) {
                    if (D.cl2[2 * (size_t)(s0 + i) + 1] == k) continue;
                    out.occ_pair.push_back(D.cl2[2 * (size_t)(s0 + i)]);
                }
        }
        if (out.occ_clause.size() & 1) {
            out.occ_clause.push_back(-1);
            if (out.uniform3) {
                out.occ_pair.push_back(-1);
                out.occ_pair.push_back(-1);
            }
        }
        out.orange[2 * f + 1] = (int32_t)out.occ_clause.size();
    }
    if (out.occ_clause.empty()) out.occ_clause.push_back(-1);
    if (!want_buckets || !out.uniform3 || n_lit_ids + 2 > ((size_t)1 << kBucketLitBits)) return;
    const uint32_t pad = (uint32_t)n_lit_ids + 1u;   // 2n + 1: the sentinel's positive literal
    out.bucket.assign(16 * (n_lit_ids + 2), 0u);
    auto put = [](uint32_t *w, int bit, uint32_t v) {
        w[bit >> 5] |= v << (bit & 31);
        if ((bit & 31) + kBucketLitBits > 32) w[(bit >> 5) + 1] |= v >> (32 - (bit & 31));
    };
    for (size_t f = 0; f < n_lit_ids + 2; f++) {
        uint32_t *w = out.bucket.data() + 16 * f;
        const int32_t os = f < n_lit_ids ? out.orange[2 * f] : 0;
        const int32_t cnt = f < n_lit_ids ? D.ostart[f + 1] - D.ostart[f] : 0;
        // occurrences (8 bits, saturating) | half the (even) index of the list's first entry (24 bits, all ones: does not fit)
        const uint32_t half = (uint32_t)os >> 1;
        w[0] = (uint32_t)std::min<int32_t>(cnt, 255) | ((half < 0xFFFFFFu ? half : 0xFFFFFFu) << 8);
        for (int j = 0; j < kBucketEntries; j++) {
            const int bit = bucket_entry_bit(j);
            put(w, bit, j < cnt ? (uint32_t)out.occ_pair[2 * (size_t)(os + j)] : pad);
            put(w, bit + kBucketLitBits, j < cnt ? (uint32_t)out.occ_pair[2 * (size_t)(os + j) + 1] : pad);
        }
    }
}

void order_cubes_for_sweep(const DeviceFormula &D, int32_t n_cubes, const int64_t *cube_offsets, const int32_t *cube_lits,
                           std::vector<int32_t> &sorted, std::vector<int32_t> &info)
{
    const int64_t base = n_cubes > 0 ? cube_offsets[0] : 0;
    const int64_t total = n_cubes > 0 ? cube_offsets[n_cubes] - base : 0;
    sorted.assign((size_t)total, 0);
    info.assign((size_t)std::max(n_cubes, 0), 0);
    std::vector<int32_t> seen_in((size_t)std::max(D.n_vars, 1), -1);   // last cube each variable was seen in
    const int kClasses = 12;                                           // 0 .. 10 occurrences, 11 and more
    auto cls = [&](int32_t x) {
        const size_t f = (size_t)(x ^ 1);
        return std::min<int32_t>(D.ostart[f + 1] - D.ostart[f], kClasses - 1);
    };
    for (int32_t j = 0; j < n_cubes; j++) {
        const int64_t b = cube_offsets[j] - base, e = cube_offsets[j + 1] - base;
        int64_t at[kClasses + 1] = {0};
        bool distinct = (e - b) < ((int64_t)1 << 30);
        for (int64_t i = b; i < e; i++) {
            const int32_t x = cube_lits[base + i];
            at[cls(x) + 1]++;
            if (seen_in[(size_t)(x >> 1)] == j) distinct = false;
            seen_in[(size_t)(x >> 1)] = j;
        }
        for (int c = 0; c < kClasses; c++) at[c + 1] += at[c];
        info[(size_t)j] = (int32_t)std::min<int64_t>(at[6], ((int64_t)1 << 30) - 1) | (distinct ? 1 << 30 : 0);
        for (int64_t i = b; i < e; i++) {
            const int32_t x = cube_lits[base + i];
            sorted[(size_t)(b + at[cls(x)]++)] = x;
        }
    }
}

}  // namespace gpsat_host

// ---------------------------------------------------------------------------------------------------------------
// C ABI (host part)
// ---------------------------------------------------------------------------------------------------------------
extern "C" {

const char *gpsat_last_error(void) { return gpsat_host::last_error(); }
const char *gpsat_version(void) { return "gpupsat-b200 0.1.0 (drop-in for gpupsat v0.0.1 hot path)"; }

int gpsat_cnf_read(const char *path, gpsat_cnf **out)
{
    if (!path || !out) {
        gpsat_host::set_error("null argument");
        return GPSAT_E_ARG;
    }
    gpsat_cnf *f = new gpsat_cnf();
    int rc = gpsat_host::read_dimacs(path, *f);
    if (rc != GPSAT_OK) {
        delete f;
        *out = nullptr;
        return rc;
    }
    *out = f;
    return GPSAT_OK;
}

int gpsat_cnf_from_arrays(int64_t n_clauses, const int64_t *offsets, const int32_t *lits, gpsat_cnf **out)
{
    if (!out || n_clauses < 0 || (n_clauses > 0 && (!offsets || !lits))) {
        gpsat_host::set_error("bad arguments");
        return GPSAT_E_ARG;
    }
    gpsat_cnf *f = new gpsat_cnf();
    if (n_clauses > 0) {
        const int64_t base = offsets[0];
        for (int64_t c = 0; c < n_clauses; c++) {
            if (offsets[c + 1] < offsets[c]) {
                delete f;
                gpsat_host::set_error("offsets not monotone");
                return GPSAT_E_ARG;
            }
            for (int64_t i = offsets[c]; i < offsets[c + 1]; i++) {
                if (lits[i] < 0) {
                    delete f;
                    gpsat_host::set_error("negative literal code");
                    return GPSAT_E_ARG;
                }
                f->lits.push_back(lits[i]);
            }
            f->offsets.push_back(offsets[c + 1] - base);
        }
    }
    gpsat_host::finish_raw(*f);
    f->n_lines = (int32_t)n_clauses + 1;
    *out = f;
    return GPSAT_OK;
}

void gpsat_cnf_free(gpsat_cnf *f) { delete f; }

int gpsat_cnf_preprocess(const gpsat_cnf *in, gpsat_cnf **out)
{
    if (!in || !out) {
        gpsat_host::set_error("null argument");
        return GPSAT_E_ARG;
    }
    gpsat_cnf *f = new gpsat_cnf();
    int rc = gpsat_host::preprocess(*in, *f);
    if (rc != GPSAT_OK) {
        delete f;
        return rc;
    }
    *out = f;
    return GPSAT_OK;
}

int32_t gpsat_cnf_n_vars(const gpsat_cnf *f) { return f->n_vars; }
int64_t gpsat_cnf_n_clauses(const gpsat_cnf *f) { return f->n_clauses(); }
int64_t gpsat_cnf_n_lits(const gpsat_cnf *f) { return (int64_t)f->lits.size(); }
const int64_t *gpsat_cnf_offsets(const gpsat_cnf *f) { return f->offsets.data(); }
const int32_t *gpsat_cnf_lits(const gpsat_cnf *f) { return f->lits.data(); }
int32_t gpsat_cnf_status(const gpsat_cnf *f) { return f->status; }
int32_t gpsat_cnf_n_solved(const gpsat_cnf *f) { return (int32_t)f->solved.size(); }
const int32_t *gpsat_cnf_solved(const gpsat_cnf *f) { return f->solved.data(); }
int32_t gpsat_cnf_header_vars(const gpsat_cnf *f) { return f->header_vars; }
int64_t gpsat_cnf_header_clauses(const gpsat_cnf *f) { return f->header_clauses; }
int32_t gpsat_cnf_largest_clause(const gpsat_cnf *f) { return f->largest_clause; }
int32_t gpsat_cnf_most_common_var(const gpsat_cnf *f) { return f->most_common_var; }
int32_t gpsat_cnf_most_common_freq(const gpsat_cnf *f) { return f->most_common_freq; }
int32_t gpsat_cnf_n_lines(const gpsat_cnf *f) { return f->n_lines; }

int gpsat_choose_cubes(const gpsat_cnf *pre, int32_t blocks, int32_t threads, int32_t strategy, int32_t *vars_per_job,
                       int32_t *n_cubes, int32_t *cube_lits, int64_t cube_lits_cap)
{
    if (!pre || !vars_per_job || !n_cubes || blocks < 1 || threads < 1) {
        gpsat_host::set_error("bad arguments");
        return GPSAT_E_ARG;
    }
    const int64_t live = (int64_t)pre->n_vars - (int64_t)pre->solved.size();
    if (strategy == GPSAT_STRATEGY_SIMPLE) {
        // SimpleJobChooser::evaluate / addJobs (JobsManager/SimpleJobChooser.cu:22-75; USE_SIMPLE_JOBS_GENERATION,
        // SATSolver/Configs.cuh:44, off as shipped): the first min(live, UNIFORM_NUMBER_OF_VARS = 7) variables in index
        // order that preprocessing did not solve, positive branch first at every depth
        const int k = (int)std::min<int64_t>(std::max<int64_t>(live, 0), 7);
        *vars_per_job = k;
        *n_cubes = 1 << k;
        if (!cube_lits) return GPSAT_OK;
        if (cube_lits_cap < (int64_t)k << k) {
            gpsat_host::set_error("cube buffer too small");
            return GPSAT_E_CAPACITY;
        }
        std::vector<char> dead((size_t)std::max(pre->n_vars, 1), 0);
        for (int32_t x : pre->solved) dead[(size_t)(x >> 1)] = 1;
        std::vector<int32_t> vars;
        for (int32_t v = 0; v < pre->n_vars && (int)vars.size() < k; v++)
            if (!dead[(size_t)v]) vars.push_back(v);
        for (int64_t j = 0; j < (int64_t)1 << k; j++)
            for (int i = 0; i < k; i++)
                cube_lits[j * k + i] = 2 * vars[(size_t)i] + ((((j >> (k - 1 - i)) & 1) == 0) ? 1 : 0);
        return GPSAT_OK;
    }
    int k = gpsat_host::vars_per_job(live, blocks, threads, strategy);
    // fewer than 3 live variables: the reference's size_t arithmetic wraps (live - MIN_FREE_VARS) and asks for more cube
    // variables than exist (its main() forces the sequential path before that, SATSolver/main.cu:133-139); a cube cannot
    // hold more variables than are live
    if ((int64_t)k > std::max<int64_t>(live, 0)) k = (int)std::max<int64_t>(live, 0);
    *vars_per_job = k;
    *n_cubes = 1 << k;
    if (!cube_lits) return GPSAT_OK;
    if (cube_lits_cap < (int64_t)k << k) {
        gpsat_host::set_error("cube buffer too small");
        return GPSAT_E_CAPACITY;
    }
    std::vector<int32_t> vars;
    int rc = gpsat_host::choose_cube_vars(*pre, k, vars);
    if (rc != GPSAT_OK) return rc;
    // MaxClauseJobChooser::addJobs (JobsManager/JobChooser.cu:75-90): positive branch first at every depth
    for (int64_t j = 0; j < (int64_t)1 << k; j++)
        for (int i = 0; i < k; i++) {
            const bool positive = ((j >> (k - 1 - i)) & 1) == 0;
            cube_lits[j * k + i] = 2 * vars[(size_t)i] + (positive ? 1 : 0);
        }
    return GPSAT_OK;
}

}  // extern "C"
