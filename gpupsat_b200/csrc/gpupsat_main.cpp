// gpupsat — command-line front end with the reference's surface (SATSolver/main.cu:109-329,
// FileManager/ParametersManager.cpp:31-185, SATSolver/Results.cu:62-156): same options, same stdout lines, same exit
// codes; the work behind "About to invoke kernel..." is libgpsat's sm_100a path.  Boost-free.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "../../include/gpsat.h"

namespace {

struct Options {
    std::string input_file = "task.cnf";
    std::string output_file = "solution.txt";
    int n_threads = 32, n_blocks = 32, verbosity = 1;     // ParametersManager.cpp:36-44
    std::string strategy = "distributed";
    bool sequential_as_parallel = false, write_log = false;
    bool preprocess_unary = true;
    int n_gpus = 1;   // extension: GPUs of this box to solve on (0 = all visible); the reference drives one
};

const char *kHelp =
    "runsat: ./runsat <file> [options]:\n"
    "  --help                             Displays this help message\n"
    "  --version                          Displays version number\n"
    "  -i [ --input-file ] arg            Input file\n"
    "  -o [ --output-file ] arg           Output file\n"
    "  -t [ --number-of-threads ] arg     Number of threads\n"
    "  -b [ --number-of-blocks ] arg      Number of blocks\n"
    "  -v [ --verbosity-level ] arg       Set verbosity level\n"
    "  -s [ --strategy ] arg              Strategy for generating the jobs: distributed or \n"
    "                                     uniform\n"
    "  -p [ --sequential-as-parallel ]    Forces the execution of the parallel strategy, \n"
    "                                     with jobs creation, but running with 1 thread \n"
    "                                     and 1 block\n"
    "  -u [ --preprocess-unary-clauses ]  Turns on the pre-processing of unary clauses\n"
    "  -l [ --write-log ]                 Prints file,threads,blocks,ms to autolog.txt\n"
    "  -g [ --gpus ] arg                  (gpupsat_b200) GPUs of this box to use, 0 = all \n"
    "                                     visible; default 1 like the reference\n";

[[noreturn]] void bad_option(const std::string &msg)
{
    std::fprintf(stderr, "%s\n", msg.c_str());
    std::exit(0);   // ParametersManager.cpp:69-72: po::error -> message on stderr, exit(0)
}

Options parse(int argc, char **argv)
{
    Options o;
    bool any = false;
    auto need_value = [&](int &i, const std::string &name, const char *inline_val) -> std::string {
        if (inline_val) return inline_val;
        if (i + 1 >= argc) bad_option("the required argument for option '--" + name + "' is missing");
        return argv[++i];
    };
    auto to_int = [&](const std::string &v, const std::string &name) -> int {
        char *end = nullptr;
        long r = std::strtol(v.c_str(), &end, 10);
        if (end == v.c_str() || *end) bad_option("the argument ('" + v + "') for option '--" + name + "' is invalid");
        return (int)r;
    };
    struct Spec { const char *lng; char shrt; bool has_value; };
    static const Spec specs[] = {{"help", 0, false}, {"version", 0, false}, {"input-file", 'i', true},
                                 {"output-file", 'o', true}, {"number-of-threads", 't', true},
                                 {"number-of-blocks", 'b', true}, {"verbosity-level", 'v', true},
                                 {"strategy", 's', true}, {"sequential-as-parallel", 'p', false},
                                 {"preprocess-unary-clauses", 'u', false}, {"write-log", 'l', false},
                                 {"gpus", 'g', true}};
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        const Spec *sp = nullptr;
        std::string inline_store;
        const char *inline_val = nullptr;
        if (a.rfind("--", 0) == 0) {
            std::string name = a.substr(2);
            size_t eq = name.find('=');
            if (eq != std::string::npos) {
                inline_store = name.substr(eq + 1);
                inline_val = inline_store.c_str();
                name = name.substr(0, eq);
            }
            for (auto &s : specs)
                if (name == s.lng) sp = &s;
            if (!sp) bad_option("unrecognised option '" + a + "'");
        } else if (a.size() >= 2 && a[0] == '-' && !(a[1] >= '0' && a[1] <= '9')) {
            for (auto &s : specs)
                if (s.shrt && a[1] == s.shrt) sp = &s;
            if (!sp) bad_option("unrecognised option '" + a + "'");
            if (a.size() > 2) {
                inline_store = a.substr(2);
                inline_val = inline_store.c_str();
            }
        } else {
            o.input_file = a;   // positional = input file
            any = true;
            continue;
        }
        any = true;
        std::string name = sp->lng;
        if (name == "help") {
            std::fputs(kHelp, stdout);
            std::exit(0);
        } else if (name == "version") {
            std::puts("v0.0.1");   // FileManager/Version.h:3
            std::exit(0);
        }
        std::string v = sp->has_value ? need_value(i, name, inline_val) : std::string();
        if (name == "input-file") o.input_file = v;
        else if (name == "output-file") o.output_file = v;
        else if (name == "number-of-threads") o.n_threads = to_int(v, name);
        else if (name == "number-of-blocks") o.n_blocks = to_int(v, name);
        else if (name == "verbosity-level") o.verbosity = to_int(v, name);
        else if (name == "strategy") o.strategy = v;
        else if (name == "sequential-as-parallel") o.sequential_as_parallel = true;
        else if (name == "preprocess-unary-clauses") o.sequential_as_parallel = true;   // sic: ParametersManager.cpp:116-118
        else if (name == "write-log") o.write_log = true;
        else if (name == "gpus") o.n_gpus = to_int(v, name);
    }
    if (!any) {   // vars.size() == 0 -> help, exit(0)
        std::fputs(kHelp, stdout);
        std::exit(0);
    }
    if (o.input_file.empty()) std::fputs(kHelp, stdout);
    std::printf("input file:\t\t\t%s\n", o.input_file.c_str());
    if (o.strategy != "distributed" && o.strategy != "uniform" && o.strategy != "simple") {   // "simple" = SimpleJobChooser
        std::fprintf(stderr, "Strategy must be either distributed or uniform!\n");
        std::exit(0);
    }
    return o;
}

// Results::print_results / print_sat_results (SATSolver/Results.cu:62-156), with a SOUND verification against the
// clauses the solver saw (the reference's check ignores literal signs: Results.cu:143-155)
// Statistics surface (≙ RuntimeStatistics::print_function_time_statistics, Statistics/RuntimeStatistics.cu:299-360 — the
// call is commented out in the reference's main(), SATSolver/main.cu:301; printed here from -v 2 on).  Same section titles
// in the same order; one line per section instead of one per (block, thread): the kernel keeps one total per GPU.
void print_phase(const char *title, int64_t ns, int64_t count)
{
    std::printf("%s\n", title);
    if (count <= 0) {
        std::printf("Average:\n\tAll warps: not run\nTotal:\n\tAll warps: not run\n");
        return;
    }
    std::printf("Average:\n\tAll warps: %15.3f\n", (double)ns / (double)count);
    std::printf("Total:\n\tAll warps: %lld, run %lld times\n", (long long)ns, (long long)count);
}

void print_statistics(const gpsat_phase_stats &p)
{
    std::printf("\n*****Statistics******\n");
    print_phase("Total job's time:", p.job_ns, p.jobs);
    print_phase("Pre-processing time:", p.ns[GPSAT_PHASE_RESET] + p.ns[GPSAT_PHASE_IMPORT], p.count[GPSAT_PHASE_RESET]);
    print_phase("Decision time:", p.ns[GPSAT_PHASE_DECIDE], p.count[GPSAT_PHASE_DECIDE]);
    print_phase("Conflict analyzing time:", p.ns[GPSAT_PHASE_PROPAGATE] + p.ns[GPSAT_PHASE_ANALYZE], p.count[GPSAT_PHASE_PROPAGATE]);
    print_phase("Backtracking time:", p.ns[GPSAT_PHASE_BACKTRACK], p.count[GPSAT_PHASE_BACKTRACK]);
    print_phase("Structures reset time:", p.ns[GPSAT_PHASE_RESET], p.count[GPSAT_PHASE_RESET]);
    print_phase("Creating structures time:", 0, 0);
    print_phase("Next job time:", p.idle_ns, p.jobs);
    print_phase("Add jobs to assumptions time:", 0, 0);
    print_phase("Processing results time:", 0, 0);
    std::printf("Average backtracked levels:\nAverage:\n");
    if (p.count[GPSAT_PHASE_BACKTRACK] > 0)
        std::printf("\tAll warps: %15.3f\n", (double)p.backtracked_levels / (double)p.count[GPSAT_PHASE_BACKTRACK]);
    else std::printf("\tAll warps: not run\n");
    print_phase("Pre-processing - handling assumptions time:", 0, 0);
    print_phase("Pre-processing - adding assumptions to graph time:", 0, 0);
    print_phase("Pre-processing - adding handling vars time:", 0, 0);
    // what the reference has no phase for (gpupsat_b200 only)
    print_phase("BCP (two watched literals) time:", p.ns[GPSAT_PHASE_PROPAGATE], p.count[GPSAT_PHASE_PROPAGATE]);
    print_phase("First-UIP analysis time:", p.ns[GPSAT_PHASE_ANALYZE], p.count[GPSAT_PHASE_ANALYZE]);
    print_phase("Cube splitting time:", p.ns[GPSAT_PHASE_SPLIT], p.count[GPSAT_PHASE_SPLIT]);
    print_phase("Split-off cube import time:", p.ns[GPSAT_PHASE_IMPORT], p.count[GPSAT_PHASE_IMPORT]);
    print_phase("Learnt database reduction time:", p.ns[GPSAT_PHASE_REDUCE], p.count[GPSAT_PHASE_REDUCE]);
}

void print_results(int verdict, int n_vars, const std::vector<uint8_t> &model, const gpsat_cnf *pre,
                   const gpsat_cnf *raw, const std::string &out_path)
{
    if (verdict == GPSAT_UNDEF) std::printf("UNDEFINED\n");
    else if (verdict == GPSAT_UNSAT) std::printf("UNSATISFIABLE\n");
    if (verdict != GPSAT_SAT) return;
    std::printf("SATISFIABLE\n");
    if (n_vars < 0) n_vars = 0;                  // a formula without clauses has no variables
    std::vector<int> value((size_t)n_vars, 1);   // unassigned variables print positive (Results.cu:130-134)
    for (int v = 0; v < n_vars && v < (int)model.size(); v++) value[(size_t)v] = model[(size_t)v] ? 1 : 0;
    const int32_t *solved = gpsat_cnf_solved(pre);
    for (int i = 0; i < gpsat_cnf_n_solved(pre); i++) value[(size_t)(solved[i] >> 1)] = solved[i] & 1;
    std::string line;
    for (int v = 0; v < n_vars; v++) line += (value[(size_t)v] ? "" : "-") + std::to_string(v + 1) + " ";
    line += "0\n";
    std::fputs(line.c_str(), stdout);
    bool ok = true;
    const int64_t *off = gpsat_cnf_offsets(raw);
    const int32_t *lits = gpsat_cnf_lits(raw);
    for (int64_t c = 0; c < gpsat_cnf_n_clauses(raw) && ok; c++) {
        bool sat = false;
        for (int64_t i = off[c]; i < off[c + 1]; i++)
            if (value[(size_t)(lits[i] >> 1)] == (lits[i] & 1)) sat = true;
        ok = sat;
    }
    std::printf("Solution %s", ok ? "was verified\n" : "was not verified\n");
    if (!out_path.empty() && out_path != "solution.txt") {   // -o is parsed but unused by the reference; honour it when given
        std::ofstream out(out_path);
        out << "s SATISFIABLE\nv " << line;
    }
}

}  // namespace

int main(int argc, char **argv)
{
    Options pm = parse(argc, argv);
    {
        FILE *f = std::fopen(pm.input_file.c_str(), "r");
        if (!f) {
            std::printf("The specified CNF file (%s) was not found!\n", pm.input_file.c_str());
            std::exit(-1);
        }
        std::fclose(f);
    }
    gpsat_cnf *raw = nullptr, *pre = nullptr;
    if (gpsat_cnf_read(pm.input_file.c_str(), &raw) != GPSAT_OK) {
        std::printf("Error parsing inputs.\n");
        std::exit(-1);
    }
    const int n_vars = std::max(gpsat_cnf_n_vars(raw), 0);   // "p cnf 0 0": FormulaData leaves n_vars at -1
    if (n_vars > gpsat_cnf_header_vars(raw))
        std::printf("header claims %d vars, but highest var found is %d. Using %d...", gpsat_cnf_header_vars(raw), n_vars, n_vars);
    if (gpsat_cnf_n_clauses(raw) != gpsat_cnf_header_clauses(raw))
        std::printf("header claims %d clauses, but found %d clauses. Using %d...", (int)gpsat_cnf_header_clauses(raw),
                    (int)gpsat_cnf_n_clauses(raw), (int)gpsat_cnf_n_clauses(raw));
    if (gpsat_cnf_preprocess(raw, &pre) != GPSAT_OK) {
        std::printf("Error parsing inputs.\n");
        std::exit(-1);
    }
    int n_threads = pm.n_threads, n_blocks = pm.n_blocks;
    bool seq_as_par = pm.sequential_as_parallel;
    if (n_vars < 3 && (n_blocks > 1 || n_threads > 1 || seq_as_par)) {   // main.cu:133-139
        std::printf("Warning: There are %d vars in the formula and at least %d are necessary to parallelize. "
                    "Forcing sequential execution!\n", n_vars, 3);
        n_blocks = n_threads = 1;
        seq_as_par = false;
    }
    const int n_clauses = (int)gpsat_cnf_n_clauses(pre);
    const int max_impl = std::max(gpsat_cnf_largest_clause(raw), 100);
    if (n_blocks < 1) std::printf("Invalid number of blocks: %d\n", n_blocks);
    if (n_threads < 1) {
        std::printf("Invalid number of threads: %d\n", n_threads);
        std::exit(1);
    }
    const bool sequential = n_threads == 1 && n_blocks == 1 && !seq_as_par;
    if (pm.verbosity >= 1) {   // print_info, main.cu:31-82
        std::printf("Solver configuration:\n");
        std::printf("Input file: %s\n", pm.input_file.c_str());
        std::printf("Formula has %d vars and %d clauses\n", n_vars, n_clauses);
        std::printf("Variable '%d', the most frequent, has been found %d times.\n", gpsat_cnf_most_common_var(raw) + 1,
                    gpsat_cnf_most_common_freq(raw));
        if (n_threads == 1 && !seq_as_par) {
            std::printf("Parallelization strategy: SEQUENTIAL RUN\n");
        } else {
            std::printf("Parallelization strategy: Divide and Conquer\n");
            std::printf("Number of blocks: %d\n", n_blocks);
            std::printf("Number of threads: %d\n", n_threads);
            std::printf("Job creation strategy: %s\n", pm.strategy == "uniform" ? "uniform" : pm.strategy == "simple" ? "simple" : "distribution per thread");
        }
        std::printf("Conflict analysis: ON without forward edges\n");
        std::printf("Capacity of edges = %d\n", max_impl);
        std::printf("Assumptions are stored in a statically allocated vector.\n");
        std::printf("Formula clauses are stored in several allocations.\n");
        std::printf("Unary clauses pre-processing is %s\n", "ON");
        std::printf("Conflict analysis is two wached literals\n");
    }
    std::printf("VSIDS is ON\n");
    std::printf("Restart is ON\n");
    std::printf("Clause learning is ON with learnt clause capacity of %d\n", 16384);
    std::printf("Simple jobs generation is %s\n", pm.strategy == "simple" ? "ON" : "OFF");

    std::vector<uint8_t> model((size_t)std::max(n_vars, 1), 1);
    if (gpsat_cnf_status(pre) != GPSAT_UNDEF) {   // main.cu:154-163
        std::printf("Solved in pre-processing.\n");
        print_results(gpsat_cnf_status(pre), n_vars, model, pre, raw, pm.output_file);
        return 0;
    }

    {   // an empty clause survives the reference's host preprocessing (status stays UNDEF) and its device code has no
        // defined behaviour for it; two watched literals need two literals: the formula is unsatisfiable, say so
        const int64_t *off = gpsat_cnf_offsets(pre);
        for (int64_t c = 0; c < gpsat_cnf_n_clauses(pre); c++)
            if (off[c + 1] == off[c]) {
                std::printf("Solved in pre-processing.\n");
                print_results(GPSAT_UNSAT, n_vars, model, pre, raw, pm.output_file);
                return 0;
            }
    }
    gpsat_opts opts;
    gpsat_opts_default(&opts);
    if (const char *e = std::getenv("GPSAT_SHARE_LEARNTS")) opts.share_learnts = std::atoi(e);
    if (const char *e = std::getenv("GPSAT_DECISION")) opts.decision = std::atoi(e);
    opts.phase_stats = pm.verbosity >= 2 ? 1 : 0;
    // one GPU: a plain handle (the reference's shape); several: the multi-GPU host (one host thread per GPU, the GPUs
    // meshed over NVLink peer memory, results reduced with NCCL)
    int n_gpus = pm.n_gpus;
    if (const char *e = std::getenv("GPSAT_GPUS")) n_gpus = std::atoi(e);
    gpsat_t *h = nullptr;
    gpsat_multi_t *mh = nullptr;
    int rc;
    if (n_gpus == 1) {
        rc = gpsat_create(&h, n_vars, gpsat_cnf_n_clauses(pre), gpsat_cnf_offsets(pre), gpsat_cnf_lits(pre), &opts);
    } else {
        opts.phase_stats = 0;
        rc = gpsat_multi_create(&mh, n_gpus, nullptr, n_vars, gpsat_cnf_n_clauses(pre), gpsat_cnf_offsets(pre),
                                gpsat_cnf_lits(pre), &opts);
        if (rc == GPSAT_OK && pm.verbosity >= 1) std::printf("Number of GPUs: %d\n", gpsat_multi_n_gpus(mh));
    }
    if (rc != GPSAT_OK) {   // CudaMemoryErrorHandler.cu:3-10: message, exit(1)
        std::printf("Error on %s, description: %s\n", "creating the solver", gpsat_last_error());
        std::exit(1);
    }
    int32_t verdict = GPSAT_UNDEF, backend = 0;
    gpsat_stats st;
    std::memset(&st, 0, sizeof(st));
    auto set_cubes = [&](int32_t n, const int64_t *off, const int32_t *lits) {
        return mh ? gpsat_multi_set_cubes(mh, n, off, lits) : gpsat_set_cubes(h, n, off, lits);
    };
    auto solve = [&]() {
        return mh ? gpsat_multi_solve(mh, &verdict, model.data(), &st, &backend) : gpsat_solve(h, &verdict, model.data(), &st);
    };
    if (sequential) {
        std::printf("About to call sequential kernel!\n");
        rc = set_cubes(0, nullptr, nullptr);
        if (rc == GPSAT_OK) rc = solve();
    } else {
        int32_t k = 0, n_jobs = 0;
        const int strategy = pm.strategy == "uniform" ? GPSAT_STRATEGY_UNIFORM
                           : pm.strategy == "simple" ? GPSAT_STRATEGY_SIMPLE : GPSAT_STRATEGY_DISTRIBUTED;
        rc = gpsat_choose_cubes(pre, n_blocks, n_threads, strategy, &k, &n_jobs, nullptr, 0);
        std::vector<int32_t> cube_lits((size_t)n_jobs * (size_t)k + 1);
        std::vector<int64_t> cube_off((size_t)n_jobs + 1);
        if (rc == GPSAT_OK)
            rc = gpsat_choose_cubes(pre, n_blocks, n_threads, strategy, &k, &n_jobs, cube_lits.data(), (int64_t)cube_lits.size());
        for (int j = 0; j <= n_jobs; j++) cube_off[(size_t)j] = (int64_t)j * k;
        std::printf("Number of jobs = %d\n", n_jobs);
        std::printf("About to invoke kernel...\n");
        if (rc == GPSAT_OK) rc = set_cubes(n_jobs, cube_off.data(), cube_lits.data());
        if (rc == GPSAT_OK) rc = solve();
        std::printf("Kernel was invoked %zu times\n", (size_t)st.kernel_launches);
        std::printf("Jobs size = %d\n", n_jobs);
        std::printf("There were %d jobs created.\nThere were %d solved jobs\n", n_jobs, (int)st.jobs_done);
    }
    if (rc != GPSAT_OK) {
        std::printf("Error on %s, description: %s\n", "solving", gpsat_last_error());
        std::exit(1);
    }
    std::printf("Total time on GPU: %f ms\n", st.kernel_ms);
    if (pm.verbosity >= 2)
        std::printf("c jobs %lld decisions %lld implications %lld conflicts %lld learnt %lld restarts %lld watchers %lld "
                    "(%d blocks x %d warps, state in %s)\n",
                    (long long)st.jobs_done, (long long)st.decisions, (long long)st.implications, (long long)st.conflicts,
                    (long long)st.learnt_clauses, (long long)st.restarts, (long long)st.watchers_visited, st.blocks,
                    st.warps_per_block, st.state_in_smem ? "shared memory" : "global memory");
    if (pm.verbosity >= 2 && h) {
        gpsat_phase_stats ps;
        if (gpsat_get_phase_stats(h, &ps) == GPSAT_OK) print_statistics(ps);
    }
    if (pm.verbosity >= 2 && mh)
        std::printf("c %d GPUs, %lld cubes taken over NVLink, results reduced with %s\n", gpsat_multi_n_gpus(mh),
                    (long long)st.steals, backend ? "ncclAllReduce" : "the host");
    print_results(verdict, n_vars, model, pre, raw, pm.output_file);
    if (pm.write_log) {   // main.cu:316-321
        char buf[512];
        std::snprintf(buf, sizeof(buf), "%s,%d,%d,%f\n", pm.input_file.c_str(), n_threads, n_blocks, st.kernel_ms);
        std::ofstream out("autolog.txt", std::ios_base::app);
        out << buf;
    }
    gpsat_destroy(h);
    gpsat_multi_destroy(mh);
    gpsat_cnf_free(pre);
    gpsat_cnf_free(raw);
    return 0;
}
