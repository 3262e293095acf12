/* TEST INFRASTRUCTURE ONLY (oracle/).
 *
 * Boost-free stand-ins for the two translation units of the reference that need Boost (which this image does not
 * have): FileManager/CnfReader.cpp (Boost.Spirit DIMACS parser) and FileManager/ParametersManager.cpp
 * (Boost.Program_options).  They implement the SAME class interfaces (FileManager/CnfReader.h:13-48,
 * FileManager/ParametersManager.h:13-94) so that the UNMODIFIED rest of the reference — every .cu of
 * CMakeLists.txt:34-112 — links into a native sm_100a binary (oracle/_ref/gpupsat_ref_native) for the same-silicon
 * baseline of SURVEY.md §8(d).  Written from the headers and the observable behaviour, not copied.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include "FileManager/CnfReader.h"
#include "FileManager/ParametersManager.h"

void CnfManager::add_clause() { formula_data->add_clause(current_clause_lits, current_clause_size); }

void CnfManager::add_lit(Var v, bool sign)
{
    current_clause_lits[current_clause_size++] = mkLit(v, sign);
    if (v + 1 > max_var) max_var = v + 1;
    const int n = ++(*occurrences)[v];
    if (n > max_occurrence_of_var) {
        max_occurrence_of_var = n;
        var_with_more_occurrences = v;
    }
}

bool CnfManager::read_cnf(const char *file, FormulaData &data)
{
    std::ifstream input(file);
    if (!input) return false;
    max_var = -1;
    var_with_more_occurrences = -1;
    max_occurrence_of_var = -1;
    formula_data = &data;
    int n_vars = 0, n_clauses = 0;
    bool have_header = false;
    std::string line;
    start_new_clause();
    while (std::getline(input, line)) {
        size_t p = line.find_first_not_of(" \t\r");
        if (p == std::string::npos || line[p] == 'c' || line[p] == '%') continue;
        if (line[p] == 'p') {
            char fmt[16];
            if (std::sscanf(line.c_str() + p, "p %15s %d %d", fmt, &n_vars, &n_clauses) != 3) return false;
            have_header = true;
            continue;
        }
        std::istringstream ls(line);
        long long x;
        while (ls >> x) {
            if (x == 0) {
                if (current_clause_size > 0) add_clause();
                start_new_clause();
            } else {
                add_lit((Var)(std::llabs(x) - 1), x > 0);
            }
        }
    }
    if (!have_header) return false;
    set_header(n_vars, n_clauses);
    data.set_n_vars(max_var);
    formula_data->copy_host_clauses_to_dev();
    formula_data->set_most_common_var(var_with_more_occurrences, max_occurrence_of_var);
    return true;
}

ParametersManager::ParametersManager(int argc, char **argv)
    : correct{false}, has_help{false}, input_file{}, output_file{}, n_threads{1}, n_blocks{1}, unknown_parameter{'\0'},
      verbosity_level{0}, strategy{ChoosingStrategy::DISTRIBUTE_JOBS_PER_THREAD}, sequential_as_parallel{false},
      preprocess_unary_clauses{true}, write_log{false}
{
    process(argc, argv);
}

void ParametersManager::process(int argc, char **argv)
{
    input_file = "task.cnf";
    output_file = "solution.txt";
    n_threads = 32;
    n_blocks = 32;
    verbosity_level = 1;
    std::string strat = "distributed";
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto val = [&]() -> const char * { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "-i" || a == "--input-file") input_file = val();
        else if (a == "-o" || a == "--output-file") output_file = val();
        else if (a == "-t" || a == "--number-of-threads") n_threads = std::atoi(val());
        else if (a == "-b" || a == "--number-of-blocks") n_blocks = std::atoi(val());
        else if (a == "-v" || a == "--verbosity-level") verbosity_level = std::atoi(val());
        else if (a == "-s" || a == "--strategy") strat = val();
        else if (a == "-p" || a == "--sequential-as-parallel") sequential_as_parallel = true;
        else if (a == "-u" || a == "--preprocess-unary-clauses") sequential_as_parallel = true;   /* sic: ParametersManager.cpp:116-118 */
        else if (a == "-l" || a == "--write-log") write_log = true;
        else if (a == "--help" || a == "--version") { std::printf("gpupsat (reference, Boost-free front end)\n"); std::exit(0); }
        else if (!a.empty() && a[0] != '-') input_file = a;
    }
    std::printf("input file:\t\t\t%s\n", input_file.c_str());
    if (strat != "distributed" && strat != "uniform") {
        std::cerr << "Strategy must be either distributed or uniform!\n";
        std::exit(0);
    }
    strategy = strat == "uniform" ? ChoosingStrategy::UNIFORM : ChoosingStrategy::DISTRIBUTE_JOBS_PER_THREAD;
    correct = true;
}

void ParametersManager::force_sequential_configuration()
{
    n_blocks = 1;
    n_threads = 1;
    sequential_as_parallel = false;
}
