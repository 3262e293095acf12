// Host-side formula container, DIMACS reader, preprocessing and cube generation (pure C++, no CUDA).
// Mirrors the reference's L4/L3 host layers (FileManager/, Preprocessing/, JobsManager/) in behaviour, not in code.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

struct gpsat_cnf {
    int32_t n_vars = 0;                 // highest variable + 1
    std::vector<int64_t> offsets{0};    // n_clauses + 1
    std::vector<int32_t> lits;          // reference encoding 2*var + positive
    int32_t status = 2;                 // sat_status after preprocessing (2 = UNDEF)
    std::vector<int32_t> solved;        // literals fixed by preprocessing, discovery order
    int32_t header_vars = -1;
    int64_t header_clauses = -1;
    int32_t largest_clause = 0;         // before preprocessing
    int32_t most_common_var = -1;
    int32_t most_common_freq = -1;
    int32_t n_lines = 0;

    int64_t n_clauses() const { return (int64_t)offsets.size() - 1; }
};

namespace gpsat_host {

void set_error(const std::string &msg);
const char *last_error();

int read_dimacs(const char *path, gpsat_cnf &out);
void finish_raw(gpsat_cnf &f);   // n_vars, largest clause, most common var from the raw clauses
int preprocess(const gpsat_cnf &in, gpsat_cnf &out);
int vars_per_job(int64_t n_working_vars, int64_t blocks, int64_t threads, int strategy);
int choose_cube_vars(const gpsat_cnf &pre, int k, std::vector<int32_t> &vars);

// Static index the kernels read (built once per formula).
struct DeviceFormula {
    int32_t n_vars = 0;
    int64_t n_clauses = 0;
    int64_t n_lits = 0;
    std::vector<int32_t> cstart;     // n_clauses + 1 : clause -> header slot in cl2
    std::vector<int32_t> cl2;        // 2*(n_lits + n_clauses): per clause a (len, index) header then (literal, occurrence slot) pairs
    std::vector<int32_t> ostart;     // 2*n_vars + 1  : literal -> first occurrence slot
    std::vector<int32_t> occ2;       // 2 * n_lits    : (clause first literal slot, clause length) per occurrence slot
    std::vector<uint32_t> wbits0;    // ceil(n_lits/32): initial watch bitmap over occurrence slots (positions 0 and 1)
    std::vector<int32_t> vsids0;     // 2*n_vars      : initial per-literal VSIDS counters
    std::vector<uint8_t> val0;       // n_vars        : 2 = occurs (unassigned), 4 = absent from the formula
    int32_t max_clause_len = 0;
};
int build_device_formula(int32_t n_vars, int64_t n_clauses, const int64_t *offsets, const int32_t *lits,
                         DeviceFormula &out);

}  // namespace gpsat_host
