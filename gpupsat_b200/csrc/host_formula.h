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

// Occurrence index of the large-database sweep kernels (GPSAT_BCP_OCCURRENCE), built from a DeviceFormula.
//   orange / occ_clause / occ_pair  PADDED occurrence lists: every literal's list starts at an even entry index and is
//       padded to an even length with -1, so a kernel reads it as 16-byte loads of two entries and finds (begin, end)
//       with ONE 8-byte load.  Per entry the clause index and, for pure 3-SAT, the two other literals of the clause.
//   bucket  (pure 3-SAT whose literal ids, with those of a sentinel variable n, fit 21 bits) the list packed into its
//       head, one 64-byte bucket (16 words) per literal id 0 .. 2n+1: word 0 = min(occurrence count, 255) | (index of the list in occ_pair / 2) << 8
//       (0xFFFFFF when that does not fit), entries 0..4 from
//       bit 32 and 5..10 from bit 256 at 42 bits each (other literal a in the low 21 bits, b above it); unused
//       entries hold the sentinel's positive literal 2n+1 twice.
struct SweepIndex {
    int32_t uniform3 = 0;
    std::vector<int32_t> orange;       // 2 per literal: (begin, end) entry index
    std::vector<int32_t> occ_clause;   // per entry
    std::vector<int32_t> occ_pair;     // 2 per entry (uniform3 only)
    std::vector<uint32_t> bucket;      // 16 per literal id, 2n + 2 of them (empty when the database does not qualify)
};
enum { kBucketEntries = 11, kBucketLitBits = 21 };
void build_sweep_index(const DeviceFormula &D, bool want_buckets, SweepIndex &out);
// bit offset of entry j of a bucket
static inline int bucket_entry_bit(int j) { return j < 5 ? 32 + 42 * j : 256 + 42 * (j - 5); }

// Order every cube's literals by the occurrence count of their NEGATION (stable counting sort, classes 0..10 and
// "11 or more"): the ternary sweep kernel scans the buckets of 32 trail literals in lock step and stops at the longest
// list among them.  info[j]: bits 0..29 = literals whose list has at most 5 entries (one 32-byte sector each),
// bit 30 = no variable occurs twice in cube j.
void order_cubes_for_sweep(const DeviceFormula &D, int32_t n_cubes, const int64_t *cube_offsets, const int32_t *cube_lits,
                           std::vector<int32_t> &sorted, std::vector<int32_t> &info);

}  // namespace gpsat_host
