/* TEST INFRASTRUCTURE ONLY (oracle/).
 *
 * C API over the UNMODIFIED reference solver classes (nvzoll/gpupsat, /root/reference/src), compiled as
 * host C++ through oracle/ref/stub/ by oracle/ref/build_ref.sh into oracle/_ref/libgpsat_ref.so.
 * It exists to (1) pin oracle/gpsat_oracle.cpp (the restatement) and the CUDA path against the real
 * reference and (2) be the "reference on host cores" CPU baseline of bench.py.  The product never links it.
 *
 * What each entry drives (reference file:line):
 *   ref_open       FormulaData::add_clause / set_n_vars / copy_host_clauses_to_dev  (FileManager/FormulaData.cu:19-38,82-105)
 *                  exactly as CnfManager::read_cnf does (FileManager/CnfReader.cpp:86-135) minus the Boost parser
 *   ref_cubes      MaxClauseJobChooser::evaluate/getJobs (JobsManager/JobChooser.cu:52-90) through a JobsQueue
 *   ref_cubes_simple  SimpleJobChooser::evaluate/getJobs (JobsManager/SimpleJobChooser.cu:22-75), the generator behind
 *                  USE_SIMPLE_JOBS_GENERATION (SATSolver/Configs.cuh:44, off as shipped)
 *   ref_clause_status  VariablesStateHandler::clause_status (SATSolver/VariablesStateHandler.cu:180-206) over the whole
 *                  formula under an arbitrary partial assignment
 *   ref_propagate  VariablesStateHandler::set_assumptions + ConflictAnalyzerWithWatchedLits::set_assumptions
 *                  = SATSolver::preprocess (SATSolver/SATSolver.cu:231-246), then the KernelContext::finished reset
 *                  (SATSolver/Parallelizer.cu:60-74)
 *   ref_solve      SATSolver::solve(GPUStaticVec<Lit>*) (SATSolver/SATSolver.cu:67-218) as KernelContext / run_sequential
 *                  build it (SATSolver/Parallelizer.cu:27-47,230-278); NOT solve(): see SURVEY.md §8(c) "known trap"
 */
#include <vector>
#include <algorithm>
#include <cstdint>
#include "SATSolver/Configs.cuh"
#include "SATSolver/SolverTypes.cuh"
#include "FileManager/FormulaData.cuh"
#include "JobsManager/JobChooser.cuh"
#include "JobsManager/SimpleJobChooser.cuh"
#include "SATSolver/JobsQueue.cuh"
#include "SATSolver/SATSolver.cuh"
#include "Statistics/RuntimeStatistics.cuh"

/* counters fed by the --wrap'd VariablesStateHandler::new_implication / new_decision (build_ref.sh) */
extern "C" {
long long gpsat_ref_n_implications = 0;
long long gpsat_ref_n_decisions = 0;
}
extern "C" void __real__ZN21VariablesStateHandler15new_implicationE8Decision(VariablesStateHandler *, Decision);
extern "C" void __wrap__ZN21VariablesStateHandler15new_implicationE8Decision(VariablesStateHandler *self, Decision d)
{
    gpsat_ref_n_implications++;
    __real__ZN21VariablesStateHandler15new_implicationE8Decision(self, d);
}
extern "C" void __real__ZN21VariablesStateHandler12new_decisionE8Decision(VariablesStateHandler *, Decision);
extern "C" void __wrap__ZN21VariablesStateHandler12new_decisionE8Decision(VariablesStateHandler *self, Decision d)
{
    gpsat_ref_n_decisions++;
    __real__ZN21VariablesStateHandler12new_decisionE8Decision(self, d);
}

namespace {

struct RefHandle {
    FormulaData *fdata = nullptr;
    int n_vars = 0;
    int max_impl = 0;
    std::vector<Var> dead_host;
    Var *dead_buf = nullptr;
    GPUVecView<Var> dead_view;
    CUDAClauseVec *formula_dev = nullptr;   /* persistent copy: solver classes keep a pointer to it */
    RuntimeStatistics *stats = nullptr;
    watched_clause_node_t *repo = nullptr;
    /* lazily built solver (ref_solve) and propagator (ref_propagate) */
    Var *fv = nullptr; Decision *dec = nullptr; Decision *imp = nullptr;
    SATSolver *solver = nullptr;
    Var *pfv = nullptr; Decision *pdec = nullptr; Decision *pimp = nullptr;
    DecisionMaker *pdm = nullptr;
    VariablesStateHandler *pvh = nullptr;
    ConflictAnalyzerWithWatchedLits *pca = nullptr;
    GPUStaticVec<Lit> cube;
    GPUStaticVec<Lit> pcube;
};

void ensure_common(RefHandle *h)
{
    if (h->stats) return;
    h->stats = new RuntimeStatistics(1, 1, nullptr);
    h->repo = new watched_clause_node_t(MAX_NUMBER_OF_NODES);
}

} // namespace

extern "C" {

/* lits use the reference encoding x = 2*var + (positive ? 1 : 0); capacity = the file's line count (main.cu:120-125). */
void *ref_open(int64_t n_clauses, const int64_t *offsets, const int32_t *lits, int capacity)
{
    RefHandle *h = new RefHandle();
    int cap = capacity > (int)n_clauses ? capacity : (int)n_clauses;
    h->fdata = new FormulaData(cap > 0 ? cap : 1, true);
    int max_var = -1;
    std::vector<Lit> tmp;
    for (int64_t c = 0; c < n_clauses; c++) {
        tmp.clear();
        for (int64_t k = offsets[c]; k < offsets[c + 1]; k++) {
            Lit l; l.x = lits[k];
            tmp.push_back(l);
            if (var(l) + 1 > max_var) max_var = var(l) + 1;
        }
        h->fdata->add_clause(tmp.data(), (int)tmp.size());
    }
    h->fdata->set_n_vars(max_var);
    h->fdata->copy_host_clauses_to_dev();
    h->n_vars = h->fdata->get_n_vars();
    h->max_impl = std::max(h->fdata->get_largest_clause_size(), MIN_IMPLICATION_PER_VAR);   /* main.cu:150 */
    for (Lit l : h->fdata->get_solved_literals()) h->dead_host.push_back(var(l));
    h->dead_buf = (Var *)malloc(sizeof(Var) * (h->dead_host.size() + 1));
    for (size_t i = 0; i < h->dead_host.size(); i++) h->dead_buf[i] = h->dead_host[i];
    h->dead_view = GPUVecView<Var>(h->dead_buf, h->dead_host.size(), h->dead_host.size());
    h->formula_dev = new CUDAClauseVec(h->fdata->get_formula_dev());
    return h;
}

int ref_n_vars(void *hv) { return ((RefHandle *)hv)->n_vars; }
int ref_status_after_preprocessing(void *hv) { return (int)((RefHandle *)hv)->fdata->get_status_after_preprocessing(); }
int ref_n_clauses(void *hv) { return (int)((RefHandle *)hv)->fdata->get_formula_host()->size(); }
int ref_n_solved_literals(void *hv) { return (int)((RefHandle *)hv)->fdata->get_solved_literals().size(); }
void ref_get_solved_literals(void *hv, int32_t *out)
{
    int i = 0;
    for (Lit l : ((RefHandle *)hv)->fdata->get_solved_literals()) out[i++] = l.x;
}
int64_t ref_n_literals(void *hv)
{
    int64_t n = 0;
    for (Clause const &c : *((RefHandle *)hv)->fdata->get_formula_host()) n += c.n_lits;
    return n;
}
/* the formula after the reference's host preprocessing, as CSR */
void ref_get_formula(void *hv, int64_t *offsets, int32_t *lits)
{
    int64_t k = 0; int64_t ci = 0;
    offsets[0] = 0;
    for (Clause const &c : *((RefHandle *)hv)->fdata->get_formula_host()) {
        for (unsigned i = 0; i < c.n_lits; i++) lits[k++] = c.literals[i].x;
        offsets[++ci] = k;
    }
}

/* strategy: 0 = distributed, 1 = uniform.  Returns n_jobs, writes vars_per_job; if out != NULL fills n_jobs*k lits. */
int ref_cubes(void *hv, int blocks, int threads, int strategy, int *vars_per_job, int32_t *out, int64_t out_cap)
{
    RefHandle *h = (RefHandle *)hv;
    MaxClauseJobChooser chooser(*h->fdata->get_formula_host(), (size_t)h->n_vars, h->dead_host.size(),
                                (size_t)threads, (size_t)blocks,
                                strategy ? ChoosingStrategy::UNIFORM : ChoosingStrategy::DISTRIBUTE_JOBS_PER_THREAD);
    chooser.evaluate();
    int n_jobs = (int)chooser.get_n_jobs();
    unsigned counter = 0;
    JobsQueue queue((size_t)n_jobs, &counter);
    chooser.getJobs(queue);
    queue.close();
    int k = (int)queue.largest_job_size();
    *vars_per_job = k;
    if (out) {
        for (int j = 0; j < n_jobs; j++) {
            Job job = queue.next_job();
            for (size_t i = 0; i < job.n_literals; i++) {
                int64_t pos = (int64_t)j * k + (int64_t)i;
                if (pos < out_cap) out[pos] = job.literals[i].x;
            }
            free(job.literals);
        }
    }
    return n_jobs;
}


/* SimpleJobChooser: the first min(live vars, UNIFORM_NUMBER_OF_VARS) live variables, positive branch first. */
int ref_cubes_simple(void *hv, int *vars_per_job, int32_t *out, int64_t out_cap)
{
    RefHandle *h = (RefHandle *)hv;
    SimpleJobChooser chooser((size_t)h->n_vars, h->dead_host);
    chooser.evaluate();
    int n_jobs = (int)chooser.get_n_jobs();
    unsigned counter = 0;
    JobsQueue queue((size_t)n_jobs, &counter);
    chooser.getJobs(queue);
    queue.close();
    int k = (int)queue.largest_job_size();
    *vars_per_job = k;
    if (out) {
        for (int j = 0; j < n_jobs; j++) {
            Job job = queue.next_job();
            for (size_t i = 0; i < job.n_literals; i++) {
                int64_t pos = (int64_t)j * k + (int64_t)i;
                if (pos < out_cap) out[pos] = job.literals[i].x;
            }
            free(job.literals);
        }
    }
    return n_jobs;
}

/* clause_status of every clause under `assignment` (per variable: 0 = true, 1 = false, 2 = unassigned — the reference's
 * sat_status).  status[c] in {0 SAT, 1 UNSAT, 2 UNDEF}; unit[c] = the single unassigned literal of a unit clause, else -1. */
int ref_clause_status(void *hv, const uint8_t *assignment, int32_t *status, int32_t *unit)
{
    RefHandle *h = (RefHandle *)hv;
    ensure_common(h);
    Var *fv = (Var *)malloc(sizeof(Var) * (h->n_vars + 1));
    Decision *dec = (Decision *)malloc(sizeof(Decision) * (h->n_vars + 1));
    Decision *imp = (Decision *)malloc(sizeof(Decision) * (h->n_vars + 1));
    DecisionMaker *dm = new DecisionMaker(h->formula_dev, (size_t)h->n_vars);
    VariablesStateHandler *vh = new VariablesStateHandler(h->n_vars, &h->dead_view, dm, fv, dec, imp);
    dm->set_vars_handler(vh);
    for (int v = 0; v < h->n_vars; v++) {
        if (assignment[v] > 1) continue;
        bool dead = false;
        for (Var d : h->dead_host) dead = dead || d == v;
        if (dead) continue;
        Decision d;
        d.literal = mkLit(v, assignment[v] == 0);
        d.decision_level = 1;
        d.implicated_from_formula = false;
        vh->new_implication(d);
    }
    int m = h->formula_dev->size_of();
    for (int c = 0; c < m; c++) {
        Lit l;
        l.x = -1;
        sat_status st = vh->clause_status(h->formula_dev->get((size_t)c), &l);
        status[c] = (int)st;
        if (unit) unit[c] = l.x;
    }
    delete vh;
    delete dm;
    free(fv);
    free(dec);
    free(imp);
    return m;
}

/* BCP of one cube from an empty trail. status: 0 SAT, 1 UNSAT (conflict), 2 UNDEF. implied = literals in discovery order. */
int ref_propagate(void *hv, const int32_t *cube, int k, int32_t *status, int32_t *implied, int32_t *n_implied)
{
    RefHandle *h = (RefHandle *)hv;
    ensure_common(h);
    if (k > MAX_VARS) return -1;
    if (!h->pca) {
        h->pfv = (Var *)malloc(sizeof(Var) * (h->n_vars + 1));
        h->pdec = (Decision *)malloc(sizeof(Decision) * (h->n_vars + 1));
        h->pimp = (Decision *)malloc(sizeof(Decision) * (h->n_vars + 1));
        h->pdm = new DecisionMaker(h->formula_dev, (size_t)h->n_vars);
        h->pvh = new VariablesStateHandler(h->n_vars, &h->dead_view, h->pdm, h->pfv, h->pdec, h->pimp);
        h->pdm->set_vars_handler(h->pvh);
        h->pca = new ConflictAnalyzerWithWatchedLits(h->n_vars, h->formula_dev, h->pvh, true, h->max_impl,
                                                     h->pdm, h->stats, h->repo);
    }
    for (int i = 0; i < k; i++) { Lit l; l.x = cube[i]; h->pcube.add(l); }
    h->pvh->set_assumptions(&h->pcube);
    sat_status st = h->pca->set_assumptions(&h->pcube);
    *status = (int)st;
    int n = (int)h->pvh->n_implications();
    for (int i = 0; i < n; i++) implied[i] = h->pvh->get_implication(i)->literal.x;
    *n_implied = n;
    h->pvh->reset();
    h->pca->reset();
    h->pcube.remove_n_last(k);
    return 0;
}

/* One job.  verdict: 0 SAT, 1 UNSAT, 2 UNDEF (MAX_ITERATIONS cap as shipped).  model: lits (decisions, cube, implications). */
int ref_solve(void *hv, const int32_t *cube, int k, int32_t *verdict, int32_t *model, int32_t *n_model)
{
    RefHandle *h = (RefHandle *)hv;
    ensure_common(h);
    if (k > MAX_VARS) return -1;
    if (!h->solver) {
        h->fv = (Var *)malloc(sizeof(Var) * (h->n_vars + 1));
        h->dec = (Decision *)malloc(sizeof(Decision) * (h->n_vars + 1));
        h->imp = (Decision *)malloc(sizeof(Decision) * (h->n_vars + 1));
        h->solver = new SATSolver(h->formula_dev, h->n_vars, h->max_impl, &h->dead_view, h->stats, h->repo,
                                  h->fv, h->dec, h->imp);
    }
    for (int i = 0; i < k; i++) { Lit l; l.x = cube[i]; h->cube.add(l); }
    sat_status st = h->solver->solve(&h->cube);
    *verdict = (int)st;
    *n_model = 0;
    if (st == sat_status::SAT) {
        int n = (int)h->solver->get_results_size();
        std::vector<Lit> buf(n + 1);
        h->solver->get_results(buf.data());
        for (int i = 0; i < n; i++) model[i] = buf[i].x;
        *n_model = n;
    }
    /* KernelContext::finished(): reset + drop the cube; learnt ring and level-0 facts persist (Appendix B.3) */
    h->solver->reset();
    h->cube.remove_n_last(k);
    return 0;
}

long long ref_counter_implications(void) { return gpsat_ref_n_implications; }
long long ref_counter_decisions(void) { return gpsat_ref_n_decisions; }
void ref_counters_reset(void) { gpsat_ref_n_implications = 0; gpsat_ref_n_decisions = 0; }
int ref_max_iterations(void)
{
#ifdef MAX_ITERATIONS
    return MAX_ITERATIONS;
#else
    return 0;
#endif
}

void ref_close(void *hv)
{
    /* the reference never frees its device structures (main.cu:304-326); neither do we beyond the handle */
    delete (RefHandle *)hv;
}

} /* extern "C" */
