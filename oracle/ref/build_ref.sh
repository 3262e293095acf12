#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY (oracle/).  Builds the reference's own solver classes as host C++:
#   oracle/_ref/libgpsat_ref.so        as shipped  (MAX_ITERATIONS 1000 -> UNDEF on hard jobs)
#   oracle/_ref/libgpsat_ref_nocap.so  same sources with SATSolver/Configs.cuh:23 (MAX_ITERATIONS) undefined
# Sources are compiled where they lie under $GPSAT_REFERENCE_SRC (default /root/reference/src); nothing is
# copied into the repo.  Two debug printers use <<<1,1>>> launches g++ cannot parse; sed-neutralised copies
# of those two files go to oracle/_ref/gen/ (git-ignored).  Recipe: SURVEY.md Appendix A.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="${GPSAT_REFERENCE_SRC:-/root/reference/src}"
OUT="$HERE/../_ref"
[ -d "$SRC" ] || { echo "reference sources not found at $SRC" >&2; exit 3; }
mkdir -p "$OUT/gen" "$OUT/obj" "$OUT/obj_nocap"

sed 's/print_clause_kernel <<< 1, 1>>>(c);/print_clause_kernel(c);/' "$SRC/SATSolver/SolverTypes.cu" > "$OUT/gen/SolverTypes.cu"
sed 's/print_dev <<< 1, 1>>>(\*this);/print_dev(*this);/'            "$SRC/Utils/CUDAClauseVec.cu"   > "$OUT/gen/CUDAClauseVec.cu"
# cap-lifted configuration: same header with the MAX_ITERATIONS line commented out; pre-included so its
# include guard shadows the original
sed 's|^#define MAX_ITERATIONS 1000|// #define MAX_ITERATIONS 1000  (lifted by oracle/ref/build_ref.sh)|' "$SRC/SATSolver/Configs.cuh" > "$OUT/gen/Configs_nocap.cuh"

FILES="
BCPStrategy/ClauseListStructure.cu BCPStrategy/WatchedClausesList.cu
ClauseLearning/LearntClauseRepository.cu ClauseLearning/LearntClausesManager.cu
ConflictAnalysis/CUDAListGraph.cu ConflictAnalysis/ConflictAnalyzer.cu ConflictAnalysis/ConflictAnalyzerWithWatchedLits.cu
ConflictAnalysis/GraphAnalyzer.cu ConflictAnalysis/GraphStructure.cu
DecisionStrategy/VSIDS.cu
Restarts/GeometricRestartsManager.cu
SATSolver/SATSolver.cu SATSolver/VariablesStateHandler.cu SATSolver/DecisionMaker.cu SATSolver/Backtracker.cu SATSolver/JobsQueue.cu
Statistics/RuntimeStatistics.cu ErrorHandler/CudaMemoryErrorHandler.cu
Preprocessing/RepeatedLiteralsRemover.cu Preprocessing/UnaryClausesRemover.cu
FileManager/FormulaData.cu FileManager/FileUtils.cu
JobsManager/JobChooser.cu JobsManager/VariableChooser.cu JobsManager/SimpleJobChooser.cu
"
CXX="${CXX:-g++}"
FLAGS="-std=c++20 -O2 -w -fpermissive -fPIC -D__CUDA_ARCH__=1000 -x c++ -include $HERE/stub/cuda_runtime.h -I$HERE/stub -I$SRC -I$SRC/SATSolver -I$SRC/Utils"

build_variant() {  # $1 = obj dir, $2 = extra flags, $3 = output .so
  local objdir="$1" extra="$2" so="$3" objs="" pids=""
  for f in $FILES; do
    o="$objdir/$(echo "$f" | tr '/' '_').o"; objs="$objs $o"
    $CXX $FLAGS $extra -c "$SRC/$f" -o "$o" &
  done
  for g in SolverTypes CUDAClauseVec; do
    o="$objdir/gen_$g.o"; objs="$objs $o"
    $CXX $FLAGS $extra -c "$OUT/gen/$g.cu" -o "$o" &
  done
  o="$objdir/ref_driver.o"; objs="$objs $o"
  $CXX $FLAGS $extra -c "$HERE/ref_driver.cpp" -o "$o" &
  wait
  $CXX -shared -o "$so" $objs \
      -Wl,--wrap=_ZN21VariablesStateHandler15new_implicationE8Decision \
      -Wl,--wrap=_ZN21VariablesStateHandler12new_decisionE8Decision
}
build_variant "$OUT/obj"       ""                                   "$OUT/libgpsat_ref.so"
build_variant "$OUT/obj_nocap" "-include $OUT/gen/Configs_nocap.cuh" "$OUT/libgpsat_ref_nocap.so"
echo "built $OUT/libgpsat_ref.so $OUT/libgpsat_ref_nocap.so"
