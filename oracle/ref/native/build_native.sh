#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY (oracle/).  Native sm_100a build of the UNMODIFIED reference (every .cu of its
# CMakeLists.txt:34-112, compiled where it lies under $GPSAT_REFERENCE_SRC) with the Boost-free front end of
# oracle/ref/native/boostfree_frontend.cu in place of FileManager/CnfReader.cpp and ParametersManager.cpp.
# Output: oracle/_ref/gpupsat_ref_native (git-ignored; travels to the GPU box).  nvcc cross-compiles without a GPU.
# Purpose: the same-silicon baseline of SURVEY.md section 8(d) (tools/run_native_reference.py).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="${GPSAT_REFERENCE_SRC:-/root/reference/src}"
OUT="$HERE/../../_ref"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
[ -d "$SRC" ] || { echo "reference sources not found at $SRC" >&2; exit 3; }
mkdir -p "$OUT/obj_native"
FILES=$(sed -n '34,112p' "$SRC/CMakeLists.txt" | grep -o '[A-Za-z/_0-9-]*\.cu' | sort -u)
FLAGS="-std=c++20 -O2 -w -DNDEBUG -gencode arch=compute_100a,code=sm_100a -rdc=true -I$SRC -I$SRC/SATSolver -I$SRC/Utils"
objs=""
n=0
for f in $FILES; do
  [ -f "$SRC/$f" ] || continue
  o="$OUT/obj_native/$(echo "$f" | tr '/' '_').o"; objs="$objs $o"
  $NVCC $FLAGS -c "$SRC/$f" -o "$o" &
  n=$((n + 1)); if [ $((n % 8)) -eq 0 ]; then wait; fi
done
o="$OUT/obj_native/boostfree_frontend.o"; objs="$objs $o"
$NVCC $FLAGS -c "$HERE/boostfree_frontend.cu" -o "$o" &
wait
$NVCC -gencode arch=compute_100a,code=sm_100a -rdc=true $objs -o "$OUT/gpupsat_ref_native" -lcudart -lcuda
echo "built $OUT/gpupsat_ref_native"
