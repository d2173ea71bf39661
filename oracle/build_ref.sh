#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Builds the *unmodified* reference (CmdStan 2.36 / Stan / Stan Math
# sources where they lie under /root/reference) into oracle/_ref/ (git-ignored, travels with gpurun).
# Recipe follows SURVEY.md Appendix C (flags: stan_math make/compiler_flags O=3, STAN_THREADS).
# Nothing from the reference tree is copied into the repository; only binaries land in oracle/_ref/.
#
#   oracle/build_ref.sh [model ...]      (default: every model under oracle/models/*.hpp)
#
# Products: oracle/_ref/tbb/libtbb.so.2, oracle/_ref/main.o, oracle/_ref/<model> executables
#           (link-time hook new_model(), C/src/cmdstan/command.hpp:71-73), oracle/_ref/stansummary.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
R="${REFERENCE_ROOT:-/root/reference}"
W="$HERE/_ref"
if [ ! -d "$R/src/cmdstan" ]; then
  echo "build_ref: $R not present -- using prebuilt oracle/_ref as is" >&2
  exit 0
fi
M="$R/stan/lib/stan_math"
mkdir -p "$W"
INC="-I $R/src -I $R/stan/src -I $R/stan/lib/rapidjson_1.1.0 -I $R/lib/CLI11-1.9.1 -I $M -I $M/lib/eigen_3.4.0 \
 -I $M/lib/boost_1.84.0 -I $M/lib/sundials_6.1.1/include -I $M/lib/sundials_6.1.1/src/sundials -I $M/lib/tbb_2020.3/include"
FL="-std=c++17 -O3 -DSTAN_THREADS -D_REENTRANT -DBOOST_DISABLE_ASSERTS -D_USE_MATH_DEFINES -w"
LINK="-Wl,-L,$W/tbb -Wl,-rpath,\$ORIGIN/tbb -ltbb -lpthread"

# 1. vendored TBB 2020.3 -> libtbb.so.2 (needs a writable copy of its build tree; lives in _ref only)
if [ ! -f "$W/tbb/libtbb.so.2" ]; then
  rm -rf "$W/tbb_src" && cp -r "$M/lib/tbb_2020.3" "$W/tbb_src" && chmod -R u+w "$W/tbb_src"
  mkdir -p "$W/tbb"
  (cd "$W/tbb" && tbb_root=../tbb_src CXX=g++ CC=gcc make -j8 -r -f "$W/tbb_src/build/Makefile.tbb" \
     compiler=gcc cfg=release stdver=c++17 \
     CXXFLAGS="-Wno-unknown-warning-option -Wno-deprecated-copy -Wno-missing-attributes -Wno-class-memaccess -Wno-sized-deallocation" >/dev/null)
  rm -rf "$W/tbb_src"; find "$W/tbb" -name '*.o' -delete; find "$W/tbb" -name '*.d' -delete
fi
# 2. cmdstan main.o (once for all models)
[ -f "$W/main.o" ] || g++ $FL $INC -c "$R/src/cmdstan/main.cpp" -o "$W/main.o" &
# 3. stansummary
[ -f "$W/stansummary" ] || g++ $FL $INC "$R/src/cmdstan/stansummary.cpp" $LINK -o "$W/stansummary" &
# 4. models: hand-written model_base_crtp classes (stanc is not available), compiled against the reference headers
MODELS=("$@")
if [ ${#MODELS[@]} -eq 0 ]; then for f in "$HERE"/models/*.hpp; do [ -f "$f" ] && MODELS+=("$(basename "${f%.hpp}")"); done; fi
for m in "${MODELS[@]}"; do
  src="$HERE/models/$m.hpp"
  if [ ! -f "$W/$m.o" ] || [ "$src" -nt "$W/$m.o" ] || [ "$HERE/models/model_common.h" -nt "$W/$m.o" ]; then
    g++ $FL $INC -I "$HERE/models" -x c++ -c "$src" -o "$W/$m.o" &
  fi
done
wait
for m in "${MODELS[@]}"; do
  g++ "$W/$m.o" "$W/main.o" $LINK -o "$W/$m"
done
# 5. small helper tools compiled against the reference headers (RNG KATs, sampler traces)
for t in "$HERE"/ref_tools/*.cpp; do
  [ -f "$t" ] || continue
  b="$W/$(basename "${t%.cpp}")"
  if [ ! -f "$b" ] || [ "$t" -nt "$b" ]; then g++ $FL $INC -I "$HERE/models" "$t" $LINK -o "$b" & fi
done
wait
echo "build_ref: done -> $W"
