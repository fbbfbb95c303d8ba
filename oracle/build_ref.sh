#!/usr/bin/env bash
# Build the UNMODIFIED reference (JohnGBaker/ptmcmc) from the sources where they lie
# under $PTMCMC_REFERENCE (default /root/reference) plus the trace drivers in
# oracle/ref_drivers/ (our own code, written against the reference's public API).
# Outputs go ONLY to oracle/_ref/ (git-ignored, travels to the GPU box with gpurun).
# Flags = the reference's shipped flags (Makefile.ac:1-2) minus MPI, plus
# -ffp-contract=off so the CPU arithmetic is unfused IEEE fp64.
# Nothing in the product links against these outputs: test/bench infrastructure only.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
R="${PTMCMC_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$R" ]; then
  echo "build_ref: $R not present; keeping prebuilt files in $OUT" >&2
  exit 0
fi
E="$R/eigen-eigen-323c052e1731"
mkdir -p "$OUT/obj"
# NB: the image exports CXX=/opt/gcc/bin/g++, which has no libgomp; use the system g++.
CXX="${REF_CXX:-/usr/bin/g++}"
CF="-O2 -fPIC -fopenmp -std=c++11 -ffp-contract=off -w -I$R -I$R/ProbabilityDist -I$E"
objs=()
build_obj() { # src obj
  if [ ! -f "$2" ] || [ "$1" -nt "$2" ]; then $CXX -c $CF "$1" -o "$2"; fi
  objs+=("$2")
}
pids=()
for f in ProbabilityDist newran1 newran2 myexcept simpstr extreal; do
  build_obj "$R/ProbabilityDist/$f.cxx" "$OUT/obj/$f.o" &
  pids+=($!)
done
for f in states chain probability_function proposal_distribution ptmcmc; do
  build_obj "$R/$f.cc" "$OUT/obj/$f.o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
rm -f "$OUT/libptmcmc_ref.a"
ar rc "$OUT/libptmcmc_ref.a" "$OUT"/obj/*.o
ENG="$HERE/../ptmcmc_b200/csrc"
for d in "$HERE"/ref_drivers/*.cc; do
  [ -e "$d" ] || continue
  b="$(basename "${d%.cc}")"
  if [ "$b" = "gpu_dropin" ]; then
    # the reference-side binding of the engine compiled against the real reference headers (include/gpu_parallel_tempering_chains.hh) and
    # linked with the engine's C ABI; needs the engine library to be built first
    if [ -f "$ENG/libptmcmc_b200.so" ]; then
      $CXX $CF -std=c++11 -fno-access-control -I"$HERE/../include" "$d" -o "$OUT/$b" "$OUT/libptmcmc_ref.a" -L"$ENG" -lptmcmc_b200 -Wl,-rpath,"\$ORIGIN/../../ptmcmc_b200/csrc"
    else
      echo "build_ref: skipping $b (ptmcmc_b200/csrc/libptmcmc_b200.so not built yet)" >&2
    fi
    continue
  fi
  $CXX $CF -fno-access-control "$d" -o "$OUT/$b" "$OUT/libptmcmc_ref.a"
done
echo "build_ref: built $(ls "$OUT" | tr '\n' ' ')"
