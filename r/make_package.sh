#!/bin/sh
# make_package.sh — assemble the accelerated NNLM R package from a checkout of the reference package.
#
#   r/make_package.sh <path-to-NNLM-checkout> <output-dir>
#
# The reference's R code, man pages, data and tests are taken from the checkout UNCHANGED (nothing of them lives in this
# repository); only the native layer is swapped:
#   src/*.cpp, src/nnlm.h  (Rcpp + Armadillo + OpenMP)  ->  src/shim.c + src/Makevars  (plain C over libnnlm_b200.so)
#   DESCRIPTION            drop Rcpp from Imports and the whole LinkingTo field (Rcpp, RcppArmadillo, RcppProgress)
#   NAMESPACE              drop import(Rcpp); useDynLib(NNLM, .registration = TRUE) stays (R_init_NNLM registers
#                          _NNLM_c_nnmf/17 and _NNLM_c_nnlm/9 exactly like src/RcppExports.cpp:56-65)
#   R/RcppExports.R        unchanged: it only does .Call(`_NNLM_c_nnmf`, ...) / .Call(`_NNLM_c_nnlm`, ...)
# Then:  NNLM_B200_HOME=<this repo> R CMD INSTALL <output-dir>   (needs R, a B200 and `make -C nnlm_b200/csrc` first).
set -eu
SRC=${1:?usage: make_package.sh <NNLM checkout> <output dir>}
OUT=${2:?usage: make_package.sh <NNLM checkout> <output dir>}
HERE=$(cd "$(dirname "$0")" && pwd)
[ -f "$SRC/DESCRIPTION" ] && grep -q '^Package: NNLM' "$SRC/DESCRIPTION" || { echo "$SRC is not an NNLM checkout" >&2; exit 1; }
mkdir -p "$OUT"
for d in R man data tests vignettes; do [ -d "$SRC/$d" ] && cp -r "$SRC/$d" "$OUT/"; done
for f in LICENSE NEWS README.md .Rbuildignore; do [ -f "$SRC/$f" ] && cp "$SRC/$f" "$OUT/"; done
mkdir -p "$OUT/src"
cp "$HERE/src/shim.c" "$HERE/src/Makevars" "$OUT/src/"
# DESCRIPTION: remove the `Rcpp (>= ...)` import line and the LinkingTo block (field line + its continuation lines)
awk '
  /^LinkingTo:/ { skip = 1; next }
  skip && /^[ \t]/ { next }
  { skip = 0 }
  /^[ \t]+Rcpp[ \t]*(\(.*\))?,?[ \t]*$/ { next }
  { print }
' "$SRC/DESCRIPTION" > "$OUT/DESCRIPTION"
printf 'SystemRequirements: CUDA 12.9+, an NVIDIA B200 (sm_100a), libnnlm_b200.so (NNLM_B200_HOME)\n' >> "$OUT/DESCRIPTION"
grep -v '^import(Rcpp)' "$SRC/NAMESPACE" > "$OUT/NAMESPACE"
echo "assembled $OUT (native layer: src/shim.c over libnnlm_b200.so)"
