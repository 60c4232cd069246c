#!/bin/bash
# builds libmshgnn_b200 of another commit into morphsym-hgnn_b200/lib/ab/ for same-box A/B runs (MSHGNN_LIB=<path>): tools/build_ab_lib.sh <commit>
set -e
C=$1
ROOT=$(cd "$(dirname "$0")/.." && pwd)
S=$ROOT/gpurun_out/ab_src_$C
rm -rf "$S"; mkdir -p "$S/a/b/csrc" "$S/a/include" "$ROOT/morphsym-hgnn_b200/lib/ab"
for f in $(git -C "$ROOT" ls-tree --name-only "$C" morphsym-hgnn_b200/csrc/); do git -C "$ROOT" show "$C:$f" > "$S/a/b/csrc/$(basename "$f")"; done
git -C "$ROOT" show "$C:include/mshgnn_b200.h" > "$S/a/include/mshgnn_b200.h"
(cd "$S/a/b/csrc" && nvcc -std=c++17 -O3 -lineinfo -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -o "$ROOT/morphsym-hgnn_b200/lib/ab/libmshgnn_b200_$C.so" api.cu plan.cu)
echo "$ROOT/morphsym-hgnn_b200/lib/ab/libmshgnn_b200_$C.so"
