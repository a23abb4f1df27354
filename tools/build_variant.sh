#!/bin/bash
# Builds var/librb_NAME.so = the library with extra nvcc flags applied to the FR3 kernels only (tuning experiments;
# every other object is reused from the main build).   usage: tools/build_variant.sh NAME "-DRB_...=..." [ptxas]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
C=$ROOT/rigidbody_rs_b200/csrc
NAME=$1; FLAGS=$2
make -s -j8 -C "$C"
rm -rf "$C/build_$NAME"; mkdir -p "$C/build_$NAME" "$ROOT/var"
cp "$C"/build/*.o "$C/build_$NAME/"; rm -f "$C/build_$NAME/rb_kernels_fr3.o"
if [ "$3" = ptxas ]; then
  make -s -j8 -C "$C" B=build_$NAME OUT=$ROOT/var/librb_$NAME.so EXTRA_NVFLAGS="$FLAGS -Xptxas -v" 2>&1 | grep -A2 "TabFr3dE" | grep "Compiling\|spill\|Used" | sed 's/ptxas info    : //' | cut -c1-140
else
  make -s -j8 -C "$C" B=build_$NAME OUT=$ROOT/var/librb_$NAME.so EXTRA_NVFLAGS="$FLAGS" 2>&1 | grep -i "error" || true
fi
rm -rf "$C/build_$NAME"
