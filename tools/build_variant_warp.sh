#!/bin/bash
# var/librb_NAME.so = the library with extra nvcc flags applied to rb_kernels_warp.cu only (long-chain FD experiments).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
C=$ROOT/rigidbody_rs_b200/csrc
NAME=$1; FLAGS=$2
make -s -j8 -C "$C"
rm -rf "$C/build_$NAME"; mkdir -p "$C/build_$NAME" "$ROOT/var"
cp "$C"/build/*.o "$C/build_$NAME/"; rm -f "$C/build_$NAME/rb_kernels_warp.o"
make -s -j8 -C "$C" B=build_$NAME OUT=$ROOT/var/librb_$NAME.so EXTRA_NVFLAGS="$FLAGS -Xptxas -v" 2>&1 | grep -A2 "rb[hq]_fd_kernel" | grep "spill\|Used" | sed 's/ptxas info    : //' | cut -c1-140
rm -rf "$C/build_$NAME"
