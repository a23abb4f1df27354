#!/bin/bash
# Counts FP64-pipe SASS instructions per kernel in the built library (straight-line kernels: the static count is
# the per-thread dynamic count, bar the rare slow paths of sincos / reciprocal).
LIB=${1:-rigidbody_rs_b200/librigidbody_b200.so}
cuobjdump -sass "$LIB" | awk '
/Function :/ { if (name != "") printf "%-60s DFMA %5d DADD %4d DMUL %4d other-F64 %4d | FP64 total %5d  MUFU %3d  all %6d\n", name, f, a, m, o, f+a+m+o, mu, all; name=$3; f=a=m=o=mu=all=0 }
/^\s+\/\*[0-9a-f]+\*\/\s+[A-Z@]/ { all++ }
/ DFMA/ {f++} / DADD/ {a++} / DMUL/ {m++} / DSETP| DMNMX| F2F\.F64| I2F\.F64| F2I.*F64| D2I| I2D/ {o++} / MUFU/ {mu++}
END { printf "%-60s DFMA %5d DADD %4d DMUL %4d other-F64 %4d | FP64 total %5d  MUFU %3d  all %6d\n", name, f, a, m, o, f+a+m+o, mu, all }' | sed -E 's/_Z[0-9]+//' | cut -c1-200
