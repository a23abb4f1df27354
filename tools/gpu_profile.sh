#!/bin/bash
# One `ncu --set full` capture per kernel named on the command line, through tools/kbench.py.
#   tools/gpu_profile.sh TAG  "urdf|op|kernel-regex|states" ...
# Writes gpurun_out/TAG_<kernel>.ncu-rep and gpurun_out/TAG_<kernel>_summary.json (tools/ncu_summary.py).
# Run only after the same kbench command has exited 0 without ncu (B200_PROFILING.md).
TAG=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
    IFS='|' read -r urdf op kre states <<< "$spec"
    name=$(echo "$kre" | tr -c 'A-Za-z0-9_' '_' | sed 's/_*$//')
    out=gpurun_out/${TAG}_${name}
    timeout 600 python tools/kbench.py --urdf "$urdf" --ops "$op" --states "$states" --iters 2 --tag plain > "$out.plain.log" 2>&1 || { echo "plain run failed: $spec"; tail -5 "$out.plain.log"; continue; }
    timeout 900 ncu --set full --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,sm__inst_executed_pipe_fp64.sum --clock-control none --import-source on -k "regex:$kre" --launch-skip 2 --launch-count 1 -f -o "$out" \
        python tools/kbench.py --urdf "$urdf" --ops "$op" --states "$states" --iters 2 --tag ncu > "$out.ncu.log" 2>&1 || { echo "ncu failed: $spec"; tail -5 "$out.ncu.log"; continue; }
    python tools/ncu_summary.py "$out.ncu-rep" "${out}_summary.json" || echo "summary failed: $spec"
    # gpurun brings back at most 64 MiB: keep only the reports named in $KEEP_REP (regex) for source-level reading
    if [ -z "$KEEP_REP" ] || ! echo "$name" | grep -Eq "$KEEP_REP"; then rm -f "$out.ncu-rep"; fi
done
