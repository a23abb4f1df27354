timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "chain32 or huge" 2>&1 | tail -3
K=gpurun_out/r2_kb10.jsonl; : > $K
python tools/kbench.py --urdf assets/chain32.urdf --ops rnea --states 1048576 --tag streamed2 >> $K 2>gpurun_out/kb.err
cat $K; tail -3 gpurun_out/kb.err
KEEP_REP='none' tools/gpu_profile.sh r2_streamed "assets/chain32.urdf|rnea|rb_long_rnea_kernel|1048576"
cat gpurun_out/r2_streamed_rb_long_rnea_kernel_summary.json
