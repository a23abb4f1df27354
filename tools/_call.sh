set -x
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ub3 tools/ubench_fp64_peak.cu && /tmp/ub3 > gpurun_out/r2_ubench_fp64_peak.txt 2>&1
cat gpurun_out/r2_ubench_fp64_peak.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "rollout or not_spd or golden" 2>&1 | tail -5
RIGIDBODY_B200_LIB=var/librb_ws.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "rollout_launch_modes or not_spd" 2>&1 | tail -3
K=gpurun_out/r2_kb7.jsonl; : > $K
for t in 65536 32768 16384 8192; do
python tools/kbench.py --ops rollout --traj $t --tag even_$t >> $K 2>>gpurun_out/kb.err
RIGIDBODY_B200_ROLLOUT=greedy python tools/kbench.py --ops rollout --traj $t --tag greedy_$t >> $K 2>>gpurun_out/kb.err
RIGIDBODY_B200_LIB=var/librb_nointer.so python tools/kbench.py --ops rollout --traj $t --tag nointer_even_$t >> $K 2>>gpurun_out/kb.err
done
cat $K
tail -3 gpurun_out/kb.err
