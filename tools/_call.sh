set -x
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ub tools/ubench_fp64_halfwarp.cu && /tmp/ub > gpurun_out/r2_ubench_fp64_halfwarp.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "rollout or rnea_fd or not_spd or golden or smoke" 2>&1 | tail -15
K=gpurun_out/r2_kb1.jsonl; : > $K
python tools/kbench.py --ops rnea,fd,rnea_fd,rollout --tag default >> $K 2>gpurun_out/kb.err
python tools/kbench.py --ops rollout --traj 8192 --tag default_8192 >> $K 2>>gpurun_out/kb.err
python tools/kbench.py --ops rollout --traj 16384 --tag default_16384 >> $K 2>>gpurun_out/kb.err
python tools/kbench.py --ops rollout --traj 32768 --tag default_32768 >> $K 2>>gpurun_out/kb.err
RIGIDBODY_B200_ROLLOUT=thread python tools/kbench.py --ops rollout --tag thread >> $K 2>>gpurun_out/kb.err
RIGIDBODY_B200_ROLLOUT=thread python tools/kbench.py --ops rollout --traj 8192 --tag thread_8192 >> $K 2>>gpurun_out/kb.err
for v in va vb vc vd; do
  RIGIDBODY_B200_LIB=var/librb_$v.so python tools/kbench.py --ops rnea_fd,rollout --tag $v >> $K 2>>gpurun_out/kb.err
  RIGIDBODY_B200_LIB=var/librb_$v.so python tools/kbench.py --ops rollout --traj 8192 --tag ${v}_8192 >> $K 2>>gpurun_out/kb.err
done
cat $K
tail -3 gpurun_out/kb.err
cat gpurun_out/r2_ubench_fp64_halfwarp.txt
