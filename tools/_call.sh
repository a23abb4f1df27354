set -x
nvidia-smi -L
timeout 1700 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python tools/e2e_probe.py --devices 0 > gpurun_out/r2_e2e_probe.jsonl 2>gpurun_out/e2e.err
python tools/e2e_probe.py --devices 0,1 >> gpurun_out/r2_e2e_probe.jsonl 2>>gpurun_out/e2e.err
python tools/e2e_probe.py --devices 0 --pageable --states 4194304 >> gpurun_out/r2_e2e_probe.jsonl 2>>gpurun_out/e2e.err
cat gpurun_out/r2_e2e_probe.jsonl; tail -3 gpurun_out/e2e.err
