set -x
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; (numactl -H || lscpu | grep -i "numa\|socket\|model name\|^CPU(s)") >> gpurun_out/topo.txt 2>&1; cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c >> gpurun_out/topo.txt; nvidia-smi --query-gpu=index,pci.bus_id --format=csv >> gpurun_out/topo.txt; for d in $(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader); do b=$(echo $d | tr 'A-Z' 'a-z' | sed 's/^0000//'); echo "$d numa $(cat /sys/bus/pci/devices/$b/numa_node 2>/dev/null) cpus $(cat /sys/bus/pci/devices/$b/local_cpulist 2>/dev/null)" >> gpurun_out/topo.txt; done; free -g >> gpurun_out/topo.txt
cat gpurun_out/topo.txt
python tools/e2e_probe.py --devices 0 --states 8388608 > gpurun_out/r2_e2e_probe2.jsonl 2>gpurun_out/e2e.err
python tools/e2e_probe.py --devices 0,1 --states 16777216 >> gpurun_out/r2_e2e_probe2.jsonl 2>>gpurun_out/e2e.err
python tools/e2e_probe.py --devices 1 --states 8388608 >> gpurun_out/r2_e2e_probe2.jsonl 2>>gpurun_out/e2e.err
cat gpurun_out/r2_e2e_probe2.jsonl
python bench.py --steps 10 --warmup 3 --no-cpu --no-chain32 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('peak', d['roofline']['peak'], 'frac', d['roofline']['frac'], d['roofline_rnea']['frac'], d['configs']['fused_rnea_fd']['roofline']['frac'], d['configs']['rollout']['roofline']['frac'])"
