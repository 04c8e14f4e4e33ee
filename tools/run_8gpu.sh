#!/bin/bash
# One multi-GPU session: the in-process multi-device tests, the host<->device copy ceilings of the box
# (tools/pcie_matrix.sh) and bench.py at N = 8 and N = 4 (torchrun, one rank per GPU).  Everything lands in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
G=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${G}gpu.txt 2>&1
lscpu | egrep "Model name|^CPU\(s\)|NUMA|Socket" >> gpurun_out/topo_${G}gpu.txt
free -g | head -2 >> gpurun_out/topo_${G}gpu.txt
timeout 300 python -m pytest tests/test_gpu_round2.py -x -q -k "distinct_devices or own_device" 2>&1 | tail -3
timeout 400 tools/pcie_matrix.sh $G > gpurun_out/pcie_${G}gpu.jsonl 2> gpurun_out/pcie_${G}gpu.err
tail -3 gpurun_out/pcie_${G}gpu.jsonl
for N in $G 4; do
  [ "$N" -le "$G" ] || continue
  extra=""
  [ "$N" -eq "$G" ] || extra="--no-configs"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --steps 5 --warmup 3 $extra > gpurun_out/bench_r2_${N}gpu.json 2> gpurun_out/bench_r2_${N}gpu.err
  echo "bench N=$N rc=$? $(wc -c < gpurun_out/bench_r2_${N}gpu.json) bytes"
  tail -2 gpurun_out/bench_r2_${N}gpu.err
done
