#!/bin/bash
# tools/pcie_matrix.sh [max_gpus] -> one JSON line per (gpus, mode, dir): the host<->device copy ceilings of this box
# (tools/pcie_probe.cu).  Used for profiles/pcie_*.md.
set -u
cd "$(dirname "$0")/.."
[ -x tools/pcie_probe ] || nvcc -O2 -std=c++17 -o tools/pcie_probe tools/pcie_probe.cu -lpthread
MAXG=${1:-$(nvidia-smi -L | wc -l)}
for g in 1 2 4 8; do
  [ "$g" -le "$MAXG" ] || continue
  for mode in proc thread; do
    for dir in h2d d2h bidir; do
      tools/pcie_probe --gpus $g --mode $mode --dir $dir --mb 2048 --reps 6 --chunk-mb 256
    done
  done
done
