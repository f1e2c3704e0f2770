#!/bin/bash
# Which GPUs share a host bridge?  (8-GPU box: gpurun --gpus 8 -- 'bash tools/pcie_diag2.sh')
# Pairs {0,k}, then 2 and 4 GPUs spread over the box, then the bench at 2 and 4 ranks with the spread rank -> GPU order.
O=gpurun_out
timeout 300 tools/microbench/pcie_peak 256 6 pairs hostalloc > $O/pcie_pairs.jsonl 2> $O/pcie_pairs.err
timeout 300 tools/microbench/pcie_peak 256 6 spread hostalloc > $O/pcie_spread.jsonl 2>> $O/pcie_pairs.err
tail -2 $O/pcie_pairs.err
for N in 2 4; do
  timeout 600 python bench.py --gpus $N --steps 20 --warmup 3 > $O/bench_n${N}_spread.json 2> $O/bench_n${N}_spread.err; echo "bench N=$N rc=$?"; tail -2 $O/bench_n${N}_spread.err
done
