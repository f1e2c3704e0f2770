#!/bin/bash
# Host-link diagnosis for the e2e number (run on an 8-GPU box: gpurun --gpus 8 -- 'bash tools/pcie_diag.sh'):
# topology, NUMA layout, the concurrent pinned-copy ceiling at 1/2/4/8 GPUs for three ways of pinning host memory, and
# then the bench's own e2e leg at 8 ranks.
O=gpurun_out
{
  echo "== nproc $(nproc)"; lscpu | egrep "Model name|Socket|NUMA|Thread|Core" ;
  echo "== numa nodes: $(cat /sys/devices/system/node/online 2>/dev/null)"; for n in /sys/devices/system/node/node*; do echo "$n: cpus $(cat $n/cpulist) $(grep MemTotal $n/meminfo)"; done
  echo "== gpu pci numa"; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ]; then echo "$(basename $d) class $(cat $d/class) numa $(cat $d/numa_node) link $(cat $d/current_link_speed 2>/dev/null) x$(cat $d/current_link_width 2>/dev/null)"; fi; done
  echo "== thp: $(cat /sys/kernel/mm/transparent_hugepage/enabled 2>/dev/null)"
  nvidia-smi topo -m | sed 's/\x1b\[[0-9;]*m//g'
  free -g | head -2
} > $O/pcie_topology.txt 2>&1
make -C tools/microbench pcie_peak > /dev/null 2>&1
timeout 600 tools/microbench/pcie_peak 256 6 > $O/pcie_peak.jsonl 2> $O/pcie_peak.err
tail -3 $O/pcie_peak.err
N=$(nvidia-smi -L | wc -l)
if [ "$N" -ge 2 ]; then
  timeout 900 python bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench N=$N rc=$?"; tail -3 $O/bench_n$N.err
fi
