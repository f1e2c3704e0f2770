#!/bin/bash
# Round-2 evidence run (ONE gpurun call): ncu --set full captures of the kernels VERDICT r1 asked for (f64 real N=2048/4096,
# f64 c2c N=4096) plus the two f32 kernels under work.  Every command runs to exit 0 WITHOUT ncu first.
set -u
O=gpurun_out
SPECS="${SPECS:-r2c_f64:2048 r2c_f64:4096 c2c_f64:4096 c2c_split:4096 c2c_split:2048 stft:2048 stft:1024}"
timeout 300 python tools/prof_one.py $SPECS > $O/plain_prof.log 2>&1 || { echo "plain prof_one failed"; tail -5 $O/plain_prof.log; exit 1; }
for s in $SPECS; do
  R=$O/full_${s/:/_}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_ -s 2 -c 1 -f -o $R python tools/prof_one.py $s > $O/ncu_full_${s/:/_}.log 2>&1
  ncu -i $R.ncu-rep --page raw --csv > $R.raw.csv 2>/dev/null
  ncu -i $R.ncu-rep --page source --csv 2>/dev/null | gzip > $R.source.csv.gz
  rm -f $R.ncu-rep
done
ls -la $O | tail -20
