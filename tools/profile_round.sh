#!/bin/bash
# Round-end evidence run (ONE gpurun call):  /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/profile_round.sh'
# then, back in the container:               python tools/make_profiles.py r01
# Every command runs to exit 0 WITHOUT ncu before it runs under ncu; numbers printed under ncu are never bench values.
set -u
O=gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
rm -f $O/sweep_burst.jsonl $O/sweep_inv.jsonl $O/sweep_sustained.jsonl $O/full_*

timeout 1200 python tools/sweep.py --kinds c2c_split,c2c_il,r2c,c2r,c2c_f64,r2c_f64,c2r_f64,stft --sizes 8,16,32,64,128,256,512,1024,2048,4096,8192 --out $O/sweep_burst.jsonl > $O/sweep_burst.log 2>&1
timeout 600 python tools/sweep.py --inverse --kinds c2c_split,c2c_il,c2c_f64 --out $O/sweep_inv.jsonl > $O/sweep_inv.log 2>&1
timeout 900 python tools/sweep.py --sustain ${SUSTAIN:-1.0} --kinds ${SUSTAIN_KINDS:-c2c_split,r2c,c2r} --sizes 16,32,64,128,256,512,1024,2048,4096 --out $O/sweep_sustained.jsonl > $O/sweep_sustained.log 2>&1

timeout 600 $B > $O/plain_bench.log 2>&1 || { echo "plain bench failed"; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file $O/launches.csv $B > $O/ncu_launches.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:^k_ -c 54 --csv --log-file $O/traffic.csv $B > $O/ncu_traffic.log 2>&1

SPECS="${SPECS:-c2c_split:16 c2c_split:64 c2c_split:1024 c2c_split:4096 c2c_split_inv:4096 r2c:1024 r2c:4096 c2r:4096 c2c_f64:1024 stft:1024}"
timeout 300 python tools/prof_one.py $SPECS > $O/plain_prof.log 2>&1 || { echo "plain prof_one failed"; exit 1; }
for s in $SPECS; do
  R=$O/full_${s/:/_}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_ -s 2 -c 1 -f -o $R python tools/prof_one.py $s > $O/ncu_full_${s/:/_}.log 2>&1
  # the reports are 10-40 MB each and gpurun brings back at most 64 MiB: export the two pages the summaries use
  ncu -i $R.ncu-rep --page raw --csv > $R.raw.csv 2>/dev/null
  ncu -i $R.ncu-rep --page source --csv 2>/dev/null | gzip > $R.source.csv.gz
  rm -f $R.ncu-rep
done
ls -la $O | tail -30
