# A = grouped dense side (default), B = -DWFB_REAL_DENSE=1: tests on A, sustained + burst rates and bank-conflict counters on both
O=gpurun_out
B=$PWD/wat-fft_b200/libwatfft_b200_B.so
timeout 300 python -m pytest tests/test_gpu_variants.py -k "real" tests/test_gpu_multitile.py -k "real" tests/test_gpu_parity.py -k "r2c or c2r" tests/test_gpu_fixtures.py -m gpu -x -q 2>&1 | tail -1
for L in A B; do
  if [ $L = B ]; then export WFB_LIB=$B; else unset WFB_LIB; fi
  for r in 1 2; do
    timeout 200 python tools/sweep.py --sustain 0.5 --kinds r2c,c2r --sizes 256 --out $O/rd_${L}${r}_s.jsonl > /dev/null 2>&1
    timeout 200 python tools/sweep.py --sustain 0.5 --kinds r2c_f64,c2r_f64 --sizes 128 --out $O/rd_${L}${r}_d.jsonl > /dev/null 2>&1
    timeout 200 python tools/sweep.py --kinds r2c,c2r --sizes 256 --out $O/rd_${L}${r}_b.jsonl > /dev/null 2>&1
  done
  timeout 100 ncu --metrics l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,gpu__time_duration.sum --clock-control none -k regex:^k_ -s 2 -c 1 --csv --log-file $O/rd_$L.csv python tools/prof_one.py r2c:256 > /dev/null 2>&1
  grep -E "bank_conflicts|wavefronts_mem_shared|time_duration" $O/rd_$L.csv | awk -F'","' '{print "'$L' r2c:256", $(NF-2), $NF}'
done
unset WFB_LIB
python - <<'PY'
import json
def load(f):
    d={}
    for l in open(f):
        if l.startswith('{'):
            r=json.loads(l); d.setdefault((r['kind'],r['n']), (r['variant'], r['frac']))
    return d
for mode in ('s','d','b'):
    A=[load(f'gpurun_out/rd_A{r}_{mode}.jsonl') for r in (1,2)]; B=[load(f'gpurun_out/rd_B{r}_{mode}.jsonl') for r in (1,2)]
    for k in A[0]: print(mode, k, A[0][k][0], 'A(grouped)', A[0][k][1], A[1][k][1], 'B(dense)', B[0][k][1], B[1][k][1])
PY
