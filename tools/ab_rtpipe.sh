# A/B of two builds on the thread-per-row real kernels (WFB_LIB selects the build)
O=gpurun_out
B=$PWD/wat-fft_b200/libwatfft_b200_B.so
WFB_LIB=$B timeout 300 python -m pytest tests/test_gpu_variants.py -k "real" tests/test_gpu_multitile.py -k "real" tests/test_gpu_parity.py -k "r2c or c2r" -m gpu -x -q 2>&1 | tail -1
for L in A B; do
  if [ $L = B ]; then export WFB_LIB=$B; else unset WFB_LIB; fi
  for r in 1 2; do
    timeout 200 python tools/sweep.py --sustain 0.5 --kinds r2c,c2r --sizes 64,128 --out $O/rt_${L}${r}_s.jsonl > /dev/null 2>&1
    timeout 200 python tools/sweep.py --sustain 0.5 --kinds r2c_f64,c2r_f64 --sizes 16,32,64 --out $O/rt_${L}${r}_d.jsonl > /dev/null 2>&1
  done
done
unset WFB_LIB
python - <<'PY'
import json
def load(f):
    d={}
    for l in open(f):
        if l.startswith('{'):
            r=json.loads(l); d[(r['kind'],r['n'],r['variant'])]=r['frac']
    return d
for mode in ('s','d'):
    A=[load(f'gpurun_out/rt_A{r}_{mode}.jsonl') for r in (1,2)]; B=[load(f'gpurun_out/rt_B{r}_{mode}.jsonl') for r in (1,2)]
    for k in A[0]:
        if 'rtpipe' in k[2]: print(mode, k, 'A', A[0][k], A[1][k], 'B', B[0].get(k), B[1].get(k))
PY
