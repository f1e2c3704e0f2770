# A/B of two builds on the thread-per-row c2c kernels (WFB_LIB selects the build): sustained + burst rates, executed instructions
O=gpurun_out
B=$PWD/wat-fft_b200/libwatfft_b200_B.so
WFB_LIB=$B timeout 300 python -m pytest tests/test_gpu_variants.py -k "c2c_f32" tests/test_gpu_multitile.py -k "c2c_f32_many_tiles" -m gpu -x -q 2>&1 | tail -1
for L in A B; do
  if [ $L = B ]; then export WFB_LIB=$B; else unset WFB_LIB; fi
  for r in 1 2; do
    timeout 200 python tools/sweep.py --sustain 0.6 --kinds c2c_split --sizes 64 --out $O/tp_${L}${r}_s.jsonl > /dev/null 2>&1
    timeout 200 python tools/sweep.py --sustain 0.6 --inverse --kinds c2c_split --sizes 64 --out $O/tp_${L}${r}_si.jsonl > /dev/null 2>&1
    timeout 200 python tools/sweep.py --kinds c2c_split --sizes 64 --out $O/tp_${L}${r}_b.jsonl > /dev/null 2>&1
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
for mode in ('s','si','b'):
    A=[load(f'gpurun_out/tp_A{r}_{mode}.jsonl') for r in (1,2)]; B=[load(f'gpurun_out/tp_B{r}_{mode}.jsonl') for r in (1,2)]
    for k in A[0]:
        if 'tpipe' in k[2]: print(mode, k, 'A', A[0][k], A[1][k], 'B', B[0].get(k), B[1].get(k))
PY
