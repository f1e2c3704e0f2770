# A/B/A/B sweep of two builds of the library (tools/sweep.py; WFB_LIB selects the build)
set -e
S="${AB_SWEEP:---kinds c2c_split,c2c_il,r2c,c2r,c2c_f64,r2c_f64 --sizes 64,128,256,512,1024,2048,4096,8192}"
for r in 1 2; do
  timeout 500 python tools/sweep.py $S --out gpurun_out/ab_A$r.jsonl > gpurun_out/ab_A$r.log 2>&1
  WFB_LIB=$PWD/wat-fft_b200/libwatfft_b200_B.so timeout 500 python tools/sweep.py $S --out gpurun_out/ab_B$r.jsonl > gpurun_out/ab_B$r.log 2>&1
done
