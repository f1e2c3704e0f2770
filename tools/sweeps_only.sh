O=gpurun_out
rm -f $O/sweep_burst.jsonl $O/sweep_inv.jsonl $O/sweep_sustained.jsonl
timeout 1200 python tools/sweep.py --kinds c2c_split,c2c_il,r2c,c2r,c2c_f64,r2c_f64,c2r_f64,stft --sizes 8,16,32,64,128,256,512,1024,2048,4096,8192 --out $O/sweep_burst.jsonl > $O/sweep_burst.log 2>&1
timeout 600 python tools/sweep.py --inverse --kinds c2c_split,c2c_il,c2c_f64 --out $O/sweep_inv.jsonl > $O/sweep_inv.log 2>&1
timeout 900 python tools/sweep.py --sustain 1.0 --kinds c2c_split,r2c,c2r --sizes 16,32,64,128,256,512,1024,2048,4096 --out $O/sweep_sustained.jsonl > $O/sweep_sustained.log 2>&1
