O=gpurun_out
timeout 200 python tools/prof_one.py stft:1024 > $O/ps1.log 2>&1 || exit 1
WFB_STFT_PIPE_MIN_N=64 timeout 200 python tools/prof_one.py stft:1024 > $O/ps2.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_stft -s 2 -c 1 -f -o $O/stft_direct python tools/prof_one.py stft:1024 > $O/ncu_s1.log 2>&1
WFB_STFT_PIPE_MIN_N=64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_stft -s 2 -c 1 -f -o $O/stft_pipe python tools/prof_one.py stft:1024 > $O/ncu_s2.log 2>&1
for R in $O/stft_direct $O/stft_pipe; do ncu -i $R.ncu-rep --page raw --csv > $R.raw.csv 2>/dev/null; ncu -i $R.ncu-rep --page source --csv 2>/dev/null | gzip > $R.source.csv.gz; rm -f $R.ncu-rep; done
