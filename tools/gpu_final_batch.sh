out=gpurun_out/r2h
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -12 > $out/gpu_tests.log; tail -4 $out/gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -2 $out/smoke.log
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/bench.json 2> $out/bench.err; tail -c 2500 $out/bench.json
timeout 300 ncu --set full --clock-control none -k regex:decode_kernel --launch-skip 18 --launch-count 1 -o $out/v2_q2k_full_token -f python tools/prof_token.py v2 q2_k 0 20 > $out/ncu_full.log 2>&1
tail -2 $out/ncu_full.log
ncu -i $out/v2_q2k_full_token.ncu-rep --page raw --csv > $out/full_raw.csv 2>/dev/null
DSK_TSTAMP=1 timeout 200 python tools/prof_token.py v2 q2_k 0 20 mla timeline > $out/mla_timeline.txt 2>&1; tail -25 $out/mla_timeline.txt
