out=gpurun_out/final2
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $out/gpu_tests.log; tail -3 $out/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -3 $out/smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > $out/ncu_bench.json 2> $out/ncu_bench.err
grep -c decode_kernel $out/launches.csv
timeout 600 ncu --set full --clock-control none -k regex:decode_kernel --launch-skip 18 --launch-count 1 -o $out/v2_q2k_full_token -f python tools/prof_token.py v2 q2_k 0 20 > $out/ncu_full.log 2>&1
tail -2 $out/ncu_full.log
ncu -i $out/v2_q2k_full_token.ncu-rep --page raw --csv > $out/full_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/final2/full_raw.csv')))
h,u,v=rows[0],rows[1],rows[2]
for i,x in enumerate(h):
    if x in ('dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum','smsp__issue_active.avg.pct_of_peak_sustained_active'): print(x,u[i],v[i])
PY
