out=gpurun_out/final3
mkdir -p $out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"decode_kernel|sample_kernel|q3k_repack|add_vec_kernel" -c 400 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > $out/ncu_bench.json 2> $out/ncu_bench.err
grep -c decode_kernel $out/launches.csv
