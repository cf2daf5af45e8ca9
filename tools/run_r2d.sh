set -x
out=gpurun_out/r2d
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_realdims.py -q --maxfail=12 -s 2>&1 | grep -v "^loading\|^$" | tail -60 > $out/t.log
tail -12 $out/t.log
timeout 500 python bench.py --workload v2 --quant q2_k --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --profile-token > $out/b_v2_q2k.json 2> $out/b_v2_q2k.err
tail -c 300 $out/b_v2_q2k.json
timeout 300 python bench.py --workload v2lite --quant f8e5m2 --steps 2 --warmup 3 --no-cpu-baseline --profile-token > $out/b_v2lite_f8.json 2> $out/b_v2lite_f8.err
tail -c 200 $out/b_v2lite_f8.json
timeout 300 python bench.py --workload v2lite --quant q2_k --steps 2 --warmup 3 --no-cpu-baseline --profile-token > $out/b_v2lite_q2k.json 2> $out/b_v2lite_q2k.err
tail -c 200 $out/b_v2lite_q2k.json
