out=gpurun_out/mg2f
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_multigpu.py -q -x -k "(tiny_v2_tp-f8e5m2 or tiny_v3_tp-q2_k or tiny_v2lite-fp32) and not -0]" 2>&1 | tail -6 > $out/pytest_tp.log
tail -3 $out/pytest_tp.log
