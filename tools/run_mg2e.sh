out=gpurun_out/mg2e
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_multigpu.py -q -x -k "_tp" 2>&1 | tail -8 > $out/pytest_tp.log
tail -4 $out/pytest_tp.log
