out=gpurun_out/r2i
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -12 > $out/gpu_tests.log; tail -4 $out/gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -3 $out/smoke.log
