out=gpurun_out/r2g
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -15 > $out/t_all.log; tail -6 $out/t_all.log
