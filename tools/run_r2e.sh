set -x
out=gpurun_out/r2e
mkdir -p $out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:decode_kernel --launch-skip 20 --launch-count 1 -o $out/v2_q2k_token -f python tools/prof_token.py v2 q2_k 8 24 > $out/ncu.log 2>&1
tail -5 $out/ncu.log
ls -la $out
