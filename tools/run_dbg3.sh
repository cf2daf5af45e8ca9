out=gpurun_out/dbg3
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_model.py -q -x -k "teacher_forced and fp32 and tiny_v2lite" 2>&1 | tail -30 > $out/t.log; cat $out/t.log | head -40
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x 2>&1 | tail -5
timeout 300 compute-sanitizer --tool memcheck python tools/debug_dims.py tiny_v2lite fp32 2>&1 | grep -aE "^pos|Invalid|ERROR|at |by " | head -30 > $out/san.log; cat $out/san.log
