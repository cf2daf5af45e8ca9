out=gpurun_out/mg2b
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_multigpu.py -q -x -k "_tp" 2>&1 | tail -25 > $out/pytest_tp.log
tail -12 $out/pytest_tp.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 2 --warmup 3 > $out/b_v2_q2k_n2_tp.json 2> $out/b_v2_q2k_n2_tp.err
tail -c 300 $out/b_v2_q2k_n2_tp.json; grep "value\|e2e\|Error\|error" $out/b_v2_q2k_n2_tp.err | tail -5
