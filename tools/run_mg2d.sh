out=gpurun_out/mg2d
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_multigpu.py -q -x 2>&1 | tail -25 > $out/pytest_multigpu.log
tail -6 $out/pytest_multigpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 2 --warmup 3 > $out/b_v2_q2k_n2_tp.json 2> $out/b_v2_q2k_n2_tp.err
grep "value\|e2e" $out/b_v2_q2k_n2_tp.err | tail -3
DSK_CHECK_SHARDED=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 2 --warmup 3 --workload v2lite --quant q2_k > $out/b_v2lite_q2k_n2_tp.json 2> $out/b_v2lite_q2k_n2_tp.err
grep "value\|e2e\|sharded" $out/b_v2lite_q2k_n2_tp.err | tail -4
