out=gpurun_out/mg2
mkdir -p $out
nvidia-smi -L > $out/gpus.txt
timeout 900 python -m pytest tests/test_gpu_multigpu.py -q -x 2>&1 | tail -15 > $out/pytest_multigpu.log
tail -5 $out/pytest_multigpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 2 --warmup 3 > $out/b_v2_q2k_n2.json 2> $out/b_v2_q2k_n2.err
tail -c 400 $out/b_v2_q2k_n2.json; tail -3 $out/b_v2_q2k_n2.err
