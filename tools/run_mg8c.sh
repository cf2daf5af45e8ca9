out=gpurun_out/mg8d
mkdir -p $out
run() { name=$1; shift
  env $1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 8 --steps 2 --warmup 3 --profile-token ${@:3} > $out/$name.json 2> $out/$name.err
  grep -a "value\|sharded" $out/$name.err | tail -1
}
run v2_q2k_n8_tp DSK_TP=1 29701
run v3_q2k_n8_tp DSK_TP=1 29703 --workload v3 --quant q2_k
