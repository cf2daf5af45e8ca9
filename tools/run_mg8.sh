out=gpurun_out/mg8
mkdir -p $out
nvidia-smi -L > $out/gpus.txt
run() { # name, env, args
  name=$1; shift
  env $1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 8 --steps 2 --warmup 3 --profile-token ${@:3} > $out/$name.json 2> $out/$name.err
  grep -a "value\|e2e " $out/$name.err | tail -2
}
run v2_q2k_n8_tp DSK_TP=1 29701
run v2_q2k_n8_ep DSK_TP=0 29702
run v3_q2k_n8_tp DSK_TP=1 29703 --workload v3 --quant q2_k
run v3_q2k_n8_ep DSK_TP=0 29704 --workload v3 --quant q2_k
