out=gpurun_out/mg2c
mkdir -p $out
DSK_TP=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 tools/debug_tp.py tiny_v2lite fp32 > $out/dbg_fp32_tp0.log 2>&1
grep -aE "^0 .*(sharding|layer|logits)" $out/dbg_fp32_tp0.log | head -12
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 tools/debug_tp.py tiny_v2lite q2_k > $out/dbg_q2k.log 2>&1
grep -aE "^0 .*(sharding|layer|logits)" $out/dbg_q2k.log | head -12
