out=gpurun_out/dbg2
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_model.py -q -x 2>&1 | tail -5 > $out/t.log; cat $out/t.log
python tools/debug_dims.py tiny_v2lite fp32 2>&1 | grep -a "^pos" | head -3
python tools/debug_dims.py tiny_v2lite fp32 n_heads=8 2>&1 | grep -a "^pos" | head -3
python tools/debug_dims.py tiny_v2lite fp32 kv_lora_rank=512 2>&1 | grep -a "^pos" | head -3
