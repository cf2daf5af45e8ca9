out=gpurun_out/dbg1
mkdir -p $out
python tools/debug_dims.py tiny_v2lite fp32 qk_nope_head_dim=128 qk_rope_head_dim=64 v_head_dim=128 2>&1 | grep -a "^pos" > $out/a.log; cat $out/a.log
python tools/debug_dims.py tiny_v2lite fp32 v_head_dim=128 2>&1 | grep -a "^pos" > $out/b.log; cat $out/b.log
python tools/debug_dims.py tiny_v2lite fp32 qk_nope_head_dim=128 2>&1 | grep -a "^pos" > $out/c.log; cat $out/c.log
python tools/debug_dims.py tiny_v2lite fp32 qk_rope_head_dim=64 2>&1 | grep -a "^pos" > $out/d.log; cat $out/d.log
