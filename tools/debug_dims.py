"""single GPU: per-buffer comparison of one layer against the oracle for a preset with overridden head dims"""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "oracle"), os.path.join(REPO, "deepseek.cpp_b200"), REPO]
import dsk, mint
import oracle as O
dsk.init(0)
preset, quant = sys.argv[1], sys.argv[2]
kw = dict(a.split("=") for a in sys.argv[3:])
kw = {k: int(v) for k, v in kw.items()}
d = f"/dev/shm/dbgd_{preset}_{quant}"
mint.mint(d, preset, quant, fast=True, **kw)
m = dsk.Model.from_dir(d)
o = O.open_session(d)
rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
for pos, tok in enumerate([0, 9]):
    o.copy_embedding(tok); m.copy_embedding(tok)
    for l in range(m.cfg.n_layers):
        m.set_buffer("x", o.buffer("x").copy())
        o.block(l, pos, 0, pos, pos + 1)
        m.block(l, pos, 0, pos, pos + 1)
        c = m.cfg
        nq = c.n_heads * (c.qk_nope_head_dim + c.qk_rope_head_dim)
        print("pos", pos, "layer", l, "x", rel(m.buffer("x"), o.buffer("x")), "q", rel(m.buffer("q"), o.buffer("q")[:nq]),
              "kv_b", rel(m.buffer("kv_b"), o.buffer("kv_b")), "xb2", rel(m.buffer("xb2")[:c.n_heads * c.v_head_dim], o.buffer("xb2")[:c.n_heads * c.v_head_dim]), flush=True)
