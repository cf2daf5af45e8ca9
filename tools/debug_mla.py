"""Localises an MLA mismatch on the GPU box: per-buffer rel-L2 of one block against oracle/_ref after identical inputs."""
import os, sys, tempfile, shutil
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("oracle", "deepseek.cpp_b200", "tests"):
    sys.path.insert(0, os.path.join(REPO, p))
import oracle as O, mint, dsk


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


dsk.init(0)
for preset, quant in [("tiny_v2", "fp32"), ("tiny_v3", "fp32"), ("tiny_v2", "f8e5m2"), ("tiny_v3", "q2_k"), ("tiny_v2", "q3_k")]:
    d = tempfile.mkdtemp(prefix="dbg_mla_")
    kw = dict(v_head_dim=128) if quant == "f8e5m2" else {}
    try:
        mint.mint(d, preset, quant, use_mla=True, fast=True, **kw)
        m = dsk.Model.from_dir(d); o = O.open_session(d)
        c = m.cfg
        nh, L, R, vh = c.n_heads, c.kv_lora_rank, c.qk_rope_head_dim, c.v_head_dim
        for pos, tok in enumerate([3, 77, 512]):
            o.copy_embedding(tok); m.copy_embedding(tok)
            for l in range(2):
                m.set_buffer("x", o.buffer("x").copy())
                for w in (0, 1):
                    m.set_kv_cache(l, w, o.kv_cache(l, w))
                o.block(l, pos, 0, pos, pos + 1); m.block(l, pos, 0, pos, pos + 1)
                out = [f"{preset}/{quant} pos {pos} layer {l}:"]
                out.append(f"kv_a.latent {rel(m.buffer('kv_a')[:L], o.buffer('kv_a')[:L]):.1e}")
                out.append(f"q_rope {rel(m.buffer('q')[:nh * R], o.buffer('q_rope')):.1e}")
                out.append(f"q_c {rel(m.buffer('q_c'), o.buffer('q_c')):.1e}")
                out.append(f"xb2 {rel(m.buffer('xb2')[:nh * L], o.buffer('xb2')[:nh * L]):.1e}")
                out.append(f"kv_b {rel(m.buffer('kv_b')[:nh * vh], o.buffer('kv_b')[:nh * vh]):.1e}")
                for w, width in ((0, L), (1, R)):
                    a = m.kv_cache(l, w)[:(pos + 1) * width].view(np.float16).astype(np.float32)
                    b = o.kv_cache(l, w)[:(pos + 1) * width].view(np.float16).astype(np.float32)
                    out.append(f"cache{w} {rel(a, b):.1e}")
                out.append(f"x {rel(m.buffer('x'), o.buffer('x')):.1e}")
                print(" ".join(out), flush=True)
        m.close(); o.close()
    except Exception as e:   # keep going: one call should tell as much as possible
        print(f"{preset}/{quant}: EXC {type(e).__name__}: {e}", flush=True)
    finally:
        shutil.rmtree(d, ignore_errors=True)
