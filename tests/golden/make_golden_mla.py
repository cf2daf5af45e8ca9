"""Generates tests/golden/e2e_mla.npz from the UNMODIFIED reference (oracle/_ref/libdsref.so): teacher-forced logits, the
latent / rope KV cache rows and the per-block intermediates of true-MLA checkpoints (use_mla=1, BlockMLA src/infer.cpp:
1051-1141) that can be minted without the reference (oracle/mint.py, seeded).

Run in the build container (needs oracle/_ref):   python tests/golden/make_golden_mla.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "oracle"))
import oracle as O  # noqa: E402
import mint  # noqa: E402

CASES = (("tiny_v3", "q2_k", {}), ("tiny_v2", "f8e5m2", {"v_head_dim": 128}), ("tiny_v2", "fp32", {"original_max_position": 5}))
TOKENS = [0, 11, 500, 3, 77, 1023, 64, 9]

if __name__ == "__main__":
    out = {}
    with tempfile.TemporaryDirectory() as td:
        for preset, quant, kw in CASES:
            key = f"{preset}_{quant}"
            d = os.path.join(td, key)
            mint.mint(d, preset, quant, fast=True, seed=78, use_mla=True, **kw)
            s = O.RefSession(d)
            logits = []
            for p, t in enumerate(TOKENS):
                s.forward(t, p)
                logits.append(s.buffer("logits").copy())
            out[key + "_tokens"] = np.array(TOKENS, np.int32)
            out[key + "_logits"] = np.stack(logits)
            out[key + "_q_c_last"] = s.buffer("q_c").copy()          # last layer, last token
            out[key + "_kv_b_last"] = s.buffer("kv_b").copy()
            n = len(TOKENS)
            c = O.config_from_metadata(O.dseek.read_dir(d)[0])
            out[key + "_latent_cache_l0"] = s.kv_cache(0, 0)[:n * c["kv_lora_rank"]].copy()
            out[key + "_rope_cache_l0"] = s.kv_cache(0, 1)[:n * c["qk_rope_head_dim"]].copy()
            s.close()
    np.savez_compressed(os.path.join(HERE, "e2e_mla.npz"), **out)
    print("golden written:", {k: v.shape for k, v in out.items()})
