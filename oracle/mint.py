"""oracle/mint.py — synthetic `.dseek` checkpoint minting for parity tests and the CPU reference arm.

TEST INFRASTRUCTURE.  No real DeepSeek weights exist in this environment (no network), so every
checkpoint is synthetic: weights N(0,1)/sqrt(fan_in), embedding N(0,1), norm weights 1+0.1 N(0,1),
gate N(0,1)/sqrt(dim)*gate_gain, V3 gate bias 0.01 N(0,1), quantised exactly the way
/root/reference/convert.py does:
  * fp16: cast;  * f8e5m2: per 128x128 block scale=57344/amax, RNE cast, stored scale=1/scale
    (convert.py:216-275);  * q2_k/q3_k: the reference's own quantize_row_q{2,3}_K_ref through
    oracle/_ref/libdsref.so (quantizer.cpp:4-66) — or, with fast=True / no _ref, random *valid* blocks.
Metadata keys are those of convert.py:123-170; tensor names those of src/model.cpp:766-871.
"""
from __future__ import annotations

import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import oracle as O  # noqa: E402
from oracle import dseek  # noqa: E402

PRESETS = {
    # V2-Lite-shaped miniature: no q_lora, softmax/greedy gate, 2 shared experts, 1 dense layer first
    "tiny_v2lite": dict(arch="DeepseekV2ForCausalLM", dim=512, hidden_dim=1024, n_layers=3, n_heads=4,
                        vocab_size=1024, max_seq_len=256, qk_nope_head_dim=64, qk_rope_head_dim=32, v_head_dim=64,
                        kv_lora_rank=256, q_lora_rank=0, first_k_dense_replace=1, n_shared_experts=2,
                        n_routed_experts=8, n_active_routed=3, moe_intermediate_size=256, routed_scaling_factor=1.0,
                        n_group=1, topk_group=1, norm_topk_prob=False, scoring_func="softmax", topk_method="greedy"),
    # V2-236B-shaped miniature: q_lora, group_limited_greedy, scaling 16
    "tiny_v2": dict(arch="DeepseekV2ForCausalLM", dim=512, hidden_dim=768, n_layers=3, n_heads=4,
                    vocab_size=1024, max_seq_len=256, qk_nope_head_dim=64, qk_rope_head_dim=32, v_head_dim=64,
                    kv_lora_rank=256, q_lora_rank=256, first_k_dense_replace=1, n_shared_experts=2,
                    n_routed_experts=16, n_active_routed=4, moe_intermediate_size=256, routed_scaling_factor=16.0,
                    n_group=4, topk_group=2, norm_topk_prob=False, scoring_func="softmax",
                    topk_method="group_limited_greedy"),
    # V3-shaped miniature: sigmoid + bias, norm_topk_prob, interleaved rope, scaling 2.5
    "tiny_v3": dict(arch="DeepseekV3ForCausalLM", dim=512, hidden_dim=768, n_layers=3, n_heads=4,
                    vocab_size=1024, max_seq_len=256, qk_nope_head_dim=64, qk_rope_head_dim=32, v_head_dim=64,
                    kv_lora_rank=256, q_lora_rank=256, first_k_dense_replace=1, n_shared_experts=1,
                    n_routed_experts=16, n_active_routed=4, moe_intermediate_size=256, routed_scaling_factor=2.5,
                    n_group=4, topk_group=2, norm_topk_prob=True, scoring_func="sigmoid",
                    topk_method="group_limited_greedy"),
    # the real V2-Lite shapes (K-quants need the 1408->1536 / 10944->11008 zero padding)
    "v2lite": dict(arch="DeepseekV2ForCausalLM", dim=2048, hidden_dim=10944, n_layers=27, n_heads=16,
                   vocab_size=102400, max_seq_len=1024, qk_nope_head_dim=128, qk_rope_head_dim=64, v_head_dim=128,
                   kv_lora_rank=512, q_lora_rank=0, first_k_dense_replace=1, n_shared_experts=2,
                   n_routed_experts=64, n_active_routed=6, moe_intermediate_size=1408, routed_scaling_factor=1.0,
                   n_group=1, topk_group=1, norm_topk_prob=False, scoring_func="softmax", topk_method="greedy"),
}


def pad_for_kquant(cfg: dict) -> dict:
    c = dict(cfg)
    up = lambda v: (v + 255) // 256 * 256
    c["moe_intermediate_size"] = up(c["moe_intermediate_size"])
    c["hidden_dim"] = up(c["hidden_dim"])
    return c


def f8e5m2_blockwise(w: np.ndarray, bs=(128, 128)):
    """convert.py:216-275 vectorised.  Returns (uint8 payload, fp32 stored scales)."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(w, dtype=np.float32))
    R, Cc = t.shape
    Rp, Cp = -(-R // bs[0]) * bs[0], -(-Cc // bs[1]) * bs[1]
    tp = torch.zeros(Rp, Cp)
    tp[:R, :Cc] = t
    blk = tp.view(Rp // bs[0], bs[0], Cp // bs[1], bs[1])
    # amax over the *real* elements of each block (padding is zero so it never wins)
    amax = blk.abs().amax(dim=(1, 3))
    scale = 57344.0 / amax.clamp(min=1e-12)
    q = (blk * scale[:, None, :, None]).clamp(min=-57344.0, max=57344.0).to(torch.float8_e5m2)
    q = q.view(Rp, Cp)[:R, :Cc].contiguous()
    return q.view(torch.uint8).numpy(), scale.float().reciprocal().numpy().astype(np.float32)


def kquant_rows(w: np.ndarray, quant: str, fast: bool, rng: np.random.Generator) -> np.ndarray:
    rows, cols = w.shape
    bb = O.BLOCK_BYTES[quant]
    out = np.zeros((rows, cols // 256 * bb), dtype=np.uint8)
    L = O.ref_lib()
    if L is not None and not fast:
        w = np.ascontiguousarray(w, dtype=np.float32)
        L.ref_quantize_rows(O._fp(w), O._vp(out), rows, cols, 1 if quant == "q3_k" else 0)
        return out
    # random valid blocks: any byte pattern with finite fp16 d/dmin is a valid block
    out[:] = rng.integers(0, 256, size=out.shape, dtype=np.uint8)
    blk = out.reshape(rows, cols // 256, bb)
    scale = np.float16(1.0 / np.sqrt(cols) / 6.0)
    if quant == "q2_k":
        blk[:, :, 80:82] = np.frombuffer(scale.tobytes(), dtype=np.uint8)
        blk[:, :, 82:84] = np.frombuffer(np.float16(scale * 1.5).tobytes(), dtype=np.uint8)
    else:
        blk[:, :, 108:110] = np.frombuffer(np.float16(scale / 8).tobytes(), dtype=np.uint8)
    return out


def tokens_tensor(vocab_size: int) -> np.ndarray:
    """ids 2..257 are the 256 single bytes (NUL stored as \\x07, convert.py:210); fillers never match text."""
    toks = [b"<unk>", b"<s>"]
    for b in range(256):
        toks.append(bytes([7]) if b == 0 else bytes([b]))
    for i in range(258, vocab_size):
        toks.append(b"\xff\xfe#" + str(i).encode())
    toks = toks[:vocab_size]
    return np.frombuffer(b"\0".join(toks) + b"\0", dtype=np.uint8).copy()


def mint(dirname: str, preset: str = "tiny_v2lite", quant: str = "fp32", seed: int = 1234, fast: bool = False,
         gate_gain: float = 4.0, original_max_position: int = 4096, use_mla: bool = False, **overrides) -> dict:
    """use_mla: BlockMLA tensors as convert.py --mla writes them (convert.py:384-: wc, wq_rope_b, wv_b instead of wq_b, wkv_b)."""
    cfg = dict(PRESETS[preset])
    cfg.update(overrides)
    if quant in ("q2_k", "q3_k"):
        cfg = pad_for_kquant(cfg)
    rng = np.random.default_rng(seed)
    if os.path.isdir(dirname):
        shutil.rmtree(dirname)
    os.makedirs(dirname)
    dim, nh = cfg["dim"], cfg["n_heads"]
    nope, rope, vh = cfg["qk_nope_head_dim"], cfg["qk_rope_head_dim"], cfg["v_head_dim"]
    hd = nope + rope
    md = {
        "arch": cfg["arch"], "use_mla": "1" if use_mla else "0", "quant": quant, "dim": dim, "hidden_dim": cfg["hidden_dim"],
        "n_layers": cfg["n_layers"], "n_heads": nh, "vocab_size": cfg["vocab_size"], "max_seq_len": cfg["max_seq_len"],
        "bos_token_id": 0, "eos_token_id": 1, "rope_theta": 10000.0, "norm_eps": 1e-6, "norm_type": "rmsnorm",
        "act_type": "silu", "first_k_dense_replace": cfg["first_k_dense_replace"],
        "kv_lora_rank": cfg["kv_lora_rank"], "q_lora_rank": cfg["q_lora_rank"], "qk_nope_head_dim": nope,
        "qk_rope_head_dim": rope, "v_head_dim": vh, "n_shared_experts": cfg["n_shared_experts"],
        "n_routed_experts": cfg["n_routed_experts"], "n_active_routed": cfg["n_active_routed"],
        "moe_intermediate_size": cfg["moe_intermediate_size"], "routed_scaling_factor": cfg["routed_scaling_factor"],
        "n_group": cfg["n_group"], "norm_topk_prob": str(bool(cfg["norm_topk_prob"])),
        "scoring_func": cfg["scoring_func"], "topk_group": cfg["topk_group"], "topk_method": cfg["topk_method"],
        "rope_scaling_beta_fast": 32, "rope_scaling_beta_slow": 1, "rope_scaling_factor": 40.0,
        "rope_scaling_mscale": 1.0, "rope_scaling_mscale_all_dim": 1.0,
        "rope_scaling_original_max_position_embeddings": original_max_position,
    }
    if quant == "f8e5m2":
        md["quantization_block_size_0"] = 128
        md["quantization_block_size_1"] = 128

    def randw(rows, cols, real_rows=None, real_cols=None, std=None):
        """N(0,1)/sqrt(fan_in) with optional zero padding of trailing rows/cols (K-quant padding)."""
        rr, rc = real_rows or rows, real_cols or cols
        w = np.zeros((rows, cols), dtype=np.float32)
        w[:rr, :rc] = rng.standard_normal((rr, rc), dtype=np.float32) * (std if std is not None else rc ** -0.5)
        return w

    def put(out, name, w):
        """w: (rows, cols) or (E, rows, cols) fp32 -> quantised tensor(s) under `name`.weight[/.scale]."""
        if quant == "fp32":
            out[name + ".weight"] = ("F32", np.ascontiguousarray(w, dtype=np.float32))
        elif quant == "fp16":
            out[name + ".weight"] = ("F16", np.ascontiguousarray(w).astype(np.float16))
        elif quant == "f8e5m2":
            if w.ndim == 3:
                qs, ss = zip(*(f8e5m2_blockwise(w[e]) for e in range(w.shape[0])))
                out[name + ".weight"] = ("F8_E5M2", np.stack(qs))
                out[name + ".scale"] = ("F32", np.stack(ss))
            else:
                q, s = f8e5m2_blockwise(w)
                out[name + ".weight"] = ("F8_E5M2", q)
                out[name + ".scale"] = ("F32", s)
        else:
            if w.ndim == 3:
                out[name + ".weight"] = ("U8", np.stack([kquant_rows(w[e], quant, fast, rng) for e in range(w.shape[0])]))
            else:
                out[name + ".weight"] = ("U8", kquant_rows(w, quant, fast, rng))

    def norm_w(n):
        return ("F32", (1.0 + 0.1 * rng.standard_normal(n)).astype(np.float32))

    base = PRESETS[preset]
    real_mi = overrides.get("moe_intermediate_size", base["moe_intermediate_size"])
    real_hidden = overrides.get("hidden_dim", base["hidden_dim"])
    mi, hidden = cfg["moe_intermediate_size"], cfg["hidden_dim"]
    E, ns = cfg["n_routed_experts"], cfg["n_shared_experts"]

    shard_idx, shard = 0, {}
    shard["tokenizer.tokens"] = ("U8", tokens_tensor(cfg["vocab_size"]))
    put(shard, "model.embed", rng.standard_normal((cfg["vocab_size"], dim), dtype=np.float32))
    first_md = md
    for l in range(cfg["n_layers"]):
        if l % 8 == 0 and l > 0:  # convert.py:374-377: new shard every 8 layers
            dseek.write_shard(os.path.join(dirname, f"shard_{shard_idx:03d}.dseek"), shard, first_md)
            first_md, shard, shard_idx = None, {}, shard_idx + 1
        p = f"model.layers.{l}."
        shard[p + "attn.norm.weight"] = norm_w(dim)
        shard[p + "mlp.norm.weight"] = norm_w(dim)
        shard[p + "attn.kv_a_norm.weight"] = norm_w(cfg["kv_lora_rank"])
        if use_mla:
            assert cfg["q_lora_rank"] > 0, "BlockMLA requires q_lora_rank > 0 (src/infer.cpp:1057)"
            shard[p + "attn.q_a_norm.weight"] = norm_w(cfg["q_lora_rank"])
            put(shard, p + "attn.wq_a", randw(cfg["q_lora_rank"], dim))
            # wc = k_nope_b^T . q_nope_b absorbed: (n_heads * kv_lora_rank, q_lora_rank); its entries carry two fan-ins
            put(shard, p + "attn.wc", randw(nh * cfg["kv_lora_rank"], cfg["q_lora_rank"], std=(cfg["q_lora_rank"] * cfg["kv_lora_rank"]) ** -0.5 * nope ** 0.5))
            put(shard, p + "attn.wq_rope_b", randw(nh * rope, cfg["q_lora_rank"]))
            put(shard, p + "attn.wkv_a", randw(cfg["kv_lora_rank"] + rope, dim))
            put(shard, p + "attn.wv_b", randw(nh * vh, cfg["kv_lora_rank"]))
        else:
            if cfg["q_lora_rank"] > 0:
                shard[p + "attn.q_a_norm.weight"] = norm_w(cfg["q_lora_rank"])
                put(shard, p + "attn.wq_a", randw(cfg["q_lora_rank"], dim))
                put(shard, p + "attn.wq_b", randw(nh * hd, cfg["q_lora_rank"]))
            else:
                put(shard, p + "attn.wq", randw(nh * hd, dim))
            put(shard, p + "attn.wkv_a", randw(cfg["kv_lora_rank"] + rope, dim))
            put(shard, p + "attn.wkv_b", randw(nh * (nope + vh), cfg["kv_lora_rank"]))
        put(shard, p + "attn.wo", randw(dim, nh * vh))
        if E > 0 and l >= cfg["first_k_dense_replace"]:
            shard[p + "moegate.weight"] = ("F32", (rng.standard_normal((E, dim), dtype=np.float32) * dim ** -0.5 * gate_gain))
            if cfg["arch"] == "DeepseekV3ForCausalLM":
                shard[p + "moegate.bias"] = ("F32", (0.01 * rng.standard_normal(E)).astype(np.float32))
            put(shard, p + "mlp.w1", np.stack([randw(mi, dim, real_rows=real_mi) for _ in range(E)]))
            put(shard, p + "mlp.w2", np.stack([randw(dim, mi, real_cols=real_mi) for _ in range(E)]))
            put(shard, p + "mlp.w3", np.stack([randw(mi, dim, real_rows=real_mi) for _ in range(E)]))
            if ns > 0:
                put(shard, p + "shared_mlp.w1", randw(ns * mi, dim, real_rows=ns * real_mi))
                put(shard, p + "shared_mlp.w2", randw(dim, ns * mi, real_cols=ns * real_mi))
                put(shard, p + "shared_mlp.w3", randw(ns * mi, dim, real_rows=ns * real_mi))
        else:
            put(shard, p + "mlp.w1", randw(hidden, dim, real_rows=real_hidden))
            put(shard, p + "mlp.w2", randw(dim, hidden, real_cols=real_hidden))
            put(shard, p + "mlp.w3", randw(hidden, dim, real_rows=real_hidden))
    shard["model.norm.weight"] = norm_w(dim)
    put(shard, "model.output", randw(cfg["vocab_size"], dim))
    dseek.write_shard(os.path.join(dirname, f"shard_{shard_idx:03d}.dseek"), shard, first_md)
    return cfg


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("dir")
    ap.add_argument("--preset", default="tiny_v2lite")
    ap.add_argument("--quant", default="fp32")
    ap.add_argument("--n-layers", type=int, default=None)
    ap.add_argument("--fast", action="store_true")
    a = ap.parse_args()
    kw = {}
    if a.n_layers:
        kw["n_layers"] = a.n_layers
    print(mint(a.dir, a.preset, a.quant, fast=a.fast, **kw))
