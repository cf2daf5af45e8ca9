"""Generates tests/golden/*.npz|json from the UNMODIFIED reference (oracle/_ref/libdsref.so).

Run in the build container (needs oracle/_ref, i.e. `make -C oracle ref` with /root/reference present):
    python tests/golden/make_golden.py
The fixtures are committed; they pin the oracle port (CPU tests) and the CUDA kernels (GPU tests) on
machines where the reference cannot be run.  Inputs are seeded; nothing here reads real model weights.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "oracle"))
import oracle as O  # noqa: E402
import mint  # noqa: E402

R = O.Ops("ref")
rng = np.random.default_rng(20240924)
out = {}

# ---- known answers carried by the reference's own test (src/test.cpp:132-165) --------------------
x16 = np.array([2.0624e-01, 1.6975e+00, 8.4918e-01, -1.7186e-01, -9.0164e-01, 6.1108e-01, 2.2116e-01, 1.0412e+00,
                -1.6616e-03, 8.2840e-01, 2.2667e-01, -1.3993e+00, 4.1013e-01, -1.2223e+00, 2.2723e-01, 6.3558e-01],
               dtype=np.float32)
w16 = np.array([-1.1210, -0.0235, -1.3527, 0.6300, 0.2566, -0.4517, -0.3528, 0.4422, -0.4032, -1.0949, -0.7834, 1.1425,
                0.6263, -0.3680, 0.3226, -0.2984, 0.1176, -1.1462, -0.8181, -2.0047, 0.0932, 1.4665, -0.8682, -0.8490,
                -1.3017, -1.0068, -0.2890, 0.0167, 1.1607, 0.7196, 1.7701, 0.2891], dtype=np.float32).reshape(2, 16)
w16_f16 = w16.astype(np.float16)
w16_f8 = (w16_f16.view(np.uint16) >> 8).astype(np.uint8)  # float_to_float8e5m2 truncates (src/codec.h:49-58)
kat = {
    "testcpp_x": x16.tolist(), "testcpp_w": w16.tolist(),
    "testcpp_expect_f32_f16": [-3.7454, -3.2738],  # src/test.cpp:150-160, tol 1e-4 / 1e-3
    "testcpp_ref_f32": R.matmul(x16, w16, "fp32", 2, 16).tolist(),
    "testcpp_ref_f16": R.matmul(x16, w16_f16, "fp16", 2, 16).tolist(),
    "testcpp_ref_f8": R.matmul(x16, w16_f8, "f8e5m2", 2, 16).tolist(),
    # SURVEY §8(c) known answers produced from the reference's static functions
    "rmsnorm_in": [1, -2, 3, -4, 5, -6, 7, -8], "rmsnorm_w": [1, 1, 1, 1, 2, 2, 2, 2], "rmsnorm_eps": 1e-6,
    "rope_in": [1, 2, 3, 4, 5, 6, 7, 8], "rope_pos": 3, "rope_theta": 1e4,
    "gate_logits": [0.1, 2, -1, 2, 0.5, 1.5, -0.3, 0],
}
kat["rmsnorm_out"] = R.rmsnorm(np.array(kat["rmsnorm_in"], np.float32), np.array(kat["rmsnorm_w"], np.float32), 1e-6).tolist()
kat["rope_v2_out"] = R.rope(np.array(kat["rope_in"], np.float32), 8, 3, 1e4, False).tolist()
kat["rope_v3_out"] = R.rope(np.array(kat["rope_in"], np.float32), 8, 3, 1e4, True).tolist()
i1, w1, _ = R.moe_gate(np.array(kat["gate_logits"], np.float32), None, 3, False, 2.0, False, 0, 1, 1)
i2, w2, _ = R.moe_gate(np.array(kat["gate_logits"], np.float32), None, 3, True, 1.0, True, 1, 4, 1)
kat["gate_softmax_greedy"] = {"idx": i1.tolist(), "w": w1.tolist()}
kat["gate_sigmoid_group"] = {"idx": i2.tolist(), "w": w2.tolist()}
kat["silu_1p5"] = R.silu(1.5)
json.dump(kat, open(os.path.join(HERE, "kat.json"), "w"), indent=1)

# ---- op-level vectors ----------------------------------------------------------------------------
ops = {}
x = (rng.standard_normal(1024) * 3).astype(np.float32)
x[256:512] = 0  # an all-zero block (d = 0 path)
x[512] = -x[513] if abs(x[513]) > abs(x[512]) else x[512]  # a +/- tie on |x| near the front of block 2
ops["q8k_x"] = x
ops["q8k_blocks"] = R.quantize_q8k(x)
for quant in ("q2_k", "q3_k"):
    d, n = 24, 768
    w = (rng.standard_normal((d, n)) * n ** -0.5).astype(np.float32)
    wb = mint.kquant_rows(w, quant, False, rng)
    xv = rng.standard_normal(n).astype(np.float32)
    ops[f"{quant}_w"] = wb
    ops[f"{quant}_x"] = xv
    ops[f"{quant}_out"] = R.matmul(xv, wb, quant, d, n)
    ops[f"{quant}_deq_row0"] = R.dequantize(wb[0], quant, n)
d, n = 200, 384  # ragged vs the 128x128 scale blocks
w = (rng.standard_normal((d, n)) * n ** -0.5).astype(np.float32)
q8, sc = mint.f8e5m2_blockwise(w)
xv = rng.standard_normal(n).astype(np.float32)
ops["f8_w"], ops["f8_scale"], ops["f8_x"] = q8, sc, xv
ops["f8_out"] = R.matmul(xv, q8, "f8e5m2", d, n, sc)
w = (rng.standard_normal((40, 256)) * 0.1).astype(np.float32)
xv = rng.standard_normal(256).astype(np.float32)
ops["f16_w"], ops["f16_x"] = w.astype(np.float16), xv
ops["f16_out"] = R.matmul(xv, w.astype(np.float16), "fp16", 40, 256)
ops["f32_w"] = w
ops["f32_out"] = R.matmul(xv, w, "fp32", 40, 256)
# gate: V2-Lite (64/6 softmax greedy), V2 (160/6 group-limited 8/3, x16), V3 (256/8 sigmoid+bias 8/4, norm, x2.5)
for name, (E, K, sig, method, ng, tg, norm, scale) in {
        "v2lite": (64, 6, 0, 0, 1, 1, 0, 1.0), "v2": (160, 6, 0, 1, 8, 3, 0, 16.0), "v3": (256, 8, 1, 1, 8, 4, 1, 2.5)}.items():
    lg = (rng.standard_normal(E) * 2).astype(np.float32)
    bias = (0.01 * rng.standard_normal(E)).astype(np.float32) if sig else None
    idx, wts, sc_out = R.moe_gate(lg, bias, K, norm, scale, sig, method, ng, tg)
    ops[f"gate_{name}_logits"] = lg
    if bias is not None:
        ops[f"gate_{name}_bias"] = bias
    ops[f"gate_{name}_idx"], ops[f"gate_{name}_w"], ops[f"gate_{name}_scores"] = idx, wts, sc_out
# rope at a large position, both layouts, 64 rotary dims
v = rng.standard_normal(64).astype(np.float32)
ops["rope_x"] = v
ops["rope_v2_p1234"] = R.rope(v, 64, 1234, 1e4, False)
ops["rope_v3_p1234"] = R.rope(v, 64, 1234, 1e4, True)
vh = rng.standard_normal(64).astype(np.float16).view(np.uint16)
ops["rope16_x"] = vh
ops["rope16_v2_p1"] = R.rope_f16(vh, 64, 1, 1e4, False)
ops["rope16_v3_p1"] = R.rope_f16(vh, 64, 1, 1e4, True)
# attention: 3 heads x (hd 48, vh 32) x 37 positions
nh, hd, vhd, T = 3, 48, 32, 37
q = rng.standard_normal(nh * hd).astype(np.float32)
kc = rng.standard_normal(T * nh * hd).astype(np.float16).view(np.uint16)
vc = rng.standard_normal(T * nh * vhd).astype(np.float16).view(np.uint16)
ops["attn_q"], ops["attn_k"], ops["attn_v"] = q, kc, vc
ops["attn_out"] = np.concatenate([R.attn(q[h * hd:(h + 1) * hd], kc[h * hd:], vc[h * vhd:], hd, vhd, nh, T) for h in range(nh)])
np.savez_compressed(os.path.join(HERE, "ops.npz"), **ops)

# ---- end-to-end: tiny V3-shaped Q2_K checkpoint of random valid blocks (mintable without the reference) ----
import tempfile
e2e = {}
with tempfile.TemporaryDirectory() as td:
    for preset, quant in (("tiny_v3", "q2_k"), ("tiny_v2lite", "f8e5m2")):
        d = os.path.join(td, f"{preset}_{quant}")
        mint.mint(d, preset, quant, fast=True, seed=77)
        s = O.RefSession(d)
        toks = [0, 11, 500, 3, 77, 1023]
        logits = []
        for p, t in enumerate(toks):
            s.forward(t, p)
            logits.append(s.buffer("logits").copy())
        e2e[f"{preset}_{quant}_tokens"] = np.array(toks, np.int32)
        e2e[f"{preset}_{quant}_logits"] = np.stack(logits)
        e2e[f"{preset}_{quant}_experts_last"] = s.active_experts()
        s.close()
np.savez_compressed(os.path.join(HERE, "e2e.npz"), **e2e)
print("golden written:", os.listdir(HERE))
