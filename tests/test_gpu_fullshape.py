"""GPU parity at BASELINE.json's real shapes (DeepSeek-V2-Lite dims: dim 2048, 16 heads, 64 routed experts top-6 + 2
shared, moe_intermediate 1408, vocab 102400), depth truncated to 1 dense + 2 MoE layers so the CPU checker finishes in
seconds.  The tiny presets of test_gpu_model.py exercise every branch of the algorithm; this file exercises the
*production tile shapes* of the persistent interpreter (warp-per-tile tensor-core F8 rows of 2064 B, register-resident
K-quant activations, multi-piece DOWN accumulation, 102400-row LM head + on-device argmax) against the unmodified
reference (oracle/_ref) or, where that is absent, the C restatement."""
import os
import shutil
import tempfile

import numpy as np
import pytest

import oracle as O
from conftest import rel_l2

pytestmark = pytest.mark.gpu

# same bounds as the tiny end-to-end test: fp32 summation order (F8), reference-vs-reference K-quant floor
TOL = {"f8e5m2": 5e-4, "q2_k": 8e-2, "q3_k": 8e-2}
TOKENS = [0, 9, 40011, 33, 100201, 77, 5, 64000]


@pytest.fixture(scope="module")
def dsk():
    import dsk as d
    d.init(0)
    return d


@pytest.mark.parametrize("quant", ["f8e5m2", "q2_k", "q3_k"])
def test_v2lite_shapes_teacher_forced(dsk, quant):
    import bench
    w = bench.workload_cfg("v2lite", quant, n_layers=3, max_seq_len=64)
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
    d = tempfile.mkdtemp(prefix="dsk_fullshape_", dir=base)
    try:
        bench.mint_cpu_truncated(w, d, 3)
        m = dsk.Model.from_dir(d)
        o = O.open_session(d)
        worst = 0.0
        for pos, tok in enumerate(TOKENS):
            logits, am = m.forward(tok, pos)
            o.forward(tok, pos)
            exp = o.buffer("logits")
            assert np.all(np.isfinite(logits))
            err = rel_l2(logits, exp)
            worst = max(worst, err)
            assert err < TOL[quant], (quant, pos, err)
            assert am == int(np.argmax(logits))          # on-device argmax over 102400 rows, lowest index on ties
            top2 = np.sort(exp)[-2:]
            if top2[1] - top2[0] > 2 * np.max(np.abs(logits - exp)):
                assert am == o.argmax()
            if quant == "f8e5m2":
                assert sorted(m.active_experts().tolist()) == sorted(o.active_experts().tolist())
        print(f"full-shape {quant}: worst logits rel-L2 {worst:.3e}")
        # the device-resident greedy loop must reproduce the host-driven loop token for token at these shapes
        pos = len(TOKENS)
        m2 = dsk.Model.from_dir(d)
        for p, t in enumerate(TOKENS):
            _, am2 = m2.forward(t, p, want_logits=False)
        host = []
        for _ in range(8):
            host.append(am)
            _, am = m.forward(am, pos, want_logits=False)
            pos += 1
        dev, ms = m2.decode_greedy(len(TOKENS), 8)
        assert dev.tolist() == host and ms > 0
        m.close(); m2.close(); o.close()
    finally:
        shutil.rmtree(d, ignore_errors=True)
