"""Conditioning of a synthetic checkpoint, measured with the UNMODIFIED reference against itself: the same tokens through the
checkpoint and through a copy whose layer-0 attention-norm weights are perturbed by ~1 ulp (relative 1e-7 gaussian).  The
logits rel-L2 between the two reference runs is the floor any other implementation (different fp32 summation order) can be
held to on that checkpoint.  usage: ref_sensitivity.py <workload> <quant> <mla 0|1> [eps]   (CPU only; needs oracle/_ref)"""
import json, os, shutil, struct, sys, tempfile
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("", "oracle", "tests"):
    sys.path.insert(0, os.path.join(REPO, p))
import bench, oracle as O

wl, quant, mla = sys.argv[1], sys.argv[2], int(sys.argv[3])
eps = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-7
TOKENS = [0, 9, 40011, 33, 100201, 77]
w = bench.workload_cfg(wl, quant, n_layers=2, max_seq_len=64)
w["first_k_dense_replace"] = 1
if mla:
    w["use_mla"] = 1
base = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
d0, d1 = tempfile.mkdtemp(prefix="sens0_", dir=base), tempfile.mkdtemp(prefix="sens1_", dir=base)
try:
    bench.mint_cpu_truncated(w, d0, 2)
    shutil.copy(os.path.join(d0, "shard_000.dseek"), os.path.join(d1, "shard_000.dseek"))
    mm = np.memmap(os.path.join(d1, "shard_000.dseek"), dtype=np.uint8, mode="r+")
    hl = struct.unpack("<Q", mm[:8].tobytes())[0]
    hdr = json.loads(mm[8:8 + hl].tobytes())
    t = hdr["model.layers.0.attn.norm.weight"]
    a, b = t["data_offsets"]
    v = mm[8 + hl + a:8 + hl + b].view(np.float32)
    v *= (1.0 + eps * np.random.default_rng(3).standard_normal(v.size)).astype(np.float32)
    mm.flush(); del mm
    r0, r1 = O.open_session(d0), O.open_session(d1)
    errs = []
    for pos, tok in enumerate(TOKENS):
        r0.forward(tok, pos); r1.forward(tok, pos)
        l0, l1 = r0.buffer("logits"), r1.buffer("logits")
        errs.append(float(np.linalg.norm(l0 - l1) / np.linalg.norm(l0)))
    print(f"{wl}/{quant} mla={mla} eps={eps:g}: reference-vs-perturbed-reference logits rel-L2 per position:",
          " ".join(f"{e:.2e}" for e in errs))
finally:
    shutil.rmtree(d0, ignore_errors=True); shutil.rmtree(d1, ignore_errors=True)
