"""N>1 host logic on CPU (gloo, world_size 2): expert-shard placement and the per-layer partial-sum all-reduce
(SURVEY §8(e)).  Each rank evaluates only its own experts of one MoE layer with the oracle port, rank 0 adds the
shared expert, the partials are all-reduced and must equal the single-rank layer update."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as O


def shard(E, rank, n_ranks):
    """Same rule as dsk_model_create: expert e lives on rank e // ceil(E / n_ranks)."""
    per = -(-E // n_ranks)
    first = min(E, rank * per)
    return first, max(0, min(per, E - first))


def test_shard_rule_partitions_all_experts():
    for E in (8, 64, 160, 256, 7):
        for n in (1, 2, 3, 4, 8):
            got = []
            for r in range(n):
                f, c = shard(E, r, n)
                got += list(range(f, f + c))
            assert got == list(range(E))


def moe_update(d, x, rank, n_ranks):
    """x-update of MoE layer 1 restricted to this rank's experts (+ shared expert on rank 0)."""
    s = O.PortSession(d)
    c, L, lay = s.c, s.L, s.layers[1]
    import ctypes as C
    buf = s.buf
    buf["x"][:] = x
    L.ork_rmsnorm(O._fp(buf["xb"]), O._fp(buf["x"]), lay.rms_ffn, c["dim"], C.c_float(c["norm_eps"]))
    gate = O.OrkTensor(0, 0, c["n_routed_experts"], c["dim"], C.cast(lay.moegate, C.c_void_p), None)
    L.ork_matmul(O._fp(buf["moe_weights"]), O._fp(buf["xb"]), C.byref(gate), -1, 0, 0)
    L.ork_moe_gate(O._fp(buf["active_experts_weights"]), lay.moegate_bias, s.active.ctypes.data_as(O.i32p), O._fp(buf["moe_weights"]),
                   c["n_routed_experts"], c["n_active_routed"], c["norm_topk_prob"], C.c_float(c["routed_scaling_factor"]),
                   c["scoring_sigmoid"], c["topk_method"], c["n_group"], c["topk_group"])
    first, count = shard(c["n_routed_experts"], rank, n_ranks)
    mi = c["moe_intermediate_size"]
    part = np.zeros(c["dim"], np.float32)
    for k in range(c["n_active_routed"]):
        e = int(s.active[k])
        if not (first <= e < first + count):
            continue
        L.ork_matmul(O._fp(buf["hb"]), O._fp(buf["xb"]), C.byref(lay.w1), e, c["bs0"], c["bs1"])
        L.ork_matmul(O._fp(buf["hb2"]), O._fp(buf["xb"]), C.byref(lay.w3), e, c["bs0"], c["bs1"])
        h = buf["hb"][:mi]
        buf["hb"][:mi] = (h / (1.0 + np.exp(-h.astype(np.float64)))).astype(np.float32) * buf["hb2"][:mi]
        L.ork_matmul(O._fp(buf["xb2"]), O._fp(buf["hb"]), C.byref(lay.w2), e, c["bs0"], c["bs1"])
        part += buf["xb2"][: c["dim"]] * buf["active_experts_weights"][k]
    if rank == 0 and c["n_shared_experts"] > 0:
        sh = c["n_shared_experts"] * mi
        L.ork_matmul(O._fp(buf["hb"]), O._fp(buf["xb"]), C.byref(lay.sw1), -1, c["bs0"], c["bs1"])
        L.ork_matmul(O._fp(buf["hb2"]), O._fp(buf["xb"]), C.byref(lay.sw3), -1, c["bs0"], c["bs1"])
        h = buf["hb"][:sh]
        buf["hb"][:sh] = (h / (1.0 + np.exp(-h.astype(np.float64)))).astype(np.float32) * buf["hb2"][:sh]
        L.ork_matmul(O._fp(buf["xb2"]), O._fp(buf["hb"]), C.byref(lay.sw2), -1, c["bs0"], c["bs1"])
        part += buf["xb2"][: c["dim"]]
    return part


def _worker(rank, world, port, d, x, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    part = torch.from_numpy(moe_update(d, x, rank, world))
    dist.all_reduce(part, op=dist.ReduceOp.SUM)        # the one collective of the path
    if rank == 0:
        np.save(out_path, part.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_moe_partial_allreduce_equals_full(ckpt, tmp_path):
    d = ckpt("tiny_v2", "f8e5m2")
    rng = np.random.default_rng(4)
    x = rng.standard_normal(512).astype(np.float32)
    full = moe_update(d, x, 0, 1)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "sum.npy")
    mp.spawn(_worker, args=(2, port, d, x, out), nprocs=2, join=True)
    got = np.load(out)
    assert np.allclose(got, full, rtol=1e-5, atol=1e-6)   # summation order differs across ranks (fp32 re-association)


# ---- tensor-parallel shards (round 2): the rules of dsk_model_create and the identities the in-kernel exchanges rely on ----------

def tp_shard(cfg, rank, n_ranks):
    """Same arithmetic as dsk_model_create (csrc/dsk_engine.cu): block splits [r B / N, (r + 1) B / N)."""
    lo = lambda B, r: r * B // n_ranks
    nh = cfg["n_heads"] // n_ranks
    sh = cfg["n_shared_experts"] * cfg["moe_intermediate_size"]
    vb = -(-cfg["vocab_size"] // 128)
    return dict(h0=rank * nh, nh=nh,
                sh0=lo(sh // 256, rank) * 256, sh1=lo(sh // 256, rank + 1) * 256,
                hid0=lo(cfg["hidden_dim"] // 256, rank) * 256, hid1=lo(cfg["hidden_dim"] // 256, rank + 1) * 256,
                v0=lo(vb, rank) * 128, v1=min(cfg["vocab_size"], lo(vb, rank + 1) * 128))


def test_tp_shard_rules_partition_every_dimension():
    cfgs = [dict(n_heads=128, n_shared_experts=2, moe_intermediate_size=1536, hidden_dim=12288, vocab_size=102400),
            dict(n_heads=128, n_shared_experts=1, moe_intermediate_size=2048, hidden_dim=18432, vocab_size=129280),
            dict(n_heads=16, n_shared_experts=2, moe_intermediate_size=1536, hidden_dim=11008, vocab_size=102400),
            dict(n_heads=4, n_shared_experts=1, moe_intermediate_size=256, hidden_dim=768, vocab_size=1024)]
    for cfg in cfgs:
        for n in (2, 4, 8):
            if cfg["n_heads"] % n:
                continue
            heads, sh, hid, voc = [], [], [], []
            for r in range(n):
                s = tp_shard(cfg, r, n)
                heads += list(range(s["h0"], s["h0"] + s["nh"]))
                sh.append((s["sh0"], s["sh1"])); hid.append((s["hid0"], s["hid1"])); voc.append((s["v0"], s["v1"]))
                assert s["sh0"] % 256 == 0 and s["hid0"] % 256 == 0 and s["v0"] % 128 == 0   # K-quant blocks / f8 scale rows stay whole
            assert heads == list(range(cfg["n_heads"]))
            for spans, total in ((sh, cfg["n_shared_experts"] * cfg["moe_intermediate_size"]), (hid, cfg["hidden_dim"]), (voc, cfg["vocab_size"])):
                assert spans[0][0] == 0 and spans[-1][1] == total
                assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))       # contiguous, no gap, no overlap (empty slices allowed)


def _tp_worker(rank, world, port, seed, out_path):
    """Each rank computes, from its slices only, what the tensor-parallel program exchanges: the wo partial sum (its heads'
    columns), the dense-FFN partial sum (its hidden units) and its LM-head rows; the all-reduced / gathered results must equal
    the unsharded layer (fp32 re-association aside)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(seed)
    cfg = dict(n_heads=4, n_shared_experts=1, moe_intermediate_size=256, hidden_dim=768, vocab_size=1024)
    dim, vh = 512, 128
    wo = rng.standard_normal((dim, cfg["n_heads"] * vh)).astype(np.float32) * 0.05
    att = rng.standard_normal(cfg["n_heads"] * vh).astype(np.float32)
    w1 = rng.standard_normal((cfg["hidden_dim"], dim)).astype(np.float32) * 0.05
    w3 = rng.standard_normal((cfg["hidden_dim"], dim)).astype(np.float32) * 0.05
    w2 = rng.standard_normal((dim, cfg["hidden_dim"])).astype(np.float32) * 0.05
    wcls = rng.standard_normal((cfg["vocab_size"], dim)).astype(np.float32) * 0.05
    x = rng.standard_normal(dim).astype(np.float32)
    s = tp_shard(cfg, rank, world)
    c0, c1 = s["h0"] * vh, (s["h0"] + s["nh"]) * vh
    p_wo = torch.from_numpy(wo[:, c0:c1] @ att[c0:c1])                                      # column-sharded wo -> partial sum
    h1, h3 = w1[s["hid0"]:s["hid1"]] @ x, w3[s["hid0"]:s["hid1"]] @ x                       # row-sharded w1 / w3 -> local hidden units
    h = (h1 / (1.0 + np.exp(-h1))) * h3
    p_ffn = torch.from_numpy((w2[:, s["hid0"]:s["hid1"]] @ h).astype(np.float32))           # column-sharded w2 -> partial sum
    logits = torch.zeros(cfg["vocab_size"])
    logits[s["v0"]:s["v1"]] = torch.from_numpy(wcls[s["v0"]:s["v1"]] @ x)                   # row-sharded LM head -> own rows
    key = torch.tensor([float(logits[s["v0"]:s["v1"]].max())])                               # local arg-max key (value part)
    dist.all_reduce(p_wo); dist.all_reduce(p_ffn); dist.all_reduce(logits); dist.all_reduce(key, op=dist.ReduceOp.MAX)
    if rank == 0:
        h1f, h3f = w1 @ x, w3 @ x
        full_ffn = w2 @ ((h1f / (1.0 + np.exp(-h1f))) * h3f)
        ok = (np.allclose(p_wo.numpy(), wo @ att, rtol=1e-4, atol=1e-5) and np.allclose(p_ffn.numpy(), full_ffn, rtol=1e-4, atol=1e-5)
              and np.allclose(logits.numpy(), wcls @ x, rtol=1e-4, atol=1e-5) and abs(float(key) - float((wcls @ x).max())) < 1e-5)
        np.save(out_path, np.array([1.0 if ok else 0.0]))
    dist.barrier()
    dist.destroy_process_group()


def test_tp_partial_sums_equal_full_layer(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "tp_ok.npy")
    mp.spawn(_tp_worker, args=(2, port, 11, out), nprocs=2, join=True)
    assert np.load(out)[0] == 1.0
