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
