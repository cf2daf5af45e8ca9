"""torchrun --nproc-per-node 2 tools/debug_tp.py <preset> <quant>: per-layer residual-stream error of the sharded engine vs the oracle."""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "oracle"), os.path.join(REPO, "deepseek.cpp_b200"), REPO]
import torch, torch.distributed as dist
import dsk, mint
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
dsk.init(rank)
preset, quant = sys.argv[1], sys.argv[2]
d = f"/dev/shm/dbg_{preset}_{quant}"
if rank == 0:
    mint.mint(d, preset, quant, fast=True, qk_nope_head_dim=128, qk_rope_head_dim=64, v_head_dim=128)
dist.barrier()
m = dsk.Model.from_dir(d, rank=rank, n_ranks=world, device=rank)
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid = torch.frombuffer(bytearray(dsk.Model.comm_unique_id()), dtype=torch.uint8).cuda()
dist.broadcast(uid, 0)
m.comm_init(bytes(uid.cpu().numpy().tobytes()))
print(rank, "sharding", m.sharding(), flush=True)
import oracle as O
o = O.open_session(d)
for pos, tok in enumerate([0, 9, 400]):
    o.copy_embedding(tok); m.copy_embedding(tok)
    for l in range(m.cfg.n_layers):
        m.set_buffer("x", o.buffer("x").copy())
        o.block(l, pos, 0, pos, pos + 1)
        m.block(l, pos, 0, pos, pos + 1)
        a, b = m.buffer("x"), o.buffer("x")
        print(rank, "pos", pos, "layer", l, "rel", float(np.linalg.norm(a - b) / np.linalg.norm(b)), flush=True)
    lg, am = m.forward(tok, pos)
    o.forward(tok, pos)
    e = o.buffer("logits")
    print(rank, "pos", pos, "logits rel", float(np.linalg.norm(lg - e) / np.linalg.norm(e)), "argmax", am, o.argmax(), flush=True)
dist.barrier()
m.close()
