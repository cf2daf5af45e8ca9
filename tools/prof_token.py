"""One-token driver for ncu: mints a layer-truncated workload on the GPU, hydrates a few tokens, then runs single-token
forwards (each forward = one decode_kernel launch).  usage: prof_token.py <workload> <quant> <n_layers> <n_forwards> [mla]
(DSK_TSTAMP=1 + "timeline" as 6th argument prints the per-stage timeline of the last token)"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "deepseek.cpp_b200"))
import bench  # noqa: E402
import dsk  # noqa: E402
import torch  # noqa: E402

wl, quant, nl, nf = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
torch.cuda.set_device(0)
dsk.init(0)
w = bench.workload_cfg(wl, quant, nl or None)
if len(sys.argv) > 5 and sys.argv[5] == "mla":
    w["use_mla"] = 1
m = bench.mint_on_gpu(dsk, w, 0, 1, 0)
pr = bench.prompt_ids(w["vocab_size"])
for p in range(nf):
    m.forward(pr[p % len(pr)], p, dsk.OUTPUT_LOGITS, want_logits=False)
if len(sys.argv) > 6 and sys.argv[6] == "timeline":
    m.profile_token(pr[0], nf)
    print(m.profile_token(pr[1], nf + 1))
print("done", m.active_bytes_per_token() / 1e9, "GB/token")
m.close()
