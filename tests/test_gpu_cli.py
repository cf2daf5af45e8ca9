"""The C++ host (`deepseek.cpp_b200/main`, the reference's `main <checkpoint_dir>` CLI surface over the C-ABI):
.dseek loader + trie tokenizer + greedy completion must reproduce the reference's generated token ids."""
import os
import re
import subprocess

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("quant,mla", [("fp32", False), ("f8e5m2", False), ("fp32", True)])
def test_cli_completion_matches_reference(repo, ckpt, quant, mla):
    exe = os.path.join(repo, "deepseek.cpp_b200", "main")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.dirname(exe), "main"])
    d = ckpt("tiny_v2", quant, use_mla=True) if mla else ckpt("tiny_v2lite", quant)   # mla: a `convert.py --mla` style checkpoint
    prompt, steps = "hello world", 20
    out = subprocess.run([exe, d, "-i", prompt, "-n", str(steps), "-t", "0"], capture_output=True, text=True, errors="replace", timeout=120)
    assert out.returncode == 0, out.stderr[-2000:]
    enc = [int(v) for v in re.search(r"^\[([0-9,]+)\]$", out.stdout, re.M).group(1).split(",")]
    # trie encoder over the synthetic vocab: BOS then byte b -> id b + 2 (src/tokenizer.cpp:57-94)
    assert enc == [0] + [b + 2 for b in prompt.encode()]
    ids = [int(v) for v in re.search(r"generated ids:([ 0-9]+)", out.stdout).group(1).split()]
    assert ids[: len(enc)] == enc
    gen = ids[len(enc):]
    o = O.open_session(d)
    for p, t in enumerate(enc):
        o.forward(t, p, p + 1 == len(enc))
    ref, pos = [], len(enc)
    for _ in range(len(gen)):
        t = o.argmax()
        ref.append(t)
        if t == 1:   # eos_token_id of the synthetic checkpoints
            break
        o.forward(t, pos)
        pos += 1
    o.close()
    assert gen == ref
    assert "Generation stats:" in out.stdout and "tok/s" in out.stdout


def test_cli_rejects_bad_input(repo, tmp_path):
    exe = os.path.join(repo, "deepseek.cpp_b200", "main")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.dirname(exe), "main"])
    bad = tmp_path / "empty"
    bad.mkdir()
    out = subprocess.run([exe, str(bad), "-i", "x"], capture_output=True, text=True, timeout=60)
    assert out.returncode != 0 and "no files found" in out.stderr
    (bad / "junk.dseek").write_bytes(b"\x10\x00\x00\x00\x00\x00\x00\x00not json at all!")
    out = subprocess.run([exe, str(bad), "-i", "x"], capture_output=True, text=True, timeout=60)
    assert out.returncode != 0
