"""oracle/oracle.py — ctypes bindings for the two CPU checkers + synthetic checkpoint minting.

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's CPU legs;
never from the product (deepseek.cpp_b200/).

  RefLib   -> oracle/_ref/libdsref.so : the UNMODIFIED reference compiled from /root/reference (ref_shim.cpp)
  PortLib  -> oracle/libdsk_oracle.so : the plain-C restatement (dsk_oracle.c)

`/root/reference` is never read at run time: libdsref.so is prebuilt in the build container and
travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from typing import Dict, Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
# The reference opens one OpenMP parallel region per 128-row scale band (src/infer.cpp:255-258); on a 128-thread
# host that fork-join cost dominates small models.  Checker runs default to a modest team (OMP's default active
# waiting is kept: a passive wait policy makes the reference ~8x slower).
DEFAULT_THREADS = int(os.environ.get("DSK_ORACLE_THREADS", str(min(16, os.cpu_count() or 1))))
os.environ.setdefault("OMP_NUM_THREADS", str(DEFAULT_THREADS))
sys.path.insert(0, os.path.join(REPO, "deepseek.cpp_b200"))
import dseek  # noqa: E402

QUANT_IDS = {"fp32": 0, "fp16": 1, "f8e5m2": 2, "q2_k": 3, "q3_k": 4}
QK_K = 256
Q8K_BYTES = 292
BLOCK_BYTES = {"q2_k": 84, "q3_k": 110}

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int)
u16p = C.POINTER(C.c_uint16)


def _fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(f32p)


def _vp(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


# ----------------------------------------------------------------------------------------------
# library loading
# ----------------------------------------------------------------------------------------------

def build_port(force: bool = False) -> str:
    so = os.path.join(HERE, "libdsk_oracle.so")
    src = os.path.join(HERE, "dsk_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "port"], stdout=subprocess.DEVNULL)
    return so


def build_ref() -> Optional[str]:
    """Builds oracle/_ref when /root/reference is present (build container); else keeps the prebuilt files."""
    subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)
    so = os.path.join(HERE, "_ref", "libdsref.so")
    return so if os.path.exists(so) else None


_ref = None
_port = None


def ref_lib():
    """The unmodified reference, or None when oracle/_ref was not built."""
    global _ref
    if _ref is None:
        so = os.path.join(HERE, "_ref", "libdsref.so")
        if not os.path.exists(so):
            return None
        L = C.CDLL(so)
        L.ref_session_create.restype = C.c_void_p
        L.ref_session_create.argtypes = [C.c_char_p, C.c_int]
        L.ref_session_destroy.argtypes = [C.c_void_p]
        L.ref_config_int.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_forward.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.ref_copy_embedding.argtypes = [C.c_void_p, C.c_int]
        L.ref_block.argtypes = [C.c_void_p] + [C.c_int] * 5
        L.ref_argmax.argtypes = [C.c_void_p]
        L.ref_state_buffer.restype = C.c_long
        L.ref_state_buffer.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(f32p)]
        L.ref_active_experts.argtypes = [C.c_void_p, i32p]
        L.ref_kv_cache.restype = C.c_long
        L.ref_kv_cache.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(u16p)]
        L.ref_timed_decode.restype = C.c_double
        L.ref_timed_decode.argtypes = [C.c_void_p, i32p, C.c_int, C.c_int, i32p]
        L.ref_silu.restype = C.c_float
        L.ref_silu.argtypes = [C.c_float]
        L.ref_gelu.restype = C.c_float
        L.ref_gelu.argtypes = [C.c_float]
        L.ref_half_to_float.restype = C.c_float
        L.ref_half_to_float.argtypes = [C.c_uint16]
        L.ref_float_to_half.restype = C.c_uint16
        L.ref_float_to_half.argtypes = [C.c_float]
        L.ref_matmul.argtypes = [f32p, f32p, C.c_void_p, C.c_int, C.c_int, C.c_int, f32p, C.c_int, C.c_int]
        L.ref_matmul_expert.argtypes = [f32p, f32p, C.c_void_p] + [C.c_int] * 5 + [f32p, C.c_int, C.c_int]
        L.ref_moe_gate_padded.argtypes = [f32p, f32p, i32p, f32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                          C.c_int, C.c_int, C.c_int]
        L.ref_rmsnorm.argtypes = [f32p, f32p, f32p, C.c_int, C.c_float]
        for nm in ("ref_rope", "ref_rope_v3"):
            getattr(L, nm).argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_float]
        for nm in ("ref_rope_f16", "ref_rope_v3_f16"):
            getattr(L, nm).argtypes = [u16p, C.c_int, C.c_int, C.c_int, C.c_float]
        L.ref_attn.argtypes = [f32p, f32p, f32p, u16p, u16p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_softmax.argtypes = [f32p, f32p, C.c_int]
        L.ref_quantize_q8_K.argtypes = [f32p, C.c_void_p, C.c_long]
        L.ref_quantize_rows.argtypes = [f32p, C.c_void_p, C.c_long, C.c_long, C.c_int]
        L.ref_dequantize_q2_K.argtypes = [C.c_void_p, f32p, C.c_long]
        L.ref_dequantize_q3_K.argtypes = [C.c_void_p, f32p, C.c_long]
        L.ref_vec_dot_q2_K.argtypes = [C.c_int, f32p, C.c_void_p, C.c_void_p]
        L.ref_vec_dot_q3_K.argtypes = [C.c_int, f32p, C.c_void_p, C.c_void_p]
        L.ref_set_num_threads.argtypes = [C.c_int]
        L.ref_set_num_threads(DEFAULT_THREADS)
        _ref = L
    return _ref


class OrkTensor(C.Structure):
    _fields_ = [("quant", C.c_int), ("n_experts", C.c_int), ("rows", C.c_int), ("cols", C.c_int),
                ("data", C.c_void_p), ("scale", f32p)]


class OrkConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("dim", "hidden_dim", "n_layers", "n_heads", "vocab_size", "max_seq_len")] + \
               [("rope_theta", C.c_float), ("norm_eps", C.c_float), ("act_silu", C.c_int),
                ("first_k_dense_replace", C.c_int)] + \
               [(n, C.c_int) for n in ("n_shared_experts", "n_routed_experts", "n_active_routed",
                                       "moe_intermediate_size")] + \
               [("routed_scaling_factor", C.c_float)] + \
               [(n, C.c_int) for n in ("n_group", "norm_topk_prob", "scoring_sigmoid", "topk_group", "topk_method",
                                       "is_v3", "kv_lora_rank", "q_lora_rank", "qk_nope_head_dim",
                                       "qk_rope_head_dim", "v_head_dim", "head_dim", "quant", "bs0", "bs1",
                                       "original_max_position", "use_mla")]


class OrkLayer(C.Structure):
    _fields_ = [("rms_att", f32p), ("rms_ffn", f32p), ("rms_q_a", f32p), ("rms_kv_a", f32p)] + \
               [(n, OrkTensor) for n in ("wq", "wq_a", "wq_b", "wkv_a", "wkv_b", "wo", "wc", "wq_rope_b", "wv_b")] + \
               [("is_moe", C.c_int)] + \
               [(n, OrkTensor) for n in ("w1", "w2", "w3", "sw1", "sw2", "sw3")] + \
               [("moegate", f32p), ("moegate_bias", f32p), ("key_cache", u16p), ("value_cache", u16p)]


class OrkModel(C.Structure):
    _fields_ = [("cfg", OrkConfig), ("layers", C.POINTER(OrkLayer)), ("embed", OrkTensor), ("rms_final", f32p),
                ("wcls", OrkTensor)]


class OrkState(C.Structure):
    _fields_ = [(n, f32p) for n in ("x", "xb", "xb2", "hb", "hb2", "q_a", "q", "kv_a", "kv_b", "att", "moe_weights",
                                    "active_experts_weights", "logits")] + [("active_experts", i32p)] + \
               [("q_c", f32p), ("q_rope", f32p)]


def port_lib():
    global _port
    if _port is None:
        L = C.CDLL(build_port())
        L.ork_half_to_float.restype = C.c_float
        L.ork_half_to_float.argtypes = [C.c_uint16]
        L.ork_float_to_half.restype = C.c_uint16
        L.ork_float_to_half.argtypes = [C.c_float]
        L.ork_f8e5m2_to_float.restype = C.c_float
        L.ork_f8e5m2_to_float.argtypes = [C.c_uint8]
        L.ork_quantize_row_q8_K.argtypes = [f32p, C.c_void_p, C.c_long]
        L.ork_dequantize_row_q2_K.argtypes = [C.c_void_p, f32p, C.c_long]
        L.ork_dequantize_row_q3_K.argtypes = [C.c_void_p, f32p, C.c_long]
        L.ork_vec_dot_q2_K_q8_K.restype = C.c_float
        L.ork_vec_dot_q2_K_q8_K.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.ork_vec_dot_q3_K_q8_K.restype = C.c_float
        L.ork_vec_dot_q3_K_q8_K.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.ork_matmul.argtypes = [f32p, f32p, C.POINTER(OrkTensor), C.c_int, C.c_int, C.c_int]
        L.ork_softmax.argtypes = [f32p, f32p, C.c_int]
        L.ork_moe_gate.argtypes = [f32p, f32p, i32p, f32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,
                                   C.c_int, C.c_int]
        L.ork_rmsnorm.argtypes = [f32p, f32p, f32p, C.c_int, C.c_float]
        for nm in ("ork_rope", "ork_rope_v3"):
            getattr(L, nm).argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_float]
        for nm in ("ork_rope_f16", "ork_rope_v3_f16"):
            getattr(L, nm).argtypes = [u16p, C.c_int, C.c_int, C.c_int, C.c_float]
        L.ork_silu.restype = C.c_float
        L.ork_silu.argtypes = [C.c_float]
        L.ork_gelu.restype = C.c_float
        L.ork_gelu.argtypes = [C.c_float]
        L.ork_attn.argtypes = [f32p, f32p, f32p, u16p, u16p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ork_copy_embedding.argtypes = [C.POINTER(OrkModel), C.POINTER(OrkState), C.c_int]
        L.ork_block.argtypes = [C.POINTER(OrkModel), C.POINTER(OrkState)] + [C.c_int] * 5
        L.ork_forward.argtypes = [C.POINTER(OrkModel), C.POINTER(OrkState), C.c_int, C.c_int, C.c_int]
        L.ork_argmax.argtypes = [f32p, C.c_int]
        L.ork_set_num_threads.argtypes = [C.c_int]
        L.ork_set_num_threads(DEFAULT_THREADS)
        _port = L
    return _port


# ----------------------------------------------------------------------------------------------
# op-level convenience wrappers (numpy in / numpy out), same signature for both checkers
# ----------------------------------------------------------------------------------------------

class Ops:
    """Uniform op-level API over either checker.  which = 'ref' | 'port'."""

    def __init__(self, which: str):
        self.which = which
        self.L = ref_lib() if which == "ref" else port_lib()
        if self.L is None:
            raise RuntimeError("oracle/_ref/libdsref.so not built")
        self.p = "ref_" if which == "ref" else "ork_"

    def quantize_q8k(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.zeros(x.size // QK_K * Q8K_BYTES, dtype=np.uint8)
        fn = self.L.ref_quantize_q8_K if self.which == "ref" else self.L.ork_quantize_row_q8_K
        fn(_fp(x), _vp(out), x.size)
        return out

    def dequantize(self, blocks: np.ndarray, quant: str, k: int) -> np.ndarray:
        out = np.zeros(k, dtype=np.float32)
        blocks = np.ascontiguousarray(blocks)
        if self.which == "ref":
            fn = self.L.ref_dequantize_q2_K if quant == "q2_k" else self.L.ref_dequantize_q3_K
        else:
            fn = self.L.ork_dequantize_row_q2_K if quant == "q2_k" else self.L.ork_dequantize_row_q3_K
        fn(_vp(blocks), _fp(out), k)
        return out

    def matmul(self, x: np.ndarray, w: np.ndarray, quant: str, d: int, n: int, scale: Optional[np.ndarray] = None,
               bs=(128, 128)) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        w = np.ascontiguousarray(w)
        out = np.zeros(d, dtype=np.float32)
        sc = _fp(np.ascontiguousarray(scale, dtype=np.float32)) if scale is not None else None
        if self.which == "ref":
            self.L.ref_matmul(_fp(out), _fp(x), _vp(w), QUANT_IDS[quant], d, n, sc, bs[0], bs[1])
        else:
            t = OrkTensor(QUANT_IDS[quant], 0, d, n, w.ctypes.data, sc)
            self.L.ork_matmul(_fp(out), _fp(x), C.byref(t), -1, bs[0] if scale is not None else 0,
                              bs[1] if scale is not None else 0)
        return out

    def rmsnorm(self, x, w, eps):
        x = np.ascontiguousarray(x, dtype=np.float32)
        w = np.ascontiguousarray(w, dtype=np.float32)
        o = np.zeros_like(x)
        getattr(self.L, self.p + "rmsnorm")(_fp(o), _fp(x), _fp(w), x.size, C.c_float(eps))
        return o

    def rope(self, vec, head_dim, pos, theta, v3: bool):
        v = np.ascontiguousarray(vec, dtype=np.float32).copy()
        getattr(self.L, self.p + ("rope_v3" if v3 else "rope"))(_fp(v), v.size, head_dim, pos, C.c_float(theta))
        return v

    def rope_f16(self, vec_u16, head_dim, pos, theta, v3: bool):
        v = np.ascontiguousarray(vec_u16, dtype=np.uint16).copy()
        getattr(self.L, self.p + ("rope_v3_f16" if v3 else "rope_f16"))(v.ctypes.data_as(u16p), v.size, head_dim, pos,
                                                                        C.c_float(theta))
        return v

    def moe_gate(self, logits, bias, n_active, norm_topk_prob, scale, sigmoid, topk_method, n_group, topk_group):
        x = np.ascontiguousarray(logits, dtype=np.float32).copy()
        w = np.zeros(n_active, dtype=np.float32)
        idx = np.zeros(n_active, dtype=np.int32)
        b = _fp(np.ascontiguousarray(bias, dtype=np.float32)) if bias is not None else None
        fn = self.L.ref_moe_gate_padded if self.which == "ref" else self.L.ork_moe_gate
        fn(_fp(w), b, idx.ctypes.data_as(i32p), _fp(x), x.size, n_active, int(norm_topk_prob), C.c_float(scale),
           int(sigmoid), int(topk_method), n_group, topk_group)
        return idx, w, x

    def attn(self, q, kcache_u16, vcache_u16, head_dim, v_head_dim, n_heads, kv_len):
        q = np.ascontiguousarray(q, dtype=np.float32)
        out = np.zeros(v_head_dim, dtype=np.float32)
        att = np.zeros(kv_len, dtype=np.float32)
        k = np.ascontiguousarray(kcache_u16, dtype=np.uint16)
        v = np.ascontiguousarray(vcache_u16, dtype=np.uint16)
        getattr(self.L, self.p + "attn")(_fp(out), _fp(att), _fp(q), k.ctypes.data_as(u16p), v.ctypes.data_as(u16p),
                                         head_dim, v_head_dim, n_heads, kv_len)
        return out

    def silu(self, v: float) -> float:
        return float(getattr(self.L, self.p + "silu")(C.c_float(v)))


# ----------------------------------------------------------------------------------------------
# model-level: the reference session, and the port driven from a .dseek directory
# ----------------------------------------------------------------------------------------------

class RefSession:
    """Model + InferenceState of the unmodified reference (src/main.cpp:71-83 minus sampler/tokenizer)."""

    def __init__(self, dirname: str, context: int = 0):
        self.L = ref_lib()
        if self.L is None:
            raise RuntimeError("oracle/_ref/libdsref.so not built")
        # the reference prints loader chatter on stdout; keep it off the bench's JSON line
        sys.stdout.flush()
        fd = os.dup(1)
        devnull = os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 1)
        try:
            self.h = self.L.ref_session_create(dirname.encode(), context)
        finally:
            os.dup2(fd, 1)
            os.close(fd)
            os.close(devnull)

    def cfg(self, key: str) -> int:
        return self.L.ref_config_int(self.h, key.encode())

    def forward(self, token: int, pos: int, output_logits: bool = True):
        self.L.ref_forward(self.h, token, pos, 1 if output_logits else 0)

    def copy_embedding(self, token: int):
        self.L.ref_copy_embedding(self.h, token)

    def block(self, layer, pos, kv_sink, kv_pos, kv_len):
        self.L.ref_block(self.h, layer, pos, kv_sink, kv_pos, kv_len)

    def buffer(self, name: str) -> np.ndarray:
        """Live numpy view of an InferenceState buffer (writes go through)."""
        p = f32p()
        n = self.L.ref_state_buffer(self.h, name.encode(), C.byref(p))
        if n == 0:
            raise KeyError(name)
        return np.ctypeslib.as_array(p, shape=(n,))

    def active_experts(self) -> np.ndarray:
        out = np.zeros(64, dtype=np.int32)
        k = self.L.ref_active_experts(self.h, out.ctypes.data_as(i32p))
        return out[:k].copy()

    def kv_cache(self, layer: int, which: int) -> np.ndarray:
        p = u16p()
        n = self.L.ref_kv_cache(self.h, layer, which, C.byref(p))
        return np.ctypeslib.as_array(p, shape=(n,))

    def argmax(self) -> int:
        return self.L.ref_argmax(self.h)

    def timed_decode(self, prompt, steps):
        pr = np.ascontiguousarray(prompt, dtype=np.int32)
        out = np.zeros(steps, dtype=np.int32)
        secs = self.L.ref_timed_decode(self.h, pr.ctypes.data_as(i32p), pr.size, steps, out.ctypes.data_as(i32p))
        return secs, out

    def close(self):
        if self.h:
            self.L.ref_session_destroy(self.h)
            self.h = None


def config_from_metadata(md: Dict[str, str], context: int = 0) -> dict:
    """Config::from_yalm (src/model.cpp:22-127) restated for the python tools."""
    g = lambda k, d=None: md.get(k, d)
    c = dict(
        dim=int(md["dim"]), hidden_dim=int(md["hidden_dim"]), n_layers=int(md["n_layers"]), n_heads=int(md["n_heads"]),
        vocab_size=int(md["vocab_size"]), max_seq_len=int(md["max_seq_len"]), rope_theta=float(md["rope_theta"]),
        norm_eps=float(g("norm_eps", "1e-5")), act_silu=1 if g("act_type", "gelu") == "silu" else 0,
        first_k_dense_replace=int(g("first_k_dense_replace", "0")), n_shared_experts=int(g("n_shared_experts", "0")),
        n_routed_experts=int(g("n_routed_experts", "0")), n_active_routed=int(g("n_active_routed", "0")),
        moe_intermediate_size=int(g("moe_intermediate_size", "0")),
        routed_scaling_factor=float(g("routed_scaling_factor", "1.0")), n_group=int(g("n_group", "1")),
        norm_topk_prob=1 if g("norm_topk_prob", "False") == "True" else 0,
        scoring_sigmoid=1 if g("scoring_func", "softmax") == "sigmoid" else 0, topk_group=int(g("topk_group", "0")),
        topk_method=1 if g("topk_method", "") == "group_limited_greedy" else 0,
        is_v3=1 if md["arch"] == "DeepseekV3ForCausalLM" else 0, kv_lora_rank=int(g("kv_lora_rank", "0")),
        q_lora_rank=int(g("q_lora_rank", "0")), qk_nope_head_dim=int(g("qk_nope_head_dim", "0")),
        qk_rope_head_dim=int(g("qk_rope_head_dim", "0")), v_head_dim=int(g("v_head_dim", "0")),
        quant=QUANT_IDS[md["quant"]], bs0=int(g("quantization_block_size_0", "0")),
        bs1=int(g("quantization_block_size_1", "0")),
        original_max_position=int(md["rope_scaling_original_max_position_embeddings"]),
        use_mla=1 if int(g("use_mla", "0")) else 0,
    )
    c["head_dim"] = c["qk_nope_head_dim"] + c["qk_rope_head_dim"]
    if context:
        c["max_seq_len"] = min(c["max_seq_len"], context)
    return c


class PortSession:
    """The plain-C restatement driven from a .dseek directory (tensor names: src/model.cpp:766-871)."""

    def __init__(self, dirname: str, context: int = 0):
        self.L = port_lib()
        md, T = dseek.read_dir(dirname)
        self.md, self.T = md, T
        c = config_from_metadata(md, context)
        self.c = c
        self._keep = []
        m = OrkModel()
        for k, v in c.items():
            setattr(m.cfg, k, v)
        qn = md["quant"]
        f8 = qn == "f8e5m2"

        def fptr(name):
            a = np.ascontiguousarray(T[name].data, dtype=np.float32)
            self._keep.append(a)
            return _fp(a)

        def tensor(prefix, rows, cols, n_experts=0):
            t = T[prefix + ".weight"]
            arr = np.ascontiguousarray(t.data)
            self._keep.append(arr)
            sc = fptr(prefix + ".scale") if f8 else None
            return OrkTensor(c["quant"], n_experts, rows, cols, arr.ctypes.data, sc)

        self.layers = (OrkLayer * c["n_layers"])()
        nope = c["qk_nope_head_dim"]
        self.kcache, self.vcache = [], []
        for l in range(c["n_layers"]):
            p = f"model.layers.{l}."
            L = self.layers[l]
            L.rms_att = fptr(p + "attn.norm.weight")
            L.rms_ffn = fptr(p + "mlp.norm.weight")
            L.rms_kv_a = fptr(p + "attn.kv_a_norm.weight")
            if c["use_mla"]:    # BlockMLA tensors (src/model.cpp:809-821)
                L.rms_q_a = fptr(p + "attn.q_a_norm.weight")
                L.wq_a = tensor(p + "attn.wq_a", c["q_lora_rank"], c["dim"])
                L.wc = tensor(p + "attn.wc", c["n_heads"] * c["kv_lora_rank"], c["q_lora_rank"])
                L.wq_rope_b = tensor(p + "attn.wq_rope_b", c["n_heads"] * c["qk_rope_head_dim"], c["q_lora_rank"])
                L.wv_b = tensor(p + "attn.wv_b", c["v_head_dim"], c["kv_lora_rank"], c["n_heads"])   # 3-D view (model.cpp:580)
            elif c["q_lora_rank"] > 0:
                L.rms_q_a = fptr(p + "attn.q_a_norm.weight")
                L.wq_a = tensor(p + "attn.wq_a", c["q_lora_rank"], c["dim"])
                L.wq_b = tensor(p + "attn.wq_b", c["n_heads"] * c["head_dim"], c["q_lora_rank"])
            else:
                L.wq = tensor(p + "attn.wq", c["n_heads"] * c["head_dim"], c["dim"])
            L.wkv_a = tensor(p + "attn.wkv_a", c["kv_lora_rank"] + c["qk_rope_head_dim"], c["dim"])
            if not c["use_mla"]:
                L.wkv_b = tensor(p + "attn.wkv_b", c["n_heads"] * (nope + c["v_head_dim"]), c["kv_lora_rank"])
            L.wo = tensor(p + "attn.wo", c["dim"], c["n_heads"] * c["v_head_dim"])
            moe = c["n_routed_experts"] > 0 and l >= c["first_k_dense_replace"]
            L.is_moe = 1 if moe else 0
            if moe:
                E, mi = c["n_routed_experts"], c["moe_intermediate_size"]
                L.moegate = fptr(p + "moegate.weight")
                if c["is_v3"]:
                    L.moegate_bias = fptr(p + "moegate.bias")
                L.w1 = tensor(p + "mlp.w1", mi, c["dim"], E)
                L.w2 = tensor(p + "mlp.w2", c["dim"], mi, E)
                L.w3 = tensor(p + "mlp.w3", mi, c["dim"], E)
                if c["n_shared_experts"] > 0:
                    sh = c["n_shared_experts"] * mi
                    L.sw1 = tensor(p + "shared_mlp.w1", sh, c["dim"])
                    L.sw2 = tensor(p + "shared_mlp.w2", c["dim"], sh)
                    L.sw3 = tensor(p + "shared_mlp.w3", sh, c["dim"])
            else:
                L.w1 = tensor(p + "mlp.w1", c["hidden_dim"], c["dim"])
                L.w2 = tensor(p + "mlp.w2", c["dim"], c["hidden_dim"])
                L.w3 = tensor(p + "mlp.w3", c["hidden_dim"], c["dim"])
            if c["use_mla"]:   # latent rows + rope keys (src/model.cpp:618-619)
                kc = np.zeros(c["max_seq_len"] * c["kv_lora_rank"], dtype=np.uint16)
                vc = np.zeros(c["max_seq_len"] * c["qk_rope_head_dim"], dtype=np.uint16)
            else:
                kc = np.zeros(c["max_seq_len"] * c["n_heads"] * c["head_dim"], dtype=np.uint16)
                vc = np.zeros(c["max_seq_len"] * c["n_heads"] * c["v_head_dim"], dtype=np.uint16)
            self.kcache.append(kc)
            self.vcache.append(vc)
            L.key_cache = kc.ctypes.data_as(u16p)
            L.value_cache = vc.ctypes.data_as(u16p)
        m.layers = C.cast(self.layers, C.POINTER(OrkLayer))
        m.embed = tensor("model.embed", c["vocab_size"], c["dim"])
        m.rms_final = fptr("model.norm.weight")
        m.wcls = tensor("model.output", c["vocab_size"], c["dim"]) if "model.output.weight" in T else m.embed
        self.m = m
        # InferenceState (src/model.cpp:677-726)
        sz = dict(
            x=c["dim"], xb=c["dim"], xb2=max(c["dim"], c["n_heads"] * c["v_head_dim"], c["n_heads"] * c["kv_lora_rank"]),
            hb=max(c["dim"], c["hidden_dim"]), hb2=c["hidden_dim"], q_a=max(1, c["q_lora_rank"]),
            q=c["n_heads"] * c["head_dim"], kv_a=c["kv_lora_rank"] + c["qk_rope_head_dim"],
            kv_b=c["n_heads"] * (nope + c["v_head_dim"]), att=c["n_heads"] * c["max_seq_len"],
            moe_weights=max(1, c["n_routed_experts"]), active_experts_weights=max(1, c["n_active_routed"]),
            logits=c["vocab_size"], q_c=max(1, c["n_heads"] * c["kv_lora_rank"] * c["use_mla"]),
            q_rope=max(1, c["n_heads"] * c["qk_rope_head_dim"] * c["use_mla"]))
        self.buf = {k: np.zeros(v, dtype=np.float32) for k, v in sz.items()}
        self.active = np.zeros(max(1, c["n_active_routed"]), dtype=np.int32)
        s = OrkState()
        for k, a in self.buf.items():
            setattr(s, k, _fp(a))
        s.active_experts = self.active.ctypes.data_as(i32p)
        self.s = s

    def forward(self, token, pos, output_logits=True):
        self.L.ork_forward(C.byref(self.m), C.byref(self.s), token, pos, 1 if output_logits else 0)

    def copy_embedding(self, token):
        self.L.ork_copy_embedding(C.byref(self.m), C.byref(self.s), token)

    def block(self, layer, pos, kv_sink, kv_pos, kv_len):
        self.L.ork_block(C.byref(self.m), C.byref(self.s), layer, pos, kv_sink, kv_pos, kv_len)

    def buffer(self, name):
        return self.buf[name]

    def active_experts(self):
        return self.active[: self.c["n_active_routed"]].copy()

    def kv_cache(self, layer, which):
        return self.kcache[layer] if which == 0 else self.vcache[layer]

    def argmax(self):
        return int(self.L.ork_argmax(_fp(self.buf["logits"]), self.c["vocab_size"]))

    def close(self):
        pass


def open_session(dirname: str, context: int = 0, prefer: str = "ref"):
    """The strongest available checker: the unmodified reference when oracle/_ref exists, else the port."""
    if prefer == "ref" and ref_lib() is not None:
        return RefSession(dirname, context)
    return PortSession(dirname, context)
