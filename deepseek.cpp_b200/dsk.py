"""dsk.py — ctypes binding of libdsk.so (include/dsk.h), mirroring the reference's host call surface.

Names follow the reference (andrewkchan/deepseek.cpp @ 8db9e56): `Model(dir).forward(state, token, pos, mode)`
(src/model.cpp:874-883), `Block::block` -> `Model.block(...)` (src/model.cpp:290-322), `InferenceState`
buffers by name (src/model.h:101-179).  The product path is the CUDA library: importing this module on a
machine without libdsk.so, or calling it without a GPU, raises — there is no CPU fallback here.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "libdsk.so")

import sys  # noqa: E402
sys.path.insert(0, HERE)
import dseek  # noqa: E402

QUANT_IDS = {"fp32": 0, "fp16": 1, "f8e5m2": 2, "q2_k": 3, "q3_k": 4}
HYDRATE_KV_CACHE, OUTPUT_LOGITS = 0, 1

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)
u16p = C.POINTER(C.c_uint16)


class DskError(RuntimeError):
    pass


class Config(C.Structure):
    """dsk_config (include/dsk.h) == the reference's Config (src/model.h:47-96)."""
    _fields_ = [(n, C.c_int) for n in ("dim", "hidden_dim", "n_layers", "n_heads", "vocab_size", "max_seq_len")] + \
               [("rope_theta", C.c_float), ("norm_eps", C.c_float), ("act_silu", C.c_int),
                ("first_k_dense_replace", C.c_int)] + \
               [(n, C.c_int) for n in ("n_shared_experts", "n_routed_experts", "n_active_routed",
                                       "moe_intermediate_size")] + \
               [("routed_scaling_factor", C.c_float)] + \
               [(n, C.c_int) for n in ("n_group", "norm_topk_prob", "scoring_sigmoid", "topk_group", "topk_method",
                                       "is_v3", "kv_lora_rank", "q_lora_rank", "qk_nope_head_dim",
                                       "qk_rope_head_dim", "v_head_dim", "quant", "bs0", "bs1",
                                       "original_max_position", "use_mla")]

    @staticmethod
    def from_metadata(md: Dict[str, str], context: int = 0) -> "Config":
        """Config::from_yalm (src/model.cpp:22-127)."""
        g = lambda k, d=None: md.get(k, d)
        c = Config()
        c.dim, c.hidden_dim, c.n_layers = int(md["dim"]), int(md["hidden_dim"]), int(md["n_layers"])
        c.n_heads, c.vocab_size, c.max_seq_len = int(md["n_heads"]), int(md["vocab_size"]), int(md["max_seq_len"])
        if context:
            c.max_seq_len = min(c.max_seq_len, context)
        c.rope_theta, c.norm_eps = float(md["rope_theta"]), float(g("norm_eps", "1e-5"))
        c.act_silu = 1 if g("act_type", "gelu") == "silu" else 0
        c.first_k_dense_replace = int(g("first_k_dense_replace", "0"))
        c.n_shared_experts, c.n_routed_experts = int(g("n_shared_experts", "0")), int(g("n_routed_experts", "0"))
        c.n_active_routed, c.moe_intermediate_size = int(g("n_active_routed", "0")), int(g("moe_intermediate_size", "0"))
        c.routed_scaling_factor = float(g("routed_scaling_factor", "1.0"))
        c.n_group = int(g("n_group", "1"))
        c.norm_topk_prob = 1 if g("norm_topk_prob", "False") == "True" else 0
        c.scoring_sigmoid = 1 if g("scoring_func", "softmax") == "sigmoid" else 0
        c.topk_group = int(g("topk_group", "0"))
        tm = g("topk_method", "")
        if tm == "noaux_tc":
            raise DskError("topk_method noaux_tc is unsupported (the reference asserts, src/model.cpp:51-53)")
        c.topk_method = 1 if tm == "group_limited_greedy" else 0
        c.is_v3 = 1 if md["arch"] == "DeepseekV3ForCausalLM" else 0
        c.kv_lora_rank, c.q_lora_rank = int(g("kv_lora_rank", "0")), int(g("q_lora_rank", "0"))
        c.qk_nope_head_dim, c.qk_rope_head_dim = int(g("qk_nope_head_dim", "0")), int(g("qk_rope_head_dim", "0"))
        c.v_head_dim = int(g("v_head_dim", "0"))
        c.use_mla = 1 if int(g("use_mla", "0")) else 0      # BlockMLA checkpoints (convert.py --mla)
        c.quant = QUANT_IDS[md["quant"]]
        c.bs0, c.bs1 = int(g("quantization_block_size_0", "0")), int(g("quantization_block_size_1", "0"))
        c.original_max_position = int(md["rope_scaling_original_max_position_embeddings"])
        return c


_lib = None
ABI_VERSION = 3   # DSK_ABI_VERSION (include/dsk.h)


def build(force: bool = False) -> str:
    """Compiles libdsk.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    if os.environ.get("DSK_LIB"):       # explicit library (A/B runs of two builds); never rebuilt
        if not os.path.exists(os.environ["DSK_LIB"]):
            raise DskError(f"DSK_LIB={os.environ['DSK_LIB']} does not exist")
        return os.environ["DSK_LIB"]
    srcs = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc")) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(REPO, "include", "dsk.h"))
    stale = not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if (force or stale) and os.path.exists("/usr/local/cuda/bin/nvcc"):
        subprocess.check_call(["make", "-C", HERE, "libdsk.so"], stdout=subprocess.DEVNULL)
    if not os.path.exists(LIB_PATH):
        raise DskError(f"{LIB_PATH} is missing and cannot be built here — the CUDA extension is required")
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        if L.dsk_abi_version() != ABI_VERSION:      # dsk_config's layout is part of the ABI
            raise DskError(f"libdsk.so speaks ABI {L.dsk_abi_version()}, these bindings ABI {ABI_VERSION}: rebuild (make -C deepseek.cpp_b200)")
        L.dsk_last_error.restype = C.c_char_p
        L.dsk_model_create.restype = C.c_void_p
        L.dsk_model_create.argtypes = [C.POINTER(Config), C.c_int, C.c_int]
        L.dsk_model_destroy.argtypes = [C.c_void_p]
        L.dsk_upload_tensor.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int64), C.c_void_p, C.c_size_t, C.c_int]
        L.dsk_model_finalize.argtypes = [C.c_void_p]
        L.dsk_model_sharding.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.dsk_model_resident_bytes.restype = C.c_size_t
        L.dsk_model_resident_bytes.argtypes = [C.c_void_p]
        L.dsk_model_active_bytes_per_token.restype = C.c_double
        L.dsk_model_active_bytes_per_token.argtypes = [C.c_void_p]
        L.dsk_state_create.restype = C.c_void_p
        L.dsk_state_create.argtypes = [C.c_void_p]
        L.dsk_state_destroy.argtypes = [C.c_void_p]
        L.dsk_state_read.argtypes = [C.c_void_p, C.c_char_p, f32p, C.c_size_t]
        L.dsk_state_write.argtypes = [C.c_void_p, C.c_char_p, f32p, C.c_size_t]
        L.dsk_state_read_i32.argtypes = [C.c_void_p, C.c_char_p, i32p, C.c_size_t]
        L.dsk_kv_read.argtypes = [C.c_void_p, C.c_int, C.c_int, u16p, C.c_size_t]
        L.dsk_kv_write.argtypes = [C.c_void_p, C.c_int, C.c_int, u16p, C.c_size_t]
        L.dsk_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, f32p, C.POINTER(C.c_int)]
        L.dsk_copy_embedding.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.dsk_block_forward.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 5
        L.dsk_decode_greedy.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, i32p, f32p]
        L.dsk_profile_token.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_size_t]
        L.dsk_launches_per_forward.argtypes = [C.c_void_p, C.c_int]
        L.dsk_comm_unique_id.argtypes = [C.c_void_p]
        L.dsk_comm_init.argtypes = [C.c_void_p, C.c_void_p]
        L.dsk_device_info.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_size_t)]
        L.dsk_gemv.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, f32p, C.c_int, C.c_int, f32p, f32p]
        L.dsk_quantize_q8k.argtypes = [f32p, C.c_int, C.c_void_p]
        L.dsk_stage_input.argtypes = [C.c_int, f32p, f32p, C.c_int, C.c_float, f32p, C.c_void_p]
        L.dsk_gate_logits.argtypes = [C.c_int, C.c_int, C.c_int, f32p, f32p, f32p, C.c_float, f32p, f32p]
        L.dsk_sample.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.POINTER(C.c_int)]
        L.dsk_sample_prob.argtypes = [C.c_void_p, C.c_void_p, C.c_int, f32p]
        L.dsk_dequantize_row.argtypes = [C.c_int, C.c_void_p, C.c_int, f32p]
        L.dsk_rmsnorm.argtypes = [f32p, f32p, C.c_int, C.c_float, f32p]
        L.dsk_rope.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int]
        L.dsk_moe_gate.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                                   i32p, f32p]
        L.dsk_bench_gemv.argtypes = [C.c_int] * 6 + [f32p, C.POINTER(C.c_double)]
        L.dsk_attn.argtypes = [f32p, u16p, u16p, C.c_int, C.c_int, C.c_int, C.c_int, f32p]
        _lib = L
    return _lib


def _ck(rc: int):
    if rc != 0:
        raise DskError(lib().dsk_last_error().decode())


def _fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(f32p)


_inited = None


def init(device: int = 0):
    global _inited
    if _inited != device:
        _ck(lib().dsk_init(device))
        _inited = device


def device_info():
    name = C.create_string_buffer(128)
    sm, mem = C.c_int(), C.c_size_t()
    _ck(lib().dsk_device_info(name, C.byref(sm), C.byref(mem)))
    return name.value.decode(), sm.value, mem.value


_DT = {"F32": 0, "F16": 1, "BF16": 2, "F8_E5M2": 3, "F8_E4M3": 4, "I32": 5, "I16": 6, "I8": 7, "U8": 8}   # CodecDType (src/codec.h:62-72)


class Model:
    """Device-resident model (the reference's YALMData + Model + InferenceState + Device::CUDA)."""

    def __init__(self, cfg: Config, rank: int = 0, n_ranks: int = 1, device: Optional[int] = None):
        init(rank if device is None else device)
        self.L = lib()
        self.cfg = cfg
        self.h = self.L.dsk_model_create(C.byref(cfg), rank, n_ranks)
        if not self.h:
            raise DskError(self.L.dsk_last_error().decode())
        self.s = None
        self.rank, self.n_ranks = rank, n_ranks

    @classmethod
    def from_dir(cls, dirname: str, context: int = 0, rank: int = 0, n_ranks: int = 1, device: Optional[int] = None):
        md, tensors = dseek.read_dir(dirname)
        m = cls(Config.from_metadata(md, context), rank, n_ranks, device)
        m.metadata = md
        for name, t in tensors.items():
            m.upload(name, t.dtype, t.shape, t.data)
        m.finalize()
        return m

    def upload(self, name: str, dtype: str, shape, data: np.ndarray):
        arr = np.ascontiguousarray(data)
        shp = (C.c_int64 * 4)(*(list(shape) + [0] * 4)[:4])
        if dtype not in _DT:
            raise DskError(f"tensor {name}: unknown dtype {dtype}")
        _ck(self.L.dsk_upload_tensor(self.h, name.encode(), _DT[dtype], shp, C.c_void_p(arr.ctypes.data),
                                     arr.nbytes, 0))

    def upload_device(self, name: str, dtype: str, shape, dev_ptr: int, nbytes: int):
        """`dev_ptr` is a CUDA device pointer (e.g. torch tensor .data_ptr()) — GPU-side minting."""
        shp = (C.c_int64 * 4)(*(list(shape) + [0] * 4)[:4])
        _ck(self.L.dsk_upload_tensor(self.h, name.encode(), _DT[dtype], shp, C.c_void_p(dev_ptr), nbytes, 1))

    def finalize(self):
        _ck(self.L.dsk_model_finalize(self.h))
        self.s = self.L.dsk_state_create(self.h)
        if not self.s:
            raise DskError(self.L.dsk_last_error().decode())
        self._logits = np.zeros(self.cfg.vocab_size, dtype=np.float32)

    def comm_init(self, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        _ck(self.L.dsk_comm_init(self.h, buf))

    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _ck(lib().dsk_comm_unique_id(buf))
        return buf.raw

    # ---- the reference's call surface ---------------------------------------------------------
    def forward(self, token: int, pos: int, mode: int = OUTPUT_LOGITS, want_logits: bool = True):
        """Model::forward.  Returns (logits view | None, argmax | None)."""
        am = C.c_int(-1)
        lp = _fp(self._logits) if (mode and want_logits) else None
        _ck(self.L.dsk_forward(self.h, self.s, token, pos, mode, lp, C.byref(am) if mode else None))
        return (self._logits if lp is not None else None), (am.value if mode else None)

    def copy_embedding(self, token: int):
        _ck(self.L.dsk_copy_embedding(self.h, self.s, token))

    def block(self, layer: int, pos: int, kv_sink: int, kv_pos: int, kv_len: int):
        _ck(self.L.dsk_block_forward(self.h, self.s, layer, pos, kv_sink, kv_pos, kv_len))

    def decode_greedy(self, start_pos: int, n_steps: int):
        out = np.zeros(n_steps, dtype=np.int32)
        ms = C.c_float(0)
        _ck(self.L.dsk_decode_greedy(self.h, self.s, start_pos, n_steps, out.ctypes.data_as(i32p), C.byref(ms)))
        return out, ms.value

    def sample(self, temperature: float = 1.0, top_p: float = 0.95, coin: float = 0.0) -> int:
        """Sampler::sample on the device (src/sampler.cpp:41-75); `coin` = the host's rand()/RAND_MAX."""
        tok = C.c_int(-1)
        _ck(self.L.dsk_sample(self.h, self.s, C.c_float(temperature), C.c_float(top_p), C.c_float(coin), C.byref(tok)))
        return tok.value

    def sample_prob(self, index: int) -> float:
        """Sampler::sample_prob on the device (src/sampler.cpp:12-26)."""
        pr = C.c_float(0)
        _ck(self.L.dsk_sample_prob(self.h, self.s, index, C.byref(pr)))
        return pr.value

    def buffer(self, name: str, n: Optional[int] = None) -> np.ndarray:
        sizes = self.buffer_sizes()
        n = n or sizes[name]
        out = np.zeros(n, dtype=np.float32)
        _ck(self.L.dsk_state_read(self.s, name.encode(), _fp(out), n))
        return out

    def set_buffer(self, name: str, data: np.ndarray):
        a = np.ascontiguousarray(data, dtype=np.float32)
        _ck(self.L.dsk_state_write(self.s, name.encode(), _fp(a), a.size))

    def buffer_sizes(self):
        c = self.cfg
        hd = c.qk_nope_head_dim + c.qk_rope_head_dim
        return {"x": c.dim, "xb2": max(c.dim, c.n_heads * c.v_head_dim, c.n_heads * c.kv_lora_rank if c.use_mla else 0),
                "q_c": c.n_heads * c.kv_lora_rank if c.use_mla else 0,
                "hb": max(c.hidden_dim, c.n_shared_experts * c.moe_intermediate_size), "q": c.n_heads * hd,
                "kv_a": c.kv_lora_rank + c.qk_rope_head_dim, "kv_b": c.n_heads * (c.qk_nope_head_dim + c.v_head_dim),
                "moe_weights": c.n_routed_experts, "active_experts_weights": c.n_active_routed, "logits": c.vocab_size}

    def active_experts(self) -> np.ndarray:
        out = np.zeros(self.cfg.n_active_routed, dtype=np.int32)
        _ck(self.L.dsk_state_read_i32(self.s, b"active_experts", out.ctypes.data_as(i32p), out.size))
        return out

    def kv_cache(self, layer: int, which: int, n: Optional[int] = None) -> np.ndarray:
        c = self.cfg
        hd = c.qk_nope_head_dim + c.qk_rope_head_dim
        if c.use_mla:   # latent rows / rope keys (BlockMLA, src/model.h:451-452)
            n = n or c.max_seq_len * (c.kv_lora_rank if which == 0 else c.qk_rope_head_dim)
        n = n or c.max_seq_len * c.n_heads * (hd if which == 0 else c.v_head_dim)
        out = np.zeros(n, dtype=np.uint16)
        _ck(self.L.dsk_kv_read(self.h, layer, which, out.ctypes.data_as(u16p), n))
        return out

    def set_kv_cache(self, layer: int, which: int, data: np.ndarray):
        a = np.ascontiguousarray(data, dtype=np.uint16)
        _ck(self.L.dsk_kv_write(self.h, layer, which, a.ctypes.data_as(u16p), a.size))

    def profile_token(self, token: int, pos: int) -> str:
        buf = C.create_string_buffer(1 << 16)
        _ck(self.L.dsk_profile_token(self.h, self.s, token, pos, buf, len(buf)))
        return buf.value.decode()

    def sharding(self):
        """(tensor_parallel, local attention heads, local routed experts) of this rank."""
        tp, nh, ne = C.c_int(0), C.c_int(0), C.c_int(0)
        _ck(self.L.dsk_model_sharding(self.h, C.byref(tp), C.byref(nh), C.byref(ne)))
        return bool(tp.value), nh.value, ne.value

    def resident_bytes(self) -> int:
        return self.L.dsk_model_resident_bytes(self.h)

    def active_bytes_per_token(self) -> float:
        return self.L.dsk_model_active_bytes_per_token(self.h)

    def launches_per_forward(self, mode: int = OUTPUT_LOGITS) -> int:
        return self.L.dsk_launches_per_forward(self.h, mode)

    def close(self):
        if self.s:
            self.L.dsk_state_destroy(self.s)
            self.s = None
        if self.h:
            self.L.dsk_model_destroy(self.h)
            self.h = None


# ---- kernel-level hooks (mirror the reference's test-exposed statics) ------------------------------

def gemv(quant: str, w: np.ndarray, x: np.ndarray, d: int, n: int, scale: Optional[np.ndarray] = None, bs=(128, 128)):
    """matmul (src/infer.cpp:381-417)"""
    init(_inited or 0)
    w = np.ascontiguousarray(w)
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.zeros(d, dtype=np.float32)
    sc = _fp(np.ascontiguousarray(scale, dtype=np.float32)) if scale is not None else None
    _ck(lib().dsk_gemv(QUANT_IDS[quant], d, n, C.c_void_p(w.ctypes.data), sc, bs[0], bs[1], _fp(x), _fp(out)))
    return out


def quantize_q8k(x: np.ndarray) -> np.ndarray:
    init(_inited or 0)
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.zeros(x.size // 256 * 292, dtype=np.uint8)
    _ck(lib().dsk_quantize_q8k(_fp(x), x.size, C.c_void_p(out.ctypes.data)))
    return out


def stage_input(model_quant: str, x: np.ndarray, norm_w: Optional[np.ndarray] = None, eps: float = 1e-6) -> np.ndarray:
    """The activation vector as the tile loop of a `model_quant` model reads it (RMSNorm fused when norm_w is given):
    K-quants -> uint8 block_q8_K records; F32/F16 -> fp32 vector; F8E5M2 -> re-assembled fp16 hi/lo split."""
    init(_inited or 0)
    x = np.ascontiguousarray(x, dtype=np.float32)
    w = _fp(np.ascontiguousarray(norm_w, dtype=np.float32)) if norm_w is not None else None
    if model_quant in ("q2_k", "q3_k"):
        out = np.zeros(x.size // 256 * 292, dtype=np.uint8)
        _ck(lib().dsk_stage_input(QUANT_IDS[model_quant], _fp(x), w, x.size, C.c_float(eps), None, C.c_void_p(out.ctypes.data)))
    else:
        out = np.zeros(x.size, dtype=np.float32)
        _ck(lib().dsk_stage_input(QUANT_IDS[model_quant], _fp(x), w, x.size, C.c_float(eps), _fp(out), None))
    return out


def gate_logits(model_quant: str, gate_w: np.ndarray, x: np.ndarray, norm_w: Optional[np.ndarray] = None, eps: float = 1e-6):
    """MoE gate logits of a `model_quant` model (F32 gate rows on rmsnorm(x)); returns (logits, normalised x)."""
    init(_inited or 0)
    gw = np.ascontiguousarray(gate_w, dtype=np.float32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    E, n = gw.shape
    w = _fp(np.ascontiguousarray(norm_w, dtype=np.float32)) if norm_w is not None else None
    out, xn = np.zeros(E, dtype=np.float32), np.zeros(n, dtype=np.float32)
    _ck(lib().dsk_gate_logits(QUANT_IDS[model_quant], E, n, _fp(gw), _fp(x), w, C.c_float(eps), _fp(out), _fp(xn)))
    return out, xn


def dequantize_row(quant: str, blocks: np.ndarray, k: int) -> np.ndarray:
    init(_inited or 0)
    b = np.ascontiguousarray(blocks)
    out = np.zeros(k, dtype=np.float32)
    _ck(lib().dsk_dequantize_row(QUANT_IDS[quant], C.c_void_p(b.ctypes.data), k, _fp(out)))
    return out


def rmsnorm(x, w, eps):
    init(_inited or 0)
    x = np.ascontiguousarray(x, dtype=np.float32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    out = np.zeros_like(x)
    _ck(lib().dsk_rmsnorm(_fp(x), _fp(w), x.size, C.c_float(eps), _fp(out)))
    return out


def rope(vec, head_dim, pos, theta, v3):
    init(_inited or 0)
    v = np.ascontiguousarray(vec, dtype=np.float32).copy()
    _ck(lib().dsk_rope(_fp(v), v.size, head_dim, pos, C.c_float(theta), int(v3)))
    return v


def moe_gate(logits, bias, n_active, norm_topk_prob, scale, sigmoid, topk_method, n_group, topk_group):
    init(_inited or 0)
    x = np.ascontiguousarray(logits, dtype=np.float32).copy()
    b = _fp(np.ascontiguousarray(bias, dtype=np.float32)) if bias is not None else None
    idx = np.zeros(n_active, dtype=np.int32)
    w = np.zeros(n_active, dtype=np.float32)
    _ck(lib().dsk_moe_gate(_fp(x), b, x.size, n_active, int(norm_topk_prob), C.c_float(scale), int(sigmoid),
                           int(topk_method), n_group, topk_group, idx.ctypes.data_as(i32p), _fp(w)))
    return idx, w, x


def attn(q, kcache_u16, vcache_u16, n_heads, head_dim, v_head_dim, kv_len):
    init(_inited or 0)
    q = np.ascontiguousarray(q, dtype=np.float32)
    k = np.ascontiguousarray(kcache_u16, dtype=np.uint16)
    v = np.ascontiguousarray(vcache_u16, dtype=np.uint16)
    out = np.zeros(n_heads * v_head_dim, dtype=np.float32)
    _ck(lib().dsk_attn(_fp(q), k.ctypes.data_as(u16p), v.ctypes.data_as(u16p), n_heads, head_dim, v_head_dim, kv_len,
                       _fp(out)))
    return out


def bench_gemv(quant: str, d: int, n: int, n_mats: int = 2, warmup: int = 3, iters: int = 20):
    """Isolated GEMV kernel timing (CUDA events).  Returns (avg ms per launch, algorithmic bytes per launch)."""
    init(_inited or 0)
    ms, b = C.c_float(0), C.c_double(0)
    _ck(lib().dsk_bench_gemv(QUANT_IDS[quant], d, n, n_mats, warmup, iters, C.byref(ms), C.byref(b)))
    return ms.value, b.value
