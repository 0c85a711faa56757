"""CPU oracle for the Pocket-TTS per-frame generation path — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this package. The product (pocket-tts.cpp_b200/) never does.

PARITY UNPINNED: the reference has no golden vectors and cannot be built here (see
oracle/ptts_oracle.cpp header and DESIGN.md); the C++ restatement is cross-checked against an
independent PyTorch restatement instead (tests/torch_second_opinion.py).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
_SO = os.path.join(_HERE, "_build", "libptts_oracle.so")

from .text_oracle import (  # noqa: E402,F401
    StrProcessor, count_words, frames_after_eos_guess, max_gen_len_for, SentencePieceOracle,
)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ptts_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        vp, ci, fp, ip = ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)
        L.oracle_create.restype = vp; L.oracle_create.argtypes = [ci, ci]
        L.oracle_set_threads.argtypes = [ci]
        L.oracle_enable_taps.argtypes = [vp, ci]
        L.oracle_set_tensor.argtypes = [vp, ctypes.c_char_p, fp, ctypes.POINTER(ctypes.c_int64), ci]
        L.oracle_finalize.argtypes = [vp]; L.oracle_finalize.restype = ci
        L.oracle_destroy.argtypes = [vp]
        L.oracle_get_t_combined.argtypes = [vp, fp]
        L.oracle_mimi_bias_row.argtypes = [vp, ci, ci, fp]
        L.oracle_stream_create.restype = vp; L.oracle_stream_create.argtypes = [vp, fp, ci, ci]
        L.oracle_stream_destroy.argtypes = [vp]
        L.oracle_stream_clone.restype = vp; L.oracle_stream_clone.argtypes = [vp]
        L.oracle_sentence_init.argtypes = [vp, ip, ci, ci, ci]
        L.oracle_step.restype = ci; L.oracle_step.argtypes = [vp, fp, fp, fp, fp]
        L.oracle_flowlm_rows.argtypes = [vp, fp, ci]
        L.oracle_input_linear.argtypes = [vp, fp, fp]
        L.oracle_flow_head.argtypes = [vp, fp, fp, fp, fp]
        L.oracle_mimi_reset.argtypes = [vp]
        L.oracle_mimi_frame.argtypes = [vp, fp, fp]
        L.oracle_current_end.restype = ci; L.oracle_current_end.argtypes = [vp]
        L.oracle_set_backbone_input.argtypes = [vp, fp]
        L.oracle_mimi_offset.restype = ci; L.oracle_mimi_offset.argtypes = [vp]
        L.oracle_set_mimi_offset.argtypes = [vp, ci]
        L.oracle_get_kv.argtypes = [vp, ci, ci, fp]
        L.oracle_get_tap.restype = ci; L.oracle_get_tap.argtypes = [vp, ctypes.c_char_p, fp, ci]
        L.oracle_num_threads_available.restype = ci
        _lib = L
    return _lib


def _fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _read_safetensors(path):
    sys.path.insert(0, os.path.join(_REPO, "tools"))
    from make_assets import read_safetensors  # plain file reader, no product code
    return read_safetensors(path)


def _file_dtype(path) -> str:
    import json, struct
    with open(path, "rb") as f:
        (n,) = struct.unpack("<Q", f.read(8))
        h = json.loads(f.read(n))
    for k, v in h.items():
        if k.endswith("in_proj.weight"):
            return v["dtype"]
    return "F32"


class Oracle:
    """Weights of one model directory, following the reference's ptts_init (src/pocket_tts.cpp:273-322)."""

    def __init__(self, model_dir: str, threads: int = 0, taps: bool = False):
        L = lib()
        self.model_dir = model_dir
        path = os.path.join(model_dir, "tts_b6369a24.safetensors")
        self.file_dtype = _file_dtype(path)
        self.weights = _read_safetensors(path)
        if threads <= 0:
            threads = max(1, (os.cpu_count() or 1))
        self.threads = threads
        self.h = L.oracle_create(1 if self.file_dtype == "BF16" else 0, threads)
        for name, arr in self.weights.items():
            arr = np.ascontiguousarray(arr, dtype=np.float32)
            shape = (ctypes.c_int64 * arr.ndim)(*arr.shape)
            L.oracle_set_tensor(self.h, name.encode(), _fp(arr), shape, arr.ndim)
        assert L.oracle_finalize(self.h) == 0
        if taps:
            L.oracle_enable_taps(self.h, 1)
        self.tokenizer = SentencePieceOracle(os.path.join(model_dir, "tokenizer.model"))

    def set_threads(self, n: int):
        self.threads = n
        lib().oracle_set_threads(n)

    def t_combined(self) -> np.ndarray:
        out = np.zeros(512, np.float32)
        lib().oracle_get_t_combined(self.h, _fp(out))
        return out

    def mimi_bias(self, offset: int) -> np.ndarray:
        out = np.zeros((16, 250), np.float32)
        for j in range(16):
            row = np.zeros(250, np.float32)
            lib().oracle_mimi_bias_row(self.h, offset, j, _fp(row))
            out[j] = row
        return out

    def voice_prompt(self, voice: str) -> np.ndarray:
        path = voice if os.path.exists(voice) else os.path.join(self.model_dir, "embeddings", voice + ".safetensors")
        t = _read_safetensors(path)["audio_prompt"]
        return np.ascontiguousarray(t.reshape(-1, 1024), dtype=np.float32)

    def stream(self, voice="cosette", kv_capacity: int = 1000) -> "OracleStream":
        prompt = voice if isinstance(voice, np.ndarray) else self.voice_prompt(voice)
        return OracleStream(self, prompt, kv_capacity)

    def input_linear(self, latent: np.ndarray) -> np.ndarray:
        out = np.zeros(1024, np.float32)
        lib().oracle_input_linear(self.h, _fp(np.ascontiguousarray(latent, np.float32)), _fp(out))
        return out

    def flow_head(self, h: np.ndarray, noise: np.ndarray):
        lat = np.zeros(32, np.float32); e = np.zeros(1, np.float32)
        lib().oracle_flow_head(self.h, _fp(np.ascontiguousarray(h, np.float32)), _fp(np.ascontiguousarray(noise, np.float32)), _fp(lat), _fp(e))
        return lat, float(e[0])


class OracleStream:
    """One utterance stream (reference ptts_stream_t, src/pocket_tts.cpp:333-394)."""

    def __init__(self, oracle: Oracle, prompt, kv_capacity: int, _handle=None):
        self.o = oracle
        if _handle is not None:
            self.h = _handle
            return
        prompt = np.ascontiguousarray(prompt, np.float32)
        self.h = lib().oracle_stream_create(oracle.h, _fp(prompt), prompt.shape[0], kv_capacity)

    def clone(self) -> "OracleStream":
        """Same voice-conditioned state without repeating the voice prefill (deterministic, so identical to a fresh stream)."""
        return OracleStream(self.o, None, 0, _handle=lib().oracle_stream_clone(self.h))

    def __del__(self):
        try:
            lib().oracle_stream_destroy(self.h)
        except Exception:
            pass

    def sentence_init(self, text: str):
        """_stream_sentence_init with the receive-side bookkeeping of src/pocket_tts.cpp:500-512."""
        tokens = self.o.tokenizer.encode(text)
        self.sentence_init_tokens(tokens, max_gen_len_for(text), frames_after_eos_guess(text))
        return tokens

    def sentence_init_tokens(self, tokens, max_gen_len: int, frames_after_eos: int):
        arr = (ctypes.c_int * len(tokens))(*tokens)
        lib().oracle_sentence_init(self.h, arr, len(tokens), max_gen_len, frames_after_eos)

    def step(self, noise=None):
        lat = np.zeros(32, np.float32); pcm = np.zeros(1920, np.float32); e = np.zeros(1, np.float32)
        n = None if noise is None else _fp(np.ascontiguousarray(noise, np.float32))
        ok = lib().oracle_step(self.h, n, _fp(lat), _fp(pcm), _fp(e))
        return bool(ok), lat, pcm, float(e[0])

    def flowlm_rows(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, np.float32).copy()
        lib().oracle_flowlm_rows(self.h, _fp(x), x.shape[0])
        return x

    def mimi_reset(self):
        lib().oracle_mimi_reset(self.h)

    def mimi_frame(self, latent: np.ndarray) -> np.ndarray:
        pcm = np.zeros(1920, np.float32)
        lib().oracle_mimi_frame(self.h, _fp(np.ascontiguousarray(latent, np.float32)), _fp(pcm))
        return pcm

    @property
    def current_end(self) -> int:
        return lib().oracle_current_end(self.h)

    @property
    def mimi_offset(self) -> int:
        return lib().oracle_mimi_offset(self.h)

    def set_mimi_offset(self, off: int):
        lib().oracle_set_mimi_offset(self.h, off)

    def kv(self, layer: int, which: int) -> np.ndarray:
        out = np.zeros((self.current_end, 1024), np.float32)
        lib().oracle_get_kv(self.h, layer, which, _fp(out))
        return out

    def tap(self, name: str):
        n = lib().oracle_get_tap(self.h, name.encode(), None, 0)
        if n < 0:
            return None
        out = np.zeros(n, np.float32)
        lib().oracle_get_tap(self.h, name.encode(), _fp(out), n)
        return out
