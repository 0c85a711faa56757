"""oracle/_ref — the REFERENCE'S OWN SOURCES (src/pocket_tts.cpp + the headers it includes, src/json.cpp, src/safetensor.cpp) compiled
where they lie under /root/reference against the stand-in headers of oracle/ggml_shim (TEST INFRASTRUCTURE; built by `make -C oracle ref`).

What runs here is the reference's graph-building, state and driver code; only ggml's op kernels (and the SentencePiece call, served
by the repo's unigram encoder) are restated. Used to pin the hand-written oracle (tests/test_ref_vs_oracle.py), to generate golden
vectors (tools/make_golden_ref.py) and as the CPU baseline (`cpu_baseline.kind = "reference"`). The product never imports this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libptts_ref.so")
REFERENCE_ROOT = os.environ.get("PTTS_REFERENCE_ROOT", "/root/reference")


def build(force: bool = False) -> str | None:
    """(Re)builds oracle/_ref when the reference sources are present; returns the library path or None."""
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "src")):
        subprocess.check_call(["make", "-C", _HERE, "ref", f"REF={REFERENCE_ROOT}"] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return _SO if os.path.exists(_SO) else None


def available() -> bool:
    return os.path.exists(_SO)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libptts_ref.so has not been built (make -C oracle ref; needs /root/reference)")
        L = ctypes.CDLL(_SO)
        vp, ci, cf, cp = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_char_p
        fp, ip = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)
        for name, res, args in (
            ("ref_init", vp, [cp, ci]), ("ref_set_threads", None, [vp, ci]), ("ref_set_seed", None, [ctypes.c_uint]), ("ref_stream", vp, [vp, cp, cf]),
            ("ref_stream_reset", None, [vp]), ("ref_send", None, [vp, cp]), ("ref_flush", None, [vp]), ("ref_receive", ci, [vp, fp]),
            ("ref_sample_rate", ci, [vp]), ("ref_frame_size", ci, [vp]), ("ref_tokenize", ci, [vp, cp, ip, ci]), ("ref_count_words", ci, [cp]),
            ("ref_pending_sentences", ci, [vp, ci, cp, ci]), ("ref_current_end", ci, [vp]), ("ref_generation_step", ci, [vp]), ("ref_max_gen_len", ci, [vp]),
            ("ref_frames_after_eos", ci, [vp]), ("ref_mimi_offset", ci, [vp]), ("ref_read_kv", ci, [vp, ci, ci, ci, fp]), ("ref_get_latent", None, [vp, fp]),
            ("ref_set_latent", None, [vp, fp]), ("ref_inject_noise", None, [fp, ci]), ("ref_noise_pending", ci, []),
        ):
            fn = getattr(L, name); fn.restype = res; fn.argtypes = args
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


class Ref:
    """ptts_init of the reference (src/pocket_tts.cpp:273-322)."""

    def __init__(self, model_dir: str, threads: int = 0):
        if not model_dir.endswith("/"):
            model_dir += "/"
        self.threads = threads or (os.cpu_count() or 1)
        self.h = lib().ref_init(model_dir.encode(), self.threads)

    def tokenize(self, text: str):
        buf = (ctypes.c_int * 8192)()
        n = lib().ref_tokenize(self.h, text.encode(), buf, 8192)
        return list(buf[:n])

    def stream(self, voice="cosette", temp=0.7) -> "RefStream":
        return RefStream(self, voice, temp)


class RefStream:
    """ptts_stream_t driven through the reference's own send / flush / receive."""

    def __init__(self, ref: Ref, voice: str, temp: float):
        self.ref = ref
        self.h = lib().ref_stream(ref.h, voice.encode(), float(temp))

    def send(self, chunk: str):
        lib().ref_send(self.h, chunk.encode())

    def flush(self):
        lib().ref_flush(self.h)

    def reset(self):
        lib().ref_stream_reset(self.h)

    def receive(self, noise=None):
        """One ptts_stream_receive. noise [32] (already scaled N(0, temp)) is injected for this frame's draw: use a stream created with
        temp = 1 so that the reference applies it unscaled. A receive that starts a sentence draws 32 values for the prefill FIRST
        (src/pocket_tts/models/flow_lm.h:131-133 runs in every forward), so 32 zeros are queued ahead in that case."""
        if noise is not None:
            starts = lib().ref_generation_step(self.h) >= lib().ref_max_gen_len(self.h)
            z = np.ascontiguousarray(np.concatenate([np.zeros(32, np.float32), noise]) if starts else noise, np.float32)
            lib().ref_inject_noise(_fp(z), len(z))
        buf = np.zeros(1920, np.float32)
        ok = lib().ref_receive(self.h, _fp(buf))
        return buf if ok else None

    @property
    def current_end(self):
        return lib().ref_current_end(self.h)

    @property
    def mimi_offset(self):
        return lib().ref_mimi_offset(self.h)

    @property
    def max_gen_len(self):
        return lib().ref_max_gen_len(self.h)

    def kv(self, layer: int, which: int, n_pos: int | None = None) -> np.ndarray:
        n = self.current_end if n_pos is None else n_pos
        out = np.zeros((n, 1024), np.float32)
        assert lib().ref_read_kv(self.h, layer, which, n, _fp(out)) == 0
        return out

    def latent(self) -> np.ndarray:
        out = np.zeros(32, np.float32)
        lib().ref_get_latent(self.h, _fp(out))
        return out

    def set_latent(self, lat):
        lib().ref_set_latent(self.h, _fp(np.ascontiguousarray(lat, np.float32)))

    def pending(self):
        n = lib().ref_pending_sentences(self.h, -1, None, 0)
        out = []
        for i in range(n):
            b = ctypes.create_string_buffer(65536)
            lib().ref_pending_sentences(self.h, i, b, 65536)
            out.append(b.value.decode())
        return out


def bench_sentence_fps(model_dir: str, text: str, threads: int, max_frames: int = 0) -> float:
    """The reference's --bench protocol (demos/pocket-tts.cpp:456-520): temp 0, feed 15 characters per loop, time send + receive only,
    frames * 1000 / sum(ms). max_frames bounds the sample."""
    r = Ref(model_dir, threads)
    s = r.stream("cosette", 0.0)                  # stream creation (voice prefill) is outside the reference's timed region as well
    rest, frames, ms, active = text, 0, 0.0, True
    while active and (max_frames <= 0 or frames < max_frames):
        active = False
        t0 = time.perf_counter()
        if rest:
            s.send(rest[:15]); rest = rest[15:]
            if not rest:
                s.flush()
            active = True
        f = s.receive()
        ms += (time.perf_counter() - t0) * 1e3
        if f is not None:
            frames += 1; active = True
    return frames * 1000.0 / ms
