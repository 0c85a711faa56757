"""CPU: the C-ABI library loads without a GPU and exports every symbol include/ptts_b200.h declares, plus the
reference's ten C++ API functions under their original mangled names (drop-in for demos/pocket-tts.cpp)."""
import ctypes
import os
import re
import subprocess

from conftest import REPO

REFERENCE_MANGLED = [
    "_Z13ptts_set_seedj", "_Z13ptts_get_seedv", "_Z9ptts_initP12ggml_backendS0_PKc", "_Z20ptts_get_sample_rateP14ptts_context_t",
    "_Z19ptts_get_frame_sizeP14ptts_context_t", "_Z28ptts_stream_from_safetensorsP14ptts_context_tPKcf",
    "_Z17ptts_stream_resetP13ptts_stream_t", "_Z17ptts_stream_flushP13ptts_stream_t",
    "_Z16ptts_stream_sendP13ptts_stream_tPKc", "_Z19ptts_stream_receiveP13ptts_stream_tPf",
]


def test_exports(P):
    hdr = open(os.path.join(REPO, "include", "ptts_b200.h")).read()
    names = re.findall(r"B200_API\s+[\w\s\*]+?\b(\w+)\s*\(", hdr)
    assert len(names) > 40
    L = ctypes.CDLL(P.LIB_PATH)
    for n in names + REFERENCE_MANGLED:
        assert hasattr(L, n), n


def test_loads_without_gpu_and_fails_loudly(P):
    import torch
    if torch.cuda.is_available():
        return
    cfg = P.default_config(max_slots=1)
    h = ctypes.c_void_p()
    rc = P.lib().b200_engine_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc != 0 and not h.value          # no CUDA device -> error, never a CPU fallback


def test_constants(P):
    assert P.FRAME == 1920 and P.SAMPLE_RATE == 24000


def test_product_never_touches_oracle():
    pkg = os.path.join(REPO, "pocket-tts.cpp_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(root, f), errors="ignore").read()
                assert "oracle" not in src.replace("no CPU fallback", ""), os.path.join(root, f)


def test_gemm_planner_invariants(P):
    """The tile/split-K cost model (host code, no GPU): valid tile widths, no empty splits, workspace bound, and the decode shapes of the
    model never fall back to more partial planes than the reduction kernels were measured with."""
    import ctypes
    L = P.lib()
    L.b200_debug_gemm_plan.restype = ctypes.c_int
    L.b200_debug_gemm_plan.argtypes = [ctypes.c_int] * 5 + [ctypes.POINTER(ctypes.c_int)] * 2
    bn, sp = ctypes.c_int(), ctypes.c_int()
    shapes = [(1024, 3072), (1024, 1024), (1024, 4096), (4096, 1024), (512, 512), (1024, 512), (512, 10240), (512, 32), (512, 1536), (2048, 512), (3584, 512)]
    for R in (3, 16, 64, 256, 512, 4096, 24576):
        for K, N in shapes:
            for ln in (0, 1):
                assert L.b200_debug_gemm_plan(R, N, K, 148, ln, ctypes.byref(bn), ctypes.byref(sp)) == 0
                assert bn.value in (32, 64, 128) and N % bn.value == 0
                assert 1 <= sp.value <= 16
                if sp.value > 1:
                    kbps = -(-(K // 64) // sp.value)
                    assert kbps >= 4 and sp.value * R * N <= 32 << 20          # >= 4 k-blocks per split, partial planes fit the workspace
                if R > 512:
                    assert sp.value == 1                                       # large-M GEMMs are never split
    assert L.b200_debug_gemm_plan(0, 64, 64, 148, 0, ctypes.byref(bn), ctypes.byref(sp)) != 0
