"""GPU: the GEMM family in isolation — tcgen05/TMEM/TMA kernel vs the CUDA-core kernel vs a numpy fp32 reference of the same
op (bf16/f16-rounded operands, fp32 accumulation), on the shapes the engine uses (linears, causal-conv windows, 2-tap transposed convs)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _round(x, f16):
    if f16:
        return x.astype(np.float16).astype(np.float32)
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return u.astype(np.uint32).view(np.float32)


def _ref(A, W, T, taps, f16, bias):
    A = _round(A, f16); W = _round(W, f16)
    n_slots, rows_buf, C = A.shape
    win = np.stack([A[:, t:t + taps].reshape(n_slots, taps * C) for t in range(T)], axis=1).reshape(n_slots * T, taps * C)
    out = win.astype(np.float64) @ W.astype(np.float64).T
    if bias is not None:
        out = out + bias
    return out.astype(np.float32)


@pytest.fixture(scope="module")
def eng(P, model_dir):
    return P.Context(model_dir, max_slots=1, kv_capacity=64).engine


CASES = [
    # n_slots, T, taps, C, N, f16     (name)
    (1, 256, 1, 1024, 3072, False),    # FlowLM in_proj at batch 256
    (1, 300, 1, 1024, 1024, False),    # ragged row count (M tail, TMA zero fill)
    (1, 256, 1, 4096, 1024, False),    # linear2 (long K)
    (1, 64, 1, 512, 32, False),        # flow-head final linear (N = 32)
    (8, 16, 7, 512, 512, True),        # SEANet conv0: k=7 window, tiles span slots
    (3, 96, 3, 256, 128, True),        # resblock conv k=3, T = 96 (32-row chunks)
    (2, 480, 2, 256, 640, True),       # transposed conv as 2-tap GEMM, N = 5*128
    (1, 1920, 3, 64, 32, True),        # r9a: C = 64, N = 32
    (2, 1920, 1, 64, 64, True),        # r9b (padded channels)
    (1, 16, 7, 512, 512, True),        # SEANet conv0 at batch 1 (16 rows, split-K)
    (1, 16, 1, 512, 1536, False),      # Mimi in_proj at batch 1
    (1, 32, 1, 1024, 4096, False),     # FlowLM linear1 at batch 32
]


@pytest.mark.parametrize("n_slots,T,taps,C,N,f16", CASES)
def test_tc_gemm_matches_reference(eng, n_slots, T, taps, C, N, f16):
    rng = np.random.default_rng(n_slots * 1000 + T + N)
    rows_buf = T + taps - 1 + (1 if taps > 1 else 0)
    A = rng.standard_normal((n_slots, rows_buf, C)).astype(np.float32)
    W = (rng.standard_normal((N, taps * C)) / np.sqrt(taps * C)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    ref = _ref(A, W, T, taps, f16, bias)
    out_tc, out2, used_tc = eng.debug_gemm(A, W, T, taps, f16=f16, bias=bias, path=0, want_out2=True)
    out_cc, _, used_cc = eng.debug_gemm(A, W, T, taps, f16=f16, bias=bias, path=1)
    assert used_tc and not used_cc
    # fp32 accumulation of exactly-representable products: only summation order differs
    assert np.abs(out_cc - ref).max() < 2e-4
    assert np.abs(out_tc - ref).max() < 2e-4, np.abs(out_tc - ref).max()
    elu = np.where(ref > 0, ref, np.expm1(ref))
    assert np.abs(out2 - elu).max() < 2e-2 * max(1.0, np.abs(elu).max())   # out2 is bf16/f16-rounded ELU(v)
