"""GPU tests of the round-2 fused kernels, through the C ABI: the flow-head cluster kernel (csrc/head_fused.cuh) and the SEANet tail kernel
(csrc/seanet_tail.cuh). Each is compared with the oracle AND with the unfused launch chain it replaces (same engine, switch in the environment).
Tolerances as in test_gpu_parity.py (latents max-abs <= 4e-2 / rel <= 1.5e-2, waveform SNR >= 40 dB); fused vs unfused is much tighter."""
import os

import numpy as np
import pytest

from conftest import BENCH_SENTENCE, snr_db

pytestmark = pytest.mark.gpu
TEXTS = [BENCH_SENTENCE, "Hello world, this is a test of the head.", "One two three four five six seven."]


class _Env:
    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update({k: str(v) for k, v in self.kv.items()})

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _run_batch(P, model_dir, B, frames, seed=1, **env):
    with _Env(**env):
        ctx = P.Context(model_dir, max_slots=B, kv_capacity=512)
    eng = ctx.engine
    st = ctx.stream("cosette", temp=0.7)
    toks = [ctx.tokenize(TEXTS[i % 3]) for i in range(B)]
    eng.begin_sentences(list(range(B)), [st.voice] * B, toks, [600] * B, [1 << 20] * B, [0.7] * B)
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(frames):
        noise = (rng.standard_normal((B, 32)) * np.sqrt(0.7)).astype(np.float32)
        pcm, prod, lat, eos = eng.step(0, B, noise)
        out.append((noise, lat.copy(), pcm.copy()))
    return out


@pytest.mark.parametrize("B", [19, 3])
def test_fused_head_vs_oracle_and_unfused(P, model_dir, orc, B):
    """B = 19: one full 16-row cluster tile + a ragged one (rows clamped, never stored); B = 3: the smallest batch that takes the cluster kernel."""
    fused = _run_batch(P, model_dir, B, 2)
    plain = _run_batch(P, model_dir, B, 2, PTTS_B200_FUSED_HEAD=0)
    # same inputs up to the head in frame 0: only the summation order inside the head's dot products differs
    assert np.abs(fused[0][1] - plain[0][1]).max() < 1.5e-2
    assert np.isfinite(fused[1][1]).all()
    for row in sorted({0, B - 1}):
        s = orc.stream("cosette", kv_capacity=512)
        s.sentence_init(TEXTS[row % 3])
        for f in range(2):
            ok, lat, pcm, e = s.step(fused[f][0][row])
            assert ok
            assert np.abs(fused[f][1][row] - lat).max() < 4e-2, (row, f)
            assert np.linalg.norm(fused[f][1][row] - lat) / np.linalg.norm(lat) < 1.5e-2, (row, f)
            assert snr_db(pcm, fused[f][2][row]) > 40.0, (row, f)


def test_fused_tail_vs_unfused_across_frames(P, model_dir):
    """Mimi decode only, 5 consecutive frames: the output conv's carried state (two rows of tap products instead of two a3 rows) crosses the
    frame boundary; fused and unfused differ in f32 summation order only."""
    B = 5
    outs = {}
    for mode in (0, 1):
        with _Env(PTTS_B200_FUSED_TAIL=mode):
            ctx = P.Context(model_dir, max_slots=B, kv_capacity=64)
        eng = ctx.engine
        rng = np.random.default_rng(3)
        eng.mimi_reset(0, B)
        outs[mode] = [eng.mimi_decode(0, B, rng.standard_normal((B, 32)).astype(np.float32) * 2.0).copy() for _ in range(5)]
    for f in range(5):
        for b in range(B):
            assert snr_db(outs[0][f][b], outs[1][f][b]) > 80.0, (f, b)
        # the first samples of a frame depend on the previous frame's last rows
        assert np.abs(outs[0][f][:, :4] - outs[1][f][:, :4]).max() < 1e-3, f


def test_fused_tail_state_resets_with_the_sentence(P, model_dir):
    """mimi_reset zeroes the tail's carried rows like every other conv state: the same latent decodes to the same first frame again."""
    B = 2
    ctx = P.Context(model_dir, max_slots=B, kv_capacity=64)
    eng = ctx.engine
    lat = np.random.default_rng(7).standard_normal((B, 32)).astype(np.float32)
    eng.mimi_reset(0, B)
    first = eng.mimi_decode(0, B, lat).copy()
    eng.mimi_decode(0, B, lat * 0.5)
    eng.mimi_reset(0, B)
    again = eng.mimi_decode(0, B, lat)
    assert np.array_equal(first, again)


def test_steps_enqueue_equals_repeated_step_enqueue(P, model_dir):
    """b200_steps_enqueue(count) is count native b200_step_enqueue calls: same positions and, with the device RNG, the same next frame."""
    outs = []
    for native in (False, True):
        ctx = P.Context(model_dir, max_slots=4, kv_capacity=512)
        eng = ctx.engine
        st = ctx.stream("cosette", temp=0.7)
        toks = [ctx.tokenize(TEXTS[i % 3]) for i in range(4)]
        eng.set_seed(11)
        eng.begin_sentences([0, 1, 2, 3], [st.voice] * 4, toks, [600] * 4, [1 << 20] * 4, [0.7] * 4)
        if native:
            eng.steps_enqueue(0, 4, 3)
        else:
            for _ in range(3):
                eng.step_enqueue(0, 4)
        eng.sync()
        pos = [eng.slot_position(s) for s in range(4)]
        pcm, prod, lat, eos = eng.step(0, 4, None)
        outs.append((pos, lat.copy(), pcm.copy()))
    assert outs[0][0] == outs[1][0]
    assert np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2])
