"""CPU: pins the hand-written oracle (oracle/ptts_oracle.cpp) against oracle/_ref = the REFERENCE'S OWN SOURCES (src/pocket_tts.cpp and
the headers it includes) compiled against the ggml stand-in headers (oracle/ggml_shim, `make -C oracle ref`). The graph structure,
state handling and driver loop that run in _ref are the reference's; only ggml's op kernels are restated. Skipped when the library
has not been built (it needs /root/reference, which does not exist on the GPU box; the prebuilt .so travels there).

  * F32 checkpoint: no bf16 rounding points inside the linears -> oracle and reference agree to ~1e-4 (f16-table GELU / f16 im2col
    rounding flips are all that is left): KV rows, latents, waveform.
  * BF16 checkpoint (what the engine computes in): identical rounding points, different summation order -> bf16 rounding flips; the
    agreement is the one of two correct implementations of this pipeline (same tolerances as engine vs oracle).
  * driver semantics bit-exact: token ids, positions, sentence splitting, caps, frames_after_eos, frame counts, Mimi offsets.
"""
import os

import numpy as np
import pytest

from conftest import BENCH_SENTENCE, snr_db

ref_mod = pytest.importorskip("oracle.ref")
if ref_mod.build() is None:
    pytest.skip("oracle/_ref not built and /root/reference absent", allow_module_level=True)

THREADS = min(8, os.cpu_count() or 1)


@pytest.mark.parametrize("dtype,lat_tol,kv_tol,snr_min", [("F32", 1e-3, 2e-3, 58.0), ("BF16", 4e-2, 6e-2, 40.0)])
def test_oracle_matches_reference_sources(dtype, lat_tol, kv_tol, snr_min, oracle_mod):
    from make_assets import default_model_dir
    d = default_model_dir(eos_mode="never", dtype=dtype)
    r = ref_mod.Ref(d, threads=THREADS)
    s = r.stream("cosette", 1.0)                       # temp 1: injected noise arrives unscaled (oracle/ref_src/inject_normal.h)
    o = oracle_mod.Oracle(d, threads=THREADS)
    os_ = o.stream("cosette", kv_capacity=1000)
    toks = os_.sentence_init(BENCH_SENTENCE)
    assert toks == r.tokenize(BENCH_SENTENCE)
    s.send(BENCH_SENTENCE); s.flush()
    rng = np.random.default_rng(0)
    n0 = 125 + len(toks)
    for i in range(6):
        noise = (rng.standard_normal(32) * np.sqrt(0.7)).astype(np.float32)
        pcm = s.receive(noise)
        ok, lat, opcm, e = os_.step(noise)
        assert ok and pcm is not None
        assert s.current_end == os_.current_end == n0 + i + 1
        assert s.mimi_offset == os_.mimi_offset == 16 * (i + 1)
        if i == 0:
            assert s.max_gen_len == oracle_mod.max_gen_len_for(BENCH_SENTENCE) == 137
            for layer in (0, 3, 5):
                for kv in (0, 1):
                    assert np.abs(s.kv(layer, kv, n0) - os_.kv(layer, kv)[:n0]).max() < kv_tol, (layer, kv)
            assert np.abs(s.kv(0, 0, n0) - os_.kv(0, 0)[:n0]).max() < 1e-5         # layer 0 keys: LN + in_proj + RoPE only
        assert np.abs(s.latent() - lat).max() < lat_tol, i
        assert snr_db(opcm, pcm) > snr_min, i
        if dtype == "BF16":                                                       # keep the comparison per frame (chaotic latent feedback)
            s.set_latent(lat)


def test_reference_driver_loop_semantics(oracle_mod):
    """Sentence splitting, caps, frames_after_eos, roll-over to the next sentence and frame counts (EOS-firing checkpoint)."""
    from make_assets import default_model_dir
    d = default_model_dir(eos_mode="mid")
    r = ref_mod.Ref(d, threads=THREADS)
    o = oracle_mod.Oracle(d, threads=THREADS)
    text = "hello there.   how are you?  fine"
    s = r.stream("cosette", 1.0)
    for i in range(0, len(text), 7):
        s.send(text[i:i + 7])
    s.flush()
    sp = oracle_mod.StrProcessor(); sp.ingest(text); sp.flush()
    sentences = [x.decode() if isinstance(x, bytes) else x for x in sp.sentences]
    assert s.pending() == sentences == ["Hello there.", "How are you?", "Fine."]
    rng = np.random.default_rng(3)
    counts = []
    for sent in sentences:
        s1 = r.stream("cosette", 1.0)                                             # one sentence per stream: a finished sentence then returns "no frame"
        s1.send(sent); s1.flush()
        os_ = o.stream("cosette", kv_capacity=1000)
        os_.sentence_init(sent)
        n = 0
        while True:
            noise = (rng.standard_normal(32) * np.sqrt(0.7)).astype(np.float32)
            pcm = s1.receive(noise)
            ok, lat, opcm, e = os_.step(noise)
            if n == 0:
                assert s1.max_gen_len == oracle_mod.max_gen_len_for(sent)
                assert ref_mod.lib().ref_frames_after_eos(s1.h) == oracle_mod.frames_after_eos_guess(sent)
            if pcm is None or not ok:
                assert pcm is None and not ok, (sent, n)                          # both stop at the same frame (EOS rule / cap)
                break
            n += 1
            s1.set_latent(lat)                                                    # same latent stream on both sides: the EOS decisions must agree
        assert 1 <= n <= oracle_mod.max_gen_len_for(sent), sent
        counts.append(n)
    assert any(c < oracle_mod.max_gen_len_for(t) for c, t in zip(counts, sentences))           # the EOS rule did end a sentence early
    # roll-over inside ONE stream (src/pocket_tts.cpp:494-519), temp 0 = deterministic: the frames of the whole text are the sentences' frames
    def count(stream):
        k = 0
        while stream.receive() is not None:
            k += 1
        return k
    per = []
    for sent in sentences:
        s2 = r.stream("cosette", 0.0); s2.send(sent); s2.flush(); per.append(count(s2))
    s3 = r.stream("cosette", 0.0); s3.send(text); s3.flush()
    assert count(s3) == sum(per)
