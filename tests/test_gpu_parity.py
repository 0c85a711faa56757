"""GPU parity tests (run on the B200 box: pytest -m gpu). Everything goes through the C ABI of libptts_b200.so.

Tolerances (north_star: bit-exact ids / frame counts; latents max-abs + relative error; waveform SNR >= 40 dB):
  * latents: max-abs <= 4e-2 (latent scale ~4), relative L2 <= 1.5e-2 per frame (measured worst case 2.5e-2 / 6.8e-3)  — two *correct* implementations of this
    bf16/f16 pipeline already differ by ~1.5e-2 max-abs through rounding flips (see tests/test_oracle.py)
  * waveform: SNR >= 40 dB per frame against the oracle under identical injected noise
Free-running comparisons use injected noise (temp 0.7): with temp 0 and RANDOM weights the latent feedback loop is
chaotic (any two implementations diverge after ~10 frames), so the temp-0 --bench configuration is compared
teacher-forced (oracle latents fed to both) and on its first frames against the golden fixture.
"""
import json
import os

import numpy as np
import pytest

from conftest import BENCH_SENTENCE, REPO, snr_db

pytestmark = pytest.mark.gpu
GOLD = os.path.join(REPO, "tests", "golden")
LAT_MAXABS, LAT_REL, SNR_MIN = 4e-2, 1.5e-2, 40.0


@pytest.fixture(scope="module")
def ctx(P, model_dir):
    return P.Context(model_dir, max_slots=8, kv_capacity=1024)


@pytest.fixture(scope="module")
def ctx_f32kv(P, model_dir):
    return P.Context(model_dir, max_slots=2, kv_capacity=1024, kv_f32=1, gemm_path=1)


def _begin(ctx, oracle_mod, st, text, temp):
    toks = ctx.tokenize(text)
    ctx.engine.begin_sentence(st.slot, st.voice, toks, oracle_mod.max_gen_len_for(text), oracle_mod.frames_after_eos_guess(text), temp)
    return toks


def test_token_ids_bit_exact(ctx, orc):
    g = json.load(open(os.path.join(GOLD, "text_golden.json")))
    for text, ids in g["token_ids"].items():
        assert ctx.tokenize(text) == ids == orc.tokenizer.encode(text)


@pytest.mark.parametrize("which", ["bf16kv", "f32kv"])
def test_prefill_kv_and_free_running_noise(which, ctx, ctx_f32kv, orc, oracle_mod):
    c = ctx if which == "bf16kv" else ctx_f32kv
    st = c.stream("cosette", temp=0.7)
    toks = _begin(c, oracle_mod, st, BENCH_SENTENCE, 0.7)
    os_ = orc.stream("cosette", kv_capacity=1024)
    assert os_.sentence_init(BENCH_SENTENCE) == toks
    n_pos = os_.current_end
    assert c.engine.slot_position(st.slot) == n_pos                     # voice-embedding indexing / positions bit-exact
    tol = 3e-2 if which == "f32kv" else 6e-2                            # KV entries (scale ~8): bf16 cache adds 2^-9 relative rounding
    for layer in (0, 3, 5):
        for kv in (0, 1):
            assert np.abs(c.engine.read_kv(st.slot, layer, kv, n_pos) - os_.kv(layer, kv)).max() < tol
    rng = np.random.default_rng(0)
    for i in range(24):
        noise = (rng.standard_normal(32) * np.sqrt(0.7)).astype(np.float32)
        ok, lat, pcm, e = os_.step(noise)
        gp, prod, glat, geos = c.engine.step(st.slot, 1, noise[None])
        assert ok and prod[0] == 1
        assert np.abs(glat[0] - lat).max() < LAT_MAXABS, i
        assert np.linalg.norm(glat[0] - lat) / np.linalg.norm(lat) < LAT_REL, i
        assert abs(geos[0] - e) < 5e-2
        assert snr_db(pcm, gp[0]) > SNR_MIN, i


def test_bench_config_temp0_teacher_forced_and_golden(ctx, orc, oracle_mod):
    g = np.load(os.path.join(GOLD, "bench_temp0.npz"))
    st = ctx.stream("cosette", temp=0.0)
    _begin(ctx, oracle_mod, st, BENCH_SENTENCE, 0.0)
    os_ = orc.stream("cosette", kv_capacity=1024)
    os_.sentence_init(BENCH_SENTENCE)
    for i in range(40):
        ok, lat, pcm, e = os_.step(None)
        gp, prod, glat, geos = ctx.engine.step(st.slot, 1, None)
        assert ok and prod[0] == 1
        assert np.abs(glat[0] - lat).max() < LAT_MAXABS, i
        assert np.linalg.norm(glat[0] - lat) / np.linalg.norm(lat) < LAT_REL, i
        if i == 0:
            assert np.abs(glat[0] - g["latents"][0]).max() < LAT_MAXABS
            assert snr_db(g["pcm_full_frame0"], gp[0]) > SNR_MIN
        # Mimi is compared on the SAME latent stream: decode the oracle latent on the engine as well
        ctx.engine.debug_set_latent(st.slot, 1, lat[None])               # teacher forcing for the next FlowLM step


def test_mimi_only_40_frames(ctx, orc):
    """BASELINE config 3 shape: synthetic latents -> PCM; 40 frames crosses the ring wrap and the offset>250 mask regime."""
    rng = np.random.default_rng(1)
    lats = rng.standard_normal((40, 32)).astype(np.float32)
    s = orc.stream("cosette", kv_capacity=256)
    s.mimi_reset()
    ctx.engine.mimi_reset(4, 3)
    for f in range(40):
        ref = s.mimi_frame(lats[f])
        got = ctx.engine.mimi_decode(4, 3, np.stack([lats[f]] * 3))
        assert snr_db(ref, got[0]) > 55.0, f
        assert np.array_equal(got[0], got[1]) and np.array_equal(got[0], got[2])      # batch rows are independent + deterministic


def test_frame_counts_bit_exact(P, model_dir_eos, orc_eos, oracle_mod):
    counts = json.load(open(os.path.join(GOLD, "frame_counts_eos_mid.json")))
    c = P.Context(model_dir_eos, max_slots=2, kv_capacity=1024)
    st = c.stream("cosette", temp=0.7)
    for text, info in counts.items():
        _begin(c, oracle_mod, st, text, 0.7)
        rng = np.random.default_rng(info["seed"])
        n = 0
        while True:
            noise = (rng.standard_normal(32) * np.sqrt(0.7)).astype(np.float32)
            gp, prod, glat, geos = c.engine.step(st.slot, 1, noise[None])
            if not prod[0]:
                break
            n += 1
            assert n <= info["max_gen_len"]
        assert n == info["frames"], text


def test_batched_equals_single(ctx, oracle_mod):
    """Utterances are independent: a ragged batch of 4 sentences gives the same frames as each sentence alone."""
    texts = ["Hello world.", BENCH_SENTENCE, "One two three four five six seven eight nine ten eleven twelve.", "Short one."]
    eng = ctx.engine
    st = ctx.stream("cosette", temp=0.7)
    voice = st.voice
    toks = [ctx.tokenize(t) for t in texts]
    mg = [oracle_mod.max_gen_len_for(t) for t in texts]
    fae = [oracle_mod.frames_after_eos_guess(t) for t in texts]
    rng = np.random.default_rng(7)
    noise = (rng.standard_normal((6, 4, 32)) * np.sqrt(0.7)).astype(np.float32)
    eng.begin_sentences([0, 1, 2, 3], [voice] * 4, toks, mg, fae, [0.7] * 4)
    batch = [eng.step(0, 4, noise[i]) for i in range(6)]
    for k in range(4):
        eng.begin_sentence(5, voice, toks[k], mg[k], fae[k], 0.7)
        for i in range(6):
            pcm, prod, lat, eos = eng.step(5, 1, noise[i, k][None])
            assert prod[0] == batch[i][1][k] == 1
            # different GEMM kernels at different batch sizes (GEMV / CUDA-core tile / tcgen05): same bound as engine-vs-oracle
            assert np.abs(lat[0] - batch[i][2][k]).max() < LAT_MAXABS
            assert snr_db(batch[i][0][k], pcm[0]) > SNR_MIN


def test_stream_api_matches_reference_driver_loop(P, model_dir, orc, oracle_mod):
    """The reference's own call sequence (demos/pocket-tts.cpp:456-520): feed 15 chars at a time, poll receive."""
    c = P.Context(model_dir, max_slots=1, kv_capacity=1024)
    P.set_seed(0)
    st = c.stream("cosette", temp=0.0)
    text = "Hello there. How are you?"
    frames = []
    rest = text
    active = True
    while active:
        active = False
        if rest:
            st.send(rest[:15]); rest = rest[15:]
            if not rest:
                st.flush()
            active = True
        f = st.receive()
        if f is not None:
            frames.append(f); active = True
    # never-EOS checkpoint: each sentence runs to its cap int((words+2)*12.5)  (src/pocket_tts.cpp:429-430)
    sp = oracle_mod.StrProcessor(); sp.ingest(text); sp.flush()
    want = sum(oracle_mod.max_gen_len_for(s) for s in sp.sentences)
    assert len(frames) == want
    # first frame of the first sentence == oracle
    os_ = orc.stream("cosette", kv_capacity=1024)
    os_.sentence_init(sp.sentences[0])
    ok, lat, pcm, e = os_.step(None)
    assert snr_db(pcm, frames[0]) > SNR_MIN
    assert P.get_seed() == 0 and c.sample_rate == 24000 and c.frame_size == 1920


def test_device_rng_statistics(ctx, oracle_mod):
    """temp > 0 without injected noise: counter-based device RNG, seeded, reproducible, N(0, temp)."""
    st = ctx.stream("cosette", temp=0.7)
    eng = ctx.engine
    import ctypes
    outs = []
    for rep in range(2):
        eng.set_seed(1234)
        _begin(ctx, oracle_mod, st, BENCH_SENTENCE, 0.7)
        pcm, prod, lat, eos = eng.step(st.slot, 1, None)
        outs.append(lat.copy())
    assert np.array_equal(outs[0], outs[1])
    eng.set_seed(99)
    _begin(ctx, oracle_mod, st, BENCH_SENTENCE, 0.7)
    pcm, prod, lat2, eos = eng.step(st.slot, 1, None)
    assert not np.array_equal(outs[0], lat2)


def test_pipelined_submit_collect_matches_step(ctx, oracle_mod):
    """b200_submit/b200_collect (two frames in flight, Mimi of frame t overlapping FlowLM of frame t+1) returns exactly the frames
    of the synchronous b200_step, in order."""
    eng = ctx.engine
    st = ctx.stream("cosette", temp=0.7)
    texts = ["Hello world.", BENCH_SENTENCE, "Short one."]
    toks = [ctx.tokenize(t) for t in texts]
    mg = [oracle_mod.max_gen_len_for(t) for t in texts]
    fae = [oracle_mod.frames_after_eos_guess(t) for t in texts]
    rng = np.random.default_rng(11)
    noise = (rng.standard_normal((8, 3, 32)) * np.sqrt(0.7)).astype(np.float32)
    for trial in range(3):                                        # trial 0 runs eagerly / captures the graphs, later trials replay them
        eng.begin_sentences([0, 1, 2], [st.voice] * 3, toks, mg, fae, [0.7] * 3)
        ref = [eng.step(0, 3, noise[i]) for i in range(8)]
        eng.begin_sentences([0, 1, 2], [st.voice] * 3, toks, mg, fae, [0.7] * 3)
        pcm = np.zeros((3, 1920), np.float32); prod = np.zeros(3, np.int32)
        got = []
        eng.submit(0, 3, noise[0])
        eng.submit(0, 3, noise[1])
        for i in range(2, 8):
            eng.submit(0, 3, noise[i])                            # submits stay two frames ahead of collects
            assert eng.collect_into(pcm, prod) == 3
            got.append((pcm.copy(), prod.copy()))
        for _ in range(2):
            assert eng.collect_into(pcm, prod) == 3
            got.append((pcm.copy(), prod.copy()))
        for i in range(8):
            assert np.array_equal(got[i][1], ref[i][1])
            assert np.array_equal(got[i][0], ref[i][0]), (trial, i)


def test_batch64_tensor_core_path_matches_oracle(P, model_dir, orc, oracle_mod):
    """At batch 64 every GEMM of the step (FlowLM decode included) runs on the tcgen05 kernel and the KV split / merge path of the
    streaming attention is exercised: slots 0 and 63 (same sentence, same injected noise) must match the oracle and each other."""
    B = 64
    c = P.Context(model_dir, max_slots=B, kv_capacity=512)
    eng = c.engine
    st = c.stream("cosette", temp=0.7)
    toks = c.tokenize(BENCH_SENTENCE)
    eng.begin_sentences(list(range(B)), [st.voice] * B, [toks] * B, [oracle_mod.max_gen_len_for(BENCH_SENTENCE)] * B,
                        [oracle_mod.frames_after_eos_guess(BENCH_SENTENCE)] * B, [0.7] * B)
    os_ = orc.stream("cosette", kv_capacity=512)
    assert os_.sentence_init(BENCH_SENTENCE) == toks
    rng = np.random.default_rng(3)
    for i in range(8):
        noise = (rng.standard_normal(32) * np.sqrt(0.7)).astype(np.float32)
        ok, lat, pcm, e = os_.step(noise)
        gp, prod, glat, geos = eng.step(0, B, np.stack([noise] * B))
        assert ok and prod.all()
        for k in (0, B - 1):
            assert np.abs(glat[k] - lat).max() < LAT_MAXABS, (i, k)
            assert np.linalg.norm(glat[k] - lat) / np.linalg.norm(lat) < LAT_REL, (i, k)
            assert snr_db(pcm, gp[k]) > SNR_MIN, (i, k)
        # slot independent and deterministic. Bitwise for slots prefilled in the same chunk; rows that fall into different prefill chunks
        # (max_prefill_rows) may see another split-K plan, i.e. another (equally valid) summation order, hence the tolerance for slot 63
        assert np.array_equal(gp[0], gp[1]) and np.array_equal(glat[0], glat[1])
        assert np.abs(glat[0] - glat[B - 1]).max() < LAT_MAXABS and snr_db(gp[0], gp[B - 1]) > SNR_MIN


@pytest.mark.parametrize("opts", [dict(convt_split=1), dict(kv_f32=1), dict(overlap=0, cuda_graphs=0), dict(pdl=0), dict(pdl=2)],
                         ids=["convt_split", "kv_f32_tc", "eager_single_stream", "no_pdl", "pdl_everywhere"])
def test_engine_options_keep_parity(opts, P, model_dir, orc, oracle_mod):
    """Every engine option runs the full pipeline within the stated tolerances (batch 16: tensor-core GEMMs everywhere)."""
    B = 16
    c = P.Context(model_dir, max_slots=B, kv_capacity=512, **opts)
    eng = c.engine
    st = c.stream("cosette", temp=0.7)
    toks = c.tokenize(BENCH_SENTENCE)
    eng.begin_sentences(list(range(B)), [st.voice] * B, [toks] * B, [oracle_mod.max_gen_len_for(BENCH_SENTENCE)] * B,
                        [oracle_mod.frames_after_eos_guess(BENCH_SENTENCE)] * B, [0.7] * B)
    os_ = orc.stream("cosette", kv_capacity=512)
    os_.sentence_init(BENCH_SENTENCE)
    rng = np.random.default_rng(5)
    for i in range(5):
        noise = (rng.standard_normal(32) * np.sqrt(0.7)).astype(np.float32)
        ok, lat, pcm, e = os_.step(noise)
        gp, prod, glat, geos = eng.step(0, B, np.stack([noise] * B))
        assert ok and prod.all()
        assert np.abs(glat[B - 1] - lat).max() < LAT_MAXABS and np.linalg.norm(glat[B - 1] - lat) / np.linalg.norm(lat) < LAT_REL, i
        assert snr_db(pcm, gp[B - 1]) > SNR_MIN, i


def test_pipelined_large_batch_matches_step(P, model_dir, oracle_mod):
    """Same check at batch 48, where the Mimi chunks of frame t are interleaved with (and gated behind) the FlowLM segments of frame t+1."""
    B = 48
    c = P.Context(model_dir, max_slots=B, kv_capacity=512)
    eng = c.engine
    st = c.stream("cosette", temp=0.7)
    toks = c.tokenize(BENCH_SENTENCE)
    args = (list(range(B)), [st.voice] * B, [toks] * B, [oracle_mod.max_gen_len_for(BENCH_SENTENCE)] * B, [oracle_mod.frames_after_eos_guess(BENCH_SENTENCE)] * B, [0.7] * B)
    rng = np.random.default_rng(13)
    noise = (rng.standard_normal((6, B, 32)) * np.sqrt(0.7)).astype(np.float32)
    pcm = np.zeros((B, 1920), np.float32); prod = np.zeros(B, np.int32)
    for trial in range(3):
        eng.begin_sentences(*args)
        ref = [eng.step(0, B, noise[i])[0] for i in range(6)]
        eng.begin_sentences(*args)
        got = []
        eng.submit(0, B, noise[0]); eng.submit(0, B, noise[1])
        for i in range(2, 6):
            eng.submit(0, B, noise[i]); eng.collect_into(pcm, prod); got.append(pcm.copy())
        for _ in range(2):
            eng.collect_into(pcm, prod); got.append(pcm.copy())
        for i in range(6):
            assert np.array_equal(got[i], ref[i]), (trial, i)


def test_stream_lookahead_is_transparent(P, model_dir, monkeypatch):
    """The streaming API's one-frame look-ahead (FlowLM of frame t+1 overlapping Mimi of frame t) returns the same frames, same count."""
    text = "Hello there. How are you?"
    outs = []
    for la in ("1", "0"):
        monkeypatch.setenv("PTTS_B200_LOOKAHEAD", la)
        c = P.Context(model_dir, max_slots=1, kv_capacity=1024)
        P.set_seed(3)
        st = c.stream("cosette", temp=0.7)
        for rep in range(2):                                      # second repetition replays captured graphs
            st.reset(); st.send(text); st.flush()
            frames = []
            while True:
                f = st.receive()
                if f is None:
                    break
                frames.append(f.copy())
        outs.append(frames)
    assert len(outs[0]) == len(outs[1]) > 0
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
