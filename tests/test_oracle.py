"""CPU: pins the oracle (a) against the committed golden fixtures, (b) against the independent PyTorch restatement."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import BENCH_SENTENCE, REPO, snr_db

GOLD = os.path.join(REPO, "tests", "golden")


def test_oracle_matches_golden_temp0(orc):
    g = np.load(os.path.join(GOLD, "bench_temp0.npz"))
    s = orc.stream("cosette", kv_capacity=1000)
    toks = s.sentence_init(BENCH_SENTENCE)
    assert toks == g["tokens"].tolist()
    for i in range(3):
        ok, lat, pcm, e = s.step(None)
        assert ok
        np.testing.assert_allclose(lat, g["latents"][i], atol=2e-3)
        np.testing.assert_allclose(e, g["eos"][i], atol=2e-3)
        np.testing.assert_allclose(pcm[::8], g["pcm"][i], atol=5e-3)


def test_oracle_matches_golden_noise(orc):
    g = np.load(os.path.join(GOLD, "bench_noise.npz"))
    s = orc.stream("cosette", kv_capacity=1000)
    s.sentence_init(BENCH_SENTENCE)
    for i in range(2):
        ok, lat, pcm, e = s.step(g["noise"][i])
        assert ok
        np.testing.assert_allclose(lat, g["latents"][i], atol=2e-3)


def test_oracle_vs_torch_second_opinion(orc):
    """Streaming C++ oracle == non-streaming PyTorch restatement (teacher-forced), incl. Mimi frames past the
    offset>250 mask quirk (frame 16+). Tolerances: two correct implementations differ by bf16/f16 rounding flips."""
    from torch_second_opinion import SecondOpinion
    so = SecondOpinion(orc.weights, file_bf16=(orc.file_dtype == "BF16"))
    np.testing.assert_allclose(orc.t_combined(), so.t_combined().numpy(), atol=1e-5)
    for off in (0, 16, 240, 250, 256, 272, 496, 512, 752):
        assert np.array_equal(np.isinf(orc.mimi_bias(off)), np.isinf(so.mimi_attention_bias(off).numpy())), off
    s = orc.stream("cosette", kv_capacity=1000)
    toks = s.sentence_init(BENCH_SENTENCE)
    rng = np.random.default_rng(0)
    N = 20
    noises = (rng.standard_normal((N, 32)) * np.sqrt(0.7)).astype(np.float32)
    lats, pcms, eoss = [], [], []
    for i in range(N):
        ok, lat, pcm, e = s.step(noises[i])
        assert ok
        lats.append(lat); pcms.append(pcm); eoss.append(e)
    lats = np.array(lats); pcm = np.concatenate(pcms)
    W = so.W
    rows = torch.cat([torch.from_numpy(orc.voice_prompt("cosette")), W["flow_lm.conditioner.embed.weight"][toks],
                      so.lin("flow_lm.input_linear", torch.cat([W["flow_lm.bos_emb"][None], torch.from_numpy(lats[:-1])]))])
    h = so.flowlm(rows)
    lat2, eos2 = so.head(h[-N:], torch.from_numpy(noises))
    assert np.abs(lat2.numpy() - lats).max() < 5e-2
    assert np.linalg.norm(lat2.numpy() - lats) / np.linalg.norm(lats) < 1e-2
    assert np.abs(eos2.numpy() - np.array(eoss)).max() < 3e-2
    pcm2 = so.mimi(torch.from_numpy(lats)).numpy()
    for f in range(N):
        sl = slice(f * 1920, (f + 1) * 1920)
        assert snr_db(pcm[sl], pcm2[sl]) > 55.0, f


def test_oracle_stop_rule(orc_eos):
    """Frame counts of the EOS-mid checkpoint are reproduced (bit-exact integer bookkeeping)."""
    counts = json.load(open(os.path.join(GOLD, "frame_counts_eos_mid.json")))
    text = "Hello world."
    s = orc_eos.stream("cosette", kv_capacity=1000)
    s.sentence_init(text)
    rng = np.random.default_rng(counts[text]["seed"])
    n = 0
    while True:
        ok, *_ = s.step((rng.standard_normal(32) * np.sqrt(0.7)).astype(np.float32))
        if not ok:
            break
        n += 1
    assert n == counts[text]["frames"]
    assert n < counts[text]["max_gen_len"]


def test_oracle_matches_reference_sources_golden(oracle_mod):
    """The oracle against golden vectors produced by the reference's own sources (oracle/_ref, tools/make_golden_ref.py): runs wherever
    the fixtures are, without /root/reference. F32 checkpoint: agreement ~1e-4; BF16: bf16 rounding flips (stated tolerances)."""
    import json
    import os
    import numpy as np
    from conftest import BENCH_SENTENCE, REPO, snr_db
    from make_assets import default_model_dir
    gold = os.path.join(REPO, "tests", "golden")
    for dtype, lat_tol, snr_min in (("f32", 1e-3, 58.0), ("bf16", 4e-2, 40.0)):
        g = np.load(os.path.join(gold, f"ref_bench_noise_{dtype}.npz"))
        o = oracle_mod.Oracle(default_model_dir(eos_mode="never", dtype=dtype.upper()), threads=os.cpu_count() or 1)
        s = o.stream("cosette", kv_capacity=1000)
        assert s.sentence_init(BENCH_SENTENCE) == list(g["tokens"])
        for i in range(8):
            ok, lat, pcm, e = s.step(g["noise"][i])
            assert ok and s.current_end == int(g["current_end"][i])
            assert np.abs(lat - g["latents"][i]).max() < lat_tol, (dtype, i)
            assert snr_db(g["pcm"][i], pcm) > snr_min, (dtype, i)
            s_lat = g["latents"][i]
            oracle_mod.lib().oracle_set_backbone_input(s.h, s_lat.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_float)))
    ref_counts = json.load(open(os.path.join(gold, "ref_frame_counts_eos_mid.json")))
    own_counts = json.load(open(os.path.join(gold, "frame_counts_eos_mid.json")))
    assert {k: v["frames"] for k, v in ref_counts.items()} == {k: v["frames"] for k, v in own_counts.items()}
