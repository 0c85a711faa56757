#!/usr/bin/env python3
"""Parity numbers (engine vs CPU oracle) under different engine options, printed as JSON lines. GPU box only.

  mimi   Mimi decoder only, 24 frames of synthetic latents: min / mean waveform SNR per frame
  full   full pipeline at batch B (tcgen05 path for every GEMM when B >= 16), identical injected noise, free running:
         per-frame latent max-abs / rel-L2 and waveform SNR of utterances 0 and B-1 against the oracle
Usage: python tools/parity_report.py [--batch 64] [--frames 12] [--opts convt_split=0 ...]
"""
import argparse
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tools")); sys.path.insert(0, os.path.join(REPO, "tests"))
from make_assets import default_model_dir  # noqa: E402

SENT = "The quick brown fox jumped over the sleeping dog."


def snr_db(ref, x):
    err = ref.astype(np.float64) - x.astype(np.float64)
    return float(10 * np.log10((ref.astype(np.float64) ** 2).sum() / max((err ** 2).sum(), 1e-30)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--opts", nargs="*", default=[])
    a = ap.parse_args()
    import oracle
    import ptts_b200 as P
    oracle.build()
    d = default_model_dir(eos_mode="never")
    opts = {k: int(v) for k, v in (o.split("=") for o in a.opts)}
    orc = oracle.Oracle(d, threads=os.cpu_count() or 1)

    # ---- Mimi only ----
    ctx = P.Context(d, max_slots=max(a.batch, 4), kv_capacity=512, **opts)
    eng = ctx.engine
    rng = np.random.default_rng(1)
    lats = rng.standard_normal((24, 32)).astype(np.float32)
    s = orc.stream("cosette", kv_capacity=256); s.mimi_reset()
    eng.mimi_reset(0, a.batch)
    snrs = []
    for f in range(24):
        ref = s.mimi_frame(lats[f])
        got = eng.mimi_decode(0, a.batch, np.stack([lats[f]] * a.batch))
        snrs.append(snr_db(ref, got[0]))
        assert np.array_equal(got[0], got[a.batch - 1])
    print(json.dumps({"case": "mimi", "opts": opts, "batch": a.batch, "snr_min": round(min(snrs), 2), "snr_mean": round(float(np.mean(snrs)), 2)}), flush=True)

    # ---- full pipeline, batch B, the same sentence and noise in every slot ----
    st = ctx.stream("cosette", temp=0.7)
    toks = ctx.tokenize(SENT)
    B = a.batch
    eng.begin_sentences(list(range(B)), [st.voice] * B, [toks] * B, [oracle.max_gen_len_for(SENT)] * B, [oracle.frames_after_eos_guess(SENT)] * B, [0.7] * B)
    os_ = orc.stream("cosette", kv_capacity=512)
    assert os_.sentence_init(SENT) == toks
    rng = np.random.default_rng(0)
    rows = []
    for i in range(a.frames):
        noise = (rng.standard_normal(32) * np.sqrt(0.7)).astype(np.float32)
        ok, lat, pcm, e = os_.step(noise)
        gp, prod, glat, geos = eng.step(0, B, np.stack([noise] * B))
        for k in (0, B - 1):
            rows.append((float(np.abs(glat[k] - lat).max()), float(np.linalg.norm(glat[k] - lat) / np.linalg.norm(lat)), snr_db(pcm, gp[k])))
    r = np.array(rows)
    print(json.dumps({"case": "full", "opts": opts, "batch": B, "frames": a.frames, "lat_maxabs_max": round(float(r[:, 0].max()), 4),
                      "lat_rel_max": round(float(r[:, 1].max()), 4), "snr_min": round(float(r[:, 2].min()), 2), "snr_mean": round(float(r[:, 2].mean()), 2)}), flush=True)


if __name__ == "__main__":
    main()
