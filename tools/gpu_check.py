#!/usr/bin/env python3
"""Diagnostic parity run on a GPU box: engine vs oracle, prints error figures (used to calibrate the pytest tolerances)."""
import argparse
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tools"))

from make_assets import default_model_dir  # noqa: E402
import oracle  # noqa: E402
import ptts_b200 as P  # noqa: E402


def snr_db(ref, x):
    err = (ref - x).astype(np.float64)
    return 10 * np.log10((ref.astype(np.float64) ** 2).sum() / max((err ** 2).sum(), 1e-30))


def run(kv_f32, gemm_path, convt_split, frames, eos_mode="never", temp=0.7, text="The quick brown fox jumped over the sleeping dog."):
    d = default_model_dir(eos_mode=eos_mode)
    o = oracle.Oracle(d, threads=os.cpu_count())
    ctx = P.Context(d, max_slots=4, kv_capacity=1024, kv_f32=kv_f32, gemm_path=gemm_path, convt_split=convt_split)
    eng = ctx.engine
    st = ctx.stream("cosette", temp=temp)
    toks = ctx.tokenize(text)
    assert toks == o.tokenizer.encode(text), (toks, o.tokenizer.encode(text))
    os_ = o.stream("cosette", kv_capacity=1024)
    os_.sentence_init(text)
    mg, fae = oracle.max_gen_len_for(text), oracle.frames_after_eos_guess(text)
    eng.begin_sentence(st.slot, st.voice, toks, mg, fae, temp)
    # prefill KV parity
    n_pos = os_.current_end
    for layer in (0, 5):
        for which in (0, 1):
            a = eng.read_kv(st.slot, layer, which, n_pos); b = os_.kv(layer, which)
            print(f"  kv layer{layer} {'KV'[which]} maxabs {np.abs(a - b).max():.4e} (scale {np.abs(b).max():.3f})")
    rng = np.random.default_rng(0)
    worst_lat, worst_rel, snrs = 0, 0, []
    nf = 0
    for i in range(frames):
        noise = (rng.standard_normal(32) * np.sqrt(temp)).astype(np.float32)
        ok, lat, pcm, e = os_.step(noise)
        gp, prod, glat, geos = eng.step(st.slot, 1, noise[None])
        if bool(prod[0]) != ok:
            print(f"  frame {i}: produced mismatch oracle={ok} engine={prod[0]} (eos oracle {e:.4f} engine {geos[0]:.4f})")
            break
        if not ok:
            print(f"  sentence finished at frame {i} (both)")
            break
        nf += 1
        dl = np.abs(glat[0] - lat).max(); rl = np.linalg.norm(glat[0] - lat) / np.linalg.norm(lat)
        worst_lat = max(worst_lat, dl); worst_rel = max(worst_rel, rl)
        s = snr_db(pcm, gp[0]); snrs.append(s)
        if i < 3 or i % 10 == 0:
            print(f"  frame {i}: latent maxabs {dl:.3e} rel {rl:.3e} eos {e:.4f}/{geos[0]:.4f} pcm snr {s:.1f} dB")
    print(f"kv_f32={kv_f32} gemm_path={gemm_path} split={convt_split}: frames {nf} worst latent maxabs {worst_lat:.3e} rel {worst_rel:.3e} "
          f"min snr {min(snrs):.1f} dB mean {np.mean(snrs):.1f} dB")
    return ctx, o


def mimi_only(ctx, o, frames=40):
    eng = ctx.engine
    rng = np.random.default_rng(1)
    lats = rng.standard_normal((frames, 32)).astype(np.float32)
    s = o.stream("cosette", kv_capacity=256)
    s.mimi_reset()
    eng.mimi_reset(0, 2)
    snrs = []
    for f in range(frames):
        ref = s.mimi_frame(lats[f])
        got = eng.mimi_decode(0, 2, np.stack([lats[f], lats[f]]))
        snrs.append(snr_db(ref, got[0]))
        assert np.array_equal(got[0], got[1]), "batch rows differ"
    print("mimi-only snr per frame:", " ".join(f"{x:.1f}" for x in snrs))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=40)
    a = ap.parse_args()
    t = time.time()
    ctx, o = run(kv_f32=1, gemm_path=1, convt_split=1, frames=a.frames)
    mimi_only(ctx, o)
    run(kv_f32=0, gemm_path=1, convt_split=1, frames=a.frames)
    run(kv_f32=0, gemm_path=1, convt_split=0, frames=a.frames)
    run(kv_f32=0, gemm_path=1, convt_split=1, frames=140, eos_mode="mid", temp=0.0)
    # timing of the batch-1 streaming path
    d = default_model_dir()
    ctx = P.Context(d, max_slots=1, kv_capacity=1024)
    st = ctx.stream("cosette", temp=0.0)
    st.send("The quick brown fox jumped over the sleeping dog."); st.flush()
    t0 = time.time(); n = 0
    while st.receive() is not None:
        n += 1
    dt = time.time() - t0
    print(f"batch-1 streaming API: {n} frames in {dt:.3f}s = {n / dt:.1f} frames/s, launches {ctx.engine.launch_count()}")
    print("total", time.time() - t)
