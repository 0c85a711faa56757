#!/usr/bin/env python3
"""Golden vectors produced by the REFERENCE'S OWN SOURCES (oracle/_ref: /root/reference/src/pocket_tts.cpp over the ggml stand-in headers,
`make -C oracle ref`), driven through its own ptts_stream_send / flush / receive with injected noise (oracle/ref_src/inject_normal.h).
The GPU box has no /root/reference, so these fixtures are committed; model weights are re-synthesised there from the same seeds
(tools/make_assets.py), hence they describe the same model. Run HERE:  python tools/make_golden_ref.py

  tests/golden/ref_bench_noise_bf16.npz / _f32.npz   bench sentence, 8 frames, free-running under seeded noise: tokens, noise, latents,
                                                     PCM (every sample), positions
  tests/golden/ref_frame_counts_eos_mid.json         frame counts of three sentences with the EOS-firing checkpoint (same seeds as
                                                     frame_counts_eos_mid.json, which the oracle produced: they must agree)
  tests/golden/ref_rollover_temp0.npz                two-sentence text through ONE stream at temp 0: frames per sentence + first two
                                                     frames of the second sentence
"""
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tools"))
from make_assets import default_model_dir  # noqa: E402
import oracle.ref as R  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")
BENCH = "The quick brown fox jumped over the sleeping dog."
TH = min(8, os.cpu_count() or 1)


def main():
    assert R.build() is not None, "oracle/_ref cannot be built here"
    for dtype in ("BF16", "F32"):
        d = default_model_dir(eos_mode="never", dtype=dtype)
        r = R.Ref(d, TH)
        s = r.stream("cosette", 1.0)
        s.send(BENCH); s.flush()
        rng = np.random.default_rng(0)
        noise = (rng.standard_normal((8, 32)) * np.sqrt(0.7)).astype(np.float32)
        lat, pcm, pos = [], [], []
        for i in range(8):
            p = s.receive(noise[i]); assert p is not None
            lat.append(s.latent()); pcm.append(p); pos.append(s.current_end)
        np.savez_compressed(os.path.join(GOLD, f"ref_bench_noise_{dtype.lower()}.npz"), tokens=np.array(r.tokenize(BENCH), np.int32), noise=noise,
                            latents=np.array(lat), pcm=np.array(pcm, np.float32), current_end=np.array(pos, np.int32))
    d2 = default_model_dir(eos_mode="mid")
    r2 = R.Ref(d2, TH)
    counts = {}
    for si, text in enumerate(["Hello world.", BENCH, "One two three four five six seven eight nine ten eleven twelve."]):
        s = r2.stream("cosette", 1.0)
        s.send(text); s.flush()
        rng = np.random.default_rng(100 + si)
        n = 0
        while True:
            nz = (rng.standard_normal(32) * np.sqrt(0.7)).astype(np.float32)
            if s.receive(nz) is None:
                break
            n += 1
        counts[text] = {"frames": n, "max_gen_len": s.max_gen_len, "seed": 100 + si}
    json.dump(counts, open(os.path.join(GOLD, "ref_frame_counts_eos_mid.json"), "w"), indent=1)
    print(counts)
    d = default_model_dir(eos_mode="never")
    r = R.Ref(d, TH)
    s = r.stream("cosette", 0.0)
    text = "Hello there. How are you?"
    s.send(text); s.flush()
    frames = []
    while True:
        p = s.receive()
        if p is None:
            break
        frames.append(p)
    n1 = 50                                               # "Hello there." : int((2 + 2) * 12.5)
    np.savez_compressed(os.path.join(GOLD, "ref_rollover_temp0.npz"), n_frames=np.int32(len(frames)), n_first=np.int32(n1),
                        second_sentence_frames=np.array(frames[n1:n1 + 2], np.float32), first_frame=np.array(frames[0], np.float32))
    print("rollover frames", len(frames))


if __name__ == "__main__":
    main()
