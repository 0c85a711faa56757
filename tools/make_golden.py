#!/usr/bin/env python3
"""Generates tests/golden/*.npz|json from the oracle + the upstream SentencePiece wheel.

The reference ships no golden vectors and cannot be built here, so these fixtures pin the ORACLE
(regression) and give the GPU tests machine-independent targets: token ids (bit-exact), frame counts (bit-exact),
first-frame latents / EOS logits / PCM under the --bench configuration (temperature 0) and under injected noise.
Run:  python tools/make_golden.py
"""
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tools"))
from make_assets import default_model_dir  # noqa: E402
import oracle  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")
SENTENCES = [
    "The quick brown fox jumped over the sleeping dog.",
    ".!...?",
    "Hello world.",
    "Wait... what?! Ok.",
    "Numbers like 123 and 4567 appear, too.",
    "Ünïcödé ½ ™ ＡＢＣ ﬁ.",
    "  Leading   and trailing   spaces   ",
    "A",
]


def main():
    d = default_model_dir(eos_mode="never")
    o = oracle.Oracle(d, threads=os.cpu_count())
    tok = {s: o.tokenizer.encode(s) for s in SENTENCES}
    words = {s: oracle.count_words(s) for s in SENTENCES}
    json.dump({"token_ids": tok, "count_words": words}, open(os.path.join(GOLD, "text_golden.json"), "w"), indent=1, ensure_ascii=False)

    # --bench configuration: temp 0, teacher-forced == free-running for the oracle itself
    s = o.stream("cosette", kv_capacity=1000)
    s.sentence_init(SENTENCES[0])
    lat, pcm, eos = [], [], []
    for i in range(6):
        ok, l, p, e = s.step(None); assert ok
        lat.append(l); pcm.append(p); eos.append(e)
    np.savez_compressed(os.path.join(GOLD, "bench_temp0.npz"), latents=np.array(lat), pcm=np.array(pcm)[:, ::8].astype(np.float32),
                        pcm_full_frame0=pcm[0], eos=np.array(eos, np.float32), tokens=np.array(tok[SENTENCES[0]], np.int32),
                        current_end=np.int32(s.current_end))
    # injected noise (temp 0.7), 4 frames
    s = o.stream("cosette", kv_capacity=1000)
    s.sentence_init(SENTENCES[0])
    rng = np.random.default_rng(0)
    noise = (rng.standard_normal((4, 32)) * np.sqrt(0.7)).astype(np.float32)
    lat, pcm, eos = [], [], []
    for i in range(4):
        ok, l, p, e = s.step(noise[i]); assert ok
        lat.append(l); pcm.append(p); eos.append(e)
    np.savez_compressed(os.path.join(GOLD, "bench_noise.npz"), noise=noise, latents=np.array(lat), pcm=np.array(pcm)[:, ::8].astype(np.float32), eos=np.array(eos, np.float32))

    # frame counts with the EOS-mid checkpoint under injected noise (bit-exact target)
    d2 = default_model_dir(eos_mode="mid")
    o2 = oracle.Oracle(d2, threads=os.cpu_count())
    counts = {}
    for si, text in enumerate(["Hello world.", SENTENCES[0], "One two three four five six seven eight nine ten eleven twelve."]):
        s = o2.stream("cosette", kv_capacity=1000)
        s.sentence_init(text)
        rng = np.random.default_rng(100 + si)
        n, margins = 0, []
        while True:
            nz = (rng.standard_normal(32) * np.sqrt(0.7)).astype(np.float32)
            ok, l, p, e = s.step(nz)
            margins.append(abs(e))
            if not ok:
                break
            n += 1
        counts[text] = {"frames": n, "max_gen_len": oracle.max_gen_len_for(text), "min_abs_eos_margin": float(min(margins)), "seed": 100 + si}
    json.dump(counts, open(os.path.join(GOLD, "frame_counts_eos_mid.json"), "w"), indent=1)
    print(counts)


if __name__ == "__main__":
    main()
