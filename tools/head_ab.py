"""A/B of an engine switch (environment variable AB_VAR = 0 / 1; default: the fused flow-head cluster kernel (head_fused.cuh) against the unfused launch chain: same weights, same injected noise, B utterances
(default 40: not a multiple of the 16-row cluster tile), a few free-running frames. Prints the worst latent / PCM difference per frame."""
import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import numpy as np
import ptts_b200 as P
from make_assets import default_model_dir
d = default_model_dir(eos_mode="never")
B = int(os.environ.get("B", "40"))
texts = ["The quick brown fox jumped over the sleeping dog.", "Hello world, this is a test of the head.", "One two three four five six seven."]
out = {}
for mode in ("0", "1"):
    os.environ[os.environ.get("AB_VAR", "PTTS_B200_FUSED_HEAD")] = mode
    ctx = P.Context(d, max_slots=B, kv_capacity=512)
    eng = ctx.engine
    st = ctx.stream("cosette", temp=0.7)
    toks = [ctx.tokenize(texts[i % 3]) for i in range(B)]
    eng.begin_sentences(list(range(B)), [st.voice] * B, toks, [600] * B, [1 << 20] * B, [0.7] * B)
    rng = np.random.default_rng(1)
    res = []
    for i in range(6):
        noise = (rng.standard_normal((B, 32)) * np.sqrt(0.7)).astype(np.float32)
        pcm, prod, lat, eos = eng.step(0, B, noise)
        res.append((lat.copy(), pcm.copy()))
    out[mode] = res
    del ctx
for i in range(6):
    l0, p0 = out["0"][i]; l1, p1 = out["1"][i]
    snr = 10 * np.log10((p0.astype(np.float64) ** 2).sum(1) / np.maximum(((p0.astype(np.float64) - p1) ** 2).sum(1), 1e-30))
    print(f"frame {i}: latent max-abs diff {np.abs(l0 - l1).max():.3e} (scale {np.abs(l0).max():.2f}), worst-row PCM SNR {snr.min():.1f} dB, finite {np.isfinite(l1).all()}")
