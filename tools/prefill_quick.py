#!/usr/bin/env python3
"""Sentence-start cost (voice-prefix restore + text prefill) of a full batch, wall clock. GPU box only."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import ptts_b200 as P
from make_assets import default_model_dir
from bench import synth_paragraph

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
tv = int(sys.argv[2]) if len(sys.argv) > 2 else 125
d = default_model_dir(eos_mode="never", t_voice=tv, voices=["cosette"])
mpr = int(sys.argv[3]) if len(sys.argv) > 3 else 512
ctx = P.Context(d, max_slots=B, max_voices=1, kv_capacity=2048, max_prefill_rows=mpr)
eng = ctx.engine
t0 = time.perf_counter(); st = ctx.stream("cosette", temp=0.7); eng.sync(); t_voice = time.perf_counter() - t0
toks = [ctx.tokenize(synth_paragraph(i)) for i in range(B)]
for rep in range(3):
    t0 = time.perf_counter()
    eng.begin_sentences(list(range(B)), [st.voice] * B, toks, [600] * B, [3] * B, [0.7] * B)
    eng.sync()
    dt = time.perf_counter() - t0
    print(f"PREFILL B={B} mpr={mpr} t_voice={tv} rows={sum(map(len, toks))} voice_prefill={t_voice*1e3:.1f} ms begin_sentences={dt*1e3:.1f} ms", flush=True)
