"""Batch-1 streaming breakdown (BASELINE config 2): ms per frame of the full pipeline through b200_submit/collect with one frame of look-ahead,
of the FlowLM part alone (PTTS_B200_DEBUG_SKIP_MIMI=1 in the environment) and of the Mimi decode alone."""
import os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import numpy as np, torch
import ptts_b200 as P
from make_assets import default_model_dir
d = default_model_dir(eos_mode="never")
B = int(os.environ.get("B", "1"))
ctx = P.Context(d, max_slots=max(B, 1), kv_capacity=1024)
eng = ctx.engine
st = ctx.stream("cosette", temp=0.7)
toks = ctx.tokenize("The quick brown fox jumped over the sleeping dog.")
eng.begin_sentences(list(range(B)), [st.voice] * B, [toks] * B, [600] * B, [1 << 20] * B, [0.7] * B)
pcm = np.zeros((B, 1920), np.float32); prod = np.zeros(B, np.int32)
for _ in range(20):
    eng.submit(0, B); eng.collect_into(pcm, prod)
eng.sync()
N = 300
t0 = time.perf_counter()
eng.submit(0, B)
for _ in range(N - 1):
    eng.submit(0, B); eng.collect_into(pcm, prod)
eng.collect_into(pcm, prod)
eng.sync()
dt = (time.perf_counter() - t0) / N * 1e3
ext = torch.cuda.ExternalStream(eng.stream_handle())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
lat = np.zeros((B, 32), np.float32)
eng.mimi_reset(0, B)
for _ in range(5): eng.mimi_decode(0, B, lat)
torch.cuda.synchronize(); e0.record(ext)
for _ in range(200): eng.mimi_decode_enqueue(0, B)
e1.record(ext); eng.sync()
print(f"BATCH{B} skip_mimi={os.environ.get('PTTS_B200_DEBUG_SKIP_MIMI','0')} pipeline {dt:.4f} ms/frame ({1e3/dt*B:.0f} frames/s)  mimi-only {e0.elapsed_time(e1)/200:.4f} ms")
