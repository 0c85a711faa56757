"""A few batch-B synchronous steps (eager launches, no graphs) for `ncu -k regex:head_res_cluster` (fused flow-head kernel)."""
import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import numpy as np
import ptts_b200 as P
from make_assets import default_model_dir
d = default_model_dir(eos_mode="never")
B = int(os.environ.get("B", "256"))
ctx = P.Context(d, max_slots=B, kv_capacity=512, cuda_graphs=0)
eng = ctx.engine
st = ctx.stream("cosette", temp=0.7)
toks = ctx.tokenize("The quick brown fox jumped over the sleeping dog.")
eng.begin_sentences(list(range(B)), [st.voice] * B, [toks] * B, [600] * B, [1 << 20] * B, [0.7] * B)
for _ in range(4):
    eng.step(0, B, None)
print("done")
