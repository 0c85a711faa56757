"""A few batch-1 synchronous steps for `ncu -k regex:flow_persistent` (the persistent FlowLM kernel of BASELINE config 2)."""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import numpy as np
import ptts_b200 as P
from make_assets import default_model_dir
d = default_model_dir(eos_mode="never")
ctx = P.Context(d, max_slots=1, kv_capacity=1024, cuda_graphs=0)
eng = ctx.engine
st = ctx.stream("cosette", temp=0.7)
toks = ctx.tokenize("The quick brown fox jumped over the sleeping dog.")
eng.begin_sentences([0], [st.voice], [toks], [600], [1 << 20], [0.7])
for _ in range(12):
    eng.step(0, 1, None)
print("done")
