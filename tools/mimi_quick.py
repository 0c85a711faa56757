import sys, os, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import numpy as np, torch
import ptts_b200 as P
from make_assets import default_model_dir
d = default_model_dir(eos_mode="never")
B = 256
ctx = P.Context(d, max_slots=B, kv_capacity=64)
eng = ctx.engine
lat = np.random.default_rng(B).standard_normal((B, 32)).astype(np.float32)
eng.mimi_reset(0, B)
for _ in range(3): eng.mimi_decode(0, B, lat)
ext = torch.cuda.ExternalStream(eng.stream_handle())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(ext)
for _ in range(200): eng.mimi_decode_enqueue(0, B)
e1.record(ext); eng.sync()
print("MIMI256", os.environ.get("PTTS_B200_PLAN", ""), round(e0.elapsed_time(e1) / 200, 4), "ms/step")
