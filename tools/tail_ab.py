"""A/B of the fused SEANet tail (seanet_tail.cuh: resnet block 9 + output conv) against the unfused launches: Mimi decode only, B utterances,
6 consecutive frames (the carried conv state crosses frames), same latents. Prints the worst-row PCM SNR per frame and the time per decode."""
import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import numpy as np, torch
import ptts_b200 as P
from make_assets import default_model_dir
d = default_model_dir(eos_mode="never")
B = int(os.environ.get("B", "37"))
out = {}
for mode in ("0", "1"):
    os.environ[os.environ.get("AB_VAR", "PTTS_B200_FUSED_TAIL")] = mode
    ctx = P.Context(d, max_slots=B, kv_capacity=64)
    eng = ctx.engine
    rng = np.random.default_rng(3)
    eng.mimi_reset(0, B)
    res = [eng.mimi_decode(0, B, rng.standard_normal((B, 32)).astype(np.float32) * 2.0).copy() for _ in range(6)]
    ext = torch.cuda.ExternalStream(eng.stream_handle())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(ext)
    for _ in range(100): eng.mimi_decode_enqueue(0, B)
    e1.record(ext); eng.sync()
    out[mode] = (res, e0.elapsed_time(e1) / 100)
    del ctx
for i in range(6):
    p0 = out["0"][0][i].astype(np.float64); p1 = out["1"][0][i]
    snr = 10 * np.log10((p0 ** 2).sum(1) / np.maximum(((p0 - p1) ** 2).sum(1), 1e-30))
    print(f"frame {i}: worst-row PCM SNR {snr.min():.1f} dB (rms {np.sqrt((p0 ** 2).mean()):.3e}), first samples max diff {np.abs(p0[:, :4] - p1[:, :4]).max():.2e}, finite {np.isfinite(p1).all()}")
print(f"B={B}: unfused {out['0'][1]:.4f} ms/decode, fused {out['1'][1]:.4f} ms/decode")
