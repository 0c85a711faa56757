#!/usr/bin/env python3
"""All five BASELINE.json configs in one report (JSON lines), complementing bench.py (which measures configs[3]/[4]).

  1. CPU `--bench` shape: oracle, voice cosette, the bench sentence, batch 1 (reference ggml-CPU-style path)
  2. single-GPU batch-1 streaming through the reference API: first-frame latency, frames/s
  3. Mimi decoder only: batch sweep 1..256 of synthetic latent sequences
  4. full pipeline at batch 256, KV ~1.5k           -> bench.py
  5. utterance-sharded multi-GPU                    -> bench.py --gpus N under torchrun
Usage: python tools/bench_configs.py [--skip-cpu] [--mimi-frames 200]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tools"))
from make_assets import default_model_dir  # noqa: E402

SENT = "The quick brown fox jumped over the sleeping dog."


def cfg1_cpu():
    import oracle
    d = default_model_dir(eos_mode="never")
    cores = os.cpu_count() or 1
    o = oracle.Oracle(d, threads=cores)
    s = o.stream("cosette", kv_capacity=1000)
    t0 = time.perf_counter()
    s.sentence_init(SENT)                     # text prefill is inside the reference's measured time (demos/pocket-tts.cpp:474-520)
    n = 0
    while n < 60:
        ok, *_ = s.step(None)
        if not ok:
            break
        n += 1
    dt = time.perf_counter() - t0
    return {"config": 1, "what": "CPU oracle (reference ggml-CPU-style path), batch 1, bench sentence, temp 0", "frames": n, "frames_per_s": round(n / dt, 2),
            "rtf": round(12.5 / (n / dt), 3), "threads": cores}


def cfg2_stream():
    import ptts_b200 as P
    d = default_model_dir(eos_mode="never")
    ctx = P.Context(d, max_slots=1, kv_capacity=1024)
    P.set_seed(0)
    st = ctx.stream("cosette", temp=0.0)
    out = {}
    for rep in range(3):                      # rep 0 warms kernels / captures the CUDA graph
        st.reset()
        t0 = time.perf_counter()
        st.send(SENT); st.flush()
        first = None; n = 0
        while True:
            f = st.receive()
            if f is None:
                break
            n += 1
            if first is None:
                first = time.perf_counter() - t0
        dt = time.perf_counter() - t0
        out = {"config": 2, "what": "1 GPU, batch-1 streaming through ptts_stream_send/receive, bench sentence, temp 0", "frames": n,
               "first_frame_latency_ms": round(first * 1e3, 3), "frames_per_s": round(n / dt, 1), "rtf": round(12.5 / (n / dt), 5),
               "weight_bytes_roofline_frames_per_s": round(6455.6e9 / 208e6, 0)}
    return out


def cfg3_mimi(frames):
    import torch
    import ptts_b200 as P
    d = default_model_dir(eos_mode="never")
    res = []
    for B in (1, 4, 16, 64, 256):
        ctx = P.Context(d, max_slots=B, kv_capacity=64)
        eng = ctx.engine
        rng = np.random.default_rng(B)
        lat = rng.standard_normal((B, 32)).astype(np.float32)
        eng.mimi_reset(0, B)
        eng.mimi_decode(0, B, lat); eng.mimi_decode(0, B, lat); eng.mimi_decode(0, B, lat)   # warm + graph capture
        ext = torch.cuda.ExternalStream(eng.stream_handle())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(ext)
        for _ in range(frames):
            eng.mimi_decode_enqueue(0, B)
        e1.record(ext)
        eng.sync()
        ms = e0.elapsed_time(e1)
        fps = B * frames / (ms * 1e-3)
        flops = 541.5e6 * fps
        res.append({"batch": B, "frames_per_s": round(fps, 1), "ms_per_step": round(ms / frames, 4), "tflops": round(flops / 1e12, 2),
                    "frac_bf16_sustained": round(flops / 1418e12, 4)})
        del ctx
    return {"config": 3, "what": f"Mimi decoder only, {frames} frames of synthetic latents per utterance, device-resident", "sweep": res}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--mimi-frames", type=int, default=200)
    a = ap.parse_args()
    if not a.skip_cpu:
        print(json.dumps(cfg1_cpu()), flush=True)
    print(json.dumps(cfg2_stream()), flush=True)
    print(json.dumps(cfg3_mimi(a.mimi_frames)), flush=True)
