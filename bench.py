#!/usr/bin/env python3
"""bench.py — audio frames/s of the Pocket-TTS per-frame generation path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch 256] [--kv-len 1500]

A "step" is one pass of the hot path over one batch: FlowLM step -> EOS rule -> LSD head -> Mimi -> 1920 samples, for every
utterance of the batch (BASELINE.json configs[3]: batch 256, ~40-word paragraphs, KV length ~1.5k; voice prefix padded so the
FlowLM cache holds ~kv_len positions in mid-run). One process per GPU (torchrun for N>1), weights replicated, utterances sharded,
no collective on the hot path; NCCL gathers per-rank counts/timings only. Prints ONE JSON line on rank 0.

  value     device-resident throughput: K steps enqueued on the engine stream, CUDA events, max over ranks
  e2e       same metric through the public C-ABI call b200_step() with HOST buffers (H2D noise, D2H PCM + flags inside)
  roofline  dominant kernel (FlowLM decode attention, KV-stream bound): algorithmic bytes / CUDA-event time vs measured HBM peak
  cpu_baseline  the oracle (CPU restatement of the reference's ggml path) on this box's host cores, bounded sample
  --impl reference : the same oracle as the reference arm (the real reference cannot be built: ggml et al. absent)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tools"))

METRIC = "audio_frames_per_sec"
UNIT = "frames/s"
PARAGRAPH_WORDS = 40


def synth_paragraph(i: int, n_words: int = PARAGRAPH_WORDS) -> str:
    from make_assets import COMMON_WORDS
    rng = np.random.default_rng(1000 + i)
    ws = [COMMON_WORDS[int(k)] for k in rng.integers(0, len(COMMON_WORDS), n_words)]
    return (" ".join(ws)).capitalize() + "."


def peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device: int):
        self.device = device; self.rows = []; self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.device)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
def run_reference(args, rank: int, world: int):
    """Reference arm / cpu_baseline: the CPU oracle on all host cores, one utterance of the same workload."""
    if rank != 0:
        return None
    import oracle
    from make_assets import default_model_dir
    oracle.build()
    d = default_model_dir(eos_mode="never", t_voice=args.t_voice, voices=["cosette"])
    cores = os.cpu_count() or 1
    o = oracle.Oracle(d, threads=cores)
    s = o.stream("cosette", kv_capacity=args.kv_capacity)
    text = synth_paragraph(0)
    s.sentence_init(text)
    sample = args.ref_frames_per_step
    for _ in range(args.warmup):
        for _ in range(sample):
            s.step(None)
    t0 = time.perf_counter()
    n = 0
    for _ in range(args.steps):
        for _ in range(sample):
            ok, *_ = s.step(None)
            n += int(ok)
    dt = time.perf_counter() - t0
    fps = n / dt
    desc = f"1 utterance (batch 1, ggml-CPU-style, {cores} threads), {sample} frames per step at FlowLM KV length ~{s.current_end}, temp 0"
    return {"value": fps, "ms_per_step": dt * 1e3 / args.steps, "cores": cores, "sample": desc, "frames": n}


def cpu_baseline_quick(args):
    """~10-20 s of oracle work on rank 0 (N=1 only)."""
    a = argparse.Namespace(**vars(args)); a.steps = 10; a.warmup = 1; a.ref_frames_per_step = 4
    r = run_reference(a, 0, 1)
    return {"value": round(r["value"], 2), "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}


# ------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="utterances per GPU")
    ap.add_argument("--kv-len", type=int, default=1500, help="FlowLM cache positions per utterance in mid-run")
    ap.add_argument("--kv-capacity", type=int, default=2048)
    ap.add_argument("--kv-f32", type=int, default=0)
    ap.add_argument("--gemm-path", type=int, default=0)
    ap.add_argument("--pdl", type=int, default=1)
    ap.add_argument("--cuda-graphs", type=int, default=1)
    ap.add_argument("--overlap", type=int, default=1, help="two-stream pipeline: Mimi(t) overlaps FlowLM(t+1)")
    ap.add_argument("--ref-frames-per-step", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    # tokens of a 40-word paragraph ~ 60-75; voice prefix sized so that voice + text + steps/2 ~ kv_len
    args.untimed = max(args.warmup, 100)             # extended warm-up: also gives the clock sampler a loaded window
    args.t_voice = max(16, args.kv_len - 45 - args.untimed - args.steps // 2)

    if args.impl == "reference":
        if rank != 0:
            return 0
        r = run_reference(args, rank, world)
        line = {"impl": "reference", "metric": METRIC, "value": round(r["value"], 3), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(r["ms_per_step"], 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"configs[3]: full FlowLM+head+Mimi, ~{PARAGRAPH_WORDS}-word paragraphs, KV length ~{args.kv_len}", "batch_per_gpu": 1,
                           "kv_len": args.kv_len},
                "cpu_baseline": {"value": round(r["value"], 3), "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
                "e2e": {"value": round(r["value"], 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import ptts_b200 as P
    from make_assets import default_model_dir
    if rank == 0:
        P.build()
        d = default_model_dir(eos_mode="never", t_voice=args.t_voice, voices=["cosette"])
    if world > 1:
        dist.barrier()
    d = default_model_dir(eos_mode="never", t_voice=args.t_voice, voices=["cosette"])

    B = args.batch
    ctx = P.Context(d, device=local, max_slots=B, max_voices=1, kv_capacity=args.kv_capacity, kv_f32=args.kv_f32, gemm_path=args.gemm_path, pdl=args.pdl, cuda_graphs=args.cuda_graphs, overlap=args.overlap)
    eng = ctx.engine
    st = ctx.stream("cosette", temp=0.7)            # prefill of the (long) voice prefix
    # this rank's utterance slice: global utterance id = rank * B + i
    texts = [synth_paragraph(rank * B + i) for i in range(B)]
    toks = [ctx.tokenize(t) for t in texts]
    total_steps = args.untimed + args.steps * 3 + 16  # warm-up + timed + e2e (+ its warm-up) + profiled passes
    eng.set_seed(1234 + rank)
    eng.begin_sentences(list(range(B)), [st.voice] * B, toks, [total_steps + 64] * B, [1 << 20] * B, [0.7] * B)
    eng.sync()
    L0 = [args.t_voice + len(t) for t in toks]

    ext = torch.cuda.ExternalStream(eng.stream_handle(), device=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks are sampled from the warm-up through the timed region (the timed region alone can be shorter than one sample period)
    sampler = ClockSampler(local); sampler.start()
    time.sleep(0.3)
    steps_done = 0
    for _ in range(args.untimed):
        eng.step_enqueue(0, B); steps_done += 1
    eng.join()
    eng.sync()
    barrier()
    l0 = eng.launch_count()
    if rank == 0:
        print(f"[bench] launches_before_timed_region={l0}", file=sys.stderr, flush=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ncu_range = os.environ.get("PTTS_NCU_RANGE") == "1"    # `ncu --profile-from-start off`: capture only the timed steps
    if ncu_range:
        P.lib().b200_profiler_range(1)
    e0.record(ext)
    for _ in range(args.steps):
        eng.step_enqueue(0, B)
    eng.join()                                       # the Mimi stream's last frame is inside the timed region
    e1.record(ext)
    eng.sync()
    if ncu_range:
        P.lib().b200_profiler_range(0)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - l0
    clocks = sampler.stop()
    mid_step = steps_done + args.steps / 2.0
    steps_done += args.steps
    t_ms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())

    # ---- end-to-end through the public C-ABI call with host buffers ----
    e2e = None
    if not args.no_e2e:
        rng = np.random.default_rng(5 + rank)
        noise = (rng.standard_normal((B, 32)) * np.sqrt(0.7)).astype(np.float32)
        pcm = np.zeros((B, P.FRAME), np.float32); produced = np.zeros(B, np.int32)
        eng.step_into(0, B, noise, pcm, produced); steps_done += 1
        if args.overlap:                                 # warm-up of the pipelined call pair: first use runs eagerly, second captures its graphs
            for _ in range(3):
                eng.submit(0, B, noise); eng.submit(0, B, noise)
                eng.collect_into(pcm, produced); eng.collect_into(pcm, produced)
                steps_done += 2
        barrier()
        t0 = time.perf_counter()
        if args.overlap:
            # the pipelined public call pair: submits stay two frames ahead of collects (at most three frames in flight)
            depth = min(2, args.steps)
            for _ in range(depth):
                eng.submit(0, B, noise)
            for _ in range(args.steps - depth):
                eng.submit(0, B, noise)
                eng.collect_into(pcm, produced)
            for _ in range(depth):
                eng.collect_into(pcm, produced)
        else:
            for _ in range(args.steps):
                eng.step_into(0, B, noise, pcm, produced)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        steps_done += args.steps
        assert produced.all(), "bench utterances must stay active (never-EOS checkpoint)"
        t_e = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        e2e = {"value": round(B * world * args.steps / float(t_e.item()), 2), "unit": UNIT, "h2d_bytes_per_step": int(noise.nbytes),
               "d2h_bytes_per_step": int(pcm.nbytes + produced.nbytes),
               "api": "b200_submit/b200_collect (submits kept two frames ahead of collects)" if args.overlap else "b200_step (synchronous)"}

    # ---- roofline of the dominant kernel: profiled pass with event pairs around every FlowLM attention launch ----
    eng.profile(True)
    prof_step0 = steps_done
    for _ in range(args.steps):
        eng.step_enqueue(0, B)
    eng.sync()
    prof = eng.profile_read()
    steps_done += args.steps
    hbm, tflops, which = peaks()
    elt = 4 if args.kv_f32 else 2
    attn_ms, attn_n = prof["attn_flow"]
    # algorithmic bytes per attention launch (one layer, all utterances): K and V rows [0, L] once + q in / out
    Lmid = np.array(L0, np.float64) + prof_step0 + args.steps / 2.0
    bytes_per_launch = float((2 * (Lmid + 1) * 1024 * elt).sum() + B * 1024 * (4 + 2))
    roof = None
    # DRAM traffic of one attention launch as measured by ncu (--set full capture summarised in profiles/attn_flow_traffic.json by
    # tools/ncu_kernel_report.py); scaled by the ratio of algorithmic bytes now / at capture, since it is per launch like `achieved`
    traffic = None
    tp = os.path.join(REPO, "profiles", "attn_flow_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic = int((tj["dram_bytes_read"] + tj["dram_bytes_write"]) * bytes_per_launch / tj["algorithmic_bytes_at_capture"])
    if attn_n:
        achieved = bytes_per_launch / (attn_ms / attn_n * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "attn_flow (FlowLM decode attention, one layer)", "achieved": round(achieved, 1), "peak": hbm, "unit": "GB/s",
                "frac": round(achieved / hbm, 4), "traffic": traffic, "peak_source": which, "algorithmic_bytes_per_launch": int(bytes_per_launch),
                "avg_launch_ms": round(attn_ms / attn_n, 4),
                "note": "peak = driver-measured copy bandwidth (read+write); a read-only stream such as this kernel can exceed it slightly",
                "share_of_step": round(attn_ms / max(prof["step"][0], 1e-9), 4),
                "segments_ms_per_step": {k: round(v[0] / args.steps, 4) for k, v in prof.items()}}

    frames_total = B * world * args.steps
    value = frames_total / (ms_max * 1e-3)
    # whole-step roofline bound for context: weights once + per-utterance KV/state bytes
    Lm = np.array(L0, np.float64) + mid_step
    step_bytes = 189.6e6 + float((12 * (Lm + 1) * 1024 * elt).sum()) + B * (1.09e6 + 0.12e6 + 7.7e3)
    bound_fps = B / (step_bytes / (hbm * 1e9))

    if world > 1:
        sys.path.insert(0, os.path.join(REPO, "pocket-tts.cpp_b200"))
        import sharding
        g = sharding.gather_stats([B * args.steps, ms], device="cuda")
    line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_max / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"configs[3]: full FlowLM+head+Mimi, batch {B}/GPU of ~{PARAGRAPH_WORDS}-word paragraphs, FlowLM KV length ~{int(Lm.mean())} in mid-run",
                       "batch_per_gpu": B, "kv_len": int(Lm.mean()), "kv_dtype": "f32" if args.kv_f32 else "bf16", "kv_capacity": args.kv_capacity,
                       "t_voice": args.t_voice, "parallelism": f"utterance-sharded x{world}", "l2": "inputs larger than L2 (KV stream ~%.1f GB/step)" % (step_bytes / 1e9),
                       "rtf_per_utterance": round(12.5 / (value / (B * world)), 4), "roofline_bound_frames_per_s_per_gpu": round(bound_fps, 1),
                       "frac_of_step_roofline": round((value / world) / bound_fps, 4)},
            "clocks": clocks, "gpu_launches": int(launches)}
    if e2e:
        line["e2e"] = e2e
    if roof:
        line["roofline"] = roof
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_quick(args)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
