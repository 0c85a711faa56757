"""Sentence start of the bench batch (256 paragraphs, 1345-row shared prefix) for different prefill chunk sizes (b200_config.max_prefill_rows)."""
import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import numpy as np, torch
import ptts_b200 as P
import bench
from make_assets import default_model_dir
B = 256
d = default_model_dir(eos_mode="never", t_voice=1345, voices=["cosette"])
for rows in (2048, 4096, 8192, 16384):
    ctx = P.Context(d, max_slots=B, kv_capacity=2048, max_prefill_rows=rows)
    eng = ctx.engine
    st = ctx.stream("cosette", temp=0.7)
    toks = [ctx.tokenize(bench.synth_paragraph(i)) for i in range(B)]
    def begin():
        eng.begin_sentences(list(range(B)), [st.voice] * B, toks, [2048] * B, [1 << 20] * B, [0.7] * B)
    ext = torch.cuda.ExternalStream(eng.stream_handle())
    begin(); eng.sync()
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record(ext); begin(); eng.join(); b1.record(ext); eng.sync()
    print(f"PREFILL_ROWS {rows}: sentence start {b0.elapsed_time(b1):.3f} ms for {sum(len(t) for t in toks)} token rows")
    del ctx
