"""CPU: utterance sharding logic and the world_size-2 stats gather (gloo)."""
import importlib.util
import os
import subprocess
import sys
import textwrap

from conftest import REPO


def _sharding():
    p = os.path.join(REPO, "pocket-tts.cpp_b200", "sharding.py")
    spec = importlib.util.spec_from_file_location("ptts_sharding", p)
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m


def test_shards_partition_and_balance():
    S = _sharding()
    import random
    rnd = random.Random(0)
    costs = [S.estimate_frames(rnd.randint(3, 45)) for _ in range(2048)]
    for w in (1, 2, 4, 8):
        sh = S.shard_utterances(costs, w)
        flat = sorted(i for s in sh for i in s)
        assert flat == list(range(2048))                      # every utterance exactly once
        assert max(len(s) for s in sh) - min(len(s) for s in sh) <= 8
        assert S.imbalance(costs, sh) < 0.01                  # the 7.5x/8 target tolerates 6 %
        assert sh == S.shard_utterances(costs, w)             # deterministic
    assert S.shard_utterances([], 4) == [[], [], [], []]
    assert S.estimate_frames(9) == 137                        # bench sentence: 9 words -> cap 137 frames


def test_gather_stats_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import importlib.util, os, sys, torch, torch.distributed as dist
        spec = importlib.util.spec_from_file_location("s", r"{os.path.join(REPO, 'pocket-tts.cpp_b200', 'sharding.py')}")
        S = importlib.util.module_from_spec(spec); spec.loader.exec_module(S)
        dist.init_process_group("gloo")
        r = dist.get_rank()
        shards = S.shard_utterances([10, 20, 30, 40, 50, 60], dist.get_world_size())
        mine = shards[r]
        g = S.gather_stats([float(len(mine)), float(sum(mine)), 100.0 + r])
        assert g.shape == (2, 3), g.shape
        assert g[:, 0].sum().item() == 6 and g[:, 1].sum().item() == 15, g
        assert g[:, 2].max().item() == 101.0
        dist.destroy_process_group()
        print("ok", r)
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29617", str(script)], capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


def test_ragged_job_is_rank_count_invariant_gloo_world2(tmp_path):
    """The multi-GPU path of the continuous-batching workload on CPU (gloo, world 2): the global sentence list is LPT-sharded, every rank
    runs the C++ scheduler over its shard (mock engine: the frames of a sentence are a function of its global RNG stream id only), the
    stats are gathered — and every sentence gets exactly the frames it gets in a single-rank run."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import ctypes, importlib.util, json, os, sys
        import numpy as np, torch, torch.distributed as dist
        sys.path.insert(0, r"{REPO}")
        import ptts_b200 as P
        spec = importlib.util.spec_from_file_location("s", r"{os.path.join(REPO, 'pocket-tts.cpp_b200', 'sharding.py')}")
        S = importlib.util.module_from_spec(spec); spec.loader.exec_module(S)

        N, SLOTS, FRAME = 300, 16, 2
        words = [3 + (i * 7) % 43 for i in range(N)]
        caps = [S.estimate_frames(w) for w in words]
        length = lambda gid: 5 + (gid * 37) % 200            # "EOS" frame of a sentence: depends on the sentence only

        def run(ids):
            slot = [None] * SLOTS; queue = []
            def begin(user, n, slots, voices, toks, off, mg, fae, temp, rng):
                for i in range(n):
                    slot[slots[i]] = [int(rng[i]), 0, min(length(int(rng[i])), int(mg[i]))]
                return 0
            def submit(user, s0, n):
                pcm = np.zeros((n, FRAME), np.float32); prod = np.zeros(n, np.int32)
                for s in range(n):
                    j = slot[s]
                    if j is not None and j[1] < j[2]:
                        pcm[s] = j[0]; prod[s] = 1; j[1] += 1
                queue.append((n, pcm, prod)); return 0
            def collect(user, po, fo):
                n, pcm, prod = queue.pop(0)
                ctypes.memmove(po, pcm.ctypes.data, pcm.nbytes); ctypes.memmove(fo, prod.ctypes.data, prod.nbytes); return n
            b = P.Batch(n_slots=SLOTS, ops=(begin, submit, collect), frame_size=FRAME)
            utts = {{g: b.add_tokens(0, [1, 2, 3], caps[g], 3, 0.7, rng_stream=g + 1) for g in ids}}
            total = b.run()
            return {{g: b.frames(u) for g, u in utts.items()}}, total, b.stats()

        dist.init_process_group("gloo")
        r, w = dist.get_rank(), dist.get_world_size()
        shards = S.shard_utterances(caps, w)
        frames, total, st = run(shards[r])
        want = {{g: min(length(g + 1), caps[g]) for g in shards[r]}}
        assert frames == want, (r, [(g, frames[g], want[g]) for g in frames if frames[g] != want[g]][:5])
        g = S.gather_stats([float(total), float(st["steps"]), float(sum(caps[i] for i in shards[r]))])
        assert g.shape == (w, 3)
        assert int(g[:, 0].sum()) == sum(min(length(i + 1), caps[i]) for i in range(N))      # all ranks together = the single-rank job
        assert g[:, 2].max() / g[:, 2].mean() - 1.0 < 0.01                                  # LPT balance on the caps
        dist.destroy_process_group()
        print("ok", r, total)
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29627", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2
