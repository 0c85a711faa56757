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
