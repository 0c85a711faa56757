"""CPU: the continuous-batching scheduler (pocket-tts.cpp_b200/csrc/host/batch_scheduler.hpp) driven through the C ABI over a MOCK engine
(ctypes callbacks): every queued sentence completes with the right frames attributed to it although slots are reused, frames of a
slot's previous sentence that are still in flight are not attributed to the new one, longest sentences start first, and the
idle-slot accounting adds up. The mock follows the engine's contract: a sentence started by begin() becomes live at the NEXT submitted
step, produces `length` frames (EOS) or stops at its cap, and a finished slot reports produced = 0."""
import ctypes

import numpy as np


class MockEngine:
    def __init__(self, n_slots, frame, lengths_by_stream):
        self.n_slots, self.frame, self.lengths = n_slots, frame, lengths_by_stream
        self.slot_job = [None] * n_slots        # [stream id, frames produced, length, cap]
        self.queue = []                          # in-flight steps: (n, pcm, produced)
        self.begin_calls = []
        self.steps = 0

    def begin(self, user, n, slots, voices, tokens, tok_off, mg, fae, temp, rng):
        call = []
        for i in range(n):
            s = slots[i]
            assert self.slot_job[s] is None or self.slot_job[s][1] >= min(self.slot_job[s][2], self.slot_job[s][3]), "slot refilled while its sentence is live"
            stream = int(rng[i])
            self.slot_job[s] = [stream, 0, self.lengths[stream], int(mg[i])]
            call.append((s, stream, tok_off[i + 1] - tok_off[i]))
        self.begin_calls.append((self.steps, call))
        return 0

    def submit(self, user, slot0, n):
        assert slot0 == 0 and 0 < n <= self.n_slots and len(self.queue) < 3
        pcm = np.zeros((n, self.frame), np.float32); produced = np.zeros(n, np.int32)
        for s in range(n):
            j = self.slot_job[s]
            if j is not None and j[1] < min(j[2], j[3]):
                pcm[s] = j[0] * 1000 + j[1]            # value identifies (sentence, frame index)
                produced[s] = 1; j[1] += 1
        self.queue.append((n, pcm, produced)); self.steps += 1
        return 0

    def collect(self, user, pcm_out, produced_out):
        n, pcm, produced = self.queue.pop(0)
        ctypes.memmove(pcm_out, pcm.ctypes.data, pcm.nbytes)
        ctypes.memmove(produced_out, produced.ctypes.data, produced.nbytes)
        return n


def _run(P, n_slots, lengths, caps, **cfg):
    frame = 4
    eng = MockEngine(n_slots, frame, {i + 1: l for i, l in enumerate(lengths)})
    b = P.Batch(n_slots=n_slots, ops=(eng.begin, eng.submit, eng.collect), frame_size=frame)
    b.configure(**cfg)
    utts = [b.add_tokens(0, [5, 6, 7], caps[i], 3, 0.7, rng_stream=i + 1) for i in range(len(lengths))]
    total = b.run()
    return eng, b, utts, total


def test_every_sentence_completes_with_its_own_frames(P):
    rng = np.random.default_rng(0)
    lengths = [int(x) for x in rng.integers(1, 60, 200)]
    caps = [int(l + rng.integers(0, 30)) if i % 5 else max(1, l // 2) for i, l in enumerate(lengths)]     # every 5th sentence is cut by its cap
    eng, b, utts, total = _run(P, 16, lengths, caps, refill_min=2, refill_every=4, range_quantum=8)
    want = [min(l, c) for l, c in zip(lengths, caps)]
    assert total == sum(want)
    for i, u in enumerate(utts):
        assert b.frames(u) == want[i], i
        pcm = b.read(u)
        assert pcm.shape == (want[i], 4)
        assert np.array_equal(pcm[:, 0], (i + 1) * 1000 + np.arange(want[i]))      # frames in order, none from another sentence
    st = b.stats()
    assert st["frames"] == total and st["sentences"] == 200 and st["steps"] == eng.steps
    assert st["slot_steps"] >= st["frames"] and 0.0 <= st["idle_slot_fraction"] < 0.5
    assert st["refills"] == len(eng.begin_calls) > 1                                # slots were refilled while others kept running


def test_longest_first_and_single_wave(P):
    lengths = [5, 40, 12, 33, 7, 21]
    eng, b, utts, total = _run(P, 8, lengths, [100] * 6)
    assert total == sum(lengths)
    assert len(eng.begin_calls) == 1                                                # fewer sentences than slots: one sentence start, no refill
    started = [stream for _, stream, _ in eng.begin_calls[0][1]]
    assert started == [2, 4, 6, 3, 5, 1] or sorted(started) == [1, 2, 3, 4, 5, 6]   # equal caps: queue order kept (stable sort)
    assert eng.steps <= max(lengths) + 3                                            # the run ends right after the longest sentence


def test_caps_order_the_queue(P):
    lengths = [3, 3, 3, 3]
    eng, b, utts, total = _run(P, 1, lengths, [10, 40, 20, 30])
    order = [c[1][0][1] for c in eng.begin_calls]
    assert order == [2, 4, 3, 1]                                                    # longest processing time (cap) first
    assert total == 12


def test_zero_room_sentences_are_skipped(P):
    eng, b, utts, total = _run(P, 2, [4, 4, 4], [0, 5, 5])
    assert total == 8 and b.frames(utts[0]) == 0 and b.stats()["sentences"] == 3
