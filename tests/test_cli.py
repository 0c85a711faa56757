"""The command-line driver over the reference API: same result lines as `pocket-tts --bench`
(reference demos/pocket-tts.cpp:454,517-520) and a WAV sink."""
import os
import re
import struct
import subprocess

import pytest

from conftest import REPO

CLI = os.path.join(REPO, "pocket-tts.cpp_b200", "lib", "pocket-tts-b200")


def test_cli_builds_and_prints_usage(P):
    assert os.path.exists(CLI)
    r = subprocess.run([CLI, "--help"], capture_output=True, text=True)
    assert r.returncode == 1 and "--bench" in r.stderr and "--voice" in r.stderr
    r = subprocess.run([CLI, "--nonsense"], capture_output=True, text=True)
    assert r.returncode == 1 and "unrecognized option" in r.stderr


@pytest.mark.gpu
def test_cli_bench_lines_and_wav(P, model_dir, oracle_mod, tmp_path):
    wav = str(tmp_path / "out.wav")
    r = subprocess.run([CLI, "-m", model_dir, "--bench", "-o", wav], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = r.stdout
    assert re.search(r"^seed: 0$", out, re.M)
    assert re.search(r"^done generating\. \d+\.\d+$", out, re.M)
    frames = int(re.search(r"^frame count:\s+(\d+) frames$", out, re.M).group(1))
    assert float(re.search(r"^frame rate:\s+([\d.]+) frames/s$", out, re.M).group(1)) > 0
    # never-EOS checkpoint: the bench sentence (9 words) runs to its cap int((9 + 2) * 12.5) = 137 frames (src/pocket_tts.cpp:429-430)
    assert frames == oracle_mod.max_gen_len_for("The quick brown fox jumped over the sleeping dog.") == 137
    b = open(wav, "rb").read()
    assert b[:4] == b"RIFF" and b[8:16] == b"WAVEfmt "
    fmt, ch, rate, _, _, bits = struct.unpack("<HHIIHH", b[20:36])
    assert (fmt, ch, rate, bits) == (1, 1, 24000, 16)
    assert struct.unpack("<I", b[40:44])[0] == frames * 1920 * 2 == len(b) - 44
