"""Builds libptts_b200.so (sm_100a only) in-tree with nvcc. Usage: python pocket-tts.cpp_b200/build.py [--force]"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "lib")
OUT = os.path.join(OUT_DIR, "libptts_b200.so")
CLI = os.path.join(OUT_DIR, "pocket-tts-b200")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
SOURCES = ["engine.cu", "host_api.cpp"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden",
         "-ccbin", "/usr/bin/g++", "-diag-suppress", "550,177"]


def _deps():
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files]
    out += [os.path.join(HERE, "..", "include", "ptts_b200.h"), os.path.join(HERE, "..", "include", "pocket_tts", "pocket_tts.h")]
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in _deps()):
        return OUT
    objs = []
    for s in SOURCES:
        o = os.path.join(OUT_DIR, os.path.splitext(s)[0] + ".o")
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, s), "-o", o]
        subprocess.check_call(cmd)
        objs.append(o)
    subprocess.check_call([NVCC, "-shared", "-o", OUT] + objs + ["-ccbin", "/usr/bin/g++", "-lcudart_static", "-lpthread", "-ldl", "-lrt"])
    # command-line driver over the reference API: plain C++, linked against the library just built
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", os.path.join(CSRC, "cli_main.cpp"), "-o", CLI, "-L" + OUT_DIR, "-lptts_b200",
                           "-Wl,-rpath,$ORIGIN"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
