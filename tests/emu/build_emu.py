"""TEST INFRASTRUCTURE ONLY: compile the kernel sources as host C++ against tests/emu/cnb_emu.h."""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
CSRC = ROOT / "cultionet_b200" / "csrc"
LIB_PATH = HERE / "libcnb_emu.so"


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    deps = [CSRC / "cnb_api.cu", *CSRC.glob("*.cuh"), HERE / "cnb_emu.h", ROOT / "include" / "cultionet_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False) -> Path:
    if not force and not needs_build():
        return LIB_PATH
    cmd = [
        "g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-DCNB_EMU", "-pthread",
        "-I/usr/local/cuda/include", "-include", str(HERE / "cnb_emu.h"),
        "-x", "c++", str(CSRC / "cnb_api.cu"), "-o", str(LIB_PATH),
        "-Wno-attributes", "-Wno-unknown-pragmas",
    ]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building the emulator library")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
