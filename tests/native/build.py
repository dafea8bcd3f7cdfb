"""Builds the host-side TEST library of the K7 ByteTrack core (tests/native/bt_host.cpp) with g++.
-ffp-contract=off: the device library is built with --fmad=false, the test build must round the same way."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libbt_host.so")
SRC = os.path.join(HERE, "bt_host.cpp")
CORE = os.path.join(ROOT, "hockey-vision-analytics_b200", "csrc", "k7_bytetrack_core.h")


def build() -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(SRC), os.path.getmtime(CORE)):
        return OUT
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-I", os.path.dirname(CORE), SRC, "-o", OUT]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("g++ failed:\n%s\n%s" % (p.stdout, p.stderr))
    return OUT


if __name__ == "__main__":
    print(build())
