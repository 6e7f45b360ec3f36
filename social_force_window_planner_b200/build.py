"""Build libsfw_b200.so (the C-ABI library: hand-written sm_100a kernels + host packer) in-tree.

    python -m social_force_window_planner_b200.build

nvcc cross-compiles for sm_100a without a GPU.  The .so is git-ignored but travels with the repo
snapshot to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsfw_b200.so")
SOURCES = ["sfw_kernels.cu", "sfw_crowd.cu", "sfw_abi.cu", "sfw_sensor.cu", "sfw_exchange.cu"]
HEADERS = ["sfw_dev.h", "sfw_kernels.h", "sfw_forces.cuh", "sfw_ctx.h", os.path.join("..", "..", "include", "sfw_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--shared", "-cudart", "static", "-diag-suppress", "177",
]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        [os.path.join(CSRC, f) for f in SOURCES] + ["-o", LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libsfw_b200.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


HOST_LIB = os.path.join(HERE, "libsfw_planner_host.so")
HOST_SRC = os.path.join(HERE, "host", "sfw_planner_host.cpp")


def build_host(force: bool = False) -> str:
    """libsfw_planner_host.so: the C++ mirror of the reference's SFWPlanner (host control flow + scene
    packer) linked against libsfw_b200.so."""
    build(force=False)
    sensor_src = os.path.join(HERE, "host", "sfw_sensor_host.cpp")
    node_src = os.path.join(HERE, "host", "sfw_node_host.cpp")
    deps = [HOST_SRC, HOST_SRC.replace(".cpp", ".hpp"), sensor_src, sensor_src.replace(".cpp", ".hpp"),
            node_src, node_src.replace(".cpp", ".hpp"),
            os.path.join(HERE, "..", "include", "sfw_b200.h"), LIB]
    if not force and os.path.exists(HOST_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(HOST_LIB) for d in deps):
        return HOST_LIB
    cmd = [os.environ.get("CXX", "g++"), "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-Wextra",
           "-o", HOST_LIB, HOST_SRC, sensor_src, node_src, "-L" + HERE, "-l:libsfw_b200.so", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building libsfw_planner_host.so")
    return HOST_LIB


if __name__ == "__main__":
    build_host(force="--force" in sys.argv)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
