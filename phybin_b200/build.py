"""In-tree builds of the native pieces (no pip, no JIT cache: the .so files travel with the repo
snapshot to the GPU box).

  libphybin_rf.so    CUDA library behind include/phybin_rf.h   (nvcc, sm_100a only)
  libphybin_host.so  C++ host side behind include/phybin_host.h (g++)
  phybin             C++ command-line host (links both)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from typing import List

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INC = os.path.join(ROOT, "include")
LIBDIR = os.path.join(PKG, "lib")

RF_SO = os.path.join(LIBDIR, "libphybin_rf.so")
HOST_SO = os.path.join(LIBDIR, "libphybin_host.so")
CLI = os.path.join(LIBDIR, "phybin")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-Xptxas", "-v",
]


def _newer(target: str, sources: List[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(sub: str, exts) -> List[str]:
    d = os.path.join(CSRC, sub) if sub else CSRC
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(tuple(exts)))


def _run(cmd: List[str], log: str | None = None) -> None:
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")
    return p


def build_host(force: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    src = _sources("host", [".cpp"])
    src = [s for s in src if not s.endswith("phybin_main.cpp")]
    deps = src + [os.path.join(INC, "phybin_host.h")]
    if force or _newer(HOST_SO, deps):
        _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wall", "-Wextra",
              "-fvisibility=hidden", "-I", INC, "-o", HOST_SO] + src)
    return HOST_SO


def build_cuda(force: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    cu = _sources("", [".cu"])
    deps = cu + _sources("", [".cuh", ".h"]) + [os.path.join(INC, "phybin_rf.h")]
    if force or _newer(RF_SO, deps):
        _run([nvcc_path()] + NVCC_FLAGS + ["-shared", "-I", INC, "-I", CSRC, "-o", RF_SO] + cu
             + ["-lcudart"], log=os.path.join(LIBDIR, "nvcc_build.log"))
    return RF_SO


def build_cli(force: bool = False) -> str:
    build_host(force)
    build_cuda(force)
    main = os.path.join(CSRC, "host", "phybin_main.cpp")
    if not os.path.exists(main):
        return ""
    if force or _newer(CLI, [main, RF_SO, HOST_SO]):
        _run(["g++", "-O2", "-std=c++17", "-Wall", "-Wextra", "-I", INC, "-o", CLI, main,
              "-L", LIBDIR, "-lphybin_host", "-lphybin_rf", "-Wl,-rpath,$ORIGIN", "-pthread"])
    return CLI


def build_all(force: bool = False) -> None:
    build_host(force)
    build_cuda(force)
    build_cli(force)


if __name__ == "__main__":
    build_all("--force" in sys.argv)
    print("built:", RF_SO, HOST_SO, CLI)
