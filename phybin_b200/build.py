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


def _includes(path: str, seen=None) -> List[str]:
    """Transitive quoted includes of a source file inside csrc/ and include/ (its rebuild dependencies)."""
    seen = seen if seen is not None else set()
    if path in seen or not os.path.exists(path):
        return []
    seen.add(path)
    out = [path]
    with open(path) as f:
        for ln in f:
            ln = ln.strip()
            if ln.startswith('#include "'):
                name = ln.split('"')[1]
                for d in (CSRC, INC):
                    out += _includes(os.path.join(d, name), seen)
    return out


def build_cuda(force: bool = False, extra_flags: List[str] | None = None, out: str | None = None) -> str:
    """One object per .cu translation unit, compiled in parallel (only the units whose sources changed), then linked."""
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(LIBDIR, exist_ok=True)
    out = out or RF_SO
    tag = "" if out == RF_SO else "." + os.path.basename(out)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    cu = _sources("", [".cu"])
    nvcc = nvcc_path()
    jobs, objs = [], []
    for src in cu:
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + tag + ".o")
        objs.append(obj)
        if force or _newer(obj, _includes(src)):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        _run([nvcc] + NVCC_FLAGS + (extra_flags or []) + ["-c", "-I", INC, "-I", CSRC, "-o", obj, src], log=obj[:-2] + ".log")

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(compile_one, jobs))
    if jobs or not os.path.exists(out):
        _run([nvcc, "-shared", "-o", out] + objs + ["-lcudart"])
        with open(os.path.join(LIBDIR, "nvcc_build.log"), "w") as lf:   # one log, as before: ptxas -v of every unit
            for obj in objs:
                lg = obj[:-2] + ".log"
                if os.path.exists(lg):
                    with open(lg) as f:
                        lf.write(f.read())
    return out


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
