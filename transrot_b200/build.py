"""Builds libtransrot_b200.so (sm_100a only) in-tree with nvcc.  Used by __graft_entry__.build()."""
from __future__ import annotations

import glob
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(HERE, "libtransrot_b200.so")
SOURCES = ["transrot_b200.cu", "crew_launch.cu", "crew2_launch.cu", "team_pc_launch.cu", "multi.cu"]
OBJDIR = os.path.join(HERE, "build")


def _inputs():
    """Everything the library is compiled from: every file under csrc/ and every public header, so that a new
    .cuh can never be missed by the staleness test."""
    return sorted(glob.glob(os.path.join(CSRC, "*"))) + sorted(glob.glob(os.path.join(INCLUDE, "*.h"))) + [os.path.abspath(__file__)]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in _inputs())


def nccl_paths():
    """Header and library of the NCCL the multi-GPU entry points (tr_create_multi / tr_gather_minima) link against:
    the copy bundled with torch (nvidia/nccl wheel), else the system one."""
    cands = []
    try:
        import nvidia.nccl as _n  # type: ignore

        base = os.path.dirname(_n.__file__) if getattr(_n, "__file__", None) else list(_n.__path__)[0]
        cands.append((os.path.join(base, "include"), os.path.join(base, "lib")))
    except Exception:
        pass
    for inc, lib in (("/usr/include", "/usr/lib/x86_64-linux-gnu"), ("/usr/local/cuda/include", "/usr/local/cuda/lib64")):
        cands.append((inc, lib))
    for inc, lib in cands:
        if os.path.exists(os.path.join(inc, "nccl.h")):
            libs = sorted(glob.glob(os.path.join(lib, "libnccl.so*")))
            if libs:
                return inc, lib, libs[0]
    return None


def _deps(path: str, seen=None):
    """The file plus every quoted #include reachable from it (resolved relative to the including file)."""
    seen = set() if seen is None else seen
    path = os.path.normpath(path)
    if path in seen or not os.path.exists(path):
        return seen
    seen.add(path)
    for m in re.finditer(r'^\s*#\s*include\s+"([^"]+)"', open(path).read(), re.M):
        _deps(os.path.join(os.path.dirname(path), m.group(1)), seen)
    return seen


def build(force: bool = False, verbose: bool = False) -> str:
    """One object per translation unit (compiled in parallel, only when the unit or a header it includes changed),
    then one link."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC"]
    if verbose:
        flags.append("-Xptxas=-v")
    flags += os.environ.get("TR_NVCC_FLAGS", "").split()
    nccl = nccl_paths()          # header only: the functions are bound with dlopen at run time (csrc/multi.cu)
    if nccl is not None:
        flags += ["-I", nccl[0]]
    os.makedirs(OBJDIR, exist_ok=True)
    stamp = os.path.join(OBJDIR, "flags.txt")
    if not os.path.exists(stamp) or open(stamp).read() != " ".join(flags):
        force = True
    procs, objs = [], []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        objs.append(obj)
        newest = max(os.path.getmtime(f) for f in _deps(path))
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < newest:
            log = open(obj + ".log", "w")
            procs.append((src, log, subprocess.Popen([nvcc] + flags + ["-c", path, "-o", obj], stdout=log, stderr=subprocess.STDOUT)))
    failed = []
    for src, log, p in procs:
        rc = p.wait()
        log.close()
        text = open(log.name).read()
        if verbose or rc != 0:
            sys.stderr.write(text)
        if rc != 0:
            failed.append(src)
    if failed:
        raise RuntimeError(f"nvcc failed for {failed}")
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-ldl"])
    open(stamp, "w").write(" ".join(flags))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
