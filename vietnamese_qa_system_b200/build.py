"""Build libvqa_b200.so in-tree with nvcc for sm_100a (no GPU needed to compile)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libvqa_b200.so")
SOURCES = ["api.cu", "misc_launch.cu", "scan_f32.cu", "scan_bf16.cu", "scan_f16.cu", "mma_launch.cu", "ts_launch.cu", "pair_launch.cu", "sparse_launch.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libvqa_b200.so cannot be built")


STAMP = os.path.join(BUILD, "sources.sha256")


def _sources_digest() -> str:
    """Content hash of everything the library is compiled from.  A content stamp instead of mtimes: the
    snapshot that carries the built .so to the GPU box does not preserve them, and N ranks importing the
    package there must all find the binary current instead of racing to rebuild it."""
    import hashlib

    h = hashlib.sha256()
    paths = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    paths.append(os.path.join(HERE, "..", "include", "vqa.h"))
    for p in paths:
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def is_current() -> bool:
    try:
        with open(STAMP) as f:
            return os.path.exists(LIB) and f.read().strip() == _sources_digest()
    except OSError:
        return False


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source with -gencode arch=compute_100a,code=sm_100a -lineinfo."""
    if not force and is_current():
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    import fcntl

    with open(os.path.join(BUILD, ".lock"), "w") as lockf:  # one builder at a time (torchrun ranks, xdist workers)
        fcntl.flock(lockf, fcntl.LOCK_EX)
        if not force and is_current():
            return LIB
        return _build_locked(verbose)


def _build_locked(verbose: bool) -> str:
    digest = _sources_digest()
    nvcc = _nvcc()
    flags = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr",
             "-Xptxas", "-v"] + ARCH

    def compile_one(src: str) -> str:
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        cmd = [nvcc, "-c", os.path.join(CSRC, src), "-o", obj] + flags
        r = subprocess.run(cmd, capture_output=True, text=True)
        # per-kernel registers / shared memory / spills of this translation unit (read by spill_report())
        with open(os.path.join(BUILD, src.replace(".cu", ".ptxas.log")), "w") as f:
            f.write(r.stdout + r.stderr)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        return obj

    with ThreadPoolExecutor(max_workers=min(6, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ARCH + ["-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    with open(STAMP, "w") as f:
        f.write(digest + "\n")
    for name, (regs, st, ld) in sorted(spill_report().items()):
        if st or ld:  # never silent: a spill inside a streaming loop costs real bandwidth
            sys.stderr.write(f"[vqa build] register spills: {name}: {st} B stores / {ld} B loads at {regs} registers\n")
    return LIB


def spill_report() -> dict:
    """{mangled kernel name: (registers, spill store bytes, spill load bytes)} from the last build's ptxas logs."""
    import re

    out = {}
    if not os.path.isdir(BUILD):
        return out
    for fn in sorted(os.listdir(BUILD)):
        if not fn.endswith(".ptxas.log"):
            continue
        with open(os.path.join(BUILD, fn)) as f:
            txt = f.read()
        for block in re.split(r"ptxas info\s+: Compiling entry function '", txt)[1:]:
            name = block.split("'")[0]
            regs = re.search(r"Used (\d+) registers", block)
            sp = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", block)
            if regs and sp:
                out[name] = (int(regs.group(1)), int(sp.group(1)), int(sp.group(2)))
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
