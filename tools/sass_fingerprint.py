"""Fingerprint (sha256 of the instruction stream, addresses and encodings stripped) of selected kernels in the
built objects -- used to prove that a source change left a GPU-measured kernel instruction-for-instruction
unchanged.  usage: python tools/sass_fingerprint.py [--write]   (CPU only: cuobjdump on the .o files)"""
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "vietnamese_qa_system_b200", "build")
# fingerprints of the build whose measurements are quoted in DESIGN.md / profiles/ (rewritten -- `--write` -- each time a
# GPU session measures a new build; round 1's are kept as profiles/r1_measured_kernel_sass.json)
GOLD = os.path.join(ROOT, "profiles", "r2_measured_kernel_sass.json")
# kernels whose measurements stand: (object, substring of the mangled name)
WATCH = [("ts_launch.o", "ts_topk_kernel"), ("mma_launch.o", "mma_topk_kernel")]


def nvcc_version() -> str:
    out = subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout
    m = re.search(r"release [\d.]+, V([\d.]+)", out)
    return m.group(1) if m else "?"


def fingerprints() -> dict:
    fp = {}
    for obj, pat in WATCH:
        sass = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True).stdout
        cur = None
        acc = {}
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = m.group(1) if pat in m.group(1) else None
                if cur:
                    acc[cur] = []
                continue
            if cur and re.search(r"/\*[0-9a-f]{4,6}\*/", line):
                ins = re.sub(r"/\*.*?\*/", "", line).strip()
                if ins:
                    acc[cur].append(ins)
        for name, ins in acc.items():
            fp[name] = {"instructions": len(ins), "sha256": hashlib.sha256("\n".join(ins).encode()).hexdigest()}
    return fp


if __name__ == "__main__":
    cur = {"nvcc": nvcc_version(), "kernels": fingerprints()}
    if "--write" in sys.argv:
        with open(GOLD, "w") as f:
            json.dump(cur, f, indent=1, sort_keys=True)
        print("wrote", GOLD, len(cur["kernels"]), "kernels")
    else:
        print(json.dumps(cur, indent=1, sort_keys=True))
