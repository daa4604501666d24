"""TEST INFRASTRUCTURE: builds tests/emu/_build/libvqa_emu.so -- the product's CUDA-core kernel headers compiled
for the HOST against tests/emu/cuda_emu.h (fiber emulation of thread blocks), so that kernel logic can be
checked bit for bit against the oracle on a machine without a GPU.

The product sources are not modified; they are copied into _build/gen/ through a mechanical transform:
  * ``extern __shared__ ... name[];``          -> pointer to the emulator's dynamic shared memory
  * ``kernel<<<grid, block, smem, stream>>>(args);`` -> ``emu::launch(grid, block, [&]{ kernel(args); });``
  * the inline-PTX helpers of common.cuh (griddepcontrol x2, ld.global.nc.v4) and scan.cuh (ld.acquire /
    st.release of the peer-memory flags) -> plain C++
Every transform asserts that it matched, so a change in the sources cannot silently bypass it.
"""
from __future__ import annotations

import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "vietnamese_qa_system_b200", "csrc")
OUT = os.path.join(HERE, "_build")
GEN = os.path.join(OUT, "gen")
LIB = os.path.join(OUT, "libvqa_emu.so")
SOURCES = ["consts.h", "common.cuh", "launch.h", "sparse.cuh", "sparse_launch.cu", "pool.cuh", "scan.cuh",
           "scan_launch.cuh", "scan_f32.cu", "scan_bf16.cu", "scan_f16.cu", "misc_launch.cu", "mma.cuh", "mma_launch.cu", "ts.cuh",
           "ts_launch.cu"]


def _split_top_level(s: str):
    parts, depth, cur = [], 0, []
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    parts.append("".join(cur).strip())
    return parts


def _launches(src: str) -> str:
    pat = re.compile(r"(\b[\w:]+)\s*<<<(.*?)>>>\s*\((.*?)\)\s*;", re.S)

    def repl(m):
        cfg = _split_top_level(" ".join(m.group(2).split()))
        assert len(cfg) in (2, 3, 4), cfg
        smem = cfg[2] if len(cfg) > 2 else "0"
        return f"emu::launch(dim3({cfg[0]}), (unsigned)({cfg[1]}), (size_t)({smem}), [&]() {{ {m.group(1)}({m.group(3)}); }});"

    return pat.sub(repl, src)


def transform(name: str, src: str) -> str:
    n_ext = len(re.findall(r"extern\s+__shared__", src))
    src, n = re.subn(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?unsigned char\s+(\w+)\[\];",
                     r"unsigned char *\1 = emu::cur_smem();", src)
    assert n == n_ext, f"{name}: {n_ext} extern __shared__ declarations, {n} transformed"
    if name == "common.cuh":
        src, a = re.subn(r'asm volatile\("griddepcontrol\.wait;" ::: "memory"\);', ";", src)
        src, b = re.subn(r'asm volatile\("griddepcontrol\.launch_dependents;" ::: "memory"\);', ";", src)
        src, c = re.subn(r'uint4 r;\s*asm volatile\("ld\.global\.nc\.L1::no_allocate\.v4\.u32.*?: "l"\(p\)\);\s*return r;',
                         "return *static_cast<const uint4 *>(p);", src, flags=re.S)
        assert (a, b, c) == (1, 1, 1), (a, b, c)
    if name == "scan.cuh":
        src, a = re.subn(r'unsigned long long v;\s*asm volatile\("ld\.acquire\.sys\.global\.u64.*?: "memory"\);\s*return v;',
                         "return *static_cast<const volatile unsigned long long *>(p);", src, flags=re.S)
        src, b = re.subn(r'asm volatile\("st\.release\.sys\.global\.u64.*?: "memory"\);',
                         "*static_cast<volatile unsigned long long *>(p) = v;", src, flags=re.S)
        assert (a, b) == (1, 1), (a, b)
    if name == "mma.cuh":
        src, a = re.subn(r'unsigned long long v;\s*asm volatile\("ld\.volatile\.global\.u64 %0, \[%1\];" : "=l"\(v\) : "l"\(p\)\);\s*return v;',
                         "return *static_cast<const volatile unsigned long long *>(p);", src)
        src, b = re.subn(r'asm volatile\("bar\.sync %0, %1;" ::"r"\(id\), "r"\(count\) : "memory"\);',
                         "emu::named_bar_sync(id, count);", src)
        src, c = re.subn(r'asm volatile\("bar\.arrive %0, %1;" ::"r"\(id\), "r"\(count\) : "memory"\);',
                         "emu::named_bar_arrive(id, count);", src)
        assert (a, b, c) == (1, 1, 1), (a, b, c)
        # count list updates of the register-list epilogue (lane 0 only: one event per warp-wide call)
        src, d = re.subn(r"(static __device__ __noinline__ Entry reglist_(?:insert_one|merge32)\([^)]*\) \{\n)",
                         r"\1    if ((threadIdx.x & 31) == 0) ++emu_event_counter();\n", src)
        assert d == 2, d
        src, d = re.subn(r"(uint32_t \(&ci\)\[NB\], bool cand_sorted\) \{\n)",
                         r"\1    if ((threadIdx.x & 31) == 0) emu_event_counter() += NB;\n", src)
        assert d == 1, d
    if name == "ts.cuh":
        src, a = re.subn(r'asm volatile\(\s*"tcgen05\.st\.sync\.aligned\.32x32b\.x16\.b32.*?: "memory"\);',
                         "ptx::model_tmem_st16(taddr, r);", src, flags=re.S)
        src, b = re.subn(r'asm volatile\("tcgen05\.wait::st\.sync\.aligned;" ::: "memory"\);', ";", src)
        src, c = re.subn(r'asm volatile\(\s*"\{\\n\\t"\s*"\.reg \.pred p;\\n\\t"\s*"setp\.ne\.b32 p, %4, 0;\\n\\t"\s*'
                         r'"tcgen05\.mma\.cta_group::1\.kind::f16 \[%0\], \[%1\], %2, %3, p;\\n\\t".*?: "memory"\);',
                         "ptx::model_umma_f16_ts(tmem_d, tmem_a, desc_b, idesc, accumulate);", src, flags=re.S)
        assert (a, b, c) == (1, 1, 1), (a, b, c)
    assert "asm volatile" not in src, f"{name}: untransformed inline PTX"
    if "<<<" in src:
        src = _launches(src)
        assert "<<<" not in src, name
    if name.endswith(".cu"):
        src = src.replace('#include "../../include/vqa.h"', f'#include "{os.path.join(ROOT, "include", "vqa.h")}"')
    return src


def _cuda_include() -> str:
    return os.environ.get("CUDA_INCLUDE", "/usr/local/cuda/include")


def toolchain_available() -> bool:
    import shutil

    return shutil.which("g++") is not None and os.path.exists(os.path.join(_cuda_include(), "cuda_runtime.h"))


def _install_ptx_model() -> None:
    """gen/ptx.cuh = tests/emu/ptx_emu.cuh, after checking that it models every wrapper of the real ptx.cuh."""
    with open(os.path.join(CSRC, "ptx.cuh"), encoding="utf-8") as f:
        real = f.read()
    with open(os.path.join(HERE, "ptx_emu.cuh"), encoding="utf-8") as f:
        model = f.read()
    pat = r"^(?:__host__ )?__device__ (?:__forceinline__ )?(?:constexpr )?[\w:]+ (\w+)\("
    real_fns = set(re.findall(pat, real, flags=re.M))
    model_fns = set(re.findall(r"^(?:inline|constexpr) [\w:]+ (\w+)\(", model, flags=re.M))
    missing = sorted(real_fns - model_fns)
    assert len(real_fns) >= 25 and not missing, f"ptx.cuh wrappers without a host model: {missing}"
    with open(os.path.join(GEN, "ptx.cuh"), "w", encoding="utf-8") as f:
        f.write(model)


def build(force: bool = False) -> str:
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(HERE, f) for f in ("cuda_emu.h", "ptx_emu.cuh", "emu_kernels.cpp", "build.py")] + \
        [os.path.join(CSRC, "ptx.cuh")]
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(d) for d in deps):
        return LIB
    os.makedirs(GEN, exist_ok=True)
    _install_ptx_model()
    for s in SOURCES:
        with open(os.path.join(CSRC, s), encoding="utf-8") as f:
            out = transform(s, f.read())
        with open(os.path.join(GEN, s), "w", encoding="utf-8") as f:
            f.write(out)
    cuda_inc = _cuda_include()
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
           "-Wno-attributes", "-Wno-unused-value", f"-I{cuda_inc}", f"-I{HERE}", f"-I{GEN}",
           os.path.join(HERE, "emu_kernels.cpp"), "-o", LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emulator build failed:\n" + (r.stdout + r.stderr)[-6000:])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
