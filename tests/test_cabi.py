"""The C-ABI library loads and exports every symbol include/vqa.h declares; argument and
no-device error behaviour.  No compute calls (CPU only)."""
import ctypes
import os
import re

import pytest

from tests.conftest import ROOT
from vietnamese_qa_system_b200 import _native as N


def declared_symbols():
    with open(os.path.join(ROOT, "include", "vqa.h")) as f:
        text = f.read()
    return sorted(set(re.findall(r"VQA_API\s+[\w\s\*]+?\b(vqa_\w+)\s*\(", text)))


def test_header_symbols_all_exported():
    syms = declared_symbols()
    assert len(syms) >= 15 and set(syms) == set(N.EXPORTS)
    L = N.lib()
    for s in syms:
        assert getattr(L, s) is not None, s


def test_version_and_last_error():
    L = N.lib()
    assert L.vqa_version() == 121 == N.ABI_VERSION
    assert isinstance(N.last_error(), str)


def test_tuning_knobs_are_validated_once_not_read_on_the_search_path(monkeypatch):
    """The kernel-selection knobs are a vqa_tuning_t: library defaults, VQA_* overrides parsed ONCE (index create /
    vqa_tuning_from_env) and range-checked -- a nonsense value is VQA_E_INVALID, never a silently shortened list."""
    L = N.lib()
    for v in ("VQA_TS_EXTRA", "VQA_TS_QS", "VQA_REDUCE_SELECT", "VQA_TMA_L2PROMO", "VQA_TS_KS"):
        monkeypatch.delenv(v, raising=False)
    t = N.tuning_default()
    assert t.size == ctypes.sizeof(N.Tuning) and t.ts_extra == 6 and t.ts_qs == 1 and t.reduce_select == 1
    assert t.ts_ks == -1 and t.stream_max_b == 0 and t.tma_l2promo == 3 and t.reduce_early == 1 and t.ts_m64 == 1
    assert N.tuning_default(from_env=True).as_dict() == t.as_dict()
    monkeypatch.setenv("VQA_TS_QS", "0")
    monkeypatch.setenv("VQA_TS_EXTRA", "12")
    e = N.tuning_default(from_env=True)
    assert (e.ts_qs, e.ts_extra) == (0, 12)
    for name, bad in (("VQA_TS_EXTRA", "-3"), ("VQA_TS_EXTRA", "97"), ("VQA_TMA_L2PROMO", "7"), ("VQA_TS_QS", "yes"),
                      ("VQA_TS_KS", "4x")):
        monkeypatch.setenv(name, bad)
        rc = L.vqa_tuning_from_env(ctypes.byref(N.Tuning()))
        assert rc == N.E_INVALID and name in N.last_error(), (name, bad, N.last_error())
        monkeypatch.delenv(name)
    # explicit tuning through the device-free planner: out-of-range fields and a wrong struct size are refused
    out, smem = (ctypes.c_int32 * 16)(), ctypes.c_size_t()
    good = N.tuning_default()
    assert L.vqa_plan_describe_tuned(1000, 768, N.BF16, 64, 10, N.MODE_FAST, 148, 232448, ctypes.byref(good), out,
                                     ctypes.byref(smem)) == 0
    bad = N.tuning_default().update(ts_extra=-1)
    assert L.vqa_plan_describe_tuned(1000, 768, N.BF16, 64, 10, N.MODE_FAST, 148, 232448, ctypes.byref(bad), out,
                                     ctypes.byref(smem)) == N.E_INVALID and "ts_extra" in N.last_error().lower()
    bad = N.tuning_default()
    bad.size = 8
    assert L.vqa_plan_describe_tuned(1000, 768, N.BF16, 64, 10, N.MODE_FAST, 148, 232448, ctypes.byref(bad), out,
                                     ctypes.byref(smem)) == N.E_INVALID
    with pytest.raises(ValueError):
        N.tuning_default().update(no_such_knob=1)
    # the hot path has no getenv: api.cu reads the environment in exactly one function
    with open(os.path.join(ROOT, "vietnamese_qa_system_b200", "csrc", "api.cu")) as f:
        api = f.read()
    assert api.count("getenv(") == 1
    for fn in os.listdir(os.path.join(ROOT, "vietnamese_qa_system_b200", "csrc")):
        if fn != "api.cu":
            with open(os.path.join(ROOT, "vietnamese_qa_system_b200", "csrc", fn)) as f:
                assert "getenv" not in f.read(), fn


def test_library_is_built_for_sm100a():
    import subprocess

    out = subprocess.run(["cuobjdump", "-lelf", N.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def has_gpu():
    return N.device_count() > 0


@pytest.mark.skipif(has_gpu(), reason="checks the no-device error path")
def test_no_device_fails_loudly_not_silently():
    L = N.lib()
    h = ctypes.c_void_p()
    rc = L.vqa_index_create(ctypes.byref(h), 10, 768, N.BF16, 0, 0)
    assert rc == N.E_CUDA and "no CPU fallback" in N.last_error()
    with pytest.raises(N.VqaError):
        N.check(rc)
    with pytest.raises(N.VqaError):
        N.require_cuda()
    # compute entry points also refuse (nothing is computed on the host)
    buf = (ctypes.c_float * 16)()
    ids = (ctypes.c_int64 * 16)()
    rc = L.vqa_merge_topk(buf, ids, 1, 1, 4, 4, buf, ids, 0, None)
    assert rc == N.E_CUDA
    rc = L.vqa_normalize_rows(buf, 4, 4, 4, buf, 4, None, 0, 0, 0, None)
    assert rc == N.E_CUDA
    # the sparse (BM25) leg and the fusion refuse the same way
    hs = ctypes.c_void_p()
    assert L.vqa_sparse_create(ctypes.byref(hs), 10, 2, 3, 0) == N.E_CUDA and "no CPU fallback" in N.last_error()
    dbl = (ctypes.c_double * 16)()
    assert L.vqa_hybrid_fuse(buf, ids, 4, dbl, ids, 4, 1, 0.5, 0.5, 2, dbl, ids, 0, None) == N.E_CUDA
    assert L.vqa_bm25_weights(ids, 1, ids, ids, 1, dbl, ids, 1.2, 0.75, 2.0, buf, 0, None) == N.E_CUDA
    assert L.vqa_agree_f64(ids, dbl, ids, dbl, 4, 0.4, None, None, 0, None) == N.E_CUDA
    from vietnamese_qa_system_b200.scoring import BM25

    bm = BM25({"terms": True})
    bm.build_postings([["a", "b"], ["a"]])       # host half works anywhere ...
    with pytest.raises(N.VqaError):
        bm.index([["a", "b"], ["a"]])            # ... scoring needs the device


def test_argument_validation_precedes_device_use():
    L = N.lib()
    h = ctypes.c_void_p()
    assert L.vqa_index_create(ctypes.byref(h), 10, 768, 9, 0, 0) == N.E_INVALID       # bad dtype
    assert L.vqa_index_create(ctypes.byref(h), 10, 7, N.BF16, 0, 0) == N.E_INVALID     # 14-byte rows
    assert "multiple of 16" in N.last_error()
    assert L.vqa_index_create(ctypes.byref(h), -1, 768, N.F32, 0, 0) == N.E_INVALID
    assert L.vqa_index_create(None, 1, 768, N.F32, 0, 0) == N.E_INVALID
    assert L.vqa_search(None, None, 0, 1, 1, 0, None, None, None, 0, None) == N.E_INVALID
    assert L.vqa_merge_topk(None, None, 1, 1, 1, 1, None, None, 0, None) == N.E_INVALID
    buf = (ctypes.c_float * 16)()
    ids = (ctypes.c_int64 * 16)()
    assert L.vqa_merge_topk(buf, ids, 1, 1, 4, 4096, buf, ids, 0, None) == N.E_INVALID  # k_out > 128
    assert L.vqa_pool_normalize(buf, 7, buf, N.I64, 1, 1, 8, 1, buf, 0, None) == N.E_INVALID
    hs = ctypes.c_void_p()
    assert L.vqa_sparse_create(ctypes.byref(hs), 2 ** 31, 1, 1, 0) == N.E_INVALID      # > 2^31 - 1 documents
    assert L.vqa_sparse_create(None, 1, 1, 1, 0) == N.E_INVALID
    assert L.vqa_sparse_search(None, None, None, None, 1, 1, 1, 1, 0, 0.0, None, None, None, 0, None) == N.E_INVALID
    assert L.vqa_hybrid_fuse(None, None, 1, None, None, 1, 1, 0.5, 0.5, 1, None, None, 0, None) == N.E_INVALID
    mt, mc = ctypes.c_int32(), ctypes.c_int32()
    assert L.vqa_sparse_limits(ctypes.byref(mt), ctypes.byref(mc)) == N.OK and (mt.value, mc.value) == (64, 1024)
    with pytest.raises(ValueError):
        N.check(N.E_INVALID)
    with pytest.raises(NotImplementedError):
        N.check(N.E_UNSUPPORTED)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "vietnamese_qa_system_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dp, fn), encoding="utf-8") as f:
                    src = f.read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, re.M), fn
                assert "liboracle" not in src, fn


def test_ctypes_bindings_match_header_prototypes():
    """Every prototype in include/vqa.h against the ctypes binding: same parameter count, pointer parameters
    bound as pointers, double as c_double, integer widths as declared (catches binding drift)."""
    with open(os.path.join(ROOT, "include", "vqa.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    protos = re.findall(r"VQA_API\s+[\w\s\*]+?\b(vqa_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S)
    assert len(protos) == len(N.EXPORTS)
    L = N.lib()
    for name, params in protos:
        params = " ".join(params.split())
        plist = [] if params in ("void", "") else [p.strip() for p in params.split(",")]
        fn = getattr(L, name)
        assert fn.argtypes is not None and len(fn.argtypes) == len(plist), (name, plist, fn.argtypes)
        for decl, ct in zip(plist, fn.argtypes):
            if "*" in decl:
                assert ct is ctypes.c_void_p or hasattr(ct, "contents") or ct is ctypes.c_char_p, (name, decl, ct)
            elif decl.startswith("double"):
                assert ct is ctypes.c_double, (name, decl)
            elif decl.startswith("int64_t"):
                assert ct is ctypes.c_int64, (name, decl)
            elif decl.startswith("int32_t"):
                assert ct is ctypes.c_int32, (name, decl)
            elif decl.startswith("uint64_t"):
                assert ct is ctypes.c_uint64, (name, decl)
            elif decl.startswith("size_t"):
                assert ct is ctypes.c_size_t, (name, decl)
            else:
                raise AssertionError(f"unhandled parameter type in {name}: {decl}")


def test_no_kernel_spills_registers_in_a_hot_loop():
    """ptxas' own accounting of the last build (build/*.ptxas.log): no kernel may spill more than a few bytes.
    (A fixed 4-token unroll once made the 768-dim bf16 pooling kernel spill ~400 bytes per thread inside its
    streaming loop without anyone noticing, because the warning was not surfaced.)"""
    from vietnamese_qa_system_b200 import build as vbuild

    rep = vbuild.spill_report()
    if not rep:
        pytest.skip("no ptxas logs next to the library (prebuilt .so)")
    assert len(rep) >= 80                                  # every kernel instantiation is accounted for
    worst = {k: v for k, v in rep.items() if v[1] > 32 or v[2] > 32}
    assert not worst, worst


def test_product_never_references_the_emulator():
    pkg = os.path.join(ROOT, "vietnamese_qa_system_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dp, fn), encoding="utf-8") as f:
                    src = f.read()
                assert "cuda_emu" not in src and "tests/emu" not in src and "tests.emu" not in src, fn


def test_measured_kernels_are_unchanged():
    """The tensor-core kernels are instruction-for-instruction the ones of the build that was last measured on the
    B200 (profiles/r2_measured_kernel_sass.json, rewritten by `tools/sass_fingerprint.py --write` when a GPU session
    measures a new build): a source edit after the last measurement that perturbs them fails here, so the numbers in
    DESIGN.md always belong to the committed code.  Compares SASS instruction streams of the current build; skipped
    when the toolchain differs from the one that produced the fingerprints."""
    import json
    import shutil
    import sys

    if shutil.which("cuobjdump") is None or shutil.which("nvcc") is None:
        pytest.skip("cuobjdump / nvcc not available")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_fingerprint as sf

    N.lib()  # make sure the objects exist
    with open(sf.GOLD) as f:
        gold = json.load(f)
    if sf.nvcc_version() != gold["nvcc"]:
        pytest.skip(f"nvcc {sf.nvcc_version()} != {gold['nvcc']} (fingerprints are compiler specific)")
    cur = sf.fingerprints()
    assert len(gold["kernels"]) >= 10
    # matched by instruction stream, not by name: opt-in variants are added as extra (defaulted) template
    # parameters, which changes the mangled names of the measured instantiations but must not change their code
    have = {v["sha256"]: k for k, v in cur.items()}
    for name, want in gold["kernels"].items():
        assert want["sha256"] in have, f"{name}: no kernel in the current build has the measured instruction stream"


def test_documented_environment_knobs_exist_in_the_library():
    """INTEGRATION.md section 6 lists the VQA_* knobs; each must be a string the built library actually reads
    (and every knob the library reads must be documented)."""
    import subprocess

    with open(os.path.join(ROOT, "INTEGRATION.md"), encoding="utf-8") as f:
        doc = f.read().split("## 6. Tuning knobs", 1)[1]
    documented = set(re.findall(r"`(VQA_[A-Z0-9_]+)", doc)) - {"VQA_EXPERIMENTAL", "VQA_E_INVALID"}
    N.lib()
    so = os.path.join(ROOT, "vietnamese_qa_system_b200", "libvqa_b200.so")
    strings = subprocess.run(["strings", "-n", "6", so], capture_output=True, text=True).stdout
    in_lib = set(re.findall(r"^(VQA_[A-Z0-9_]+)$", strings, flags=re.M)) - {"VQA_SPIN_LIMIT"}
    assert documented <= in_lib, f"documented but not read by the library: {sorted(documented - in_lib)}"
    assert in_lib <= documented, f"read by the library but not documented: {sorted(in_lib - documented)}"
