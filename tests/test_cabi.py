"""The C-ABI boundary: header <-> ctypes table <-> exported symbols (no compute, CPU only)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "accelrl_b200.h")


def _header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(arl_[a-z0-9_]+)\s*\(", src)))


def _built_lib():
    from accel_rl_b200 import _lib as L
    if not os.path.exists(L.LIB_PATH):
        L.build()
    return L


def test_header_declares_what_ctypes_binds():
    from accel_rl_b200 import _lib as L
    assert _header_symbols() == sorted(L.SIGNATURES.keys())


def test_library_exports_every_declared_symbol():
    L = _built_lib()
    out = subprocess.run(["nm", "-D", "--defined-only", L.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (arl_[a-z0-9_]+)", out))
    missing = [s for s in _header_symbols() if s not in exported]
    assert not missing, missing


def test_library_loads_and_types_entry_points():
    L = _built_lib()
    lib = L.load()
    for name in L.SIGNATURES:
        assert getattr(lib, name) is not None


def test_header_cites_reference_for_each_group():
    src = open(HEADER).read()
    for ref in ("atari_cnn_policy.py", "envs/atari_env.py:151-157", "overlap/sampler.py:97-151", "aac_base.py:108-145",
                "optimizers/single", "optimizers/sync/base.py:8-24", "rllab/misc/special.py:22-27"):
        assert ref in src, ref


def test_no_torch_types_in_the_abi():
    src = open(HEADER).read()
    assert "at::" not in src and "Tensor" not in src


def test_library_is_sm100a_tcgen05():
    L = _built_lib()
    r = subprocess.run(["cuobjdump", "-sass", L.LIB_PATH], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in r.stdout
    assert "UTCHMMA" in r.stdout and "LDTM" in r.stdout     # tcgen05.mma / tcgen05.ld


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "accel_rl_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), os.path.join(dp, f)


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from accel_rl_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine([32, 64, 64], [8, 4, 3], [4, 2, 1], [0, 1, 1], [512], 4, (4, 104, 80))


def test_loaded_library_matches_the_sources_in_the_tree():
    """the prebuilt .so travels to the GPU box next to its sources: the sha256 compiled into it (csrc/Makefile) must be
    the sha256 of those sources"""
    from accel_rl_b200 import _lib as L
    assert L.check_source_hash() == L.source_hash() and len(L.source_hash()) == 64
