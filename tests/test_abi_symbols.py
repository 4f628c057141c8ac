"""The C-ABI library loads on a CPU-only machine and exports every symbol include/mgn_b200.h
declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mgn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mgn_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(pkg):
    lib = ctypes.CDLL(pkg.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mgn_b200.h but not exported"


def test_binding_covers_header(pkg):
    import mgn_pkg  # noqa: F401
    from meshgraphnets_jl_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()


def test_error_convention(pkg):
    with pytest.raises(pkg.MgnError) as e:
        pkg.Model(9, 3, 2, 15, 4096, 2)  # latent out of range
    assert e.value.code == 1 and "latent" in str(e.value)
    m = pkg.Model(9, 3, 2, 15, 128, 2)
    assert m.n_params == 2877570  # SURVEY.md 8 a15 (L = 4)
    names = [n for n, *_ in m.param_layout()]
    assert names[0] == "encoder.node.dense1.weight" and names[-1] == "decoder.dense4.bias"
    assert "processor15.node.layernorm.scale" in names


def test_param_layout_matches_oracle(pkg):
    import mgn_oracle as orc
    cfg = orc.ModelConfig(9, 3, 2, 128, 15, 2)
    specs, P = orc.mlp_specs(cfg)
    m = pkg.Model(9, 3, 2, 15, 128, 2)
    assert P == m.n_params
    lay = {n: (o, r, c) for n, o, r, c in m.param_layout()}
    for s in specs:
        for l, (w, b, i, o) in enumerate(s.dense):
            assert lay[f"{s.name}.dense{l + 1}.weight"] == (w, o, i)
            assert lay[f"{s.name}.dense{l + 1}.bias"] == (b, o, 1)
        if s.ln:
            assert lay[f"{s.name}.layernorm.bias"][0] == s.ln[0]
            assert lay[f"{s.name}.layernorm.scale"][0] == s.ln[1]


def test_no_device_is_an_error_not_a_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    n = ctypes.c_int32(-1)
    assert pkg.load().mgn_device_count(ctypes.byref(n)) == 0 and n.value == 0
    h = ctypes.c_void_p()
    st = pkg.load().mgn_graph_create(4, 0, None, None, 1, None, ctypes.byref(h))
    assert st == 2  # MGN_ERR_CUDA: fails loudly, nothing is computed on the CPU


def test_product_path_fails_loudly_without_the_library(pkg, monkeypatch, tmp_path):
    """No CPU / PyTorch fallback: with the shared library missing, loading (and therefore every
    product call) raises MgnError instead of silently computing somewhere else."""
    from meshgraphnets_jl_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libmgn_b200.so"))
    with pytest.raises(_lib.MgnError):
        _lib.load()
    with pytest.raises(_lib.MgnError):
        pkg.one_hot(np.zeros(3, np.int32), 7, 1)


def test_device_entry_points_fail_without_a_gpu(pkg):
    """On a box without a CUDA device the device entry points return MGN_ERR_CUDA (2) - never a fallback."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import ctypes as C
    lib = pkg.load()
    n = C.c_int32(-1)
    assert lib.mgn_device_count(C.byref(n)) == 0 and n.value == 0
    from meshgraphnets_jl_b200._lib import ModelConfig
    h = C.c_void_p()
    st = lib.mgn_model_create(C.byref(ModelConfig(9, 3, 2, 128, 2, 2, 1e-5, 1)), C.byref(h))
    assert st == 2     # the bf16 model uploads its weight-image plan to the device: MGN_ERR_CUDA


def test_plain_c_consumer_runs_the_known_answer_tests(tmp_path):
    """tests/c/abi_kat.c is compiled as C11 against include/mgn_b200.h (the header must be C, not C++), linked with
    libmgn_b200.so and run: the SURVEY 8c known-answer tests through the host half of the ABI, the parameter table of
    the reference configuration, and - on a machine without a GPU - MGN_ERR_CUDA from the device entry points (no CPU
    fallback)."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "meshgraphnets.jl_b200", "csrc")
    if not os.path.exists(os.path.join(libdir, "libmgn_b200.so")):
        import __graft_entry__
        __graft_entry__.build()
    gcc = shutil.which("gcc")
    assert gcc, "gcc is part of the image"
    exe = str(tmp_path / "abi_kat")
    r = subprocess.run([gcc, "-std=c11", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                        os.path.join(root, "tests", "c", "abi_kat.c"), "-o", exe, "-L", libdir, "-lmgn_b200",
                        "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "all checks passed" in r.stdout, r.stdout + r.stderr
