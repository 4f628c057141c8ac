"""tcgen05 building blocks (descriptors, TMEM, commit) against torch on small GEMMs."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _probe(pkg, a, b, K, mode):
    import os
    lib = C.CDLL(os.path.join(os.path.dirname(pkg.LIB_PATH), "libmgn_b200_probe.so"))   # test-only library
    fn = lib.mgn_debug_umma_probe
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    fn.restype = C.c_int32
    out = torch.zeros(128, 128, device="cuda")
    st = fn(a.data_ptr(), b.data_ptr(), out.data_ptr(), K, mode, torch.cuda.current_stream().cuda_stream)
    assert st == 0
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("K", [64, 128, 256])
def test_umma_kmajor(pkg, K):
    g = torch.Generator(device="cuda").manual_seed(K)
    a = torch.randn(128, K, device="cuda", generator=g).bfloat16()
    b = torch.randn(128, K, device="cuda", generator=g).bfloat16()
    out = _probe(pkg, a, b, K, 0)
    ref = a.float() @ b.float().T
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-3), float((out - ref).abs().max())


@pytest.mark.parametrize("K", [16, 128, 208, 256])
def test_umma_mnmajor(pkg, K):
    g = torch.Generator(device="cuda").manual_seed(K + 1)
    a = torch.randn(K, 128, device="cuda", generator=g).bfloat16()
    b = torch.randn(K, 128, device="cuda", generator=g).bfloat16()
    out = _probe(pkg, a, b, K, 1)
    ref = a.float().T @ b.float()
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-3), float((out - ref).abs().max())
