"""Builds libmgn_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the
repo snapshot to the GPU box).  Run as ``python meshgraphnets.jl_b200/build.py`` or through
``__graft_entry__.build()``."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(CSRC, "libmgn_b200.so")
SOURCES = ["abi.cu", "device_state.cu", "device_graph.cu", "features.cu", "csr.cu", "simt_kernels.cu", "solver_kernels.cu", "pipeline.cu", "dp.cu", "tc_kernels.cu",
           "tc_bwd_kernels.cu", "tc_pipeline.cu"]
# test-only probe of the tcgen05 building blocks: its own library, never linked into the product
PROBE_LIB = os.path.join(CSRC, "libmgn_b200_probe.so")
PROBE_SOURCES = ["tc_probe.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src, verbose, extra=()):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "mgn_b200.h"))
    stamp = obj + ".sha"
    dig = _digest([os.path.join(CSRC, src)] + headers) + " ".join(extra)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, False
    cmd = [NVCC] + FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        sys.stderr.write(r.stderr)
    with open(stamp, "w") as f:
        f.write(dig)
    return obj, True


def _link(lib, objs):
    cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
           "-Xcompiler", "-fPIC", "-o", lib] + objs + ["-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")


def build(verbose=False, force=False, trace=False):
    """Compiles every CUDA source for sm_100a and links libmgn_b200.so (+ the test-only probe library).
    ``trace=True`` is the debug build with the per-role timestamp hooks (tools/trace_kernels.py)."""
    os.makedirs(OBJ, exist_ok=True)
    srcs = [s for s in SOURCES + PROBE_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    extra = ("-DMGN_ENABLE_TRACE",) if trace else ()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = dict(zip(srcs, ex.map(lambda s: _compile(s, verbose, extra), srcs)))
    main = [results[s] for s in SOURCES if s in results]
    if any(changed for _, changed in main) or not os.path.exists(LIB):
        _link(LIB, [o for o, _ in main])
    probe = [results[s] for s in PROBE_SOURCES if s in results]
    if probe and (any(changed for _, changed in probe) or not os.path.exists(PROBE_LIB)):
        _link(PROBE_LIB, [o for o, _ in probe])
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv, trace="--trace" in sys.argv))
