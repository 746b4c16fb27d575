"""The jax.ffi face of the boundary (ffi/larnd_ffi.cc + ffi/sim_b200.py).

jax is absent from the build image and from the GPU box (profiles/r2_probe_jax.txt), so the shim cannot run there; what CAN
be checked without jax is checked here on the CPU: the handlers are well-formed C++ against the XLA FFI API shape (a mock
header, ffi/mock/) with every handler signature matching its binding, the Python side names exactly the custom-call targets
the C++ side defines, and the parameter-block fields it patches exist in the ctypes mirror of include/larnd_b200.h.  Where
`import jax` works the real library is built by __graft_entry__.build() and a forward + gradient round trip runs on the GPU.
"""
import ast
import importlib.util
import os
import re
import subprocess

import pytest

import common as cm

FFI = os.path.join(cm.ROOT, "ffi")
HAVE_JAX = importlib.util.find_spec("jax") is not None


def test_handlers_compile_against_the_ffi_api_shape():
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-I" + os.path.join(FFI, "mock"), "-I" + os.path.join(cm.ROOT, "include"),
           "-I/usr/local/cuda/include", os.path.join(FFI, "larnd_ffi.cc")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_python_side_and_cpp_side_name_the_same_targets_and_fields():
    src_cc = open(os.path.join(FFI, "larnd_ffi.cc")).read()
    src_py = open(os.path.join(FFI, "sim_b200.py")).read()
    ast.parse(src_py)
    defined = set(re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\(larnd_ffi_(\w+),", src_cc))
    targets = set(ast.literal_eval(re.search(r"_TARGETS = (\([^)]*\))", src_py).group(1)))
    called = set(re.findall(r'ffi_call\(\s*"larnd_(\w+)"', src_py))
    assert defined == targets and called <= targets and len(defined) == 7
    # every C-ABI entry point the handlers call is declared in the public header
    header = open(os.path.join(cm.ROOT, "include", "larnd_b200.h")).read()
    for sym in set(re.findall(r"\b(larnd_(?:lut|fee|mc)_\w+)\(", src_cc)):
        assert re.search(r"\b%s\(" % sym, header), sym
    # the parameter-block fields patched from traced leaves exist in the ctypes mirror, 4-byte aligned
    from larndsim_b200 import _lib
    dep = ast.literal_eval(re.search(r"_DEPENDENT = (\([^)]*\))", src_py, re.S).group(1))
    names = {f[0] for f in _lib.ParamsPOD._fields_}
    assert set(dep) <= names and all(getattr(_lib.ParamsPOD, d).offset % 4 == 0 for d in dep)
    # ... and are exactly the fields sim.fill_pod_leaves sets on the torch side
    fill = open(os.path.join(cm.ROOT, "larnd-sim-jax_b200", "larndsim_b200", "sim.py")).read()
    body = fill[fill.index("def fill_pod_leaves"):fill.index("def make_pod")]
    assert set(re.findall(r"P\.(\w+)", body)) == set(dep)


@pytest.mark.gpu
@pytest.mark.skipif(not HAVE_JAX, reason="jax is not installed (neither in the build image nor on the GPU box: profiles/r2_probe_jax.txt)")
def test_shim_round_trip_matches_the_torch_binding(torch_dev):
    """simulate_wfs + simulate_stochastic through jax.ffi against the ctypes/torch binding of the same library."""
    import sys
    import numpy as np
    sys.path.insert(0, FFI)
    import jax.numpy as jnp
    import sim_b200
    import torch
    from larndsim_b200 import sim
    pp = cm.product_params(number_pix_neighbors=2, signal_length=100)
    bank = cm.synthetic_bank(32, 25, 25, 1950)
    tr = cm.small_batch(600, pad=8, precision=0.01)
    st = sim.lut_forward(pp, torch.as_tensor(bank, device=torch_dev), torch.as_tensor(tr, device=torch_dev), cm.FIELDS)
    wfs, upix = sim_b200.simulate_wfs(_RefLike(pp), jnp.asarray(bank), jnp.asarray(tr), cm.FIELDS)
    assert np.array_equal(np.asarray(upix), st.unique_pixels.cpu().numpy())
    assert np.allclose(np.asarray(wfs), st.wfs_full[:, 1:].cpu().numpy(), rtol=1e-5, atol=1e-3)


class _RefLike:
    """Attribute view of a larndsim_b200 Params object with the reference's field names (they are the same)."""

    def __init__(self, p):
        self._p = p

    def __getattr__(self, name):
        return getattr(self._p, name)

    def replace(self, **kw):
        return _RefLike(self._p.replace(**kw))
