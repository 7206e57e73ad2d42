"""The C-ABI library loads and exports every symbol include/gsn_b200.h declares
(no compute calls without a GPU); the product has no CPU fallback."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

from tests.conftest import ROOT


def _declared():
    hdr = open(os.path.join(ROOT, 'include', 'gsn_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    return sorted(set(re.findall(r'\b(gsn_[a-z0-9_]+)\s*\(', hdr)))


def test_library_builds_loads_and_exports_header_symbols():
    from gsn_b200 import _lib, build
    build.build()
    L = ctypes.CDLL(_lib.so_path())
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(L, n), f'{n} declared in include/gsn_b200.h but not exported'
    assert set(_lib.EXPORTED_SYMBOLS) == set(names)
    assert _lib.lib().gsn_abi_version() == 1


def test_sm100a_code_object():
    from gsn_b200 import _lib
    out = subprocess.run(['cuobjdump', '-lelf', _lib.so_path()], capture_output=True, text=True).stdout
    assert 'sm_100a' in out


def test_workspace_queries_validate_arguments():
    from gsn_b200 import _lib
    L = _lib.lib()
    nb = ctypes.c_size_t(0)
    assert L.gsn_graph_workspace_bytes(1000, 4000, 1, ctypes.byref(nb)) == 0 and nb.value > 1000 * 8
    assert L.gsn_graph_workspace_bytes(-1, 0, 1, ctypes.byref(nb)) == -1
    assert L.gsn_graph_workspace_bytes(10, 10, 0, ctypes.byref(nb)) == -1
    assert L.gsn_csr_workspace_bytes(10, 10, ctypes.byref(nb)) == 0
    assert L.gsn_graph_workspace_bytes(1 << 31, 10, 1, ctypes.byref(nb)) == -2


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU behaviour')
def test_no_cpu_fallback():
    """the product path refuses CPU tensors instead of computing on the host"""
    from gsn_b200 import ops
    with pytest.raises(RuntimeError):
        ops.EdgePlan(torch.zeros((2, 3), dtype=torch.int64), 4)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'gsn_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+\.*oracle', txt, flags=re.M), f
