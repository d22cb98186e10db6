"""The C-ABI boundary: libcaspr_b200.so loads without a GPU and exports exactly what
include/caspr_b200.h declares (no compute calls here)."""
import ctypes
import os
import re
import subprocess

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, 'include', 'caspr_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(caspr_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_exported(lib_built):
    names = _declared()
    assert len(names) >= 15
    lib = ctypes.CDLL(lib_built)
    for n in names:
        assert hasattr(lib, n), 'header declares %s but the library does not export it' % n
    out = subprocess.run(['nm', '-D', '--defined-only', lib_built], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r' T (caspr_[a-z0-9_]+)', out)))
    assert exported == names, 'exports not declared in the header: %s' % (set(exported) ^ set(names))
    # nothing else leaves the library: no C++-mangled internals, no data symbols (csrc/exports.map)
    others = [l for l in out.splitlines() if l.strip() and not re.search(r' T caspr_[a-z0-9_]+$', l)]
    assert others == [], others


def test_ctypes_signatures_cover_header(lib_built):
    from caspr_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    assert _lib.lib.caspr_version() == 100
    assert _lib.lib.caspr_build_arch() == b'sm_100a'
    assert 'underflow in dt' in _lib.status_string(-4)
    assert 'non-finite' in _lib.status_string(-5)


def test_sass_is_sm100a(lib_built):
    out = subprocess.run(['cuobjdump', '--list-elf', lib_built], capture_output=True, text=True).stdout
    assert 'sm_100a' in out


def test_argument_validation_without_gpu(lib_built):
    """EINVAL paths return before any CUDA call, so they are checkable on a GPU-less machine."""
    from caspr_b200._lib import lib
    assert lib.caspr_fps(None, 1, 16, 4, None, None, None) == -1
    assert lib.caspr_linear(None, 0, None, 0, None, None, 0, 0, 0, 0, 0, 0, None) == -1
    assert lib.caspr_cnf_workspace_bytes(0, 0, 0, 0, 0) == 0
    assert lib.caspr_cnf_workspace_bytes(2, 256, 512, 1600, 0) > 2 * 256 * 512 * 4 * 4
