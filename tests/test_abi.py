"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/gpb200.h declares (no compute calls - there is no GPU here), the ctypes table matches
the header, and the product package never touches the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'gpb200.h')


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(gpb_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_header_symbol():
    from gptest_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), 'libgpb200.so does not export %s' % s


def test_ctypes_table_matches_header():
    from gptest_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_symbols()
    src = re.sub(r'/\*.*?\*/', '', open(HEADER).read(), flags=re.S)
    for name, (_, args) in _lib.SIGNATURES.items():
        m = re.search(r'\b%s\s*\(([^;]*?)\)\s*;' % name, src, flags=re.S)
        assert m, name
        params = [p for p in m.group(1).split(',') if p.strip() and p.strip() != 'void']
        assert len(params) == len(args), (name, params, args)


def test_no_device_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from gptest_b200 import _lib
    with pytest.raises(_lib.GpbError):
        _lib.Handle(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'gptest_b200')
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(base, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', txt, flags=re.M), os.path.join(base, f)
                assert 'oracle.' not in re.sub(r'oracle/\w+\.py', '', txt), os.path.join(base, f)
                assert '/root/reference' not in txt.replace('/root/reference/GP', '') or f.endswith('.py')
    for f in ('GPr.py', 'GPc.py', 'GPpref.py'):
        p = os.path.join(ROOT, 'dropin', f)
        if os.path.exists(p):
            assert not re.search(r'^\s*(from|import)\s+oracle\b', open(p).read(), flags=re.M)
