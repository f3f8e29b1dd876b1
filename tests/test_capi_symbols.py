"""CPU test: the C-ABI library loads and exports every symbol include/arx_b200.h declares
(no compute calls without a GPU)."""
import ctypes
import os
import re

import arecsys_b200  # noqa: F401
from arecsys_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'arx_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(arx_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_something():
    syms = declared_symbols()
    assert 'arx_pool_fwd' in syms and 'arx_pool_bwd_apply' in syms and len(syms) >= 15


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, 'symbols declared in include/arx_b200.h but not exported: %s' % missing


def test_binding_covers_every_declared_symbol():
    bound = set(_lib.SIGNATURES) | {'arx_abi_version', 'arx_build_info'}
    assert set(declared_symbols()) == bound


def test_abi_version_and_struct_layout():
    lib = _lib.load()
    assert lib.arx_abi_version() == 4
    assert ctypes.sizeof(_lib.AttrDesc) == 88      # 8 pointers + int64 + 2 x int32 + 1 pointer
    assert ctypes.sizeof(_lib.BwdPlan) == 112      # 11 pointers + 3 x int64
    assert b'sm_100a' in lib.arx_build_info()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'a-recsys_b200')
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                txt = open(os.path.join(d, f)).read()
                assert 'import oracle' not in txt and 'from oracle' not in txt, f


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
    --impl reference legs may import it (the product path has no CPU fallback)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    offenders = []
    for top in ('a-recsys_b200', 'hmf', 'lstm', 'word2vec', 'tools'):
        for dp, _, files in os.walk(os.path.join(root, top)):
            for f in files:
                if f.endswith('.py'):
                    src = open(os.path.join(dp, f), errors='replace').read()
                    if re.search(r'^\s*(from|import)\s+oracle\b', src, re.M) or 'tf1_shim' in src:
                        offenders.append(os.path.join(dp, f))
    assert offenders == [], offenders
