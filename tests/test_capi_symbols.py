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


def test_new_struct_layouts():
    """arx_pool_push / arx_peer_seg as include/arx_b200.h declares them (the ctypes mirrors of INTEGRATION.md)."""
    assert ctypes.sizeof(_lib.PoolPush) == 40      # 2 pointers + 2 x int64 + 2 x int32
    assert ctypes.sizeof(_lib.PeerSeg) == 64       # 2 pointers + 5 x int64 + 2 x int32
    assert ctypes.sizeof(_lib.PoolReq) == 56       # 4 pointers + 2 x int64 + 2 x int32


def test_entry_points_reject_bad_arguments_before_touching_the_device():
    """Argument validation comes first in every entry point, so it can be checked without a GPU: NULL pointers, sizes out
    of range and unsupported modes return ARX_E_BADARG (-1) / ARX_E_UNSUPPORTED (-3), never a launch."""
    lib = _lib.load()
    plan = _lib.BwdPlan()                           # all-NULL plan
    assert lib.arx_pool_bwd_apply_slab(None, 1, 128, plan, None, 0, None, 0.1, None, 0, 64, None) == -1
    assert lib.arx_bwd_plan_alloc_h(None, plan, 64, None) == -1
    assert lib.arx_peer_push_many(None, 1, 2, None) == -1
    segs = (_lib.PeerSeg * 1)()
    assert lib.arx_peer_push_many(ctypes.addressof(segs), 9, 2, None) == -1          # more than 8 segments
    assert lib.arx_peer_push_many(ctypes.addressof(segs), 1, 2, None) == -1          # NULL source / destination
    assert lib.arx_peer_barrier(None, 0, 2, None, 1000, None, None) == -1
    assert lib.arx_peer_alloc(0, None) == -1
    assert lib.arx_peer_export(None, None) == -1 and lib.arx_peer_open(None, None) == -1
    assert lib.arx_pool_fwd_many_push(None, None, 1, 128, None) == -1
    assert lib.arx_score_max(None, None, 1, 1, None, None) == -1
    assert lib.arx_token_pool_fwd(None, None, 1, None, None, 1, 2, None, 1.0, None, None, None, None) == -1
    assert lib.arx_rowsum(None, 1, 1, None, 0, None) == -1
    assert lib.arx_set_tuning(b'heavy', 4) == -1 and lib.arx_set_tuning(b'no_such_knob', 1) == -1
    assert lib.arx_set_tuning(b'heavy', 64) == 0
