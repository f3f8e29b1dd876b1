"""ctypes binding of libarx_b200.so (the C ABI declared in include/arx_b200.h).

The product path fails loudly when the CUDA library is missing or a call returns an
error code: there is no CPU fallback anywhere in this package.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('ARX_LIB') or os.path.join(_HERE, 'lib', 'libarx_b200.so')     # ARX_LIB: A/B of build variants

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_f32p = ctypes.POINTER(ctypes.c_float)
vp = ctypes.c_void_p


class AttrDesc(ctypes.Structure):
    """arx_attr_desc (include/arx_b200.h)."""
    _fields_ = [('table', vp), ('table_acc', vp), ('bias', vp), ('bias_acc', vp),
                ('values', vp), ('starts', vp), ('lengths', vp), ('touch', vp),
                ('vocab', ctypes.c_int64), ('kind', ctypes.c_int32), ('reserved', ctypes.c_int32),
                ('lengths_full', vp)]


class PoolReq(ctypes.Structure):
    """arx_pool_req (include/arx_b200.h)."""
    _fields_ = [('attrs', vp), ('ent_ids', vp), ('out', vp), ('bias_out', vp), ('n', ctypes.c_int64),
                ('out_stride', ctypes.c_int64), ('n_attr', ctypes.c_int32), ('max_rows_per_entity', ctypes.c_int32)]


class PoolPush(ctypes.Structure):
    """arx_pool_push (include/arx_b200.h)."""
    _fields_ = [('peer_out', vp), ('peer_bias', vp), ('rows_per_rank', ctypes.c_int64), ('stride', ctypes.c_int64),
                ('n_ranks', ctypes.c_int32), ('reserved', ctypes.c_int32)]


class PeerSeg(ctypes.Structure):
    """arx_peer_seg (include/arx_b200.h)."""
    _fields_ = [('src', vp), ('dst', vp), ('rows', ctypes.c_int64), ('width', ctypes.c_int64), ('src_stride', ctypes.c_int64),
                ('dst_stride', ctypes.c_int64), ('row0', ctypes.c_int64), ('mode', ctypes.c_int32), ('reserved', ctypes.c_int32)]


class BwdPlan(ctypes.Structure):
    """arx_bwd_plan (include/arx_b200.h)."""
    _fields_ = [('counters', vp), ('uniq_tok', vp), ('uniq_attr', vp), ('row_base', vp),
                ('row_cnt', vp), ('bucket_src', vp), ('bucket_w', vp), ('chunk_row', vp),
                ('row_chunk0', vp), ('row_done', vp), ('partials', vp),
                ('cap_rows', ctypes.c_int64), ('cap_occ', ctypes.c_int64), ('cap_chunks', ctypes.c_int64)]


class ApplySet(ctypes.Structure):
    """arx_apply_set (include/arx_b200.h)."""
    _fields_ = [('attrs', vp), ('dout', vp), ('dbias', vp), ('dout_stride', ctypes.c_int64), ('plan', BwdPlan),
                ('n_attr', ctypes.c_int32), ('reserved', ctypes.c_int32)]


POOL_MEAN, POOL_CONCAT = 0, 1
OPT_ADAGRAD, OPT_SGD, OPT_NONE = 0, 1, 2
HEAVY_DEFAULT = 64          # default of arx_set_tuning("heavy"): keep in step with g_tune_heavy in csrc/pool.cu
LOSS_KIND = {'ce': 0, 'warp': 1, 'warp_eval': 1, 'rs': 2, 'rs-sig': 3, 'rs-sig2': 4, 'bbpr': 5, 'mw': 6}
LOSS_FUNC = {'log': 0, 'exp': 1, 'poly': 2, 'poly2': 3, 'linear': 4, 'square': 5}

i64, i32, f32 = ctypes.c_int64, ctypes.c_int, ctypes.c_float

# name -> argtypes; every function returns int except the two info calls.
SIGNATURES = {
    'arx_pool_fwd': [vp, i32, i32, vp, i64, vp, i64, i32, vp, i32, vp],
    'arx_mulhot_flat_index': [vp, i32, vp, i64, vp, vp, vp, vp],
    'arx_bwd_plan_begin': [BwdPlan, vp],
    'arx_bwd_plan_count': [vp, i32, i32, vp, i64, BwdPlan, vp],
    'arx_bwd_plan_alloc': [vp, BwdPlan, vp],
    'arx_bwd_plan_fill': [vp, i32, i32, vp, i64, i32, i64, BwdPlan, vp],
    'arx_bwd_plan_end': [vp, BwdPlan, vp],
    'arx_pool_bwd_plan': [vp, i32, vp, i64, i32, BwdPlan, vp],
    'arx_pool_bwd_apply': [vp, i32, i32, BwdPlan, vp, i64, vp, f32, vp, i32, vp, vp, vp],
    'arx_pool_bwd_apply_many': [vp, i32, i32, f32, vp, i32, vp],
    'arx_pool_bwd_sumsq': [vp, i32, BwdPlan, vp, i64, vp, vp, i32, vp],
    'arx_rows_sumsq': [vp, vp, BwdPlan, i32, vp, vp],
    'arx_set_tuning': [ctypes.c_char_p, i32],
    'arx_gemm': [vp, vp, vp, i64, i64, i64, i32, i32, vp, f32, f32, vp],
    'arx_gemm_tc': [vp, vp, vp, i64, i64, i64, i32, i32, vp, f32, f32, vp],
    'arx_ce_workspace_floats': [i64, i64, vp],
    'arx_ce_fwd': [vp, vp, vp, i64, i64, i64, vp, vp, vp],
    'arx_ce_bwd': [vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, i64, vp, vp, vp, vp],
    'arx_ce_rowloss': [vp, vp, vp, vp, vp, i64, i64, i64, vp, vp],
    'arx_mw_mask_words': [i64, vp],
    'arx_mw_mask_build': [vp, vp, vp, i64, i64, vp, i64, vp],
    'arx_mw_fwd': [vp, vp, vp, vp, vp, i64, i64, i64, i64, vp, vp, vp, vp],
    'arx_mw_bwd': [vp, vp, vp, vp, vp, vp, vp, i64, vp, vp, i64, i64, i64, vp, vp, vp, vp, vp],
    'arx_mw_bwd2': [vp, vp, vp, vp, vp, vp, vp, i64, vp, vp, i64, i64, i64, vp, vp, vp, vp, i32, vp],
    'arx_lstm_gates_fwd': [vp, vp, vp, vp, i64, i32, f32, vp],
    'arx_lstm_gates_bwd': [vp, vp, vp, vp, vp, vp, vp, i64, i32, vp],
    'arx_lstm_gates_fwd2': [vp, vp, vp, vp, vp, i64, i32, f32, vp],
    'arx_lstm_gates_bwd2': [vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, vp],
    'arx_pool_fwd_many': [vp, i32, i32, vp],
    'arx_mw_prep': [vp, vp, f32, vp, vp, vp, vp, vp, i64, i64, i32, vp, vp, vp, vp, vp, vp, vp],
    'arx_mw_post': [vp, vp, vp, vp, vp, f32, i64, i32, vp, vp, vp, vp],
    'arx_pool_fwd_many_push': [vp, vp, i32, i32, vp],
    'arx_peer_alloc': [i64, vp],
    'arx_peer_free': [vp],
    'arx_peer_export': [vp, vp],
    'arx_peer_open': [vp, vp],
    'arx_peer_close': [vp],
    'arx_peer_barrier': [vp, i32, i32, vp, i64, vp, vp],
    'arx_peer_push_rows': [vp, i64, i64, i64, vp, i64, i64, i32, i32, i32, vp],
    'arx_peer_push_many': [vp, i32, i32, vp],
    'arx_bwd_plan_alloc_h': [vp, BwdPlan, i32, vp],
    'arx_pool_bwd_apply_slab': [vp, i32, i32, BwdPlan, vp, i32, vp, f32, vp, i32, i32, vp],
    'arx_score_max': [vp, vp, i64, i64, vp, vp],
    'arx_token_pool_fwd': [vp, vp, i64, vp, vp, i64, i32, vp, f32, vp, vp, vp, vp],
    'arx_token_pool_bwd': [vp, vp, vp, i64, vp, vp, i64, i32, vp, f32, vp, vp, vp, vp, vp],
    'arx_rowsum': [vp, i64, i64, vp, i32, vp],
    'arx_gather_pairs': [vp, vp, vp, i64, vp, vp, vp],
    'arx_cbow_window_batch': [vp, vp, vp, vp, i64, i64, i64, i32, i32, i32, vp, vp, vp, vp, vp],
    'arx_lstm_pad_batch': [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp],
    'arx_gumbel_keys': [vp, i64, vp, vp, vp],
    'arx_lstm_seq_fwd': [vp, vp, vp, vp, i64, i64, i32, f32, vp],
    'arx_lstm_seq_bwd': [vp, vp, vp, vp, i64, i64, i32, vp],
    'arx_axpby_rows': [vp, vp, f32, f32, i64, i64, i32, vp, vp],
    'arx_sum_over_steps': [vp, i64, i64, i32, f32, vp, vp],
    'arx_transpose': [vp, i64, i64, vp, i32, vp],
    'arx_round_tf32': [vp, vp, i64, vp],
    'arx_colsum': [vp, i64, i64, i64, vp, vp],
    'arx_loss_rows': [vp, i64, i64, i64, vp, vp, vp, vp, vp, i32, i32, f32, vp, vp, vp, vp, vp, vp],
    'arx_rowdot_fwd': [vp, vp, vp, i64, i32, vp, vp],
    'arx_rowdot_bwd': [vp, vp, vp, i64, i32, vp, vp, vp],
    'arx_topk_rows': [vp, i64, i64, i64, i32, vp, vp, vp],
    'arx_dense_update': [vp, vp, vp, i64, f32, vp, i32, vp],
    'arx_scale_mask': [vp, vp, f32, i64, vp, vp],
}

_ERR = {-1: 'ARX_E_BADARG', -2: 'ARX_E_LAUNCH', -3: 'ARX_E_UNSUPPORTED', -4: 'ARX_E_CAPACITY'}

_lib = None
TUNING = {}        # arx_set_tuning settings taken from ARX_TUNE at load time (the plan capacity follows 'heavy')
launch_count = 0   # number of C-ABI compute calls issued (bench.py's gpu_launches evidence)


def load():
    """dlopen the in-tree library; raises (never falls back) if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError('libarx_b200.so not built: run `python -c "import __graft_entry__ as g; '
                           'g.build()"` or `make -C a-recsys_b200/csrc` (expected at %s)' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    lib.arx_abi_version.argtypes = []
    lib.arx_abi_version.restype = ctypes.c_int
    lib.arx_build_info.argtypes = []
    lib.arx_build_info.restype = ctypes.c_char_p
    _lib = lib
    for kv in filter(None, os.environ.get('ARX_TUNE', '').split(',')):     # e.g. ARX_TUNE=plan_agg=1,flat_epb=5
        k, v = kv.split('=')
        if lib.arx_set_tuning(k.encode(), int(v)) != 0:
            raise RuntimeError('ARX_TUNE: bad setting %r' % kv)
        TUNING[k] = int(v)
    return lib


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


timeline = None    # when a list: (name, tag, start_event, end_event) per call (bench.py roofline leg)
tag = ''           # free-form label the host sets to tell apart uses of one kernel (e.g. 'user'/'item')


def call(name, *args):
    """Invoke a C-ABI entry point on torch's current stream; raise on any error code."""
    global launch_count
    lib = load()
    if timeline is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args, stream())
        e1.record()
        timeline.append((name, tag, e0, e1))
    else:
        rc = getattr(lib, name)(*args, stream())
    launch_count += 1
    if rc != 0 and not (rc == -3 and name in _MAY_BE_UNSUPPORTED):
        raise RuntimeError('%s failed: %s (%d)' % (name, _ERR.get(rc, '?'), rc))
    return rc


_MAY_BE_UNSUPPORTED = ('arx_gemm_tc', 'arx_ce_fwd', 'arx_ce_bwd', 'arx_mw_fwd', 'arx_mw_bwd', 'arx_mw_bwd2', 'arx_lstm_seq_fwd',
                       'arx_lstm_seq_bwd', 'arx_pool_fwd_many', 'arx_mw_prep', 'arx_mw_post', 'arx_pool_bwd_apply_many')
exact_fp32 = False   # True: every contraction on the exact-fp32 SIMT kernel (parity anchor runs)


ROUND_LIMIT = 1 << 26   # elements; larger operands go to the tensor core un-rounded (truncated)


def _rounded(X, numel):
    if numel > ROUND_LIMIT:
        return X
    Y = torch.empty(numel, dtype=torch.float32, device=X.device)
    call('arx_round_tf32', X.data_ptr(), Y.data_ptr(), numel)
    return Y


def gemm(A, B, C, m, n, k, trans_a, trans_b, bias=None, alpha=1.0, beta=0.0, a_ready=False, b_ready=False):
    """C = alpha * op(A) op(B) + bias: tensor cores (tcgen05, tf32) when TMA can describe the
    operands, else the exact-fp32 SIMT kernel.  Both are this library's own CUDA kernels."""
    if not exact_fp32 and k % 4 == 0 and k > 0:
        # the tensor-core kernel takes K-major operands: A [m,k], B [n,k]; an operand stored the
        # other way round is staged through arx_transpose first (small next to the contraction)
        # Operands are rounded to the nearest tf32 on the way (for free inside the transpose, one
        # extra elementwise pass otherwise; skipped for operands beyond ROUND_LIMIT elements).
        if trans_a:
            Ak = torch.empty((m, k), dtype=torch.float32, device=A.device)
            call('arx_transpose', A.data_ptr(), k, m, Ak.data_ptr(), 1)
        else:
            Ak = A if a_ready else _rounded(A, m * k)
        if not trans_b:
            Bk = torch.empty((n, k), dtype=torch.float32, device=B.device)
            call('arx_transpose', B.data_ptr(), k, n, Bk.data_ptr(), 1)
        else:
            Bk = B if b_ready else _rounded(B, n * k)
        if call('arx_gemm_tc', Ak.data_ptr(), Bk.data_ptr(), C.data_ptr(), m, n, k, 0, 1, ptr(bias),
                alpha, beta) == 0:
            return
    call('arx_gemm', A.data_ptr(), B.data_ptr(), C.data_ptr(), m, n, k, trans_a, trans_b, ptr(bias), alpha, beta)


def ce_supported(M, N, d):
    """Shapes the fused scoring + cross-entropy kernels take (arx_ce_fwd / arx_ce_bwd)."""
    return (not exact_fp32) and d in (32, 64, 96, 128) and M % 4 == 0 and N % 4 == 0 and M > 0 and N > 0


def ce_fwd(U_r, P_r, beta, M, N, d):
    """lse[M] of logits = U P^T + beta, logits never materialised.  U_r / P_r: tf32-rounded operands."""
    n = ctypes.c_int64(0)
    rc = load().arx_ce_workspace_floats(M, N, ctypes.byref(n))
    if rc != 0:
        raise RuntimeError('arx_ce_workspace_floats failed (%d)' % rc)
    ws = torch.empty(n.value, dtype=torch.float32, device=U_r.device)
    lse = torch.empty(M, dtype=torch.float32, device=U_r.device)
    if call('arx_ce_fwd', U_r.data_ptr(), P_r.data_ptr(), ptr(beta), M, N, d, ws.data_ptr(), lse.data_ptr()) != 0:
        return None
    return lse


def round_tf32(X):
    """tf32-nearest copy of X (any size; the fused CE kernels want rounded operands)."""
    Y = torch.empty_like(X)
    call('arx_round_tf32', X.data_ptr(), Y.data_ptr(), X.numel())
    return Y


def ce_rowloss(U_r, P_r, beta, target, lse, M, N, d):
    loss = torch.empty(M, dtype=torch.float32, device=U_r.device)
    call('arx_ce_rowloss', U_r.data_ptr(), P_r.data_ptr(), ptr(beta), target.data_ptr(), lse.data_ptr(), M, N, d,
         loss.data_ptr())
    return loss


def ce_bwd(U_r, P_r, beta, lse, g, target, M, N, d, dP=None):
    """(dU, dP, dbeta) of sum_r g[r] * (lse[r] - logits[r, target[r]])."""
    dev = U_r.device
    UT = torch.empty((d, M), dtype=torch.float32, device=dev)
    PT = torch.empty((d, N), dtype=torch.float32, device=dev)
    call('arx_transpose', U_r.data_ptr(), M, d, UT.data_ptr(), 0)
    call('arx_transpose', P_r.data_ptr(), N, d, PT.data_ptr(), 0)
    dU = torch.empty((M, d), dtype=torch.float32, device=dev)
    if dP is None:
        dP = torch.empty((N, d), dtype=torch.float32, device=dev)
    dbeta = torch.empty((N,), dtype=torch.float32, device=dev)
    if call('arx_ce_bwd', U_r.data_ptr(), P_r.data_ptr(), UT.data_ptr(), PT.data_ptr(), ptr(beta), lse.data_ptr(),
            g.data_ptr(), target.data_ptr(), M, N, d, dU.data_ptr(), dP.data_ptr(), dbeta.data_ptr()) != 0:
        return None
    return dU, dP, dbeta


def mw_mask_words(N):
    n = ctypes.c_int64(0)
    rc = load().arx_mw_mask_words(N, ctypes.byref(n))
    if rc != 0:
        raise RuntimeError('arx_mw_mask_words failed (%d)' % rc)
    return n.value


def mw_fwd(U_r, P_r, beta, tscore, mask, mask_ld, M, N, d):
    """(hsum[M], loss[M]) of the sampled WMRB loss; scores never materialised.  None if unsupported."""
    n = ctypes.c_int64(0)
    rc = load().arx_ce_workspace_floats(M, N, ctypes.byref(n))
    if rc != 0:
        raise RuntimeError('arx_ce_workspace_floats failed (%d)' % rc)
    ws = torch.empty(n.value, dtype=torch.float32, device=U_r.device)
    hsum = torch.empty(M, dtype=torch.float32, device=U_r.device)
    loss = torch.empty(M, dtype=torch.float32, device=U_r.device)
    if call('arx_mw_fwd', U_r.data_ptr(), P_r.data_ptr(), ptr(beta), tscore.data_ptr(), ptr(mask), mask_ld, M, N, d,
            ws.data_ptr(), hsum.data_ptr(), loss.data_ptr()) != 0:
        return None
    return hsum, loss


def mw_bwd(U_r, P_r, beta, tscore, mask, mask_ld, hsum, g, M, N, d, dP=None, UT=None, PT=None, outputs=None, accumulate=False):
    """(dU, dP, dbeta, dts) of sum_r g[r] * loss[r].  UT / PT: the transposed operands when the caller already has
    them (arx_mw_prep emits them with the rounding pass)."""
    dev = U_r.device
    if UT is None:
        UT = torch.empty((d, M), dtype=torch.float32, device=dev)
        call('arx_transpose', U_r.data_ptr(), M, d, UT.data_ptr(), 0)
    if PT is None:
        PT = torch.empty((d, N), dtype=torch.float32, device=dev)
        call('arx_transpose', P_r.data_ptr(), N, d, PT.data_ptr(), 0)
    zeroed = 0
    if outputs is not None:                     # (dU, dP, dbeta, dts) zeroed by the caller, off the dependent chain
        dU, dP, dbeta, dts = outputs
        zeroed = 2 if accumulate else 1
    else:
        dU = torch.empty((M, d), dtype=torch.float32, device=dev)
        if dP is None:
            dP = torch.empty((N, d), dtype=torch.float32, device=dev)
        dbeta = torch.empty((N,), dtype=torch.float32, device=dev)
        dts = torch.empty((M,), dtype=torch.float32, device=dev)
    if call('arx_mw_bwd2', U_r.data_ptr(), P_r.data_ptr(), UT.data_ptr(), PT.data_ptr(), ptr(beta), tscore.data_ptr(),
            ptr(mask), mask_ld, hsum.data_ptr(), g.data_ptr(), M, N, d, dU.data_ptr(), dP.data_ptr(), dbeta.data_ptr(),
            dts.data_ptr(), zeroed) != 0:
        return None
    return dU, dP, dbeta, dts
