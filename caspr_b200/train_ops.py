"""Torch-tensor front-ends of the training operators of libcaspr_b200.so (include/caspr_b200.h, "encoder
training operators").  Same conventions as ``ops.py``: CUDA fp32 rows x channels views with a leading
dimension, outputs and workspaces allocated here with torch, work enqueued on the current stream.
"""
import ctypes

import torch

from ._lib import lib, check
from .ops import _p, _stream, _rows2d, _count, _f32


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def gn_moments(x, samples, rows_per_sample, groups, eps=1e-5):
    """(mean, rstd) per (sample, group) of GroupNorm(groups, C) over samples of consecutive rows."""
    x, ldx = _rows2d(x, 'x')
    C = x.shape[1]
    assert x.shape[0] == samples * rows_per_sample
    mr = torch.empty(samples, groups, 2, dtype=torch.float32, device=x.device)
    nb = lib.caspr_gn_workspace_bytes(samples, rows_per_sample, C)
    ws = _ws(nb, x.device)
    _count('gn_moments')
    check(lib.caspr_gn_moments(_p(x), ldx, samples, rows_per_sample, C, groups, float(eps), _p(mr), _p(ws), nb,
                               _stream()), 'caspr_gn_moments')
    return mr


def gn_apply(x, mr, samples, rows_per_sample, groups, gamma, beta, relu, out=None):
    x, ldx = _rows2d(x, 'x')
    C = x.shape[1]
    if out is None:
        out = torch.empty(x.shape[0], C, dtype=torch.float32, device=x.device)
    out, ldy = _rows2d(out, 'out')
    _count('gn_apply')
    check(lib.caspr_gn_apply(_p(x), ldx, _p(mr), samples, rows_per_sample, C, groups, _p(gamma), _p(beta), int(relu),
                             _p(out), ldy, _stream()), 'caspr_gn_apply')
    return out


def rowmax(y, samples, rows_per_sample, maxout=None):
    """max over each sample's rows -> (maxout (samples,C), argmax (samples,C) int32 row-in-sample)."""
    y, ldy = _rows2d(y, 'y')
    C = y.shape[1]
    if maxout is None:
        maxout = torch.empty(samples, C, dtype=torch.float32, device=y.device)
    maxout, ld_max = _rows2d(maxout, 'maxout')
    arg = torch.empty(samples, C, dtype=torch.int32, device=y.device)
    nb = lib.caspr_rowmax_workspace_bytes(samples, rows_per_sample, C)
    ws = _ws(nb, y.device)
    _count('rowmax')
    check(lib.caspr_rowmax(_p(y), ldy, samples, rows_per_sample, C, _p(maxout), ld_max, _p(arg), _p(ws), nb,
                           _stream()), 'caspr_rowmax')
    return maxout, arg


def gn_backward(x, mr, samples, rows_per_sample, groups, gamma, beta, relu, d_out=None, d_max=None, argmax=None):
    """-> (dX rows x C, dgamma, dbeta).  Output cotangent = d_out (rows,C view) + d_max (samples,C view) at argmax."""
    x, ldx = _rows2d(x, 'x')
    C = x.shape[1]
    lddy = ld_dmax = 0
    if d_out is not None:
        d_out, lddy = _rows2d(d_out, 'd_out')
        assert d_out.shape == x.shape
    if d_max is not None:
        d_max, ld_dmax = _rows2d(d_max, 'd_max')
        assert d_max.shape == (samples, C) and argmax is not None
    dx = torch.empty(x.shape[0], C, dtype=torch.float32, device=x.device)
    dgamma = torch.empty(C, dtype=torch.float32, device=x.device)
    dbeta = torch.empty(C, dtype=torch.float32, device=x.device)
    nb = lib.caspr_gn_workspace_bytes(samples, rows_per_sample, C)
    ws = _ws(nb, x.device)
    _count('gn_backward')
    check(lib.caspr_gn_backward(_p(d_out), lddy, _p(d_max), ld_dmax, _p(argmax), _p(x), ldx, _p(mr), samples,
                                rows_per_sample, C, groups, _p(gamma), _p(beta), int(relu), _p(dx), C, _p(dgamma),
                                _p(dbeta), _p(ws), nb, _stream()), 'caspr_gn_backward')
    return dx, dgamma, dbeta


WGRAD_ENGINE = 'auto'          # 'auto' | 'tc' | 'simt' (accuracy studies)


def linear_wgrad(d_y, x, relu_x=False, want_bias=True, engine=None):
    """dW (Cout,Cin) = d_y^T . act(x), db (Cout) = column sums of d_y.  Large layers (rows >= 2048, Cout, Cin >= 64)
    run on the split-K tcgen05 fp16x3 GEMM, the rest on the exact-fp32 SIMT kernel."""
    d_y, lddy = _rows2d(d_y, 'd_y')
    x, ldx = _rows2d(x, 'x')
    rows, cout = d_y.shape
    cin = x.shape[1]
    assert x.shape[0] == rows
    dW = torch.empty(cout, cin, dtype=torch.float32, device=x.device)
    db = torch.empty(cout, dtype=torch.float32, device=x.device) if want_bias else None
    engine = engine or WGRAD_ENGINE
    if engine == 'auto':
        engine = 'tc' if (rows >= 2048 and cout >= 64 and cin >= 64) else 'simt'
    if engine == 'tc':
        nb = lib.caspr_linear_wgrad_tc_workspace_bytes(rows, cout, cin)
        buf = torch.empty(nb + 1024, dtype=torch.uint8, device=x.device)
        ptr = (buf.data_ptr() + 1023) // 1024 * 1024
        _count('linear_wgrad_tc')
        check(lib.caspr_linear_wgrad_tc(_p(d_y), lddy, _p(x), ldx, rows, cout, cin, int(relu_x), _p(dW), None, None,
                                        ctypes.c_void_p(ptr), nb, _stream()), 'caspr_linear_wgrad_tc')
        if want_bias:
            colsum(d_y, db)
        return dW, db
    nb = lib.caspr_linear_wgrad_workspace_bytes(rows, cout, cin)
    ws = _ws(nb, x.device)
    _count('linear_wgrad')
    check(lib.caspr_linear_wgrad(_p(d_y), lddy, _p(x), ldx, rows, cout, cin, int(relu_x), _p(dW), _p(db), _p(ws), nb,
                                 _stream()), 'caspr_linear_wgrad')
    return dW, db


def colsum(x, out, accumulate=False):
    x, ldx = _rows2d(x, 'x')
    rows, C = x.shape
    assert out.is_contiguous() and out.numel() == C
    nb = lib.caspr_colsum_workspace_bytes(rows, C)
    ws = _ws(nb, x.device)
    _count('colsum')
    check(lib.caspr_colsum(_p(x), ldx, rows, C, _p(out), int(accumulate), _p(ws), nb, _stream()), 'caspr_colsum')
    return out


def group_points_bwd(d_out, idx, N, C, d_feat):
    """d_feat (B,N,C) channels-last view (accumulated into) from d_out rows (B*M*ns, 3+C)."""
    d_out, ld_out = _rows2d(d_out, 'd_out')
    B, M, ns = idx.shape
    assert d_feat.dim() == 3 and d_feat.stride(2) == 1 and d_feat.stride(0) == N * d_feat.stride(1)
    _count('group_points_bwd')
    check(lib.caspr_group_points_bwd(_p(d_out), ld_out, _p(idx), B, N, M, C, ns, _p(d_feat), d_feat.stride(1),
                                     _stream()), 'caspr_group_points_bwd')


def three_interp_bwd(d_out, idx, dist, m, Cp, d_prev):
    d_out, ld_out = _rows2d(d_out, 'd_out')
    B, n, _ = idx.shape
    assert d_prev.dim() == 3 and d_prev.stride(2) == 1 and d_prev.stride(0) == m * d_prev.stride(1)
    _count('three_interp_bwd')
    check(lib.caspr_three_interp_bwd(_p(d_out), ld_out, _p(idx), _p(dist), B, n, m, Cp, _p(d_prev),
                                     d_prev.stride(1), _stream()), 'caspr_three_interp_bwd')


def rows_update(src, dst, accumulate=False, relu_ref=None):
    """dst (+)= src (2-D views), optionally only where relu_ref > 0."""
    src, ld_src = _rows2d(src, 'src')
    dst, ld_dst = _rows2d(dst, 'dst')
    assert src.shape == dst.shape
    ld_ref = 0
    if relu_ref is not None:
        relu_ref, ld_ref = _rows2d(relu_ref, 'relu_ref')
    _count('rows_update')
    check(lib.caspr_rows_update(_p(src), ld_src, src.shape[0], src.shape[1], int(accumulate), _p(relu_ref), ld_ref,
                                _p(dst), ld_dst, _stream()), 'caspr_rows_update')
    return dst


def transpose(w):
    """(rows, cols) contiguous -> (cols, rows) contiguous."""
    _f32(w, 'w')
    assert w.dim() == 2 and w.is_contiguous()
    out = torch.empty(w.shape[1], w.shape[0], dtype=torch.float32, device=w.device)
    _count('transpose')
    check(lib.caspr_transpose(_p(w), w.shape[0], w.shape[1], _p(out), _stream()), 'caspr_transpose')
    return out
