"""Torch-tensor front-ends of the C-ABI entry points (include/caspr_b200.h).

Every function takes CUDA fp32 / int32 tensors, allocates outputs with torch (the library
never allocates), enqueues on torch's current stream and raises ``CasprError`` on a non-zero
status.  Activations are "rows x channels" (channels-last); a tensor's row stride is passed
as the leading dimension, so column slices of a wider buffer can be read and written in
place (that is how the concat buffers of the encoder are filled without copies).

These replace the Kaolin ops bound at reference caspr/models/pointnet2.py:7-10, the torch
Conv1d/GroupNorm calls around them, and torchdiffeq's odeint (latent_ode_model.py:98,
cnf.py:102-119).
"""
import ctypes
import os
import weakref

import torch

from . import _lib
from ._lib import lib, check

ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2
CNF_SIMT_FP32, CNF_TC_FP16X3 = 0, 1

# kernels launched through this module since the last reset (bench.py's gpu_launches claim is
# derived from the per-entry-point launch counts documented in DESIGN.md)
CALLS = {}


def _count(name):
    CALLS[name] = CALLS.get(name, 0) + 1


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _f32(t, name):
    if not (t.is_cuda and t.dtype == torch.float32):
        raise TypeError('%s must be a CUDA float32 tensor, got %s on %s' % (name, t.dtype, t.device))
    return t


def _rows2d(t, name):
    """(rows, C) view requirements: unit stride along channels; returns (tensor, ld)."""
    _f32(t, name)
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError('%s must be 2-D with unit channel stride' % name)
    return t, t.stride(0)


# ------------------------------------------------------------------------------- geometry
def fps(xyz, m):
    """furthest_point_sampling + gather (pointnet2.py:384-387): xyz (B,N,3) -> idx (B,M) int32, new_xyz (B,M,3)."""
    _f32(xyz, 'xyz')
    xyz = xyz.contiguous()
    B, N, _ = xyz.shape
    idx = torch.empty(B, m, dtype=torch.int32, device=xyz.device)
    new_xyz = torch.empty(B, m, 3, dtype=torch.float32, device=xyz.device)
    _count('fps')
    check(lib.caspr_fps(_p(xyz), B, N, m, _p(idx), _p(new_xyz), _stream()), 'caspr_fps')
    return idx, new_xyz


def ball_query2(xyz, new_xyz, r0, ns0, r1, ns1):
    """Both ball queries of one SA level in one scan -> idx0 (B,M,ns0), idx1 (B,M,ns1) int32."""
    xyz, new_xyz = _f32(xyz, 'xyz').contiguous(), _f32(new_xyz, 'new_xyz').contiguous()
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    idx0 = torch.empty(B, M, ns0, dtype=torch.int32, device=xyz.device)
    idx1 = torch.empty(B, M, ns1, dtype=torch.int32, device=xyz.device)
    _count('ball_query2')
    check(lib.caspr_ball_query2(_p(xyz), _p(new_xyz), B, N, M, float(r0), ns0, _p(idx0), float(r1), ns1,
                                _p(idx1), _stream()), 'caspr_ball_query2')
    return idx0, idx1


def group_points(xyz, new_xyz, feat, idx):
    """Grouping layer output as rows: (B*M*ns, 3+C) = [xyz[idx]-centre | feat[idx]].

    feat: (B,N,C) channels-last view with unit channel stride (row stride free) or None."""
    B, N, _ = xyz.shape
    M, ns = idx.shape[1], idx.shape[2]
    C, ld_feat = 0, 0
    if feat is not None:
        _f32(feat, 'feat')
        assert feat.dim() == 3 and feat.stride(2) == 1 and feat.stride(0) == N * feat.stride(1)
        C, ld_feat = feat.shape[2], feat.stride(1)
    out = torch.empty(B * M * ns, 3 + C, dtype=torch.float32, device=xyz.device)
    _count('group_points')
    check(lib.caspr_group_points(_p(xyz), _p(new_xyz), _p(feat), ld_feat, _p(idx), B, N, M, C, ns, _p(out),
                                 3 + C, _stream()), 'caspr_group_points')
    return out


def three_nn(unknown, known):
    """three_nn (pointnet2.py:514): -> dist (B,n,3) Euclidean, idx (B,n,3) int32."""
    unknown, known = _f32(unknown, 'unknown').contiguous(), _f32(known, 'known').contiguous()
    B, n, _ = unknown.shape
    m = known.shape[1]
    dist = torch.empty(B, n, 3, dtype=torch.float32, device=unknown.device)
    idx = torch.empty(B, n, 3, dtype=torch.int32, device=unknown.device)
    _count('three_nn')
    check(lib.caspr_three_nn(_p(unknown), _p(known), B, n, m, _p(dist), _p(idx), _stream()), 'caspr_three_nn')
    return dist, idx


def three_interp_concat(feat_prev, idx, dist, skip):
    """pointnet2.py:516-523: rows (B*n, Cp+Cs) = [inverse-distance interpolation of feat_prev | skip].

    feat_prev (B,m,Cp), skip (B,n,Cs) channels-last views (unit channel stride) or None."""
    B, n, _ = idx.shape
    m, Cp = feat_prev.shape[1], feat_prev.shape[2]
    assert feat_prev.stride(2) == 1 and feat_prev.stride(0) == m * feat_prev.stride(1)
    Cs, ld_skip = 0, 0
    if skip is not None:
        assert skip.stride(2) == 1 and skip.stride(0) == n * skip.stride(1)
        Cs, ld_skip = skip.shape[2], skip.stride(1)
    out = torch.empty(B * n, Cp + Cs, dtype=torch.float32, device=idx.device)
    _count('three_interp_concat')
    check(lib.caspr_three_interp_concat(_p(feat_prev), feat_prev.stride(1), _p(idx), _p(dist), _p(skip), ld_skip,
                                        B, n, m, Cp, Cs, _p(out), Cp + Cs, _stream()),
          'caspr_three_interp_concat')
    return out


# ------------------------------------------------------------------------------ dense ops
# layers at least this large run on the tcgen05 fp16x3 GEMM, the rest on the exact-fp32 SIMT kernel
TC_MIN_ROWS, TC_MIN_CIN, TC_MIN_COUT = 2048, 64, 64
LINEAR_ENGINE = 'auto'          # module-wide override used by accuracy studies: 'auto' | 'tc' | 'simt'


# fp16 hi/lo planes of layer weights, keyed by the weight tensor's storage and in-place version counter
_WEIGHT_PLANES = {}


def _aligned_bytes(nbytes, device):
    """uint8 buffer and a 1024-byte aligned pointer into it."""
    buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
    return buf, (buf.data_ptr() + 1023) // 1024 * 1024


def _prepared_weights(weight, w2d, owner=None):
    """fp16 hi/lo planes of `weight`, cached per tensor OBJECT (weak reference: recycled storage addresses of
    a freed model never alias), storage pointer and in-place version counter.  `owner`: the parameter a derived
    matrix (e.g. a column subset) was built from; its version then invalidates the planes."""
    if owner is not None:
        version_of = owner
    else:
        version_of = weight
    key = id(weight)
    hit = _WEIGHT_PLANES.get(key)
    if hit is not None:
        ref, ptr, version, buf, planes_ptr = hit
        if ref() is weight and ptr == w2d.data_ptr() and version == version_of._version:
            return planes_ptr
    cout, cin = w2d.shape
    nbytes = lib.caspr_linear_tc_weight_bytes(cin, cout)
    buf, planes_ptr = _aligned_bytes(nbytes, w2d.device)
    _count('linear_tc_prepare_weights')
    check(lib.caspr_linear_tc_prepare_weights(_p(w2d), cin, cin, cout, ctypes.c_void_p(planes_ptr), nbytes, _stream()),
          'caspr_linear_tc_prepare_weights')
    ref = weakref.ref(weight, lambda _r, k=key: _WEIGHT_PLANES.pop(k, None))
    _WEIGHT_PLANES[key] = (ref, w2d.data_ptr(), version_of._version, buf, planes_ptr)
    return planes_ptr


class PendingNorm(object):
    """GroupNorm (+ReLU) that has NOT been applied to a tensor yet: the statistics were accumulated by the GEMM
    that produced it; the next tensor-core linear folds the normalisation into its operand split
    (``linear(..., in_norm=...)``) or ``groupnorm(..., stats=...)`` materialises it."""

    def __init__(self, stats, samples, rows_per_sample, groups, channels, gamma, beta, relu, eps=1e-5):
        self.stats, self.samples, self.rows_per_sample, self.groups = stats, samples, rows_per_sample, groups
        self.channels, self.gamma, self.beta, self.relu, self.eps = channels, gamma, beta, relu, eps
        self._table = None

    def table(self):
        if self._table is None:
            self._table = torch.empty(self.samples * self.channels * 2, dtype=torch.float32,
                                      device=self.stats.device)
            _count('gn_table')
            check(lib.caspr_gn_table(_p(self.stats), self.samples, self.groups, self.rows_per_sample, self.channels,
                                     float(self.eps), _p(self.gamma), _p(self.beta), _p(self._table), _stream()),
                  'caspr_gn_table')
        return self._table


_DERIVED = {}


def derived_weight(param, tag, fn):
    """A matrix derived from `param` (column subsets of the head's first layer), rebuilt when the parameter's storage
    or in-place version changes; cached per parameter object."""
    key = (id(param), tag)
    hit = _DERIVED.get(key)
    if hit is not None:
        ref, ptr, version, value = hit
        if ref() is param and ptr == param.data_ptr() and version == param._version:
            return value
    value = fn(param.detach()).contiguous()
    ref = weakref.ref(param, lambda _r, k=key: _DERIVED.pop(k, None))
    _DERIVED[key] = (ref, param.data_ptr(), param._version, value)
    return value


def invalidate_weight_cache():
    """Drop every cached set of fp16 weight planes / derived matrices.  The caches are keyed by the parameters' in-place
    version counters, which `p.data.copy_()`-style edits do NOT bump: call this (and `model.encoder.reset_graphs()`)
    after modifying weights through `.data`."""
    _WEIGHT_PLANES.clear()
    _DERIVED.clear()


def tc_eligible(rows, cin, cout):
    return (LINEAR_ENGINE != 'simt' and rows >= TC_MIN_ROWS and cin >= TC_MIN_CIN and cout >= TC_MIN_COUT and
            cout % 4 == 0)


def conv_gn_relu_conv(x, conv_a, gn_a, conv_b, samples, rows_per_sample, groups, out=None, stats_b=False,
                      reduce_only=False):
    """Conv1d -> GroupNorm -> ReLU -> Conv1d on rows.  On the tensor-core path the first GEMM's epilogue
    accumulates the GroupNorm statistics and the second GEMM normalises while splitting its operand, so the
    normalised intermediate never exists in memory.  Returns y (and y's statistics if stats_b).
    reduce_only (with stats_b): if the tensor-core path applies, y itself is not produced either - returns
    (None, (statistics, per-channel extrema)) for `gn_max_from_extrema`; otherwise (y, None) as usual."""
    rows, cin = x.shape
    ca, cb = conv_a.weight.shape[0], conv_b.weight.shape[0]
    # the GEMM epilogue keeps one running GroupNorm group per 32-column chunk: groups of >= 32 channels only
    fold = (tc_eligible(rows, cin, ca) and tc_eligible(rows, ca, cb) and rows_per_sample % 32 == 0 and
            ca // groups >= 32 and (not stats_b or cb // groups >= 32))
    if fold:
        h, st = linear(x, conv_a.weight, conv_a.bias, engine='tc', out_stats=(samples, rows_per_sample, groups))
        pn = PendingNorm(st, samples, rows_per_sample, groups, ca, gn_a.weight, gn_a.bias, relu=True, eps=gn_a.eps)
        return linear(h, conv_b.weight, conv_b.bias, out=out, engine='tc', in_norm=pn,
                      out_stats=(samples, rows_per_sample, groups) if stats_b else None,
                      reduce_only=reduce_only and stats_b)
    h = linear(x, conv_a.weight, conv_a.bias)
    groupnorm(h, samples, rows_per_sample, groups, gn_a.weight, gn_a.bias, eps=gn_a.eps, relu=True)
    if stats_b and tc_eligible(rows, ca, cb) and rows_per_sample % 32 == 0 and cb // groups >= 32:
        # the first layer's groups are too narrow for the epilogue statistics, the second layer's are not
        return linear(h, conv_b.weight, conv_b.bias, out=None if reduce_only else out, engine='tc',
                      out_stats=(samples, rows_per_sample, groups), reduce_only=reduce_only)
    y = linear(h, conv_b.weight, conv_b.bias, out=out)
    return (y, None) if stats_b else y


def linear(x, weight, bias=None, out=None, act_in=ACT_NONE, act_out=ACT_NONE, engine='auto', in_norm=None,
           out_stats=None, bias_rows_per_sample=0, weight_key=None, reduce_only=False):
    """1x1 Conv1d / Linear on rows: y = act_out(act_in(x) @ W^T + b).

    x (rows, Cin) view; weight (Cout, Cin) or Conv1d-shaped (Cout, Cin, 1); out optional (rows, Cout) view.
    engine: 'auto' (tensor cores for large layers), 'tc' or 'simt'.
    Tensor-core engine only: in_norm = PendingNorm of x (normalise while splitting the operand);
    out_stats = (samples, rows_per_sample, groups): also return the fp64 GroupNorm statistics of y;
    bias_rows_per_sample > 0: `bias` is (rows / bias_rows_per_sample, Cout), one bias row per sample;
    weight_key: the parameter `weight` was derived from (its version keys the cache of split planes);
    reduce_only (with out_stats): y is not written; returns (None, (statistics, extrema)) where extrema holds the
    per-(sample, channel) max / min of y as ordered keys (`gn_max_from_extrema`)."""
    x, ldx = _rows2d(x, 'x')
    w = weight.reshape(weight.shape[0], weight.shape[1])
    _f32(w, 'weight')
    assert w.is_contiguous()
    rows, cin = x.shape
    cout = w.shape[0]
    assert w.shape[1] == cin
    if reduce_only:
        assert engine == 'tc' and out_stats is not None and out is None and act_out == ACT_NONE and cout % 4 == 0
        samples, rps, groups = out_stats
        stats_t = torch.empty(samples * groups * 2, dtype=torch.float64, device=x.device)
        ext_t = torch.empty(samples, 2, cout, dtype=torch.int32, device=x.device)
        stats_s = _lib.GnStats(stats_t.data_ptr(), rps, groups, ext_t.data_ptr())
        fold_s = None
        if in_norm is not None:
            fold_s = _lib.GnFold(in_norm.table().data_ptr(), in_norm.rows_per_sample, int(in_norm.relu))
        ws_bytes = lib.caspr_linear_tc_workspace_bytes(rows, cin, cout)
        ws, ws_ptr = _aligned_bytes(ws_bytes, x.device)
        prepared = _prepared_weights(weight, w, weight_key)
        _count('linear_tc')
        check(lib.caspr_linear_tc(_p(x), ldx, _p(w), cin, _p(bias), None, 0, rows, cin, cout, act_in, act_out,
                                  ctypes.c_void_p(prepared) if prepared else None,
                                  ctypes.byref(fold_s) if fold_s is not None else None, ctypes.byref(stats_s),
                                  int(bias_rows_per_sample), ctypes.c_void_p(ws_ptr), ws_bytes, _stream()),
              'caspr_linear_tc')
        return None, (stats_t, ext_t)
    if out is None:
        out = torch.empty(rows, cout, dtype=torch.float32, device=x.device)
    out, ldy = _rows2d(out, 'out')
    assert out.shape == (rows, cout)
    if engine == 'auto' and LINEAR_ENGINE != 'auto':
        engine = LINEAR_ENGINE if (LINEAR_ENGINE == 'simt' or cin >= 16) else 'simt'
    if engine == 'auto':
        engine = 'tc' if (rows >= TC_MIN_ROWS and cin >= TC_MIN_CIN and cout >= TC_MIN_COUT) else 'simt'
    if engine == 'tc' and (cout % 4 or ldy % 4 or out.data_ptr() % 16 or (bias is not None and bias.data_ptr() % 16)
                           or act_out == ACT_SIGMOID):
        if in_norm is not None or out_stats is not None or bias_rows_per_sample:
            raise ValueError('tensor-core linear needs Cout % 4 == 0 and 16-byte aligned output / bias rows')
        engine = 'simt'
    if engine != 'tc' and (in_norm is not None or out_stats is not None or bias_rows_per_sample):
        raise ValueError('GroupNorm folding / per-sample bias are implemented by the tensor-core linear only')
    if engine == 'tc':
        fold_s = stats_s = None
        stats_t = None
        if in_norm is not None:
            fold_s = _lib.GnFold(in_norm.table().data_ptr(), in_norm.rows_per_sample, int(in_norm.relu))
        if out_stats is not None:
            samples, rps, groups = out_stats
            stats_t = torch.empty(samples * groups * 2, dtype=torch.float64, device=x.device)
            stats_s = _lib.GnStats(stats_t.data_ptr(), rps, groups, None)
        ws_bytes = lib.caspr_linear_tc_workspace_bytes(rows, cin, cout)
        ws, ws_ptr = _aligned_bytes(ws_bytes, x.device)
        # weights are split once per (storage, in-place version); CUDA graphs that captured a call are keyed
        # by the same versions (TPointNet2._param_key), so a weight update re-captures with fresh planes
        prepared = _prepared_weights(weight, w, weight_key)
        _count('linear_tc')
        check(lib.caspr_linear_tc(_p(x), ldx, _p(w), cin, _p(bias), _p(out), ldy, rows, cin, cout, act_in, act_out,
                                  ctypes.c_void_p(prepared) if prepared else None,
                                  ctypes.byref(fold_s) if fold_s is not None else None,
                                  ctypes.byref(stats_s) if stats_s is not None else None, int(bias_rows_per_sample),
                                  ctypes.c_void_p(ws_ptr), ws_bytes, _stream()), 'caspr_linear_tc')
        return (out, stats_t) if out_stats is not None else out
    _count('linear')
    check(lib.caspr_linear(_p(x), ldx, _p(w), cin, _p(bias), _p(out), ldy, rows, cin, cout, act_in, act_out,
                           _stream()), 'caspr_linear')
    return out


def linear_gn_ball(x, weight, bias, gamma, beta, ns, relu, want_rows=True, maxout=None, eps=1e-5):
    """Fused per-ball layer: GroupNorm(16) over balls of `ns` rows (+ReLU) of x @ W^T + b; returns the
    normalised rows (want_rows) and/or writes the per-ball max into `maxout` (balls, Cout) view."""
    x, ldx = _rows2d(x, 'x')
    w = weight.reshape(weight.shape[0], weight.shape[1])
    rows, cin = x.shape
    cout = w.shape[0]
    y, ldy = None, 0
    if want_rows:
        y = torch.empty(rows, cout, dtype=torch.float32, device=x.device)
        ldy = cout
    ld_max = 0
    if maxout is not None:
        maxout, ld_max = _rows2d(maxout, 'maxout')
        assert maxout.shape == (rows // ns, cout)
    _count('linear_gn_ball')
    check(lib.caspr_linear_gn_ball(_p(x), ldx, _p(w), cin, _p(bias), _p(gamma), _p(beta), float(eps), rows, cin, cout,
                                   ns, int(relu), _p(y), ldy, _p(maxout), ld_max, _stream()), 'caspr_linear_gn_ball')
    return y


SA_MLP_TC = True                # module-wide switch: per-ball GroupNorm inside the tensor-core GEMM epilogues (SA 3-5)


def sa_mlp_tc_supported(ns, cin, widths, rows):
    return (SA_MLP_TC and LINEAR_ENGINE != 'simt' and len(widths) == 3 and rows >= TC_MIN_ROWS and
            bool(lib.caspr_sa_mlp_tc_supported(ns, cin, *widths)))


def sa_mlp_tc(rows, ns, convs, norms, out):
    """Per-ball MLP of a set-abstraction scale on the tensor cores: rows (balls*ns, Cin) grouped points ->
    out (balls, C3) view (may be a column slice).  Three [conv, per-ball GroupNorm(16)] layers, ReLU after the first two,
    max over the ball; the normalisation runs in the GEMM epilogues."""
    rows, ldx = _rows2d(rows, 'rows')
    n, cin = rows.shape
    out, ld_out = _rows2d(out, 'out')
    args = []
    for conv, gn in zip(convs, norms):
        w = conv.weight.reshape(conv.weight.shape[0], conv.weight.shape[1])
        assert w.is_contiguous() and w.dtype == torch.float32
        args += [ctypes.c_void_p(_prepared_weights(conv.weight, w)), _p(conv.bias), _p(gn.weight), _p(gn.bias),
                 w.shape[0]]
    c1, c2 = convs[0].weight.shape[0], convs[1].weight.shape[0]
    nb = lib.caspr_sa_mlp_tc_workspace_bytes(n, cin, c1, c2)
    ws, ws_ptr = _aligned_bytes(nb, rows.device)
    _count('sa_mlp_tc')
    check(lib.caspr_sa_mlp_tc(_p(rows), ldx, n, cin, ns, *args, float(norms[0].eps), _p(out), ld_out,
                              ctypes.c_void_p(ws_ptr), nb, _stream()), 'caspr_sa_mlp_tc')
    return out


def sa_mlp_tc_grouped(xyz, new_xyz, feat, idx, convs, norms, out):
    """`sa_mlp_tc` with the group gather folded into the operand split: xyz (B,N,3), new_xyz (B,M,3), feat (B,N,C) view,
    idx (B,M,ns) -> out (B*M, C3) view.  The grouped rows never exist in memory."""
    B, N, _ = xyz.shape
    M, ns = idx.shape[1], idx.shape[2]
    assert feat.dim() == 3 and feat.stride(2) == 1 and feat.stride(0) == N * feat.stride(1)
    assert xyz.is_contiguous() and new_xyz.is_contiguous() and idx.is_contiguous()
    C, ld_feat = feat.shape[2], feat.stride(1)
    out, ld_out = _rows2d(out, 'out')
    args = []
    for conv, gn in zip(convs, norms):
        w = conv.weight.reshape(conv.weight.shape[0], conv.weight.shape[1])
        assert w.is_contiguous() and w.dtype == torch.float32
        args += [ctypes.c_void_p(_prepared_weights(conv.weight, w)), _p(conv.bias), _p(gn.weight), _p(gn.bias),
                 w.shape[0]]
    c1, c2 = convs[0].weight.shape[0], convs[1].weight.shape[0]
    nb = lib.caspr_sa_mlp_tc_workspace_bytes(B * M * ns, 3 + C, c1, c2)
    ws, ws_ptr = _aligned_bytes(nb, xyz.device)
    _count('sa_mlp_tc_grouped')
    check(lib.caspr_sa_mlp_tc_grouped(_p(xyz), _p(new_xyz), _p(feat), ld_feat, C, _p(idx), B, N, M, ns, *args,
                                      float(norms[0].eps), _p(out), ld_out, ctypes.c_void_p(ws_ptr), nb, _stream()),
          'caspr_sa_mlp_tc_grouped')
    return out


def sa_mlp_tc_delayed(xyz, new_xyz, feat, idx, convs, norms, out):
    """`sa_mlp_tc_grouped` with the first layer's product taken before the gather: P = feat . W1[:, 3:]^T once per
    source point, then one kernel gathers P, adds W1[:, :3] . (xyz - centre) + b1 and normalises per ball."""
    B, N, _ = xyz.shape
    M, ns = idx.shape[1], idx.shape[2]
    assert feat.dim() == 3 and feat.stride(2) == 1 and feat.stride(0) == N * feat.stride(1)
    assert xyz.is_contiguous() and new_xyz.is_contiguous() and idx.is_contiguous()
    C = feat.shape[2]
    w1 = convs[0].weight
    c1, c2 = w1.shape[0], convs[1].weight.shape[0]
    w1_2d = w1.reshape(c1, 3 + C)
    assert w1_2d.is_contiguous() and w1.dtype == torch.float32
    w_feat = derived_weight(w1, 'sa_feature_columns', lambda w: w.reshape(w.shape[0], -1)[:, 3:])
    prod = linear(feat.reshape(B * N, C) if feat.is_contiguous() else feat.flatten(0, 1), w_feat, None, weight_key=w1)
    out, ld_out = _rows2d(out, 'out')
    args = [_p(w1_2d), 3 + C, _p(convs[0].bias), _p(norms[0].weight), _p(norms[0].bias), c1]
    for conv, gn in zip(convs[1:], norms[1:]):
        w = conv.weight.reshape(conv.weight.shape[0], conv.weight.shape[1])
        assert w.is_contiguous() and w.dtype == torch.float32
        args += [ctypes.c_void_p(_prepared_weights(conv.weight, w)), _p(conv.bias), _p(gn.weight), _p(gn.bias),
                 w.shape[0]]
    nb = lib.caspr_sa_mlp_tc_delayed_workspace_bytes(B * M * ns, c1, c2)
    ws, ws_ptr = _aligned_bytes(nb, xyz.device)
    _count('sa_mlp_tc_delayed')
    check(lib.caspr_sa_mlp_tc_delayed(_p(xyz), _p(new_xyz), _p(prod), prod.stride(0), _p(idx), B, N, M, ns, *args,
                                      float(norms[0].eps), _p(out), ld_out, ctypes.c_void_p(ws_ptr), nb, _stream()),
          'caspr_sa_mlp_tc_delayed')
    return out


SA_FUSED = True                 # module-wide switch (accuracy / timing studies): fused set-abstraction scale kernel
SA_MMA = os.environ.get('CASPR_SA_MMA', '1') != '0'     # its tensor-core (mma.sync) version, SA levels 1-2


def sa_fused_supported(ns, cin, widths):
    return SA_FUSED and len(widths) == 3 and bool(lib.caspr_sa_fused_supported(ns, cin, *widths))


def sa_fused(xyz, new_xyz, feat, idx, convs, norms, out):
    """One scale of a set-abstraction level in one kernel: gather -> 3 x [conv, per-ball GroupNorm(16), ReLU (not
    after the last)] -> max over the ball.  xyz (B,N,3), new_xyz (B,M,3), feat (B,N,C) view or None, idx (B,M,ns),
    out (B*M, C3) view (may be a column slice)."""
    B, N, _ = xyz.shape
    M, ns = idx.shape[1], idx.shape[2]
    C, ld_feat = 0, 0
    if feat is not None:
        assert feat.dim() == 3 and feat.stride(2) == 1 and feat.stride(0) == N * feat.stride(1)
        C, ld_feat = feat.shape[2], feat.stride(1)
    out, ld_out = _rows2d(out, 'out')
    args = []
    for conv, gn in zip(convs, norms):
        w = conv.weight.reshape(conv.weight.shape[0], conv.weight.shape[1])
        assert w.is_contiguous() and w.dtype == torch.float32
        args += [_p(w), _p(conv.bias), _p(gn.weight), _p(gn.bias), w.shape[0]]
    _count('sa_fused')
    check(lib.caspr_sa_fused(_p(xyz), _p(new_xyz), _p(feat), ld_feat, C, _p(idx), B, N, M, ns, *args,
                             float(norms[0].eps), _p(out), ld_out, _stream()), 'caspr_sa_fused')
    return out


def sa_mma_supported(ns, cin, widths, feat=None):
    """feat: the (B,N,C) feature view, if known - the kernel addresses rows with 32-bit offsets."""
    if feat is not None and feat.shape[0] * feat.shape[1] * max(feat.stride(1), 3) >= 2 ** 32:
        return False
    return SA_MMA and len(widths) == 3 and bool(lib.caspr_sa_mma_supported(ns, cin, *widths))


def sa_absmax(xyz, feat):
    """Device scalar max(|feat|, 2 max|xyz|): an upper bound of every entry of the gathered rows of a set-abstraction
    level, shared by its two scales (operand scale of `sa_mma`)."""
    B, N, _ = xyz.shape
    assert xyz.is_contiguous() and feat.dim() == 3 and feat.stride(2) == 1 and feat.stride(0) == N * feat.stride(1)
    out = torch.empty(1, dtype=torch.float32, device=xyz.device)
    _count('sa_absmax')
    check(lib.caspr_sa_absmax(_p(xyz), _p(feat), feat.stride(1), feat.shape[2], B, N, _p(out), _stream()),
          'caspr_sa_absmax')
    return out


def sa_mma(xyz, new_xyz, feat, idx, convs, norms, out, absmax=None):
    """`sa_fused` with the three per-ball layers on the tensor cores (mma.sync, fp16x3 split); same arguments plus
    the operand bound of `sa_absmax`."""
    B, N, _ = xyz.shape
    M, ns = idx.shape[1], idx.shape[2]
    assert feat.dim() == 3 and feat.stride(2) == 1 and feat.stride(0) == N * feat.stride(1)
    C, ld_feat = feat.shape[2], feat.stride(1)
    out, ld_out = _rows2d(out, 'out')
    args = []
    for conv, gn in zip(convs, norms):
        w = conv.weight.reshape(conv.weight.shape[0], conv.weight.shape[1])
        assert w.is_contiguous() and w.dtype == torch.float32
        args += [_p(w), _p(conv.bias), _p(gn.weight), _p(gn.bias), w.shape[0]]
    _count('sa_mma')
    check(lib.caspr_sa_mma(_p(xyz), _p(new_xyz), _p(feat), ld_feat, C, _p(idx), B, N, M, ns, *args,
                           float(norms[0].eps), _p(absmax), _p(out), ld_out, _stream()), 'caspr_sa_mma')
    return out


def groupnorm(x, samples, rows_per_sample, groups, gamma, beta, eps=1e-5, relu=False, write_back=True,
              maxout=None, stats=None):
    """In-place GroupNorm(groups, C) over `samples` blocks of consecutive rows, fused ReLU / max-pool.

    x (samples*rows_per_sample, C) view.  maxout: optional (samples, C) view receiving the max over
    the rows of each sample of the normalised (and ReLU'd) values.  stats: fp64 (sum, sum of squares)
    already accumulated by the producing GEMM (skips the statistics pass)."""
    x, ldx = _rows2d(x, 'x')
    C = x.shape[1]
    assert x.shape[0] == samples * rows_per_sample
    ld_max = 0
    if maxout is not None:
        maxout, ld_max = _rows2d(maxout, 'maxout')
        assert maxout.shape == (samples, C)
    ready = stats is not None
    if stats is None:
        stats = torch.empty(samples * groups * 2, dtype=torch.float64, device=x.device)
    _count('groupnorm')
    check(lib.caspr_groupnorm(_p(x), ldx, samples, rows_per_sample, C, groups, _p(gamma), _p(beta), float(eps),
                              int(relu), int(write_back), _p(maxout), ld_max, _p(stats), int(ready), _stream()),
          'caspr_groupnorm')
    return x


def groupnorm_project(x, samples, rows_per_sample, groups, gamma, beta, stats, weight, bias, eps=1e-5, maxout=None,
                      act=ACT_NONE):
    """GroupNorm (statistics from the producing GEMM) -> max over each sample's rows (pre-ReLU) and
    act(weight . relu(gn(x)) + bias) for a layer of <= 4 outputs, in one pass that leaves x untouched
    (tpointnet2.py:104-113: bn2, the max-pool that yields z0, conv3 + sigmoid).  Returns (rows, P)."""
    x, ldx = _rows2d(x, 'x')
    C = x.shape[1]
    assert x.shape[0] == samples * rows_per_sample
    w2d = weight.reshape(weight.shape[0], -1)
    _f32(w2d, 'weight')
    assert w2d.shape[1] == C and w2d.is_contiguous() and 1 <= w2d.shape[0] <= 4
    ld_max = 0
    if maxout is not None:
        maxout, ld_max = _rows2d(maxout, 'maxout')
        assert maxout.shape == (samples, C)
    out = torch.empty(x.shape[0], w2d.shape[0], dtype=torch.float32, device=x.device)
    _count('groupnorm_project')
    check(lib.caspr_groupnorm_project(_p(x), ldx, samples, rows_per_sample, C, groups, _p(gamma), _p(beta), float(eps),
                                      _p(maxout), ld_max, _p(stats), _p(w2d), _p(bias), w2d.shape[0], int(act),
                                      _p(out), out.shape[1], _stream()), 'caspr_groupnorm_project')
    return out


def gn_max_from_extrema(stats_ext, samples, rows_per_sample, groups, gamma, beta, maxout, eps=1e-5):
    """max over each sample's rows of GroupNorm(y) from the statistics and per-channel extrema a `linear(...,
    reduce_only=True)` call left behind (the normalisation is monotone per channel): y never existed in memory."""
    stats, ext = stats_ext
    C = ext.shape[2]
    maxout, ld_max = _rows2d(maxout, 'maxout')
    assert maxout.shape == (samples, C)
    _count('gn_max_from_extrema')
    check(lib.caspr_gn_max_from_extrema(_p(stats), _p(ext), samples, rows_per_sample, C, groups, _p(gamma), _p(beta),
                                        float(eps), _p(maxout), ld_max, _stream()), 'caspr_gn_max_from_extrema')
    return maxout


def augment_xyz(x4):
    """tpointnet2.py:79-90 with both augmentations: (R,4) -> (R,9) [x,y,z,x2,y2,z2,xz,xy,yz]."""
    _f32(x4, 'x4')
    x4 = x4.contiguous()
    out = torch.empty(x4.shape[0], 9, dtype=torch.float32, device=x4.device)
    _count('augment_xyz')
    check(lib.caspr_augment_xyz(_p(x4), x4.shape[0], _p(out), _stream()), 'caspr_augment_xyz')
    return out


def strip_time(x4):
    _f32(x4, 'x4')
    x4 = x4.contiguous()
    out = torch.empty(x4.shape[0], 3, dtype=torch.float32, device=x4.device)
    _count('strip_time')
    check(lib.caspr_strip_time(_p(x4), x4.shape[0], _p(out), _stream()), 'caspr_strip_time')
    return out


def broadcast_rows(src, rows_per_sample, dst):
    """dst[s*rows_per_sample + r, :] = src[s, :] (pointnet.py:44-46 repeat + concat)."""
    src, ld_src = _rows2d(src, 'src')
    dst, ld_dst = _rows2d(dst, 'dst')
    samples, C = src.shape
    assert dst.shape == (samples * rows_per_sample, C)
    _count('broadcast_rows')
    check(lib.caspr_broadcast_rows(_p(src), ld_src, samples, rows_per_sample, C, _p(dst), ld_dst, _stream()),
          'caspr_broadcast_rows')
    return dst


# ------------------------------------------------------------------------------ latent ODE
LATENT_MAX_TIMES = 64                    # csrc/latent_ode.cu: kMaxTimes
LATENT_MAX_STAGING_BYTES = 200 * 1024    # B * max(H, D) floats of staging per CTA


def latent_ode_solve(z0, weights, biases, times, rtol, atol):
    """dopri5 solve of the latent dynamics MLP.  z0 (B,D); times: increasing python floats /
    1-D tensor (times[0] = start).  Returns out (nT,B,D) and the info list [status,nfe,acc,rej,...]."""
    _f32(z0, 'z0')
    z0 = z0.contiguous()
    B, D = z0.shape
    H = weights[0].shape[0]
    tl = [float(t) for t in times]
    nT = len(tl)
    # limits of the cooperative solver (csrc/latent_ode.cu): the output-time table lives in shared memory
    if nT > LATENT_MAX_TIMES:
        raise ValueError('caspr_latent_ode_solve handles at most %d distinct time stamps per call, got %d '
                         '(reconstruct(timestamps=...) with more steps: query them in several calls)'
                         % (LATENT_MAX_TIMES, nT))
    if B * max(H, D) * 4 > LATENT_MAX_STAGING_BYTES:
        raise ValueError('caspr_latent_ode_solve stages B x max(H, D) floats in shared memory (<= %d KB): B = %d is too '
                         'large; split the batch' % (LATENT_MAX_STAGING_BYTES // 1024, B))
    h_times = (ctypes.c_double * nT)(*tl)
    out = torch.empty(nT, B, D, dtype=torch.float32, device=z0.device)
    info = torch.zeros(8, dtype=torch.int32, device=z0.device)
    h_info = (ctypes.c_int32 * 8)()
    ws_bytes = lib.caspr_latent_ode_workspace_bytes(B, D, H)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=z0.device)
    w = [x.contiguous() for x in weights]
    b = [x.contiguous() for x in biases]
    _count('latent_ode_solve')
    rc = lib.caspr_latent_ode_solve(_p(z0), B, D, H, _p(w[0]), _p(b[0]), _p(w[1]), _p(b[1]), _p(w[2]), _p(b[2]),
                                    _p(w[3]), _p(b[3]), h_times, nT, float(rtol), float(atol), _p(out), _p(info),
                                    h_info, _p(ws), ws_bytes, _stream())
    return out, list(h_info), rc


def latent_ode_adjoint(zs, gzs, weights, biases, times, rtol, atol):
    """Adjoint backward of latent_ode_solve: zs, gzs (nT,B,D) -> gz0 (B,D), gparams (flat: W0,b0,...,W3,b3)."""
    zs, gzs = _f32(zs, 'zs').contiguous(), _f32(gzs, 'gzs').contiguous()
    nT, B, D = zs.shape
    H = weights[0].shape[0]
    tl = [float(t) for t in times]
    assert len(tl) == nT
    h_times = (ctypes.c_double * nT)(*tl)
    gz0 = torch.empty(B, D, dtype=torch.float32, device=zs.device)
    gparams = torch.empty(int(lib.caspr_latent_ode_param_count(D, H)), dtype=torch.float32, device=zs.device)
    info = torch.zeros(8, dtype=torch.int32, device=zs.device)
    h_info = (ctypes.c_int32 * 8)()
    ws_bytes = lib.caspr_latent_ode_adjoint_workspace_bytes(B, D, H)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=zs.device)
    w = [x.detach().contiguous() for x in weights]
    b = [x.detach().contiguous() for x in biases]
    _count('latent_ode_adjoint')
    rc = lib.caspr_latent_ode_adjoint(_p(zs), _p(gzs), B, D, H, _p(w[0]), _p(b[0]), _p(w[1]), _p(b[1]), _p(w[2]),
                                      _p(b[2]), _p(w[3]), _p(b[3]), h_times, nT, float(rtol), float(atol),
                                      _p(gz0), _p(gparams), _p(info), h_info, _p(ws), ws_bytes, _stream())
    return gz0, gparams, list(h_info), rc


# -------------------------------------------------------------------------------------- CNF
class CnfWeightPack(object):
    """Keeps the tensors referenced by a caspr_cnf_weights struct alive."""

    def __init__(self, layers, hidden, ctx_dim):
        # layers: list of 4 dicts with W, b, Wgate, bgate, Wbias (contiguous CUDA fp32 tensors)
        self.keep = layers
        self.struct = _lib.CnfWeights()
        for l, d in enumerate(layers):
            for k in ('W', 'b', 'Wgate', 'bgate', 'Wbias'):
                t = d[k]
                assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
                getattr(self.struct, k)[l] = t.data_ptr()
        self.struct.hidden = hidden
        self.struct.ctx_dim = ctx_dim


def _mbn_struct(m):
    if m is None:
        return None, None
    keep = [m[k].contiguous() for k in ('weight', 'bias', 'running_mean', 'running_var')]
    s = _lib.MbnParams(*[t.data_ptr() for t in keep])
    return s, keep


class LockstepSync(object):
    """Step-control synchronisation over the ranks of `group` for caspr_cnf_flow_lockstep: holds the staging buffer
    and the ctypes callback that all-reduces it with torch.distributed on the current stream."""

    def __init__(self, n_global, device, group=None):
        import torch.distributed as dist
        self.stage = torch.zeros(2, dtype=torch.float64, device=device)
        self.calls = 0

        def _allreduce(ptr, count, user, stream):
            try:
                assert ptr == self.stage.data_ptr() and count == 2
                dist.all_reduce(self.stage, op=dist.ReduceOp.SUM, group=group)
                self.calls += 1
                return 0
            except Exception:            # never unwind through the C frames
                return 1
        self._cb = _lib.ALLREDUCE_FN(_allreduce)
        self.struct = _lib.CnfSync(int(n_global), self.stage.data_ptr(), self._cb, None)


def cnf_flow(x, logp, e, ctx, pack, mbn0, mbn2, end_time, reverse, rtol=1e-5, atol=1e-5, engine=CNF_SIMT_FP32,
             sync=None):
    """One pass through [MBN, CNF, MBN] (or its inverse).  x (F,P,3), logp (F,P) or None, e (F,P,3),
    ctx (F,ctx_dim).  Returns x_out, logp_out (or None), info list, status.  sync: LockstepSync or None."""
    x, e, ctx = _f32(x, 'x').contiguous(), _f32(e, 'e').contiguous(), _f32(ctx, 'ctx').contiguous()
    F, P, _ = x.shape
    if logp is not None:
        logp = _f32(logp, 'logp').contiguous()
    x_out = torch.empty_like(x)
    logp_out = torch.empty(F, P, dtype=torch.float32, device=x.device) if logp is not None else None
    info = torch.zeros(8, dtype=torch.int32, device=x.device)
    h_info = (ctypes.c_int32 * 8)()
    ws_bytes = lib.caspr_cnf_workspace_bytes(F, P, pack.struct.hidden, pack.struct.ctx_dim, engine)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    s0, k0 = _mbn_struct(mbn0)
    s2, k2 = _mbn_struct(mbn2)
    _count('cnf_flow')
    args = (_p(x), _p(logp), _p(e), _p(ctx), F, P, ctypes.byref(pack.struct),
            ctypes.byref(s0) if s0 is not None else None, ctypes.byref(s2) if s2 is not None else None,
            float(end_time), int(bool(reverse)), float(rtol), float(atol), int(engine),
            _p(x_out), _p(logp_out), _p(info), h_info, _p(ws), ws_bytes, _stream())
    if sync is not None:
        rc = lib.caspr_cnf_flow_lockstep(*args, ctypes.byref(sync.struct))
    else:
        rc = lib.caspr_cnf_flow(*args)
    del k0, k2
    return x_out, logp_out, list(h_info), rc


def cnf_feval(y, e, ctx, pack, t, engine=CNF_SIMT_FP32):
    """One dynamics evaluation: returns dy (F,P,3), neg_div (F,P)."""
    y, e, ctx = _f32(y, 'y').contiguous(), _f32(e, 'e').contiguous(), _f32(ctx, 'ctx').contiguous()
    F, P, _ = y.shape
    dy = torch.empty_like(y)
    nd = torch.empty(F, P, dtype=torch.float32, device=y.device)
    ws_bytes = lib.caspr_cnf_workspace_bytes(F, P, pack.struct.hidden, pack.struct.ctx_dim, engine)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=y.device)
    _count('cnf_feval')
    check(lib.caspr_cnf_feval(_p(y), _p(e), _p(ctx), F, P, ctypes.byref(pack.struct), float(t), int(engine),
                              _p(dy), _p(nd), _p(ws), ws_bytes, _stream()), 'caspr_cnf_feval')
    return dy, nd


def cnf_param_count(pack):
    return int(lib.caspr_cnf_param_count(pack.struct.hidden, pack.struct.ctx_dim))


def cnf_adjoint(x1, logp1, gx1, glogp1, e, ctx, pack, end_time, rtol=1e-5, atol=1e-5, engine=CNF_SIMT_FP32):
    """Adjoint backward of the CNF block (forward direction).  x1 (F,P,3), logp1 (F,P): block outputs at t1;
    gx1, glogp1: their gradients.  Returns gx0 (F,P,3), glogp0 (F,P), gctx (F,ctx), gparams (flat, ODEfunc
    parameters() order), gtimes (2,), info list, status."""
    x1, gx1, e, ctx = [_f32(t, n).contiguous() for t, n in ((x1, 'x1'), (gx1, 'gx1'), (e, 'e'), (ctx, 'ctx'))]
    logp1, glogp1 = _f32(logp1, 'logp1').contiguous(), _f32(glogp1, 'glogp1').contiguous()
    F, P, _ = x1.shape
    dev = x1.device
    H, C = pack.struct.hidden, pack.struct.ctx_dim
    gx0 = torch.empty_like(x1)
    glogp0 = torch.empty(F, P, dtype=torch.float32, device=dev)
    gctx = torch.empty(F, C, dtype=torch.float32, device=dev)
    gparams = torch.empty(cnf_param_count(pack), dtype=torch.float32, device=dev)
    gtimes = torch.zeros(2, dtype=torch.float32, device=dev)
    info = torch.zeros(8, dtype=torch.int32, device=dev)
    h_info = (ctypes.c_int32 * 8)()
    ws_bytes = lib.caspr_cnf_adjoint_workspace_bytes(F, P, H, C)
    ws, ws_ptr = _aligned_bytes(ws_bytes, dev)
    _count('cnf_adjoint')
    rc = lib.caspr_cnf_adjoint(_p(x1), _p(logp1), _p(gx1), _p(glogp1), _p(e), _p(ctx), F, P,
                               ctypes.byref(pack.struct), float(end_time), float(rtol), float(atol), int(engine),
                               _p(gx0), _p(glogp0), _p(gctx), _p(gparams), _p(gtimes), _p(info), h_info,
                               ctypes.c_void_p(ws_ptr), ws_bytes, _stream())
    return gx0, glogp0, gctx, gparams, gtimes, list(h_info), rc


def chamfer(a, b):
    """Squared-NN distances both ways: a (B,P,3), b (B,Q,3) -> d_ab (B,P), d_ba (B,Q)."""
    a, b = _f32(a, 'a').contiguous(), _f32(b, 'b').contiguous()
    B, P, _ = a.shape
    Q = b.shape[1]
    d_ab = torch.empty(B, P, dtype=torch.float32, device=a.device)
    d_ba = torch.empty(B, Q, dtype=torch.float32, device=a.device)
    _count('chamfer')
    check(lib.caspr_chamfer(_p(a), _p(b), B, P, Q, _p(d_ab), _p(d_ba), _stream()), 'caspr_chamfer')
    return d_ab, d_ba


def emd(a, b):
    """Approximate earth mover's distance (utils/emd.py earth_mover_distance with transpose=False): a (B,n,3),
    b (B,m,3) -> cost (B,); evaluations.py:45-46 divides it by the number of points."""
    a, b = _f32(a, 'a').contiguous(), _f32(b, 'b').contiguous()
    B, n, _ = a.shape
    m = b.shape[1]
    cost = torch.empty(B, dtype=torch.float32, device=a.device)
    nb = lib.caspr_emd_workspace_bytes(B, n, m)
    ws = torch.empty(nb, dtype=torch.uint8, device=a.device)
    _count('emd')
    check(lib.caspr_emd(_p(a), _p(b), B, n, m, _p(cost), _p(ws), nb, _stream()), 'caspr_emd')
    return cost


def tnocs_error(pred, gt):
    """evaluations.py:243-254: pred, gt (B,T,N,4) -> (space (B,T), time (B,T)): per-frame mean L2 position error and
    mean absolute time error of the T-NOCS regression."""
    pred, gt = _f32(pred, 'pred').contiguous(), _f32(gt, 'gt').contiguous()
    B, T, N, four = pred.shape
    assert four == 4 and gt.shape == pred.shape
    space = torch.empty(B, T, dtype=torch.float32, device=pred.device)
    terr = torch.empty(B, T, dtype=torch.float32, device=pred.device)
    _count('tnocs_error')
    check(lib.caspr_tnocs_error(_p(pred), _p(gt), B * T, N, _p(space), _p(terr), _stream()), 'caspr_tnocs_error')
    return space, terr


def ransac_pose(src, dst, samples, max_distance=0.015, refine=False, want_counts=False):
    """Correspondence-RANSAC rigid pose (evaluations.py:360-380): src, dst (F,N,3) corresponding points, samples
    (F,H,4) int32 correspondences of the H hypotheses -> dict with R (F,3,3), t (F,3), best (F,) int32, fitness (F,),
    inlier_rmse (F,) and, on request, counts (F,H)."""
    src, dst = _f32(src, 'src').contiguous(), _f32(dst, 'dst').contiguous()
    F, N, _ = src.shape
    assert dst.shape == src.shape and samples.dim() == 3 and samples.shape[0] == F and samples.shape[2] == 4
    samples = samples.to(device=src.device, dtype=torch.int32).contiguous()
    H = samples.shape[1]
    dev = src.device
    R = torch.empty(F, 3, 3, dtype=torch.float32, device=dev)
    t = torch.empty(F, 3, dtype=torch.float32, device=dev)
    best = torch.empty(F, dtype=torch.int32, device=dev)
    fit = torch.empty(F, dtype=torch.float32, device=dev)
    rmse = torch.empty(F, dtype=torch.float32, device=dev)
    counts = torch.empty(F, H, dtype=torch.int32, device=dev) if want_counts else None
    nb = lib.caspr_ransac_pose_workspace_bytes(F, H)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    _count('ransac_pose')
    check(lib.caspr_ransac_pose(_p(src), _p(dst), _p(samples), F, N, H, float(max_distance), int(bool(refine)), _p(R),
                                _p(t), _p(best), _p(fit), _p(rmse), _p(counts) if want_counts else None, _p(ws), nb,
                                _stream()), 'caspr_ransac_pose')
    out = {'R': R, 't': t, 'best': best, 'fitness': fit, 'inlier_rmse': rmse}
    if want_counts:
        out['counts'] = counts
    return out
