"""Oracle restatement of the Kaolin PointNet++ ops the reference imports.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference binds these names at ``caspr/models/pointnet2.py:7-10``::

    from kaolin.models.PointNet2 import separate_xyz_and_features,
        PointNet2GroupingLayer, furthest_point_sampling, fps_gather_by_index,
        three_nn, three_interpolate

Kaolin itself is NOT under ``/root/reference`` (un-vendored, unpinned git
master of 2020; its PointNet2 CUDA ops derive from
erikwijmans/Pointnet2_PyTorch).  What follows restates the published algorithm
of those kernels on the CPU with a DECLARED CANONICAL ARITHMETIC, because the
real kernels' results depend on nvcc's FMA contraction and on the thread layout
of their tree reductions:

* squared distance  d2 = ((dx*dx) + (dy*dy)) + (dz*dz)   in fp32, each product
  and sum rounded separately (no FMA contraction), evaluated left to right;
* arg-max / arg-min ties resolve to the LOWEST point index;
* FPS keeps upstream's quirk of ignoring points with |p|^2 <= 1e-3.

"Bit-exact FPS / ball-query indices" in this repository therefore means: equal
to THIS restatement on the same inputs.

Call sites in the reference: ``pointnet2.py:228`` (separate), ``:384`` (FPS),
``:385`` (gather), ``:340,:391`` (grouping layer = ball query + group gather),
``:514`` (three_nn), ``:519`` (three_interpolate).
"""
import numpy as np
import torch

FPS_ORIGIN_SKIP_MAG = np.float32(1e-3)   # upstream: `if (mag <= 1e-3) continue;`
FPS_INIT_DIST = np.float32(1e10)


def _sqdist_f32(ax, ay, az, bx, by, bz):
    """Canonical fp32 squared distance, no contraction: ((dx*dx)+(dy*dy))+(dz*dz)."""
    dx = (ax - bx).astype(np.float32)
    dy = (ay - by).astype(np.float32)
    dz = (az - bz).astype(np.float32)
    return ((dx * dx).astype(np.float32) + (dy * dy).astype(np.float32)).astype(np.float32) \
        + (dz * dz).astype(np.float32)


def separate_xyz_and_features(points):
    """(B,N,3+C) -> xyz (B,N,3) contiguous, features (B,C,N) contiguous or None.

    Reference call site: pointnet2.py:228.
    """
    assert points.dim() == 3 and points.shape[2] >= 3
    xyz = points[..., 0:3].contiguous()
    features = points[..., 3:].transpose(1, 2).contiguous() if points.shape[2] > 3 else None
    return xyz, features


def furthest_point_sampling_np(xyz, num_points_out):
    """xyz (B,N,3) float32 ndarray -> idx (B,M) int32.  Call site pointnet2.py:384.

    Upstream kernel (one CTA per cloud): temp[k]=1e10; idx[0]=0; then M-1 times:
    for every point with |p|^2 > 1e-3: temp[k]=min(temp[k], |p_k-p_last|^2), and
    the next pick is arg-max temp (strict '>' from best=-1, besti=0).
    """
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    B, N, _ = xyz.shape
    M = int(num_points_out)
    x, y, z = xyz[..., 0], xyz[..., 1], xyz[..., 2]
    mag = ((x * x).astype(np.float32) + (y * y).astype(np.float32)).astype(np.float32) \
        + (z * z).astype(np.float32)
    skip = mag <= FPS_ORIGIN_SKIP_MAG
    temp = np.full((B, N), FPS_INIT_DIST, dtype=np.float32)
    idx = np.zeros((B, M), dtype=np.int32)
    last = np.zeros((B,), dtype=np.int64)
    ar = np.arange(B)
    neg = np.float32(-1.0)
    for j in range(1, M):
        lx = x[ar, last][:, None]
        ly = y[ar, last][:, None]
        lz = z[ar, last][:, None]
        d = _sqdist_f32(x, y, z, lx, ly, lz)
        d2 = np.minimum(d, temp)
        temp = np.where(skip, temp, d2)
        cand = np.where(skip, neg, d2)
        # np.argmax returns the first (lowest-index) maximum; all-skipped -> 0
        last = np.argmax(cand, axis=1)
        idx[:, j] = last
    return idx


def furthest_point_sampling(xyz, num_points_out):
    """torch wrapper: (B,N,3) float tensor -> (B,M) int32 tensor (non-differentiable)."""
    out = furthest_point_sampling_np(xyz.detach().cpu().numpy(), num_points_out)
    return torch.from_numpy(out).to(xyz.device)


def fps_gather_by_index(features, idx):
    """features (B,C,N), idx (B,M) int -> (B,C,M).  Call site pointnet2.py:385."""
    B, C, _ = features.shape
    index = idx.long().unsqueeze(1).expand(B, C, idx.shape[1])
    return torch.gather(features, 2, index)


def ball_query_np(radius, num_samples, xyz, new_xyz):
    """idx (B,M,ns) int32: the first `ns` indices k (ascending) with d2 < r^2.

    Upstream kernel: slots are zero-initialised; on the FIRST hit every slot is
    filled with that index; later hits overwrite slot cnt++; stop at ns.
    r^2 is formed in fp32.  Part of the grouping layer, pointnet2.py:340,:391.
    """
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    new_xyz = np.ascontiguousarray(new_xyz, dtype=np.float32)
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    ns = int(num_samples)
    r = np.float32(radius)
    r2 = np.float32(r * r)
    out = np.zeros((B, M, ns), dtype=np.int32)
    slot = np.arange(ns)[None, :]
    for b in range(B):
        d2 = _sqdist_f32(new_xyz[b, :, None, 0], new_xyz[b, :, None, 1], new_xyz[b, :, None, 2],
                         xyz[b, None, :, 0], xyz[b, None, :, 1], xyz[b, None, :, 2])
        hit = d2 < r2                                            # (M,N)
        cnt = hit.sum(axis=1)                                    # (M,)
        # stable sort puts the hits first, in ascending index order
        order = np.argsort(~hit, axis=1, kind='stable')[:, :ns]  # (M,<=ns)
        if order.shape[1] < ns:
            order = np.concatenate(
                [order, np.zeros((M, ns - order.shape[1]), dtype=order.dtype)], axis=1)
        first = np.where(cnt > 0, order[:, 0], 0)[:, None]
        filled = slot < np.minimum(cnt, ns)[:, None]
        out[b] = np.where(filled, order, first).astype(np.int32)
    return out


def ball_query(radius, num_samples, xyz, new_xyz):
    out = ball_query_np(radius, num_samples, xyz.detach().cpu().numpy(),
                        new_xyz.detach().cpu().numpy())
    return torch.from_numpy(out).to(xyz.device)


def group_gather_by_index(features, idx):
    """features (B,C,N), idx (B,M,ns) -> (B,C,M,ns)."""
    B, C, _ = features.shape
    _, M, ns = idx.shape
    index = idx.long().reshape(B, 1, M * ns).expand(B, C, M * ns)
    return torch.gather(features, 2, index).reshape(B, C, M, ns)


class PointNet2GroupingLayer(torch.nn.Module):
    """Ball query + grouping; returns (B, M, 3+C, ns) with channels [dxyz | features].

    Output layout per the reference's own comment (pointnet2.py:392) and the
    `.view(-1, in_channels, num_samples)` that follows (pointnet2.py:397).
    Constructed at pointnet2.py:340-342; has no parameters.
    """

    def __init__(self, radius, num_samples, use_xyz_feature=True, use_random_ball_query=False):
        super().__init__()
        assert not use_random_ball_query, 'reference always passes use_random_ball_query=False'
        self.radius = radius
        self.num_samples = num_samples
        self.use_xyz_feature = use_xyz_feature

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query(self.radius, self.num_samples, xyz, new_xyz)
        grouped_xyz = group_gather_by_index(xyz.transpose(1, 2).contiguous(), idx)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is not None:
            grouped = group_gather_by_index(features, idx)
            if self.use_xyz_feature:
                grouped = torch.cat([grouped_xyz, grouped], dim=1)
        else:
            grouped = grouped_xyz
        return grouped.permute(0, 2, 1, 3).contiguous()


def three_nn_np(unknown, known):
    """For each unknown point the 3 nearest known points.

    Upstream kernel keeps a running best-3 of SQUARED distances with strict '<'
    while scanning k ascending (ties keep the earlier index); the Python wrapper
    returns sqrt(dist2).  Call site pointnet2.py:514.
    Returns dist (B,n,3) float32 Euclidean, idx (B,n,3) int32.
    """
    unknown = np.ascontiguousarray(unknown, dtype=np.float32)
    known = np.ascontiguousarray(known, dtype=np.float32)
    B, n, _ = unknown.shape
    dist = np.zeros((B, n, 3), dtype=np.float32)
    idx = np.zeros((B, n, 3), dtype=np.int32)
    for b in range(B):
        d2 = _sqdist_f32(unknown[b, :, None, 0], unknown[b, :, None, 1], unknown[b, :, None, 2],
                         known[b, None, :, 0], known[b, None, :, 1], known[b, None, :, 2])
        order = np.argsort(d2, axis=1, kind='stable')[:, :3]
        idx[b] = order.astype(np.int32)
        dist[b] = np.sqrt(np.take_along_axis(d2, order, axis=1)).astype(np.float32)
    return dist, idx


def three_nn(unknown, known):
    dist, idx = three_nn_np(unknown.detach().cpu().numpy(), known.detach().cpu().numpy())
    return torch.from_numpy(dist).to(unknown.device), torch.from_numpy(idx).to(unknown.device)


def three_interpolate(features, idx, weight):
    """features (B,C,m), idx (B,n,3), weight (B,n,3) -> (B,C,n) = sum_j w_j * f[idx_j].

    Accumulated in slot order j = 0,1,2.  Call site pointnet2.py:519.
    """
    B, C, _ = features.shape
    n = idx.shape[1]
    out = None
    for j in range(3):
        index = idx[:, :, j].long().unsqueeze(1).expand(B, C, n)
        term = torch.gather(features, 2, index) * weight[:, :, j].unsqueeze(1)
        out = term if out is None else out + term
    return out
