"""Oracle restatement of open3d's correspondence RANSAC as the reference calls it.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  **Parity unpinned**: open3d is not under ``/root/reference``
and the reference does not pin a version; the ``o3d.registration`` namespace it uses (``utils/evaluations.py:368``)
dates it <= 0.10.  Published algorithm of ``registration_ransac_based_on_correspondence`` in those versions::

    for itr in range(min(max_iteration, max_validation)):          # RANSACConvergenceCriteria(50000, 5000) -> 5000
        ransac_corres = [corres[rand() % len(corres)] for _ in range(ransac_n)]          # ransac_n = 4
        T = TransformationEstimationPointToPoint(False).ComputeTransformation(...)        # Umeyama / Kabsch, no scale
        fitness, inlier_rmse = evaluate(T)    # inliers: |T src_i - dst_i| < max_correspondence_distance over ALL corres
        keep T if fitness is higher, or equal with a lower inlier_rmse

The call site (``evaluations.py:360-380``) uses identity correspondences, threshold 0.015.  open3d draws the samples
with C ``rand()``; the oracle (like the CUDA kernel) takes them as an argument.  The inlier test follows the declared
fp32 arithmetic of ``caspr_b200/csrc/ransac.cu``: R, t rounded to float32, every product and sum rounded separately.
"""
import numpy as np


def kabsch(src, dst):
    """Least-squares rigid transform dst ~ R src + t (float64; SVD with the reflection fix of Umeyama 1991)."""
    src, dst = np.asarray(src, dtype=np.float64), np.asarray(dst, dtype=np.float64)
    cs, cd = src.mean(0), dst.mean(0)
    H = (dst - cd).T @ (src - cs)
    U, _, Vt = np.linalg.svd(H)
    S = np.diag([1.0, 1.0, np.sign(np.linalg.det(U) * np.linalg.det(Vt))])
    R = U @ S @ Vt
    return R, cd - R @ cs


def _sqdist_after_f32(R, t, s, d):
    f = np.float32
    R, t, s, d = R.astype(f), t.astype(f), s.astype(f), d.astype(f)
    p = [(((R[a, 0] * s[:, 0]).astype(f) + (R[a, 1] * s[:, 1]).astype(f)).astype(f)
          + (R[a, 2] * s[:, 2]).astype(f)).astype(f) + t[a] for a in range(3)]
    e = [(p[a].astype(f) - d[:, a]).astype(f) for a in range(3)]
    return (((e[0] * e[0]).astype(f) + (e[1] * e[1]).astype(f)).astype(f) + (e[2] * e[2]).astype(f)).astype(f)


def ransac_pose(src, dst, samples, max_distance=0.015):
    """src, dst (N,3); samples (H,4) int.  -> dict: R, t (float64 fit of the winner), best, fitness, inlier_rmse,
    counts (H,) inlier count of every hypothesis, Rs (H,3,3), ts (H,3)."""
    src, dst = np.asarray(src, dtype=np.float32), np.asarray(dst, dtype=np.float32)
    thr2 = np.float32(max_distance) * np.float32(max_distance)
    H = samples.shape[0]
    counts = np.zeros(H, dtype=np.int64)
    mse = np.zeros(H, dtype=np.float64)
    Rs, ts = np.zeros((H, 3, 3)), np.zeros((H, 3))
    for h in range(H):
        idx = samples[h]
        R, t = kabsch(src[idx], dst[idx])
        Rs[h], ts[h] = R, t
        d2 = _sqdist_after_f32(R, t, src, dst)
        inl = d2 < thr2
        counts[h] = int(inl.sum())
        mse[h] = d2[inl].astype(np.float64).sum() / counts[h] if counts[h] else 0.0
    best = 0
    for h in range(1, H):
        if counts[h] > counts[best] or (counts[h] == counts[best] and mse[h] < mse[best]):
            best = h
    return {'R': Rs[best], 't': ts[best], 'best': best, 'fitness': counts[best] / float(src.shape[0]),
            'inlier_rmse': float(np.sqrt(mse[best])), 'counts': counts, 'Rs': Rs, 'ts': ts}


def pose_errors(R_pred, t_pred, R_gt, t_gt, gt_nocs, pts):
    """evaluations.py:386-430 for one frame (numpy, as the reference writes it)."""
    g = gt_nocs - 0.5
    moved = np.dot(R_pred, g.T).T + t_pred
    dist = np.linalg.norm(moved - pts, axis=1)
    cosang = np.clip((np.trace(np.dot(R_pred.T, R_gt)) - 1.0) / 2.0, -1.0, 1.0)
    return {'trans': np.linalg.norm(t_pred - t_gt), 'rot': np.degrees(np.arccos(cosang)),
            'point': np.median(dist), 'point_mean': np.mean(dist)}
