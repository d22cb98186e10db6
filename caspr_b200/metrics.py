"""Evaluation metrics of the reference's test protocol on the CUDA kernels (mirror of utils/evaluations.py:36-49 and
:243-254).  Inputs are CUDA float32 tensors; nothing here touches the host until the caller asks for numpy."""
from . import ops


def eval_reconstr_frames(pred, gt):
    """evaluations.py:36-49: pred, gt (F,P,3) clouds -> (chamfer (F,), emd (F,)) as numpy arrays.
    chamfer = mean squared NN distance pred->gt + gt->pred (tk3dv ChamferDistance), emd = approximate earth mover's
    distance divided by the number of predicted points (utils/emd.py)."""
    d1, d2 = ops.chamfer(pred, gt)
    mean_dist = d1.mean(dim=1) + d2.mean(dim=1)
    cur_emd = ops.emd(pred, gt) / pred.size(1)
    return tuple(r.cpu().numpy() for r in (mean_dist, cur_emd))


def tnocs_regression_error(pred_tnocs, nocs_out):
    """evaluations.py:243-254: -> (space (B,T), time (B,T)) numpy arrays."""
    space, terr = ops.tnocs_error(pred_tnocs[..., :4], nocs_out[..., :4])
    return space.cpu().numpy(), terr.cpu().numpy()


RANSAC_DISTANCE_THRESHOLD = 0.015      # evaluations.py:367
RANSAC_HYPOTHESES = 5000               # RANSACConvergenceCriteria(50000, 5000): min(max_iteration, max_validation)


def ransac_camera_pose(pred_tnocs, pcl_in, samples=None, generator=None, refine=False):
    """evaluations.py:345-380: rigid object pose of every frame from the T-NOCS regression.

    pred_tnocs (B,T,N,>=3) predicted NOCS, pcl_in (B,T,N,>=3) observed points (same point order = the correspondences).
    `samples` (B*T, H, 4) int32 fixes the hypotheses; by default H = 5000 draws per frame from `generator` (a CUDA
    torch.Generator; open3d uses C rand()).  -> dict of CUDA tensors: R (B,T,3,3), t (B,T,3), fitness, inlier_rmse."""
    import torch
    B, T, N = pred_tnocs.shape[:3]
    src = (pred_tnocs[..., :3] - 0.5).reshape(B * T, N, 3)
    dst = pcl_in[..., :3].reshape(B * T, N, 3)
    if samples is None:
        samples = torch.randint(0, N, (B * T, RANSAC_HYPOTHESES, 4), device=src.device, generator=generator,
                                dtype=torch.int32)
    out = ops.ransac_pose(src, dst, samples, RANSAC_DISTANCE_THRESHOLD, refine=refine)
    return {'R': out['R'].view(B, T, 3, 3), 't': out['t'].view(B, T, 3), 'fitness': out['fitness'].view(B, T),
            'inlier_rmse': out['inlier_rmse'].view(B, T), 'best': out['best'].view(B, T)}


def ransac_pose_errors(R_pred, t_pred, R_gt, t_gt, gt_nocs, pcl_in):
    """evaluations.py:386-430 for all frames at once (torch on the device).  R_* (B,T,3,3), t_* (B,T,3), gt_nocs and
    pcl_in (B,T,N,>=3).  -> dict of (B,T) tensors: trans (|t_pred - t_gt|), rot (degrees), point (median) and
    point_mean distance between the observed points and the ground-truth NOCS moved by the predicted pose."""
    import math
    import torch
    # float64 like the reference's numpy code: arccos near 1 loses half the digits in float32
    R_pred, t_pred, R_gt, t_gt = R_pred.double(), t_pred.double(), R_gt.double(), t_gt.double()
    g = gt_nocs[..., :3].double() - 0.5
    moved = torch.einsum('btij,btnj->btni', R_pred, g) + t_pred.unsqueeze(2)
    dist = (moved - pcl_in[..., :3].double()).norm(dim=-1)
    tr = torch.einsum('btji,btji->bt', R_pred, R_gt)                  # trace(R_pred^T R_gt)
    cosang = ((tr - 1.0) / 2.0).clamp(-1.0, 1.0)
    return {'trans': (t_pred - t_gt).norm(dim=-1), 'rot': torch.acos(cosang) * (180.0 / math.pi),
            'point': torch.quantile(dist, 0.5, dim=-1), 'point_mean': dist.mean(dim=-1)}
