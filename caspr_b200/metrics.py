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
