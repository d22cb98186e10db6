"""CPU restatement of the approximate earth mover's distance the reference evaluates with (TEST INFRASTRUCTURE ONLY).

``utils/emd.py:11-12`` calls ``emd_cuda.approxmatch_forward`` / ``matchcost_forward``: the PyTorchEMD extension
(daerduoCarey/PyTorchEMD, a port of the approxmatch kernels of Fan, Su, Guibas, "A Point Set Generation Network", CVPR
2017).  The extension is a third-party dependency that is absent from /root/reference and unpinned there
(**parity unpinned**); what follows restates its published algorithm in float64 numpy: ten annealing levels
``-4^7 ... -4^-1, 0``, each a row normalisation, a column normalisation with saturation and a match update, then
``cost = sum match(k,l) |p_k - q_l|`` (``evaluations.py:45-46`` divides by the number of points)."""
import numpy as np


def approx_emd(xyz1, xyz2):
    """xyz1 (B,n,3), xyz2 (B,m,3) -> cost (B,) float64."""
    xyz1, xyz2 = np.asarray(xyz1, dtype=np.float64), np.asarray(xyz2, dtype=np.float64)
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    multiL, multiR = (1.0, float(n // m)) if n >= m else (float(m // n), 1.0)
    costs = np.zeros(B)
    for b in range(B):
        dd = ((xyz1[b][:, None, :] - xyz2[b][None, :, :]) ** 2).sum(-1)          # (n,m)
        dist = np.sqrt(dd)
        remainL, remainR = np.full(n, multiL), np.full(m, multiR)
        cost = 0.0
        for j in range(7, -3, -1):
            level = 0.0 if j == -2 else -(4.0 ** j)
            K = np.exp(level * dd)
            ratioL = remainL / (1e-9 + K @ remainR)
            sumr = (K.T @ ratioL) * remainR
            consumption = np.minimum(remainR / (sumr + 1e-9), 1.0)
            ratioR = consumption * remainR
            remainR = np.maximum(0.0, remainR - sumr)
            w = K * ratioL[:, None] * ratioR[None, :]
            cost += (w * dist).sum()
            remainL = np.maximum(0.0, remainL - w.sum(1))
        costs[b] = cost
    return costs
