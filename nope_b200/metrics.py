"""Geodesic pose metric of the reference's evaluation (src/model/loss.py:8-115), restated in
plain torch so `test_shapeNet.py` can report the numbers the reference logs.  Host-side,
after the hot path.

`so3_relative_angle` comes from pytorch3d in the reference (not vendored there, version not
pinned in environment.yml; call sites loss.py:3,20-48).  Restated from its published formula:
angle = acos(clamp((trace(R1 R2^T) - 1) / 2, -1 + b, 1 - b)) with cos_bound b = 1e-4.
"""
import math

import torch
import torch.nn.functional as F


def so3_relative_angle(R1, R2, cos_bound=1e-4):
    R12 = torch.bmm(R1, R2.transpose(1, 2))
    cos = (R12[:, 0, 0] + R12[:, 1, 1] + R12[:, 2, 2] - 1.0) * 0.5
    return torch.acos(cos.clamp(-1.0 + cos_bound, 1.0 - cos_bound))


def _roty180(dtype, device):
    # load_rotation_transform("y", 180)[:3, :3]  (src/poses/utils.py:136-139)
    return torch.tensor([[-1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, -1.0]], dtype=dtype, device=device)


def so3_relative_angle_with_symmetry(pred, gt, symmetry):
    """loss.py:14-71.  symmetry: 0 none, 1 two-fold about Y, 2 circular (in-plane ignored)."""
    pred, gt = pred.to(torch.float64), gt.to(torch.float64)
    err = torch.zeros(pred.shape[0], dtype=torch.float64, device=pred.device)
    m0, m1, m2 = symmetry == 0, symmetry == 1, symmetry == 2
    if m0.any():
        err[m0] = so3_relative_angle(pred[m0], gt[m0])
    if m1.any():
        a = so3_relative_angle(pred[m1], gt[m1])
        rot = _roty180(pred.dtype, pred.device).expand(int(m1.sum()), 3, 3)
        b = so3_relative_angle(torch.bmm(rot, pred[m1]), gt[m1])
        err[m1] = torch.minimum(a, b)
    if m2.any():
        flip = torch.tensor([[1.0, 0, 0], [0, -1.0, 0], [0, 0, -1.0]], dtype=pred.dtype, device=pred.device)
        pc = torch.matmul(flip, torch.linalg.inv(pred[m2]))      # object pose -> camera, OpenGL frame
        gc = torch.matmul(flip, torch.linalg.inv(gt[m2]))
        err[m2] = torch.acos(F.cosine_similarity(pc[:, 2, :3], gc[:, 2, :3]).clamp(-1, 1))
    return err


class GeodesicError:
    """loss.py:74-115: accuracy at the thresholds (degrees) and median, for the top-1 / top-3 /
    top-5 candidates.  Same result keys as the reference."""

    def __init__(self, thresholds=(15,)):
        self.thresholds = list(thresholds)

    def __call__(self, predR, gtR, symmetry):
        if predR.dim() == 3:
            predR = predR[:, None]
        B, K = predR.shape[:2]
        errors = torch.zeros((B, K), dtype=torch.float64, device=predR.device)
        results = {}
        for k in range(K):
            errors[:, k] = torch.rad2deg(so3_relative_angle_with_symmetry(predR[:, k], gtR, symmetry))
            if k in (0, 2, 4):
                top = errors[:, : k + 1].min(dim=1).values
                for t in self.thresholds:
                    results[f"top{k + 1}, accuracy_{t}"] = (top <= t).float().mean() * 100
                results[f"top{k + 1}, median"] = top.median()
        return errors[:, 0], results
