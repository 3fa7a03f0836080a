"""TEST INFRASTRUCTURE ONLY -- seeded synthetic inputs shared by the golden
generator, the parity tests and the bench's cpu_baseline leg.

Images follow the reference's batch schema (src/dataloader/shapeNet.py:348-357):
float32 NCHW in [-1, 1]; poses are 6-D rotations (first two rows of R_rel,
src/poses/rotation_conversions.py:490-503).
"""
import os
import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                          "tests", "golden")


def make_images(seed, batch, size=256):
    g = torch.Generator(device="cpu").manual_seed(3000 + seed)
    q = torch.rand((batch, 3, size, size), generator=g) * 2 - 1
    r = torch.rand((batch, 3, size, size), generator=g) * 2 - 1
    return q, r


def load_pose_fixture():
    """tests/golden/pose_grids.npz: object poses (3x3, float64) of the reference's
    icosphere grids, written by oracle/make_golden.py from
    src/poses/predefined_poses/*.npy."""
    return np.load(os.path.join(GOLDEN_DIR, "pose_grids.npz"))


def relative_rot6d(template_R, ref_R):
    """all_relativeR = rot6d(R_template @ inv(R_ref))  (shapeNet.py:243-250,302-307)."""
    rel = template_R @ np.linalg.inv(ref_R)
    rel = torch.tensor(rel, dtype=torch.float32)
    return rel[..., :2, :].reshape(*rel.shape[:-2], 6)


def make_pose_batch(grid, batch, n=None, ref_index=7):
    """[B,N,6] relative rotations: batch item b uses grid pose (ref_index + 3 b) as its
    reference view.  Returns (all_relativeR [B,N,6] f32, template_poses [N,3,3] f64)."""
    fx = load_pose_fixture()
    R = fx[grid]
    if n is not None:
        R = R[:n]
    full = fx[grid]
    rels = [relative_rot6d(R, full[(ref_index + 3 * b) % len(full)]) for b in range(batch)]
    return torch.stack(rels), torch.from_numpy(R.copy())
