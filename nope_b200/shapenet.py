"""Input side of the ShapeNet evaluation: the reference's batch schema and
dataloader keying, without the 2 TB of renders.

* `BATCH_KEYS` / `ShapeNetBatchAdapter`: the dict a test-split `ShapeNet.__getitem__` returns, collated
  (src/dataloader/shapeNet.py:325-357): query / reference [B,3,256,256] float32 in [-1,1], gt_relativeR
  [B,6], all_relativeR [B,N,6], gt_templates [B,N,3,256,256] (optional here: only the reference's
  visualisation reads it), symmetry [B,1] or [B], query_pose [B,3,3], template_poses [B,N,3,3].
  The adapter checks shapes / dtypes, fills what can be derived and moves the tensors to the device.
* `keyed_batches`: what `trainer.test` hands to `PoseConditional.test_step` -- a dict
  {"shapeNet_<category>": batch} over the test categories (model.py:550-565 splits the key on "_";
  categories: src/utils/shapeNet_utils.py:21-32).
* `SyntheticShapeNet`: a stand-in dataset with the same item schema (seeded random images, poses from
  the icosphere grid): plumbing and throughput, not accuracy.
"""
import numpy as np
import torch

from .poses import relative_rotation_6d, synthetic_pose_batch

# src/utils/shapeNet_utils.py:21-32
TEST_CATS = ["bottle", "bus", "clock", "dishwasher", "guitar", "mug", "pistol", "skateboard", "train", "washer"]
BATCH_KEYS = ("query", "reference", "gt_relativeR", "all_relativeR", "gt_templates", "symmetry", "query_pose",
              "template_poses")


class ShapeNetBatchAdapter:
    """Validates a collated test batch of the reference schema and puts it on `device`."""

    def __init__(self, device="cuda:0", image_size=256):
        self.device = torch.device(device)
        self.image_size = image_size

    def __call__(self, batch):
        missing = [k for k in BATCH_KEYS if k not in batch and k != "gt_templates"]
        if missing:
            raise KeyError(f"ShapeNet test batch misses {missing} (schema: src/dataloader/shapeNet.py:348-357)")
        out = {}
        q, r = batch["query"], batch["reference"]
        B = q.shape[0]
        for name, t in (("query", q), ("reference", r)):
            if tuple(t.shape) != (B, 3, self.image_size, self.image_size):
                raise ValueError(f"{name}: expected [B,3,{self.image_size},{self.image_size}], got {tuple(t.shape)}")
            out[name] = t.to(self.device, torch.float32)
        allR = batch["all_relativeR"]
        if allR.dim() != 3 or allR.shape[0] != B or allR.shape[2] != 6:
            raise ValueError(f"all_relativeR: expected [B,N,6], got {tuple(allR.shape)}")
        N = allR.shape[1]
        out["all_relativeR"] = allR.to(self.device, torch.float32)
        gt = batch["gt_relativeR"]
        if tuple(gt.shape) != (B, 6):
            raise ValueError(f"gt_relativeR: expected [B,6], got {tuple(gt.shape)}")
        out["gt_relativeR"] = gt.to(self.device, torch.float32)
        tp = batch["template_poses"]
        if tp.dim() == 3:                      # one shared grid: broadcast like the collated reference batch
            tp = tp[None].expand(B, -1, -1, -1)
        if tuple(tp.shape) != (B, N, 3, 3):
            raise ValueError(f"template_poses: expected [B,{N},3,3], got {tuple(tp.shape)}")
        out["template_poses"] = tp.to(self.device)
        qp = batch["query_pose"]
        if tuple(qp.shape) != (B, 3, 3):
            raise ValueError(f"query_pose: expected [B,3,3], got {tuple(qp.shape)}")
        out["query_pose"] = qp.to(self.device)
        out["symmetry"] = torch.as_tensor(batch["symmetry"]).reshape(B, -1)[:, :1].to(self.device)
        if batch.get("gt_templates") is not None:
            out["gt_templates"] = batch["gt_templates"]      # stays on the host: visualisation only
        return out


def keyed_batches(loaders):
    """loaders: {category: iterable of batches} -> yields {"shapeNet_<category>": batch} per step, stepping
    every category's loader together the way Lightning combines a dict of test dataloaders."""
    its = {c: iter(l) for c, l in loaders.items()}
    while its:
        step = {}
        for c in list(its):
            try:
                step[f"shapeNet_{c}"] = next(its[c])
            except StopIteration:
                del its[c]
        if step:
            yield step


class SyntheticShapeNet(torch.utils.data.Dataset):
    """Items with the schema of a test-split `ShapeNet.__getitem__` (shapeNet.py:338-357)."""

    def __init__(self, category, n_items=4, grid=642, seed=0, with_templates=False):
        self.category, self.n, self.with_templates = category, n_items, with_templates
        _, R = synthetic_pose_batch(grid, 1)
        self.R = R.numpy()
        self.seed = seed * 1000 + (TEST_CATS.index(category) if category in TEST_CATS else 99)

    def __len__(self):
        return self.n

    def __getitem__(self, index):
        g = torch.Generator().manual_seed(self.seed * 100 + index)
        qi = int(torch.randint(0, len(self.R), (1,), generator=g))
        ri = int(torch.randint(0, len(self.R), (1,), generator=g))
        item = {
            "query": torch.rand(3, 256, 256, generator=g) * 2 - 1,
            "reference": torch.rand(3, 256, 256, generator=g) * 2 - 1,
            "gt_relativeR": relative_rotation_6d(self.R[qi][None], self.R[ri])[0],
            "all_relativeR": relative_rotation_6d(self.R, self.R[ri]),
            "symmetry": torch.zeros(1),
            "query_pose": torch.from_numpy(self.R[qi].copy()),
            "template_poses": torch.from_numpy(self.R.copy()),
        }
        if self.with_templates:
            item["gt_templates"] = torch.zeros(len(self.R), 3, 256, 256)
        return item
