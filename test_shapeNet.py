#!/usr/bin/env python
"""Evaluation entry point the reference's README names (README.md:82) but does not ship
(SURVEY.md F1): for every test batch, predict the query pose from one reference view by
sweeping the pose grid, and report the geodesic metrics the reference logs
(src/model/model.py:268-358, src/model/loss.py:74-115).

The ShapeNet renders (~2 TB) and the trained checkpoint are not available here, so without
--data-root the script runs on SYNTHETIC batches with the reference's batch schema
(src/dataloader/shapeNet.py:348-357) and seeded random weights -- it then measures plumbing and
throughput, not accuracy, and says so.  With torchrun (one process per GPU) the pose grid is
sharded across ranks.

  python test_shapeNet.py --batches 4 --batch-size 2 --grid 642
  python test_shapeNet.py --checkpoint last.ckpt --data-root /data/shapenet ...   (needs a loader)
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch


def synthetic_batches(n_batches, batch_size, grid, seed=0):
    """Yields dicts with the keys of src/dataloader/shapeNet.py:348-357 (gt_templates omitted:
    it is only used for visualisation in the reference)."""
    from nope_b200.poses import relative_rotation_6d, synthetic_pose_batch
    g = torch.Generator().manual_seed(seed)
    _, R = synthetic_pose_batch(grid, 1)
    R = R.numpy()
    for _ in range(n_batches):
        qi = torch.randint(0, len(R), (batch_size,), generator=g)
        ri = torch.randint(0, len(R), (batch_size,), generator=g)
        yield {
            "query": torch.rand(batch_size, 3, 256, 256, generator=g) * 2 - 1,
            "reference": torch.rand(batch_size, 3, 256, 256, generator=g) * 2 - 1,
            "all_relativeR": torch.stack([relative_rotation_6d(R, R[int(i)]) for i in ri]),
            "gt_relativeR": torch.stack([relative_rotation_6d(R[int(q)][None], R[int(i)])[0]
                                         for q, i in zip(qi, ri)]),
            "query_pose": torch.from_numpy(R[qi.numpy()]),
            "template_poses": torch.from_numpy(R)[None].expand(batch_size, -1, -1, -1),
            "symmetry": torch.zeros(batch_size, 1),
        }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--checkpoint", default=None, help="Lightning .ckpt or state_dict of the reference model")
    ap.add_argument("--data-root", default=None)
    ap.add_argument("--batches", type=int, default=2)
    ap.add_argument("--batch-size", type=int, default=2)
    ap.add_argument("--grid", type=int, default=642, help="pose-grid size (642 = level 2, 'all')")
    ap.add_argument("--metric", default="l2", choices=["l2", "cosine"])
    args = ap.parse_args()
    if args.data_root is not None:
        raise SystemExit("a ShapeNet render loader is out of scope (SURVEY.md section 2); pass batches "
                         "with the reference schema to PoseConditional.predict_pose instead")
    import torch.distributed as dist
    from nope_b200.metrics import GeodesicError
    from nope_b200.model import build_model
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    model = build_model(device=f"cuda:{local}", similarity_metric=args.metric)
    if args.checkpoint:
        sd = torch.load(args.checkpoint, map_location="cpu")
        model.load_state_dict(sd.get("state_dict", sd))
        weights_desc = args.checkpoint
    else:
        from nope_b200 import synth_weights as weights
        model.load_state_dict(weights.make_full_state_dict(seed=0))
        weights_desc = "seeded random init (no trained checkpoint available)"
    if world > 1:
        from nope_b200.dist import ShardedSweep
        model.dist = ShardedSweep()
    metric = GeodesicError()
    errs, n_hyp, t0 = [], 0, time.time()
    agg = {}
    for batch in synthetic_batches(args.batches, args.batch_size, args.grid):
        R, idx, sim = model.predict_pose(batch["query"], batch["reference"], batch["all_relativeR"],
                                         batch["template_poses"], k=5)
        err, res = metric(R.cpu(), batch["query_pose"], batch["symmetry"].reshape(-1))
        errs.append(err)
        for k, v in res.items():
            agg.setdefault(k, []).append(float(v))
        n_hyp += batch["all_relativeR"].shape[0] * batch["all_relativeR"].shape[1]
    torch.cuda.synchronize()
    dt = time.time() - t0
    if rank == 0:
        print(json.dumps({
            "data": "synthetic (accuracy numbers are meaningless without the dataset + checkpoint)",
            "weights": weights_desc, "grid": args.grid, "batches": args.batches,
            "batch_size": args.batch_size, "gpus": world, "hypotheses": n_hyp,
            "hyp_per_s_incl_first_call": n_hyp / dt,
            **{k: sum(v) / len(v) for k, v in agg.items()}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
