#!/usr/bin/env python
"""Evaluation entry point the reference's README names (README.md:82) but does not ship
(SURVEY.md F1): what `trainer.test(model, dataloaders={"shapeNet_<category>": ...})` would run --
for every test batch `PoseConditional.test_step` (src/model/model.py:550-565) -> `eval_geodesic`
(model.py:268-376): validation loss under the ground-truth pose, pose-grid sweep + retrieval from one
reference view, geodesic accuracy (src/model/loss.py:74-115), predictions saved per step.

The ShapeNet renders (~2 TB) and the trained checkpoint are not available here, so without a dataset
the script runs on `SyntheticShapeNet` items -- same item schema (src/dataloader/shapeNet.py:348-357),
same "shapeNet_<category>" dataloader keys over the reference's test categories
(src/utils/shapeNet_utils.py:21-32), seeded random images / weights -- and says so: it then measures
plumbing and throughput, not accuracy.  Batches of the reference schema from any other source go through
`nope_b200.shapenet.ShapeNetBatchAdapter` the same way.  With torchrun (one process per GPU) the pose
grid is sharded across ranks (one all-gather of packed top-k records per batch).

  python test_shapeNet.py --batches 1 --batch-size 2 --grid 642 --categories bottle,mug
  python test_shapeNet.py --checkpoint last.ckpt ...        (reference Lightning checkpoint)
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch


def build_loaders(categories, batches, batch_size, grid, seed=0):
    from nope_b200.shapenet import SyntheticShapeNet
    return {c: torch.utils.data.DataLoader(SyntheticShapeNet(c, n_items=batches * batch_size, grid=grid, seed=seed),
                                           batch_size=batch_size, shuffle=False, num_workers=0)
            for c in categories}


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--checkpoint", default=None, help="Lightning .ckpt or state_dict of the reference model")
    ap.add_argument("--batches", type=int, default=1, help="batches per category")
    ap.add_argument("--batch-size", type=int, default=2)
    ap.add_argument("--grid", type=int, default=642, help="pose-grid size (642 = level 2 'all'; 26/341 'upper')")
    ap.add_argument("--categories", default="bottle,mug", help="comma list out of the reference's test_cats, or 'all'")
    ap.add_argument("--metric", default="l2", choices=["l2", "cosine", "cosine_occlusion"])
    ap.add_argument("--precision", default="fp16", choices=["fp16", "fp16_w2", "parity", "parity_fast", "bf16"])
    ap.add_argument("--save-dir", default=None, help="predictions/pred_<cat>_step<k>_rank<r>.npz are written here")
    ap.add_argument("--json-out", default=None)
    args = ap.parse_args(argv)
    import torch.distributed as dist
    from nope_b200.model import build_model
    from nope_b200.shapenet import TEST_CATS, ShapeNetBatchAdapter, keyed_batches
    from nope_b200.weight import load_checkpoint
    cats = TEST_CATS if args.categories == "all" else [c for c in args.categories.split(",") if c]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    model = build_model(device=f"cuda:{local}", similarity_metric=args.metric, precision=args.precision)
    model.save_dir = args.save_dir
    model.global_rank = rank
    from nope_b200 import synth_weights as weights
    model.load_state_dict(weights.make_full_state_dict(seed=0))
    weights_desc = "seeded random init (no trained checkpoint available)"
    if args.checkpoint:
        # the reference's own loading rule (src/utils/weight.py:6-37): prefix stripped, shape-filtered
        loaded, cannot, not_updated = load_checkpoint(model, args.checkpoint, checkpoint_key="state_dict")
        weights_desc = f"{args.checkpoint} ({len(loaded)} tensors loaded, {len(cannot)} skipped)"
    if world > 1:
        from nope_b200.dist import ShardedSweep
        model.dist = ShardedSweep()
    adapter = ShapeNetBatchAdapter(device=f"cuda:{local}")
    loaders = build_loaders(cats, args.batches, args.batch_size, args.grid)
    n_hyp, t0, top1 = 0, time.time(), {}
    for idx_batch, step in enumerate(keyed_batches(loaders)):
        step = {name: adapter(b) for name, b in step.items()}
        out = model.test_step(step, idx_batch)                      # model.py:550-565
        for name, (err, nearest_idx, sim) in out.items():
            top1.setdefault(name, []).extend(nearest_idx[:, 0].tolist())
            n_hyp += sim.numel()
    torch.cuda.synchronize()
    dt = time.time() - t0
    # Lightning reports the epoch mean of everything logged
    scores = {k: sum(v) / len(v) for k, v in model.logged.items()}
    result = {
        "data": "synthetic items with the ShapeNet test schema (accuracy numbers are meaningless without "
                "the dataset + checkpoint)",
        "weights": weights_desc, "grid": args.grid, "categories": cats, "batches_per_category": args.batches,
        "batch_size": args.batch_size, "gpus": world, "metric": args.metric, "precision": args.precision,
        "hypotheses": n_hyp, "hyp_per_s_incl_first_call": n_hyp / dt, "top1_idx": top1, "scores": scores}
    if rank == 0:
        print(json.dumps(result))
        if args.json_out:
            with open(args.json_out, "w") as f:
                json.dump(result, f)
    if world > 1 and argv is None:
        dist.destroy_process_group()
    return result


if __name__ == "__main__":
    main()
