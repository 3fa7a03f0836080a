"""GPU: the evaluation entry point -- `test_shapeNet.py` drives
`PoseConditional.test_step` / `eval_geodesic` (src/model/model.py:550-565, 268-376) over
"shapeNet_<category>" dataloaders of the reference's batch schema; its top-1 pose index, scores and
validation loss are checked against the oracle's `predict_pose` on the same items.  Also `.sample()`
(model.py:113-124), which the round-1 suite never called."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from _util import log, max_rel, rel_l2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_test_shapenet_entry_point_matches_oracle(tmp_path, seeded_state_dict):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    sys.path.insert(0, ROOT)
    import test_shapeNet
    from nope_b200.shapenet import SyntheticShapeNet
    from oracle import unet_oracle as orc
    cats = ["bottle", "mug"]
    res = test_shapeNet.main(["--batches", "1", "--batch-size", "2", "--grid", "26", "--categories", ",".join(cats),
                              "--save-dir", str(tmp_path), "--json-out", str(tmp_path / "out.json")])
    assert json.load(open(tmp_path / "out.json"))["top1_idx"] == res["top1_idx"]
    assert sorted(res["top1_idx"]) == [f"shapeNet_{c}" for c in cats]
    for c in cats:
        ds = SyntheticShapeNet(c, n_items=2, grid=26)
        items = [ds[i] for i in range(2)]
        q = torch.stack([it["query"] for it in items])
        r = torch.stack([it["reference"] for it in items])
        allR = torch.stack([it["all_relativeR"] for it in items])
        _, idx_o, sim_o, emb_o, qf, rf = orc.predict_pose(seeded_state_dict, q, r, allR, items[0]["template_poses"], k=5)
        assert res["top1_idx"][f"shapeNet_{c}"] == idx_o[:, 0].tolist(), c
        # saved predictions (model.py:361-376) hold the similarity rows
        saved = np.load(tmp_path / "predictions" / f"pred_{c}_step0_rank0.npz")
        e_sim = max_rel(torch.from_numpy(saved["similarity"]), sim_o)
        # validation loss under the GT pose (model.py:281, 106-111): l1 between predicted and query latents
        unet_sd = {k: v for k, v in seeded_state_dict.items() if not k.startswith("encoder.")}
        gt = torch.stack([it["gt_relativeR"] for it in items])
        with torch.no_grad():
            loss_o = float((orc.unet_forward(unet_sd, rf, gt) - qf).abs().flatten(1).mean(1).mean())
        loss = res["scores"][f"loss/val_{c}"]
        log("test_shapenet", category=c, sim_max_rel=e_sim, loss=loss, loss_oracle=loss_o,
            top1=res["top1_idx"][f"shapeNet_{c}"])
        assert e_sim < 1e-3 and abs(loss - loss_o) < 2e-3 * abs(loss_o)
    for c in cats:
        for key in ("top1, accuracy_15", "top1, median", "top3, accuracy_15", "top5, median"):
            assert f"{key}/val_{c}" in res["scores"]


def test_sample_matches_oracle(gpu_model, seeded_state_dict, golden_dir):
    """PoseConditional.sample(reference, relativeR) = encode_image + UNet.forward, decoder-less."""
    from oracle import inputs, unet_oracle as orc
    g = np.load(f"{golden_dir}/cfg1_b1_n6.npz")
    q, r = inputs.make_images(seed=0, batch=1)
    pose = torch.from_numpy(g["all_relativeR"][:, 3])
    feat, rgb = gpu_model.sample(r, pose)
    assert rgb is None and feat.shape == (1, 8, 32, 32)
    e = rel_l2(feat, torch.from_numpy(g["emb"][:, 3]))
    log("sample", emb_rel_l2=e)
    assert e < 2.5e-3
