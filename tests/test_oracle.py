"""CPU: the oracle restatement against the golden fixtures generated from the unmodified
reference (oracle/make_golden.py), plus host-side pose utilities."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import inputs, unet_oracle as orc, weights


def rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm())


@pytest.fixture(scope="module")
def sds(seeded_state_dict):
    sd = seeded_state_dict
    unet = {k: v for k, v in sd.items() if not k.startswith("encoder.")}
    enc = {k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}
    return unet, enc


def test_weight_recipe_is_deterministic(seeded_state_dict, golden_dir):
    meta = json.load(open(os.path.join(golden_dir, "meta.json")))
    assert abs(weights.checksum(seeded_state_dict) - meta["weights_checksum_seed0"]) < 1e-6 * abs(
        meta["weights_checksum_seed0"])
    assert meta["n_unet_tensors"] == 301 and meta["n_encoder_entries"] == 644
    assert sum(1 for k in seeded_state_dict if not k.startswith("encoder.")) == 301


def test_oracle_pinned_by_reference(golden_dir):
    meta = json.load(open(os.path.join(golden_dir, "meta.json")))
    for k, v in meta["oracle_vs_reference_rel_err"].items():
        assert v < 2e-5, (k, v)


def test_cfg1_against_golden(sds, golden_dir):
    unet, enc = sds
    g = np.load(os.path.join(golden_dir, "cfg1_b1_n6.npz"))
    q, r = inputs.make_images(seed=0, batch=1)
    relR, _ = inputs.make_pose_batch("level0_upper", batch=1, n=6)
    assert torch.equal(relR, torch.from_numpy(g["all_relativeR"]))
    with torch.no_grad():
        qf = orc.encode_image(enc, q)
        rf = orc.encode_image(enc, r)
        assert rel(qf, g["query_feat"]) < 1e-5 and rel(rf, g["ref_feat"]) < 1e-5
        emb = orc.generate_templates(unet, torch.from_numpy(g["ref_feat"]), relR)
        assert rel(emb, g["emb"]) < 1e-4
        sim = orc.l2_similarity(torch.from_numpy(g["query_feat"]), emb)
        assert rel(sim, g["similarity"]) < 1e-5
        assert np.array_equal(orc.topk_lowest_index(sim, 5).numpy(), g["nearest_idx"])


def test_unet_taps_against_golden(sds, golden_dir):
    unet, _ = sds
    g = np.load(os.path.join(golden_dir, "cfg1_b1_n6.npz"))
    t = np.load(os.path.join(golden_dir, "unet_taps.npz"))
    taps = {}
    with torch.no_grad():
        orc.unet_forward(unet, torch.from_numpy(g["ref_feat"]),
                         torch.from_numpy(g["all_relativeR"][:, 0]), taps=taps)
    for k, v in taps.items():
        st = np.array([float(v.mean()), float(v.std()), float(v.abs().max())])
        assert np.allclose(st, t[k], rtol=1e-4, atol=1e-5), (k, st, t[k])


def test_l2_similarity_matches_reference_formula_restated_in_loss_py():
    # src/model/loss.py:129-132 restates retrieval on [8,26,4,32,32] random tensors
    g = torch.Generator().manual_seed(0)
    q = torch.randn(8, 4, 32, 32, generator=g)
    t = torch.randn(8, 26, 4, 32, 32, generator=g)
    d = (q.unsqueeze(1).repeat(1, 26, 1, 1, 1) - t) ** 2
    ref = -torch.norm(d, dim=2).sum(axis=3).sum(axis=2)
    assert torch.allclose(orc.l2_similarity(q, t), ref, rtol=1e-6)


def test_topk_tie_break_lowest_index():
    sim = torch.tensor([[1.0, 3.0, 3.0, 2.0, 3.0, 0.0]])
    assert orc.topk_lowest_index(sim, 5).tolist() == [[1, 2, 4, 3, 0]]


def test_pose_grids_and_rot6d(golden_dir):
    fx = inputs.load_pose_fixture()
    assert {k: fx[k].shape[0] for k in fx.files} == {
        "level0_all": 42, "level0_upper": 26, "level2_all": 642, "level2_upper": 341}
    R = fx["level2_all"]
    assert np.abs(R @ R.transpose(0, 2, 1) - np.eye(3)).max() < 1e-6
    from nope_b200 import poses
    a = poses.relative_rotation_6d(R[:10], R[7])
    b = inputs.relative_rot6d(R[:10], R[7])
    assert torch.equal(a, b)
    # synthetic icosphere grids have the reference's sizes and are rotations
    for lvl, n in [(0, 42), (1, 162), (2, 642)]:
        S = poses.icosphere_object_rotations(lvl)
        assert S.shape == (n, 3, 3)
        assert np.abs(S @ S.transpose(0, 2, 1) - np.eye(3)).max() < 1e-9
        assert np.allclose(np.linalg.det(S), 1.0)
    p6, Rg = poses.synthetic_pose_batch(642, batch=2)
    assert p6.shape == (2, 642, 6) and Rg.shape == (642, 3, 3)


def test_encoder_module_matches_oracle(sds):
    """nope_b200.encoder (torch module, reference key names) == oracle.encode_image on CPU."""
    _, enc = sds
    from nope_b200.encoder import FeatureExtractor
    fe = FeatureExtractor(descriptor_size=8, threshold=0.2, normalize=False)
    fe.load_state_dict({k: v for k, v in enc.items()
                        if k.startswith("backbone.") or k.startswith("projector.")})
    q, _ = inputs.make_images(seed=3, batch=1)
    with torch.no_grad():
        assert rel(fe.encode_image(q), orc.encode_image(enc, q)) < 1e-5


def test_geodesic_metric_properties():
    """nope_b200.metrics (restated loss.py:14-115): known angles, symmetry handling, top-k keys."""
    import math
    from nope_b200.metrics import GeodesicError, so3_relative_angle_with_symmetry

    def rot(axis, deg):
        a = math.radians(deg)
        c, s = math.cos(a), math.sin(a)
        m = {"x": [[1, 0, 0], [0, c, -s], [0, s, c]], "y": [[c, 0, s], [0, 1, 0], [-s, 0, c]],
             "z": [[c, -s, 0], [s, c, 0], [0, 0, 1]]}[axis]
        return torch.tensor(m, dtype=torch.float64)
    gt = torch.stack([torch.eye(3, dtype=torch.float64)] * 3)
    pred = torch.stack([rot("x", 30), rot("y", 180), rot("z", 40)])
    e0 = torch.rad2deg(so3_relative_angle_with_symmetry(pred, gt, torch.zeros(3)))
    assert torch.allclose(e0, torch.tensor([30.0, 180.0, 40.0], dtype=torch.float64), atol=1.0)
    e1 = torch.rad2deg(so3_relative_angle_with_symmetry(pred, gt, torch.ones(3)))
    assert e1[1] < 1.0 and abs(float(e1[0]) - 30.0) < 1.0           # Y-180 symmetric object
    e2 = torch.rad2deg(so3_relative_angle_with_symmetry(pred, gt, torch.full((3,), 2.0)))
    assert e2[2] < 1.0                                              # in-plane rotation ignored
    predk = torch.stack([pred, pred.flip(0), gt, gt, gt], dim=1)    # [3, 5, 3, 3]
    err, res = GeodesicError()(predk, gt, torch.zeros(3))
    assert set(res) == {"top1, accuracy_15", "top1, median", "top3, accuracy_15", "top3, median",
                        "top5, accuracy_15", "top5, median"}
    assert float(res["top3, accuracy_15"]) == 100.0 and float(res["top1, accuracy_15"]) == 0.0


def test_geodesic_metric_against_reference_fixture(golden_dir):
    """nope_b200.metrics.GeodesicError == the reference's GeodesicError (loss.py:74-115) on the
    fixture oracle/make_golden.py generated from the unmodified reference module."""
    from nope_b200.metrics import GeodesicError
    g = np.load(os.path.join(golden_dir, "geodesic.npz"))
    predR, gtR = torch.from_numpy(g["predR"]), torch.from_numpy(g["gtR"])
    for name, sym in (("sym0", torch.zeros(24)), ("mixed", torch.from_numpy(g["symmetry"]))):
        err, res = GeodesicError()(predR, gtR, sym)
        assert np.allclose(err.numpy(), g[f"{name}_err"], atol=2e-3), name   # degrees (ref is fp32)
        for k, v in res.items():
            assert abs(float(v) - float(g[f"{name}|{k}"])) < 2e-3, (name, k)
