"""CPU: host-side pieces of the evaluation path -- batch schema adapter,
dataloader keying, the shape-filtered checkpoint load of src/utils/weight.py:6-37."""
import pytest
import torch


def test_synthetic_items_follow_the_reference_schema():
    from nope_b200.shapenet import BATCH_KEYS, SyntheticShapeNet
    it = SyntheticShapeNet("bottle", n_items=2, grid=26, with_templates=True)[1]
    assert set(it) == set(BATCH_KEYS)
    assert it["query"].shape == (3, 256, 256) and it["all_relativeR"].shape == (26, 6)
    assert it["template_poses"].shape == (26, 3, 3) and it["query_pose"].shape == (3, 3)
    assert float(it["query"].min()) >= -1 and float(it["query"].max()) <= 1
    # gt_relativeR is the relative rotation of the query pose w.r.t. the reference view: one grid row
    assert any(torch.allclose(it["gt_relativeR"], row, atol=1e-6) for row in it["all_relativeR"])
    # deterministic
    it2 = SyntheticShapeNet("bottle", n_items=2, grid=26)[1]
    assert torch.equal(it2["query"], it["query"])


def test_adapter_validates_and_keys_match_test_step():
    from nope_b200.shapenet import ShapeNetBatchAdapter, SyntheticShapeNet, keyed_batches
    loaders = {c: torch.utils.data.DataLoader(SyntheticShapeNet(c, n_items=3, grid=26), batch_size=2)
               for c in ("bottle", "mug")}
    steps = list(keyed_batches(loaders))
    assert [sorted(s) for s in steps] == [["shapeNet_bottle", "shapeNet_mug"]] * 2
    for name in steps[0]:
        data_name, category = name.split("_")                      # model.py:552
        assert data_name == "shapeNet" and category in ("bottle", "mug")
    ad = ShapeNetBatchAdapter(device="cpu")
    b = ad(steps[1]["shapeNet_mug"])                               # ragged last batch of 1
    assert b["query"].shape == (1, 3, 256, 256) and b["template_poses"].shape == (1, 26, 3, 3)
    bad = dict(steps[0]["shapeNet_mug"])
    bad["all_relativeR"] = bad["all_relativeR"][:, :, :5]
    with pytest.raises(ValueError):
        ad(bad)
    del bad["query_pose"]
    with pytest.raises(KeyError):
        ad(bad)


class _FakeModel:
    def __init__(self):
        self.sd = {"a.weight": torch.zeros(2, 3), "b.bias": torch.ones(4), "c": torch.zeros(1)}
        self.loaded = None

    def state_dict(self):
        return dict(self.sd)

    def load_state_dict(self, sd):
        self.loaded = sd


def test_load_checkpoint_filters_by_key_and_shape():
    """src/utils/weight.py:6-37: prefix removed, entries kept only if the key exists with the same
    shape; everything else keeps the model's value."""
    from nope_b200.weight import load_checkpoint
    m = _FakeModel()
    ckpt = {"state_dict": {"u_net.a.weight": torch.full((2, 3), 7.0), "u_net.b.bias": torch.zeros(5),
                           "u_net.zzz": torch.zeros(1)}}
    loaded, cannot, not_updated = load_checkpoint(m, ckpt, checkpoint_key="state_dict", prefix="u_net.")
    assert loaded == ["a.weight"] and sorted(cannot) == ["b.bias", "zzz"] and not_updated == ["c"]
    assert torch.equal(m.loaded["a.weight"], torch.full((2, 3), 7.0))
    assert torch.equal(m.loaded["b.bias"], torch.ones(4))          # shape mismatch: model value kept


def test_unet_mirror_state_dict_schema():
    from nope_b200.encoder import FeatureExtractor
    from nope_b200.unet import UNet
    u = UNet(u_net_dim=64, rot_representation_dim=6, encoder=FeatureExtractor(descriptor_size=8, backend="torch"),
             device="cpu")
    sd = u.state_dict()
    assert sd["init_conv.weight"].shape == (64, 8, 3, 3) and "encoder.backbone.conv1.weight" in sd
    assert len([k for k in sd if not k.startswith("encoder.")]) == 301
