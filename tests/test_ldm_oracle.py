"""CPU: the LDM-variant oracle (oracle/ldm_oracle.py) against the golden fixture generated from
the UNMODIFIED reference UNetModelPose (oracle/make_golden.py --only-ldm), the seeded weight
recipe's schema, and the algebra the CUDA engine relies on (one-token cross-attention = a vector)."""
import numpy as np
import pytest
import torch

from _util import rel_l2


@pytest.fixture(scope="module")
def ldm_sd():
    from nope_b200.synth_weights import make_ldm_state_dict
    return make_ldm_state_dict(seed=0)


def test_schema_counts(ldm_sd):
    # 628 tensors / 394.99 M parameters: UNetModelPose.state_dict() of vae_cin_ldm.yaml (encoder stubbed)
    assert len(ldm_sd) == 628
    assert sum(v.numel() for v in ldm_sd.values()) == 394987780


def test_oracle_matches_reference_golden(ldm_sd, golden_dir):
    from oracle import ldm_oracle, unet_oracle
    g = np.load(f"{golden_dir}/ldm_b1_n3.npz")
    ref = torch.from_numpy(g["ref_latent"])
    poses = torch.from_numpy(g["all_relativeR"])
    torch.set_num_threads(max(1, torch.get_num_threads()))
    taps = {}
    with torch.no_grad():
        emb = ldm_oracle.ldm_forward(ldm_sd, ref.expand(2, -1, -1, -1), poses[0, :2], taps=taps)
    assert rel_l2(emb, torch.from_numpy(g["emb"][:2])) < 2e-5
    sim = unet_oracle.l2_similarity(torch.from_numpy(g["query_latent"]), emb[None])
    assert rel_l2(sim, torch.from_numpy(g["similarity"][:, :2])) < 2e-5
    for k, v in taps.items():        # activation statistics of every block (all 3 hypotheses in the fixture)
        st = g["tap:" + k]
        assert abs(float(v.std()) - st[1]) < 0.05 * st[1] + 1e-3, k


def test_one_token_cross_attention_is_a_vector(ldm_sd):
    """attn2 with a single context token returns to_out(to_v(ctx)) for every query: the identity
    nope_ldm::make_cross folds into a per-hypothesis channel vector."""
    from oracle import ldm_oracle
    p = "input_blocks.4.1.transformer_blocks.0.attn2"
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 256, 512, generator=g)
    ctx = torch.randn(2, 1, 512, generator=g)
    out = ldm_oracle.cross_attention(x, ctx, ldm_sd, p, heads=16)
    vec = (ctx[:, 0] @ ldm_sd[p + ".to_v.weight"].t()) @ ldm_sd[p + ".to_out.0.weight"].t() \
        + ldm_sd[p + ".to_out.0.bias"]
    assert rel_l2(out, vec[:, None, :].expand_as(out)) < 1e-6


def test_block_plan_matches_reference_module_list():
    """ldm_block_plan() against the module list UNetModel.__init__ builds for vae_cin_ldm.yaml
    (recorded from the reference: input/output block widths and the skip widths popped per block)."""
    from nope_b200.synth_weights import ldm_block_plan, ldm_flops_per_hyp
    inp, mid, out = ldm_block_plan()
    assert [b[0] for b in inp] == ["conv", "res", "res", "down", "res", "res", "down", "res", "res"]
    assert [b[2] for b in inp] == [256, 256, 256, 256, 512, 512, 512, 1024, 1024]
    assert mid == 1024
    assert [(b[1], b[2], b[3], b[4]) for b in out] == [
        (2048, 1024, 1024, False), (2048, 1024, 1024, False), (1536, 1024, 512, True),
        (1536, 512, 512, False), (1024, 512, 512, False), (768, 512, 256, True),
        (768, 256, 256, False), (512, 256, 256, False), (512, 256, 256, False)]
    f = ldm_flops_per_hyp()
    assert abs(f["total"] / 1e9 - 109.73) < 0.05 and abs(f["attention"] / 1e9 - 6.14) < 0.01


def test_mirror_rejects_unsupported_configurations():
    """only the shipped vae_cin_ldm.yaml configuration resolves; no CUDA needed to find out"""
    from nope_b200.ldm import UNetModelPose
    UNetModelPose(device="cuda:0")                      # constructing does not touch the library
    for bad in (dict(channel_mult=(1, 2, 4, 8)), dict(injecting_condition_twice=True),
                dict(pose_mlp_name="two_layers"), dict(num_head_channels=64), dict(transformer_depth=2),
                dict(use_scale_shift_norm=True), dict(attention_resolutions=(4, 2))):
        with pytest.raises(ValueError):
            UNetModelPose(**bad)


def test_gemm_schedule_count():
    """tools/ldm_conv_table.py replays nope_ldm::forward_chunk's GEMM launches: 135 per chunk with the
    pose-independent prefix hoisted (bench.py --variant ldm reports the same count)."""
    import importlib.util
    import os
    import sys
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "ldm_conv_table.py")
    src = open(path).read().split("rows = list(csv.reader")[0]
    argv = sys.argv
    sys.argv = ["ldm_conv_table.py", "unused", "128"]
    try:
        ns = {"__file__": path, "__name__": "ldm_conv_table"}
        exec(compile(src, path, "exec"), ns)
    finally:
        sys.argv = argv
    sched = ns["schedule"](128)
    assert len(sched) == 135
    assert sum(1 for s in sched if s[0].endswith(".ff1g")) == 16
