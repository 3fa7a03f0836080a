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
