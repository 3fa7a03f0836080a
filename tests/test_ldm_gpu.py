"""GPU: the LDM-variant UNet (UNetModelPose) through the Python mirror / C ABI
against the oracle run live on the host CPU and the golden fixture from the unmodified reference.

Every module is also driven in isolation (nope_ldm_run_block): it gets the ORACLE's input for
that module and must reproduce the oracle's output, so one failing kernel does not hide the rest.

Tolerances (fp16 storage of activations / weights / softmax probabilities, fp32 accumulation,
statistics, softmax and LayerNorm): per module rel-L2 <= BLOCK_TOL; l2 scores max-rel <= SIM_TOL;
whole network embeddings rel-L2 <= EMB_TOL = the north-star 1e-3 in the exact-weights mode
(precision="fp16_w2": W = W_hi + W_lo K-segments), <= EMB_TOL_FP16 in the fast fp16-weights mode, whose
weight rounding alone is 0.62e-3 of its measured 1.08e-3 (CPU experiment on the oracle)."""
import numpy as np
import pytest
import torch

from _util import log, max_rel, rel_l2

pytestmark = pytest.mark.gpu

BLOCK_TOL = 1e-3    # measured 3.0e-4 .. 4.1e-4 per module
EMB_TOL = 1e-3      # exact-weights mode: the north-star tolerance
EMB_TOL_FP16 = 1.5e-3   # fp16-weights mode: measured 1.07e-3 end to end (cumulative 1.3e-3 at the bottleneck)
SIM_TOL = 1e-3      # measured 2.9e-5
N_HYP = 2


def _plan():
    from nope_b200.synth_weights import ldm_block_plan
    return ldm_block_plan()


def _modules():
    """(name, kind, input tap(s), output tap) for every module of the forward pass."""
    inp, _, out = _plan()
    mods = []
    for i, b in enumerate(inp):
        p = f"input_blocks.{i}"
        prev = f"input_blocks.{i - 1}"
        if b[0] == "res":
            mods.append((p + ".0", "res", (prev, None), p + ".0"))
            mods.append((p + ".1", "st", (p + ".0", None), p))
        elif b[0] == "down":
            mods.append((p + ".0.op", "resample", (prev, None), p))
    last_in = f"input_blocks.{len(inp) - 1}"
    mods.append(("middle_block.0", "res", (last_in, None), "middle_block.0"))
    mods.append(("middle_block.1", "st", ("middle_block.0", None), "middle_block.1"))
    mods.append(("middle_block.2", "res", ("middle_block.1", None), "middle_block"))
    prev = "middle_block"
    for i, b in enumerate(out):
        p = f"output_blocks.{i}"
        mods.append((p + ".0", "res", (prev, f"input_blocks.{len(inp) - 1 - i}"), p + ".0"))
        mods.append((p + ".1", "st", (p + ".0", None), p + ".1"))
        if b[4]:
            mods.append((p + ".2.conv", "resample", (p + ".1", None), p))
        prev = p
    return mods


@pytest.fixture(scope="module")
def ldm_sd():
    from nope_b200.synth_weights import make_ldm_state_dict
    return make_ldm_state_dict(seed=0)


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(f"{golden_dir}/ldm_b1_n3.npz")


@pytest.fixture(scope="module")
def oracle_taps(ldm_sd, golden):
    from oracle import ldm_oracle
    ref = torch.from_numpy(golden["ref_latent"])
    poses = torch.from_numpy(golden["all_relativeR"])[0, :N_HYP]
    taps = {}
    with torch.no_grad():
        emb = ldm_oracle.ldm_forward(ldm_sd, ref.expand(N_HYP, -1, -1, -1), poses, taps=taps)
    taps["emb"] = emb
    taps["poses"] = poses
    return taps


@pytest.fixture(scope="module")
def ldm_model(ldm_sd):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from nope_b200.ldm import UNetModelPose
    m = UNetModelPose(device="cuda:0", chunk=8)
    m.load_state_dict(ldm_sd)
    return m


@pytest.mark.parametrize("n_tok,C", [(64, 1024), (256, 512), (1024, 256), (320, 64)])
@pytest.mark.parametrize("impl", ["simt", "tcgen05"])
def test_mh_attention(impl, n_tok, C):
    """softmax(q k^T / sqrt(32)) v per head against torch fp32 on the fp16-rounded operands."""
    from nope_b200.ldm import mh_attention
    g = torch.Generator().manual_seed(n_tok + C)
    qkv = torch.randn(2, n_tok, 3 * C, generator=g) * 1.5
    qkv = qkv.half().float()
    out = mh_attention(qkv.cuda(), impl=impl).cpu()
    H = C // 32
    q, k, v = [t.reshape(2, n_tok, H, 32).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=-1)]
    att = (torch.einsum("bhid,bhjd->bhij", q, k) * 32 ** -0.5).softmax(dim=-1)
    ref = torch.einsum("bhij,bhjd->bhid", att, v).permute(0, 2, 1, 3).reshape(2, n_tok, C)
    e = rel_l2(out, ref)
    log("ldm_mh_attention", impl=impl, n_tok=n_tok, C=C, rel_l2=e, max_rel=max_rel(out, ref))
    assert e < 1.5e-3


@pytest.mark.parametrize("mod", _modules(), ids=lambda m: m[0])
def test_module_in_isolation(ldm_model, oracle_taps, mod):
    name, kind, (in0, in1), out_tap = mod
    x0 = oracle_taps[in0]
    x1 = oracle_taps[in1] if in1 else None
    want = oracle_taps[out_tap]
    got = ldm_model.run_block(name, x0, x1, poses=oracle_taps["poses"] if kind == "st" else None,
                              out_channels=want.shape[1], out_side=want.shape[2])
    e = rel_l2(got, want)
    log("ldm_module", module=name, kind=kind, rel_l2=e, max_rel=max_rel(got, want))
    assert e < BLOCK_TOL


@pytest.mark.parametrize("attn", ["tcgen05", "simt"])
def test_forward_vs_oracle(ldm_model, oracle_taps, golden, attn):
    ldm_model.set_impl(attn=attn)
    try:
        ref = torch.from_numpy(golden["ref_latent"])
        emb = ldm_model(ref.expand(N_HYP, -1, -1, -1), oracle_taps["poses"])
    finally:
        ldm_model.set_impl()
    e = rel_l2(emb, oracle_taps["emb"])
    log("ldm_forward_vs_oracle", attn=attn, emb_rel_l2=e, launches=ldm_model.last_launch_count)
    assert e < EMB_TOL_FP16


def test_geglu_fusion_matches_separate_kernel(ldm_model, oracle_taps):
    """the GEGLU epilogue of the projection GEMM against the separate elementwise kernel"""
    name = "input_blocks.4.1"
    x = oracle_taps["input_blocks.4.0"]
    want = oracle_taps["input_blocks.4"]
    fused = ldm_model.run_block(name, x, poses=oracle_taps["poses"], out_channels=512, out_side=16)
    ldm_model.set_option("fuse_geglu", 0)
    try:
        plain = ldm_model.run_block(name, x, poses=oracle_taps["poses"], out_channels=512, out_side=16)
    finally:
        ldm_model.set_option("fuse_geglu", 1)
    log("ldm_geglu_fusion", fused=rel_l2(fused, want), separate=rel_l2(plain, want), fused_vs_separate=rel_l2(fused, plain))
    assert rel_l2(fused, want) < BLOCK_TOL and rel_l2(plain, want) < BLOCK_TOL


def test_taps_vs_oracle(ldm_model, oracle_taps, golden):
    """cumulative error along the network (informational thresholds: 2x the end-to-end tolerance)"""
    ref = torch.from_numpy(golden["ref_latent"])
    worst = 0.0
    for tap in ["input_blocks.0", "input_blocks.1.0", "input_blocks.1", "input_blocks.3", "input_blocks.5",
                "input_blocks.8", "middle_block", "output_blocks.2", "output_blocks.5", "output_blocks.8"]:
        got = ldm_model.debug_tap(ref, oracle_taps["poses"], tap)
        e = rel_l2(got, oracle_taps[tap])
        log("ldm_tap", tap=tap, rel_l2=e)
        worst = max(worst, e)
    assert worst < 2 * EMB_TOL_FP16


def test_sweep_golden(ldm_model, golden):
    """B=1, N=3 sweep with fused score / top-k against the unmodified reference's outputs."""
    ref = torch.from_numpy(golden["ref_latent"])
    qry = torch.from_numpy(golden["query_latent"])
    poses = torch.from_numpy(golden["all_relativeR"])
    out = ldm_model.sweep(ref, poses, qry, want_emb=True, k=3)
    e_emb = rel_l2(out["emb"][0], torch.from_numpy(golden["emb"]))
    e_sim = max_rel(out["sim"], torch.from_numpy(golden["similarity"]))
    order = torch.from_numpy(golden["similarity"]).argsort(dim=1, descending=True)
    log("ldm_sweep_golden", emb_rel_l2=e_emb, sim_max_rel=e_sim, topi=out["topi"].tolist(), ref_order=order.tolist())
    assert e_emb < EMB_TOL_FP16 and e_sim < SIM_TOL
    assert torch.equal(out["topi"].cpu(), order)
    # chunking / batching invariance: two references, chunk smaller than N
    ldm_model.set_chunk(2)
    try:
        out2 = ldm_model.sweep(torch.cat([ref, qry]), torch.cat([poses, poses]), torch.cat([qry, ref]), k=1)
    finally:
        ldm_model.set_chunk(8)
    assert torch.equal(out2["sim"][0], out["sim"][0])


def test_hoisted_prefix_is_bitwise_neutral(ldm_model, golden):
    """running the pose-independent prefix once per reference changes no bit of the result"""
    ref = torch.from_numpy(golden["ref_latent"])
    qry = torch.from_numpy(golden["query_latent"])
    poses = torch.from_numpy(golden["all_relativeR"])
    a = ldm_model.sweep(torch.cat([ref, qry]), torch.cat([poses, poses]), torch.cat([qry, ref]), want_emb=True, k=1)
    n_hoist = ldm_model.last_launch_count
    ldm_model.set_option("hoist", 0)
    try:
        b = ldm_model.sweep(torch.cat([ref, qry]), torch.cat([poses, poses]), torch.cat([qry, ref]), want_emb=True, k=1)
        n_plain = ldm_model.last_launch_count
    finally:
        ldm_model.set_option("hoist", 1)
    log("ldm_hoist", launches_hoisted=n_hoist, launches_plain=n_plain,
        max_abs_diff=float((a["emb"] - b["emb"]).abs().max()))
    assert torch.equal(a["emb"], b["emb"]) and torch.equal(a["sim"], b["sim"])


def test_epilogue_residual_and_narrow_tiles(ldm_sd, oracle_taps, golden):
    """the alternative GEMM configuration (residuals added in the epilogue, 128-wide tiles) against
    the oracle and against the default engine"""
    from nope_b200.ldm import UNetModelPose
    m = UNetModelPose(device="cuda:0", chunk=4)
    m.set_option("fold_residual", 0)
    m.set_option("wide_tiles", 0)
    m.load_state_dict(ldm_sd)
    ref = torch.from_numpy(golden["ref_latent"])
    emb = m(ref.expand(N_HYP, -1, -1, -1), oracle_taps["poses"])
    e = rel_l2(emb, oracle_taps["emb"])
    log("ldm_alt_gemm_config", emb_rel_l2=e)
    assert e < EMB_TOL_FP16


@pytest.fixture(scope="module")
def ldm_model_w2(ldm_sd):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from nope_b200.ldm import UNetModelPose
    m = UNetModelPose(device="cuda:0", chunk=8, precision="fp16_w2")
    m.load_state_dict(ldm_sd)
    return m


def test_exact_weights_mode_meets_the_north_star_tolerance(ldm_model_w2, oracle_taps, golden):
    """precision="fp16_w2": embeddings within 1e-3 of the fp32 oracle AND of the unmodified reference's golden
    sweep, scores and ranking as in the fp16 mode"""
    ref = torch.from_numpy(golden["ref_latent"])
    emb = ldm_model_w2(ref.expand(N_HYP, -1, -1, -1), oracle_taps["poses"])
    e = rel_l2(emb, oracle_taps["emb"])
    qry = torch.from_numpy(golden["query_latent"])
    poses = torch.from_numpy(golden["all_relativeR"])
    out = ldm_model_w2.sweep(ref, poses, qry, want_emb=True, k=3)
    e_emb = rel_l2(out["emb"][0], torch.from_numpy(golden["emb"]))
    e_sim = max_rel(out["sim"], torch.from_numpy(golden["similarity"]))
    order = torch.from_numpy(golden["similarity"]).argsort(dim=1, descending=True)
    log("ldm_exact_weights", emb_rel_l2_vs_oracle=e, emb_rel_l2_vs_golden=e_emb, sim_max_rel=e_sim,
        launches=ldm_model_w2.last_launch_count)
    assert e < EMB_TOL and e_emb < EMB_TOL and e_sim < SIM_TOL
    assert torch.equal(out["topi"].cpu(), order)


def test_exact_weights_module_in_isolation(ldm_model_w2, oracle_taps):
    """a ResBlock (3x3 convs + folded identity / 1x1 skip K-segments) and a SpatialTransformer (q|k|v, GEGLU,
    residual linears) with the doubled segment list"""
    for name, kind, in0, in1, out_tap in [("input_blocks.4.0", "res", "input_blocks.3", None, "input_blocks.4.0"),
                                          ("input_blocks.4.1", "st", "input_blocks.4.0", None, "input_blocks.4"),
                                          ("output_blocks.2.0", "res", "output_blocks.1", "input_blocks.6", "output_blocks.2.0")]:
        want = oracle_taps[out_tap]
        got = ldm_model_w2.run_block(name, oracle_taps[in0], oracle_taps[in1] if in1 else None,
                                     poses=oracle_taps["poses"] if kind == "st" else None,
                                     out_channels=want.shape[1], out_side=want.shape[2])
        e = rel_l2(got, want)
        log("ldm_module_w2", module=name, rel_l2=e)
        assert e < BLOCK_TOL


def test_rejects_bad_input(ldm_model):
    from nope_b200 import _lib
    with pytest.raises(_lib.NopeError):
        ldm_model.run_block("no.such.module", torch.zeros(1, 256, 32, 32), out_channels=256, out_side=32)
    from nope_b200.ldm import UNetModelPose
    with pytest.raises(ValueError):
        UNetModelPose(channel_mult=(1, 2, 4, 8))
