"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch fp32) of the reference's LDM-variant
pose-conditioned UNet, `UNetModelPose.forward`
(src/model/u_net/ldm/adapt_openaimodel.py:127-158), written functionally over a plain
state_dict so it needs neither the reference tree nor its module classes at run time.

Pinned by tests/test_ldm_oracle.py against outputs of the UNMODIFIED reference module
(`oracle/make_golden.py --only-ldm` imports it from /root/reference, loads the same seeded
weights with strict=True and stores the result under tests/golden/).  The reference ships no
tests or golden vectors of its own for this path.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
import torch
import torch.nn.functional as F

from nope_b200.synth_weights import ldm_block_plan


def _gn(x, sd, p, eps):
    # normalization() = GroupNorm32(32, C), computed in fp32 (ldm/util.py:187-204);
    # Normalize() = GroupNorm(32, C, eps=1e-6) (ldm/attention.py:73-76)
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps)


def res_block(x, sd, p):
    """ResBlock._forward with emb = 0 and use_scale_shift_norm = False
    (ldm/openaimodel.py:265-286): emb_layers(0) = Linear(SiLU(0)) = its bias."""
    h = F.conv2d(F.silu(_gn(x, sd, p + ".in_layers.0", 1e-5)), sd[p + ".in_layers.2.weight"],
                 sd[p + ".in_layers.2.bias"], padding=1)
    h = h + sd[p + ".emb_layers.1.bias"][None, :, None, None]
    h = F.conv2d(F.silu(_gn(h, sd, p + ".out_layers.0", 1e-5)), sd[p + ".out_layers.3.weight"],
                 sd[p + ".out_layers.3.bias"], padding=1)
    if (p + ".skip_connection.weight") in sd:
        x = F.conv2d(x, sd[p + ".skip_connection.weight"], sd[p + ".skip_connection.bias"])
    return x + h


def cross_attention(x, ctx, sd, p, heads):
    """CrossAttention.forward (ldm/attention.py:170-195); ctx=None => self-attention."""
    b, n, c = x.shape
    d = c // heads
    ctx = x if ctx is None else ctx
    q = x @ sd[p + ".to_q.weight"].t()
    k = ctx @ sd[p + ".to_k.weight"].t()
    v = ctx @ sd[p + ".to_v.weight"].t()
    split = lambda t: t.reshape(b, t.shape[1], heads, d).permute(0, 2, 1, 3)
    q, k, v = split(q), split(k), split(v)
    sim = torch.einsum("bhid,bhjd->bhij", q, k) * d ** -0.5
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bhij,bhjd->bhid", attn, v).permute(0, 2, 1, 3).reshape(b, n, c)
    return out @ sd[p + ".to_out.0.weight"].t() + sd[p + ".to_out.0.bias"]


def transformer_block(x, ctx, sd, p, heads):
    """BasicTransformerBlock._forward (ldm/attention.py:229-233), GEGLU feed-forward (:44-68)."""
    ln = lambda t, n: F.layer_norm(t, (t.shape[-1],), sd[f"{p}.{n}.weight"], sd[f"{p}.{n}.bias"])
    x = cross_attention(ln(x, "norm1"), None, sd, p + ".attn1", heads) + x
    x = cross_attention(ln(x, "norm2"), ctx, sd, p + ".attn2", heads) + x
    hgate = ln(x, "norm3") @ sd[p + ".ff.net.0.proj.weight"].t() + sd[p + ".ff.net.0.proj.bias"]
    h, gate = hgate.chunk(2, dim=-1)
    h = h * F.gelu(gate)
    return h @ sd[p + ".ff.net.2.weight"].t() + sd[p + ".ff.net.2.bias"] + x


def spatial_transformer(x, ctx, sd, p):
    """SpatialTransformer.forward (ldm/attention.py:264-277); heads = C / 32
    (openaimodel.py:556-566 with num_head_channels = 32, legacy = True)."""
    b, c, h, w = x.shape
    x_in = x
    x = F.conv2d(_gn(x, sd, p + ".norm", 1e-6), sd[p + ".proj_in.weight"], sd[p + ".proj_in.bias"])
    x = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
    x = transformer_block(x, ctx, sd, p + ".transformer_blocks.0", c // 32)
    x = x.reshape(b, h, w, c).permute(0, 3, 1, 2)
    return F.conv2d(x, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"]) + x_in


def ldm_forward(sd, x, pose, taps=None, model_channels=256, channel_mult=(1, 2, 4)):
    """UNetModelPose.forward (adapt_openaimodel.py:127-158) with
    injecting_condition_twice = False: emb = 0, context = pose_mlp(pose)[:, None]."""
    inp, _, out = ldm_block_plan(model_channels, channel_mult)

    def tap(name, t):
        if taps is not None:
            taps[name] = t.detach().clone()

    ctx = (pose @ sd["pose_mlp.0.weight"].t() + sd["pose_mlp.0.bias"])[:, None, :]
    hs = []
    h = x
    for i, b in enumerate(inp):
        p = f"input_blocks.{i}"
        if b[0] == "conv":
            h = F.conv2d(h, sd[p + ".0.weight"], sd[p + ".0.bias"], padding=1)
        elif b[0] == "res":
            h = res_block(h, sd, p + ".0")
            tap(p + ".0", h)
            h = spatial_transformer(h, ctx, sd, p + ".1")
        else:
            h = F.conv2d(h, sd[p + ".0.op.weight"], sd[p + ".0.op.bias"], stride=2, padding=1)
        tap(p, h)
        hs.append(h)
    h = res_block(h, sd, "middle_block.0")
    tap("middle_block.0", h)
    h = spatial_transformer(h, ctx, sd, "middle_block.1")
    tap("middle_block.1", h)
    h = res_block(h, sd, "middle_block.2")
    tap("middle_block", h)
    for i, b in enumerate(out):
        p = f"output_blocks.{i}"
        h = torch.cat([h, hs.pop()], dim=1)
        h = res_block(h, sd, p + ".0")
        tap(p + ".0", h)
        h = spatial_transformer(h, ctx, sd, p + ".1")
        tap(p + ".1", h)
        if b[4]:
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            h = F.conv2d(h, sd[p + ".2.conv.weight"], sd[p + ".2.conv.bias"], padding=1)
        tap(p, h)
    h = F.conv2d(F.silu(_gn(h, sd, "out.0", 1e-5)), sd["out.2.weight"], sd["out.2.bias"], padding=1)
    tap("out", h)
    return h


def ldm_sweep(sd, ref_latent, poses, chunk=4, **kw):
    """All hypotheses of a sweep: ref_latent [B,4,32,32], poses [B,N,6] -> [B,N,4,32,32]
    (the batched form of the loop at src/model/model.py:212-222 for this UNet)."""
    B, N = poses.shape[:2]
    outs = []
    with torch.no_grad():
        for b in range(B):
            row = []
            for n0 in range(0, N, chunk):
                pz = poses[b, n0:n0 + chunk]
                row.append(ldm_forward(sd, ref_latent[b:b + 1].expand(pz.shape[0], -1, -1, -1), pz, **kw))
            outs.append(torch.cat(row, 0))
    return torch.stack(outs, 0)
