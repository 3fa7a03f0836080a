"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/* from the UNMODIFIED
reference (imported read-only from /root/reference through oracle/ref_import.py)
and pin the restatement in oracle/unet_oracle.py against it.

Run in the build container (CPU):   python -m oracle.make_golden
Writes:
  tests/golden/pose_grids.npz   3x3 object rotations of the shipped icosphere grids
                                (src/poses/predefined_poses/obj_poses_level{0,2}.npy,
                                 'all' and 'upper' per src/poses/utils.py:72-102)
  tests/golden/cfg1_b1_n6.npz   BASELINE config 1 (N=6 because retrieval() hard-codes
                                topk(5), model.py:265): feats, embeddings, scores, top-5
  tests/golden/grid26_b2.npz    B=2, N=26 'upper' level-0 grid: scores, top-5, embedding
                                digests + two full templates
  tests/golden/unet_taps.npz    per-layer activation statistics of one UNet forward
  tests/golden/ldm_b1_n3.npz    (--only-ldm) LDM-variant UNet (UNetModelPose): 3 hypotheses, embeddings,
                                scores, per-block activation statistics
  tests/golden/cfg2_b8_n2562.npz (--only-cfg2) BASELINE configs[2] at full size: 8 queries x the 2562-pose
                                level-3 grid: scores, top-5, embedding norms, three full templates
  tests/golden/meta.json        weight checksum, torch version, oracle-vs-reference errors
It asserts (hard failure) that
  * the seeded state_dict loads into the reference modules with strict=True
  * oracle.unet_forward / encode_image / l2_similarity match the reference modules
    to fp32 round-off (different summation order only)
"""
import json
import os
import sys
import time

import numpy as np
import torch

from . import inputs, unet_oracle as orc, weights
from .ref_import import REFERENCE_ROOT, build_reference_model, import_reference

OUT = inputs.GOLDEN_DIR


def rel_err(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def write_pose_grids():
    pp = os.path.join(REFERENCE_ROOT, "src/poses/predefined_poses")
    out = {}
    for level in (0, 2):
        obj = np.load(os.path.join(pp, f"obj_poses_level{level}.npy"))
        cam = np.load(os.path.join(pp, f"sphere_poses_level{level}.npy"))
        out[f"level{level}_all"] = obj[:, :3, :3].astype(np.float64)
        out[f"level{level}_upper"] = obj[cam[:, 2, 3] >= 0][:, :3, :3].astype(np.float64)
    np.savez_compressed(os.path.join(OUT, "pose_grids.npz"), **out)
    return {k: v.shape[0] for k, v in out.items()}


def write_geodesic_fixture():
    """tests/golden/geodesic.npz: the reference's GeodesicError (src/model/loss.py:74-115) on
    seeded random rotations with all three symmetry classes.  pytorch3d is absent, so its
    so3_relative_angle is the restatement in oracle/ref_import.py: this pins the reference's OWN
    symmetry / top-k logic, not pytorch3d's acos extrapolation near 0 and 180 degrees."""
    ref = import_reference()
    from src.model.loss import GeodesicError
    g = torch.Generator().manual_seed(4)
    q = torch.randn(24 * 5 + 24, 4, generator=g, dtype=torch.float64)
    R = ref.rotation_conversions.quaternion_to_matrix(q / q.norm(dim=1, keepdim=True))
    predR, gtR = R[:120].reshape(24, 5, 3, 3), R[120:]
    sym = torch.tensor([0, 1, 2] * 8, dtype=torch.float32)
    out = {"predR": predR.numpy(), "gtR": gtR.numpy(), "symmetry": sym.numpy()}
    for name, s in (("sym0", torch.zeros(24)), ("mixed", sym)):
        err, res = GeodesicError()(predR.clone(), gtR.clone(), s)
        out[f"{name}_err"] = err.double().numpy()
        for k, v in res.items():
            out[f"{name}|{k}"] = np.array(float(v))
    np.savez_compressed(os.path.join(OUT, "geodesic.npz"), **out)
    return {k: float(v) for k, v in out.items() if "|" in k}


def write_full_grid_fixture():
    """tests/golden/level2_642_b1.npz: BASELINE configs[1] at full size -- the reference's own
    UNet + encoder modules (batched 16 hypotheses per forward, which model.py:212-222's loop is
    arithmetically equal to) over the shipped 642-pose level-2 grid, its "l2" similarity
    (model.py:260-262) and topk(5)."""
    model = build_reference_model()
    sd = weights.make_full_state_dict(seed=0)
    model.u_net.load_state_dict(sd, strict=True)
    q, r = inputs.make_images(seed=2, batch=1)
    relR, tposes = inputs.make_pose_batch("level2_all", batch=1)
    with torch.no_grad():
        qf = model.u_net.encoder.encode_image(q)
        rf = model.u_net.encoder.encode_image(r)
        embs = []
        t0 = time.time()
        for s0 in range(0, 642, 16):
            p = relR[0, s0:s0 + 16]
            embs.append(model.u_net(rf.expand(p.shape[0], -1, -1, -1), p))
        emb = torch.cat(embs)[None]
        d = (qf.unsqueeze(1).repeat(1, 642, 1, 1, 1) - emb) ** 2
        sim = -torch.norm(d, dim=2).sum(axis=3).sum(axis=2)
        _, idx = sim.topk(k=5, dim=1)
        secs = time.time() - t0
    srt = torch.sort(sim, dim=1, descending=True).values
    gap = float(((srt[:, :-1] - srt[:, 1:]) / srt[:, :-1].abs())[:, :5].min())
    np.savez_compressed(os.path.join(OUT, "level2_642_b1.npz"), query_feat=qf.numpy(), ref_feat=rf.numpy(),
                        all_relativeR=relR.numpy(), similarity=sim.numpy(), nearest_idx=idx.numpy(),
                        emb_n0=emb[0, 0].numpy(), emb_n641=emb[0, 641].numpy())
    return {"reference_cpu_seconds": secs, "top5": idx.tolist(), "min_rel_gap_top6": gap}


def write_cfg2_fixture(batch=8):
    """tests/golden/cfg2_b8_n2562.npz: BASELINE configs[2] at FULL size -- 8 queries x the shipped
    2562-pose level-3 grid (src/poses/predefined_poses/obj_poses_level3.npy) through the reference's own
    encoder / UNet modules (16 hypotheses per forward), "l2" similarity (model.py:260-262) and topk(5).
    20 496 reference forwards: ~40 minutes on 8 cores.  Stored: latents, the grid, similarity rows, top-5,
    per-hypothesis embedding norms and three full templates."""
    pp = os.path.join(REFERENCE_ROOT, "src/poses/predefined_poses")
    R3 = np.load(os.path.join(pp, "obj_poses_level3.npy"))[:, :3, :3].astype(np.float64)
    n = R3.shape[0]
    ref_idx = [(7 + 311 * b) % n for b in range(batch)]
    relR = torch.stack([inputs.relative_rot6d(R3, R3[i]) for i in ref_idx])          # [B, n, 6]
    model = build_reference_model()
    sd = weights.make_full_state_dict(seed=0)
    model.u_net.load_state_dict(sd, strict=True)
    q, r = inputs.make_images(seed=7, batch=batch)
    keep = [(0, 0), (3, 1000), (batch - 1, n - 1)]
    kept = {}
    with torch.no_grad():
        qf = torch.cat([model.u_net.encoder.encode_image(q[b:b + 1]) for b in range(batch)])
        rf = torch.cat([model.u_net.encoder.encode_image(r[b:b + 1]) for b in range(batch)])
        sim = torch.zeros(batch, n)
        l2n = torch.zeros(batch, n)
        t0 = time.time()
        for b in range(batch):
            for s0 in range(0, n, 16):
                p = relR[b, s0:s0 + 16]
                emb = model.u_net(rf[b:b + 1].expand(p.shape[0], -1, -1, -1), p)        # [16, 8, 32, 32]
                d = (qf[b:b + 1] - emb) ** 2
                sim[b, s0:s0 + p.shape[0]] = -torch.norm(d, dim=1).sum(dim=(1, 2))
                l2n[b, s0:s0 + p.shape[0]] = emb.flatten(1).norm(dim=1)
                for (kb, kn) in keep:
                    if kb == b and s0 <= kn < s0 + p.shape[0]:
                        kept[f"emb_b{kb}_n{kn}"] = emb[kn - s0].numpy().copy()
            print(f"cfg2: query {b} done, {time.time() - t0:.0f} s", flush=True)
        _, idx = sim.topk(k=5, dim=1)
        secs = time.time() - t0
    srt = torch.sort(sim, dim=1, descending=True).values
    gap = ((srt[:, :-1] - srt[:, 1:]) / srt[:, :-1].abs())[:, :5]
    np.savez_compressed(os.path.join(OUT, "cfg2_b8_n2562.npz"), query_feat=qf.numpy(), ref_feat=rf.numpy(),
                        level3_all=R3, ref_idx=np.array(ref_idx), similarity=sim.numpy(), nearest_idx=idx.numpy(),
                        emb_l2=l2n.numpy(), **kept)
    return {"reference_cpu_seconds": secs, "top5": idx.tolist(), "min_rel_gap_top6": float(gap.min()),
            "hypotheses": batch * n}


def make_ldm_inputs(seed, batch, n):
    """Seeded LDM-variant inputs: VAE-like latents N(0,1) [B,4,32,32] (the diffusers VAE itself
    is absent, see ref_import.build_reference_ldm_unet) and [B,n,6] level-0 relative poses."""
    g = torch.Generator(device="cpu").manual_seed(5000 + seed)
    ref_lat = torch.randn((batch, 4, 32, 32), generator=g)
    query_lat = torch.randn((batch, 4, 32, 32), generator=g)
    relR, _ = inputs.make_pose_batch("level0_upper", batch=batch, n=n)
    return ref_lat, query_lat, relR


def write_ldm_fixture():
    """tests/golden/ldm_b1_n3.npz: three hypotheses through the UNMODIFIED UNetModelPose
    (src/model/u_net/ldm/adapt_openaimodel.py:127-158, wired as configs/model/vae_cin_ldm.yaml)
    with the seeded weights of nope_b200.synth_weights.make_ldm_state_dict loaded strict=True,
    scored with the reference's l2 formula (model.py:259-264 restated in unet_oracle), and the
    restatement in oracle/ldm_oracle.py checked against it."""
    from . import ldm_oracle
    from .ref_import import build_reference_ldm_unet
    torch.set_num_threads(os.cpu_count())
    t0 = time.time()
    m = build_reference_ldm_unet()
    sd = weights.make_ldm_state_dict(seed=0)
    print("load_state_dict strict:", m.load_state_dict(sd, strict=True), f"{time.time() - t0:.0f}s")
    ref_lat, query_lat, relR = make_ldm_inputs(0, 1, 3)
    with torch.no_grad():
        t0 = time.time()
        emb_ref = m(ref_lat.expand(3, -1, -1, -1), relR[0])
        secs = time.time() - t0
        taps = {}
        emb = ldm_oracle.ldm_forward(sd, ref_lat.expand(3, -1, -1, -1), relR[0], taps=taps)
    err = rel_err(emb, emb_ref)
    sim = orc.l2_similarity(query_lat, emb_ref[None])
    tap_stats = {"tap:" + k: np.array([float(v.mean()), float(v.std()), float(v.abs().max())])
                 for k, v in taps.items()}
    np.savez_compressed(os.path.join(OUT, "ldm_b1_n3.npz"), ref_latent=ref_lat.numpy(),
                        query_latent=query_lat.numpy(), all_relativeR=relR.numpy(),
                        emb=emb_ref.numpy(), similarity=sim.numpy(), **tap_stats)
    out = {"oracle_vs_reference_rel_err": err, "reference_cpu_seconds_3hyp": secs,
           "weights_checksum_seed0": weights.checksum(sd), "emb_std": float(emb_ref.std()),
           "emb_absmax": float(emb_ref.abs().max())}
    assert err < 2e-5, f"ldm oracle disagrees with the reference: {err}"
    return out


def main():
    if "--only-ldm" in sys.argv:
        print(json.dumps(write_ldm_fixture(), indent=1))
        return
    if "--only-full-grid" in sys.argv:
        torch.set_num_threads(os.cpu_count())
        print(json.dumps(write_full_grid_fixture(), indent=1))
        return
    if "--only-cfg2" in sys.argv:
        torch.set_num_threads(int(os.environ.get("NOPE_GOLDEN_THREADS", os.cpu_count())))
        print(json.dumps(write_cfg2_fixture(), indent=1))
        return
    if "--only-geodesic" in sys.argv:
        print(json.dumps(write_geodesic_fixture(), indent=1))
        return
    torch.set_num_threads(os.cpu_count())
    os.makedirs(OUT, exist_ok=True)
    meta = {"torch": torch.__version__, "reference_root": REFERENCE_ROOT}
    meta["pose_grids"] = write_pose_grids()

    ref = import_reference()
    model = build_reference_model()
    sd = weights.make_full_state_dict(seed=0)
    missing = model.u_net.load_state_dict(sd, strict=True)
    meta["weights_checksum_seed0"] = weights.checksum(sd)
    meta["n_unet_tensors"] = sum(1 for k in sd if not k.startswith("encoder."))
    meta["n_encoder_entries"] = sum(1 for k in sd if k.startswith("encoder."))
    print("state_dict loaded strict:", missing, meta["n_unet_tensors"], meta["n_encoder_entries"])
    unet_sd = {k: v for k, v in sd.items() if not k.startswith("encoder.")}
    enc_sd = {k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}

    errs = {}
    with torch.no_grad():
        # ---------------- cfg1: B=1, N=6 through the reference's own loop ----------
        q, r = inputs.make_images(seed=0, batch=1)
        relR, tposes = inputs.make_pose_batch("level0_upper", batch=1, n=6)
        t0 = time.time()
        emb_ref, _, _ = model.generate_templates(r, relR, gt_templates=None)
        sim_ref, idx_ref = model.retrieval(q, emb_ref)
        meta["cfg1_reference_seconds"] = time.time() - t0
        qf_ref = model.u_net.encoder.encode_image(q)
        rf_ref = model.u_net.encoder.encode_image(r)
        # oracle restatement on the same inputs
        qf = orc.encode_image(enc_sd, q)
        rf = orc.encode_image(enc_sd, r)
        emb = orc.generate_templates(unet_sd, rf, relR)
        sim = orc.l2_similarity(qf, emb)
        idx = orc.topk_lowest_index(sim, 5)
        errs["cfg1_query_feat"] = rel_err(qf, qf_ref)
        errs["cfg1_ref_feat"] = rel_err(rf, rf_ref)
        errs["cfg1_emb"] = rel_err(emb, emb_ref)
        errs["cfg1_sim"] = rel_err(sim, sim_ref)
        assert torch.equal(idx, idx_ref), (idx, idx_ref)
        # rotation_conversions.matrix_to_rotation_6d cross-check of inputs.relative_rot6d
        fx = inputs.load_pose_fixture()
        Rrel = torch.tensor(fx["level0_upper"][:6] @ np.linalg.inv(fx["level0_upper"][7]),
                            dtype=torch.float32)
        assert torch.equal(ref.rotation_conversions.matrix_to_rotation_6d(Rrel), relR[0])
        np.savez_compressed(
            os.path.join(OUT, "cfg1_b1_n6.npz"),
            query_feat=qf_ref.numpy(), ref_feat=rf_ref.numpy(), all_relativeR=relR.numpy(),
            emb=emb_ref.numpy(), similarity=sim_ref.numpy(), nearest_idx=idx_ref.numpy(),
            template_poses=tposes.numpy())

        # ---------------- unet taps: one hypothesis, oracle vs reference module -----
        taps = {}
        y = orc.unet_forward(unet_sd, rf_ref, relR[:, 0], taps=taps)
        y_ref = model.u_net(rf_ref, relR[:, 0])
        errs["unet_single"] = rel_err(y, y_ref)
        tap_stats = {k: np.array([float(v.mean()), float(v.std()), float(v.abs().max())],
                                 dtype=np.float64) for k, v in taps.items()}
        tap_stats["out"] = np.array([float(y_ref.mean()), float(y_ref.std()),
                                     float(y_ref.abs().max())])
        np.savez_compressed(os.path.join(OUT, "unet_taps.npz"), **tap_stats)

        # ---------------- grid26: B=2, N=26 ----------------------------------------
        q2, r2 = inputs.make_images(seed=1, batch=2)
        relR2, tposes2 = inputs.make_pose_batch("level0_upper", batch=2)
        t0 = time.time()
        emb2, _, _ = model.generate_templates(r2, relR2, gt_templates=None)
        sim2, idx2 = model.retrieval(q2, emb2)
        meta["grid26_reference_seconds"] = time.time() - t0
        qf2 = model.u_net.encoder.encode_image(q2)
        rf2 = model.u_net.encoder.encode_image(r2)
        emb2o = orc.generate_templates(unet_sd, orc.encode_image(enc_sd, r2), relR2)
        sim2o = orc.l2_similarity(orc.encode_image(enc_sd, q2), emb2o)
        errs["grid26_emb"] = rel_err(emb2o, emb2)
        errs["grid26_sim"] = rel_err(sim2o, sim2)
        assert torch.equal(orc.topk_lowest_index(sim2o, 5), idx2)
        srt = torch.sort(sim2, dim=1, descending=True).values
        meta["grid26_min_adjacent_gap_rel"] = float(
            ((srt[:, :-1] - srt[:, 1:]) / srt[:, :-1].abs()).min())
        np.savez_compressed(
            os.path.join(OUT, "grid26_b2.npz"),
            query_feat=qf2.numpy(), ref_feat=rf2.numpy(), all_relativeR=relR2.numpy(),
            similarity=sim2.numpy(), nearest_idx=idx2.numpy(),
            emb_mean=emb2.mean(dim=(2, 3, 4)).numpy(), emb_l2=emb2.flatten(2).norm(dim=2).numpy(),
            emb_b0_n0=emb2[0, 0].numpy(), emb_b1_n25=emb2[1, 25].numpy(),
            template_poses=tposes2.numpy())

    meta["geodesic"] = write_geodesic_fixture()
    meta["level2_642"] = write_full_grid_fixture()
    meta["oracle_vs_reference_rel_err"] = errs
    print(json.dumps(meta, indent=1))
    for k, v in errs.items():
        assert v < 2e-5, f"oracle restatement disagrees with the reference on {k}: {v}"
    with open(os.path.join(OUT, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1)


if __name__ == "__main__":
    sys.exit(main())
