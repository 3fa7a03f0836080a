#!/usr/bin/env python
"""bench.py -- pose-hypotheses/s of the NOPE hot path on B200.

A "step" = one pass of the hot path for one query: the pose-conditioned UNet over the
whole pose grid + l2 scoring + top-5 (BASELINE.json configs[1]: 256x256, 642-pose
icosphere grid, batch = 1 query, fp16 UNet).  With N GPUs the grid is sharded
(weak scaling: 642 poses per GPU, global grid = 642 N) and the only collective is the
all-gather of per-shard (score, index) top-k.

  value  hypotheses/s with the encoder latents and poses already resident in HBM
  e2e    the same metric through the public API (PoseConditional.predict_pose): pinned HOST
         images + poses -> H2D -> encoder x2 -> sweep -> fused score/top-k -> D2H result
  roofline  tensor-core convolution kernel: algorithmic FLOPs / CUDA-event launch time
  cpu_baseline  the oracle (CPU port of the reference) on the host cores, bounded sample

`--impl reference` times the reference's own CPU implementation of the path (the real
reference modules when /root/reference is mounted, else the oracle port of them).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_POSES = 642
GFLOP_PER_HYP = 35.05       # SURVEY.md 8d / BASELINE.md section 3
METRIC = "pose-hypotheses/sec @256x256 (UNet sweep + l2 score + top-5)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--poses", type=int, default=N_POSES, help="poses per GPU")
    ap.add_argument("--queries", type=int, default=1, help="queries (batch) per step")
    ap.add_argument("--chunk", type=int, default=int(os.environ.get("NOPE_CHUNK", "642")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variant", default="default", choices=["default", "ldm"],
                    help="default: template_base UNet (BASELINE configs); ldm: the LDM-variant UNetModelPose "
                         "sweep on latents (SURVEY.md 8 f2), an additional line for profiles/")
    ap.add_argument("--conv-impl", default=os.environ.get("NOPE_CONV_IMPL", "tcgen05_2cta"),
                    choices=["tcgen05", "tcgen05_2cta"])
    return ap.parse_args()


# ---------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.idx), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 8:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except ValueError:
                continue
            for n, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm),
                "reasons": sorted(reasons)}


def ncu_traffic():
    """DRAM bytes per launch of the convolution kernel from the committed ncu capture
    (profiles/roofline_traffic.json, written by tools/summarize_ncu_raw.py), or None."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops")), d.get("hbm_gbs"), "measured"
    return 1400.0, 6650.0, "fallback"     # B200_PROFILING.md fallback (sustained)


# ---------------------------------------------------------------------------------------
def pick_threads(fn, candidates=(16, 32, 64)):
    """Host boxes with >100 cores run torch-CPU convs slower at full thread count than at a
    NUMA-friendly one; probe a few counts on a tiny workload and keep the fastest."""
    import torch
    n = os.cpu_count()
    cands = sorted({min(c, n) for c in candidates} | {n})
    best, best_t = n, None
    for c in cands:
        torch.set_num_threads(c)
        fn()
        t0 = time.time()
        fn()
        dt = time.time() - t0
        if best_t is None or dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def cpu_baseline(seconds=12.0, chunk=16):
    """Oracle (CPU port of the reference path) on all host cores: UNet sweep in chunks of
    16 hypotheses + scoring, until ~`seconds` of work.  hyp/s."""
    import torch
    from oracle import inputs, unet_oracle as orc, weights
    sd = weights.make_unet_state_dict(seed=0)
    g = torch.Generator().manual_seed(0)
    rf = torch.randn(1, 8, 32, 32, generator=g) * 1.5
    qf = torch.randn(1, 8, 32, 32, generator=g) * 1.5
    from nope_b200.poses import synthetic_pose_batch
    poses, _ = synthetic_pose_batch(N_POSES, 1)
    done, t0 = 0, time.time()
    with torch.no_grad():
        threads = pick_threads(lambda: orc.generate_templates(sd, rf, poses[:, :4], chunk=4))
        t0 = time.time()
        while time.time() - t0 < seconds and done + chunk <= N_POSES:
            emb = orc.generate_templates(sd, rf, poses[:, done:done + chunk], chunk=chunk)
            orc.l2_similarity(qf, emb)
            done += chunk
    dt = time.time() - t0
    return {"value": done / dt, "unit": "hyp/s", "cores": threads, "kind": "port",
            "sample": f"first {done} of the {N_POSES}-pose grid, batched {chunk}/forward, "
                      f"fp32 torch-CPU oracle, {threads} of {os.cpu_count()} host threads "
                      f"(fastest of a probe), {dt:.1f} s"}


def eager_gpu_baseline(dev, chunk=64, reps=3):
    """the reference ships no custom kernel, so the on-box GPU baseline is the same
    module in PyTorch eager (cuDNN/cuBLAS), fp16.  /root/reference does not exist on the GPU box:
    the oracle's torch restatement of UNet.forward runs on CUDA half tensors instead, batched
    `chunk` hypotheses per forward.  hyp/s (UNet + l2 score only, no encoder)."""
    import torch
    from oracle import unet_oracle as orc, weights
    from nope_b200.poses import synthetic_pose_batch
    sd = {k: v.to(dev, torch.float16) for k, v in weights.make_unet_state_dict(seed=0).items()}
    g = torch.Generator().manual_seed(0)
    rf = (torch.randn(1, 8, 32, 32, generator=g) * 1.5).to(dev, torch.float16)
    qf = (torch.randn(1, 8, 32, 32, generator=g) * 1.5).to(dev)
    poses, _ = synthetic_pose_batch(N_POSES, 1)
    poses = poses.to(dev, torch.float16)
    x = rf.expand(chunk, -1, -1, -1).contiguous(memory_format=torch.channels_last)

    def run():
        with torch.no_grad():
            emb = orc.unet_forward(sd, x, poses[0, :chunk])
            orc.l2_similarity(qf, emb.float()[None])
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return {"value": chunk / (ms * 1e-3), "unit": "hyp/s", "kind": "oracle port, torch-eager CUDA fp16 "
            "channels_last (cuDNN/cuBLAS)", "sample": f"{chunk} hypotheses per forward, {reps} forwards, "
            "UNet + l2 score"}


def run_reference(args):
    """--impl reference: the reference's own CPU path, bounded sample per step."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 16
    from oracle import ref_import, unet_oracle as orc, weights
    from nope_b200.poses import synthetic_pose_batch
    poses, _ = synthetic_pose_batch(N_POSES, 1)
    g = torch.Generator().manual_seed(0)
    q = torch.rand(1, 3, 256, 256, generator=g) * 2 - 1
    r = torch.rand(1, 3, 256, 256, generator=g) * 2 - 1
    sd = weights.make_full_state_dict(seed=0)
    if ref_import.reference_available():
        kind = "reference"
        model = ref_import.build_reference_model()
        model.u_net.load_state_dict(sd, strict=True)

        def step():
            # the reference's retrieval path with the grid batched along dim 0 (its own
            # modules, best-effort CPU: one encoder call per image, UNet at batch `sample`)
            with torch.no_grad():
                qf = model.u_net.encoder.encode_image(q)
                rf = model.u_net.encoder.encode_image(r)
                emb = model.u_net(rf.expand(sample, -1, -1, -1), poses[0, :sample])[None]
                d = (qf.unsqueeze(1) - emb) ** 2
                sim = -torch.norm(d, dim=2).sum(3).sum(2)
                sim.topk(k=5, dim=1)
    else:
        kind = "port"
        unet_sd = {k: v for k, v in sd.items() if not k.startswith("encoder.")}
        enc_sd = {k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}

        def step():
            with torch.no_grad():
                qf = orc.encode_image(enc_sd, q)
                rf = orc.encode_image(enc_sd, r)
                emb = orc.generate_templates(unet_sd, rf, poses[:, :sample], chunk=sample)
                orc.topk_lowest_index(orc.l2_similarity(qf, emb), 5)
    with torch.no_grad():
        small = poses[0, :4]
        if kind == "reference":
            rf0 = torch.randn(4, 8, 32, 32)
            threads = pick_threads(lambda: model.u_net(rf0, small))
        else:
            rf0 = torch.randn(1, 8, 32, 32)
            threads = pick_threads(lambda: orc.generate_templates(unet_sd, rf0, poses[:, :4], chunk=4))
    for _ in range(args.warmup):
        step()
    t0 = time.time()
    for _ in range(args.steps):
        step()
    dt = (time.time() - t0) / args.steps
    v = sample / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "hyp/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"configs[1]: 256x256, {N_POSES}-pose icosphere grid per GPU, batch=1 query, "
                               "fp32 reference modules on the host CPU, l2 score + top-5",
                   "sample": f"each step = a {sample}-pose sample of the grid + 2 encoder calls",
                   "poses_per_step": sample},
        "cpu_baseline": {"value": v, "unit": "hyp/s", "cores": threads, "kind": kind,
                         "sample": f"{sample} poses + 2 encoder calls per step, {threads} of "
                                   f"{os.cpu_count()} host threads (fastest of a probe)"},
        "e2e": {"value": v, "unit": "hyp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------
def ldm_cpu_baseline(seconds=12.0, chunk=4):
    """oracle/ldm_oracle.py (CPU port of UNetModelPose.forward) + the l2 score, hyp/s."""
    import torch
    from oracle import ldm_oracle, unet_oracle as orc
    from nope_b200.synth_weights import make_ldm_state_dict
    from nope_b200.poses import synthetic_pose_batch
    sd = make_ldm_state_dict(seed=0)
    g = torch.Generator().manual_seed(0)
    rl = torch.randn(1, 4, 32, 32, generator=g)
    ql = torch.randn(1, 4, 32, 32, generator=g)
    poses, _ = synthetic_pose_batch(N_POSES, 1)
    done = 0
    with torch.no_grad():
        threads = pick_threads(lambda: ldm_oracle.ldm_sweep(sd, rl, poses[:, :2], chunk=2))
        t0 = time.time()
        while time.time() - t0 < seconds and done + chunk <= N_POSES:
            emb = ldm_oracle.ldm_sweep(sd, rl, poses[:, done:done + chunk], chunk=chunk)
            orc.l2_similarity(ql, emb)
            done += chunk
    dt = time.time() - t0
    return {"value": done / dt, "unit": "hyp/s", "cores": threads, "kind": "port",
            "sample": f"first {done} poses of the {N_POSES}-pose grid, batched {chunk}/forward, fp32 torch-CPU "
                      f"oracle of UNetModelPose, {threads} of {os.cpu_count()} host threads, {dt:.1f} s"}


def main_ldm(args):
    """LDM-variant sweep (UNetModelPose on VAE-sized latents): same metric and timing rules as main()."""
    import torch
    from nope_b200.ldm import UNetModelPose
    from nope_b200.poses import synthetic_pose_batch
    from nope_b200.synth_weights import ldm_flops_per_hyp, make_ldm_state_dict
    import torch.distributed as dist
    from nope_b200.dist import ShardedSweep
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    sharded = ShardedSweep() if world > 1 else None
    n_local = args.poses
    n = n_local * world                      # weak scaling: one grid per GPU shard
    chunk = min(args.chunk, 642)
    m = UNetModelPose(device=str(dev), chunk=chunk)
    m.load_state_dict(make_ldm_state_dict(seed=0))
    Q = args.queries
    poses, _ = synthetic_pose_batch(n_local, Q)
    poses = poses.repeat(1, world, 1)
    g = torch.Generator().manual_seed(0)
    ref_h = torch.randn(Q, 4, 32, 32, generator=g).pin_memory()
    qry_h = torch.randn(Q, 4, 32, 32, generator=g).pin_memory()
    poses_h = poses.clone().pin_memory()
    ref_d, qry_d, poses_d = ref_h.to(dev), qry_h.to(dev), poses.to(dev)
    h2d = (ref_h.numel() + qry_h.numel() + poses_h.numel()) * 4
    d2h = Q * (5 * 8 + n * 4)

    def run(ref, pz, qry):
        if sharded is not None:
            sim, topi, _ = sharded.sweep(m, ref, pz, qry, k=5, want_emb=False)
            return {"sim": sim, "topi": topi}
        return m.sweep(ref, pz, qry, want_emb=False, k=5)

    def step_resident():
        return run(ref_d, poses_d, qry_d)

    def step_e2e():
        out = run(ref_h.to(dev, non_blocking=True), poses_h.to(dev, non_blocking=True),
                  qry_h.to(dev, non_blocking=True))
        return out["topi"].cpu(), out["sim"].cpu()

    def barrier():
        if world > 1:
            dist.barrier()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        barrier()
        t = e0.elapsed_time(e1) / steps
        if world > 1:
            tt = torch.tensor([t], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt)
        return t

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(step_resident, args.steps, max(args.warmup, 3))
    launches = m.last_launch_count
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed(step_e2e, args.steps, max(args.warmup, 3))
    m.profile(True)
    step_resident()
    prof = m.profile_read()
    m.profile(False)
    peak_tf, _, peak_src = measured_peaks()
    gm, at = prof["gemm"], prof["attention"]
    gemm_tf = gm["flops"] / (gm["ms"] * 1e-3) / 1e12 if gm["ms"] > 0 else 0.0
    attn_tf = at["flops"] / (at["ms"] * 1e-3) / 1e12 if at["ms"] > 0 else 0.0
    fl = ldm_flops_per_hyp()
    value = Q * n / (ms * 1e-3)
    if rank != 0:
        dist.destroy_process_group()
        return
    cpu = None if (args.no_cpu_baseline or world > 1) else ldm_cpu_baseline()
    line = {
        "metric": METRIC, "value": value, "unit": "hyp/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
        "config": {
            "workload": f"LDM variant (SURVEY.md 8 f2): UNetModelPose of configs/model/vae_cin_ldm.yaml on 4x32x32 "
                        f"latents, {n_local}-pose grid per GPU, batch={Q} query, fp16 storage / fp32 accumulate, "
                        "l2 score + top-5",
            "poses_per_gpu": n_local, "global_poses": n, "queries": Q, "chunk": chunk,
            "parallelism": f"pose grid sharded {world}-way, all-gather of top-k" if world > 1 else "1 GPU",
            "weights": "seeded random init, reference state_dict schema (395.0 M params)",
            "gflop_per_hyp": fl["total"] / 1e9,
            "encoder": "none: the diffusers VAE of this variant is not in the reference tree; inputs are latents",
            "l2": "not flushed: each step streams 0.79 GB of fp16 weights and > 5 GB of activations, >> 126 MB L2",
        },
        "clocks": clocks,
        "e2e": {"value": Q * n / (ms_e2e * 1e-3), "unit": "hyp/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "UNetModelPose.sweep (pinned host latents + poses -> sweep -> top-5 -> host)"},
        "gpu_launches": int(launches * args.steps),
        "roofline": {
            "bound": "tensor", "kernel": "conv_tc2_kernel (tcgen05 implicit-GEMM conv / linear layers)",
            "achieved": gemm_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": gemm_tf / peak_tf if peak_tf else None,
            "peak_source": f"{peak_src} (sustained bf16 cuBLAS)", "traffic": None,
            "launches_per_step": gm["launches"], "gemm_ms_per_step": gm["ms"],
            "gemm_share_of_step": gm["ms"] / ms if ms else None,
            "algorithmic_tflop_per_step": gm["flops"] / 1e12,
            "attention": {"kernel": "ldm_attn_tc_kernel (tcgen05 QK^T / PV, softmax in registers)",
                          "achieved": attn_tf, "unit": "TFLOP/s", "ms_per_step": at["ms"],
                          "launches_per_step": at["launches"], "share_of_step": at["ms"] / ms if ms else None},
            "whole_step_tflops": value * fl["total"] / 1e12,
        },
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.variant == "ldm":
        if args.impl == "reference":
            line = {"impl": "reference", "metric": METRIC, "unit": "hyp/s", "higher_is_better": True}
            cpu = ldm_cpu_baseline(seconds=20.0)
            line.update({"value": cpu["value"], "cpu_baseline": cpu, "n_gpus": 1,
                         "config": {"workload": "LDM variant, oracle port on host cores"},
                         "e2e": {"value": cpu["value"], "unit": "hyp/s", "h2d_bytes_per_step": 0,
                                 "d2h_bytes_per_step": 0}})
            print(json.dumps(line))
            return
        main_ldm(args)
        return
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, (world, args.gpus)

    from nope_b200 import synth_weights as weights  # seeded random-init weights, reference schema
    from nope_b200.model import build_model
    from nope_b200.dist import ShardedSweep
    from nope_b200.poses import synthetic_pose_batch

    n_local = args.poses
    n_global = n_local * world
    model = build_model(device=str(dev), chunk=args.chunk)
    model.load_state_dict(weights.make_full_state_dict(seed=0)).eval()
    unet = model.u_net
    unet.set_conv_impl(args.conv_impl)
    if world > 1:
        model.dist = ShardedSweep()

    # global grid: one icosphere-642 grid per GPU shard (pose VALUES do not affect timing)
    Q = args.queries
    poses_g, tposes = synthetic_pose_batch(n_local, Q)
    poses_g = poses_g.repeat(1, world, 1)           # [Q, n_global, 6]
    g = torch.Generator().manual_seed(0)
    q_img = (torch.rand(Q, 3, 256, 256, generator=g) * 2 - 1).pin_memory()
    r_img = (torch.rand(Q, 3, 256, 256, generator=g) * 2 - 1).pin_memory()
    poses_host = poses_g.clone().pin_memory()
    h2d = q_img.numel() * 4 + r_img.numel() * 4 + poses_host.numel() * 4
    d2h = Q * (5 * 8 + n_global * 4)                # top-5 indices (int64) + similarity rows

    # ---- resident inputs for `value`
    q_feat = unet.encoder.encode_image(q_img.to(dev))
    r_feat = unet.encoder.encode_image(r_img.to(dev))
    poses_dev = poses_g.to(dev)

    def step_resident():
        if world > 1:
            return model.dist.sweep(unet, r_feat, poses_dev, q_feat, k=5, want_emb=False)
        out = unet.sweep(r_feat, poses_dev, query_feat=q_feat, want_emb=False, k=5)
        return out["sim"], out["topi"], None

    def step_e2e():
        q = q_img.to(dev, non_blocking=True)
        r = r_img.to(dev, non_blocking=True)
        p = poses_host.to(dev, non_blocking=True)
        _, idx, sim = model.predict_pose(q, r, p, None, k=5)
        return idx.cpu(), sim.cpu()                 # D2H of the step's result

    def barrier():
        if world > 1:
            dist.barrier()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        barrier()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(step_resident, args.steps, max(args.warmup, 3))
    launches_per_step = unet.last_launch_count
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed(step_e2e, args.steps, max(args.warmup, 3))

    # ---- roofline of the dominant kernel (tcgen05 convolution), CUDA events per launch
    unet.profile(True)
    step_resident()
    prof = unet.profile_read()
    unet.profile(False)
    peak_tf, peak_hbm, peak_src = measured_peaks()
    conv_tf = prof["conv_flops"] / (prof["conv_ms"] * 1e-3) / 1e12 if prof["conv_ms"] > 0 else 0.0

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None if (args.no_cpu_baseline or world > 1) else cpu_baseline()
    eager = None
    if not args.no_cpu_baseline and world == 1:
        try:
            eager = eager_gpu_baseline(dev)
        except Exception as exc:                      # informational leg only
            eager = {"unavailable": repr(exc)[:200]}
    value = Q * n_global / (ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": "hyp/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
        "config": {
            "workload": f"configs[1]: 256x256, {n_local}-pose icosphere grid per GPU, batch={Q} query, "
                        "fp16 UNet (fp32 accumulate / statistics), l2 score + top-5",
            "poses_per_gpu": n_local, "global_poses": n_global, "queries": Q, "chunk": args.chunk,
            "conv_impl": args.conv_impl,
            "weights": "seeded random init, reference state_dict schema (305.8 M params)",
            "l2": "not flushed: each step streams 0.61 GB of fp16 weights and ~1.4 GB of "
                  "activations per chunk, >> 126 MB L2",
            "parallelism": f"pose grid sharded {world}-way, all-gather of top-k" if world > 1 else "1 GPU",
        },
        "clocks": clocks,
        "e2e": {"value": Q * n_global / (ms_e2e * 1e-3), "unit": "hyp/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "PoseConditional.predict_pose (pinned host images -> encoder x2 -> sweep -> top-5 -> host)"},
        "gpu_launches": int(launches_per_step * args.steps),
        "roofline": {
            "bound": "tensor", "kernel": "conv_tc_kernel (tcgen05 implicit-GEMM conv)",
            "achieved": conv_tf, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": conv_tf / peak_tf if peak_tf else None, "peak_source": f"{peak_src} (sustained bf16 cuBLAS)",
            "traffic": ncu_traffic(), "traffic_source": "profiles/roofline_traffic.json (ncu --set full, "
            "dram__bytes_read.sum + dram__bytes_write.sum per launch, sweep convolutions)",
            "launches_per_step": prof["conv_launches"],
            "conv_ms_per_step": prof["conv_ms"], "conv_share_of_step": prof["conv_ms"] / ms if ms else None,
            "algorithmic_tflop_per_step": prof["conv_flops"] / 1e12,
            "best_single_launch_tflops": prof["max_launch_tflops"],
            "whole_step_tflops": value * GFLOP_PER_HYP / 1e3,
        },
        "cpu_baseline": cpu,
        "eager_gpu_baseline": eager,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
