"""Multi-GPU sweep: the pose grid is embarrassingly parallel over hypotheses
(src/model/model.py:212-222 -- every (reference, pose) forward is independent), so rank g
takes a contiguous slice of the N poses with replicated weights and features, and the only
collective is one all-gather of each rank's local (score, global index) top-k and,
optionally, of its slice of the similarity row.  One process per GPU,
torch.distributed (NCCL on GPUs; gloo in the CPU tests of the merge logic).  The reference
has no inference-time sharding; Lightning DDP there only shards the dataloader."""
import torch
import torch.distributed as dist


def shard_range(n_poses, rank, world):
    """Contiguous slice [lo, hi) of rank `rank`; ceil(N/W) poses per rank, last may be short
    or empty."""
    per = (n_poses + world - 1) // world
    lo = min(rank * per, n_poses)
    return lo, min(lo + per, n_poses)


def merge_topk(vals, idx, k):
    """vals/idx [W*k', B] candidates (any order, idx -1 = padding) -> top-k per batch row,
    descending score, ties -> lowest global index.  Pure torch, tiny tensors."""
    vals = vals.clone()
    vals[idx < 0] = float("-inf")
    order = torch.sort(idx, dim=0, stable=True).indices           # ascending index first ...
    v1, i1 = torch.gather(vals, 0, order), torch.gather(idx, 0, order)
    order2 = torch.sort(-v1, dim=0, stable=True).indices          # ... then stable by score
    v2, i2 = torch.gather(v1, 0, order2), torch.gather(i1, 0, order2)
    return v2[:k].t().contiguous(), i2[:k].t().contiguous()


class ShardedSweep:
    def __init__(self, group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def gather_merge(self, local_sim, local_topv, local_topi, n_poses, k, want_sim=True):
        """local_sim [B, n_local] (may be empty), local_topv/topi [B, k_local] with GLOBAL
        indices -> (sim [B,N] | None, topi [B,k]) identical on every rank."""
        B = local_topv.shape[0] if local_topv is not None else local_sim.shape[0]
        dev = local_sim.device
        kk = k
        pv = torch.full((kk, B), float("-inf"), device=dev, dtype=torch.float32)
        pi = torch.full((kk, B), -1, device=dev, dtype=torch.int64)
        if local_topv is not None and local_topv.numel() > 0:
            kl = local_topv.shape[1]
            pv[:kl] = local_topv.t()
            pi[:kl] = local_topi.t()
        gv = torch.empty((self.world * kk, B), device=dev, dtype=torch.float32)
        gi = torch.empty((self.world * kk, B), device=dev, dtype=torch.int64)
        dist.all_gather_into_tensor(gv, pv, group=self.group)
        dist.all_gather_into_tensor(gi, pi, group=self.group)
        _, topi = merge_topk(gv, gi, k)
        sim = None
        if want_sim:
            per = (n_poses + self.world - 1) // self.world
            ps = torch.full((B, per), float("-inf"), device=dev, dtype=torch.float32)
            ps[:, : local_sim.shape[1]] = local_sim
            gs = torch.empty((self.world, B, per), device=dev, dtype=torch.float32)
            dist.all_gather_into_tensor(gs.view(self.world * B, per), ps, group=self.group)
            sim = gs.permute(1, 0, 2).reshape(B, self.world * per)[:, :n_poses].contiguous()
        return sim, topi

    def sweep(self, u_net, reference_feat, all_relativeR, query_feat, k=5, metric="l2",
              want_emb=False, want_sim=True):
        from .model import score_topk
        N = all_relativeR.shape[1]
        lo, hi = shard_range(N, self.rank, self.world)
        B = reference_feat.shape[0]
        dev = u_net.device
        emb = None
        if hi > lo:
            kl = min(k, hi - lo)
            poses = all_relativeR[:, lo:hi].contiguous()
            if metric == "l2":
                out = u_net.sweep(reference_feat, poses, query_feat=query_feat, want_emb=want_emb,
                                  k=kl, idx_base=lo)
                sim_l, tv, ti, emb = out["sim"], out["topv"], out["topi"], out["emb"]
            else:
                emb = u_net.sweep(reference_feat, poses, want_emb=True)["emb"]
                sim_l, ti = score_topk(query_feat, emb, k=kl, metric=metric, idx_base=lo)
                tv = torch.gather(sim_l, 1, ti - lo)
        else:
            sim_l = torch.empty((B, 0), device=dev)
            tv = torch.empty((B, 0), device=dev)
            ti = torch.empty((B, 0), device=dev, dtype=torch.int64)
        sim, topi = self.gather_merge(sim_l, tv, ti, N, k, want_sim=want_sim)
        return sim, topi, emb
