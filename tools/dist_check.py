"""torchrun check (N GPUs, NCCL): the sharded sweep + all-gather + merge equals the
single-GPU sweep bit for bit, on every rank."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from nope_b200 import synth_weights as weights
from nope_b200.dist import ShardedSweep
from nope_b200.model import build_model
from nope_b200.poses import synthetic_pose_batch

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
model = build_model(device=str(dev))
model.load_state_dict(weights.make_full_state_dict(seed=0)).eval()
g = torch.Generator().manual_seed(3)
rf = torch.randn(2, 8, 32, 32, generator=g) * 1.5
qf = torch.randn(2, 8, 32, 32, generator=g) * 1.5
ok = True
ss = ShardedSweep()
for n in (7, 162, 643):     # 7: shards of one pose / empty shards (k > shard size); 643: uneven shards
    poses, _ = synthetic_pose_batch(642, 2)
    poses = poses[:, :n] if n <= 642 else torch.cat([poses, poses[:, :1]], dim=1)
    single = model.u_net.sweep(rf, poses, query_feat=qf, want_emb=False, k=5)
    sim, topi, _ = ss.sweep(model.u_net, rf.to(dev), poses.to(dev), qf.to(dev), k=5)
    same = torch.equal(sim, single["sim"]) and torch.equal(topi, single["topi"])
    ok = ok and same
    print(f"rank {rank} n={n}: sharded == single: {same}; top5 {topi[0].tolist()}", flush=True)
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
if rank == 0:
    print("DIST_CHECK", "PASS" if int(t) == 1 else "FAIL")
sys.exit(0 if int(t) == 1 else 1)
