"""One warm sweep + one sweep inside cudaProfilerStart/Stop, for ncu
(`ncu --profile-from-start off ...`).  642 poses, 1 query, chunk from NOPE_CHUNK."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import weights
from nope_b200.model import build_model
from nope_b200.poses import synthetic_pose_batch

n = int(os.environ.get("NOPE_POSES", "642"))
model = build_model(device="cuda:0", chunk=int(os.environ.get("NOPE_CHUNK", "642")))
model.load_state_dict(weights.make_full_state_dict(seed=0)).eval()
poses, _ = synthetic_pose_batch(n, 1)
g = torch.Generator().manual_seed(0)
rf = (torch.randn(1, 8, 32, 32, generator=g) * 1.5).cuda()
qf = (torch.randn(1, 8, 32, 32, generator=g) * 1.5).cuda()
poses = poses.cuda()
for _ in range(int(os.environ.get("NOPE_WARM", "1"))):
    model.u_net.sweep(rf, poses, query_feat=qf, want_emb=False, k=5)
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = model.u_net.sweep(rf, poses, query_feat=qf, want_emb=False, k=5)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("top5", out["topi"].tolist(), "launches", model.u_net.last_launch_count)
