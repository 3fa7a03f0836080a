"""One warm step + one step inside cudaProfilerStart/Stop, for ncu
(`ncu --profile-from-start off ...`): the public predict_pose path = native encoder (2
images) + sweep over the pose grid + fused score/top-k.  642 poses, 1 query by default; NOPE_POSES / NOPE_QUERIES select another workload
(BASELINE configs[2]: NOPE_POSES=2562 NOPE_QUERIES=8)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from nope_b200 import synth_weights as weights
from nope_b200.model import build_model
from nope_b200.poses import synthetic_pose_batch

n = int(os.environ.get("NOPE_POSES", "642"))
Q = int(os.environ.get("NOPE_QUERIES", "1"))
model = build_model(device="cuda:0", chunk=int(os.environ.get("NOPE_CHUNK", "642")))
model.load_state_dict(weights.make_full_state_dict(seed=0)).eval()
poses, _ = synthetic_pose_batch(n, Q)
g = torch.Generator().manual_seed(0)
q = (torch.rand(Q, 3, 256, 256, generator=g) * 2 - 1).cuda()
r = (torch.rand(Q, 3, 256, 256, generator=g) * 2 - 1).cuda()
poses = poses.cuda()
for _ in range(int(os.environ.get("NOPE_WARM", "1"))):
    model.predict_pose(q, r, poses, None, k=5)
torch.cuda.synchronize()
torch.cuda.profiler.start()
_, idx, sim = model.predict_pose(q, r, poses, None, k=5)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("top5", idx.tolist(), "launches", model.u_net.last_launch_count)
