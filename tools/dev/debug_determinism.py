import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, clouds
from hotrack_b200 import backbones, pointnet_utils as pu
from hotrack_b200.handtrack_path import HandTrackPointPath, init_weights
from hotrack_b200.train import TrainStep
dev = torch.device("cuda:0")
B, N = (2, 1024) if len(sys.argv) < 3 else (int(sys.argv[1]), int(sys.argv[2]))
x = torch.from_numpy(clouds.ball(B, N, seed=3)).to(dev).transpose(1, 2).contiguous()
k = torch.from_numpy(clouds.keypoints(B, 21, seed=3)).to(dev).transpose(1, 2).contiguous()
tg = [torch.randn(s, generator=torch.Generator().manual_seed(i)).to(dev) for i, s in enumerate([(B, 384, N), (B, 384, 21), (B, 384, 21)])]
loss = lambda out: sum((o - t).square().mean() for o, t in zip(out[:3], tg))
def run():
    pu.set_engine("fused"); m = HandTrackPointPath(backbones.default_cfg(dev)); pu.set_engine("ops")
    init_weights(m, seed=0)
    ts = TrainStep(m.to(dev).train(), loss, lr=1e-3, graph=False)
    ts._fwd_bwd((x, k))
    torch.cuda.synchronize()
    return {n: p.grad.clone() for n, p in m.named_parameters()}, ts
a, _ = run(); b, _ = run()
rel = lambda u, v: ((u - v).norm() / v.norm().clamp_min(1e-30)).item()
worst = sorted(((rel(a[n], b[n]), n) for n in a if a[n].abs().max() > 0), reverse=True)
print("B=%d N=%d run-to-run gradient difference, worst parameters:" % (B, N))
for r, n in worst[:12]: print("  %.3e %s" % (r, n))
print("  median %.3e" % worst[len(worst) // 2][0])
print("in network order (reverse):")
for n in reversed(list(a)):
    if a[n].abs().max() > 0 and not (n.endswith(".bias") and "conv" in n): print("  %.3e %s" % (rel(a[n], b[n]), n))
