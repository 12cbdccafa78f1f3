import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from hotrack_b200 import backbones, pointnet_utils as pu, synthetic
from hotrack_b200.handtrack_path import HandTrackPointPath, init_weights
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
def rel(a, b): return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()

# (a) a dense 3-layer stack on well-conditioned random data: fused vs torch fp32
torch.manual_seed(0)
import torch.nn as nn
from hotrack_b200 import fused
for R in (4096, 65536):
    convs = nn.ModuleList([nn.Conv1d(64, 128, 1), nn.Conv1d(128, 128, 1), nn.Conv1d(128, 64, 1)]).to(dev)
    bns = nn.ModuleList([nn.BatchNorm1d(128), nn.BatchNorm1d(128), nn.BatchNorm1d(64)]).to(dev)
    x = torch.randn(4, 64, R // 4, device=dev)
    g = torch.randn(4, 64, R // 4, device=dev)
    xa = x.clone().requires_grad_(True)
    h = xa
    for c, b in zip(convs, bns):
        h = torch.relu(b(c(h)))
    h.backward(g)
    ref = {n: p.grad.clone() for n, p in list(convs.named_parameters()) + [("bn" + n, p) for n, p in bns.named_parameters()]}
    refx = xa.grad.clone()
    for p in list(convs.parameters()) + list(bns.parameters()): p.grad = None
    xb = x.clone().requires_grad_(True)
    out = fused.dense_stack(xb, convs, bns, True)
    print("dense R=%d fwd rel %.4f" % (R, rel(out, h)))
    out.backward(g)
    for n, p in list(convs.named_parameters()) + [("bn" + n, p) for n, p in bns.named_parameters()]:
        if n.endswith("bias") and not n.startswith("bn"): continue
        print("   %-12s %.4f" % (n, rel(p.grad, ref[n])))
    print("   dx           %.4f" % rel(xb.grad, refx))

# (b) whole path, per-parameter
if len(sys.argv) > 1 and sys.argv[1] == 'a': sys.exit(0)
for (B, N) in ((4, 2048), (16, 4096)):
    models = {}
    for eng in ("ops", "fused"):
        pu.set_engine(eng)
        m = HandTrackPointPath(backbones.default_cfg(dev)); init_weights(m, seed=0)
        models[eng] = m.to(dev).train()
    pu.set_engine("ops")
    x = torch.from_numpy(synthetic.ball(B, N, seed=4)).to(dev).transpose(1, 2).contiguous()
    k = torch.from_numpy(synthetic.keypoints(B, 21, seed=4)).to(dev).transpose(1, 2).contiguous()
    for e, m in models.items():
        o = m(x, k)
        sum(v.square().mean() for v in o[:3]).backward()
    print("whole path B=%d N=%d" % (B, N))
    for (n1, p1), (n2, p2) in zip(models["ops"].named_parameters(), models["fused"].named_parameters()):
        if n1.endswith(".bias") and "conv" in n1: continue
        print("   %-40s %.4f  |g| %.3e" % (n1, rel(p2.grad, p1.grad), p1.grad.norm().item()))
