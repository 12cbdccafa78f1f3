import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, clouds
from hotrack_b200 import backbones, pointnet_utils as pu, fused
from hotrack_b200.handtrack_path import HandTrackPointPath, init_weights
dev = torch.device("cuda:0")
B, N = 32, 4096
x = torch.from_numpy(clouds.ball(B, N, seed=3)).to(dev).transpose(1, 2).contiguous()
k = torch.from_numpy(clouds.keypoints(B, 21, seed=3)).to(dev).transpose(1, 2).contiguous()
NAMES = ("bhand.sa1", "bhand.sa2", "bhand.sa3", "bhand.fp3", "bhand.fp2", "bhand", "q1", "q2")
def run(mode):
    fused.set_precise(mode)
    pu.set_engine("fused"); m = HandTrackPointPath(backbones.default_cfg(dev)); pu.set_engine("ops")
    init_weights(m, seed=0); m = m.to(dev).train()
    rec = {}
    for name, mod in m.named_modules():
        if name in NAMES:
            def hk(mod, inp, out, name=name):
                o = out[1] if isinstance(out, tuple) and name.startswith("bhand.sa") else (out[0] if isinstance(out, tuple) else out)
                rec[name] = o.detach().float().clone()
            mod.register_forward_hook(hk)
    with torch.no_grad():
        m(x, k)
    torch.cuda.synchronize()
    return rec
rel = lambda u, v: ((u - v).norm() / v.norm().clamp_min(1e-30)).item()
for mode in ("auto", "off", "all"):
    a, b = run(mode), run(mode)
    print("precision", mode, "forward run-to-run:")
    for n in NAMES: print("   %-10s rel %.3e  bitwise-equal %s  max|d| %.3e" % (n, rel(a[n], b[n]), torch.equal(a[n], b[n]), (a[n]-b[n]).abs().max().item()))
