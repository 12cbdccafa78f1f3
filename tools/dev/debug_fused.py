import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from hotrack_b200 import backbones, pointnet_utils as pu, synthetic
from hotrack_b200.handtrack_path import HandTrackPointPath, init_weights
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
def rel(a, b): return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()
B, N = 4, 2048
for train in (True, False):
    models = {}
    for eng in ("ops", "fused"):
        pu.set_engine(eng)
        m = HandTrackPointPath(backbones.default_cfg(dev)); init_weights(m, seed=0)
        models[eng] = m.to(dev).train(train)
    pu.set_engine("ops")
    x = torch.from_numpy(synthetic.ball(B, N, seed=4)).to(dev).transpose(1, 2).contiguous()
    k = torch.from_numpy(synthetic.keypoints(B, 21, seed=4)).to(dev).transpose(1, 2).contiguous()
    acts = {}
    for eng, m in models.items():
        rec = {}
        hooks = []
        for name, mod in m.named_modules():
            if name in ("bhand.sa1", "bhand.sa2", "bhand.sa3", "bhand.fp3", "bhand.fp2", "bhand.fp1", "bhand", "q1", "q2"):
                def hk(mod, inp, out, name=name):
                    o = out[1] if isinstance(out, tuple) and name.startswith("bhand.sa") else (out[0] if isinstance(out, tuple) else out)
                    rec[name] = o.detach().clone()
                hooks.append(mod.register_forward_hook(hk))
        out = m(x, k)
        for h in hooks: h.remove()
        acts[eng] = rec
    print("train" if train else "eval")
    for name in acts["ops"]:
        a, b = acts["ops"][name], acts["fused"][name]
        print("  %-10s shape %s rel %.4f  max|ops| %.3f" % (name, tuple(a.shape), rel(b, a), a.abs().max().item()))
