import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from hotrack_b200 import backbones, pointnet_utils as pu, synthetic
from hotrack_b200.handtrack_path import HandTrackPointPath, init_weights
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
def rel(a, b): return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()
NAMES = ("bhand.sa1", "bhand.sa2", "bhand.sa3", "bhand.fp3", "bhand.fp2", "bhand.fp1", "bhand", "q1", "q2")
def run(m, x, k, quant=None, autocast=False):
    rec = {}; hooks = []
    for name, mod in m.named_modules():
        if name in NAMES:
            def hk(mod, inp, out, name=name):
                o = out[1] if isinstance(out, tuple) and name.startswith("bhand.sa") else (out[0] if isinstance(out, tuple) else out)
                rec[name] = o.detach().float().clone()
                if quant and name in quant:
                    if isinstance(out, tuple) and name.startswith("bhand.sa"):
                        return (out[0], out[1].to(torch.bfloat16).float())
                    if isinstance(out, tuple):
                        return (out[0].to(torch.bfloat16).float(),) + tuple(out[1:])
                    return out.to(torch.bfloat16).float()
            hooks.append(mod.register_forward_hook(hk))
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        m(x, k)
    for h in hooks: h.remove()
    return rec
for (B, N) in ((4, 2048), (32, 4096)):
    pu.set_engine("ops")
    mo = HandTrackPointPath(backbones.default_cfg(dev)); init_weights(mo, seed=0); mo = mo.to(dev).train()
    pu.set_engine("fused")
    mf = HandTrackPointPath(backbones.default_cfg(dev)); init_weights(mf, seed=0); mf = mf.to(dev).train()
    pu.set_engine("ops")
    x = torch.from_numpy(synthetic.ball(B, N, seed=4)).to(dev).transpose(1, 2).contiguous()
    k = torch.from_numpy(synthetic.keypoints(B, 21, seed=4)).to(dev).transpose(1, 2).contiguous()
    ref = run(mo, x, k)
    q3 = run(mo, x, k, quant=("bhand.sa3",))
    qall = run(mo, x, k, quant=NAMES)
    fu = run(mf, x, k)
    print("B=%d N=%d   rel err vs fp32 ops:  [sa3->bf16]  [all module outputs->bf16]  [fused]" % (B, N))
    for n in NAMES:
        print("  %-10s %.4f  %.4f  %.4f" % (n, rel(q3[n], ref[n]), rel(qall[n], ref[n]), rel(fu[n], ref[n])))
