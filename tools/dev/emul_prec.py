"""Precision budget experiment (ops engine, fp32) -- which stacks can stay 16-bit?

Emulates the fused engine's roundings inside chosen modules of the fp32 pipeline: conv operands and weights rounded to
fp16, conv outputs stored as fp16 around their channel mean.  Prints norm-relative error of every module output against
the plain fp32 run, for several choices of "16-bit modules", plus the deviation of the TF32 default (what the reference
itself runs with on torch >= 1.12) from strict fp32.
"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import torch.nn as nn
import torch.nn.functional as F
from hotrack_b200 import backbones, pointnet_utils as pu, synthetic
from hotrack_b200.handtrack_path import HandTrackPointPath, init_weights

dev = torch.device("cuda:0")
NAMES = ("bhand.sa1", "bhand.sa2", "bhand.sa3", "bhand.fp3", "bhand.fp2", "bhand.fp1", "bhand", "q1", "q2")


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()


def h(x):
    return x.half().float()


def wrap(conv, skip_first_input=False):
    orig = conv.forward

    def fwd(x):
        f = F.conv2d if x.dim() == 4 else F.conv1d
        y = f(h(x), h(conv.weight), conv.bias)
        dims = [d for d in range(y.dim()) if d != 1]
        m = y.mean(dim=dims, keepdim=True)
        return h(y - m) + m
    conv.forward = fwd
    return orig


def run(m, x, k, low=()):
    rec, hooks, undo = {}, [], []
    for name, mod in m.named_modules():
        if name in NAMES:
            def hk(mod, inp, out, name=name):
                o = out[1] if isinstance(out, tuple) and name.startswith("bhand.sa") else (out[0] if isinstance(out, tuple) else out)
                rec[name] = o.detach().float().clone()
            hooks.append(mod.register_forward_hook(hk))
        if isinstance(mod, (nn.Conv1d, nn.Conv2d)):
            if any(name == l or name.startswith(l + ".") for l in low):
                undo.append((mod, wrap(mod)))
    with torch.no_grad():
        m(x, k)
    for hh in hooks:
        hh.remove()
    for mod, o in undo:
        mod.forward = o
    return rec


B, N = (32, 4096) if len(sys.argv) < 2 else (int(sys.argv[1]), int(sys.argv[2]))
pu.set_engine("ops")
mo = HandTrackPointPath(backbones.default_cfg(dev)); init_weights(mo, seed=0); mo = mo.to(dev).train()
x = torch.from_numpy(synthetic.ball(B, N, seed=4)).to(dev).transpose(1, 2).contiguous()
k = torch.from_numpy(synthetic.keypoints(B, 21, seed=4)).to(dev).transpose(1, 2).contiguous()
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
ref = run(mo, x, k)
torch.backends.cudnn.allow_tf32 = True; torch.backends.cuda.matmul.allow_tf32 = True
tf32 = run(mo, x, k)
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
cases = {
    "tf32(all)": tf32,
    "f16 all": run(mo, x, k, low=("bhand.sa1","bhand.sa2","bhand.sa3","bhand.fp3","bhand.fp2","bhand.fp1","bhand.conv1","q1","q2")),
    "f16 >=fp3.1": run(mo, x, k, low=("bhand.fp3.mlp_convs.1", "bhand.fp2", "bhand.fp1", "bhand.conv1", "q1", "q2")),
    "f16 >=fp2": run(mo, x, k, low=("bhand.fp2", "bhand.fp1", "bhand.conv1", "q1", "q2")),
    "f16 >=fp1": run(mo, x, k, low=("bhand.fp1", "bhand.conv1", "q1", "q2")),
    "f16 sa1": run(mo, x, k, low=("bhand.sa1",)),
    "f16 q1q2": run(mo, x, k, low=("q1", "q2")),
}
print("B=%d N=%d  rel err vs strict fp32" % (B, N))
print("  %-10s " % "" + "  ".join("%-12s" % c for c in cases))
for n in NAMES:
    print("  %-10s " % n + "  ".join("%-12.5f" % rel(v[n], ref[n]) for v in cases.values()))
