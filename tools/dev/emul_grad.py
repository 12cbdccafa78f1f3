"""Gradient-precision budget (ops engine, fp32): parameter-gradient error when every BatchNorm's incoming gradient is
rounded to bf16 / fp16 (what the fused engine's gradient rows store), against the exact fp32 backward; and the fused
engine's per-parameter error for comparison."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import torch.nn as nn
from hotrack_b200 import backbones, pointnet_utils as pu, synthetic
from hotrack_b200.handtrack_path import HandTrackPointPath, init_weights
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
def rel(a, b): return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()
B, N = (32, 4096) if len(sys.argv) < 3 else (int(sys.argv[1]), int(sys.argv[2]))
x = torch.from_numpy(synthetic.ball(B, N, seed=4)).to(dev).transpose(1, 2).contiguous()
k = torch.from_numpy(synthetic.keypoints(B, 21, seed=4)).to(dev).transpose(1, 2).contiguous()

def build(engine):
    pu.set_engine(engine)
    m = HandTrackPointPath(backbones.default_cfg(dev)); init_weights(m, seed=0)
    pu.set_engine("ops")
    return m.to(dev).train()

def grads(m, tg=None, rnd=None, scale=1.0):
    hooks = []
    if rnd is not None:
        for mod in m.modules():
            if isinstance(mod, (nn.BatchNorm1d, nn.BatchNorm2d)):
                hooks.append(mod.register_full_backward_pre_hook(lambda mod, go: tuple((g * scale).to(rnd).float() / scale for g in go)))
    for p in m.parameters(): p.grad = None
    o = m(x, k)
    if tg is None:
        g = torch.Generator(device="cpu").manual_seed(1)
        tg = [torch.randn(v.shape, generator=g).to(dev) * v.detach().std() + v.detach().mean() for v in o[:3]]
    sum((v - t).square().mean() for v, t in zip(o[:3], tg)).backward()
    for h in hooks: h.remove()
    return {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}, tg

mo = build("ops")
g0, tg = grads(mo)
gb, _ = grads(mo, tg, torch.bfloat16)
gh, _ = grads(mo, tg, torch.float16, scale=2.0 ** 14)
mf = build("fused")
gf, _ = grads(mf, tg)
from hotrack_b200 import fused
fused.set_precise("all")
mf1 = build("fused")
gf1, _ = grads(mf1, tg)
fused.set_precise("auto")
print("%-42s %9s %9s %9s %9s   |g|" % ("parameter", "bf16-dz", "fp16-dz", "fused", "fused-all"))
for n in g0:
    if n.endswith(".bias") and "conv" in n: continue
    print("%-42s %9.4f %9.4f %9.4f %9.4f   %.2e" % (n, rel(gb[n], g0[n]), rel(gh[n], g0[n]), rel(gf[n], g0[n]), rel(gf1[n], g0[n]), g0[n].norm()))

# (c) forward-only perturbation: fp16 roundings (operands, weights, centred outputs) in the modules behind FP3's first
# layer -- the fused engine's one-plane part -- with an exact fp32 backward
import torch.nn.functional as F
def h(t): return t.half().float()
def wrap(conv):
    orig = conv.forward
    def fwd(xx):
        f = F.conv2d if xx.dim() == 4 else F.conv1d
        y = f(h(xx), h(conv.weight), conv.bias)
        m = y.mean(dim=[d for d in range(y.dim()) if d != 1], keepdim=True).detach()
        return h(y - m) + m
    conv.forward = fwd
    return orig
low = ("bhand.fp3.mlp_convs.1", "bhand.fp2", "bhand.fp1", "bhand.conv1", "q1", "q2")
undo = [(mod, wrap(mod)) for name, mod in mo.named_modules() if isinstance(mod, (nn.Conv1d, nn.Conv2d))
        and any(name == l or name.startswith(l + ".") for l in low)]
gp, _ = grads(mo, tg)
for mod, o in undo: mod.forward = o
print("\nforward-only fp16 emulation behind FP3.0, exact fp32 backward:")
for n in g0:
    if n.endswith(".bias") and "conv" in n: continue
    print("%-42s %9.4f" % (n, rel(gp[n], g0[n])))
