"""Per-parameter gradient deviation of the whole path with row-form sinks on vs plain autograd (the body of
tests/test_fused_gpu.py::test_row_form_sinks_match_autograd_on_the_whole_path, printing everything), plus sink-off twice
as the noise floor."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import clouds
from hotrack_b200 import backbones, fused, pointnet_utils as pu
from hotrack_b200.handtrack_path import HandTrackPointPath, init_weights

cuda = torch.device("cuda", 0)
B, N = 4, 2048
x = torch.from_numpy(clouds.ball(B, N, seed=11)).to(cuda).transpose(1, 2).contiguous()
k = torch.from_numpy(clouds.keypoints(B, 21, seed=11)).to(cuda).transpose(1, 2).contiguous()
pu.set_engine("fused"); m = HandTrackPointPath(backbones.default_cfg(cuda)); pu.set_engine("ops")
init_weights(m, seed=0); m = m.to(cuda).train()
with torch.no_grad():
    m(x, k)
state = {n_: b_.clone() for n_, b_ in m.named_buffers()}
centers = {id(mod): mod._pn2_center.clone() for mod in m.modules() if hasattr(mod, "_pn2_center")}
res = []
for sink in (False, False, True, True):
    fused.set_sparse_grad_sink(sink)
    with torch.no_grad():
        for n_, b_ in m.named_buffers():
            b_.copy_(state[n_])
        for mod in m.modules():
            if id(mod) in centers:
                mod._pn2_center.copy_(centers[id(mod)])
    for p_ in m.parameters():
        p_.grad = None
    src2, f11, f13, _ = m(x, k)
    (src2.square().mean() + f11.square().mean() + f13.square().mean()).backward()
    res.append({n_: p_.grad.clone() for n_, p_ in m.named_parameters() if p_.grad is not None})
fused.set_sparse_grad_sink(False)
rel = lambda a, b: ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
gmax = max(g.abs().max().item() for g in res[0].values())
print("%-40s %10s %10s %10s   |g|max" % ("param", "off-off", "on-off", "on-on"))
for n_, a in res[0].items():
    if a.abs().max().item() < 1e-6 * gmax:
        continue
    d = (rel(res[1][n_], a), rel(res[2][n_], a), rel(res[3][n_], res[2][n_]))
    flag = " <<<" if d[1] > 3 * max(d[0], d[2], 0.02) else ""
    print("%-40s %10.2e %10.2e %10.2e   %.2e%s" % (n_, d[0], d[1], d[2], a.abs().max().item(), flag))
