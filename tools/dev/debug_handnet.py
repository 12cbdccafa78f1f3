import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
from oracle import ref_modules
import test_handnet_gpu as T
cuda = torch.device("cuda:0")
refnet = ref_modules.load_full("ref")
ours, theirs = T._nets(refnet, cuda, "kp", "ops")
def rel(a, b): return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()
for mode in ("eval", "train"):
    for net in (ours, theirs):
        net.train(mode == "train")
        for m in net.modules():
            if isinstance(m, torch.nn.Dropout): m.p = 0.0
            if isinstance(m, torch.nn.MultiheadAttention): m.dropout = 0.0
    data = T._data(8, 2048, 6, cuda)
    recs = []
    for net in (ours, theirs):
        rec = {}
        hs = []
        for name, mod in net.named_modules():
            if name in ("bhand", "bhand.sa1", "bhand.sa2", "bhand.sa3", "bhand.fp3", "bhand.fp2", "bhand.fp1", "q1", "r1", "q2", "r2", "transt", "c3", "final_mlp"):
                def hk(mod, inp, out, name=name):
                    o = out
                    while isinstance(o, (tuple, list)): o = o[1] if name.startswith("bhand.sa") else o[0]
                    rec[name] = o.detach().float().clone()
                hs.append(mod.register_forward_hook(hk))
        with torch.no_grad():
            ret = net(data, {"track_flag": False, "IKNet_flag": False})
        rec["xyz1"] = ret["init_kp_handframe"]; rec["xyz2"] = ret["points_handframe"]; rec["pred"] = ret["pred_kp_handframe"]
        for h in hs: h.remove()
        recs.append(rec)
    print(mode)
    for k in recs[1]:
        print("  %-10s %.3e" % (k, rel(recs[0][k], recs[1][k])), "equal" if torch.equal(recs[0][k], recs[1][k]) else "")
