"""Times pn2_knn for the few-queries shapes (21 joints, K=64) at B=1 / N=8192 and B=32 / N=4096."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hotrack_b200 import pointnet2_utils as fu

dev = torch.device("cuda", 0)
for B, N in ((1, 8192), (32, 4096), (1, 2048)):
    g = torch.Generator(device="cpu").manual_seed(0)
    known = torch.randn(B, N, 3, generator=g).to(dev)
    unknown = torch.randn(B, 21, 3, generator=g).to(dev)
    for _ in range(3):
        fu.knn(64, unknown, known)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fu.knn(64, unknown, known); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    print("knn B=%d N=%d n=21 K=64: median %.1f us" % (B, N, ts[len(ts) // 2]))
