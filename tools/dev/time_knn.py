"""Times the pn2_knn kernel (C ABI, 20 launches between two events) for the few-queries shapes (21 joints, K=64).
PN2_KNN_COOP=0|1 forces the warp-per-query / cooperative variant."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hotrack_b200 import _lib

dev = torch.device("cuda", 0)
for B, N in ((1, 8192), (1, 2048), (4, 4096), (8, 4096), (32, 4096)):
    g = torch.Generator(device="cpu").manual_seed(0)
    known = torch.randn(B, N, 3, generator=g).to(dev)
    unknown = torch.randn(B, 21, 3, generator=g).to(dev)
    d2 = torch.empty(B, 21, 64, device=dev)
    idx = torch.empty(B, 21, 64, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    run = lambda: _lib.call("pn2_knn", B, 21, N, 64, unknown.data_ptr(), known.data_ptr(), d2.data_ptr(), idx.data_ptr(), st)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        run()
    e.record(); torch.cuda.synchronize()
    print("knn B=%d N=%d n=21 K=64: %.1f us/launch (COOP=%s)" % (B, N, s.elapsed_time(e) * 1e3 / 20, os.environ.get("PN2_KNN_COOP", "auto")))
