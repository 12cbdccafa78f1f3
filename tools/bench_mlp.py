"""Micro-benchmark of the grouped-MLP C-ABI kernels at the shapes of the BASELINE step (B=32, N=4096):
CUDA-event timing with an L2 flush between iterations, achieved algorithmic GB/s per kernel.
python tools/bench_mlp.py [filter]   (development aid; bench.py is the contract)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hotrack_b200 import _lib, fused  # noqa: E402

dev = torch.device("cuda:0")
HF, BF = torch.float16, torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def st():
    return torch.cuda.current_stream().cuda_stream


def timeit(fn, iters=10, warmup=2):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


CASES = {  # name: (rows, k, n)   layer shapes of the step
    "sa1_l2": (262144, 32, 32), "sa1_l3": (262144, 32, 64), "sa2_l1": (131072, 96, 64), "sa2_l3": (131072, 64, 128),
    "fp1_l1": (131072, 160, 128), "fp1_l2": (131072, 128, 128), "conv1": (131072, 128, 384),
    "q2_l1": (43008, 800, 128), "q_l3": (43008, 128, 192),
}


def main():
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    out = []
    for name, (R, K, N) in CASES.items():
        if flt and flt not in name:
            continue
        x = torch.randn(R, K, device=dev).to(HF)
        w = (torch.randn(N, K, device=dev) / K ** 0.5).to(HF)
        wt = w.t().contiguous().to(BF)
        sc, sh = torch.rand(K, device=dev) + 0.5, torch.randn(K, device=dev) * 0.1
        y = torch.empty(R, N, dtype=HF, device=dev)
        stats = torch.zeros(2, N, device=dev)
        cen = torch.zeros(N, device=dev)
        args = ("pn2_mlp_gemm_fwd", R, K, N, x.data_ptr(), K, sc.data_ptr(), sh.data_ptr(), w.data_ptr(), cen.data_ptr(),
                y.data_ptr(), N, stats.data_ptr(), st())
        t = timeit(lambda: _lib.call(*args))
        nb = fused.alg_bytes("pn2_mlp_gemm_fwd", args[1:])
        out.append(dict(kernel="gemm_fwd", case=name, us=round(t, 1), GBs=round(nb / t / 1e3, 1)))
        print(out[-1], flush=True)
        dz = torch.randn(R, N, device=dev).to(BF)
        co = [torch.randn(N, device=dev) * 0.3 for _ in range(3)]
        pm = [torch.randn(K, device=dev) * 0.3 + 1 for _ in range(4)]
        dzp = torch.empty(R, K, dtype=BF, device=dev)
        sums = torch.zeros(2, K, device=dev)
        wb = wt.to(torch.float16)
        unscale = torch.ones(1, device=dev)
        args = ("pn2_mlp_gemm_dgrad", R, N, K, dz.data_ptr(), N, y.data_ptr(), N, wt.data_ptr(), wb.data_ptr(),
                pm[0].data_ptr(), unscale.data_ptr(), x.data_ptr(), K, pm[0].data_ptr(), pm[1].data_ptr(), pm[2].data_ptr(),
                pm[3].data_ptr(), dzp.data_ptr(), K, sums.data_ptr(), st())
        t = timeit(lambda: _lib.call(*args))
        nb = fused.alg_bytes("pn2_mlp_gemm_dgrad", args[1:])
        out.append(dict(kernel="gemm_dgrad", case=name, us=round(t, 1), GBs=round(nb / t / 1e3, 1)))
        print(out[-1], flush=True)
        dw = torch.zeros(N, K, device=dev)
        args = ("pn2_mlp_gemm_wgrad", R, N, K, K, dz.data_ptr(), N, y.data_ptr(), N, co[0].data_ptr(), co[1].data_ptr(),
                co[2].data_ptr(), x.data_ptr(), K, sc.data_ptr(), sh.data_ptr(), dw.data_ptr(), K, st())
        t = timeit(lambda: _lib.call(*args))
        nb = fused.alg_bytes("pn2_mlp_gemm_wgrad", args[1:])
        out.append(dict(kernel="gemm_wgrad", case=name, us=round(t, 1), GBs=round(nb / t / 1e3, 1)))
        print(out[-1], flush=True)
    # pooling: head (K=1, 384 ch), FP1 (K=1, 128 ch), SA1 (K=32, 64 ch), SA2 (K=32, 128 ch), q (K=64, 192 ch, 21 groups)
    for name, (B, S, K, C) in {"pool_head": (32, 4096, 1, 384), "pool_fp1": (32, 4096, 1, 128), "pool_sa1": (32, 256, 32, 64),
                               "pool_sa2": (32, 128, 32, 128), "pool_q64": (32, 21, 64, 192), "pool_sa3": (32, 1, 128, 512)}.items():
        if flt and flt not in name:
            continue
        R = B * S * K
        y = torch.randn(R, C, device=dev).to(HF)
        c4 = [torch.randn(C, device=dev) * 0.3 + 1 for _ in range(4)]
        out_cm = torch.empty(B, C, S, device=dev)
        cs = torch.zeros(C, device=dev)
        am = torch.empty(B, S, C, dtype=torch.int32, device=dev)
        args = ("pn2_pool_fwd", B, S, K, C, y.data_ptr(), C, c4[0].data_ptr(), c4[1].data_ptr(), out_cm.data_ptr(),
                cs.data_ptr() if K > 1 else 0, am.data_ptr() if K > 1 else 0, st())
        t = timeit(lambda: _lib.call(*args))
        nb = fused.alg_bytes("pn2_pool_fwd", args[1:])
        out.append(dict(kernel="pool_fwd", case=name, us=round(t, 1), GBs=round(nb / t / 1e3, 1)))
        print(out[-1], flush=True)
        dout = torch.randn(B, C, S, device=dev)
        dz = torch.empty(R, C, dtype=BF, device=dev)
        sums = torch.zeros(2, C, device=dev)
        args = ("pn2_pool_bwd", B, S, K, C, dout.data_ptr(), 0, 0, 0, y.data_ptr(), C, c4[0].data_ptr(), c4[1].data_ptr(),
                c4[2].data_ptr(), c4[3].data_ptr(), am.data_ptr() if K > 1 else 0, dz.data_ptr(), C, sums.data_ptr(), st())
        t = timeit(lambda: _lib.call(*args))
        nb = fused.alg_bytes("pn2_pool_bwd", args[1:])
        out.append(dict(kernel="pool_bwd", case=name, us=round(t, 1), GBs=round(nb / t / 1e3, 1)))
        print(out[-1], flush=True)
    for name, (B, C, N) in {"to_rows_q": (32, 192, 21), "to_rows_sa1": (32, 64, 256)}.items():
        if flt and flt not in name:
            continue
        src = torch.randn(B, C, N, device=dev)
        sums = torch.randn(C, device=dev)
        dst = torch.empty(B * N, C, dtype=HF, device=dev)
        args = ("pn2_to_rows", B, C, N, src.data_ptr(), sums.data_ptr(), 0.01, dst.data_ptr(), C, st())
        t = timeit(lambda: _lib.call(*args))
        out.append(dict(kernel="to_rows", case=name, us=round(t, 1)))
        print(out[-1], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_mlp.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
