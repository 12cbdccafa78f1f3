"""Per-kernel micro-benchmark: ours vs the reference's own kernels (oracle/_ref) on the
same GPU, at the BASELINE config-3 shapes (B=32, N=4096).  CUDA-event timing, L2 flushed
between iterations.  Prints one JSON line per op with achieved algorithmic GB/s
(SURVEY.md section 8d byte counts) -- development aid; bench.py is the contract."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import clouds  # noqa: E402
from hotrack_b200 import pointnet2_utils as futils  # noqa: E402
from oracle import ref_lib  # noqa: E402


def timeit(fn, iters=20, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    return float(np.median(ts))


def main():
    dev = torch.device("cuda:0")
    B, N = int(os.environ.get("B", 32)), int(os.environ.get("N", 4096))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    xyz = torch.from_numpy(clouds.ball(B, N, seed=0)).to(dev)
    kp = torch.from_numpy(clouds.keypoints(B, 21, seed=0)).to(dev)
    has_ref = ref_lib.available()
    rows = []

    def row(name, ours, ref, nbytes):
        t = timeit(ours, flush=flush)
        r = timeit(ref, flush=flush) if (has_ref and ref is not None) else None
        rows.append(dict(op=name, ours_us=round(t, 2), ref_us=None if r is None else round(r, 2),
                         speedup=None if r is None else round(r / t, 2), alg_MB=round(nbytes / 1e6, 3),
                         ours_GBs=round(nbytes / t / 1e3, 1)))
        print(json.dumps(rows[-1]), flush=True)

    S1, S2, K = 256, 128, 32
    row("fps_sa1", lambda: futils.furthest_point_sample(xyz, S1), lambda: ref_lib.furthest_point_sample(xyz, S1),
        B * N * 12 + B * S1 * 4)
    fps1 = futils.furthest_point_sample(xyz, S1).long()
    l1 = torch.gather(xyz, 1, fps1.unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    row("fps_sa2", lambda: futils.furthest_point_sample(l1, S2), lambda: ref_lib.furthest_point_sample(l1, S2),
        B * S1 * 12 + B * S2 * 4)
    fps2 = futils.furthest_point_sample(l1, S2).long()
    l2 = torch.gather(l1, 1, fps2.unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    row("ball_sa1", lambda: futils.ball_query(0.1, K, xyz, l1), lambda: ref_lib.ball_query(0.1, K, xyz, l1),
        B * N * 12 + B * S1 * 12 + B * S1 * K * 4)
    row("ball_sa2", lambda: futils.ball_query(0.2, K, l1, l2), lambda: ref_lib.ball_query(0.2, K, l1, l2),
        B * S1 * 12 + B * S2 * 12 + B * S2 * K * 4)
    for k in (4, 16, 64):
        row("knn_k%d" % k, lambda: futils.knn(k, kp, xyz), lambda: ref_lib.knn(k, kp, xyz),
            B * 21 * 12 + B * N * 12 + B * 21 * k * 8)
    row("three_nn_fp1", lambda: futils.three_nn(xyz, l1), lambda: ref_lib.three_nn(xyz, l1),
        B * N * 12 + B * S1 * 12 + B * N * 24)
    row("three_nn_fp2", lambda: futils.three_nn(l1, l2), lambda: ref_lib.three_nn(l1, l2),
        B * S1 * 12 + B * S2 * 12 + B * S1 * 24)
    d, idx = futils.three_nn(xyz, l1)
    w = 1.0 / (d + 1e-8)
    w = (w / w.sum(-1, keepdim=True)).contiguous()
    C = 128
    feats = torch.randn(B, C, S1, device=dev)
    row("interp_fp1", lambda: futils.three_interpolate(feats, idx, w), lambda: ref_lib.three_interpolate(feats, idx, w),
        B * C * S1 * 4 + B * N * 24 + B * C * N * 4)
    g = torch.randn(B, C, N, device=dev)
    from hotrack_b200 import pointnet2_cuda as pc
    gp = torch.zeros(B, C, S1, device=dev)
    row("interp_grad_fp1", lambda: pc.three_interpolate_grad_wrapper(B, C, N, S1, g, idx, w, gp),
        lambda: ref_lib.three_interpolate_grad(g, idx, w, S1), B * C * S1 * 4 + B * N * 24 + B * C * N * 4)
    bidx = futils.ball_query(0.2, K, l1, l2)
    f64 = torch.randn(B, 64, S1, device=dev)
    row("group_sa2", lambda: futils.grouping_operation(f64, bidx), lambda: ref_lib.group_points(f64, bidx),
        B * 64 * S1 * 4 + B * S2 * K * 4 + B * 64 * S2 * K * 4)
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "bench_ops_B%d_N%d.json" % (B, N)), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
