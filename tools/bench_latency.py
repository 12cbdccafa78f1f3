"""BASELINE config 5: per-frame tracking inference, B=1, N=8192, backbone only (SA+FP: PointNet2Msg_fast),
eval mode, sequential frames (each frame synchronised before the next starts, as the tracker's recurrence
forces).  p50 / p99 wall latency per frame: ours (fused engine, CUDA-graph replay; also eager) vs the reference
(its modules on its own kernels, eager).  python tools/bench_latency.py [frames]"""
import json
import os
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hotrack_b200 import backbones, pointnet_utils as pu, synthetic  # noqa: E402
from hotrack_b200.handtrack_path import init_weights  # noqa: E402
from hotrack_b200.train import GraphedForward  # noqa: E402


def run(fn, frames, x):
    for _ in range(20):
        fn(x)
    torch.cuda.synchronize()
    ts = []
    for _ in range(frames):
        t0 = time.perf_counter()
        fn(x)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    ts = np.array(ts)
    return {"p50_ms": round(float(np.percentile(ts, 50)), 4), "p99_ms": round(float(np.percentile(ts, 99)), 4),
            "mean_ms": round(float(ts.mean()), 4)}


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    dev = torch.device("cuda:0")
    B, N = 1, 8192
    x = torch.from_numpy(synthetic.ball(B, N, seed=7)).to(dev).transpose(1, 2).contiguous()
    res = {"config": "B=1 N=8192 backbone (PointNet2Msg_fast shallow1, out 384), eval, sequential frames", "frames": frames}
    state = None
    for eng in ("fused", "ops"):
        pu.set_engine(eng)
        m = backbones.PointNet2Msg_fast(backbones.default_cfg(dev), 384)
        init_weights(m, seed=0)
        m = m.to(dev).eval()
        state = m.state_dict()
        with torch.no_grad():
            res["ours_%s_eager" % eng] = run(lambda t: m(t), frames, x)
        res["ours_%s_graph" % eng] = run(GraphedForward(m), frames, x)
    pu.set_engine("ops")
    from oracle import ref_modules
    if ref_modules.available(cuda=True):
        rpu, rbb = ref_modules.load(cuda=True)
        r = rbb.PointNet2Msg_fast(backbones.default_cfg(dev), 384)
        r.load_state_dict(state)
        r = r.to(dev).eval()
        with torch.no_grad():
            res["reference_eager"] = run(lambda t: r(t), frames, x)
    print(json.dumps(res))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bench_latency.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
