"""Graph replay vs eager dispatch of TrainStep on a tiny batch (the situation of
tests/test_fused_gpu.py::test_train_step_graph_replay_matches_eager), printing the per-step losses and the parameter
deviation -- to tell run-to-run noise (atomics order; B=2 clouds make BatchNorm chaotic) from a real ordering bug.
Runs eager twice too: eager-vs-eager is the noise floor.  Env toggles (PN2_PDL, PN2_PRIO, PN2_WGRAD_STREAM) apply."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import clouds
    from hotrack_b200 import backbones, pointnet_utils as pu
    from hotrack_b200.handtrack_path import HandTrackPointPath, init_weights
    from hotrack_b200.train import TrainStep

    cuda = torch.device("cuda", 0)
    B, N = 2, 1024
    x = torch.from_numpy(clouds.ball(B, N, seed=6)).to(cuda).transpose(1, 2).contiguous()
    k = torch.from_numpy(clouds.keypoints(B, 21, seed=6)).to(cuda).transpose(1, 2).contiguous()
    finals = []
    for graph in (False, False, True, True):
        pu.set_engine("fused")
        m = HandTrackPointPath(backbones.default_cfg(cuda))
        pu.set_engine("ops")
        init_weights(m, seed=0)
        m = m.to(cuda).train()
        ts = TrainStep(m, lambda out: sum(v.square().mean() for v in out[:3]), lr=1e-3, graph=graph)
        losses = [float(ts(x, k)) for _ in range(5)]
        finals.append((ts.flat.data.clone(), losses))
        print("graph=%d losses %s" % (graph, " ".join("%.4f" % l for l in losses)), flush=True)
    ref = finals[0][0]
    for i in range(1, 4):
        print("run %d vs run 0: params rel %.3e" % (i, ((finals[i][0] - ref).norm() / ref.norm()).item()))
    env = {k: os.environ.get(k) for k in ("PN2_PDL", "PN2_PRIO", "PN2_WGRAD_STREAM")}
    print("env", env)


if __name__ == "__main__":
    main()
