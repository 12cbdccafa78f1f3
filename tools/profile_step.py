"""Kernel-level breakdown of one training step of the path (torch.profiler, CUDA activities):
python tools/profile_step.py [ours|reference] [engine] -> gpurun_out/profile_<impl>_<engine>.txt
Development aid (numbers under a profiler are not bench values)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    impl = sys.argv[1] if len(sys.argv) > 1 else "ours"
    engine = sys.argv[2] if len(sys.argv) > 2 else "ops"
    B, N = int(os.environ.get("B", 32)), int(os.environ.get("N", 4096))
    dev = torch.device("cuda:0")
    model = bench.build_model(impl, engine, dev)
    if impl == "ours":
        from hotrack_b200.flat import FlatAdam, FlatParams
        flat = FlatParams(model)
        opt = FlatAdam(flat)
    else:
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-4)
    xyz, kps = bench.make_inputs(B, N, 0)
    x = xyz.to(dev).transpose(1, 2).contiguous()
    k = kps.to(dev).transpose(1, 2).contiguous()

    def step():
        if impl == "ours":
            flat.zero_grad()
        else:
            opt.zero_grad(set_to_none=False)
        loss = bench.loss_fn(*model(x, k)[:3])
        loss.backward()
        if impl == "ours":
            opt.step(flat.allreduce_grads())
        else:
            opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    txt = prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=70)
    with open(os.path.join(out, "profile_%s_%s.txt" % (impl, engine)), "w") as f:
        f.write("3 steps, B=%d N=%d\n" % (B, N))
        f.write(txt)
    print(txt[:6000])


if __name__ == "__main__":
    main()
