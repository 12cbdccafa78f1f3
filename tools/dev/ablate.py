"""Where does the step time REALLY go?  Graph-replayed bench step with one kernel family switched off at a time.

The ncu launch list serialises kernels and measures each with a cold cache; inside the replayed graph kernels overlap
(side streams) and run warm.  This script times the real thing: the bench step (B=32, N=4096, graph replay, L2 flushed
between steps) once as it is and once per family with that family's C-ABI calls replaced by no-ops -- the results are
garbage, the time difference is the family's true cost on the critical path.  Development tool, not a bench value.

usage: python tools/dev/ablate.py            (runs every family in a subprocess each)
       PN2_ABLATE=wgrad python tools/dev/ablate.py --one
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

FAMILIES = {
    "none": [],
    "wgrad": ["pn2_mlp_gemm_wgrad"],
    "coefs": ["pn2_bn_bwd_coefs"],
    "dgrad": ["pn2_mlp_gemm_dgrad"],
    "fwd_gemm": ["pn2_mlp_gemm_fwd_bn_x2", "pn2_mlp_gemm_fwd_x2"],
    "build_rows": ["pn2_sa_build_rows_x2", "pn2_fp_build_rows_x2", "pn2_to_rows_x2"],
    "pool": ["pn2_pool_fwd_x2", "pn2_pool_bwd"],
    "rows_bwd": ["pn2_sa_rows_bwd", "pn2_fp_rows_bwd"],
    "center": ["pn2_mlp_center"],
    "loss_small": [],   # loss on f13 only: no dense gradient on the 201 MB backbone output
    "no_wgrad_stream": [],
    "pdl_off": [],
    "prio_off": [],
    "pdl_prio_off": [],
    "wg_min1": [], "wg_min2": [], "wg_min8": [], "wg_min16": [], "wgs2": [], "wgs3": [], "no_prefetch": [], "fps512": [], "fps256": [], "fps1024": [], "maxbn128": [],
}
ENV = {"no_wgrad_stream": {"PN2_WGRAD_STREAM": "0"}, "pdl_off": {"PN2_PDL": "0"}, "prio_off": {"PN2_PRIO": "0"},
       "pdl_prio_off": {"PN2_PDL": "0", "PN2_PRIO": "0"},
       "wg_min1": {"PN2_WG_MIN_STAGES": "1"}, "wg_min2": {"PN2_WG_MIN_STAGES": "2"}, "wg_min8": {"PN2_WG_MIN_STAGES": "8"},
       "wg_min16": {"PN2_WG_MIN_STAGES": "16"}, "no_prefetch": {"PN2_SEARCH_PREFETCH": "0"}, "wgs2": {"PN2_WGRAD_STREAMS": "2"}, "wgs3": {"PN2_WGRAD_STREAMS": "3"}, "fps512": {"PN2_FPS_THREADS": "512"}, "fps256": {"PN2_FPS_THREADS": "256"}, "fps1024": {"PN2_FPS_THREADS": "1024"}, "maxbn128": {"PN2_TC_MAXBN": "128"}}


def one():
    import torch

    import bench
    from hotrack_b200 import _lib, fused, pointnet_utils as pu
    from hotrack_b200.train import TrainStep

    name = os.environ.get("PN2_ABLATE", "none")
    skip = set(FAMILIES[name])
    if skip:
        real = _lib.call

        def call(fn, *a):
            if fn in skip:
                return
            real(fn, *a)

        _lib.call = call
        fused._lib.call = call
    dev = torch.device("cuda", 0)
    fused.set_precise("auto")
    model = bench.build_model("ours", "fused", dev)

    class FromPoints(torch.nn.Module):
        def __init__(self, path):
            super().__init__()
            self.path = path

        def forward(self, xyz, kps):
            with pu.coord_scope():
                return self.path(pu.t_contig(xyz), pu.t_contig(kps))

    if name == "loss_small":
        loss = lambda out: out[2].square().mean() + out[1].square().mean()
    else:
        loss = lambda out: bench.loss_fn(*out[:3])
    train = TrainStep(FromPoints(model), loss, lr=1e-4, weight_decay=1e-4, graph=True)
    xyz, kps = bench.make_inputs(32, 4096, 0)
    xyz, kps = xyz.to(dev), kps.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(5):
        train(xyz, kps)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        train(xyz, kps)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    print("ABLATE %-16s median %.3f ms  min %.3f ms" % (name, ts[len(ts) // 2], ts[0]), flush=True)


def main():
    if "--one" in sys.argv:
        one()
        return
    names = [a for a in sys.argv[1:] if a in FAMILIES] or list(FAMILIES)
    for name in names:
        env = dict(os.environ, PN2_ABLATE=name)
        env.update(ENV.get(name, {}))
        t0 = time.time()
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=env, capture_output=True, text=True,
                           timeout=300)
        lines = [l for l in r.stdout.splitlines() if l.startswith("ABLATE")]
        print(lines[0] if lines else "ABLATE %-16s FAILED rc=%d %s" % (name, r.returncode, r.stderr[-300:]),
              "(%.0f s)" % (time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
