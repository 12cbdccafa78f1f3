#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: point-clouds/sec (fwd+bwd, B=32 N=4096) of HOTrack's pointnet_lib
hot path (PointNet2Msg_fast backbone -> q1 -> q2 of HandTrackNet), one training step per "step"
(forward + backward + gradient all-reduce when N>1 + Adam), B=32 clouds of N=4096 points PER GPU
(weak scaling: clouds are independent, the only exchange is the gradient all-reduce).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--engine ops|fused]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...          (N > 1)

Prints ONE JSON line (rank 0).  `value` = clouds/s with the batch already resident in HBM; `e2e` =
the same step driven from pinned HOST buffers (H2D of the clouds and joints + D2H of the loss inside
the timed region).  `roofline` = the kernel of OURS with the largest share of the step, its
algorithmic bytes (SURVEY.md section 8d formulas) over its CUDA-event duration measured inside the
timed region, against MEASURED_PEAKS.json.  `cpu_baseline` = the reference's own CPU fallback path
(oracle/_ref/pyref imported with CUDA hidden) on a bounded sample, all host threads.

--impl reference runs the UNMODIFIED reference: its Python layer (oracle/_ref/pyref) on its own CUDA
kernels (oracle/_ref/libpn2_ref.so, compiled for sm_100a from /root/reference), torch.optim.Adam,
same step, same loss, same inputs.  HOTrack's pointnet_lib is GPU code, so the like-for-like
reference arm is that GPU path on the same B200; `--ref-device cpu` times its CPU fallback instead.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "point-clouds/sec (fwd+bwd, B=32 N=4096)"
UNIT = "clouds/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--engine", default="auto", choices=["auto", "ops", "fused"])
    ap.add_argument("--ref-device", default="cuda", choices=["cuda", "cpu"])
    ap.add_argument("--precision", default="auto", choices=["auto", "off", "all"],
                    help="fused engine: which stacks keep two-plane (hi+lo fp16) forward rows (hotrack_b200.fused.set_precise)")
    ap.add_argument("--batch", type=int, default=32, help="clouds per GPU")
    ap.add_argument("--points", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side-legs", action="store_true", help="skip the ops-engine / --precision all / config-5 latency legs")
    ap.add_argument("--no-graph", action="store_true", help="ours: dispatch every kernel from Python instead of replaying a CUDA graph")
    ap.add_argument("--cpu-sample", type=int, default=8, help="clouds in the cpu_baseline sample")
    ap.add_argument("--profile-mode", action="store_true",
                    help="for runs under ncu: skip the e2e leg, the per-kernel probe pass and the cpu baseline (the printed "
                         "line is then NOT a bench value)")
    return ap.parse_args()


# ------------------------------------------------------------------ helpers -------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms.  Started BEFORE the warm-up steps: nvidia-smi's
    own start-up (device enumeration, ~0.1-0.3 s) stalls the GPU it queries for milliseconds, and when it ran into the
    timed region it showed up as a 1.5-2x slower end-to-end leg on some boxes.  Only the samples whose timestamp lies
    inside the timed window are reported (all of them if the clock-skew-free match finds none)."""

    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def window_begin(self):
        self.t0 = time.time()

    def window_end(self):
        self.t1 = time.time()

    @staticmethod
    def _ts(text):
        import datetime
        for fmt in ("%Y/%m/%d %H:%M:%S.%f", "%Y/%m/%d %H:%M:%S"):
            try:
                return datetime.datetime.strptime(text, fmt).timestamp()
            except ValueError:
                pass
        return None

    def parse(self, text):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for line in text.splitlines():
            c = [t.strip() for t in line.split(",")]
            if len(c) < 10:
                continue
            try:
                rows.append((self._ts(c[0]), float(c[2]), float(c[3]),
                             {nm for nm, v in zip(names, c[6:10]) if v.lower().startswith("active")}))
            except ValueError:
                continue
        inside = [r for r in rows if r[0] is not None and self.t0 is not None and self.t1 is not None
                  and self.t0 - 0.2 <= r[0] <= self.t1 + 0.2]
        use = inside or rows
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if use:
            sm = sorted(r[1] for r in use)
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(r[2] for r in use),
                       reasons=sorted(set().union(*[r[3] for r in use])), samples=len(use),
                       window="timed region" if inside else "whole run")
        return out

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        text = self.f.read()
        os.unlink(self.f.name)
        return self.parse(text)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def alg_bytes(name, a):
    """Algorithmic bytes of one C-ABI call; `a` = the call's argument tuple."""
    if name == "pn2_furthest_point_sampling":
        b, n, m = a[:3]
        return b * n * 12 + b * m * 4
    if name == "pn2_ball_query":
        b, n, m, _, ns = a[:5]
        return b * n * 12 + b * m * 12 + b * m * ns * 4
    if name == "pn2_knn":
        b, n, m, k = a[:4]
        return b * n * 12 + b * m * 12 + b * n * k * 8
    if name == "pn2_three_nn":
        b, n, m = a[:3]
        return b * n * 12 + b * m * 12 + b * n * 24
    if name == "pn2_three_interpolate":
        b, c, m, n = a[:4]
        return b * c * m * 4 + b * n * 24 + b * c * n * 4
    if name == "pn2_three_interpolate_grad":
        b, c, n, m = a[:4]
        return b * c * m * 4 + b * n * 24 + b * c * n * 4
    if name in ("pn2_group_points", "pn2_group_points_grad"):
        b, c, n, s, k = a[:5]
        return b * c * n * 4 + b * s * k * 4 + b * c * s * k * 4
    if name in ("pn2_gather_points", "pn2_gather_points_grad"):
        b, c, n, s = a[:4]
        return b * c * n * 4 + b * s * 4 + b * c * s * 4
    if name == "pn2_adam_step":
        return a[0] * 28
    try:
        from hotrack_b200 import fused
        return fused.alg_bytes(name, a)
    except Exception:
        return None


_MS = None
_MS_CONST = {}


def loss_fn(src2, f11, f13):
    """mean(src2^2) + mean(f11^2) + mean(f13^2), as ONE autograd.Function of plain torch ops: each tensor is read once
    forward (one reduction kernel; the backbone output alone is 201 MB) and its gradient 2 t / numel is one elementwise
    kernel backward, the scalar bookkeeping is three small kernels forward and one backward -- torch's own composite
    (square().mean() per term, summed) takes five passes per tensor and a dozen scalar kernels.  The loss is harness, not
    the path, and both arms pay for it."""
    global _MS
    if _MS is None:
        import torch

        class MeanSquareSum(torch.autograd.Function):
            @staticmethod
            def forward(ctx, *xs):
                ctx.save_for_backward(*xs)
                key = (xs[0].device, tuple(x.numel() for x in xs))
                if key not in _MS_CONST:  # first call happens outside any graph capture (warm-up steps)
                    _MS_CONST[key] = torch.tensor([1.0 / x.numel() for x in xs], dtype=torch.float32, device=xs[0].device)
                ctx.inv = _MS_CONST[key]
                n = torch.stack([torch.linalg.vector_norm(x) for x in xs])
                return torch.dot(n * n, ctx.inv)

            @staticmethod
            def backward(ctx, g):
                s = ctx.inv * (g * 2.0)
                # x * (0-dim CUDA tensor) dispatches a broadcasting, non-vectorised elementwise kernel (3.2 TB/s on the
                # 201 MB tensor); the multi-tensor-apply kernel behind _foreach_mul streams it with 16-byte accesses
                return tuple(torch._foreach_mul([x], s[i])[0] for i, x in enumerate(ctx.saved_tensors))

        _MS = MeanSquareSum
    return _MS.apply(src2, f11, f13)


def make_inputs(B, N, rank):
    import importlib.util

    import torch

    # hotrack_b200/synthetic.py loaded as a plain file: importing the PACKAGE would dlopen libpn2b200.so, which the
    # reference arm must not do
    spec = importlib.util.spec_from_file_location("_pn2_synthetic", os.path.join(ROOT, "hotrack_b200", "synthetic.py"))
    synthetic = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(synthetic)
    xyz = torch.from_numpy(synthetic.ball(B, N, seed=1000 + rank))       # (B,N,3) canonicalised hand cloud
    kps = torch.from_numpy(synthetic.keypoints(B, 21, seed=1000 + rank))  # (B,21,3) jittered joints
    return xyz, kps


def build_model(impl, engine, dev):
    if impl == "ours":
        from hotrack_b200 import backbones, pointnet_utils as pu
        from hotrack_b200.handtrack_path import HandTrackPointPath, init_weights
        pu.set_engine(engine)
        model = HandTrackPointPath(backbones.default_cfg(dev))
        pu.set_engine("ops")
        init_weights(model, seed=0)
    else:
        # the reference's own classes on its own kernels; nothing of hotrack_b200 is imported on this arm
        from oracle import ref_path
        model = ref_path.RefPointPath(dev, cuda=(dev.type == "cuda"))
        ref_path.xavier_init(model, seed=0)
    model = model.to(dev)
    model.train()
    return model


# ------------------------------------------------------------------ CPU reference ------------
def cpu_reference(sample, N, steps=3, warmup=1):
    """The reference's CPU fallback path (pointnet_utils.py CUDA=False branch) on `sample` clouds."""
    import torch

    dev = torch.device("cpu")
    model = build_model("reference", None, dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-4)
    xyz, kps = make_inputs(sample, N, 0)
    x, k = xyz.transpose(1, 2).contiguous(), kps.transpose(1, 2).contiguous()
    ts = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        loss = loss_fn(*model(x, k)[:3])
        loss.backward()
        opt.step()
        float(loss.detach())
        if it >= warmup:
            ts.append(time.perf_counter() - t0)
    per_step = sum(ts) / len(ts)
    return {"value": round(sample / per_step, 3), "unit": UNIT, "cores": torch.get_num_threads(),
            "kind": "reference", "host_cpus": os.cpu_count(),
            "sample": "%d clouds x N=%d, %d fwd+bwd+Adam steps of the same path through the reference's own CPU "
                      "fallback (pointnet_utils.py CUDA=False branch + torch CPU), %.2f s/step" % (sample, N, steps,
                                                                                                  per_step)}


# ------------------------------------------------------------------ side legs ----------------
def _time_steps(torch, fn, n, flush):
    """mean device ms of n calls of fn(), L2 flushed (untimed) before each"""
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / n


def side_legs_ours(torch, dev, B, N, xyz_d, kps_d, flush, FromPoints):
    """Numbers quoted NEXT to the headline (rank 0, N=1): the same step with the fp32 'ops' engine (our index / gather /
    interpolate kernels + torch.nn fp32 convolutions: the 1e-5 parity configuration), the fused engine with two-plane
    rows everywhere (--precision all: fp32-class forward), and BASELINE config 5 (B=1, N=8192 backbone, eval, sequential
    frames, p50 wall latency, one CUDA-graph replay per frame)."""
    import numpy as np

    from hotrack_b200 import backbones, fused, pointnet_utils as pu, synthetic
    from hotrack_b200.handtrack_path import init_weights
    from hotrack_b200.train import GraphedForward, TrainStep

    out = {}
    for tag, engine, mode, graph, n in (("engine_ops_fp32", "ops", "auto", False, 5), ("precision_all", "fused", "all", True, 10)):
        fused.set_precise(mode)
        try:
            model = build_model("ours", engine, dev)
            ts = TrainStep(FromPoints(model), lambda o: loss_fn(*o[:3]), lr=1e-4, weight_decay=1e-4, graph=graph)
            for _ in range(3):
                ts(xyz_d, kps_d)
            torch.cuda.synchronize()
            ms = _time_steps(torch, lambda: ts(xyz_d, kps_d), n, flush)
            out[tag] = {"ms_per_step": round(ms, 4), "value": round(B / ms * 1e3, 2), "unit": UNIT, "steps": n,
                        "dispatch": "cuda-graph replay" if graph else "eager"}
            del ts, model
        finally:
            fused.set_precise("auto")
    pu.set_engine("fused")
    try:
        bb = backbones.PointNet2Msg_fast(backbones.default_cfg(dev), 384)
    finally:
        pu.set_engine("ops")
    init_weights(bb, seed=0)
    bb = bb.to(dev).eval()
    x5 = torch.from_numpy(synthetic.ball(1, 8192, seed=7)).to(dev).transpose(1, 2).contiguous()
    out["latency_config5"] = _latency(torch, np, GraphedForward(bb), x5)
    out["latency_config5"]["what"] = "B=1 N=8192 PointNet2Msg_fast forward, eval, fused engine, one CUDA-graph replay per frame"
    return out


def _palm(B):
    """a hand-sized 6-point palm template (wrist + five finger bases), metres -- what the reference takes from the MANO
    layer at zero pose (track_network.py:151-153)"""
    import numpy as np
    import torch
    base = np.array([[0, 0, 0], [0.03, 0.02, 0.005], [0.09, 0.03, 0.0], [0.095, 0.01, 0.002], [0.09, -0.01, 0.0],
                     [0.08, -0.03, -0.003]], dtype=np.float32)
    return torch.from_numpy(np.repeat(base[None], B, 0))


def full_network_legs(torch, dev, impl, B, N, flush):
    """The network AROUND the path, timed on both arms: (1) a full HandTrackNet training
    step -- handframe 'kp', loss = 10 kp + r + t as handtracknet_train_SimGrasp.yml:25-30, Adam -- at B x N; (2) one tracked
    frame (B=1, N=8192, eval, recurrence of track_network.py:159-217), p50 wall latency with a synchronisation per frame.
    Reference arm: its own hand_network.py on its own kernels, eager, CPU SVD and discarded attention included."""
    import numpy as np

    xyz, kps = make_inputs(B, N, 0)
    hp = (xyz * 0.1 + 0.3).to(dev)                       # metres, as the loader delivers them (scale 0.2: hand_network.py:99)
    jk = (kps * 0.1 + 0.3).to(dev)
    gk = (make_inputs(B, N, 1)[1] * 0.1 + 0.3).to(dev)
    palm = _palm(B).to(dev)
    flags = {"track_flag": False, "IKNet_flag": False}
    out = {}
    if impl == "ours":
        from hotrack_b200 import hand_network, pointnet_utils as pu
        from hotrack_b200.handtrack_path import init_weights
        from hotrack_b200.track import HandTracker
        from hotrack_b200.train import TrainStep
        from hotrack_b200.backbones import default_cfg
        cfg = default_cfg(dev)
        cfg["network"]["handframe"] = "kp"
        pu.set_engine("fused")
        try:
            net = hand_network.HandTrackNet(cfg)
        finally:
            pu.set_engine("ops")
        init_weights(net, seed=0)
        net = net.to(dev).train()

        class Full(torch.nn.Module):
            def __init__(self, net):
                super().__init__()
                self.net = net

            def forward(self, hp, jk, gk, palm):
                data = {"hand_points": hp, "jittered_hand_kp": jk, "gt_hand_kp": gk, "gt_hand_pose": {"palm_template": palm}}
                ret = self.net(data, flags)
                loss, _ = self.net.compute_loss(data, ret, flags)
                return 10 * loss["hand_pred_kp_loss"] + loss["hand_pred_r_loss"] + loss["hand_pred_t_loss"]

        ts = TrainStep(Full(net), lambda total: total, lr=1e-4, weight_decay=1e-4, graph=True)
        for _ in range(3):
            ts(hp, jk, gk, palm)
        torch.cuda.synchronize()
        ms = _time_steps(torch, lambda: ts(hp, jk, gk, palm), 10, flush)
        out["handtracknet_step"] = {"ms_per_step": round(ms, 4), "value": round(B / ms * 1e3, 2), "unit": UNIT, "steps": 10,
                                    "what": "full HandTrackNet (handframe kp) fwd + loss (10 kp + r + t) + bwd + Adam, B=%d N=%d, "
                                            "fused engine, CUDA-graph replay, GPU Kabsch" % (B, N)}
        net.eval()
        x1, k1 = make_inputs(1, 8192, 7)
        pts = (x1 * 0.1 + 0.3).to(dev)
        tr = HandTracker(net, _palm(1).to(dev), graph=True)
        tr.reset((k1 * 0.1 + 0.3).to(dev), pts)
        out["tracker_frame"] = _latency(torch, np, lambda t: tr.step(t), pts)
        out["tracker_frame"]["what"] = "one tracked frame: recurrence + full HandTrackNet forward, B=1 N=8192, one CUDA-graph replay"
    else:
        from oracle import ref_modules, ref_path
        _, _, rhn = ref_modules.load_full("ref")
        net = rhn.HandTrackNet(ref_modules.handtracknet_cfg(dev, "kp"))
        ref_path.xavier_init(net, seed=0)
        net = net.to(dev).train()
        opt = torch.optim.Adam(net.parameters(), lr=1e-4, weight_decay=1e-4)
        data = {"hand_points": hp, "jittered_hand_kp": jk, "gt_hand_kp": gk, "gt_hand_pose": {"palm_template": palm}}

        def step():
            opt.zero_grad(set_to_none=False)
            ret = net(data, flags)
            loss, _ = net.compute_loss(data, ret, flags)
            (10 * loss["hand_pred_kp_loss"] + loss["hand_pred_r_loss"] + loss["hand_pred_t_loss"]).backward()
            opt.step()

        for _ in range(2):
            step()
        torch.cuda.synchronize()
        ms = _time_steps(torch, step, 3, flush)
        out["handtracknet_step"] = {"ms_per_step": round(ms, 4), "value": round(B / ms * 1e3, 2), "unit": UNIT, "steps": 3,
                                    "what": "full HandTrackNet (handframe kp) fwd + loss + bwd + Adam, B=%d N=%d, reference "
                                            "hand_network.py on reference kernels, eager" % (B, N)}
        net.eval()
        x1, k1 = make_inputs(1, 8192, 7)
        pts = (x1 * 0.1 + 0.3).to(dev)
        state = {"last": (k1 * 0.1 + 0.3).to(dev) - pts.mean(dim=-2, keepdim=True)}
        palm1 = _palm(1).to(dev)

        def frame(t):  # track_network.py:159-217, branch without IKNet
            c = t.mean(dim=-2, keepdim=True)
            ret = net({"pred_palm_template": palm1, "hand_points": t, "jittered_hand_kp": state["last"] + c},
                      {"track_flag": True, "test_flag": True, "IKNet_flag": False})
            state["last"] = ret["pred_kp"] - c
            return ret["pred_kp"]

        out["tracker_frame"] = _latency(torch, np, frame, pts, frames=100)
        out["tracker_frame"]["what"] = "one tracked frame: recurrence + full HandTrackNet forward, B=1 N=8192, reference, eager"
    return out


def _latency(torch, np, fn, x, frames=300):
    import time as _t
    with torch.no_grad():
        for _ in range(20):
            fn(x)
        torch.cuda.synchronize()
        ts = []
        for _ in range(frames):
            t0 = _t.perf_counter()
            fn(x)
            torch.cuda.synchronize()  # frame t+1 needs frame t (track_network.py:163,217)
            ts.append((_t.perf_counter() - t0) * 1e3)
    ts = np.array(ts)
    return {"p50_ms": round(float(np.percentile(ts, 50)), 4), "p99_ms": round(float(np.percentile(ts, 99)), 4), "frames": frames}


# ------------------------------------------------------------------ main ----------------------
def main():
    args = parse()
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    B, N, K, W = args.batch, args.points, args.steps, max(args.warmup, 3)

    if args.impl == "reference" and rank != 0:
        return  # the reference is single-GPU: rank 0 alone runs it

    if args.impl == "reference" and args.ref_device == "cpu":
        cb = cpu_reference(args.cpu_sample, N, steps=max(1, min(K, 5)), warmup=1)
        line = {"metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": 0, "steps": K, "warmup": W,
                "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "impl": "reference", "cpu_baseline": cb,
                "config": {"workload": "HandTrackNet pointnet_lib path fwd+bwd+Adam, bounded CPU sample", "points": N},
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path for --impl ours)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ddp = world > 1 and args.impl == "ours"
    if ddp:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    engine = args.engine
    if args.impl == "ours" and engine == "auto":
        try:
            from hotrack_b200 import fused  # noqa: F401
            engine = "fused"
        except ImportError:
            engine = "ops"

    if args.impl == "ours" and engine == "fused":
        from hotrack_b200 import fused
        fused.set_precise(args.precision)
    model = build_model(args.impl, engine, dev)
    if args.impl == "ours":
        from hotrack_b200 import _lib, pointnet_utils as pu
        from hotrack_b200.train import TrainStep

        class FromPoints(torch.nn.Module):  # (B,N,3)/(B,21,3) -> the (B,3,N) layout canonicalize() hands the backbone
            def __init__(self, path):
                super().__init__()
                self.path = path

            def forward(self, xyz, kps):
                # t_contig = .transpose(1, 2).contiguous(), remembered for the duration of the scope: the modules inside
                # want the point-major twin back (FPS, ball query, kNN, three-NN) and find the original instead of
                # transposing a second time
                with pu.coord_scope():
                    return self.path(pu.t_contig(xyz), pu.t_contig(kps))

        train = TrainStep(FromPoints(model), lambda out: loss_fn(*out[:3]), lr=1e-4, weight_decay=1e-4,
                          graph=not args.no_graph)
        flat = train.flat
    else:
        _lib = None
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-4)

    xyz_h, kps_h = make_inputs(B, N, rank)
    xyz_h, kps_h = xyz_h.pin_memory(), kps_h.pin_memory()
    xyz_d, kps_d = xyz_h.to(dev), kps_h.to(dev)
    loss_h = torch.zeros(1).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step(xyz, kps):
        if args.impl == "ours":
            return train(xyz, kps)
        x = xyz.transpose(1, 2).contiguous()   # (B,3,N), the layout canonicalize() hands the backbone
        k = kps.transpose(1, 2).contiguous()
        opt.zero_grad(set_to_none=False)
        loss = loss_fn(*model(x, k)[:3])
        loss.backward()
        opt.step()
        return loss.detach()

    def barrier():
        if ddp:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n_steps, from_host):
        """Device time of n_steps steps: first step's start event -> last step's end event, minus the L2-flush
        kernels in between (each bracketed by its own events).  GPU idle gaps caused by a slow host ARE
        counted.  from_host: every step copies its inputs from pinned host memory and writes its loss back to
        pinned host memory (asynchronously; the host reads it one step late, nothing waits inside the loop)."""
        marks, flushes = [], []
        barrier()
        for _ in range(n_steps):
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            flush.zero_()  # evict L2 between steps (excluded from the step time)
            f1.record()
            flushes.append((f0, f1))
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            if from_host:
                if args.impl == "ours" and train.use_graph:
                    # H2D straight into the replayed graph's input buffers (TrainStep copies whatever it is given)
                    loss = step(xyz_h, kps_h)
                else:
                    loss = step(xyz_h.to(dev, non_blocking=True), kps_h.to(dev, non_blocking=True))
                loss_h.copy_(loss.reshape(1), non_blocking=True)
            else:
                step(xyz_d, kps_d)
            e.record()
            marks.append((s, e))
        barrier()
        ms = marks[0][0].elapsed_time(marks[-1][1]) - sum(a.elapsed_time(b) for a, b in flushes[1:])
        if ddp:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()  # before the warm-up: its start-up cost must not land in the timed region
    for _ in range(W):
        step(xyz_d, kps_d)
    barrier()

    if sampler:
        sampler.window_begin()
    ms_dev = timed(K, from_host=False)
    ms_e2e = ms_dev if args.profile_mode else timed(K, from_host=True)
    if sampler:
        sampler.window_end()
    clocks = sampler.stop() if sampler else None

    # Per-kernel device times of OUR kernels, CUDA events around every C-ABI call.  A replayed CUDA graph
    # cannot be bracketed kernel by kernel, so this pass dispatches the same step eagerly (same kernels, same
    # shapes, same process, right after the timed region); the launch count per step comes from it too.
    probe, launches = None, 0
    if _lib and not args.profile_mode:  # every rank steps (the gradient all-reduce needs all of them), rank 0 records
        n_probe = 3
        train.use_graph, was = False, train.use_graph
        step(xyz_d, kps_d)
        barrier()
        n0 = _lib.lib.pn2_launch_count()
        if rank == 0:
            _lib.PROBE = {}
        for _ in range(n_probe):
            step(xyz_d, kps_d)
        barrier()
        probe, _lib.PROBE = _lib.PROBE, None
        launches = (_lib.lib.pn2_launch_count() - n0) // n_probe * K * (world if ddp else 1)
        train.use_graph = was

    if rank != 0:
        if ddp:
            dist.destroy_process_group()
        return

    clouds_per_step = B * (world if ddp else 1)
    value = clouds_per_step * K / (ms_dev / 1e3)
    e2e = clouds_per_step * K / (ms_e2e / 1e3)
    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world if ddp else 1, "steps": K, "warmup": W,
        "ms_per_step": round(ms_dev / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        # the arithmetic the grouped MLP runs in (not a precision claim): fp16 forward rows -- as two-plane hi+lo pairs in
        # SA1-SA3 and FP3's first layer --, bf16 gradient rows, fp32 accumulation and BatchNorm; indices / gathers fp32
        "dtype": ("f16 fwd rows (two-plane f16x2 in SA1-3, FP3.0) / bf16 grad rows / f32 accumulate"
                  if (args.impl == "ours" and engine == "fused") else "f32(tf32 conv)" if args.impl == "reference" else "f32"),
        "data": "synthetic",
        "config": {"workload": "HandTrackNet pointnet_lib path (PointNet2Msg_fast shallow1 backbone -> q1 -> q2, 21 joints), "
                               "train step fwd+bwd+Adam, B=%d N=%d per GPU" % (B, N),
                   "clouds_per_gpu": B, "points": N, "engine": engine if args.impl == "ours" else "reference-cuda",
                   "precision": args.precision if (args.impl == "ours" and engine == "fused") else None,
                   "parallelism": "dp%d" % (world if ddp else 1),
                   "dispatch": ("cuda-graph replay" if (args.impl == "ours" and not args.no_graph) else "eager"),
                   "l2": "256 MiB flush between timed steps (untimed); per-step working set >> 126 MB L2"},
        "e2e": {"value": round(e2e, 2), "unit": UNIT, "ms_per_step": round(ms_e2e / K, 4),
                "h2d_bytes_per_step": int(xyz_h.numel() * 4 + kps_h.numel() * 4) * (world if ddp else 1),
                "d2h_bytes_per_step": 4 * (world if ddp else 1)},
        "gpu_launches": int(launches),  # kernels of libpn2b200.so inside the K timed steps (all ranks), counted by the library
        "clocks": clocks,
    }
    if args.profile_mode:
        line["invalid"] = "profile mode: e2e/probe/cpu legs skipped, not a bench value"
    if args.impl == "reference":
        line["impl"] = "reference"
        line["cpu_baseline"] = {"value": line["value"], "unit": UNIT, "cores": torch.get_num_threads(), "kind": "reference",
                                "sample": "full workload; NOTE the reference's pointnet_lib is CUDA code: this arm runs its "
                                          "own kernels (compiled for sm_100a) + its Python layer + torch.optim.Adam on the "
                                          "same B200, host threads only drive launches"}
        line["e2e"]["note"] = "reference GPU path driven from pinned host buffers"
        if not args.profile_mode:
            import numpy as np
            bb = model.bhand.eval()
            x5, _ = make_inputs(1, 8192, 7)
            x5 = x5.to(dev).transpose(1, 2).contiguous()
            line["latency_config5"] = _latency(torch, np, lambda t: bb(t), x5)
            line["latency_config5"]["what"] = "B=1 N=8192 PointNet2Msg_fast forward, eval, reference modules on reference kernels, eager"
            if not args.no_side_legs:
                try:
                    del model, opt
                    torch.cuda.empty_cache()
                    line.update(full_network_legs(torch, dev, "reference", B, N, flush))
                except Exception as ex:
                    line["side_legs_error"] = repr(ex)[:300]
    else:
        # dominant kernel of OURS inside the timed region
        peak, peak_src = peaks()
        best = None
        torch.cuda.synchronize()
        tot = {}
        for name, recs in (probe or {}).items():
            by_shape = {}
            for a, s, e in recs:
                # group by the size arguments only (device pointers and float scalars differ from call to call)
                key = tuple(x for x in a if isinstance(x, int) and abs(x) < (1 << 31))[:6]
                by_shape.setdefault(key, []).append((a, s.elapsed_time(e)))
            for key, lst in by_shape.items():
                # median x count: one hiccup (a host-side stall between the two event records) must not promote a
                # 10 us kernel to the top of the table
                ts_ = sorted(x[1] for x in lst)
                t = ts_[len(ts_) // 2] * len(ts_)
                tot[(name, key)] = (t, lst)
        shares = []
        ours_ms = sum(t for t, _ in tot.values())
        for (name, key), (t, lst) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
            shares.append({"kernel": name, "shape": list(key), "calls_per_step": len(lst) // 3,
                           "us_per_call": round(t / len(lst) * 1e3, 2),
                           "share_of_our_kernels": round(t / ours_ms, 4)})
            if best is None:
                nb = alg_bytes(name, lst[0][0])
                # the roofline line is about the dominant DATA-MOVING kernel: launches that touch < 8 MB are
                # launch-latency bound and their event timing is dominated by host submission jitter
                if nb and nb >= (8 << 20):
                    avg_ms = t / len(lst)
                    ach = nb / (avg_ms * 1e-3) / 1e9
                    best = {"bound": "hbm", "kernel": name, "shape": list(key), "achieved": round(ach, 2), "peak": peak,
                            "unit": "GB/s", "frac": round(ach / peak, 5), "traffic": None,
                            "alg_bytes_per_launch": int(nb), "avg_launch_us": round(avg_ms * 1e3, 3), "peak_source": peak_src,
                            "share_of_our_kernels": round(t / ours_ms, 4),
                            "timing": "CUDA events around each launch, eager probe pass after the timed region"}
        if best is not None:
            # dram__bytes_read + write of the same kernel / shape from the committed ncu --set full capture
            # (profiles/rNN_traffic.json: "<entry point>:<size args>" -> bytes per launch), null when not captured
            try:
                tr = {}
                for tag in ("r01", "r02"):  # later rounds override
                    fn = os.path.join(ROOT, "profiles", tag + "_traffic.json")
                    if os.path.exists(fn):
                        tr.update(json.load(open(fn)))
                key = "%s:%s" % (best["kernel"], ",".join(str(v) for v in best["shape"]))
                if key in tr:
                    best["traffic"] = tr[key]["dram_bytes"]
                    best["traffic_source"] = tr[key].get("source")
            except (OSError, ValueError):
                pass
        line["roofline"] = best
        line["kernel_shares"] = shares[:12]
        # the same probe aggregated per kernel ENTRY POINT (all shapes): which function the step spends its time in, and
        # what that function achieves over all of its launches
        fam = {}
        for (name, key), (t, lst) in tot.items():
            f = fam.setdefault(name, [0.0, 0, 0])
            f[0] += t
            f[1] += len(lst)
            nb = alg_bytes(name, lst[0][0])
            f[2] += (nb or 0) * len(lst)
        line["kernel_families"] = [
            {"kernel": name, "launches_per_step": cnt // 3, "us_per_step": round(t / 3 * 1e3, 1),
             "share_of_our_kernels": round(t / ours_ms, 4),
             "achieved_gbs": round(nbytes / (t * 1e-3) / 1e9, 1) if nbytes else None,
             "frac_of_hbm_peak": round(nbytes / (t * 1e-3) / 1e9 / peak, 4) if nbytes else None}
            for name, (t, cnt, nbytes) in sorted(fam.items(), key=lambda kv: -kv[1][0])][:10]
        if not ddp and not args.profile_mode and not args.no_side_legs and engine == "fused":
            try:
                line.update(side_legs_ours(torch, dev, B, N, xyz_d, kps_d, flush, FromPoints))
                line.update(full_network_legs(torch, dev, "ours", B, N, flush))
            except Exception as ex:  # a side leg must never take the headline down with it
                line["side_legs_error"] = repr(ex)[:300]
        if not ddp and not args.no_cpu_baseline and not args.profile_mode:  # rank 0 at N=1 only
            try:
                line["cpu_baseline"] = cpu_reference(args.cpu_sample, N, steps=3, warmup=1)
            except Exception as ex:  # oracle/_ref not staged
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                                        "sample": "unavailable: %s" % ex}
    print(json.dumps(line), flush=True)
    if ddp:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
