"""Generate tests/golden/*.npz by running the REFERENCE's own kernels (oracle/_ref)
on a GPU.  Run on the B200 box:

    python tools/make_golden.py gpurun_out/golden

then copy the files to tests/golden/ and commit them.  Inputs are stored with the
outputs so the fixtures do not depend on any RNG implementation.  Every output
is also checked against the CPU oracle before it is written (a mismatch aborts).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import clouds  # noqa: E402
from oracle import pn2_oracle as orc  # noqa: E402
from oracle import ref_lib  # noqa: E402


def case(name, xyz, npoint, radius, nsample, kps, k, C, out_dir):
    dev = torch.device("cuda:0")
    x = torch.from_numpy(xyz).to(dev)
    B, N, _ = xyz.shape
    fps, temp = ref_lib.furthest_point_sample(x, npoint, return_temp=True)
    fps_np = fps.cpu().numpy()
    new_xyz = np.stack([xyz[b][fps_np[b]] for b in range(B)])
    nx = torch.from_numpy(new_xyz).to(dev)
    ball = ref_lib.ball_query(radius, nsample, x, nx)
    kp = torch.from_numpy(kps).to(dev)
    knn_d2, knn_idx = ref_lib.knn(k, kp, x)
    nn_d2, nn_idx = ref_lib.three_nn(x, nx)
    rng = np.random.RandomState(5)
    feats = rng.randn(B, C, npoint).astype(np.float32)
    d = torch.sqrt(nn_d2)
    w = 1.0 / (d + 1e-8)
    w = (w / w.sum(-1, keepdim=True)).contiguous()
    interp = ref_lib.three_interpolate(torch.from_numpy(feats).to(dev), nn_idx, w)
    grouped = ref_lib.group_points(x.transpose(1, 2).contiguous(), ball)
    gathered = ref_lib.gather_points(x.transpose(1, 2).contiguous(), fps)
    torch.cuda.synchronize()
    out = dict(xyz=xyz, npoint=npoint, radius=np.float32(radius), nsample=nsample, kps=kps, k=k,
               fps_idx=fps_np, fps_temp=temp.cpu().numpy(), new_xyz=new_xyz, ball_idx=ball.cpu().numpy(),
               knn_d2=knn_d2.cpu().numpy(), knn_idx=knn_idx.cpu().numpy(), nn_d2=nn_d2.cpu().numpy(),
               nn_idx=nn_idx.cpu().numpy(), feats=feats, weight=w.cpu().numpy(), interp=interp.cpu().numpy(),
               grouped_xyz=grouped.cpu().numpy(), gathered_xyz=gathered.cpu().numpy())
    # the restatement must agree with the reference before the fixture is accepted
    np.testing.assert_array_equal(orc.furthest_point_sample(xyz, npoint), out["fps_idx"])
    np.testing.assert_array_equal(orc.ball_query(radius, nsample, xyz, new_xyz), out["ball_idx"])
    od2, oidx = orc.knn(k, kps, xyz)
    np.testing.assert_array_equal(oidx, out["knn_idx"])
    np.testing.assert_array_equal(od2, out["knn_d2"])
    od2, oidx = orc.three_nn(xyz, new_xyz)
    np.testing.assert_array_equal(oidx, out["nn_idx"])
    np.testing.assert_array_equal(od2, out["nn_d2"])
    np.testing.assert_array_equal(orc.three_interpolate(feats, out["nn_idx"], out["weight"]), out["interp"])
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), **out)
    print("wrote", name, {k_: np.asarray(v).shape for k_, v in out.items()})


def main():
    out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    assert ref_lib.available(), "oracle/_ref/libpn2_ref.so missing: run `make -C oracle ref` first"
    # BASELINE.json config 1
    case("config1_ball_b2_n1024", clouds.ball(2, 1024, seed=1), 256, 0.1, 32, clouds.keypoints(2, 21, seed=1), 16, 8, out_dir)
    # exact ties / duplicates (FPS tie rule, strict radius, stable kNN order); reference block size 512
    case("lattice_b2_n1000", clouds.lattice(2, 1000, seed=2), 128, 0.125, 16, clouds.lattice(2, 21, seed=3), 32, 4, out_dir)
    # SA2-like small cloud, block size 256
    case("shell_b2_n256", clouds.shell(2, 256, seed=4), 128, 0.2, 32, clouds.keypoints(2, 21, seed=4), 4, 4, out_dir)


if __name__ == "__main__":
    main()
