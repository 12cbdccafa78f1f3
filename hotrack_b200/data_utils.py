"""Data-loader farthest-point down-sampling: mirror of the reference's datasets/data_utils.py:226-245
(SURVEY.md section 8f, row N1 -- the first caller outside the network that sits on the hot path).

Every dataset calls it twice per frame (hand and object cloud: datasets/SimGrasp_dataset.py:60,68,
HO3D_dataset.py:173,176, DexYCB_dataset.py:169,174) with B = 1, up to 5*npoint points, npoint = 512.
Same behaviour: clouds larger than 5*npoint are first cut to a random 5*npoint subset (numpy's global RNG, as the
reference), then FPS runs on the GPU and indices into the ORIGINAL array come back as numpy.  ``sample_batch``
is the addition: it down-samples a list of clouds with ONE kernel launch and one D2H copy instead of one launch
and two PCIe round trips per cloud (the FPS kernel takes a batch; ragged clouds are padded by repeating point 0,
which can never be selected before every real point has been).
"""
import numpy as np
import torch

from . import pointnet2_utils as futils


def farthest_point_sample(xyz, npoint, device):
    """xyz: (N,3) numpy -> (npoint,) numpy indices.  Reference datasets/data_utils.py:226-245.
    There is no CPU branch (the reference falls back to RANDOM sampling without a GPU)."""
    xyz = np.asarray(xyz)
    if len(xyz) > 5 * npoint:
        idx = np.random.permutation(len(xyz))[:5 * npoint]
        t = torch.as_tensor(xyz[idx], dtype=torch.float32).to(device).reshape(1, -1, 3)
        sel = futils.furthest_point_sample(t, npoint).long().cpu().numpy().reshape(-1)
        return idx[sel]
    t = torch.as_tensor(xyz, dtype=torch.float32).to(device).reshape(1, -1, 3)
    return futils.furthest_point_sample(t, npoint).long().reshape(-1).cpu().numpy()


def sample_batch(clouds, npoint, device):
    """clouds: list of (N_i,3) numpy arrays -> list of (npoint,) numpy index arrays, one FPS launch for all.
    Each result equals ``farthest_point_sample(cloud_i, npoint, device)`` run on the same (sub-sampled) points
    provided N_i >= npoint distinct points exist; the random 5*npoint pre-selection is drawn per cloud in order."""
    subs, picks = [], []
    for c in clouds:
        c = np.asarray(c, dtype=np.float32)
        if len(c) > 5 * npoint:
            idx = np.random.permutation(len(c))[:5 * npoint]
            subs.append(c[idx])
            picks.append(idx)
        else:
            subs.append(c)
            picks.append(None)
    n_max = max(len(s) for s in subs)
    batch = np.empty((len(subs), n_max, 3), dtype=np.float32)
    for i, s in enumerate(subs):
        batch[i, :len(s)] = s
        batch[i, len(s):] = s[0]  # padding = copies of point 0 (distance 0 from the first sample: never farthest)
    t = torch.from_numpy(batch).to(device)
    sel = futils.furthest_point_sample(t, npoint).long().cpu().numpy()
    out = []
    for i, s in enumerate(subs):
        ii = sel[i]
        out.append(ii if picks[i] is None else picks[i][ii])
    return out
