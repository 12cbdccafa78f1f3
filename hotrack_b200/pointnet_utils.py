"""SA / FP modules and functional shims: mirror of the reference's network/models/pointnet_utils.py.

Same class names, constructor arguments, ``forward`` signatures, tensor layouts and ``state_dict``
keys (``conv_blocks.i.j``, ``bn_blocks.i.j``, ``mlp_convs.j``, ``mlp_bns.j``; reference
pointnet_utils.py:352-365,415-421,473-479,518-533), so ``backbones.py`` / ``hand_network.py`` and
the reference's checkpoints work unchanged when this module is the ``pointnet_utils`` they import.

Two engines sit behind the same modules:

* ``"ops"``   -- every index / gather / interpolate op is one of the sm_100a kernels of
                 libpn2b200.so; the 1x1 convolutions and BatchNorm stay ``torch.nn`` fp32 exactly as in
                 the reference (pointnet_utils.py:399-403).  This is the fp32 parity configuration.
* ``"fused"`` -- the whole grouped MLP (gather -> conv/BN/ReLU stack -> max-pool, and three-NN
                 interpolate -> concat -> conv/BN/ReLU stack) runs in hand-written 16-bit (fp16 forward / bf16 gradient) tensor-core
                 kernels (``hotrack_b200.fused``).  Indices are identical (they depend on coordinates
                 only); features agree to bf16 rounding.

There is no CPU path (the reference's pure-torch fallback, pointnet_utils.py:26-32,40-43,50-53,
126-137,156-167, is a different algorithm: random FPS start, inclusive radius test); CPU tensors
raise.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pointnet2_utils as futils

_ENGINE = "ops"


def set_engine(name):
    """Select the default engine for modules constructed afterwards ('ops' or 'fused')."""
    global _ENGINE
    if name not in ("ops", "fused"):
        raise ValueError("engine must be 'ops' or 'fused'")
    _ENGINE = name


def get_engine():
    return _ENGINE


# ------------------------------------------------------------------ functional shims ----------
def knn_point(k, pos2, pos1):
    """k nearest ``pos1`` (B,N,3) points of every ``pos2`` (B,M,3) query -> (dist (B,M,k), idx int64).
    Reference pointnet_utils.py:12-24."""
    val, idx = futils.knn(k, pos2, pos1)
    return val, idx.long()


def three_nn(xyz1, xyz2):
    """Reference pointnet_utils.py:35-38: (dist (B,N,3), idx int64 (B,N,3)) of xyz1 in xyz2."""
    dists, idx = futils.three_nn(xyz1, xyz2)
    return dists, idx.long()


def three_interpolate(points, idx, weight):
    """points (B,C,M), idx (B,N,3), weight (B,N,3) -> (B,C,N).  Reference pointnet_utils.py:46-48."""
    return futils.three_interpolate(points, idx.int(), weight)


def square_distance(src, dst):
    """(B,N,M) squared distances by the expansion the reference uses (pointnet_utils.py:56-78)."""
    d = -2.0 * torch.matmul(src, dst.transpose(1, 2))
    d = d + (src * src).sum(-1).unsqueeze(-1)
    return d + (dst * dst).sum(-1).unsqueeze(-2)


def index_points(points, idx):
    """points (B,N,C), idx (B,S) or (B,S,K) -> (B,S[,K],C).  Exact copies (pointnet_utils.py:80-97);
    here through the gather kernels of libpn2b200.so rather than advanced indexing."""
    B = points.shape[0]
    flat = idx.reshape(B, -1)
    out = futils.gather_operation(points.transpose(1, 2).contiguous(), flat.int())  # (B,C,S*K)
    return out.transpose(1, 2).reshape(*idx.shape, points.shape[-1])


def gather_operation(feature, idx):
    """feature (B,C,N), idx (B,S) -> (B,C,S).  Reference pointnet_utils.py:100-103."""
    return futils.gather_operation(feature.contiguous(), idx.int())


def group_operation(feature, idx):
    """feature (B,C,N), idx (B,S,K) -> (B,C,S,K).  Reference pointnet_utils.py:106-109."""
    return futils.grouping_operation(feature.contiguous(), idx.int())


def farthest_point_sample(xyz, npoint):
    """xyz (B,N,3) -> (B,npoint) int64, first index 0.  Reference pointnet_utils.py:112-125."""
    return futils.furthest_point_sample(xyz, npoint).long()


def query_ball_point(radius, nsample, xyz, new_xyz):
    """(B,S,nsample) int64 ball-query indices.  Reference pointnet_utils.py:140-154."""
    return futils.ball_query(radius, nsample, xyz, new_xyz).long()


def sample_and_group_all(xyz, points):
    """One group holding every point, channel order [xyz, points] (pointnet_utils.py:170-186)."""
    B, N, C = xyz.shape
    new_xyz = torch.zeros(B, 1, C, device=xyz.device, dtype=xyz.dtype)
    grouped = xyz.view(B, 1, N, C)
    if points is not None:
        grouped = torch.cat([grouped, points.view(B, 1, N, -1)], dim=-1)
    return new_xyz, grouped


def _carry(src, dst):
    """Keep the fused engine's bf16 row form (if any) attached across a reshape of the same values."""
    r = getattr(src, "_pn2_rows", None)
    if r is not None and dst.numel() == r.numel:
        dst._pn2_rows = r
    return dst


# Coordinate layouts.  The reference API hands coordinates around channel-major, (B,3,N), while the neighbour-search
# kernels take point-major (B,N,3): every module transposes what it is given (pointnet_utils.py:380,383,437 of the
# reference), so one forward pass copies the same three coordinate sets again and again (l0 four times, l1 three times).
# Inside a ``coord_scope()`` -- one forward pass, during which coordinates do not change -- the transposed twin of a tensor
# is made once and found again by (address, shape, strides, version); entries keep their source alive, so an address
# cannot be handed to another tensor while the scope is open.  Outside a scope nothing is remembered.
_MEMO = None


class coord_scope:
    def __enter__(self):
        global _MEMO
        self.outer = _MEMO is not None
        if not self.outer:
            _MEMO = {}
        return self

    def __exit__(self, *exc):
        global _MEMO
        if not self.outer:
            _MEMO = None
        return False


def _memo_key(t):
    return (t.data_ptr(), tuple(t.shape), tuple(t.stride()), t._version, t.dtype)


def t_contig(t):
    """``t.transpose(1, 2).contiguous()`` of a 3-D tensor, memoised while a ``coord_scope`` is open."""
    v = t.transpose(1, 2)
    if v.is_contiguous():
        return v
    if _MEMO is None or t.requires_grad:
        return v.contiguous()
    e = _MEMO.get(_memo_key(t))
    if e is not None:
        return e[1]
    out = v.contiguous()
    _MEMO[_memo_key(t)] = (t, out)
    if t.is_contiguous():
        _MEMO[_memo_key(out)] = (out, t)  # and back
    return out


def memo_call(tag, tensors, scalars, fn):
    """``fn()`` -- a neighbour search that depends on coordinates only -- remembered while a ``coord_scope`` is open, keyed
    by the operation, its scalar arguments and the identity of its input tensors.  Lets a caller run the SAME search
    early on another stream (backbones.PointNet2Msg_fast prefetches SA2's sampling / ball query and the FP layers' three-NN
    while SA1's MLP runs): the module that needs the result later finds it here and makes its stream wait for the event
    recorded behind the search."""
    if _MEMO is None or any(t.requires_grad for t in tensors):
        return fn()
    key = (tag, scalars) + tuple(_memo_key(t) for t in tensors)
    e = _MEMO.get(key)
    if e is None:
        out = fn()
        ev = st = None
        if tensors[0].is_cuda:
            st = torch.cuda.current_stream(tensors[0].device)
            ev = torch.cuda.Event()
            ev.record(st)
        _MEMO[key] = (tensors, out, ev, st)  # the inputs stay alive with the entry: their addresses cannot be reused
        return out
    _, out, ev, st = e
    if ev is not None:
        cur = torch.cuda.current_stream(tensors[0].device)
        if cur != st:
            cur.wait_event(ev)
            for t in (out if isinstance(out, (tuple, list)) else (out,)):
                t.record_stream(cur)
    return out


# ------------------------------------------------------------------ building blocks -----------
def _make_stack(in_channel, widths, conv_cls, bn_cls):
    convs, bns = nn.ModuleList(), nn.ModuleList()
    last = in_channel
    for w in widths:
        convs.append(conv_cls(last, w, 1))
        bns.append(bn_cls(w))
        last = w
    return convs, bns, last


def _run_stack(x, convs, bns):
    for conv, bn in zip(convs, bns):
        x = F.relu(bn(conv(x)))
    return x


def _neighbour_idx(use_knn, radius, K, xyz_t, new_xyz_t):
    """int32 (B,S,K) group indices: kNN or ball query on (B,N,3) / (B,S,3) coordinates."""
    if use_knn:
        return futils.knn(K, new_xyz_t, xyz_t)[1]
    return futils.ball_query(radius, K, xyz_t, new_xyz_t)


def _as_long(idx32):
    """int64 copy of an int32 index tensor (the type the reference API returns, pointnet_utils.py:12-24,140-154) that
    remembers its int32 source: the kernels want int32, and q1's indices go back in twice per scale (q1, then q2)."""
    out = idx32.long()
    out._pn2_i32 = (idx32, out._version)
    return out


def _as_int(idx):
    twin = getattr(idx, "_pn2_i32", None)
    if twin is not None and twin[1] == idx._version and twin[0].shape == idx.shape:
        return twin[0]
    return idx.int()


class _EngineMixin:
    def _fused(self):
        return getattr(self, "engine", _ENGINE) == "fused"


class _MsgBase(nn.Module, _EngineMixin):
    """Shared constructor of the three multi-scale-grouping SA classes."""

    def _build(self, radius_list, nsample_list, in_channel, mlp_list, knn):
        self.radius_list = radius_list
        self.nsample_list = nsample_list
        self.conv_blocks = nn.ModuleList()
        self.bn_blocks = nn.ModuleList()
        self.out_channel = 0
        for widths in mlp_list:
            convs, bns, last = _make_stack(in_channel, widths, nn.Conv2d, nn.BatchNorm2d)
            self.conv_blocks.append(convs)
            self.bn_blocks.append(bns)
            self.out_channel += last
        self.knn = knn
        self.engine = _ENGINE

    def _scale(self, i, xyz, points, new_xyz, idx, centre_feat=None):
        """One scale: group -> [points, xyz - centre(, centre_feat)] -> MLP -> max over K.
        xyz (B,3,N), points (B,D,N)|None, new_xyz (B,3,S), idx (B,S,K) int32 -> (B,D',S)."""
        if self._fused():
            from . import fused
            return fused.sa_scale(xyz, points, new_xyz, idx, centre_feat, self.conv_blocks[i], self.bn_blocks[i],
                                  self.training, precise=getattr(self, "precise_layers", 0))
        grouped = futils.grouping_operation(xyz.contiguous(), idx) - new_xyz.unsqueeze(-1)
        if points is not None:
            grouped = torch.cat([futils.grouping_operation(points.contiguous(), idx), grouped], dim=1)
        if centre_feat is not None:
            grouped = torch.cat([grouped, centre_feat.unsqueeze(-1).expand(-1, -1, -1, grouped.shape[-1])], dim=1)
        return _run_stack(grouped, self.conv_blocks[i], self.bn_blocks[i]).max(dim=-1)[0]


class PointNetSetAbstractionMsg(_MsgBase):
    """Reference pointnet_utils.py:189-250.  xyz (B,3,N), points (B,D,N)|None -> (new_xyz (B,3,S), (B,D',S))."""

    def __init__(self, npoint, radius_list, nsample_list, in_channel, mlp_list, knn=False):
        super().__init__()
        self.npoint = npoint
        self._build(radius_list, nsample_list, in_channel, mlp_list, knn)

    def forward(self, xyz, points):
        xyz_t = t_contig(xyz)
        fps_idx = futils.furthest_point_sample(xyz_t, self.npoint)
        new_xyz = futils.gather_operation(xyz.contiguous(), fps_idx)
        new_xyz_t = t_contig(new_xyz)
        outs = []
        for i, radius in enumerate(self.radius_list):
            idx = _neighbour_idx(self.knn, radius, self.nsample_list[i], xyz_t, new_xyz_t)
            outs.append(self._scale(i, xyz, points, new_xyz, idx))
        return new_xyz, torch.cat(outs, dim=1)


class PointNetSetAbstractionMsg_fast(_MsgBase):
    """Reference pointnet_utils.py:346-409: as above with a part dimension P.
    xyz (B,P,3,N), points (B,P,D,N) (D may be 0) -> (new_xyz (B,P,3,S), (B,P,D',S)).
    FPS and the neighbour search run on part 0 and are shared by all parts (:380-392)."""

    def __init__(self, npoint, radius_list, nsample_list, in_channel, mlp_list, knn=False):
        super().__init__()
        self.npoint = npoint
        self._build(radius_list, nsample_list, in_channel, mlp_list, knn)

    def search(self, xyz):
        """Sampling and grouping of ``forward`` -- everything that depends on the coordinates alone: xyz (B,P,3,N) ->
        (xyz0 (B,3,N), new_xyz (B,3,S), [idx (B,S,K) int32 per scale]).  Inside a ``coord_scope`` the results are
        remembered (``memo_call``), so a caller may run it ahead of time, on another stream."""
        S = self.npoint
        xyz0 = xyz[:, 0].contiguous()
        xyz_t = t_contig(xyz0)
        fps_idx = memo_call("fps", (xyz_t,), (S,), lambda: futils.furthest_point_sample(xyz_t, S))
        new_xyz = memo_call("gather", (xyz0, fps_idx), (), lambda: futils.gather_operation(xyz0, fps_idx))
        new_xyz_t = t_contig(new_xyz)
        idxs = [memo_call("group", (xyz_t, new_xyz_t), (bool(self.knn), float(radius), int(k)),
                          lambda radius=radius, k=k: _neighbour_idx(self.knn, radius, k, xyz_t, new_xyz_t))
                for radius, k in zip(self.radius_list, self.nsample_list)]
        return xyz0, new_xyz, idxs

    def forward(self, xyz, points):
        B, P, C, N = xyz.shape
        S = self.npoint
        xyz0, new_xyz, idxs = self.search(xyz)
        feats = None
        if points is not None and points.shape[-2] > 0:
            feats = _carry(points, points.reshape(B * P, -1, N))
        rep = (lambda t: t) if P == 1 else (lambda t: t.repeat_interleave(P, dim=0))
        outs = []
        for i, radius in enumerate(self.radius_list):
            idx = idxs[i]
            outs.append(self._scale(i, rep(xyz0), feats, rep(new_xyz), rep(idx)))
        out = outs[0] if len(outs) == 1 else torch.cat(outs, dim=1)
        return new_xyz.unsqueeze(1).expand(B, P, C, S), _carry(out, out.reshape(B, P, -1, S))


class PointNetSetAbstractionMsg_GivenCenterPoints(_MsgBase):
    """Reference pointnet_utils.py:515-590: SA-MSG around GIVEN centres (no FPS), optionally
    concatenating a per-centre feature broadcast over the K neighbours (:574-575), and able to
    return / re-use the group indices (HandTrackNet q1 -> q2, hand_network.py:132-134)."""

    def __init__(self, radius_list, nsample_list, mlp_list, in_channel, knn=False):
        super().__init__()
        self._build(radius_list, nsample_list, in_channel, mlp_list, knn)

    def group_indices(self, xyz, new_xyz):
        """The per-scale group indices (int64 (B,S,K) each) ``forward`` would compute: they depend on the coordinates
        only, so a caller may ask for them ahead of time (e.g. on a side stream, while the backbone runs) and pass
        them back as ``pre_group_idx``."""
        xyz_t = t_contig(xyz)
        new_xyz_t = t_contig(new_xyz)
        if self.knn:
            # the k nearest come back ascending with a stable tie rule (interpolate_gpu.cu:30-56), so the K nearest
            # are the first K columns of the max(K) nearest: one search serves every scale
            knn_all = _neighbour_idx(True, self.radius_list[0], max(self.nsample_list), xyz_t, new_xyz_t)
            return [_as_long(knn_all[..., :k].contiguous()) for k in self.nsample_list]
        return [_as_long(_neighbour_idx(False, r, k, xyz_t, new_xyz_t)) for r, k in zip(self.radius_list, self.nsample_list)]

    def forward(self, xyz, points, new_xyz, new_points, return_4nn=False, pre_group_idx=None,
                return_group_idx=False):
        outs, idx_list = [], []
        idx = None
        if pre_group_idx is None:
            pre_group_idx = self.group_indices(xyz, new_xyz)
        for i, radius in enumerate(self.radius_list):
            idx = pre_group_idx[i]
            idx_list.append(idx)
            outs.append(self._scale(i, xyz, points, new_xyz, _as_int(idx), new_points))
        out = torch.cat(outs, dim=1)
        if return_4nn:
            rel = futils.grouping_operation(xyz.contiguous(), idx[..., :4].int().contiguous()) - new_xyz.unsqueeze(-1)
            return out, rel.norm(dim=1, keepdim=True).mean(dim=-1)
        if return_group_idx:
            return out, idx_list
        return out


class _GroupAllBase(nn.Module, _EngineMixin):
    def _build(self, npoint, radius, nsample, in_channel, mlp, group_all, knn):
        self.npoint, self.radius, self.nsample = npoint, radius, nsample
        self.mlp_convs, self.mlp_bns, self.out_channel = _make_stack(in_channel, mlp, nn.Conv2d, nn.BatchNorm2d)
        self.group_all = group_all
        self.knn = knn
        self.engine = _ENGINE

    def _pool_all(self, xyz, points):
        """xyz (B,3,N), points (B,D,N)|None -> (B,D',1): MLP over [xyz, points], max over all N points."""
        assert self.group_all, "Not Implemented"  # as the reference (pointnet_utils.py:502)
        if self._fused():
            from . import fused
            return fused.sa_group_all(xyz, points, self.mlp_convs, self.mlp_bns, self.training,
                                      precise=getattr(self, "precise_layers", 0))
        x = xyz if points is None else torch.cat([xyz, points], dim=1)
        return _run_stack(x.unsqueeze(-1), self.mlp_convs, self.mlp_bns).max(dim=2)[0]


class PointNetSetAbstraction(_GroupAllBase):
    """Reference pointnet_utils.py:298-343 (group_all only).  xyz (B,3,N), points (B,D,N)|None."""

    def __init__(self, npoint, radius, nsample, in_channel, mlp, group_all, knn=False):
        super().__init__()
        self._build(npoint, radius, nsample, in_channel, mlp, group_all, knn)

    def forward(self, xyz, points):
        B, C, _ = xyz.shape
        return torch.zeros(B, C, 1, device=xyz.device, dtype=xyz.dtype), self._pool_all(xyz, points)


class PointNetSetAbstraction_fast(_GroupAllBase):
    """Reference pointnet_utils.py:467-512.  xyz (B,P,3,N), points (B,P,D,N) -> ((B,P,3,1), (B,P,D',1))."""

    def __init__(self, npoint, radius, nsample, in_channel, mlp, group_all, knn=False):
        super().__init__()
        self._build(npoint, radius, nsample, in_channel, mlp, group_all, knn)

    def forward(self, xyz, points):
        B, P, C, N = xyz.shape
        feats = _carry(points, points.reshape(B * P, -1, N)) if points is not None else None
        if feats is not None and feats.shape[1] == 0:
            feats = None
        out = self._pool_all(xyz.reshape(B * P, C, N), feats)
        return torch.zeros(B, P, C, 1, device=xyz.device, dtype=xyz.dtype), _carry(out, out.reshape(B, P, -1, 1))


class _FpBase(nn.Module, _EngineMixin):
    def _build(self, in_channel, mlp):
        self.mlp_convs, self.mlp_bns, self.out_channel = _make_stack(in_channel, mlp, nn.Conv1d, nn.BatchNorm1d)
        self.engine = _ENGINE

    def _propagate(self, xyz1_t, xyz2_t, points1, points2, reps=1):
        """xyz1_t (B,N,3), xyz2_t (B,S,3) coordinates (shared by ``reps`` parts); points1 (B*reps,D1,N)|None,
        points2 (B*reps,D2,S) -> (B*reps,D',N).  Weights: 1/(dist+1e-8) normalised over the three
        neighbours (pointnet_utils.py:446-449) -- the epsilon is added to the DISTANCE."""
        N, S = xyz1_t.shape[1], xyz2_t.shape[1]
        if self._fused():
            from . import fused
            return fused.fp_layer(xyz1_t, xyz2_t, points1, points2, self.mlp_convs, self.mlp_bns, self.training, reps,
                                  rows_only=getattr(self, "_rows_only", False),
                                  precise=getattr(self, "precise_layers", 0))
        if S == 1:
            interpolated = points2.expand(-1, -1, N)
        else:
            u_, k_ = xyz1_t.contiguous(), xyz2_t.contiguous()
            dist, idx = memo_call("three_nn_ops", (u_, k_), (), lambda: futils.three_nn(u_, k_))
            recip = 1.0 / (dist + 1e-8)
            weight = recip / recip.sum(dim=2, keepdim=True)
            if reps > 1:
                weight, idx = weight.repeat_interleave(reps, dim=0), idx.repeat_interleave(reps, dim=0)
            interpolated = futils.three_interpolate(points2, idx, weight)
        x = interpolated if points1 is None else torch.cat([points1, interpolated], dim=-2)
        return _run_stack(x, self.mlp_convs, self.mlp_bns)


class PointNetFeaturePropagation(_FpBase):
    """Reference pointnet_utils.py:253-295.  xyz1 (B,3,N), xyz2 (B,3,S), points1 (B,D1,N)|None, points2 (B,D2,S)."""

    def __init__(self, in_channel, mlp):
        super().__init__()
        self._build(in_channel, mlp)

    def forward(self, xyz1, xyz2, points1, points2):
        return self._propagate(t_contig(xyz1), t_contig(xyz2), points1, points2)


class PointNetFeaturePropagation_fast(_FpBase):
    """Reference pointnet_utils.py:412-464.  xyz1 (B,P,3,N), xyz2 (B,P,3,S), points1 (B,P,D1,N),
    points2 (B,P,D2,S) -> (B,P,D',N); three_nn on part 0 only (:437-451)."""

    def __init__(self, in_channel, mlp):
        super().__init__()
        self._build(in_channel, mlp)

    def forward(self, xyz1, xyz2, points1, points2):
        B, P, _, N = xyz1.shape
        S = xyz2.shape[-1]
        p1 = _carry(points1, points1.reshape(B * P, -1, N)) if points1 is not None else None
        out = self._propagate(t_contig(xyz1[:, 0]), t_contig(xyz2[:, 0]), p1,
                              _carry(points2, points2.reshape(B * P, -1, S)), reps=P)
        return _carry(out, out.reshape(B, P, -1, N))
