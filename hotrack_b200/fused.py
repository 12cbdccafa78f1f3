"""The "fused" engine: the grouped per-point MLP of the SA / FP modules on 16-bit tensor-core kernels
(fp16 forward rows, bf16 gradient rows, fp32 accumulation and BatchNorm arithmetic).

Host side of include/pn2b200_mlp.h.  One ``autograd.Function`` (``_MlpStack``) runs a whole
  build rows (gather / interpolate / concat)  ->  [GEMM + BatchNorm statistics] x L  ->  BN+ReLU(+max-pool)
stack and its backward; ``sa_scale`` / ``sa_group_all`` / ``fp_layer`` / ``dense_stack`` are the
entry points ``pointnet_utils.py`` / ``backbones.py`` call when the engine is "fused".  They take
and return the reference's tensor layouts (channel-major fp32), so the modules stay drop-in; the
16-bit row form of every output is additionally attached to the returned tensor (``_pn2_rows``) and
picked up by the next fused consumer, which then never touches the fp32 copy.

What replaces what (reference network/models/pointnet_utils.py):
  :389-403  group_operation x2, "-= centre", cat, 3 x relu(bn(conv)), max   -> sa_scale
  :484-512  sample_and_group_all, 3 x relu(bn(conv)), max                    -> sa_group_all
  :443-463  three_nn, weights, three_interpolate, cat, L x relu(bn(conv))    -> fp_layer
  backbones.py:131-132  relu(bn1(conv1(x)))                                  -> dense_stack
BatchNorm semantics are nn.BatchNorm's: batch statistics (biased variance) in training with the
running-statistics update (momentum, unbiased variance, num_batches_tracked), running statistics
in eval.  Training-mode conv biases cancel in BatchNorm: they only enter the running mean, and
their gradient is exactly zero.

Kernels behind it: forward and input-gradient GEMMs on tcgen05 / TMEM (csrc/mlp_gemm_tc.cu; BatchNorm
finalisation in the forward GEMM's tail, BatchNorm-backward coefficients folded into the weights for the backward one),
weight gradient on tcgen05 too (csrc/mlp_wgrad_tc.cu; the warp-level mma.sync kernel of csrc/mlp_gemm.cu is the cross-check,
PN2_WGRAD_IMPL=mma), row builders / poolers / scatters in csrc/mlp_rows.cu.  Gradients between fused stacks travel as fp32
rows in the producer's sink (_Sink), the row form of a pooled output is made when a fused consumer first asks (_LazyPooledRows).
Step-level helpers owned by train.TrainStep: WeightPlan (one weight-conversion launch per step) and ZeroArena (the
accumulators of the step, zeroed with the gradients by one memset).  Training leaves one piece of state outside state_dict: ``bn._pn2_center``, the
centring constant of each fused BatchNorm (reset_center_state() forgets it).
"""
import torch
from torch.autograd import Function

from . import _lib
from . import pointnet2_cuda as pc

_BF16 = torch.bfloat16  # gradient rows
_F16 = torch.float16    # forward rows (see csrc/mma_common.cuh)


def _stream(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def _p(t):
    return 0 if t is None else t.data_ptr()


def _padk(c):
    """GEMM reduction-dimension padding: one 32-wide chunk, else whole 64-wide (128-byte) chunks."""
    return 32 if c <= 32 else (c + 63) // 64 * 64


# Row-form gradient hand-over ("sink").  A consumer that GATHERS rows of a fused K=1 stack's output (the q
# modules gathering 21 x 80 neighbours per cloud from the (B,384,N) backbone output) has a sparse gradient
# w.r.t. that output.  Through autograd it must be returned dense and channel-major: a 201 MB memset, an
# uncoalesced atomic scatter and a 600 MB-traffic add kernel per consumer.  With the sink enabled the consumer
# instead accumulates its rows (coalesced atomics) into ONE fp32 row buffer owned by the producer and returns
# None; the producer's backward -- which autograd runs after all of its consumers -- adds that buffer while it
# builds dz.  Opt-in (TrainStep enables it): code that asks autograd for d(consumer)/d(producer output)
# directly (torch.autograd.grad on intermediate tensors) must leave it off.
SPARSE_GRAD_SINK = False


def set_sparse_grad_sink(on):
    global SPARSE_GRAD_SINK
    SPARSE_GRAD_SINK = bool(on)


# Parameter gradients written straight into ``param.grad`` (atomics / "+=" inside the kernels) instead of being returned
# to autograd.  Opt-in, for a training step that owns the whole backward pass and keeps plain fp32 .grad buffers
# (train.TrainStep over flat.FlatParams sets it for the duration of its step): with it on, autograd hooks on the
# parameters (DDP's reducer), torch.autograd.grad and backward(inputs=...) do not see these gradients.
DIRECT_PARAM_GRADS = False


class WeightPlan:
    """fp16 copies of the conv weights of a step, all converted by ONE launch at the top of the step.

    The first step records which (weight, padded width) pairs the stacks ask for (and converts them one by one, as
    the engine does without a plan); from then on ``prepare()`` converts all of them with a single
    pn2_mlp_prep_weights_multi launch instead of one small launch in front of every GEMM.  ``finish()`` invalidates
    the copies (the optimiser is about to change the weights).  Owned by TrainStep."""

    def __init__(self):
        self.entries = {}      # (data_ptr, kp, two) -> (weight view, fp16 buffer [1 or 2 planes][cout][kp])
        self.table = None      # device array of PrepDesc records, rebuilt when an entry is added
        self.ready = False

    def prepare(self):
        self.ready = False
        if not self.entries:
            return
        if self.table is None:
            import struct
            raw = b"".join(struct.pack("<QQiiii", w.data_ptr(), buf.data_ptr(), w.shape[0], w.shape[1], kp, 1 if two else 0)
                           for (_, kp, two), (w, buf) in self.entries.items())
            dev = next(iter(self.entries.values()))[0].device
            self.table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
        _lib.call("pn2_mlp_prep_weights_multi", len(self.entries), self.table.data_ptr(), _stream())
        self.ready = True

    def lookup(self, w, kp, two=False):
        """fp16 [planes][cout][kp] copy of ``w`` for this step (converted now if it was not planned); planes = 2
        (hi, lo) when ``two``."""
        key = (w.data_ptr(), kp, bool(two))
        e = self.entries.get(key)
        if e is not None and self.ready:
            return e[1]
        buf = e[1] if e is not None else _prep_weight(w, kp, two, convert=False)
        _prep_weight(w, kp, two, buf=buf)
        if e is None:
            # a detached alias: holding the autograd-tracked view would keep the parameter's grad accumulator of
            # THAT step alive (with the stream it was created on), which a later CUDA-graph capture trips over
            self.entries[key] = (w.detach(), buf)
            self.table = None
        return buf

    def finish(self):
        self.ready = False


def _prep_weight(w, kp, two, buf=None, convert=True):
    if buf is None:
        buf = torch.empty(2 if two else 1, w.shape[0], kp, dtype=_F16, device=w.device)
    if convert:
        if two:
            _lib.call("pn2_mlp_prep_weights_x2", w.shape[0], w.shape[1], kp, w.data_ptr(), buf[0].data_ptr(),
                      buf[1].data_ptr(), _stream())
        else:
            _lib.call("pn2_mlp_prep_weights", w.shape[0], w.shape[1], kp, w.data_ptr(), buf.data_ptr(), 0, _stream())
    return buf


ACTIVE_PLAN = None  # set by TrainStep around its forward pass

# Two-plane ("x2") forward precision.  A stack run with precise=L keeps its first L layers' operands and outputs as
# (hi, lo) fp16 pairs (include/pn2b200_mlp.h).  backbones.PointNet2Msg_fast asks for it on SA1-SA3 and FP3's first
# layer: measured on BASELINE config 3 (tools/dev/emul_prec.py), plain fp16 rows in SA1 ALONE put 7 % on the last
# module's output (FP3 normalises a broadcast global feature, x40 amplification by the end of the network), with those
# stacks two-plane the whole path is within 4e-3 of strict fp32.  PRECISE_MODE "off" ignores the requests (all fp16,
# the round-1 engine), "all" runs every layer of every stack two-plane (fp32-class forward, for strict comparisons).
PRECISE_MODE = "auto"   # "auto": as the modules ask | "off": never | "all": every layer of every stack

# Max-pool of the SA stacks taken in the last layer's GEMM epilogue (pn2_mlp_gemm_fwd[_bn]_pool + pn2_pool_finalize) instead of
# a pn2_pool_fwd pass over the layer's output.  Validated (tests/test_fused_gpu.py::test_gemm_epilogue_pooling) but OFF by
# default: measured on B200 it LOSES 0.33 ms per step -- these last layers have K = 32..128, their tiles are epilogue-bound
# already, and one redux.sync + ballot per column per 32-row quarter (512 of each per 128x128 tile) more than doubles the
# epilogue's issue time, while the vectorised pool_fwd pass it saves costs 0.17 ms in total.
EPILOGUE_POOL = False


def set_precise(mode):
    """'auto' (default) | 'off' | 'all'; booleans are accepted for on/off."""
    global PRECISE_MODE
    if mode is True:
        mode = "auto"
    elif mode is False:
        mode = "off"
    if mode not in ("auto", "off", "all"):
        raise ValueError("precision mode must be 'auto', 'off' or 'all'")
    PRECISE_MODE = mode


class ZeroArena:
    """One zero-filled fp32 buffer per step for the accumulators of every stack (BatchNorm statistics, channel sums,
    BatchNorm-backward sums, last-CTA counters): ONE memset at the top of the step instead of two small ones per stack.
    Bump allocation, nothing is returned; a stack that finds the arena absent or exhausted falls back to torch.zeros.
    Owned by TrainStep (static addresses: CUDA-graph friendly)."""

    def __init__(self, device=None, floats=1 << 18, buf=None):
        # buf: a buffer somebody else zeroes at the top of every step (the tail of FlatParams' gradient buffer: one
        # memset for gradients and accumulators together)
        self.owned = buf is None
        self.buf = torch.zeros(floats, dtype=torch.float32, device=device) if buf is None else buf
        self.off = 0

    def begin(self):
        if self.owned:
            self.buf.zero_()
        self.off = 0

    def take(self, n):
        n = (n + 3) // 4 * 4  # 16-byte granules
        if self.off + n > self.buf.numel():
            return None
        t = self.buf[self.off:self.off + n]
        self.off += n
        return t


ACTIVE_ARENA = None  # set by TrainStep for the whole step (forward and backward)

# Weight gradients on a side stream.  A layer's weight gradient feeds nothing but the optimiser, while the backward pass
# proper is a CHAIN (coefficients -> input gradient -> next layer's coefficients ...) full of small latency-bound launches
# that leave most SMs idle; with this on (TrainStep sets it for its step; needs DIRECT_PARAM_GRADS) every
# pn2_mlp_gemm_wgrad is forked onto a second stream and joined once, before the optimiser (join_wgrad()).  The operands
# of the in-flight kernels are kept alive until the join.
WGRAD_SIDE_STREAM = False
WGRAD_STREAMS = max(1, int(__import__("os").environ.get("PN2_WGRAD_STREAMS", "1")))  # side streams, used round-robin
_WGRAD_SIDE = {}  # device index -> [streams, keep-alive list, launch counter]


def _wgrad_side(dev):
    e = _WGRAD_SIDE.get(dev.index)
    if e is None:
        e = _WGRAD_SIDE[dev.index] = [[torch.cuda.Stream(device=dev) for _ in range(WGRAD_STREAMS)], [], 0]
    e[2] += 1
    return e[0][e[2] % len(e[0])], e[1]


def join_wgrad():
    """Make the current stream of every device wait for the weight-gradient stream(s) (call once after backward)."""
    for idx, (sides, keep, _) in _WGRAD_SIDE.items():
        if keep:
            for side in sides:
                torch.cuda.current_stream(torch.device("cuda", idx)).wait_stream(side)
            keep.clear()


def _zeros(n, dev):
    if ACTIVE_ARENA is not None and ACTIVE_ARENA.buf.device == dev:
        t = ACTIVE_ARENA.take(n)
        if t is not None:
            return t[:n]
    return torch.zeros(n, dtype=torch.float32, device=dev)


def reset_center_state(module):
    """Forget the centring constants training left on the BatchNorm modules of ``module`` (``_pn2_center``): the next
    training forward re-estimates them from 16 sampled rows, exactly as a first step does."""
    for m in module.modules():
        if hasattr(m, "_pn2_center"):
            del m._pn2_center


class _Sink:
    __slots__ = ("rows", "c", "buf", "rows16", "k1")

    def __init__(self, rows, c, k1=True):
        self.rows, self.c, self.buf = rows, c, None
        self.rows16 = None  # (bf16 [rows][ld] tensor, ld): the input gradient of a dense consumer, left in row form
        self.k1 = k1        # producer is a K=1 stack (pn2_pool_bwd takes rows16 only there)


class Rows:
    """fp16 row matrix [rows][ld] with c valid channels; scale/shift (fp32, length >= c) mean the
    consumer must read relu(y*scale + shift) -- the producer's BatchNorm+ReLU applied on the fly."""

    __slots__ = ("y", "lo", "c", "ld", "scale", "shift", "offset", "sink", "numel", "version")

    def __init__(self, y, c, ld, scale=None, shift=None, offset=None, sink=None, lo=None):
        self.lo = lo  # second fp16 plane of two-plane rows (value = y + lo), same shape; None for plain rows
        # offset ((fp32 sums [c], scale)): the rows are stored CENTRED, true value = y + sums * scale (pooled features)
        self.y, self.c, self.ld, self.scale, self.shift, self.offset = y, c, ld, scale, shift, offset
        self.sink = sink  # _Sink of the producing K=1 stack (row-form gradient hand-over), or None
        self.numel = self.version = None


class _LazyPooledRows:
    """Row form of a pooled stack's output, converted (pn2_to_rows_x2: fp32 (B,C,S) -> centred fp16 rows) when the first
    fused consumer asks for it."""

    __slots__ = ("out", "chan_sums", "inv", "dims", "two", "sink", "numel", "version", "rows")

    def __init__(self, out, chan_sums, inv, B, C, groups, two, sink):
        self.out, self.chan_sums, self.inv, self.dims, self.two, self.sink = out.detach(), chan_sums, inv, (B, C, groups), two, sink
        self.numel = self.version = self.rows = None

    def materialize(self):
        if self.rows is None:
            B, C, groups = self.dims
            buf = torch.empty(2 if self.two else 1, B * groups, C, dtype=_F16, device=self.out.device)
            with torch.cuda.device(self.out.device):
                _lib.call("pn2_to_rows_x2", B, C, groups, self.out.data_ptr(), self.chan_sums.data_ptr(), self.inv,
                          buf[0].data_ptr(), buf[1].data_ptr() if self.two else 0, C, _stream(self.out.device))
            self.rows = Rows(buf[0], C, C, offset=(self.chan_sums, self.inv), lo=buf[1] if self.two else None, sink=self.sink)
            self.rows.numel, self.rows.version = self.numel, self.version
            self.out = None
        return self.rows


def attach_rows(t, rows):
    rows.numel, rows.version = t.numel(), t._version
    t._pn2_rows = rows
    return t


def carry_rows(src, dst):
    """Propagate the attached row form across a reshape/view of the same values."""
    r = getattr(src, "_pn2_rows", None)
    if r is not None and dst.numel() == r.numel:
        dst._pn2_rows = r
    return dst


def rows_of(t, two=False):
    """Row source of a (B,C,N) fp32 tensor: the attached one if still valid, else a conversion (two-plane if asked)."""
    r = getattr(t, "_pn2_rows", None)
    if r is not None and r.numel == t.numel() and r.version == t._version:
        return r.materialize() if isinstance(r, _LazyPooledRows) else r
    B, C, N = t.shape
    t = t.contiguous()
    if t.dtype != torch.float32:
        raise TypeError("fused engine expects fp32 feature tensors")
    ld = (C + 7) // 8 * 8
    y = torch.empty(2 if two else 1, B * N, ld, dtype=_F16, device=t.device)
    _lib.call("pn2_to_rows_x2", B, C, N, t.data_ptr(), 0, 0.0, y[0].data_ptr(), y[1].data_ptr() if two else 0, ld, _stream())
    return Rows(y[0], C, ld, lo=y[1] if two else None)


class _Layer:
    __slots__ = ("w", "cin", "kp", "cout", "y", "scale", "shift", "mean", "rstd")  # w, y: hi planes (what backward reads)


def _bn_momentum(bn):
    """nn.BatchNorm's update factor for THIS step: ``momentum``, or 1 / (batches seen so far, this one included) when
    momentum is None (cumulative moving average).  The latter reads num_batches_tracked on the host: such a module
    cannot be captured in a CUDA graph (the reference never builds one: trainer.py:167-190 sets a float momentum)."""
    if bn.momentum is not None:
        return float(bn.momentum)
    seen = int(bn.num_batches_tracked.item()) if bn.num_batches_tracked is not None else 0
    return 1.0 / float(seen + 1)


class _MlpStack(Function):
    """kind 'sa'    : meta = (xyz (B,3,N), new_xyz (B,3,S)|None, idx (B,S,K) int32|None, xyz_first)
                      a = features (B,D,N)|None, b = centre features (B,E,S)|None   -> (B,Cout,S)
       kind 'fp'    : meta = (idx (B,N,3) int32|None, dist2 (B,N,3)|None, N, S)
                      a = skip (B,D1,N)|None, b = coarse (B,D2,S)                    -> (B,Cout,N)
       kind 'dense' : a = x (B,C,N)                                                  -> (B,Cout,N)"""

    @staticmethod
    def forward(ctx, kind, meta, bns, pobjs, training, precise, a, b, *params):
        dev = params[0].device
        st = _stream(dev)
        nl = len(params) // 4
        # leading layers with two-plane operands
        precise = nl if PRECISE_MODE == "all" else (min(int(precise), nl) if PRECISE_MODE == "auto" else 0)
        two0 = precise > 0
        ra = rows_of(a, two0) if a is not None else None
        rb = rows_of(b, two0) if b is not None else None

        # ---- layer-0 input rows
        if kind == "sa":
            xyz, new_xyz, idx, xyz_first = meta
            B, _, N = xyz.shape
            if idx is not None:
                S, K = idx.shape[1], idx.shape[2]
            else:
                S, K = 1, N
            fc = ra.c if ra is not None else 0
            cc = rb.c if rb is not None else 0
            cin = fc + 3 + cc
            R, groups, pool_k = B * S * K, S, K
            x0p = torch.empty(2 if two0 else 1, R, _padk(cin), dtype=_F16, device=dev)
            x0, x0_lo = x0p[0], (x0p[1] if two0 else None)
            _lib.call("pn2_sa_build_rows_x2", B, N, S, K, xyz.data_ptr(), _p(new_xyz), _p(idx),
                      _p(ra.y) if ra else 0, _p(ra.lo) if ra else 0, fc, ra.ld if ra else 0, _p(ra.scale) if ra else 0,
                      _p(ra.shift) if ra else 0,
                      _p(rb.y) if rb else 0, _p(rb.lo) if rb else 0, cc, rb.ld if rb else 0, _p(rb.scale) if rb else 0,
                      _p(rb.shift) if rb else 0,
                      1 if xyz_first else 0, x0.data_ptr(), _p(x0_lo), x0.shape[1], st)
        elif kind == "fp":
            idx, dist2, N, S = meta[:4]
            B = b.shape[0]
            sc = ra.c if ra is not None else 0
            cin = sc + rb.c
            R, groups, pool_k = B * N, N, 1
            x0p = torch.empty(2 if two0 else 1, R, _padk(cin), dtype=_F16, device=dev)
            x0, x0_lo = x0p[0], (x0p[1] if two0 else None)
            _lib.call("pn2_fp_build_rows_x2", B, N, S, _p(ra.y) if ra else 0, _p(ra.lo) if ra else 0, sc, ra.ld if ra else 0,
                      _p(ra.scale) if ra else 0, _p(ra.shift) if ra else 0, rb.y.data_ptr(), _p(rb.lo), rb.c, rb.ld,
                      _p(rb.scale), _p(rb.shift), _p(idx), _p(dist2), x0.data_ptr(), _p(x0_lo), x0.shape[1], st)
        else:
            B, cin, N = a.shape
            R, groups, pool_k = B * N, N, 1
            x0 = x0_lo = None  # the dense stack reads its input rows in place

        # ---- L x (GEMM + batch statistics)
        layers = []
        if x0 is not None:
            x, x_lo, x_ld, xs, xh, kp = x0, x0_lo, x0.shape[1], None, None, x0.shape[1]
        else:
            if ra.ld == _padk(ra.c):
                x, x_lo, x_ld, xs, xh, kp = ra.y, ra.lo, ra.ld, ra.scale, ra.shift, ra.ld
            else:  # re-pad the row form to the GEMM's chunk granularity
                kp = _padk(ra.c)
                x = torch.zeros(R, kp, dtype=_F16, device=dev)
                x[:, :ra.c] = ra.y[:, :ra.c]
                x_lo = None
                if ra.lo is not None:
                    x_lo = torch.zeros(R, kp, dtype=_F16, device=dev)
                    x_lo[:, :ra.c] = ra.lo[:, :ra.c]
                x_ld, xs, xh = kp, ra.scale, ra.shift
            if xs is not None and xs.numel() < kp:
                xs = torch.cat([xs, xs.new_zeros(kp - xs.numel())])
                xh = torch.cat([xh, xh.new_zeros(kp - xh.numel())])
        in0 = (x, x_ld, xs, xh)
        # per-channel constants the layer-0 input rows were centred by (pooled features): up to two column segments, each
        # (sums, scale, first column, channels) -- handed to pn2_mlp_center as they are (no assembly kernels)
        segs = []
        if kind == "sa":
            segs = [(ra, 3 if meta[3] else 0), (rb, (ra.c if ra is not None else 0) + 3)]
        elif kind == "fp":
            segs = [(ra, 0), (rb, ra.c if ra is not None else 0)]
        elif ra.offset is not None:
            segs = [(ra, 0)]  # centred pooled rows fed straight to a dense stack
        in_off = [(r.offset[0], float(r.offset[1]), start, r.c) for r, start in segs if r is not None and r.offset is not None]
        in_off = in_off or None
        # one zero-filled arena for every accumulator of the forward pass (one memset per stack)
        widths = [params[4 * l].shape[0] for l in range(nl)]
        arena = _zeros(2 * sum(widths) + widths[-1] + nl, dev)
        counters = arena[2 * sum(widths) + widths[-1]:]  # one zeroed word per layer (GEMM tail: BatchNorm finalisation)
        a_off = 0
        # last layer of a pooled stack: the max over the pool_k rows of a group is taken in the GEMM's epilogue
        epi_pool = EPILOGUE_POOL and pool_k in (16, 32, 64, 128) and R % pool_k == 0
        pool_val = pool_arg = None
        for l in range(nl):
            w, bias, gamma, beta = params[4 * l: 4 * l + 4]
            bn = bns[l]
            L = _Layer()
            L.cout, L.cin, L.kp = w.shape[0], w.shape[1], kp
            if L.cout % 32:
                raise ValueError("fused engine needs layer widths that are multiples of 32 (got %d)" % L.cout)
            # two-plane layer: BatchNorm'd two-plane inputs are limited to 256 columns by the kernel (wider: fp16 planes)
            two = l < precise and not (xs is not None and kp > 256)
            if two and x_lo is None:  # a one-plane producer in front of a two-plane layer
                x_lo = torch.zeros_like(x)
            wbuf = ACTIVE_PLAN.lookup(w, kp, two) if (ACTIVE_PLAN is not None and training) else _prep_weight(w, kp, two)
            L.w, w_lo = wbuf[0], (wbuf[1] if two else None)
            pooled_here = epi_pool and l == nl - 1
            # (the pooled layer's lo plane would only feed the pooling: the epilogue pools the fp32 accumulators instead)
            keep_lo = two and not pooled_here
            yp = torch.empty(2 if keep_lo else 1, R, L.cout, dtype=_F16, device=dev)
            L.y, y_lo = yp[0], (yp[1] if keep_lo else None)
            if pooled_here:
                pool_val = torch.empty(R // pool_k, L.cout, dtype=torch.float32, device=dev)
                pool_arg = torch.empty(B, groups, L.cout, dtype=torch.int32, device=dev)
            consts = torch.empty(6, L.cout, dtype=torch.float32, device=dev)
            L.scale, L.shift, L.mean, L.rstd, cen, cen_true = (consts[i] for i in range(6))
            # Centring constant.  Training keeps it as state on the BatchNorm module: the GEMM tail of step t leaves the
            # batch mean of the un-centred output there, and step t+1 centres by it (BatchNorm is shift-invariant, any
            # constant near the mean serves) -- the 16-row estimate (pn2_mlp_center) then only runs on the first step, in
            # eval mode, and for layers whose input rows carry a per-channel offset (its W.offset term is needed too).
            stateful = training and not (l == 0 and in_off is not None)
            state = getattr(bn, "_pn2_center", None) if stateful else None
            if state is not None and (state.device != dev or state.numel() != L.cout):
                state = None
            next_cen = None
            if state is not None:
                cen = cen_true = next_cen = state
            else:
                o0 = in_off[0] if (l == 0 and in_off) else (None, 0.0, 0, 0)
                o1 = in_off[1] if (l == 0 and in_off and len(in_off) > 1) else (None, 0.0, 0, 0)
                _lib.call("pn2_mlp_center", R, kp, L.cout, x.data_ptr(), x_ld, _p(xs), _p(xh), L.w.data_ptr(),
                          _p(o0[0]), o0[1], o0[2], o0[3], _p(o1[0]), o1[1], o1[2], o1[3], cen.data_ptr(), cen_true.data_ptr(), st)
                if stateful:
                    next_cen = bn._pn2_center = torch.empty(L.cout, dtype=torch.float32, device=dev)
            if training:
                stats = arena[a_off:a_off + 2 * L.cout]
                a_off += 2 * L.cout
                track = bn.track_running_stats and bn.running_mean is not None
                args = (R, kp, L.cout, x.data_ptr(), _p(x_lo) if two else 0, x_ld, _p(xs), _p(xh),
                        L.w.data_ptr(), _p(w_lo), cen.data_ptr(), L.y.data_ptr(), _p(y_lo), L.cout, stats.data_ptr(),
                        counters[l].data_ptr(),
                        bn.weight.data_ptr(), bn.bias.data_ptr(), _p(bias), cen_true.data_ptr(), _bn_momentum(bn),
                        float(bn.eps), _p(bn.running_mean) if track else 0, _p(bn.running_var) if track else 0,
                        _p(bn.num_batches_tracked) if track else 0, L.scale.data_ptr(), L.shift.data_ptr(),
                        L.mean.data_ptr(), L.rstd.data_ptr(), _p(next_cen))
                if pooled_here:
                    _lib.call("pn2_mlp_gemm_fwd_bn_pool", *args, pool_k, pool_val.data_ptr(), pool_arg.data_ptr(), st)
                else:
                    _lib.call("pn2_mlp_gemm_fwd_bn_x2", *args, st)
            else:
                _lib.call("pn2_bn_eval_affine", L.cout, bn.weight.data_ptr(), bn.bias.data_ptr(), _p(bias),
                          cen_true.data_ptr(), bn.running_mean.data_ptr(), bn.running_var.data_ptr(), float(bn.eps),
                          L.scale.data_ptr(), L.shift.data_ptr(), st)
                args = (R, kp, L.cout, x.data_ptr(), _p(x_lo) if two else 0, x_ld, _p(xs), _p(xh),
                        L.w.data_ptr(), _p(w_lo), cen.data_ptr(), L.y.data_ptr(), _p(y_lo), L.cout, 0)
                if pooled_here:
                    _lib.call("pn2_mlp_gemm_fwd_pool", *args, pool_k, bn.weight.data_ptr(), pool_val.data_ptr(),
                              pool_arg.data_ptr(), st)
                else:
                    _lib.call("pn2_mlp_gemm_fwd_x2", *args, st)
            layers.append(L)
            x, x_lo, x_ld, xs, xh, kp = L.y, y_lo, L.cout, L.scale, L.shift, L.cout
            last_two = two

        # ---- BN + ReLU (+ max over the group) -> module output
        last = layers[-1]
        C = last.cout
        out = torch.empty(B, C, groups, dtype=torch.float32, device=dev)
        rows_only = pool_k == 1 and ((kind == "fp" and len(meta) > 4 and meta[4]) or (kind == "dense" and bool(meta and meta[0])))
        chan_sums = argmax = None
        if pool_k > 1:
            chan_sums = arena[2 * sum(widths):2 * sum(widths) + widths[-1]]
            argmax = pool_arg if pool_arg is not None else (torch.empty(B, groups, C, dtype=torch.int32, device=dev)
                                                            if training else None)
        if pool_val is not None:
            _lib.call("pn2_pool_finalize", B, groups, pool_k, C, pool_val.data_ptr(), pool_arg.data_ptr(), last.y.data_ptr(), C,
                      last.scale.data_ptr(), last.shift.data_ptr(), out.data_ptr(), _p(chan_sums), st)
        elif not rows_only:  # rows_only: the caller promises that only the attached row form is read (backbone FP1 -> head)
            _lib.call("pn2_pool_fwd_x2", B, groups, pool_k, C, last.y.data_ptr(), _p(x_lo), C, last.scale.data_ptr(),
                      last.shift.data_ptr(), out.data_ptr(), _p(chan_sums), _p(argmax), st)
        if pool_k > 1:
            # pooled features go to the next fused consumer as bf16 rows centred on their channel mean
            inv = 1.0 / (B * groups)
            two_out = last_two  # the last layer was two-plane: so are the pooled rows
            # pooled outputs offer a row-form gradient sink too: SA2 / SA3 gathering from SA1 / SA2's output and the FP
            # layers' skip connections accumulate coalesced fp32 rows instead of strided channel-major atomics
            ctx.out_sink = _Sink(B * groups, C, k1=False) if training else None
            # the conversion itself waits for the first fused consumer (rows_of): the outputs of q1 / q2's scales are
            # concatenated by torch and never read in this form
            _MlpStack.last_rows = _LazyPooledRows(out, chan_sums, inv, B, C, groups, two_out, ctx.out_sink)
        else:
            ctx.out_sink = _Sink(R, C) if (training and C <= 1024) else None
            _MlpStack.last_rows = Rows(last.y, C, C, last.scale, last.shift, sink=ctx.out_sink, lo=x_lo)
        # gather-type consumer of a producer that offers a sink: deliver the feature gradient in row form
        ctx.feat_sink = ra.sink if (kind in ("sa", "dense") and SPARSE_GRAD_SINK and training and ra is not None
                                    and ra.sink is not None and a is not None and a.requires_grad) else None
        # FP layer on top of a fused K=1 producer (FP2 -> FP1, FP3 -> FP2): the coarse features' gradient is scattered
        # straight into the producer's row-form sink instead of a zero-filled buffer that is then transposed to (B,C,S)
        # for autograd and transposed back to rows by the producer's backward
        ctx.skip_sink = ra.sink if (kind == "fp" and SPARSE_GRAD_SINK and training and ra is not None and ra.sink is not None
                                    and a is not None and ra.sink.rows == a.shape[0] * a.shape[2] and ra.sink.c == ra.c
                                    and a.requires_grad) else None
        ctx.coarse_sink = rb.sink if (kind == "fp" and SPARSE_GRAD_SINK and training and rb is not None and rb.sink is not None
                                      and rb.sink.rows == b.shape[0] * b.shape[2] and b.requires_grad) else None
        del x_lo, x0_lo  # lo planes are forward-only: backward reads the hi planes

        ctx.kind, ctx.meta, ctx.training = kind, meta, training
        ctx.dims = (B, groups, pool_k, R, cin)
        ctx.layers, ctx.in0, ctx.argmax = layers, in0, argmax
        ctx.a_shape = None if a is None else tuple(a.shape)
        ctx.b_shape = None if b is None else tuple(b.shape)
        ctx.gammas = [params[4 * l + 2] for l in range(nl)]
        ctx.pobjs = pobjs
        ctx.set_materialize_grads(False)
        return out

    @staticmethod
    def backward(ctx, dout):
        nl = len(ctx.layers)
        none_params = [None] * (4 * nl)
        extra = ctx.out_sink.buf if ctx.out_sink is not None else None
        extra16 = ctx.out_sink.rows16 if ctx.out_sink is not None else None
        if ctx.out_sink is not None:
            ctx.out_sink.buf = ctx.out_sink.rows16 = None
        if dout is None and extra is None and extra16 is None:
            return (None, None, None, None, None, None, None, None, *none_params)
        if not ctx.training:
            raise NotImplementedError("fused engine: backward through eval-mode BatchNorm is not implemented; "
                                      "use engine 'ops' for that")
        dev = ctx.layers[0].y.device
        st = _stream(dev)
        B, groups, pool_k, R, cin = ctx.dims
        layers = ctx.layers
        if dout is not None:
            dout = dout.contiguous().float()
        last = layers[-1]
        C = last.cout
        dz = torch.empty(R, C, dtype=_BF16, device=dev)
        # one zero-filled arena for the BatchNorm-backward sums of every layer
        arena = _zeros(2 * sum(L.cout for L in layers), dev)
        a_off = 2 * C
        sums = arena[:a_off]
        _lib.call("pn2_pool_bwd", B, groups, pool_k, C, _p(dout), _p(extra), _p(extra16[0]) if extra16 else 0,
                  extra16[1] if extra16 else 0, last.y.data_ptr(), C, last.scale.data_ptr(),
                  last.shift.data_ptr(), last.mean.data_ptr(), last.rstd.data_ptr(), _p(ctx.argmax), dz.data_ptr(), C,
                  sums.data_ptr(), st)
        need_a = ctx.needs_input_grad[6] and ctx.a_shape is not None
        need_b = ctx.needs_input_grad[7] and ctx.b_shape is not None
        grads = [None] * (4 * nl)
        dx0 = None
        for l in range(nl - 1, -1, -1):
            L = layers[l]
            # Parameter gradients go STRAIGHT into the parameters' .grad when the step opted in (DIRECT_PARAM_GRADS) and
            # those exist as plain fp32 buffers (FlatParams keeps them so): autograd's per-parameter "+=" kernels
            # disappear.  Otherwise they are returned to autograd the usual way.
            pw, pb, pg, pbt = ctx.pobjs[4 * l: 4 * l + 4]
            direct = DIRECT_PARAM_GRADS and all(q.grad is not None and q.grad.is_contiguous() and q.grad.dtype == torch.float32
                                                for q in (pw, pg, pbt))
            coefs = torch.empty(5, L.cout, dtype=torch.float32, device=dev)
            # the input-gradient GEMM multiplies dz and y as stored: the BatchNorm-backward coefficients are folded into
            # two copies of the weights (and a bias) by the same small kernel that derives them
            want_dx = l > 0 or need_a or need_b
            wa = wb = negbias = unscale = None
            if want_dx:
                wa = torch.empty(L.kp, L.cout, dtype=_BF16, device=dev)
                wb = torch.empty(L.kp, L.cout, dtype=_F16, device=dev)
                negbias = torch.empty(L.kp + 1, dtype=torch.float32, device=dev)
                unscale = negbias[L.kp:]
            _lib.call("pn2_bn_bwd_coefs", L.cout, R, sums.data_ptr(), ctx.gammas[l].data_ptr(), L.mean.data_ptr(),
                      L.rstd.data_ptr(), coefs[0].data_ptr(), coefs[1].data_ptr(), coefs[2].data_ptr(),
                      pg.grad.data_ptr() if direct else coefs[3].data_ptr(),
                      pbt.grad.data_ptr() if direct else coefs[4].data_ptr(), 1 if direct else 0,
                      pw.data_ptr() if want_dx else 0, L.cin, L.kp, _p(wa), _p(wb), _p(negbias), _p(unscale), st)
            if l > 0:
                P = layers[l - 1]
                x, x_ld, xs, xh = P.y, P.cout, P.scale, P.shift
            else:
                x, x_ld, xs, xh = ctx.in0
            dw = None if direct else torch.zeros(L.cout, L.cin, dtype=torch.float32, device=dev)
            wargs = (R, L.cout, L.kp, L.cin, dz.data_ptr(), L.cout, L.y.data_ptr(), L.cout,
                     coefs[0].data_ptr(), coefs[1].data_ptr(), coefs[2].data_ptr(), x.data_ptr(), x_ld, _p(xs), _p(xh),
                     pw.grad.data_ptr() if direct else dw.data_ptr(), L.cin)
            if direct and WGRAD_SIDE_STREAM:
                side, keep = _wgrad_side(dev)
                side.wait_stream(torch.cuda.current_stream(dev))  # dz, the coefficients and everything before them
                with torch.cuda.stream(side):
                    _lib.call("pn2_mlp_gemm_wgrad", *wargs, side.cuda_stream)
                keep.extend((dz, L.y, x, xs, xh, coefs))
            else:
                _lib.call("pn2_mlp_gemm_wgrad", *wargs, st)
            if not direct:
                grads[4 * l] = dw
                grads[4 * l + 2] = coefs[3]
                grads[4 * l + 3] = coefs[4]
            if not direct or pb.grad is None:
                # conv bias: cancelled by train-mode BatchNorm, gradient exactly zero
                grads[4 * l + 1] = torch.zeros(L.cout, dtype=torch.float32, device=dev)
            if l > 0:
                P = layers[l - 1]
                dzp = torch.empty(R, P.cout, dtype=_BF16, device=dev)
                sums_p = arena[a_off:a_off + 2 * P.cout]
                a_off += 2 * P.cout
                _lib.call("pn2_mlp_gemm_dgrad", R, L.cout, P.cout, dz.data_ptr(), L.cout, L.y.data_ptr(), L.cout,
                          wa.data_ptr(), wb.data_ptr(), negbias.data_ptr(), unscale.data_ptr(),
                          P.y.data_ptr(), P.cout, P.scale.data_ptr(), P.shift.data_ptr(), P.mean.data_ptr(),
                          P.rstd.data_ptr(), dzp.data_ptr(), P.cout, sums_p.data_ptr(), st)
                dz, sums = dzp, sums_p
            elif need_a or need_b:
                # gradient w.r.t. the VALUES the stack consumed (post-activation when the input rows were
                # still pre-BatchNorm: the producer's own backward applies its ReLU mask)
                dx0 = torch.empty(R, L.kp, dtype=_BF16, device=dev)
                _lib.call("pn2_mlp_gemm_dgrad", R, L.cout, L.kp, dz.data_ptr(), L.cout, L.y.data_ptr(), L.cout,
                          wa.data_ptr(), wb.data_ptr(), negbias.data_ptr(), unscale.data_ptr(), 0, 0, 0, 0, 0,
                          0, dx0.data_ptr(), L.kp, 0, st)
        # weight grads come back in the conv weight's own shape
        da = db = None
        if dx0 is not None:
            if ctx.kind == "sa":
                xyz, new_xyz, idx, xyz_first = ctx.meta
                N = xyz.shape[2]
                S, K = (idx.shape[1], idx.shape[2]) if idx is not None else (1, N)
                fc = ctx.a_shape[1] if ctx.a_shape is not None else 0
                cc = ctx.b_shape[1] if ctx.b_shape is not None else 0
                sink = ctx.feat_sink if need_a else None
                if sink is not None:
                    if sink.buf is None:
                        sink.buf = torch.zeros(sink.rows, sink.c, dtype=torch.float32, device=dev)
                    dfeat, rows_major = sink.buf, 1  # da stays None: the producer's backward picks the buffer up
                elif need_a:
                    da = torch.zeros(ctx.a_shape, dtype=torch.float32, device=dev)
                    dfeat, rows_major = da, 0
                else:
                    dfeat, rows_major = None, 0
                if need_b:
                    db = torch.zeros(ctx.b_shape, dtype=torch.float32, device=dev)
                _lib.call("pn2_sa_rows_bwd", B, N, S, K, _p(idx), dx0.data_ptr(), dx0.shape[1], fc, _p(dfeat), rows_major,
                          cc, _p(db), 1 if xyz_first else 0, st)
            elif ctx.kind == "fp":
                idx, dist2, N, S = ctx.meta[:4]
                sc = ctx.a_shape[1] if ctx.a_shape is not None else 0
                c2 = ctx.b_shape[1]
                dskip, skip_rows = None, 0
                if need_a and ctx.skip_sink is not None:
                    if ctx.skip_sink.buf is None:
                        ctx.skip_sink.buf = torch.zeros(ctx.skip_sink.rows, ctx.skip_sink.c, dtype=torch.float32, device=dev)
                    dskip, skip_rows = ctx.skip_sink.buf, 1  # da stays None (see the sink hand-over above)
                elif need_a:
                    dskip = da = torch.empty(ctx.a_shape, dtype=torch.float32, device=dev)
                sink = ctx.coarse_sink if need_b else None
                if sink is not None:
                    if sink.buf is None:
                        sink.buf = torch.zeros(sink.rows, sink.c, dtype=torch.float32, device=dev)
                    dcr = sink.buf  # db stays None: the producer's backward adds the buffer while it builds its dz
                else:
                    dcr = torch.zeros(B * S, c2, dtype=torch.float32, device=dev) if need_b else None
                _lib.call("pn2_fp_rows_bwd", B, N, S, _p(idx), _p(dist2), dx0.data_ptr(), dx0.shape[1], sc, _p(dskip), skip_rows,
                          c2, _p(dcr), st)
                if need_b and sink is None:
                    db = dcr.view(B, S, c2).transpose(1, 2).contiguous()
            elif ctx.feat_sink is not None and ctx.feat_sink.k1 and need_a:
                # the producer's backward reads it as rows; da stays None (autograd still runs the producer's node,
                # with an undefined gradient: ctx.set_materialize_grads(False))
                if ctx.feat_sink.rows16 is not None:
                    raise RuntimeError("fused engine: two dense consumers of one fused output in row form are not "
                                       "supported (the second would overwrite the first one's gradient)")
                ctx.feat_sink.rows16 = (dx0, dx0.shape[1])
            else:
                Bc, Cc, Nc = ctx.a_shape
                da = dx0.view(Bc, Nc, -1)[:, :, :Cc].transpose(1, 2).float().contiguous()
        return (None, None, None, None, None, None, da, db, *grads)


def _params(convs, bns):
    ps = []
    for conv, bn in zip(convs, bns):
        if conv.bias is None or bn.weight is None:
            raise ValueError("fused engine expects Conv(bias=True) + affine BatchNorm, as the reference builds them")
        ps += [conv.weight, conv.bias, bn.weight, bn.bias]
    return ps


def _check_cuda(t):
    if not t.is_cuda:
        raise ValueError("hotrack_b200 has no CPU path")


def _no_coordinate_grad(*tensors):
    """The fused stacks return no gradient for coordinates (HandTrackNet's are inputs); asking for one must not pass
    silently."""
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise NotImplementedError("fused engine: gradients w.r.t. xyz / new_xyz are not implemented; use engine 'ops'")


def _run(kind, meta, convs, bns, training, a, b, precise=0):
    ps = _params(convs, bns)
    views = []
    for i, p in enumerate(ps):
        views.append(p.view(p.shape[0], -1) if i % 4 == 0 else p)  # conv weight (Cout,Cin,1[,1]) -> (Cout,Cin)
    with torch.cuda.device(ps[0].device):  # kernels launch on the CURRENT device: make it the tensors' one
        out = _MlpStack.apply(kind, meta, list(bns), ps, bool(training), int(precise), a, b, *views)
    rows, _MlpStack.last_rows = _MlpStack.last_rows, None  # set by forward (single-threaded hand-over)
    return attach_rows(out, rows) if rows is not None else out


_MlpStack.last_rows = None


def sa_scale(xyz, points, new_xyz, idx, centre_feat, convs, bns, training, precise=0):
    """One SA scale.  xyz (B,3,N), points (B,D,N)|None, new_xyz (B,3,S), idx (B,S,K) int32,
    centre_feat (B,E,S)|None -> (B,Cout,S)."""
    _check_cuda(xyz)
    if points is not None and points.shape[1] == 0:
        points = None
    _no_coordinate_grad(xyz, new_xyz)
    meta = (xyz.contiguous().float(), new_xyz.contiguous().float(), idx.contiguous().int(), False)
    return _run("sa", meta, convs, bns, training, points, centre_feat, precise)


def sa_group_all(xyz, points, convs, bns, training, precise=0):
    """Group-all SA.  xyz (B,3,N), points (B,D,N)|None -> (B,Cout,1); channel order [xyz, points]."""
    _check_cuda(xyz)
    _no_coordinate_grad(xyz)
    meta = (xyz.contiguous().float(), None, None, True)
    return _run("sa", meta, convs, bns, training, points, None, precise)


def three_nn_sq(xyz1_t, xyz2_t):
    """Squared distances (B,N,3) and int32 indices (B,N,3) of the three nearest xyz2 points of every xyz1 point --
    coordinates only, so remembered per forward pass (pointnet_utils.memo_call: the backbone asks for FP1's and FP2's ahead
    of time, on a side stream)."""
    from . import pointnet_utils as pu

    u, k = xyz1_t.contiguous().float(), xyz2_t.contiguous().float()

    def run():
        B, N, _ = u.shape
        d2 = torch.empty(B, N, 3, dtype=torch.float32, device=u.device)
        ix = torch.empty(B, N, 3, dtype=torch.int32, device=u.device)
        pc.three_nn_wrapper(B, N, k.shape[1], u, k, d2, ix)
        return d2, ix

    return pu.memo_call("three_nn", (u, k), (), run)


def fp_layer(xyz1_t, xyz2_t, points1, points2, convs, bns, training, reps=1, rows_only=False, precise=0):
    """FP layer.  xyz1_t (B,N,3), xyz2_t (B,S,3), points1 (B*reps,D1,N)|None, points2 (B*reps,D2,S) -> (B*reps,Cout,N).
    rows_only: the returned fp32 tensor is left UNINITIALISED (only its attached row form is valid) -- for a caller that
    hands it straight to another fused stack (the backbone's FP1 -> conv1 head) and saves a 67 MB transpose."""
    _check_cuda(points2)
    B, N, _ = xyz1_t.shape
    S = xyz2_t.shape[1]
    idx = dist2 = None
    if S > 1:
        dist2, idx = three_nn_sq(xyz1_t, xyz2_t)
        if reps > 1:
            dist2, idx = dist2.repeat_interleave(reps, dim=0), idx.repeat_interleave(reps, dim=0)
    return _run("fp", (idx, dist2, N, S, bool(rows_only)), convs, bns, training, points1, points2, precise)


def dense_stack(x, convs, bns, training, precise=0, rows_only=False):
    """relu(bn(conv1x1(x))) stack on (B,C,N) -> (B,Cout,N).  rows_only: as in fp_layer -- the returned fp32 tensor is left
    UNINITIALISED and only its attached row form is valid (a caller whose consumers are all fused stacks)."""
    _check_cuda(x)
    return _run("dense", (bool(rows_only),), convs, bns, training, x, None, precise)


def alg_bytes(name, a):
    """Algorithmic bytes of one C-ABI call of the fused kernels (bench.py's roofline line).  Two-plane ("_x2") calls
    move two fp16 planes where their lo pointers are set."""
    if name in ("pn2_mlp_gemm_fwd", "pn2_mlp_gemm_fwd_bn"):
        rows, kdim, n = a[:3]
        return rows * (kdim + n) * 2 + n * kdim * 2
    if name in ("pn2_mlp_gemm_fwd_x2", "pn2_mlp_gemm_fwd_bn_x2", "pn2_mlp_gemm_fwd_pool", "pn2_mlp_gemm_fwd_bn_pool"):
        rows, kdim, n = a[:3]
        pin, pout = (2 if a[4] else 1), (2 if a[12] else 1)
        return rows * (kdim * pin + n * pout) * 2 + n * kdim * 2 * pin
    if name == "pn2_mlp_gemm_dgrad":
        rows, n_red, k_out = a[:3]
        masked = a[11] != 0
        return rows * (2 * n_red + k_out + (k_out if masked else 0)) * 2 + n_red * k_out * 2
    if name == "pn2_mlp_gemm_wgrad":
        rows, n, kp = a[:3]
        return rows * (2 * n + kp) * 2 + n * kp * 4
    if name in ("pn2_pool_fwd", "pn2_pool_fwd_x2"):
        b, s, k, c = a[:4]
        planes = 2 if (name.endswith("_x2") and a[5]) else 1
        return b * s * k * c * 2 * planes + b * s * c * 4 + (b * s * c * 4 if k > 1 else 0)
    if name == "pn2_pool_bwd":
        b, s, k, c = a[:4]
        extra = (b * s * c * 4 if a[5] else 0) + (b * s * c * 2 if a[6] else 0)  # row-form gradients of fused consumers
        return (b * s * c * 4 if a[4] else 0) + extra + b * s * k * c * 2 + (b * s * c * (4 + 2) if k > 1 else b * s * c * 2)
    if name == "pn2_sa_build_rows":
        b, n, s, k = a[:4]
        return b * s * k * (a[19] * 2 + 4) + b * s * k * 2 * (a[8] + a[13])
    if name == "pn2_sa_build_rows_x2":
        b, n, s, k = a[:4]
        pin_f, pin_c, pout = (2 if a[8] else 1), (2 if a[14] else 1), (2 if a[21] else 1)
        return b * s * k * (a[22] * 2 * pout + 4) + b * s * k * 2 * (a[9] * pin_f + a[15] * pin_c)
    if name == "pn2_fp_build_rows":
        b, n, s = a[:3]
        return b * n * (a[17] * 2 + 24 + 2 * a[4]) + b * s * a[9] * 2
    if name == "pn2_fp_build_rows_x2":
        b, n, s = a[:3]
        pin_s, pin_c, pout = (2 if a[4] else 1), (2 if a[10] else 1), (2 if a[18] else 1)
        return b * n * (a[19] * 2 * pout + 24 + 2 * a[5] * pin_s) + b * s * a[11] * 2 * pin_c
    if name == "pn2_sa_rows_bwd":
        b, n, s, k = a[:4]
        return b * s * k * (a[6] * 2 + 4)
    if name == "pn2_fp_rows_bwd":
        b, n, s = a[:3]
        return b * n * (a[6] * 2 + 24)
    if name == "pn2_to_rows":
        b, c, n = a[:3]
        return b * c * n * 4 + b * n * a[7] * 2
    if name == "pn2_to_rows_x2":
        b, c, n = a[:3]
        return b * c * n * 4 + b * n * a[8] * 2 * (2 if a[7] else 1)
    return None
