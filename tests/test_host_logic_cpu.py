"""Host-side logic added in round 2 that needs no GPU: the per-forward coordinate-transpose memo, the int32 twins of group
indices, the shared gradient + arena buffer of FlatParams / ZeroArena."""
import torch

from hotrack_b200 import fused, pointnet_utils as pu
from hotrack_b200.flat import FlatParams


def test_t_contig_remembers_only_inside_a_scope():
    x = torch.randn(2, 3, 16)
    a, b = pu.t_contig(x), pu.t_contig(x)
    assert a.data_ptr() != b.data_ptr() and torch.equal(a, x.transpose(1, 2))  # no scope: plain copies
    with pu.coord_scope():
        a, b = pu.t_contig(x), pu.t_contig(x)
        assert a.data_ptr() == b.data_ptr() and a.is_contiguous() and torch.equal(a, x.transpose(1, 2))
        assert pu.t_contig(a).data_ptr() == x.data_ptr()          # ... and back: the original, not a third copy
        assert pu.t_contig(x[:]).data_ptr() == a.data_ptr()       # a fresh view object of the same memory finds it too
        with pu.coord_scope():                                    # nested scopes share the outer memo
            assert pu.t_contig(x).data_ptr() == a.data_ptr()
        assert pu.t_contig(x).data_ptr() == a.data_ptr()          # ... and do not close it
        x.add_(1.0)                                               # an in-place change bumps the version: fresh copy
        c = pu.t_contig(x)
        assert c.data_ptr() != a.data_ptr() and torch.equal(c, x.transpose(1, 2))
    assert pu._MEMO is None
    with pu.coord_scope():
        v = x.transpose(1, 2).contiguous().transpose(1, 2)        # already the transpose of a contiguous tensor: a view
        assert pu.t_contig(v).data_ptr() == v.data_ptr()
        g = x.clone().requires_grad_(True)                        # tensors autograd tracks are never memoised
        assert pu.t_contig(g).data_ptr() != pu.t_contig(g).data_ptr()


def test_group_index_twins():
    i32 = torch.randint(0, 100, (2, 21, 16), dtype=torch.int32)
    i64 = pu._as_long(i32)
    assert i64.dtype == torch.int64 and torch.equal(i64, i32.long())
    assert pu._as_int(i64).data_ptr() == i32.data_ptr()            # the kernels get the original back, no conversion
    i64[0, 0, 0] += 1                                              # modified by the caller: the twin is stale
    back = pu._as_int(i64)
    assert back.data_ptr() != i32.data_ptr() and torch.equal(back, i64.int())
    plain = torch.randint(0, 100, (2, 21, 16))
    assert torch.equal(pu._as_int(plain), plain.int())


def test_flat_params_tail_is_zeroed_with_the_gradients():
    m = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    flat = FlatParams(m, tail=64)
    assert flat.tail.numel() == 64 and flat.grad.numel() == flat.numel
    assert flat.tail.data_ptr() == flat.grad.data_ptr() + 4 * flat.numel and flat.numel % 4 == 0
    arena = fused.ZeroArena(buf=flat.tail)
    a = arena.take(10)[:10]  # take() hands out whole 16-byte granules
    a.fill_(3.0)
    for p in m.parameters():
        p.grad.fill_(2.0)
    arena.begin()                              # a borrowed buffer is NOT zeroed by the arena ...
    assert float(flat.tail.sum()) == 30.0
    flat.zero_grad()                           # ... but by the one memset of the gradients
    assert float(flat.tail.abs().sum()) == 0.0 and all(float(p.grad.abs().sum()) == 0.0 for p in m.parameters())
    assert all(p.grad.data_ptr() == flat.grad.data_ptr() + 4 * o for p, o in zip(flat.params, flat.offsets))
    own = fused.ZeroArena(torch.device("cpu"), floats=16)
    own.take(8).fill_(1.0)
    own.begin()                                # an arena that owns its buffer zeroes it itself
    assert float(own.buf.sum()) == 0.0 and own.take(20) is None


def test_memo_call_remembers_by_operation_scalars_and_tensor_identity():
    x, y = torch.randn(2, 16, 3), torch.randn(2, 4, 3)
    calls = []

    def search():
        calls.append(1)
        return torch.cdist(y, x).argmin(-1)

    assert pu.memo_call("nn", (x, y), (1,), search) is not pu.memo_call("nn", (x, y), (1,), search)  # no scope: no memo
    assert len(calls) == 2
    with pu.coord_scope():
        a = pu.memo_call("nn", (x, y), (1,), search)
        assert pu.memo_call("nn", (x, y), (1,), search) is a and len(calls) == 3
        assert pu.memo_call("nn", (x, y), (2,), search) is not a and len(calls) == 4       # other scalars
        assert pu.memo_call("other", (x, y), (1,), search) is not a and len(calls) == 5    # other operation
        assert pu.memo_call("nn", (x[:], y), (1,), search) is a                            # same memory, new view object
        assert pu.memo_call("nn", (x.clone(), y), (1,), search) is not a and len(calls) == 6
        g = x.clone().requires_grad_(True)
        pu.memo_call("nn", (g, y), (1,), search), pu.memo_call("nn", (g, y), (1,), search)
        assert len(calls) == 8                                                             # autograd-tracked inputs: never
    assert pu._MEMO is None
