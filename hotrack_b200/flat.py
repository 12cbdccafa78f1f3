"""Flat parameter / gradient buffers, one-bucket gradient all-reduce and the fused Adam step.

The reference is single-GPU; the data-parallel training path this package adds
shards the batch over ranks (one process per GPU, identical replicas) and needs exactly one
exchange per step: a SUM all-reduce of the gradients.  All parameters are re-pointed into ONE flat
fp32 buffer and their ``.grad`` into a second one, so that exchange is a single NCCL call over
NVLink/NVSwitch on the whole buffer and the optimiser is a single kernel (csrc/optim.cu).
BatchNorm statistics stay per rank, as in the single-GPU reference (no SyncBN there).
"""
import torch
import torch.distributed as dist


class FlatParams:
    """Re-points every parameter of ``module`` (and its .grad) into flat fp32 buffers.

    Parameters that never receive a gradient simply keep a zero gradient slice.  Offsets are
    16-byte aligned so the Adam kernel can use float4 accesses on each tensor boundary too.
    """

    def __init__(self, module, align=4, tail=0):
        self.params = [p for p in module.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("module has no trainable parameters")
        dev, dt = self.params[0].device, torch.float32
        self.offsets, off = [], 0
        for p in self.params:
            if p.dtype != dt or p.device != dev:
                raise ValueError("FlatParams needs fp32 parameters on one device")
            self.offsets.append(off)
            off += (p.numel() + align - 1) // align * align
        self.numel = off
        self.data = torch.zeros(off, dtype=dt, device=dev)
        # ``tail`` extra floats behind the gradients, zeroed by the same memset (TrainStep's accumulator arena)
        self._gbuf = torch.zeros(off + tail, dtype=dt, device=dev)
        self.grad = self._gbuf[:off]
        self.tail = self._gbuf[off:]
        for p, o in zip(self.params, self.offsets):
            n = p.numel()
            self.data[o:o + n].copy_(p.data.reshape(-1))
            p.data = self.data[o:o + n].view_as(p)
            p.grad = self.grad[o:o + n].view_as(p)

    def zero_grad(self):
        """One memset; keeps every p.grad aliased to the flat buffer (autograd accumulates in place)."""
        self._gbuf.zero_()
        for p, o in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:
                p.grad = self.grad[o:o + p.numel()].view_as(p)

    def broadcast(self, src=0):
        """Make every replica start from rank ``src``'s weights."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.broadcast(self.data, src)

    def allreduce_grads(self):
        """The one exchange of the data-parallel step: SUM over ranks of the whole gradient buffer.
        Returns the factor the optimiser must scale gradients by (1/world) to get the mean."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM)
            return 1.0 / dist.get_world_size()
        return 1.0


class FlatAdam:
    """torch.optim.Adam semantics (reference trainer.py:66-73: lr 1e-4, weight_decay 1e-4) over a
    FlatParams, one CUDA kernel per step (pn2_adam_step).  CUDA only: no CPU fallback."""

    def __init__(self, flat, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4):
        self.flat, self.lr, self.betas, self.eps, self.weight_decay = flat, lr, betas, eps, weight_decay
        self.exp_avg = torch.zeros_like(flat.data)
        self.exp_avg_sq = torch.zeros_like(flat.data)
        # the step number lives on the device so that a captured step can be replayed (CUDA graphs)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=flat.data.device)

    @property
    def t(self):
        return int(self.step_dev.item())

    def step(self, grad_scale=1.0):
        from . import _lib

        f = self.flat
        if not f.data.is_cuda:
            raise RuntimeError("FlatAdam runs on CUDA tensors only (hotrack_b200 has no CPU path)")
        _lib.call("pn2_adam_step", f.numel, f.data.data_ptr(), f.grad.data_ptr(), self.exp_avg.data_ptr(),
                  self.exp_avg_sq.data_ptr(), self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                  0, self.step_dev.data_ptr(), float(grad_scale), torch.cuda.current_stream().cuda_stream)


def shard_batch(n_items, rank, world):
    """Contiguous, balanced split of ``n_items`` clouds over ``world`` ranks -> (start, stop) of ``rank``."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)
