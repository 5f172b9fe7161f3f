"""Data-parallel training step: one process per GPU, ONE flat-buffer gradient all-reduce per step over NCCL/NVLink.

Replaces the reference's single-process ``nn.DataParallel`` wrapper (nn/train.py:123-125; used by nn/trainer.py:96 and
nn/experiment.py:234), which scatters the batch, re-broadcasts all parameters on every forward, gathers outputs and
reduces gradients to device 0.  Clouds are independent, so the batch dimension is the only natural
partition: every rank runs the full model on its shard; the only exchange is the gradient sum.

Semantics kept from the reference wrapper:
  * ``.module`` and ``.device_ids`` attributes (read by nn/trainer.py:62,96 and nn/data/wrapper.py:222);
  * BatchNorm batch statistics are PER REPLICA (DataParallel has no SyncBN); running statistics / num_batches_tracked are
    taken from rank 0 (``sync_buffers()``; DataParallel keeps replica 0's);
  * state_dict keys carry the ``module.`` prefix (checkpoints written through the wrapper, nn/trainer.py:275-291).
Differences: the loss is evaluated per rank on the local shard (mean of per-rank means == the gathered-batch mean for equal
shards); parameters that receive no gradient (``feature_extractor.lin`` with local attention) are simply
all-reduced as zeros -- no unused-parameter search is needed because the buffer is flat and static.
"""
import torch
import torch.distributed as dist
import torch.nn as nn


class FlatDataParallel(nn.Module):
    def __init__(self, module, device_ids=None, process_group=None, average=True, auto_reduce=True):
        """auto_reduce: all-reduce automatically at the end of every backward() (queued engine callback, the way
        torch's DDP finalises), so an unmodified training loop (nn/trainer.py:96-99: loss.backward();
        optimizer.step()) needs no extra call.  With auto_reduce=False call reduce_gradients() yourself."""
        super().__init__()
        self.module = module
        p0 = next(module.parameters())
        self.device_ids = list(device_ids) if device_ids is not None else [p0.device]
        self.process_group = process_group
        self.average = average
        self.auto_reduce = auto_reduce
        self._callback_queued = False
        self._build_flat_grads()
        if self.world_size > 1:
            self.sync_parameters()
        if auto_reduce:
            for p in self._params:
                p.register_post_accumulate_grad_hook(self._on_grad)

    def _on_grad(self, param):
        if not self._callback_queued and self.world_size > 1:
            self._callback_queued = True
            torch.autograd.Variable._execution_engine.queue_callback(self._finalize_backward)

    def _finalize_backward(self):
        self._callback_queued = False
        self.reduce_gradients()

    # ---- process-group helpers ---------------------------------------------------------------------------
    @property
    def world_size(self):
        return dist.get_world_size(self.process_group) if dist.is_available() and dist.is_initialized() else 1

    @property
    def rank(self):
        return dist.get_rank(self.process_group) if dist.is_available() and dist.is_initialized() else 0

    # ---- flat gradient buffer ----------------------------------------------------------------------------
    def _build_flat_grads(self):
        """All parameter gradients become views into one contiguous fp32 buffer, so the step needs exactly one
        collective and no packing copies (1 842 589 elements = 7.4 MB for the attention model)."""
        params = [p for p in self.module.parameters() if p.requires_grad]
        total = sum(p.numel() for p in params)
        self.flat_grad = torch.zeros(total, dtype=params[0].dtype, device=params[0].device)
        off = 0
        for p in params:
            n = p.numel()
            p.grad = self.flat_grad[off:off + n].view_as(p)
            off += n
        self._params = params

    def zero_grad(self, set_to_none=False):
        """Keeps the flat views alive (never sets .grad to None)."""
        self.flat_grad.zero_()

    def sync_parameters(self, src=0):
        """Broadcast parameters and buffers from `src` (done once at wrap time; DataParallel does it every forward)."""
        for t in list(self.module.parameters()) + list(self.module.buffers()):
            dist.broadcast(t.data, src=src, group=self.process_group)

    def sync_buffers(self, src=0):
        """BN running statistics follow rank 0 (what DataParallel's replica 0 would have accumulated)."""
        if self.world_size > 1:
            for b in self.module.buffers():
                dist.broadcast(b.data, src=src, group=self.process_group)

    def reduce_gradients(self, async_op=False):
        """Sum (and average) the flat gradient over all ranks.  Call after backward(), before optimizer.step()."""
        # optimizer.zero_grad(set_to_none=True) (torch's default, and what the reference Trainer calls) drops the views;
        # re-adopt such gradients into the flat buffer (one copy) instead of failing
        lo = self.flat_grad.data_ptr()
        hi = lo + self.flat_grad.numel() * self.flat_grad.element_size()
        off = 0
        for p in self._params:
            n = p.numel()
            view = self.flat_grad[off:off + n].view_as(p)
            if p.grad is None:
                view.zero_()
                p.grad = view
            elif not (lo <= p.grad.data_ptr() < hi):
                view.copy_(p.grad)
                p.grad = view
            off += n
        if self.world_size == 1:
            return None
        if self.average:
            self.flat_grad.div_(self.world_size)
        return dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.process_group, async_op=async_op)

    # ---- nn.Module surface ---------------------------------------------------------------------------------
    def forward(self, *inputs, **kwargs):
        return self.module(*inputs, **kwargs)

    def shard(self, batch_tensor):
        """This rank's slice of a global batch (rank r takes clouds [r*B/G, (r+1)*B/G))."""
        B = batch_tensor.shape[0]
        if B % self.world_size != 0:
            raise RuntimeError('global batch {} is not divisible by world size {}'.format(B, self.world_size))
        per = B // self.world_size
        return batch_tensor[self.rank * per:(self.rank + 1) * per]
