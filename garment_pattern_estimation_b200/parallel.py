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

from . import ops


class FlatDataParallel(nn.Module):
    def __init__(self, module, device_ids=None, process_group=None, average=True, auto_reduce=True, flatten_parameters=True):
        """auto_reduce: all-reduce automatically at the end of every backward() (queued engine callback, the way
        torch's DDP finalises), so an unmodified training loop (nn/trainer.py:96-99: loss.backward();
        optimizer.step()) needs no extra call.  With auto_reduce=False call reduce_gradients() yourself.
        flatten_parameters: also move the parameter VALUES into one contiguous buffer (``flat_param``; every ``p.data`` becomes a
        view of it, state_dict / optimizers see no difference) so that ``FlatAdam`` can update the whole model with one kernel."""
        super().__init__()
        self.module = module
        p0 = next(module.parameters())
        self.device_ids = list(device_ids) if device_ids is not None else [p0.device]
        self.process_group = process_group
        self.average = average
        self.auto_reduce = auto_reduce
        self._callback_queued = False
        self._auto_suspended = False        # set by GraphedTrainStep, which reduces the gradients itself
        self.flat_param = None
        if flatten_parameters:
            self._build_flat_params()
        self._build_flat_grads()
        if self.world_size > 1:
            self.sync_parameters()
        if auto_reduce:
            for p in self._params:
                p.register_post_accumulate_grad_hook(self._on_grad)

    def _on_grad(self, param):
        if not self._callback_queued and self.world_size > 1 and not self._auto_suspended:
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

    # ---- flat parameter / gradient buffers ----------------------------------------------------------------
    def _build_flat_params(self):
        """Parameter values as views into one contiguous fp32 buffer, in the order (and with the offsets) of the flat gradient
        buffer; every segment starts at a multiple of 4 elements so that 16-byte vector accesses never straddle two tensors."""
        params = [p for p in self.module.parameters() if p.requires_grad]
        if any(p.dtype != torch.float32 for p in params):
            return
        offsets, total = [], 0
        for p in params:
            offsets.append(total)
            total += (p.numel() + 3) // 4 * 4
        flat = torch.zeros(total, dtype=torch.float32, device=params[0].device)
        for p, off in zip(params, offsets):
            flat[off:off + p.numel()].copy_(p.data.reshape(-1))
            p.data = flat[off:off + p.numel()].view_as(p)
        self.flat_param, self._offsets = flat, offsets

    def _build_flat_grads(self):
        """All parameter gradients become views into one contiguous fp32 buffer, so the step needs exactly one
        collective and no packing copies (1 842 589 elements = 7.4 MB for the attention model)."""
        params = [p for p in self.module.parameters() if p.requires_grad]
        if self.flat_param is None:
            offsets, total = [], 0
            for p in params:
                offsets.append(total)
                total += p.numel()
            self._offsets = offsets
        else:
            total = self.flat_param.numel()
        self.flat_grad = torch.zeros(total, dtype=params[0].dtype, device=params[0].device)
        for p, off in zip(params, self._offsets):
            p.grad = self.flat_grad[off:off + p.numel()].view_as(p)
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

    def adopt_gradients(self):
        """Bring every ``.grad`` back into the flat buffer (one copy per stray gradient; nothing to do in the steady state)."""
        lo = self.flat_grad.data_ptr()
        hi = lo + self.flat_grad.numel() * self.flat_grad.element_size()
        for p, off in zip(self._params, self._offsets):
            if p.grad is not None and lo <= p.grad.data_ptr() < hi:
                continue
            view = self.flat_grad[off:off + p.numel()].view_as(p)
            if p.grad is None:
                view.zero_()
            else:
                view.copy_(p.grad)
            p.grad = view

    def reduce_gradients(self, async_op=False):
        """Sum (and average) the flat gradient over all ranks.  Call after backward(), before optimizer.step()."""
        # optimizer.zero_grad(set_to_none=True) (torch's default, and what the reference Trainer calls) drops the views;
        # re-adopt such gradients into the flat buffer (one copy) instead of failing
        self.adopt_gradients()
        if self.world_size == 1:
            return None
        if self.average:
            self.flat_grad.div_(self.world_size)
        return dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.process_group, async_op=async_op)

    def sum_gradients(self):
        """The collective alone (no division): FlatAdam folds the 1 / world_size of the average into its update kernel."""
        self.adopt_gradients()
        if self.world_size > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.process_group)

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


class FlatAdam(torch.optim.Optimizer):
    """torch.optim.Adam (nn/trainer.py:64; amsgrad off) for a FlatDataParallel-wrapped model: ONE sm_100a kernel per step
    (``nt_adam_step``) updates the flat parameter buffer from the flat gradient buffer -- including the 1 / world_size of the
    data-parallel average and the zeroing of the gradients -- instead of ~110 small per-parameter kernels.

    It is a ``torch.optim.Optimizer`` (one param group holding the flat buffer), so ``torch.optim.lr_scheduler.OneCycleLR``
    (nn/trainer.py:73-80) drives it unchanged: the scheduler writes a Python float into ``param_groups[0]['lr']``; ``step()``
    mirrors it into a device scalar that the kernel reads, which keeps a CUDA-graph replay of the step correct while the
    learning rate moves (call ``sync_lr()`` before replaying a captured step).  The step counter lives on the device."""

    def __init__(self, wrapper, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if wrapper.flat_param is None:
            raise RuntimeError('FlatAdam needs FlatDataParallel(..., flatten_parameters=True) with fp32 parameters')
        if not wrapper.flat_param.is_cuda:
            raise RuntimeError('FlatAdam runs on a CUDA device (the B200 hot path has no CPU fallback)')
        flat = torch.nn.Parameter(wrapper.flat_param, requires_grad=True)
        flat.grad = wrapper.flat_grad
        super().__init__([flat], dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.wrapper = wrapper
        dev = wrapper.flat_param.device
        # moments and step counter live in self.state like torch's Adam keeps them, so optimizer.state_dict() /
        # load_state_dict() (checkpoints: nn/trainer.py:275-291) carry them; 'step' is [steps taken, internal counter]
        self.state[flat] = dict(step=torch.zeros(2, dtype=torch.float32, device=dev), exp_avg=torch.zeros_like(wrapper.flat_param),
                                exp_avg_sq=torch.zeros_like(wrapper.flat_param))
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self._lr_host = float(lr)

    def sync_lr(self):
        lr = float(self.param_groups[0]['lr'])
        if lr != self._lr_host:
            self.lr_dev.fill_(lr)
            self._lr_host = lr

    @torch.no_grad()
    def step(self, closure=None, zero_grad=False, sync_lr=True):
        """sync_lr=False inside a CUDA-graph capture (the host-side mirror must not be baked into the graph)."""
        if sync_lr:
            self.sync_lr()
        g = self.param_groups[0]
        w = self.wrapper
        st = self.state[g['params'][0]]
        scale = 1.0 / w.world_size if (w.average and w.world_size > 1) else 1.0
        ops.adam_step(w.flat_param, w.flat_grad, st['exp_avg'], st['exp_avg_sq'], self.lr_dev, st['step'], betas=g['betas'],
                      eps=g['eps'], weight_decay=g['weight_decay'], grad_scale=scale, zero_grad=zero_grad)

    def zero_grad(self, set_to_none=False):
        self.wrapper.zero_grad()

    @property
    def steps_taken(self):
        return int(self.state[self.param_groups[0]['params'][0]]['step'][0].item())


class GraphedTrainStep:
    """One training step (forward -> loss -> backward [-> gradient all-reduce] -> optimizer.step -> zero_grad) captured ONCE
    into a CUDA graph and replayed per batch: the ~450 kernel launches of a step (libnt_b200 kernels, cuDNN LSTM, loss,
    Adam) cost one host call, so the step time is the GPU time even when the host is slow or shared between ranks.

    The reference's loop (nn/trainer.py:92-103) is the eager equivalent; numerics are identical because the graph replays
    exactly the kernels the eager step launches.  Requirements of CUDA-graph capture:
      * fixed batch shape (the shape of ``example_x`` / ``example_gt``); other shapes fall back to ``eager_step``;
      * an optimizer whose step is capturable (``torch.optim.Adam(..., capturable=True)``); a learning-rate schedule (the
        reference uses OneCycleLR, nn/trainer.py:73-80) must then act on a TENSOR learning rate
        (``lr=torch.tensor(2e-3, device=...)``), because a Python-float lr is baked into the captured kernels;
      * ``wrapper`` is a FlatDataParallel (its flat gradient buffer keeps every ``.grad`` at a static address).
    With more than one rank the NCCL all-reduce of the flat gradient buffer is captured INTO the graph as well (PyTorch's NCCL
    process group supports stream capture), followed by the optimizer step; if that capture fails on a given software stack the
    step falls back to graph = forward + loss + backward with the collective and the update launched eagerly after the replay.
    With ``parallel.FlatAdam`` the update is one kernel and the gradient average / zeroing are folded into it.
    The wrapper's ``auto_reduce`` hook is suspended: this class issues the one collective of the step itself.

    Construction runs ``warmup`` eager forward/backward passes on the example batch on a side stream (required before
    capture; BatchNorm running statistics see these batches like any other training batch) and creates the optimizer state
    with one step on all-zero gradients (parameters do not move for Adam / SGD without weight decay).
    """

    def __init__(self, wrapper, optimizer, example_x, example_gt, warmup=3, capture_update=True, forward_kwargs=None):
        if not example_x.is_cuda:
            raise RuntimeError('GraphedTrainStep needs CUDA tensors (the B200 hot path has no CPU fallback)')
        self.wrapper, self.optimizer = wrapper, optimizer
        self.model = wrapper.module
        self.forward_kwargs = dict(forward_kwargs or {})
        self.capture_update = bool(capture_update)
        self.flat_adam = isinstance(optimizer, FlatAdam)
        wrapper._auto_suspended = True      # never reduce twice (hook + explicit call) and never from inside backward here
        self.static_x = example_x.clone()
        self.static_gt = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in example_gt.items()}
        side = torch.cuda.Stream(device=example_x.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._forward_backward()
                wrapper.zero_grad()
            if wrapper.world_size > 1:
                wrapper.sum_gradients()      # communicator warm-up outside of any capture (gradients are zero here)
            if self.capture_update and not self.flat_adam:
                # Optimizer state (Adam moments, step counters) must EXIST before capture, otherwise its lazy initialisation
                # would be recorded into the graph and re-run on every replay.  One step on all-zero gradients creates it
                # without moving the parameters (true for Adam / SGD without weight decay); step counters are rewound.
                optimizer.step()
                for st in optimizer.state.values():
                    if torch.is_tensor(st.get('step')):
                        st['step'].zero_()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(self.graph):
                self.static_loss = self._forward_backward()
                if self.capture_update:
                    self._update(in_capture=True)
        except Exception:
            if not (self.capture_update and wrapper.world_size > 1):
                raise
            # the collective could not be captured on this stack: graph = forward + loss + backward only
            torch.cuda.synchronize()
            wrapper.zero_grad()
            self.capture_update = False
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.static_loss = self._forward_backward()

    def _loss(self, out, gt):
        """The training loss only: the no-grad quality metrics (diagnostics; they need a host read) are switched off for the
        step, like ``loss.with_quality_eval = False`` in the reference."""
        loss_obj = self.model.loss
        had = getattr(loss_obj, 'with_quality_eval', None)
        if had:
            loss_obj.with_quality_eval = False
        try:
            return loss_obj(out, gt)[0]
        finally:
            if had:
                loss_obj.with_quality_eval = True

    def _forward_backward(self):
        out = self.wrapper(self.static_x, **self.forward_kwargs)
        loss = self._loss(out, self.static_gt)
        loss.backward()
        return loss.detach()

    def _update(self, in_capture=False):
        if self.flat_adam:
            self.wrapper.sum_gradients()                                  # the 1 / world_size is folded into the update kernel
            self.optimizer.step(zero_grad=True, sync_lr=not in_capture)
        else:
            self.wrapper.reduce_gradients()
            self.optimizer.step()
            self.wrapper.zero_grad()

    def eager_step(self, x, gt):
        out = self.wrapper(x, **self.forward_kwargs)
        loss = self._loss(out, gt)
        loss.backward()
        self._update()
        return loss.detach()

    def __call__(self, x, gt):
        """x / gt: device (or pinned host) tensors of the captured shapes.  Returns the loss tensor of this step (a static
        buffer: read it before the next call)."""
        if tuple(x.shape) != tuple(self.static_x.shape):
            return self.eager_step(x.to(self.static_x.device), gt)
        self.static_x.copy_(x, non_blocking=True)
        for k, v in self.static_gt.items():
            if torch.is_tensor(v):
                v.copy_(gt[k], non_blocking=True)
        if self.flat_adam:
            self.optimizer.sync_lr()          # a scheduler may have moved param_groups[0]['lr'] since the last step
        self.graph.replay()
        if not self.capture_update:
            self._update()
        return self.static_loss
