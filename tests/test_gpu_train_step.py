"""GPU parity of the training-step kernels of csrc/train_step.cu: the fused pattern loss against
the torch formulation of nn/metrics/composed_loss.py:301-321 + nn/metrics/losses.py:19-51, nt_adam_step against torch.optim.Adam
(nn/trainer.py:64) under a stepping OneCycleLR (nn/trainer.py:73-80), and the CUDA-graph step built on them."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('components', [['shape', 'loop', 'rotation', 'translation'], ['shape'], ['loop', 'translation']])
def test_pattern_loss_kernel_matches_torch(cuda_device, components):
    from garment_pattern_estimation_b200 import ops
    from garment_pattern_estimation_b200.losses import panel_loop_loss
    from oracle import model as om
    dev = cuda_device
    B = 5
    gt = om.synthetic_ground_truth(B, seed=3, device=dev)
    g = torch.Generator().manual_seed(4)
    # predictions as strided views, like the slices of the decoder output (nets.forward_panel_decode)
    dec = torch.randn(B, 23, 14, 8, generator=g).to(dev).requires_grad_(True)
    place = torch.randn(B * 23, 7, generator=g).to(dev).requires_grad_(True)
    dec2, place2 = dec.detach().clone().requires_grad_(True), place.detach().clone().requires_grad_(True)
    pad = torch.tensor([0.25, -0.5])

    def preds(d, p):
        return d[:, :, :, :4], p[:, :4].view(B, 23, 4), p[:, 4:].view(B, 23, 3)

    o, r, t = preds(dec, place)
    got = ops.pattern_loss(o, r, t, gt['outlines'], gt['rotations'], gt['translations'], gt['num_edges'], components, 0.7,
                           (0.25, -0.5))
    o2, r2, t2 = preds(dec2, place2)
    want = {'shape': F.mse_loss(o2, gt['outlines']), 'loop': panel_loop_loss(o2, gt['num_edges'].int().view(-1), pad),
            'rotation': F.mse_loss(r2, gt['rotations']), 'translation': F.mse_loss(t2, gt['translations'])}
    total = sum((0.7 if k == 'loop' else 1.0) * v for k, v in want.items() if k in components)
    assert abs(float(got[0]) - float(total)) <= 1e-5 * abs(float(total))
    for i, k in enumerate(('shape', 'loop', 'rotation', 'translation')):
        if k in components:
            assert abs(float(got[i + 1]) - float(want[k])) <= 1e-5 * abs(float(want[k])), k
        else:
            assert float(got[i + 1]) == 0.0
    (got[0] * 1.5).backward()
    (total * 1.5).backward()
    for a, b, name in ((dec.grad, dec2.grad, 'decoder output'), (place.grad, place2.grad, 'placement')):
        if b is None:                                      # term not requested: the kernel returns exact zeros
            assert float(a.abs().max()) == 0.0, name
            continue
        assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max()) + 1e-12, name


@pytest.mark.parametrize('n,weight_decay', [(1842589, 0.0), (1003, 0.01)])
def test_adam_kernel_matches_torch_adam_under_onecycle(cuda_device, n, weight_decay):
    from garment_pattern_estimation_b200 import ops
    dev = cuda_device
    g = torch.Generator().manual_seed(n)
    p0 = torch.randn(n, generator=g).to(dev)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=2e-3, weight_decay=weight_decay)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=2e-3, epochs=2, steps_per_epoch=5, cycle_momentum=False)
    pad = (n + 3) // 4 * 4
    p = torch.zeros(pad, device=dev); p[:n] = p0
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    grad = torch.zeros_like(p)
    lr = torch.zeros(1, device=dev)
    state = torch.zeros(2, device=dev)
    for step in range(8):
        gstep = (torch.randn(n, generator=g) * (10.0 ** (step % 3 - 2))).to(dev)
        ref.grad = gstep.clone()
        lr.fill_(opt.param_groups[0]['lr'])
        grad[:n] = gstep * 4.0                                        # the kernel applies grad_scale = 1/4 (data-parallel average)
        ops.adam_step(p, grad, m, v, lr, state, weight_decay=weight_decay, grad_scale=0.25, zero_grad=True)
        opt.step()
        sched.step()
        assert float(grad.abs().max()) == 0.0                         # zero_grad folded into the kernel
        err = float((p[:n] - ref.detach()).abs().max())
        assert err <= 2e-6 * float(ref.detach().abs().max()), (step, err)
    assert int(state[0].item()) == 8


def test_graphed_step_with_flat_adam_follows_the_scheduler(cuda_device):
    """parallel.GraphedTrainStep + parallel.FlatAdam + OneCycleLR: the replayed graph must use the learning rate the scheduler set
    for THIS step (device scalar), take the same steps as the eager loop of the reference (torch Adam + OneCycleLR), and leave the
    flat gradient buffer zeroed."""
    import garment_pattern_estimation_b200 as gpe
    from garment_pattern_estimation_b200.parallel import FlatAdam, FlatDataParallel, GraphedTrainStep
    from oracle import model as om
    dev = cuda_device
    dc, nc = dict(om.ATT_DATA_CONFIG), dict(om.ATT_NN_CONFIG)
    lc = {'loss_components': ['shape', 'loop', 'rotation', 'translation'], 'quality_components': [],
          'panel_origin_invariant_loss': False, 'panel_order_inariant_loss': False}
    B, N = 2, 256
    batches = []
    for i in range(4):
        x = torch.randn(B, N, 3, generator=torch.Generator().manual_seed(30 + i)).to(dev)
        batches.append((x, {k: v.to(dev) for k, v in om.synthetic_ground_truth(B, seed=40 + i).items()}))
    torch.manual_seed(7)
    state = (om.init_state(3, B * 23, 250).to(dev), om.init_state(3, B * 23, 250).to(dev))

    def run(graphed):
        torch.manual_seed(5)
        model = gpe.GarmentSegmentPattern3D(dict(dc), dict(nc), dict(lc)).to(dev).train()
        wrapper = FlatDataParallel(model, device_ids=[dev], auto_reduce=False)
        if graphed:
            opt = FlatAdam(wrapper, lr=2e-3)
        else:
            opt = torch.optim.Adam(model.parameters(), lr=2e-3)
        sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=2e-3, epochs=1, steps_per_epoch=4, cycle_momentum=False)
        step = GraphedTrainStep(wrapper, opt, batches[0][0], batches[0][1], warmup=2, forward_kwargs={'lstm_state': state}) if graphed else None
        losses, lrs = [], []
        for x, gt in batches:
            lrs.append(opt.param_groups[0]['lr'])
            if graphed:
                losses.append(float(step(x, gt)))
                assert abs(float(opt.lr_dev) - lrs[-1]) <= 1e-9 + 1e-6 * lrs[-1]
            else:
                out = wrapper(x, lstm_state=state)
                loss = model.loss(out, gt)[0]
                loss.backward()
                opt.step()
                wrapper.zero_grad()
                losses.append(float(loss))
            sched.step()
        torch.cuda.synchronize()
        if graphed:
            assert float(wrapper.flat_grad.abs().max()) == 0.0 and opt.steps_taken == 4
        # parameters only: the BatchNorm running statistics of the graphed arm also saw the warm-up passes of the capture
        return losses, lrs, {k: v.detach().clone() for k, v in model.named_parameters()}

    l_ref, lr_ref, sd_ref = run(False)
    l_g, lr_g, sd_g = run(True)
    assert lr_ref == lr_g and len(set(lr_g)) == 4
    assert abs(l_ref[0] - l_g[0]) <= 1e-5 * abs(l_ref[0])
    for a, b in zip(l_ref, l_g):          # Adam's first steps move every weight by ~lr * sign(g): chaotic amplification (see test_gpu_model)
        assert abs(a - b) <= 1.5e-1 * abs(a), (l_ref, l_g)
    for k in sd_ref:
        if sd_ref[k].dtype.is_floating_point:
            assert float((sd_g[k] - sd_ref[k]).abs().max()) <= 1.6e-2 + 2e-2 * float(sd_ref[k].abs().max()), k
