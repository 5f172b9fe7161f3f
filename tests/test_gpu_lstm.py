"""GPU parity of the persistent tcgen05 LSTM decoder (csrc/lstm.cu) against torch.nn.LSTM in
fp32 (cuDNN with TF32 off -- tests/conftest.py), the operator the reference calls at nn/net_blocks.py:373,393.

Tolerances: outputs max|a-b| / max|b| <= 1e-3 (BASELINE.json north_star; the BF16x3 products deliver ~2e-5), gradients relative
L2 <= 1e-3 (the LSTM has no ReLU kinks, so unlike the EdgeConv gradients these are smooth and tight)."""
import pytest
import torch

from helpers import assert_close

pytestmark = pytest.mark.gpu


def _reference(lstm, x, h0, c0, T):
    seq = x.unsqueeze(1).repeat(1, T, 1)                    # nn/net_blocks.py:388
    out, _ = lstm(seq, (h0, c0))
    return out                                              # [R, T, H]


@pytest.mark.parametrize('R,T,L,H,E', [
    (736, 14, 3, 250, 250),       # C2: 32 clouds x 23 panels -> three 256-row tiles
    (184, 14, 3, 250, 250),       # C3 at 8 ranks: 8 clouds per GPU -> two 128-row tiles
    (32, 23, 2, 250, 250),        # baseline model's pattern decoder (rows = clouds, 23 steps, 2 layers)
    (50, 3, 1, 100, 37),          # small / odd sizes, one layer
    (1500, 5, 3, 250, 250),       # more row tiles than fit one launch (two cooperative launches)
])
@pytest.mark.parametrize('training', [True, False])
def test_lstm_decoder_matches_torch_lstm(cuda_device, R, T, L, H, E, training):
    from garment_pattern_estimation_b200 import ops
    dev = cuda_device
    g = torch.Generator().manual_seed(R + T)
    lstm = torch.nn.LSTM(E, H, L, batch_first=True).to(dev)
    with torch.no_grad():
        for name, p in lstm.named_parameters():
            if 'weight' in name:
                torch.nn.init.kaiming_normal_(p)            # nn/net_blocks.py:318-333
            else:
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
    x = torch.randn(R, E, generator=g).to(dev)
    h0 = (0.3 * torch.randn(L, R, H, generator=g)).to(dev)
    c0 = (0.3 * torch.randn(L, R, H, generator=g)).to(dev)
    params = [getattr(lstm, '{}_l{}'.format(n, l)) for l in range(L) for n in ('weight_ih', 'weight_hh', 'bias_ih', 'bias_hh')]
    if not training:
        with torch.no_grad():
            got = ops.lstm_decoder(x, h0, c0, T, params).transpose(0, 1)
            want = _reference(lstm, x, h0, c0, T)
        assert_close(got, want, what='lstm forward (no grad)')
        return
    x1, x2 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    mine = [p.detach().clone().requires_grad_(True) for p in params]
    got = ops.lstm_decoder(x1, h0, c0, T, mine).transpose(0, 1)
    want = _reference(lstm, x2, h0, c0, T)
    assert got.shape == want.shape == (R, T, H)
    assert_close(got, want, what='lstm forward')
    gout = torch.randn(want.shape, generator=g).to(dev)
    got.backward(gout)
    want.backward(gout)
    torch.cuda.synchronize()

    def l2(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))

    assert l2(x1.grad, x2.grad) <= 1e-3, 'dx {:.2e}'.format(l2(x1.grad, x2.grad))
    for p_mine, p_ref, (name, _) in zip(mine, params, [(n, None) for l in range(L) for n in ('w_ih', 'w_hh', 'b_ih', 'b_hh')]):
        err = l2(p_mine.grad, p_ref.grad)
        assert err <= 1e-3, 'grad {} relative L2 {:.2e}'.format(name, err)


def test_lstm_decoder_module_keeps_the_reference_surface(cuda_device):
    """LSTMDecoderModule: state_dict keys of nn.LSTM, [rows, out_len, out] result, injected states, gradients into lstm.* / lin.*."""
    from garment_pattern_estimation_b200 import net_blocks as nb
    dev = cuda_device
    torch.manual_seed(3)
    dec = nb.LSTMDecoderModule(encoding_size=250, hidden_size=250, out_elem_size=8, n_layers=3, out_len=14).to(dev)
    assert {'lstm.weight_ih_l0', 'lstm.weight_hh_l2', 'lstm.bias_ih_l1', 'lstm.bias_hh_l0', 'lin.weight', 'lin.bias'} <= set(dec.state_dict())
    enc = torch.randn(46, 250, device=dev, requires_grad=True)
    state = (0.01 * torch.randn(3, 46, 250, device=dev), 0.01 * torch.randn(3, 46, 250, device=dev))
    out = dec(enc, 14, lstm_state=state)
    assert out.shape == (46, 14, 8)
    seq, _ = dec.lstm(enc.detach().unsqueeze(1).repeat(1, 14, 1), state)            # cuDNN on the same parameters
    want = dec.lin(seq)
    assert_close(out, want, what='decoder module forward')
    out.sum().backward()
    assert enc.grad is not None and all(p.grad is not None for p in dec.parameters())
    out2 = dec(enc, 14)                                                              # random states drawn on the device
    assert out2.shape == (46, 14, 8) and torch.isfinite(out2).all()
