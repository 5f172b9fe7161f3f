"""ORACLE (test infrastructure): the LSTM decoder of the reference (nn/net_blocks.py:363-402 -- ``nn.LSTM(batch_first=True)``
on an encoding repeated ``out_len`` times) restated as the sequence of row GEMMs and element-wise cell updates that a custom
(non-cuDNN) kernel implementation has to perform, forward AND backward, with the buffers such kernels would keep.

Purpose: SURVEY.md section 8 row a10 is the one library call left on the B200 path (cuDNN); this file is the checked
specification for its replacement.  ``tests/test_oracle.py::test_decomposed_lstm_*`` compare it with
``torch.nn.LSTM`` + autograd on CPU.  Nothing in the product imports it.

Data flow (R rows = B * 23 sequences, T steps, H hidden, gates in PyTorch order i, f, g, o):

  forward, layer l
    Gx    = X_l . W_ih^T + (b_ih + b_hh)         one GEMM over all T*R rows; for l = 0 the input is the SAME vector at every step
                                                 (the reference repeats the encoding), so Gx0 is [R, 4H], computed once
    t = 0..T-1:   G = Gx[t] + h_{t-1} . W_hh^T   recurrent GEMM [R, H] x [H, 4H]
                  i, f, o = sigmoid, g = tanh;  c_t = f c_{t-1} + i g;  h_t = o tanh(c_t)       (cell kernel)
    kept for the backward: the activated gates [T, R, 4H], c_t [T, R, H], h_t [T+1, R, H] (slot 0 = h0)

  backward, layer l (from the top layer down)
    t = T-1..0:   dh = dH_l[t] + dh_rec;  do = dh tanh(c_t);  dc = dc_rec + dh o (1 - tanh(c_t)^2)
                  di = dc g;  dg = dc i;  df = dc c_{t-1};  dc_rec = dc f
                  dG[t] = [di i(1-i), df f(1-f), dg (1-g^2), do o(1-o)]                          (cell kernel)
                  dh_rec = dG[t] . W_hh                                                          recurrent GEMM [R, 4H] x [4H, H]
    dW_hh = dG^T . H_prev     dW_ih = dG^T . X_l     db_ih = db_hh = column sums of dG           time-batched GEMMs / reductions
    dX_l  = dG . W_ih         (= dH_{l-1}; for l = 0 only sum_t dG[t] is needed: dx = (sum_t dG[t]) . W_ih, dW_ih = (sum_t dG[t])^T . x)
"""
import torch


def _cell_forward(G, c_prev):
    H = c_prev.shape[1]
    i, f, g, o = torch.sigmoid(G[:, :H]), torch.sigmoid(G[:, H:2 * H]), torch.tanh(G[:, 2 * H:3 * H]), torch.sigmoid(G[:, 3 * H:])
    c = f * c_prev + i * g
    h = o * torch.tanh(c)
    return h, c, torch.cat([i, f, g, o], dim=1)


def _cell_backward(dh, dc_rec, act, c, c_prev):
    H = c.shape[1]
    i, f, g, o = act[:, :H], act[:, H:2 * H], act[:, 2 * H:3 * H], act[:, 3 * H:]
    tc = torch.tanh(c)
    do = dh * tc
    dc = dc_rec + dh * o * (1 - tc * tc)
    dG = torch.cat([dc * g * i * (1 - i), dc * c_prev * f * (1 - f), dc * i * (1 - g * g), do * o * (1 - o)], dim=1)
    return dG, dc * f


def lstm_forward(x, weights, h0, c0, T):
    """x: [R, I] (the encoding, repeated T times by the reference).  weights: list over layers of (W_ih, W_hh, b_ih, b_hh).
    h0, c0: [L, R, H].  Returns (out [T, R, H] time-major, saved)."""
    saved = []
    X = None                                     # [T, R, H] input of layers >= 1
    for l, (W_ih, W_hh, b_ih, b_hh) in enumerate(weights):
        R, H = h0.shape[1], h0.shape[2]
        bias = b_ih + b_hh
        if l == 0:
            Gx = (x @ W_ih.t() + bias).unsqueeze(0).expand(T, R, 4 * H)          # computed once, read T times
        else:
            Gx = (X.reshape(T * R, -1) @ W_ih.t() + bias).view(T, R, 4 * H)
        Hs = torch.empty(T + 1, R, H, dtype=x.dtype)
        Cs = torch.empty(T + 1, R, H, dtype=x.dtype)
        act = torch.empty(T, R, 4 * H, dtype=x.dtype)
        Hs[0], Cs[0] = h0[l], c0[l]
        for t in range(T):
            G = Gx[t] + Hs[t] @ W_hh.t()
            Hs[t + 1], Cs[t + 1], act[t] = _cell_forward(G, Cs[t])
        saved.append((Hs, Cs, act))
        X = Hs[1:]
    return X, saved


def lstm_backward(dout, x, weights, saved, T):
    """dout: [T, R, H] gradient of the top layer's outputs.  Returns (dx [R, I], grads) with grads = list over layers of
    (dW_ih, dW_hh, db_ih, db_hh)."""
    L = len(weights)
    grads = [None] * L
    dH = dout
    dx = None
    for l in range(L - 1, -1, -1):
        W_ih, W_hh, _, _ = weights[l]
        Hs, Cs, act = saved[l]
        R, H = Hs.shape[1], Hs.shape[2]
        dG = torch.empty(T, R, 4 * H, dtype=dout.dtype)
        dh_rec = torch.zeros(R, H, dtype=dout.dtype)
        dc_rec = torch.zeros(R, H, dtype=dout.dtype)
        for t in range(T - 1, -1, -1):
            dG[t], dc_rec = _cell_backward(dH[t] + dh_rec, dc_rec, act[t], Cs[t + 1], Cs[t])
            dh_rec = dG[t] @ W_hh
        flat = dG.reshape(T * R, 4 * H)
        dW_hh = flat.t() @ Hs[:T].reshape(T * R, H)
        db = flat.sum(0)
        if l == 0:
            sumG = dG.sum(0)                                                     # the input is the same at every step
            dW_ih = sumG.t() @ x
            dx = sumG @ W_ih
        else:
            X = saved[l - 1][0][1:].reshape(T * R, -1)
            dW_ih = flat.t() @ X
            dH = (flat @ W_ih).view(T, R, -1)
        grads[l] = (dW_ih, dW_hh, db, db.clone())
    return dx, grads
