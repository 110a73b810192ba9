"""The lazily evaluated optimiser schedule (oracle/lazy_adam.py, the CPU restatement of
gags_adam_lazy_rows) against the dense schedule and against torch.optim.Adam — no GPU needed."""
import numpy as np
import torch

from oracle import lazy_adam as LA


def test_lazy_schedule_is_bit_identical_to_dense_schedule():
    rng = np.random.default_rng(0)
    N, D, T = 60, 8, 12
    p0 = rng.standard_normal((N, D)).astype(np.float32)
    pd, md, vd = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    lz = LA.LazyTable(p0)
    for t in range(1, T + 1):
        lr = 1e-2 / t
        touched = np.nonzero((rng.random(N) < 0.2) & (np.arange(N) % 5 != 1))[0]
        g = np.zeros_like(p0)
        g[touched] = rng.standard_normal((len(touched), D)).astype(np.float32)
        pd, md, vd = LA.dense_step(pd, g, md, vd, t, lr)
        if t % 4 == 0:                                   # a view is about to read these rows
            lz.catch_up(np.nonzero(rng.random(N) < 0.3)[0])
        lz.apply(g, touched, lr)
    assert not np.array_equal(lz.p, pd)                  # rows really are behind
    lz.flush()
    assert np.array_equal(lz.p, pd) and np.array_equal(lz.m, md) and np.array_equal(lz.v, vd)
    assert (lz.last == T).all()


def test_restated_update_matches_torch_adam():
    """the explicit-rounding update is torch.optim.Adam(eps=1e-15) to ~1e-6 (different, equally valid
    fp32 roundings of the same formula)."""
    g = torch.Generator().manual_seed(1)
    p0 = torch.randn(40, 6, generator=g)
    pt = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([pt], lr=1e-3, eps=1e-15)
    p, m, v = p0.numpy().copy(), np.zeros((40, 6), np.float32), np.zeros((40, 6), np.float32)
    for t in range(1, 6):
        gr = torch.randn(40, 6, generator=g) * t
        pt.grad = gr.clone()
        opt.step()
        p, m, v = LA.dense_step(p, gr.numpy(), m, v, t, 1e-3)
    ref = pt.detach().numpy()
    assert np.abs(p - ref).max() <= 1e-6 * np.abs(ref).max()
    st = opt.state[pt]
    assert np.abs(m - st["exp_avg"].numpy()).max() <= 1e-6 * np.abs(m).max()
    assert np.abs(v - st["exp_avg_sq"].numpy()).max() <= 1e-6 * np.abs(v).max()
