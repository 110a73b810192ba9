"""CPU restatement (test infrastructure — never imported by the product) of the optimiser step on the
hot path and of its lazily evaluated form.

The reference steps `torch.optim.Adam(lr, eps=1e-15)` on `_semantic_feature`
(/root/reference/scene/gaussian_model.py:199,208; /root/reference/train.py:222-223).  The product's
kernels (csrc/train_ops.cu, adam_update1) spell that update with explicit fp32 roundings:

    m = fma(1-b1, g, b1*m);  v = fma(1-b2, g*g, b2*v)
    p = fma(-lr/(1-b1^t), m / fma(sqrt(v), 1/sqrt(1-b2^t), eps), p)

`dense_step` applies it to every row; `LazyTable` keeps, per row, the step it is current to and
replays the missed zero-gradient steps when the row is next visited — the schedule of
gags_adam_lazy_rows.  Both use numpy float32 arithmetic (fma emulated in float64, which is exact for
fp32 operands up to the final rounding), so equality between the two SCHEDULES is bit-exact here just
as it is between the two CUDA kernels."""
import numpy as np


def _fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def step_consts(lr, b1, b2, t):
    """(lr / (1 - b1^t), 1 / sqrt(1 - b2^t)) as float32, formed in double like gags_adam_step."""
    return np.float32(lr / (1.0 - b1 ** t)), np.float32(1.0 / np.sqrt(1.0 - b2 ** t))


def update(p, g, m, v, step_size, inv_sqrt_bc2, b1, b2, eps):
    b1f, b2f = np.float32(b1), np.float32(b2)
    omb1, omb2 = np.float32(1.0 - b1), np.float32(1.0 - b2)
    m = _fma(np.full_like(g, omb1), g, (b1f * m).astype(np.float32))
    v = _fma(np.full_like(g, omb2), (g * g).astype(np.float32), (b2f * v).astype(np.float32))
    denom = _fma(np.sqrt(v).astype(np.float32), np.full_like(v, inv_sqrt_bc2), np.full_like(v, np.float32(eps)))
    p = _fma(np.full_like(p, -step_size), (m / denom).astype(np.float32), p)
    return p, m, v


def dense_step(p, g, m, v, t, lr, b1=0.9, b2=0.999, eps=1e-15):
    ss, isb = step_consts(lr, b1, b2, t)
    return update(p, g, m, v, ss, isb, b1, b2, eps)


class LazyTable:
    """[N, D] table updated lazily: `last[r]` = the step row r is current to."""

    def __init__(self, p, b1=0.9, b2=0.999, eps=1e-15):
        self.p = p.copy()
        self.m = np.zeros_like(p)
        self.v = np.zeros_like(p)
        self.last = np.zeros(p.shape[0], dtype=np.int64)
        self.consts = [None]                      # consts[t] for t >= 1
        self.b1, self.b2, self.eps = b1, b2, eps

    def _catch_up_row(self, r, t_to):
        z = np.zeros_like(self.p[r])
        for s in range(int(self.last[r]) + 1, t_to + 1):
            ss, isb = self.consts[s]
            self.p[r], self.m[r], self.v[r] = update(self.p[r], z, self.m[r], self.v[r], ss, isb,
                                                     self.b1, self.b2, self.eps)
        self.last[r] = max(int(self.last[r]), t_to)

    def catch_up(self, rows):
        t = len(self.consts) - 1
        for r in rows:
            self._catch_up_row(r, t)

    def apply(self, grad, rows, lr):
        """optimiser step t = len(consts) on `rows` only (grad rows elsewhere must be zero)."""
        t = len(self.consts)
        self.consts.append(step_consts(lr, self.b1, self.b2, t))
        for r in rows:
            self._catch_up_row(r, t - 1)
            ss, isb = self.consts[t]
            self.p[r], self.m[r], self.v[r] = update(self.p[r], grad[r], self.m[r], self.v[r], ss, isb,
                                                     self.b1, self.b2, self.eps)
            self.last[r] = t

    def flush(self):
        self.catch_up(range(self.p.shape[0]))
