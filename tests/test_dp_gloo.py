"""View-parallel plumbing on CPU: world_size 2 over gloo (the kernels themselves need a GPU; here
the gradient comes from the oracle so that the sharding + all-reduce logic is what is tested)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _grad_for_view(v: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(100 + v)
    return torch.randn(64, 8, generator=g)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from gags_b200 import parallel
    r, w, _ = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    p = torch.nn.Parameter(torch.zeros(64, 8))
    mine = parallel.views_for_rank(3, rank, world, 2, 16)
    p.grad = sum(_grad_for_view(v) for v in mine)
    parallel.allreduce_grads([p], world)
    # the pipelined variant, on a host-side optimiser exposing step_chunks (the product's FusedAdam
    # needs a GPU): slices are reduced asynchronously and handed to the optimiser one by one
    q = torch.nn.Parameter(torch.zeros(64, 8))
    q.grad = sum(_grad_for_view(v) for v in mine)

    class _Sgd:
        def __init__(self):
            self.seen = []

        def step_chunks(self, prm, ready):
            n = ready(None, None)
            per = -(-prm.numel() // n)
            per += (-per) % 4
            for i in range(n):
                ready(i, n)
                self.seen.append(i)
                sl = slice(i * per, min(prm.numel(), (i + 1) * per))
                prm.data.view(-1)[sl] -= prm.grad.view(-1)[sl]

    opt = _Sgd()
    parallel.allreduce_and_step(opt, q, world, chunks=4)
    mx = parallel.max_over_ranks(float(rank + 1), world, "cpu")
    sm = parallel.sum_over_ranks(1.0, world, "cpu")
    if rank == 0:
        torch.save({"grad": p.grad, "mine": mine, "max": mx, "sum": sm, "q": q.data,
                    "seen": opt.seen}, out)
    dist.destroy_process_group()


def test_view_sharding_is_disjoint_and_covers():
    from gags_b200 import parallel
    seen = []
    for step in range(2):
        for r in range(4):
            seen += parallel.views_for_rank(step, r, 4, 2, 64)
    assert sorted(seen) == list(range(16))


def test_allreduce_equals_single_process_sum(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    # single process accumulating the same G*k views (App. C-14)
    expect = sum(_grad_for_view(v) for v in [12, 13, 14, 15])
    assert res["mine"] == [12, 13]
    assert torch.allclose(res["grad"], expect, atol=1e-6)
    assert res["max"] == 2.0 and res["sum"] == 2.0
    # pipelined all-reduce + step: every slice reduced before it was consumed
    assert res["seen"] == [0, 1, 2, 3]
    assert torch.allclose(res["q"], -expect, atol=1e-6)


def test_peer_slices_cover_the_table_in_float4_units():
    """ownership split of parallel.PeerAdam: equal, 16-byte aligned, contiguous, complete."""
    from gags_b200.parallel import peer_slices
    for numel in (0, 1, 7, 640448, 2_000_000 * 256, 10007 * 64):
        for world in (1, 2, 3, 4, 8, 16):
            padded, per = peer_slices(numel, world)
            assert per * world == padded >= numel and padded - numel < 4 * world
            assert per % 4 == 0
            starts = [r * per for r in range(world)]
            assert all(s % 4 == 0 for s in starts)
            assert starts[-1] + per == padded
    import pytest
    with pytest.raises(ValueError):
        peer_slices(10, 0)


def test_views_for_rank_one_view_per_rank_is_a_partition():
    """bench.py's default at every N: step s, rank r renders view (s*G + r) mod n_views."""
    from gags_b200.parallel import views_for_rank
    n_views = 64
    for world in (1, 2, 4, 8):
        seen = []
        for step in range(n_views // world):
            for r in range(world):
                v = views_for_rank(step, r, world, 1, n_views)
                assert v == [(step * world + r) % n_views]
                seen += v
        assert sorted(seen) == list(range(n_views))


def test_row_word_ranges_partition_the_flag_words():
    """ownership split of parallel.SparsePeerAdam / gags_grad_allreduce_rows: contiguous word ranges
    that cover [0, ceil(rows / 4)) exactly once (the C side computes the same split)."""
    from gags_b200.parallel import row_word_ranges
    for rows in (0, 1, 3, 4, 5, 10007, 2_000_000, 4099):
        for world in (1, 2, 3, 4, 8, 16):
            words, rng = row_word_ranges(rows, world)
            assert words == (rows + 3) // 4 and len(rng) == world
            assert rng[0][0] == 0 and rng[-1][1] == words
            for (a0, a1), (b0, b1) in zip(rng[:-1], rng[1:]):
                assert a0 <= a1 == b0 <= b1
    import pytest
    with pytest.raises(ValueError):
        row_word_ranges(4, 0)
