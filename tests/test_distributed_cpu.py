"""Host-side logic of the multi-GPU post-correction on CPU: world_size 2, gloo.
The per-shard compute is stubbed (no GPU here); sharding + the single gather are real."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _stub_compute(model, mcmc_output, particles, is_type="is2", seed=1, draw_range=None):
    b, e = draw_range
    d = np.arange(b, e, dtype=np.float64)
    return {"weights": np.exp(-0.01 * d) + seed, "posterior": -0.5 * d,
            "alpha": np.stack([d, 2 * d])[None, :, :].repeat(3, axis=0)}  # (3, 2, nd): last axis = draws


def _worker(rank, world, port, n_stored, out_q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from bssm_b200 import distributed as D
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    mo = {"counts": np.ones(n_stored, dtype=np.uint32), "output_type": 1}
    res = D.post_correct_distributed(None, mo, 16, is_type="is2", seed=4, compute=_stub_compute)
    if rank == 0:
        out_q.put((res["weights"], res["posterior"], res["alpha"]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    from bssm_b200.distributed import shard_range
    for n in (1, 7, 15000, 15001):
        for world in (1, 2, 3, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            sizes = [e - b for b, e in edges]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gather_equals_single_process():
    n_stored, world = 37, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_stored, q)) for r in range(world)]
    for p in procs:
        p.start()
    w, post, alpha = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ref = _stub_compute(None, None, 16, seed=4, draw_range=(0, n_stored))
    assert np.array_equal(w, ref["weights"])
    assert np.array_equal(post, ref["posterior"])
    assert np.array_equal(alpha, ref["alpha"])
