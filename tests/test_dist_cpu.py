"""World-size-2 gloo test of the sharding + end-of-batch gather logic (no GPU)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hover_net_b200.dist import compact_rows, gather_tables, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_tiles, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_tiles, rank, world)
    b = n_tiles // world  # equal per-rank batch (weak scaling, as bench.py)
    rng = np.random.default_rng(100 + rank)
    table = torch.zeros((b, 7, 10), dtype=torch.int64)
    nrows = torch.zeros((b,), dtype=torch.int32)
    for i in range(b):
        k = int(rng.integers(0, 8))
        nrows[i] = k
        table[i, :k] = torch.from_numpy(rng.integers(0, 1000, (k, 10))) + 10000 * (rank * b + i)
    t_all, n_all = gather_tables(table, nrows)
    if rank == 0:
        rows = compact_rows(t_all, n_all)
        out.put(([r.tolist() for r in rows], n_all.tolist(), (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 65):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_gather_tables_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 8, q)) for r in range(2)]
    for p in procs:
        p.start()
    rows, n_all, span = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert span == (0, 4) and len(rows) == 8 and len(n_all) == 8
    for i, r in enumerate(rows):
        assert len(r) == n_all[i]
        assert all(10000 * i <= v < 10000 * i + 1000 for row in r for v in row), "rows must stay in global tile order"


def _tile_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hover_net_b200.infer import tile
    img = np.random.default_rng(5).integers(0, 256, (500, 333, 3), dtype=np.uint8)
    padded, pinfo, _ = tile._prepare_patching(img, 256, 164, True)

    def fake_step(batch):  # a per-patch function of the patch content only
        b = batch.astype(np.float32)
        return np.stack([b[:, 46:210, 46:210, 0], b[:, 46:210, 46:210, 1], b[:, 46:210, 46:210, 2],
                         b[:, 46:210, 46:210].sum(-1)], -1)

    outs = tile.run_patches(padded, pinfo, 256, fake_step, 5)
    m = tile._stitch(pinfo, outs, img.shape)
    if rank == 0:
        out.put(m)
    dist.barrier()
    dist.destroy_process_group()


def test_tile_patches_sharded_gloo_world3_equals_single():
    """config[3] plumbing: patches of one image sharded over 3 ranks, gathered, stitched == 1-rank result
    (here the valid-conv crop of the padded image must reproduce the source image exactly)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_tile_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    m = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    img = np.random.default_rng(5).integers(0, 256, (500, 333, 3), dtype=np.uint8).astype(np.float32)
    assert m.shape == (500, 333, 4)
    assert np.array_equal(m[..., :3], img) and np.array_equal(m[..., 3], img.sum(-1))
