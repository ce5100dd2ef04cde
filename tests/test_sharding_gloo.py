"""N>1 host logic on CPU: two gloo ranks derive the shard layout independently (no collective on
the data path), fill their slices and gather them at the XX writer."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import INPUTS, example_zmat


def _worker(rank, world, port, tmp, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import myqc_b200 as Q
    s = Q.make_job(os.path.join(tmp, f"r{rank}"), example_zmat("h2o_4"), INPUTS)
    off = Q.shard_layout(s, world)
    # every rank computed the same layout without talking to the others
    t = torch.from_numpy(off.copy())
    gathered = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    same = all(torch.equal(g, t) for g in gathered)
    # host gather to the writer (rank 0): each rank contributes exactly its slice
    mine = np.full(int(off[rank + 1] - off[rank]), float(rank + 1))
    full = None
    if rank == 0:
        full = np.zeros(s.nunique)
        full[off[0]:off[1]] = mine
        for r in range(1, world):
            buf = torch.zeros(int(off[r + 1] - off[r]), dtype=torch.float64)
            dist.recv(buf, src=r)
            full[off[r]:off[r + 1]] = buf.numpy()
    else:
        dist.send(torch.from_numpy(mine), dst=0)
    dist.barrier()
    if rank == 0:
        covered = bool(np.all(full > 0)) and float(full.sum()) == float(sum((r + 1) * (off[r + 1] - off[r]) for r in range(world)))
        q.put((same, covered, off.tolist(), s.nunique))
    dist.destroy_process_group()


def test_two_ranks_agree_on_layout_and_tile_the_array(tmp_path):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    same, covered, off, nunique = q.get(timeout=10)
    assert same and covered
    assert off[0] == 0 and off[-1] == nunique and 0 < off[1] < nunique
