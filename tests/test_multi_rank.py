"""The N>1 path on CPU: world_size-2 gloo.  The data path has no collective (worlds are independent, SURVEY.md
section 8e); what the ranks share is the partition rule and the diagnostics gather.  Each rank builds its block of
an ensemble with avbd-demo3d_b200/scenes.py, steps it with the CPU oracle (no GPU here), and all_gathers the
per-world diagnostics; rank 0 checks the gathered result equals the unsharded run bit for bit."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world_size, port, out):
    import torch
    import torch.distributed as dist
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    from _libs import Oracle
    from avbd_demo3d_b200 import scenes
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    total, steps = 4, 30
    per = total // world_size
    base = scenes.scene("Stack")
    n = len(base["size"])

    def run(first, count):
        ens = scenes.ensemble(base, count, first_world=first)
        rows = []
        for k in range(count):          # the oracle steps one world at a time
            o = Oracle("port").create()
            sl = slice(k * n, (k + 1) * n)
            for i in range(n):
                o.add_body(ens["size"][sl][i], float(ens["density"][sl][i]), float(ens["friction"][sl][i]), ens["pos"][sl][i], ens["quat"][sl][i])
            o.step(steps)
            d = o.diagnostics()
            rows.append([d["maxPen"], d["maxLin"], d["contacts"], d["manifolds"], float(o.state()[:, 1].sum())])
            o.close()
        return torch.tensor(rows, dtype=torch.float64)

    mine = run(rank * per, per)
    gathered = [torch.zeros_like(mine) for _ in range(world_size)]
    dist.all_gather(gathered, mine)
    if rank == 0:
        whole = run(0, total)
        out.put(bool(torch.equal(torch.cat(gathered), whole)) and len({tuple(r.tolist()) for r in whole}) == total)
    dist.barrier()
    dist.destroy_process_group()


def test_block_partition_and_diagnostics_gather_world_size_2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    ok = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
    assert ok
