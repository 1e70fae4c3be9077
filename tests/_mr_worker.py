"""Worker for tests/test_multirank.py (launched by torch.distributed.run, gloo backend, CPU only).

Exercises the N>1 host logic without a GPU: every rank renders ITS image tiles (same 16x16 tile
ownership rule as librl_b200: tile (tx,ty) -> rank (tx+ty) % nranks) with the CPU emulator of the
device arithmetic, the partial frames are summed with ONE all-reduce (the stand-in for the
library's single ncclReduce), and rank 0 checks the result against the full single-rank frame.
Also exercises the unique-id broadcast pattern bench.py uses."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import emu_binding as eb  # noqa: E402
from conftest import load_cbox  # noqa: E402
from oracle import binding as ob  # noqa: E402
from rustlight_b200 import _abi  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    ids = [os.urandom(128) if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    assert isinstance(ids[0], bytes) and len(ids[0]) == 128
    sc = load_cbox(80, 48)
    for integ in (_abi.path_desc(), _abi.direct_desc(1, 1)):
        part, st = eb.EmuScene(sc).render(integ, 3, seed=7, rank=rank, nranks=world)
        opart, ost = ob.OracleScene(sc).render(integ, 3, seed=7, cfg=ob.config(
            estimator=ob.EST_STREAM, accel_mode=ob.ACCEL_BVH, rank=rank, nranks=world))
        assert np.array_equal(part, opart) and st.segments == ost.segments
        t = torch.from_numpy(part.copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        counts = torch.tensor([st.samples, st.segments], dtype=torch.int64)
        dist.all_reduce(counts)
        owned = torch.from_numpy((part != 0).any(axis=2).astype(np.int32))
        dist.all_reduce(owned)
        if rank == 0:
            full, sf = eb.EmuScene(sc).render(integ, 3, seed=7)
            assert np.array_equal(t.numpy(), full), "sum of rank frames != single-rank frame"
            assert counts.tolist() == [sf.samples, sf.segments]
            assert int(owned.max()) <= 1, "two ranks wrote the same pixel"
    dist.barrier()
    if rank == 0:
        print("MULTIRANK_OK", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
