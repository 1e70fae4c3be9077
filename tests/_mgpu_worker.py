"""Worker for tests/test_gpu_multi.py: one rank per GPU, NCCL inside librl_b200 (dlopen), torchrun launch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import load_cbox  # noqa: E402
from rustlight_b200 import _abi  # noqa: E402
from rustlight_b200.device import Context, DeviceScene, nccl_unique_id  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ids = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx = Context(local, nranks=world, rank=rank, nccl_id=ids[0])
    sc = load_cbox(400, 304)
    dev = DeviceScene(ctx, sc)
    for integ in (_abi.path_desc(), _abi.direct_desc(1, 1)):
        img, st = dev.render(integ, 8, seed=3)  # one ncclReduce inside; full frame lands on rank 0
        tot = torch.tensor([st.samples, st.segments], dtype=torch.int64, device="cuda")
        dist.all_reduce(tot)
        if rank == 0:
            solo = Context(local)
            ref, sr = DeviceScene(solo, sc).render(integ, 8, seed=3)
            assert np.array_equal(img, ref), "reduced frame != single-GPU frame"
            assert tot.tolist() == [sr.samples, sr.segments]
            assert st.ms_reduce > 0
            solo.close()
    dist.barrier()
    if rank == 0:
        print("MGPU_OK", world)
    dev.close()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
