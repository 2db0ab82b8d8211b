"""Worker of tests/test_gpu_nccl.py: one process per GPU under torchrun (NCCL).  Every rank runs
runner.run_sharded on the same 256 DISTINCT frames (rank r takes i % W == r), the records are gathered over
NCCL, and rank 0 compares them with a single-rank run of all frames on its own GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import litepi_b200
    from litepi_b200 import runner, synth
    from helpers import model_paths
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_frames = int(os.environ.get("LP_NCCL_FRAMES", "256"))
    param, binp = model_paths("vntsr")
    pipe = litepi_b200.B200Pipeline(param, binp, None, "shufflenetv2", num_classes=49, device=local, max_batch=32, seed=0)
    frames = [synth.vn_frame(i) for i in range(n_frames)]
    got = runner.run_sharded(pipe, frames, 0.25, 0.45, 50, rank, world)           # gathers over NCCL
    cnt = torch.tensor([int((got[:, 0] % world == rank).sum())], dtype=torch.int64, device=pipe.device)
    dist.all_reduce(cnt)
    assert int(cnt) == got.shape[0], f"rank {rank}: gathered {got.shape[0]} records, ranks hold {int(cnt)}"
    if rank == 0:
        # the same frames on ONE rank, no collective (the process group stays initialised: world-1 semantics via rank/world args)
        local_only = []
        B = pipe.max_batch
        from litepi_b200.detector import FrameBatch
        for i in range(0, n_frames, B):
            ids = list(range(i, min(i + B, n_frames)))
            fb = FrameBatch.from_host([frames[j] for j in ids], pipe.device)
            n = pipe.run_device(fb, 0.25, 0.45, 50, torch.tensor(ids, dtype=torch.int32, device=pipe.device))
            local_only.append(pipe.records[:n].cpu().numpy().copy())
        want = runner.sort_records(np.concatenate(local_only))
        assert want.shape == got.shape and np.array_equal(want, got), "sharded + NCCL gather differs from the 1-rank run"
        assert got.shape[0] > n_frames                                              # several detections per frame
        print(f"NCCL_OK world={world} frames={n_frames} records={got.shape[0]}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
