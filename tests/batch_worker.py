"""Run under torchrun (one rank per GPU) by tests/test_zz_gpu_batch.py: a batch of files dealt over the ranks,
tables all-gathered over NCCL, every rank checks the full table against the oracle."""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from batch_common import FILTER, TIMESTEP, assert_tables_match, detector, make_files, oracle_tables, segmenter  # noqa: E402
from pypore_b200.batch import FileBatch, assign_files  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 11
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    files = make_files(n_files)
    owned = set(assign_files(n_files, rank, world))
    mine = [f if i in owned else None for i, f in enumerate(files)]
    b = FileBatch(device=local, workers=2, rank=rank, world=world)
    for _ in range(2):
        got = b.parse(mine, TIMESTEP, detector(), segmenter(), FILTER)
    assert sorted(b.local) == sorted(owned)
    assert_tables_match(got, oracle_tables(files))
    dist.barrier()
    if rank == 0:
        print("BATCH OK world=%d files=%d events=%d segments=%d" % (world, n_files, got.n_events, got.n_segments),
              flush=True)
    b.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
