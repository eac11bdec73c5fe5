"""One rank of test_host_logic.py::test_rank_parts_merge_in_rank_order (gloo, launched by torchrun)."""
import os
import sys

import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepsignal_plant_b200 import call_modifications as cm  # noqa: E402

if __name__ == "__main__":
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    result = sys.argv[1]
    with open("%s.part%05d" % (result, rank), "wb") as f:          # parts of very different sizes, one of them empty
        f.write((b"rank %d line\n" % rank) * (0 if rank == 1 else 1000 * (rank + 1) + 7))
    cm._merge_parts(result, rank, world)
    dist.destroy_process_group()
