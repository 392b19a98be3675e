"""torchrun --nproc-per-node N scripts/peer_selftest.py : collective self test of the peer-memory link (csrc/peer.h)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openifem_b200 as ifem

rank, world = ifem.init_distributed()
bad = ifem.peer_selftest(500)
print(f"[rank {rank}/{world}] peer self test: {bad} mismatches (-1 = link inactive)", flush=True)
if world > 1:
    import torch.distributed as dist

    dist.barrier()
    ifem.comm_finalize()
    dist.destroy_process_group()
