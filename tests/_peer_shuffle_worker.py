"""Worker of tests/test_peer_shuffle_procs_gpu.py: ONE rank of the partitioned hash join with the
fused peer-memory shuffle (real CUDA-IPC mapping between processes; handles travel over gloo)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    out_dir = sys.argv[1]
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from databend_b200 import abi, lib
    from databend_b200.block import Column, DataBlock
    from databend_b200.distributed import partitioned_hash_join_peer
    from databend_b200.exchange import owner_of
    from databend_b200.transforms import to_device
    n_dev = lib.require_device()
    dev = rank % n_dev
    rng = np.random.default_rng(1234)  # every rank generates the same tables and takes its row range
    n_dim, n_fact = 50_000, 400_000
    dk = rng.permutation(n_dim).astype(np.int64) * 3 - 7000
    dv = rng.integers(-2**40, 2**40, n_dim).astype(np.int64)
    fk = dk[rng.integers(0, n_dim, n_fact)].copy()
    fk[::50] = 10**12  # keys without a match
    fv = rng.integers(0, 2**31, n_fact).astype(np.int32)
    b_lo, b_hi = n_dim * rank // world, n_dim * (rank + 1) // world
    p_lo, p_hi = n_fact * rank // world, n_fact * (rank + 1) // world
    build = DataBlock([to_device(Column.from_data(dk[b_lo:b_hi]), dev), to_device(Column.from_data(dv[b_lo:b_hi]), dev)], b_hi - b_lo)
    probe = DataBlock([to_device(Column.from_data(fk[p_lo:p_hi]), dev), to_device(Column.from_data(fv[p_lo:p_hi]), dev)], p_hi - p_lo)
    stats = {}
    # small rounds: several send/recv rounds per side, so both parities of the regions are reused
    outs, j, shufs = partitioned_hash_join_peer(build, probe, 0, 0, dev, rank, world, round_rows=40_000, stats=stats)
    cols = [np.concatenate([o.columns[i].values() for o in outs]) if outs else np.empty(0, dtype=np.int64) for i in range(4)]
    # every joined row must sit on the owner of its key
    assert (owner_of(cols[0].view(np.uint64), np.zeros(len(cols[0]), np.int64), world) == rank).all()
    np.savez(os.path.join(out_dir, f"join_r{rank}.npz"), fk=cols[0], fv=cols[1], dk=cols[2], dv=cols[3])
    dist.barrier()
    for s in shufs:
        s.close()
    j.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
