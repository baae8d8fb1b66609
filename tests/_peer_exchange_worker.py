"""Worker of tests/test_peer_exchange_procs_gpu.py: ONE rank of the real multi-process
peer-memory exchange (CUDA-IPC mapping of the peers' receive buffers, release/acquire flags at
system scope, epoch parity).  Launched once per rank; handles travel over gloo."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    out_dir = sys.argv[1]
    grouped = sys.argv[2] == "grouped"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from databend_b200 import abi, expr as E, lib
    from databend_b200.block import Column, DataBlock
    from databend_b200.exchange import PeerExchange
    from databend_b200.transforms import AggregatorParams, TransformFinalAggregate, TransformPartialAggregate, to_device
    from oracle import oracle as orc
    n_dev = lib.require_device()
    dev = rank % n_dev  # fewer GPUs than ranks: the ranks share a device (IPC works within one GPU too)
    shared = n_dev < world
    filt = E.eq(E.col(1) % E.lit(3), E.lit(0))
    types = [abi.I64, abi.I64, abi.F64]
    if not grouped:
        # no GROUP BY: per-rank single states, gathered and merged in rank order (dbx_agg_final_merge_rows)
        from databend_b200.distributed import allreduce_single_state
        params = AggregatorParams([], [("sum", 1), ("count", 1), ("avg", 2), ("min", 1), ("max", 2)])
        rows = 300_000
        k = orc.synth_fill(0, 7, 1000, 0, rows)
        v = orc.synth_fill(1, 8, 0, 0, rows)
        x = orc.synth_fill(3, 9, 0, 0, rows)  # non-integer doubles: the merge order matters, and is fixed
        lo, hi = rows * rank // world, rows * (rank + 1) // world
        part = TransformPartialAggregate(params, types, filt, dev)
        fin = TransformFinalAggregate(params, types, dev)
        part.transform(DataBlock([Column.from_data(k[lo:hi]), Column.from_data(v[lo:hi]), Column.from_data(x[lo:hi])]))
        out = allreduce_single_state(part, fin, dev)
        np.savez(os.path.join(out_dir, f"single_r{rank}.npz"), **{f"c{i}": out.columns[i].values() for i in range(5)},
                 **{f"v{i}": out.columns[i].valid_mask() for i in range(5)})
        part.close(); fin.close()
        dist.barrier()
        dist.destroy_process_group()
        return
    params = AggregatorParams([0], [("sum", 1), ("count", 1), ("avg", 2)])
    part = TransformPartialAggregate(params, types, filt, dev)
    fin = TransformFinalAggregate(params, types, dev)
    x = PeerExchange(part, rank, world)
    x.connect()  # collective: all-gather of the IPC handles + cudaIpcOpenMemHandle
    for epoch, (rows, keys) in enumerate([(400_000, 30_000), (250_000, 90_000), (123_457, 1_000), (300_000, 50_000)]):
        k = orc.synth_fill(0, 100 + epoch, keys, 0, rows)
        v = orc.synth_fill(1, 101 + epoch, 0, 0, rows)
        xs = orc.synth_fill(2, 102 + epoch, 20, 0, rows)
        k[:3] = -(2**63)  # the key equal to the table's EMPTY sentinel travels through the exchange too
        lo, hi = rows * rank // world, rows * (rank + 1) // world
        blk = DataBlock([to_device(Column.from_data(k[lo:hi]), dev), to_device(Column.from_data(v[lo:hi]), dev),
                         to_device(Column.from_data(xs[lo:hi]), dev)], hi - lo)
        part.reset()
        fin.reset()
        part.transform(blk)
        part.on_finish()
        x.scatter(part)
        if shared:  # ranks time-slice ONE GPU: a merge must not spin while a peer's scatter still waits for the device
            part.synchronize()
            dist.barrier()
        x.merge(fin)
        out = fin.on_finish()[0]
        ph = x.phase_ms()
        assert ph["scatter"] >= 0 and ph["merge"] >= 0
        np.savez(os.path.join(out_dir, f"e{epoch}_r{rank}.npz"), **{f"c{i}": out.columns[i].values() for i in range(4)},
                 **{f"v{i}": out.columns[i].valid_mask() for i in range(4)})
    dist.barrier()
    x.close()
    part.close(); fin.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
