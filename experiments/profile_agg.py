"""Times the fused filter->hash-agg kernel alone on device-resident synthetic columns.
usage: python experiments/profile_agg.py [rows] [n_keys] [steps]   (env DBX_AGG_DEBUG bisects)"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from databend_b200 import abi, build, lib, expr as E
from databend_b200.block import Column, DataBlock
from databend_b200.transforms import AggregatorParams, DeviceBuffer, TransformFinalAggregate, TransformPartialAggregate

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
n_keys = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
build.build()
L = lib.load()
lib.require_device()
bufs = [DeviceBuffer(rows * 8) for _ in range(3)]
lib.check(L.dbx_synth_fill(0, 0, 42, n_keys, 0, rows, bufs[0].ptr))
lib.check(L.dbx_synth_fill(0, 1, 43, 0, 0, rows, bufs[1].ptr))
lib.check(L.dbx_synth_fill(0, 2, 44, 20, 0, rows, bufs[2].ptr))
blk = DataBlock([Column.device(abi.I64, rows, bufs[0].ptr), Column.device(abi.I64, rows, bufs[1].ptr),
                 Column.device(abi.F64, rows, bufs[2].ptr)], rows)
params = AggregatorParams([0], [("sum", 1), ("count", 1), ("avg", 2)])
filt = E.eq(E.col(1) % E.lit(3), E.lit(0))
types = [abi.I64, abi.I64, abi.F64]
part = TransformPartialAggregate(params, types, filt)
fin = TransformFinalAggregate(params, types)
for i in range(steps):
    part.reset(); fin.reset()
    part.transform(blk)
    ms = part.last_kernel_ms()
    fin.transform(part.on_finish())
    out = fin.on_finish(abi.MEM_DEVICE)
    g = out[0].num_rows
    L.dbx_block_release(C.byref(out[0]))
    print(f"step {i}: rows {rows} keys {n_keys} groups {g} kernel {ms:.3f} ms  {rows/ms/1e6:.2f} Grows/s  {24*rows/ms/1e6:.1f} GB/s  dbg={os.environ.get('DBX_AGG_DEBUG','0')}", flush=True)
