"""DataBlock kernels on the device: take / take_ranges / scatter / concat.

Host-side mirror of the reference's block kernels
(src/query/expression/src/kernels/take.rs:43-60, take_ranges.rs:40, scatter.rs:21, concat.rs:62);
each forwards to one libdbx entry point (include/dbx.h: dbx_block_*).  No compute happens here."""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import numpy as np

from . import abi
from .block import DataBlock
from .lib import check, load
from .transforms import _block_from_c


def take(block: DataBlock, indices: Sequence[int], device: int = 0) -> DataBlock:
    """DataBlock::take(indices): out row i = block row indices[i]."""
    idx = np.ascontiguousarray(indices, dtype=np.uint32)
    b, keep = block.as_c()
    out = abi.Block()
    check(load().dbx_block_take(device, C.byref(b), idx.ctypes.data, len(idx), abi.MEM_HOST, abi.MEM_HOST, C.byref(out)))
    return _block_from_c(out, device)


def take_ranges(block: DataBlock, ranges: Sequence[Sequence[int]], device: int = 0) -> DataBlock:
    """DataBlock::take_ranges(ranges): the rows of every [start, end) range, range after range."""
    starts = np.ascontiguousarray([r[0] for r in ranges], dtype=np.uint32)
    lens = np.ascontiguousarray([r[1] - r[0] for r in ranges], dtype=np.uint32)
    b, keep = block.as_c()
    out = abi.Block()
    check(load().dbx_block_take_ranges(device, C.byref(b), starts.ctypes.data, lens.ctypes.data, len(starts), abi.MEM_HOST, C.byref(out)))
    return _block_from_c(out, device)


def scatter(block: DataBlock, indices: Sequence[int], scatter_size: int, device: int = 0) -> List[DataBlock]:
    """DataBlock::scatter(indices, scatter_size): row i goes to output indices[i]; row order is kept."""
    idx = np.ascontiguousarray(indices, dtype=np.uint32)
    assert len(idx) == block.num_rows
    b, keep = block.as_c()
    outs = (abi.Block * scatter_size)()
    check(load().dbx_block_scatter(device, C.byref(b), idx.ctypes.data, abi.MEM_HOST, scatter_size, abi.MEM_HOST, outs))
    return [_block_from_c(outs[i], device) for i in range(scatter_size)]


def concat(blocks: Sequence[DataBlock], device: int = 0) -> DataBlock:
    """DataBlock::concat(blocks)."""
    arr = (abi.Block * len(blocks))()
    keep = []
    for i, blk in enumerate(blocks):
        b, k = blk.as_c()
        arr[i] = b
        keep.append(k)
    out = abi.Block()
    check(load().dbx_block_concat(device, arr, len(blocks), abi.MEM_HOST, C.byref(out)))
    return _block_from_c(out, device)
