"""DataBlock::take / take_ranges / scatter / concat on the device against numpy restatements and
the reference's own golden cases (src/query/expression/tests/it/kernel.rs:50-200 with
testdata/kernel-pass.txt, numeric columns transcribed into tests/golden/kernel.json)."""
import json
import os

import numpy as np
import pytest

from databend_b200 import abi
from databend_b200.block import Column, DataBlock
from databend_b200.kernels import concat, scatter, take, take_ranges
from databend_b200.transforms import to_device

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DT = {"I32": abi.I32, "U8": abi.U8, "I64": abi.I64}


def mixed_block(n, seed=0, bit_off=0):
    rng = np.random.default_rng(seed)
    cols = [
        Column.from_data(rng.integers(-2**40, 2**40, n).astype(np.int64)),
        Column.from_data(rng.integers(0, 200, n).astype(np.uint8), validity=rng.random(n) > 0.3, validity_bit_offset=bit_off),
        Column.from_data(rng.normal(size=n).astype(np.float32)),
        Column.from_data(rng.random(n) > 0.5, abi.BOOL, validity=rng.random(n) > 0.1),
        Column.from_data(rng.integers(-30000, 30000, n).astype(np.int16)),
        Column.vector(rng.normal(size=(n, 5)).astype(np.float32)),
        Column.new_const(abi.I32, 7, n),
        Column.new_const(abi.F64, None, n),
    ]
    return DataBlock(cols, n)


def expect_rows(block, rows):
    out = []
    for c in block.columns:
        v = c.values()
        out.append((v[rows] if len(rows) else v[:0], c.valid_mask()[rows] if len(rows) else c.valid_mask()[:0]))
    return out


def assert_block_rows(got: DataBlock, exp):
    assert got.num_columns() == len(exp)
    for c, (v, m) in zip(got.columns, exp):
        assert c.length == len(m)
        np.testing.assert_array_equal(c.valid_mask(), m)
        gv = c.values()
        if gv.dtype.kind == "f":
            np.testing.assert_array_equal(np.asarray(gv)[m].view(np.uint32 if gv.itemsize == 4 else np.uint64),
                                          np.asarray(v)[m].view(np.uint32 if gv.itemsize == 4 else np.uint64))
        else:
            np.testing.assert_array_equal(np.asarray(gv)[m], np.asarray(v)[m])


@pytest.mark.parametrize("n,m", [(1, 0), (5, 3), (1000, 4000), (100_003, 50_000)])
def test_take(gpu, n, m):
    blk = mixed_block(n, seed=n, bit_off=3)
    idx = np.random.default_rng(m).integers(0, n, m)
    out = take(blk, idx)
    assert out.num_rows == m
    assert_block_rows(out, expect_rows(blk, idx))
    assert out.columns[6].is_const and out.columns[7].is_const  # BlockEntry::Const stays const


def test_take_device_resident_and_bad_index(gpu):
    from databend_b200.lib import DbxError
    blk = mixed_block(3000, seed=1)
    dev = DataBlock([to_device(c) if not c.is_const else c for c in blk.columns], blk.num_rows)
    idx = np.random.default_rng(2).integers(0, 3000, 777)
    assert_block_rows(take(dev, idx), expect_rows(blk, idx))
    with pytest.raises(DbxError):
        take(blk, [3000])


def test_take_ranges(gpu):
    blk = mixed_block(50_000, seed=4, bit_off=5)
    ranges = [(10, 20), (0, 1), (49_990, 50_000), (300, 300), (1000, 30_000)]
    rows = np.concatenate([np.arange(a, b) for a, b in ranges])
    out = take_ranges(blk, ranges)
    assert_block_rows(out, expect_rows(blk, rows))
    assert take_ranges(blk, []).num_rows == 0


@pytest.mark.parametrize("n,parts", [(0, 3), (17, 1), (10_000, 7), (200_000, 300)])
def test_scatter(gpu, n, parts):
    blk = mixed_block(max(n, 1), seed=n + parts).slice(0, n)
    idx = np.random.default_rng(parts).integers(0, parts, n)
    outs = scatter(blk, idx, parts)
    assert len(outs) == parts and sum(o.num_rows for o in outs) == n
    for q, o in enumerate(outs):
        assert_block_rows(o, expect_rows(blk, np.nonzero(idx == q)[0]))  # row order preserved inside a target


def test_concat(gpu):
    a, b, c = mixed_block(1000, seed=1, bit_off=1), mixed_block(1, seed=2), mixed_block(70_001, seed=3, bit_off=6)
    out = concat([a, b, c])
    exp = []
    for i in range(a.num_columns()):
        exp.append((np.concatenate([x.columns[i].values() for x in (a, b, c)]), np.concatenate([x.columns[i].valid_mask() for x in (a, b, c)])))
    assert_block_rows(out, exp)
    assert out.columns[6].is_const and out.columns[7].is_const  # the same constant in every block
    # different constants (and const next to a full column) are materialised; NULL constants become NULL rows
    x = DataBlock([Column.new_const(abi.I32, 1, 3), Column.new_const(abi.F64, None, 3), Column.from_data(np.arange(3, dtype=np.int64))])
    y = DataBlock([Column.new_const(abi.I32, 2, 2), Column.from_data(np.array([1.5, 2.5])), Column.new_const(abi.I64, 9, 2)])
    out = concat([x, y])
    assert not out.columns[0].is_const
    np.testing.assert_array_equal(out.columns[0].values(), [1, 1, 1, 2, 2])
    np.testing.assert_array_equal(out.columns[1].valid_mask(), [False, False, False, True, True])
    np.testing.assert_array_equal(out.columns[1].values()[3:], [1.5, 2.5])
    np.testing.assert_array_equal(out.columns[2].values(), [0, 1, 2, 9, 9])


def test_reference_goldens(gpu):
    with open(os.path.join(GOLD, "kernel.json")) as f:
        g = json.load(f)

    def blk(cols):
        return DataBlock([Column.from_data(c["values"], DT[c["dtype"]], validity=c.get("validity")) for c in cols])

    def check(out, result):
        for col, r in zip(out.columns, result):
            m = np.array(r["validity"], dtype=bool)
            np.testing.assert_array_equal(col.valid_mask(), m)
            np.testing.assert_array_equal(np.asarray(col.values())[m], np.array(r["values"])[m])

    for c in (g["take"] if isinstance(g["take"], list) else [g["take"]]):
        check(take(blk(c["columns"]), c["indices"]), c["result"])
    for c in g.get("concat", []):
        check(concat([blk(b) for b in c["blocks"]]), c["result"])
    for c in g.get("scatter", []):
        outs = scatter(blk(c["columns"]), c["indices"], c["scatter_size"])
        for o, r in zip(outs, c["results"]):
            check(o, r)
