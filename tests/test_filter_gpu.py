"""Parity of the standalone filter operator (TransformFilter = FilterExecutor::filter) with the CPU
oracle: bit-exact values and validity, rows in input order.

Mirrors src/query/expression/tests/it/kernel.rs:50-70 (golden, tests/golden/kernel.json) and the
differential fuzz of src/query/service/tests/it/pipelines/filter/filter_executor.rs:18-70."""
import json
import os

import numpy as np
import pytest

from databend_b200 import abi, expr as E
from databend_b200.block import Column, DataBlock
from databend_b200.lib import DbxError
from databend_b200.transforms import TransformFilter, schema_types, to_device

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DT = {"I64": abi.I64, "U64": abi.U64, "F64": abi.F64, "I32": abi.I32, "U8": abi.U8}


def oracle():
    from oracle import oracle as orc
    return orc


def check_against_oracle(blk, pred, device_resident=False, split=None):
    cpred = E.build_predicate(pred)
    exp = oracle().filter_block(blk, cpred)
    op = TransformFilter(pred, schema_types(blk))
    parts = blk.split_by_rows(split) if split else [blk]
    outs = []
    for p in parts:
        if device_resident:
            p = DataBlock([to_device(c) for c in p.columns], p.num_rows)
        outs.append(op.transform(p))
    op.close()
    n_out = sum(o.num_rows for o in outs)
    for ci, c in enumerate(blk.columns):
        if c.dtype == abi.BOOL or c.is_const:
            continue
        vals = np.concatenate([o.columns[ci].values() for o in outs]) if outs else np.empty(0)
        valid = np.concatenate([o.columns[ci].valid_mask() for o in outs]) if outs else np.empty(0, bool)
        ev, em = exp[ci]
        assert len(vals) == len(ev) == n_out
        np.testing.assert_array_equal(valid, em)
        if vals.dtype.kind == "f":
            w = np.uint64 if vals.itemsize == 8 else np.uint32
            np.testing.assert_array_equal(vals.view(w)[em], ev.view(w)[em])
        else:
            np.testing.assert_array_equal(vals[em], ev[em])
    return outs


def test_filter_golden(gpu):
    """kernel.rs:54-68 / kernel-pass.txt:1-18: filter_with_bitmap on an Int32 and a Nullable(UInt8) column."""
    with open(os.path.join(GOLD, "kernel.json")) as f:
        g = json.load(f)["filter"]
    cols = [Column.from_data(g["bitmap"], abi.BOOL)]
    for c in g["columns"]:
        cols.append(Column.from_data(c["values"], DT[c["dtype"]], validity=c.get("validity")))
    blk = DataBlock(cols)
    op = TransformFilter(E.bool_column(0), schema_types(blk))
    out = op.transform(blk)
    op.close()
    assert out.num_rows == sum(g["bitmap"])
    assert list(out.columns[0].values()) == [True] * out.num_rows
    for i, r in enumerate(g["result"]):
        got = out.columns[1 + i]
        assert list(got.valid_mask()) == r["validity"], g["src"]
        for v, e, ok in zip(got.values(), r["values"], r["validity"]):
            if ok:
                assert v == e, g["src"]


@pytest.mark.parametrize("n", [0, 1, 5, 1023, 1024, 1025, 70_001])
def test_filter_modulo_config1_shape(gpu, n):
    """`WHERE number % 3 = 0` (config 1) on a UInt64 column plus payload columns of other widths."""
    rng = np.random.default_rng(n)
    blk = DataBlock([Column.from_data(np.arange(n, dtype=np.uint64)),
                     Column.from_data(rng.integers(-2**31, 2**31, n).astype(np.int32)),
                     Column.from_data(rng.standard_normal(n)),
                     Column.from_data(rng.integers(0, 200, n).astype(np.uint8), validity=rng.random(n) > 0.25 if n else None),
                     Column.from_data(rng.integers(-300, 300, n).astype(np.int16))], n)
    pred = E.eq(E.col(0) % E.lit(3), E.lit(0))
    outs = check_against_oracle(blk, pred)
    assert sum(o.num_rows for o in outs) == (n + 2) // 3


def test_filter_differential_random_predicates(gpu):
    """Random And/Or trees of comparisons over nullable columns, host and device inputs, blocks
    split at non-aligned boundaries (validity bitmaps with bit offsets)."""
    rng = np.random.default_rng(2026)
    n = 20_000
    a = rng.integers(-50, 50, n).astype(np.int64)
    b = rng.integers(0, 1000, n).astype(np.uint64)
    c = rng.standard_normal(n)
    c[rng.random(n) < 0.01] = np.nan
    blk = DataBlock([Column.from_data(a, validity=rng.random(n) > 0.1), Column.from_data(b), Column.from_data(c, validity=rng.random(n) > 0.2),
                     Column.from_data(rng.random(n) > 0.5, abi.BOOL, validity=rng.random(n) > 0.1)], n)
    preds = [
        E.gt(E.col(0), E.lit(3)),
        E.and_(E.ge(E.col(0), E.lit(-10)), E.lt(E.col(1) % E.lit(7), E.lit(3))),
        E.or_(E.lt(E.col(2), E.lit(-0.5)), E.and_(E.bool_column(3), E.ne(E.col(0) % E.lit(5), E.lit(0)))),
        E.or_(E.eq(E.col(1), E.lit(999999)), E.bool_scalar(False)),
        E.le(E.col(2), E.lit(float("nan"))),
        None,
    ]
    for i, pred in enumerate(preds):
        check_against_oracle(blk, pred, device_resident=(i % 2 == 1), split=[None, 3001, 777][i % 3])


def test_filter_const_columns_and_errors(gpu):
    n = 100
    blk = DataBlock([Column.from_data(np.arange(n, dtype=np.int64)), Column.new_const(abi.I32, 7, n), Column.new_const(abi.F64, None, n)], n)
    op = TransformFilter(E.lt(E.col(0), E.lit(10)), schema_types(blk))
    out = op.transform(blk)
    op.close()
    assert out.num_rows == 10
    assert out.columns[1].is_const and out.columns[1].const_value == 7 and out.columns[1].length == 10
    assert out.columns[2].is_const and out.columns[2].const_value is None
    with pytest.raises(DbxError, match="Division by zero"):
        op = TransformFilter(E.eq(E.col(0) % E.lit(0), E.lit(0)), schema_types(blk))
        op.transform(blk)
