"""Parity of the streaming device top-k against the CPU oracle (ORDER BY key LIMIT k).

Mirrors src/query/expression/tests/it/sort.rs:29-100 (golden) and the OrderedFloat total order
(ordered_float.rs:147-201).  Bit-exact: the returned row ids must equal the oracle's, with ties
broken by ascending row id on both sides."""
import json
import os

import numpy as np
import pytest

from databend_b200 import abi
from databend_b200.block import Column, DataBlock
from databend_b200.transforms import TransformTopN, schema_types, to_device

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def oracle():
    from oracle import oracle as orc
    return orc


def run_topk(col, asc, nulls_first, k, split=None, device_resident=False):
    blk = DataBlock([col])
    op = TransformTopN(0, asc, nulls_first, k, schema_types(blk))
    blocks = blk.split_by_rows(split) if split else [blk]
    for b in blocks:
        if device_resident:
            b = DataBlock([to_device(c) for c in b.columns], b.num_rows)
        op.transform(b)
    out = op.on_finish()
    op.close()
    ref = oracle().topk(col, asc, nulls_first, k)
    got_rows = out.columns[1].values()
    np.testing.assert_array_equal(got_rows, ref)
    # the key column carries the original values of those rows
    valid = col.valid_mask()[ref]
    np.testing.assert_array_equal(out.columns[0].valid_mask() if col.validity is not None else np.ones(len(ref), bool), valid)
    src = col.values()[ref]
    got = out.columns[0].values()
    if src.dtype.kind == "f":
        np.testing.assert_array_equal(got[valid].view(np.uint64 if src.itemsize == 8 else np.uint32),
                                      src[valid].view(np.uint64 if src.itemsize == 8 else np.uint32))
    else:
        np.testing.assert_array_equal(got[valid], src[valid])
    return out


def test_sort_goldens(gpu):
    DT = {"I64": abi.I64}
    with open(os.path.join(GOLD, "sort.json")) as f:
        for c in json.load(f)["cases"]:
            col = Column.from_data(c["values"], DT[c["dtype"]])
            k = c["limit"] if c["limit"] is not None else len(c["values"])
            out = run_topk(col, c["asc"], c["nulls_first"], k)
            assert list(out.columns[1].values()) == c["rows"], c["src"]
            assert list(out.columns[0].values()) == c["sorted"], c["src"]


@pytest.mark.parametrize("n,k", [(1, 1), (10, 100), (1000, 10), (65536, 1000), (3_000_000, 1000)])
@pytest.mark.parametrize("asc", [True, False])
def test_config4_uniform_f64(gpu, n, k, asc):
    x = oracle().synth_fill(3, 4242, 0, 0, n)
    run_topk(Column.from_data(x), asc, False, k, split=65536 * 4)


def test_device_resident_single_push(gpu):
    x = oracle().synth_fill(3, 7, 0, 0, 5_000_000)
    run_topk(Column.from_data(x), True, False, 1000, device_resident=True)


def test_adversarial_floats(gpu):
    """1% NaN, +-0, +-inf, heavy duplicates: key-sequence parity with row-id tiebreak."""
    rng = np.random.default_rng(5)
    n = 400_000
    x = rng.integers(-50, 50, n).astype(np.float64) / 4.0
    x[rng.random(n) < 0.01] = np.nan
    x[rng.random(n) < 0.01] = -0.0
    x[rng.random(n) < 0.001] = np.inf
    x[rng.random(n) < 0.001] = -np.inf
    for asc in (True, False):
        run_topk(Column.from_data(x), asc, False, 1000, split=50_000)


def test_sorted_inputs_worst_case(gpu):
    """Descending input with ASC order: every row beats the boundary (worst case for the filter)."""
    n = 600_000
    x = np.arange(n, 0, -1).astype(np.float64)
    run_topk(Column.from_data(x), True, False, 500, split=200_000)
    run_topk(Column.from_data(x[::-1].copy()), True, False, 500)


def test_nulls_first_and_last(gpu):
    rng = np.random.default_rng(9)
    n = 100_000
    x = rng.normal(size=n)
    valid = rng.random(n) > 0.001
    for nulls_first in (True, False):
        for asc in (True, False):
            run_topk(Column.from_data(x, validity=valid), asc, nulls_first, 300, split=30_000)
    # fewer non-null rows than k: NULLs fill the tail (nulls last) / the head (nulls first)
    few = Column.from_data(x[:50], validity=[i % 5 != 0 for i in range(50)])
    run_topk(few, True, False, 45)
    run_topk(few, True, True, 45)


@pytest.mark.parametrize("dtype", [abi.I8, abi.I16, abi.I32, abi.I64, abi.U8, abi.U32, abi.U64, abi.F32])
def test_key_dtypes(gpu, dtype):
    from databend_b200.block import np_dtype
    rng = np.random.default_rng(dtype)
    nd = np_dtype(dtype)
    n = 200_000
    if nd.kind == "f":
        vals = rng.normal(size=n).astype(nd)
    else:
        info = np.iinfo(nd)
        vals = rng.integers(info.min, info.max, n, dtype=nd, endpoint=True)
    for asc in (True, False):
        run_topk(Column.from_data(vals), asc, False, 777, split=64_000)


def test_more_nulls_than_k(gpu):
    """More NULL keys than k (nulls first): the k NULL rows with the SMALLEST row ids are returned,
    in row order (ties broken by ascending row id, like the oracle) — also across pushes."""
    rng = np.random.default_rng(17)
    n = 300_000
    x = rng.normal(size=n)
    valid = rng.random(n) > 0.4  # ~120 000 NULLs, k = 500
    for asc in (True, False):
        run_topk(Column.from_data(x, validity=valid), asc, True, 500, split=40_000)
        run_topk(Column.from_data(x, validity=valid), asc, True, 500)
        run_topk(Column.from_data(x, validity=valid), asc, False, 500, split=40_000)


def test_large_k_radix_sort_path(gpu):
    """k above the one-CTA rank sort (4096): the final order comes from the hand-written radix sort."""
    x = oracle().synth_fill(3, 99, 0, 0, 2_000_000)
    x[::7] = x[3]  # heavy ties: the row-id order decides
    run_topk(Column.from_data(x), True, False, 50_000, split=300_000)
    run_topk(Column.from_data(x), False, False, 20_000, device_resident=True)


def test_small_candidate_list_forces_replay(gpu, monkeypatch):
    """A candidate list far smaller than the block and adversarial (sorted) input: the optimistic
    scan overflows, its appends are dropped, and the range is replayed in pieces that fit."""
    monkeypatch.setenv("DBX_TOPK_CAP", "20000")
    n = 3_000_000
    x = np.arange(n, 0, -1).astype(np.float64)  # ASC top-k over descending data: every row beats the boundary
    run_topk(Column.from_data(x), True, False, 100, device_resident=True)
    y = oracle().synth_fill(3, 5, 0, 0, n)
    run_topk(Column.from_data(y), True, False, 100, device_resident=True)


def run_sort(col, asc, nulls_first, split=None, device_resident=False):
    """ORDER BY without LIMIT (limit = 0): the full permutation must equal the oracle's stable sort."""
    blk = DataBlock([col])
    op = TransformTopN(0, asc, nulls_first, 0, schema_types(blk))
    blocks = blk.split_by_rows(split) if split else [blk]
    for b in blocks:
        if device_resident:
            b = DataBlock([to_device(c) for c in b.columns], b.num_rows)
        op.transform(b)
    out = op.on_finish()
    op.close()
    ref = oracle().topk(col, asc, nulls_first, col.length)
    np.testing.assert_array_equal(out.columns[1].values(), ref)
    valid = col.valid_mask()[ref]
    if col.validity is not None:
        np.testing.assert_array_equal(out.columns[0].valid_mask(), valid)
    src, got = col.values()[ref], out.columns[0].values()
    if src.dtype.kind == "f":
        w = np.uint64 if src.itemsize == 8 else np.uint32
        np.testing.assert_array_equal(got[valid].view(w), src[valid].view(w))
    else:
        np.testing.assert_array_equal(got[valid], src[valid])


@pytest.mark.parametrize("n", [0, 1, 2, 4095, 4096, 4097, 100_003, 1_500_000])
def test_full_sort_f64(gpu, n):
    x = oracle().synth_fill(3, 31 + n, 0, 0, n)
    for asc in (True, False):
        run_sort(Column.from_data(x), asc, False, split=250_000 if n > 250_000 else None)


def test_full_sort_adversarial_and_nulls(gpu):
    rng = np.random.default_rng(21)
    n = 700_000
    x = rng.integers(-30, 30, n).astype(np.float64) / 4.0  # heavy ties: stability = row order
    x[rng.random(n) < 0.01] = np.nan
    x[rng.random(n) < 0.01] = -0.0
    x[rng.random(n) < 0.001] = np.inf
    x[rng.random(n) < 0.001] = -np.inf
    valid = rng.random(n) > 0.05
    for asc in (True, False):
        run_sort(Column.from_data(x), asc, False, split=100_000)
        for nulls_first in (True, False):
            run_sort(Column.from_data(x, validity=valid), asc, nulls_first, split=90_000)
    run_sort(Column.from_data(x), True, False, device_resident=True)


@pytest.mark.parametrize("dtype", [abi.I8, abi.I32, abi.I64, abi.U16, abi.U64, abi.F32])
def test_full_sort_dtypes(gpu, dtype):
    from databend_b200.block import np_dtype
    rng = np.random.default_rng(dtype + 100)
    nd = np_dtype(dtype)
    n = 300_000
    if nd.kind == "f":
        vals = rng.normal(size=n).astype(nd)
    else:
        info = np.iinfo(nd)
        vals = rng.integers(info.min, info.max, n, dtype=nd, endpoint=True)
    for asc in (True, False):
        run_sort(Column.from_data(vals), asc, False, split=77_000)


@pytest.mark.parametrize("device_resident", [False, True])
def test_order_by_several_keys(gpu, device_resident):
    """ORDER BY a [dir] [nulls], b [dir] [nulls], c: ties on the earlier keys are broken by the later
    ones (each with its own direction and NULL placement) and finally by input order — one stable
    device radix sort per key, least significant first.  Row ids must equal the oracle's permutation
    exactly; with a LIMIT the sorted result is cut."""
    from oracle import sort_oracle
    rng = np.random.default_rng(31)
    n = 300_001
    a = rng.integers(-3, 4, n).astype(np.int8)
    av = rng.random(n) > 0.1
    b = np.where(rng.random(n) < 0.05, np.nan, rng.integers(-2, 3, n) * 0.5)
    b[rng.random(n) < 0.05] = -0.0
    bv = rng.random(n) > 0.15
    c = rng.integers(0, 2**64, n, dtype=np.uint64) >> np.uint64(58)
    d = rng.standard_normal(n)
    blk = DataBlock([Column.from_data(a, validity=av), Column.from_data(b, validity=bv), Column.from_data(c), Column.from_data(d)])
    blocks = blk.split_by_rows(70_000)
    if device_resident:
        blocks = [DataBlock([to_device(col) for col in bb.columns], bb.num_rows) for bb in blocks]
    for (asc0, nf0), (asc1, nf1), (asc2, nf2), limit in [((True, False), (False, True), (True, False), 0),
                                                          ((False, True), (True, False), (False, False), 0),
                                                          ((True, True), (True, True), (True, True), 1234)]:
        op = TransformTopN(0, asc0, nf0, limit, schema_types(blk), extra_keys=[(1, asc1, nf1), (2, asc2, nf2)])
        for bb in blocks:
            op.transform(bb)
        out = op.on_finish()
        op.close()
        exp = sort_oracle.sort_permutation([(a, av, asc0, nf0), (b, bv, asc1, nf1), (c, None, asc2, nf2)], limit)
        np.testing.assert_array_equal(out.columns[1].values(), exp)
        np.testing.assert_array_equal(out.columns[0].valid_mask(), av[exp])
        np.testing.assert_array_equal(out.columns[0].values()[av[exp]], a[exp][av[exp]])
    # two keys, second one only: every pair of directions on a float key with NaN / -0
    for asc1 in (True, False):
        op = TransformTopN(2, True, False, 0, schema_types(blk), extra_keys=[(1, asc1, False)])
        for bb in blocks:
            op.transform(bb)
        out = op.on_finish()
        op.close()
        exp = sort_oracle.sort_permutation([(c, None, True, False), (b, bv, asc1, False)])
        np.testing.assert_array_equal(out.columns[1].values(), exp)
