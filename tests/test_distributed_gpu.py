"""Single-GPU checks of the pieces the multi-GPU drivers are made of: dbx_hash_partition (owner
rule = host restatement, rows preserved as a multiset) and the top-k candidate merge."""
import numpy as np
import pytest

from databend_b200 import abi
from databend_b200.block import Column, DataBlock
from databend_b200.transforms import TransformTopN, schema_types, to_device

pytestmark = pytest.mark.gpu


def oracle():
    from oracle import oracle as orc
    return orc


@pytest.mark.parametrize("n,parts", [(0, 3), (1, 1), (1000, 8), (300_001, 5)])
def test_hash_partition(gpu, n, parts):
    from databend_b200.distributed import hash_partition
    from databend_b200.exchange import owner_of
    rng = np.random.default_rng(n + parts)
    k = rng.integers(-2**40, 2**40, n).astype(np.int64)
    v = rng.integers(0, 2**31, n).astype(np.int32)
    x = rng.standard_normal(n)
    blk = DataBlock([to_device(Column.from_data(k)), to_device(Column.from_data(v)), to_device(Column.from_data(x))], n)
    outs, offs = hash_partition(blk, 0, parts)
    assert offs[0] == 0 and offs[-1] == n
    ko = outs[0].cpu().numpy()[: n * 8].view(np.int64)
    vo = outs[1].cpu().numpy()[: n * 4].view(np.int32)
    xo = outs[2].cpu().numpy()[: n * 8].view(np.float64)
    for p in range(parts):
        seg = ko[offs[p]:offs[p + 1]]
        assert (owner_of(seg.view(np.uint64), np.zeros(len(seg), np.int64), parts) == p).all()
    # same rows, as a multiset
    a = np.lexsort((x, v, k))
    b = np.lexsort((xo, vo, ko))
    np.testing.assert_array_equal(k[a], ko[b])
    np.testing.assert_array_equal(v[a], vo[b])
    np.testing.assert_array_equal(x[a].view(np.uint64), xo[b].view(np.uint64))


def test_topk_merge_equals_global_topk(gpu):
    """Row-range shards -> local top-k -> merge == top-k over the whole column (row ids global)."""
    from databend_b200.distributed import topk_merge
    rng = np.random.default_rng(5)
    n, k, world = 200_000, 100, 4
    x = rng.integers(0, 5000, n).astype(np.float64)  # many ties: the row-id order matters
    x[rng.random(n) < 0.001] = np.nan
    col = Column.from_data(x)
    ref = oracle().topk(col, True, False, k)
    # simulate ranks one after another in this process (world_size 1 merge of pre-gathered parts)
    parts = []
    for r in range(world):
        lo, hi = n * r // world, n * (r + 1) // world
        op = TransformTopN(0, True, False, k, [abi.F64])
        op.transform(DataBlock([Column.from_data(x[lo:hi])]))
        out = op.on_finish()
        op.close()
        parts.append((out, lo))
    vals = np.concatenate([p.columns[0].values() for p, _ in parts])
    rows = np.concatenate([p.columns[1].values() + lo for p, lo in parts])
    merged = topk_merge(DataBlock([Column.from_data(vals), Column.from_data(rows)], len(vals)), 0, k, True, False)
    np.testing.assert_array_equal(merged.columns[1].values(), ref)
