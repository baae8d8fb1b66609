"""Partial aggregate states in the reference's spill / wire layout (AggregatorParams::spill_schema):
dbx_agg_partial_serialize against the numpy restatement of the reference's StateSerde (oracle/
spill_oracle.py), and dbx_agg_final_merge_serialized fed by (a) a GPU partial's own serialisation and
(b) a CPU partial's states — the final answer must equal the oracle's for the whole input."""
import numpy as np
import pytest

from databend_b200 import abi, expr as E
from databend_b200.block import Column, DataBlock
from databend_b200.transforms import AggregatorParams, TransformFinalAggregate, TransformPartialAggregate, schema_types
from helpers import assert_group_results_equal, sorted_group_result_from_block, sorted_group_result_from_oracle

pytestmark = pytest.mark.gpu


def oracle():
    from oracle import oracle as orc
    return orc


def spill_oracle():
    from oracle import spill_oracle
    return spill_oracle


def make_block(n, seed, n_keys=300):
    rng = np.random.default_rng(seed)
    k = Column.from_data(rng.integers(-5, n_keys, n).astype(np.int32), validity=rng.random(n) > 0.05)
    v = Column.from_data(rng.integers(-10**6, 10**6, n).astype(np.int64), validity=rng.random(n) > 0.3)
    x = Column.from_data(rng.standard_normal(n) * 100)
    u = Column.from_data(rng.integers(0, 60000, n).astype(np.uint16))
    f = Column.from_data((rng.standard_normal(n) * 10).astype(np.float32), validity=rng.random(n) > 0.5)
    return DataBlock([k, v, x, u, f])


PARAMS = AggregatorParams([0], [("sum", 1), ("count", None), ("count", 1), ("avg", 2), ("min", 1), ("max", 4), ("sum", 3), ("avg", 4)])
KINDS = ["sum", "count", "count", "avg", "min", "max", "sum", "avg"]
ARGS = [1, None, 1, 2, 1, 4, 3, 4]


def key_of(cols, valids, i):
    return tuple((cols[j][i].item() if valids[j][i] else None) for j in range(len(cols)))


def fields_as_dict(field_cols, arity, key_cols):
    """{group key: [tuple of field values per aggregate]}"""
    n = len(key_cols[0][0]) if key_cols else len(field_cols[0])
    out = {}
    for i in range(n):
        k = key_of([kc[0] for kc in key_cols], [kc[1] for kc in key_cols], i) if key_cols else ()
        pos, tup = 0, []
        for a in arity:
            tup.append(tuple(field_cols[pos + j][i].item() for j in range(a)))
            pos += a
        assert k not in out
        out[k] = tup
    return out


def expected_fields(blk, key_cols):
    so = spill_oracle()
    keys = [(blk.columns[c].values(), blk.columns[c].valid_mask() if blk.columns[c].validity is not None else None) for c in key_cols]
    args = [None if a is None else (blk.columns[a].values(), blk.columns[a].valid_mask() if blk.columns[a].validity is not None else None) for a in ARGS]
    fields, arity, okeys = so.group_states(keys, args, KINDS)
    flat = [f for fs in fields for f in fs]
    return flat, arity, okeys


def test_serialize_matches_reference_state_layout(gpu):
    blk = make_block(120_000, 5)
    part = TransformPartialAggregate(PARAMS, schema_types(blk))
    for b in blk.split_by_rows(50_000):
        part.transform(b)
    part.on_finish()
    out, arity = part.serialize()
    part.close()
    flat, exp_arity, okeys = expected_fields(blk, [0])
    assert arity == exp_arity == [3, 1, 1, 3, 4, 4, 2, 4]
    assert out.num_columns() == sum(arity) + 1
    # column types: Sum<T>, UInt64 counts, Boolean flags, T values
    assert [c.dtype for c in out.columns] == [abi.I64, abi.BOOL, abi.BOOL, abi.U64, abi.U64, abi.F64, abi.U64, abi.BOOL,
                                              abi.BOOL, abi.I64, abi.BOOL, abi.BOOL, abi.BOOL, abi.F32, abi.BOOL, abi.BOOL,
                                              abi.U64, abi.BOOL, abi.F64, abi.U64, abi.BOOL, abi.BOOL, abi.I32]
    kc = out.columns[-1]
    got = fields_as_dict([c.values() for c in out.columns[:-1]], arity, [(kc.values(), kc.valid_mask())])
    exp = fields_as_dict(flat, exp_arity, okeys)
    assert got.keys() == exp.keys()
    for k in exp:
        for a, (g, e) in enumerate(zip(got[k], exp[k])):
            for j, (gv, ev) in enumerate(zip(g, e)):
                if isinstance(ev, float):
                    assert gv == pytest.approx(ev, rel=1e-9, abs=1e-9) or (np.isnan(gv) and np.isnan(ev)), (k, a, j, gv, ev)
                else:
                    assert gv == ev, (k, a, j, gv, ev)


def final_equals_oracle(fin, blk, params, float_tol=1e-9):
    out = fin.on_finish()[0]
    ref = oracle().filter_group_agg(blk, params.to_c(None), threads=4)
    n_aggs, n_keys = len(params.aggregate_functions), len(params.group_columns)
    g = sorted_group_result_from_block(out, n_aggs, n_keys)
    o = sorted_group_result_from_oracle(ref, [blk.columns[c].dtype for c in params.group_columns])
    assert_group_results_equal(g, o, float_exact=False, rtol=float_tol)


def test_gpu_partial_serialised_into_gpu_final(gpu):
    """partial(A) -> serialize -> merge_serialized, partial(B) adopted directly: final == oracle(A + B)."""
    blk = make_block(200_000, 6)
    a, b = blk.split_by_rows(120_000)
    types = schema_types(blk)
    pa, pb = TransformPartialAggregate(PARAMS, types), TransformPartialAggregate(PARAMS, types)
    pa.transform(a); pb.transform(b)
    pa.on_finish(); pb.on_finish()
    spill, _ = pa.serialize()
    fin = TransformFinalAggregate(PARAMS, types)
    fin.transform(pb)
    fin.merge_serialized(spill)
    final_equals_oracle(fin, blk, PARAMS)
    for op in (pa, pb, fin):
        op.close()


def test_cpu_partial_states_into_gpu_final(gpu):
    """The states of a CPU partial aggregate (numpy restatement of the reference's serialisation) for
    part A merged into a GPU final that also gets a GPU partial for part B."""
    blk = make_block(150_000, 7)
    a, b = blk.split_by_rows(75_000)
    flat, arity, okeys = expected_fields(a, [0])
    cols = [Column.from_data(f, abi.BOOL if f.dtype == np.bool_ else None) for f in flat]
    cols.append(Column.from_data(okeys[0][0], validity=okeys[0][1]))
    spill = DataBlock(cols)
    types = schema_types(blk)
    pb = TransformPartialAggregate(PARAMS, types)
    pb.transform(b)
    pb.on_finish()
    fin = TransformFinalAggregate(PARAMS, types)
    fin.merge_serialized(spill)
    fin.transform(pb)
    final_equals_oracle(fin, blk, PARAMS)
    # a block that does not have the spill schema is refused
    from databend_b200.lib import DbxError
    fin2 = TransformFinalAggregate(PARAMS, types)
    with pytest.raises(DbxError, match="spill schema"):
        fin2.merge_serialized(DataBlock(cols[:-2]))
    for op in (pb, fin, fin2):
        op.close()


def test_round_trip_multi_column_float_and_no_group_by(gpu):
    rng = np.random.default_rng(8)
    n = 90_000
    k1 = Column.from_data(rng.integers(0, 40, n).astype(np.int16), validity=rng.random(n) > 0.1)
    k2 = Column.from_data(rng.integers(0, 7, n).astype(np.uint8))
    v = Column.from_data(rng.integers(-1000, 1000, n).astype(np.int32), validity=rng.random(n) > 0.2)
    x = Column.from_data(rng.standard_normal(n))
    fk = Column.from_data(np.where(rng.random(n) < 0.1, np.nan, rng.integers(-3, 3, n) * 0.5))
    blk = DataBlock([k1, k2, v, x, fk])
    types = schema_types(blk)
    for params in (AggregatorParams([0, 1], [("sum", 2), ("min", 2), ("max", 3), ("count", 2), ("avg", 3)]),
                   AggregatorParams([4], [("sum", 2), ("count", None)]),
                   AggregatorParams([], [("sum", 2), ("avg", 3), ("max", 2), ("count", None)])):
        a, b = blk.split_by_rows(50_000)
        pa, pb = TransformPartialAggregate(params, types), TransformPartialAggregate(params, types)
        pa.transform(a); pb.transform(b)
        pa.on_finish(); pb.on_finish()
        sa, arity = pa.serialize()
        sb, _ = pb.serialize()
        assert sa.num_columns() == sum(arity) + len(params.group_columns)
        fin = TransformFinalAggregate(params, types)
        fin.merge_serialized(sa)
        fin.merge_serialized(sb)
        if params.group_columns == [4]:  # float keys: compare as {bits: aggregates}
            out = fin.on_finish()[0]
            ref = oracle().filter_group_agg(blk, params.to_c(None), threads=4)
            got = {np.float64(kv).tobytes() if not np.isnan(kv) else b"nan": (int(out.columns[0].values()[i]), int(out.columns[1].values()[i]))
                   for i, kv in enumerate(out.columns[2].values())}
            rk = ref[0][0].view(np.float64)
            exp = {np.float64(kv).tobytes() if not np.isnan(kv) else b"nan": (int(ref[2][0][i]), int(ref[2][1][i])) for i, kv in enumerate(rk)}
            assert got == exp
        elif params.group_columns:
            from test_agg_gpu import _group_dict
            out = fin.on_finish()[0]
            keys, kvalid, aggs, avalid, _ = oracle().filter_group_agg(blk, params.to_c(None), threads=4)
            exp = _group_dict([k.view(np.int64) for k in keys], kvalid, aggs, avalid)
            na = len(params.aggregate_functions)
            gk = [out.columns[na + j] for j in range(2)]
            got = _group_dict([k.values().astype(np.int64) for k in gk], [k.valid_mask() for k in gk],
                              [out.columns[i].values() for i in range(na)], [out.columns[i].valid_mask() for i in range(na)])
            assert got.keys() == exp.keys()
            for k in exp:
                for gv, ev in zip(got[k], exp[k]):
                    assert gv == ev or (isinstance(ev, float) and gv == pytest.approx(ev, rel=1e-9)), (k, got[k], exp[k])
        else:
            final_equals_oracle(fin, blk, params)
        for op in (pa, pb, fin):
            op.close()
