"""The expression oracle (oracle/eval_oracle.py) against the reference's own printed results:
tests/golden/arithmetic.json is transcribed from functions/tests/it/scalars/testdata/
{arithmetic,cast,boolean,comparison}.txt by tests/golden/make_arith_golden.py."""
import json
import math
import os

import pytest

from oracle import eval_oracle as eo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "arithmetic.json")


def load_cases():
    with open(GOLD) as f:
        return json.load(f)


def tree(e):
    """json lists -> the tuples eval_oracle takes."""
    if e[0] == "col":
        return ("col", e[1])
    if e[0] == "lit":
        return ("lit", e[1], e[2])
    if e[0] == "cast":
        return ("cast", tree(e[1]), e[2], e[3])
    return ("call", e[1]) + tuple(tree(a) for a in e[2:])


def columns_of(case):
    cols = [(c["type"], [float(v) if c["type"][0] == "F" else v for v in c["values"]], c["valid"]) for c in case["columns"]]
    if not cols:
        cols = [("U8", [0] * case["rows"], None)]
    return cols


def same_value(t, got, exp):
    if t[0] == "F":
        exp = float(exp)
        return (math.isnan(got) and math.isnan(exp)) or got == exp or abs(got - exp) <= 1e-12 * abs(exp)  # the golden prints shortest round-trip digits
    return int(got) == int(exp)


@pytest.mark.parametrize("case", load_cases()["cases"], ids=lambda c: c["src"])
def test_oracle_matches_reference_output(case):
    t, nullable, vals, oks = eo.evaluate(tree(case["expr"]), columns_of(case))
    assert t == case["out_type"], case["checked"]
    exp_valid = case["out_valid"] or [1] * case["rows"]
    assert [int(o) for o in oks] == [int(v) for v in exp_valid[:case["rows"]]], case["checked"]
    for r in range(case["rows"]):
        if exp_valid[r]:
            assert same_value(t, vals[r], case["out_values"][r]), (case["checked"], r, vals[r], case["out_values"][r])


@pytest.mark.parametrize("case", load_cases()["errors"], ids=lambda c: c["src"])
def test_oracle_raises_reference_errors(case):
    with pytest.raises(eo.EvalFailure) as ei:
        eo.evaluate(tree(case["expr"]), columns_of(case))
    assert ei.value.msg == case["error"] and ei.value.row == case["row"]


def test_result_type_table():
    """arithmetics_type.rs:240-265 spot checks (the golden outputs' types cover the rest)."""
    assert eo.t_add_mul("I8", "I16") == "I32" and eo.t_add_mul("U8", "U8") == "U16" and eo.t_add_mul("U64", "I8") == "I64"
    assert eo.t_minus("U8", "U8") == "I16" and eo.t_minus("U32", "F64") == "F64"
    assert eo.t_intdiv("U32", "F64") == "I64" and eo.t_intdiv("U8", "U32") == "U32"
    assert eo.t_modulo("I8", "I8") == "I16" and eo.t_modulo("U16", "U8") == "U8" and eo.t_modulo("U8", "I8") == "U8"
    assert eo.t_negate("U8") == "I16" and eo.t_negate("F32") == "F32" and eo.t_negate("U64") == "I64"


def test_sort_oracle_matches_reference_golden_permutations():
    """oracle/sort_oracle.py (multi-column ORDER BY restatement) on the reference's single-key sort
    goldens (tests/golden/sort.json, from expression/tests/it/sort.rs)."""
    import numpy as np
    from oracle import sort_oracle
    with open(os.path.join(os.path.dirname(GOLD), "sort.json")) as f:
        cases = json.load(f)["cases"]
    np_dt = {"I64": np.int64, "F64": np.float64, "I32": np.int32, "U64": np.uint64, "F32": np.float32}
    for c in cases:
        vals = c["values"]
        valid = None
        if any(v is None for v in vals):
            valid = [v is not None for v in vals]
            vals = [0 if v is None else v for v in vals]
        arr = np.asarray(vals, dtype=np_dt.get(c["dtype"], np.float64))
        perm = sort_oracle.sort_permutation([(arr, valid, c["asc"], c["nulls_first"])], c["limit"] or 0)
        if c.get("rows") is not None:
            assert perm.tolist() == c["rows"], c["src"]


def test_spill_oracle_states_finalize_to_the_c_oracle_results():
    """oracle/spill_oracle.py (serialised partial states, restated from the reference's StateSerde) has no
    reference golden; this ties it to the pinned C oracle instead: finalising its states (sum, count,
    sum / count, min / max with the NULL flags) must give exactly the C oracle's final results."""
    import numpy as np
    from databend_b200 import abi
    from databend_b200.block import Column, DataBlock
    from databend_b200.transforms import AggregatorParams
    from oracle import oracle as orc
    from oracle import spill_oracle
    rng = np.random.default_rng(4)
    n = 50_000
    k = rng.integers(-3, 200, n).astype(np.int32)
    kv = rng.random(n) > 0.05
    v = rng.integers(-10**6, 10**6, n).astype(np.int64)
    vv = rng.random(n) > 0.3
    x = rng.integers(0, 1 << 20, n).astype(np.float64)
    f = (rng.integers(-500, 500, n) * 0.25).astype(np.float32)
    fv = rng.random(n) > 0.5
    blk = DataBlock([Column.from_data(k, validity=kv), Column.from_data(v, validity=vv), Column.from_data(x), Column.from_data(f, validity=fv)])
    kinds = ["sum", "count", "count", "avg", "min", "max", "avg"]
    args = [1, None, 1, 2, 1, 3, 3]
    params = AggregatorParams([0], list(zip(kinds, args)))
    cols = {1: (v, vv), 2: (x, None), 3: (f, fv)}
    fields, arity, okeys = spill_oracle.group_states([(k, kv)], [None if a is None else cols[a] for a in args], kinds)
    keys, kvalid, aggs, avalid, _ = orc.filter_group_agg(blk, params.to_c(None), threads=2)
    exp = {}
    for i in range(len(aggs[0])):
        key = int(keys[0].view(np.int64)[i]) if kvalid[0][i] else None
        exp[key] = [(aggs[a][i].item() if avalid[a][i] else None) for a in range(len(kinds))]
    assert len(okeys[0][0]) == len(exp)
    for g in range(len(okeys[0][0])):
        key = int(okeys[0][0][g]) if okeys[0][1][g] else None
        got = []
        for a, kind in enumerate(kinds):
            fs = [fld[g] for fld in fields[a]]
            if kind == "count":
                got.append(int(fs[0]))
            elif kind == "sum":
                got.append(fs[0].item() if fs[-1] else None)
            elif kind == "avg":
                got.append(float(np.float64(fs[0]) / np.float64(fs[1])) if fs[-1] else None)
            else:
                got.append(fs[1].item() if fs[0] else None)
        assert got == exp[key], (key, got, exp[key])


def test_expression_builder_flattens_to_postfix():
    """Host logic of databend_b200/scalar_expr.py (no GPU): operator overloads build the tree, flatten()
    emits the postfix program the C-ABI takes, oversized trees and unknown functions are refused."""
    from databend_b200 import abi, scalar_expr as sx
    from databend_b200.lib import DbxError
    e = sx.call("and", sx.call("gt", sx.cast((sx.col(0) * sx.col(1) + sx.col(1)) % sx.lit(7, abi.U8), abi.I64), sx.lit(3, abi.I64)), sx.call("not", sx.call("is_null", sx.col(2))))
    prog = sx.flatten(e)
    kinds = [prog.nodes[i].kind for i in range(prog.n_nodes)]
    funcs = [prog.nodes[i].func for i in range(prog.n_nodes) if prog.nodes[i].kind == abi.EXPR_CALL]
    assert kinds == [abi.EXPR_COLUMN, abi.EXPR_COLUMN, abi.EXPR_CALL, abi.EXPR_COLUMN, abi.EXPR_CALL, abi.EXPR_CONST, abi.EXPR_CALL, abi.EXPR_CAST,
                     abi.EXPR_CONST, abi.EXPR_CALL, abi.EXPR_COLUMN, abi.EXPR_CALL, abi.EXPR_CALL, abi.EXPR_CALL]
    assert funcs == [abi.FN_MULTIPLY, abi.FN_PLUS, abi.FN_MODULO, abi.FN_GT, abi.FN_IS_NULL, abi.FN_NOT, abi.FN_AND]
    assert prog.nodes[7].cast_to == abi.I64 and prog.nodes[5].c.dtype == abi.U8 and prog.nodes[5].c.v.u64 == 7
    big = sx.col(0)
    for _ in range(abi.MAX_EXPR_NODES):
        big = big + sx.col(0)
    with pytest.raises(DbxError, match="too large"):
        sx.flatten(big)
    with pytest.raises(DbxError, match="not built"):
        sx.call("sqrt", sx.col(0))
