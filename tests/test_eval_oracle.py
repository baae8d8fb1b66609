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
