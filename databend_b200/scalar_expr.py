"""Scalar expressions over a DataBlock: host-side mirror of the reference's `Expr` tree
(src/query/expression/src/expression.rs: ColumnRef / Constant / Cast / FunctionCall) for the numeric
and boolean functions libdbx evaluates on the device (include/dbx.h: dbx_eval_scalar).  Trees are
flattened to the postfix program the C-ABI takes; no compute happens here."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

from . import abi
from .block import Column, DataBlock, make_scalar
from .lib import DbxError, check, load

FUNCS = {"plus": abi.FN_PLUS, "minus": abi.FN_MINUS, "multiply": abi.FN_MULTIPLY, "divide": abi.FN_DIVIDE, "div": abi.FN_DIV,
         "modulo": abi.FN_MODULO, "negate": abi.FN_NEGATE, "eq": abi.FN_EQ, "noteq": abi.FN_NOTEQ, "lt": abi.FN_LT, "lte": abi.FN_LTE,
         "gt": abi.FN_GT, "gte": abi.FN_GTE, "and": abi.FN_AND, "or": abi.FN_OR, "not": abi.FN_NOT, "is_null": abi.FN_IS_NULL,
         "is_not_null": abi.FN_IS_NOT_NULL}
UNARY = {"negate", "not", "is_null", "is_not_null"}


@dataclass
class SExpr:
    kind: int
    func: str = ""
    col: int = 0
    dtype: int = 0
    value: object = None
    try_cast: bool = False
    args: List["SExpr"] = field(default_factory=list)

    def __add__(self, o): return call("plus", self, o)
    def __sub__(self, o): return call("minus", self, o)
    def __mul__(self, o): return call("multiply", self, o)
    def __truediv__(self, o): return call("divide", self, o)
    def __floordiv__(self, o): return call("div", self, o)
    def __mod__(self, o): return call("modulo", self, o)
    def __neg__(self): return call("negate", self)


def col(i: int) -> SExpr:
    return SExpr(abi.EXPR_COLUMN, col=i)


def lit(value, dtype: int) -> SExpr:
    """Scalar literal of an explicit type (the reference's binder picks the smallest integer type)."""
    return SExpr(abi.EXPR_CONST, dtype=dtype, value=value)


def cast(e: SExpr, dtype: int, try_cast: bool = False) -> SExpr:
    return SExpr(abi.EXPR_CAST, dtype=dtype, try_cast=try_cast, args=[e])


def call(name: str, *args: SExpr) -> SExpr:
    if name not in FUNCS:
        raise DbxError(abi.ERR_UNSUPPORTED, f"function {name} is not built")
    assert len(args) == (1 if name in UNARY else 2)
    return SExpr(abi.EXPR_CALL, func=name, args=list(args))


def flatten(e: SExpr) -> abi.Expr:
    nodes: List[SExpr] = []

    def walk(x: SExpr):
        for a in x.args:
            walk(a)
        nodes.append(x)
    walk(e)
    if len(nodes) > abi.MAX_EXPR_NODES:
        raise DbxError(abi.ERR_UNSUPPORTED, "expression too large")
    out = abi.Expr()
    out.n_nodes = len(nodes)
    for i, x in enumerate(nodes):
        n = out.nodes[i]
        n.kind = x.kind
        if x.kind == abi.EXPR_COLUMN:
            n.col = x.col
        elif x.kind == abi.EXPR_CONST:
            n.c = make_scalar(x.dtype, x.value)
        elif x.kind == abi.EXPR_CAST:
            n.cast_to, n.try_cast = x.dtype, int(x.try_cast)
        else:
            n.func = FUNCS[x.func]
    return out


class EvalError(DbxError):
    def __init__(self, status, message, row):
        super().__init__(status, message)
        self.row = row


def eval_scalar(block: DataBlock, e: SExpr, device: int = 0, out_mem: int = abi.MEM_HOST):
    """Evaluator::run(expr) over `block` -> (result column, its dtype | NULLABLE flag).  With
    out_mem = MEM_DEVICE the result stays in HBM: (library-owned dbx_block to release, dtype)."""
    from .transforms import _block_from_c
    ce = flatten(e)
    b, keep = block.as_c()
    out = abi.Block()
    odt, erow = C.c_int32(0), C.c_int64(-1)
    st = load().dbx_eval_scalar(device, C.byref(ce), C.byref(b), out_mem, C.byref(out), C.byref(odt), C.byref(erow))
    if st != abi.OK:
        msg = (load().dbx_last_error(None) or b"").decode("utf-8", "replace")
        raise EvalError(st, msg, erow.value)
    if out_mem != abi.MEM_HOST:
        return out, odt.value
    res = _block_from_c(out, device)
    return res.columns[0], odt.value
