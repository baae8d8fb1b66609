"""ctypes binding of the CPU oracle (oracle/libdbx_oracle.so) — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (databend_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

from databend_b200 import abi
from databend_b200.block import Column, DataBlock

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libdbx_oracle.so")


class AggResult(C.Structure):
    _fields_ = [
        ("n_groups", C.c_int64),
        ("n_group_cols", C.c_int32),
        ("n_aggs", C.c_int32),
        ("key_bits", C.POINTER(C.c_uint64) * abi.MAX_GROUP_COLS),
        ("key_valid", C.POINTER(C.c_uint8) * abi.MAX_GROUP_COLS),
        ("agg_bits", C.POINTER(C.c_uint64) * abi.MAX_AGGS),
        ("agg_valid", C.POINTER(C.c_uint8) * abi.MAX_AGGS),
        ("agg_dtype", C.c_int32 * abi.MAX_AGGS),
    ]


def build() -> str:
    """Compile the C restatement (gcc).  Building the checker is not using it."""
    src = os.path.join(_HERE, "dbx_oracle.c")
    if (not os.path.exists(_SO)) or os.path.getmtime(_SO) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "dbx_oracle.h")),
            os.path.getmtime(os.path.join(_HERE, "..", "include", "dbx.h"))):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_filter_select.argtypes = [C.POINTER(abi.Block), C.POINTER(abi.Predicate), C.c_void_p,
                                        C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_take_column.argtypes = [C.POINTER(abi.Column), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_agg_hash_u64.argtypes = [C.c_uint64]
        L.orc_agg_hash_u64.restype = C.c_uint64
        L.orc_filter_group_agg.argtypes = [C.POINTER(abi.Block), C.POINTER(abi.AggParams), C.c_int,
                                           C.POINTER(AggResult), C.POINTER(C.c_int64)]
        L.orc_agg_result_free.argtypes = [C.POINTER(AggResult)]
        L.orc_hash_join_inner.argtypes = [C.POINTER(abi.Column), C.POINTER(abi.Column),
                                          C.POINTER(C.POINTER(C.c_int64)), C.POINTER(C.POINTER(C.c_int64)),
                                          C.POINTER(C.c_int64)]
        L.orc_hash_join.argtypes = [C.c_int, C.POINTER(abi.Column), C.POINTER(abi.Column),
                                    C.POINTER(C.POINTER(C.c_int64)), C.POINTER(C.POINTER(C.c_int64)), C.POINTER(C.c_int64)]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_topk.argtypes = [C.POINTER(abi.Column), C.c_int, C.c_int, C.c_int64, C.c_void_p, C.POINTER(C.c_int64)]
        L.orc_cosine_distance.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_cosine_distance.restype = C.c_float
        L.orc_l2_distance.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_l2_distance.restype = C.c_float
        L.orc_distance_rows.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int64, C.c_int64,
                                        C.c_void_p, C.c_int]
        L.orc_distance_rows.restype = None
        L.orc_synth_fill.argtypes = [C.c_int, C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int]
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


class OracleError(Exception):
    def __init__(self, status: int, row: int = -1):
        self.status, self.row = status, row
        msg = "Division by zero" if status == abi.ERR_BAD_ARGUMENTS else f"oracle status {status}"
        super().__init__(f"{msg} (row {row})")


def num_threads() -> int:
    return int(lib().orc_num_threads())


def agg_hash(x: int) -> int:
    return int(lib().orc_agg_hash_u64(x & 0xFFFFFFFFFFFFFFFF))


def filter_select(block: DataBlock, pred: abi.Predicate) -> np.ndarray:
    """FilterExecutor::select -> true_selection[..count] (ascending u32 row ids)."""
    b, keep = block.as_c()
    sel = np.empty(max(1, block.num_rows), dtype=np.uint32)
    n, err = C.c_int64(0), C.c_int64(-1)
    st = lib().orc_filter_select(C.byref(b), C.byref(pred), sel.ctypes.data, C.byref(n), C.byref(err))
    if st != abi.OK:
        raise OracleError(st, err.value)
    return sel[: n.value].copy()


def take(column: Column, sel: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    from databend_b200.block import np_dtype
    c = column.as_c()
    out = np.empty(len(sel), dtype=np_dtype(column.dtype))
    valid = np.empty(len(sel), dtype=np.uint8)
    sel = np.ascontiguousarray(sel, dtype=np.uint32)
    st = lib().orc_take_column(C.byref(c), sel.ctypes.data, len(sel), out.ctypes.data, valid.ctypes.data)
    if st != abi.OK:
        raise OracleError(st)
    return out, valid.astype(bool)


def filter_block(block: DataBlock, pred: abi.Predicate) -> List[Tuple[np.ndarray, np.ndarray]]:
    """FilterExecutor::filter = select + take for every column."""
    sel = filter_select(block, pred)
    return [take(c, sel) if c.dtype != abi.BOOL else None for c in block.columns]


def _bits_to_array(bits: np.ndarray, dtype: int) -> np.ndarray:
    from databend_b200.block import np_dtype
    nd = np_dtype(dtype)
    if nd.itemsize == 8:
        return bits.view(nd)
    if nd.kind == "f":  # f32 min/max: bit pattern of the f64-widened value is stored by the oracle
        return bits.view(np.float64).astype(nd)
    return bits.astype(nd)  # truncation of the sign-extended word


def filter_group_agg(block: DataBlock, params: abi.AggParams, threads: int = 1):
    """Returns (keys, key_valid, aggs, agg_valid, agg_dtypes): lists of numpy arrays, one entry
    per group, arbitrary order (assert_block_value_sort_eq compares sorted)."""
    b, keep = block.as_c()
    res = AggResult()
    err = C.c_int64(-1)
    st = lib().orc_filter_group_agg(C.byref(b), C.byref(params), threads, C.byref(res), C.byref(err))
    if st != abi.OK:
        raise OracleError(st, err.value)
    n = res.n_groups
    keys, kvalid, aggs, avalid, adt = [], [], [], [], []
    for k in range(res.n_group_cols):
        keys.append(np.ctypeslib.as_array(res.key_bits[k], shape=(max(n, 1),))[:n].copy())
        kvalid.append(np.ctypeslib.as_array(res.key_valid[k], shape=(max(n, 1),))[:n].astype(bool))
    for a in range(res.n_aggs):
        bits = np.ctypeslib.as_array(res.agg_bits[a], shape=(max(n, 1),))[:n].copy()
        dt = res.agg_dtype[a]
        if dt in (abi.F32,):
            arr = bits.view(np.float64).astype(np.float32)
        else:
            arr = _bits_to_array(bits, dt)
        aggs.append(arr)
        avalid.append(np.ctypeslib.as_array(res.agg_valid[a], shape=(max(n, 1),))[:n].astype(bool))
        adt.append(dt)
    lib().orc_agg_result_free(C.byref(res))
    return keys, kvalid, aggs, avalid, adt


def hash_join_inner(build_key: Column, probe_key: Column) -> Tuple[np.ndarray, np.ndarray]:
    bk, pk = build_key.as_c(), probe_key.as_c()
    pp, pb = C.POINTER(C.c_int64)(), C.POINTER(C.c_int64)()
    n = C.c_int64(0)
    st = lib().orc_hash_join_inner(C.byref(bk), C.byref(pk), C.byref(pp), C.byref(pb), C.byref(n))
    if st != abi.OK:
        raise OracleError(st)
    m = n.value
    probe = np.ctypeslib.as_array(pp, shape=(max(m, 1),))[:m].copy()
    build = np.ctypeslib.as_array(pb, shape=(max(m, 1),))[:m].copy()
    lib().orc_free(pp)
    lib().orc_free(pb)
    return probe, build


def hash_join(kind: int, build_key: Column, probe_key: Column) -> Tuple[np.ndarray, np.ndarray]:
    """kind: abi.JOIN_INNER / JOIN_LEFT_SEMI / JOIN_LEFT_ANTI / JOIN_LEFT.  Returns (probe_idx, build_idx),
    build_idx = -1 where the output row carries no build row."""
    bk, pk = build_key.as_c(), probe_key.as_c()
    pp, pb = C.POINTER(C.c_int64)(), C.POINTER(C.c_int64)()
    n = C.c_int64(0)
    st = lib().orc_hash_join(kind, C.byref(bk), C.byref(pk), C.byref(pp), C.byref(pb), C.byref(n))
    if st != abi.OK:
        raise OracleError(st)
    m = n.value
    probe = np.ctypeslib.as_array(pp, shape=(max(m, 1),))[:m].copy()
    build = np.ctypeslib.as_array(pb, shape=(max(m, 1),))[:m].copy()
    lib().orc_free(pp)
    lib().orc_free(pb)
    return probe, build


def topk(key: Column, asc: bool, nulls_first: bool, k: int) -> np.ndarray:
    c = key.as_c()
    out = np.empty(max(1, min(k, key.length)), dtype=np.int64)
    n = C.c_int64(0)
    st = lib().orc_topk(C.byref(c), int(asc), int(nulls_first), k, out.ctypes.data, C.byref(n))
    if st != abi.OK:
        raise OracleError(st)
    return out[: n.value].copy()


def cosine_distance(a: Sequence[float], b: Sequence[float]) -> np.float32:
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    if len(a) != len(b):
        raise ValueError(f"Vector length not equal: {len(a)} != {len(b)}")  # distance.rs:20-26
    return np.float32(lib().orc_cosine_distance(a.ctypes.data, b.ctypes.data, len(a)))


def l2_distance(a: Sequence[float], b: Sequence[float]) -> np.float32:
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    if len(a) != len(b):
        raise ValueError(f"Vector length not equal: {len(a)} != {len(b)}")
    return np.float32(lib().orc_l2_distance(a.ctypes.data, b.ctypes.data, len(a)))


def distance_rows(kind: int, lhs: np.ndarray, rhs: np.ndarray, threads: int = 1) -> np.ndarray:
    """calculate_distance: lhs/rhs are [rows, dim] or [dim] (const side)."""
    lhs = np.ascontiguousarray(lhs, dtype=np.float32)
    rhs = np.ascontiguousarray(rhs, dtype=np.float32)
    lc, rc = lhs.ndim == 1, rhs.ndim == 1
    dim = lhs.shape[-1]
    rows = rhs.shape[0] if lc else lhs.shape[0]
    out = np.empty(rows, dtype=np.float32)
    lib().orc_distance_rows(kind, lhs.ctypes.data, int(lc), rhs.ctypes.data, int(rc), rows, dim, out.ctypes.data, threads)
    return out


def synth_fill(kind: int, seed: int, a: int, first_row: int, length: int, threads: int = 0) -> np.ndarray:
    dt = {0: np.int64, 1: np.int64, 2: np.float64, 3: np.float64, 4: np.float32, 5: np.int64}[kind]
    out = np.empty(length, dtype=dt)
    lib().orc_synth_fill(kind, seed, a, first_row, length, out.ctypes.data, threads or num_threads())
    return out
