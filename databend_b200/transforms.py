"""Host-side mirror of the reference's operator interface for the hot path.

Each class corresponds to a reference Processor / Transform and forwards to one `dbx_op`
handle of libdbx through the C-ABI (include/dbx.h).  Names, argument meaning and error
behaviour follow the reference so the parity tests read like the reference's own tests:

  TransformFilter             src/query/pipeline/transforms/src/processors/transforms/filters/filter_predicate.rs:35-104
  AggregatorParams            src/query/service/src/pipelines/processors/transforms/aggregator/aggregator_params.rs:30-78
  TransformPartialAggregate   .../aggregator/transform_aggregate_partial.rs:117-304  (AccumulatingTransform)
  TransformFinalAggregate     .../aggregator/transform_aggregate_final.rs:67-330
  (no GROUP BY)               .../aggregator/transform_single_key.rs:42-279

In the reference a worker thread calls Processor::process(); here the caller drives
transform()/on_finish() directly (the adaptor contract of transform_accumulating.rs:30-37).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import abi, expr as E
from .block import Column, DataBlock, np_dtype
from .lib import DbxError, check, load

_AGG_NAMES = {"sum": abi.AGG_SUM, "count": abi.AGG_COUNT, "avg": abi.AGG_AVG, "min": abi.AGG_MIN, "max": abi.AGG_MAX}


class DeviceBuffer:
    """A cudaMalloc'ed buffer owned by Python (tests / bench)."""

    def __init__(self, nbytes: int, device: int = 0):
        self.device, self.nbytes = device, nbytes
        p = C.c_void_p()
        check(load().dbx_device_alloc(device, nbytes, C.byref(p)))
        self.ptr = p.value

    def free(self):
        if self.ptr:
            load().dbx_device_free(self.device, self.ptr)
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def upload(self, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        check(load().dbx_memcpy_h2d(self.device, self.ptr, arr.ctypes.data, arr.nbytes))

    def download(self, dtype, count: int) -> np.ndarray:
        out = np.empty(count, dtype=dtype)
        check(load().dbx_memcpy_d2h(self.device, out.ctypes.data, self.ptr, out.nbytes))
        return out


def to_device(col: Column, device: int = 0) -> Column:
    """Copy a host column into HBM (device-resident pipelines)."""
    if col.is_const or col.data is None:
        return col
    buf = DeviceBuffer(max(1, col.data.nbytes), device)
    buf.upload(col.data)
    out = Column(col.dtype, col.length, dev_ptr=buf.ptr, vec_dim=col.vec_dim, data_bit_offset=col.data_bit_offset)
    out._keep.append(buf)
    if col.validity is not None:
        vb = DeviceBuffer(max(1, col.validity.nbytes), device)
        vb.upload(col.validity)
        out.dev_validity = vb.ptr
        out.validity_bit_offset = col.validity_bit_offset
        out._keep.append(vb)
    return out


def _block_from_c(b: abi.Block, device: int) -> DataBlock:
    """Copy a library-owned HOST output block into numpy-backed columns, then release it."""
    cols = []
    for i in range(b.num_cols):
        c = b.cols[i]
        n = c.len
        if c.is_const:  # BlockEntry::Const stays const
            k = c.konst
            if k.is_null:
                v = None
            elif c.dtype in (abi.F32, abi.F64):
                v = k.v.f64
            elif c.dtype in (abi.U8, abi.U16, abi.U32, abi.U64, abi.BOOL):
                v = k.v.u64
            else:
                v = k.v.i64
            cols.append(Column.new_const(c.dtype, v, n))
            continue
        assert c.mem == abi.MEM_HOST
        if c.dtype == abi.BOOL:
            nb = (c.data_bit_offset + n + 7) // 8
            arr = np.zeros(max(nb, 1), dtype=np.uint8)
            if nb:
                C.memmove(arr.ctypes.data, c.data, nb)
            col = Column(abi.BOOL, n, data=arr, data_bit_offset=c.data_bit_offset)
        elif c.dtype == abi.VEC_F32:
            arr = np.empty((n, c.vec_dim), dtype=np.float32)
            if n:
                C.memmove(arr.ctypes.data, c.data, arr.nbytes)
            col = Column(abi.VEC_F32, n, data=arr, vec_dim=c.vec_dim)
        else:
            nd = np_dtype(c.dtype)
            arr = np.empty(n, dtype=nd)
            if n:
                C.memmove(arr.ctypes.data, c.data, arr.nbytes)
            col = Column(c.dtype, n, data=arr)
        if c.validity:
            nb = (c.validity_bit_offset + n + 7) // 8
            v = np.empty(max(nb, 1), dtype=np.uint8)
            if nb:
                C.memmove(v.ctypes.data, c.validity, nb)
            col.validity = v
            col.validity_bit_offset = c.validity_bit_offset
        cols.append(col)
    rows = b.num_rows
    check(load().dbx_block_release(C.byref(b)))
    return DataBlock(cols, rows)


class _Op:
    """Owns one dbx_op handle."""

    def __init__(self, kind: int, params, input_types: Sequence[int], device: int):
        self._h = C.c_void_p()
        self.device = device
        types = (C.c_int32 * max(1, len(input_types)))(*input_types)
        self._params = params
        check(load().dbx_op_create(kind, C.cast(C.byref(params), C.c_void_p), types, len(input_types), device,
                                   C.byref(self._h)))

    @property
    def handle(self):
        return self._h

    def push(self, block: DataBlock):
        b, keep = block.as_c()
        check(load().dbx_op_push(self._h, C.byref(b)), self._h)

    def finish(self):
        check(load().dbx_op_finish(self._h), self._h)

    def pull_c(self, out_mem: int = abi.MEM_HOST) -> Optional[abi.Block]:
        b = abi.Block()
        has = C.c_int32(0)
        check(load().dbx_op_pull(self._h, out_mem, C.byref(b), C.byref(has)), self._h)
        return b if has.value else None

    def reset(self):
        check(load().dbx_op_reset(self._h), self._h)

    def synchronize(self):
        check(load().dbx_op_synchronize(self._h), self._h)

    def last_kernel_ms(self) -> float:
        ms = C.c_float(0)
        check(load().dbx_op_last_kernel_ms(self._h, C.byref(ms)), self._h)
        return ms.value

    def kernel_ms(self, back: int = 0) -> float:
        """Device time of the dominant kernel(s) of an earlier push (0 = last, 1 = the one before ...)."""
        ms = C.c_float(0)
        check(load().dbx_op_kernel_ms(self._h, back, C.byref(ms)), self._h)
        return ms.value

    def kernel_variant(self) -> str:
        """"specialised" (kernel compiled for this plan) or "precompiled kernels (<why>)"."""
        buf = C.create_string_buffer(2048)
        check(load().dbx_op_kernel_variant(self._h, buf, 2048), self._h)
        return buf.value.decode("utf-8", "replace")

    def inputs_consumed(self):
        """Block until every pushed block has been read completely (pinned/device inputs may be reused)."""
        check(load().dbx_op_inputs_consumed(self._h), self._h)

    def close(self):
        if self._h:
            load().dbx_op_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class TransformFilter(_Op):
    """TransformFilter (filters/filter_predicate.rs:35-104) = FilterExecutor::filter
    (filter_executor.rs:82-160): `transform(block)` returns the rows for which the predicate is
    true, in input order, for every column.  NULL predicate values count as false."""

    def __init__(self, predicate: Optional[E.Node], input_types: Sequence[int], device: int = 0):
        super().__init__(abi.OP_FILTER, E.build_predicate(predicate), input_types, device)

    def transform(self, block: DataBlock) -> DataBlock:
        self.push(block)
        b = self.pull_c(abi.MEM_HOST)
        return _block_from_c(b, self.device)


@dataclass
class AggregatorParams:
    """AggregatorParams::try_create (aggregator_params.rs:44-78): group column offsets and the
    aggregate functions with their argument offsets.  `aggregate_functions` entries are
    (factory name, argument column offset or None for count())."""

    group_columns: List[int]
    aggregate_functions: List[Tuple[str, Optional[int]]]
    expected_groups: int = 0

    def to_c(self, filter_expr: Optional[E.Node] = None) -> abi.AggParams:
        p = abi.AggParams()
        p.n_group_cols = len(self.group_columns)
        for i, g in enumerate(self.group_columns):
            p.group_cols[i] = g
        p.n_aggs = len(self.aggregate_functions)
        for i, (name, arg) in enumerate(self.aggregate_functions):
            if name not in _AGG_NAMES:
                raise DbxError(abi.ERR_UNSUPPORTED, f"Unknown aggregate function {name}")  # factory.get error
            p.aggs[i].kind = _AGG_NAMES[name]
            p.aggs[i].arg_col = -1 if arg is None else arg
        p.filter = E.build_predicate(filter_expr)
        p.expected_groups = self.expected_groups
        return p


def schema_types(block_or_types, nullable: Optional[Sequence[bool]] = None) -> List[int]:
    """DataSchema -> input_types[] with the DBX_NULLABLE flag for Nullable(T) columns."""
    if isinstance(block_or_types, DataBlock):
        out = []
        for c in block_or_types.columns:
            nul = c.validity is not None or c.dev_validity != 0 or (c.is_const and c.const_value is None)
            out.append(c.dtype | (abi.NULLABLE if nul else 0))
        return out
    types = list(block_or_types)
    if nullable:
        types = [t | (abi.NULLABLE if n else 0) for t, n in zip(types, nullable)]
    return types


class TransformPartialAggregate(_Op):
    """[TransformFilter ->] TransformPartialAggregate fused into one device pass.

    transform(block) accumulates (AccumulatingTransform::transform returns no blocks while
    the hash table has room); on_finish() yields the payload reference that
    TransformFinalAggregate consumes (AggregateMeta::AggregatePayload)."""

    def __init__(self, params: AggregatorParams, input_types: Sequence[int], filter_expr: Optional[E.Node] = None,
                 device: int = 0):
        self.params = params
        super().__init__(abi.OP_AGG_PARTIAL, params.to_c(filter_expr), input_types, device)

    def transform(self, block: DataBlock) -> List[DataBlock]:
        self.push(block)
        return []

    def on_finish(self):
        self.finish()
        return self  # the payload stays in HBM; the handle is the AggregateMeta

    def serialize(self):
        """The partial's groups in the reference's spill / wire layout (AggregatorParams::spill_schema):
        (block with every `agg_i` Tuple flattened into consecutive columns followed by the group
        columns, [arity of agg_0, ...])."""
        out = abi.Block()
        arity = (C.c_int32 * abi.MAX_AGGS)()
        check(load().dbx_agg_partial_serialize(self._h, abi.MEM_HOST, C.byref(out), arity), self._h)
        return _block_from_c(out, self.device), list(arity[:len(self.params.aggregate_functions)])


class TransformFinalAggregate(_Op):
    def __init__(self, params: AggregatorParams, input_types: Sequence[int], device: int = 0):
        self.params = params
        super().__init__(abi.OP_AGG_FINAL, params.to_c(None), input_types, device)

    def transform(self, partial: TransformPartialAggregate) -> List[DataBlock]:
        """handle_meta -> combine_payload (transform_aggregate_final.rs:201-303)."""
        check(load().dbx_agg_final_merge_partial(self._h, partial.handle), self._h)
        return []

    def merge_serialized(self, block: DataBlock):
        """Merge a spill-schema block (flattened tuples + group columns) produced by a CPU partial
        aggregate or by TransformPartialAggregate.serialize() elsewhere."""
        b, keep = block.as_c()
        check(load().dbx_agg_final_merge_serialized(self._h, C.byref(b)), self._h)
        del keep

    def merge_rows(self, dev_rows_ptr: int, n_rows: int):
        check(load().dbx_agg_final_merge_rows(self._h, dev_rows_ptr, n_rows), self._h)

    def on_finish(self, out_mem: int = abi.MEM_HOST):
        """merge_result: returns the result blocks [aggs..., group keys...]."""
        self.finish()
        if out_mem == abi.MEM_HOST:
            b = self.pull_c(abi.MEM_HOST)
            return [] if b is None else [_block_from_c(b, self.device)]
        b = self.pull_c(abi.MEM_DEVICE)
        return [] if b is None else [b]


class TransformTopN(_Op):
    """ORDER BY key LIMIT k: TransformSortPartial + limit-aware merge / TransformPartialTopN+FinalTopN
    (sorts/sort_partial.rs:24-60, top_n/transform_partial_top_n.rs:73-130) as one streaming device
    top-k.  SortColumnDescription{offset, asc, nulls_first} + LimitType::LimitRows(k)
    (kernels/sort.rs:41-63).  The result block is [key column, row_id Int64] in output order; ties
    are broken by ascending row id."""

    def __init__(self, offset: int, asc: bool, nulls_first: bool, limit: int, input_types: Sequence[int], device: int = 0,
                 extra_keys: Sequence[Tuple[int, bool, bool]] = ()):
        """extra_keys: further SortColumnDescriptions (offset, asc, nulls_first) that break ties of the
        earlier keys (ORDER BY a, b, c); the result still is [first key column, row_id]."""
        p = abi.TopkParams()
        p.key_col, p.asc, p.nulls_first, p.limit = offset, int(asc), int(nulls_first), limit
        p.n_extra_keys = len(extra_keys)
        for i, (c, a, nf) in enumerate(extra_keys):
            p.extra_key_cols[i], p.extra_asc[i], p.extra_nulls_first[i] = c, int(a), int(nf)
        super().__init__(abi.OP_TOPK, p, input_types, device)

    def transform(self, block: DataBlock) -> List[DataBlock]:
        self.push(block)
        return []

    def on_finish(self) -> DataBlock:
        self.finish()
        b = self.pull_c(abi.MEM_HOST)
        return _block_from_c(b, self.device)


class HashJoin(_Op):
    """Inner hash join behind the reference's `Join` trait (new_hash_join/join.rs:26-53):
    add_block(build block) / final_build() / probe_block(block) -> joined blocks.
    Output columns = probe columns then build columns (inner_join.rs:236-245); output row order is
    unspecified (compare as multisets)."""

    def __init__(self, build_types: Sequence[int], probe_types: Sequence[int], build_key: int, probe_key: int,
                 device: int = 0, kind: int = abi.JOIN_INNER, expected_build_rows: int = 0):
        """kind: abi.JOIN_INNER, JOIN_LEFT_SEMI or JOIN_LEFT_ANTI (probe side = left; semi/anti
        emit probe columns only: left_join_semi.rs / left_join_anti.rs)."""
        p = abi.JoinParams()
        p.kind, p.build_key_col, p.probe_key_col, p.n_build_cols = kind, build_key, probe_key, len(build_types)
        p.expected_build_rows = expected_build_rows
        super().__init__(abi.OP_JOIN, p, list(build_types) + list(probe_types), device)

    def add_block(self, block: DataBlock):
        self.push(block)

    def final_build(self):
        self.finish()

    def probe_block(self, block: DataBlock, out_mem: int = abi.MEM_HOST) -> List[DataBlock]:
        b, keep = block.as_c()
        check(load().dbx_join_probe(self._h, C.byref(b)), self._h)
        out = []
        while True:
            ob = self.pull_c(out_mem)
            if ob is None:
                break
            out.append(_block_from_c(ob, self.device) if out_mem == abi.MEM_HOST else ob)
        return out


def filter_group_aggregate(blocks: Sequence[DataBlock], params: AggregatorParams, filter_expr: Optional[E.Node] = None,
                           input_types: Optional[Sequence[int]] = None, device: int = 0,
                           n_partials: int = 1) -> DataBlock:
    """Convenience pipeline: Filter -> Partial x n_partials -> Final, like
    PipelineBuilder::build_aggregate_partial/final would wire it."""
    types = list(input_types) if input_types is not None else schema_types(blocks[0])
    partials = [TransformPartialAggregate(params, types, filter_expr, device) for _ in range(n_partials)]
    for i, b in enumerate(blocks):
        partials[i % n_partials].transform(b)
    final = TransformFinalAggregate(params, types, device)
    for p in partials:
        final.transform(p.on_finish())
    out = final.on_finish()
    for p in partials:
        p.close()
    final.close()
    return out[0]
