"""CPU ORACLE (test infrastructure only) for the spill_schema layout of partial aggregate states.

numpy restatement of what the reference's partial aggregate serialises per group
(AggregatorParams::spill_schema, src/query/service/src/pipelines/processors/transforms/aggregator/
aggregator_params.rs:103-117; StateSerde::serialize_type / batch_serialize of
  count   src/query/functions/src/aggregates/aggregate_count.rs:170-190      (UInt64)
  sum     .../aggregate_sum.rs:155-170                                        (Sum<T>)
  avg     .../aggregate_avg.rs:106-127                                        (Sum<T>, UInt64)
  min/max .../aggregate_min_max_any.rs:315-345                                (Boolean, T)
  + one Boolean per adaptor: adaptors/aggregate_null_adaptor.rs:508-517 (Nullable argument),
    adaptors/aggregate_ornull_adaptor.rs:184-190 (every function but count,
    aggregate_function_factory.rs:219-249)).
Parity of this file with the reference is by reading those sources ("parity unpinned": the reference
has no golden file for the spill block and cannot run here)."""
import numpy as np


def sum_dtype(dt):
    dt = np.dtype(dt)
    return np.float64 if dt.kind == "f" else (np.int64 if dt.kind == "i" else np.uint64)


def group_states(keys, args, aggs):
    """keys: list of (values, valid or None); args: list of (values, valid or None) per aggregate (None
    for count(*)); aggs: list of kind names.  Returns (flattened field arrays per aggregate, arities,
    group key arrays [(values, valid)]) with groups in first-appearance order."""
    n = len(keys[0][0])
    ident = []
    for v, ok in keys:
        ok = np.ones(n, bool) if ok is None else np.asarray(ok, bool)
        vv = np.asarray(v)
        if vv.dtype.kind == "f":  # floats group by bits, every NaN one group (group_hash.rs:599-619)
            bits = np.where(np.isnan(vv), np.float64("nan"), vv).astype(vv.dtype).view(np.uint64 if vv.dtype.itemsize == 8 else np.uint32).astype(np.uint64)
        else:
            bits = vv.astype(np.int64).view(np.uint64) if vv.dtype.kind == "i" else vv.astype(np.uint64)
        ident.append(np.where(ok, bits, 0))
        ident.append(ok.astype(np.uint64))
    mat = np.stack(ident, axis=1)
    _, first, inv = np.unique(mat, axis=0, return_index=True, return_inverse=True)
    order = np.argsort(first)
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    gid = rank[inv.reshape(-1)]
    g = len(order)
    out_keys = []
    for v, ok in keys:
        ok = np.ones(n, bool) if ok is None else np.asarray(ok, bool)
        rep = first[order]
        out_keys.append((np.where(ok[rep], np.asarray(v)[rep], 0).astype(np.asarray(v).dtype), ok[rep]))
    fields, arity = [], []
    for kind, arg in zip(aggs, args):
        fs = []
        if arg is None:
            fs.append(np.bincount(gid, minlength=g).astype(np.uint64))
            fields.append(fs); arity.append(1)
            continue
        v, ok = arg
        v = np.asarray(v)
        nullable = ok is not None
        ok = np.ones(n, bool) if ok is None else np.asarray(ok, bool)
        cnt = np.bincount(gid, weights=ok.astype(np.float64), minlength=g).astype(np.uint64)
        has = cnt > 0
        if kind == "count":
            fields.append([cnt]); arity.append(1)
            continue
        if kind in ("sum", "avg"):
            st = sum_dtype(v.dtype)
            acc = np.zeros(g, dtype=st)
            with np.errstate(over="ignore"):
                np.add.at(acc, gid[ok], v[ok].astype(st))
            fs.append(acc)
            if kind == "avg":
                fs.append(cnt)
        else:
            val = np.zeros(g, dtype=v.dtype)
            vv, gg = v[ok], gid[ok]
            if v.dtype.kind == "f":  # OrderedFloat: NaN is the greatest value
                key = np.where(np.isnan(vv), np.inf, vv)
                nanmask = np.isnan(vv)
            for j in range(g):
                sel = vv[gg == j]
                if len(sel) == 0:
                    continue
                if v.dtype.kind == "f":
                    nn = sel[~np.isnan(sel)]
                    if kind == "min":
                        val[j] = nn.min() if len(nn) else np.nan
                    else:
                        val[j] = np.nan if np.isnan(sel).any() else nn.max()
                else:
                    val[j] = sel.min() if kind == "min" else sel.max()
            fs.append(has.copy())
            fs.append(val)
        if nullable:
            fs.append(has.copy())
        fs.append(has.copy())
        fields.append(fs); arity.append(len(fs))
    return fields, arity, out_keys
