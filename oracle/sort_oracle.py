"""CPU ORACLE (test infrastructure only) for multi-column ORDER BY.

Restates the order the reference's sort produces for a list of SortColumnDescription{offset, asc,
nulls_first} (src/query/expression/src/kernels/sort.rs:41-111, sort_compare.rs:197-296): rows compare
by the first key, ties by the next, ...; each key with its own direction and NULL placement; floats
compare as OrderedFloat (src/common/base/src/base/ordered_float.rs:147-201: NaN greatest and equal to
itself, -0.0 == +0.0); remaining ties keep input order (the permutation is built stably).  Pinned for
one key by tests/golden/sort.json (expression/tests/it/sort.rs); several keys: by those rules."""
import numpy as np


def column_rank(values, valid, asc, nulls_first):
    """-> (null_key, rank) integer arrays: ascending lexicographic order of (null_key, rank) is the column's order."""
    v = np.asarray(values)
    n = len(v)
    ok = np.ones(n, bool) if valid is None else np.asarray(valid, bool)
    img = v.astype(np.float64) + 0.0 if v.dtype.kind == "f" else v
    _, inv = np.unique(img, return_inverse=True)  # NaNs sort last and collapse into one value; -0.0 == 0.0
    rank = inv.reshape(-1).astype(np.int64)
    if not asc:
        rank = rank.max(initial=0) - rank
    rank = np.where(ok, rank, 0)
    null_key = np.where(ok, 1, 0) if nulls_first else np.where(ok, 0, 1)
    return null_key.astype(np.int64), rank


def sort_permutation(columns, limit=0):
    """columns: list of (values, valid or None, asc, nulls_first), most significant first.  Returns row ids in output order."""
    n = len(columns[0][0])
    keys = [np.arange(n)]
    for values, valid, asc, nulls_first in reversed(columns):
        nk, rk = column_rank(values, valid, asc, nulls_first)
        keys += [rk, nk]
    perm = np.lexsort(tuple(keys))
    return perm[:limit] if limit else perm
