"""Key-count and skew variants of configs[1] (filter v % 3 = 0 + sum / count / avg GROUP BY k) on one GPU:
uniform keys from 1e5 to 1e7 distinct values, and log-uniform ("Zipf-like", P(k) ~ 1/k) keys over 1e6.
Every variant's group counts are checked against torch.bincount over the selected rows.
usage: python experiments/agg_variants.py [rows]"""
import json
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from databend_b200 import abi, build, lib, expr as E
from databend_b200.block import Column, DataBlock
from databend_b200.transforms import AggregatorParams, TransformFinalAggregate, TransformPartialAggregate

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
build.build()
L = lib.load()
lib.require_device()
dev = torch.device("cuda:0")
g = torch.Generator(device=dev)
g.manual_seed(7)
v = torch.randint(0, 1 << 40, (rows,), dtype=torch.int64, device=dev, generator=g)
x = torch.randint(0, 1 << 20, (rows,), dtype=torch.int64, device=dev, generator=g).to(torch.float64)
sel = (v % 3) == 0
types = [abi.I64, abi.I64, abi.F64]
filt = E.eq(E.col(1) % E.lit(3), E.lit(0))


def keys_uniform(n_keys):
    return torch.randint(0, n_keys, (rows,), dtype=torch.int64, device=dev, generator=g)


def keys_log_uniform(n_keys):
    out = torch.empty(rows, dtype=torch.int64, device=dev)
    step = 1 << 27
    for i in range(0, rows, step):
        u = torch.rand(min(step, rows - i), device=dev, generator=g, dtype=torch.float64)
        out[i:i + step] = torch.clamp((u * math.log(n_keys)).exp().to(torch.int64) - 1, 0, n_keys - 1)
    return out


def run(name, k, n_keys, hint):
    blk = DataBlock([Column.device(abi.I64, rows, k.data_ptr()), Column.device(abi.I64, rows, v.data_ptr()), Column.device(abi.F64, rows, x.data_ptr())], rows)
    params = AggregatorParams([0], [("sum", 1), ("count", 1), ("avg", 2)], expected_groups=hint)
    part = TransformPartialAggregate(params, types, filt)
    fin = TransformFinalAggregate(params, types)
    best, res = 1e9, None
    for i in range(3):
        part.reset(); fin.reset()
        part.transform(blk)
        ms = part.last_kernel_ms()
        fin.transform(part.on_finish())
        res = fin.on_finish()[0]
        best = min(best, ms)
    variant = part.kernel_variant()
    part.close(); fin.close()
    cnt = torch.bincount(k[sel], minlength=n_keys)
    exp_groups = int((cnt > 0).sum())
    keys_out = torch.from_numpy(res.columns[3].values().astype(np.int64)).to(dev)
    got = torch.zeros(n_keys, dtype=torch.int64, device=dev)
    got[keys_out] = torch.from_numpy(res.columns[1].values().astype(np.int64)).to(dev)
    ok = bool(res.num_rows == exp_groups and torch.equal(got, cnt))
    top = float(cnt.max()) / float(cnt.sum())
    print(json.dumps({"variant": name, "rows": rows, "distinct_keys": exp_groups, "size_hint": hint, "kernel_ms": round(best, 3),
                      "frac_of_hbm_roofline": round(24.0 * rows / (best * 1e-3) / 1e9 / 6572.2, 4), "hottest_key_share": round(top, 5),
                      "counts_match_bincount": ok, "kernel": variant}), flush=True)
    del cnt, got, keys_out


for nk in (100_000, 1_000_000, 2_000_000, 4_000_000, 10_000_000):
    k = keys_uniform(nk)
    run(f"uniform {nk}", k, nk, nk)
    del k
k = keys_log_uniform(1_000_000)
run("log-uniform (P(k) ~ 1/k) over 1e6", k, 1_000_000, 1_000_000)
run("log-uniform over 1e6, no size hint", k, 1_000_000, 0)
