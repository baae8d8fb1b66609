"""kNN micro-benchmark: device-generated N(0,1) corpus, one batch of queries; prints the
tensor-core pass time (CUDA events inside the library) and the whole-search wall time."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from databend_b200 import abi  # noqa: E402
from databend_b200.block import Column  # noqa: E402
from databend_b200.lib import check, load  # noqa: E402
from databend_b200.transforms import DeviceBuffer  # noqa: E402
from databend_b200.vector import VectorTopN  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=2_000_000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--nq", type=int, default=1024)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--fn", default="cosine_distance")
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
L = load()
buf = DeviceBuffer(a.n * a.dim * 4)
check(L.dbx_synth_fill(0, 4, 42, 0, 0, a.n * a.dim, buf.ptr))
qb = DeviceBuffer(a.nq * a.dim * 4)
check(L.dbx_synth_fill(0, 4, 43, 0, 0, a.nq * a.dim, qb.ptr))
t0 = time.time()
op = VectorTopN(a.fn, Column.device(abi.VEC_F32, a.n, buf.ptr, vec_dim=a.dim))
t_create = time.time() - t0
q = Column.device(abi.VEC_F32, a.nq, qb.ptr, vec_dim=a.dim)
for r in range(a.reps):
    t0 = time.time()
    idx, dist = op.search(q, a.k)
    wall = time.time() - t0
    ms, launches = op.last_gemm_ms()
    flop = 2.0 * a.nq * a.n * a.dim
    print(json.dumps({"n": a.n, "dim": a.dim, "nq": a.nq, "k": a.k, "fn": a.fn, "create_s": round(t_create, 3),
                      "search_wall_ms": round(wall * 1e3, 2), "gemm_ms": round(ms, 3), "gemm_launches": launches,
                      "gemm_tflops": round(flop / (ms * 1e-3) / 1e12, 1) if ms > 0 else None,
                      "qps": round(a.nq / wall, 1), "stats": op.stats(), "first": [int(idx[0, 0]), float(dist[0, 0])]}), flush=True)
