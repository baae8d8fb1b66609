"""Partitioned hash join over REAL ranks with the fused peer-memory shuffle: one process per rank,
receive regions mapped through CUDA IPC, several send/recv rounds per side (both region parities
reused).  The union of the ranks' results must equal the oracle's inner join as a multiset, and
every joined row must be on the owner of its key."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from databend_b200.block import Column

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [1, 2, 3])
def test_partitioned_join_between_processes(gpu, tmp_path, world):
    from oracle import oracle as orc
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), DBX_EXCH_SPIN_MS="30000")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_peer_shuffle_worker.py"), str(tmp_path)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    try:
        for p in procs:
            outs.append(p.communicate(timeout=600)[0])
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o[-3000:]}"
    rng = np.random.default_rng(1234)
    n_dim, n_fact = 50_000, 400_000
    dk = rng.permutation(n_dim).astype(np.int64) * 3 - 7000
    dv = rng.integers(-2**40, 2**40, n_dim).astype(np.int64)
    fk = dk[rng.integers(0, n_dim, n_fact)].copy()
    fk[::50] = 10**12
    fv = rng.integers(0, 2**31, n_fact).astype(np.int32)
    pi, bi = orc.hash_join_inner(Column.from_data(dk), Column.from_data(fk))
    exp = np.stack([fk[pi], fv[pi].astype(np.int64), dk[bi], dv[bi]], axis=1)
    parts = [np.load(os.path.join(tmp_path, f"join_r{r}.npz")) for r in range(world)]
    got = np.stack([np.concatenate([d[k].astype(np.int64) for d in parts]) for k in ("fk", "fv", "dk", "dv")], axis=1)
    assert len(got) == len(exp)
    np.testing.assert_array_equal(got[np.lexsort(got.T[::-1])], exp[np.lexsort(exp.T[::-1])])
