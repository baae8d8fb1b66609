"""The multi-GPU aggregate path on REAL ranks: one process per rank, receive buffers mapped
through CUDA IPC (dbx_agg_exchange_connect with handles, not local pointers), system-scope
release/acquire flags, four consecutive queries (epoch parity alternates).  With fewer GPUs than
ranks the ranks share a device — the IPC mapping, flag protocol and region layout are the same.
Results of all ranks together must equal the oracle's, and every group must sit on its owner."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from databend_b200 import abi, expr as E
from databend_b200.block import Column, DataBlock
from databend_b200.transforms import AggregatorParams
from helpers import assert_group_results_equal, sorted_group_result_from_block, sorted_group_result_from_oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _launch(world, out_dir, mode):
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   DBX_EXCH_SPIN_MS="20000")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_peer_exchange_worker.py"), str(out_dir), mode],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
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


@pytest.mark.parametrize("world", [2, 3])
def test_peer_exchange_between_processes(gpu, tmp_path, world):
    from databend_b200.exchange import owner_of
    from oracle import oracle as orc
    _launch(world, tmp_path, "grouped")
    params = AggregatorParams([0], [("sum", 1), ("count", 1), ("avg", 2)])
    filt = E.eq(E.col(1) % E.lit(3), E.lit(0))
    for epoch, (rows, keys) in enumerate([(400_000, 30_000), (250_000, 90_000), (123_457, 1_000), (300_000, 50_000)]):
        k = orc.synth_fill(0, 100 + epoch, keys, 0, rows)
        v = orc.synth_fill(1, 101 + epoch, 0, 0, rows)
        x = orc.synth_fill(2, 102 + epoch, 20, 0, rows)
        k[:3] = -(2**63)
        blk = DataBlock([Column.from_data(k), Column.from_data(v), Column.from_data(x)])
        parts = [np.load(os.path.join(tmp_path, f"e{epoch}_r{r}.npz")) for r in range(world)]
        for r, d in enumerate(parts):
            kk = d["c3"]
            kind = np.where(kk == -(2**63), 1, 0)
            assert (owner_of(kk.view(np.uint64), kind, world) == r).all(), "a group was merged on a rank that does not own it"
        merged = DataBlock([Column.from_data(np.concatenate([d[f"c{i}"] for d in parts]),
                                             validity=np.concatenate([d[f"v{i}"] for d in parts])) for i in range(4)])
        g = sorted_group_result_from_block(merged, 3, 1)
        o = sorted_group_result_from_oracle(orc.filter_group_agg(blk, params.to_c(filt), 4), [abi.I64])
        assert_group_results_equal(g, o)


def test_single_state_allreduce_between_processes(gpu, tmp_path):
    """No GROUP BY on 2 ranks (FinalSingleStateAggregator, transform_single_key.rs:232-278):
    every rank ends with the same state; integer aggregates equal the oracle bit for bit, the f64
    sum equals the rank-ordered sum of the per-rank partial sums."""
    from oracle import oracle as orc
    world = 2
    _launch(world, tmp_path, "single")
    params = AggregatorParams([], [("sum", 1), ("count", 1), ("avg", 2), ("min", 1), ("max", 2)])
    filt = E.eq(E.col(1) % E.lit(3), E.lit(0))
    rows = 300_000
    k = orc.synth_fill(0, 7, 1000, 0, rows)
    v = orc.synth_fill(1, 8, 0, 0, rows)
    x = orc.synth_fill(3, 9, 0, 0, rows)
    blk = DataBlock([Column.from_data(k), Column.from_data(v), Column.from_data(x)])
    _, _, aggs, avalid, _ = orc.filter_group_agg(blk, params.to_c(filt), 1)
    res = [np.load(os.path.join(tmp_path, f"single_r{r}.npz")) for r in range(world)]
    for i in range(5):
        np.testing.assert_array_equal(res[0][f"c{i}"].view(np.uint64), res[1][f"c{i}"].view(np.uint64), err_msg="ranks disagree")
    for i in (0, 1, 3, 4):  # sum(v), count(v), min(v), max(x): exact
        np.testing.assert_array_equal(res[0][f"c{i}"].view(np.uint64), aggs[i].view(np.uint64))
    np.testing.assert_allclose(res[0]["c2"], aggs[2], rtol=1e-12)  # avg over non-integer doubles: order differs from the CPU's
