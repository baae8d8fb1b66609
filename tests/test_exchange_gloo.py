"""World-size-2 gloo test of the partial->final exchange plumbing (CPU, no GPU).

Each rank holds partial group rows [key][kind][rowcount][sum] for its row range, partitions them
by owner (host restatement of the device rule), exchanges them with ONE variable-size
all-to-all, merges, and the union over ranks must equal the single-process aggregation."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _partials(rank, world, n=20000, n_keys=300):
    rng = np.random.default_rng(123)
    k = rng.integers(-n_keys // 2, n_keys // 2, n).astype(np.int64)
    v = rng.integers(-1000, 1000, n).astype(np.int64)
    lo, hi = n * rank // world, n * (rank + 1) // world
    ks, inv = np.unique(k[lo:hi], return_inverse=True)
    cnt = np.bincount(inv, minlength=len(ks)).astype(np.uint64)
    sm = np.zeros(len(ks), dtype=np.int64)
    np.add.at(sm, inv, v[lo:hi])
    return ks.view(np.uint64), cnt, sm.view(np.uint64), (k, v)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from databend_b200.exchange import all_to_all_rows, owner_of
    keys, cnt, sm, _ = _partials(rank, world)
    kind = np.zeros(len(keys), dtype=np.uint64)
    own = owner_of(keys, kind, world)
    order = np.argsort(own, kind="stable")
    rows = np.stack([keys, kind, cnt, sm], axis=1)[order]          # 4 words = 32-byte rows
    send_counts = np.bincount(own, minlength=world).tolist()
    send = torch.from_numpy(rows.reshape(-1).view(np.uint8).copy())
    recv, recv_counts = all_to_all_rows(send, send_counts, 32)
    got = recv.numpy()[: sum(recv_counts) * 32].view(np.uint64).reshape(-1, 4)
    # every received key must be owned by this rank
    assert (owner_of(got[:, 0], got[:, 1], world) == rank).all()
    ks, inv = np.unique(got[:, 0], return_inverse=True)
    mc = np.zeros(len(ks), dtype=np.uint64)
    ms = np.zeros(len(ks), dtype=np.uint64)
    np.add.at(mc, inv, got[:, 2])
    np.add.at(ms, inv, got[:, 3])
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.stack([ks, mc, ms], axis=1))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_exchange_matches_single_process(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(os.path.join(tmp_path, f"r{r}.npy")) for r in range(world)]
    merged = np.concatenate(parts)
    assert len(np.unique(merged[:, 0])) == len(merged), "a group ended up on two ranks"
    _, _, _, (k, v) = _partials(0, 1)
    ks, inv = np.unique(k, return_inverse=True)
    cnt = np.bincount(inv).astype(np.uint64)
    sm = np.zeros(len(ks), dtype=np.int64)
    np.add.at(sm, inv, v)
    order = np.argsort(merged[:, 0].view(np.int64))
    np.testing.assert_array_equal(merged[order, 0].view(np.int64), ks)
    np.testing.assert_array_equal(merged[order, 1], cnt)
    np.testing.assert_array_equal(merged[order, 2].view(np.int64), sm)


def _join_worker(rank, world, port, out_dir):
    """Host logic of the partitioned hash join shuffle: rows partitioned by the owner rule (host
    restatement of dbx_hash_partition), one all-to-all per column, every received key owned here."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from databend_b200.distributed import all_to_all_columns
    from databend_b200.exchange import owner_of
    rng = np.random.default_rng(77)
    n = 30000
    k_all = rng.integers(-5000, 5000, n).astype(np.int64)
    v_all = rng.integers(0, 2**31, n).astype(np.int32)
    lo, hi = n * rank // world, n * (rank + 1) // world
    k, v = k_all[lo:hi], v_all[lo:hi]
    own = owner_of(k.view(np.uint64), np.zeros(len(k), np.int64), world)
    order = np.argsort(own, kind="stable")
    offs = np.concatenate([[0], np.cumsum(np.bincount(own, minlength=world))]).tolist()
    cols = [torch.from_numpy(k[order].view(np.uint8).copy()), torch.from_numpy(v[order].view(np.uint8).copy())]
    recv, n_recv = all_to_all_columns(cols, [8, 4], offs)
    rk = recv[0].numpy()[: n_recv * 8].view(np.int64)
    rv = recv[1].numpy()[: n_recv * 4].view(np.int32)
    assert (owner_of(rk.view(np.uint64), np.zeros(n_recv, np.int64), world) == rank).all()
    np.save(os.path.join(out_dir, f"j{rank}.npy"), np.stack([rk, rv.astype(np.int64)], axis=1))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_join_shuffle_keeps_every_row(tmp_path):
    world = 2
    mp.spawn(_join_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(os.path.join(tmp_path, f"j{r}.npy")) for r in range(world)])
    rng = np.random.default_rng(77)
    k_all = rng.integers(-5000, 5000, 30000).astype(np.int64)
    v_all = rng.integers(0, 2**31, 30000).astype(np.int64)
    exp = np.stack([k_all, v_all], axis=1)
    np.testing.assert_array_equal(got[np.lexsort((got[:, 1], got[:, 0]))], exp[np.lexsort((exp[:, 1], exp[:, 0]))])
