#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path: filter -> hash-aggregate.

Workload (BASELINE.json configs[1]):
    SELECT k, sum(v), count(v), avg(x) FROM t WHERE v % 3 = 0 GROUP BY k
    t = 1e9 rows, k Int64 uniform [0,1e6), v Int64 uniform [-2^31,2^31), x Float64 integer-valued [0,2^20)
    synthetic, counter-based generator (dbx_synth_fill / orc_synth_fill, seeds 42/43/44).

One "step" = one full query over the 1e9-row batch: table reset, fused filter+partial
aggregation, final merge, result materialisation.

  value  rows/s with the three columns already resident in HBM (CUDA events on the operator's
         stream; includes table re-initialisation and result finalisation, excludes nothing)
  e2e    the same query through the public operator API with HOST (pinned) columns pushed in
         blocks: host->device copies and the device->host copy of the result are inside the
         timed region
  roofline  the fused kernel alone: 24 algorithmic bytes per row / its CUDA-event duration
            against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the CPU oracle (reference-algorithm restatement, OpenMP) on a bounded sample

`--impl reference` times the CPU oracle on the host cores (the Rust reference cannot be
built in this image: no cargo/rustc).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rows/sec filter->hash-agg (sum,count,avg GROUP BY 1e6 int64 keys) over int64/float64 columns"
SEEDS = (42, 43, 44)
N_KEYS = 1_000_000
BYTES_PER_ROW = 24.0  # three 8-byte columns, each read exactly once
NCU_TRAFFIC_PER_LAUNCH = 7.094290e9 + 0.544827e9  # bytes; one 2^28-row launch of the fused kernel (profiles/r01b_filter_group_agg_ncu_full.csv)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi during the timed region."""

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for nm, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_query():
    from databend_b200 import expr as E
    from databend_b200.transforms import AggregatorParams
    params = AggregatorParams([0], [("sum", 1), ("count", 1), ("avg", 2)])
    filt = E.eq(E.col(1) % E.lit(3), E.lit(0))
    return params, filt


# ---------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """CPU arm: the oracle port on all host threads, each step a bounded sample of the workload."""
    import numpy as np
    from databend_b200.block import Column, DataBlock
    from oracle import oracle as orc
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = len(os.sched_getaffinity(0))  # torchrun pins OMP_NUM_THREADS=1: ask for every host core explicitly
    n = args.cpu_rows
    params, filt = make_query()
    k = orc.synth_fill(0, SEEDS[0], N_KEYS, 0, n)
    v = orc.synth_fill(1, SEEDS[1], 0, 0, n)
    x = orc.synth_fill(2, SEEDS[2], 20, 0, n)
    blk = DataBlock([Column.from_data(k), Column.from_data(v), Column.from_data(x)])
    cp = params.to_c(filt)
    for _ in range(args.warmup):
        orc.filter_group_agg(blk, cp, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.filter_group_agg(blk, cp, threads=threads)
    dt = (time.perf_counter() - t0) / args.steps
    val = n / dt
    sample = f"{n} rows of the same synthetic columns per step (reference-algorithm CPU restatement in C/OpenMP; the Rust reference cannot be built here)"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": "filter+hash-agg sum/count/avg GROUP BY 1e6 int64 keys, WHERE v%3=0", "rows": n},
        "cpu_baseline": {"value": val, "unit": "rows/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if not args.no_knn:  # second half of the metric, on the same host cores: row-wise cosine_distance + top-k
        from databend_b200 import abi
        sn, sq, dim, kk = min(args.knn_rows, args.knn_cpu_rows), 8, args.knn_dim, args.knn_k
        rng = np.random.default_rng(0)
        c = rng.standard_normal((sn, dim)).astype(np.float32)
        qs = rng.standard_normal((sq, dim)).astype(np.float32)
        orc.distance_rows(abi.DIST_COSINE, c, qs[0], threads=threads)
        t0 = time.perf_counter()
        for i in range(sq):
            d = orc.distance_rows(abi.DIST_COSINE, c, qs[i], threads=threads)
            np.argpartition(d, min(kk, sn - 1))[:kk]
        dt_k = time.perf_counter() - t0
        line["knn"] = {"metric": "kNN QPS @768d (cosine_distance, brute force, exact top-k)", "impl": "reference",
                       "value": sq / dt_k * sn / args.knn_rows, "unit": "queries/s",
                       "cpu_baseline": {"value": sq / dt_k * sn / args.knn_rows, "unit": "queries/s", "cores": threads, "kind": "port",
                                        "sample": f"{sq} queries x {sn} rows x {dim} dims, row-wise cosine_distance (oracle, OpenMP) + top-{kk}, scaled by {sn}/{args.knn_rows} rows"}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------- kNN leg
def run_knn(args, L, dev, rank, world, barrier):
    """configs[4]: cosine_distance brute-force kNN, corpus sharded by rows across ranks, queries
    replicated; per-rank top-k all-gathered and merged.  Returns the "knn" object of the JSON line."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from databend_b200 import abi, lib
    from databend_b200.block import Column
    from databend_b200.transforms import DeviceBuffer
    from databend_b200.vector import VectorTopN

    n_total, dim, nq, k = args.knn_rows, args.knn_dim, args.knn_queries, args.knn_k
    r0, r1 = n_total * rank // world, n_total * (rank + 1) // world
    n = r1 - r0
    cbuf = DeviceBuffer(n * dim * 4, dev)
    lib.check(L.dbx_synth_fill(dev, 4, 42, 0, r0 * dim, n * dim, cbuf.ptr))
    qbuf = DeviceBuffer(nq * dim * 4, dev)
    lib.check(L.dbx_synth_fill(dev, 4, 43, 0, 0, nq * dim, qbuf.ptr))
    t0 = time.perf_counter()
    op = VectorTopN("cosine_distance", Column.device(abi.VEC_F32, n, cbuf.ptr, vec_dim=dim), dev)
    create_s = time.perf_counter() - t0
    q_dev = Column.device(abi.VEC_F32, nq, qbuf.ptr, vec_dim=dim)
    q_host = Column.vector(qbuf.download(np.float32, nq * dim).reshape(nq, dim))

    def search(q):
        idx, d = op.search(q, k)
        if world == 1:
            return idx, d
        # one collective: [nq, k] global row ids and the distance bits, packed into one int64 tensor
        pack = np.empty((2, nq, k), dtype=np.int64)
        pack[0] = idx + r0
        pack[1] = d.view(np.int32)
        t = torch.from_numpy(pack).to(f"cuda:{dev}", non_blocking=True)
        g = torch.empty((world,) + tuple(t.shape), dtype=torch.int64, device=f"cuda:{dev}")
        dist.all_gather_into_tensor(g, t)
        ai = g[:, 0].permute(1, 0, 2).reshape(nq, world * k)
        ad = g[:, 1].permute(1, 0, 2).reshape(nq, world * k).to(torch.int32).view(torch.float32)
        # merge: ascending (distance, row id); NaN last like OrderedFloat
        key = torch.where(torch.isnan(ad), torch.full_like(ad, float("inf")), ad)
        o1 = torch.argsort(ai, dim=1, stable=True)
        key1, ai1, ad1 = key.gather(1, o1), ai.gather(1, o1), ad.gather(1, o1)
        o2 = torch.argsort(key1, dim=1, stable=True)[:, :k]
        return ai1.gather(1, o2).cpu().numpy(), ad1.gather(1, o2).cpu().numpy()

    def timed(q, steps, warmup):
        for _ in range(warmup):
            search(q)
        barrier()
        gemm = []
        t0 = time.perf_counter()
        for _ in range(steps):
            res = search(q)
            gemm.append(op.last_gemm_ms()[0])
        barrier()
        ms = (time.perf_counter() - t0) * 1e3 / steps
        t = torch.tensor([ms, sum(gemm) / len(gemm)], dtype=torch.float64, device=f"cuda:{dev}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist(), res

    launches0 = L.dbx_kernel_launch_count()
    (ms_dev, gemm_ms), res = timed(q_dev, args.steps, args.warmup)
    launches = L.dbx_kernel_launch_count() - launches0
    (ms_host, _), _ = timed(q_host, max(1, min(args.steps, 3)), 1)
    stats = op.stats()
    op.close()
    if rank != 0:
        return None
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, src = 1400.0, "fallback (B200_PROFILING.md sustained)"
    if os.path.exists(p):
        with open(p) as f:
            peak, src = json.load(f).get("bf16_tflops_sustained", 1400.0), "measured sustained cuBLAS bf16 (MEASURED_PEAKS.json)"
    flop = 2.0 * nq * n * dim  # per rank and batch: the similarity GEMM
    achieved = flop / (gemm_ms * 1e-3) / 1e12
    out = {
        "metric": "kNN QPS @768d (cosine_distance, brute force, exact top-k)", "value": nq / (ms_dev * 1e-3), "unit": "queries/s",
        "ms_per_batch": ms_dev, "n_gpus": world, "scaling": "strong", "dtype": "bf16 candidate GEMM (f32 accumulate) + exact f32 re-rank",
        "config": {"workload": "configs[4]", "corpus_rows": n_total, "rows_per_gpu": n, "dim": dim, "queries": nq, "k": k,
                   "data": "synthetic N(0,1), device-generated", "create_s": create_s,
                   "parallelism": f"corpus rows x{world}" + ("" if world == 1 else " + all-gather of per-GPU top-k")},
        "gpu_launches_per_batch": int(launches // max(1, args.steps)),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": None, "kernel": f"knn_gemm_filter_kernel<{stats['cluster']}>", "kernel_ms": gemm_ms,
                     "flop_per_launch_set": flop, "peak_source": src},
        "certified_queries": stats["certified"], "exact_fallback_queries": stats["exact_fallback"],
        "e2e": {"value": nq / (ms_host * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": nq * dim * 4,
                "d2h_bytes_per_step": nq * k * 12, "ms_per_batch": ms_host,
                "timing": "host wall clock around VectorTopN.search() with HOST query vectors, max over ranks"},
    }
    if world == 1 and not args.no_cpu:
        from oracle import oracle as orc
        threads = len(os.sched_getaffinity(0))
        sn, sq = min(n_total, 1_000_000), 8
        rng = np.random.default_rng(0)
        c = rng.standard_normal((sn, dim)).astype(np.float32)
        qs = rng.standard_normal((sq, dim)).astype(np.float32)
        orc.distance_rows(abi.DIST_COSINE, c, qs[0], threads=threads)
        t0 = time.perf_counter()
        for i in range(sq):
            d = orc.distance_rows(abi.DIST_COSINE, c, qs[i], threads=threads)
            np.argpartition(d, k)[:k]
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": sq / dt * sn / n_total, "unit": "queries/s", "cores": threads, "kind": "port",
                               "sample": f"{sq} queries x {sn} rows, row-wise cosine_distance (oracle, OpenMP) + top-k, scaled by {sn}/{n_total} rows"}
    return out

# ---------------------------------------------------------------------------------- GPU arm
def run_dbx(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from databend_b200 import abi, build, lib
    from databend_b200.block import Column, DataBlock
    from databend_b200.exchange import all_to_all_rows
    from databend_b200.transforms import (DeviceBuffer, TransformFinalAggregate, TransformPartialAggregate)

    build.build()
    L = lib.load()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    lib.require_device()
    dev = local_rank
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))

    total_rows = args.rows
    # strong scaling: the 1e9-row table is split into `world` contiguous row ranges
    r_begin = total_rows * rank // world
    r_end = total_rows * (rank + 1) // world
    n = r_end - r_begin
    params, filt = make_query()
    types = [abi.I64, abi.I64, abi.F64]

    bufs = [DeviceBuffer(n * 8, dev) for _ in range(3)]
    lib.check(L.dbx_synth_fill(dev, 0, SEEDS[0], args.keys, r_begin, n, bufs[0].ptr))
    lib.check(L.dbx_synth_fill(dev, 1, SEEDS[1], 0, r_begin, n, bufs[1].ptr))
    lib.check(L.dbx_synth_fill(dev, 2, SEEDS[2], 20, r_begin, n, bufs[2].ptr))
    dblock = DataBlock([Column.device(abi.I64, n, bufs[0].ptr), Column.device(abi.I64, n, bufs[1].ptr),
                        Column.device(abi.F64, n, bufs[2].ptr)], n)

    part = TransformPartialAggregate(params, types, filt, dev)
    fin = TransformFinalAggregate(params, types, dev)
    sp = C.c_void_p()
    lib.check(L.dbx_op_stream(part.handle, C.byref(sp)))
    part_stream = torch.cuda.ExternalStream(sp.value, device=dev)
    lib.check(L.dbx_op_stream(fin.handle, C.byref(sp)))
    fin_stream = torch.cuda.ExternalStream(sp.value, device=dev)

    kernel_ms = []

    use_peer = world > 1 and os.environ.get("DBX_EXCHANGE", "peer") == "peer"
    xchg = None
    if use_peer:
        from databend_b200.exchange import PeerExchange
        ok = 1
        try:
            xchg = PeerExchange(part, rank, world)
            xchg.connect()
        except Exception as e:  # e.g. no peer access between these GPUs: all ranks fall back together
            ok = 0
            print(f"[bench] rank {rank}: peer-memory exchange unavailable ({e}); using the NCCL all-to-all", file=sys.stderr)
        t_ok = torch.tensor([ok], dtype=torch.int32, device=f"cuda:{dev}")
        dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
        if int(t_ok.item()) == 0:
            if xchg is not None:
                xchg.close()
            xchg = None
            use_peer = False

    _k = C.c_float(0)

    def last_kernel_ms():
        lib.check(L.dbx_op_last_kernel_ms(part.handle, C.byref(_k)), part.handle)
        return _k.value

    def exchange_and_finish(out_mem):
        """partial -> (N>1: hash-partition + exchange) -> final -> result block"""
        part.on_finish()
        if world == 1:
            fin.transform(part)
        elif use_peer:
            # rows go straight into the owners' HBM over NVLink; the merge kernel waits on the
            # sources' flags on the device: no NCCL call, staging copy or host sync in between
            xchg.scatter(part)
            part.reset()  # re-arm the partial now: its table is cleared while the owners merge
            xchg.merge(fin)
        else:
            rows_ptr = C.c_void_p()
            offs = (C.c_int64 * (world + 1))()
            rb = C.c_int32(0)
            lib.check(L.dbx_agg_partial_partition(part.handle, world, C.byref(rows_ptr), offs, C.byref(rb)), part.handle)
            row_bytes = rb.value
            send_counts = [offs[i + 1] - offs[i] for i in range(world)]
            total_send = offs[world]
            send = torch.empty(max(total_send, 1) * row_bytes, dtype=torch.uint8, device=f"cuda:{dev}")
            if total_send:
                lib.check(L.dbx_memcpy_d2d(dev, send.data_ptr(), rows_ptr.value, total_send * row_bytes))
            lib.check(L.dbx_device_free(dev, rows_ptr))
            recv, recv_counts = all_to_all_rows(send, send_counts, row_bytes)
            torch.cuda.current_stream().synchronize()
            fin.merge_rows(recv.data_ptr(), sum(recv_counts))
        return fin.on_finish(out_mem)

    def step_device():
        if not use_peer:
            part.reset()
        fin.reset()
        part.transform(dblock)
        out = exchange_and_finish(abi.MEM_DEVICE)
        # read the kernel's event pair only now: asking earlier blocks the host until the kernel
        # has finished and exposes the launch latency of everything behind it
        kernel_ms.append(last_kernel_ms())
        rows_out = out[0].num_rows
        L.dbx_block_release(C.byref(out[0]))
        return rows_out

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        groups = step_device()
    kernel_ms.clear()
    barrier()
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()
    launches0 = L.dbx_kernel_launch_count()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record(part_stream)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        groups = step_device()
    ev1.record(fin_stream)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    launches = L.dbx_kernel_launch_count() - launches0
    step_ms = max(dev_ms, 0.0) / args.steps
    t = torch.tensor([step_ms, wall * 1e3 / args.steps], dtype=torch.float64, device=f"cuda:{dev}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, wall_ms = t.tolist()
    k_ms = sum(kernel_ms) / max(1, len(kernel_ms))

    # ---- e2e: host (pinned) columns pushed through the operator API, result pulled to the host
    e2e = None
    if not args.no_e2e:
        e_rows = min(n, args.e2e_rows // world if args.e2e_rows else n)
        try:
            import psutil
            avail = psutil.virtual_memory().available
            while e_rows * 24 > 0.5 * avail and e_rows > 1_000_000:
                e_rows //= 2
        except Exception:
            pass
        hp = []
        for i in range(3):
            p = C.c_void_p()
            lib.check(L.dbx_host_alloc(e_rows * 8, C.byref(p)))
            lib.check(L.dbx_memcpy_d2h(dev, p, bufs[i].ptr, e_rows * 8))
            hp.append(p)
        nd = [np.int64, np.int64, np.float64]
        harr = [np.ctypeslib.as_array(C.cast(hp[i], C.POINTER(C.c_int64 if i < 2 else C.c_double)), shape=(e_rows,)) for i in range(3)]
        hblock = DataBlock([Column.from_data(harr[0]), Column.from_data(harr[1]), Column.from_data(harr[2])], e_rows)
        hblocks = hblock.split_by_rows(args.block_rows)

        def step_host():
            if not use_peer:
                part.reset()
            fin.reset()
            for b in hblocks:
                part.transform(b)
            out = exchange_and_finish(abi.MEM_HOST)
            return out[0]

        res = None
        for _ in range(max(1, args.warmup - 1)):
            res = step_host()
        barrier()
        e_steps = max(1, min(args.steps, 3))
        ev0.record(part_stream)
        t0 = time.perf_counter()
        for _ in range(e_steps):
            res = step_host()
        ev1.record(fin_stream)
        barrier()
        e_wall_ms = (time.perf_counter() - t0) * 1e3 / e_steps
        te = torch.tensor([e_wall_ms], dtype=torch.float64, device=f"cuda:{dev}")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e_wall_ms = te.item()
        d2h = sum(c.data.nbytes for c in res.columns)
        e2e = {"value": (e_rows * world) / (e_wall_ms * 1e-3), "unit": "rows/s", "h2d_bytes_per_step": int(e_rows * 24),
               "d2h_bytes_per_step": int(d2h), "rows": int(e_rows * world), "block_rows": args.block_rows,
               "ms_per_step": e_wall_ms, "timing": "host wall clock around push..pull incl. stream sync, max over ranks"}
        for p in hp:
            L.dbx_host_free(p)

    knn = None
    if xchg is not None:
        barrier()
        xchg.close()
    if not args.no_knn:
        part.close()
        fin.close()
        for b_ in bufs:
            b_.free()
        knn = run_knn(args, L, dev, rank, world, barrier)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only)
    cpu = None
    if world == 1 and not args.no_cpu:
        from oracle import oracle as orc
        threads = len(os.sched_getaffinity(0))
        cn = args.cpu_rows
        k = orc.synth_fill(0, SEEDS[0], N_KEYS, 0, cn)
        v = orc.synth_fill(1, SEEDS[1], 0, 0, cn)
        x = orc.synth_fill(2, SEEDS[2], 20, 0, cn)
        cblk = DataBlock([Column.from_data(k), Column.from_data(v), Column.from_data(x)])
        cp = params.to_c(filt)
        orc.filter_group_agg(cblk, cp, threads=threads)
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            orc.filter_group_agg(cblk, cp, threads=threads)
        cdt = (time.perf_counter() - t0) / reps
        cpu = {"value": cn / cdt, "unit": "rows/s", "cores": threads, "kind": "port",
               "sample": f"first {cn} rows of the same columns, reference-algorithm C/OpenMP restatement (oracle), {reps} reps"}

    peak, peak_src = peaks()
    achieved = BYTES_PER_ROW * n / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    line = {
        "metric": METRIC, "value": total_rows / (step_ms * 1e-3), "unit": "rows/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "int64", "data": "synthetic",
        "config": {"workload": "configs[1]: filter(v%3=0) + hash-agg sum(v),count(v),avg(x) GROUP BY k; 1e6 int64 keys",
                   "rows": total_rows, "rows_per_gpu": n, "groups_out": int(groups) * (1 if world == 1 else world),
                   "columns": "k:int64 v:int64 x:float64", "l2": "inputs (24 B/row x rows) far larger than the 126 MB L2",
                   "timing": "CUDA events on the operators' stream, max over ranks; wall_ms_per_step alongside",
                   "parallelism": f"row-range x{world}" + ("" if world == 1 else (" + peer-memory (NVLink) scatter of partial groups" if use_peer else " + NCCL all-to-all of partial groups"))},
        "wall_ms_per_step": wall_ms, "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": NCU_TRAFFIC_PER_LAUNCH, "traffic_note": "dram read+write per 2^28-row launch from profiles/r01b_filter_group_agg_ncu_full.csv (ncu --set full)",
                     "achieved_per_launch_bytes": BYTES_PER_ROW * min(n, 1 << 28), "kernel": "filter_group_agg_kernel<3,FAST=1,INDIRECT=0,BULK=0>", "kernel_ms": k_ms,
                     "algorithmic_bytes_per_row": BYTES_PER_ROW, "peak_source": peak_src},
        "cpu_baseline": cpu, "e2e": e2e, "knn": knn,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="dbx", choices=["dbx", "reference"])
    ap.add_argument("--rows", type=int, default=1_000_000_000)
    ap.add_argument("--e2e-rows", type=int, default=0, help="0 = same as --rows")
    ap.add_argument("--block-rows", type=int, default=1 << 22, help="rows per pushed host block in the e2e leg (max_block_size)")
    ap.add_argument("--cpu-rows", type=int, default=50_000_000)
    ap.add_argument("--keys", type=int, default=N_KEYS, help="distinct group keys (the named config uses 1e6)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-knn", action="store_true", help="skip the kNN leg (second half of BASELINE.json's metric)")
    ap.add_argument("--knn-rows", type=int, default=10_000_000)
    ap.add_argument("--knn-dim", type=int, default=768)
    ap.add_argument("--knn-queries", type=int, default=1024)
    ap.add_argument("--knn-k", type=int, default=10)
    ap.add_argument("--knn-cpu-rows", type=int, default=1_000_000, help="corpus rows of the CPU sample in the reference arm")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_dbx(args)


if __name__ == "__main__":
    main()
