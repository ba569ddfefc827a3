#!/usr/bin/env python
"""bench.py -- headline benchmark: queries/sec of exact flat top-k search (k=100, d=768).

  python bench.py --gpus 1 --steps K --warmup W        # C2: 4M x 768 corpus, batch 4096, 1 B200
  torchrun ... bench.py --gpus N ...                    # C4: 16M x 768 row-sharded over N GPUs, batch 8192
  python bench.py --impl reference ...                  # the reference's CPU flat search on the host cores

A "step" is one ``index.search`` of one query batch against the resident corpus -- the window the
reference times (retrieve/retrieve_faiss.py:69-72).  One JSON line on stdout (rank 0).

  value     q/s, queries already resident in HBM, results left in HBM (CUDA events, max over ranks)
  e2e       q/s through the public API with pinned HOST buffers (H2D of queries, D2H of D and I inside)
  roofline  dominant kernel (K2 tcgen05 scoring main pass): 2*B*N*d flop / its CUDA-event duration
  cpu_baseline  the oracle (FAISS restatement: sgemm + heap) on the host cores, bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D_MODEL = 768
K = 100
METRIC_NAME = "queries/sec @k=100, 768-d, exact flat inner-product top-k"


def workload(n_gpus, args):
    if n_gpus == 1:
        rows, batch, name = 4_000_000, 4096, "C2: 4M x 768 fp32 corpus, batch 4096, k=100, IndexFlatIP, 1 GPU"
    else:
        rows, batch = 16_000_000, 8192
        name = f"C4: 16M x 768 fp32 corpus row-sharded over {n_gpus} GPUs, batch 8192, k=100, all-gather + device merge"
    if args.rows:
        rows = args.rows
        name += f" [rows overridden to {rows}]"
    if args.batch:
        batch = args.batch
        name += f" [batch overridden to {batch}]"
    return rows, batch, name


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            m = json.load(f)
        return {"tf_burst": m["bf16_tflops"], "tf_sust": m.get("bf16_tflops_sustained", m["bf16_tflops"]),
                "hbm": m["hbm_gbs"], "src": "measured"}
    return {"tf_burst": 1590.0, "tf_sust": 1400.0, "hbm": 6650.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.lines, self.proc, self.first, self.last = gpu_index, [], None, 0, None

    def mark(self):
        """Samples taken from now on are the ones reported (call right before the timed region)."""
        self.first = len(self.lines)
        self.last = None

    def end(self):
        """... up to now (call right after the timed region; one trailing sample is let in)."""
        time.sleep(0.06)
        self.last = len(self.lines)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[self.first:self.last]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        # the median over samples taken under load (idle samples at the edges sit at the floor clock)
        loaded = [c for c in sm if mx and c > 0.4 * mx] or sm
        med = loaded[len(loaded) // 2] if loaded else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_qps(xb_sample, xq_sample, rows_full, k, steps=1, warmup=0):
    """Time the oracle's FAISS-restatement BLAS path on all host threads over a bounded sample and
    scale by the row ratio (flat search is linear in the number of rows)."""
    from oracle import cpu_flat as oracle
    for _ in range(warmup):
        oracle.search_blas(xb_sample, xq_sample[:64], k, 0)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle.search_blas(xb_sample, xq_sample, k, 0)
    dt = (time.perf_counter() - t0) / steps
    qps_sample = xq_sample.shape[0] / dt
    return qps_sample * xb_sample.shape[0] / rows_full, dt


def host_threads():
    try:
        import threadpoolctl
        n = max([p["num_threads"] for p in threadpoolctl.threadpool_info()] or [os.cpu_count()])
        return int(n)
    except Exception:
        return os.cpu_count()


def run_reference(args):
    """--impl reference: FAISS CPU flat search (the oracle restatement; real faiss is not installable
    here: un-vendored, un-pinned, no network) on the box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun pins OMP_NUM_THREADS=1 for its workers; the reference arm is meant to use every host core
    if os.environ.get("OMP_NUM_THREADS", "") in ("", "1"):
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    import numpy as np
    rows, batch, name = workload(args.gpus, args)
    sample_rows, sample_q = min(rows, 400_000), min(batch, 1024)
    rng = np.random.default_rng(1234)
    xb = rng.standard_normal((sample_rows, D_MODEL), dtype=np.float32)
    xq = np.random.default_rng(4321).standard_normal((sample_q, D_MODEL), dtype=np.float32)
    import numpy  # noqa: F401  (touch BLAS threads before timing)
    kind = "port"
    try:
        import faiss  # noqa: F401
        if not getattr(faiss, "__version__", "").startswith("textreact_b200"):
            kind = "reference"
    except Exception:
        faiss = None
    if kind == "reference":
        index = faiss.IndexFlatIP(D_MODEL)
        index.add(xb)
        for _ in range(args.warmup):
            index.search(xq[:64], K)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            index.search(xq, K)
        dt = (time.perf_counter() - t0) / args.steps
        qps = sample_q / dt * sample_rows / rows
        cores = faiss.omp_get_max_threads()
    else:
        qps, dt = cpu_reference_qps(xb, xq, rows, K, steps=args.steps, warmup=min(args.warmup, 1))
        cores = host_threads()
    sample = f"{sample_q} queries x {sample_rows} rows per step; q/s scaled by {sample_rows}/{rows} rows"
    out = {"impl": "reference", "metric": METRIC_NAME, "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 * (batch / sample_q) * (rows / sample_rows),
           "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": name, "rows": rows, "batch": batch, "k": K, "d": D_MODEL},
           "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


def stage(msg):
    """Progress line on stderr (rank-tagged): a stalled run shows where it stopped."""
    if os.environ.get("TRX_BENCH_QUIET"):
        return
    print(f"[bench rank {os.environ.get('RANK', '0')} +{time.perf_counter() - T_START:7.1f}s] {msg}",
          file=sys.stderr, flush=True)


T_START = time.perf_counter()
# Only the JSON line goes to the real stdout: libraries (NCCL prints its version there when NCCL_DEBUG is
# VERSION/WARN) are pointed at stderr for the whole run.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(obj):
    _REAL_STDOUT.write(json.dumps(obj) + "\n")
    _REAL_STDOUT.flush()


def main():
    import faulthandler
    wd = float(os.environ.get("TRX_BENCH_WATCHDOG", "900"))
    if wd > 0:   # dump every thread's stack and exit if the run is still alive after `wd` seconds (0 disables)
        faulthandler.dump_traceback_later(wd, exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=0, help="override corpus rows (debug)")
    ap.add_argument("--batch", type=int, default=0, help="override query batch (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="sharded", choices=["sharded", "replicas"],
                    help="N > 1 only: row-sharded corpus (north_star's contract, default) or the whole corpus on every "
                         "GPU with the queries split (measurement beside it)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import textreact_b200 as trx
    from textreact_b200.sharded import ShardedIndexFlat, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    n_gpus = world
    assert torch.cuda.is_available(), "bench.py needs a GPU: textreact_b200 has no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    rows, batch, name = workload(n_gpus, args)
    lo, hi = shard_bounds(rows, world, rank)
    replicas = world > 1 and args.mode == "replicas"
    if replicas:
        lo, hi = 0, rows
        name += " [REPLICAS mode: every GPU holds all rows, queries split]"
    stage(f"process group up; shard rows [{lo}, {hi}) batch {batch}")

    # ---- corpus: dist G (iid N(0,1)), generated on the device per shard, seeded ------------------
    ridx = None
    if replicas:
        from textreact_b200.sharded import ReplicatedIndexFlat
        ridx = ReplicatedIndexFlat(D_MODEL, trx.METRIC_INNER_PRODUCT, device=local_rank)
        sidx, local = None, ridx.local
    elif world > 1:
        sidx = ShardedIndexFlat(D_MODEL, trx.METRIC_INNER_PRODUCT, device=local_rank)
        local = sidx.local
    else:
        sidx = None
        local = trx.IndexFlatIP(D_MODEL, device=local_rank)
    local.reserve(hi - lo)
    for key, env in (("target_candidates", "TRX_TARGET"), ("sample_rate", "TRX_SAMPLE_RATE"), ("path", "TRX_PATH"), ("thr_bias", "TRX_THR_BIAS"),
                     ("umma_pair", "TRX_UMMA_PAIR")):
        if os.environ.get(env):          # tuning knobs for experiments; defaults are what is reported
            local.set_option(key, float(os.environ[env]))
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + (0 if replicas else rank))
    chunk = 500_000
    first_rows = None
    for c0 in range(lo, hi, chunk):
        c1 = min(hi, c0 + chunk)
        x = torch.randn((c1 - c0, D_MODEL), generator=gen, device=dev, dtype=torch.float32)
        local.add(x)
        if first_rows is None and rank == 0:
            first_rows = x[:400_000].cpu().numpy()
        del x
    if sidx is not None:
        local.set_id_offset(lo)
        sidx._lo, sidx._ntotal_global = lo, rows
    index = ridx if ridx is not None else (sidx if sidx is not None else local)
    torch.cuda.synchronize()
    stage("corpus resident")

    nb = args.steps + args.warmup
    qgen = torch.Generator(device=dev)
    qgen.manual_seed(4321)
    queries = [torch.randn((batch, D_MODEL), generator=qgen, device=dev, dtype=torch.float32) for _ in range(min(nb, 8))]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, drain=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        if drain is not None:
            drain()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    # ---- device-resident throughput ---------------------------------------------------------------
    # N > 1: the exchange (all-gather + merge) of a step completes before the next step starts.  Leaving it pending on a
    # side stream while the next local search runs (ShardedIndexFlat.search_async) was measured SLOWER (93 vs 72 ms per
    # step at N = 2): the NCCL kernel of rank A takes SMs and spins for rank B, whose own NCCL kernel is queued behind its
    # persistent, statically partitioned K2 -- the displaced K2 CTAs of A then run their whole share late.
    def step_dev(i):
        index.search(queries[i % len(queries)], K)

    def drain():
        pass

    # nvidia-smi needs a few hundred ms before its first sample: started ahead of the warm-up steps (same load) so that
    # the timed region is covered from its first step
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    for i in range(args.warmup):
        step_dev(i)
        drain()
        torch.cuda.synchronize()
        stage(f"warm-up step {i} done")
    launches0 = local.stats()["launches"]
    sampler.mark()
    ms = timed(step_dev, args.steps, drain)
    if rank == 0:
        sampler.end()
    clocks = sampler.stop() if rank == 0 else None
    launches = local.stats()["launches"] - launches0
    qps = batch * args.steps / (ms * 1e-3)
    stage(f"device-resident timing done: {ms / args.steps:.2f} ms/step")

    # ---- end to end through the public API with pinned host buffers -------------------------------
    hq = [torch.empty((batch, D_MODEL), dtype=torch.float32).pin_memory() for _ in range(2)]
    for i, h in enumerate(hq):
        h.copy_(queries[i % len(queries)])
    hD = torch.empty((batch, K), dtype=torch.float32).pin_memory()
    hI = torch.empty((batch, K), dtype=torch.int64).pin_memory()

    if world == 1:
        def step_e2e(i):
            local.search(hq[i % 2].numpy(), K, D=hD.numpy(), I=hI.numpy())
    else:
        dq = torch.empty((batch, D_MODEL), dtype=torch.float32, device=dev)

        def step_e2e(i):
            dq.copy_(hq[i % 2], non_blocking=True)           # H2D of the replicated query batch
            Dm, Im = index.search(dq, K)
            hD.copy_(Dm, non_blocking=True); hI.copy_(Im, non_blocking=True)   # D2H of the merged result
            torch.cuda.current_stream().synchronize()
    for i in range(2):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    ms_e2e_dev = timed(step_e2e, args.steps)
    wall_e2e = time.perf_counter() - t0
    qps_e2e = batch * args.steps / (ms_e2e_dev * 1e-3)
    stage(f"end-to-end timing done: {ms_e2e_dev / args.steps:.2f} ms/step")

    # ---- roofline of the dominant kernel (K2 main pass), CUDA events inside the library -----------
    local.set_option("timing", 1)
    pre_ms, tot_ms = [], []
    for i in range(min(args.steps, 5)):
        step_dev(i)
        drain()
        s = local.stats()
        pre_ms.append(s["last_prefilter_ms"]); tot_ms.append(s["last_total_ms"])
    local.set_option("timing", 0)
    st = local.stats()
    pk = peaks()

    # ---- multi-GPU: where the step goes -- per-rank local search time, and the exchange (all-gather + K5) alone
    multi = None
    if world > 1 and sidx is not None:
        mine = torch.tensor([sum(tot_ms) / len(tot_ms), sum(pre_ms) / len(pre_ms)], device=dev)
        allr = torch.empty((world, 2), device=dev)
        dist.all_gather_into_tensor(allr, mine)
        Dl, Il = local.search(queries[0], K)

        def exchange(i):
            sidx.exchange(Dl, Il)
        mode = sidx._exchange_mode
        ex_ms = {}
        for m in ([mode, "nccl"] if mode == "peer" else [mode]):      # the mode in use, and NCCL beside it
            sidx._exchange_mode = m
            for i in range(3):
                exchange(i)
            ex_ms[m] = timed(exchange, 10) / 10
        sidx._exchange_mode = mode
        multi = {"local_ms_per_rank": [round(float(v), 3) for v in allr[:, 0].tolist()],
                 "k2_ms_per_rank": [round(float(v), 3) for v in allr[:, 1].tolist()],
                 "exchange_mode": mode, "exchange_ms": ex_ms[mode], "exchange_ms_by_mode": ex_ms,
                 "exchange": {"peer": "ONE kernel: all-gather fused into the k-way merge over NVLink peer memory "
                                      "(CUDA IPC export buffers, flag protocol, no NCCL in the data path)",
                              "nccl": "NCCL all-gather of per-shard (D, I) + device k-way merge"}[mode]}

    kern_ms = sum(pre_ms) / len(pre_ms)
    flops = 2.0 * (batch / world if replicas else batch) * (hi - lo) * D_MODEL     # per rank, per K2 launch
    achieved_tf = flops / (kern_ms * 1e-3) / 1e12 if kern_ms > 0 else 0.0
    traffic = None   # dram bytes per launch of the dominant kernel: from the committed ncu --set full capture (C2 only)
    tp = os.path.join(ROOT, "profiles", "k2_traffic.json")
    if world == 1 and not args.rows and not args.batch and os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get("traffic_bytes_per_launch")
    roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                "frac": achieved_tf / pk["tf_sust"], "traffic": traffic,
                "traffic_source": "profiles/k2_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum)" if traffic else None,
                "kernel": "k2_umma_kernel<THRESH> (tcgen05 bf16 scoring + fused threshold select)",
                "kernel_ms": kern_ms, "batch_ms_on_device": sum(tot_ms) / len(tot_ms),
                "peak_kind": f"{pk['src']} cuBLAS bf16 sustained; burst {pk['tf_burst']}",
                "frac_of_burst": achieved_tf / pk["tf_burst"],
                "algorithmic_flops_per_launch": flops}

    # ---- CPU baseline (rank 0, N=1 only) ------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        xq_s = queries[0][:1024].cpu().numpy()
        v, dt = cpu_reference_qps(first_rows, xq_s, rows, K, steps=1, warmup=1)
        cpu = {"value": v, "unit": "queries/s", "cores": host_threads(), "kind": "port",
               "sample": f"{xq_s.shape[0]} queries x {first_rows.shape[0]} rows ({dt:.2f} s); q/s scaled by "
                         f"{first_rows.shape[0]}/{rows} rows"}

    if rank == 0:
        out = {"metric": METRIC_NAME, "value": qps, "unit": "queries/s", "n_gpus": n_gpus, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
               "scaling": "strong" if n_gpus > 1 else "weak", "vs_baseline": None, "dtype": "bf16",
               "data": "synthetic",
               "config": {"workload": name, "rows": rows, "rows_per_gpu": hi - lo, "batch": batch, "k": K, "d": D_MODEL,
                          "dtype_detail": "bf16 tcgen05 prefilter (fp32 accumulate) + exact fp32 rescore with certificate: "
                                          "results are the fp32 flat-search answer",
                          "dist": "iid N(0,1) fp32, seeds 1234+rank / 4321", "l2_policy": "inputs_exceed_l2 "
                          f"(bf16 corpus shard {2 * (hi - lo) * D_MODEL / 1e9:.1f} GB >> 126 MB L2)",
                          "scored_pairs_per_s": qps * rows,
                          "exchange": None if sidx is None else f"{sidx._exchange_mode} exchange after every local search "
                                      "(peer = gather fused into the merge kernel over NVLink peer memory)"},
               "clocks": clocks,
               "e2e": {"value": qps_e2e, "unit": "queries/s", "h2d_bytes_per_step": batch * D_MODEL * 4,
                       "d2h_bytes_per_step": batch * K * 12, "ms_per_step": ms_e2e_dev / args.steps,
                       "wall_ms_per_step": wall_e2e * 1e3 / args.steps},
               "gpu_launches": int(launches),
               "roofline": roofline,
               "multi_gpu": multi,
               "cpu_baseline": cpu,
               "engine": {k: st[k] for k in ("queries", "queries_exact", "queries_uncert", "queries_overflow",
                                             "rescored", "candidates", "last_path")}}
        emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
