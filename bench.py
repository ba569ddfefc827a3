#!/usr/bin/env python
"""bench.py -- headline benchmark: queries/sec of exact flat top-k search (k=100, d=768).

  python bench.py --gpus 1 --steps K --warmup W        # C2: 4M x 768 corpus, batch 4096, 1 B200
  torchrun ... bench.py --gpus N ...                    # C4: 16M x 768 row-sharded over N GPUs, batch 8192
  python bench.py --impl reference ...                  # the reference's CPU flat search on the host cores

A "step" is one ``index.search`` of one query batch against the resident corpus -- the window the
reference times (retrieve/retrieve_faiss.py:69-72).  One JSON line on stdout (rank 0).

  value     q/s, queries already resident in HBM, results left in HBM (CUDA events, max over ranks)
  e2e       q/s through the public API with pinned HOST buffers (H2D of queries, D2H of D and I inside)
  roofline  dominant kernel (K2 tcgen05 scoring main pass): 2*B*N*d flop / its CUDA-event duration, the events
            recorded inside trx_search on the launching stream DURING the timed steps (trx_stats sums)
  parity_checked  after the timed region: 32 queries of the batch against a float64 top-k computed on the
            device over the same (regenerated) rows, north_star rule (ids identical wherever the fp64 gap at a
            rank exceeds 1e-5 relative, scores within 1e-5 |q||x|); N > 1: over all shards, for both exchanges
  cpu_baseline  the oracle (FAISS restatement: sgemm + k-heap) on the host cores: 1,024 queries x the FULL corpus

N = 1 adds, on the resident C2 corpus unless said otherwise (driver-witnessed legs of BASELINE.json's configs):
  c5           small-batch latency (configs[4]): batch 1 / 8 / 64, 200 calls each, p50 / p99 of the whole call with
               host buffers, and the scoring kernel against the HBM roofline (2*N*d bytes per batch)
  c3           gold-removed mode (configs[2]): C2 + per-query exclusion mask, q/s and excluded rows returned (0)
  dists        C2 on Dist U (clustered unit-norm) and Dist A (scores ascend with the row id): q/s + fallback counters
  strong_base  16M x 768, batch 8192 on ONE GPU: the origin of the N = 2/4/8 strong-scaling curve (configs[3])
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
CHUNK = 500_000          # corpus rows generated / added / re-generated per piece
NCENT = 4096             # Dist U centroids
PARITY_Q = 32            # queries checked against the float64 arbiter after the timed region
PARITY_EXTRA = 8         # fp64 ranks kept beyond k (tie groups at the boundary)


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
                "hbm": m["hbm_gbs"], "src": "MEASURED_PEAKS.json"}
    # /opt/skills/guides/B200_PROFILING.md fallbacks
    return {"tf_burst": 1590.0, "tf_sust": 1400.0, "hbm": 6650.0, "src": "B200_PROFILING.md fallback"}


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


def host_threads():
    try:
        import threadpoolctl
        n = max([p["num_threads"] for p in threadpoolctl.threadpool_info()] or [os.cpu_count()])
        return int(n)
    except Exception:
        return os.cpu_count()


def host_blas():
    """Which BLAS the CPU arm's sgemm runs on (torch.mm -> MKL when torch was built with it)."""
    try:
        import torch
        cfg = torch.__config__.show()
        return "torch.mm (" + ("MKL" if "USE_MKL=ON" in cfg or "mkl" in cfg.lower() else "torch default BLAS") + ")"
    except Exception:
        return "numpy (OpenBLAS)"


def host_corpus_rows(rows, reserve_gb=8.0):
    """Rows of the fp32 corpus that fit the host's free memory (the CPU arm wants the FULL corpus)."""
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 32 << 30
    fit = int(max(0.0, avail - reserve_gb * (1 << 30)) // (D_MODEL * 4))
    return max(100_000, min(rows, fit))


def cpu_flat_search_timed(xb, xq, k, steps, warmup):
    """The CPU arm: oracle restatement of FAISS's BLAS path (sgemm blocks + per-query k-heap, all host threads), or the
    real faiss when it is importable.  Returns (seconds per step, kind, cores)."""
    kind = "port"
    try:
        import faiss  # noqa: F401
        if not str(getattr(faiss, "__version__", "")).startswith("textreact_b200"):
            kind = "reference"
    except Exception:
        faiss = None
    if kind == "reference":
        index = faiss.IndexFlatIP(xb.shape[1])
        index.add(xb)
        run = lambda q: index.search(q, k)           # noqa: E731
        cores = faiss.omp_get_max_threads()
    else:
        from oracle import cpu_flat as oracle
        run = lambda q: oracle.search_blas(xb, q, k, 0, gemm="torch")     # noqa: E731
        cores = host_threads()
    for _ in range(warmup):
        run(xq[:64])
    t0 = time.perf_counter()
    for _ in range(steps):
        run(xq)
    return (time.perf_counter() - t0) / steps, kind, cores


def run_reference(args):
    """--impl reference: the reference's CPU flat search (real faiss when importable; else the oracle restatement --
    faiss is un-vendored, un-pinned and unobtainable offline, profiles/round2_a_try_faiss_on_gpu_box.log) on the box's host
    cores, on the FULL corpus of the configuration.  A step is a bounded sample of one batch's QUERIES (1,024 of 4,096
    at N = 1; scaled down with the corpus size at N > 1 so a step stays ~6 s); nothing is extrapolated in rows, and
    ms_per_step is the measured time of what was run."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun pins OMP_NUM_THREADS=1 for its workers; the reference arm is meant to use every host core
    if os.environ.get("OMP_NUM_THREADS", "") in ("", "1"):
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    os.environ.setdefault("MKL_NUM_THREADS", os.environ["OMP_NUM_THREADS"])
    import numpy as np
    import torch
    torch.set_num_threads(int(os.environ["OMP_NUM_THREADS"]))
    rows, batch, name = workload(args.gpus, args)
    host_rows = host_corpus_rows(rows)
    sample_q = args.ref_queries or max(64, min(batch, int(1024 * 4_000_000 / max(rows, 1))))
    # Host corpus, Dist G.  CPU randn is the slow part of this arm (~10 s per million rows), so only the first
    # BASE rows are drawn; every further block is a column-permuted, sign-flipped copy of it -- still iid N(0,1) rows,
    # all distinct, and the search work (sgemm + heap updates) is what it would be on fresh draws.
    BASE = 2_000_000
    g = torch.Generator().manual_seed(1234)
    xb = torch.empty((host_rows, D_MODEL), dtype=torch.float32)
    for c0 in range(0, min(host_rows, BASE), CHUNK):
        c1 = min(host_rows, BASE, c0 + CHUNK)
        torch.randn((c1 - c0, D_MODEL), generator=g, out=xb[c0:c1])
    for b0 in range(BASE, host_rows, BASE):
        b1 = min(host_rows, b0 + BASE)
        perm = torch.randperm(D_MODEL, generator=g)
        sign = (torch.randint(0, 2, (D_MODEL,), generator=g) * 2 - 1).to(torch.float32)
        for c0 in range(b0, b1, CHUNK):
            c1 = min(b1, c0 + CHUNK)
            torch.mul(xb[c0 - b0:c1 - b0][:, perm], sign, out=xb[c0:c1])
    xb = xb.numpy()
    xq = torch.randn((sample_q, D_MODEL), generator=torch.Generator().manual_seed(4321)).numpy()
    dt, kind, cores = cpu_flat_search_timed(xb, xq, K, args.steps, min(args.warmup, 2))
    qps = sample_q / dt * (host_rows / rows)
    extrap = {"queries": batch / sample_q}
    if host_rows < rows:
        extrap["rows"] = rows / host_rows       # only when the host cannot hold the corpus; q/s scaled by it
    sample = (f"{sample_q} of the {batch} queries of a batch x {host_rows} of {rows} corpus rows per step "
              f"({dt:.2f} s measured); BLAS: {host_blas() if kind == 'port' else 'faiss'}; host corpus: N(0,1), rows beyond the "
              f"first {BASE} are column-permuted sign-flipped copies of them (draw time only)")
    out = {"impl": "reference", "metric": METRIC_NAME, "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": name, "rows": rows, "batch": batch, "k": K, "d": D_MODEL,
                      "step": f"{sample_q} queries (a bounded sample of one batch) against the corpus"},
           "extrapolated": extrap,
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


# ---- synthetic corpora, generated on the device piece by piece and re-generated identically by the
# ---- float64 arbiter (same seed, same sequence of generator calls) --------------------------------------------
def corpus_pieces(dist, lo, hi, rows_total, dev, seed):
    """Yields (first_row, fp32 rows) covering rows [lo, hi) of the corpus named `dist`."""
    import torch
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    cent = dirn = None
    if dist == "U":
        cg = torch.Generator(device=dev); cg.manual_seed(777)
        cent = torch.randn((NCENT, D_MODEL), generator=cg, device=dev)
    if dist == "A":
        dg = torch.Generator(device=dev); dg.manual_seed(99)
        dirn = torch.randn((D_MODEL,), generator=dg, device=dev)
        dirn *= (D_MODEL ** 0.5) / dirn.norm()
    for c0 in range(lo, hi, CHUNK):
        c1 = min(hi, c0 + CHUNK)
        x = torch.randn((c1 - c0, D_MODEL), generator=gen, device=dev, dtype=torch.float32)
        if dist == "U":      # 4096 Gaussian centroids + 0.3 noise, rows L2-normalised
            pick = torch.randint(0, NCENT, (c1 - c0,), generator=gen, device=dev)
            x = cent[pick] + 0.3 * x
            x /= x.norm(dim=1, keepdim=True)
        elif dist == "A":    # scores ascend with the row id for every query: a drift along one direction
            ramp = torch.arange(c0, c1, device=dev, dtype=torch.float32) / float(rows_total)
            x += 0.5 * ramp[:, None] * dirn[None, :]
        yield c0, x
        del x


def make_queries(dist, n, dev, seed=4321):
    import torch
    g = torch.Generator(device=dev); g.manual_seed(seed)
    q = torch.randn((n, D_MODEL), generator=g, device=dev, dtype=torch.float32)
    if dist == "U":
        cg = torch.Generator(device=dev); cg.manual_seed(777)
        cent = torch.randn((NCENT, D_MODEL), generator=cg, device=dev)
        pick = torch.randint(0, NCENT, (n,), generator=g, device=dev)
        q = cent[pick] + 0.3 * q
        q /= q.norm(dim=1, keepdim=True)
    elif dist == "A":
        dg = torch.Generator(device=dev); dg.manual_seed(99)
        dirn = torch.randn((D_MODEL,), generator=dg, device=dev)
        dirn *= (D_MODEL ** 0.5) / dirn.norm()
        q += 0.5 * dirn[None, :]
    return q


DIST_TEXT = {"G": "Dist G: iid N(0,1) fp32", "U": "Dist U: 4096 Gaussian centroids + 0.3 N(0,1), rows L2-normalised",
             "A": "Dist A: N(0,1) + 0.5 (row/N) u, queries N(0,1) + 0.5 u (|u|^2 = d): scores ascend with the row id"}


def fp64_topk_local(dist, lo, hi, rows_total, dev, seed, xq, kk):
    """float64 top-kk of xq against rows [lo, hi) (re-generated), with GLOBAL ids and the rows' norms."""
    import torch
    q64 = xq.double()
    bestD = torch.full((xq.shape[0], 0), 0.0, dtype=torch.float64, device=dev)
    bestI = torch.zeros((xq.shape[0], 0), dtype=torch.int64, device=dev)
    bestN = torch.zeros((xq.shape[0], 0), dtype=torch.float64, device=dev)
    for c0, x in corpus_pieces(dist, lo, hi, rows_total, dev, seed):
        for s0 in range(0, x.shape[0], 250_000):          # 250K x 768 fp64 = 1.5 GB at a time
            x64 = x[s0:s0 + 250_000].double()
            s = q64 @ x64.T
            d, i = torch.topk(s, min(kk, s.shape[1]), dim=1)
            nrm = (x64 * x64).sum(1).sqrt()
            bestD = torch.cat([bestD, d], 1); bestI = torch.cat([bestI, i + (c0 + s0)], 1)
            bestN = torch.cat([bestN, nrm[i]], 1)
            if bestD.shape[1] > kk:
                d2, j = torch.topk(bestD, kk, dim=1)
                bestD, bestI, bestN = d2, torch.gather(bestI, 1, j), torch.gather(bestN, 1, j)
            del x64, s
    return bestD, bestI, bestN


def pct(v, p):
    v = sorted(v)
    return v[min(len(v) - 1, int(round(p / 100.0 * (len(v) - 1))))]


def main():
    import faulthandler
    wd = float(os.environ.get("TRX_BENCH_WATCHDOG", "1500"))
    if wd > 0:   # dump every thread's stack and exit if the run is still alive after `wd` seconds (0 disables)
        faulthandler.dump_traceback_later(wd, exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=0, help="override corpus rows (debug)")
    ap.add_argument("--batch", type=int, default=0, help="override query batch (debug)")
    ap.add_argument("--dist", default="G", choices=["G", "U", "A"], help="distribution of the headline leg (SURVEY 8d)")
    ap.add_argument("--legs", default=None,
                    help="comma list of the extra N=1 legs: c5,c3,dists,strong_base,cpu (default: all; 'none')")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-queries", type=int, default=0, help="--impl reference: queries per step (default: bounded)")
    ap.add_argument("--merge", default="slice", choices=["slice", "full"],
                    help="N > 1: 'slice' = the merge is sharded too, rank g ends up with the merged top-k of queries "
                         "[g*B/N, (g+1)*B/N) (1/N of the exchange traffic; every answer exists on one rank); 'full' = every "
                         "rank ends up with every query's result")
    ap.add_argument("--shards", default="even", choices=["calibrated", "even"],
                    help="N > 1: equal shards (default), or sizes proportional to each GPU's measured scoring speed "
                         "(ShardedIndexFlat.calibrate; measured: the intrinsic spread is +-1.5 %, no gain at N = 8)")
    ap.add_argument("--two-phase", type=int, default=1,
                    help="N > 1: 1 = the shards exchange their best prefilter scores first and rescore only what can reach "
                         "the global top-k (default); 0 = every shard computes its full local top-k")
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
    from textreact_b200.check import north_star_rule
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
    legs = set((args.legs if args.legs is not None else "c5,c3,dists,strong_base,cpu").split(",")) - {"none", ""}
    if args.no_cpu_baseline:
        legs.discard("cpu")
    if world > 1 or args.rows or args.batch:
        legs &= {"cpu"} if world == 1 else set()
    pk = peaks()
    settle_s = float(os.environ.get("TRX_BENCH_SETTLE_S", "1.0"))

    rows, batch, name = workload(n_gpus, args)
    replicas = world > 1 and args.mode == "replicas"
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
    # Shards: equal by default.  --shards calibrated: each rank first measures its scoring speed on a throw-away index
    # (1.5 s) and the rows are split in proportion (ShardedIndexFlat.calibrate).
    weights = None
    if sidx is not None:
        sidx._two_phase = bool(args.two_phase)
    if sidx is not None and args.shards == "calibrated":
        weights = sidx.calibrate(batch=batch)
    lo, hi = shard_bounds(rows, world, rank, weights)
    if replicas:
        lo, hi = 0, rows
        name += " [REPLICAS mode: every GPU holds all rows, queries split]"
    stage(f"process group up; shard rows [{lo}, {hi}) batch {batch}" + (f"; speed weights {[round(w, 4) for w in weights]}" if weights else ""))
    want_base = "strong_base" in legs
    local.reserve(16_000_000 if want_base else hi - lo)
    for key, env in (("target_candidates", "TRX_TARGET"), ("sample_rate", "TRX_SAMPLE_RATE"), ("path", "TRX_PATH"), ("thr_bias", "TRX_THR_BIAS"),
                     ("umma_pair", "TRX_UMMA_PAIR"), ("second_pass", "TRX_SECOND_PASS"), ("thr_margin", "TRX_THR_MARGIN")):
        if os.environ.get(env):          # tuning knobs for experiments; defaults are what is reported
            local.set_option(key, float(os.environ[env]))
    seed = 1234 + (0 if replicas else rank)

    def build(d_name, r_lo, r_hi, r_total):
        for _, x in corpus_pieces(d_name, r_lo, r_hi, r_total, dev, seed):
            local.add(x)
        torch.cuda.synchronize()

    build(args.dist, lo, hi, rows)
    if sidx is not None:
        local.set_id_offset(lo)
        sidx._lo, sidx._ntotal_global = lo, rows
    index = ridx if ridx is not None else (sidx if sidx is not None else local)
    stage("corpus resident")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    def settle(fn, min_steps):
        """>= min_steps untimed steps AND >= settle_s seconds of them: the first steps after an idle period run at
        boost clocks the power cap then takes away; what is timed afterwards is the sustained state."""
        for i in range(min_steps):
            t0 = time.perf_counter()
            fn(i)
            torch.cuda.synchronize()
        per = torch.tensor([time.perf_counter() - t0], device=dev)       # the last one: set-up costs are behind it
        if world > 1:       # every rank must run the same number of (collective) steps: agree on the step time
            dist.all_reduce(per, op=dist.ReduceOp.MAX)
        extra = max(0, min(2000, int(settle_s / max(float(per.item()), 1e-4)) + 1 - min_steps))
        for i in range(extra):
            fn(min_steps + i)
            torch.cuda.synchronize()
        return min_steps + extra

    def stage_delta(s0, s1):
        nb = max(1, s1["timed_batches"] - s0["timed_batches"])
        return {k2: (s1["sum_" + k2 + "_ms"] - s0["sum_" + k2 + "_ms"]) / nb for k2 in ("sample", "prefilter", "rescore", "total")}, nb

    def engine_delta(s0, s1):
        return {k2: s1[k2] - s0[k2] for k2 in ("queries", "queries_exact", "queries_second_pass", "queries_uncert",
                                                "queries_overflow", "rescored", "candidates")}

    def throughput_leg(idx, qs, nsteps, bsz, **kw):
        """settle + K timed device-resident steps + the stage times of exactly those steps."""
        def step(i):
            idx.search(qs[i % len(qs)], K, **kw)
        n_settle = settle(step, args.warmup)
        s0 = local.stats()
        ms = timed(step, nsteps)
        s1 = local.stats()
        stages, nb = stage_delta(s0, s1)
        return {"qps": bsz * nsteps / (ms * 1e-3), "ms_per_step": ms / nsteps, "stages_ms": stages,
                "engine": engine_delta(s0, s1), "launches": s1["launches"] - s0["launches"], "settle_steps": n_settle}

    # ---- headline: device-resident throughput -------------------------------------------------------------------
    nq_bufs = min(args.steps + args.warmup, 8)
    queries = [make_queries(args.dist, batch, dev, 4321 + i) for i in range(nq_bufs)]
    # N > 1: the exchange (all-gather + merge) of a step completes before the next step starts (search_async measured
    # slower with the persistent K2).
    sampler = ClockSampler(local_rank)
    if rank == 0:        # nvidia-smi needs a few hundred ms before its first sample: started ahead of the settle steps
        sampler.start()
        time.sleep(0.3)

    merge_kw = {"result": args.merge} if sidx is not None else {}

    def step_dev(i):
        index.search(queries[i % len(queries)], K, **merge_kw)
    n_settle = settle(step_dev, args.warmup)
    stage(f"{n_settle} warm-up / settle steps done")
    s0 = local.stats()
    sampler.mark()
    ms = timed(step_dev, args.steps)
    if rank == 0:
        sampler.end()
    clocks = sampler.stop() if rank == 0 else None
    s1 = local.stats()
    launches = s1["launches"] - s0["launches"]
    stages, timed_batches = stage_delta(s0, s1)
    eng = engine_delta(s0, s1)
    qps = batch * args.steps / (ms * 1e-3)
    stage(f"device-resident timing done: {ms / args.steps:.2f} ms/step; stages {stages}")

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
            Dm, Im = index.search(dq, K, **merge_kw)
            n_out = Dm.shape[0]                               # D2H of the merged result (this rank's slice of it)
            hD[:n_out].copy_(Dm, non_blocking=True); hI[:n_out].copy_(Im, non_blocking=True)
            torch.cuda.current_stream().synchronize()
    settle(step_e2e, 2)
    barrier()
    t0 = time.perf_counter()
    ms_e2e_dev = timed(step_e2e, args.steps)
    wall_e2e = time.perf_counter() - t0
    qps_e2e = batch * args.steps / (ms_e2e_dev * 1e-3)
    stage(f"end-to-end timing done: {ms_e2e_dev / args.steps:.2f} ms/step")

    # ---- parity of what was just timed: 32 queries against a float64 top-k over the same rows -----------------
    def parity(idx_search, d_name, r_lo, r_hi, r_total, qbuf, modes=("default",)):
        kk = K + PARITY_EXTRA
        pq = torch.arange(PARITY_Q, device=dev) * (qbuf.shape[0] // PARITY_Q)     # spread over the batch (and the slices)
        xq = qbuf[pq].contiguous()
        D64, I64, N64 = fp64_topk_local(d_name, r_lo, r_hi, r_total, dev, seed, xq, kk)
        if world > 1 and not replicas:
            gD = torch.empty((world,) + tuple(D64.shape), dtype=D64.dtype, device=dev)
            gI = torch.empty((world,) + tuple(I64.shape), dtype=I64.dtype, device=dev)
            gN = torch.empty((world,) + tuple(N64.shape), dtype=N64.dtype, device=dev)
            dist.all_gather_into_tensor(gD, D64.contiguous()); dist.all_gather_into_tensor(gI, I64.contiguous())
            dist.all_gather_into_tensor(gN, N64.contiguous())
            cD = gD.permute(1, 0, 2).reshape(PARITY_Q, -1); cI = gI.permute(1, 0, 2).reshape(PARITY_Q, -1)
            cN = gN.permute(1, 0, 2).reshape(PARITY_Q, -1)
            D64, j = torch.topk(cD, kk, dim=1)
            I64, N64 = torch.gather(cI, 1, j), torch.gather(cN, 1, j)
        qn = xq.double().norm(dim=1).cpu().numpy()
        out = {"queries": PARITY_Q, "against": f"float64 top-{kk} on the device over the re-generated rows"
                                               + (f" of all {world} shards (all-gathered)" if world > 1 and not replicas else ""),
               "rule": "ids identical wherever the fp64 relative gap at a rank > 1e-5; scores within 1e-5 |q||x|"}
        ok = True
        for m in modes:
            D, I = idx_search(qbuf, m)
            rows = pq
            sel = torch.ones(PARITY_Q, dtype=torch.bool, device=dev)
            if m.endswith("_slice"):       # this rank holds the merged rows [q_lo, q_hi) only
                q_lo, q_hi = sidx.query_slice(qbuf.shape[0])
                sel = (pq >= q_lo) & (pq < q_hi)
                rows = pq[sel] - q_lo
            sel_c = sel.cpu().numpy()
            r = north_star_rule(D[rows].cpu().numpy(), I[rows].cpu().numpy(), D64.cpu().numpy()[sel_c],
                                I64.cpu().numpy()[sel_c], N64.cpu().numpy()[sel_c], qn[sel_c], K)
            if m.endswith("_slice"):
                r["queries_on_this_rank"] = int(sel.sum().item())
            ok = ok and r["ok"]
            if len(modes) == 1:
                out.update(r)
            else:
                out[m] = r
        out["ok"] = ok
        if world > 1:      # a violation seen by any rank fails the line
            t = torch.tensor([1 if ok else 0], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            out["ok"] = bool(int(t.item()))
        return out

    if sidx is not None:
        def search_mode(qbuf, m):
            keep = sidx._exchange_mode
            sidx._exchange_mode = m.replace("_slice", "")
            try:
                return sidx.search(qbuf, K, result="slice" if m.endswith("_slice") else "full")
            finally:
                sidx._exchange_mode = keep
        modes = [sidx._exchange_mode] + (["nccl"] if sidx._exchange_mode == "peer" else [])
        modes += [sidx._exchange_mode + "_slice"]
        parity_checked = parity(search_mode, args.dist, lo, hi, rows, queries[0], tuple(modes))
    else:
        parity_checked = parity(lambda qbuf, m: index.search(qbuf, K), args.dist, lo, hi, rows, queries[0])
    stage(f"parity check done: ok={parity_checked['ok']}")

    # ---- multi-GPU: where the step goes -- per-rank stage times of the timed steps, and the exchange alone
    multi = None
    if world > 1 and sidx is not None:
        mine = torch.tensor([stages["total"], stages["prefilter"], stages["rescore"], stages["sample"]], device=dev)
        allr = torch.empty((world, 4), device=dev)
        dist.all_gather_into_tensor(allr, mine)
        # two-phase local search on / off, alternating in this process (run-to-run differences between launches are
        # larger than the effect): median ms/step of 3 timed regions each, and the rows K4 rescored per query per shard
        ab = {"on": [], "off": []}
        resc = {}
        keep_tp = sidx._two_phase
        for rep in range(3):
            for name, flag in (("on", True), ("off", False)):
                sidx._two_phase = flag
                step_dev(0); step_dev(1)
                r0 = local.stats()["rescored"]
                ab[name].append(timed(step_dev, args.steps) / args.steps)
                resc[name] = (local.stats()["rescored"] - r0) / (args.steps * batch)
        sidx._two_phase = keep_tp
        two_phase_ab = {"ms_per_step_on": sorted(ab["on"])[1], "ms_per_step_off": sorted(ab["off"])[1],
                        "all_ms_on": [round(v, 3) for v in ab["on"]], "all_ms_off": [round(v, 3) for v in ab["off"]],
                        "rescored_rows_per_query_per_shard_on": resc["on"], "rescored_rows_per_query_per_shard_off": resc["off"]}
        # the same steps without the exchange: what the coupling of the ranks (exchange + waiting for the slowest) costs
        def step_local(i):
            local.search(queries[i % len(queries)], K)
        for i in range(3):
            step_local(i)
        local_only_ms = timed(step_local, args.steps) / args.steps
        Dl, Il = local.search(queries[0], K)

        mode = sidx._exchange_mode
        ex_ms = {}
        for m in ([mode + "_slice", mode, "nccl"] if mode == "peer" else [mode + "_slice", mode]):   # in use, and beside it
            sidx._exchange_mode = m.replace("_slice", "")

            def exchange(i, res="slice" if m.endswith("_slice") else "full"):
                sidx.exchange(Dl, Il, result=res)
            for i in range(3):
                exchange(i)
            ex_ms[m] = timed(exchange, 10) / 10
        sidx._exchange_mode = mode
        mode_used = mode + ("_slice" if args.merge == "slice" else "")
        multi = {"shards": {"policy": args.shards, "speed_weights": weights,
                            "rows_per_rank": [shard_bounds(rows, world, r, weights)[1] - shard_bounds(rows, world, r, weights)[0]
                                              for r in range(world)]},
                 "two_phase": bool(sidx._two_phase), "two_phase_ab": two_phase_ab,
                 "local_ms_per_rank": [round(float(v), 3) for v in allr[:, 0].tolist()],
                 "k2_ms_per_rank": [round(float(v), 3) for v in allr[:, 1].tolist()],
                 "k4_ms_per_rank": [round(float(v), 3) for v in allr[:, 2].tolist()],
                 "sample_pass_ms_per_rank": [round(float(v), 3) for v in allr[:, 3].tolist()],
                 "local_only_ms_per_step": local_only_ms,
                 "exchange_mode": mode_used, "exchange_ms": ex_ms[mode_used], "exchange_ms_by_mode": ex_ms,
                 "merge": {"slice": "the merge is sharded: rank g merges (and keeps) queries [g*B/N, (g+1)*B/N)",
                           "full": "every rank merges every query"}[args.merge],
                 "exchange": {"peer": "ONE kernel: all-gather fused into the k-way merge over NVLink peer memory "
                                      "(CUDA IPC export buffers, flag protocol, no NCCL in the data path)",
                              "nccl": "NCCL all-gather of per-shard (D, I) + device k-way merge"}[mode]}

    kern_ms = stages["prefilter"]
    flops = 2.0 * (batch / world if replicas else batch) * (hi - lo) * D_MODEL     # per rank, per K2 launch
    achieved_tf = flops / (kern_ms * 1e-3) / 1e12 if kern_ms > 0 else 0.0
    traffic = None   # dram bytes per launch of the dominant kernel: from the committed ncu --set full capture (C2 only)
    tp = os.path.join(ROOT, "profiles", "k2_traffic.json")
    if world == 1 and not args.rows and not args.batch and os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get("traffic_bytes_per_launch")
    roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": pk["tf_burst"], "unit": "TFLOP/s",
                "frac": achieved_tf / pk["tf_burst"], "traffic": traffic,
                "traffic_source": "profiles/k2_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum)" if traffic else None,
                "kernel": "k2_umma_kernel<THRESH> (tcgen05 bf16 scoring + fused threshold select) + hit scatter",
                "kernel_ms": kern_ms,
                "kernel_ms_source": f"CUDA events around the launch inside trx_search, mean over the {timed_batches} "
                                    "batches of the timed region",
                "stages_ms": stages, "ms_per_step": ms / args.steps,
                "peak_kind": f"{pk['src']}: cuBLAS bf16 burst (SURVEY 8d); sustained {pk['tf_sust']} beside it",
                "frac_of_sustained": achieved_tf / pk["tf_sust"],
                "algorithmic_flops_per_launch": flops}

    out = {"metric": METRIC_NAME, "value": qps, "unit": "queries/s", "n_gpus": n_gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
           "data": "synthetic",
           "config": {"workload": name, "rows": rows, "rows_per_gpu": hi - lo, "batch": batch, "k": K, "d": D_MODEL,
                      "dtype_detail": "bf16 tcgen05 prefilter (fp32 accumulate) + exact fp32 rescore with certificate: "
                                      "results are the fp32 flat-search answer",
                      "dist": f"{DIST_TEXT[args.dist]}, seeds 1234+rank / 4321+i", "l2_policy": "inputs_exceed_l2 "
                      f"(bf16 corpus shard {2 * (hi - lo) * D_MODEL / 1e9:.1f} GB >> 126 MB L2)",
                      "settle": f"{n_settle} untimed steps (>= {settle_s} s) before the timed region",
                      "scored_pairs_per_s": qps * rows,
                      "scaling_note": "strong scaling on the fixed 16M x 768, batch 8192 workload for N >= 2; its 1-GPU "
                                      "origin is the strong_base leg of the N = 1 line (value at N = 1 is C2, the "
                                      "configuration the metric is quoted on)",
                      "exchange": None if sidx is None else f"{sidx._exchange_mode} exchange after every local search "
                                  f"(peer = gather fused into the merge kernel over NVLink peer memory), merge = {args.merge}"},
           "clocks": clocks,
           "e2e": {"value": qps_e2e, "unit": "queries/s", "h2d_bytes_per_step": batch * D_MODEL * 4,
                   "d2h_bytes_per_step": batch * K * 12, "ms_per_step": ms_e2e_dev / args.steps,
                   "d2h_note": None if sidx is None else ("summed over ranks: each rank reads back its slice" if args.merge == "slice"
                                                         else "per rank: every rank reads back the full result"),
                   "wall_ms_per_step": wall_e2e * 1e3 / args.steps},
           "gpu_launches": int(launches),
           "roofline": roofline,
           "parity_checked": parity_checked,
           "multi_gpu": multi,
           "engine": eng}

    # =============================== N = 1: the other configurations, driver-witnessed ===============================
    if world == 1 and legs:
        N0 = hi - lo
        # ---- C5: small-batch latency on the resident 4M corpus (BASELINE.json configs[4]) ----
        if "c5" in legs:
            c5 = []
            bytes_pass = 2.0 * N0 * D_MODEL
            for B in (1, 8, 64):
                hg = torch.Generator().manual_seed(555 + B)
                hqs = [torch.randn((B, D_MODEL), generator=hg).pin_memory() for _ in range(4)]
                hDb = torch.empty((B, K), dtype=torch.float32).pin_memory()
                hIb = torch.empty((B, K), dtype=torch.int64).pin_memory()
                for i in range(20):
                    local.search(hqs[i % 4].numpy(), K, D=hDb.numpy(), I=hIb.numpy())
                lat = []
                for i in range(200):
                    t0 = time.perf_counter()
                    local.search(hqs[i % 4].numpy(), K, D=hDb.numpy(), I=hIb.numpy())     # complete on return
                    lat.append((time.perf_counter() - t0) * 1e3)
                # the scoring kernel alone: plain launches with event records (graphs off), device time of the main pass
                local.set_option("timing", 1)
                a0 = local.stats()
                for i in range(20):
                    local.search(hqs[i % 4].numpy(), K, D=hDb.numpy(), I=hIb.numpy())
                a1 = local.stats()
                local.set_option("timing", 0)
                st5, _ = stage_delta(a0, a1)
                p50 = pct(lat, 50)
                c5.append({"batch": B, "calls": 200, "call_ms_p50": p50, "call_ms_p99": pct(lat, 99),
                           "qps_at_p50": B / (p50 * 1e-3),
                           "kernel_ms": st5["prefilter"], "kernel_frac_hbm": bytes_pass / (st5["prefilter"] * 1e-3) / 1e9 / pk["hbm"],
                           "call_frac_hbm": bytes_pass / (p50 * 1e-3) / 1e9 / pk["hbm"],
                           "path": {trx.PATH_STREAM: "K3 CUDA-core streaming", trx.PATH_UMMA: "K2 tcgen05"}.get(a1["last_path"], str(a1["last_path"]))})
            out["c5"] = {"workload": f"C5: batch 1/8/64 over {N0} x 768, k=100, host buffers in/out, whole call (graph replay)",
                         "bytes_per_pass": bytes_pass, "hbm_peak_gbs": pk["hbm"], "legs": c5}
            stage(f"c5 done: {[(c['batch'], round(c['call_ms_p50'], 3), round(c['call_ms_p99'], 3)) for c in c5]}")

        # ---- C3: gold-removed mode (BASELINE.json configs[2]) ----
        if "c3" in legs:
            groups = (torch.arange(N0, device=dev, dtype=torch.int64) // 5).to(torch.int32)
            local.set_groups(groups)
            eg = torch.Generator(device=dev); eg.manual_seed(77)
            excls = [groups[torch.randint(0, N0, (batch,), generator=eg, device=dev)].contiguous() for _ in range(len(queries))]

            def step_c3(i):
                local.search(queries[i % len(queries)], K, exclude=excls[i % len(queries)])
            settle(step_c3, args.warmup)
            a0 = local.stats()
            ms3 = timed(step_c3, args.steps)
            a1 = local.stats()
            Dm, Im = local.search(queries[0], K, exclude=excls[0])
            returned_excluded = int((groups[Im.clamp(min=0)] == excls[0][:, None]).sum().item())
            # mask == post-filter of a deeper list (textreact/dataset.py:74-76) on a sample of the batch
            ns = 256
            Dd, Id = local.search(queries[0][:ns], K + 5)
            keep = groups[Id] != excls[0][:ns, None]
            same = 0
            Im_c, Id_c, keep_c = Im[:ns].cpu().numpy(), Id.cpu().numpy(), keep.cpu().numpy()
            for r in range(ns):
                same += int((Id_c[r][keep_c[r]][:K] == Im_c[r]).all())
            st3, _ = stage_delta(a0, a1)
            out["c3"] = {"workload": f"C3: C2 + per-query exclusion (group = row // 5, excl = group of a random row), batch {batch}",
                         "value": batch * args.steps / (ms3 * 1e-3), "unit": "queries/s", "ms_per_step": ms3 / args.steps,
                         "stages_ms": st3, "excluded_rows_returned": returned_excluded,
                         "mask_equals_post_filter": {"queries": ns, "identical": same}, "engine": engine_delta(a0, a1)}
            local.set_groups(None)
            stage(f"c3 done: {ms3 / args.steps:.2f} ms/step, excluded returned {returned_excluded}, post-filter identical {same}/{ns}")

        # ---- strong-scaling base: 16M x 768, batch 8192 on one GPU (BASELINE.json configs[3] at G = 1) ----
        if want_base and args.dist == "G":
            NB, BB = 16_000_000, 8192
            # rows [N0, 16M) continue the generator stream of seed 1234: re-create it and skip what is resident
            for c0, x in corpus_pieces("G", 0, NB, NB, dev, seed):
                if c0 >= N0:
                    local.add(x)
            torch.cuda.synchronize()
            qb = [make_queries("G", BB, dev, 5321 + i) for i in range(4)]
            leg = throughput_leg(local, qb, max(5, args.steps // 2), BB)
            fl = 2.0 * BB * NB * D_MODEL
            tf = fl / (leg["stages_ms"]["prefilter"] * 1e-3) / 1e12
            pc = parity(lambda qbuf, m: local.search(qbuf, K), "G", 0, NB, NB, qb[0])
            out["strong_base"] = {"workload": "16M x 768 fp32 corpus on ONE GPU, batch 8192, k=100 (C4's shape at G = 1)",
                                  "value": leg["qps"], "unit": "queries/s", "ms_per_step": leg["ms_per_step"],
                                  "steps": max(5, args.steps // 2), "stages_ms": leg["stages_ms"],
                                  "roofline_frac": tf / pk["tf_burst"], "achieved_tflops": tf,
                                  "scored_pairs_per_s": leg["qps"] * NB, "engine": leg["engine"], "parity_checked": pc,
                                  "use": "efficiency(N) = value(N) / (N x this value) for the N = 2/4/8 lines"}
            stage(f"strong_base done: {leg['ms_per_step']:.2f} ms/step, parity ok={pc['ok']}")

        # ---- the workload the retriever really has: Dist U (clustered, unit norm) and Dist A (ascending scores) ----
        if "dists" in legs:
            out["dists"] = {}
            for dn in ("U", "A"):
                if dn == args.dist:
                    continue
                local.reset()
                build(dn, 0, rows, rows)
                qd = [make_queries(dn, batch, dev, 4321 + i) for i in range(4)]
                leg = throughput_leg(local, qd, args.steps, batch)
                pc = parity(lambda qbuf, m: local.search(qbuf, K), dn, 0, rows, rows, qd[0])
                out["dists"][dn] = {"dist": DIST_TEXT[dn], "value": leg["qps"], "unit": "queries/s",
                                    "ms_per_step": leg["ms_per_step"], "vs_dist_G": leg["qps"] / qps,
                                    "stages_ms": leg["stages_ms"], "engine": leg["engine"], "parity_checked": pc}
                stage(f"dist {dn} done: {leg['ms_per_step']:.2f} ms/step, engine {leg['engine']}, parity ok={pc['ok']}")

    # ---- CPU baseline (rank 0, N=1 only): the FULL corpus on the host, a bounded sample of one batch's queries ----
    if rank == 0 and world == 1 and "cpu" in legs:
        local.reset()
        torch.cuda.empty_cache()
        host_rows = host_corpus_rows(rows)
        xb = np.empty((host_rows, D_MODEL), dtype=np.float32)
        for c0, x in corpus_pieces("G", 0, host_rows, rows, dev, seed):
            xb[c0:c0 + x.shape[0]] = x.cpu().numpy()
        xq_s = make_queries("G", 1024, dev, 4321).cpu().numpy()
        torch.set_num_threads(os.cpu_count() or 1)
        dt, kind, cores = cpu_flat_search_timed(xb, xq_s, K, 1, 1)
        out["cpu_baseline"] = {"value": xq_s.shape[0] / dt * host_rows / rows, "unit": "queries/s", "cores": cores, "kind": kind,
                               "sample": f"{xq_s.shape[0]} of the {batch} queries of a batch x {host_rows} of {rows} corpus rows "
                                         f"({dt:.2f} s measured" + ("" if host_rows == rows else f"; q/s scaled by {host_rows}/{rows} rows")
                                         + f"); BLAS: {host_blas() if kind == 'port' else 'faiss'}"}
        del xb
    else:
        out["cpu_baseline"] = None

    if rank == 0:
        emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
