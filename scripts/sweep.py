#!/usr/bin/env python
"""scripts/sweep.py -- the non-headline measurements of SURVEY.md section 8d on one B200:

  C5  small-batch latency sweep: batch 1..256 over a 4M x 768 corpus, p50/p99 of the whole
      `index.search` call (host buffers, so H2D/D2H are inside) and of the device-side batch, plus
      the scoring kernel alone against the HBM roofline (2*N*d bytes of bf16 corpus per batch),
      for the CUDA-core streaming prefilter (K3) and the tcgen05 prefilter (K2): the measured
      crossover between the two is what `stream_max_batch` defaults to.
  C3  gold-removed mode: C2 shape with a per-query exclusion group (group = row // 5).
  C2 full job (--full-job 700000): the whole USPTO-scale query set through one `index.search` call.

  python scripts/sweep.py [--rows N] [--out gpurun_out/sweep.json] [--reps 200]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import textreact_b200 as trx  # noqa: E402

D_MODEL, K = 768, 100


def pct(v, p):
    v = sorted(v)
    return v[min(len(v) - 1, int(round(p / 100.0 * (len(v) - 1))))]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=4_000_000)
    ap.add_argument("--reps", type=int, default=200)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.json"))
    ap.add_argument("--skip-c3", action="store_true")
    ap.add_argument("--full-job", type=int, default=0, help="also run this many queries (C2: 700000) as one call")
    ap.add_argument("--batches", default="1,2,4,8,16,32,64,128,192,256,384,512,1024")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}

    idx = trx.IndexFlatIP(D_MODEL, device=0)
    idx.reserve(args.rows)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234)
    for c0 in range(0, args.rows, 500_000):
        x = torch.randn((min(500_000, args.rows - c0), D_MODEL), generator=gen, device=dev, dtype=torch.float32)
        idx.add(x)
        del x
    qgen = torch.Generator(device=dev)
    qgen.manual_seed(4321)
    out = {"rows": args.rows, "d": D_MODEL, "k": K, "hbm_peak_gbs": peaks["hbm_gbs"], "c5": [], "c3": None}
    corpus_bytes = 2.0 * args.rows * D_MODEL

    # ---- C5 ---------------------------------------------------------------------------------------
    for B in [int(b) for b in args.batches.split(",")]:
        qs = [torch.randn((B, D_MODEL), generator=qgen, device=dev, dtype=torch.float32).cpu().numpy() for _ in range(8)]
        D = np.empty((B, K), np.float32)
        I = np.empty((B, K), np.int64)
        for name, path in (("stream", trx.PATH_STREAM), ("umma", trx.PATH_UMMA), ("umma_single_cta", trx.PATH_UMMA)):
            if name == "stream" and B > 8:
                continue
            if name == "umma_single_cta" and B <= 128:
                continue                      # batches up to 128 queries are single-CTA tiles anyway
            idx.set_option("path", path)
            idx.set_option("umma_pair", 0 if name == "umma_single_cta" else 1)
            idx.set_option("timing", 1)
            for i in range(5):
                idx.search(qs[i % 8], K, D=D, I=I)
            wall, devtot, kern = [], [], []
            for i in range(args.reps):
                t0 = time.perf_counter()
                idx.search(qs[i % 8], K, D=D, I=I)
                wall.append((time.perf_counter() - t0) * 1e3)
                s = idx.stats()
                devtot.append(s["last_total_ms"]); kern.append(s["last_prefilter_ms"])
            idx.set_option("timing", 0)
            wall_nt = []
            for i in range(args.reps):
                t0 = time.perf_counter()
                idx.search(qs[i % 8], K, D=D, I=I)
                wall_nt.append((time.perf_counter() - t0) * 1e3)
            kmed = pct(kern, 50)
            rec = {"batch": B, "path": name, "reps": args.reps,
                   "call_ms_p50": pct(wall_nt, 50), "call_ms_p99": pct(wall_nt, 99),
                   "device_ms_p50": pct(devtot, 50), "device_ms_p99": pct(devtot, 99),
                   "kernel_ms_p50": kmed, "kernel_ms_p99": pct(kern, 99),
                   "kernel_gbs": corpus_bytes / (kmed * 1e-3) / 1e9,
                   "kernel_frac_hbm": corpus_bytes / (kmed * 1e-3) / 1e9 / peaks["hbm_gbs"],
                   "call_frac_hbm": corpus_bytes / (pct(wall_nt, 50) * 1e-3) / 1e9 / peaks["hbm_gbs"],
                   "qps_p50": B / (pct(wall_nt, 50) * 1e-3),
                   "queries_exact": idx.stats()["queries_exact"]}
            out["c5"].append(rec)
            print(json.dumps(rec), flush=True)
    idx.set_option("path", trx.PATH_AUTO)
    idx.set_option("umma_pair", 1)

    # ---- C3 ---------------------------------------------------------------------------------------
    if not args.skip_c3:
        B = 4096
        groups = (torch.arange(args.rows, device=dev, dtype=torch.int64) // 5).to(torch.int32)
        idx.set_groups(groups)
        qs = [torch.randn((B, D_MODEL), generator=qgen, device=dev, dtype=torch.float32) for _ in range(4)]
        ex = [groups[torch.randint(0, args.rows, (B,), device=dev, generator=qgen)].contiguous() for _ in range(4)]
        res = {}
        for name, use in (("masked", True), ("unmasked", False)):
            for i in range(3):
                idx.search(qs[i % 4], K, exclude=ex[i % 4] if use else None)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            steps = 10
            for i in range(steps):
                idx.search(qs[i % 4], K, exclude=ex[i % 4] if use else None)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            res[name] = {"ms_per_batch": ms, "qps": B / (ms * 1e-3)}
        # correctness of the mask at full size: no returned row belongs to the excluded group
        Dm, Im = idx.search(qs[0], K, exclude=ex[0])
        bad = int((groups[Im.clamp(min=0)] == ex[0][:, None]).sum().item())
        res["excluded_rows_returned"] = bad
        res["stats"] = {k: v for k, v in idx.stats().items() if k.startswith("queries")}
        out["c3"] = res
        print(json.dumps({"c3": res}), flush=True)

    # ---- C2 as one job: 700K queries through ONE index.search call, pageable host arrays in and out ----
    if args.full_job:
        nq = args.full_job
        idx.set_groups(None)
        idx.set_option("max_batch", 4096)
        hq = np.empty((nq, D_MODEL), np.float32)
        for c0 in range(0, nq, 65536):
            c1 = min(nq, c0 + 65536)
            hq[c0:c1] = torch.randn((c1 - c0, D_MODEL), generator=qgen, device=dev, dtype=torch.float32).cpu().numpy()
        idx.search(hq[:8192], K)
        idx.set_option("pipeline", 0)
        t0 = time.perf_counter()
        idx.search(hq[:nq // 4], K)
        dt_serial = (time.perf_counter() - t0) * 4
        idx.set_option("pipeline", 1)
        s0 = idx.stats()
        t0 = time.perf_counter()
        Dj, Ij = idx.search(hq, K)
        dt = time.perf_counter() - t0
        s1 = idx.stats()
        res = {"queries": nq, "batch": 4096, "seconds": dt, "qps": nq / dt, "host_buffers": "pageable numpy",
               "qps_serial_batches": nq / dt_serial,
               "h2d_bytes": hq.nbytes, "d2h_bytes": Dj.nbytes + Ij.nbytes,
               "queries_exact": s1["queries_exact"] - s0["queries_exact"],
               "queries_uncert": s1["queries_uncert"] - s0["queries_uncert"]}
        # self-consistency of the big job: a late slice re-run alone returns the same rows
        D2, I2 = idx.search(hq[nq - 1000:], K)
        res["tail_slice_identical"] = bool((I2 == Ij[nq - 1000:]).all() and (D2 == Dj[nq - 1000:]).all())
        out["c2_full_job"] = res
        print(json.dumps({"c2_full_job": res}), flush=True)
        idx.set_option("max_batch", 8192)

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    idx.close()


if __name__ == "__main__":
    main()
