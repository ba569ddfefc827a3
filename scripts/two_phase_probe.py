#!/usr/bin/env python
"""Per-kernel cost of the two-phase local search at the N = 8 shard shape (2M x 768 rows, batch 8192, k = 100) on ONE GPU:
the floor is synthesised from the shard's own payload (its 13th best prefilter score ~ the global 100th of 8 such
shards), so K4 sees the candidate counts it sees at N = 8.  Run under `ncu --metrics gpu__time_duration.sum`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import textreact_b200 as trx  # noqa: E402

rows, B, K, nb = int(os.environ.get("ROWS", 2_000_000)), 8192, 100, 32
dev = torch.device("cuda", 0)
idx = trx.IndexFlatIP(768, device=0)
g = torch.Generator(device=dev); g.manual_seed(1)
for c0 in range(0, rows, 500_000):
    idx.add(torch.randn((min(500_000, rows - c0), 768), generator=g, device=dev))
q = torch.randn((B, 768), generator=g, device=dev)
for it in range(3):
    s0 = idx.stats()
    payload = idx.search_begin(q, K, nb)
    floor = (payload[:, 12] - 2.0 * payload[:, nb]).contiguous()
    D, I = idx.search_finish(floor)
    s1 = idx.stats()
    print("two-phase: rescored/query", (s1["rescored"] - s0["rescored"]) / B, "valid/query", float((I >= 0).sum()) / B)
    s0 = idx.stats()
    D, I = idx.search(q, K)
    s1 = idx.stats()
    print("plain: rescored/query", (s1["rescored"] - s0["rescored"]) / B)
torch.cuda.synchronize()
