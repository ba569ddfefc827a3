"""The acceptance rule of BASELINE.json's north_star as an executable check (numpy only: no engine, no CPU restatement):
"returned ids must be identical wherever the fp32 score gap at rank k exceeds a stated tolerance of 1e-5 relative, and
scores must agree within that same tolerance".  The caller supplies a float64 ranking of the same corpus (bench.py
computes one on the device after every timed region; the tests bring their own CPU arbiter)."""
import numpy as np


def north_star_rule(D, I, D64, I64, N64, qnorm, k, rtol=1e-5):
    """BASELINE.json north_star, executable (numpy; inner product): returned ids identical to the float64 ranking
    wherever the fp64 gap at a rank exceeds rtol relative; scores within rtol * |q||x| of the fp64 score of the id
    returned.  D64/I64/N64: fp64 top-(k+extra) scores, ids and row norms.  -> dict of counts; ok False on violation."""
    forced = tied = 0
    bad = []
    for i in range(D.shape[0]):
        ref = {int(r): (float(s), float(n)) for r, s, n in zip(I64[i], D64[i], N64[i])}
        for j in range(k):
            rid = int(I[i, j])
            if rid not in ref:
                bad.append(f"q{i} rank{j}: id {rid} is not in the fp64 top-{D64.shape[1]}")
                continue
            s64, xn = ref[rid]
            if abs(float(D[i, j]) - s64) > rtol * max(qnorm[i] * xn, 1e-30):
                bad.append(f"q{i} rank{j}: score {float(D[i, j])!r} vs fp64 {s64!r}")
        if np.any(D[i, :-1] < D[i, 1:]):
            bad.append(f"q{i}: D not descending")
        for j in range(k):
            a, b = float(D64[i, j]), float(D64[i, j + 1])
            if abs(a - b) / max(abs(a), abs(b), 1e-30) > rtol:
                forced += 1
                if set(I[i, :j + 1].tolist()) != set(I64[i, :j + 1].tolist()):
                    bad.append(f"q{i}: top-{j + 1} id set differs from fp64 across a gap > {rtol}")
            else:
                tied += 1
    return {"ok": not bad, "forced_ranks": forced, "tied_ranks": tied, "violations": bad[:5]}
