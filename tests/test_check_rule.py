"""CPU: the in-run parity checker of bench.py (textreact_b200/check.py) accepts what the north_star rule accepts and
flags what it forbids -- a checker that cannot fail would make every `parity_checked.ok` meaningless."""
import numpy as np

from textreact_b200.check import north_star_rule


def _case(nq=6, k=10, extra=4, seed=0):
    rng = np.random.default_rng(seed)
    D64 = -np.sort(-rng.standard_normal((nq, k + extra)) * 10, axis=1)           # well separated, descending
    I64 = np.stack([rng.permutation(1000)[:k + extra] for _ in range(nq)]).astype(np.int64)
    N64 = np.full((nq, k + extra), 20.0)
    qn = np.full(nq, 20.0)
    return D64, I64, N64, qn, k


def test_exact_answer_passes():
    D64, I64, N64, qn, k = _case()
    r = north_star_rule(D64[:, :k].astype(np.float32), I64[:, :k], D64, I64, N64, qn, k)
    assert r["ok"] and r["forced_ranks"] == D64.shape[0] * k and r["tied_ranks"] == 0


def test_wrong_id_across_a_gap_is_flagged():
    D64, I64, N64, qn, k = _case()
    I = I64[:, :k].copy()
    I[2, k - 1] = I64[2, k]                                  # rank k+1 returned instead of rank k
    D = D64[:, :k].astype(np.float32)
    D[2, k - 1] = D64[2, k]
    r = north_star_rule(D, I, D64, I64, N64, qn, k)
    assert not r["ok"] and any("id set differs" in v for v in r["violations"])


def test_score_off_by_more_than_the_tolerance_is_flagged():
    D64, I64, N64, qn, k = _case()
    D = D64[:, :k].astype(np.float32)
    D[0, 3] += 1e-5 * 400 * 3                                # tolerance is 1e-5 * |q||x| = 4e-3
    r = north_star_rule(D, I64[:, :k], D64, I64, N64, qn, k)
    assert not r["ok"] and any("score" in v for v in r["violations"])


def test_swap_inside_a_tie_group_is_accepted_but_foreign_ids_are_not():
    D64, I64, N64, qn, k = _case()
    D64[1, 4] = D64[1, 3] * (1 - 1e-7)                        # ranks 4 and 5 tie within 1e-5 relative
    I = I64[:, :k].copy()
    I[1, 3], I[1, 4] = I64[1, 4], I64[1, 3]
    D = D64[:, :k].astype(np.float32)
    r = north_star_rule(D, I, D64, I64, N64, qn, k)
    assert r["ok"] and r["tied_ranks"] == 1
    I[1, 0] = 999_999                                         # an id that is nowhere near the top
    assert not north_star_rule(D, I, D64, I64, N64, qn, k)["ok"]


def test_unsorted_scores_are_flagged():
    D64, I64, N64, qn, k = _case()
    D = D64[:, :k].astype(np.float32)
    I = I64[:, :k].copy()
    D[3, [1, 2]] = D[3, [2, 1]]
    I[3, [1, 2]] = I[3, [2, 1]]
    assert not north_star_rule(D, I, D64, I64, N64, qn, k)["ok"]
