"""Comparators shared by the parity tests (tie-aware, tolerance stated at the call site)."""
from __future__ import annotations

import torch

# north_star: top-k indices must agree exactly except for near-ties within this score gap
NEAR_TIE_GAP = 1e-3


def make_inputs(Q, V, D, seed, dist="normal", dtype=torch.bfloat16, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    if dist == "normal":
        q = torch.randn(Q, D, generator=g)
        t = torch.randn(V, D, generator=g)
    elif dist == "aniso":   # LM-table-like: small noise around a shared mean direction
        mean = torch.randn(1, D, generator=g)
        q = torch.randn(Q, D, generator=g) * 0.02 + mean
        t = torch.randn(V, D, generator=g) * 0.02 + mean
    else:
        raise ValueError(dist)
    return q.to(dtype).to(device), t.to(dtype).to(device)


def check_topk(val, idx, ref_scores, k, *, rtol, atol=1e-6, gap=NEAR_TIE_GAP, index_base=0,
               exact_ties_lowest=False):
    """val/idx: [Q,k] from the implementation; ref_scores: [Q,V] ground truth (fp64)."""
    val = val.detach().cpu().double()
    idx = idx.detach().cpu().long() - index_base
    ref = ref_scores.detach().cpu().double()
    Q, V = ref.shape
    assert val.shape == (Q, k) and idx.shape == (Q, k)
    assert (idx >= 0).all() and (idx < V).all(), "index out of range"
    # descending values, unique indices
    assert (val[:, 1:] <= val[:, :-1] + 1e-12).all(), "values not sorted descending"
    assert all(len(set(r.tolist())) == k for r in idx), "duplicate index in a row"
    got = torch.gather(ref, 1, idx)
    err = (val - got).abs()
    tol = atol + rtol * got.abs()
    assert (err <= tol).all(), f"value mismatch: max err {err.max():.3e} (tol {tol.max():.3e})"
    ref_sorted, ref_order = torch.sort(ref, dim=1, descending=True, stable=True)
    kth = ref_sorted[:, k - 1:k]
    assert (got >= kth - gap).all(), "returned an item that is not within the near-tie gap of the top-k"
    # every reference top-k item that is clearly above the boundary must be returned
    nxt = ref_sorted[:, k:k + 1] if V > k else torch.full_like(kth, -float("inf"))
    for i in range(Q):
        clear = ref_order[i, :k][ref_sorted[i, :k] > nxt[i] + gap]
        missing = set(clear.tolist()) - set(idx[i].tolist())
        assert not missing, f"row {i}: missing clear top-k items {sorted(missing)[:5]}"
    if exact_ties_lowest:
        assert torch.equal(idx, ref_order[:, :k]), "exact order (lowest index wins) differs"


def check_stats(stats, ref, *, rtol, atol=1e-5):
    """stats [Q,4] vs an oracle ScanResult (m, s, sum_z, z_label) through lse / sum_z / z_label."""
    st = stats.detach().cpu().double()
    lse = st[:, 0] + torch.log(st[:, 1])
    for name, got, want in (("lse", lse, ref.lse.double()), ("sum_z", st[:, 2], ref.sum_z.double()),
                            ("z_label", st[:, 3], ref.z_label.double()),
                            ("m", st[:, 0], ref.m.double())):
        err = (got - want).abs()
        scale = want.abs().max().clamp_min(1.0) if name == "sum_z" else want.abs()
        tol = atol + rtol * scale
        assert (err <= tol).all(), f"{name}: max err {err.max():.3e}"
