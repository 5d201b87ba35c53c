"""Numerical model (numpy, CPU) of the online softmax of the persistent hidden-layer kernel (csrc/attn_hidden.cu, NOMAX
form) against TransformerConv's segment softmax (oracle/transformer_conv.py: exp(a - max) / (sum + 1e-16)):

  * 64-source blocks; the reference point m of a row is the masked maximum of the FIRST block in which the row has an edge
    and only moves when a block's row sum leaves [0, 2^60) -- then to the maximum of that block, rescaling l and O;
  * P is split into bf16 hi + lo (round to nearest even, lo = bf16(p - hi)), V likewise; O accumulates
    P_hi V_hi + P_hi V_lo + P_lo V_hi in fp32;  l accumulates the fp32 p;  out = O / (l + 1e-16).

The claims checked: any reference point cancels; l >= 1 as soon as the row has an edge, so the reference's 1e-16 stays
negligible; the split-bf16 products keep fp32-level accuracy at any magnitude of p; rows without any edge give exactly 0.
The CUDA kernel itself is held to the oracle in tests/test_gpu_benchmark_config.py.
"""
import numpy as np
import pytest


def _bf16(x):
    """fp32 -> bf16 (round to nearest even) -> fp32."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)


def _kernel_model(S, mask, V, C):
    n, ns = S.shape
    c_log2 = np.float32(1.4426950408889634 / np.sqrt(C))
    v_hi = _bf16(V)
    v_lo = _bf16(V - v_hi)
    m = np.full(n, -np.inf, dtype=np.float32)
    l = np.zeros(n, dtype=np.float32)
    O = np.zeros((n, V.shape[1]), dtype=np.float32)
    n_rescales = 0
    for b0 in range(0, ns, 64):
        s, mk = S[:, b0:b0 + 64].astype(np.float32), mask[:, b0:b0 + 64]
        sm = np.where(mk, s, -np.inf).astype(np.float32)
        first = b0 == 0
        if first:   # pre-pass of the first block
            m = np.maximum(m, sm.max(1))
        while True:
            m_sub = np.where(np.isinf(m), np.float32(0), m * c_log2).astype(np.float32)
            with np.errstate(over="ignore", invalid="ignore"):
                p = np.exp2(sm * c_log2 - m_sub[:, None]).astype(np.float32)
            p = np.where(mk, p, np.float32(0))
            lsum = p.sum(1, dtype=np.float32)
            exceeded = ~(lsum < np.float32(2.0 ** 60)) | (np.isinf(m) & mk.any(1))
            if not exceeded.any():
                break
            n_rescales += 1
            bmax = np.maximum(sm.max(1), m)
            m_new = np.where(exceeded, bmax, m).astype(np.float32)
            with np.errstate(invalid="ignore"):
                alpha = np.where(np.isinf(m), np.float32(0), np.exp2((m - m_new) * c_log2)).astype(np.float32)
            O *= alpha[:, None]
            l *= alpha
            m = m_new
        l = (l + lsum).astype(np.float32)
        p_hi = _bf16(p)
        p_lo = _bf16(p - p_hi)
        vh, vl = v_hi[b0:b0 + 64], v_lo[b0:b0 + 64]
        O = (O + p_hi @ vh + p_hi @ vl + p_lo @ vh).astype(np.float32)
    return O / (l + np.float32(1e-16))[:, None], l, n_rescales


def _reference(S, mask, V, C):
    a = np.where(mask, S.astype(np.float64) / np.sqrt(C), -np.inf)
    mx = np.where(mask.any(1), a.max(1), 0.0)
    e = np.where(mask, np.exp(a - mx[:, None]), 0.0)
    return (e / (e.sum(1) + 1e-16)[:, None]) @ V.astype(np.float64)


@pytest.mark.parametrize("case", ["typical", "late_large_scores", "first_blocks_empty", "huge_jump", "isolated_rows"])
def test_max_free_online_softmax_model(case):
    rng = np.random.default_rng(3)
    n, ns, C = 128, 960, 32
    S = (rng.standard_normal((n, ns)) * 12.0).astype(np.float32)     # raw q.k scores, scale 1/sqrt(32) applied in the loop
    mask = rng.random((n, ns)) < 0.6
    V = rng.standard_normal((ns, C)).astype(np.float32)
    if case == "late_large_scores":       # every later block beats the first one's maximum by far: p up to 2^40, no rescale needed
        S += (np.arange(ns)[None, :] // 64 * 12.0).astype(np.float32)
    elif case == "first_blocks_empty":    # band structure: rows meet their first edge in different blocks
        for i in range(n):
            mask[i, : 64 * (i % 9)] = False
    elif case == "huge_jump":             # a score 2^100 above the reference point: the row sum overflows the 2^60 window
        S[:, 700] = 500.0
        mask[:, 700] = True
    elif case == "isolated_rows":
        mask[::7] = False
    got, l, n_rescales = _kernel_model(S, mask, V, C)
    want = _reference(S, mask, V, C)
    has_edge = mask.any(1)
    assert (l[has_edge] >= 1.0 - 1e-6).all()          # the first edge block contributes p = 1 at its maximum
    assert (got[~has_edge] == 0).all()
    # split-bf16 products: p_lo and v_lo carry 8 more bits each, so a product is good to ~2^-17 (the dropped lo x lo term is
    # 2^-18): ~1e-5 on a row dominated by one edge, far inside the 1e-4 bar of the whole step
    err = np.abs(got - want).max() / np.abs(want).max()
    assert err < 3e-5, (case, err)
    if case == "huge_jump":
        assert n_rescales >= 1
    if case == "late_large_scores":
        assert n_rescales == 0 and l.max() > 2.0 ** 30
