"""Seeded inputs of the headline-shape parity cases (BASELINE.json configs[1] mixture: 200 classes x 10 prototypes,
800-row banks; D = 128, and the D = 256 / K = 20 / K = 40 variants).  Pure numpy (PCG64 streams are stable across
machines), shared by tests/golden/make_golden_headline.py -- which feeds them to the UNMODIFIED reference and stores
its outputs -- and by the GPU parity tests, which regenerate them on the GPU box and feed them to the CUDA path and
to the numpy oracle.  Test infrastructure only."""
import numpy as np


def l2n(x, axis):
    return x / np.maximum(np.sqrt((x * x).sum(axis=axis, keepdims=True)), 1e-12)


def mixture(C, K, D, seed=2, sigma_mode="init", pi_mode="rand"):
    """mu as model.py:148-149 (l2-normalised uniform), sigma 1/sqrt(2 pi) (model.py:151) or per-prototype /
    per-dimension random, pi rows on the simplex."""
    r = np.random.default_rng(seed)
    mu = l2n(r.random((C, K, D), dtype=np.float64), 2).astype(np.float32)
    if sigma_mode == "init":
        sg = np.full((C, K, D), 1.0 / np.sqrt(2 * np.pi), np.float32)
    elif sigma_mode == "iso":                                   # constant over d inside each prototype
        sg = np.repeat((0.3 + 0.3 * r.random((C, K, 1))).astype(np.float32), D, axis=2)
    else:                                                       # general diagonal
        sg = (0.3 + 0.3 * r.random((C, K, D))).astype(np.float32)
    if pi_mode == "rand":
        e = np.exp(r.standard_normal((C, K)))
        pi = (e / e.sum(1, keepdims=True)).astype(np.float32)
    else:
        pi = np.full((C, K), 1.0 / K, np.float32)
    wt = np.zeros((C, C * K), np.float32)
    for c in range(C):
        wt[c, c * K:(c + 1) * K] = pi[c]
    return mu, sg, wt


def head_batch(B, C, K, D, H, W, mu, seed=1, gt_fixed=()):
    """Add-on feature maps [B,D,H,W]: noise plus, on a third of the patches, a pull towards a random prototype
    (so top-k gaps vary); labels hit the classes in gt_fixed first (classes whose prototypes straddle the
    128-row tensor-core tiles)."""
    r = np.random.default_rng(seed)
    x = r.standard_normal((B, D, H, W)).astype(np.float32)
    P = C * K
    pick = r.integers(0, P, size=(B, H, W))
    pull = (r.random((B, H, W)) < 0.33).astype(np.float32) * (2.0 + 4.0 * r.random((B, H, W))).astype(np.float32)
    x = x + np.transpose(mu.reshape(P, D)[pick], (0, 3, 1, 2)) * pull[:, None] * np.sqrt(D).astype(np.float32)
    x = x * (0.5 + r.random((B, 1, H, W))).astype(np.float32)
    gt = r.integers(0, C, size=(B,)).astype(np.int64)
    for i, c in enumerate(gt_fixed):
        if i < B:
            gt[i] = c
    return x, gt


def bank_rows(C, K, D, cap, mu, seed=6):
    """rows = l2_normalize(mu_c,k + 0.3 randn), every class full."""
    r = np.random.default_rng(seed)
    kk = r.integers(0, K, size=(C, cap))
    rows = mu[np.arange(C)[:, None], kk] + 0.3 * r.standard_normal((C, cap, D)).astype(np.float32)
    return l2n(rows.astype(np.float32), 2)


def em_state(C, K, D, seed=7, n_active=(160, 140), n_short=5, step0=1000):
    """Pre-seeded Adam state (as after `step0` steps), the update flags of two successive update_GMM calls, and a few
    flagged classes whose bank is not full (their flag is cleared without an update, model.py:287-289)."""
    r = np.random.default_rng(seed)
    m = (1e-3 * r.standard_normal((C, K, D))).astype(np.float32)
    v = ((1e-3) ** 2 * (0.1 + r.random((C, K, D)))).astype(np.float32)
    flags = []
    for n in n_active:
        f = np.zeros(C, bool)
        f[r.choice(C, size=min(n, C), replace=False)] = True
        flags.append(f)
    short = r.choice(C, size=min(n_short, C), replace=False)
    for f in flags:
        f[short] = True
    return m, v, flags, short, step0
