"""Diagnostic (GPU box): tcgen05 first-layer kernels vs the fp32 SIMT kernels, stage by stage."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from locator_b200 import model  # noqa: E402


def mk(impl, K, seed=3):
    os.environ["LOC_L1_IMPL"] = impl
    m = model.LocatorModel(K, seed=seed)
    os.environ.pop("LOC_L1_IMPL")
    return m


def stats(name, a, b):
    d = np.abs(a - b)
    den = np.abs(b).max() + 1e-30
    print(f"  {name:10s} max|d|={d.max():.3e}  max|ref|={den:.3e}  rel={d.max() / den:.3e}  "
          f"frac(|d|>1e-3*max)={(d > 1e-3 * den).mean():.4f}", flush=True)
    return d.max() / den


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 4096 + 40
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    rng = np.random.default_rng(0)
    n = 80
    p = rng.uniform(0.02, 0.98, K)
    x = rng.binomial(2, p, size=(n, K)).astype(np.uint8)
    y = rng.normal(size=(n, 2)).astype(np.float32)
    ms = {impl: mk(impl, K) for impl in ("simt", "tcgen05")}
    print("impls:", {k: v.impl for k, v in ms.items()}, "K", K, "nb", nb, flush=True)
    rows = rng.permutation(n)[:nb]
    ws = ms["simt"].get_weights()
    ws[0] = rng.uniform(0.5, 1.5, K).astype(np.float32)
    ws[1] = rng.normal(0, 0.2, K).astype(np.float32)
    for m in ms.values():
        m.set_weights(ws)
        m.bind_train(x, y)
        m.set_schedule(patience=100)
    # ---- forward ----
    Z = {}
    for impl, m in ms.items():
        m.debug_stage(0, rows)
        Z[impl] = m.debug_read(0).sum(0)
    print("forward (Z1 = sum of partial tiles):")
    r = stats("Z1", Z["tcgen05"], Z["simt"])
    if r > 5e-3:
        print("  simt[0,:8]", Z["simt"][0, :8])
        print("  tc  [0,:8]", Z["tcgen05"][0, :8])
        print("  simt[:8,0]", Z["simt"][:8, 0])
        print("  tc  [:8,0]", Z["tcgen05"][:8, 0])
    w_s, w_t = ms["simt"].get_weights(), ms["tcgen05"].get_weights()
    stats("mov.mean", w_t[2], w_s[2])
    stats("mov.var", w_t[3], w_s[3])
    # ---- hidden + backward (feed the SAME dZ1 to both: run hidden on each model's own partials) ----
    for impl, m in ms.items():
        m.debug_stage(1, rows)
    d_s, d_t = ms["simt"].debug_read(1)[0], ms["tcgen05"].debug_read(1)[0]
    print("hidden (dZ1 from each model's own Z1):")
    stats("dZ1", d_t, d_s)
    for impl, m in ms.items():
        m.debug_stage(2, rows)
    w_s, w_t = ms["simt"].get_weights(), ms["tcgen05"].get_weights()
    print("backward + Adam (updates relative to lr = 1e-3):")
    for idx, nm in ((4, "W1"), (0, "gamma"), (1, "beta")):
        stats(nm + " upd", w_t[idx] - ws[idx], w_s[idx] - ws[idx])
    for idx, nm in ((4, "W1"), (0, "gamma"), (1, "beta")):
        (m_s, v_s), (m_t, v_t) = ms["simt"].get_adam(idx), ms["tcgen05"].get_adam(idx)
        stats(nm + ".m", m_t, m_s)
        stats(nm + ".v", v_t, v_s)


if __name__ == "__main__":
    main()
