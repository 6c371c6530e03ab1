"""How much of the EMD Bid scan could a block-level bound skip?  CPU simulation (numpy) of the auction on the bench's C5
inputs (uniform pairs, n = 8192, eps 0.005, 50 iterations): per iteration, the number of bidders and the fraction of
G-target blocks (targets sorted along a Morton curve) whose upper bound 3 - boxdist - min price reaches the bidder's
second-best value.  Statistics only: ties are resolved arbitrarily here."""
import sys
import numpy as np
import os
PMINB = int(os.environ.get("PMINB", "1"))


def morton(p, bits):
    lo, hi = p.min(0), p.max(0)
    q = np.minimum(((p - lo) / np.maximum(hi - lo, 1e-30) * (1 << bits)).astype(np.int64), (1 << bits) - 1)
    code = np.zeros(len(p), np.int64)
    for b in range(bits):
        for a in range(3):
            code |= ((q[:, a] >> b) & 1) << (3 * b + a)
    return code


def main(n=8192, G=64, iters=50, eps=0.005, seed=0, kind="uniform"):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        x, y = rng.random((n, 3), np.float32), rng.random((n, 3), np.float32)
    else:
        sys.path.insert(0, ".")
        from genpc_b200 import synthetic
        x, y = synthetic.superquadric_pair(seed, n, n)
        x, y = x.astype(np.float32), y.astype(np.float32)
    order = np.argsort(morton(y, 10), kind="stable")
    ys = y[order]
    nb = n // G
    lo = ys.reshape(nb, G, 3).min(1)
    hi = ys.reshape(nb, G, 3).max(1)
    price = np.zeros(n, np.float32)   # sorted space
    asg = np.full(n, -1)
    inv = np.full(n, -1)
    tot_pairs = tot_visit = 0
    for it in range(iters):
        un = np.nonzero(asg < 0)[0]
        U = len(un)
        if U == 0:
            break
        frac = []
        best_k = np.empty(U, np.int64)
        inc = np.empty(U, np.float32)
        pminb = price.reshape(nb, G).min(1) if PMINB else np.zeros(nb, np.float32)
        for c0 in range(0, U, 512):
            q = x[un[c0:c0 + 512]]
            d = np.sqrt(((q[:, None, :] - ys[None]) ** 2).sum(-1))
            v = 3.0 - d - price[None]
            part = np.partition(v, n - 2, axis=1)
            best, better = part[:, n - 1], part[:, n - 2]
            best_k[c0:c0 + 512] = v.argmax(1)
            inc[c0:c0 + 512] = best - better + eps
            bd = np.maximum(np.maximum(lo[None] - q[:, None], q[:, None] - hi[None]), 0)
            ub = 3.0 - np.sqrt((bd ** 2).sum(-1)) - pminb[None]
            frac.append((ub >= better[:, None] - 1e-4).mean(1))
        frac = np.concatenate(frac)
        tot_pairs += U * n
        tot_visit += frac.sum() * n
        print(f"it {it:2d} U {U:5d} blocks passing: mean {frac.mean():.3f} median {np.median(frac):.3f} p90 {np.quantile(frac, .9):.3f} max {frac.max():.3f}")
        # GetMax / Assign
        win = {}
        for u in np.argsort(-inc, kind="stable"):
            k = best_k[u]
            if k not in win:
                win[k] = u
        for k, u in win.items():
            if inv[k] >= 0:
                asg[inv[k]] = -1
            inv[k] = un[u]
            asg[un[u]] = k
            price[k] += inc[u]
    print(f"total (bidder, target) evaluations {tot_pairs:.3e}; inside passing blocks {tot_visit:.3e} = {tot_visit / tot_pairs:.3f}")


if __name__ == "__main__":
    main(G=int(sys.argv[1]) if len(sys.argv) > 1 else 64, kind=sys.argv[2] if len(sys.argv) > 2 else "uniform")
