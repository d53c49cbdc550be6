"""Generates the committed fixtures under tests/golden/.

  python tests/golden/make_golden.py

1. rosenbrock_mp_trace.json — an INDEPENDENT restatement of the reference algorithm
   (QuasiNewtonMethods.jl:34-69, 157-232, 237-292; Rosenbrock from test/runtests.jl:4-33)
   in 60-digit mpmath arithmetic: the first 5 iterates from seeded starts.  Rounding
   order is immaterial at 60 digits, so this pins what any faithful fp64
   implementation (the Julia reference included) must reproduce to ~1e-13 for the
   first iterations, independently of the C oracle's canonical summation order.
2. oracle_vectors.npz — outputs of the C oracle (oracle/qn_oracle.c) for seeded
   starts: bitwise regression vectors that travel to the GPU box.

The Julia reference itself cannot be run (no Julia runtime in this image), so no
fixture here comes from the reference's own code.
"""
import json
import os
import sys

import mpmath as mp
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

mp.mp.dps = 60
SEED = 20261018


def rosen(th):
    n = len(th); N = n >> 1
    s = mp.mpf(0)
    for i in range(N):
        s -= 100 * (th[i + N] - th[i] ** 2) ** 2 + (1 - th[i]) ** 2
    if n & 1:
        s -= (1 - th[-1]) ** 2
    return s


def rosen_grad(th):
    n = len(th); N = n >> 1
    g = [mp.mpf(0)] * n
    for i in range(N):
        g[i] = 400 * (th[i + N] - th[i] ** 2) * th[i] + 2 * (1 - th[i])
        g[i + N] = 200 * (th[i] ** 2 - th[i + N])
    if n & 1:
        g[-1] = 2 * (1 - th[-1])
    return rosen(th), g


def nanmin(a, b):
    return a if a < b else b


def nanmax(a, b):
    return b if a < b else a


def linesearch(x, s, l0, m, order, c1=mp.mpf("1e-4"), rho_hi=mp.mpf("0.5"), rho_lo=mp.mpf("0.1")):
    step = lambda a: rosen([xi + a * si for xi, si in zip(x, s)])
    a1 = a2 = mp.mpf(1)
    lx0 = lx1 = l0
    lx1 = step(a1)
    nfe = 1
    it = 0
    while not (lx1 >= l0 + a2 * c1 * m):
        it += 1
        if it > 1000:
            return mp.mpf(0), nfe
        if order == 2 or it == 1:
            at = -(m * a2 * a2) / (2 * (lx1 - l0 - m * a2))
        else:
            denom = 1 / (a1 * a1 * a2 * a2 * (a2 - a1))
            a = (a1 * a1 * (lx1 - l0 - m * a2) - a2 * a2 * (lx0 - l0 - m * a1)) * denom
            b = (-a1 * a1 * a1 * (lx1 - l0 - m * a2) + a2 * a2 * a2 * (lx0 - l0 - m * a1)) * denom
            if abs(a) <= mp.mpf(2) ** -52 + mp.mpf(2) ** -26 * abs(a):
                at = m / (2 * b)
            else:
                d = nanmax(b * b - 3 * a * m, mp.mpf(0))
                at = (mp.sqrt(d) + b) / (-3 * a)
        a1 = a2
        at = nanmin(at, a2 * rho_hi)
        a2 = nanmax(at, a2 * rho_lo)
        lx0, lx1 = lx1, step(a2)
        nfe += 1
    return a2, nfe


def bfgs_trace(x0, order, niter):
    n = len(x0)
    x = [mp.mpf(float(v)) for v in x0]
    H = None
    g_old = s = None
    out, alphas = [], []
    for it in range(1, niter + 1):
        l0, g = rosen_grad(x)
        if it == 1:
            m = mp.mpf(-1)
        else:
            y = [a - b for a, b in zip(g_old, g)]
            sy = sum(a * b for a, b in zip(s, y))
            Hy = [sum(H[c][r] * y[r] for r in range(n)) for c in range(n)]
            u = [t / sy for t in Hy]
            yHy = sum(a * b for a, b in zip(Hy, y))
            c1 = (1 + yHy / sy) / sy
            for c in range(n):
                for r in range(n):
                    H[c][r] = H[c][r] + c1 * s[r] * s[c] - u[r] * s[c] - u[c] * s[r]
            d = [sum(H[c][r] * g[r] for r in range(n)) for c in range(n)]
            m = sum(a * b for a, b in zip(d, g))
            s = d
        if m <= 0:
            H = [[mp.mpf(1 if r == c else 0) for r in range(n)] for c in range(n)]
            s = list(g)
            m = sum(a * a for a in g)
        alpha, _ = linesearch(x, s, l0, m, order)
        s = [alpha * v for v in s]
        x = [a + b for a, b in zip(x, s)]
        g_old = g
        out.append([float(v) for v in x])
        alphas.append(float(alpha))
    return out, alphas


def main():
    rng = np.random.Generator(np.random.Philox(key=SEED))
    cases = []
    for n, order in ((60, 2), (60, 3), (7, 2), (24, 3), (2, 2)):
        for rep in range(2):
            x0 = rng.standard_normal(n)
            tr, al = bfgs_trace(x0, order, 5)
            cases.append(dict(n=n, order=order, x0=[float(v).hex() for v in x0], trace=tr, alpha=al))
            print("mp trace n=%d order=%d alphas=%s" % (n, order, al))
    json.dump(dict(seed=SEED, dps=mp.mp.dps, cases=cases), open(os.path.join(HERE, "rosenbrock_mp_trace.json"), "w"))

    from oracle import binding as ob
    vec = {}
    rng = np.random.Generator(np.random.Philox(key=SEED + 1))
    for n in (2, 3, 7, 16, 24, 33, 60, 100):
        for order in (2, 3):
            Theta = rng.standard_normal((8, n)).T.copy(order="F")
            r = ob.optimize_batch(ob.rosenbrock(), Theta, ls=ob.backtracking(order=order), ntrace=5)
            k = "n%d_o%d_" % (n, order)
            vec[k + "theta0"] = Theta
            for f in ("x", "g", "ell", "iters", "fevals", "status", "trace"):
                vec[k + f] = r[f]
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **vec)
    print("wrote fixtures")


if __name__ == "__main__":
    main()
