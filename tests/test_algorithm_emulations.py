"""CPU emulations (numpy, small sizes) of the two non-obvious reformulations the round-2 kernels rely on.  They restate the
device algorithms step by step and check them against the oracle, so the mathematics is pinned independently of a GPU:

  * ENTROPY in rank space (csrc/k_entropy.cu:k_entropy_rank): sort; for a tolerance the matches of a sample are a
    contiguous rank interval [lo, hi]; lo from ONE search against a float32 threshold corrected by at most one ulp with
    numpy's own float64 predicate; hi + 1 from the prefix sum of the histogram of lo; bit row = T[hi+1] & ~T[lo].
  * lag products on the FP64 tensor cores (csrc/k_basic.cu:lag_products_dmma): with A[i][u] = x[8(b+u)+i] and
    B[u][j] = x[8(b+t+u)+j] the 8 x 8 accumulator of tile t holds lag 8t + j - i on its diagonals.
"""
import numpy as np

from oracle import calculators as C


def _key(f):
    u = np.float32(f).view(np.uint32)
    if u == 0x80000000:
        u = np.uint32(0)
    return np.uint32(~u) if (u & 0x80000000) else np.uint32(u | 0x80000000)


def _unkey(k):
    k = np.uint32(k)
    u = np.uint32(k & 0x7FFFFFFF) if (k & 0x80000000) else np.uint32(~k)
    return u.view(np.float32)


def _round_up_f32(z):
    f = np.float32(z)
    if np.float64(f) < z:
        f = np.nextafter(f, np.float32(np.inf))
    return f


def rank_space_counts(x32, tau):
    """template-match counts c2(i), c3(i) exactly as k_entropy_rank forms them"""
    n = len(x32)
    order = np.argsort(x32, kind="stable")
    s = x32[order]
    rk = np.empty(n, dtype=np.int64)
    rk[order] = np.arange(n)
    sk = np.array([_key(v) for v in s], dtype=np.uint64)
    lo = np.zeros(n, dtype=np.int64)
    for r in range(n):
        sr = np.float64(_unkey(np.uint32(sk[r])))
        Lf = _round_up_f32(sr - tau)
        if not ((sr - np.float64(Lf)) <= tau):
            Lf = _unkey(_key(Lf) + np.uint32(1))
        else:
            Lp = _unkey(_key(Lf) - np.uint32(1))
            if (sr - np.float64(Lp)) <= tau:
                Lf = Lp
        lo[r] = int(np.sum(sk < np.uint64(_key(Lf))))
    hip1 = np.cumsum(np.bincount(lo, minlength=n + 1))[:n]           # hi + 1 = #{a : lo(a) <= r}
    T = np.array([rk < k for k in range(n + 1)])                       # T[k][j] = [rank(j) < k], time order
    R = np.array([T[hip1[rk[i]]] & ~T[lo[rk[i]]] for i in range(n)])
    n2, n3 = n - 1, n - 2
    c2 = np.array([np.sum(R[i, :n2] & R[i + 1, 1:n2 + 1]) for i in range(n2)])
    c3 = np.array([np.sum(R[i, :n3] & R[i + 1, 1:n3 + 1] & R[i + 2, 2:n3 + 2]) for i in range(n3)])
    return c2, c3


def test_entropy_in_rank_space_equals_the_oracle():
    rng = np.random.default_rng(1)
    for n, kind in ((6, "normal"), (37, "rounded"), (64, "walk"), (130, "normal")):
        x = rng.standard_normal(n)
        if kind == "rounded":
            x = np.round(x * 3) / 3
        if kind == "walk":
            x = x.cumsum()
        x32 = x.astype(np.float32)
        x32[0] = np.float32(-0.0) if n == 37 else x32[0]              # -0.0 and +0.0 are one value
        x64 = x32.astype(np.float64)
        sd = np.std(x64)
        c2, c3 = rank_space_counts(x32, 0.2 * sd)
        with np.errstate(all="ignore"):
            se = -np.log(np.sum(c3 - 1) / np.sum(c2 - 1))
        want = C.sample_entropy(x64)
        assert (np.isnan(se) and np.isnan(want)) or se == want or abs(se - want) <= 1e-12 * abs(want)
        for r in (0.1, 0.5, 0.9):
            c2, c3 = rank_space_counts(x32, r * sd)
            ae = abs(np.sum(np.log(c2 / (n - 1))) / (n - 1) - np.sum(np.log(c3 / (n - 2))) / (n - 2))
            assert abs(ae - C.approximate_entropy(x64, 2, r)) <= 1e-12


def test_lag_products_from_8x8_tiles():
    rng = np.random.default_rng(2)
    for n, K in ((256, 40), (100, 40), (37, 10), (64, 41)):
        x = rng.standard_normal(n)
        x -= x.mean()
        ntiles = K // 8 + 1 + (1 if K % 8 else 0)
        pad = np.concatenate([x, np.zeros(32 * ((n + 31) // 32) + 8 * ntiles + 64 - n)])
        G = (n + 31) // 32
        lag = np.zeros(K + 1)
        for t in range(ntiles):
            D = np.zeros((8, 8))
            for g in range(G):                                     # one m8n8k4 MMA per group of four 8-sample blocks
                A = np.array([[pad[8 * (4 * g + u) + i] for u in range(4)] for i in range(8)])
                B = np.array([[pad[8 * (4 * g + t + u) + j] for j in range(8)] for u in range(4)])
                D += A @ B
            for d in range(-7, 8):                                 # diagonal j - i = d of tile t is lag 8 t + d
                k = 8 * t + d
                if 0 <= k <= K:
                    lag[k] += sum(D[i, i + d] for i in range(8) if 0 <= i + d < 8)
        want = np.array([np.dot(x[:n - k], x[k:]) if k < n else 0.0 for k in range(K + 1)])
        np.testing.assert_allclose(lag, want, rtol=1e-11, atol=1e-11)
