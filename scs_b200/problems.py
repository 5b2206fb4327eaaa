"""Synthetic conic problems with a known optimum, O(nnz) to generate.

Same construction as the reference's generator (test/problem_utils.h:22-81:
z ~ U[-1,1]^m, y = Pi_{K*}(z), s = y - z, x ~ U[-1,1]^n, b = A x + s,
c = -A' y  =>  (x, y, s) is primal-dual optimal with objective c'x), but the
sparsity pattern is drawn per column in O(col_nnz) instead of the reference's
O(m) Knuth selection loop, which cannot reach n = 1e6.

The named configurations of BASELINE.json live in CONFIGS.
"""
import numpy as np

SQRT2 = np.sqrt(2.0)


def random_sparse_csc(m, n, col_nnz, rng):
    """n columns, each with col_nnz distinct sorted rows uniform in [0, m); values U[-1,1]."""
    col_nnz = int(min(col_nnz, m))
    rows = rng.integers(0, m, size=(n, col_nnz), dtype=np.int64)
    rows.sort(axis=1)
    # resolve duplicates inside a column (rare for col_nnz << m)
    for _ in range(64):
        dup = np.zeros_like(rows, dtype=bool)
        dup[:, 1:] = rows[:, 1:] == rows[:, :-1]
        nd = int(dup.sum())
        if nd == 0:
            break
        rows[dup] = rng.integers(0, m, size=nd, dtype=np.int64)
        rows.sort(axis=1)
    else:
        raise RuntimeError("could not draw distinct rows")
    data = rng.uniform(-1.0, 1.0, size=n * col_nnz)
    indices = rows.reshape(-1).astype(np.int32)
    indptr = (np.arange(n + 1, dtype=np.int64) * col_nnz).astype(np.int32)
    return data, indices, indptr, (m, n)


def csc_matvec(A, x):
    data, indices, indptr, (m, n) = A
    counts = np.diff(indptr)
    xx = np.repeat(x, counts)
    return np.bincount(indices, weights=data * xx, minlength=m)


def csc_rmatvec(A, y):
    data, indices, indptr, (m, n) = A
    prod = data * y[indices]
    cs = np.concatenate(([0.0], np.cumsum(prod)))
    return cs[indptr[1:]] - cs[indptr[:-1]]


# ------------------------------------------------------------------ cone projections (numpy)
def _proj_soc(v):
    if v.size == 0:
        return v
    if v.size == 1:
        return np.maximum(v, 0.0)
    t, s = v[0], float(np.linalg.norm(v[1:]))
    if s <= t:
        return v.copy()
    if s <= -t:
        return np.zeros_like(v)
    a = 0.5 * (s + t)
    out = np.empty_like(v)
    out[0] = a
    out[1:] = v[1:] * (a / s)
    return out


def psd_vec_to_mat(v, k):
    X = np.zeros((k, k))
    idx = np.tril_indices(k)
    # column-major lower triangle: iterate columns
    pos = 0
    for j in range(k):
        ln = k - j
        X[j:, j] = v[pos:pos + ln]
        pos += ln
    X = X + np.tril(X, -1).T
    off = ~np.eye(k, dtype=bool)
    X[off] /= SQRT2
    return X


def psd_mat_to_vec(X, k):
    out = np.empty(k * (k + 1) // 2)
    pos = 0
    for j in range(k):
        ln = k - j
        col = X[j:, j].copy()
        col[1:] *= SQRT2
        out[pos:pos + ln] = col
        pos += ln
    return out


def _proj_psd(v, k):
    if k == 0:
        return v
    if k == 1:
        return np.maximum(v, 0.0)
    X = psd_vec_to_mat(v, k)
    w, V = np.linalg.eigh(X)
    w = np.maximum(w, 0.0)
    return psd_mat_to_vec((V * w) @ V.T, k)


def _proj_box(tx, bl, bu, t0=1.0):
    """Euclidean projection onto {(t, s): t*bl <= s <= t*bu, t >= 0} (reference cones.c:1182-1245)."""
    t, x = tx[0], tx[1:]
    tt = t0
    for _ in range(25):
        tp = tt
        hi = x > tt * bu
        lo = x < tt * bl
        with np.errstate(invalid="ignore"):
            gt = (tt - t) + np.sum(np.where(hi, (tt * bu - x) * bu, 0.0)) + np.sum(np.where(lo, (tt * bl - x) * bl, 0.0))
            ht = 1.0 + np.sum(np.where(hi, bu * bu, 0.0)) + np.sum(np.where(lo, bl * bl, 0.0))
        tt = max(tt - gt / max(ht, 1e-8), 0.0)
        if abs(gt / max(ht, 1e-6)) < 1e-12 * max(tt, 1.0) or abs(tt - tp) < 1e-11 * max(tt, 1.0):
            break
    out = np.empty_like(tx)
    with np.errstate(invalid="ignore"):
        out[1:] = np.where(x > tt * bu, tt * bu, np.where(x < tt * bl, tt * bl, x))
    out[0] = tt
    return out


def proj_cone(v, cone):
    """Euclidean projection onto the primal cone K (zero, LP, box, SOC, PSD)."""
    out = v.copy()
    pos = 0
    z = int(cone.get("z", 0))
    out[pos:pos + z] = 0.0
    pos += z
    l = int(cone.get("l", 0))
    out[pos:pos + l] = np.maximum(v[pos:pos + l], 0.0)
    pos += l
    bu = cone.get("bu")
    if bu is not None and len(bu) > 0:
        bs = len(bu) + 1
        out[pos:pos + bs] = _proj_box(v[pos:pos + bs], np.asarray(cone["bl"], float), np.asarray(bu, float))
        pos += bs
    q = np.asarray(cone.get("q", []) or [], dtype=np.int64)
    if q.size:
        if q.size > 16 and np.all(q == q[0]) and q[0] >= 2:
            qq = int(q[0])
            blk = v[pos:pos + q.size * qq].reshape(q.size, qq)
            t = blk[:, 0]
            s = np.linalg.norm(blk[:, 1:], axis=1)
            a = 0.5 * (s + t)
            res = blk.copy()
            below = (s > t) & (s <= -t)
            proj = (s > t) & ~below
            res[below] = 0.0
            with np.errstate(invalid="ignore", divide="ignore"):
                res[proj, 0] = a[proj]
                res[proj, 1:] = blk[proj, 1:] * (a[proj] / s[proj])[:, None]
            out[pos:pos + q.size * qq] = res.reshape(-1)
            pos += q.size * qq
        else:
            for qq in q:
                out[pos:pos + qq] = _proj_soc(v[pos:pos + qq])
                pos += int(qq)
    for k in (cone.get("s", []) or []):
        ln = int(k) * (int(k) + 1) // 2
        out[pos:pos + ln] = _proj_psd(v[pos:pos + ln], int(k))
        pos += ln
    assert pos == v.size, (pos, v.size)
    return out


def proj_dual_cone(v, cone):
    """Pi_{K*}(v) = v + Pi_K(-v)  (Moreau; reference cones.c:1552-1596 with r_y = NULL)."""
    return v + proj_cone(-v, cone)


# ------------------------------------------------------------------ problem assembly
def make_problem(m, n, col_nnz, cone, seed):
    rng = np.random.default_rng(seed)
    A = random_sparse_csc(m, n, col_nnz, rng)
    zz = rng.uniform(-1.0, 1.0, size=m)
    y = proj_dual_cone(zz, cone)
    s = y - zz
    x = rng.uniform(-1.0, 1.0, size=n)
    b = csc_matvec(A, x) + s
    c = -csc_rmatvec(A, y)
    return {"A": A, "b": b, "c": c, "cone": cone, "x_opt": x, "y_opt": y, "s_opt": s,
            "opt": float(c @ x), "m": m, "n": n, "nnz": int(A[2][-1])}


def socp_cone(m, p_f, p_l, n_soc):
    z = int(np.floor(m * p_f))
    l = int(np.floor(m * p_l))
    rest = m - z - l
    base = rest // n_soc
    q = [base] * n_soc
    q[-1] += rest - base * n_soc
    return {"z": z, "l": l, "q": q}


def config(name, scale=1.0, seed=1234):
    """The named workloads of BASELINE.json (`scale` shrinks n, m, nnz proportionally)."""
    if name == "C1":  # reference demo size: n=1000, m=4000, 32 nnz/col
        n = max(int(1000 * scale), 20)
        m = 4 * n
        col = int(np.ceil(np.sqrt(n)))
        cone = socp_cone(m, 0.1, 0.3, 11)
        return make_problem(m, n, col, cone, seed)
    if name == "C2":  # SOCP n=1e6 m=3e6 nnz=1e7, 50 SOC cones
        n = max(int(1_000_000 * scale), 50)
        m = 3 * n
        cone = socp_cone(m, 0.1, 0.3, 50)
        return make_problem(m, n, 10, cone, seed)
    if name == "C3":  # LP n=5e6 m=5e6 nnz=5e7, box + positive cone
        n = max(int(5_000_000 * scale), 50)
        m = n
        l = m // 2
        bs = m - l
        rng = np.random.default_rng(seed + 1)
        bl = -rng.uniform(0.5, 1.5, size=bs - 1)
        bu = rng.uniform(0.5, 1.5, size=bs - 1)
        cone = {"l": l, "bl": bl, "bu": bu}
        return make_problem(m, n, 10, cone, seed)
    if name == "C4":  # SDP: 200 PSD(100) + 1e5 linear rows, nnz=2e7, n=1e5
        nb = max(int(200 * scale), 2)
        k = 100 if scale >= 0.05 else 12
        l = max(int(100_000 * scale), 10)
        m = l + nb * (k * (k + 1) // 2)
        n = max(int(100_000 * scale), 10)
        col = max(min(int(round(2e7 * scale / n)), m // 2), 2)
        cone = {"l": l, "s": [k] * nb}
        return make_problem(m, n, col, cone, seed)
    if name == "C5":  # mixed SOCP+SDP n=2e6 nnz=3e7, m=6e6
        n = max(int(2_000_000 * scale), 50)
        m = 3 * n
        nb = max(int(100 * scale), 1)
        k = 100 if scale >= 0.05 else 10
        psd_rows = nb * (k * (k + 1) // 2)
        z = int(0.1 * m)
        l = int(0.3 * m)
        rest = m - z - l - psd_rows
        nq = 50
        base = rest // nq
        q = [base] * nq
        q[-1] += rest - base * nq
        cone = {"z": z, "l": l, "q": q, "s": [k] * nb}
        return make_problem(m, n, 15, cone, seed)
    raise KeyError(name)
