"""CPU emulation of the PUSH-based "sharded-x" multi-GPU CG iteration (scs_b200/csrc/kernels/cg.cu:
k_cgx_iteration + cg_iteration_shard_x, scs_b200/csrc/kernels/spmv.cu B200_HOOK_P2P_ROUTE). Every rank is a Python
thread; the exchange allocations (inbox [G][S], the p vector), flag lines and scalar slots are shared numpy arrays laid
out like the device buffers. Data moves only by STORES into the peer's arrays, exactly as on the device: the K2 SpMV
pushes each output row into the owner's inbox as a SELF-VALIDATING element (value + the sequence number of the exchange,
common.cuh ll_store: no "partial ready" flag exists, the reader waits per element), the partial scalars travel the same
way, and so do the new p slices (into the peers' p-boxes, unpacked by the receiver): no flag and no fence anywhere. The
emulation follows the kernel phase by phase -- same rank-order sums, same stop logic, NO double buffering of inbox,
messages or p, writes of one push deliberately split in two halves with a delay between them -- and checks: no deadlock, no torn reads under random delays (ranks run up to an iteration apart), all
ranks hold bit-identical p / alpha / beta / stop decisions, and the iterates agree with a plain single-process
preconditioned CG on the same operator (reference linsys/cpu/indirect/private.c:133-217).
It validates the PROTOCOL (what the blocks of one rank do together is one thread here); the CUDA-level parts
(fences, co-residency) are for the multi-GPU run (tests/mgpu_check.py)."""
import threading
import time

import numpy as np
import pytest
import scipy.sparse as sp


class Rank(threading.Thread):
    def __init__(self, me, G, n, A_g, d_g, rx, M, b, shared, iters, tol, max_its):
        super().__init__(daemon=True)
        self.me, self.G, self.n = me, G, n
        self.A, self.d, self.rx, self.M = A_g, d_g, rx, M
        self.sh = shared
        self.iters_to_run, self.tol, self.max_its = iters, tol, max_its
        self.lo, self.hi = n * me // G, n * (me + 1) // G
        # k_cg_init (replicated): r = b, x = 0, z = M r, p = z
        self.x = np.zeros(n)
        self.r = b.copy()
        self.z = self.r * M
        self.p = shared["p"][me]            # the rank's p lives in its exchange allocation: peers store into it
        self.p[:] = self.z
        self.Gp = np.zeros(n)
        self.ctl = dict(ztr=float(self.z @ self.r), rnorm=float(np.abs(self.r).max()), iters=0, done=0,
                        alpha=0.0, beta=0.0)
        self.error = None
        self.history = []
        self.jitter = np.random.default_rng(1000 + me) if shared.get("jitter") else None

    def pause(self):
        """random delays at the phase boundaries: ranks run up to a whole iteration apart"""
        if self.jitter is not None and self.jitter.random() < 0.5:
            time.sleep(self.jitter.random() * 2e-3)

    def wait(self, first, skip_rank, seq):
        line = self.sh["flags"][self.me]
        t0 = time.time()
        for q in range(self.G):
            if q == skip_rank:
                continue
            while line[first + q] < seq:
                if time.time() - t0 > 20:
                    raise TimeoutError(f"rank {self.me} waiting for slot {first + q} >= {seq}")
                time.sleep(0)

    def scal(self, owner, rnd, parity, frm):
        base = ((rnd * 2 + parity) * 8 + frm) * 2
        return self.sh["scal"][owner], base

    def iteration(self, seq):
        G, me, n, sh = self.G, self.me, self.n, self.sh
        if self.ctl["done"]:
            return
        lo, hi = self.lo, self.hi
        S = (n + G - 1) // G
        # K1 / K2: the rank's partial, every output row PUSHED into the inbox of its owner (lane `me`) as (value, seq);
        # no flag follows. The push is split in two halves with a pause in between: readers must cope with rows that
        # have not arrived yet
        part = self.A.T @ (self.d * (self.A @ self.p))
        for half in (0, 1):
            for q in range(G):
                qlo, qhi = n * q // G, n * (q + 1) // G
                mid = (qhi - qlo) // 2
                a, b = (0, mid) if half == 0 else (mid, qhi - qlo)
                sh["inbox_val"][q][me * S + a:me * S + b] = part[qlo + a:qlo + b]
                sh["inbox_seq"][q][me * S + a:me * S + b] = seq
            self.pause()
        ztr_old, iters_old = self.ctl["ztr"], self.ctl["iters"]
        # 1: reduce my slice from MY inbox in rank order, waiting per element for rows still in flight
        s = np.zeros(hi - lo)
        for q in range(G):
            self.wait_elems(sh["inbox_seq"][me], q * S, q * S + (hi - lo), seq)
            s = s + sh["inbox_val"][me][q * S:q * S + (hi - lo)]
        self.Gp[lo:hi] = self.rx[lo:hi] * self.p[lo:hi] + s
        mine = float(self.p[lo:hi] @ self.Gp[lo:hi])
        for q in range(G):
            sh["msg_val"][q][0 * 16 + me * 2] = mine
            sh["msg_seq"][q][0 * 16 + me * 2] = seq
        self.pause()
        # 2: all partial scalars -> alpha
        pGp = 0.0
        for q in range(G):
            self.wait_elems(sh["msg_seq"][me], q * 2, q * 2 + 1, seq)
            pGp += sh["msg_val"][me][q * 2]
        alpha = ztr_old / pGp
        self.pause()
        # 3: K3 on the slice
        self.x[lo:hi] += alpha * self.p[lo:hi]
        self.r[lo:hi] -= alpha * self.Gp[lo:hi]
        self.z[lo:hi] = self.r[lo:hi] * self.M[lo:hi]
        t0 = float(self.z[lo:hi] @ self.r[lo:hi])
        t1 = float(np.abs(self.r[lo:hi]).max()) if hi > lo else 0.0
        for q in range(G):
            base = 16 + me * 2
            sh["msg_val"][q][base], sh["msg_val"][q][base + 1] = t0, t1
            sh["msg_seq"][q][base], sh["msg_seq"][q][base + 1] = seq, seq
        self.pause()
        # 4: z'r, ||r||_inf, stop decision, beta
        ztr, rn = 0.0, 0.0
        for q in range(G):
            base = 16 + q * 2
            self.wait_elems(sh["msg_seq"][me], base, base + 2, seq)
            ztr += sh["msg_val"][me][base]
            rn = max(rn, sh["msg_val"][me][base + 1])
        done, beta = 0, 0.0
        if rn < self.tol:
            done = 1
        elif ztr_old == 0.0:
            done = 1
        else:
            beta = ztr / ztr_old
            if iters_old + 1 >= self.max_its:
                done = 1
        self.pause()
        # 5: K4 on the slice: into the local p and, as (value, seq) elements, into every peer's p-box (two halves)
        if not done:
            pn = self.z[lo:hi] + beta * self.p[lo:hi]
            self.p[lo:hi] = pn
            mid = (hi - lo) // 2
            for a, b in ((0, mid), (mid, hi - lo)):
                for q in range(G):
                    if q != me:
                        sh["pbox_val"][q][lo + a:lo + b] = pn[a:b]
                        sh["pbox_seq"][q][lo + a:lo + b] = seq
                self.pause()
        self.ctl.update(ztr=ztr, rnorm=rn, iters=iters_old + 1, alpha=alpha, beta=beta if not done else self.ctl["beta"],
                        done=done)
        self.history.append((alpha, beta, ztr, rn, done))
        # 6: unpack the peers' slices from my p-box, waiting per element
        if done:
            return
        for q in range(G):
            if q == me:
                continue
            qlo, qhi = n * q // G, n * (q + 1) // G
            self.wait_elems(sh["pbox_seq"][me], qlo, qhi, seq)
            self.p[qlo:qhi] = sh["pbox_val"][me][qlo:qhi]

    def wait_elems(self, seqarr, a, b, seq):
        t0 = time.time()
        while not np.all(seqarr[a:b] == seq):
            if time.time() - t0 > 20:
                raise TimeoutError(f"rank {self.me} waiting for elements [{a}, {b}) of sequence {seq}")
            time.sleep(0)

    def run(self):
        try:
            for it in range(self.iters_to_run):
                self.iteration(it + 1)         # the host increments seq before every iteration, starting at 1
        except Exception as e:  # noqa: BLE001 -- reported by the test
            self.error = e


def serial_pcg(Afull, d, rx, M, b, iters, tol, max_its):
    n = len(b)
    x, r = np.zeros(n), b.copy()
    z = r * M
    p = z.copy()
    ztr = float(z @ r)
    hist = []
    for it in range(iters):
        Gp = rx * p + Afull.T @ (d * (Afull @ p))
        alpha = ztr / float(p @ Gp)
        x += alpha * p
        r -= alpha * Gp
        z = r * M
        ztr_new, rn = float(z @ r), float(np.abs(r).max())
        hist.append((alpha, ztr_new, rn))
        if rn < tol or ztr == 0.0:
            break
        beta = ztr_new / ztr
        ztr = ztr_new
        if it + 1 >= max_its:
            break
        p = z + beta * p
    return x, hist


@pytest.mark.parametrize("jitter", [False, True])
@pytest.mark.parametrize("G,n,m,iters,tol", [(2, 40, 150, 12, 0.0), (4, 103, 400, 25, 0.0), (8, 64, 300, 30, 0.0),
                                             (3, 50, 200, 200, 1e-9), (5, 7, 60, 40, 1e-12)])
def test_sharded_x_protocol(G, n, m, iters, tol, jitter):
    rng = np.random.default_rng(G * 1000 + n)
    A = sp.random(m, n, density=min(1.0, 6.0 / n), format="csr", random_state=np.random.RandomState(n)) + \
        sp.csr_matrix((np.ones(min(m, n)), (np.arange(min(m, n)), np.arange(min(m, n)))), shape=(m, n))
    A = sp.csr_matrix(A)
    d = 1.0 / rng.uniform(0.5, 2.0, m)                 # R_y^-1
    rx = np.full(n, 1e-3)
    Mdiag = 1.0 / (rx + np.asarray((A.multiply(A)).T @ d).ravel())
    b = rng.standard_normal(n)
    offs = [m * g // G for g in range(G + 1)]           # contiguous row blocks
    S = (n + G - 1) // G
    shared = {"inbox_val": [np.zeros(G * S) for _ in range(G)], "inbox_seq": [np.zeros(G * S, dtype=np.int64) for _ in range(G)],
              "msg_val": [np.zeros(32) for _ in range(G)], "msg_seq": [np.zeros(32, dtype=np.int64) for _ in range(G)],
              "p": [np.zeros(n) for _ in range(G)],
              "pbox_val": [np.zeros(n) for _ in range(G)], "pbox_seq": [np.zeros(n, dtype=np.int64) for _ in range(G)],
              "flags": [np.zeros(64, dtype=np.uint64) for _ in range(G)],
              "scal": [np.zeros(64) for _ in range(G)], "jitter": jitter}
    max_its = 10 * n
    ranks = [Rank(g, G, n, A[offs[g]:offs[g + 1]], d[offs[g]:offs[g + 1]], rx, Mdiag, b, shared, iters, tol, max_its)
             for g in range(G)]
    for r in ranks:
        r.start()
    for r in ranks:
        r.join(timeout=60)
    for r in ranks:
        assert not r.is_alive(), "deadlock: a rank is still waiting on a flag"
        assert r.error is None, r.error
    # all ranks: identical scalars and decisions, bit for bit, and identical p after every completed iteration
    for r in ranks[1:]:
        assert r.history == ranks[0].history
        assert np.array_equal(np.asarray(r.p), np.asarray(ranks[0].p))
        assert r.ctl == ranks[0].ctl
    # x is owned by slices: assemble it and compare with the single-process PCG
    x = np.concatenate([r.x[r.lo:r.hi] for r in ranks])
    xs, hist = serial_pcg(A, d, rx, Mdiag, b, iters, tol, max_its)
    steps = len(ranks[0].history)
    if tol > 0:
        # CG trajectories are not reproducible across summation orders (the step lengths of the two runs agree to
        # 1e-16 for the first iterations, drift apart in the ill-conditioned middle phase and meet again): the
        # stop test may fire a few iterations apart, the solutions agree
        assert ranks[0].ctl["done"] == 1 and hist[-1][2] < tol and abs(steps - len(hist)) <= 5
        assert np.abs(x - xs).max() <= 1e-7 * max(1.0, np.abs(xs).max())
    else:
        assert steps == len(hist) == iters and ranks[0].ctl["done"] == 0
        assert np.abs(x - xs).max() <= 1e-9 * max(1.0, np.abs(xs).max())
        a_mine = np.array([h[0] for h in ranks[0].history])[:3]
        a_ser = np.array([h[0] for h in hist])[:3]
        assert np.abs(a_mine - a_ser).max() <= 1e-12 * np.abs(a_ser).max()
    # the reduced system (R_x + A' R_y^-1 A) x = b is solved as far as the single-process PCG got
    res = rx * x + A.T @ (d * (A @ x)) - b
    assert np.abs(res).max() <= 10 * hist[-1][2] + 1e-9 * max(1.0, np.abs(b).max())
