"""CPU, world_size = 2 over gloo: the row-sharded KKT operator of SURVEY 8(e).

Each rank keeps the row block of A that the library's own partitioner
(scs_b200_row_partition, host logic in the C-ABI library) assigns to it, applies
the local part of  G p = R_x p + sum_g A_g' R_g^-1 A_g p , and the n-vector is
summed with an all-reduce -- the exact exchange pattern the CUDA path runs over
NCCL (kernels/cg.cu mat_vec_sharded). The result must equal the unsharded operator."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from scs_b200 import capi, problems
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lib = capi.load()
    rng = np.random.default_rng(5)            # same matrix on every rank
    m, n = 900, 250
    A = problems.random_sparse_csc(m, n, 7, rng)
    data, idx, ptr, _ = A
    offs = np.zeros(world + 1, dtype=np.int32)
    assert lib.scs_b200_row_partition(m, n, capi.iptr(np.ascontiguousarray(ptr)), capi.iptr(np.ascontiguousarray(idx)),
                                      world, capi.iptr(offs)) == 0
    assert offs[0] == 0 and offs[-1] == m and np.all(np.diff(offs) > 0)
    r0, r1 = int(offs[rank]), int(offs[rank + 1])
    rx = np.full(n, 1e-6)
    ry = np.full(m, 10.0)
    ry[:90] = 0.01
    p = rng.standard_normal(n)
    # local block: entries with row in [r0, r1)
    mask = (idx >= r0) & (idx < r1)
    col = np.repeat(np.arange(n), np.diff(ptr))
    z_loc = np.bincount(idx[mask] - r0, weights=data[mask] * p[col[mask]], minlength=r1 - r0) / ry[r0:r1]
    part = np.bincount(col[mask], weights=data[mask] * z_loc[idx[mask] - r0], minlength=n)
    t = torch.from_numpy(part)
    dist.all_reduce(t)                         # the one collective of a CG iteration
    Gp = rx * p + t.numpy()
    full = rx * p + problems.csc_rmatvec(A, problems.csc_matvec(A, p) / ry)
    err = np.abs(Gp - full).max() / np.abs(full).max()
    # all-gather of the y block (back-substitution y = R_y^-1 (A x - r_y))
    y_loc = torch.from_numpy(z_loc.copy())
    sizes = [int(offs[r + 1] - offs[r]) for r in range(world)]
    outs = [torch.zeros(s, dtype=torch.float64) for s in sizes]
    dist.all_gather(outs, y_loc) if len(set(sizes)) == 1 else None
    if len(set(sizes)) != 1:
        for r in range(world):
            buf = y_loc if r == rank else torch.zeros(sizes[r], dtype=torch.float64)
            dist.broadcast(buf, src=r)
            outs[r] = buf
    y_full = torch.cat(outs).numpy()
    err_y = np.abs(y_full - problems.csc_matvec(A, p) / ry).max()
    # balance: nonzeros per rank within 25 % of the mean
    nnz_loc = int(mask.sum())
    q.put((rank, float(err), float(err_y), nnz_loc, len(data)))
    dist.destroy_process_group()


def test_row_sharded_operator_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, err_y, nnz_loc, nnz in res:
        assert err <= 1e-13, (rank, err)
        assert err_y <= 1e-13, (rank, err_y)
        assert abs(nnz_loc - nnz / 2) <= 0.25 * nnz / 2


def test_row_partition_edge_cases():
    sys.path.insert(0, ROOT)
    from scs_b200 import capi
    lib = capi.load()
    # all nonzeros in one row, more ranks than busy rows, empty rows
    n, m = 6, 10
    ptr = np.arange(n + 1, dtype=np.int32)
    idx = np.full(n, 3, dtype=np.int32)
    for world in (1, 2, 4, 8):
        offs = np.zeros(world + 1, dtype=np.int32)
        assert lib.scs_b200_row_partition(m, n, capi.iptr(ptr), capi.iptr(idx), world, capi.iptr(offs)) == 0
        assert offs[0] == 0 and offs[-1] == m
        assert np.all(np.diff(offs) >= 0)


@pytest.mark.parametrize("threads", ["1", "3", "16"])
def test_row_block_extraction_is_thread_count_independent(threads, monkeypatch):
    """every rank cuts its row block out of the user's CSC on the host (threaded over column ranges of equal nnz,
    host/linsys_b200.c restrict_rows) before the device builder takes over: same arrays as numpy for any thread count,
    and the per-row counts behind scs_b200_row_partition (threaded, atomic integer counts) give the same offsets."""
    sys.path.insert(0, ROOT)
    from scs_b200 import capi, problems
    lib = capi.load()
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    lib.scs_b200_restrict_rows_csc.restype = C.c_int
    lib.scs_b200_restrict_rows_csc.argtypes = [C.c_int, ip, ip, dp, C.c_int, C.c_int, C.POINTER(ip), C.POINTER(ip), C.POINTER(dp)]
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    rng = np.random.default_rng(5)
    for m, n, per in ((50, 20, 3), (90000, 60000, 10), (300, 70000, 8)):
        data, idx, ptr, _ = problems.random_sparse_csc(m, n, per, rng)
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        ptr = np.ascontiguousarray(ptr, dtype=np.int32)
        data = np.ascontiguousarray(data)
        monkeypatch.setenv("SCS_B200_HOST_THREADS", "1")
        base = np.zeros(4, dtype=np.int32)
        assert lib.scs_b200_row_partition(m, n, capi.iptr(ptr), capi.iptr(idx), 3, capi.iptr(base)) == 0
        monkeypatch.setenv("SCS_B200_HOST_THREADS", threads)
        offs = np.zeros(4, dtype=np.int32)
        assert lib.scs_b200_row_partition(m, n, capi.iptr(ptr), capi.iptr(idx), 3, capi.iptr(offs)) == 0
        assert np.array_equal(offs, base)
        cols = np.repeat(np.arange(n), np.diff(ptr))
        for g in range(3):
            r0, r1 = int(offs[g]), int(offs[g + 1])
            lp, li, lx = ip(), ip(), dp()
            assert lib.scs_b200_restrict_rows_csc(n, ptr.ctypes.data_as(ip), idx.ctypes.data_as(ip), data.ctypes.data_as(dp),
                                                  r0, r1, C.byref(lp), C.byref(li), C.byref(lx)) == 0
            keep = (idx >= r0) & (idx < r1)
            want_p = np.concatenate([[0], np.cumsum(np.bincount(cols[keep], minlength=n))])
            got_p = np.ctypeslib.as_array(lp, (n + 1,)).copy()
            cnt = int(got_p[-1])
            got_i = np.ctypeslib.as_array(li, (max(cnt, 1),))[:cnt].copy()
            got_x = np.ctypeslib.as_array(lx, (max(cnt, 1),))[:cnt].copy()
            for q in (lp, li, lx):
                libc.free(C.cast(q, C.c_void_p))
            assert np.array_equal(got_p, want_p)
            assert np.array_equal(got_i, idx[keep] - r0) and np.array_equal(got_x, data[keep])
