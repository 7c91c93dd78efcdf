"""The C-side bootstrap channel of the multi-GPU path (kmcb200_rdv_*: all-gather / barrier of the node's ranks through a
shared directory) -- host logic, runs without a GPU."""
import ctypes as C
import importlib
import multiprocessing as mp
import os
import sys

import numpy as np

from conftest import PKG, ROOT


def _worker(rank, size, d, q):
    sys.path.insert(0, ROOT)
    kmc = importlib.import_module(PKG)
    lib = kmc.load_library()
    h = C.c_void_p()
    assert lib.kmcb200_rdv_open(d.encode(), rank, size, C.byref(h)) == 0
    ok = True
    for rnd, nbytes in enumerate((64, 1, 100003, 8)):      # IPC handle, barrier byte, a need map, a small record
        mine = np.full(nbytes, (rank * 37 + rnd) % 251, dtype=np.uint8)
        allb = np.zeros(size * nbytes, dtype=np.uint8)
        rc = lib.kmcb200_rdv_allgather(h, mine.ctypes.data_as(C.c_void_p), nbytes, allb.ctypes.data_as(C.c_void_p))
        ok = ok and rc == 0 and all((allb[r * nbytes:(r + 1) * nbytes] == (r * 37 + rnd) % 251).all() for r in range(size))
        ok = ok and lib.kmcb200_rdv_barrier(h) == 0
    lib.kmcb200_rdv_close(h)
    q.put((rank, ok))


def test_rendezvous_allgather_and_barrier(tmp_path):
    size = 3
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, size, str(tmp_path / "rdv"), q)) for r in range(size)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(size))
    for p in ps:
        p.join(timeout=60)
    assert res == [(r, True) for r in range(size)]
    assert len(os.listdir(tmp_path / "rdv")) <= size * 3          # old rounds are garbage collected


def test_rendezvous_times_out_instead_of_hanging(kmc, tmp_path):
    lib = kmc.load_library()
    os.environ["KMCB200_RDV_TIMEOUT_S"] = "0.5"
    try:
        h = C.c_void_p()
        assert lib.kmcb200_rdv_open(str(tmp_path / "lonely").encode(), 0, 2, C.byref(h)) == 0
        b = np.zeros(2, dtype=np.uint8)
        rc = lib.kmcb200_rdv_allgather(h, b.ctypes.data_as(C.c_void_p), 1, b.ctypes.data_as(C.c_void_p))
        assert rc == -6 and b"waited" in lib.kmcb200_last_error()      # KMCB200_E_COMM: the peer never showed up
        lib.kmcb200_rdv_close(h)
    finally:
        del os.environ["KMCB200_RDV_TIMEOUT_S"]
