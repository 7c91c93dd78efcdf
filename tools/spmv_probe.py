"""SpMV bottleneck probe: same CSR shape, different column patterns (gather locality)."""
import sys, importlib, numpy as np, torch
sys.path.insert(0, '.')
PKG = "accelerated-kinetic-monte-carlo-simulations-of-atomistically-resolved-resistive-memory-arrays_b200"
kmc = importlib.import_module(PKG)
syn = importlib.import_module(PKG + ".synthetic")
ctx = kmc.Context(0)
order = sys.argv[1] if len(sys.argv) > 1 else "file"
s = syn.crossbar_standin("tests/golden/5nm_device/parameters.txt", 8, 8, order=order)
x, y, z = ctx.dev_d(s.x), ctx.dev_d(s.y), ctx.dev_d(s.z)
K = ctx.initialize_sparsity_K(x, y, z, s.lattice, s.pbc, s.nn_dist, s.N_left, s.N_right)
h = K.to_host()
n = K.rows; nnz = K.nnz
rp = h["row_ptr"]; col = h["col"]
rows = np.repeat(np.arange(n, dtype=np.int32), np.diff(rp))
val = np.random.default_rng(0).standard_normal(nnz)
def run(name, c):
    Kc = ctx.kmat_from_csr(ctx.dev_i(rp), ctx.dev_i(c), ctx.dev_d(val))
    xv = ctx.empty_d(n, 1.0); yv = ctx.empty_d(n, 0.0)
    for _ in range(3): ctx.spmv(Kc, xv, yv)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): ctx.spmv(Kc, xv, yv)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print(f"{order:8s} {name:28s} {ms*1e3:8.1f} us  {(12.0*nnz+20.0*n)/ms/1e6:8.1f} GB/s")
    Kc.close()
run("real columns", col)
run("col = row (1 line/row)", rows)
run("col = row + k (contiguous)", np.minimum(rows + (np.arange(nnz) - np.repeat(rp[:-1], np.diff(rp))).astype(np.int32), n - 1).astype(np.int32))
run("col = 0 (single address)", np.zeros(nnz, np.int32))
run("random columns", np.random.default_rng(1).integers(0, n, nnz).astype(np.int32))
