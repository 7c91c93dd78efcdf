"""Time kmcb200_spmv on the 8x8 stand-in matrix (ORDER=file|xsorted|lex|brick[B]); checks y against numpy."""
import os, sys, importlib, numpy as np, torch
sys.path.insert(0, '.')
PKG = "accelerated-kinetic-monte-carlo-simulations-of-atomistically-resolved-resistive-memory-arrays_b200"
kmc = importlib.import_module(PKG)
syn = importlib.import_module(PKG + ".synthetic")
ctx = kmc.Context(0)
s = syn.crossbar_standin("tests/golden/5nm_device/parameters.txt", 8, 8, order=os.environ.get("ORDER", "file"))
x, y, z = ctx.dev_d(s.x), ctx.dev_d(s.y), ctx.dev_d(s.z)
K0 = ctx.initialize_sparsity_K(x, y, z, s.lattice, s.pbc, s.nn_dist, s.N_left, s.N_right)
h = K0.to_host()
n, nnz = K0.rows, K0.nnz
rp = h["row_ptr"].astype(np.int64); col = h["col"]
val = np.random.default_rng(0).standard_normal(nnz)
xh = np.random.default_rng(1).standard_normal(n)
# reference result (float64 numpy, different summation order: compare loosely)
rows = np.repeat(np.arange(n), np.diff(rp))
yref = np.zeros(n); np.add.at(yref, rows, val * xh[col])
K = ctx.kmat_from_csr(ctx.dev_i(rp.astype(np.int32)), ctx.dev_i(col), ctx.dev_d(val))
xv = ctx.dev_d(xh); yv = ctx.empty_d(n, 0.0)
for _ in range(3): ctx.spmv(K, xv, yv)
torch.cuda.synchronize()
err = np.abs(yv.cpu().numpy() - yref).max()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(30): ctx.spmv(K, xv, yv)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 30
print(f"spmv {ms*1e3:8.1f} us  {(12.0*nnz+20.0*n)/ms/1e6:8.1f} GB/s   max|y-yref|={err:.2e}")
