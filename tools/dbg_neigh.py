import sys, importlib, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from conftest import make_synthetic, PKG
kmc = importlib.import_module(PKG)
from oracle import binding as orc
ctx = kmc.Context(0)
s = make_synthetic(kmc)
x, y, z = ctx.dev_d(s.x), ctx.dev_d(s.y), ctx.dev_d(s.z)
for nn, rs, rc in ((52, 0, s.N), (6, 17, 200), (3, s.N - 5, 5)):
    got = ctx.compute_neighbor_list(x, y, z, 3.5, nn, rs, rc).cpu().numpy()
    want = orc.neighbor_list(s.x, s.y, s.z, 3.5, nn, rs, rc)
    bad = np.where((got != want).any(1))[0]
    print(nn, rs, rc, "bad rows", len(bad))
    for b in bad[:3]:
        print(" row", b, "got", got[b], "want", want[b])
