#!/usr/bin/env python3
"""Diagnostic: lattice after one sweep, TMA-staged two-phase kernel vs the buffered kernel (plain loads); prints where they differ."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_parity import make_model

dims = tuple(int(x) for x in sys.argv[1].split("x"))
tabs = []
for fused in (False, True):
    m = make_model(3, dims, 1, ints=0, prng="RANLUX3", fused=True, iters=2, beta=5.7)
    if not fused:
        m.GPU0.set_fused_prng(False)
    m.lattice_start()
    m.lattice_sweep()
    tabs.append(m.table_numpy())
    rowsize = m.lattice_table_row_size
    m.close()
a, b = tabs
d = (a != b)
print(dims, "differing reals:", int(d.sum()), "of", d.size)
if d.any():
    idx = np.argwhere(d.reshape(12, rowsize, 4).any(axis=2))
    n1, n2, n3, n4 = dims
    seen = 0
    for plane, gid in idx[:12]:
        x = gid // (n2 * n3 * n4); r = gid % (n2 * n3 * n4); t = r // (n2 * n3); r %= n2 * n3; z = r // n2; y = r % n2
        print("  plane", plane, "dir", plane % 4, "site", (x, y, z, t), "max diff", float(np.abs(a.reshape(12, rowsize, 4)[plane, gid] - b.reshape(12, rowsize, 4)[plane, gid]).max()))
    ys = idx[:, 1] % n2
    print("  y histogram of differing sites:", np.bincount(ys, minlength=n2)[:8], "...", np.bincount(ys, minlength=n2)[-8:])
    print("  dirs:", np.bincount(idx[:, 0] % 4, minlength=4))
