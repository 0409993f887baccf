#!/usr/bin/env python3
"""Diagnostic (build with QCD_NVCC_EXTRA=-DQCD_TRACE): per-block phase time stamps of the two-phase update kernel.
Prints the median cycles a block spends before the generator phase (state loads, addresses), in the generator phase and in
the physics phase, and how many blocks of the same SM are in the generator phase together."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    m = bench.build_model((64, 64, 64, 64))
    g = m.GPU0
    for _ in range(2):
        m.lattice_sweep()
    g.synchronize()
    # one more X-direction pass: even X is the 5th launch of a sweep; run a full sweep, the table keeps the last X launch (even)
    m.lattice_sweep()
    g.synchronize()
    n = 65536
    buf = np.zeros(5 * n, dtype=np.uint64)
    r = g.L.qgpu_debug_trace(buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.size))
    assert r == 0, r
    t = buf.reshape(n, 5).astype(np.int64)
    sm = t[:, 4]
    wait, gen, phys = t[:, 1] - t[:, 0], t[:, 2] - t[:, 1], t[:, 3] - t[:, 2]
    print("median cycles: head %d  generator %d  physics %d  total %d" %
          (np.median(wait), np.median(gen), np.median(phys), np.median(t[:, 3] - t[:, 0])))
    print("   percentiles generator", np.percentile(gen, [10, 50, 90]).astype(int), "physics", np.percentile(phys, [10, 50, 90]).astype(int),
          "wait", np.percentile(wait, [10, 50, 90]).astype(int))
    # overlap: for each block, how many other blocks of the same SM have a generator interval overlapping its generator midpoint
    over = []
    for s in np.unique(sm)[:16]:
        idx = np.flatnonzero(sm == s)
        a, b = t[idx, 1], t[idx, 2]
        mid = (a + b) // 2
        over.extend(((a[None, :] <= mid[:, None]) & (b[None, :] > mid[:, None])).sum(axis=1))
    print("   blocks of the same SM in the generator phase at a block's generator midpoint (incl. itself): mean %.2f" % np.mean(over))
    m.close()


if __name__ == "__main__":
    main()
