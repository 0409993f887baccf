#!/usr/bin/env python3
"""Diagnostic (build with QCD_NVCC_EXTRA=-DQCD_WS_TRACE): time line of the warp-specialised update kernel, per unit (one generator warp
and its three physics warps) and round: where the generator warp and the physics warps spend their cycles and who waits for whom."""
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
    for _ in range(3):
        m.lattice_sweep()
    g.synchronize()
    n = 592 * 64 * 24
    buf = np.zeros(n, dtype=np.uint64)
    r = g.L.qgpu_debug_ws_trace(buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.size))
    assert r == 0, r
    t = buf.reshape(592, 64, 24).astype(np.int64)
    t = t[:, 4:60, :]                       # steady state rounds
    med = lambda x: int(np.median(x))
    print("generator warp, median cycles per round:")
    print("  ticket %d  state loads %d  wait for free buffers %d  park (to full signal) %d  rest of the pass + store %d  | round %d" % (
        med(t[:, :, 1] - t[:, :, 0]), med(t[:, :, 2] - t[:, :, 1]), med(t[:, :, 3] - t[:, :, 2]), med(t[:, :, 4] - t[:, :, 3]),
        med(t[:, :, 5] - t[:, :, 4]), med(t[:, 1:, 0] - t[:, :-1, 0])))
    for c in range(3):
        b = 6 + 4 * c
        print("physics warp %d: staples %d  wait for samples %d  heat bath %d  | item %d, wait for the next ticket %d" % (
            c, med(t[:, :, b + 1] - t[:, :, b]), med(t[:, :, b + 2] - t[:, :, b + 1]), med(t[:, :, b + 3] - t[:, :, b + 2]),
            med(t[:, :, b + 3] - t[:, :, b]), med(t[:, 1:, b] - t[:, :-1, b + 3])))
    u = t[5]
    print("unit 5, rounds 4..9, cycles relative to the generator's round start (gen: ticket/loaded/free/parked/stored; phys c: ticket/staples/samples/done):")
    for r_ in range(6):
        base = u[r_, 0]
        print("  ", [int(x - base) for x in u[r_, 1:6]], [[int(x - base) for x in u[r_, 6 + 4 * c:10 + 4 * c]] for c in range(3)])
    m.close()


if __name__ == "__main__":
    main()
