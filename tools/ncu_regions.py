#!/usr/bin/env python3
"""Where a kernel's time goes, from the source page of an ncu capture (`ncu --set full --import-source on`, built with -lineinfo):

    python tools/ncu_regions.py gpurun_out/x.ncu-rep [instructions per bin] > profiles/rNN_regions_x.txt

Bins the kernel's SASS in program order and prints, per bin, its share of the warp stall samples (= of the warps' life time), the
instructions a warp executes in it and the leading stall reasons.  For the two-phase update kernel the bins fall into prologue,
generator phase (IADD3.X chains), staples (DFMA blocks between TMA waits) and the Cabibbo-Marinari / Kennedy-Pendleton loop."""
import csv
import io
import subprocess
import sys

KEYS = ["stall_wait", "stall_no_inst", "stall_long_sb", "stall_short_sb", "stall_math", "stall_dispatch", "stall_not_selected",
        "stall_selected", "stall_branch_resolving"]


def main():
    rep = sys.argv[1]
    step = int(sys.argv[2]) if len(sys.argv) > 2 else 250
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    name, hdr, data = rows[0][1], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, KeyError, IndexError):
            return 0.0
    tot = sum(f(r, "# Samples") for r in data)
    grid_warps = max(f(r, "Instructions Executed") for r in data[:40]) or 1.0      # the first instructions run once per warp
    print(name)
    print(f"{len(data)} SASS instructions, {tot:.0f} stall samples, {sum(f(r, 'Instructions Executed') for r in data) / grid_warps:.0f} instructions per warp")
    print("  first  samples  instr/warp  " + "  ".join(k.replace("stall_", "")[:9].rjust(9) for k in KEYS) + "   first instruction of the bin")
    for a in range(0, len(data), step):
        blk = data[a:a + step]
        s = sum(f(r, "# Samples") for r in blk)
        if s / tot < 5e-4:
            continue
        e = sum(f(r, "Instructions Executed") for r in blk) / grid_warps
        print("%7d  %6.2f%%  %10.0f  " % (a, 100 * s / tot, e) + "  ".join("%8.1f%%" % (100 * sum(f(r, k) for r in blk) / tot) for k in KEYS)
              + "   " + blk[0][ix["Source"]].strip()[:44])


if __name__ == "__main__":
    main()
