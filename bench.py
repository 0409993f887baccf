#!/usr/bin/env python3
"""bench.py -- SU(3) link updates/s of the gauge-update hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

One step = one full heat-bath sweep (8 (parity, direction) passes with their RANLUX3 random numbers +
reunitarisation, which the SU(3) update kernels apply to the links they write: qgpu_set_fold_reunitarisation)
of the workload below.  N = 1: BASELINE config C5, SU(3) 64^4 fp64, beta = 6, NHIT = 10,
RANLUX3, cold start.  N > 1 (torchrun): BigLattice x-slab split, one 64-slice slab per GPU (weak scaling,
global lattice (64 N) x 64^3), halo exchange after every pass.

The JSON line also carries: `roofline` for the update kernel (algorithmic bytes 4.5 links x 96 B per updated
link = 864 B, SURVEY.md 8d, against MEASURED_PEAKS.json), `e2e` (the same sweeps driven through the C ABI
with host-side inputs/outputs copied inside the timed region), `cpu_baseline` (the UNMODIFIED reference's
CPU_RUN program on this box's host cores) and `clocks`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(group=3, dims=(64, 64, 64, 64), precision=1, beta=6.0, nhit=10, prng="RANLUX3", randseries=54, ints=1)
BYTES_PER_LINK_UPDATE = 9 * 96      # SU(3) 2-row fp64: 4 V link reads + V/2 link writes per pass, 8 passes / 4 V links
NCU_DRAM_BYTES_PER_LAUNCH = 7.985194e9 + 1.903730e9   # profiles/r02_ncu_update_jump.json
FLOPS_PER_LINK_UPDATE = 4700


def workload_config(n_gpus: int, dims=None) -> dict:
    """`config` of the JSON line: the workload, named the same way by both arms (native and --impl reference) for a given N."""
    d = tuple(dims) if dims else WORKLOAD["dims"]
    table_gb = 4 * d[0] * d[1] * d[2] * d[3] * 96 / 1e9
    if n_gpus <= 1:
        w = (f"SU(3) {'x'.join(map(str, d))} fp64 heat-bath sweeps (8 passes + reunitarisation), beta={WORKLOAD['beta']}, NHIT={WORKLOAD['nhit']}, "
             f"RANLUX3, cold start")
    else:
        full = (d[0] * n_gpus, d[1], d[2], d[3])
        w = (f"SU(3) {'x'.join(map(str, full))} fp64 heat-bath sweeps (8 passes + reunitarisation), BigLattice x-slab split over {n_gpus} GPUs "
             f"({d[0]} slices per GPU, halo exchange after each pass), beta={WORKLOAD['beta']}, NHIT={WORKLOAD['nhit']}, RANLUX3, cold start")
    return {"workload": w, "l2_policy": f"inputs_larger_than_L2 (link table {table_gb:.2f} GB per GPU vs 126 MB L2)"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------- reference arm
def run_reference_cpu(n_procs: int, ls: int, sweeps: int):
    """Time the UNMODIFIED reference (oracle/_ref/QCDGPU_cpu, its sequential CPU_RUN configuration) on
    `n_procs` host cores, one independent SU(3) ls^4 run per core.  Returns (link updates/s, seconds)."""
    binary = os.path.join(ROOT, "oracle", "_ref", "QCDGPU_cpu")
    if not os.path.exists(binary):
        return None, None
    tmp = tempfile.mkdtemp(prefix="qcdref_")
    init = ("PLATFORM = 0\nDEVICE = 0\nFINISHPATH=./\nOUTPUTPATH=./\nGROUP = 3\nND = 4\n"
            f"LS = {ls}\nLT = {ls}\nITER = {sweeps + 1}\nNITER = 1\nNHIT = 10\nBETA = 6.0\nNAV = 0\nINTS = 1\n"
            "PRNG = RANLUX3\nRANDSERIES = 54\n")
    dirs = []
    for i in range(n_procs):
        d = os.path.join(tmp, str(i))
        os.makedirs(d)
        open(os.path.join(d, "init.dat"), "w").write(init.replace("RANDSERIES = 54", f"RANDSERIES = {54 + i}"))
        dirs.append(d)
    t0 = time.time()
    procs = [subprocess.Popen([binary, "init.dat"], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for d in dirs]
    for p in procs:
        p.wait()
    dt = time.time() - t0
    links = 4 * ls ** 4 * sweeps * n_procs
    return links / dt, dt


def reference_main(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref/QCDGPU_cpu = the UNMODIFIED reference built
    in its CPU_RUN configuration), on all host cores.  One step = one bounded sample of the workload: every core runs an independent
    SU(3) 12^4 lattice for 2 sweeps (link updates/s is volume independent to first order; 64^4 does not fit the reference's int sizes,
    SURVEY.md 0 row 6).  W warm-up steps, then exactly K timed steps; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    ls, sweeps = 12, 2
    if run_reference_cpu(1, 4, 1)[0] is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/QCDGPU_cpu missing (run make -C oracle ref where /root/reference exists)"}))
        return
    for _ in range(args.warmup):
        run_reference_cpu(cores, ls, sweeps)
    t0 = time.time()
    for _ in range(args.steps):
        run_reference_cpu(cores, ls, sweeps)
    dt = time.time() - t0
    links_per_step = 4 * ls ** 4 * sweeps * cores
    v = links_per_step * args.steps / dt
    sample = (f"reference CPU_RUN binary, per step {cores} independent SU(3) {ls}^4 cold-start runs (one per core), {sweeps} sweeps each incl. "
              f"start-up and per-sweep plaquette+action measurement; {args.steps} steps in {dt:.1f} s")
    out = {"impl": "reference", "metric": "SU(3) link updates/sec", "value": v, "unit": "link-updates/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(int(os.environ.get("WORLD_SIZE", str(args.gpus))), tuple(int(x) for x in args.dims.split("x")) if args.dims else None),
           "cpu_baseline": {"value": v, "unit": "link-updates/s", "cores": cores, "kind": "reference", "sample": sample},
           "e2e": {"value": v, "unit": "link-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


# ---------------------------------------------------------------------------------------------- native arm
def build_model(dims, fused=True, device=0, group=None, precision=None, phi=0.0, or_steps=0, hot=False):
    from qcdgpu_b200 import Model
    m = Model()
    m.desired_device = device
    m.lattice_group = WORKLOAD["group"] if group is None else group
    m.lattice_full_size = list(dims)
    m.precision = WORKLOAD["precision"] if precision is None else precision
    m.PHI = phi
    m.or_steps = or_steps
    if m.lattice_group == 2 and group is not None:
        m.BETA = 2.5
    m.ints = 0 if hot else WORKLOAD["ints"]
    if not (m.lattice_group == 2 and group is not None):
        m.BETA = WORKLOAD["beta"]
    m.NHIT = WORKLOAD["nhit"]
    m.PRNG0.PRNG_generator = WORKLOAD["prng"]
    m.PRNG0.PRNG_randseries = WORKLOAD["randseries"]
    m.fused_prng = fused
    m.PL_level = 0
    m.ITER = 4096
    m.lattice_init()
    m.lattice_start()
    return m


def native_single(args):
    import numpy as np
    dims = tuple(int(x) for x in args.dims.split("x")) if args.dims else WORKLOAD["dims"]
    m = build_model(dims, fused=not args.buffered, group=args.group, precision=args.precision, phi=args.phi, or_steps=args.or_steps, hot=args.hot)
    diag = args.group is not None or args.precision is not None or args.phi != 0.0 or args.or_steps != 0 or args.hot
    g = m.GPU0
    if args.materialize:
        g.set_fused_prng(False)      # PRNG_instances == PRNGSTEP, but every produce() writes the PRN buffer
    V = m.lattice_domain_site
    links_per_step = 4 * V
    for _ in range(args.warmup):
        m.lattice_sweep()
    g.synchronize()
    # ---- device-timed throughput (inputs resident in HBM)
    clocks = ClockSampler(0)
    clocks.start()
    l0 = g.launch_count()
    # an event pair per launch, recorded without waiting: per-kernel times INSIDE the timed region.  Only where a launch is long against
    # the two event records (64^4: 3.5 ms); on the small diagnostic lattices (16^4: 13 us per pass) they would be half of what is timed.
    if V >= (1 << 20):
        g.set_debug(profiling=2)
    g.timer_start()
    for _ in range(args.steps):
        m.lattice_sweep()
    ms = g.timer_stop()
    g.profile_collect()
    g.set_debug(profiling=False)
    launches = g.launch_count() - l0
    tk_in = [g.kernel_get_execution_time(k) for k in m.sun_update_id.values()]
    upd_ns_in = sum(t[0] * t[2] for t in tk_in) / max(1, sum(t[2] for t in tk_in))
    value = links_per_step * args.steps / (ms * 1e-3)
    # ---- e2e: the run loop as a user drives it through the C ABI -- per step the host-side inputs (parameter
    # buffer) go H2D from pinned memory, the sweep and the plaquette measurement run, the result comes back D2H
    params = g.buffer_map(m.lattice_parameters, np.uint8)
    h2d = int(params.size)
    d2h = 16
    t0 = time.perf_counter()
    plq = None
    for i in range(args.steps):
        g._chk(g.L.qgpu_buffer_unmap(g.h, m.lattice_parameters, params.ctypes.data), "unmap")   # H2D
        m.lattice_sweep()
        g.kernel_run(m.sun_measurement_plq_id)
        g.kernel_init_constant_reset(m.sun_measurement_plq_reduce_id, 1, m.argument_plq_index)
        g.kernel_run(m.sun_measurement_plq_reduce_id)
        p = g.L.qgpu_buffer_map_range(g.h, m.lattice_energies_plq, 16, 16)                     # D2H (+sync)
        plq = np.frombuffer((__import__("ctypes").c_char * 16).from_address(p), dtype=np.float64).copy()
    e2e_s = time.perf_counter() - t0
    e2e = links_per_step * args.steps / e2e_s
    ck = clocks.stop()
    # ---- dominant kernel: average launch duration of the update kernel, CUDA events around each launch
    g.set_debug(profiling=True)
    for _ in range(2):
        m.lattice_sweep()
    g.set_debug(profiling=False)
    tk = [g.kernel_get_execution_time(k) for k in m.sun_update_id.values()]
    # the statistics accumulate: take out what the timed region contributed
    iso_sum = sum(t[0] * t[2] for t in tk) - sum(t[0] * t[2] for t in tk_in)
    iso_n = sum(t[2] for t in tk) - sum(t[2] for t in tk_in)
    upd_ns_iso = iso_sum / max(1, iso_n)
    upd_ns = upd_ns_in if upd_ns_in > 0 else upd_ns_iso          # the roofline line is quoted on the kernel as it runs in the timed region
    gs_ns = g.kernel_get_execution_time(m.sun_GramSchmidt_id)[0]
    peak, peak_kind = measured_peaks()
    link_bytes = (12 if m.lattice_group == 3 else 4) * (8 if m.precision == 1 else 4)
    bytes_per_launch = 4.5 * V * link_bytes
    achieved = bytes_per_launch / (upd_ns * 1e-9) / 1e9
    # DRAM bytes of one launch of that kernel from the committed ncu --set full capture (same command, 64^4 fused only)
    traffic, traffic_src = None, None
    if tuple(dims) == (64, 64, 64, 64) and not args.buffered and not args.materialize and not diag:
        traffic, traffic_src = NCU_DRAM_BYTES_PER_LAUNCH, "profiles/r02_ncu_update_jump.json (dram__bytes_read.sum + dram__bytes_write.sum)"
    regs = g.update_kernel_registers(m.program_update, 0, 2 if not args.buffered else 0)
    kname = ("update_kernel_rlx<%d,%s,MU,%s,223>" if not args.buffered else "update_kernel<%d,%s,MU,%s,PrnBuffered>") % (
        m.lattice_group, "double" if m.precision == 1 else "float", "true" if args.phi != 0.0 else "false")
    if not args.buffered and m.lattice_group == 3 and dims[1] % 64 == 0 and (V // 2) % 128 == 0 and os.environ.get("QCDGPU_TMA", "1") != "0":
        kname = kname.replace("update_kernel_rlx<3,", "update_kernel_rlx_tma<")      # links staged by TMA (update.cuh, Su3StageTma)
    plq_total = float(plq.sum() / (6.0 * V)) if plq is not None else None
    # ---- CPU baseline: the reference itself, 1 core, bounded sample
    cpu_v, cpu_dt = (None, None) if args.no_cpu else run_reference_cpu(1, 16, 8)
    out = {
        "metric": "SU(3) link updates/sec", "value": value, "unit": "link-updates/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64" if m.precision == 1 else "f32", "data": "synthetic",
        "config": (workload_config(1, dims) if not diag and not args.buffered and not args.materialize else
                   {"workload": f"DIAGNOSTIC: SU({m.lattice_group}) {'x'.join(map(str, dims))} {'fp64' if m.precision == 1 else 'fp32'} heat-bath sweeps (8 passes + reunitarisation), beta={m.BETA}, "
                                f"NHIT=10, RANLUX3 ({'fused in-register' if not args.buffered else 'buffered'}), {'hot' if args.hot else 'cold'} start" + (f", {args.or_steps} overrelaxation sweeps after each heat-bath sweep" if args.or_steps else ""),
                    "l2_policy": (f"inputs_larger_than_L2 (link table {4 * V * link_bytes / 1e9:.2f} GB vs 126 MB L2)" if 4 * V * link_bytes > 126e6
                                  else f"diagnostic configuration: link table {4 * V * link_bytes / 1e6:.1f} MB fits the L2, no flush")}),
        "details": {"sweeps_per_s": args.steps / (ms * 1e-3), "plaquette_after_run": plq_total, "update_kernel_registers": regs,
                    "prng": "RANLUX3, one generator per update thread, run inside the update kernel (fused)" if not args.buffered else "RANLUX3, stand-alone producer + PRN buffer"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_kind,
                     "kernel": kname,
                     "avg_launch_us": upd_ns / 1e3, "algorithmic_bytes_per_launch": bytes_per_launch,
                     "avg_launch_us_how": "CUDA events around each of the update launches of the timed region, on the launch stream, no waiting in between "
                                          "(board at its power cap); avg_launch_us_one_by_one: the same kernel with a wait after every launch (unthrottled clocks)",
                     "avg_launch_us_one_by_one": upd_ns_iso / 1e3, "frac_one_by_one": bytes_per_launch / (upd_ns_iso * 1e-9) / 1e9 / peak if upd_ns_iso > 0 else None,
                     "fp64_gflops": FLOPS_PER_LINK_UPDATE * (V / 2) / (upd_ns * 1e-9) / 1e9,
                     "reunitarisation_launch_us": gs_ns / 1e3,
                     "reunitarisation": ("folded into the update kernels (lattice_GramSchmidt kernel_run skipped: the table already holds its result)"
                                         if gs_ns < 50e3 else "separate lattice_GramSchmidt launch per sweep")},   # an event pair around no launch: a few us
        "e2e": {"value": e2e, "unit": "link-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "what": "per step: parameter buffer H2D, sweep, plaquette measurement + reduction, result D2H, all through the C ABI"},
        "gpu_launches": int(launches),
        "clocks": ck,
        "cpu_baseline": ({"value": cpu_v, "unit": "link-updates/s", "cores": 1, "kind": "reference",
                          "sample": f"oracle/_ref/QCDGPU_cpu (unmodified reference, CPU_RUN build), SU(3) 16^4 cold start, 8 sweeps incl. start-up and per-sweep measurements, {cpu_dt:.1f} s"}
                         if cpu_v else {"value": None, "unit": "link-updates/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}),
    }
    print(json.dumps(out))
    m.close()


def native_on(args):
    """Diagnostic (BASELINE config 4): O(4) scalar model, Metropolis sweeps (odd + even pass) with XOR128 in registers.
    Prints site updates/s; the bound is the FP64 pipe (4 log + exp + sqrt per proposal, NHIT proposals per site), not HBM."""
    import numpy as np
    from qcdgpu_b200.on_model import OnModel
    dims = tuple(int(x) for x in args.dims.split("x")) if args.dims else (32, 32, 32, 32)
    m = OnModel()
    m.lattice_full_size = list(dims)
    m.precision = 1 if args.precision is None else args.precision
    m.ints, m.NHIT, m.ITER = 0, 10, 64
    m.get_acceptance_rate = not args.no_acc
    m.fused_prng = not args.buffered
    m.PRNG0.PRNG_generator, m.PRNG0.PRNG_randseries = "XOR128", 45
    m.lattice_init()
    m.lattice_start()
    g = m.GPU0
    for _ in range(args.warmup):
        m.lattice_update()
    g.synchronize()
    l0 = g.launch_count()
    g.timer_start()
    for _ in range(args.steps):
        m.lattice_update()
    ms = g.timer_stop()
    launches = g.launch_count() - l0
    g.set_debug(profiling=True)
    for _ in range(3):
        m.lattice_update()
    g.set_debug(profiling=False)
    tk = [g.kernel_get_execution_time(k)[0] for k in m.sun_update_id.values()]
    m.lattice_measure()
    m.ITER_counter += 1
    A = m.lattice_analysis()
    V = m.lattice_full_site
    # CPU baseline: the reference has no CPU implementation of this update; the C restatement (1 core) on a 12^4 sample
    cpu = None
    if not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O
        L = O.OnLattice((12, 12, 12, 12), precision="f64" if m.precision == 1 else "f32")
        L.init_cold()
        rng = np.random.default_rng(1)
        prns = rng.random((6 * L.prngstep, 4), dtype=np.float32)
        t0 = time.perf_counter(); n = 0
        while time.perf_counter() - t0 < 5.0:
            L.update(prns, 1); L.update(prns, 0); n += 1
        cpu = n * L.sites / (time.perf_counter() - t0)
    out = {"metric": "O(4) site updates/sec (diagnostic, not the BASELINE metric)", "value": V * args.steps / (ms * 1e-3), "unit": "site-updates/s",
           "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "dtype": "f64" if m.precision == 1 else "f32",
           "config": {"workload": f"O(4) scalar model {'x'.join(map(str, dims))}, NHIT=10 multi-hit Metropolis, XOR128 ({'in registers' if m.fused_prng else 'buffered'}), hot start",
                      "acceptance_sums": m.get_acceptance_rate, "S_after_run": float(A["S"]["data"][1]), "phi_after_run": float(A["Phi"]["data"][1])},
           "update_kernel_us": [t / 1e3 for t in tk], "proposals_per_s": V * 10 * args.steps / (ms * 1e-3), "gpu_launches": int(launches),
           "cpu_baseline": {"value": cpu, "unit": "site-updates/s", "cores": 1, "kind": "port", "sample": "oracle/on_oracle.c, 12^4, 5 s"}}
    print(json.dumps(out))
    m.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--dims", default="")
    ap.add_argument("--buffered", action="store_true", help="reference-layout PRN buffer instead of the fused generator")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--group", type=int, default=None, help="diagnostic: 2 or 3 (default: the BASELINE workload, SU(3))")
    ap.add_argument("--precision", type=int, default=None, help="diagnostic: 0 single, 1 double, 2 mixed")
    ap.add_argument("--phi", type=float, default=0.0, help="diagnostic: twist angle (turns -D TBC on)")
    ap.add_argument("--materialize", action="store_true", help="one generator per update thread, PRN buffer written (diagnostic)")
    ap.add_argument("--hot", action="store_true", help="diagnostic: hot start (INTS=0: every link random) instead of the cold start of the BASELINE workload")
    ap.add_argument("--or-steps", type=int, default=0, help="diagnostic (BASELINE config 2): overrelaxation sweeps after each heat-bath sweep")
    ap.add_argument("--on", action="store_true", help="diagnostic: the O(N) scalar branch (BASELINE config 4) instead of the gauge sweep")
    ap.add_argument("--no-acc", action="store_true", help="diagnostic (--on): without the -D ACC_RATE acceptance sums")
    args = ap.parse_args()
    if args.on:
        return native_on(args)
    if args.impl == "reference":
        return reference_main(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 or args.gpus > 1:
        from qcdgpu_b200 import biglattice
        return biglattice.bench_main(args, WORKLOAD)
    return native_single(args)


if __name__ == "__main__":
    main()
