#!/usr/bin/env python
"""bench.py -- multipole points/sec of the all-effects signal path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl coffe_b200|reference] [--nz 32] [--nsep 1000]

One "step" = one pass of the hot path over the whole cfg5 grid: 32 z_mean x 1000 separations x l = 0,2,4,
all effects (den, rsd, len, d1, d2, g1..g5), full-sky, = 96 000 multipole points, sharded over the N
GPUs as contiguous blocks of (z, r) work items (no data-path collective; one NCCL gather of the result).

  value  : points/s with the grid and its quadrature items already resident in HBM (kernels only,
           CUDA events on the launching stream, max over ranks)
  e2e    : points/s through the reference-facing C ABI call coffe_b200_multipoles_range with HOST
           buffers: item construction, host->device copies and the device->host copy inside the timed region
  roofline: FP64, for the dominant kernel di_items_smem_kernel (integrated x integrated piece, ~88 % of the step):
           achieved = canonical algorithmic flops (F_DI = 450 per integrand evaluation x the
           evaluations one launch performs) / that kernel's launch duration (CUDA events on its stream, averaged
           over the timed steps); peak = the DFMA probe kernel measured on the same device in the same run;
           `step` repeats the accounting for the whole step (NI 700, SI 1000, DI 450 flop per evaluation)
  covariance: cfg4 (4 z x (3 x 100)^2 entries, flat-sky Gaussian covariance with shot noise) entries/s, rank 0
  cpu_baseline: the reference's own code (oracle/_ref, -O3 -ffast-math like its wheels) on a bounded sample

With --impl reference the reference's CPU implementation is timed instead (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

ALL = ("den", "rsd", "len", "d1", "d2", "g1", "g2", "g3", "g4", "g5")
H = 0.67
F_NI, F_SI, F_DI = 700.0, 1000.0, 450.0        # SURVEY.md 8d canonical accounting
METRIC = "multipole points/sec (z,r,ell; all effects)"


def grid(nz, nsep):
    return np.linspace(0.5, 2.0, nz), np.linspace(10.0, 300.0, nsep) / H, np.array([0, 2, 4], dtype=np.int32)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    def __init__(self, index):
        self.index, self.rows, self.stop, self.thread = index, [], threading.Event(), None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows if len(r) > 3 + i)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "samples": len(self.rows), "reasons": reasons}


def reference_arm(args, z, sep, ls, rank):
    """the reference's own CPU implementation (oracle/_ref, fast build) on a bounded sample of the workload"""
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the reference's OpenMP path gets all the host cores it can use
    if "LOCAL_RANK" in os.environ:
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    from oracle import coffe_oracle as O
    fast = O.available(fast=True)
    ref = O.RefCoffe(fast=fast)
    ref.set_power_spectrum(O.load_pk())
    cores = ref.lib.ref_nthreads(ref.h_)
    # bounded sample: n_s stratified (z, r) pairs x 3 multipoles; every all-effects point costs ~1-2 core-seconds
    n_s = max(1, min(cores, 16) // 3 * 1) if args.sample_points <= 0 else args.sample_points
    iz = np.linspace(0, len(z) - 1, 2).round().astype(int)
    ir = np.linspace(0, len(sep) - 1, max(n_s, 2)).round().astype(int)
    zs, ss = z[iz], sep[ir]
    ref.set_contributions(ALL)
    ref.set_z_mean(zs); ref.set_deltaz([0.3] * len(zs)); ref.set_sep(ss); ref.set_l(ls)
    ref.lib.ref_integrals(ref.h_)           # tables prebuilt, outside the timed region
    n_pts = len(zs) * len(ss) * len(ls)
    times = []
    for i in range(args.warmup + args.steps):
        t = time.perf_counter()
        ref.multipoles()
        if i >= args.warmup:
            times.append(time.perf_counter() - t)
    tot = sum(times)
    v = n_pts * len(times) / tot
    sample = f"{len(zs)} z x {len(ss)} r x {len(ls)} l = {n_pts} points per step, stratified over the cfg5 grid, tables prebuilt"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "points/s", "n_gpus": args.gpus, "steps": len(times),
            "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"cfg5 bulk sweep {len(z)} z_mean x {len(sep)} sep x l=0,2,4, all effects incl. lensing, full-sky (bounded sample)"},
            "cpu_baseline": {"value": v, "unit": "points/s", "cores": int(cores), "kind": "reference" if fast else "reference (-O2)", "sample": sample},
            "e2e": {"value": v, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline(z, sep, ls, budget_points=72):
    from oracle import coffe_oracle as O
    if not O.available(fast=True) and not O.available():
        return None
    fast = O.available(fast=True)
    ref = O.RefCoffe(fast=fast)
    ref.set_power_spectrum(O.load_pk())
    cores = ref.lib.ref_nthreads(ref.h_)
    n_r = max(2, budget_points // 6)
    zs = z[[0, len(z) - 1]]
    ss = sep[np.linspace(0, len(sep) - 1, n_r).round().astype(int)]
    ref.set_contributions(ALL)
    ref.set_z_mean(zs); ref.set_deltaz([0.3] * 2); ref.set_sep(ss); ref.set_l(ls)
    ref.lib.ref_integrals(ref.h_)
    t = time.perf_counter()
    ref.multipoles()
    dt = time.perf_counter() - t
    n = len(zs) * len(ss) * len(ls)
    return {"value": n / dt, "unit": "points/s", "cores": int(cores), "kind": "reference",
            "sample": f"{len(zs)} z x {len(ss)} r x {len(ls)} l = {n} points of the same grid in {dt:.1f} s, "
                      f"reference sources built -O3{' -ffast-math' if fast else ''} -fopenmp, tables prebuilt"}


# dram__bytes_read.sum + dram__bytes_write.sum of one di_items_smem_kernel launch on this very workload (cfg5, one GPU), from
# the ncu --set full capture summarised in profiles/r1_di_kernel_cfg5_v41_ncu.txt: 19.04 MB read + 15 KB written in a 169 ms
# launch.  Algorithmic bytes of that launch: the (point, mu) items in (12 B each), one double per item out, the 6 MB of tables
# once = ~45 MB (the items are expanded on the device by the step's own expand_items_kernel and are still in L2); either way
# ~0.1 GB/s -- the kernel does not stream HBM (L2 hit rate 99.97 %).  Its ceiling is on-chip: the same capture has the L1 /
# shared-memory data pipe at 77 % of peak (81 % before the knots were packed) and the FP64 pipe at 36 %.
TRAFFIC_BYTES_PER_LAUNCH_CFG5_1GPU = 1.9051e7
NCU_PIPES_CFG5_1GPU = {"source": "profiles/r1_di_kernel_cfg5_v41_ncu.txt (ncu --set full of this kernel on this workload)",
                       "l1tex_data_pipe_pct_of_peak": 76.7, "shared_memory_pct_of_peak": 58.2, "fp64_pipe_pct_of_peak": 35.6,
                       "issue_active_pct": 43.0}


def covariance_cfg4(eng, args):
    """cfg4 of BASELINE.json: flat-sky Gaussian covariance of l = 0, 2, 4, 4 z-bins, 100 x 100 separation blocks,
    density + RSD with shot noise.  Host buffers in and out (the C ABI call)."""
    import coffe_b200
    k, pk = np.loadtxt(Path(coffe_b200.__file__).resolve().parent / "data" / "PkL_CLASS.dat", unpack=True)
    k, pk = k * H, pk / H ** 3
    zc = [0.6, 1.0, 1.4, 1.8]
    sepc = np.linspace(20.0, 218.0, 100) / H
    eng.set_options(("den", "rsd"))
    kw = dict(z_mean=zc, deltaz=[0.1] * 4, window=False)
    argsc = ((k, pk), sepc, [0, 2, 4], [0.2] * 4, [1e-3 * H ** 3] * 4, [1e-3 * H ** 3] * 4, [2.0 / H] * 4)
    eng.covariance(*argsc, **kw)
    t0 = time.perf_counter()
    n_rep = 3
    for _ in range(n_rep):
        out = eng.covariance(*argsc, **kw)
    dt = (time.perf_counter() - t0) / n_rep
    ms, launches = eng.last_timing()
    assert np.all(np.isfinite(out)) and out.shape == (4, 3, 100, 3, 100)
    # the reference's other integration method: 2-D FFTLog at its default sampling (8000 x 8000 per multipole pair and power)
    t0 = time.perf_counter()
    out_f = eng.covariance(*argsc, method="fftlog", bins=8000, **kw)
    dt_f = time.perf_counter() - t0
    ms_f, launches_f = eng.last_timing()
    eng.set_options(ALL)
    return {"value": out.size / dt, "unit": "entries/s", "entries": int(out.size), "wall_ms": dt * 1e3, "kernel_ms": ms, "gpu_launches": launches,
            "workload": "cfg4: 4 z x (3 x 100)^2, den + rsd, cosmic + mixed + Poisson, unbinned, covariance_integration_method = standard",
            "fftlog": {"value": out_f.size / dt_f, "unit": "entries/s", "wall_ms": dt_f * 1e3, "kernel_ms": ms_f, "gpu_launches": launches_f,
                       "sampling": 8000}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="coffe_b200")
    ap.add_argument("--nz", type=int, default=32)
    ap.add_argument("--nsep", type=int, default=1000)
    ap.add_argument("--sample-points", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-covariance", action="store_true")
    ap.add_argument("--variant", type=int, default=-1, help="items kernel variant (experiments; default = library default)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    z, sep, ls = grid(args.nz, args.nsep)

    if args.impl == "reference":
        reference_arm(args, z, sep, ls, rank)
        return

    import torch
    import torch.distributed as dist
    import coffe_b200
    from coffe_b200.default_tables import load_default_tableset

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: coffe_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = coffe_b200.Engine(local)
    eng.upload_tables(load_default_tableset())
    eng.set_options(ALL)
    if args.variant >= 0:
        eng.set_kernel_variant(args.variant)

    from coffe_b200 import sharding
    n_items = len(z) * len(sep)
    # contiguous blocks of (z, r) items with equal estimated cost (evaluations counted by the library on a few rows and
    # columns of the grid, outside the timed region); every rank derives the same boundaries
    bounds = sharding.weighted_bounds(sharding.separable_cost(eng, z, sep, ls), world) if world > 1 else [0, n_items]
    begin, end = bounds[rank], bounds[rank + 1]
    n_pts_total = n_items * len(ls)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- kernel-only arm: inputs resident in HBM ------------------------------------------------------
    plan = eng.plan_multipoles(z, sep, ls, begin, end)
    for _ in range(args.warmup):
        plan.execute()
    barrier()
    with ClockSampler(local) as clocks:
        t0 = time.perf_counter()
        dev_ms = 0.0
        si_ms = di_ms = 0.0
        for _ in range(args.steps):
            dev_ms += plan.execute()           # CUDA events on the launching stream, synchronised each step
            a, b = eng.last_kernel_times()
            si_ms += a; di_ms += b
        barrier()
        wall = time.perf_counter() - t0
    launches = eng.last_timing()[1] * args.steps
    res_local = plan.fetch()
    t_dev = torch.tensor([dev_ms, wall * 1e3, si_ms, di_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max, si_ms_max, di_ms_max = t_dev.tolist()
    n_work, n_ni, n_si, n_di = plan.count()
    cnt = torch.tensor([n_work, n_ni, n_si, n_di, launches], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    n_work, n_ni, n_si, n_di, launches = cnt.tolist()

    # one gather of the result slices (NCCL over NVLink); checks the sharded result is complete
    full = sharding.gather_ranges(res_local, n_items, dist, device="cuda", bounds=bounds) if world > 1 else res_local
    plan.close()

    # ---- end-to-end arm: the C ABI call with host buffers ---------------------------------------------
    for _ in range(min(args.warmup, 1)):
        eng.multipoles_range(z, sep, ls, begin, end)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = eng.multipoles_range(z, sep, ls, begin, end)
    barrier()
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_s = t_e.item()
    # bytes per step crossing PCIe: per point z, sep, rule index and the two item offsets; per separation the two mu rules
    # (nodes + weights; the (point, mu, weight) items themselves are expanded on the device); per-point pieces out
    h2d = int((end - begin) * 28 + n_work / max(len(z), 1) * 16 + (len(sep) + 1) * 8)
    d2h = int((end - begin) * 3 * len(ls) * 8)

    peak_tf = eng.fp64_peak_tflops()
    if rank == 0:
        assert np.all(np.isfinite(full)) and full.shape == (len(ls), n_items)
        assert np.array_equal(out, res_local), "C-ABI call and prepared plan must agree bit for bit"
        flops_per_step = F_NI * n_ni + F_SI * n_si + F_DI * n_di
        achieved_step = flops_per_step / world / (dev_ms_max / args.steps * 1e-3) / 1e12
        # dominant kernel: di_items_smem_kernel, one launch per step and rank
        di_launch_ms = di_ms_max / args.steps
        achieved = F_DI * n_di / world / (di_launch_ms * 1e-3) / 1e12 if di_launch_ms > 0 else None
        value = n_pts_total * args.steps / (dev_ms_max * 1e-3)
        cov = None
        if not args.no_covariance:
            try:
                cov = covariance_cfg4(eng, args)
            except Exception as e:
                cov = {"value": None, "unit": "entries/s", "error": str(e)}
        line = {
            "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"cfg5 bulk sweep {len(z)} z_mean x {len(sep)} sep x l=0,2,4, all effects incl. lensing, full-sky",
                       "points": n_pts_total, "work_items_z_r_mu": int(n_work), "integrand_evaluations": {"NI": n_ni, "SI": n_si, "DI": n_di},
                       "tables": "DEFAULT_SETTINGS cosmology, PkL_CLASS.dat (coffe_b200/data/default_tables.npz)",
                       "sharding": f"contiguous cost-balanced blocks of (z, r) items over {world} GPU(s), tables replicated, one NCCL gather",
                       "block_items": [bounds[r + 1] - bounds[r] for r in range(world)],
                       "l2_policy": "compute-bound FP64 kernel; working set (tables 2.6 MB) is L2-resident by design, no HBM streaming to flush",
                       "quadrature": eng.get_quadrature(), "wall_ms_per_step": wall_ms_max / args.steps},
            "roofline": {"bound": "fp64", "kernel": "di_items_smem_kernel", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved / peak_tf if peak_tf and achieved else None,
                         "traffic": TRAFFIC_BYTES_PER_LAUNCH_CFG5_1GPU if (world == 1 and args.nz == 32 and args.nsep == 1000) else None,
                         "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, profiles/r1_di_kernel_cfg5_v41_ncu.txt",
                         "ncu_pipes": NCU_PIPES_CFG5_1GPU if (world == 1 and args.nz == 32 and args.nsep == 1000) else None,
                         "peak_source": "DFMA probe kernel measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                         "launch_ms": di_launch_ms, "share_of_step": di_ms_max / dev_ms_max if dev_ms_max else None,
                         "flops_per_launch": F_DI * n_di / world, "accounting": "SURVEY.md 8d: 450 flop per DI integrand evaluation",
                         "step": {"achieved": achieved_step, "frac": achieved_step / peak_tf if peak_tf else None, "flops": flops_per_step / world,
                                  "si_kernel_ms": si_ms_max / args.steps, "accounting": "700/1000/450 flop per NI/SI/DI evaluation"}},
            "e2e": {"value": n_pts_total * args.steps / e2e_s, "unit": "points/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "clocks": clocks.summary(),
        }
        if cov is not None:
            line["covariance"] = cov
        if not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(z, sep, ls)
            except Exception as e:       # the baseline is a report, never a reason to lose the GPU numbers
                line["cpu_baseline"] = {"value": None, "unit": "points/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
