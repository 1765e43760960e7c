#!/usr/bin/env python
"""Headline benchmark: DF inlet planes/s @512^2 x 3 (+ PODFS snapshots/s) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # this implementation
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

A DF "step" is one time shard of the round-jet inlet of BASELINE config 3 (512 x 512 x 3, nf = 10,
ln = 5, circular-hyperbolic-tangent profile => adapt2d): 2500 planes per GPU, i.e. the 20 000-step run
of config 3 at 8 GPUs (weak scaling: per-GPU work fixed).  The random field is generated on the
device (Philox, inside the timed region), filtered in time and in the plane, Lund-transformed and
written once to HBM.  The PODFS part runs BASELINE config 4's per-GPU row shard (393 216 rows x
10 000 snapshots, 100 modes): mean removal, FP64 tensor-core Gram (+ NCCL all-reduce for N > 1),
eigen-solve, projection, Fourier fit.

One JSON line is printed by rank 0.  Timing: CUDA events on the launching stream, barrier +
synchronize on both sides, max over ranks.  Every step writes 15.7 GB of output (>> 126 MB L2), so
no L2 flush is needed between iterations.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DF_CHUNK = 1256      # planes per pass of the time loop (2500-plane shard = 2 passes; see profiles/r01_chunk_sweep.json)
CFG3 = dict(J=512, K=512, nf=10, ln=5.0, profile="circular-hyperbolic-tangent", U0=1.0, u_dash=0.02, ring=0.5,
            planes_per_gpu=2500, total_steps_at_8_gpus=20000)
CFG4 = dict(rows_per_gpu=393216, snapshots=10000, modes=100, et=0.9, dt=1.0e-3)
FP64_DMMA_PEAK_TFLOPS = 37.0   # measured on this pool's B200 with tools/probe_fp64.cu (profiles/probe_fp64_r01.json)
FP64_DFMA_PEAK_TFLOPS = 34.0


# stdout carries exactly ONE JSON line: everything else any library prints there (the NCCL version banner, the
# reference-style progress prints) is diverted to stderr for the whole run.
_REAL_STDOUT = None


def _divert_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        sys.stdout = sys.stderr


def _emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def kplane_traffic(nf, planes_per_launch):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (nf = 10 only), scaled to
    this run's planes per launch; None when no capture matches."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_kplane_traffic.json")) as f:
            d = json.load(f)
        if nf != 10:
            return None
        return d["dram_bytes_per_launch"] * planes_per_launch / d["planes_per_launch"]
    except Exception:
        return None


def gram_traffic(M, N):
    """DRAM bytes per launch of the Gram kernel from the committed ncu capture (bench size only)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_gram_traffic.json")) as f:
            d = json.load(f)
        return d["dram_bytes_per_launch"] if (d["M"], d["N"]) == (M, N) else None
    except Exception:
        return None


def df_algorithmic(J, K, nfx, nfy, nfz):
    """bytes and flops per output plane."""
    Jp, Kp = J + 2 * nfy, K + 2 * nfz
    b_plane_kernel = 24.0 * (Jp * Kp + J * K)       # K-plane: T1 read once + output written once
    b_time_kernel = 24.0 * Jp * Kp                  # K-time (generated mode): T1 written once
    flops = 6.0 * ((2 * nfy + 1) * J * Kp + (2 * nfz + 1) * J * K + (2 * nfx + 1) * J * K) + 12.0 * J * K + 18.0 * J * K
    flops_plane_kernel = 6.0 * ((2 * nfz + 1) * Jp * K + (2 * nfy + 1) * J * K) + 18.0 * J * K
    return b_plane_kernel, b_time_kernel, flops, flops_plane_kernel


class ClockSampler(object):
    """nvidia-smi clocks line of B200_PROFILING.md, sampled during the timed region."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.rows = []          # (time, sm_mhz, reasons bitmask, power_w)
        self.smax = None
        self.stop_flag = False
        self.mode = None
        self.proc = None
        try:                    # NVML in-process: a sample every few ms (the timed region is ~0.1-0.3 s)
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.mode = "nvml"
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.mode = None
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.mode = "nvidia-smi"
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv, h = self.nv, self.h
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                try:
                    watts = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                except Exception:
                    watts = None
                self.rows.append((time.time(), mhz, mask, watts))
            except Exception:
                pass
            time.sleep(0.003)

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            if len(f) < 9:
                continue
            try:
                self.smax = float(f[2])
                mask = 0
                for (name, bit), v in zip(self.REASONS, f[5:9]):
                    if v.lower().startswith("active"):
                        mask |= bit
                self.rows.append((time.time(), float(f[1]), mask, float(f[3])))
            except ValueError:
                continue

    def stop(self, t0, t1):
        if self.mode is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        if self.mode == "nvidia-smi":
            time.sleep(0.15)
            self.proc.terminate()
        self.stop_flag = True
        pad = 0.0 if self.mode == "nvml" else 0.15
        rows = [r for r in self.rows if t0 - 0.02 <= r[0] <= t1 + pad]
        reasons = set()
        for _, _, mask, _ in rows:
            for name, bit in self.REASONS:
                if mask & bit:
                    reasons.add(name)
        sm = [r[1] for r in rows]
        watts = [r[3] for r in rows if r[3] is not None]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": float(min(sm)) if sm else None,
                "sm_max_mhz": self.smax, "reasons": sorted(reasons), "samples": len(sm), "source": self.mode,
                "power_w_median": float(np.median(watts)) if watts else None}


def us_per_plane_fmt(x):
    return float("%.4f" % x)


def cfg3_lund(pkg):
    c = CFG3
    df = pkg.digitalfilters
    U, uu, vv, ww, uw = df.build_profile(c["profile"], "top-hat", c["U0"], c["u_dash"], c["K"])
    return pkg._lund.lund_2d(U, uu, vv, ww, uw, c["J"], c["K"], c["profile"], c["ring"])


# --------------------------------------------------------------------------------------------------
def cpu_df_baseline(nplanes, loops=True):
    """The oracle port of the reference's per-step path (DF:1403-1477) on the host, config 3 shape."""
    from oracle import df_oracle as dfo
    c = CFG3
    J, K, nf, ln = c["J"], c["K"], c["nf"], c["ln"]
    U, uu, vv, ww, uw = dfo.build_profile(c["profile"], "top-hat", c["U0"], c["u_dash"], K)
    rs = np.random.RandomState(2)
    X = rs.uniform(-dfo.SQRT3, dfo.SQRT3, (3, nplanes + 2 * nf, J + 2 * nf, K + 2 * nf))
    lund = ("2d", U, uu, vv, ww, uw, c["profile"], c["ring"])
    tm = {}
    t0 = time.perf_counter()
    dfo.generate(X, nplanes, J, K, ln, ln, ln, nf, nf, nf, lund, rotate=(1.0, 0.0, 0.0), loops=loops, timers=tm)
    dt = time.perf_counter() - t0
    return nplanes / dt, dt, tm


def cpu_pod_baseline(M, N, nm):
    """The oracle port of POD + Fourier fit (POD:1294-1333, 1523-1593) on the host, reduced sample."""
    from oracle import pod_oracle as podo
    A = podo.synthetic_snapshots(M, N, seed=1)
    t0 = time.perf_counter()
    mean = A.mean(1)
    A -= mean[:, None]
    t1 = time.perf_counter()
    r = podo.pod(A, N, M // 3, 3, "false", [], 1.0e-15, nm)
    t2 = time.perf_counter()
    podo.fourier_coefficients(r["temporal_modes"], N, r["nm"], CFG4["dt"], CFG4["et"], direct=True)
    t3 = time.perf_counter()
    return N / (t3 - t0), {"mean_s": t1 - t0, "gram_eig_project_s": t2 - t1, "fourier_loop_s": t3 - t2}


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; the Python-2
    reference cannot run here) on the box's host cores.  Rank 0 only."""
    if rank != 0:
        return
    import multiprocessing
    ncores = multiprocessing.cpu_count()
    rates = []
    for i in range(args.warmup + args.steps):
        if i >= 1 and i < args.warmup:
            continue        # one warm-up plane is plenty for a CPU loop; keeps the run within minutes
        r, dt, tm = cpu_df_baseline(1, loops=True)
        if i >= args.warmup:
            rates.append((r, dt))
        if sum(d for _, d in rates) > 150:
            break
    nsteps = len(rates)
    total = sum(d for _, d in rates)
    value = nsteps / total
    pod = None
    if not args.no_pod:
        pv, pt = cpu_pod_baseline(24576, 1000, 20)
        pod = {"metric": "pod_snapshots_per_s", "value": pv, "unit": "snapshots/s",
               "sample": "oracle POD+Fourier on a 24576 x 1000 snapshot matrix, 20 modes (numpy BLAS threads = all cores, "
                         "Python DFT loop single-threaded), not extrapolated", "parts": pt}
    line = {
        "impl": "reference", "metric": "df_planes_per_s", "value": value, "unit": "planes/s", "n_gpus": args.gpus,
        "steps": nsteps, "warmup": args.warmup, "ms_per_step": 1e3 * total / nsteps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg3 round-jet inlet 512x512x3, nf=10, ln=5, adapt2d: one plane per step on the host CPU",
                   "timing": "host wall clock"},
        "cpu_baseline": {"value": value, "unit": "planes/s", "cores": 1, "kind": "port",
                         "sample": "%d planes of config 3 through oracle/df_oracle.py generate(loops=True): scipy direct "
                                   "convolutions + the reference's per-point Python loops (single-threaded, as the reference is); "
                                   "host has %d cores" % (nsteps, ncores)},
        "e2e": {"value": value, "unit": "planes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "pod": pod,
    }
    _emit(line)



# --------------------------------------------------------------------------------------------------
CFG1 = dict(J=64, K=64, nf=10, ln=5.0, ns=1000, nm=20, et=0.9, profile="hyperbolic-tangent", U0=1.0, u_dash=0.02, res=0.1)


def run_cfg1(args, pkg, torch, dev, local_rank, cpu_too=True):
    """BASELINE config 1 IN FULL (64 x 64 x 3, 1000 steps, half-width 10, ln = 5, built tanh profile, 20 modes, et = 0.9):
    `digitalfilters.main()`'s compute -- DF:1403-1510 without the file writers -- through the device-resident pipeline,
    results (mean, 20 modes, energies, Fourier coefficients) copied to the host inside the timed region; next to the
    oracle port of the reference on the host cores, faithful (per-point Python loops) and vectorised, with the BLAS
    thread count at all cores and at 1."""
    capi, df, pipeline = pkg._capi, pkg.digitalfilters, pkg.pipeline
    c = CFG1
    J, K, nf, ln, ns, nm = c["J"], c["K"], c["nf"], c["ln"], c["ns"], c["nm"]
    U, uu, vv, ww, uw = df.build_profile(c["profile"], "top-hat", c["U0"], c["u_dash"], K)
    dt = c["res"] / np.mean(U[U != 0])
    taps = df._taps(nf, ln)

    def ours(seed):
        plan = capi.DfPlan(J, K, nf, nf, nf, taps, taps, taps, device=local_rank)
        plan.set_coefficients(*pkg._lund.pack(*pkg._lund.lund_1d(U, uu, vv, ww, uw, J, K), J, K))   # H2D of the inputs
        res = pipeline.df_to_podfs(plan, ns, nm, dt, c["et"], seed=seed, device_index=local_rank)
        out = dict(mean=res["mean"].cpu().numpy(), Phi=res["Phi"].cpu().numpy(), energy=res["energy"], T=res["T"],
                   c_count=res["c_count"], route=res.get("route"))
        plan.close()
        return out

    ours(1)                                                      # warm-up (workspaces, cuSOLVER handle if route 0)
    torch.cuda.synchronize(dev)
    reps = 5
    each = []                                                    # (plan creation / destruction allocate and free device
    t0 = time.perf_counter()                                     #  memory every time: 16 ms to 60 ms per run were seen)
    for r in range(reps):
        t1 = time.perf_counter()
        o = ours(2 + r)
        torch.cuda.synchronize(dev)
        each.append(time.perf_counter() - t1)
    wall = (time.perf_counter() - t0) / reps
    d2h = o["mean"].nbytes + o["Phi"].nbytes + 8 * ns
    out = {"workload": "cfg1 in full: 64x64x3 planes, 1000 steps, nf=10, ln=5, adapt1d + identity rotation, PODFS 20 modes, et=0.9",
           "ours": {"wall_s": wall, "wall_each_s": each, "wall_median_s": float(np.median(each)), "planes_per_s": ns / wall, "snapshots_per_s": ns / wall, "reps": reps,
                    "api": "DfPlan + pipeline.df_to_podfs (what digitalfilters.main() runs when not verbose), results on the host",
                    "h2d_bytes": 5 * J * K * 8, "d2h_bytes": int(d2h), "modes_kept": int(len(o["c_count"])),
                    "fourier_coefficients_kept": int(np.sum(o["c_count"]))}}
    if cpu_too:
        from oracle import df_oracle as dfo
        from oracle import pod_oracle as podo
        import multiprocessing
        try:
            from threadpoolctl import threadpool_limits
        except Exception:
            threadpool_limits = None
        rs = np.random.RandomState(0)
        X = rs.uniform(-dfo.SQRT3, dfo.SQRT3, (3, ns + 2 * nf, J + 2 * nf, K + 2 * nf))
        lund = ("1d", U, uu, vv, ww, uw)
        cpu = {"host_cores": multiprocessing.cpu_count()}
        A = None
        for name, loops in (("faithful", True), ("vectorised", False)):
            t0 = time.perf_counter()
            A = dfo.generate(X, ns, J, K, ln, ln, ln, nf, nf, nf, lund, rotate=(1.0, 0.0, 0.0), loops=loops)
            cpu["df_%s_s" % name] = time.perf_counter() - t0
        for label, limit in (("all", None), ("1", 1)):
            ctx = threadpool_limits(limits=limit) if (threadpool_limits and limit) else None
            if ctx is not None:
                ctx.__enter__()
            try:
                t0 = time.perf_counter()
                mean, A0 = dfo.remove_mean(np.array(A))
                r = podo.pod(A0, ns, J * K, 3, "false", [], 1.0e-15, nm)
                t1 = time.perf_counter()
                podo.fourier_coefficients(r["temporal_modes"], ns, r["nm"], dt, c["et"], direct=True)
                t2 = time.perf_counter()
            finally:
                if ctx is not None:
                    ctx.__exit__(None, None, None)
            cpu["pod_blas_threads_%s_s" % label] = {"mean_gram_eig_project": t1 - t0, "fourier_python_loop": t2 - t1}
        for name in ("faithful", "vectorised"):
            tot = cpu["df_%s_s" % name] + sum(cpu["pod_blas_threads_all_s"].values())
            cpu["%s_total_s" % name] = tot
            cpu["%s_planes_per_s" % name] = ns / tot
        cpu["kind"] = "port"
        cpu["note"] = ("oracle port of DF:1403-1510 (scipy direct convolutions; faithful = the reference's per-point Python "
                       "loops, vectorised = same arithmetic in numpy), POD via numpy BLAS/LAPACK, Fourier fit = the reference's "
                       "O(ns^2) Python DFT loop; the DF loop is single-threaded either way, BLAS threads only touch the POD part")
        out["cpu"] = cpu
    return out


def run_multi_gpu(args, pkg, torch, dist, dev, local_rank, rank, world, barrier, max_over_ranks):
    """Driver-visible multi-GPU evidence (N > 1): (a) parity on a small case -- generated time shards equal the
    single-GPU run bit for bit, the injected-mode halo received over NCCL equals the neighbour's slabs bit for bit,
    the row-sharded Gram matrix after the all-reduce equals the one-GPU Gram matrix; (b) the two collectives of the
    path timed at config-3 size: the filter-width halo exchange and the DF -> PODFS all-to-all re-shard."""
    capi, df, pipeline = pkg._capi, pkg.digitalfilters, pkg.pipeline
    out = {}
    # ---- (a) parity, 64 x 48 plane, nf = 4, 16 steps per rank
    J, K, nf, ln = 64, 48, 4, 2.0
    ns = 16 * world
    taps = df._taps(nf, ln)
    plan = capi.DfPlan(J, K, nf, nf, nf, taps, taps, taps, device=local_rank)
    U, uu, vv, ww, uw = df.build_profile("hyperbolic-tangent", "top-hat", 1.0, 0.05, K)
    plan.set_coefficients(*pkg._lund.pack(*pkg._lund.lund_1d(U, uu, vv, ww, uw, J, K), J, K))
    st = torch.cuda.current_stream(dev).cuda_stream
    seed = 99
    t0, t1 = pipeline.shard_range(ns, rank, world, 8)
    whole = torch.empty((ns, 3, J, K), dtype=torch.float64, device=dev)
    plan.generate_seeded_dev(seed, 0, ns, whole.data_ptr(), st)                       # the one-GPU run, on every rank
    mine = torch.empty((t1 - t0, 3, J, K), dtype=torch.float64, device=dev)
    plan.generate_seeded_dev(seed, t0, t1 - t0, mine.data_ptr(), st)
    torch.cuda.synchronize(dev)
    shard_ok = bool(torch.equal(mine, whole[t0:t1]))
    Jp, Kp = J + 2 * nf, K + 2 * nf
    nown = (t1 - t0) if rank < world - 1 else (t1 - t0) + 2 * nf                       # the last rank owns the tail slabs
    own = torch.empty((nown, 3, Jp, Kp), dtype=torch.float64, device=dev)
    plan.random_slabs_dev(seed, t0, nown, own.data_ptr(), st)
    torch.cuda.synchronize(dev)
    ext = pipeline.exchange_halo(own, 2 * nf, rank, world)
    want = torch.empty((t1 - t0 + 2 * nf, 3, Jp, Kp), dtype=torch.float64, device=dev)
    plan.random_slabs_dev(seed, t0, t1 - t0 + 2 * nf, want.data_ptr(), st)
    torch.cuda.synchronize(dev)
    halo_ok = bool(torch.equal(ext, want))
    inj = torch.empty((t1 - t0, 3, J, K), dtype=torch.float64, device=dev)
    plan.generate_injected_dev(ext.data_ptr(), ext.shape[0], inj.data_ptr(), st)
    torch.cuda.synchronize(dev)
    inj_err = float((inj - whole[t0:t1]).abs().max() / whole.abs().max())
    ovl = torch.empty((t1 - t0, 3, J, K), dtype=torch.float64, device=dev)       # halo hidden behind the interior steps
    pipeline.generate_injected_sharded(plan, own, ovl, rank, world)
    torch.cuda.synchronize(dev)
    ovl_ok = bool(torch.equal(ovl, inj))
    # Gram: rows sharded, one all-reduce
    M = 3 * J * K
    a_full = whole.view(ns, M)
    L = capi.lib()
    C1 = torch.empty((ns, ns), dtype=torch.float64, device=dev)
    capi.check(L.podsdf_gram_dev(local_rank, a_full.data_ptr(), M, ns, 1, None, 0, 1.0 / ns, C1.data_ptr(), st))
    m0, m1 = pipeline.shard_range(M, rank, world)
    a_rows = a_full[:, m0:m1].contiguous()
    Cp = torch.empty((ns, ns), dtype=torch.float64, device=dev)
    capi.check(L.podsdf_gram_dev(local_rank, a_rows.data_ptr(), m1 - m0, ns, 1, None, 0, 1.0 / ns, Cp.data_ptr(), st))
    pipeline.allreduce_gram(Cp)
    torch.cuda.synchronize(dev)
    gram_err = float((Cp - C1).abs().max() / C1.abs().max())
    # re-shard parity: all-to-all of the time shards gives every rank its rows of the one-GPU matrix
    a_loc, rows = pipeline.reshard_time_to_rows(mine.view(t1 - t0, M))
    torch.cuda.synchronize(dev)
    resh_ok = bool(torch.equal(a_loc, a_full[:, rows[0]:rows[1]]))
    flags = torch.tensor([float(shard_ok), float(halo_ok), float(resh_ok), -inj_err, -gram_err, float(ovl_ok)],
                         dtype=torch.float64, device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    out["parity"] = {"case": "64x48x3, nf=4, %d steps, %d ranks" % (ns, world),
                     "generated_shards_bitexact": bool(flags[0] > 0.5), "halo_exchange_bitexact": bool(flags[1] > 0.5),
                     "reshard_all_to_all_bitexact": bool(flags[2] > 0.5),
                     "overlapped_halo_equals_blocking_bitexact": bool(flags[5] > 0.5),
                     "injected_vs_one_gpu_max_rel": float(-flags[3]), "gram_allreduce_vs_one_gpu_max_rel": float(-flags[4])}
    plan.close()
    del whole, mine, own, ext, want, inj, ovl, C1, Cp, a_rows, a_loc
    torch.cuda.empty_cache()
    # ---- (b) collectives at config-3 size (512^2 x 3, nf = 10, 2500 planes per GPU)
    c = CFG3
    J, K, nf = c["J"], c["K"], c["nf"]
    Jp, Kp = J + 2 * nf, K + 2 * nf
    halo = torch.randn((2 * nf + 1, 3, Jp, Kp), dtype=torch.float64, device=dev)     # >= 2 nfx slabs: what a neighbour needs
    for it in range(3):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ext = pipeline.exchange_halo(halo, 2 * nf, rank, world)
        e1.record()
        barrier()
        halo_ms = max_over_ranks(e0.elapsed_time(e1))
        del ext
    hb = 2 * nf * 3 * Jp * Kp * 8
    out["halo_exchange"] = {"ms": halo_ms, "bytes_per_edge": hb, "gbs_per_edge": hb / (halo_ms * 1e-3) / 1e9,
                            "what": "2 nfx = 20 raw random slabs of 3 x 532 x 532 doubles from the right neighbour "
                                    "(NCCL send/recv + concatenation), injected mode only"}
    del halo
    n_loc = min(args.df_planes, 2500)
    M = 3 * J * K
    planes = torch.empty((n_loc, M), dtype=torch.float64, device=dev)
    planes.normal_()
    for it in range(2):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a_loc, rows = pipeline.reshard_time_to_rows(planes)
        e1.record()
        barrier()
        resh_ms = max_over_ranks(e0.elapsed_time(e1))
        del a_loc
    nbytes = n_loc * M * 8
    out["reshard_all_to_all"] = {"ms": resh_ms, "bytes_per_gpu": nbytes, "gbs_per_gpu": nbytes / (resh_ms * 1e-3) / 1e9,
                                 "sent_off_gpu_gbs": nbytes * (world - 1) / world / (resh_ms * 1e-3) / 1e9,
                                 "what": "time-sharded planes [t_local][row] -> row shards [t][rows_local]: pack, one NCCL "
                                         "all-to-all, concatenate (pipeline.reshard_time_to_rows), %d planes per GPU" % n_loc}
    del planes
    torch.cuda.empty_cache()
    return out

# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8,
                    help="timed steps; the default 8 shards of 2500 planes are config 3's whole 20000-step job")
    ap.add_argument("--sustained-steps", type=int, default=40, help="extra back-to-back steps for the sustained figure (0 = off)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--df-planes", type=int, default=CFG3["planes_per_gpu"])
    ap.add_argument("--nf", type=int, default=CFG3["nf"], help="half filter width (config 5 sweeps 4..64)")
    ap.add_argument("--rng", default="philox", choices=["philox", "injected"],
                    help="philox: random field generated on the device inside the timed region (headline); "
                         "injected: device-resident random slabs read from HBM")
    ap.add_argument("--e2e-planes", type=int, default=512)
    ap.add_argument("--pod-rows", type=int, default=CFG4["rows_per_gpu"])
    ap.add_argument("--pod-snapshots", type=int, default=CFG4["snapshots"])
    ap.add_argument("--pod-modes", type=int, default=CFG4["modes"])
    ap.add_argument("--pod-steps", type=int, default=2)
    ap.add_argument("--no-pod", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cfg1", action="store_true", help="skip the config-1 leg (64x64x3, 1000 steps, 20 modes, in full)")
    ap.add_argument("--no-multi", action="store_true", help="skip the multi-GPU parity / collective timing leg (N > 1)")
    ap.add_argument("--impl-kernels", type=int, default=0, help="0 auto, 1 generic, 2 tuned")
    args = ap.parse_args()

    _divert_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import podsdf_loader
    pkg = podsdf_loader.load()
    capi, df = pkg._capi, pkg.digitalfilters
    capi.require_device()                       # no GPU => loud failure, never a CPU fallback
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    launches0 = capi.launch_count()
    c = CFG3
    J, K, nf = c["J"], c["K"], args.nf
    ln = c["ln"] * nf / c["nf"]
    n_planes = args.df_planes
    taps = df._taps(nf, ln)
    plan = capi.DfPlan(J, K, nf, nf, nf, taps, taps, taps, device=local_rank)
    plan.set_impl(args.impl_kernels)
    plan.set_chunk(DF_CHUNK)
    n_passes = -(-n_planes // plan.info()["chunk"])          # K-time + K-plane launches per step
    planes_per_launch = n_planes / float(n_passes)
    coef, mask = pkg._lund.pack(*cfg3_lund(pkg), J, K)      # default normal (1,0,0): the rotation is the identity
    plan.set_coefficients(coef, mask)
    shard = pkg.pipeline.DfShard(plan, n_planes, dev)
    seed, t0_rank = 20261018, rank * n_planes
    stream = torch.cuda.current_stream(dev)
    slabs = None
    if args.rng == "injected":
        slabs = torch.empty((n_planes + 2 * nf, 3, J + 2 * nf, K + 2 * nf), dtype=torch.float64, device=dev)
        plan.random_slabs_dev(seed, t0_rank, n_planes + 2 * nf, slabs.data_ptr(), stream.cuda_stream)

    def df_step():
        if slabs is None:
            shard.generate_seeded(seed, t0_rank, stream)
        else:
            shard.generate_injected(slabs, stream)

    # the clock sampler starts BEFORE the warm-up, so that it is certainly polling when the timed region begins
    # (started after it, the first NVML sample sometimes arrived only after the ~0.1 s region had ended)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler is not None and sampler.mode == "nvidia-smi":
        time.sleep(0.3)                                  # the CLI sampler needs time to start; NVML does not
    for _ in range(args.warmup):
        df_step()
    barrier()
    if sampler is not None and sampler.mode == "nvml" and not sampler.rows:
        t_wait = time.time()                             # (rare) the poller has not delivered yet: wait, then re-warm
        while not sampler.rows and time.time() - t_wait < 0.5:
            time.sleep(0.002)
        for _ in range(2):
            df_step()
        barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    launches_before = capi.launch_count()
    e0.record(stream)
    for _ in range(args.steps):
        df_step()
    e1.record(stream)
    launches_timed = capi.launch_count() - launches_before
    barrier()
    w1 = time.time()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    t_time_ms, t_plane_ms = plan.last_kernel_ms()          # kernels of the last step, events on `stream`
    clocks = sampler.stop(w0, w1) if sampler else None
    ms_per_step = ms_total / args.steps
    value = world * n_planes * args.steps / (ms_total * 1e-3)

    # ---- sustained leg: the same step back to back for ~0.6 s, long enough for the board's power controller to
    # settle (FP64-heavy kernels run into sw_power_cap on this pool); reported next to `value`, never instead of it
    sustained = None
    if args.sustained_steps > 0:
        barrier()
        sampler2 = ClockSampler(local_rank) if rank == 0 else None
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        v0 = time.time()
        s0.record(stream)
        for _ in range(args.sustained_steps):
            df_step()
        s1.record(stream)
        barrier()
        v1 = time.time()
        ms_sus = max_over_ranks(s0.elapsed_time(s1))
        tt_s, tp_s = plan.last_kernel_ms()
        sustained = {"value": world * n_planes * args.sustained_steps / (ms_sus * 1e-3), "unit": "planes/s",
                     "steps": args.sustained_steps, "ms_per_step": ms_sus / args.sustained_steps,
                     "time_filter_ms": tt_s, "plane_filter_ms": tp_s,
                     "clocks": sampler2.stop(v0, v1) if sampler2 else None}

    info = plan.info()
    assert info["passes"] == n_passes, (info, n_passes)
    # ---- the same shard with a rotated plane normal: the rotation couples the components (DF:1119-1131), so the
    # three-component plane kernel runs instead of the per-component one (short leg: 1 warm-up + 2 timed steps)
    rotated = None
    if args.rng == "philox":
        plan.set_rotation(df.prof_rotation_matrix(0.6, -0.48, 0.64))
        df_step()
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record(stream)
        for _ in range(2):
            df_step()
        r1.record(stream)
        barrier()
        ms_rot = max_over_ranks(r0.elapsed_time(r1)) / 2
        tt_r, tp_r = plan.last_kernel_ms()
        rotated = {"value": world * n_planes / (ms_rot * 1e-3), "unit": "planes/s", "ms_per_step": ms_rot, "normal": [0.6, -0.48, 0.64],
                   "time_filter_ms": tt_r, "plane_filter_ms": tp_r,
                   "plane_kernel": "df_plane_panel_kernel (three components per CTA, rotation applied after the Lund step)",
                   "plane_kernel_frac_of_hbm": 24.0 * ((J + 2 * nf) * (K + 2 * nf) + J * K) * n_planes / (tp_r * 1e-3) / 1e9 / measured_peaks()[0]}
        plan.set_rotation(None)
    # ---- roofline of the dominant kernel (K-plane: yz filter + Lund, one launch per pass of <= DF_CHUNK planes)
    b_plane, b_time, fl_total, fl_plane = df_algorithmic(J, K, nf, nf, nf)
    if args.rng == "injected":
        b_time = 24.0 * (J + 2 * nf) * (K + 2 * nf) * 2
    hbm_peak, hbm_src = measured_peaks()
    achieved = b_plane * n_planes / (t_plane_ms * 1e-3) / 1e9
    roofline = {
        "kernel": ("df_plane_panel1_kernel<%d,%d> (z-pass + y-pass + Lund + mean fused, one CTA per strip/plane/component)"
                   % (2 * nf + 1, 2 * nf + 1) if info["diagonal"] else
                   "df_plane_panel_kernel<%d,%d> (z-pass + y-pass + coupled Lund/rotation + mean fused)" % (2 * nf + 1, 2 * nf + 1))
        if info["tuned_plane"] else "df generic plane kernels",
        "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
        "peak_source": hbm_src, "traffic": kplane_traffic(nf, planes_per_launch),
        "algorithmic_bytes_per_plane": b_plane, "planes_per_launch": planes_per_launch,
        "algorithmic_bytes_per_launch": b_plane * planes_per_launch,
        "launches_per_step": n_passes, "avg_launch_ms": t_plane_ms / n_passes,
        "kernel_ms_per_step": t_plane_ms,
        "fp64": {"achieved_tflops": fl_plane * n_planes / (t_plane_ms * 1e-3) / 1e12, "peak_tflops": FP64_DFMA_PEAK_TFLOPS,
                 "frac": fl_plane * n_planes / (t_plane_ms * 1e-3) / 1e12 / FP64_DFMA_PEAK_TFLOPS,
                 "note": "at nf=10 the kernel is balanced between HBM (%.2f us/plane) and the DFMA pipe (%.2f us/plane)"
                         % (b_plane / hbm_peak / 1e3, fl_plane / FP64_DFMA_PEAK_TFLOPS / 1e6)},
        "time_filter_kernel": {"ms_per_step": t_time_ms, "achieved_gbs": b_time * n_planes / (t_time_ms * 1e-3) / 1e9,
                               "algorithmic_bytes_per_plane": b_time},
    }

    # ---- whole step against SURVEY 8(d): max(algorithmic bytes / HBM peak, flops / FP64 peak) / measured time
    us_plane = 1e3 * ms_per_step / n_planes
    b_step = 24.0 * J * K if args.rng == "philox" else b_plane
    step_roofline = {
        "us_per_plane": us_per_plane_fmt(us_plane),
        "algorithmic_bytes_per_plane": b_step, "algorithmic_flops_per_plane": fl_total,
        "hbm_floor_us": b_step / hbm_peak / 1e3, "fp64_floor_us": fl_total / FP64_DFMA_PEAK_TFLOPS / 1e6,
        "frac": max(b_step / hbm_peak / 1e3, fl_total / FP64_DFMA_PEAK_TFLOPS / 1e6) / us_plane,
        "bound": "fp64" if fl_total / FP64_DFMA_PEAK_TFLOPS / 1e6 > b_step / hbm_peak / 1e3 else "hbm",
        "note": "generated mode writes each plane once (24 J K bytes); the time-filtered intermediate T1 makes one round "
                "trip through HBM (6.95 MB/plane written + read at nf=10) that is not algorithmic -- DESIGN.md section 3 "
                "shows why it cannot stay on chip",
    }

    # ---- e2e: public host API (pinned host output, D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        n_e = min(args.e2e_planes, n_planes)
        out_h, owner = capi.pinned_empty((n_e, 3, J, K))
        plan.generate_seeded(seed, t0_rank, n_e, out=out_h)                    # warm-up (allocates staging)
        barrier()
        t0 = time.perf_counter()
        reps = max(2, min(args.steps, 5))
        for _ in range(reps):
            plan.set_coefficients(coef, mask)                                  # H2D of the step's inputs
            plan.generate_seeded(seed, t0_rank, n_e, out=out_h)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        present = bin(mask).count("1")
        e2e = {"value": world * n_e * reps / dt, "unit": "planes/s",
               "h2d_bytes_per_step": present * J * K * 8, "d2h_bytes_per_step": n_e * 3 * J * K * 8,
               "planes_per_step": n_e, "steps": reps,
               "api": "DfPlan.set_coefficients + DfPlan.generate_seeded -> podsdf_df_generate_seeded_host (pinned host output)",
               "timer": "host wall clock around the synchronous C-ABI call, barrier both sides, max over ranks"}
        owner.free()
    del shard, slabs
    torch.cuda.empty_cache()

    # ---- PODFS on config 4's per-GPU row shard
    pod = None
    if not args.no_pod:
        pod = run_pod(args, pkg, torch, dist, dev, local_rank, rank, world, barrier, max_over_ranks)
    launches_all = capi.launch_count() - launches0

    multi = None
    if world > 1 and not args.no_multi:
        multi = run_multi_gpu(args, pkg, torch, dist, dev, local_rank, rank, world, barrier, max_over_ranks)
    cfg1 = None
    if rank == 0 and world == 1 and not args.no_cfg1:
        cfg1 = run_cfg1(args, pkg, torch, dev, local_rank, cpu_too=not args.no_cpu)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, dt, tm = cpu_df_baseline(2, loops=True)
        vv, dtv, tmv = cpu_df_baseline(2, loops=False)
        import multiprocessing
        cpu = {"value": v, "unit": "planes/s", "cores": 1, "kind": "port",
               "sample": "2 planes of config 3 through the oracle port of the reference loop (scipy direct convolutions + "
                         "per-point Python Lund/rotate loops, single-threaded like the reference), %.1f s" % dt,
               "parts_s": tm, "vectorised_numpy_value": vv, "host_cores": multiprocessing.cpu_count()}

    if rank == 0:
        line = {
            "metric": "df_planes_per_s", "value": value, "unit": "planes/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "cfg3 round-jet inlet 512x512x3, nf=%d, ln=%g, adapt2d Lund transform (default plane normal: the "
                                   "rotation is the identity and is skipped, DF:1476-1477): %d-plane time shard per GPU "
                                   "(config 3's 20000 steps at 8 GPUs)" % (nf, ln, n_planes),
                       "rng": args.rng, "planes_per_step_per_gpu": n_planes, "parallelism": "time-sharded x%d" % world,
                       "l2": "every step writes %.1f GB per GPU (>> 126 MB L2); no flush needed" % (n_planes * 3 * J * K * 8 / 1e9),
                       "kernels": "tuned" if (info["tuned_time"] and info["tuned_plane"]) else "generic",
                       "planes_per_pass": info["chunk"]},
            "roofline": roofline, "step_roofline": step_roofline, "cpu_baseline": cpu, "e2e": e2e, "cfg1": cfg1,
            "multi_gpu": multi, "rotated_normal": rotated, "gpu_launches": launches_timed,
            "gpu_launches_total": launches_all, "clocks": clocks, "sustained": sustained, "pod": pod,
        }
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_pod(args, pkg, torch, dist, dev, local_rank, rank, world, barrier, max_over_ranks):
    capi = pkg._capi
    M, N, nm = args.pod_rows, args.pod_snapshots, args.pod_modes
    # synthetic snapshots: low-rank structure with a separated spectrum + noise,
    # built on the device in row chunks (data generation, untimed)
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    A = torch.empty((M, N), dtype=torch.float64, device=dev)
    R = 200
    t = torch.arange(N, dtype=torch.float64, device=dev) / N
    r = torch.arange(1, R + 1, dtype=torch.float64, device=dev)
    G = (torch.cos(2 * np.pi * r[:, None] * t[None, :] + 0.7 * r[:, None])
         + 0.5 * torch.sin(2 * np.pi * (r[:, None] + 3) * t[None, :])) * (r ** -0.8)[:, None]
    CH = 32768
    for m0 in range(0, M, CH):
        m1 = min(M, m0 + CH)
        p = (torch.arange(rank * M + m0, rank * M + m1, dtype=torch.float64, device=dev) + 0.5) / (M * world)
        phi = torch.sin(np.pi * r[None, :] * p[:, None] + 0.3 * r[None, :])
        A[m0:m1] = phi @ G
        A[m0:m1] += 1e-3 * (torch.rand((m1 - m0, N), dtype=torch.float64, device=dev, generator=g) - 0.5)
        A[m0:m1] += 1.0 + p[:, None]              # a mean field to remove
    del G
    group = None
    timers = {}
    res = pkg.pipeline.pod_resident(A, M, N, 0, nm, CFG4["dt"], CFG4["et"], local_rank, group, None)   # warm-up
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = max(1, args.pod_steps)
    e0.record()
    gram_ms = 0.0
    step_ms = []                                   # per-step device time (the host parts -- ranking of the Fourier
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]   # coefficients -- can hiccup)
    marks[0].record()
    for i in range(steps):
        res = pkg.pipeline.pod_resident(A, M, N, 0, nm, CFG4["dt"], CFG4["et"], local_rank, group, timers)
        g_ms = np.zeros(1)
        capi.check(capi.lib().podsdf_gram_last_ms(capi.ptr(g_ms)))
        gram_ms += float(g_ms[0])
        marks[i + 1].record()
    e1.record()
    torch.cuda.synchronize(dev)
    step_ms = [marks[i].elapsed_time(marks[i + 1]) for i in range(steps)]
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    gram_ms /= steps
    flops = float(M) * N * (N + 1)
    tf = flops / (gram_ms * 1e-3) / 1e12
    phi = res["Phi"]
    orth = float((phi.T @ phi - torch.eye(res["nm"], dtype=torch.float64, device=dev)).abs().max()) if world == 1 else None
    return {
        "metric": "pod_snapshots_per_s", "value": N / (ms * 1e-3), "unit": "snapshots/s", "ms_per_step": ms, "steps": steps,
        "ms_each_step": step_ms,
        "config": {"workload": "cfg4 row shard: %d rows x %d snapshots per GPU (x%d GPUs), %d modes, et=0.9" % (M, N, world, nm)},
        "parts_ms": {k: (v / steps if not isinstance(v, list) else v) for k, v in timers.items()},
        "roofline": {"kernel": "gram_dmma_kernel (FP64 mma.sync m8n8k4, upper-triangular tiles)", "bound": "tensor",
                     "achieved": tf, "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": tf / FP64_DMMA_PEAK_TFLOPS,
                     "peak_source": "measured FP64 DMMA peak on this pool (tools/probe_fp64.cu); MEASURED_PEAKS.json has no FP64 entry",
                     "frac_of_nominal_40": tf / 40.0,
                     "algorithmic_flops": flops, "kernel_ms": gram_ms, "traffic": gram_traffic(M, N),
                     "algorithmic_bytes": 8.0 * M * N,
                     "traffic_note": "DRAM re-reads of the operand panels (L2 catches half) at 12 % of the DRAM peak: the kernel is "
                                     "bound by the FP64 tensor pipe (86.7 % active in the capture), not by memory"},
        "modes_kept": int(res["nm"]), "orthonormality_max_abs": orth,
        "fourier_coefficients_kept": int(res["c_count"].sum()),
    }


if __name__ == "__main__":
    main()
