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
    ap.add_argument("--pod-steps", type=int, default=1)
    ap.add_argument("--no-pod", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
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

    for _ in range(args.warmup):
        df_step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler is not None and sampler.mode == "nvidia-smi":
        time.sleep(0.3)                                  # the CLI sampler needs time to start; NVML does not
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
            "config": {"workload": "cfg3 round-jet inlet 512x512x3, nf=%d, ln=%g, adapt2d Lund + rotation: %d-plane time shard per GPU "
                                   "(config 3's 20000 steps at 8 GPUs)" % (nf, ln, n_planes),
                       "rng": args.rng, "planes_per_step_per_gpu": n_planes, "parallelism": "time-sharded x%d" % world,
                       "l2": "every step writes %.1f GB per GPU (>> 126 MB L2); no flush needed" % (n_planes * 3 * J * K * 8 / 1e9),
                       "kernels": "tuned" if (info["tuned_time"] and info["tuned_plane"]) else "generic",
                       "planes_per_pass": info["chunk"]},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches_timed,
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
    for _ in range(steps):
        res = pkg.pipeline.pod_resident(A, M, N, 0, nm, CFG4["dt"], CFG4["et"], local_rank, group, timers)
        g_ms = np.zeros(1)
        capi.check(capi.lib().podsdf_gram_last_ms(capi.ptr(g_ms)))
        gram_ms += float(g_ms[0])
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    gram_ms /= steps
    flops = float(M) * N * (N + 1)
    tf = flops / (gram_ms * 1e-3) / 1e12
    phi = res["Phi"]
    orth = float((phi.T @ phi - torch.eye(res["nm"], dtype=torch.float64, device=dev)).abs().max()) if world == 1 else None
    return {
        "metric": "pod_snapshots_per_s", "value": N / (ms * 1e-3), "unit": "snapshots/s", "ms_per_step": ms, "steps": steps,
        "config": {"workload": "cfg4 row shard: %d rows x %d snapshots per GPU (x%d GPUs), %d modes, et=0.9" % (M, N, world, nm)},
        "parts_ms": {k: v / steps for k, v in timers.items()},
        "roofline": {"kernel": "gram_dmma_kernel (FP64 mma.sync m8n8k4, upper-triangular tiles)", "bound": "tensor",
                     "achieved": tf, "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": tf / FP64_DMMA_PEAK_TFLOPS,
                     "peak_source": "measured FP64 DMMA peak on this pool (tools/probe_fp64.cu); MEASURED_PEAKS.json has no FP64 entry",
                     "algorithmic_flops": flops, "kernel_ms": gram_ms, "traffic": None},
        "modes_kept": int(res["nm"]), "orthonormality_max_abs": orth,
        "fourier_coefficients_kept": int(res["c_count"].sum()),
    }


if __name__ == "__main__":
    main()
