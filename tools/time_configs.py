"""Device-resident DF timing of BASELINE configs 1-3 (and injected vs generated RNG) on one GPU.
Prints one JSON line per case; used for profiles/ only (bench.py is the contract benchmark)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import podsdf_loader  # noqa: E402
from oracle import df_oracle as dfo  # noqa: E402  (profile definitions only)


def run(pkg, name, J, K, nf3, ln3, fac_mean, rot, nplanes, rng):
    capi, df = pkg._capi, pkg.digitalfilters
    plan = capi.DfPlan(J, K, nf3[0], nf3[1], nf3[2], df._taps(nf3[0], ln3[0]), df._taps(nf3[1], ln3[1]), df._taps(nf3[2], ln3[2]))
    coef, mask = pkg._lund.pack(fac_mean[0], fac_mean[1], J, K)
    plan.set_coefficients(coef, mask)
    plan.set_rotation(rot)
    dev = torch.device("cuda", 0)
    shard = pkg.pipeline.DfShard(plan, nplanes, dev)
    slabs = None
    if rng == "injected":
        slabs = torch.empty((nplanes + 2 * nf3[0], 3, J + 2 * nf3[1], K + 2 * nf3[2]), dtype=torch.float64, device=dev)
        plan.random_slabs_dev(1, 0, slabs.shape[0], slabs.data_ptr(), torch.cuda.current_stream().cuda_stream)
    step = (lambda: shard.generate_injected(slabs)) if slabs is not None else (lambda: shard.generate_seeded(1, 0))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    tt, tp = plan.last_kernel_ms()
    Jp, Kp = J + 2 * nf3[1], K + 2 * nf3[2]
    b_plane = 24.0 * (Jp * Kp + J * K)
    print(json.dumps({"case": name, "rng": rng, "J": J, "K": K, "nf": nf3, "planes": nplanes, "mask": mask,
                      "planes_per_s": nplanes / (ms * 1e-3), "ms": ms, "k_time_ms": tt, "k_plane_ms": tp,
                      "k_plane_hbm_frac": b_plane * nplanes / (tp * 1e-3) / 1e9 / 6447.2}))
    plan.close()


def main():
    pkg = podsdf_loader.load()
    lund, df = pkg._lund, pkg.digitalfilters
    # config 1: 64x64, nf=10, hyperbolic-tangent profile (adapt1d), default normal -> diagonal transform
    U, uu, vv, ww, uw = dfo.build_profile("hyperbolic-tangent", "top-hat", 1.0, 0.02, 64)
    run(pkg, "cfg1 64x64 nf=10 adapt1d", 64, 64, (10, 10, 10), (5.0, 5.0, 5.0), lund.lund_1d(U, uu, vv, ww, uw, 64, 64), None, 1000, "philox")
    # config 2: 256x128 channel, anisotropic (16,8,4), shear stress -> coupled transform (PAT 1)
    prof = dfo.channel_profile(128)
    for rng in ("philox", "injected"):
        run(pkg, "cfg2 256x128 nf=(16,8,4) adapt1d+uw", 256, 128, (16, 8, 4), (8.0, 4.0, 2.0), lund.lund_1d(*prof, 256, 128), None, 10000, rng)
    # config 3: 512x512 round jet, nf=10 (the bench workload), generated and injected; plus a rotated normal (full 3x3)
    U, uu, vv, ww, uw = dfo.build_profile("circular-hyperbolic-tangent", "top-hat", 1.0, 0.02, 512)
    fm = lund.lund_2d(U, uu, vv, ww, uw, 512, 512, "circular-hyperbolic-tangent", 0.5)
    for rng in ("philox", "injected"):
        run(pkg, "cfg3 512x512 nf=10 adapt2d", 512, 512, (10, 10, 10), (5.0, 5.0, 5.0), fm, None, 1024, rng)
    run(pkg, "cfg3 rotated normal (0.6,-0.48,0.64): full 3x3", 512, 512, (10, 10, 10), (5.0, 5.0, 5.0), fm,
        df.prof_rotation_matrix(0.6, -0.48, 0.64), 1024, "philox")


if __name__ == "__main__":
    main()
