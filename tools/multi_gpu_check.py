"""Multi-GPU parity check (run under torchrun on a box with >= 2 B200s):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/multi_gpu_check.py

1. DF, injected mode: the global random stream is split into per-rank slab shards, every rank fetches its
   2*nfx-slab halo from the right neighbour over NCCL (pipeline.exchange_halo), generates its time shard, and
   the concatenation must equal the single-GPU run of the whole stream BIT FOR BIT.
2. DF, generated mode: time shards regenerate their halo from the Philox stream (no collective).
3. PODFS: rows of a snapshot matrix are sharded, partial Grams are all-reduced (NCCL), the replicated
   eigen-solve / projection / Fourier fit must match the single-GPU result.
4. DF -> PODFS: time-sharded planes are re-sharded by rows with one all-to-all and fed to the row-sharded PODFS.
Rank 0 prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import podsdf_loader  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pkg = podsdf_loader.load()
    capi, df, pl = pkg._capi, pkg.digitalfilters, pkg.pipeline
    out = {"world": world}

    # ---- 1 + 2: DF time shards
    J, K, nf, ln, ns = 96, 128, 10, 5.0, 40 * world + 3
    taps = df._taps(nf, ln)
    plan = capi.DfPlan(J, K, nf, nf, nf, taps, taps, taps, device=local)
    U, uu, vv, ww, uw = df.build_profile("hyperbolic-tangent", "top-hat", 1.0, 0.05, K)
    plan.set_coefficients(*pkg._lund.pack(*pkg._lund.lund_1d(U, uu, vv, ww, uw, J, K), J, K, None))
    seed = 77
    t0, t1 = pl.shard_range(ns, rank, world, 8)
    own_hi = t1 if rank < world - 1 else ns + 2 * nf
    mine = torch.empty((own_hi - t0, 3, J + 2 * nf, K + 2 * nf), dtype=torch.float64, device=dev)
    plan.random_slabs_dev(seed, t0, own_hi - t0, mine.data_ptr(), torch.cuda.current_stream().cuda_stream)
    win = pl.exchange_halo(mine, 2 * nf)                       # NCCL send/recv of the filter-width halo
    shard = pl.DfShard(plan, t1 - t0, dev)
    part_inj = shard.generate_injected(win).clone()
    part_gen = shard.generate_seeded(seed, t0).clone()
    torch.cuda.synchronize()
    gathered = [torch.empty((pl.shard_range(ns, r, world, 8)[1] - pl.shard_range(ns, r, world, 8)[0], 3, J, K),
                            dtype=torch.float64, device=dev) for r in range(world)]
    dist.all_gather(gathered, part_inj)
    whole_inj = torch.cat(gathered, 0)
    dist.all_gather(gathered, part_gen)
    whole_gen = torch.cat(gathered, 0)
    single = pl.DfShard(plan, ns, dev).generate_seeded(seed, 0)
    all_slabs = torch.empty((ns + 2 * nf, 3, J + 2 * nf, K + 2 * nf), dtype=torch.float64, device=dev)
    plan.random_slabs_dev(seed, 0, ns + 2 * nf, all_slabs.data_ptr(), torch.cuda.current_stream().cuda_stream)
    single_inj = pl.DfShard(plan, ns, dev).generate_injected(all_slabs)
    torch.cuda.synchronize()
    out["df_injected_halo_bitexact"] = bool(torch.equal(whole_inj, single_inj))
    out["df_injected_vs_generated_max_abs"] = float((whole_inj - single).abs().max())
    out["df_generated_shards_bitexact"] = bool(torch.equal(whole_gen, single))

    # ---- 3: PODFS row shards (the DF planes above are the snapshots: layout [t][row])
    M, N, nm = 3 * J * K, ns, 8
    A_full = single.view(N, M).t().contiguous()                # (M, N) reference layout
    m0, m1 = pl.shard_range(M, rank, world)
    ref = pl.pod_resident(A_full.clone(), M, N, 0, nm, 0.01, 0.9, local, group=dist.new_group([rank]))
    res = pl.pod_resident(A_full[m0:m1].clone(), m1 - m0, N, 0, nm, 0.01, 0.9, local)
    sign = np.sign(np.sum(res["T"] * ref["T"], axis=0))
    out["pod_energy_rel_err"] = float(np.abs(res["energy"][:nm] - ref["energy"][:nm]).max() / ref["energy"][0])
    out["pod_temporal_modes_rel_err"] = float(np.abs(res["T"] * sign - ref["T"]).max() / np.abs(ref["T"]).max())
    phi_ref = ref["Phi"][m0:m1].cpu().numpy()
    out["pod_spatial_modes_rel_err"] = float(np.abs(res["Phi"].cpu().numpy() * sign - phi_ref).max() / np.abs(phi_ref).max())
    out["pod_fourier_counts_equal"] = bool(np.array_equal(res["c_count"], ref["c_count"]))
    # ---- 4: DF -> PODFS across ranks: time shards, one all-to-all re-shard by rows, row-sharded PODFS
    sh = pl.df_to_podfs_sharded(plan, ns, nm, 0.01, 0.9, seed, local)
    r0, r1 = sh["rows"]
    out["reshard_max_abs_err"] = float((sh["planes_rows"] + sh["mean"][None, :] - single.view(N, M)[:, r0:r1]).abs().max())
    sign2 = np.sign(np.sum(sh["T"] * ref["T"], axis=0))
    out["pipeline_temporal_modes_rel_err"] = float(np.abs(sh["T"] * sign2 - ref["T"]).max() / np.abs(ref["T"]).max())
    phi_ref2 = ref["Phi"][r0:r1].cpu().numpy()
    out["pipeline_spatial_modes_rel_err"] = float(np.abs(sh["Phi"].cpu().numpy() * sign2 - phi_ref2).max() / np.abs(phi_ref2).max())
    ok = (out["df_injected_halo_bitexact"] and out["reshard_max_abs_err"] < 1e-12 and out["pipeline_temporal_modes_rel_err"] < 1e-8
          and out["pipeline_spatial_modes_rel_err"] < 1e-8 and out["df_generated_shards_bitexact"] and out["pod_energy_rel_err"] < 1e-12
          and out["pod_temporal_modes_rel_err"] < 1e-8 and out["pod_spatial_modes_rel_err"] < 1e-8
          and out["pod_fourier_counts_equal"])
    flags = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    out["ok_on_all_ranks"] = bool(flags.item() == 1.0)
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()
    sys.exit(0 if out["ok_on_all_ranks"] else 1)


if __name__ == "__main__":
    main()
