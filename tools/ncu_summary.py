"""Summarise an .ncu-rep (one kernel launch) into a small text file for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/name.txt [--source-top N]"""
import csv, io, re, subprocess, sys, collections

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index("--source-top") + 1]) if "--source-top" in sys.argv else 12
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    lines = []
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        lines.append("kernel: %s" % name)
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append("  %-78s %s %s" % (k, vals[i], units[i]))
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv"]))))
    if len(src) > 2:
        h = src[1]
        ix = {n: i for i, n in enumerate(h)}
        data = src[2:]
        stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        agg = {s[6:]: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
        tot = sum(agg.values()) or 1
        lines.append("warp stall sampling (share of samples): " + ", ".join(
            "%s %.1f%%" % (k, 100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
        mix = collections.Counter()
        for r in data:
            op = re.sub(r"^@!?U?P\w+\s+", "", r[ix["Source"]].strip()).split(" ")[0].split(".")[0]
            try:
                mix[op] += float(r[ix["Instructions Executed"]])
            except ValueError:
                pass
        t = sum(mix.values()) or 1
        lines.append("executed instruction mix: " + ", ".join("%s %.1f%%" % (k, 100 * v / t) for k, v in mix.most_common(12)))
        lines.append("top stall sites (SASS):")
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:top]:
            st = {s[6:]: int(r[ix[s]] or 0) for s in stalls if int(r[ix[s]] or 0) > 0}
            main3 = sorted(st.items(), key=lambda kv: -kv[1])[:2]
            lines.append("  %5.1f%%  %-60s %s" % (100.0 * int(r[ix["# Samples"]]) / tot, r[ix["Source"]].strip()[:60], main3))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
