# BASELINE config 5: filter-width sweep at 512^2 x 3 (one JSON line per nf into gpurun_out/r02_config5.jsonl)
rm -f gpurun_out/r02_config5.jsonl
for nf in 4 6 8 10 12 16 24 32 48 64; do
  python bench.py --steps 3 --warmup 3 --no-pod --no-cpu --no-e2e --no-cfg1 --sustained-steps 0 --nf $nf --df-planes 512 >> gpurun_out/r02_config5.jsonl 2>> gpurun_out/r02_config5.err
done
python - <<'PY'
import json
print("| nf | taps | planes/s | ms/step | K-plane ms | K-time ms | K-plane frac of HBM | K-plane frac of DFMA peak | whole-step frac |")
print("|---|---|---|---|---|---|---|---|---|")
for l in open("gpurun_out/r02_config5.jsonl"):
    d = json.loads(l)
    r = d["roofline"]
    nf = int(d["config"]["workload"].split("nf=")[1].split(",")[0])
    print("| %d | %d | %.0f | %.3f | %.3f | %.3f | %.3f | %.3f | %.3f |" % (nf, 2 * nf + 1, d["value"], d["ms_per_step"], r["kernel_ms_per_step"],
          r["time_filter_kernel"]["ms_per_step"], r["frac"], r["fp64"]["frac"], d["step_roofline"]["frac"]))
PY
