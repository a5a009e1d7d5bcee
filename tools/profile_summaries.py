"""Writes the round's ncu summaries into profiles/: python tools/profile_summaries.py r2_cfg2 r2 2 32768 65536"""
import collections, csv, io, json, subprocess, sys
tag, rnd, cfg, per_launch, per_step = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
out = {}
dram = 0.0
for k in ("factor", "solve", "certify"):
    rep = "gpurun_out/%s_%s.ncu-rep" % (tag, k)
    a = subprocess.run([sys.executable, "tools/ncu_lines.py", rep, "24"], capture_output=True, text=True).stdout
    b = subprocess.run([sys.executable, "tools/ncu_funcs.py", rep, "qppvm_b200/csrc/qp_kernel.cuh", str(per_launch)], capture_output=True, text=True).stdout
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw))); h = rows[0]
    get = lambda name: float(rows[2][h.index(name)].replace(",", ""))
    unit = lambda name: rows[1][h.index(name)]
    rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    rd *= scale[unit("dram__bytes_read.sum")]; wr *= scale[unit("dram__bytes_write.sum")]
    dram += rd + wr
    extra = ""
    for m in ("gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
              "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__warps_eligible.avg.per_cycle_active"):
        if m in h:
            extra += "%s %s %s\n" % (m, rows[1][h.index(m)], rows[2][h.index(m)])
    open("profiles/%s_cfg%d_qp_%s_kernel_summary.txt" % (rnd, cfg, k), "w").write(
        "# ncu --set full --import-source on --clock-control none -k regex:qp_%s -s 6 -c 1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-latency\n"
        "# one launch = %d problems (one prepare-workspace pass of the %d-record step)\n" % (k, per_launch, per_step) + a + extra + "\n# warp-instructions by source function\n" + b)
    out[k] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "time_us": get("gpu__time_duration.sum") * (1e3 if unit("gpu__time_duration.sum") == "ms" else 1.0)}
json.dump({"config": cfg, "records_per_launch": per_launch, "dram_bytes_per_launch": dram, "kernels": out,
           "what": "dram__bytes_read.sum + dram__bytes_write.sum of one launch each of qp_factor_kernel, qp_solve_kernel, qp_certify_kernel (ncu --set full), i.e. one %d-problem pass" % per_launch},
          open("profiles/traffic_config%d.json" % cfg, "w"), indent=1)
# launch list -> per-kernel share
rows = [r for r in csv.reader(open("gpurun_out/%s_launches.csv" % tag)) if len(r) > 10]
h = rows[0]; ik, iv = h.index("Kernel Name"), h.index("Metric Value")
d = collections.defaultdict(list)
for r in rows[1:]:
    d[r[ik].split("(")[0][:60]].append(float(r[iv].replace(",", "")))
tot = sum(sum(v) for v in d.values())
with open("profiles/%s_cfg%d_launches_summary.txt" % (rnd, cfg), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-latency\n")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        f.write("%-62s launches %4d  avg %9.1f us  share %5.1f %%\n" % (k, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
print(open("profiles/%s_cfg%d_launches_summary.txt" % (rnd, cfg)).read()); print(json.dumps(out))
