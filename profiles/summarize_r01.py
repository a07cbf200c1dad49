"""Turns the artefacts of profiles/capture_r01.sh (gpurun_out/r01_*) into the tracked
summaries under profiles/: per-kernel ncu metrics + stall reasons + source hot spots,
launch lists restricted to libsbx kernels, the DRAM traffic bench.py reports, bench lines."""
import collections, csv, json, shutil, subprocess, sys

SRC = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"


def raw(rep):
  out = subprocess.run(f"ncu -i {rep} --page raw --csv", shell=True, capture_output=True, text=True).stdout
  rows = list(csv.reader(out.splitlines()))
  return rows[0], rows[1], rows[2:]


def summarize(rep, out, title, source_kernel, ntop=25):
  hdr, units, data = raw(rep)
  lines = [title, ""]
  want = ["Kernel Name", "launch__grid_size", "launch__block_size", "gpu__time_duration.sum",
          "dram__bytes_read.sum", "dram__bytes_write.sum",
          "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
          "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_fma.sum",
          "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum"]
  for w in want:
    for i, h in enumerate(hdr):
      if h == w:
        lines.append(f"{w} [{units[i]}]: " + " | ".join(r[i][:48] for r in data))
  lines += ["", "warp stall reasons (cycles per issued instruction, > 0.25):"]
  for i, h in enumerate(hdr):
    if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h and "_not_issued" not in h:
      v = [r[i] for r in data]
      try:
        if max(float(x) for x in v) > 0.25:
          name = h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")
          lines.append("  " + name + ": " + " | ".join(f"{float(x):.2f}" for x in v))
      except ValueError:
        pass
  src = subprocess.run(f"ncu -i {rep} --page source --csv --print-source cuda,sass --kernel-name regex:{source_kernel} "
                       "--launch-count 1", shell=True, capture_output=True, text=True).stdout
  cur, agg = None, collections.defaultdict(lambda: [0, 0, ""])
  for r in csv.reader(src.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
      cur = r[1].split("/")[-1]
    elif len(r) >= 8 and r[0].isdigit():
      try:
        inst, samp = int(r[7]), int(r[4])
      except ValueError:
        continue
      k = (cur, int(r[0]))
      agg[k][0] += inst
      agg[k][1] += samp
      agg[k][2] = r[1].strip()[:90]
  tot = sum(v[0] for v in agg.values()) or 1
  tots = sum(v[1] for v in agg.values()) or 1
  lines += ["", f"source hot spots of {source_kernel} (share of executed instructions / of stall samples):"]
  for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:ntop]:
    lines.append(f"  {v[0] / tot * 100:5.1f}% inst {v[1] / tots * 100:5.1f}% samp  {k[0]}:{k[1]}  {v[2]}")
  open(out, "w").write("\n".join(lines) + "\n")
  return hdr, data


def col(hdr, data, name, row):
  return float(data[row][hdr.index(name)])


h1, d1 = summarize(f"{SRC}/r01_randomized.ncu-rep", "profiles/r01_ncu_full_randomized.txt",
                   "ncu --set full --clock-control none --import-source on, one step of `python bench.py` (32768 randomized "
                   "64x96 buildings, 1 x B200): k_pre<8>, k_resident_step<4>, k_post<8> (profiles/capture_r01.sh)",
                   "k_resident_step")
h2, d2 = summarize(f"{SRC}/r01_office.ncu-rep", "profiles/r01_ncu_full_office.txt",
                   "ncu --set full --clock-control none --import-source on, `python bench.py --workload office "
                   "--envs-per-gpu 512` (744x1004 plan, 512 copies so that the ~40 replays stay short): two k_sweep<4> "
                   "launches and k_zone_reduce<4> (profiles/capture_r01.sh)", "k_sweep")
for name in ("randomized", "office"):
  rows = list(csv.reader(open(f"{SRC}/r01_launches_{name}.csv")))
  out = [[r[0], r[4].replace("sbx::", ""), r[7], r[8], r[-3], r[-2], r[-1]]
         for r in rows if len(r) > 10 and (r[0] == "ID" or "sbx::" in r[4])]
  csv.writer(open(f"profiles/r01_launches_{name}.csv", "w", newline="")).writerows(out)
ir = next(i for i, r in enumerate(d1) if "k_resident_step" in r[h1.index("Kernel Name")])
unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
hdr_u = raw(f"{SRC}/r01_randomized.ncu-rep")[1]
hdr_o = raw(f"{SRC}/r01_office.ncu-rep")[1]
traffic = {
    "randomized": {"kernel": "k_resident_step<4>", "launch_envs": 32768,
                   "dram_bytes_read": col(h1, d1, "dram__bytes_read.sum", ir) * unit[hdr_u[h1.index("dram__bytes_read.sum")]],
                   "dram_bytes_write": col(h1, d1, "dram__bytes_write.sum", ir) * unit[hdr_u[h1.index("dram__bytes_write.sum")]],
                   "source": "profiles/r01_ncu_full_randomized.txt (ncu --set full, 32768 buildings per launch)"},
    "office": {"kernel": "k_sweep<4>", "launch_envs": 512,
               "dram_bytes_read": col(h2, d2, "dram__bytes_read.sum", 1) * unit[hdr_o[h2.index("dram__bytes_read.sum")]],
               "dram_bytes_write": col(h2, d2, "dram__bytes_write.sum", 1) * unit[hdr_o[h2.index("dram__bytes_write.sum")]],
               "source": "profiles/r01_ncu_full_office.txt (ncu --set full, 512 buildings per launch, a sweep >= 2: reads "
                         "T_est and T_prev)"}}
json.dump(traffic, open("profiles/traffic.json", "w"), indent=1)
for f in ("r01_bench.json", "r01_bench_reference.json"):
  shutil.copy(f"{SRC}/{f}", f"profiles/{f}")
print(json.dumps(traffic, indent=1))
