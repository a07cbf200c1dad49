"""Turns the artefacts of profiles/capture_r02.sh (gpurun_out/r02_*) into the tracked
summaries under profiles/: per-kernel ncu metrics + stall reasons + source hot spots + the
instruction / stall / shared-memory-wavefront share of every barrier-delimited segment, launch
lists restricted to libsbx kernels, the DRAM traffic bench.py reports, the bench lines, and a
SASS mnemonic listing per kernel of the shipped libsbx.so (UTMALDG / UBLKCP / FFMA2 / REDUX ...)."""
import collections, csv, json, os, re, shutil, subprocess, sys

SRC = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def run(cmd):
  return subprocess.run(cmd, shell=True, capture_output=True, text=True).stdout


def raw(rep):
  rows = list(csv.reader(run(f"ncu -i {rep} --page raw --csv").splitlines()))
  return rows[0], rows[1], rows[2:]


WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def segments(rep, kernel, envs):
  """Instructions / stall samples / shared-memory wavefronts per barrier-delimited segment."""
  rows = list(csv.reader(run(f"ncu -i {rep} --page source --csv --print-source sass --kernel-name regex:{kernel} "
                             "--launch-count 1").splitlines()))
  if len(rows) < 3:
    return []
  hdr = rows[1]
  ia, isrc, ie, iss = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
  iw, iwi = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
  data, base = [], None
  for r in rows[2:]:
    if len(r) <= ie:
      continue
    if r[ia] == "Address":      # a second matching launch: keep the first
      break
    a = int(r[ia], 16)
    base = a if base is None else base
    data.append((a - base, r[isrc].strip(), int(r[ie]), int(r[iss]), int(r[iw] or 0), int(r[iwi] or 0)))
  tot = sum(d[2] for d in data) or 1
  tots = sum(d[3] for d in data) or 1
  out = [f"  per building: {tot / envs:.0f} warp instructions, {sum(d[4] for d in data) / envs:.0f} shared-memory wavefronts "
         f"({sum(d[5] for d in data) / envs:.0f} ideal)"]
  cur, start = [0, 0, 0, 0, 0], 0
  for off, src, e, s, w, wi in data:
    cur = [cur[0] + e, cur[1] + s, cur[2] + w, cur[3] + wi, cur[4] + 1]
    if "BAR." in src:
      out.append(f"  {start:#06x}-{off:#06x}  {cur[4]:4d} static  {cur[0] / envs:7.0f} exec/bldg ({cur[0] / tot * 100:4.1f}%)  "
                 f"stall samples {cur[1] / tots * 100:4.1f}%  smem wavefronts/bldg {cur[2] / envs:6.0f} (ideal {cur[3] / envs:6.0f})"
                 f"  ends: {src.split()[0]}")
      cur, start = [0, 0, 0, 0, 0], off + 16
  out.append(f"  {start:#06x}-end      {cur[4]:4d} static  {cur[0] / envs:7.0f} exec/bldg ({cur[0] / tot * 100:4.1f}%)  "
             f"stall samples {cur[1] / tots * 100:4.1f}%")
  return out


def summarize(rep, out, title, source_kernel, envs=None, ntop=20):
  hdr, units, data = raw(rep)
  lines = [title, ""]
  for w in WANT:
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
  src = run(f"ncu -i {rep} --page source --csv --print-source cuda,sass --kernel-name regex:{source_kernel} --launch-count 1")
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
  if envs:
    lines += ["", f"{source_kernel}: segments between barriers (SASS addresses)"] + segments(rep, source_kernel, envs)
  open(out, "w").write("\n".join(lines) + "\n")
  return hdr, units, data


def col(hdr, units, data, name, row):
  i = hdr.index(name)
  return float(data[row][i]) * UNIT.get(units[i], 1.0)


def row_of(hdr, data, kernel):
  return next(i for i, r in enumerate(data) if kernel in r[hdr.index("Kernel Name")])


h1, u1, d1 = summarize(f"{SRC}/r02_randomized.ncu-rep", "profiles/r02_ncu_full_randomized.txt",
                       "ncu --set full --clock-control none --import-source on, one step of `python bench.py` (32768 randomized "
                       "64x96 buildings, 1 x B200): k_pre<4>, k_resident_step<4>, k_post<4> (profiles/capture_r02.sh)",
                       "k_resident_step", envs=32768)
if os.path.exists(f"{SRC}/r02_randomized_v3.ncu-rep"):
  summarize(f"{SRC}/r02_randomized_v3.ncu-rep", "profiles/r02_ncu_full_randomized_v3.txt",
            "ncu --set full --clock-control none --import-source on, SBX_RESIDENT_V3=1 `python bench.py` (32768 randomized "
            "64x96 buildings, 1 x B200): k_resident_step3, the opt-in half-tile kernel (profiles/capture_r02.sh)",
            "k_resident_step3", envs=32768)
h2, u2, d2 = summarize(f"{SRC}/r02_office.ncu-rep", "profiles/r02_ncu_full_office.txt",
                       "ncu --set full --clock-control none --import-source on, `python bench.py --workload office "
                       "--envs-per-gpu 512` (calibrated sb1 plan 744x1004, 512 copies so that the ~40 replays stay short, "
                       "device-RNG convection as shipped; host-polled sweep loop, see profiles/capture_r02_office.sh): k_sweep_tma "
                       "(TMA-staged tiles, the default sweep), k_convect_reduce (profiles/capture_r02_office.sh)",
                       "k_sweep")
for name in ("randomized", "office"):
  rows = list(csv.reader(open(f"{SRC}/r02_launches_{name}.csv")))
  out = [[r[0], r[4].replace("sbx::", ""), r[7], r[8], r[-3], r[-2], r[-1]]
         for r in rows if len(r) > 10 and (r[0] == "ID" or "sbx::" in r[4])]
  csv.writer(open(f"profiles/r02_launches_{name}.csv", "w", newline="")).writerows(out)
ir = row_of(h1, d1, "k_resident_step")
sweeps = [i for i, r in enumerate(d2) if "k_sweep" in r[h2.index("Kernel Name")]]
isw = max(sweeps, key=lambda i: col(h2, u2, d2, "dram__bytes_read.sum", i))   # a sweep >= 2 (reads T_est and T_prev)
traffic = {
    "randomized": {"kernel": "k_resident_step<4>", "launch_envs": 32768,
                   "dram_bytes_read": col(h1, u1, d1, "dram__bytes_read.sum", ir),
                   "dram_bytes_write": col(h1, u1, d1, "dram__bytes_write.sum", ir),
                   "source": "profiles/r02_ncu_full_randomized.txt (ncu --set full, 32768 buildings per launch)"},
    "office": {"kernel": "k_sweep_tma<0>", "launch_envs": 512,
               "dram_bytes_read": col(h2, u2, d2, "dram__bytes_read.sum", isw),
               "dram_bytes_write": col(h2, u2, d2, "dram__bytes_write.sum", isw),
               "source": "profiles/r02_ncu_full_office.txt (ncu --set full, 512 buildings per launch, a sweep >= 2: reads "
                         "T_est and T_prev)"}}
json.dump(traffic, open("profiles/traffic.json", "w"), indent=1)
for f in ("r02_bench.json", "r02_bench_reference.json"):
  if os.path.exists(f"{SRC}/{f}"):
    shutil.copy(f"{SRC}/{f}", f"profiles/{f}")

# SASS mnemonics per kernel of the shipped library
sass = run("cuobjdump -sass sbsim_b200/lib/libsbx.so")
want = ["UTMALDG", "UTMASTG", "UBLKCP", "UBLKPF", "SYNCS", "FFMA2", "FMUL2", "FADD2", "REDUX", "CREDUX", "ATOMS", "LDS", "STS",
        "LDG", "STG", "BAR", "LDL", "STL", "CCTL", "SHFL"]
out, cur, cnt = ["SASS mnemonic counts per kernel of sbsim_b200/lib/libsbx.so (cuobjdump -sass; static instruction counts)", ""], None, None
for line in sass.splitlines():
  m = re.search(r"Function : (\S+)", line)
  if m:
    if cur:
      out.append(f"{cur}: total {cnt['_total']}  " + "  ".join(f"{k} {cnt[k]}" for k in want if cnt[k]))
    cur, cnt = run(f"c++filt {m.group(1)}").strip()[:90], collections.Counter()
    continue
  m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
  if m and cur:
    cnt["_total"] += 1
    op = m.group(1)
    for k in want:
      if op.startswith(k):
        cnt[k] += 1
        break
if cur:
  out.append(f"{cur}: total {cnt['_total']}  " + "  ".join(f"{k} {cnt[k]}" for k in want if cnt[k]))
open("profiles/r02_sass_mnemonics.txt", "w").write("\n".join(out) + "\n")
print(json.dumps(traffic, indent=1))
