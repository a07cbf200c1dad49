#!/usr/bin/env python
"""Benchmark of the sbsim hot path (one Environment.step() over a batch of buildings).

  python bench.py --gpus N --steps K --warmup W             # libsbx on B200
  python bench.py --impl reference --gpus N --steps K ...    # CPU restatement of the reference

Metric: building-env-steps/s (BASELINE.json).  Workload (config.workload):
"randomized" = BASELINE.json configs[3] sharded per GPU (32768 randomised 64x96
buildings per GPU; N=8 is exactly configs[3], N=1 with --envs-per-gpu 65536 is
configs[2]); "office" = configs[1]: copies of the calibrated sb1 building
(744x1004 CVs, 126 zones, reset_temps.npy, Moffett Field replay weather, the shipped
randomized occupancy; sim_config.gin:160-196) from the committed fixture
tests/golden/sb1_calibrated.npz; with --envs-per-gpu 1 it is configs[0].

One JSON line is printed by rank 0.  Keys follow the driver's contract:
value = whole-job steps/s with inputs resident in HBM (actions already on the
device, outputs left on the device), e2e = the same metric through the public
host API `Environment.step(np.ndarray)` (pinned staging, H2D + D2H inside the
timed region), roofline = achieved algorithmic HBM bytes/s of the step kernel
against the measured copy bandwidth, cpu_baseline = the oracle timed on the
host cores for a bounded sample.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CALIBRATED_FIXTURE = os.path.join(ROOT, "tests", "golden", "sb1_calibrated.npz")
METRIC = "building_env_steps_per_sec"
UNIT = "env-steps/s"


def parse_args():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=200)
  ap.add_argument("--warmup", type=int, default=5)
  ap.add_argument("--impl", choices=["sbx", "reference"], default="sbx")
  ap.add_argument("--workload", choices=["randomized", "office", "sac_rollout"], default="randomized")
  ap.add_argument("--envs-per-gpu", type=int, default=None)
  ap.add_argument("--layouts", type=int, default=1024)
  ap.add_argument("--histogram", type=int, default=1)
  ap.add_argument("--allgather", type=int, default=0,
                  help="all-gather obs/reward/step_type to every rank each step (NCCL)")
  ap.add_argument("--cpu-sample-envs", type=int, default=None)
  ap.add_argument("--cpu-sample-steps", type=int, default=24)
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--no-e2e", action="store_true")
  ap.add_argument("--others", type=int, default=1,
                  help="also measure BASELINE.json's other single-GPU configurations (N=1 only)")
  ap.add_argument("--path", choices=["auto", "streaming", "resident"], default="auto")
  ap.add_argument("--convergence-threshold", type=float, default=0.1,
                  help="experiments only; BASELINE uses 0.1 K (sim_config.gin:161)")
  ap.add_argument("--iteration-limit", type=int, default=100)
  ap.add_argument("--prefetch", type=int, default=-1,
                  help="SBX_OPT_L2_PREFETCH_DISTANCE override (-1 = the library's choice)")
  ap.add_argument("--convection", choices=["default", "off", "device"], default="default",
                  help="stochastic convection: the office workload runs the shipped model "
                       "(p=1, distance=5, seed=5; sim_config.gin:37-39) in device-RNG mode by "
                       "default, the randomized workload has it off (SURVEY 8d config 3)")
  ap.add_argument("--zone-means", choices=["exact-integer", "numpy"], default="exact-integer",
                  help="numpy: SBX_OPT_NUMPY_MEANS, means summed in the order of the reference's np.mean (bit-identical)")
  ap.add_argument("--host-shares", type=int, default=0,
                  help="SBX_OPT_HOST_SHARES override (0 = the library's choice)")
  ap.add_argument("--chunks", type=int, default=0,
                  help="SBX_OPT_PIPELINE_CHUNKS override (0 = the library's choice)")
  return ap.parse_args()


def dist_env():
  return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
          int(os.environ.get("WORLD_SIZE", 1)))


class ClockSampler(threading.Thread):
  """Samples SM clocks and throttle reasons with nvidia-smi during the timed region."""

  Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
       "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
       "clocks_event_reasons.sw_power_cap")

  def __init__(self, index: int):
    super().__init__(daemon=True)
    self.index = index
    self.samples = []
    self._stop_evt = threading.Event()

  def _nvml_loop(self) -> bool:
    """NVML sampling (~0.1 ms per query instead of ~0.3 s per nvidia-smi process)."""
    try:
      import pynvml
      pynvml.nvmlInit()
      visible = os.environ.get("CUDA_VISIBLE_DEVICES")
      index = int(visible.split(",")[self.index]) if visible else self.index
      dev = pynvml.nvmlDeviceGetHandleByIndex(index)
      bits = {"hw_slowdown": pynvml.nvmlClocksThrottleReasonHwSlowdown,
              "hw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
              "sw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonSwThermalSlowdown,
              "sw_power_cap": pynvml.nvmlClocksThrottleReasonSwPowerCap}
      mx = pynvml.nvmlDeviceGetMaxClockInfo(dev, pynvml.NVML_CLOCK_SM)
    except Exception:  # pylint: disable=broad-except
      return False
    while not self._stop_evt.is_set():
      try:
        sm = pynvml.nvmlDeviceGetClockInfo(dev, pynvml.NVML_CLOCK_SM)
        reasons = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(dev)
        power = pynvml.nvmlDeviceGetPowerUsage(dev) / 1000.0
        flags = ["Active" if reasons & bits[n] else "Not Active"
                 for n in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")]
        self.samples.append([str(sm), str(mx), str(power)] + flags)
      except Exception:  # pylint: disable=broad-except
        pass
      self._stop_evt.wait(0.01)
    return True

  def run(self):
    if self._nvml_loop():
      return
    while not self._stop_evt.is_set():
      try:
        out = subprocess.run(
            ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
             "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        parts = [x.strip() for x in out.strip().split(",")]
        if len(parts) >= 7:
          self.samples.append(parts)
      except Exception:  # pylint: disable=broad-except
        pass
      self._stop_evt.wait(0.05)

  def stop(self):
    self._stop_evt.set()
    self.join(timeout=3)

  def summary(self):
    if not self.samples:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
    sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
    mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = sorted({n for s in self.samples for n, v in zip(names, s[3:7]) if v == "Active"})
    return {"sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
            "samples": len(self.samples)}


def measured_peak_gbs():
  path = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(path):
    with open(path) as f:
      return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
  return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def algorithmic_bytes_per_env_step(n_cv, n_zones, obs_dim, n_actions, desc_bytes, sweeps, resident):
  """SURVEY.md section 8d.  Resident: N*(4 read + 4 write) + N*s + 64*Z + 4*(D+A+3).
  Streaming: n*N*12 + N*s (+ the same small terms)."""
  small = 64 * n_zones + 4 * (obs_dim + n_actions + 3)
  if resident:
    return n_cv * 8 + n_cv * desc_bytes + small
  return sweeps * n_cv * 12 + n_cv * desc_bytes + small


def build_env(args, rank, local_rank):
  import sbsim_b200 as sbx
  from sbsim_b200 import floorplan, workloads
  path = {"auto": sbx.PATH_AUTO, "streaming": sbx.PATH_STREAMING,
          "resident": sbx.PATH_RESIDENT}[args.path]
  episode = args.warmup + args.steps + 16 + (24 if dist_env()[2] > 1 else 0)
  if args.workload in ("randomized", "sac_rollout"):
    n = args.envs_per_gpu or (1024 if args.workload == "sac_rollout" else 32768)
    if args.workload == "sac_rollout":
      episode = max(episode, 288)
    env, wl = workloads.make_randomized_env(
        n, seed=2024 + rank, episode_steps=episode, n_layouts=args.layouts,
        histogram=bool(args.histogram), device=local_rank, kernel_path=path,
        convergence_threshold=args.convergence_threshold, iteration_limit=args.iteration_limit,
        numpy_zone_means=args.zone_means == "numpy")
    if args.convection == "device":
      env.handle.set_device_convection(*workloads.CALIBRATED_CONVECTION)
    desc = {"workload": ("SAC rollout: randomized-64x96 buildings, one-day (288-step) episode, actions from a "
                         "fixed-seed tanh-Gaussian policy stub drawn on the device (BASELINE.json configs[4] per-GPU shard)"
                         if args.workload == "sac_rollout" else
                         "randomized-64x96 (BASELINE.json configs[3] per-GPU shard)"),
            "envs_per_gpu": n, "grid": [64, 96], "layouts": wl.n_layouts,
            "plans": "per-env descriptor, materials, weather, T0, actions",
            "stochastic_convection": "device-rng mode" if args.convection == "device" else "off",
            "zone_means": args.zone_means}
    return env, wl, desc
  n = args.envs_per_gpu or 4096
  cal = workloads.load_calibrated(CALIBRATED_FIXTURE)
  cp = cal.plan
  conv_on = args.convection in ("default", "device")
  conv = (sbx.StochasticConvectionSimulator(*workloads.CALIBRATED_CONVECTION, mode="device")
          if conv_on else None)
  env = workloads.make_calibrated_env(cal, n, episode_steps=episode,
                                      histogram=bool(args.histogram), device=local_rank,
                                      kernel_path=path, convection=conv,
                                      numpy_zone_means=args.zone_means == "numpy")
  desc = {"workload": "calibrated sb1 building 744x1004 x copies (BASELINE.json configs[1]; "
                      "sim_config.gin:160-196: TF-Jacobi, reset_temps.npy, Moffett replay weather, "
                      "US/Pacific schedule, RandomizedArrivalDepartureOccupancy seed 17321)",
          "envs_per_gpu": n, "grid": [cp.height, cp.width], "zones": cp.n_zones,
          "plans": "one shared descriptor",
          "stochastic_convection": ("device-rng mode, p=1 distance=5 seed=5 (sim_config.gin:37-39): "
                                    "k_convect_reduce fuses it with the zone reduction") if conv_on else "off",
          "zone_means": args.zone_means}
  return env, None, desc


def run_sbx(args):
  import copy
  import torch
  import torch.distributed as dist
  rank, local_rank, world = dist_env()
  if world != args.gpus and world > 1:
    raise SystemExit(f"WORLD_SIZE {world} != --gpus {args.gpus}")
  if not torch.cuda.is_available():
    raise SystemExit("bench.py needs a CUDA device; sbsim_b200 has no CPU fallback")
  if args.warmup < 3:
    raise SystemExit("--warmup must be >= 3")
  torch.cuda.set_device(local_rank)
  dev = torch.device("cuda", local_rank)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)
  line = _measure(args, torch, dist, rank, local_rank, world, dev, with_cpu=True)
  # The other single-GPU configurations of BASELINE.json, measured briefly beside the
  # headline (N=1 only): configs[1]'s size class (4096 copies of one 744x1004 plan,
  # streaming path) and configs[2] (65536 randomised buildings on one GPU).
  if args.others and args.workload == "randomized" and args.envs_per_gpu is None:
    others = []
    todo = [("sac_rollout", 1024, 280, "default")]          # configs[4]: at every N
    if world == 1:
      todo = [("office", 4096, 16, "default"), ("office", 4096, 16, "off"),
              ("randomized", 65536, 50, "default")] + todo
    for wl_name, envs, steps, conv in todo:
      a2 = copy.copy(args)
      a2.workload, a2.envs_per_gpu, a2.steps, a2.warmup, a2.convection = wl_name, envs, steps, 3, conv
      try:
        l2 = _measure(a2, torch, dist, rank, local_rank, world, dev,
                      with_cpu=(wl_name == "office" and conv == "default"))
        others.append({k: l2[k] for k in ("value", "unit", "steps", "warmup", "ms_per_step",
                                          "config", "gpu_launches", "roofline", "e2e",
                                          "cpu_baseline", "with_allgather") if k in l2})
      except Exception as e:  # pylint: disable=broad-except
        others.append({"config": {"workload": wl_name, "envs_per_gpu": envs},
                       "error": f"{type(e).__name__}: {e}"})
    line["other_configs"] = others
  if rank == 0:
    emit(line)
  if world > 1:
    dist.destroy_process_group()


def _measure(args, torch, dist, rank, local_rank, world, dev, with_cpu):
  env, wl, cfg_desc = build_env(args, rank, local_rank)
  from sbsim_b200 import _lib
  if args.chunks > 0:
    env.handle.set_option(_lib.OPT_PIPELINE_CHUNKS, args.chunks)
  if args.prefetch >= 0:
    env.handle.set_option(_lib.OPT_L2_PREFETCH_DISTANCE, args.prefetch)
  if args.host_shares > 0:
    env.handle.set_option(_lib.OPT_HOST_SHARES, args.host_shares)
  B = env.batch_size
  D = env.observation_spec().shape[0]
  A = env.action_spec().shape[0]
  W, K = args.warmup, args.steps

  gen = torch.Generator(device=dev)
  gen.manual_seed(3000 + rank)
  n_extra = 24 if world > 1 else 0       # steps of the all-gather variant timed after the main loop
  if args.workload == "sac_rollout":     # tanh-Gaussian policy stub (the TF-Agents learner is not in this image)
    actions = torch.tanh(torch.randn((W + K + n_extra, B, A), device=dev, generator=gen))
  else:
    actions = torch.rand((W + K + n_extra, B, A), device=dev, generator=gen) * 2.0 - 1.0
  from sbsim_b200 import distributed
  packed = distributed.PackedTimeStep(B, D, dev)   # [obs | reward | step_type | discount] in one block
  obs, rew, st, dis = packed.obs, packed.reward, packed.step_type, packed.discount
  stream = torch.cuda.current_stream(dev)
  sptr = stream.cuda_stream
  gather = distributed.TimeStepGather(packed) if world > 1 else None
  gathered = gather if (world > 1 and args.allgather) else None

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize(dev)

  def one_step(i):
    env.step_device(actions[i], obs, rew, st, dis, stream=sptr)
    if gathered is not None:
      gathered()

  # ---- device-resident timing (value) ----
  env.reset_device(obs, rew, st, dis, stream=sptr)
  for i in range(W):
    one_step(i)
  barrier()
  info0 = env.handle.info()
  sampler = ClockSampler(local_rank)
  sampler.start()
  ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
  barrier()
  ev[0].record(stream)
  for i in range(K):
    one_step(W + i)
    ev[i + 1].record(stream)
  barrier()
  info1 = env.handle.info()
  per_step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(K)]
  total_ms = ev[0].elapsed_time(ev[K])
  t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  max_ms = float(t.item())
  launches = int(info1.kernel_launches - info0.kernel_launches)
  sweeps = int(info1.sweeps_total - info0.sweeps_total)
  mean_sweeps = sweeps / float(B * K)
  value = world * B * K / (max_ms / 1e3)

  # ---- the optional collective (BASELINE.json configs[3]: "optional NCCL obs all-gather"):
  # the same steps continued with ONE ncclAllGather of the packed TimeStep block per step,
  # and the collective alone for its bus bandwidth ----
  allgather = None
  if gather is not None and not args.allgather:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(n_extra):
      env.step_device(actions[W + K + i], obs, rew, st, dis, stream=sptr)
      gather()
    e1.record(stream)
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ag_ms = float(t.item()) / n_extra
    barrier()
    e0.record(stream)
    for i in range(20):
      gather()
    e1.record(stream)
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    coll_ms = float(t.item()) / 20
    total_bytes = gather.bytes_per_rank * world
    allgather = {"value": world * B / (ag_ms / 1e3), "unit": UNIT, "ms_per_step": ag_ms, "steps": n_extra,
                 "collective": "ncclAllGather (torch.distributed.all_gather_into_tensor) of one packed "
                               "[obs | reward | step_type | discount] block per rank, every step, to every rank",
                 "bytes_per_rank": gather.bytes_per_rank, "collective_ms": coll_ms,
                 "algbw_gbs": total_bytes / (coll_ms / 1e3) / 1e9,
                 "busbw_gbs": total_bytes * (world - 1) / world / (coll_ms / 1e3) / 1e9}

  # ---- kernel-timing pass: CUDA events recorded by the library around every solve
  # launch, on the stream the kernel is launched on (sbx_timing_begin / _end) ----
  # The SAME trajectory again (reset, same actions: the step is deterministic, so the
  # sweep counts are identical to the value pass), this time with the events on.
  env.reset_device(obs, rew, st, dis, stream=sptr)
  for i in range(W):
    one_step(i)
  barrier()
  info1b = env.handle.info()
  env.handle.timing_begin()
  for i in range(K):
    one_step(W + i)
  barrier()
  tm = env.handle.timing_end()
  sampler.stop()                      # sampled over both timed loops (value pass + event pass)
  solve_ms = tm.solve_ms / max(tm.n_steps, 1)           # all solve launches of one step
  timed_step_ms = tm.step_ms / max(tm.n_steps, 1)
  info2 = env.handle.info()
  sweeps_t = int(info2.sweeps_total - info1b.sweeps_total) / float(B * max(tm.n_steps, 1))

  # ---- end-to-end through the public host API (e2e) ----
  e2e = None
  if not args.no_e2e:
    a_host = (np.random.default_rng(4000 + rank).uniform(-1, 1, (W + K, B, A))
              .astype(np.float32))
    env.reset()
    for i in range(W):
      env.step(a_host[i])
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
      ts = env.step(a_host[W + i])
    torch.cuda.synchronize(dev)
    el = time.perf_counter() - t0
    t = torch.tensor([el], device=dev, dtype=torch.float64)
    if world > 1:
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
    el = float(t.item())
    e2e = {"value": world * B * K / el, "unit": UNIT,
           "h2d_bytes_per_step": int(B * A * 4),
           "d2h_bytes_per_step": int(B * (D + 3) * 4),
           "api": "sbsim_b200.Environment.step(np.ndarray) -> TimeStep (sbx_step_host)",
           "ms_per_step": el / K * 1e3}
    del ts

  # ---- roofline of the dominant kernel ----
  resident = env.kernel_path == 2
  n_cv = env.building.height * env.building.width
  Z = env.building.n_zones
  desc_bytes = 2 if len(env.building.plans) > 1 else 0
  bytes_step = algorithmic_bytes_per_env_step(n_cv, Z, D, A, desc_bytes, mean_sweeps, resident)
  peak, peak_src = measured_peak_gbs()
  # achieved = algorithmic bytes of one step's solve launches / their measured duration
  bytes_step_t = algorithmic_bytes_per_env_step(n_cv, Z, D, A, desc_bytes, sweeps_t, resident)
  bytes_launch = bytes_step_t * B
  launch_ms = solve_ms
  achieved = bytes_launch / (launch_ms / 1e3) / 1e9
  if resident:
    kname = {1: "k_resident_step", 2: "k_resident_step2", 3: "k_resident_step3"}.get(
        int(env.handle.info().resident_kernel), "k_resident_step")
    kernel = ("%s (the whole diffusion solve of every building; %d launch(es) per "
              "step, one per pipelined share of the batch)" % (kname, tm.n_chunks))
  else:
    kernel = "k_sweep_tma / k_sweep (one Jacobi sweep of every active building per launch; %.2f launches per step)" % (
        tm.n_solve_launches / max(tm.n_steps, 1))
  traffic = None
  try:  # DRAM bytes per launch from the committed ncu --set full capture, scaled to this B
    with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
      tr = json.load(f).get(args.workload)
    if tr and (resident or args.workload == "office"):
      # per launch of the dominant kernel over all B buildings (streaming: one sweep >= 2)
      traffic = (tr["dram_bytes_read"] + tr["dram_bytes_write"]) * B / tr["launch_envs"]
  except Exception:  # pylint: disable=broad-except
    traffic = None
  roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak,
              "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
              "traffic": traffic, "algorithmic_bytes_per_launch": bytes_launch,
              "algorithmic_bytes_per_env_step": bytes_step_t,
              "mean_sweeps_per_step": sweeps_t,
              "launch_ms": launch_ms, "timed_steps": int(tm.n_steps),
              "timing": "cudaEvent pairs around each solve launch on its stream (sbx_timing_*)",
              # the same bytes over the WHOLE step (HVAC prologue / epilogue included)
              "step_ms": timed_step_ms, "kernel_share_of_step": launch_ms / timed_step_ms,
              "frac_whole_step": bytes_step * B / (statistics.mean(per_step_ms) / 1e3) / 1e9 / peak}

  # ---- CPU baseline beside it (rank 0, N=1 only) ----
  cpu = None
  if with_cpu and rank == 0 and world == 1 and not args.no_cpu_baseline:
    if args.workload == "randomized":
      cpu = cpu_baseline(args, wl, K=args.cpu_sample_steps)
    else:
      cpu = cpu_baseline_calibrated(min(args.cpu_sample_steps, 8), convection=args.convection != "off")

  line = {
      "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
      "warmup": W, "ms_per_step": max_ms / K, "higher_is_better": True, "scaling": "weak",
      "vs_baseline": None, "dtype": "f32", "data": "synthetic",
      "config": dict(cfg_desc, global_envs=world * B,
                     kernel_path="resident" if resident else "streaming",
                     obs_dim=D, l2_policy="inputs larger than L2 (state %.0f MB per GPU)"
                     % (B * n_cv * 4 / 1e6),
                     allgather=bool(gathered is not None), episode_steps=288,
                     time_step_sec=300, mean_sweeps_per_step=mean_sweeps),
      "clocks": sampler.summary(), "gpu_launches": launches,
      "roofline": roofline,
  }
  if allgather is not None:
    line["with_allgather"] = allgather
  if e2e is not None:
    line["e2e"] = e2e
  if cpu is not None:
    line["cpu_baseline"] = cpu
  env.close()
  del env, actions, obs, rew, st, dis
  torch.cuda.empty_cache()
  return line


def _oracle_specs(wl, n, episode_steps):
  from sbsim_b200 import workloads
  specs = []
  for b in range(n):
    specs.append((wl.plans[b], float(wl.weather_low[b]), float(wl.weather_high[b]),
                  float(wl.convection[b]), float(wl.initial_temp[b]),
                  workloads.NORMALIZATION, workloads.HISTOGRAM, episode_steps,
                  workloads.DEFAULT_START))
  return specs


def cpu_baseline(args, wl, K):
  """The oracle (CPU restatement of the reference) on a bounded sample."""
  from oracle import bench_support
  cores = bench_support.host_cores()
  n = args.cpu_sample_envs or min(max(cores * 128, 256), 4096)   # ~10 s of CPU work incl. set-up
  n = min(n, len(wl.plans))
  rate, total, wall, sweeps = bench_support.time_oracle(
      _oracle_specs(wl, n, K + 8), K, n_procs=cores)
  return {"value": rate, "unit": UNIT, "cores": min(cores, n), "kind": "port",
          "sample": f"first {n} buildings of the same randomized workload x {K} steps "
                    f"({total} env-steps, {wall:.1f} s wall, {sweeps:.2f} sweeps/step), "
                    "oracle/ NumPy-fp32 restatement, one process per core"}


def cpu_baseline_calibrated(K, warmup=2, convection=True):
  """The oracle on the calibrated building (BASELINE.json configs[0]): one building per
  host core, K timed steps each after `warmup` untimed ones (the first step after
  reset_temps.npy takes 7 sweeps, the steady state about 2)."""
  from oracle import bench_support
  from sbsim_b200 import workloads
  cores = bench_support.host_cores()
  warmup = max(warmup, 1)       # the first convection call builds the reference's candidate cache (~25 s)
  rate, total, wall, sweeps = bench_support.time_calibrated_oracle(
      CALIBRATED_FIXTURE, workloads.NORMALIZATION, workloads.HISTOGRAM, K, warmup, cores,
      convection=workloads.CALIBRATED_CONVECTION if convection else None)
  return {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
          "sample": f"{cores} copies of the calibrated sb1 building (one per core) x {K} steps "
                    f"after {warmup} warm-up steps ({total} env-steps, {wall:.1f} s wall, "
                    f"{sweeps:.2f} sweeps/step), oracle/ NumPy-fp32 restatement of "
                    "TFSimulator + Environment, stochastic convection "
                    + ("as shipped (p=1, distance=5: the reference's per-CV Python loop)"
                       if convection else "off")}


def run_reference(args):
  """--impl reference: the CPU restatement of the reference's own path (oracle
  port; the reference is Python with TensorFlow, which this image lacks) on all
  host cores, on the same workload definition; each step is a bounded sample."""
  rank, _, world = dist_env()
  if rank != 0:
    return
  from oracle import bench_support
  from sbsim_b200 import workloads
  if args.workload != "randomized":
    # calibrated building: ~7 env-steps/s/core -> bound the run to ~12 timed steps
    K_eff = max(1, min(args.steps, 8))
    cpu = cpu_baseline_calibrated(K_eff, warmup=min(max(args.warmup, 1), 2), convection=args.convection != "off")
    n = cpu["cores"]
    emit({
        "impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": K_eff, "warmup": min(max(args.warmup, 0), 3),
        "ms_per_step": 1e3 * n / cpu["value"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "calibrated sb1 building 744x1004 (BASELINE.json configs[0/1])",
                   "sample_envs": n, "grid": [744, 1004]},
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0}})
    return
  cores = bench_support.host_cores()
  n = args.cpu_sample_envs or min(max(cores * 128, 256), 4096)
  W, K = args.warmup, args.steps
  # bound the run: ~250 env-steps/s/core => keep total env-steps near 30 s of work
  budget = int(250 * cores * 30)
  K_eff = max(1, min(K, budget // max(n, 1)))
  wl = workloads.randomized(n, seed=2024, n_layouts=min(args.layouts, n))
  rate, total, wall, sweeps = bench_support.time_oracle(
      _oracle_specs(wl, n, K_eff + W + 8), K_eff, n_procs=cores)
  line = {
      "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT,
      "n_gpus": args.gpus, "steps": K_eff, "warmup": 0, "ms_per_step": 1e3 * n / rate,
      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
      "data": "synthetic",
      "config": {"workload": "randomized-64x96 (BASELINE.json configs[3] per-GPU shard)",
                 "sample_envs": n, "grid": [64, 96]},
      "cpu_baseline": {"value": rate, "unit": UNIT, "cores": min(cores, n), "kind": "port",
                       "sample": f"{n} buildings x {K_eff} steps ({total} env-steps, "
                                 f"{wall:.1f} s wall, {sweeps:.2f} sweeps/step)"},
      "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
  }
  emit(line)


_REAL_STDOUT = None


def protect_stdout():
  """stdout carries exactly ONE JSON line.  Libraries print there too (NCCL's version
  banner with NCCL_DEBUG=VERSION, as on the GPU box), so file descriptor 1 is pointed at
  stderr for the whole run and the JSON line goes to the saved descriptor."""
  global _REAL_STDOUT
  sys.stdout.flush()
  _REAL_STDOUT = os.dup(1)
  os.dup2(2, 1)


def emit(line):
  data = (json.dumps(line) + "\n").encode()
  if _REAL_STDOUT is None:
    sys.stdout.write(data.decode())
    sys.stdout.flush()
  else:
    os.write(_REAL_STDOUT, data)


def main():
  args = parse_args()
  protect_stdout()
  if args.impl == "reference":
    run_reference(args)
  else:
    run_sbx(args)


if __name__ == "__main__":
  main()
