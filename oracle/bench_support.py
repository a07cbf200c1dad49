"""Oracle-side twin of sbsim_b200.workloads (TEST / BASELINE INFRASTRUCTURE ONLY).

Builds `OracleEnvironment`s for the randomized workload so that bench.py's
`cpu_baseline` and `--impl reference` legs time the CPU restatement of the
reference on the same buildings, weather and action streams as the GPU arm.
"""

from __future__ import annotations

import multiprocessing as mp
import os
import time
from typing import List, Tuple

import numpy as np
import pandas as pd

import random

from oracle import convection as oconv
from oracle import env as oenv
from oracle import exogenous as oex
from oracle import hvac as ohvac
from oracle import reward as orew
from oracle import tf_jacobi


def oracle_plan(cp, floor_height_cm: float = 300.0) -> tf_jacobi.OraclePlan:
  rooms = []
  for zi, name in enumerate(cp.zone_names):
    r, c = cp.zone_indices(zi)
    rooms.append((name, r, c))
  return tf_jacobi.OraclePlan(
      exterior_space=cp.exterior_space, conductivity=cp.dense_material(0),
      heat_capacity=cp.dense_material(1), density=cp.dense_material(2),
      diffusers=cp.diffuser_weight, rooms=rooms, cv_size_cm=cp.cv_size_m * 100.0,
      floor_height_cm=floor_height_cm)


def f32(x):
  return float(np.float32(x))


def make_oracle_env(cp, low: float, high: float, convection: float, initial_temp: float,
                    normalization, histogram, episode_steps: int, start: str,
                    occupancy_norm: float = 125.0) -> oenv.OracleEnvironment:
  """Calibrated device / reward parameters (sim_config.gin:101-225)."""
  cfg = oenv.OracleEnvConfig(
      plan=oracle_plan(cp), start_timestamp=pd.Timestamp(start),
      weather=oex.WeatherController(low, high, convection_coefficient=convection),
      schedule=ohvac.SetpointSchedule(6, 19, (294, 297), (289, 298)),
      occupancy=oex.StepFunctionOccupancy(pd.Timedelta(9, unit="h"),
                                          pd.Timedelta(17, unit="h"), 1.0, 0.1),
      reward_function=orew.SetpointEnergyCarbonRegretFunction(
          300.0, 100.0, 160000, 400000, 0.5, 4.3, oex.ElectricityEnergyCost(),
          oex.NaturalGasEnergyCost(), 0.2, 0.4, 0.4),
      solver="tf", initial_temp=float(initial_temp),
      normalization={k: (f32(m), f32(v)) for k, (m, v) in normalization.items()},
      histogram=histogram, discount_factor=0.9, num_timesteps_in_episode=episode_steps,
      occupancy_normalization_constant=occupancy_norm)
  return oenv.OracleEnvironment(cfg)


def make_calibrated_oracle_env(fixture_path: str, normalization, histogram, episode_steps: int,
                               occupancy: str = "randomized", convection=None
                               ) -> oenv.OracleEnvironment:
  """BASELINE.json configs[0]: the calibrated sb1 building (sim_config.gin:160-196) --
  TF-Jacobi solver, reset_temps.npy, Moffett replay weather, US/Pacific schedule, the
  shipped randomized occupancy -- from the committed fixture of its resource files."""
  from sbsim_b200 import workloads      # plan compiler + the gin constants (host-side data only)
  cal = workloads.load_calibrated(fixture_path)
  occ = (oex.RandomizedArrivalDepartureOccupancy(1, 7, 12, 13, 18, 300, seed=17321,
                                                 time_zone=workloads.CALIBRATED_TZ)
         if occupancy == "randomized" else oex.ConstantOccupancy(0.7))
  cfg = oenv.OracleEnvConfig(
      plan=oracle_plan(cal.plan), start_timestamp=pd.Timestamp(workloads.CALIBRATED_START),
      weather=oex.ReplayWeatherController(cal.weather_time_sec, cal.weather_temp_f, 100.0),
      schedule=ohvac.SetpointSchedule(6, 19, (294, 297), (289, 298),
                                      time_zone=workloads.CALIBRATED_TZ),
      occupancy=occ,
      reward_function=orew.SetpointEnergyCarbonRegretFunction(
          300.0, 100.0, 160000, 400000, 0.5, 4.3, oex.ElectricityEnergyCost(),
          oex.NaturalGasEnergyCost(), 0.2, 0.4, 0.4),
      solver="tf", initial_temp=294.0, reset_temp_values=cal.reset_temps,
      convection=(oconv.StochasticConvectionSimulator(convection[0], convection[1], convection[2],
                                                      rng=random.Random(convection[2]))
                  if convection else None),
      normalization={k: (f32(m), f32(v)) for k, (m, v) in normalization.items()},
      histogram=histogram, discount_factor=0.9, num_timesteps_in_episode=episode_steps,
      occupancy_normalization_constant=125.0)
  return oenv.OracleEnvironment(cfg)


def _calibrated_worker(args) -> Tuple[int, float, int]:
  fixture, normalization, histogram, steps, warmup, seed, convection = args
  env = make_calibrated_oracle_env(fixture, normalization, histogram, steps + warmup + 8,
                                   convection=convection)
  rng = np.random.default_rng(seed)
  env.reset()
  for _ in range(warmup):
    env.step(rng.uniform(-1, 1, 2).astype(np.float32))
  sweeps = 0
  t0 = time.perf_counter()
  for _ in range(steps):
    env.step(rng.uniform(-1, 1, 2).astype(np.float32))
    sweeps += env.info["n_sweeps"]
  return steps, time.perf_counter() - t0, sweeps


def time_calibrated_oracle(fixture: str, normalization, histogram, steps: int, warmup: int,
                           n_procs: int, convection=None):
  """One calibrated building per host core, env b driven by default_rng(1000 + b)
  (SURVEY 8d config 2's action streams).  Returns like time_oracle."""
  jobs = [(fixture, normalization, histogram, steps, warmup, 1000 + i, convection) for i in range(n_procs)]
  t0 = time.perf_counter()
  if n_procs == 1:
    res = [_calibrated_worker(jobs[0])]
  else:
    ctx = mp.get_context("fork")
    with ctx.Pool(n_procs) as pool:
      res = pool.map(_calibrated_worker, jobs)
  wall = time.perf_counter() - t0
  total = sum(r[0] for r in res)
  busy = max(r[1] for r in res)
  return total / busy, total, wall, sum(r[2] for r in res) / max(total, 1)


_WORKER_ENVS: List[oenv.OracleEnvironment] = []


def _worker(args) -> Tuple[int, float, int]:
  """Steps a slice of buildings `steps` times; returns (env_steps, seconds, sweeps)."""
  specs, steps, seed = args
  envs = [make_oracle_env(*s) for s in specs]
  rng = np.random.default_rng(seed)
  for e in envs:
    e.reset()
  sweeps = 0
  t0 = time.perf_counter()
  for _ in range(steps):
    for e in envs:
      e.step(rng.uniform(-1, 1, 2).astype(np.float32))
      sweeps += e.info["n_sweeps"]
  return len(envs) * steps, time.perf_counter() - t0, sweeps


def time_oracle(specs: list, steps: int, n_procs: int, seed: int = 0):
  """Runs the oracle on `specs` (one tuple of make_oracle_env args per building),
  fanned out over n_procs processes.  Returns (env_steps_per_sec, total env steps,
  wall seconds, mean sweeps per step)."""
  n_procs = max(1, min(n_procs, len(specs)))
  chunks = [specs[i::n_procs] for i in range(n_procs)]
  jobs = [(c, steps, seed + i) for i, c in enumerate(chunks) if c]
  t0 = time.perf_counter()
  if n_procs == 1:
    res = [_worker(jobs[0])]
  else:
    ctx = mp.get_context("fork")
    with ctx.Pool(len(jobs)) as pool:
      res = pool.map(_worker, jobs)
  wall = time.perf_counter() - t0
  total = sum(r[0] for r in res)
  # steady-state throughput: every worker runs concurrently; the slowest bounds it
  busy = max(r[1] for r in res)
  sweeps = sum(r[2] for r in res)
  return total / busy, total, wall, sweeps / max(total, 1)


def host_cores() -> int:
  try:
    return len(os.sched_getaffinity(0))
  except AttributeError:
    return os.cpu_count() or 1
