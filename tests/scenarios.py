"""Shared scenario builders: the SAME parameters feed the product
(`sbsim_b200.Environment`, CUDA) and the checker (`oracle.env.OracleEnvironment`)."""

from __future__ import annotations

import dataclasses
import random
from typing import List, Optional, Sequence

import numpy as np
import pandas as pd

import sbsim_b200 as sbx
from sbsim_b200 import floorplan
from oracle import convection as oconv
from oracle import env as oenv
from oracle import exogenous as oex
from oracle import hvac as ohvac
from oracle import reward as orew
from oracle import tf_jacobi

# plan of tf_simulator_test.py:36-52 (re-typed; non-rectangular outline)
TF_TEST_PLAN = np.array([
    [2, 2, 2, 2, 2, 2, 2, 2, 2],
    [2, 1, 1, 2, 2, 2, 1, 1, 2],
    [2, 1, 0, 1, 2, 1, 0, 1, 2],
    [2, 1, 0, 0, 1, 0, 0, 1, 2],
    [2, 1, 0, 0, 1, 0, 0, 1, 2],
    [2, 1, 1, 1, 1, 1, 1, 1, 2],
    [2, 1, 0, 0, 1, 0, 0, 1, 2],
    [2, 1, 0, 0, 1, 0, 0, 1, 2],
    [2, 1, 1, 1, 1, 1, 1, 1, 2],
    [2, 2, 2, 2, 2, 2, 2, 2, 2],
])

NORMALIZATION = {  # subset of sim_config.gin:253-580 that matches simulated device fields
    "supply_water_setpoint": (320.261985, 240.195517),
    "supply_water_temperature_sensor": (321.520315, 658.413066),
    "outside_air_temperature_sensor": (291.244931, 12.904175),
    "outside_air_flowrate_sensor": (3.701930, 20.300565),
    "supply_air_cooling_temperature_setpoint": (289.329414, 3.186769),
    "supply_air_heating_temperature_setpoint": (289.329414, 3.186769),
    "supply_fan_speed_percentage_command": (26.543748, 575.094979),
    "differential_pressure_setpoint": (83810.269540, 14889040.603647),
    "cooling_request_count": (100.0, 25.0),
    "zone_air_temperature_sensor": (290.0, 25.0),
}
HISTOGRAM = (  # sim_config.gin:586-590, zone temperature bins in normalised units here
    ("zone_air_temperature_sensor", tuple(np.linspace(-1.03, 2.57, 19))),
    ("supply_air_damper_percentage_command", (0.0, 0.2, 0.4, 0.6, 0.8, 1.0)),
    ("supply_air_flowrate_setpoint", (0., 0.05, .1, .2, .3, .4, .5, .7, .9)),
)


def f32(x):
  return float(np.float32(x))


@dataclasses.dataclass
class Scenario:
  floor_plan: np.ndarray
  cv_size_cm: float = 20.0
  floor_height_cm: float = 300.0
  air: tuple = (50.0, 700.0, 1.0)
  wall: tuple = (2.0, 1000.0, 1800.0)
  exterior: tuple = (0.05, 1000.0, 3000.0)
  buffer_from_walls: int = 2
  initial_temp: float = 292.0
  reset_temp_values: Optional[np.ndarray] = None
  start: str = "2023-07-06 05:00:00"
  weather_low: float = 275.0
  weather_high: float = 290.0
  convection_coefficient: float = 60.0
  time_step_sec: float = 300.0
  convergence_threshold: float = 0.1
  iteration_limit: int = 100
  num_days: float = 1.0
  discount: float = 0.9
  occupancy_norm: float = 3.0
  histogram: bool = False
  reward: str = "regret"          # or "energy_carbon" (setpoint_energy_carbon_reward.py)
  schedule_tz: str = "UTC"
  occupancy: str = "step"      # "step" | "const"
  convection: Optional[tuple] = None   # (p, distance, seed) of StochasticConvectionSimulator

  def compiled(self) -> floorplan.CompiledPlan:
    return floorplan.compile_plan(
        self.floor_plan, None, cv_size_cm=self.cv_size_cm,
        inside_air=floorplan.MaterialProperties(*self.air),
        inside_wall=floorplan.MaterialProperties(*self.wall),
        building_exterior=floorplan.MaterialProperties(*self.exterior),
        buffer_from_walls=self.buffer_from_walls)

  @property
  def start_timestamp(self) -> pd.Timestamp:
    return pd.Timestamp(self.start)

  @property
  def n_steps(self) -> int:
    return int(pd.Timedelta(self.num_days, unit="days") / pd.Timedelta(5, unit="minutes"))


def oracle_plan(cp: floorplan.CompiledPlan, floor_height_cm: float) -> tf_jacobi.OraclePlan:
  rooms = []
  for zi, name in enumerate(cp.zone_names):
    r, c = cp.zone_indices(zi)
    rooms.append((name, r, c))
  return tf_jacobi.OraclePlan(
      exterior_space=cp.exterior_space, conductivity=cp.dense_material(0),
      heat_capacity=cp.dense_material(1), density=cp.dense_material(2),
      diffusers=cp.diffuser_weight, rooms=rooms, cv_size_cm=cp.cv_size_m * 100.0,
      floor_height_cm=floor_height_cm)


def make_oracle(sc: Scenario, cp: Optional[floorplan.CompiledPlan] = None,
                weather=None, initial_temp=None, solver="tf") -> oenv.OracleEnvironment:
  cp = cp or sc.compiled()
  if sc.occupancy == "step":
    occ = oex.StepFunctionOccupancy(pd.Timedelta(9, unit="h"), pd.Timedelta(17, unit="h"),
                                    1.0, 0.1)
  else:
    occ = oex.ConstantOccupancy(0.7)
  cfg = oenv.OracleEnvConfig(
      plan=oracle_plan(cp, sc.floor_height_cm), start_timestamp=sc.start_timestamp,
      weather=weather or oex.WeatherController(sc.weather_low, sc.weather_high,
                                               convection_coefficient=sc.convection_coefficient),
      schedule=ohvac.SetpointSchedule(6, 19, (294, 297), (289, 298), time_zone=sc.schedule_tz),
      occupancy=occ,
      reward_function=(orew.SetpointEnergyCarbonRewardFunction(
          300.0, 0.5, 4.3, oex.ElectricityEnergyCost(), oex.NaturalGasEnergyCost(),
          1.5, 2.0, 0.3, 25.0, 400.0) if sc.reward == "energy_carbon" else
                       orew.SetpointEnergyCarbonRegretFunction(
          300.0, 100.0, 160000, 400000, 0.5, 4.3, oex.ElectricityEnergyCost(),
          oex.NaturalGasEnergyCost(), 0.2, 0.4, 0.4)),
      solver=solver, time_step_sec=sc.time_step_sec,
      convergence_threshold=sc.convergence_threshold, iteration_limit=sc.iteration_limit,
      initial_temp=sc.initial_temp if initial_temp is None else initial_temp,
      reset_temp_values=sc.reset_temp_values,
      convection=(oconv.StochasticConvectionSimulator(
          sc.convection[0], sc.convection[1], sc.convection[2],
          rng=random.Random(sc.convection[2])) if sc.convection else None),
      normalization={k: (f32(m), f32(v)) for k, (m, v) in NORMALIZATION.items()},
      histogram=HISTOGRAM if sc.histogram else None,
      discount_factor=sc.discount, num_timesteps_in_episode=sc.n_steps,
      occupancy_normalization_constant=sc.occupancy_norm)
  return oenv.OracleEnvironment(cfg)


def make_env(sc: Scenario, n_envs: int = 1, plans=None, weather=None, initial_temp=None,
             kernel_path: int = sbx.PATH_AUTO, device: int = 0,
             solver: str = "tf_jacobi", legacy_hvac_coordinates=None, **env_kwargs) -> sbx.Environment:
  plans = plans or sc.compiled()
  schedule = sbx.SetpointSchedule(6, 19, (294, 297), (289, 298), time_zone=sc.schedule_tz)
  weather = weather or sbx.WeatherController(sc.weather_low, sc.weather_high,
                                             convection_coefficient=sc.convection_coefficient)
  hvac_args = dict(
      air_handler=sbx.AirHandler(0.3, 285, 298, 10000.0, 0.9, max_air_flow_rate=8.67),
      boiler=sbx.Boiler(360.0, 6.0, 0.98, heating_rate=0.5, cooling_rate=0.1),
      schedule=schedule, vav_max_air_flow_rate=0.035, vav_reheat_max_water_flow_rate=0.03)
  if legacy_hvac_coordinates is not None:       # the deprecated Hvac (simulator/hvac.py:35-123)
    hvac = sbx.Hvac(zone_coordinates=legacy_hvac_coordinates, **hvac_args)
  else:
    hvac = sbx.FloorPlanBasedHvac(**hvac_args)
  if sc.occupancy == "step":
    occ = sbx.StepFunctionOccupancy(pd.Timedelta(9, unit="h"), pd.Timedelta(17, unit="h"),
                                    1.0, 0.1)
  else:
    occ = sbx.ConstantOccupancy(0.7)
  building = sbx.SimulatorBuilding(
      plans, hvac, weather, occ, n_envs=n_envs, time_step_sec=sc.time_step_sec,
      convergence_threshold=sc.convergence_threshold, iteration_limit=sc.iteration_limit,
      start_timestamp=sc.start_timestamp, floor_height_cm=sc.floor_height_cm,
      initial_temp=sc.initial_temp if initial_temp is None else initial_temp,
      reset_temp_values=sc.reset_temp_values,
      convection_simulator=(sbx.StochasticConvectionSimulator(*sc.convection)
                            if sc.convection else None), solver=solver)
  if sc.reward == "energy_carbon":
    reward = sbx.SetpointEnergyCarbonRewardFunction(
        300.0, 0.5, 4.3, sbx.ElectricityEnergyCost(), sbx.NaturalGasEnergyCost(),
        1.5, 2.0, 0.3, 25.0, 400.0)
  else:
    reward = sbx.SetpointEnergyCarbonRegretFunction(
        300.0, 100.0, 160000, 400000, 0.5, 4.3, sbx.ElectricityEnergyCost(),
        sbx.NaturalGasEnergyCost(), 0.2, 0.4, 0.4)
  return sbx.Environment(
      building, reward, sbx.StandardScoreObservationNormalizer(NORMALIZATION),
      sbx.ActionConfig({
          "supply_water_setpoint": sbx.BoundedActionNormalizer(310, 355.0),
          "supply_air_heating_temperature_setpoint": sbx.BoundedActionNormalizer(285, 300.0)}),
      discount_factor=sc.discount, num_days_in_episode=sc.num_days,
      occupancy_normalization_constant=sc.occupancy_norm,
      observation_histogram_reducer=sbx.HistogramReducer(HISTOGRAM) if sc.histogram else None,
      device=device, kernel_path=kernel_path, **env_kwargs)


def small_plan(h=24, w=34) -> np.ndarray:
  plan = np.full((h, w), 2, dtype=np.int64)
  plan[2:h - 2, 2:w - 2] = 1
  plan[3:h - 3, 3:w - 3] = 0
  plan[h // 2, 3:w - 3] = 1
  plan[3:h - 3, w // 2] = 1
  return plan
