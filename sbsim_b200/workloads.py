"""Synthetic workloads of BASELINE.json (configs 2-5), built from seeds.

  randomized(...)   configs 3/4/5: B buildings on 64x96 grids with randomised
                    floor plans, wall/exterior materials, CV size, initial
                    temperature, sinusoid weather and convection coefficient
                    ("Config 3").
  make_calibrated_env(...)
                    configs 1/2: B copies of the calibrated sb1 building
                    (744x1004 CVs, 126 zones; sim_config.gin:160-196): floor plan,
                    reset temperatures and the Moffett Field weather rows come from
                    a committed fixture of the reference's resource files
                    (tests/golden/sb1_calibrated.npz, written by
                    oracle/make_golden.py), the devices, schedule, occupancy model,
                    reward and normalisation constants from the gin file.
"""

from __future__ import annotations

import dataclasses
from typing import List, Optional, Tuple

import numpy as np
import pandas as pd

import sbsim_b200 as sbx
from sbsim_b200 import exogenous, floorplan

DEFAULT_START = "2023-07-06 07:00:00"   # sim_config.gin:164 (tz-naive for the sinusoid weather)

# sim_config.gin:253-580, entries whose measurement names the simulated devices expose
NORMALIZATION = {
    "cooling_request_count": (100.0, 25.0),
    "differential_pressure_setpoint": (83810.269540, 14889040.603647),
    "outside_air_flowrate_sensor": (3.701930, 20.300565),
    "outside_air_temperature_sensor": (291.244931, 12.904175),
    "supply_air_cooling_temperature_setpoint": (289.329414, 3.186769),
    "supply_air_heating_temperature_setpoint": (289.329414, 3.186769),
    "supply_fan_speed_percentage_command": (26.543748, 575.094979),
    "supply_water_setpoint": (320.261985, 240.195517),
    "supply_water_temperature_sensor": (321.520315, 658.413066),
}
HISTOGRAM = (  # sim_config.gin:586-590
    ("zone_air_temperature_sensor", (285., 286., 287., 288, 289., 290., 291., 292., 293., 294.,
                                     295., 296., 297., 298., 299., 300., 301, 302, 303)),
    ("supply_air_damper_percentage_command", (0.0, 0.2, 0.4, 0.6, 0.8, 1.0)),
    ("supply_air_flowrate_setpoint", (0., 0.05, .1, .2, .3, .4, .5, .7, .9)),
)


def calibrated_hvac(schedule_tz="UTC") -> sbx.FloorPlanBasedHvac:
  """sim_config.gin:101-157."""
  return sbx.FloorPlanBasedHvac(
      air_handler=sbx.AirHandler(recirculation=0.3, heating_air_temp_setpoint=285.0,
                                 cooling_air_temp_setpoint=298.0,
                                 fan_differential_pressure=10000.0, fan_efficiency=0.9,
                                 max_air_flow_rate=8.67),
      boiler=sbx.Boiler(reheat_water_setpoint=360.0, water_pump_differential_head=6.0,
                        water_pump_efficiency=0.98, heating_rate=0.5, cooling_rate=0.1),
      schedule=sbx.SetpointSchedule(6, 19, (294, 297), (289, 298), time_zone=schedule_tz),
      vav_max_air_flow_rate=0.035, vav_reheat_max_water_flow_rate=0.03)


def calibrated_reward() -> sbx.SetpointEnergyCarbonRegretFunction:
  """sim_config.gin:201-225."""
  return sbx.SetpointEnergyCarbonRegretFunction(
      max_productivity_personhour_usd=300.0, min_productivity_personhour_usd=100.0,
      max_electricity_rate=160000, max_natural_gas_rate=400000,
      productivity_midpoint_delta=0.5, productivity_decay_stiffness=4.3,
      electricity_energy_cost=sbx.ElectricityEnergyCost(),
      natural_gas_energy_cost=sbx.NaturalGasEnergyCost(),
      productivity_weight=0.2, energy_cost_weight=0.4, carbon_emission_weight=0.4)


def calibrated_action_config() -> sbx.ActionConfig:
  """sim_config.gin:229-244."""
  return sbx.ActionConfig({
      "supply_water_setpoint": sbx.BoundedActionNormalizer(310, 355.0),
      "supply_air_heating_temperature_setpoint": sbx.BoundedActionNormalizer(285, 300.0)})


@dataclasses.dataclass
class RandomizedWorkload:
  plans: List[floorplan.CompiledPlan]      # one per env (layouts tiled from a pool)
  weather_low: np.ndarray
  weather_high: np.ndarray
  convection: np.ndarray
  initial_temp: np.ndarray
  n_layouts: int


def randomized(n_envs: int, seed: int = 2024, n_layouts: int = 1024,
               spec: floorplan.RandomPlanSpec = floorplan.RandomPlanSpec()) -> RandomizedWorkload:
  """Config 3 generator.  `n_layouts` distinct floor-plan layouts are compiled and
  tiled over the batch (the host-side compile is ~3 ms per layout); every env still
  gets its OWN descriptor copy, materials, CV size, initial temperature and weather."""
  rng = np.random.default_rng(seed)
  n_layouts = min(n_layouts, n_envs)
  layouts = [floorplan.random_floor_plan(rng, spec) for _ in range(n_layouts)]
  base = []
  for lay in layouts:
    air, wall, ext = floorplan.random_materials(rng)
    cv = float(rng.choice([10.0, 20.0]))
    base.append(floorplan.compile_plan(lay, None, cv_size_cm=cv, inside_air=air,
                                       inside_wall=wall, building_exterior=ext))
  plans = []
  for b in range(n_envs):
    src = base[b % n_layouts]
    if b < n_layouts:
      plans.append(src)
    else:  # same layout, fresh materials
      air, wall, ext = floorplan.random_materials(rng)
      mat = np.array([[m.conductivity, m.heat_capacity, m.density]
                      for m in (air, wall, ext)])
      plans.append(dataclasses.replace(src, material=mat))
  low = rng.uniform(268.0, 288.0, n_envs)
  high = low + rng.uniform(5.0, 15.0, n_envs)
  return RandomizedWorkload(plans=plans, weather_low=low, weather_high=high,
                            convection=rng.uniform(12.0, 100.0, n_envs),
                            initial_temp=rng.uniform(290.0, 296.0, n_envs).astype(np.float32),
                            n_layouts=n_layouts)


def make_randomized_env(n_envs: int, seed: int = 2024, episode_steps: int = 288,
                        n_layouts: int = 1024, histogram: bool = False, device: int = 0,
                        kernel_path: int = sbx.PATH_AUTO, numpy_zone_means: bool = False, start: str = DEFAULT_START,
                        workload: Optional[RandomizedWorkload] = None,
                        convergence_threshold: float = 0.1, iteration_limit: int = 100
                        ) -> Tuple[sbx.Environment, RandomizedWorkload]:
  wl = workload or randomized(n_envs, seed, n_layouts)
  weather = sbx.BatchedWeather(
      lambda ts: exogenous.sinusoid_weather_tables(ts, wl.weather_low, wl.weather_high),
      wl.convection)
  occ = sbx.StepFunctionOccupancy(pd.Timedelta(9, unit="h"), pd.Timedelta(17, unit="h"),
                                  1.0, 0.1)
  building = sbx.SimulatorBuilding(
      wl.plans, calibrated_hvac(), weather, occ, n_envs=n_envs, time_step_sec=300.0,
      convergence_threshold=convergence_threshold, iteration_limit=iteration_limit,
      start_timestamp=pd.Timestamp(start),
      floor_height_cm=300.0, initial_temp=wl.initial_temp)
  env = sbx.Environment(
      building, calibrated_reward(), sbx.StandardScoreObservationNormalizer(NORMALIZATION),
      calibrated_action_config(), discount_factor=0.9,
      num_days_in_episode=(episode_steps + 0.5) * 300.0 / 86400.0,
      occupancy_normalization_constant=125.0,
      observation_histogram_reducer=sbx.HistogramReducer(HISTOGRAM) if histogram else None,
      device=device, kernel_path=kernel_path, numpy_zone_means=numpy_zone_means)
  return env, wl


CALIBRATED_START = "2023-07-06 07:00:00+00:00"   # sim_config.gin:164
CALIBRATED_TZ = "US/Pacific"                       # sim_config.gin:20
CALIBRATED_CONVECTION = (1.0, 5, 5)                # StochasticConvectionSimulator p, distance, seed (:37-39)


@dataclasses.dataclass
class CalibratedBuilding:
  plan: floorplan.CompiledPlan
  reset_temps: np.ndarray          # fp64 [744, 1004], reset_temps.npy
  weather_time_sec: np.ndarray     # Moffett Field CSV rows covering the episode
  weather_temp_f: np.ndarray


def load_calibrated(fixture_path: str) -> CalibratedBuilding:
  """The calibrated building of sim_config.gin:42-99 from the committed fixture."""
  g = np.load(fixture_path)
  cp = floorplan.compile_plan(
      g["floor_plan"].astype(np.int64), None, cv_size_cm=10.0,
      inside_air=floorplan.MaterialProperties(50.0, 700.0, 1.0),
      inside_wall=floorplan.MaterialProperties(50.0, 1.0, 700.0),
      building_exterior=floorplan.MaterialProperties(0.05, 700.0, 1.0),
      buffer_from_walls=3)
  return CalibratedBuilding(cp, np.asarray(g["reset_temps"], dtype=np.float64),
                            np.asarray(g["weather_time_sec"], dtype=np.float64),
                            np.asarray(g["weather_temp_f"], dtype=np.float64))


def calibrated_occupancy(kind: str = "randomized"):
  """sim_config.gin:179-196 ("randomized", the shipped model) or a constant 0.7
  occupants per zone ("const", the deterministic stand-in of the parity tests: the
  reference's StepFunctionOccupancy raises on the tz-aware start timestamp)."""
  if kind == "randomized":
    return exogenous.RandomizedArrivalDepartureOccupancy(
        1, 7, 12, 13, 18, time_step_sec=300, seed=17321, time_zone=CALIBRATED_TZ)
  if kind == "const":
    return exogenous.ConstantOccupancy(0.7)
  raise ValueError(kind)


def make_calibrated_env(cal: CalibratedBuilding, n_envs: int, episode_steps: int = 288,
                        histogram: bool = True, device: int = 0,
                        kernel_path: int = sbx.PATH_AUTO, numpy_zone_means: bool = False, occupancy: str = "randomized",
                        convection=None) -> sbx.Environment:
  """Configs 1/2: n_envs copies of the calibrated building (one shared descriptor),
  Moffett replay weather, reset_temps.npy, schedule in US/Pacific."""
  weather = sbx.ReplayWeatherController(times_utc_sec=cal.weather_time_sec,
                                        temps_f=cal.weather_temp_f,
                                        convection_coefficient=100.0)
  building = sbx.SimulatorBuilding(
      cal.plan, calibrated_hvac(CALIBRATED_TZ), weather, calibrated_occupancy(occupancy),
      n_envs=n_envs, time_step_sec=300.0, convergence_threshold=0.1, iteration_limit=100,
      start_timestamp=pd.Timestamp(CALIBRATED_START), floor_height_cm=300.0,
      initial_temp=294.0, reset_temp_values=cal.reset_temps.astype(np.float32),
      convection_simulator=convection)
  return sbx.Environment(
      building, calibrated_reward(), sbx.StandardScoreObservationNormalizer(NORMALIZATION),
      calibrated_action_config(), discount_factor=0.9,
      num_days_in_episode=(episode_steps + 0.5) * 300.0 / 86400.0,
      occupancy_normalization_constant=125.0,
      observation_histogram_reducer=sbx.HistogramReducer(HISTOGRAM) if histogram else None,
      time_zone=CALIBRATED_TZ, device=device, kernel_path=kernel_path, numpy_zone_means=numpy_zone_means)


def synthetic_office_plan(height: int = 744, width: int = 1004, rooms_y: int = 9,
                          rooms_x: int = 14, seed: int = 7) -> np.ndarray:
  """A large office-like plan of the calibrated building's size class (126 zones)."""
  spec = floorplan.RandomPlanSpec(height=height, width=width, rooms_y=(rooms_y,),
                                  rooms_x=(rooms_x,), min_room=12)
  return floorplan.random_floor_plan(np.random.default_rng(seed), spec)


def make_shared_plan_env(plan: floorplan.CompiledPlan, n_envs: int, episode_steps: int = 288,
                         reset_temp_values=None, initial_temp=294.0, weather=None,
                         histogram: bool = True, device: int = 0,
                         kernel_path: int = sbx.PATH_AUTO, numpy_zone_means: bool = False, start: str = DEFAULT_START,
                         solver: str = "tf_jacobi", iteration_limit: int = 100) -> sbx.Environment:
  """Config 2: n_envs copies of one plan (descriptor shared, L2-resident)."""
  weather = weather or sbx.WeatherController(283.0, 296.0, convection_coefficient=100.0)
  occ = sbx.StepFunctionOccupancy(pd.Timedelta(9, unit="h"), pd.Timedelta(17, unit="h"),
                                  1.0, 0.1)
  building = sbx.SimulatorBuilding(
      plan, calibrated_hvac(), weather, occ, n_envs=n_envs, time_step_sec=300.0,
      convergence_threshold=0.1, iteration_limit=iteration_limit, start_timestamp=pd.Timestamp(start),
      floor_height_cm=300.0, initial_temp=initial_temp, reset_temp_values=reset_temp_values, solver=solver)
  return sbx.Environment(
      building, calibrated_reward(), sbx.StandardScoreObservationNormalizer(NORMALIZATION),
      calibrated_action_config(), discount_factor=0.9,
      num_days_in_episode=(episode_steps + 0.5) * 300.0 / 86400.0,
      occupancy_normalization_constant=125.0,
      observation_histogram_reducer=sbx.HistogramReducer(HISTOGRAM) if histogram else None,
      device=device, kernel_path=kernel_path, numpy_zone_means=numpy_zone_means)
