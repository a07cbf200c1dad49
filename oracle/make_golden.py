"""Generates tests/golden/*.npz by running the UNMODIFIED reference (via
oracle/refshim.py) in a container where /root/reference is mounted.

  python -m oracle.make_golden            # rewrites every fixture

TEST INFRASTRUCTURE ONLY.  The fixtures are what travels to the GPU box (the
reference tree does not); tests/test_oracle_golden.py pins the oracle to them and
tests/test_gpu_parity.py pins the CUDA path to the oracle and to them.

Fixtures
  sb1_calibrated.npz   the calibrated building's inputs (data, not source):
                       floor plan {0,1,2} as uint8, reset temperatures, the
                       2023-07-05..08 slice of the Moffett Field weather CSV, plus
                       reference-derived structure counts for the static compiler.
  ref_gs_golden.npz    arrays of the reference's own golden scenario
                       (simulator_flexible_floor_plan_test.py:1275-1312, expected
                       return water temperature 301.895482) and the reference's result.
  ref_env_tf.npz       40-step rollout of the unmodified Environment + SimulatorBuilding
  ref_env_gs.npz       + regret reward on a 24x34 plan, TF-Jacobi (TFSimulator on the
                       NumPy provider of the TF primitives) and Gauss-Seidel solvers:
                       actions, observations, rewards, zone temperatures, final field.
  ref_env_hist.npz     same with the histogram observation reducer (D = 53 layout).
  ref_tf_calibrated.npz  one reset + 2 FD steps of TFSimulator on the calibrated plan.
  ref_env_ecr.npz      ref_env_tf's scenario with SetpointEnergyCarbonRewardFunction
                       (reward/setpoint_energy_carbon_reward.py) instead of the regret reward.
  ref_env_conv.npz     ref_env_tf's scenario with the shipped StochasticConvectionSimulator
                       (p = 1, distance = 5, seed = 5; sim_config.gin:37-39) in the building.
"""

from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
SB1 = "/root/reference/smart_control/configs/resources/sb1"

NORMALIZATION = {
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
HISTOGRAM = (
    ("zone_air_temperature_sensor", tuple(np.linspace(-1.03, 2.57, 19))),
    ("supply_air_damper_percentage_command", (0.0, 0.2, 0.4, 0.6, 0.8, 1.0)),
    ("supply_air_flowrate_setpoint", (0., 0.05, .1, .2, .3, .4, .5, .7, .9)),
)


def small_plan(h=24, w=34):
  plan = np.full((h, w), 2, dtype=np.int64)
  plan[2:h - 2, 2:w - 2] = 1
  plan[3:h - 3, 3:w - 3] = 0
  plan[h // 2, 3:w - 3] = 1
  plan[3:h - 3, w // 2] = 1
  return plan


def _ref_modules():
  from oracle import refshim
  refshim.install()
  import importlib
  names = dict(
      tfs="simulator.tf_simulator", building="simulator.building",
      hvac="simulator.hvac_floorplan_based", ah="simulator.air_handler",
      bl="simulator.boiler", ss="simulator.setpoint_schedule",
      wc="simulator.weather_controller", sb="simulator.simulator_building",
      occ="simulator.step_function_occupancy", sffp="simulator.simulator_flexible_floor_plan",
      conv="simulator.stochastic_convection_simulator",
      regret="reward.setpoint_energy_carbon_regret", ecr="reward.setpoint_energy_carbon_reward", elec="reward.electricity_energy_cost",
      gas="reward.natural_gas_energy_cost", env="environment.environment",
      onorm="utils.observation_normalizer", anorm="utils.bounded_action_normalizer",
      npb="proto.smart_control_normalization_pb2", hist="utils.histogram_reducer")
  return {k: importlib.import_module("smart_buildings.smart_control." + v)
          for k, v in names.items()}


def build_reference_env(m, plan, solver, histogram=False, reward="regret", convection=None,
                        occupancy=None):
  """The scenario of tests/scenarios.py:Scenario() built from reference classes."""
  b = m["building"].FloorPlanBasedBuilding(
      cv_size_cm=20.0, floor_height_cm=300.0, initial_temp=292.0,
      inside_air_properties=m["building"].MaterialProperties(50.0, 700.0, 1.0),
      inside_wall_properties=m["building"].MaterialProperties(2.0, 1000.0, 1800.0),
      building_exterior_properties=m["building"].MaterialProperties(0.05, 1000.0, 3000.0),
      floor_plan=plan, zone_map=plan.copy(), buffer_from_walls=2,
      convection_simulator=(m["conv"].StochasticConvectionSimulator(*convection)
                            if convection else None))
  sched = m["ss"].SetpointSchedule(6, 19, (294, 297), (289, 298))
  weather = m["wc"].WeatherController(275.0, 290.0, convection_coefficient=60.0)
  boiler = m["bl"].Boiler(360.0, 6.0, 0.98, heating_rate=0.5, cooling_rate=0.1,
                          device_id="boiler_id_x")
  ahu = m["ah"].AirHandler(0.3, 285, 298, 10000.0, 0.9, max_air_flow_rate=8.67,
                           device_id="air_handler_id_x", sim_weather_controller=weather)
  hv = m["hvac"].FloorPlanBasedHvac(air_handler=ahu, boiler=boiler, schedule=sched,
                                    vav_max_air_flow_rate=0.035,
                                    vav_reheat_max_water_flow_rate=0.03)
  start = pd.Timestamp("2023-07-06 05:00:00")
  cls = m["tfs"].TFSimulator if solver == "tf" else m["sffp"].SimulatorFlexibleGeometries
  sim = cls(b, hv, weather, 300.0, 0.1, 100, 30, start)
  occ = m["occ"].StepFunctionOccupancy(pd.Timedelta(9, unit="h"), pd.Timedelta(17, unit="h"),
                                       1.0, 0.1)
  sb = m["sb"].SimulatorBuilding(sim, occupancy or occ)
  rf = m["regret"].SetpointEnergyCarbonRegretFunction(
      300.0, 100.0, 160000, 400000, 0.5, 4.3, m["elec"].ElectricityEnergyCost(),
      m["gas"].NaturalGasEnergyCost(), 0.2, 0.4, 0.4)
  if reward == "energy_carbon":      # the parameters of tests/scenarios.py
    rf = m["ecr"].SetpointEnergyCarbonRewardFunction(
        300.0, 0.5, 4.3, m["elec"].ElectricityEnergyCost(), m["gas"].NaturalGasEnergyCost(),
        1.5, 2.0, 0.3, 25.0, 400.0)
  norm = {k: m["npb"].ContinuousVariableInfo(id=k, sample_mean=mu, sample_variance=var)
          for k, (mu, var) in NORMALIZATION.items()}
  on = m["onorm"].StandardScoreObservationNormalizer(norm)
  ac = m["env"].ActionConfig({
      "supply_water_setpoint": m["anorm"].BoundedActionNormalizer(310, 355.0),
      "supply_air_heating_temperature_setpoint": m["anorm"].BoundedActionNormalizer(285, 300.0)})
  hr = None
  if histogram:
    hr = m["hist"].HistogramReducer.__new__(m["hist"].HistogramReducer)
    # the constructor only needs a reader for expand(); the step path uses the bins
    hr._normalize_reduce = True
    hr._histogram_parameters = {k: np.array(v) for k, v in HISTOGRAM}
    hr._histogram_assignments = {}
  env = m["env"].Environment(sb, rf, on, ac, discount_factor=0.9, num_days_in_episode=1,
                             occupancy_normalization_constant=3.0,
                             observation_histogram_reducer=hr, time_zone="UTC")
  return env, b


def make_env_rollout(m, solver, histogram, n_steps, seed, reward="regret", convection=None):
  plan = small_plan()
  env, b = build_reference_env(m, plan, solver, histogram, reward, convection)
  rng = np.random.default_rng(seed)
  ts = env.reset()
  obs, rew, stype, disc, zts, acts, blr, ahu = [ts.observation], [0.0], [0], [1.0], [], [], [], []
  zts.append([float(v) for v in b.get_zone_average_temps().values()])
  for _ in range(n_steps):
    a = rng.uniform(-1, 1, 2).astype(np.float32)
    ts = env.step(a)
    acts.append(a)
    obs.append(ts.observation)
    rew.append(float(ts.reward))
    stype.append(int(ts.step_type))
    disc.append(float(ts.discount))
    zts.append([float(v) for v in b.get_zone_average_temps().values()])
    ri = env.building.reward_info
    bi = list(ri.boiler_reward_infos.values())[0]
    ai = list(ri.air_handler_reward_infos.values())[0]
    blr.append([bi.natural_gas_heating_energy_rate, bi.pump_electrical_energy_rate])
    ahu.append([ai.blower_electrical_energy_rate, ai.air_conditioning_electrical_energy_rate])
  return dict(floor_plan=plan.astype(np.uint8), actions=np.array(acts),
              observations=np.array(obs), rewards=np.array(rew, dtype=np.float64),
              step_types=np.array(stype), discounts=np.array(disc),
              zone_temps=np.array(zts), boiler_rates=np.array(blr), ahu_rates=np.array(ahu),
              final_temp=np.asarray(b.temp, dtype=np.float64),
              field_names=np.array(env._field_names))


def make_occupancy_golden(m, n_steps=200):
  """The shipped RandomizedArrivalDepartureOccupancy (sim_config.gin:185-196, two
  occupants per zone here) under the unmodified reference Environment: every
  average_zone_occupancy call the Environment makes is logged in order, so the
  host generator of sbsim_b200.exogenous can be checked call for call."""
  import importlib
  rado = importlib.import_module(
      "smart_buildings.smart_control.simulator.randomized_arrival_departure_occupancy")
  occ = rado.RandomizedArrivalDepartureOccupancy(2, 7, 12, 13, 18, 300, seed=17321)
  log = []
  inner = occ.average_zone_occupancy

  def logged(zone_id, start, end):
    v = inner(zone_id, start, end)
    log.append((zone_id, start, end, v))
    return v

  occ.average_zone_occupancy = logged
  env, b = build_reference_env(m, small_plan(), "tf", occupancy=occ)
  rng = np.random.default_rng(7)
  ts = env.reset()
  n_occ = [float(ts.observation[-1])]
  rewards = [0.0]
  for _ in range(n_steps):
    ts = env.step(rng.uniform(-1, 1, 2).astype(np.float32))
    n_occ.append(float(ts.observation[-1]))
    rewards.append(float(ts.reward))
  epoch = pd.Timestamp("1970-01-01")
  zones = sorted({z for z, _, _, _ in log})
  return dict(call_zone=np.array([zones.index(z) for z, _, _, _ in log]),
              call_start_sec=np.array([(s - epoch).total_seconds() for _, s, _, _ in log]),
              call_end_sec=np.array([(e - epoch).total_seconds() for _, _, e, _ in log]),
              call_value=np.array([v for _, _, _, v in log]), zone_ids=np.array(zones),
              num_occupants_feature=np.array(n_occ), rewards=np.array(rewards),
              start="2023-07-06 05:00:00", n_steps=n_steps)


def make_gs_golden(m):
  """simulator_flexible_floor_plan_test.py:1275-1312 through the reference's own
  test fixture methods."""
  import importlib
  t = importlib.import_module(
      "smart_buildings.smart_control.simulator.simulator_flexible_floor_plan_test")
  case = t.FlexibleFloorplanSimulatorTest()
  building = case._create_scenario_building(initial_temp=200.0, match_old_diffusers=True)
  hvac = case._create_scenario_hvac(zone_identifier=building._room_dict.keys())
  weather = m["wc"].WeatherController(296.0, 296.0)
  sim = m["sffp"].SimulatorFlexibleGeometries(building, hvac, weather, 300.0, 0.1, 100, 10,
                                              pd.Timestamp("12-21-2012"))
  rooms = [k for k in building._room_dict if k.startswith("room")]
  room_rows = [np.array([c[0] for c in building._room_dict[k]]) for k in rooms]
  room_cols = [np.array([c[1] for c in building._room_dict[k]]) for k in rooms]
  ah, bl = hvac.air_handler, hvac.boiler
  vav = next(iter(hvac.vavs.values()))
  sched = vav.thermostat.get_setpoint_schedule()
  sim.step_sim()
  return dict(
      exterior_space=(building._exterior_space == -1), conductivity=building.conductivity,
      heat_capacity=building.heat_capacity, density=building.density,
      diffusers=building.diffusers, room_names=np.array(rooms),
      room_sizes=np.array([len(r) for r in room_rows]),
      room_rows=np.concatenate(room_rows), room_cols=np.concatenate(room_cols),
      cv_size_cm=building.cv_size_cm, floor_height_cm=building.floor_height_cm,
      hvac_params=np.array([ah._init_recirculation, ah._init_heating_air_temp_setpoint,
                            ah._init_cooling_air_temp_setpoint,
                            ah._init_fan_differential_pressure, ah._init_fan_efficiency,
                            ah._init_max_air_flow_rate, bl._init_reheat_water_setpoint,
                            bl._init_water_pump_differential_head,
                            bl._init_water_pump_efficiency, vav._init_max_air_flow_rate,
                            vav._init_reheat_max_water_flow_rate], dtype=np.float64),
      schedule=np.array([sched.morning_start_hour, sched.evening_start_hour,
                         *sched.comfort_temp_window, *sched.eco_temp_window], dtype=np.float64),
      expected_return_water_temperature=301.895482,
      reference_return_water_temperature=bl.return_water_temperature_sensor,
      reference_temp_after_step=np.asarray(building.temp, dtype=np.float64))


def make_sb1(m):
  plan = np.load(os.path.join(SB1, "double_resolution_zone_1_2.npy"))
  reset = np.load(os.path.join(SB1, "reset_temps.npy"))
  df = pd.read_csv(os.path.join(SB1, "local_weather_moffett_field_20230701_20231122.csv"))
  times = np.array([(pd.Timestamp(t, tz="UTC") - pd.Timestamp("1970-01-01", tz="UTC"))
                    .total_seconds() for t in df["Time"]])
  lo = (pd.Timestamp("2023-07-05", tz="UTC") - pd.Timestamp("1970-01-01", tz="UTC")).total_seconds()
  hi = (pd.Timestamp("2023-07-09", tz="UTC") - pd.Timestamp("1970-01-01", tz="UTC")).total_seconds()
  sel = (times >= lo) & (times <= hi)
  b = m["building"].FloorPlanBasedBuilding(
      cv_size_cm=10.0, floor_height_cm=300.0, initial_temp=294.0,
      inside_air_properties=m["building"].MaterialProperties(50.0, 700.0, 1.0),
      inside_wall_properties=m["building"].MaterialProperties(50.0, 1.0, 700.0),
      building_exterior_properties=m["building"].MaterialProperties(0.05, 700.0, 1.0),
      floor_plan=plan, zone_map=plan.copy(), buffer_from_walls=3, reset_temp_values=reset)
  rooms = [k for k in b._room_dict if k.startswith("room")]
  ln = np.array([[len(x) for x in row] for row in b.neighbors])
  sb1 = dict(floor_plan=plan.astype(np.uint8), reset_temps=reset,
             weather_time_sec=times[sel], weather_temp_f=np.asarray(df["TempF"])[sel],
             room_names=np.array(rooms),
             room_sizes=np.array([len(b._room_dict[k]) for k in rooms]),
             n_diffusers=int((b.diffusers > 0).sum()),
             neighbor_count_hist=np.bincount(ln.ravel(), minlength=5),
             material_checksum=np.array([b.conductivity.sum(), b.heat_capacity.sum(),
                                         b.density.sum(), b.diffusers.sum()]))
  # two FD steps of the unmodified TFSimulator (NumPy provider) from the reset field
  hv = m["hvac"].FloorPlanBasedHvac(
      air_handler=m["ah"].AirHandler(0.3, 285, 298, 10000.0, 0.9),
      boiler=m["bl"].Boiler(360.0, 6.0, 0.98, heating_rate=0.5, cooling_rate=0.1),
      schedule=m["ss"].SetpointSchedule(6, 19, (294, 297), (289, 298)),
      vav_max_air_flow_rate=0.035, vav_reheat_max_water_flow_rate=0.03)
  weather = m["wc"].WeatherController(283.0, 296.0, convection_coefficient=100.0)
  sim = m["tfs"].TFSimulator(b, hv, weather, 300.0, 0.1, 100, 30, pd.Timestamp("2023-07-06 07:00:00"))
  out = {}
  rng = np.random.default_rng(0)
  qz = rng.uniform(-200, 3000, len(rooms))
  for zi, k in enumerate(rooms):
    b.apply_thermal_power_zone(k, qz[zi])
  ambients = [288.5, 289.25]
  zone_means = []
  for amb in ambients:
    sim.finite_differences_timestep(ambient_temperature=amb, convection_coefficient=100.0)
    zone_means.append([float(v) for v in b.get_zone_average_temps().values()])
  t = np.asarray(b.temp, dtype=np.float32)
  out = dict(q_zone=qz, ambients=np.array(ambients), zone_means=np.array(zone_means),
             temp_sum=float(t.astype(np.float64).sum()), temp_mean_f32=float(b.temp.mean()),
             temp_sample=t[::31, ::37].copy(), temp_rows=t[[100, 372, 600], :].copy())
  return sb1, out


def main():
  warnings.filterwarnings("ignore")
  os.makedirs(OUT, exist_ok=True)
  m = _ref_modules()
  np.savez_compressed(os.path.join(OUT, "ref_env_tf.npz"), **make_env_rollout(m, "tf", False, 40, 0))
  np.savez_compressed(os.path.join(OUT, "ref_env_gs.npz"), **make_env_rollout(m, "gs", False, 25, 1))
  np.savez_compressed(os.path.join(OUT, "ref_env_hist.npz"), **make_env_rollout(m, "tf", True, 40, 2))
  np.savez_compressed(os.path.join(OUT, "ref_env_ecr.npz"),
                      **make_env_rollout(m, "tf", False, 60, 3, reward="energy_carbon"))
  np.savez_compressed(os.path.join(OUT, "ref_env_conv.npz"),
                      **make_env_rollout(m, "tf", False, 30, 4, convection=(1.0, 5, 5)))
  np.savez_compressed(os.path.join(OUT, "ref_occupancy.npz"), **make_occupancy_golden(m))
  np.savez_compressed(os.path.join(OUT, "ref_gs_golden.npz"), **make_gs_golden(m))
  sb1, cal = make_sb1(m)
  np.savez_compressed(os.path.join(OUT, "sb1_calibrated.npz"), **sb1)
  np.savez_compressed(os.path.join(OUT, "ref_tf_calibrated.npz"), **cal)
  for f in sorted(os.listdir(OUT)):
    print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
  sys.path.insert(0, ROOT)
  main()
