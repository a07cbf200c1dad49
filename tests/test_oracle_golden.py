"""Pins the oracle (oracle/) to outputs of the UNMODIFIED reference, committed as
fixtures under tests/golden/ by oracle/make_golden.py (the reference itself does
not exist on the GPU box)."""

import os

import numpy as np
import pandas as pd
import pytest

import scenarios as S
from oracle import env as oenv
from oracle import exogenous as oex
from oracle import hvac as ohvac
from oracle import reward as orew
from oracle import tf_jacobi
from sbsim_b200 import floorplan

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
  return np.load(os.path.join(GOLD, name), allow_pickle=False)


@pytest.mark.parametrize("name,solver,histogram,reward", [
    ("ref_env_tf.npz", "tf", False, "regret"), ("ref_env_hist.npz", "tf", True, "regret"),
    ("ref_env_gs.npz", "gs", False, "regret"), ("ref_env_ecr.npz", "tf", False, "energy_carbon"),
    ("ref_env_conv.npz", "tf", False, "regret")])
def test_oracle_env_matches_reference_rollout(name, solver, histogram, reward):
  g = _load(name)
  sc = S.Scenario(floor_plan=g["floor_plan"].astype(np.int64), histogram=histogram,
                  reward=reward, convection=(1.0, 5, 5) if "conv" in name else None)
  o = S.make_oracle(sc, solver=solver)
  ts = o.reset()
  np.testing.assert_allclose(ts[3], g["observations"][0], rtol=1e-6, atol=1e-7)
  np.testing.assert_array_equal([float(t) for t in o.zone_average_temps()], g["zone_temps"][0])
  for i, a in enumerate(g["actions"]):
    ts = o.step(a)
    assert int(ts[0]) == int(g["step_types"][i + 1])
    np.testing.assert_allclose(ts[3], g["observations"][i + 1], rtol=1e-6, atol=1e-7,
                               err_msg=f"obs step {i}")
    np.testing.assert_allclose(float(ts[1]), g["rewards"][i + 1], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(float(ts[2]), g["discounts"][i + 1], rtol=1e-6)
    np.testing.assert_array_equal([float(t) for t in o.zone_average_temps()],
                                  g["zone_temps"][i + 1])
    info = o.info["reward_info"]
    np.testing.assert_allclose(
        [info.natural_gas_heating_energy_rate, info.pump_electrical_energy_rate],
        g["boiler_rates"][i], rtol=1e-6)
    np.testing.assert_allclose(
        [info.blower_electrical_energy_rate, info.air_conditioning_electrical_energy_rate],
        g["ahu_rates"][i], rtol=1e-6, atol=1e-6)
  # the temperature field itself: bit-identical to the reference's
  np.testing.assert_array_equal(np.asarray(o.temp, dtype=np.float64), g["final_temp"])


def test_oracle_gs_reproduces_reference_golden_return_water_temperature():
  """simulator_flexible_floor_plan_test.py:1275-1312: 301.895482 +- 1e-5."""
  g = _load("ref_gs_golden.npz")
  sizes = g["room_sizes"]
  offs = np.concatenate([[0], np.cumsum(sizes)])
  rooms = [(str(n), g["room_rows"][offs[i]:offs[i + 1]], g["room_cols"][offs[i]:offs[i + 1]])
           for i, n in enumerate(g["room_names"])]
  plan = tf_jacobi.OraclePlan(
      exterior_space=g["exterior_space"], conductivity=g["conductivity"],
      heat_capacity=g["heat_capacity"], density=g["density"], diffusers=g["diffusers"],
      rooms=rooms, cv_size_cm=float(g["cv_size_cm"]), floor_height_cm=float(g["floor_height_cm"]))
  hp = g["hvac_params"]
  sch = g["schedule"]
  cfg = oenv.OracleEnvConfig(
      plan=plan, start_timestamp=pd.Timestamp("12-21-2012"),
      weather=oex.WeatherController(296.0, 296.0),
      schedule=ohvac.SetpointSchedule(int(sch[0]), int(sch[1]), (sch[2], sch[3]), (sch[4], sch[5])),
      occupancy=oex.ConstantOccupancy(1.0),
      reward_function=orew.SetpointEnergyCarbonRegretFunction(
          300.0, 100.0, 160000, 400000, 0.5, 4.3, oex.ElectricityEnergyCost(),
          oex.NaturalGasEnergyCost(), 0.2, 0.4, 0.4),
      solver="gs", time_step_sec=300.0, convergence_threshold=0.1, iteration_limit=100,
      initial_temp=200.0, ahu_recirculation=hp[0], ahu_heating_setpoint=hp[1],
      ahu_cooling_setpoint=hp[2], ahu_fan_differential_pressure=hp[3],
      ahu_fan_efficiency=hp[4], ahu_max_air_flow_rate=hp[5], boiler_setpoint=hp[6],
      boiler_pump_head=hp[7], boiler_pump_efficiency=hp[8], boiler_heating_rate=0,
      boiler_cooling_rate=0, vav_max_air_flow_rate=hp[9],
      vav_reheat_max_water_flow_rate=hp[10])
  o = oenv.OracleEnvironment(cfg)
  o._setup_step_sim()          # Simulator.step_sim simulator.py:578-594
  o._execute_step_sim()
  got = o.boiler.return_water_temperature_sensor
  assert abs(got - float(g["expected_return_water_temperature"])) < 1e-5
  assert got == float(g["reference_return_water_temperature"])
  np.testing.assert_array_equal(o.temp, g["reference_temp_after_step"])


@pytest.fixture(scope="module")
def calibrated_plan():
  g = _load("sb1_calibrated.npz")
  cp = floorplan.compile_plan(
      g["floor_plan"].astype(np.int64), None, cv_size_cm=10.0,
      inside_air=floorplan.MaterialProperties(50.0, 700.0, 1.0),
      inside_wall=floorplan.MaterialProperties(50.0, 1.0, 700.0),     # sim_config.gin:78-81 (swapped)
      building_exterior=floorplan.MaterialProperties(0.05, 700.0, 1.0))
  return g, cp


def test_static_compiler_on_calibrated_plan(calibrated_plan):
  g, cp = calibrated_plan
  assert (cp.height, cp.width) == (744, 1004)
  assert cp.zone_names == [str(n) for n in g["room_names"]]
  np.testing.assert_array_equal(cp.zone_ncv, g["room_sizes"])
  assert int(cp.zone_ndiff.sum()) == int(g["n_diffusers"]) == 4263
  hist = np.bincount(tf_jacobi.neighbor_counts(cp.exterior_space).ravel(), minlength=5)
  np.testing.assert_array_equal(hist, g["neighbor_count_hist"])
  chk = g["material_checksum"]
  np.testing.assert_allclose([cp.dense_material(0).sum(), cp.dense_material(1).sum(),
                              cp.dense_material(2).sum(), cp.diffuser_weight.sum()], chk, rtol=1e-12)
  # 490172 interior / 6060 edge / 304 corner CVs, 126 zones
  cls = cp.cv_class
  assert int((cls == 1).sum()) == 490172
  assert int(((cls >= 2) & (cls <= 5)).sum()) == 6060
  assert int((cls >= 6).sum()) == 304
  assert cp.n_zones == 126


def test_oracle_tf_jacobi_on_calibrated_plan_matches_reference(calibrated_plan):
  """Two finite_differences_timestep calls of the unmodified TFSimulator (NumPy
  provider of the TF primitives) on the 744x1004 plan, from reset_temps.npy."""
  g, cp = calibrated_plan
  r = _load("ref_tf_calibrated.npz")
  jac = tf_jacobi.TFJacobi(S.oracle_plan(cp, 300.0), 300.0, 0.1, 100)
  q = np.zeros((cp.height, cp.width))
  for zi in range(cp.n_zones):
    m = (cp.zone_id == zi) & (cp.diffuser_weight > 0)
    q[m] = r["q_zone"][zi] * cp.diffuser_weight[m]
  temp = g["reset_temps"].copy()
  for i, amb in enumerate(r["ambients"]):
    temp, n, conv, _ = jac.fd_step(temp, q, float(amb), 100.0)
    zm = [float(np.mean(temp[cp.zone_indices(zi)])) for zi in range(cp.n_zones)]
    np.testing.assert_array_equal(zm, r["zone_means"][i])
  np.testing.assert_array_equal(temp[::31, ::37], r["temp_sample"])
  np.testing.assert_array_equal(temp[[100, 372, 600], :], r["temp_rows"])
  assert float(temp.astype(np.float64).sum()) == float(r["temp_sum"])


class _FlatEnergyCost:
  """TestEnergyCost of setpoint_energy_carbon_reward_test.py:232-254."""

  def __init__(self, usd_per_kwh, kg_per_kwh):
    self._p = usd_per_kwh / 3600.0 / 1000.0
    self._c = kg_per_kwh / 3600.0 / 1000.0

  def cost(self, start, end, rate):
    return self._p * rate * (end - start).total_seconds()

  def carbon(self, start, end, rate):
    return self._c * rate * (end - start).total_seconds()


@pytest.mark.parametrize("temp,occ,blower,ac,gas,pump,want", [
    # setpoint_energy_carbon_reward_test.py:30-112 (reward, productivity, elec cost, gas cost,
    # carbon emitted, carbon cost); the reference's info has two of every device
    (293.0, 10.0, 500.0, 2500.0, 20000.0, 250.0, (0.11660, 5000.0 / 6, 0.102916665, 0.1, 0.63208336, 0.12641667)),
    (293.0, 0.0, 500.0, 2500.0, 20000.0, 250.0, (-0.050065, 0.0, 0.102916665, 0.1, 0.63208336, 0.12641667)),
    (292.0, 10.0, 500.0, 2500.0, 20000.0, 250.0, (0.099212, 746.3906, 0.102916665, 0.1, 0.63208336, 0.12641667)),
    (299.0, 10.0, 500.0, 2500.0, 20000.0, 250.0, (-0.0326773, 86.942687, 0.102916665, 0.1, 0.63208336, 0.12641667)),
    (293.0, 10.0, 0.0, 0.0, 0.0, 0.0, (0.116666, 5000.0 / 6, 0.0, 0.0, 0.0, 0.0)),
])
def test_energy_carbon_reward_reference_known_answers(temp, occ, blower, ac, gas, pump, want):
  fn = orew.SetpointEnergyCarbonRewardFunction(
      500.0, 1.5, 4.3, _FlatEnergyCost(0.19, 0.01), _FlatEnergyCost(0.03, 0.188),
      1.0, 1.0, 0.2, 250.0, 500.0 * 10)
  f32 = orew.f32
  zone = orew.ZoneRewardInfo(f32(293.0), f32(297.0), f32(temp), f32(0.013), f32(0.012), f32(occ))
  info = orew.RewardInfo(
      pd.Timestamp("2021-05-03 12:13:00-05:00"), pd.Timestamp("2021-05-03 12:18:00-05:00"),
      [zone, zone], 2 * blower, 2 * ac, 2 * gas, 2 * pump)
  r = fn.compute_reward(info)
  got = (r.agent_reward_value, r.productivity_reward, r.electricity_energy_cost,
         r.natural_gas_energy_cost, r.carbon_emitted, r.carbon_cost)
  for g, w in zip(got, want):
    assert abs(g - w) < 0.5e-4, (got, want)     # assertAlmostEqual(..., 4)


def _legacy_cp(exact_materials=False):
  return floorplan.legacy_building(
      20.0, (20, 30), (3, 3), floorplan.MaterialProperties(50.0, 700.0, 1.0),
      floorplan.MaterialProperties(5.0, 800.0, 1800.0), floorplan.MaterialProperties(5.0, 800.0, 3000.0),
      exact_materials=exact_materials)


def test_legacy_rectangular_building_equals_reference_arrays():
  """the deprecated `Building(room_shape, building_shape)`
  (building.py:394-505) through legacy_building(exact_materials=False): materials, exterior
  space, rooms and the four diffusers per room equal the arrays of the reference's own
  floor-plan twin of it (the fixture: FloorPlanBasedBuilding with the OLD Building's diffusers
  copied over, simulator_flexible_floor_plan_test.py:467-471).  The old Building itself is
  compared live in tests/test_oracle_vs_reference.py."""
  g = _load("ref_gs_golden.npz")
  cp = _legacy_cp()
  assert (cp.height, cp.width) == g["exterior_space"].shape == (68, 98)
  np.testing.assert_array_equal(cp.exterior_space, g["exterior_space"])
  for k, name in enumerate(("conductivity", "heat_capacity", "density")):
    np.testing.assert_array_equal(cp.dense_material(k), g[name])
  np.testing.assert_array_equal(cp.diffuser_weight, g["diffusers"].astype(np.float64))
  assert cp.zone_names == [str(n) for n in g["room_names"]]
  np.testing.assert_array_equal(cp.zone_ncv, g["room_sizes"])
  assert list(cp.zone_ndiff) == [4] * 9


@pytest.mark.parametrize("exact_materials", [False, True])
def test_oracle_gs_golden_through_legacy_building(exact_materials):
  """The reference's golden return-water temperature 301.895482
  (simulator_test.py:955-987 on the old Building, simulator_flexible_floor_plan_test.py:1275-1312
  on its twin) from legacy_building() + the oracle's Gauss-Seidel step, for the old
  Building's exact materials and for the twin's."""
  g = _load("ref_gs_golden.npz")
  hp, sch = g["hvac_params"], g["schedule"]
  cfg = oenv.OracleEnvConfig(
      plan=S.oracle_plan(_legacy_cp(exact_materials), 300.0), start_timestamp=pd.Timestamp("12-21-2012"),
      weather=oex.WeatherController(296.0, 296.0),
      schedule=ohvac.SetpointSchedule(int(sch[0]), int(sch[1]), (sch[2], sch[3]), (sch[4], sch[5])),
      occupancy=oex.ConstantOccupancy(1.0),
      reward_function=orew.SetpointEnergyCarbonRegretFunction(
          300.0, 100.0, 160000, 400000, 0.5, 4.3, oex.ElectricityEnergyCost(),
          oex.NaturalGasEnergyCost(), 0.2, 0.4, 0.4),
      solver="gs", time_step_sec=300.0, convergence_threshold=0.1, iteration_limit=100,
      initial_temp=200.0, ahu_recirculation=hp[0], ahu_heating_setpoint=hp[1],
      ahu_cooling_setpoint=hp[2], ahu_fan_differential_pressure=hp[3],
      ahu_fan_efficiency=hp[4], ahu_max_air_flow_rate=hp[5], boiler_setpoint=hp[6],
      boiler_pump_head=hp[7], boiler_pump_efficiency=hp[8], boiler_heating_rate=0,
      boiler_cooling_rate=0, vav_max_air_flow_rate=hp[9],
      vav_reheat_max_water_flow_rate=hp[10])
  o = oenv.OracleEnvironment(cfg)
  o._setup_step_sim()
  o._execute_step_sim()
  got = o.boiler.return_water_temperature_sensor
  assert abs(got - 301.895482) < 1e-5
  if exact_materials:      # 8 wall CVs at the shell differ in density from the twin's
    np.testing.assert_allclose(o.temp, g["reference_temp_after_step"], rtol=2e-4)
  else:
    np.testing.assert_array_equal(o.temp, g["reference_temp_after_step"])
