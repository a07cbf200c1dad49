"""Live checks against the UNMODIFIED reference (only where /root/reference is
mounted; skipped on the GPU box, which relies on the committed goldens instead)."""

import warnings

import numpy as np
import pandas as pd
import pytest

from oracle import refshim
from oracle import tf_jacobi
from sbsim_b200 import floorplan
import scenarios as S

pytestmark = pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ref():
  warnings.filterwarnings("ignore")
  from oracle import make_golden
  return make_golden._ref_modules()


def _ref_building(ref, plan, cv=20.0, bfw=3):
  bp = ref["building"]
  return bp.FloorPlanBasedBuilding(
      cv_size_cm=cv, floor_height_cm=300.0, initial_temp=292.0,
      inside_air_properties=bp.MaterialProperties(50.0, 700.0, 1.0),
      inside_wall_properties=bp.MaterialProperties(2.0, 1000.0, 1800.0),
      building_exterior_properties=bp.MaterialProperties(0.05, 1000.0, 3000.0),
      floor_plan=plan, zone_map=plan.copy(), buffer_from_walls=bfw)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_static_compiler_matches_reference_building(ref, seed):
  rng = np.random.default_rng(seed)
  spec = floorplan.RandomPlanSpec(40, 56, (1, 2), (2, 3), 5) if seed % 2 else floorplan.RandomPlanSpec()
  plan = floorplan.random_floor_plan(rng, spec).astype(np.int64)
  b = _ref_building(ref, plan, cv=10.0)
  cp = S.Scenario(floor_plan=plan, cv_size_cm=10.0, buffer_from_walls=3).compiled()
  np.testing.assert_array_equal(cp.exterior_space, b._exterior_space == -1)
  np.testing.assert_array_equal(cp.dense_material(0), b.conductivity)
  np.testing.assert_array_equal(cp.dense_material(1), b.heat_capacity)
  np.testing.assert_array_equal(cp.dense_material(2), b.density)
  np.testing.assert_array_equal(cp.diffuser_weight, b.diffusers)
  rooms = [k for k in b._room_dict if k.startswith("room")]
  assert rooms == cp.zone_names
  for zi, k in enumerate(rooms):
    r, c = cp.zone_indices(zi)
    assert [tuple(x) for x in b._room_dict[k]] == list(zip(r.tolist(), c.tolist()))
  # classes against the reference's classify_cv
  tfs = ref["tfs"]
  codes = {("EDGE", "TOP"): 2, ("EDGE", "BOTTOM"): 3, ("EDGE", "LEFT"): 4, ("EDGE", "RIGHT"): 5,
           ("CORNER", "TOP_LEFT"): 6, ("CORNER", "BOTTOM_LEFT"): 7, ("CORNER", "TOP_RIGHT"): 8,
           ("CORNER", "BOTTOM_RIGHT"): 9}
  for (i, j), t in tfs.get_cv_mapping(b.neighbors, tfs.CVPositionType.BOUNDARY).items():
    assert cp.cv_class[i, j] == codes[(t.boundary.name, (t.edge or t.corner).name)]
  ext = tfs.get_cv_mapping(b.neighbors, tfs.CVPositionType.EXTERIOR)
  assert set(ext) == set(map(tuple, np.argwhere(cp.cv_class == 0)))


def test_tf_sweep_restatement_is_bit_identical_to_reference_code(ref):
  """Unmodified TFSimulator.update_temperature_estimates (tf_simulator.py:573-853) on
  the NumPy provider of its TF primitives vs oracle/tf_jacobi.py, random fields."""
  rng = np.random.default_rng(7)
  plan = S.small_plan()
  b = _ref_building(ref, plan, bfw=2)
  hv = ref["hvac"].FloorPlanBasedHvac(
      air_handler=ref["ah"].AirHandler(0.3, 285, 298, 10000.0, 0.9),
      boiler=ref["bl"].Boiler(360.0, 6.0, 0.98), schedule=ref["ss"].SetpointSchedule(
          6, 19, (294, 297), (289, 298)),
      vav_max_air_flow_rate=0.035, vav_reheat_max_water_flow_rate=0.03)
  w = ref["wc"].WeatherController(275.0, 290.0)
  sim = ref["tfs"].TFSimulator(b, hv, w, 300.0, 0.1, 100, 30, pd.Timestamp("2023-07-06"))
  cp = S.Scenario(floor_plan=plan).compiled()
  jac = tf_jacobi.TFJacobi(S.oracle_plan(cp, 300.0), 300.0, 0.1, 100)
  for _ in range(5):
    b.temp = rng.uniform(280, 300, plan.shape)
    b.input_q = np.where(b.diffusers > 0, rng.uniform(-100, 900, plan.shape), 0.0)
    est = rng.uniform(280, 300, plan.shape)
    amb, h = float(rng.uniform(270, 300)), float(rng.uniform(5, 100))
    want, wmd = sim.update_temperature_estimates(est.copy(), amb, h)
    got, gmd = jac.sweep(est, b.temp, b.input_q, amb, h)
    np.testing.assert_array_equal(got, want)
    assert gmd == wmd


def test_oracle_env_matches_live_reference_env(ref):
  from oracle import make_golden
  env, b = make_golden.build_reference_env(ref, S.small_plan(), "tf", histogram=True)
  o = S.make_oracle(S.Scenario(floor_plan=S.small_plan(), histogram=True))
  rts, ots = env.reset(), o.reset()
  np.testing.assert_allclose(ots[3], rts.observation, rtol=1e-6, atol=1e-7)
  rng = np.random.default_rng(5)
  for _ in range(15):
    a = rng.uniform(-1, 1, 2).astype(np.float32)
    rts, ots = env.step(a), o.step(a)
    np.testing.assert_allclose(ots[3], rts.observation, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(float(ots[1]), float(rts.reward), rtol=1e-6, atol=1e-9)
    np.testing.assert_array_equal(np.asarray(o.temp), np.asarray(b.temp))


@pytest.mark.parametrize("room_shape,building_shape", [((20, 30), (3, 3)), ((5, 7), (2, 4)),
                                                       ((9, 4), (1, 1)), ((6, 11), (4, 2))])
def test_legacy_building_plan_matches_reference_building(ref, room_shape, building_shape):
  """legacy_building() against the reference's deprecated rectangular Building
  (building.py:394-505) for several shapes: materials and diffusers, one ring of exterior
  space added around the old grid."""
  bp = ref["building"]
  air = (50.0, 700.0, 1.0)
  wall = (5.0, 800.0, 1800.0)
  ext = (0.5, 900.0, 3000.0)
  old = bp.Building(20.0, 300.0, room_shape, building_shape, 292.0, bp.MaterialProperties(*air),
                    bp.MaterialProperties(*wall), bp.MaterialProperties(*ext))
  cp = floorplan.legacy_building(20.0, room_shape, building_shape, floorplan.MaterialProperties(*air),
                                 floorplan.MaterialProperties(*wall), floorplan.MaterialProperties(*ext))
  assert (cp.height, cp.width) == (old.temp.shape[0] + 2, old.temp.shape[1] + 2)
  for k, arr in enumerate((old.conductivity, old.heat_capacity, old.density)):
    np.testing.assert_array_equal(cp.dense_material(k)[1:-1, 1:-1], arr)
  np.testing.assert_array_equal(cp.diffuser_weight[1:-1, 1:-1] > 0, old.diffusers > 0)
  np.testing.assert_allclose(cp.diffuser_weight[1:-1, 1:-1], old.diffusers.astype(np.float64))
  assert cp.n_zones == building_shape[0] * building_shape[1]
  # zones in the old building's (zone_x, zone_y) row-major order, air CVs only
  zi = 0
  for zx in range(building_shape[0]):
    for zy in range(building_shape[1]):
      x0 = zx * (room_shape[0] + 1) + 2
      y0 = zy * (room_shape[1] + 1) + 2
      block = cp.zone_id[1:-1, 1:-1][x0:x0 + room_shape[0], y0:y0 + room_shape[1]]
      assert np.all(block == zi), (zx, zy)
      zi += 1


def test_legacy_hvac_naming_matches_reference(ref):
  """sbx.Hvac + legacy_building() name devices and zones like the deprecated
  simulator/hvac.py:Hvac (`vav_<i>_<j>`, `zone_id_(i, j)`), room k <-> coordinate k row-major."""
  import importlib
  import sbsim_b200 as sbx
  hvac_mod = importlib.import_module("smart_buildings.smart_control.simulator.hvac")
  coords = [(i, j) for i in range(2) for j in range(3)]
  weather = ref["wc"].WeatherController(275.0, 290.0, convection_coefficient=60.0)
  sched = ref["ss"].SetpointSchedule(6, 19, (294, 297), (289, 298))
  boiler = ref["bl"].Boiler(360.0, 6.0, 0.98, heating_rate=0.5, cooling_rate=0.1, device_id="boiler_id_x")
  ahu = ref["ah"].AirHandler(0.3, 285, 298, 10000.0, 0.9, max_air_flow_rate=8.67, device_id="air_handler_id_x",
                             sim_weather_controller=weather)
  rh = hvac_mod.Hvac(coords, ahu, boiler, sched, 0.035, 0.03)
  air = floorplan.MaterialProperties(50.0, 700.0, 1.0)
  wall = floorplan.MaterialProperties(5.0, 800.0, 1800.0)
  cp = floorplan.legacy_building(20.0, (6, 8), (2, 3), air, wall, wall)
  b = sbx.SimulatorBuilding(
      cp, sbx.Hvac(coords, sbx.AirHandler(0.3, 285, 298, 10000.0, 0.9, max_air_flow_rate=8.67),
                   sbx.Boiler(360.0, 6.0, 0.98), sbx.SetpointSchedule(6, 19, (294, 297), (289, 298)), 0.035, 0.03),
      sbx.WeatherController(275.0, 290.0), sbx.ConstantOccupancy(1.0),
      start_timestamp=pd.Timestamp("2023-07-06 07:00:00"), solver="gauss_seidel")
  assert ["vav_" + n for n in b.plans[0].zone_names] == [rh.vavs[z].device_id() for z in coords]
  assert b.zone_ids == [rh.zone_infos[z].zone_id for z in coords]
