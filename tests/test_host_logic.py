"""Host-side logic of sbsim_b200 (no GPU): exogenous tables against the oracle's
scalar restatements, config validation, floor-plan compiler invariants."""

import numpy as np
import pandas as pd
import pytest

import sbsim_b200 as sbx
from sbsim_b200 import exogenous, floorplan, workloads
from oracle import exogenous as oex
from oracle import hvac as ohvac
from oracle import tf_jacobi
import scenarios as S


def _ts(start, n, tz=None):
  t0 = pd.Timestamp(start, tz=tz)
  return exogenous.step_timestamps(t0, 300.0, n)


def test_sinusoid_weather_table_matches_oracle_scalar():
  ts = _ts("2023-07-06 07:00:00", 400)
  w = sbx.WeatherController(275.0, 290.0, special_days={188: (270.0, 280.0)})
  o = oex.WeatherController(275.0, 290.0, special_days={188: (270.0, 280.0)})
  np.testing.assert_array_equal(w.table(ts), [o.get_current_temp(t) for t in ts])
  # reference KATs (weather_controller_test.py:79-122): low at midnight, high at noon
  w2 = sbx.WeatherController(273.0, 283.0)
  assert w2.get_current_temp(pd.Timestamp("2021-05-09 00:00")) == pytest.approx(273.0)
  assert w2.get_current_temp(pd.Timestamp("2021-05-09 12:00")) == pytest.approx(283.0)
  assert w2.get_current_temp(pd.Timestamp("2021-05-09 06:00")) == pytest.approx(278.0)


def test_batched_sinusoid_tables_match_per_building_controllers():
  ts = _ts("2023-07-06 07:00:00", 300)
  rng = np.random.default_rng(0)
  low = rng.uniform(268, 288, 5)
  high = low + rng.uniform(5, 15, 5)
  tab = exogenous.sinusoid_weather_tables(ts, low, high)
  for b in range(5):
    o = oex.WeatherController(low[b], high[b])
    np.testing.assert_array_equal(tab[b], [o.get_current_temp(t) for t in ts])


def test_replay_weather_matches_oracle_and_raises_out_of_range():
  g = np.load(S.__file__.replace("scenarios.py", "golden/sb1_calibrated.npz"))
  w = sbx.ReplayWeatherController(times_utc_sec=g["weather_time_sec"],
                                  temps_f=g["weather_temp_f"], convection_coefficient=100.0)
  o = oex.ReplayWeatherController(g["weather_time_sec"], g["weather_temp_f"], 100.0)
  ts = _ts("2023-07-06 07:00:00", 288, tz="UTC")
  np.testing.assert_array_equal(w.table(ts), [o.get_current_temp(t) for t in ts])
  with pytest.raises(ValueError, match="before"):
    w.get_current_temp(pd.Timestamp("2023-06-01 00:00:00", tz="UTC"))
  with pytest.raises(ValueError, match="after"):
    w.get_current_temp(pd.Timestamp("2023-12-01 00:00:00", tz="UTC"))


@pytest.mark.parametrize("tz,start_tz", [("UTC", None), ("US/Pacific", "UTC")])
def test_schedule_table_matches_oracle(tz, start_tz):
  ts = _ts("2023-07-06 07:00:00", 3 * 288, tz=start_tz)   # Thursday .. Sunday
  s = sbx.SetpointSchedule(6, 19, (294, 297), (289, 298), time_zone=tz)
  o = ohvac.SetpointSchedule(6, 19, (294, 297), (289, 298), time_zone=tz)
  tab = s.table(ts)
  assert list(tab) == [int(o.is_comfort_mode(t)) for t in ts]
  assert 0 < tab.sum() < len(tab)
  with pytest.raises(ValueError, match="morning_start_hour"):
    sbx.SetpointSchedule(20, 19, (294, 297), (289, 298))


def test_occupancy_and_energy_tables_match_oracle():
  ts = _ts("2023-07-06 05:00:00", 600)
  occ = sbx.StepFunctionOccupancy(pd.Timedelta(9, unit="h"), pd.Timedelta(17, unit="h"), 1.0, 0.1)
  o = oex.StepFunctionOccupancy(pd.Timedelta(9, unit="h"), pd.Timedelta(17, unit="h"), 1.0, 0.1)
  zones = ["zone_id_1", "zone_id_2", "zone_id_3"]
  rew, obs, obs_zone = exogenous.occupancy_tables(occ, zones, ts, 300.0, per_zone=False)
  assert obs_zone.shape == (600, 1)
  dt = pd.Timedelta(300, unit="s")
  five = pd.Timedelta(5, unit="minute")
  for s in range(0, 600, 7):
    assert rew[s, 0] == o.average_zone_occupancy("z", ts[s], ts[s] + dt)
    n = 0.0
    for _ in zones:
      n += o.average_zone_occupancy("z", ts[s] - five, ts[s])
    assert obs[s] == int(n)
    assert obs_zone[s, 0] == o.average_zone_occupancy("z", ts[s] - five, ts[s])
  pe, ce, pg = exogenous.energy_tables(sbx.ElectricityEnergyCost(), sbx.NaturalGasEnergyCost(), ts)
  oe, og = oex.ElectricityEnergyCost(), oex.NaturalGasEnergyCost()
  for s in range(0, 600, 11):
    utc = pd.Timestamp(int(ts[s].timestamp()), unit="s", tz="UTC")
    assert pe[s] * 1000.0 * 300.0 == oe.cost(utc, utc + dt, 1000.0)
    assert ce[s] * 1000.0 * 300.0 == oe.carbon(utc, utc + dt, 1000.0)
    assert pg[s] * (1000.0 * 300.0) == og.cost(utc, utc + dt, 1000.0)
  # reference KATs (electricity_energy_cost_test.py / natural_gas_energy_cost_test.py values)
  assert sbx.NaturalGasEnergyCost().carbon_rate == pytest.approx(53.12 / 293.07107 / 3.6e6)


def test_time_features_match_oracle():
  ts = _ts("2023-07-06 07:00:00", 300, tz="UTC")
  tab = exogenous.time_feature_table(ts)
  for s in range(0, 300, 13):
    hod = oex.expand_time_features(1, oex.get_radian_time(ts[s], "hod"))
    dow = oex.expand_time_features(1, oex.get_radian_time(ts[s], "dow"))
    np.testing.assert_array_equal(tab[s], hod + dow)


def test_config_validation_errors_mirror_reference():
  with pytest.raises(ValueError, match="cooling_air_temp_setpoint must greater"):
    sbx.AirHandler(0.3, 298, 285, 10000.0, 0.9)
  with pytest.raises(ValueError, match="default_low_temp"):
    sbx.WeatherController(300.0, 290.0)
  n = sbx.BoundedActionNormalizer(310, 355.0)
  assert n.setpoint_value(np.float32(-1.0)) == 310.0 and n.setpoint_value(np.float32(1.0)) == 355.0
  assert n.agent_value(332.5) == pytest.approx(0.0)
  with pytest.raises(ValueError, match="not within bounds"):
    n.setpoint_value(np.float32(1.1))
  with pytest.raises(ValueError, match="expected to be scalar"):
    n.setpoint_value(np.zeros(2, dtype=np.float32))
  with pytest.raises(ValueError, match="not within bounds"):
    n.agent_value(400.0)
  norm = sbx.StandardScoreObservationNormalizer({"a": (1.1, 2.2)})
  assert norm.get("a") == (float(np.float32(1.1)), float(np.float32(2.2)))   # proto float fields
  assert norm.get("unknown") == (0.0, 1.0)


def test_floorplan_compiler_invariants_and_classes():
  cp = S.Scenario(floor_plan=S.TF_TEST_PLAN, buffer_from_walls=0).compiled()
  # classes by the rules of tf_simulator.py:208-243 on the plan of tf_simulator_test.py:36-52
  assert cp.cv_class[1, 1] == floorplan.CV_CORNER_TL       # has (2,1) and (1,2)
  assert cp.cv_class[1, 2] == floorplan.CV_CORNER_TR       # has (2,2) and (1,1); (1,3) is exterior
  assert cp.cv_class[2, 1] == floorplan.CV_EDGE_LEFT       # missing (2,0)
  assert cp.cv_class[8, 3] == floorplan.CV_EDGE_BOTTOM     # missing (9,3)
  assert cp.cv_class[3, 3] == floorplan.CV_INTERIOR
  assert cp.cv_class[0, 0] == floorplan.CV_EXTERIOR
  np.testing.assert_array_equal(cp.cv_class, tf_jacobi.classify(cp.exterior_space))
  assert cp.n_zones == 4 and list(cp.zone_ncv) == [5, 5, 4, 4]
  # descriptor packing round trip
  d = cp.desc.astype(np.int64)
  np.testing.assert_array_equal(d & 15, cp.cv_class)
  np.testing.assert_array_equal((d >> 4) & 3, cp.material_id)
  np.testing.assert_array_equal(((d >> 8) & 255), np.where(cp.zone_id >= 0, cp.zone_id, 255))
  np.testing.assert_array_equal((d & 0x40) != 0, cp.diffuser_weight > 0)
  for zi in range(cp.n_zones):
    w = cp.diffuser_weight[cp.zone_id == zi]
    if cp.zone_ndiff[zi]:
      assert w.sum() == pytest.approx(1.0)
  # a 2-neighbour CV whose neighbours are opposite has no class (tf_simulator.py:219-221)
  bad = np.full((5, 5), 2)
  bad[1:4, 2] = 1
  with pytest.raises(ValueError, match="wasn't able to determine"):
    floorplan.classify_cvs(bad == 2)
  # walls touching the frame get an air ring (building_utils.py:144-219)
  tight = np.ones((6, 6), dtype=np.int64)
  tight[1:5, 1:5] = 0
  assert floorplan.guarantee_air_padding_in_frame(tight).shape == (8, 8)


def test_randomized_workload_is_seeded_and_packs():
  a = workloads.randomized(16, seed=5, n_layouts=4)
  b = workloads.randomized(16, seed=5, n_layouts=4)
  assert all(np.array_equal(x.desc, y.desc) for x, y in zip(a.plans, b.plans))
  np.testing.assert_array_equal(a.weather_low, b.weather_low)
  packed = floorplan.pack_plans(a.plans)
  assert packed["plan_desc"].shape == (16, 64, 96)
  assert packed["zone_ncv"].shape[0] == 16 and (packed["zone_ncv"].sum(1) > 0).all()
  assert not np.array_equal(a.plans[0].material, a.plans[4].material)  # same layout, own materials


def test_environment_rejects_bad_discount_before_touching_the_gpu():
  sc = S.Scenario(floor_plan=S.small_plan(), discount=1.5)
  with pytest.raises(ValueError, match=r"Discount factor must be in \(0,1\]"):
    S.make_env(sc)


def test_randomized_occupancy_matches_reference_call_log():
  """RandomizedArrivalDepartureOccupancy (randomized_arrival_departure_occupancy.py:36-238):
  the host generator reproduces, call for call, what the unmodified reference
  Environment drew from the shipped model (fixture ref_occupancy.npz, written by
  oracle/make_golden.py:make_occupancy_golden)."""
  import os
  from sbsim_b200 import exogenous
  g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                           "ref_occupancy.npz"))
  zone_ids = [str(z) for z in g["zone_ids"]]
  n = int(g["n_steps"])
  occ = exogenous.RandomizedArrivalDepartureOccupancy(2, 7, 12, 13, 18, 300, seed=17321)
  ts = exogenous.step_timestamps(pd.Timestamp(str(g["start"])), 300.0, n + 1)
  rew, obs, obs_zone = exogenous.occupancy_tables(occ, zone_ids, ts, 300.0, per_zone=True)
  # replay the reference's call order from the tables
  Z = len(zone_ids)
  values, starts = [], []
  epoch = pd.Timestamp("1970-01-01")
  for s in range(n + 1):
    values += [obs_zone[s, z] for z in range(Z)]
    starts += [(ts[s] - pd.Timedelta(5, unit="minute") - epoch).total_seconds()] * Z
    if s > 0:
      values += [rew[s, z] for z in range(Z)]
      starts += [(ts[s] - epoch).total_seconds()] * Z
  np.testing.assert_array_equal(np.array(starts), g["call_start_sec"])
  np.testing.assert_array_equal(np.array(values), g["call_value"])
  # and the observation feature the Environment derived from it (environment.py:952-956)
  np.testing.assert_allclose((obs - 3.0) / 4.0, g["num_occupants_feature"], rtol=1e-6)
  assert 0 < g["call_value"].sum() < 2 * len(g["call_value"])


def test_randomized_occupancy_tz_aware_matches_oracle_restatement():
  """Same model on the calibrated config's tz-aware clock (sim_config.gin:164, :20): the
  batched host generator against the occupant-by-occupant restatement in oracle/."""
  from sbsim_b200 import exogenous
  from oracle import exogenous as oex
  zone_ids = [f"zone_id_{i}" for i in range(5)]
  ts = exogenous.step_timestamps(pd.Timestamp("2023-07-06 07:00:00+00:00"), 300.0, 400)
  a = exogenous.RandomizedArrivalDepartureOccupancy(3, 7, 12, 13, 18, 300, seed=17321,
                                                    time_zone="US/Pacific")
  b = oex.RandomizedArrivalDepartureOccupancy(3, 7, 12, 13, 18, 300, seed=17321,
                                              time_zone="US/Pacific")
  rew, obs, obs_zone = exogenous.occupancy_tables(a, zone_ids, ts, 300.0, per_zone=True)
  five, dt = pd.Timedelta(5, unit="minute"), pd.Timedelta(300, unit="s")
  for s, t in enumerate(ts):
    for z, zid in enumerate(zone_ids):
      assert obs_zone[s, z] == b.average_zone_occupancy(zid, t - five, t)
    if s > 0:
      for z, zid in enumerate(zone_ids):
        assert rew[s, z] == b.average_zone_occupancy(zid, t, t + dt)
  assert obs.max() > 5 and obs.min() == 0


def test_legacy_hvac_names_zones_by_coordinates():
  """The deprecated Hvac (simulator/hvac.py:35-123): devices `vav_<i>_<j>`, zone ids
  `zone_id_(i, j)` (conversion_utils.zone_coordinates_to_id), rooms of the rectangular Building
  in row-major order; observation slots follow the sorted device ids."""
  import sbsim_b200 as sbx
  from sbsim_b200 import floorplan
  air = floorplan.MaterialProperties(50.0, 700.0, 1.0)
  wall = floorplan.MaterialProperties(5.0, 800.0, 1800.0)
  cp = floorplan.legacy_building(20.0, (6, 8), (3, 4), air, wall, wall)
  coords = [(i, j) for i in range(3) for j in range(4)]
  hvac = sbx.Hvac(coords[::-1], sbx.AirHandler(0.3, 285, 298, 10000.0, 0.9, max_air_flow_rate=8.67),
                  sbx.Boiler(360.0, 6.0, 0.98), sbx.SetpointSchedule(6, 19, (294, 297), (289, 298)), 0.035, 0.03)
  b = sbx.SimulatorBuilding(cp, hvac, sbx.WeatherController(280.0, 290.0), sbx.ConstantOccupancy(1.0),
                            start_timestamp=pd.Timestamp("2023-07-06 07:00:00"), solver="gauss_seidel")
  assert b.plans[0].zone_names == [f"{i}_{j}" for i, j in coords]
  assert b.zone_ids == [f"zone_id_({i}, {j})" for i, j in coords]
  # room k of the plan (raster order of its first CV) is coordinate k in row-major order
  zid = b.plans[0].zone_id
  for k, (i, j) in enumerate(coords):
    r, c = np.argwhere(zid == k)[0]
    assert (r, c) == (3 + 7 * i, 3 + 9 * j)
  assert list(b.plans[0].obs_zone_order) == sorted(range(12), key=lambda k: "vav_%d_%d" % coords[k])
  with pytest.raises(ValueError):
    sbx.SimulatorBuilding(cp, sbx.Hvac(coords[:5], hvac.air_handler, hvac.boiler, hvac.schedule, 0.035, 0.03),
                          sbx.WeatherController(280.0, 290.0), sbx.ConstantOccupancy(1.0),
                          start_timestamp=pd.Timestamp("2023-07-06 07:00:00"))
