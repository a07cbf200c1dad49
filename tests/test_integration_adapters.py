"""integration/cuda_simulator.py: libsbx behind the reference's B2 / B3 seams.

The adapter classes subclass reference classes, and the reference tree is not on the GPU box,
so they are covered in two halves: (CPU, live reference) everything that turns a reference
building into libsbx inputs, against `FloorPlanBasedBuilding` itself; (GPU) the solve /
observation functions the classes call, on a duck-typed building, against the oracle."""

import types
import warnings

import numpy as np
import pandas as pd
import pytest

from integration import cuda_simulator as cs
from oracle import refshim
from oracle import tf_jacobi
from sbsim_b200 import _lib, floorplan
import scenarios as S

needs_ref = pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ref():
  warnings.filterwarnings("ignore")
  from oracle import make_golden
  return make_golden._ref_modules()


def _ref_building(ref, plan, cv=20.0, bfw=2):
  bp = ref["building"]
  return bp.FloorPlanBasedBuilding(
      cv_size_cm=cv, floor_height_cm=300.0, initial_temp=292.0,
      inside_air_properties=bp.MaterialProperties(50.0, 700.0, 1.0),
      inside_wall_properties=bp.MaterialProperties(2.0, 1000.0, 1800.0),
      building_exterior_properties=bp.MaterialProperties(0.05, 1000.0, 3000.0),
      floor_plan=plan, zone_map=plan.copy(), buffer_from_walls=bfw)


@needs_ref
@pytest.mark.parametrize("which", ["small", "random"])
def test_plan_from_reference_building_equals_the_compiler(ref, which):
  """The adapter reads the reference building's own arrays; the result is the plan
  floorplan.compile_plan builds from the same floor plan (B3 / B2 plan packing)."""
  if which == "small":
    plan, cv, bfw = S.small_plan(), 20.0, 2
  else:
    plan, cv, bfw = floorplan.random_floor_plan(np.random.default_rng(4)).astype(np.int64), 10.0, 3
  b = _ref_building(ref, plan, cv, bfw)
  got = cs.plan_from_reference_building(b)
  want = S.Scenario(floor_plan=plan, cv_size_cm=cv, buffer_from_walls=bfw).compiled()
  np.testing.assert_array_equal(got.desc, want.desc)
  np.testing.assert_array_equal(got.material, want.material)
  assert got.zone_names == want.zone_names and got.cv_size_m == want.cv_size_m
  np.testing.assert_array_equal(got.zone_ncv, want.zone_ncv)
  np.testing.assert_array_equal(got.zone_ndiff, want.zone_ndiff)
  np.testing.assert_array_equal(got.obs_zone_order, want.obs_zone_order)
  packed_got, packed_want = floorplan.pack_plans([got]), floorplan.pack_plans([want])
  assert sorted(packed_got) == sorted(packed_want)
  for k in packed_got:
    np.testing.assert_array_equal(packed_got[k], packed_want[k], err_msg=k)


@needs_ref
def test_diffuser_heat_per_zone_reads_apply_thermal_power_zone(ref):
  """building.apply_thermal_power_zone (building.py:873-889) spreads a zone's power over its
  diffusers; the adapter hands libsbx the per-diffuser value of each zone."""
  b = _ref_building(ref, S.small_plan())
  plan = cs.plan_from_reference_building(b)
  powers = [350.0, -120.0, 0.0, 77.5]
  for name, pw in zip(plan.zone_names, powers):
    b.apply_thermal_power_zone(name, pw)
  q = cs.diffuser_heat_per_zone(b, plan)
  assert q.shape == (1, plan.n_zones) and q.dtype == np.float32
  for zi, pw in enumerate(powers):
    assert q[0, zi] == np.float32(pw / plan.zone_ndiff[zi])
    dense = np.asarray(b.input_q)[(plan.zone_id == zi) & (plan.diffuser_weight > 0)]
    assert np.all(dense == dense[0])


@needs_ref
def test_adapter_classes_subclass_the_reference_seams(ref):
  """make_cuda_simulator / make_cuda_simulator_building produce subclasses of the reference's
  SimulatorFlexibleGeometries (seam B3) and BaseBuilding (seam B2) with every abstract member
  implemented."""
  from smart_buildings.smart_control.models import base_building
  from smart_buildings.smart_control.simulator import simulator_flexible_floor_plan as sffp
  sim_cls = cs.make_cuda_simulator()
  assert issubclass(sim_cls, sffp.SimulatorFlexibleGeometries)
  assert "finite_differences_timestep" in sim_cls.__dict__
  bld_cls = cs.make_cuda_simulator_building()
  assert issubclass(bld_cls, base_building.BaseBuilding)
  assert not getattr(bld_cls, "__abstractmethods__", None)


def test_make_sbx_config_is_a_solve_only_handle_configuration():
  cp = S.Scenario(floor_plan=S.small_plan()).compiled()
  cfg = cs.make_sbx_config(cp, time_step_sec=300.0, convergence_threshold=0.1, iteration_limit=100,
                           floor_height_cm=300.0, solver="gauss_seidel")
  assert (cfg.n_envs, cfg.height, cfg.width, cfg.n_zones) == (1, cp.height, cp.width, cp.n_zones)
  assert cfg.solver == _lib.SOLVER_GAUSS_SEIDEL and cfg.abi_version == _lib.ABI_VERSION
  assert (cfg.time_step_sec, cfg.floor_height_m, cfg.convergence_threshold) == (300.0, 3.0, 0.1)


@pytest.mark.gpu
@pytest.mark.parametrize("plan_name", ["small", "random"])
def test_cuda_fd_timestep_matches_the_oracle(plan_name):
  """What CudaSimulator.finite_differences_timestep runs (seam B3): building.temp / input_q in,
  the converged field back in building.temp -- bit-identical to the TF-Jacobi oracle."""
  if plan_name == "small":
    sc = S.Scenario(floor_plan=S.small_plan())
  else:
    sc = S.Scenario(floor_plan=floorplan.random_floor_plan(np.random.default_rng(4)).astype(np.int64),
                    cv_size_cm=10.0, buffer_from_walls=3)
  cp = sc.compiled()
  cfg = cs.make_sbx_config(cp, time_step_sec=sc.time_step_sec, convergence_threshold=sc.convergence_threshold,
                           iteration_limit=sc.iteration_limit, floor_height_cm=sc.floor_height_cm)
  h = _lib.Handle(cfg, 0)
  try:
    cs.upload_plan(h, cp)
    rng = np.random.default_rng(2)
    temp = rng.uniform(286, 299, (cp.height, cp.width)).astype(np.float32)
    q = np.zeros((cp.height, cp.width))
    per_zone = rng.uniform(-40, 300, cp.n_zones)
    for zi in range(cp.n_zones):
      q[(cp.zone_id == zi) & (cp.diffuser_weight > 0)] = np.float32(per_zone[zi])
    building = types.SimpleNamespace(temp=temp.copy(), input_q=q)
    jac = tf_jacobi.TFJacobi(S.oracle_plan(cp, sc.floor_height_cm), sc.time_step_sec,
                             sc.convergence_threshold, sc.iteration_limit)
    for step in range(3):
      want, n, _, _ = jac.fd_step(np.asarray(building.temp, dtype=np.float32), q, 281.0 + step, 20.0)
      converged = cs.cuda_fd_timestep(h, cp, building, "tf_jacobi", 281.0 + step, 20.0,
                                      sc.iteration_limit, sc.convergence_threshold)
      assert converged and int(h.download("n_sweeps", (1,))[0]) == n
      np.testing.assert_array_equal(building.temp, want)
  finally:
    h.close()


@pytest.mark.gpu
def test_native_observations_are_the_denormalized_observation():
  """CudaSimulatorBuilding.request_observations (seam B2) serves native values from the step's
  diagnostics; normalizing them reproduces the observation the environment returned."""
  sc = S.Scenario(floor_plan=S.small_plan(), occupancy="step", start="2023-07-06 08:50:00")
  env = S.make_env(sc, n_envs=1)
  try:
    env.reset()
    rng = np.random.default_rng(5)
    for _ in range(3):
      ts = env.step(rng.uniform(-1, 1, (1, 2)).astype(np.float32))
    native = cs.native_observations(env)
    norm = S.NORMALIZATION
    checked = 0
    for i, name in enumerate(env.field_names):
      for (dev, field), v in native.items():
        if name == f"{dev}_{field}" and field in norm:
          mean, var = norm[field]
          np.testing.assert_allclose((v - mean) / np.sqrt(var), ts.observation[0, i], rtol=1e-5, atol=1e-6,
                                     err_msg=name)
          checked += 1
    assert checked >= 10          # every field the scenario has normalization constants for
  finally:
    env.close()
