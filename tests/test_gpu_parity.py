"""GPU parity: libsbx (CUDA, through the C ABI) against the oracle.

Bars
  * temperature field of a diffusion solve: BIT-EXACT (same fp32 op order as the
    reference's eager TF arithmetic), identical sweep counts;
  * zone temperatures, energy rates, reward, observations: 1e-4 relative
    (BASELINE.json north_star), fp64 device algebra vs the oracle's Python floats.
"""

import numpy as np
import pandas as pd
import pytest

import sbsim_b200 as sbx
from sbsim_b200 import _lib, floorplan, workloads
import scenarios as S
from oracle import tf_jacobi

pytestmark = pytest.mark.gpu

RTOL = 1e-4  # north_star tolerance (fp32 relative)
PATHS = {"streaming": sbx.PATH_STREAMING, "resident": sbx.PATH_RESIDENT}


def _plans():
  rng = np.random.default_rng(11)
  return {
      "tf_test_10x9": (S.TF_TEST_PLAN, 0),
      "small_24x34": (S.small_plan(), 2),
      "odd_23x31": (S.small_plan(23, 31), 2),
      "rand_64x96": (floorplan.random_floor_plan(rng).astype(np.int64), 3),
      # wider than a TMA tensor-map box (256): the resident kernel copies row by row
      "wide_12x320": (S.small_plan(12, 320), 1),
  }


def _dense_q(cp, qcv):
  q = np.zeros(cp.desc.shape, dtype=np.float64)
  for zi in range(cp.n_zones):
    m = (cp.zone_id == zi) & (cp.diffuser_weight > 0)
    q[m] = qcv[zi]
  return q


@pytest.fixture(params=["k_resident_step", "k_resident_step2", "k_resident_step3"])
def resident_kernel(request, monkeypatch):
  """The resident path has three kernels: k_resident_step (the default), the record-driven
  k_resident_step2 (SBX_RESIDENT_V2=1) and k_resident_step3 (half-tiles in registers,
  SBX_RESIDENT_V3=1, grids with even height and width a multiple of 4); the switches are read
  when a handle is created."""
  monkeypatch.delenv("SBX_RESIDENT_V2", raising=False)
  monkeypatch.delenv("SBX_RESIDENT_V3", raising=False)
  if request.param == "k_resident_step2":
    monkeypatch.setenv("SBX_RESIDENT_V2", "1")
  elif request.param == "k_resident_step3":
    monkeypatch.setenv("SBX_RESIDENT_V3", "1")
  return request.param


def test_resident_kernels_agree_bit_for_bit(resident_kernel):
  """Both resident kernels against the oracle on random 64x96 plans with per-env plans,
  weather and actions: bit-identical fields and identical sweep counts on the first step,
  1e-4 afterwards (the workload of BASELINE.json configs[2..4])."""
  from oracle import bench_support
  B = 6
  wl = workloads.randomized(B, seed=5, n_layouts=B)
  env, _ = workloads.make_randomized_env(B, workload=wl, episode_steps=16, histogram=True,
                                         kernel_path=sbx.PATH_RESIDENT)
  try:
    oracles = [bench_support.make_oracle_env(
        wl.plans[b], float(wl.weather_low[b]), float(wl.weather_high[b]), float(wl.convection[b]),
        float(wl.initial_temp[b]), workloads.NORMALIZATION, workloads.HISTOGRAM, 16,
        workloads.DEFAULT_START) for b in range(B)]
    env.reset()
    for o in oracles:
      o.reset()
    rng = np.random.default_rng(9)
    for step in range(6):
      a = rng.uniform(-1, 1, (B, 2)).astype(np.float32)
      ts = env.step(a)
      temp = env.handle.download("temp", (B, 64, 96))
      sweeps = env.handle.download("n_sweeps", (B,))
      assert env.handle.info().resident_kernel == int(resident_kernel[-1:].replace("p", "1")), resident_kernel
      for b, o in enumerate(oracles):
        ots = o.step(a[b])
        assert sweeps[b] == o.info["n_sweeps"], (step, b)
        np.testing.assert_allclose(ts.observation[b], ots[3], rtol=RTOL, atol=2e-5)
        # free-running: the AC power is a cancellation of two ~290 K temperatures, one of them an
        # fp32 zone mean that may sit 1 ulp from the reference's pairwise mean (see _compare_step)
        np.testing.assert_allclose(ts.reward[b], ots[1], rtol=RTOL, atol=5e-5)
        if step == 0:
          np.testing.assert_array_equal(temp[b], np.asarray(o.temp, dtype=np.float32))
        else:
          np.testing.assert_allclose(temp[b], np.asarray(o.temp, dtype=np.float64), rtol=RTOL)
  finally:
    env.close()


@pytest.mark.parametrize("sweep", ["k_sweep_tma", "k_sweep"])
@pytest.mark.parametrize("plan_name", ["rand_64x96", "wide_12x320"])
def test_streaming_sweep_kernels_bit_exact(sweep, plan_name, monkeypatch):
  """The streaming path has two sweep kernels for 4-CV vectors: k_sweep_tma (TMA-staged tiles,
  the default) and k_sweep (plain loads, SBX_SWEEP_TMA=0; also what 1-CV vectors run).  Both
  against the oracle: a solve with per-building ambient / convection / heat input bit for bit,
  then a free-running rollout whose observations must be IDENTICAL between the two kernels."""
  monkeypatch.delenv("SBX_SWEEP_TMA", raising=False)
  if sweep == "k_sweep":
    monkeypatch.setenv("SBX_SWEEP_TMA", "0")
  plan, bfw = _plans()[plan_name]
  sc = S.Scenario(floor_plan=plan, buffer_from_walls=bfw, cv_size_cm=10.0 if "rand" in plan_name else 20.0)
  cp = sc.compiled()
  B = 5
  env = S.make_env(sc, n_envs=B, plans=cp, kernel_path=sbx.PATH_STREAMING)
  try:
    env.reset()
    rng = np.random.default_rng(3)
    H, W, Z = cp.height, cp.width, env.building.n_zones
    temp = (rng.uniform(285, 300, (B, H, W))).astype(np.float32)
    qcv = rng.uniform(-50, 400, (B, Z)).astype(np.float32)
    ambient = rng.uniform(270, 300, B)
    conv = rng.uniform(5, 100, B)
    env.handle.upload("temp", temp)
    env.handle.upload("q_cv", qcv)
    env.handle.fd_step(ambient, conv)
    got = env.handle.download("temp", (B, H, W))
    sweeps = env.handle.download("n_sweeps", (B,))
    jac = tf_jacobi.TFJacobi(S.oracle_plan(cp, sc.floor_height_cm), sc.time_step_sec,
                             sc.convergence_threshold, sc.iteration_limit)
    for b in range(B):
      want, n, _, _ = jac.fd_step(temp[b], _dense_q(cp, qcv[b]), ambient[b], conv[b])
      assert sweeps[b] == n, (b, sweeps[b], n)
      np.testing.assert_array_equal(got[b], want, err_msg=f"env {b}")
    # a rollout from a fresh reset: a digest of everything the step returns, compared across the
    # two parametrizations through a module-level dict
    env.reset()
    arng = np.random.default_rng(8)
    digest = []
    for _ in range(6):
      ts = env.step(arng.uniform(-1, 1, (B, 2)).astype(np.float32))
      digest.append((ts.observation.tobytes(), ts.reward.tobytes(), env.handle.download("temp", (B, H, W)).tobytes()))
    other = _SWEEP_DIGESTS.setdefault(plan_name, {})
    other[sweep] = digest
    if len(other) == 2:
      assert other["k_sweep_tma"] == other["k_sweep"]
  finally:
    env.close()


_SWEEP_DIGESTS = {}


@pytest.mark.parametrize("path", list(PATHS))
@pytest.mark.parametrize("plan_name", list(_plans()))
def test_fd_step_bit_exact(path, plan_name):
  """B3 seam: finite_differences_timestep (simulator.py:318-371) on the TF sweep."""
  plan, bfw = _plans()[plan_name]
  sc = S.Scenario(floor_plan=plan, buffer_from_walls=bfw, cv_size_cm=10.0 if "rand" in plan_name else 20.0)
  cp = sc.compiled()
  B = 5
  env = S.make_env(sc, n_envs=B, plans=cp, kernel_path=PATHS[path])
  try:
    env.reset()
    rng = np.random.default_rng(3)
    H, W, Z = cp.height, cp.width, env.building.n_zones
    temp = (rng.uniform(285, 300, (B, H, W))).astype(np.float32)
    temp[0] = 292.0           # mirrors tf_simulator_test.py:661-723 (uniform 292 K, T_inf 285, h 12)
    qcv = rng.uniform(-50, 400, (B, Z)).astype(np.float32)
    qcv[0] = 0
    ambient = rng.uniform(270, 300, B)
    conv = rng.uniform(5, 100, B)
    ambient[0], conv[0] = 285.0, 12.0
    env.handle.upload("temp", temp)
    env.handle.upload("q_cv", qcv)
    env.handle.fd_step(ambient, conv)
    got = env.handle.download("temp", (B, H, W))
    sweeps = env.handle.download("n_sweeps", (B,))
    md = env.handle.download("max_delta", (B,))
    jac = tf_jacobi.TFJacobi(S.oracle_plan(cp, sc.floor_height_cm), sc.time_step_sec,
                             sc.convergence_threshold, sc.iteration_limit)
    for b in range(B):
      want, n, _, wmd = jac.fd_step(temp[b], _dense_q(cp, qcv[b]), ambient[b], conv[b])
      assert sweeps[b] == n, (b, sweeps[b], n)
      np.testing.assert_array_equal(got[b], want, err_msg=f"env {b}")
      assert md[b] == np.float32(wmd)
    if plan_name == "tf_test_10x9":
      # reference's own assertions for this plan: max_delta >= 7 on sweep 1,
      # convergence within 4 sweeps (tf_simulator_test.py:691, 722)
      _, first = jac.sweep(temp[0], temp[0], np.zeros((H, W)), 285.0, 12.0)
      assert first >= 7.0
      assert sweeps[0] <= 4
  finally:
    env.close()


def test_resident_kernels_solve_only_and_many_rooms(resident_kernel):
  """sbx_fd_step (no zone sums) on a random 64x96 plan and the plan with several rooms per 4-CV
  vector, under each resident kernel."""
  test_fd_step_bit_exact("resident", "rand_64x96")
  test_many_rooms_per_vector("resident")


def _compare_step(env, oracles, ts, B, cp, step, check_temp=True):
  Z = cp.n_zones
  H, W = cp.height, cp.width
  temp = env.handle.download("temp", (B, H, W))
  sweeps = env.handle.download("n_sweeps", (B,))
  diag = env.handle.download("step_diag", (B, _lib.DIAG_N))
  zmean = env.handle.download("zone_mean", (B, env.building.n_zones))
  stats = dict(temp_exact=0, sweeps_equal=0)
  for b, o in enumerate(oracles):
    ots = o._current_time_step
    assert int(ts.step_type[b]) == int(ots[0]), (step, b)
    np.testing.assert_allclose(ts.discount[b], ots[2], rtol=1e-6)
    np.testing.assert_allclose(ts.observation[b], ots[3], rtol=RTOL, atol=2e-5,
                               err_msg=f"obs step {step} env {b}")
    if int(ots[0]) != 0:
      np.testing.assert_allclose(ts.reward[b], ots[1], rtol=RTOL, atol=1e-6,
                                 err_msg=f"reward step {step} env {b}")
      info = o.info["reward_info"]
      want = dict(blower_w=info.blower_electrical_energy_rate,
                  ac_w=info.air_conditioning_electrical_energy_rate,
                  gas_w=info.natural_gas_heating_energy_rate,
                  pump_w=info.pump_electrical_energy_rate)
      got = {k: diag[b, _lib.DIAG[k]] for k in want}
      # HVAC electrical / gas energy within 1e-4 relative (north_star)
      elec_w = want["blower_w"] + abs(want["ac_w"]) + want["pump_w"]
      elec_g = got["blower_w"] + abs(got["ac_w"]) + got["pump_w"]
      np.testing.assert_allclose(elec_g, elec_w, rtol=RTOL, err_msg=f"electricity step {step} env {b}")
      np.testing.assert_allclose(got["gas_w"], want["gas_w"], rtol=RTOL, atol=1e-3,
                                 err_msg=f"gas step {step} env {b}")
      # Components: AC power is flow*c_p*(supply - mixed), a difference of two ~290 K
      # temperatures one of which comes from an fp32 grid mean; the reference's own
      # fp32 pairwise mean and our fp64-accumulated mean may differ by 1 ulp
      # (3.05e-5 K), hence an absolute floor of 2 ulp * flow * c_p.
      ulp = 3.05e-5
      flow = diag[b, _lib.DIAG["ahu_flow"]]
      bflow = diag[b, _lib.DIAG["boiler_flow"]]
      atol = dict(blower_w=1e-6, pump_w=1e-6, ac_w=2 * ulp * 1006.0 * flow + 1e-6,
                  gas_w=2 * ulp * 4180.0 * bflow + 1e-3)
      for name in want:
        np.testing.assert_allclose(got[name], want[name], rtol=RTOL, atol=atol[name],
                                   err_msg=f"{name} step {step} env {b}")
      assert sweeps[b] == o.info["n_sweeps"], (step, b, sweeps[b], o.info["n_sweeps"])
    zt = np.array([float(t) for t in o.zone_average_temps()])
    np.testing.assert_allclose(zmean[b, :Z], zt, rtol=RTOL, err_msg=f"zone temps step {step}")
    if check_temp:
      np.testing.assert_allclose(temp[b], np.asarray(o.temp, dtype=np.float64), rtol=RTOL,
                                 err_msg=f"temp step {step} env {b}")
      stats["temp_exact"] += int(np.array_equal(temp[b], np.asarray(o.temp, dtype=np.float32)))
  return stats


@pytest.mark.parametrize("path", list(PATHS))
@pytest.mark.parametrize("plan_name,histogram", [("small_24x34", False), ("small_24x34", True),
                                                 ("odd_23x31", False), ("rand_64x96", False)])
def test_env_rollout_matches_oracle(path, plan_name, histogram):
  """Free-running Environment.reset()/step() rollout, 3 envs x 60 steps."""
  plan, bfw = _plans()[plan_name]
  sc = S.Scenario(floor_plan=plan, buffer_from_walls=bfw, histogram=histogram,
                  cv_size_cm=10.0 if "rand" in plan_name else 20.0)
  cp = sc.compiled()
  B, N = 3, 60
  env = S.make_env(sc, n_envs=B, plans=cp, kernel_path=PATHS[path])
  try:
    assert env.kernel_path == PATHS[path]
    oracles = [S.make_oracle(sc, cp) for _ in range(B)]
    ts = env.reset()
    for o in oracles:
      o.reset()
    assert ts.observation.shape == (B, env.observation_spec().shape[0])
    _compare_step(env, oracles, ts, B, cp, -1)
    rngs = [np.random.default_rng(1000 + b) for b in range(B)]
    exact = 0
    for step in range(N):
      a = np.stack([r.uniform(-1, 1, 2).astype(np.float32) for r in rngs])
      ts = env.step(a)
      for b, o in enumerate(oracles):
        o.step(a[b])
      st = _compare_step(env, oracles, ts, B, cp, step)
      exact += st["temp_exact"]
    # Free-running, the field stays within RTOL (asserted above) but not bit-identical
    # for ever: the reference's fp32 pairwise zone mean and our fp64-accumulated mean
    # may differ by 1 ulp, which perturbs the next step's diffuser heat.  The first
    # steps (zero heat input) must be exact; test_fd_step_bit_exact covers the rest.
    print(f"bit-identical temperature fields: {exact}/{B * N}")
    assert exact >= B, exact
  finally:
    env.close()


def _strip_rooms_plan(h=64, w=96):
  """One-CV-wide rooms separated by one-CV walls: every 4-CV vector touches two rooms
  and a wall, so the resident kernel's zone-sum list overflows its capacity and the
  generic per-CV loop runs instead."""
  plan = np.full((h, w), 2, dtype=np.int64)
  plan[2:h - 2, 2:w - 2] = 1
  for c in range(3, w - 3, 2):
    plan[3:h - 3, c] = 0
  return plan


@pytest.mark.parametrize("path", list(PATHS))
def test_many_rooms_per_vector(path):
  sc = S.Scenario(floor_plan=_strip_rooms_plan(), buffer_from_walls=0, cv_size_cm=20.0)
  cp = sc.compiled()
  assert cp.n_zones == 45
  B, N = 2, 6
  env = S.make_env(sc, n_envs=B, plans=cp, kernel_path=PATHS[path])
  try:
    oracles = [S.make_oracle(sc, cp) for _ in range(B)]
    ts = env.reset()
    for o in oracles:
      o.reset()
    _compare_step(env, oracles, ts, B, cp, -1)
    rng = np.random.default_rng(5)
    for step in range(N):
      a = rng.uniform(-1, 1, (B, 2)).astype(np.float32)
      ts = env.step(a)
      for b, o in enumerate(oracles):
        o.step(a[b])
      _compare_step(env, oracles, ts, B, cp, step)
  finally:
    env.close()


def test_episode_boundaries_and_auto_reset():
  """step_type / discount sequence and the auto-reset of environment.py:1252, 1313-1368."""
  n = 5
  sc = S.Scenario(floor_plan=S.small_plan(), num_days=(n + 0.5) * 300.0 / 86400.0)
  assert sc.n_steps == n
  cp = sc.compiled()
  env = S.make_env(sc, n_envs=2, plans=cp)
  try:
    o = S.make_oracle(sc, cp)
    a = np.zeros((2, 2), dtype=np.float32)
    first = env.step(a)                       # step before reset == reset [TF-Agents]
    o.step(a[0])
    assert list(first.step_type) == [0, 0] and list(first.discount) == [1.0, 1.0]
    types = []
    for _ in range(2 * (n + 2)):
      ts = env.step(a)
      ots = o.step(a[0])
      types.append(int(ts.step_type[0]))
      assert int(ts.step_type[0]) == int(ots[0])
      np.testing.assert_allclose(ts.observation[0], ots[3], rtol=RTOL, atol=2e-5)
      np.testing.assert_allclose(ts.reward[0], ots[1], rtol=RTOL, atol=1e-6)
    assert types[:n + 2] == [1] * n + [2, 0]
    assert types[n + 2:] == [1] * n + [2, 0]
  finally:
    env.close()


def test_action_out_of_bounds_raises():
  sc = S.Scenario(floor_plan=S.small_plan())
  env = S.make_env(sc, n_envs=2)
  try:
    env.reset()
    with pytest.raises(ValueError, match="not within bounds"):
      env.step(np.array([[0.0, 1.5], [0.0, 0.0]], dtype=np.float32))
    env.step(np.array([[1.0 + 5e-6, -1.0], [0.0, 0.0]], dtype=np.float32))  # inside the 1e-5 tolerance
    # the check lives in the C ABI (sbx_step_host): a non-Python caller gets SBX_E_INVALID with the
    # reference's message, and nothing has been stepped
    before = env.handle.download("episode", (4,)).copy()
    bad = np.array([[0.0, np.nan], [0.0, 0.0]], dtype=np.float32)
    with pytest.raises(_lib.SbxLibraryError, match="not within bounds"):
      env.handle.step_host(bad, env._obs, env._reward, env._step_type, env._discount)
    np.testing.assert_array_equal(env.handle.download("episode", (4,)), before)
    with pytest.raises(ValueError, match="action shape"):
      env.step(np.zeros((2, 3), dtype=np.float32))
  finally:
    env.close()


def test_device_pointer_api_matches_host_api():
  import torch
  sc = S.Scenario(floor_plan=S.small_plan())
  cp = sc.compiled()
  B = 4
  env_h = S.make_env(sc, n_envs=B, plans=cp)
  env_d = S.make_env(sc, n_envs=B, plans=cp)
  try:
    D = env_h.observation_spec().shape[0]
    dev = torch.device("cuda:0")
    obs = torch.zeros(B, D, device=dev)
    rew = torch.zeros(B, device=dev)
    st = torch.zeros(B, dtype=torch.int32, device=dev)
    dis = torch.zeros(B, device=dev)
    ts = env_h.reset()
    env_d.reset_device(obs, rew, st, dis)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(obs.cpu().numpy(), ts.observation)
    rng = np.random.default_rng(0)
    for _ in range(10):
      a = rng.uniform(-1, 1, (B, 2)).astype(np.float32)
      ts = env_h.step(a)
      env_d.step_device(torch.from_numpy(a).to(dev), obs, rew, st, dis)
      torch.cuda.synchronize()
      np.testing.assert_array_equal(obs.cpu().numpy(), ts.observation)
      np.testing.assert_array_equal(rew.cpu().numpy(), ts.reward)
      np.testing.assert_array_equal(st.cpu().numpy(), ts.step_type)
  finally:
    env_h.close()
    env_d.close()


@pytest.mark.parametrize("path", list(PATHS))
def test_stochastic_convection_replay_matches_oracle(path):
  """Shipped calibrated setting p=1, distance=5, seed=5 (sim_config.gin:37-39): the
  host-drawn gather map applied on the GPU equals the oracle's in-place swaps."""
  sc = S.Scenario(floor_plan=S.small_plan(), convection=(1.0, 5, 5))
  cp = sc.compiled()
  B = 2
  env = S.make_env(sc, n_envs=B, plans=cp, kernel_path=PATHS[path])
  try:
    oracles = [S.make_oracle(sc, cp) for _ in range(B)]
    ts = env.reset()
    for o in oracles:
      o.reset()
    rng = np.random.default_rng(3)
    exact = 0
    for step in range(12):
      a = rng.uniform(-1, 1, (B, 2)).astype(np.float32)
      ts = env.step(a)
      for b, o in enumerate(oracles):
        o.step(a[b])
      st = _compare_step(env, oracles, ts, B, cp, step)
      exact += st["temp_exact"]
    assert exact >= B          # first step: bit-identical field, permutation included
    # convection really moved temperatures around: consecutive room CVs differ
    temp = env.handle.download("temp", (B, cp.height, cp.width))
    assert np.abs(np.diff(temp[0][5:10, 5:15], axis=1)).max() > 0
  finally:
    env.close()


def test_gauss_seidel_with_stochastic_convection_matches_oracle():
  """The legacy config runs the fp64 Gauss-Seidel solver WITH StochasticConvectionSimulator
  (sim_config_legacy.gin:39-41, 101, 182): the gather map is applied after the solve, before
  the write-back and the zone sums."""
  sc = S.Scenario(floor_plan=S.small_plan(), convection=(1.0, 5, 5))
  cp = sc.compiled()
  B = 2
  env = S.make_env(sc, n_envs=B, plans=cp, solver="gauss_seidel")
  try:
    oracles = [S.make_oracle(sc, cp, solver="gs") for _ in range(B)]
    ts = env.reset()
    for o in oracles:
      o.reset()
    rng = np.random.default_rng(4)
    for step in range(8):
      a = rng.uniform(-1, 1, (B, 2)).astype(np.float32)
      ts = env.step(a)
      t64 = env.handle.download("temp64", (B, cp.height, cp.width))
      for b, o in enumerate(oracles):
        ots = o.step(a[b])
        np.testing.assert_allclose(ts.observation[b], ots[3], rtol=RTOL, atol=2e-5)
        np.testing.assert_allclose(ts.reward[b], ots[1], rtol=RTOL, atol=1e-6)
        np.testing.assert_allclose(t64[b], o.temp, rtol=1e-9)
        if step == 0:
          np.testing.assert_array_equal(t64[b], o.temp)     # permutation included, bit for bit
    assert np.abs(np.diff(t64[0][5:10, 5:15], axis=1)).max() > 0
  finally:
    env.close()


@pytest.mark.parametrize("plan_name", ["tf_test_10x9", "small_24x34", "odd_23x31", "rand_64x96"])
def test_gauss_seidel_fd_step_bit_exact_fp64(plan_name):
  """the legacy fp64 in-place raster Gauss-Seidel sweep
  (simulator.py:98-316) as anti-diagonal wavefronts, bit-identical in fp64."""
  from oracle import gs_solver
  plan, bfw = _plans()[plan_name]
  sc = S.Scenario(floor_plan=plan, buffer_from_walls=bfw, cv_size_cm=10.0 if "rand" in plan_name else 20.0)
  cp = sc.compiled()
  B = 3
  env = S.make_env(sc, n_envs=B, plans=cp, solver="gauss_seidel")
  try:
    env.reset()
    rng = np.random.default_rng(5)
    H, W, Z = cp.height, cp.width, env.building.n_zones
    temp = rng.uniform(285, 300, (B, H, W))
    qcv = rng.uniform(-50, 400, (B, Z))
    ambient = rng.uniform(270, 300, B)
    conv = rng.uniform(5, 100, B)
    env.handle.upload("temp64", temp)
    env.handle.upload("q_cv64", qcv)
    env.handle.fd_step(ambient, conv)
    got = env.handle.download("temp64", (B, H, W))
    sweeps = env.handle.download("n_sweeps", (B,))
    gs = gs_solver.GaussSeidel(S.oracle_plan(cp, sc.floor_height_cm), sc.time_step_sec,
                               sc.convergence_threshold, sc.iteration_limit)
    for b in range(B):
      want, n, _, _ = gs.fd_step(temp[b], _dense_q(cp, qcv[b]), ambient[b], conv[b])
      assert sweeps[b] == n, (b, sweeps[b], n)
      np.testing.assert_array_equal(got[b], want, err_msg=f"env {b}")
    np.testing.assert_array_equal(env.handle.download("temp", (B, H, W)), got.astype(np.float32))
  finally:
    env.close()


def test_gauss_seidel_env_matches_unmodified_reference_fixture():
  """Whole Environment.step() with the GS solver against a rollout of the
  100 %-unmodified reference stack (tests/golden/ref_env_gs.npz)."""
  import os
  g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_env_gs.npz"))
  sc = S.Scenario(floor_plan=g["floor_plan"].astype(np.int64))
  cp = sc.compiled()
  env = S.make_env(sc, n_envs=2, plans=cp, solver="gauss_seidel")
  try:
    ts = env.reset()
    np.testing.assert_allclose(ts.observation[0], g["observations"][0], rtol=RTOL, atol=2e-5)
    for i, a in enumerate(g["actions"]):
      ts = env.step(np.stack([a, a]))
      np.testing.assert_allclose(ts.observation[1], g["observations"][i + 1], rtol=RTOL, atol=2e-5,
                                 err_msg=f"obs step {i}")
      np.testing.assert_allclose(ts.reward[0], g["rewards"][i + 1], rtol=RTOL, atol=1e-6)
      zm = env.handle.download("zone_mean", (2, env.building.n_zones))
      np.testing.assert_allclose(zm[0, :cp.n_zones], g["zone_temps"][i + 1], rtol=RTOL)
    t64 = env.handle.download("temp64", (2, cp.height, cp.width))
    np.testing.assert_allclose(t64[0], g["final_temp"], rtol=1e-6)
  finally:
    env.close()


def test_pipelined_step_equals_single_launch_step_and_timing():
  """SBX_OPT_PIPELINE_CHUNKS: cutting the step into shares of the batch that overlap on
  two streams must not change a single bit; sbx_timing_* reports every step / solve."""
  B = 96
  wl = workloads.randomized(B, seed=5, n_layouts=24)
  envs = []
  try:
    for chunks in (1, 3):
      env, _ = workloads.make_randomized_env(B, workload=wl, episode_steps=16, histogram=True)
      env.handle.set_option(_lib.OPT_PIPELINE_CHUNKS, chunks)
      envs.append(env)
    rng = np.random.default_rng(1)
    ts = [e.reset() for e in envs]
    np.testing.assert_array_equal(ts[0].observation, ts[1].observation)
    for e in envs:
      e.handle.timing_begin()
    for _ in range(6):
      a = rng.uniform(-1, 1, (B, 2)).astype(np.float32)
      ts = [e.step(a) for e in envs]
      np.testing.assert_array_equal(ts[0].observation, ts[1].observation)
      np.testing.assert_array_equal(ts[0].reward, ts[1].reward)
    t = [e.handle.timing_end() for e in envs]
    assert [x.n_steps for x in t] == [6, 6]
    assert [x.n_solve_launches for x in t] == [6, 18]
    assert [x.n_chunks for x in t] == [1, 3]
    assert all(0.0 < x.solve_ms <= x.step_ms * 1.5 for x in t)
    np.testing.assert_array_equal(envs[0].handle.download("temp", (B, 64, 96)),
                                  envs[1].handle.download("temp", (B, 64, 96)))
  finally:
    for e in envs:
      e.close()


def test_energy_carbon_reward_function_matches_oracle():
  """SetpointEnergyCarbonRewardFunction (setpoint_energy_carbon_reward.py:127-190, SURVEY
  section 8f rank 4): productivity - w_e * cost - w_c * carbon cost, shifted and scaled."""
  sc = S.Scenario(floor_plan=S.small_plan(), reward="energy_carbon", occupancy="step",
                  start="2023-07-06 07:30:00")
  cp = sc.compiled()
  B = 2
  env = S.make_env(sc, n_envs=B, plans=cp)
  try:
    oracles = [S.make_oracle(sc, cp) for _ in range(B)]
    env.reset()
    for o in oracles:
      o.reset()
    rng = np.random.default_rng(3)
    seen = set()
    for step in range(40):
      a = rng.uniform(-1, 1, (B, 2)).astype(np.float32)
      ts = env.step(a)
      for b, o in enumerate(oracles):
        ots = o.step(a[b])
        np.testing.assert_allclose(ts.reward[b], ots[1], rtol=RTOL, atol=1e-6,
                                   err_msg=f"reward step {step} env {b}")
        np.testing.assert_allclose(ts.observation[b], ots[3], rtol=RTOL, atol=2e-5)
        seen.add(round(float(ots[1]), 3))
    assert len(seen) > 5          # the reward actually varies over the rollout
  finally:
    env.close()


def test_gauss_seidel_on_the_legacy_rectangular_building():
  """The deprecated rectangular `Building` (3 x 3 rooms of 20 x 30 CVs)
  through legacy_building(): fp64 Gauss-Seidel solve bit-identical to the oracle."""
  from oracle import gs_solver
  cp = floorplan.legacy_building(
      20.0, (20, 30), (3, 3), floorplan.MaterialProperties(50.0, 700.0, 1.0),
      floorplan.MaterialProperties(5.0, 800.0, 1800.0), floorplan.MaterialProperties(5.0, 800.0, 3000.0))
  sc = S.Scenario(floor_plan=np.zeros((4, 4), dtype=np.int64), cv_size_cm=20.0)
  B = 2
  env = S.make_env(sc, n_envs=B, plans=cp, solver="gauss_seidel")
  try:
    env.reset()
    rng = np.random.default_rng(8)
    H, W, Z = cp.height, cp.width, env.building.n_zones
    assert (H, W, Z) == (68, 98, 9)
    temp = rng.uniform(285, 300, (B, H, W))
    qcv = rng.uniform(-50, 400, (B, Z))
    ambient = rng.uniform(270, 300, B)
    conv = rng.uniform(5, 100, B)
    env.handle.upload("temp64", temp)
    env.handle.upload("q_cv64", qcv)
    env.handle.fd_step(ambient, conv)
    got = env.handle.download("temp64", (B, H, W))
    sweeps = env.handle.download("n_sweeps", (B,))
    gs = gs_solver.GaussSeidel(S.oracle_plan(cp, sc.floor_height_cm), sc.time_step_sec,
                               sc.convergence_threshold, sc.iteration_limit)
    for b in range(B):
      want, n, _, _ = gs.fd_step(temp[b], _dense_q(cp, qcv[b]), ambient[b], conv[b])
      assert sweeps[b] == n
      np.testing.assert_array_equal(got[b], want)
  finally:
    env.close()


def test_legacy_hvac_on_the_legacy_building_changes_names_not_numbers():
  """The deprecated `Hvac(zone_coordinates, ...)` (simulator/hvac.py:35-123) on
  the rectangular Building: devices `vav_<i>_<j>`, zone ids `zone_id_(i, j)`; the same devices and
  algebra as FloorPlanBasedHvac, so a rollout gives the same numbers."""
  cp = floorplan.legacy_building(
      20.0, (8, 10), (2, 3), floorplan.MaterialProperties(50.0, 700.0, 1.0),
      floorplan.MaterialProperties(5.0, 800.0, 1800.0), floorplan.MaterialProperties(5.0, 800.0, 3000.0))
  sc = S.Scenario(floor_plan=np.zeros((4, 4), dtype=np.int64), cv_size_cm=20.0, occupancy="step",
                  start="2023-07-06 08:30:00")
  coords = [(i, j) for i in range(2) for j in range(3)]
  outs = {}
  for name, kw in (("floor_plan", {}), ("legacy", dict(legacy_hvac_coordinates=coords))):
    env = S.make_env(sc, n_envs=2, plans=cp, solver="gauss_seidel", **kw)
    try:
      env.reset()
      rng = np.random.default_rng(4)
      rows = []
      for _ in range(12):
        ts = env.step(rng.uniform(-1, 1, (2, 2)).astype(np.float32))
        rows.append((ts.observation.copy(), ts.reward.copy()))
      outs[name] = (rows, list(env.field_names), env.building.zone_ids)
    finally:
      env.close()
  for (oa, ra), (ob, rb) in zip(outs["floor_plan"][0], outs["legacy"][0]):
    np.testing.assert_array_equal(oa, ob)
    np.testing.assert_array_equal(ra, rb)
  assert "vav_0_0_zone_air_temperature_sensor" in outs["legacy"][1]
  assert "vav_room_1_zone_air_temperature_sensor" in outs["floor_plan"][1]
  assert outs["legacy"][2][:2] == ["zone_id_(0, 0)", "zone_id_(0, 1)"]


def test_gauss_seidel_on_a_grid_larger_than_shared_memory():
  """SURVEY row a9 at the legacy config's scale (sim_config_legacy.gin:182, 208 runs the
  Gauss-Seidel simulator on the 744x1004 plan): grids whose fp64 planes do not fit one SM run
  the same anti-diagonal wavefront on the field in global memory.  136x200 = 27 200 CVs here
  (the shared-memory variant ends at ~13 k): field bit-identical to the oracle's sequential
  raster sweep, same sweep count, and a short Environment rollout within 1e-4."""
  from oracle import gs_solver
  plan = S.small_plan(136, 200)
  plan[40, 3:197] = 1
  plan[3:133, 66] = 1
  plan[3:133, 131] = 1
  sc = S.Scenario(floor_plan=plan, buffer_from_walls=2, occupancy="step", start="2023-07-06 08:40:00")
  cp = sc.compiled()
  B = 2
  env = S.make_env(sc, n_envs=B, plans=cp, solver="gauss_seidel")
  try:
    env.reset()
    rng = np.random.default_rng(12)
    H, W, Z = cp.height, cp.width, env.building.n_zones
    temp = rng.uniform(289, 295, (B, H, W))
    qcv = rng.uniform(-20, 200, (B, Z))
    ambient = rng.uniform(275, 290, B)
    conv = rng.uniform(10, 60, B)
    env.handle.upload("temp64", temp)
    env.handle.upload("q_cv64", qcv)
    env.handle.fd_step(ambient, conv)
    got = env.handle.download("temp64", (B, H, W))
    sweeps = env.handle.download("n_sweeps", (B,))
    gs = gs_solver.GaussSeidel(S.oracle_plan(cp, sc.floor_height_cm), sc.time_step_sec,
                               sc.convergence_threshold, sc.iteration_limit)
    for b in range(B):
      want, n, _, _ = gs.fd_step(temp[b], _dense_q(cp, qcv[b]), ambient[b], conv[b])
      assert sweeps[b] == n and n >= 2
      np.testing.assert_array_equal(got[b], want)
  finally:
    env.close()
  # and through Environment.step: 3 free-running steps against the oracle environment
  env = S.make_env(sc, n_envs=1, plans=cp, solver="gauss_seidel")
  try:
    oracle = S.make_oracle(sc, cp, solver="gs")
    env.reset()
    oracle.reset()
    rng = np.random.default_rng(13)
    for step in range(3):
      a = rng.uniform(-1, 1, (1, 2)).astype(np.float32)
      ts = env.step(a)
      ots = oracle.step(a[0])
      np.testing.assert_allclose(ts.observation[0], ots[3], rtol=RTOL, atol=2e-5)
      np.testing.assert_allclose(ts.reward[0], ots[1], rtol=RTOL, atol=5e-5)
      assert int(env.handle.download("n_sweeps", (1,))[0]) == oracle.info["n_sweeps"]
  finally:
    env.close()
