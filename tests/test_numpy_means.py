"""SBX_OPT_NUMPY_MEANS: zone and grid means in the summation order of the reference's np.mean
(building.py:845-871, float32 pairwise sums over each room's raster-ordered CVs).

CPU: the summation tree the library builds (sbx_pairwise.cuh, build_pairwise_tables) restated in
Python and pinned on NumPy itself.  GPU: the library's means against np.mean of the field it
returns, bit for bit, on both kernel paths; free-running rollouts against the oracle."""

import os

import numpy as np
import pytest

import sbsim_b200 as sbx
from sbsim_b200 import workloads

f32 = np.float32
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def pairwise_sum(a):
  """NumPy's float32 add.reduce over a contiguous vector: leaves of <= 128 values summed on 8
  strided accumulators, combined ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), tail one by one; longer
  vectors split at n/2 rounded down to a multiple of 8."""
  n = len(a)
  if n < 8:
    r = f32(0.0)
    for x in a:
      r = f32(r + x)
    return r
  if n <= 128:
    r = [f32(a[j]) for j in range(8)]
    i = 8
    while i < n - n % 8:
      for j in range(8):
        r[j] = f32(r[j] + a[i + j])
      i += 8
    res = f32(f32(f32(r[0] + r[1]) + f32(r[2] + r[3])) + f32(f32(r[4] + r[5]) + f32(r[6] + r[7])))
    while i < n:
      res = f32(res + a[i])
      i += 1
    return res
  n2 = n // 2
  n2 -= n2 % 8
  return f32(pairwise_sum(a[:n2]) + pairwise_sum(a[n2:]))


def test_the_restated_tree_is_numpy_s():
  rng = np.random.default_rng(0)
  for n in list(range(1, 300)) + [511, 1000, 4097, 6144, 10007, 93372]:
    a = rng.uniform(280, 300, n).astype(f32)
    assert pairwise_sum(a).tobytes() == np.add.reduce(a).tobytes(), n
    assert f32(pairwise_sum(a) / f32(n)).tobytes() == np.mean(a).tobytes(), n
  g = rng.uniform(280, 300, (64, 96)).astype(f32)       # building.temp.mean(): the grid as one vector
  assert f32(pairwise_sum(g.ravel()) / f32(g.size)).tobytes() == g.mean().tobytes()


def _bits(x):
  return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


def _check_means(env, plans, B, H, W):
  temp = env.handle.download("temp", (B, H, W))
  zm = env.handle.download("zone_mean", (B, env.building.n_zones))
  gm = env.handle.download("global_mean", (B,))
  for b in range(B):
    cp = plans[b] if isinstance(plans, (list, tuple)) else plans
    for zi in range(cp.n_zones):
      want = np.mean(temp[b][cp.zone_id == zi])          # raster order, like _room_dict
      assert _bits(zm[b, zi]) == _bits(want), (b, zi, zm[b, zi], want)
    assert _bits(gm[b]) == _bits(temp[b].mean()), (b, gm[b], temp[b].mean())


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["streaming", "resident", "resident-separate"])
def test_means_are_np_mean_of_the_field_bit_for_bit(path, monkeypatch):
  """resident: k_resident_step<4, true> takes the means itself from the field in shared memory
  (3 launches per step: k_pre, the solve, k_post); resident-separate (SBX_PW_SEPARATE=1) and
  streaming: k_pw_leaves + k_pw_combine on the field in global memory."""
  monkeypatch.delenv("SBX_PW_SEPARATE", raising=False)
  if path == "resident-separate":
    monkeypatch.setenv("SBX_PW_SEPARATE", "1")
  B = 6
  wl = workloads.randomized(B, seed=5, n_layouts=B)
  env, _ = workloads.make_randomized_env(
      B, workload=wl, episode_steps=16, histogram=True, numpy_zone_means=True,
      kernel_path=sbx.PATH_STREAMING if path == "streaming" else sbx.PATH_RESIDENT)
  try:
    env.reset()
    _check_means(env, wl.plans, B, 64, 96)
    rng = np.random.default_rng(9)
    for step in range(4):
      n0 = env.handle.info().kernel_launches
      env.step(rng.uniform(-1, 1, (B, 2)).astype(np.float32))
      if path != "streaming" and step > 0:       # the first step also prepares the plans
        assert env.handle.info().kernel_launches - n0 == (3 if path == "resident" else 5)
      _check_means(env, wl.plans, B, 64, 96)
  finally:
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["streaming", "resident"])
def test_free_running_rollout_against_the_oracle(path):
  """A free-running rollout against the oracle with the reference's means: every zone mean the
  oracle reports is the library's bit for bit, so observations are IDENTICAL and the fields agree
  to the last bit nearly everywhere.  They are not identical for ever: from the second step on
  the reference's zone temperatures are np.float32 scalars, and under NumPy 2's promotion
  rules its VAV algebra then runs in float32 where the library (like the reference under NumPy 1)
  computes in float64 -- the heat per diffuser CV can differ by an fp32 ulp, and with it a
  handful of CVs around the diffusers."""
  from oracle import bench_support
  B, steps = 6, 24
  wl = workloads.randomized(B, seed=5, n_layouts=B)
  env, _ = workloads.make_randomized_env(
      B, workload=wl, episode_steps=steps + 4, histogram=True, numpy_zone_means=True,
      kernel_path=sbx.PATH_RESIDENT if path == "resident" else sbx.PATH_STREAMING)
  try:
    oracles = [bench_support.make_oracle_env(
        wl.plans[b], float(wl.weather_low[b]), float(wl.weather_high[b]), float(wl.convection[b]),
        float(wl.initial_temp[b]), workloads.NORMALIZATION, workloads.HISTOGRAM, steps + 4,
        workloads.DEFAULT_START) for b in range(B)]
    env.reset()
    for o in oracles:
      o.reset()
    rng = np.random.default_rng(9)
    identical = 0
    for step in range(steps):
      a = rng.uniform(-1, 1, (B, 2)).astype(np.float32)
      ts = env.step(a)
      temp = env.handle.download("temp", (B, 64, 96))
      zm = env.handle.download("zone_mean", (B, env.building.n_zones))
      sweeps = env.handle.download("n_sweeps", (B,))
      for b, o in enumerate(oracles):
        ots = o.step(a[b])
        assert sweeps[b] == o.info["n_sweeps"], (step, b)
        diff = temp[b] != o.temp
        identical += int(not diff.any())
        assert diff.mean() <= 0.01, (step, b, int(diff.sum()))
        np.testing.assert_allclose(temp[b], o.temp, rtol=0, atol=1.3e-4, err_msg=f"step {step} env {b}")   # 4 ulp at 290 K
        # the means of the library's own field in the reference's order (the oracle's field may
        # sit an ulp away at a few CVs, see above)
        for zi in range(wl.plans[b].n_zones):
          assert _bits(zm[b, zi]) == _bits(np.mean(temp[b][wl.plans[b].zone_id == zi])), (step, b, zi)
        np.testing.assert_allclose(ts.observation[b], ots[3], rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(ts.reward[b], ots[1], rtol=1e-4, atol=1e-6)
    assert identical >= B * steps // 2, identical      # most env-steps: the whole field, bit for bit
  finally:
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["streaming", "resident"])
@pytest.mark.parametrize("plan_name", ["small_24x34", "odd_23x31", "tf_test_10x9"])
def test_means_on_shared_plans_with_scalar_vectors(path, plan_name):
  """One plan for the whole batch; widths that are not a multiple of 4 (one CV per vector,
  k_resident_step<1>: the resident path runs the separate kernels), with stochastic convection
  replayed on top for the 24x34 plan (the means are taken after the swaps)."""
  import scenarios as S
  plan, bfw = {"small_24x34": (S.small_plan(), 2), "odd_23x31": (S.small_plan(23, 31), 2),
               "tf_test_10x9": (S.TF_TEST_PLAN, 0)}[plan_name]
  sc = S.Scenario(floor_plan=plan, buffer_from_walls=bfw, cv_size_cm=20.0,
                  convection=(0.3, 2, 11) if plan_name == "small_24x34" else None)
  cp = sc.compiled()
  B = 3
  env = S.make_env(sc, n_envs=B, plans=cp, numpy_zone_means=True,
                   kernel_path=sbx.PATH_RESIDENT if path == "resident" else sbx.PATH_STREAMING)
  try:
    env.reset()
    _check_means(env, cp, B, cp.height, cp.width)
    rng = np.random.default_rng(2)
    for _ in range(3):
      env.step(rng.uniform(-1, 1, (B, 2)).astype(np.float32))
      _check_means(env, cp, B, cp.height, cp.width)
  finally:
    env.close()


@pytest.mark.gpu
def test_means_on_the_calibrated_building():
  """744x1004, 126 zones, one shared plan (u32 CV list, 8192 grid leaves, 13 levels)."""
  cal = workloads.load_calibrated(os.path.join(GOLD, "sb1_calibrated.npz"))
  B = 2
  env = workloads.make_calibrated_env(cal, B, episode_steps=8, numpy_zone_means=True)
  try:
    env.reset()
    _check_means(env, cal.plan, B, cal.plan.height, cal.plan.width)
    env.step(np.zeros((B, env.action_spec().shape[-1]), dtype=np.float32))
    _check_means(env, cal.plan, B, cal.plan.height, cal.plan.width)
  finally:
    env.close()
