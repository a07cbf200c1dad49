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
