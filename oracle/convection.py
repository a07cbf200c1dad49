"""Oracle: stochastic in-room convection (TEST INFRASTRUCTURE ONLY).

Restates StochasticConvectionSimulator
(/root/reference/smart_control/simulator/stochastic_convection_simulator.py:34-145):
after every FD step, inside each room, a CV is selected with probability p and
swapped with a random partner of the same room within `distance`; the swaps are
shuffled and applied sequentially.  The reference draws from Python's
process-global `random` (seeded in __init__, :59-60); pass `rng=random.Random(s)`
to give an environment its own stream (what a separate reference process has).
"""

from __future__ import annotations

import collections
import copy
import random as _random


class StochasticConvectionSimulator:

  def __init__(self, p: float, distance: int, seed=None, rng=None):
    self._p = p
    self._distance = distance
    self._cache = collections.defaultdict(lambda: {})
    self._rng = rng if rng is not None else _random
    if seed is not None:
      self._rng.seed(seed)                                         # :59-60

  def apply_convection(self, room_dict, temp) -> None:             # :62-82
    p, distance = self._p, self._distance
    if p == 0 or distance == 0:
      return
    for k, v in room_dict.items():
      if k in ("exterior_space", "interior_wall"):
        continue
      if distance == -1 and p == 1:
        self._shuffle_no_max_dist(v, temp)
      else:
        self._shuffle_max_dist(p, v, distance, temp)

  def _shuffle_no_max_dist(self, v, temp):                         # :84-99
    v = copy.deepcopy(v)
    vals = {cv: temp[cv[0], cv[1]] for cv in v}
    v_shuffle = copy.deepcopy(v)
    self._rng.shuffle(v_shuffle)
    for i, cv in enumerate(v_shuffle):
      temp[cv[0], cv[1]] = vals[v[i]]

  def _shuffle_max_dist(self, p, v, max_dist, temp):               # :101-145
    if max_dist == -1:
      max_dist = 1000
    in_v = {val: True for val in v}
    swap_list = []
    for val in v:
      if self._rng.uniform(0, 1) > p:
        continue
      if max_dist in self._cache and val in self._cache[max_dist]:
        candidates = self._cache[max_dist][val]
      else:
        candidates = []
        for c0 in range(val[0] - max_dist, val[0] + max_dist):     # excludes +max_dist
          for c1 in range(val[1] - max_dist, val[1] + max_dist):
            c2 = (c0, c1)
            if c2 not in in_v:
              continue
            if (val[0] - c0) ** 2 + (val[1] - c1) ** 2 <= max_dist:  # dist^2 <= d
              candidates.append(c2)
        self._cache[max_dist][val] = candidates
      swap_list.append((val, self._rng.choice(candidates)))
    self._rng.shuffle(swap_list)
    for a, b in swap_list:
      t = temp[a[0], a[1]]
      temp[a[0], a[1]] = temp[b[0], b[1]]
      temp[b[0], b[1]] = t
