"""Host side of the stochastic in-room convection model.

The reference mutates building.temp after every FD step with a sequence of
random, sequentially dependent swaps drawn from Python's Mersenne Twister
(simulator/stochastic_convection_simulator.py:62-145 of the reference).  Swapping
temperatures is a permutation of the grid, so the host draws the swaps with the
same generator calls in the same order and composes them into one gather index
per control volume; the GPU applies `T_new[i] = T_old[src[i]]` between the
diffusion solve and the zone reductions (libsbx field SBX_F_CONVECTION_PERM).

This is the exact-replay mode.  It costs a Python loop
over every room CV per building per step, so it is meant for parity and small
batches.  Large throughput batches use `mode="device"`: the same (p, distance, seed)
drive a counter-based generator on the GPU and a swap pattern that needs no
sequential pass (libsbx `sbx_set_device_convection`, kernel k_convect_reduce): in-room
moves of squared length <= distance, participation probability p, zone and grid sums
unchanged -- the reference model's invariants, not its Mersenne-Twister sequence.
"""

from __future__ import annotations

import random as _random
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


class StochasticConvectionSimulator:
  """Same constructor as the reference: (p, distance, seed).

  Every environment of a batch gets its own `random.Random(seed)` stream -- what
  B separate reference processes built from the same config would have."""

  def __init__(self, p: float, distance: int, seed: Optional[int], mode: str = "replay"):
    if mode not in ("replay", "device"):
      raise ValueError("mode must be 'replay' (exact host replay) or 'device' (device RNG)")
    self._p = p
    self._distance = distance
    self._seed = seed
    self.mode = mode
    # swap candidates per (room list, max distance, CV).  The reference keeps one simulator
    # -- and so one cache -- per building (stochastic_convection_simulator.py:120-134); a
    # batch with one plan per env must not reuse env 0's candidates for another plan's room
    # at the same coordinates, so the cache is keyed by the identity of the room list.
    self._cache: Dict[Tuple[int, int], Dict[Tuple[int, int], List[Tuple[int, int]]]] = {}

  def make_stream(self) -> _random.Random:
    return _random.Random(self._seed) if self._seed is not None else _random.Random()

  def gather_index(self, rooms: Sequence[Sequence[Tuple[int, int]]], shape,
                   rng: _random.Random) -> np.ndarray:
    """Flat source index per CV for one application of the convection model.

    `rooms`: the room CV lists in room-dict order (building.py:863-871).  Applying
    the reference's swaps to an array that holds each CV's own flat index yields
    exactly the gather map."""
    h, w = shape
    idx = np.arange(h * w, dtype=np.int64).reshape(h, w)
    p, distance = self._p, self._distance
    if p == 0 or distance == 0:
      return idx.ravel().astype(np.int32)
    for v in rooms:
      if distance == -1 and p == 1:
        self._shuffle_no_max_dist(v, idx, rng)
      else:
        self._shuffle_max_dist(p, v, distance, idx, rng, cache_key=id(v))
    return idx.ravel().astype(np.int32)

  @staticmethod
  def _shuffle_no_max_dist(v, arr, rng):
    v = list(v)
    vals = {cv: arr[cv[0], cv[1]] for cv in v}
    shuffled = list(v)
    rng.shuffle(shuffled)
    for i, cv in enumerate(shuffled):
      arr[cv[0], cv[1]] = vals[v[i]]

  def _shuffle_max_dist(self, p, v, max_dist, arr, rng, cache_key=0):
    if max_dist == -1:
      max_dist = 1000
    members = set(v)
    cache = self._cache.setdefault((cache_key, max_dist), {})
    swaps = []
    for val in v:
      if rng.uniform(0, 1) > p:
        continue
      cand = cache.get(val)
      if cand is None:
        cand = []
        for c0 in range(val[0] - max_dist, val[0] + max_dist):
          for c1 in range(val[1] - max_dist, val[1] + max_dist):
            if (c0, c1) in members and (val[0] - c0) ** 2 + (val[1] - c1) ** 2 <= max_dist:
              cand.append((c0, c1))
        cache[val] = cand
      swaps.append((val, rng.choice(cand)))
    rng.shuffle(swaps)
    for a, b in swaps:
      t = arr[a[0], a[1]]
      arr[a[0], a[1]] = arr[b[0], b[1]]
      arr[b[0], b[1]] = t
