"""Oracle: fp32 Jacobi control-volume diffusion (TEST INFRASTRUCTURE ONLY).

Restates, op for op and in the reference's evaluation order, the TensorFlow
eager arithmetic of `TFSimulator.update_temperature_estimates`
(/root/reference/smart_control/simulator/tf_simulator.py:573-853) with NumPy
float32 arrays, plus the statics built in `TFSimulator.__init__`
(tf_simulator.py:531-559) and the convergence loop of
`Simulator.finite_differences_timestep` (simulator.py:318-371).

Every TF op in that function is an IEEE-754 fp32 elementwise add / multiply /
divide / subtract / abs / select, each rounded separately (eager mode, nothing
fused), so NumPy fp32 in the same order reproduces it bit for bit.
"""

from __future__ import annotations

import dataclasses
from typing import List, Sequence, Tuple

import numpy as np

# CV classes (tf_simulator.py:43-86).  Codes are this repo's own.
EXTERIOR = 0
INTERIOR = 1
EDGE_TOP = 2      # missing neighbour (i-1, j)          tf_simulator.py:235
EDGE_BOTTOM = 3   # missing neighbour (i+1, j)          tf_simulator.py:237
EDGE_LEFT = 4     # missing neighbour (i, j-1)          tf_simulator.py:239
EDGE_RIGHT = 5    # missing neighbour (i, j+1)          tf_simulator.py:241
CORNER_TL = 6     # has (i+1, j), (i, j+1)              tf_simulator.py:209
CORNER_BL = 7     # has (i-1, j), (i, j+1)              tf_simulator.py:211
CORNER_TR = 8     # has (i+1, j), (i, j-1)              tf_simulator.py:213
CORNER_BR = 9     # has (i-1, j), (i, j-1)              tf_simulator.py:215

# (horizontal scale -> U, vertical scale -> V); tf_simulator.py:57-85, 144-174
_SCALE = {
    EXTERIOR: (1.0, 1.0), INTERIOR: (1.0, 1.0),
    EDGE_TOP: (1.0, 0.5), EDGE_BOTTOM: (1.0, 0.5),
    EDGE_LEFT: (0.5, 1.0), EDGE_RIGHT: (0.5, 1.0),
    CORNER_TL: (0.5, 0.5), CORNER_BL: (0.5, 0.5),
    CORNER_TR: (0.5, 0.5), CORNER_BR: (0.5, 0.5),
}
# which sides face ambient: (left, right, top, bottom); tf_simulator.py:354-379
_SIDES = {
    EDGE_TOP: (0, 0, 1, 0), EDGE_BOTTOM: (0, 0, 0, 1),
    EDGE_LEFT: (1, 0, 0, 0), EDGE_RIGHT: (0, 1, 0, 0),
    CORNER_TL: (1, 0, 1, 0), CORNER_TR: (0, 1, 1, 0),
    CORNER_BL: (1, 0, 0, 1), CORNER_BR: (0, 1, 0, 1),
}


@dataclasses.dataclass
class OraclePlan:
  """Static per-building arrays, as `FloorPlanBasedBuilding` holds them.

  Mirrors building.py:634-766: `exterior_space` is the cv_type ==
  'exterior_space' mask (building.py:290-297), conductivity / heat_capacity /
  density the per-CV material arrays (building.py:727-749), `diffusers` the
  per-CV weights 1/n_diffusers(zone) (building.py:349-351), `rooms` the
  `_room_dict` entries whose key starts with "room" in dict order
  (building.py:863-871), each as (name, rows, cols).
  """

  exterior_space: np.ndarray          # bool [H, W]
  conductivity: np.ndarray            # f64 [H, W]
  heat_capacity: np.ndarray           # f64 [H, W]
  density: np.ndarray                 # f64 [H, W]
  diffusers: np.ndarray               # f64 [H, W]
  rooms: List[Tuple[str, np.ndarray, np.ndarray]]
  cv_size_cm: float
  floor_height_cm: float

  @property
  def shape(self) -> Tuple[int, int]:
    return self.exterior_space.shape


def neighbor_counts(exterior_space: np.ndarray) -> np.ndarray:
  """len(neighbors[i][j]) of building.py:794-813 as an int array.

  Exterior-space CVs get no neighbours at all (building.py:805-806); for any
  other CV the in-bounds 4-neighbours that are not exterior space count.
  """
  ext = exterior_space.astype(bool)
  non_ext = (~ext).astype(np.int32)
  p = np.pad(non_ext, 1, constant_values=0)
  n = p[:-2, 1:-1] + p[2:, 1:-1] + p[1:-1, :-2] + p[1:-1, 2:]
  n = np.where(ext, 0, n)
  return n


def classify(exterior_space: np.ndarray) -> np.ndarray:
  """CV class map; restates classify_cv / get_cv_mapping tf_simulator.py:180-280."""
  ext = exterior_space.astype(bool)
  h, w = ext.shape
  ok = np.pad(~ext, 1, constant_values=False)
  up = ok[:-2, 1:-1]      # (i-1, j) is a neighbour
  down = ok[2:, 1:-1]     # (i+1, j)
  left = ok[1:-1, :-2]    # (i, j-1)
  right = ok[1:-1, 2:]    # (i, j+1)
  n = neighbor_counts(ext)
  cls = np.full((h, w), -1, dtype=np.int8)
  cls[n <= 1] = EXTERIOR
  cls[n == 4] = INTERIOR
  e = n == 3
  cls[e & ~up] = EDGE_TOP
  cls[e & ~down] = EDGE_BOTTOM
  cls[e & ~left] = EDGE_LEFT
  cls[e & ~right] = EDGE_RIGHT
  c = n == 2
  cls[c & down & right] = CORNER_TL
  cls[c & up & right] = CORNER_BL
  cls[c & down & left] = CORNER_TR
  cls[c & up & left] = CORNER_BR
  cls[ext] = EXTERIOR
  if (cls < 0).any():
    i, j = np.argwhere(cls < 0)[0]
    # tf_simulator.py:219-221 raises for opposite-neighbour 2-patterns.
    raise ValueError(f"wasn't able to determine which corner the CV {(i, j)} is.")
  return cls


def _shift_left(x, pad):   # tf_simulator.py:467-472: out[i, j] = x[i, j+1]
  return np.concatenate([x[:, 1:], np.full((x.shape[0], 1), pad, x.dtype)], 1)


def _shift_right(x, pad):  # tf_simulator.py:459-464: out[i, j] = x[i, j-1]
  return np.concatenate([np.full((x.shape[0], 1), pad, x.dtype), x[:, :-1]], 1)


def _shift_up(x, pad):     # tf_simulator.py:475-480: out[i, j] = x[i+1, j]
  return np.concatenate([x[1:, :], np.full((1, x.shape[1]), pad, x.dtype)], 0)


def _shift_down(x, pad):   # tf_simulator.py:483-488: out[i, j] = x[i-1, j]
  return np.concatenate([np.full((1, x.shape[1]), pad, x.dtype), x[:-1, :]], 0)


class TFJacobi:
  """Statics + one sweep + the convergence loop, fp32."""

  def __init__(self, plan: OraclePlan, time_step_sec: float,
               convergence_threshold: float, iteration_limit: int):
    self.plan = plan
    self.dt = np.float32(time_step_sec)                      # :788
    self.threshold = convergence_threshold
    self.iteration_limit = int(iteration_limit)
    self.cls = classify(plan.exterior_space)
    self.exterior_mask = self.cls == EXTERIOR                 # :561-571
    dx = plan.cv_size_cm / 100.0                              # :547
    shape = plan.shape
    # get_cv_dimension_tensors :283-329 (fp32 array filled with a Python float;
    # boundary values computed in fp64 then stored to fp32).
    u = np.full(shape, dx, dtype=np.float32)
    v = np.full(shape, dx, dtype=np.float32)
    for code, (hs, vs) in _SCALE.items():
      if code in (EXTERIOR, INTERIOR):
        continue
      m = self.cls == code
      u[m] = dx * hs
      v[m] = dx * vs
    self.u, self.v = u, v
    # get_oriented_conductivity_tensors :401-456
    k = np.asarray(plan.conductivity, dtype=np.float64)
    kl, kr, kt, kb = k.copy(), k.copy(), k.copy(), k.copy()
    self.side_masks = {}
    for code, (sl, sr, st, sb) in _SIDES.items():
      m = self.cls == code
      if sl:
        kl[m] = 0.0
      if sr:
        kr[m] = 0.0
      if st:
        kt[m] = 0.0
      if sb:
        kb[m] = 0.0
    self.k_left = kl.astype(np.float32)
    self.k_right = kr.astype(np.float32)
    self.k_top = kt.astype(np.float32)
    self.k_bottom = kb.astype(np.float32)
    self._left_m = np.isin(self.cls, [EDGE_LEFT, CORNER_TL, CORNER_BL])
    self._right_m = np.isin(self.cls, [EDGE_RIGHT, CORNER_TR, CORNER_BR])
    self._top_m = np.isin(self.cls, [EDGE_TOP, CORNER_TL, CORNER_TR])
    self._bottom_m = np.isin(self.cls, [EDGE_BOTTOM, CORNER_BL, CORNER_BR])
    self.density = np.asarray(plan.density).astype(np.float32)        # :615
    self.heat_capacity = np.asarray(plan.heat_capacity).astype(np.float32)
    self.z = np.float32(plan.floor_height_cm / 100.0)                  # :619

  def _h_tensors(self, h: float):
    """get_oriented_convection_coefficient_tensors :332-398 (rebuilt per sweep)."""
    shape = self.plan.shape
    out = []
    for m in (self._left_m, self._right_m, self._top_m, self._bottom_m):
      a = np.zeros(shape, dtype=np.float32)
      a[m] = h
      out.append(a)
    return out

  def sweep(self, t_est: np.ndarray, temp: np.ndarray, input_q: np.ndarray,
            ambient_temperature: float, convection_coefficient: float):
    """One Jacobi sweep.  Returns (new estimate fp32 [H,W], max_delta).

    `temp` is building.temp (previous step), `t_est` the running estimate.
    Line numbers refer to tf_simulator.py.
    """
    f32 = np.float32
    t = np.asarray(t_est, dtype=f32)                        # :611-612
    t_minus = np.asarray(temp, dtype=f32)                   # :613
    q = np.asarray(input_q, dtype=f32)                      # :614
    rho, c, z, dt = self.density, self.heat_capacity, self.z, self.dt
    u, v = self.u, self.v
    h_l, h_r, h_t, h_b = self._h_tensors(convection_coefficient)  # :768-777
    t_inf = f32(ambient_temperature)                        # :785
    # padding value is the Python float ambient, cast to the array dtype :636-646
    t_left = _shift_left(t, f32(ambient_temperature))
    t_right = _shift_right(t, f32(ambient_temperature))
    t_above = _shift_down(t, f32(ambient_temperature))
    t_below = _shift_up(t, f32(ambient_temperature))
    uz = z * u                                              # :791
    vz = z * v                                              # :792
    k1 = self.k_left / u                                    # :795
    k3 = self.k_right / u                                   # :796
    k2 = self.k_bottom / v                                  # :797
    k4 = self.k_top / v                                     # :798
    # denominator :669-690
    d1 = k1 + k3
    d1 = d1 + h_l
    d1 = d1 + h_r
    d1 = vz * d1
    d2 = k2 + k4
    d2 = d2 + h_b
    d2 = d2 + h_t
    d2 = uz * d2
    d3 = rho * u
    d3 = d3 * v
    d3 = d3 * c
    d3 = z * d3
    d3 = d3 * c
    d3 = d3 / dt
    den = d1 + d2
    den = den + d3
    # numerator :719-754
    a1 = k1 * t_left
    a3 = k3 * t_right
    a2 = k2 * t_below
    a4 = k4 * t_above
    hl_t = t_inf * h_l
    hr_t = t_inf * h_r
    ha_t = t_inf * h_t
    hb_t = t_inf * h_b
    n1 = a1 + a3
    n1 = n1 + hl_t
    n1 = n1 + hr_t
    n1 = vz * n1
    n2 = a2 + a4
    n2 = n2 + hb_t
    n2 = n2 + ha_t
    n2 = uz * n2
    n3 = rho * u
    n3 = n3 * v
    n3 = n3 * c
    n3 = z * n3
    n3 = n3 * c
    n3 = n3 * t_minus
    n3 = n3 / dt
    num = n1 + n2
    num = num + n3
    num = num + q
    with np.errstate(divide="ignore", invalid="ignore"):
      new = num / den                                       # :843
    new = np.where(self.exterior_mask, t_inf, new)          # :847-849
    assert new.dtype == np.float32
    max_delta = np.max(np.abs(new - t))                     # :851-853
    return new, max_delta

  def fd_step(self, temp: np.ndarray, input_q: np.ndarray,
              ambient_temperature: float, convection_coefficient: float):
    """finite_differences_timestep simulator.py:318-371.

    Returns (new temp fp32 [H,W], n_sweeps, converged, last max_delta).
    """
    est = np.array(temp, copy=True)
    converged = False
    n = 0
    max_delta = np.float32(0)
    for _ in range(self.iteration_limit):
      est, max_delta = self.sweep(est, temp, input_q, ambient_temperature,
                                  convection_coefficient)
      n += 1
      if max_delta <= self.threshold:                       # simulator.py:362
        converged = True
        break
    return est, n, converged, float(max_delta)
