"""Host-side static compiler: floor plan -> packed per-CV descriptors.

Setup-time only (runs once per plan on the host).  It turns a floor-plan array
(0 interior space, 1 wall, 2 exterior space -- simulator/constants.py:24-40 of
the reference) into the data the CUDA kernels consume: a uint16 descriptor per
control volume (class, material, diffuser flag, zone id), the per-zone CV and
diffuser counts, and the material table.

The geometry semantics follow the reference's setup pipeline so that the same
plan yields the same arrays (paths relative to /root/reference/smart_control/simulator/):
  padding / exterior space / wall shell / interior walls / rooms
      building_utils.py:144-251, 322-373, 254-292, 376-414, 437-482
  exterior-wall enlargement        building.py:183-229, building_utils.py:485-509
  materials, cv types, neighbours  building.py:232-297, 727-749, 794-813
  diffuser placement               building.py:299-353, thermal_diffuser_utils.py:36-262
  CV classes                       tf_simulator.py:180-280
It is written from those semantics with NumPy/SciPy (no OpenCV): connected
components by scipy.ndimage.label, the 3x3-mask L2 chamfer distance <= 2 of
cv2.distanceTransform as a fixed 13-cell structuring element.
"""

from __future__ import annotations

import dataclasses
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
from scipy import ndimage

from sbsim_b200 import _lib

# CV classes (include/sbx.h)
CV_EXTERIOR, CV_INTERIOR = 0, 1
CV_EDGE_TOP, CV_EDGE_BOTTOM, CV_EDGE_LEFT, CV_EDGE_RIGHT = 2, 3, 4, 5
CV_CORNER_TL, CV_CORNER_BL, CV_CORNER_TR, CV_CORNER_BR = 6, 7, 8, 9

MAT_AIR, MAT_INTERIOR_WALL, MAT_EXTERIOR_WALL = 0, 1, 2


@dataclasses.dataclass(frozen=True)
class MaterialProperties:
  """building.py:37-42."""
  conductivity: float
  heat_capacity: float
  density: float


@dataclasses.dataclass
class CompiledPlan:
  height: int
  width: int
  desc: np.ndarray                 # uint16 [H, W]
  material: np.ndarray             # float64 [3, 3]: rows air / interior wall / exterior wall; cols k, c, rho
  cv_size_m: float
  zone_names: List[str]            # room_dict order (building.py:863-871)
  zone_ncv: np.ndarray             # int32 [n_zones]
  zone_ndiff: np.ndarray           # int32 [n_zones]
  obs_zone_order: np.ndarray       # int32 [n_zones]: zones sorted by device id 'vav_<name>'
  # dense views (tests / oracle / diagnostics)
  cv_class: np.ndarray             # int8 [H, W]
  material_id: np.ndarray          # int8 [H, W]
  zone_id: np.ndarray              # int16 [H, W], -1 = none
  diffuser_weight: np.ndarray      # float64 [H, W]
  exterior_space: np.ndarray       # bool [H, W]

  @property
  def n_zones(self) -> int:
    return len(self.zone_names)

  def zone_indices(self, zi: int) -> Tuple[np.ndarray, np.ndarray]:
    """(rows, cols) of zone zi in raster order, as the reference's _room_dict lists them."""
    return np.nonzero(self.zone_id == zi)

  def dense_material(self, col: int) -> np.ndarray:
    return self.material[self.material_id.astype(np.int64), col]


def guarantee_air_padding_in_frame(plan: np.ndarray) -> np.ndarray:
  """building_utils.py:144-219: add an exterior row/column where walls touch the frame."""
  if 1 in plan.shape or 0 in plan.shape:
    raise ValueError("floor plan is a 1 dimensional array")
  if np.any(plan[0, :] == 1):
    plan = np.concatenate((np.full((1, plan.shape[1]), 2, plan.dtype), plan), axis=0)
  if np.any(plan[:, 0] == 1):
    plan = np.concatenate((np.full((plan.shape[0], 1), 2, plan.dtype), plan), axis=1)
  if np.any(plan[-1, :] == 1):
    plan = np.concatenate((plan, np.full((1, plan.shape[1]), 2, plan.dtype)), axis=0)
  if np.any(plan[:, -1] == 1):
    plan = np.concatenate((plan, np.full((plan.shape[0], 1), 2, plan.dtype)), axis=1)
  return plan


_CROSS = ndimage.generate_binary_structure(2, 1)
# cells whose 3x3-mask chamfer L2 distance (0.955 / 1.3693 weights of
# cv2.DIST_L2 with maskSize 3) rounds to <= 2: straight 1 and 2, diagonal 1.
_ENLARGE = np.array([[0, 0, 1, 0, 0],
                     [0, 1, 1, 1, 0],
                     [1, 1, 1, 1, 1],
                     [0, 1, 1, 1, 0],
                     [0, 0, 1, 0, 0]], dtype=bool)


def classify_cvs(exterior_space: np.ndarray) -> np.ndarray:
  """CV class from the 4-neighbourhood (building.py:794-813, tf_simulator.py:180-262)."""
  ext = exterior_space.astype(bool)
  ok = np.pad(~ext, 1, constant_values=False)
  up, down = ok[:-2, 1:-1], ok[2:, 1:-1]
  left, right = ok[1:-1, :-2], ok[1:-1, 2:]
  n = up.astype(np.int8) + down + left + right
  cls = np.full(ext.shape, -1, dtype=np.int8)
  cls[n <= 1] = CV_EXTERIOR
  cls[n == 4] = CV_INTERIOR
  e = n == 3
  cls[e & ~up] = CV_EDGE_TOP
  cls[e & ~down] = CV_EDGE_BOTTOM
  cls[e & ~left] = CV_EDGE_LEFT
  cls[e & ~right] = CV_EDGE_RIGHT
  c = n == 2
  cls[c & down & right] = CV_CORNER_TL
  cls[c & up & right] = CV_CORNER_BL
  cls[c & down & left] = CV_CORNER_TR
  cls[c & up & left] = CV_CORNER_BR
  cls[ext] = CV_EXTERIOR
  if (cls < 0).any():
    i, j = np.argwhere(cls < 0)[0]
    raise ValueError(f"wasn't able to determine which corner the CV {(int(i), int(j))} is.")
  return cls


def _evenly_spaced_inds_from_domain(start: int, end: int, spacing: int) -> List[int]:
  """thermal_diffuser_utils.py:36-67."""
  ind_len = end - start
  if ind_len == 0:
    return [start]
  n_diffusers = np.max((1, np.round(ind_len / spacing)))
  placement = np.arange(start, end, ind_len / (n_diffusers + 1))[1:]
  return [int(math.ceil(i)) for i in placement]


def _diffusers_for_room(rows: np.ndarray, cols: np.ndarray, spacing: int,
                        buffer_from_walls: int,
                        interior_walls: np.ndarray) -> List[Tuple[int, int]]:
  """diffuser_allocation_switch thermal_diffuser_utils.py:193-262."""
  num_cvs = len(rows)
  sx, ex, sy, ey = rows.min(), rows.max(), cols.min(), cols.max()
  rect = max(ex - sx, 1) * max(ey - sy, 1)
  if num_cvs / rect > 0.1:                                 # _rectangularity_test :69-101
    if ex - sx > 2 * buffer_from_walls:                    # only the row range is buffered :163-166
      sx, ex = sx + buffer_from_walls, ex - buffer_from_walls
    px = set(_evenly_spaced_inds_from_domain(int(sx), int(ex), spacing))
    py = set(_evenly_spaced_inds_from_domain(int(sy), int(ey), spacing))
    inds = [(int(r), int(c)) for r, c in zip(rows, cols) if r in px and c in py]
  else:                                                    # :103-133
    rng = np.random.default_rng(23)
    n = int(np.max((1, np.round(num_cvs / (spacing * spacing)))))
    pick = rng.choice(np.stack([rows, cols], axis=1), n, replace=False)
    inds = [(int(r), int(c)) for r, c in pick]
  return [ind for ind in inds if interior_walls[ind[0], ind[1]] == 0]


def compile_plan(floor_plan: np.ndarray, zone_map: Optional[np.ndarray] = None, *,
                 cv_size_cm: float,
                 inside_air: MaterialProperties,
                 inside_wall: MaterialProperties,
                 building_exterior: MaterialProperties,
                 buffer_from_walls: int = 3,
                 diffuser_spacing: int = 10,
                 diffuser_mask: Optional[np.ndarray] = None,
                 material_id: Optional[np.ndarray] = None) -> CompiledPlan:
  """FloorPlanBasedBuilding.__init__ (building.py:634-766) + TFSimulator statics.

  `diffuser_mask` (bool [H, W], optional) places the diffusers explicitly instead of the
  spacing rule -- what the reference's tests do with `building.diffusers = ...`
  (simulator_flexible_floor_plan_test.py:467-471); every zone's diffusers share its power
  equally (building.py:349-351).  `material_id` (int [H, W] of MAT_AIR / MAT_INTERIOR_WALL /
  MAT_EXTERIOR_WALL, optional) overrides the wall classification (legacy_building)."""
  floor_plan = np.asarray(floor_plan)
  zone_map = floor_plan if zone_map is None else np.asarray(zone_map)
  plan = guarantee_air_padding_in_frame(floor_plan)
  zmap = guarantee_air_padding_in_frame(zone_map)
  if plan.shape != zmap.shape:
    raise ValueError("floor_plan and zone_map shapes differ after padding")
  ext = plan == 2                                          # :222-251
  shell = ndimage.binary_dilation(ext, structure=_CROSS) & ~ext   # :322-356
  interior_walls = (plan == 1) & ~shell                    # :359-373
  # enlarge_exterior_walls building.py:183-229
  near_shell = ndimage.binary_dilation(shell, structure=_ENLARGE)
  exterior_walls = shell | (near_shell & interior_walls)
  interior_walls_shrunk = interior_walls & ~exterior_walls
  if material_id is None:
    material_id = np.where(interior_walls_shrunk, MAT_INTERIOR_WALL,
                           np.where(exterior_walls, MAT_EXTERIOR_WALL, MAT_AIR)).astype(np.int8)
  else:
    material_id = np.asarray(material_id, dtype=np.int8)
    if material_id.shape != floor_plan.shape:
      raise ValueError("material_id must have the floor plan's shape")
    pad = (plan.shape[0] - floor_plan.shape[0]) // 2
    if pad:
      material_id = np.pad(material_id, pad, mode="constant", constant_values=MAT_AIR)
  # rooms: 4-connected components of the zone map's interior space :254-292, 376-414
  labels, n_rooms = ndimage.label(zmap == 0, structure=_CROSS)
  labels = np.where(zmap == 2, -1, labels)                 # exterior space set negative :294-320
  zone_id = np.full(plan.shape, -1, dtype=np.int16)
  zone_names: List[str] = []
  # room_dict insertion order = first appearance in raster order :396-413
  flat = labels.ravel()
  pos = flat > 0
  first = {}
  if pos.any():
    vals, idx = np.unique(flat[pos], return_index=True)
    order = vals[np.argsort(idx)]
    for zi, lab in enumerate(order):
      first[int(lab)] = zi
      zone_names.append(f"room_{int(lab)}")
  for lab, zi in first.items():
    zone_id[labels == lab] = zi
  n_zones = len(zone_names)
  if n_zones > 254:
    raise ValueError(f"{n_zones} zones; the packed descriptor holds at most 254")
  diffuser_weight = np.zeros(plan.shape, dtype=np.float64)
  zone_ncv = np.zeros(n_zones, dtype=np.int32)
  zone_ndiff = np.zeros(n_zones, dtype=np.int32)
  iw_for_check = interior_walls.astype(np.int8)            # pre-shrink walls, building.py:751-757
  dmask = None
  if diffuser_mask is not None:
    dmask = np.asarray(diffuser_mask, dtype=bool)
    if dmask.shape != floor_plan.shape:
      raise ValueError("diffuser_mask must have the floor plan's shape")
    pad = (plan.shape[0] - floor_plan.shape[0]) // 2       # ring added by guarantee_air_padding_in_frame
    if pad:
      dmask = np.pad(dmask, pad, mode="constant", constant_values=False)
  for zi in range(n_zones):
    rows, cols = np.nonzero(zone_id == zi)
    zone_ncv[zi] = len(rows)
    if diffuser_mask is not None:
      inds = [(int(r), int(c)) for r, c in zip(rows, cols) if dmask[r, c]]
    else:
      inds = _diffusers_for_room(rows, cols, diffuser_spacing, buffer_from_walls, iw_for_check)
    zone_ndiff[zi] = len(inds)
    for r, c in inds:
      diffuser_weight[r, c] = 1.0 / float(len(inds))       # building.py:349-351
  cv_class = classify_cvs(ext)
  desc = (cv_class.astype(np.uint16)
          | (material_id.astype(np.uint16) << _lib.DESC_MATERIAL_SHIFT)
          | np.where(diffuser_weight > 0, _lib.DESC_DIFFUSER, 0).astype(np.uint16)
          | (np.where(zone_id >= 0, zone_id, _lib.ZONE_NONE).astype(np.uint16)
             << _lib.DESC_ZONE_SHIFT))
  material = np.array(
      [[m.conductivity, m.heat_capacity, m.density]
       for m in (inside_air, inside_wall, building_exterior)], dtype=np.float64)
  obs_order = np.array(sorted(range(n_zones), key=lambda i: "vav_" + zone_names[i]),
                       dtype=np.int32)
  return CompiledPlan(
      height=plan.shape[0], width=plan.shape[1], desc=desc.astype(np.uint16),
      material=material, cv_size_m=cv_size_cm / 100.0, zone_names=zone_names,
      zone_ncv=zone_ncv, zone_ndiff=zone_ndiff, obs_zone_order=obs_order,
      cv_class=cv_class, material_id=material_id, zone_id=zone_id,
      diffuser_weight=diffuser_weight, exterior_space=ext)


# ----------------------------------------------------------------------------
# legacy rectangular Building
# ----------------------------------------------------------------------------


def legacy_building_plan(room_shape: Tuple[int, int], building_shape: Tuple[int, int]):
  """The reference's deprecated rectangular `Building` (building.py:394-505: rooms of
  room_shape air CVs in a building_shape grid, one-CV interior walls, a two-CV exterior
  shell) as (floor_plan, diffuser_mask) for compile_plan.

  The equivalence is the reference's own: its tests rebuild the old building as a floor plan
  -- the old grid with one more ring of exterior space around it -- and copy the old
  diffusers over (simulator_flexible_floor_plan_test.py:103-160, 467-471).  The four
  diffusers per room follow generate_thermal_diffusers (building.py:102-157)."""
  rs0, rs1 = int(room_shape[0]), int(room_shape[1])
  nrows = (rs0 + 1) * int(building_shape[0]) + 3           # building.py:463-464
  ncols = (rs1 + 1) * int(building_shape[1]) + 3
  old = np.zeros((nrows, ncols), dtype=np.int8)            # 0 air, 1 wall
  for x in range(rs0 + 2, nrows - 2, rs0 + 1):             # assign_interior_wall_values :94-99
    old[x, 2:ncols - 2] = 1
  for y in range(rs1 + 2, ncols - 2, rs1 + 1):
    old[2:nrows - 2, y] = 1
  old[:, [0, 1, -2, -1]] = 1                               # assign_building_exterior_values :73-74
  old[[0, 1, -2, -1], :] = 1
  diff = np.zeros((nrows, ncols), dtype=bool)              # generate_thermal_diffusers :117-156
  d1x = (rs0 - 2) // 3
  d2x = rs0 - d1x - 1
  d1y = (rs1 - 2) // 3
  d2y = rs1 - d1y - 1
  for sx in range(2, nrows - 3, rs0 + 1):
    for sy in range(2, ncols - 3, rs1 + 1):
      for dx in (d1x, d2x):
        for dy in (d1y, d2y):
          diff[sx + dx, sy + dy] = True
  plan = np.pad(old, 1, mode="constant", constant_values=2)
  return plan, np.pad(diff, 1, mode="constant", constant_values=False)


def legacy_building(cv_size_cm: float, room_shape: Tuple[int, int], building_shape: Tuple[int, int],
                    inside_air: MaterialProperties, inside_wall: MaterialProperties,
                    building_exterior: MaterialProperties, exact_materials: bool = True) -> CompiledPlan:
  """`Building(cv_size_cm, floor_height_cm, room_shape, building_shape, initial_temp, ...)`
  of the reference (building.py:419-505) compiled for libsbx; floor height and initial
  temperature are arguments of SimulatorBuilding here.

  exact_materials=True reproduces the old Building's material arrays exactly (two exterior
  layers, interior wall lines up to them); False gives the reference's own floor-plan twin
  (simulator_flexible_floor_plan_test.py:431-472), whose wall classification turns the 2 x
  (rooms - 1) x 2 interior-wall CVs that touch the shell into exterior wall."""
  plan, dm = legacy_building_plan(room_shape, building_shape)
  mat = None
  if exact_materials:
    old = plan[1:-1, 1:-1]
    m = np.where(old == 1, MAT_INTERIOR_WALL, MAT_AIR).astype(np.int8)
    m[:, [0, 1, -2, -1]] = MAT_EXTERIOR_WALL                # assign_building_exterior_values :73-74
    m[[0, 1, -2, -1], :] = MAT_EXTERIOR_WALL
    mat = np.pad(m, 1, mode="constant", constant_values=MAT_AIR)
  return compile_plan(plan.astype(np.int64), None, cv_size_cm=cv_size_cm, inside_air=inside_air,
                      inside_wall=inside_wall, building_exterior=building_exterior,
                      buffer_from_walls=0, diffuser_mask=dm, material_id=mat)


# ----------------------------------------------------------------------------
# randomised rectangular-room plans (BASELINE.json configs 3-5)
# ----------------------------------------------------------------------------


@dataclasses.dataclass
class RandomPlanSpec:
  height: int = 64
  width: int = 96
  rooms_y: Tuple[int, ...] = (1, 2, 3)      # room rows
  rooms_x: Tuple[int, ...] = (2, 3, 4)      # room columns
  min_room: int = 6


def random_floor_plan(rng: np.random.Generator, spec: RandomPlanSpec = RandomPlanSpec()) -> np.ndarray:
  """One exterior-air ring, a wall shell, Ky x Kx rooms split by jittered wall lines."""
  h, w = spec.height, spec.width
  plan = np.full((h, w), 2, dtype=np.int8)
  plan[1:h - 1, 1:w - 1] = 1
  plan[2:h - 2, 2:w - 2] = 0
  ky = int(rng.choice(spec.rooms_y))
  kx = int(rng.choice(spec.rooms_x))

  def cuts(lo: int, hi: int, k: int) -> List[int]:
    # k rooms need k-1 wall lines in [lo, hi); rooms at least min_room wide
    span = hi - lo
    out = []
    for i in range(1, k):
      centre = lo + span * i / k
      jitter = max(0, int(span / k / 2) - spec.min_room)
      c = int(round(centre + (rng.integers(-jitter, jitter + 1) if jitter > 0 else 0)))
      out.append(c)
    return out

  for r in cuts(2, h - 2, ky):
    plan[r, 2:w - 2] = 1
  for c in cuts(2, w - 2, kx):
    plan[2:h - 2, c] = 1
  return plan


def random_materials(rng: np.random.Generator):
  """air as calibrated, random wall / exterior."""
  air = MaterialProperties(50.0, 700.0, 1.0)
  wall = MaterialProperties(float(rng.uniform(2, 50)), float(rng.uniform(500, 1000)),
                            float(rng.uniform(1, 1800)))
  ext = MaterialProperties(float(rng.uniform(0.05, 5)), float(rng.uniform(500, 1000)),
                           float(rng.uniform(1, 1800)))
  return air, wall, ext


def pack_plans(plans: Sequence[CompiledPlan], n_zones: Optional[int] = None) -> Dict[str, np.ndarray]:
  """Stacks compiled plans into the arrays `sbx_upload` expects ([P, ...])."""
  h, w = plans[0].height, plans[0].width
  for p in plans:
    if (p.height, p.width) != (h, w):
      raise ValueError("all plans of a batch must share one grid size")
  z = max(p.n_zones for p in plans) if n_zones is None else n_zones
  z = max(z, 1)
  n = len(plans)
  out = {
      "plan_desc": np.stack([p.desc for p in plans]).astype(np.uint16),
      "plan_material": np.stack([p.material for p in plans]).astype(np.float64),
      "plan_cv_size": np.array([p.cv_size_m for p in plans], dtype=np.float64),
      "zone_ncv": np.zeros((n, z), dtype=np.int32),
      "zone_ndiff": np.zeros((n, z), dtype=np.int32),
      "obs_zone_order": np.full((n, z), -1, dtype=np.int32),
  }
  for i, p in enumerate(plans):
    if p.n_zones > z:
      raise ValueError(f"plan {i} has {p.n_zones} zones > {z}")
    out["zone_ncv"][i, :p.n_zones] = p.zone_ncv
    out["zone_ndiff"][i, :p.n_zones] = p.zone_ndiff
    out["obs_zone_order"][i, :p.n_zones] = p.obs_zone_order
  return out
