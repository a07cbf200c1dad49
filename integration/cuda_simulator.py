"""libsbx behind the UNMODIFIED reference's seams, one building (B = 1).

  make_cuda_simulator()           -> class CudaSimulator(SimulatorFlexibleGeometries)      seam B3
      overrides finite_differences_timestep (simulator.py:318-371): the reference's step loop,
      devices, thermostats, reward and Environment run untouched around the CUDA solve;
      `SimulatorBuilding(simulator=CudaSimulator(...), occupancy=...)` accepts it
      (simulator_building.py:49-53).
  make_cuda_simulator_building()  -> class CudaSimulatorBuilding(BaseBuilding)              seam B2
      the whole building on the GPU (models/base_building.py:27-95): proto requests in, proto
      responses out, so that the unmodified `Environment` (environment.py:355) drives it.

Both classes subclass reference classes, so they are built by factory functions that import
the reference (package `smart_buildings.smart_control`) when called; everything that does not
need it -- turning a reference `FloorPlanBasedBuilding` into libsbx's static plan, the
solver configuration, the per-zone diffuser heat -- is plain functions over NumPy arrays and
is tested on the CPU against the reference building (tests/test_integration_adapters.py).
"""

from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import pandas as pd

import sbsim_b200 as sbx
from sbsim_b200 import _lib, config as config_lib, floorplan

ROOM_PREFIX = "room"     # constants.ROOM_STRING_DESIGNATOR


# ---------------------------------------------------------------------------
# reference building -> static plan (no reference import needed: attribute access only)
# ---------------------------------------------------------------------------


def plan_from_reference_building(building) -> floorplan.CompiledPlan:
  """Static plan of a reference `FloorPlanBasedBuilding` (building.py:609-893) from its own
  arrays: exterior space, the three materials, diffuser weights and the room dict.

  Raises ValueError when the building's (conductivity, heat_capacity, density) arrays do not
  decompose into air / interior wall / exterior wall (libsbx keeps three materials per plan)."""
  ext = np.asarray(building._exterior_space) == -1                     # building.py:294-320
  shape = ext.shape
  k = np.asarray(building.conductivity, dtype=np.float64)
  c = np.asarray(building.heat_capacity, dtype=np.float64)
  rho = np.asarray(building.density, dtype=np.float64)
  material_id = np.full(shape, floorplan.MAT_AIR, dtype=np.int8)
  material_id[np.asarray(building._interior_walls) != 0] = floorplan.MAT_INTERIOR_WALL
  material_id[np.asarray(building._exterior_walls) != 0] = floorplan.MAT_EXTERIOR_WALL
  material = np.zeros((3, 3), dtype=np.float64)
  for m in range(3):
    sel = material_id == m
    if not sel.any():
      continue
    vals = np.stack([k[sel], c[sel], rho[sel]], axis=1)
    if not np.all(vals == vals[0]):
      raise ValueError("the building's material arrays are not uniform per material class")
    material[m] = vals[0]
  rooms = [(name, cvs) for name, cvs in building._room_dict.items() if name.startswith(ROOM_PREFIX)]
  if len(rooms) > 254:
    raise ValueError(f"{len(rooms)} zones; the packed descriptor holds at most 254")
  zone_id = np.full(shape, -1, dtype=np.int16)
  zone_names: List[str] = []
  for zi, (name, cvs) in enumerate(rooms):
    zone_names.append(name)
    rows, cols = zip(*cvs) if len(cvs) else ((), ())
    zone_id[np.asarray(rows, dtype=np.int64), np.asarray(cols, dtype=np.int64)] = zi
  diffusers = np.asarray(building.diffusers, dtype=np.float64)
  zone_ncv = np.array([len(cvs) for _, cvs in rooms], dtype=np.int32)
  zone_ndiff = np.array([int(((zone_id == zi) & (diffusers > 0)).sum()) for zi in range(len(rooms))],
                        dtype=np.int32)
  cv_class = floorplan.classify_cvs(ext)
  desc = (cv_class.astype(np.uint16)
          | (material_id.astype(np.uint16) << _lib.DESC_MATERIAL_SHIFT)
          | np.where(diffusers > 0, _lib.DESC_DIFFUSER, 0).astype(np.uint16)
          | (np.where(zone_id >= 0, zone_id, _lib.ZONE_NONE).astype(np.uint16) << _lib.DESC_ZONE_SHIFT))
  order = np.array(sorted(range(len(rooms)), key=lambda i: "vav_" + zone_names[i]), dtype=np.int32)
  return floorplan.CompiledPlan(
      height=shape[0], width=shape[1], desc=desc.astype(np.uint16), material=material,
      cv_size_m=float(building.cv_size_cm) / 100.0, zone_names=zone_names, zone_ncv=zone_ncv,
      zone_ndiff=zone_ndiff, obs_zone_order=order, cv_class=cv_class, material_id=material_id,
      zone_id=zone_id, diffuser_weight=diffusers, exterior_space=ext)


def diffuser_heat_per_zone(building, plan: floorplan.CompiledPlan) -> np.ndarray:
  """`input_q` at a diffuser CV of each zone, float32 [1, Z]: what apply_thermal_power_zone
  (building.py:873-889) scattered -- every diffuser CV of a zone holds the same value."""
  q = np.zeros((1, max(1, plan.n_zones)), dtype=np.float32)
  input_q = np.asarray(building.input_q)
  for zi in range(plan.n_zones):
    m = (plan.zone_id == zi) & (plan.diffuser_weight > 0)
    if m.any():
      vals = input_q[m]
      q[0, zi] = np.float32(vals.flat[0])
  return q


def make_sbx_config(plan: floorplan.CompiledPlan, *, time_step_sec: float, convergence_threshold: float,
                    iteration_limit: int, floor_height_cm: float, solver: str = "tf_jacobi"
                    ) -> "_lib.SbxConfig":
  """sbx_config of a solve-only handle (sbx_fd_step, B = 1): the FD parameters of
  Simulator.__init__ (simulator.py:45-78); device / reward fields take the calibrated defaults
  and are not used by sbx_fd_step."""
  cfg = _lib.SbxConfig()
  cfg.abi_version = _lib.ABI_VERSION
  cfg.n_envs, cfg.height, cfg.width = 1, plan.height, plan.width
  cfg.n_zones = max(1, plan.n_zones)
  cfg.n_plans = cfg.n_weather = cfg.n_occ_zones = 1
  cfg.n_reset = 0
  cfg.episode_steps, cfg.n_table_steps = 1, 3
  cfg.kernel_path = _lib.PATH_AUTO
  cfg.solver = _lib.SOLVER_GAUSS_SEIDEL if solver == "gauss_seidel" else _lib.SOLVER_TF_JACOBI
  cfg.iteration_limit = int(iteration_limit)
  cfg.time_step_sec = float(time_step_sec)
  cfg.floor_height_m = float(floor_height_cm) / 100.0
  cfg.convergence_threshold = float(convergence_threshold)
  cfg.comfort_heat, cfg.comfort_cool, cfg.eco_heat, cfg.eco_cool = 294.0, 297.0, 289.0, 298.0
  cfg.ahu_recirculation, cfg.ahu_init_heating_setpoint, cfg.ahu_init_cooling_setpoint = 0.3, 285.0, 298.0
  cfg.ahu_fan_differential_pressure, cfg.ahu_fan_efficiency, cfg.ahu_max_air_flow_rate = 1e4, 0.9, 8.67
  cfg.boiler_init_setpoint, cfg.boiler_pump_head, cfg.boiler_pump_efficiency = 360.0, 6.0, 0.98
  cfg.boiler_heating_rate, cfg.boiler_cooling_rate = 0.5, 0.1
  b = config_lib.Boiler(360.0, 6.0, 0.98)
  cfg.boiler_convection_coefficient, cfg.boiler_tank_length = b.convection_coefficient, b.tank_length
  cfg.boiler_tank_radius, cfg.boiler_water_capacity = b.tank_radius, b.water_capacity
  cfg.boiler_insulation_conductivity, cfg.boiler_insulation_thickness = b.insulation_conductivity, b.insulation_thickness
  cfg.vav_max_air_flow_rate, cfg.vav_reheat_max_water_flow_rate = 0.035, 0.03
  cfg.max_productivity_personhour_usd, cfg.min_productivity_personhour_usd = 300.0, 100.0
  cfg.max_electricity_rate, cfg.max_natural_gas_rate = 160000.0, 400000.0
  cfg.productivity_midpoint_delta, cfg.productivity_decay_stiffness = 0.5, 4.3
  cfg.productivity_weight, cfg.energy_cost_weight, cfg.carbon_emission_weight = 0.2, 0.4, 0.4
  cfg.reward_kind = _lib.REWARD_REGRET
  cfg.discount_factor = 1.0
  cfg.n_actions = 0
  cfg.obs_mode = _lib.OBS_RAW
  return cfg


def upload_plan(handle: "_lib.Handle", plan: floorplan.CompiledPlan) -> None:
  """SBX_F_PLAN_DESC / _MATERIAL / _CV_SIZE / ZONE_NCV / ZONE_NDIFF / OBS_ZONE_ORDER."""
  for name, arr in floorplan.pack_plans([plan], n_zones=max(1, plan.n_zones)).items():
    handle.upload(name, arr)


# ---------------------------------------------------------------------------
# B3: finite_differences_timestep on libsbx
# ---------------------------------------------------------------------------


def cuda_fd_timestep(handle: "_lib.Handle", plan: floorplan.CompiledPlan, building, solver: str,
                     ambient_temperature: float, convection_coefficient: float,
                     iteration_limit: int, convergence_threshold: float) -> bool:
  """finite_differences_timestep (simulator.py:318-371) of ONE building on libsbx: uploads
  `building.temp` and the per-zone diffuser heat of `building.input_q`, solves to convergence,
  stores the result in `building.temp`; returns the reference's `converged` flag.  `building`
  only needs the attributes temp / input_q, so this is testable without the reference."""
  shape = (1, plan.height, plan.width)
  if solver == "gauss_seidel":
    handle.upload("temp64", np.ascontiguousarray(building.temp, dtype=np.float64)[None])
    handle.upload("q_cv64", diffuser_heat_per_zone(building, plan).astype(np.float64))
  else:
    handle.upload("temp", np.ascontiguousarray(building.temp, dtype=np.float32)[None])   # tf.convert_to_tensor(.., float32)
    handle.upload("q_cv", diffuser_heat_per_zone(building, plan))
  handle.fd_step(np.array([float(ambient_temperature)]), np.array([float(convection_coefficient)]))
  building.temp = (handle.download("temp64", shape)[0] if solver == "gauss_seidel"
                   else handle.download("temp", shape)[0])                                # simulator.py:369
  n = int(handle.download("n_sweeps", (1,))[0])
  return n < iteration_limit or float(handle.download("max_delta", (1,))[0]) <= convergence_threshold


def make_cuda_simulator(solver: str = "tf_jacobi", device: int = 0):
  """Returns class CudaSimulator(SimulatorFlexibleGeometries).  solver='tf_jacobi' reproduces
  TFSimulator.update_temperature_estimates (tf_simulator.py:573-853, fp32 Jacobi),
  'gauss_seidel' the base class's own sweep (simulator.py:98-316, fp64)."""
  from smart_buildings.smart_control.simulator import simulator_flexible_floor_plan as sffp

  class CudaSimulator(sffp.SimulatorFlexibleGeometries):
    """The reference simulator with its diffusion solve on the GPU (B = 1)."""

    def __init__(self, *args, **kwargs):
      super().__init__(*args, **kwargs)
      b = self._building
      self._sbx_plan = plan_from_reference_building(b)
      cfg = make_sbx_config(self._sbx_plan, time_step_sec=self._time_step_sec,
                            convergence_threshold=self._convergence_threshold,
                            iteration_limit=self._iteration_limit,
                            floor_height_cm=b.floor_height_cm, solver=solver)
      self._sbx = _lib.Handle(cfg, device)
      upload_plan(self._sbx, self._sbx_plan)
      self._sbx_solver = solver

    def finite_differences_timestep(self, *, ambient_temperature: float,
                                    convection_coefficient: float) -> bool:
      return cuda_fd_timestep(self._sbx, self._sbx_plan, self._building, self._sbx_solver,
                              ambient_temperature, convection_coefficient, self._iteration_limit,
                              self._convergence_threshold)

  return CudaSimulator


# ---------------------------------------------------------------------------
# B2: BaseBuilding over a B = 1 sbsim_b200.Environment-like handle
# ---------------------------------------------------------------------------

AHU_FIELDS = sbx.environment.AHU_FIELDS
BOILER_FIELDS = sbx.environment.BOILER_FIELDS
VAV_FIELDS = sbx.environment.VAV_FIELDS


def native_observations(env: "sbx.Environment", b: int = 0) -> Dict[tuple, float]:
  """{(device_id, measurement_name): native value} of building b after the last step, from
  libsbx's diagnostics -- what SimulatorBuilding.request_observations reads off the device
  objects (simulator_building.py:151-202, smart_device.py:149-170)."""
  h, cfg, bld = env.handle, env._cfg, env.building
  B, Z = bld.n_envs, bld.n_zones
  diag = h.download("step_diag", (B, _lib.DIAG_N))[b]
  D = _lib.DIAG
  pre = h.download("pre_zone_mean", (B, Z))[b]
  mode = h.download("thermostat_mode", (B, Z))[b]
  heat_sp = h.download("ahu_heating_sp", (B,))[b]
  cool_sp = h.download("ahu_cooling_sp", (B,))[b]
  boiler_sp = h.download("boiler_sp", (B,))[b]
  flow = diag[D["ahu_flow"]]
  t_index = env._time_index
  amb = env._ambient_table[0 if env._ambient_table.shape[0] == 1 else b, t_index]
  fan_pct = flow / cfg.ahu_max_air_flow_rate
  out = {}
  ahu = dict(zip(AHU_FIELDS, (diag[D["cooling_requests"]], cfg.ahu_fan_differential_pressure, fan_pct,
                              (1.0 - cfg.ahu_recirculation) * flow, amb, cool_sp, flow, heat_sp, fan_pct)))
  for f, v in ahu.items():
    out[("air_handler", f)] = float(v)
  for f, v in zip(BOILER_FIELDS, (diag[D["heating_requests"]], boiler_sp, diag[D["tank_temp"]])):
    out[("boiler", f)] = float(v)
  plan = bld.plans[0 if len(bld.plans) == 1 else b]
  for zi, name in enumerate(plan.zone_names):
    damper = 1.0 if mode[zi] in (1, 2) else 0.1
    for f, v in zip(VAV_FIELDS, (damper, cfg.vav_max_air_flow_rate, pre[zi])):
      out[("vav_" + name, f)] = float(v)
  return out


def make_cuda_simulator_building():
  """Returns class CudaSimulatorBuilding(BaseBuilding): one building of a sbsim_b200.Environment
  behind the reference's BaseBuilding interface, protos in and out."""
  from smart_buildings.smart_control.models import base_building
  from smart_buildings.smart_control.proto import smart_control_building_pb2 as bpb
  from smart_buildings.smart_control.proto import smart_control_reward_pb2 as rpb
  from smart_buildings.smart_control.utils import conversion_utils

  class CudaSimulatorBuilding(base_building.BaseBuilding):
    """BaseBuilding (models/base_building.py:27-95) over libsbx, B = 1.

    request_action stores the native setpoints; wait_time runs one sbx step with them mapped
    back to the agent's [-1, 1] range (the library applies BoundedActionNormalizer itself);
    request_observations / reward_info read the step's diagnostics."""

    def __init__(self, env: "sbx.Environment", building_id: str = "US-SIM-001"):
      if env.batch_size != 1:
        raise ValueError("CudaSimulatorBuilding wraps a one-building environment")
      self._env = env
      self._building_id = building_id
      self._pending = {}
      self._plan = env.building.plans[0]
      env.reset()

    # -- identity ---------------------------------------------------------
    @property
    def devices(self):
      devs = [bpb.DeviceInfo(device_id="air_handler", namespace="SIM", code="AHU",
                             device_type=bpb.DeviceInfo.AHU,
                             observable_fields={f: bpb.DeviceInfo.VALUE_CONTINUOUS for f in AHU_FIELDS},
                             action_fields={"supply_air_heating_temperature_setpoint": bpb.DeviceInfo.VALUE_CONTINUOUS,
                                            "supply_air_cooling_temperature_setpoint": bpb.DeviceInfo.VALUE_CONTINUOUS}),
              bpb.DeviceInfo(device_id="boiler", namespace="SIM", code="BLR", device_type=bpb.DeviceInfo.BLR,
                             observable_fields={f: bpb.DeviceInfo.VALUE_CONTINUOUS for f in BOILER_FIELDS},
                             action_fields={"supply_water_setpoint": bpb.DeviceInfo.VALUE_CONTINUOUS})]
      for name in self._plan.zone_names:
        devs.append(bpb.DeviceInfo(
            device_id="vav_" + name, namespace="SIM", code="VAV", device_type=bpb.DeviceInfo.VAV,
            zone_id=conversion_utils.floor_plan_based_zone_identifier_to_id(name),
            observable_fields={f: bpb.DeviceInfo.VALUE_CONTINUOUS for f in VAV_FIELDS}))
      return devs

    @property
    def zones(self):
      return [bpb.ZoneInfo(zone_id=conversion_utils.floor_plan_based_zone_identifier_to_id(n),
                           building_id=self._building_id, zone_description="Simulated zone",
                           devices=["vav_" + n], zone_type=bpb.ZoneInfo.ROOM, floor=0)
              for n in self._plan.zone_names]

    @property
    def current_timestamp(self) -> pd.Timestamp:
      return self._env.current_simulation_timestamp

    @property
    def time_step_sec(self) -> float:
      return self._env.building.time_step_sec

    def is_comfort_mode(self, current_time: pd.Timestamp) -> bool:
      return bool(self._env.building.hvac.schedule.is_comfort_mode(current_time))

    @property
    def num_occupants(self) -> int:
      return int(self._env._occ_obs_table[self._env._time_index])

    def render(self, path: str) -> None:
      raise NotImplementedError("Rendering not currently supported on simulated building")

    # -- step ---------------------------------------------------------------
    def reset(self) -> None:
      self._env.reset()
      self._pending = {}

    def request_action(self, action_request):
      ts = conversion_utils.pandas_to_proto_timestamp(self.current_timestamp)
      response = bpb.ActionResponse(timestamp=ts, request=action_request)
      for r in action_request.single_action_requests:
        self._pending[r.setpoint_name] = float(r.continuous_value)
        response.single_action_responses.append(bpb.SingleActionResponse(
            request=r, response_type=bpb.SingleActionResponse.ACCEPTED))
      return response

    def wait_time(self) -> None:
      env = self._env
      a = np.zeros((1, len(env._action_fields)), dtype=np.float32)
      for i, ((_, setpoint), fid) in enumerate(zip(env._action_fields, env._action_names)):
        norm = env._action_normalizers[fid]
        native = self._pending.get(setpoint)
        if native is None:
          raise RuntimeError(f"no action requested for {setpoint}")
        a[0, i] = 2.0 * (native - norm.setpoint_min) / (norm.setpoint_max - norm.setpoint_min) - 1.0
      self._last = env.step(a)

    def request_observations(self, observation_request):
      ts = conversion_utils.pandas_to_proto_timestamp(self.current_timestamp)
      native = native_observations(self._env)
      response = bpb.ObservationResponse(timestamp=ts, request=observation_request)
      for r in observation_request.single_observation_requests:
        key = (r.device_id, r.measurement_name)
        valid = key in native
        response.single_observation_responses.append(bpb.SingleObservationResponse(
            timestamp=ts, single_observation_request=r, observation_valid=valid,
            continuous_value=native.get(key, 0.0)))
      return response

    def request_observations_within_time_interval(self, observation_request, start_timestamp, end_timestamp):
      raise NotImplementedError("historical observations are not kept by the simulated building")

    @property
    def reward_info(self):
      env, h = self._env, self._env.handle
      Z = env.building.n_zones
      diag = h.download("step_diag", (1, _lib.DIAG_N))[0]
      D = _lib.DIAG
      zmean = h.download("zone_mean", (1, Z))[0]
      ts = self.current_timestamp
      window = env.building.hvac.schedule.get_temperature_window(ts)
      dt = pd.Timedelta(self.time_step_sec, unit="s")
      info = rpb.RewardInfo(
          start_timestamp=conversion_utils.pandas_to_proto_timestamp(ts),
          end_timestamp=conversion_utils.pandas_to_proto_timestamp(ts + dt))
      occ = env._occ_reward_table[env._time_index]
      for zi, name in enumerate(self._plan.zone_names):
        zid = conversion_utils.floor_plan_based_zone_identifier_to_id(name)
        info.zone_reward_infos[zid].CopyFrom(rpb.RewardInfo.ZoneRewardInfo(
            heating_setpoint_temperature=window[0], cooling_setpoint_temperature=window[1],
            zone_air_temperature=float(zmean[zi]),
            air_flow_rate_setpoint=env._cfg.vav_max_air_flow_rate, air_flow_rate=float(diag[D["ahu_flow"]]),
            average_occupancy=float(occ[0 if occ.shape[0] == 1 else zi])))
      info.air_handler_reward_infos["air_handler"].CopyFrom(rpb.RewardInfo.AirHandlerRewardInfo(
          blower_electrical_energy_rate=float(diag[D["blower_w"]]),
          air_conditioning_electrical_energy_rate=float(diag[D["ac_w"]])))
      info.boiler_reward_infos["boiler"].CopyFrom(rpb.RewardInfo.BoilerRewardInfo(
          natural_gas_heating_energy_rate=float(diag[D["gas_w"]]),
          pump_electrical_energy_rate=float(diag[D["pump_w"]])))
      return info

  return CudaSimulatorBuilding
