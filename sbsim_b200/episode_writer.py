"""Episode logs in the reference's on-disk format.

The reference logs every step's RewardInfo / RewardResponse / ObservationResponse /
ActionResponse through `ProtoWriter` (`smart_control/utils/controller_writer.py:35-170`):
hourly shard files `<prefix>_<YYYY.MM.DD.HH>`, each record a 4-byte little-endian length
followed by the serialized proto; its readers and plotting notebooks (`SAC_Demo.ipynb`
cells 5, 22) consume those files.  `EpisodeWriter` produces the same files for a sampled
subset of the batched env's buildings.

protoc and the generated `*_pb2` modules are not part of this repository, so the messages
are serialized directly in the protobuf wire format, field numbers and types as published in
`smart_control/proto/smart_control_reward.proto:12-89` and
`smart_control_building.proto:83-165` (proto3: implicit-presence scalars are omitted when
zero, `oneof` members are always written).  `tests/test_episode_writer.py` parses the
output back with the reference's own runtime message classes where the reference tree is
mounted.  Host-side only: values come from the handle's diagnostic downloads.
"""

from __future__ import annotations

import os
import struct
from typing import Dict, Optional, Sequence

import numpy as np
import pandas as pd

from sbsim_b200 import _lib

OBSERVATION_RESPONSE_FILE_PREFIX = "observation_response"   # utils/constants.py:51-56
ACTION_RESPONSE_FILE_PREFIX = "action_response"
REWARD_INFO_PREFIX = "reward_info"
REWARD_RESPONSE_PREFIX = "reward_response"

AHU_FIELDS = (   # air_handler.py:66-95, sorted as request_observations does
    "cooling_request_count", "differential_pressure_setpoint",
    "discharge_fan_speed_percentage_command", "outside_air_flowrate_sensor",
    "outside_air_temperature_sensor", "supply_air_cooling_temperature_setpoint",
    "supply_air_flowrate_sensor", "supply_air_heating_temperature_setpoint",
    "supply_fan_speed_percentage_command")
BOILER_FIELDS = ("heating_request_count", "supply_water_setpoint",
                 "supply_water_temperature_sensor")          # boiler.py:69-79
VAV_FIELDS = ("supply_air_damper_percentage_command", "supply_air_flowrate_setpoint",
              "zone_air_temperature_sensor")                 # vav.py:54-64


# ---- protobuf wire format ------------------------------------------------------


def _varint(v: int) -> bytes:
  v &= (1 << 64) - 1                     # negative int32 / int64 are 10-byte varints
  out = bytearray()
  while True:
    b = v & 0x7F
    v >>= 7
    if v:
      out.append(b | 0x80)
    else:
      out.append(b)
      return bytes(out)


def _tag(field: int, wire_type: int) -> bytes:
  return _varint((field << 3) | wire_type)


def _f32(field: int, value, always: bool = False) -> bytes:
  raw = struct.pack("<f", float(np.float32(value)))
  if not always and raw == b"\x00\x00\x00\x00":       # proto3 default: not on the wire
    return b""
  return _tag(field, 5) + raw


def _int(field: int, value: int, always: bool = False) -> bytes:
  if not always and int(value) == 0:
    return b""
  return _tag(field, 0) + _varint(int(value))


def _bytes(field: int, payload: bytes, always: bool = True) -> bytes:
  if not always and not payload:
    return b""
  return _tag(field, 2) + _varint(len(payload)) + payload


def _str(field: int, s: str, always: bool = False) -> bytes:
  return _bytes(field, s.encode("utf-8"), always=always)


def _timestamp(ts: pd.Timestamp) -> bytes:
  """google.protobuf.Timestamp from a pandas timestamp: whole seconds, as
  conversion_utils.pandas_to_proto_timestamp (conversion_utils.py:39-49) keeps them."""
  return _int(1, int(ts.timestamp()))


# ---- messages -----------------------------------------------------------------


def encode_reward_info(start: pd.Timestamp, end: pd.Timestamp, zones: Dict[str, Sequence[float]],
                       air_handlers: Dict[str, Sequence[float]], boilers: Dict[str, Sequence[float]],
                       agent_id: str = "", scenario_id: str = "") -> bytes:
  """RewardInfo (smart_control_reward.proto:12-51).  zones[zone_id] = (heating_setpoint,
  cooling_setpoint, zone_air_temperature, air_flow_rate_setpoint, air_flow_rate,
  average_occupancy); air_handlers[id] = (blower W, air conditioning W); boilers[id] =
  (natural gas W, pump W)."""
  out = _bytes(1, _timestamp(start)) + _bytes(2, _timestamp(end))
  out += _str(3, agent_id) + _str(4, scenario_id)
  for fld, table in ((5, zones), (6, air_handlers), (7, boilers)):
    for key, vals in table.items():
      msg = b"".join(_f32(i + 1, v) for i, v in enumerate(vals))
      out += _bytes(fld, _str(1, key, always=False) + _bytes(2, msg))    # map entry {key, value}
  return out


REWARD_RESPONSE_FIELDS = (   # smart_control_reward.proto:53-85
    "agent_reward_value", "productivity_reward", "electricity_energy_cost",
    "natural_gas_energy_cost", "carbon_emitted", "carbon_cost", "productivity_weight",
    "energy_cost_weight", "carbon_emission_weight", "person_productivity", "total_occupancy",
    "reward_scale", "reward_shift", "productivity_regret", "normalized_productivity_regret",
    "normalized_energy_cost", "normalized_carbon_emission")


def encode_reward_response(start: pd.Timestamp, end: pd.Timestamp, **values) -> bytes:
  out = b""
  for i, name in enumerate(REWARD_RESPONSE_FIELDS):
    if name in values:
      out += _f32(i + 1, values.pop(name))
  if values:
    raise ValueError(f"unknown RewardResponse fields: {sorted(values)}")
  return out + _bytes(18, _timestamp(start)) + _bytes(19, _timestamp(end))


def _single_observation_request(device_id: str, measurement_name: str) -> bytes:
  return _str(1, device_id) + _str(2, measurement_name)


def encode_observation_response(ts: pd.Timestamp, observations: Sequence[tuple]) -> bytes:
  """ObservationResponse (smart_control_building.proto:98-120); observations = sequence
  of (device_id, measurement_name, continuous_value), all valid."""
  stamp = _bytes(1, _timestamp(ts))
  request = stamp + b"".join(_bytes(2, _single_observation_request(d, m)) for d, m, _ in observations)
  out = stamp + _bytes(2, request)
  for d, m, v in observations:
    single = (stamp + _bytes(2, _single_observation_request(d, m)) + _int(3, 1)
              + _f32(4, v, always=True))                      # oneof member: always written
    out += _bytes(3, single)
  return out


def encode_action_response(ts: pd.Timestamp, actions: Sequence[tuple]) -> bytes:
  """ActionResponse (smart_control_building.proto:122-165); actions = sequence of
  (device_id, setpoint_name, continuous_value), every one ACCEPTED (the simulated devices
  accept any in-range action, simulator_building.py:204-263)."""
  stamp = _bytes(1, _timestamp(ts))

  def single_request(d, s, v):
    return _str(1, d) + _str(2, s) + _f32(3, v, always=True)

  request = stamp + b"".join(_bytes(2, single_request(*a)) for a in actions)
  out = stamp + _bytes(2, request)
  for a in actions:
    out += _bytes(3, _bytes(1, single_request(*a)) + _int(2, 1))   # response_type ACCEPTED = 1
  return out


# ---- writer -------------------------------------------------------------------


class ProtoShardWriter:
  """File layout of ProtoWriter (controller_writer.py:118-147): hourly shards, records
  prefixed with their 4-byte little-endian size."""

  def __init__(self, output_dir: str):
    self.output_dir = output_dir
    os.makedirs(output_dir, exist_ok=True)

  def _path(self, prefix: str, ts: pd.Timestamp) -> str:
    return os.path.join(self.output_dir, "%s_%s" % (prefix, ts.strftime("%Y.%m.%d.%H")))

  def write(self, prefix: str, ts: pd.Timestamp, payload: bytes) -> None:
    with open(self._path(prefix, ts), "ab") as f:
      f.write(len(payload).to_bytes(4, "little"))
      f.write(payload)


def read_shard(path: str):
  """Yields the serialized records of one shard file (the reader side of the format)."""
  with open(path, "rb") as f:
    data = f.read()
  pos = 0
  while pos < len(data):
    size = int.from_bytes(data[pos:pos + 4], "little")
    yield data[pos + 4:pos + 4 + size]
    pos += 4 + size


class EpisodeWriter:
  """Logs the steps of selected buildings of a batched `sbsim_b200.Environment`.

  Call `log_step(action)` after every `env.step(action)`; one sub-directory per building
  (`env_<index>/`) receives the four shard families of the reference's ProtoWriter."""

  def __init__(self, env, output_dir: str, env_indices: Sequence[int] = (0,),
               air_handler_id: str = "air_handler_id_0", boiler_id: str = "boiler_id_0"):
    self.env = env
    self.indices = [int(i) for i in env_indices]
    self.writers = {b: ProtoShardWriter(os.path.join(output_dir, f"env_{b}")) for b in self.indices}
    self.air_handler_id, self.boiler_id = air_handler_id, boiler_id
    b = env.building
    self._Z = b.n_zones
    h = env.handle
    T = h.cfg.n_table_steps
    self._comfort = h.download("comfort", (T,))
    self._occ = h.download("occ_reward", (T, h.cfg.n_occ_zones))
    self._ambient = h.download("ambient", (h.cfg.n_weather, T))
    s = b.hvac.schedule
    self._windows = {1: tuple(float(x) for x in s.comfort_temp_window),
                     0: tuple(float(x) for x in s.eco_temp_window)}

  def _plan(self, b):
    plans = self.env.building.plans
    return plans[b] if len(plans) > 1 else plans[0]

  def log_step(self, action: np.ndarray, time_step=None) -> None:
    env, h, Z = self.env, self.env.handle, self._Z
    B = env.batch_size
    s1 = env._time_index                                    # post-step table row
    dt = pd.Timedelta(env.building.time_step_sec, unit="s")
    ts = env.current_simulation_timestamp                    # post-step timestamp
    diag = h.download("step_diag", (B, _lib.DIAG_N))
    zone_mean = h.download("zone_mean", (B, Z))
    pre_mean = h.download("pre_zone_mean", (B, Z))
    mode = h.download("thermostat_mode", (B, Z))
    heat_sp = h.download("ahu_heating_sp", (B,))
    cool_sp = h.download("ahu_cooling_sp", (B,))
    boiler_sp = h.download("boiler_sp", (B,))
    window = self._windows[int(self._comfort[s1])]
    cfg = h.cfg
    D = _lib.DIAG
    reward = None if time_step is None else np.asarray(time_step.reward)
    for b in self.indices:
      plan, w = self._plan(b), self.writers[b]
      flow = diag[b, D["ahu_flow"]]
      zones, obs = {}, []
      amb = self._ambient[b if self._ambient.shape[0] > 1 else 0, s1]
      ahu_vals = (diag[b, D["cooling_requests"]], cfg.ahu_fan_differential_pressure,
                  flow / cfg.ahu_max_air_flow_rate, (1.0 - cfg.ahu_recirculation) * flow, amb,
                  cool_sp[b], flow, heat_sp[b], flow / cfg.ahu_max_air_flow_rate)
      obs += [(self.air_handler_id, n, v) for n, v in zip(AHU_FIELDS, ahu_vals)]
      obs += [(self.boiler_id, n, v) for n, v in zip(
          BOILER_FIELDS, (diag[b, D["heating_requests"]], boiler_sp[b], diag[b, D["tank_temp"]]))]
      zone_ids = env.building.zone_ids if plan is env.building.plans[0] else [
          "zone_id_" + n.replace("room_", "") for n in plan.zone_names]
      for zi, name in enumerate(plan.zone_names):
        zid = zone_ids[zi]
        occ = self._occ[s1, zi if self._occ.shape[1] > 1 else 0]
        zones[zid] = (window[0], window[1], zone_mean[b, zi], cfg.vav_max_air_flow_rate, flow, occ)
        damper = 1.0 if mode[b, zi] in (1, 2) else 0.1         # vav.py:230-243
        obs += [("vav_" + name, n, v) for n, v in zip(
            VAV_FIELDS, (damper, cfg.vav_max_air_flow_rate, pre_mean[b, zi]))]
      w.write(REWARD_INFO_PREFIX, ts, encode_reward_info(
          ts, ts + dt, zones,
          {self.air_handler_id: (diag[b, D["blower_w"]], diag[b, D["ac_w"]])},
          {self.boiler_id: (diag[b, D["gas_w"]], diag[b, D["pump_w"]])}))
      resp = dict(productivity_reward=diag[b, D["productivity"]], total_occupancy=diag[b, D["total_occ"]],
                  normalized_productivity_regret=diag[b, D["regret"]],
                  normalized_energy_cost=diag[b, D["norm_cost"]],
                  normalized_carbon_emission=diag[b, D["norm_carbon"]])
      if reward is not None:
        resp["agent_reward_value"] = reward[b]
      w.write(REWARD_RESPONSE_PREFIX, ts, encode_reward_response(ts, ts + dt, **resp))
      w.write(OBSERVATION_RESPONSE_FILE_PREFIX, ts, encode_observation_response(ts, obs))
      # ActionResponse (environment.py:834-871): one SingleActionRequest per action field,
      # device id + BARE setpoint name + native value (kelvin), as the reference logs it
      native = {"supply_water_setpoint": boiler_sp[b],
                "supply_air_heating_temperature_setpoint": heat_sp[b],
                "supply_air_cooling_temperature_setpoint": cool_sp[b]}
      acts = []
      for device, setpoint in env._action_fields:
        dev = self.boiler_id if device == "boiler" else self.air_handler_id
        acts.append((dev, setpoint, float(native[setpoint])))
      w.write(ACTION_RESPONSE_FILE_PREFIX, ts - dt, encode_action_response(ts - dt, acts))
