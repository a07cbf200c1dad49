"""Episode logs in the reference's ProtoWriter format."""

import os

import numpy as np
import pandas as pd
import pytest

from oracle import refshim
from sbsim_b200 import episode_writer as ew

TS = pd.Timestamp("2023-07-06 09:05:00+00:00")
DT = pd.Timedelta(300, unit="s")
ZONES = {"zone_id_1": (294.0, 297.0, 293.25, 0.035, 0.21, 1.0),
         "zone_id_2": (294.0, 297.0, 295.5, 0.035, 0.21, 0.0)}
OBS = [("air_handler_id_0", "supply_air_flowrate_sensor", 0.21),
       ("boiler_id_0", "heating_request_count", 0.0),          # a zero in a oneof is still written
       ("vav_room_1", "zone_air_temperature_sensor", 293.25)]
ACTS = [("boiler_id_0", "supply_water_setpoint", 340.0),
        ("air_handler_id_0", "supply_air_heating_temperature_setpoint", 285.0)]


def _decode(buf):
  """Minimal wire-format reader: list of (field, wire_type, value)."""
  out, pos = [], 0
  while pos < len(buf):
    key, shift = 0, 0
    while True:
      b = buf[pos]; pos += 1
      key |= (b & 0x7F) << shift; shift += 7
      if not b & 0x80:
        break
    field, wt = key >> 3, key & 7
    if wt == 0:
      v, shift = 0, 0
      while True:
        b = buf[pos]; pos += 1
        v |= (b & 0x7F) << shift; shift += 7
        if not b & 0x80:
          break
    elif wt == 5:
      v = np.frombuffer(buf[pos:pos + 4], "<f4")[0]; pos += 4
    elif wt == 2:
      n, shift = 0, 0
      while True:
        b = buf[pos]; pos += 1
        n |= (b & 0x7F) << shift; shift += 7
        if not b & 0x80:
          break
      v = buf[pos:pos + n]; pos += n
    else:
      raise AssertionError(wt)
    out.append((field, wt, v))
  return out


def test_wire_format_round_trip_and_shard_layout(tmp_path):
  payload = ew.encode_reward_info(TS, TS + DT, ZONES, {"air_handler_id_0": (1200.0, -350.5)},
                                  {"boiler_id_0": (5000.0, 12.5)})
  top = _decode(payload)
  assert [f for f, _, _ in top] == [1, 2, 5, 5, 6, 7]
  assert _decode(top[0][2]) == [(1, 0, int(TS.timestamp()))]
  entry = _decode(top[2][2])
  assert entry[0][2] == b"zone_id_1"
  vals = [v for _, _, v in _decode(entry[1][2])]
  np.testing.assert_array_equal(vals, np.float32(ZONES["zone_id_1"]))
  # proto3: the zero occupancy of zone 2 is not on the wire
  assert [f for f, _, _ in _decode(_decode(top[3][2])[1][2])] == [1, 2, 3, 4, 5]
  w = ew.ProtoShardWriter(str(tmp_path))
  w.write(ew.REWARD_INFO_PREFIX, TS, payload)
  w.write(ew.REWARD_INFO_PREFIX, TS + DT, payload)
  path = os.path.join(str(tmp_path), "reward_info_2023.07.06.09")     # controller_writer.py:118-123
  assert os.path.exists(path)
  assert list(ew.read_shard(path)) == [payload, payload]
  with pytest.raises(ValueError):
    ew.encode_reward_response(TS, TS + DT, not_a_field=1.0)


@pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted")
def test_messages_parse_with_the_reference_proto_classes():
  """The bytes are read back by the reference's own message classes (built at runtime from
  its .proto files) and every field lands where the reference's readers expect it."""
  refshim.install()
  from smart_buildings.smart_control.proto import smart_control_building_pb2 as bpb
  from smart_buildings.smart_control.proto import smart_control_reward_pb2 as rpb
  info = rpb.RewardInfo()
  info.ParseFromString(ew.encode_reward_info(
      TS, TS + DT, ZONES, {"air_handler_id_0": (1200.0, -350.5)}, {"boiler_id_0": (5000.0, 12.5)},
      agent_id="a", scenario_id="s"))
  assert info.start_timestamp.seconds == int(TS.timestamp())
  assert info.end_timestamp.seconds == int((TS + DT).timestamp())
  assert (info.agent_id, info.scenario_id) == ("a", "s")
  assert sorted(info.zone_reward_infos) == ["zone_id_1", "zone_id_2"]
  z = info.zone_reward_infos["zone_id_1"]
  got = (z.heating_setpoint_temperature, z.cooling_setpoint_temperature, z.zone_air_temperature,
         z.air_flow_rate_setpoint, z.air_flow_rate, z.average_occupancy)
  np.testing.assert_array_equal(np.float32(got), np.float32(ZONES["zone_id_1"]))
  assert info.zone_reward_infos["zone_id_2"].average_occupancy == 0.0
  ah = info.air_handler_reward_infos["air_handler_id_0"]
  assert (ah.blower_electrical_energy_rate, ah.air_conditioning_electrical_energy_rate) == (1200.0, -350.5)
  bl = info.boiler_reward_infos["boiler_id_0"]
  assert (bl.natural_gas_heating_energy_rate, bl.pump_electrical_energy_rate) == (5000.0, 12.5)
  # the reference re-serializes to the same number of bytes (same fields present)
  assert info.ByteSize() == len(ew.encode_reward_info(
      TS, TS + DT, ZONES, {"air_handler_id_0": (1200.0, -350.5)}, {"boiler_id_0": (5000.0, 12.5)},
      agent_id="a", scenario_id="s"))

  resp = rpb.RewardResponse()
  resp.ParseFromString(ew.encode_reward_response(
      TS, TS + DT, agent_reward_value=-0.125, productivity_reward=50.0, total_occupancy=2.0,
      normalized_energy_cost=0.25))
  assert (resp.agent_reward_value, resp.productivity_reward, resp.total_occupancy,
          resp.normalized_energy_cost) == (-0.125, 50.0, 2.0, 0.25)
  assert resp.end_timestamp.seconds - resp.start_timestamp.seconds == 300

  o = bpb.ObservationResponse()
  raw = ew.encode_observation_response(TS, OBS)
  o.ParseFromString(raw)
  assert o.timestamp.seconds == int(TS.timestamp())
  assert len(o.request.single_observation_requests) == len(o.single_observation_responses) == 3
  for single, (d, m, v) in zip(o.single_observation_responses, OBS):
    assert single.single_observation_request.device_id == d
    assert single.single_observation_request.measurement_name == m
    assert single.observation_valid and single.WhichOneof("observation_value") == "continuous_value"
    assert single.continuous_value == np.float32(v)
  assert o.SerializeToString() == raw        # no maps here: byte-identical to the reference's

  a = bpb.ActionResponse()
  raw = ew.encode_action_response(TS, ACTS)
  a.ParseFromString(raw)
  assert [r.request.setpoint_name for r in a.single_action_responses] == [x[1] for x in ACTS]
  assert all(r.response_type == bpb.SingleActionResponse.ACCEPTED for r in a.single_action_responses)
  assert a.request.single_action_requests[0].continuous_value == 340.0
  assert a.SerializeToString() == raw


@pytest.mark.gpu
def test_episode_writer_on_the_cuda_environment(tmp_path):
  import scenarios as S
  sc = S.Scenario(floor_plan=S.small_plan(), occupancy="step", start="2023-07-06 08:50:00")
  env = S.make_env(sc, n_envs=3)
  try:
    wr = ew.EpisodeWriter(env, str(tmp_path), env_indices=(0, 2))
    env.reset()
    rng = np.random.default_rng(0)
    rewards, zmeans = [], []
    for _ in range(4):
      a = rng.uniform(-1, 1, (3, 2)).astype(np.float32)
      ts = env.step(a)
      wr.log_step(a, ts)
      rewards.append(float(ts.reward[2]))
      zmeans.append(env.handle.download("zone_mean", (3, env.building.n_zones))[2].copy())
    d = os.path.join(str(tmp_path), "env_2")
    files = sorted(os.listdir(d))
    assert any(f.startswith("reward_info_2023.07.06") for f in files)
    assert {f.split("_2023")[0] for f in files} == {"reward_info", "reward_response",
                                                    "observation_response", "action_response"}
    resp = [r for f in files if f.startswith("reward_response") for r in ew.read_shard(os.path.join(d, f))]
    assert len(resp) == 4
    got = [v for f, wt, v in _decode(resp[-1]) if f == 1 and wt == 5][0]
    assert got == np.float32(rewards[-1])
    infos = [r for f in files if f.startswith("reward_info") for r in ew.read_shard(os.path.join(d, f))]
    zones = [_decode(v) for f, wt, v in _decode(infos[-1]) if f == 5]
    temps = [[x for ff, _, x in _decode(z[1][2]) if ff == 3][0] for z in zones]
    np.testing.assert_array_equal(np.float32(temps), zmeans[-1][:len(temps)])
    # ActionResponse: bare setpoint names and NATIVE values (kelvin), device order boiler, AHU
    # (environment.py:834-871); the last action was `a`
    acts = [r for f in files if f.startswith("action_response") for r in ew.read_shard(os.path.join(d, f))]
    assert len(acts) == 4
    singles = [_decode(v) for f, wt, v in _decode(acts[-1]) if f == 3]
    got_acts = []
    for sresp in singles:
      req = _decode([v for f, wt, v in sresp if f == 1][0])
      got_acts.append(([v for f, _, v in req if f == 1][0].decode(), [v for f, _, v in req if f == 2][0].decode(),
                       float([v for f, wt, v in req if f == 3 and wt == 5][0])))
    assert [g[1] for g in got_acts] == ["supply_water_setpoint", "supply_air_heating_temperature_setpoint"]
    assert got_acts[0][0].startswith("boiler") and got_acts[1][0].startswith("air_handler")
    want = [np.float32(np.float32((a[2, 0] + 1) / 2) * np.float32(45.0) + np.float32(310.0)),
            np.float32(np.float32((a[2, 1] + 1) / 2) * np.float32(15.0) + np.float32(285.0))]
    np.testing.assert_allclose([g[2] for g in got_acts], want, rtol=1e-6)
  finally:
    env.close()


@pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted")
def test_shards_read_back_with_the_reference_proto_reader(tmp_path):
  """Shards written by ProtoShardWriter are read by the reference's own ProtoReader
  (utils/controller_reader.py:39-175): file names, record framing and messages."""
  refshim.install()
  from smart_buildings.smart_control.utils import controller_reader
  w = ew.ProtoShardWriter(str(tmp_path))
  for k in range(3):                        # 09:05, 09:55, 10:45 UTC: two hourly shards
    ts = TS + k * pd.Timedelta(50, unit="min")
    w.write(ew.REWARD_INFO_PREFIX, ts, ew.encode_reward_info(
        ts, ts + DT, ZONES, {"air_handler_id_0": (1200.0 + k, -350.5)}, {"boiler_id_0": (5000.0, 12.5)}))
    w.write(ew.REWARD_RESPONSE_PREFIX, ts, ew.encode_reward_response(ts, ts + DT, agent_reward_value=-0.25 * k))
    w.write(ew.OBSERVATION_RESPONSE_FILE_PREFIX, ts, ew.encode_observation_response(ts, OBS))
    w.write(ew.ACTION_RESPONSE_FILE_PREFIX, ts, ew.encode_action_response(ts, ACTS))
  reader = controller_reader.ProtoReader(str(tmp_path))
  naive = TS.tz_localize(None)             # shard serials are wall-clock hours without a zone
  t0, t1 = naive - pd.Timedelta(1, unit="h"), naive + pd.Timedelta(3, unit="h")
  infos = reader.read_reward_infos(t0, t1)
  assert [i.air_handler_reward_infos["air_handler_id_0"].blower_electrical_energy_rate for i in infos] == [
      1200.0, 1201.0, 1202.0]
  assert [r.agent_reward_value for r in reader.read_reward_responses(t0, t1)] == [0.0, -0.25, -0.5]
  obs = reader.read_observation_responses(t0, t1)
  assert len(obs) == 3 and obs[0].single_observation_responses[2].continuous_value == np.float32(293.25)
  acts = reader.read_action_responses(t0, t1)
  assert [a.request.single_action_requests[0].setpoint_name for a in acts] == ["supply_water_setpoint"] * 3
  # the reader selects shards by hour: a window that ends before 10:00 sees the first two records
  assert len(reader.read_reward_infos(t0, naive + pd.Timedelta(30, unit="min"))) == 2


@pytest.mark.gpu
def test_environment_metrics_path_writes_the_episode_logs(tmp_path):
  """Environment(metrics_path=...) (environment.py:362, 1181-1198): reset() opens
  <metrics_path>/<label>_<stamp>/ and every step() logs the chosen buildings; the reference's
  other constructor arguments are accepted."""
  import scenarios as S
  sc = S.Scenario(floor_plan=S.small_plan(), occupancy="step", start="2023-07-06 08:50:00")
  with pytest.warns(UserWarning, match="device_action_tuples"):
    env = S.make_env(sc, n_envs=2, metrics_path=str(tmp_path), label="run", metrics_env_indices=(1,),
                     device_action_tuples=[("BLR", "supply_water_setpoint")], metrics_reporting_interval=10)
  try:
    env.reset()
    rng = np.random.default_rng(1)
    for _ in range(3):
      ts = env.step(rng.uniform(-1, 1, (2, 2)).astype(np.float32))
    runs = os.listdir(str(tmp_path))
    assert len(runs) == 1 and runs[0].startswith("run_")
    d = os.path.join(str(tmp_path), runs[0], "env_1")
    files = sorted(os.listdir(d))
    assert {f.split("_2023")[0] for f in files} == {"reward_info", "reward_response",
                                                    "observation_response", "action_response"}
    resp = [r for f in files if f.startswith("reward_response") for r in ew.read_shard(os.path.join(d, f))]
    assert len(resp) == 3
    got = [v for f, wt, v in _decode(resp[-1]) if f == 1 and wt == 5][0]
    assert got == np.float32(ts.reward[1])
  finally:
    env.close()
