"""Oracle: SetpointEnergyCarbonRegretFunction (TEST INFRASTRUCTURE ONLY).

Restates /root/reference/smart_control/reward/setpoint_energy_carbon_regret.py:142-291
and base_setpoint_energy_carbon_reward.py:54-172.  All RewardInfo scalars are
proto `float` (smart_control_reward.proto:14-35), i.e. rounded to fp32 before the
reward function sees them; `f32()` marks those roundings.
"""

from __future__ import annotations

import dataclasses
from typing import List

import numpy as np
import pandas as pd

from oracle import exogenous


def f32(x) -> float:
  """proto `float` field round trip: Python float -> fp32 -> Python float."""
  return float(np.float32(x))


@dataclasses.dataclass
class ZoneRewardInfo:                      # smart_control_reward.proto:14-21
  heating_setpoint_temperature: float
  cooling_setpoint_temperature: float
  zone_air_temperature: float
  air_flow_rate_setpoint: float
  air_flow_rate: float
  average_occupancy: float


@dataclasses.dataclass
class RewardInfo:
  start_timestamp: pd.Timestamp
  end_timestamp: pd.Timestamp
  zones: List[ZoneRewardInfo]
  blower_electrical_energy_rate: float
  air_conditioning_electrical_energy_rate: float
  natural_gas_heating_energy_rate: float
  pump_electrical_energy_rate: float


@dataclasses.dataclass
class RewardResponse:                       # smart_control_reward.proto:53-85
  agent_reward_value: float
  productivity_reward: float
  electricity_energy_cost: float
  natural_gas_energy_cost: float
  carbon_emitted: float
  total_occupancy: float
  productivity_regret: float
  normalized_productivity_regret: float
  normalized_energy_cost: float
  normalized_carbon_emission: float
  carbon_cost: float = 0.0


class SetpointEnergyCarbonRegretFunction:

  def __init__(self, max_productivity_personhour_usd, min_productivity_personhour_usd,
               max_electricity_rate, max_natural_gas_rate,
               productivity_midpoint_delta, productivity_decay_stiffness,
               electricity_energy_cost, natural_gas_energy_cost,
               productivity_weight, energy_cost_weight, carbon_emission_weight):
    self.pmax = max_productivity_personhour_usd
    self.pmin = min_productivity_personhour_usd
    self.emax = max_electricity_rate
    self.gmax = max_natural_gas_rate
    self.delta = productivity_midpoint_delta
    self.stiff = productivity_decay_stiffness
    self.elec = electricity_energy_cost
    self.gas = natural_gas_energy_cost
    self.u = productivity_weight
    self.v = energy_cost_weight
    self.w = carbon_emission_weight
    assert self.pmax > self.pmin

  def _zone_productivity(self, heat, cool, temp, dt, occ):     # base:83-123
    x0low = heat - self.delta
    x0high = cool + self.delta
    if temp < heat:
      p = self.pmax / (1.0 + np.exp(-self.stiff * (temp - x0low)))
    elif temp > cool:
      p = self.pmax * (1.0 - 1.0 / (1.0 + np.exp(-self.stiff * (temp - x0high))))
    else:
      p = self.pmax
    return p * occ * dt / 3600.0

  def compute_reward(self, info: RewardInfo) -> RewardResponse:  # regret:142-291
    start, end = info.start_timestamp, info.end_timestamp
    dt = (end - start).total_seconds()
    actual = 0.0
    total_occ = 0.0
    for z in info.zones:                                       # base:54-81
      total_occ += z.average_occupancy
      actual += self._zone_productivity(
          z.heating_setpoint_temperature, z.cooling_setpoint_temperature,
          z.zone_air_temperature, dt, z.average_occupancy)
    max_p = self.pmax * total_occ * dt / 3600.0
    min_p = self.pmin * total_occ * dt / 3600.0
    actual = max(actual, min_p)
    if total_occ > 0.0:
      regret = (actual - min_p) / (max_p - min_p) - 1.0
    else:
      regret = 0.0
    elec_rate = (info.blower_electrical_energy_rate
                 + np.abs(info.air_conditioning_electrical_energy_rate)
                 + info.pump_electrical_energy_rate)           # base:138-158
    capped_e = min(elec_rate, self.emax)
    cost_e = self.elec.cost(start, end, capped_e)
    cost_e_max = self.elec.cost(start, end, self.emax)
    carb_e = self.elec.carbon(start, end, capped_e)
    carb_e_max = self.elec.carbon(start, end, self.emax)
    capped_g = min(info.natural_gas_heating_energy_rate, self.gmax)
    cost_g = self.gas.cost(start, end, capped_g)
    cost_g_max = self.gas.cost(start, end, self.gmax)
    carb_g = self.gas.carbon(start, end, capped_g)
    carb_g_max = self.gas.carbon(start, end, self.gmax)
    n_cost = (cost_e + cost_g) / (cost_e_max + cost_g_max)
    n_carbon = (carb_e + carb_g) / (carb_e_max + carb_g_max)
    raw = regret * self.u - n_cost * self.v - n_carbon * self.w
    value = raw / (self.u + self.v + self.w)
    return RewardResponse(
        agent_reward_value=f32(value),
        productivity_reward=f32(actual),
        electricity_energy_cost=f32(cost_e),
        natural_gas_energy_cost=f32(cost_g),
        carbon_emitted=f32(carb_e + carb_g),
        total_occupancy=f32(total_occ),
        productivity_regret=f32(actual - max_p),
        normalized_productivity_regret=f32(regret),
        normalized_energy_cost=f32(n_cost),
        normalized_carbon_emission=f32(n_carbon),
    )


class SetpointEnergyCarbonRewardFunction(SetpointEnergyCarbonRegretFunction):
  """/root/reference/smart_control/reward/setpoint_energy_carbon_reward.py:103-190:
  productivity - w_e * (electricity + gas cost) - w_c * carbon cost, shifted and scaled.
  Shares the zone-productivity logistic with the regret function (base:54-123)."""

  def __init__(self, max_productivity_personhour_usd, productivity_midpoint_delta,
               productivity_decay_stiffness, electricity_energy_cost,
               natural_gas_energy_cost, energy_cost_weight, carbon_cost_weight,
               carbon_cost_factor, reward_normalizer_shift=0.0, reward_normalizer_scale=1.0):
    self.pmax = max_productivity_personhour_usd
    self.delta = productivity_midpoint_delta
    self.stiff = productivity_decay_stiffness
    self.elec = electricity_energy_cost
    self.gas = natural_gas_energy_cost
    self.v = energy_cost_weight
    self.w = carbon_cost_weight
    self.factor = carbon_cost_factor
    self.shift = reward_normalizer_shift
    self.scale = reward_normalizer_scale

  def compute_reward(self, info: RewardInfo) -> RewardResponse:   # :127-190
    start, end = info.start_timestamp, info.end_timestamp
    dt = (end - start).total_seconds()
    productivity = 0.0
    total_occ = 0.0
    for z in info.zones:                                         # base:54-81
      total_occ += z.average_occupancy
      productivity += self._zone_productivity(
          z.heating_setpoint_temperature, z.cooling_setpoint_temperature,
          z.zone_air_temperature, dt, z.average_occupancy)
    elec_rate = (info.blower_electrical_energy_rate
                 + np.abs(info.air_conditioning_electrical_energy_rate)
                 + info.pump_electrical_energy_rate)             # base:125-148
    cost_e = self.elec.cost(start, end, elec_rate)               # :141-150
    carb_e = self.elec.carbon(start, end, elec_rate)
    gas_rate = info.natural_gas_heating_energy_rate              # :152-164
    cost_g = self.gas.cost(start, end, gas_rate)
    carb_g = self.gas.carbon(start, end, gas_rate)
    combined = carb_e + carb_g
    carbon_cost = f32(combined * self.factor)                    # :174-176 proto float, read back below
    raw = productivity - self.v * (cost_e + cost_g) - self.w * carbon_cost   # :178-183
    value = (raw - self.shift) / self.scale                      # :185-187
    return RewardResponse(
        agent_reward_value=f32(value), productivity_reward=f32(productivity),
        electricity_energy_cost=f32(cost_e), natural_gas_energy_cost=f32(cost_g),
        carbon_emitted=f32(combined), total_occupancy=f32(total_occ),
        productivity_regret=0.0, normalized_productivity_regret=0.0,
        normalized_energy_cost=0.0, normalized_carbon_emission=0.0,
        carbon_cost=carbon_cost)
