"""Oracle: HVAC device algebra (TEST INFRASTRUCTURE ONLY).

Scalar Python restatement of the reference devices, same expressions and the
same evaluation order so NumPy scalar-type promotion behaves identically:
  SetpointSchedule  simulator/setpoint_schedule.py:29-128
  Thermostat        simulator/thermostat.py:39-148
  Vav               simulator/vav.py:29-264
  AirHandler        simulator/air_handler.py:29-320
  Boiler            simulator/boiler.py:30-333
Paths are relative to /root/reference/smart_control/.
"""

from __future__ import annotations

import datetime
import zoneinfo

import numpy as np
import pandas as pd

AIR_HEAT_CAPACITY = 1006.0     # utils/constants.py:21
WATER_HEAT_CAPACITY = 4180.0   # utils/constants.py:22
WATER_DENSITY = 1000.0         # utils/constants.py:46
GRAVITY = 9.8                  # utils/constants.py:47

_TZ_ALIAS = {"US/Pacific": "America/Los_Angeles", "US/Eastern": "America/New_York"}


def tzinfo(name):
  if name is None or name == "UTC":
    return datetime.timezone.utc
  if isinstance(name, datetime.tzinfo):
    return name
  return zoneinfo.ZoneInfo(_TZ_ALIAS.get(name, name))


class SetpointSchedule:
  """setpoint_schedule.py:55-128."""

  def __init__(self, morning_start_hour, evening_start_hour, comfort_temp_window,
               eco_temp_window, holidays=None, time_zone="UTC"):
    if morning_start_hour > evening_start_hour:
      raise ValueError("morning_start_hour must be less than evening_start_hour")
    self.morning_start_hour = morning_start_hour
    self.evening_start_hour = evening_start_hour
    self.comfort_temp_window = comfort_temp_window
    self.eco_temp_window = eco_temp_window
    self.holidays = set(holidays) if holidays else set()
    self._tz = tzinfo(time_zone)

  def _local(self, ts: pd.Timestamp) -> pd.Timestamp:          # :101-107
    if ts.tz is not None:
      return ts.tz_convert(self._tz)
    return ts.tz_localize(datetime.timezone.utc)

  def is_weekend(self, ts) -> bool:                            # :109-116
    return self._local(ts).day_name() in ("Saturday", "Sunday")

  def is_comfort_mode(self, ts) -> bool:                       # :86-99
    lt = self._local(ts)
    return (lt.hour >= self.morning_start_hour
            and lt.hour < self.evening_start_hour
            and lt.dayofyear not in self.holidays
            and not self.is_weekend(lt))

  def get_temperature_window(self, ts):                        # :118-128
    return (self.comfort_temp_window if self.is_comfort_mode(ts)
            else self.eco_temp_window)


OFF, HEAT, COOL, PASSIVE_COOL = 0, 1, 2, 3   # thermostat.py:63-66


class Thermostat:
  """thermostat.py:68-148.  Not reset by Vav.reset (vav.py:98)."""

  def __init__(self, schedule: SetpointSchedule):
    self.schedule = schedule
    self.previous_timestamp = None
    self.mode = OFF

  def _default_control(self, zone_temp, window):               # :76-112
    heat, cool = window
    mid = 0.5 * (cool - heat) + heat
    if zone_temp < heat:
      self.mode = HEAT
    elif zone_temp > cool:
      self.mode = COOL
    elif zone_temp < mid and self.mode == HEAT:
      self.mode = HEAT
    elif zone_temp > mid and self.mode == COOL:
      self.mode = COOL
    else:
      self.mode = OFF
    return self.mode

  def update(self, zone_temp, ts):                             # :114-148
    window = self.schedule.get_temperature_window(ts)
    if self.schedule.is_comfort_mode(ts):
      self._default_control(zone_temp, window)
    elif (self.previous_timestamp is not None
          and self.schedule.is_comfort_mode(self.previous_timestamp)):
      self.mode = PASSIVE_COOL
    else:
      if self.mode == PASSIVE_COOL and zone_temp > window[0]:
        self.mode = PASSIVE_COOL
      else:
        self._default_control(zone_temp, window)
    self.previous_timestamp = ts
    return self.mode


class Boiler:
  """boiler.py:53-333."""

  def __init__(self, reheat_water_setpoint, water_pump_differential_head,
               water_pump_efficiency, heating_rate=0, cooling_rate=0,
               convection_coefficient=5.6, tank_length=2.0, tank_radius=0.5,
               water_capacity=1.5, insulation_conductivity=0.067,
               insulation_thickness=0.06):
    self._init_setpoint = reheat_water_setpoint
    self.head = water_pump_differential_head
    self.efficiency = water_pump_efficiency
    self.heating_rate = heating_rate
    self.cooling_rate = cooling_rate
    self.convection_coefficient = convection_coefficient
    self.tank_length = tank_length
    self.tank_radius = tank_radius
    self.water_capacity = water_capacity
    self.insulation_conductivity = insulation_conductivity
    self.insulation_thickness = insulation_thickness
    # smart_device.py:72-73 -- never cleared by reset()
    self.action_timestamp = None
    self.observation_timestamp = None
    self.reset()

  def reset(self):                                             # :112-123
    self.reset_demand()
    self.reheat_water_setpoint = self._init_setpoint
    self.return_water_temperature_sensor = 0.0
    self.current_temperature = self._init_setpoint
    self.step_tank_temperature_change = 0.0
    self.last_step_duration_sec = 0.0

  def reset_demand(self):                                      # :154-156
    self.total_flow_rate = 0.0
    self.heating_request_count = 0

  def add_demand(self, flow_rate):                             # :219-231
    if flow_rate <= 0:
      raise ValueError("Flow rate must be positive")
    self.total_flow_rate += flow_rate
    self.heating_request_count += 1

  def observe_supply_water_temperature(self, observation_timestamp):
    """supply_water_temperature_sensor :146-148 via get_observation."""
    self.observation_timestamp = observation_timestamp
    self._set_current_temperature()
    return self.current_temperature

  def _set_current_temperature(self):                          # :158-183
    if self.action_timestamp:
      self.last_step_duration_sec = (
          self.observation_timestamp - self.action_timestamp).total_seconds()
    else:
      self.action_timestamp = self.observation_timestamp
    if self.action_timestamp and self.cooling_rate > 0.0 and self.heating_rate > 0.0:
      begin = self.current_temperature
      self.current_temperature = self._adjust_temperature(
          self.reheat_water_setpoint, begin, self.last_step_duration_sec)
      self.step_tank_temperature_change = self.current_temperature - begin
    else:
      self.current_temperature = self.reheat_water_setpoint

  def _adjust_temperature(self, setpoint, actual, seconds):    # :185-217
    if setpoint > actual:
      return min(actual + self.heating_rate * seconds / 60.0, setpoint)
    elif setpoint < actual:
      return max(actual - self.cooling_rate * seconds / 60.0, setpoint)
    return setpoint

  def compute_thermal_energy_rate(self, return_water_temp, outside_temp):  # :233-273
    if self.reheat_water_setpoint > return_water_temp:
      supply = self.reheat_water_setpoint
    else:
      supply = return_water_temp
    flow_rate = WATER_HEAT_CAPACITY * self.total_flow_rate * (
        supply - return_water_temp)
    dissipation = self.compute_thermal_dissipation_rate(supply, outside_temp)
    if self.last_step_duration_sec > 0:
      tank = (WATER_HEAT_CAPACITY * self.water_capacity
              * self.step_tank_temperature_change / self.last_step_duration_sec)
    else:
      tank = 0
    return flow_rate + dissipation + tank

  def compute_thermal_dissipation_rate(self, water_temp, outside_temp):    # :275-320
    assert water_temp >= outside_temp
    delta = water_temp - outside_temp
    numerator = self.tank_length * 2.0 * np.pi * delta
    r1 = self.tank_radius
    r2 = r1 + self.insulation_thickness
    conduction = np.log(r2 / r1) / self.insulation_conductivity
    convection = 1.0 / self.convection_coefficient / r2
    return numerator / (conduction + convection)

  def compute_pump_power(self):                                # :322-333
    return (self.total_flow_rate * WATER_DENSITY * GRAVITY * self.head
            / self.efficiency)


class AirHandler:
  """air_handler.py:51-320."""

  def __init__(self, recirculation, heating_air_temp_setpoint,
               cooling_air_temp_setpoint, fan_differential_pressure,
               fan_efficiency, max_air_flow_rate=8.67):
    if cooling_air_temp_setpoint <= heating_air_temp_setpoint:
      raise ValueError("cooling_air_temp_setpoint must greater than"
                       " heating_air_temp_setpoint")
    self._init = (recirculation, heating_air_temp_setpoint,
                  cooling_air_temp_setpoint, fan_differential_pressure,
                  fan_efficiency, max_air_flow_rate)
    self.reset()

  def reset(self):                                             # :127-135
    (self.recirculation, self.heating_air_temp_setpoint,
     self.cooling_air_temp_setpoint, self.fan_differential_pressure,
     self.fan_efficiency, self.max_air_flow_rate) = self._init
    self.air_flow_rate = 0.0
    self.cooling_request_count = 0

  def get_mixed_air_temp(self, recirculation_temp, ambient_temp):  # :204-216
    return (self.recirculation * recirculation_temp
            + (1 - self.recirculation) * ambient_temp)

  def get_supply_air_temp(self, recirculation_temp, ambient_temp):  # :218-233
    mixed = self.get_mixed_air_temp(recirculation_temp, ambient_temp)
    if mixed > self.cooling_air_temp_setpoint:
      return self.cooling_air_temp_setpoint
    elif mixed < self.heating_air_temp_setpoint:
      return self.heating_air_temp_setpoint
    return mixed

  @property
  def ambient_flow_rate(self):                                 # :236-238
    return (1.0 - self.recirculation) * self.air_flow_rate

  @property
  def supply_fan_speed_percentage(self):                       # :246-248
    return self.air_flow_rate / self.max_air_flow_rate

  def reset_demand(self):                                      # :250-252
    self.air_flow_rate = 0.0
    self.cooling_request_count = 0

  def add_demand(self, flow_rate):                             # :254-268
    if flow_rate <= 0:
      raise ValueError("Flow rate must be positive")
    self.air_flow_rate += flow_rate
    if self.air_flow_rate > self.max_air_flow_rate:
      self.air_flow_rate = self.max_air_flow_rate
    self.cooling_request_count += 1

  def compute_thermal_energy_rate(self, recirculation_temp, ambient_temp):  # :270-285
    mixed = self.get_mixed_air_temp(recirculation_temp, ambient_temp)
    supply = self.get_supply_air_temp(recirculation_temp, ambient_temp)
    return self.air_flow_rate * AIR_HEAT_CAPACITY * (supply - mixed)

  def compute_fan_power(self, flow_rate, dp, eff):             # :287-302
    return flow_rate * dp / eff

  def compute_intake_fan_energy_rate(self):                    # :304-310
    return self.compute_fan_power(self.air_flow_rate,
                                  self.fan_differential_pressure,
                                  self.fan_efficiency)

  def compute_exhaust_fan_energy_rate(self):                   # :312-320
    return self.compute_fan_power(self.air_flow_rate * (1.0 - self.recirculation),
                                  self.fan_differential_pressure,
                                  self.fan_efficiency)


class Vav:
  """vav.py:48-264."""

  def __init__(self, max_air_flow_rate, reheat_max_water_flow_rate,
               therm: Thermostat, boiler: Boiler):
    self.max_air_flow_rate = max_air_flow_rate
    self.reheat_max_water_flow_rate = reheat_max_water_flow_rate
    self.thermostat = therm
    self.boiler = boiler
    self.reset()

  def reset(self):                                             # :93-99
    self.reheat_valve_setting = 0.0
    self.damper_setting = 0.1
    self.zone_air_temperature = 0

  @property
  def flow_rate_demand(self):                                  # :139-141
    return self.damper_setting * self.max_air_flow_rate

  @property
  def reheat_demand(self):                                     # :143-145
    return self.reheat_valve_setting * self.reheat_max_water_flow_rate

  def compute_zone_supply_temp(self, supply_air_temp, input_water_temp):  # :168-195
    assert self.damper_setting > 0
    reheat_flow_rate = self.reheat_valve_setting * self.reheat_max_water_flow_rate
    air_flow_rate = self.damper_setting * self.max_air_flow_rate
    heat_difference = (AIR_HEAT_CAPACITY * air_flow_rate
                       - WATER_HEAT_CAPACITY * reheat_flow_rate)
    input_water_heat = input_water_temp * WATER_HEAT_CAPACITY * reheat_flow_rate
    return ((supply_air_temp * heat_difference + input_water_heat)
            / air_flow_rate / AIR_HEAT_CAPACITY)

  def compute_energy_applied_to_zone(self, zone_temp, supply_air_temp,
                                     input_water_temp):        # :197-217
    if self.damper_setting == 0 or self.max_air_flow_rate == 0:
      return 0
    zone_supply_temp = self.compute_zone_supply_temp(supply_air_temp,
                                                     input_water_temp)
    air_flow_rate = self.damper_setting * self.max_air_flow_rate
    return air_flow_rate * AIR_HEAT_CAPACITY * (zone_supply_temp - zone_temp)

  def update_settings(self, zone_temp, ts):                    # :219-243
    self.zone_air_temperature = zone_temp
    mode = self.thermostat.update(zone_temp, ts)
    if mode == HEAT:
      self.damper_setting, self.reheat_valve_setting = 1.0, 1.0
    elif mode == COOL:
      self.damper_setting, self.reheat_valve_setting = 1.0, 0.0
    else:  # OFF, PASSIVE_COOL
      self.damper_setting, self.reheat_valve_setting = 0.1, 0.0

  def output(self, zone_temp, supply_air_temp):                # :245-264
    self.zone_air_temperature = zone_temp
    q_zone = self.compute_energy_applied_to_zone(
        zone_temp, supply_air_temp, self.boiler.reheat_water_setpoint)
    temp_vav_supply = self.compute_zone_supply_temp(
        supply_air_temp, self.boiler.reheat_water_setpoint)
    return q_zone, temp_vav_supply
