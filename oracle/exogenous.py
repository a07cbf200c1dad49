"""Oracle: exogenous inputs of a step (TEST INFRASTRUCTURE ONLY).

Scalar restatements, paths relative to /root/reference/smart_control/:
  WeatherController / ReplayWeatherController  simulator/weather_controller.py:47-218
  StepFunctionOccupancy                        simulator/step_function_occupancy.py:36-173
  ElectricityEnergyCost                        reward/electricity_energy_cost.py:29-224
  NaturalGasEnergyCost                         reward/natural_gas_energy_cost.py:28-138
  is_work_day / get_radian_time                utils/conversion_utils.py:62-134
  expand_time_features                         utils/regression_building_utils.py:97-125

`holidays.US()` (third-party `holidays`, unpinned, absent from the image) is
restated with pandas' USFederalHolidayCalendar; the two agree on every federal
holiday, which is all `is_work_day` can observe on the dates used here.
"""

from __future__ import annotations

import functools
import math

import numpy as np
import pandas as pd

_SECONDS_IN_A_DAY = 24 * 3600
_EPOCH = pd.Timestamp("1970-01-01", tz="UTC")


# ----------------------------------------------------------------------------
# weather
# ----------------------------------------------------------------------------


class WeatherController:
  """Sinusoid: low at midnight, high at noon.  weather_controller.py:47-132."""

  def __init__(self, default_low_temp, default_high_temp, special_days=None,
               convection_coefficient=12.0):
    if default_low_temp > default_high_temp:
      raise ValueError("default_low_temp cannot be greater than default_high_temp.")
    self.default_low_temp = default_low_temp
    self.default_high_temp = default_high_temp
    self.special_days = special_days if special_days else {}
    self.convection_coefficient = convection_coefficient

  def seconds_to_rads(self, seconds_in_day):                   # :83-91
    min_rad, max_rad = -math.pi / 2.0, 3.0 * math.pi / 2.0
    return (seconds_in_day / _SECONDS_IN_A_DAY) * (max_rad - min_rad) + min_rad

  def get_current_temp(self, ts: pd.Timestamp) -> float:       # :93-123
    today = ts.dayofyear
    tomorrow = (today + 1) % 365
    if today in self.special_days:
      today_low, today_high = self.special_days[today]
    else:
      today_low, today_high = self.default_low_temp, self.default_high_temp
    if tomorrow in self.special_days:
      tomorrow_low, _ = self.special_days[tomorrow]
    else:
      tomorrow_low = self.default_low_temp
    high = today_high
    low = today_low if ts.hour < 12 else tomorrow_low
    seconds_in_day = (ts - pd.Timestamp(ts.date())).total_seconds()
    rad = self.seconds_to_rads(seconds_in_day)
    return 0.5 * (math.sin(rad) + 1) * (high - low) + low

  def get_air_convection_coefficient(self, ts) -> float:
    return self.convection_coefficient


class ReplayWeatherController:
  """Hourly CSV (columns Time, TempF) interpolated.  weather_controller.py:164-218."""

  def __init__(self, times_utc_sec: np.ndarray, temps_f: np.ndarray,
               convection_coefficient=12.0):
    self.times = np.asarray(times_utc_sec, dtype=np.float64)
    self.temps_f = np.asarray(temps_f, dtype=np.float64)
    self.convection_coefficient = convection_coefficient

  @classmethod
  def from_csv(cls, path, convection_coefficient=12.0):
    df = pd.read_csv(path)
    times = np.array([(pd.Timestamp(t, tz="UTC") - _EPOCH).total_seconds()
                      for t in df["Time"]])
    return cls(times, np.asarray(df["TempF"], dtype=np.float64),
               convection_coefficient)

  def get_current_temp(self, ts: pd.Timestamp) -> float:       # :188-214
    ts = ts.tz_convert("UTC")
    target = (ts - _EPOCH).total_seconds()
    if target < self.times.min():
      raise ValueError(f"Attempting to get weather data at {ts}, before the "
                       "latest timestamp.")
    if target > self.times.max():
      raise ValueError(f"Attempting to get weather data at {ts}, after the "
                       "latest timestamp.")
    temp_f = np.interp(target, self.times, self.temps_f)
    return fahrenheit_to_kelvin(temp_f)

  def get_air_convection_coefficient(self, ts) -> float:
    return self.convection_coefficient


def fahrenheit_to_kelvin(fahrenheit):                          # conversion_utils.py:155-170
  celsius = (fahrenheit - 32.0) * 5.0 / 9.0
  return celsius + 273.15


# ----------------------------------------------------------------------------
# calendar
# ----------------------------------------------------------------------------


@functools.cache
def _us_holidays():
  from pandas.tseries.holiday import USFederalHolidayCalendar
  cal = USFederalHolidayCalendar()
  return frozenset(d.date() for d in cal.holidays(start="2000-01-01",
                                                  end="2040-12-31"))


def is_work_day(ts: pd.Timestamp) -> bool:                     # conversion_utils.py:67-70
  return ts.weekday() < 5 and ts.date() not in _us_holidays()


def get_radian_time(ts: pd.Timestamp, interval: str) -> float:  # conversion_utils.py:107-134
  day_local = pd.Timestamp(year=ts.year, month=ts.month, day=ts.day, tz=ts.tz)
  if interval == "dow":
    frac = float(day_local.weekday()) / 7.0
  elif interval == "hod":
    frac = (ts - day_local).total_seconds() / 86400.0
  else:
    raise ValueError(interval)
  return 2.0 * np.pi * frac


def expand_time_features(n: int, rad: float):                  # regression_building_utils.py:97-125
  """Returns [cos_0..cos_{n-1}, sin_0..sin_{n-1}]."""
  phase = rad + (np.arange(n) / n * 2.0 * np.pi)
  return list(np.cos(phase)) + list(np.sin(phase))


# ----------------------------------------------------------------------------
# occupancy
# ----------------------------------------------------------------------------


class StepFunctionOccupancy:
  """step_function_occupancy.py:36-173 (zone-independent)."""

  def __init__(self, work_start_time: pd.Timedelta, work_end_time: pd.Timedelta,
               work_occupancy: float, nonwork_occupancy: float):
    self._work_start_time = work_start_time
    self._work_end_time = work_end_time
    self._work_occupancy = work_occupancy
    self._nonwork_occupancy = nonwork_occupancy

  def average_zone_occupancy(self, zone_id, start_time, end_time) -> float:  # :63-117
    if start_time >= end_time:
      raise ValueError("End time may not occur before start time.")
    work_seconds = 0.0
    nonwork_seconds = 0.0
    day = pd.Timestamp(year=start_time.year, month=start_time.month,
                       day=start_time.day)
    if start_time.tz is not None:
      # The reference subtracts a naive midnight from the timestamp, which only
      # works for tz-naive inputs ("local time w/o TZ", :78-79); callers with
      # tz-aware timestamps hit a TypeError there.  Keep that contract.
      raise TypeError("StepFunctionOccupancy needs tz-naive local timestamps")
    current_time = start_time - day
    while day <= end_time:
      day_end = min(pd.Timedelta(1, unit="day"), end_time - day)
      if is_work_day(day):
        before, during, after = self._split_workday(current_time, day_end)
        work_seconds += during
        nonwork_seconds += before + after
      else:
        nonwork_seconds += (day_end - current_time).total_seconds()
      day += pd.Timedelta(1.0, unit="day")
      current_time = pd.Timedelta(0.0, unit="sec")
    return (work_seconds * self._work_occupancy
            + nonwork_seconds * self._nonwork_occupancy) / (
                work_seconds + nonwork_seconds)

  def _split_workday(self, start_time, end_time):              # :119-165
    before = during = after = 0.0
    current = start_time
    interval_end = min(end_time, pd.Timedelta(24, unit="hour"))
    next_step = min(interval_end, self._work_start_time)
    if current < next_step:
      before = (next_step - current).total_seconds()
      current = max(current, next_step)
    next_step = min(interval_end, self._work_end_time)
    if current < next_step:
      during = (next_step - current).total_seconds()
      current = next_step
    next_step = interval_end
    if current < next_step:
      after = (next_step - current).total_seconds()
    return before, during, after


class RandomizedArrivalDepartureOccupancy:
  """randomized_arrival_departure_occupancy.py:36-238, occupant by occupant (the
  product's generator, sbsim_b200/exogenous.py, batches the draws of one call)."""

  def __init__(self, zone_assignment, earliest_expected_arrival_hour,
               latest_expected_arrival_hour, earliest_expected_departure_hour,
               latest_expected_departure_hour, time_step_sec, seed=17321, time_zone="UTC"):
    self._n = zone_assignment
    self._arr = (earliest_expected_arrival_hour, latest_expected_arrival_hour)
    self._dep = (earliest_expected_departure_hour, latest_expected_departure_hour)
    step = pd.Timedelta(time_step_sec, unit="second")
    self._p_arr = 1.0 / (pd.Timedelta(self._arr[1] - self._arr[0], unit="hour") / step / 2.0)   # :93-104
    self._p_dep = 1.0 / (pd.Timedelta(self._dep[1] - self._dep[0], unit="hour") / step / 2.0)
    self._rs = np.random.RandomState(seed)
    self._tz = time_zone
    self._work = {}              # zone_id -> list of bool (occupant is at WORK)

  def _local(self, ts):                                        # :86-91
    if ts.tz is None:
      return ts
    import zoneinfo
    tz = {"US/Pacific": "America/Los_Angeles"}.get(self._tz, self._tz)
    return ts.tz_convert(zoneinfo.ZoneInfo(tz) if isinstance(tz, str) and tz != "UTC" else tz)

  def average_zone_occupancy(self, zone_id, start_time, end_time) -> float:  # :198-238
    if zone_id not in self._work:
      self._work[zone_id] = [False] * self._n
    states = self._work[zone_id]
    local = self._local(start_time)
    day = pd.Timestamp(year=local.year, month=local.month, day=local.day)
    count = 0.0
    for i in range(self._n):                                   # ZoneOccupant.peek :127-147
      if not is_work_day(day):
        states[i] = False
      elif not states[i]:
        if not (local.hour < self._arr[0] or local.hour > self._arr[1]):     # :106-117
          if self._rs.rand() < self._p_arr:
            states[i] = True
      else:
        if not local.hour < self._dep[0]:                      # :119-125
          if self._rs.rand() < self._p_dep:
            states[i] = False
      if states[i]:
        count += 1.0
    return count


class ConstantOccupancy:
  """Trivial occupancy used by tz-aware scenarios (a BaseOccupancy, models/base_occupancy.py:27-46)."""

  def __init__(self, value: float):
    self.value = value

  def average_zone_occupancy(self, zone_id, start_time, end_time) -> float:
    return self.value


# ----------------------------------------------------------------------------
# energy cost
# ----------------------------------------------------------------------------

# electricity_energy_cost.py:40-123
CARBON_EMISSION_BY_HOUR = (
    88.19666493, 87.79190866, 87.87607686, 87.83054163, 88.00279618,
    88.19648183, 89.70663283, 93.97947901, 98.85868291, 100.7853521,
    101.3866866, 101.7795612, 102.5919168, 103.4403736, 104.1380294,
    104.7359292, 102.0714466, 97.04226176, 93.57895651, 92.46355045,
    91.72914657, 90.69209747, 89.76552213, 88.99950995,
)
WEEKDAY_PRICE_BY_HOUR = (16.0,) * 6 + (18.0,) * 6 + (20.0,) * 7 + (16.0,) * 5
WEEKEND_PRICE_BY_HOUR = (16.0,) * 24
# natural_gas_energy_cost.py:30-43
GAS_PRICE_BY_MONTH_SOURCE = (9.02, 8.35, 7.77, 7.26, 6.69, 6.86, 6.77, 6.76,
                             6.99, 7.19, 7.96, 8.98)
KWH_PER_KFT3_GAS = 293.07107   # utils/constants.py:39
JOULES_PER_KWH = 3.6e6         # utils/constants.py:26
GAS_CO2 = 53.12                # utils/constants.py:43


class ElectricityEnergyCost:
  """electricity_energy_cost.py:127-224 (pint units are tags only)."""

  def __init__(self, weekday_energy_prices=WEEKDAY_PRICE_BY_HOUR,
               weekend_energy_prices=WEEKEND_PRICE_BY_HOUR,
               carbon_emission_rates=CARBON_EMISSION_BY_HOUR):
    if (len(weekday_energy_prices) != 24 or len(weekend_energy_prices) != 24
        or len(carbon_emission_rates) != 24):
      raise ValueError("rates must have 24 entries.")
    self._carbon = np.array(carbon_emission_rates) / 1.0e6 / 3600.0      # :145-147
    self._weekday = np.array(weekday_energy_prices) / 100.0 / 1000.0 / 3600.0
    self._weekend = np.array(weekend_energy_prices) / 100.0 / 1000.0 / 3600.0

  def cost(self, start_time, end_time, energy_rate) -> float:  # :166-194
    dt = (end_time - start_time).total_seconds()
    hour = start_time.hour
    price = self._weekday[hour] if is_work_day(start_time) else self._weekend[hour]
    return price * np.abs(energy_rate) * dt

  def carbon(self, start_time, end_time, energy_rate) -> float:  # :196-224
    dt = (end_time - start_time).total_seconds()
    return self._carbon[start_time.hour] * np.abs(energy_rate) * dt


class NaturalGasEnergyCost:
  """natural_gas_energy_cost.py:48-138."""

  def __init__(self, gas_price_by_month=GAS_PRICE_BY_MONTH_SOURCE):
    assert len(gas_price_by_month) == 12
    self._month_gas_price = (np.array(gas_price_by_month) / KWH_PER_KFT3_GAS
                             / JOULES_PER_KWH)
    self._carbon_rate = GAS_CO2 / KWH_PER_KFT3_GAS / JOULES_PER_KWH

  def cost(self, start_time, end_time, energy_rate) -> float:  # :75-110
    if energy_rate < 0.0:
      energy_rate = 0.0
    dt = (end_time - start_time).total_seconds()
    return self._month_gas_price[start_time.month - 1] * (energy_rate * dt)

  def carbon(self, start_time, end_time, energy_rate) -> float:  # :112-138
    if energy_rate < 0.0:
      energy_rate = 0.0
    dt = (end_time - start_time).total_seconds()
    return self._carbon_rate * (energy_rate * dt)
