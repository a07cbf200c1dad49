"""Host-side exogenous inputs: weather, schedule, occupancy, energy prices.

The reference evaluates these per call, per building, in Python
(`weather_controller.get_current_temp(ts)` four times per step, schedules with
pandas tz conversions, ...).  None of them depends on building state, so here
they are evaluated ONCE per episode into per-step tables that the kernels index
by the simulation step.  Classes keep the reference's
names and constructor arguments; each gains a `table(timestamps)` method.

Reference (paths relative to /root/reference/smart_control/):
  WeatherController / ReplayWeatherController  simulator/weather_controller.py:47-218
  SetpointSchedule                             simulator/setpoint_schedule.py:29-128
  StepFunctionOccupancy                        simulator/step_function_occupancy.py:36-173
  ElectricityEnergyCost                        reward/electricity_energy_cost.py:29-224
  NaturalGasEnergyCost                         reward/natural_gas_energy_cost.py:28-138
  is_work_day / get_radian_time                utils/conversion_utils.py:62-134
  expand_time_features                         utils/regression_building_utils.py:97-125
"""

from __future__ import annotations

import datetime
import functools
import math
import zoneinfo
from typing import Mapping, Optional, Sequence, Set, Tuple

import numpy as np
import pandas as pd

_EPOCH = pd.Timestamp("1970-01-01", tz="UTC")
_TZ_ALIAS = {"US/Pacific": "America/Los_Angeles", "US/Eastern": "America/New_York",
             "US/Central": "America/Chicago", "US/Mountain": "America/Denver"}


def resolve_timezone(tz):
  """Accepts tzinfo objects, 'UTC' or IANA / legacy 'US/...' names."""
  if tz is None:
    return datetime.timezone.utc
  if isinstance(tz, datetime.tzinfo):
    return tz
  if tz in ("UTC", "utc"):
    return datetime.timezone.utc
  return zoneinfo.ZoneInfo(_TZ_ALIAS.get(tz, tz))


def step_timestamps(start: pd.Timestamp, time_step_sec: float, n: int):
  """[start + s * dt for s in range(n)] (simulator.py:455 increments by Timedelta)."""
  dt = pd.Timedelta(time_step_sec, unit="s")
  return [start + s * dt for s in range(n)]


# ----------------------------------------------------------------------------
# weather
# ----------------------------------------------------------------------------


class WeatherController:
  """Sinusoid weather, low at midnight and high at noon (weather_controller.py:47-132)."""

  def __init__(self, default_low_temp: float, default_high_temp: float,
               special_days: Optional[Mapping[int, Tuple[float, float]]] = None,
               convection_coefficient: float = 12.0):
    self.default_low_temp = default_low_temp
    self.default_high_temp = default_high_temp
    self.special_days = special_days if special_days else {}
    self.convection_coefficient = convection_coefficient
    if self.default_low_temp > self.default_high_temp:
      raise ValueError("default_low_temp cannot be greater than default_high_temp.")
    for day, (low, high) in self.special_days.items():
      if low > high:
        raise ValueError(
            f"Low temp cannot be greater than high temp for special day: {day}.")

  @staticmethod
  def seconds_to_rads(seconds_in_day: float) -> float:
    lo, hi = -math.pi / 2.0, 3.0 * math.pi / 2.0
    return (seconds_in_day / 86400) * (hi - lo) + lo

  def get_current_temp(self, timestamp: pd.Timestamp) -> float:
    today = timestamp.dayofyear
    tomorrow = (today + 1) % 365
    today_low, today_high = self.special_days.get(
        today, (self.default_low_temp, self.default_high_temp))
    tomorrow_low = self.special_days.get(tomorrow, (self.default_low_temp, None))[0]
    low = today_low if timestamp.hour < 12 else tomorrow_low
    seconds = (timestamp - pd.Timestamp(timestamp.date())).total_seconds()
    rad = self.seconds_to_rads(seconds)
    return 0.5 * (math.sin(rad) + 1) * (today_high - low) + low

  def get_air_convection_coefficient(self, timestamp=None) -> float:
    return self.convection_coefficient

  def table(self, timestamps: Sequence[pd.Timestamp]) -> np.ndarray:
    return np.array([self.get_current_temp(t) for t in timestamps], dtype=np.float64)


def sinusoid_weather_tables(timestamps: Sequence[pd.Timestamp], low: np.ndarray,
                            high: np.ndarray) -> np.ndarray:
  """Per-building sinusoid series [B, T]: same expression as
  WeatherController.get_current_temp, evaluated with the shared sin(rad) of each
  timestamp and broadcast over buildings' (low, high)."""
  s = np.array([math.sin(WeatherController.seconds_to_rads(
      (t - pd.Timestamp(t.date())).total_seconds())) for t in timestamps])
  low = np.asarray(low, dtype=np.float64)[:, None]
  high = np.asarray(high, dtype=np.float64)[:, None]
  return 0.5 * (s[None, :] + 1) * (high - low) + low


class ReplayWeatherController:
  """Linear interpolation of an hourly CSV with columns Time, TempF
  (weather_controller.py:164-218)."""

  def __init__(self, local_weather_path: Optional[str] = None,
               convection_coefficient: float = 12.0, *, times_utc_sec=None, temps_f=None):
    if local_weather_path is not None:
      df = pd.read_csv(local_weather_path)
      self._times = np.array(
          [(pd.Timestamp(t, tz="UTC") - _EPOCH).total_seconds() for t in df["Time"]])
      self._temps_f = np.asarray(df["TempF"], dtype=np.float64)
    else:
      self._times = np.asarray(times_utc_sec, dtype=np.float64)
      self._temps_f = np.asarray(temps_f, dtype=np.float64)
    self.convection_coefficient = convection_coefficient

  def get_current_temp(self, timestamp: pd.Timestamp) -> float:
    timestamp = timestamp.tz_convert("UTC")
    target = (timestamp - _EPOCH).total_seconds()
    if target < self._times.min():
      raise ValueError(f"Attempting to get weather data at {timestamp}, before the"
                       " latest timestamp.")
    if target > self._times.max():
      raise ValueError(f"Attempting to get weather data at {timestamp}, after the"
                       " latest timestamp.")
    temp_f = np.interp(target, self._times, self._temps_f)
    return (temp_f - 32.0) * 5.0 / 9.0 + 273.15      # conversion_utils.py:155-170

  def get_air_convection_coefficient(self, timestamp=None) -> float:
    return self.convection_coefficient

  def table(self, timestamps: Sequence[pd.Timestamp]) -> np.ndarray:
    return np.array([self.get_current_temp(t) for t in timestamps], dtype=np.float64)


# ----------------------------------------------------------------------------
# schedule
# ----------------------------------------------------------------------------


class SetpointSchedule:
  """Comfort / eco temperature windows (setpoint_schedule.py:29-128)."""

  def __init__(self, morning_start_hour: int, evening_start_hour: int,
               comfort_temp_window: Tuple[float, float],
               eco_temp_window: Tuple[float, float],
               holidays: Optional[Set[int]] = None, time_zone="UTC"):
    if morning_start_hour > evening_start_hour:
      raise ValueError("morning_start_hour must be less than evening_start_hour")
    if comfort_temp_window[0] > comfort_temp_window[1]:
      raise ValueError("comfort_temp_window[0] must be less than comfort_temp_window[1]")
    if eco_temp_window[0] > eco_temp_window[1]:
      raise ValueError("eco_temp_window[0] must be less than eco_temp_window[1]")
    self.morning_start_hour = morning_start_hour
    self.evening_start_hour = evening_start_hour
    self.comfort_temp_window = comfort_temp_window
    self.eco_temp_window = eco_temp_window
    self.holidays = set(holidays) if holidays else set()
    self._time_zone = resolve_timezone(time_zone)

  def _local(self, ts: pd.Timestamp) -> pd.Timestamp:
    if ts.tz is not None:
      return ts.tz_convert(self._time_zone)
    return ts.tz_localize(datetime.timezone.utc)

  def is_weekend(self, ts: pd.Timestamp) -> bool:
    return self._local(ts).day_name() in ("Saturday", "Sunday")

  def is_comfort_mode(self, ts: pd.Timestamp) -> bool:
    lt = self._local(ts)
    return (lt.hour >= self.morning_start_hour and lt.hour < self.evening_start_hour
            and lt.dayofyear not in self.holidays and not self.is_weekend(lt))

  def get_temperature_window(self, ts: pd.Timestamp):
    return self.comfort_temp_window if self.is_comfort_mode(ts) else self.eco_temp_window

  def table(self, timestamps: Sequence[pd.Timestamp]) -> np.ndarray:
    return np.array([self.is_comfort_mode(t) for t in timestamps], dtype=np.uint8)


# ----------------------------------------------------------------------------
# calendar helpers
# ----------------------------------------------------------------------------


@functools.cache
def _us_holidays_package():
  """`holidays.US()` as the reference uses it (conversion_utils.py:62-70; the object fills in
  a year's holidays on the first lookup of a date of that year), or None when the package is
  not installed -- it is not in this image."""
  try:
    import holidays      # pylint: disable=g-import-not-at-top
    return holidays.US()
  except ImportError:
    return None


@functools.cache
def _us_federal_holidays_of_year(year: int):
  """Without the `holidays` package: the federal holidays of pandas' USFederalHolidayCalendar,
  which is what plain `holidays.US()` contains (observed-day rules may differ between versions
  of either package: the deviation is confined to this lookup)."""
  from pandas.tseries.holiday import USFederalHolidayCalendar      # pylint: disable=g-import-not-at-top
  cal = USFederalHolidayCalendar()
  return frozenset(d.date() for d in cal.holidays(start=f"{year}-01-01", end=f"{year}-12-31"))


def is_work_day(ts: pd.Timestamp) -> bool:
  if ts.weekday() >= 5:
    return False
  cal = _us_holidays_package()
  if cal is not None:
    return ts.date() not in cal
  return ts.date() not in _us_federal_holidays_of_year(ts.year)


def time_feature_table(timestamps: Sequence[pd.Timestamp]) -> np.ndarray:
  """[T, 4] = hod cos, hod sin, dow cos, dow sin with one feature pair each
  (environment.py:916-940, conversion_utils.py:107-134)."""
  out = np.zeros((len(timestamps), 4), dtype=np.float64)
  for i, ts in enumerate(timestamps):
    day = pd.Timestamp(year=ts.year, month=ts.month, day=ts.day, tz=ts.tz)
    hod = 2.0 * np.pi * ((ts - day).total_seconds() / 86400.0)
    dow = 2.0 * np.pi * (float(day.weekday()) / 7.0)
    # expand_time_features with n == 1: phase = rad + 0
    out[i] = (np.cos(hod + 0.0), np.sin(hod + 0.0), np.cos(dow + 0.0), np.sin(dow + 0.0))
  return out


# ----------------------------------------------------------------------------
# occupancy
# ----------------------------------------------------------------------------


class StepFunctionOccupancy:
  """Constant occupancy inside / outside working hours (step_function_occupancy.py:36-173)."""

  def __init__(self, work_start_time: pd.Timedelta, work_end_time: pd.Timedelta,
               work_occupancy: float, nonwork_occupancy: float):
    for t in (work_start_time, work_end_time):
      if t > pd.Timedelta(24, unit="hour") or t.total_seconds() < 0.0:
        raise ValueError("Time delta must be positive and less than one day.")
    self._work_start_time = work_start_time
    self._work_end_time = work_end_time
    self._work_occupancy = work_occupancy
    self._nonwork_occupancy = nonwork_occupancy

  def average_zone_occupancy(self, zone_id: str, start_time: pd.Timestamp,
                             end_time: pd.Timestamp) -> float:
    if start_time >= end_time:
      raise ValueError("End time may not occur before start time.")
    day = pd.Timestamp(year=start_time.year, month=start_time.month, day=start_time.day)
    current = start_time - day          # raises for tz-aware input, as the reference does
    work = nonwork = 0.0
    while day <= end_time:
      day_end = min(pd.Timedelta(1, unit="day"), end_time - day)
      if is_work_day(day):
        b, d, a = self._split_workday(current, day_end)
        work += d
        nonwork += b + a
      else:
        nonwork += (day_end - current).total_seconds()
      day += pd.Timedelta(1.0, unit="day")
      current = pd.Timedelta(0.0, unit="sec")
    return (work * self._work_occupancy + nonwork * self._nonwork_occupancy) / (work + nonwork)

  def _split_workday(self, start: pd.Timedelta, end: pd.Timedelta):
    before = during = after = 0.0
    current = start
    interval_end = min(end, pd.Timedelta(24, unit="hour"))
    nxt = min(interval_end, self._work_start_time)
    if current < nxt:
      before = (nxt - current).total_seconds()
      current = max(current, nxt)
    nxt = min(interval_end, self._work_end_time)
    if current < nxt:
      during = (nxt - current).total_seconds()
      current = nxt
    if current < interval_end:
      after = (interval_end - current).total_seconds()
    return before, during, after


class ConstantOccupancy:
  """A BaseOccupancy (models/base_occupancy.py:27-46) that is constant in time."""

  def __init__(self, value: float):
    self.value = value

  def average_zone_occupancy(self, zone_id, start_time, end_time) -> float:
    return self.value


class RandomizedArrivalDepartureOccupancy:
  """The shipped occupancy model (randomized_arrival_departure_occupancy.py:36-238;
  sim_config.gin:185-196): `zone_assignment` occupants per zone arrive between
  the earliest and latest arrival hour and leave after the earliest departure
  hour, one Bernoulli draw per occupant per call from ONE shared
  numpy RandomState(seed).

  The model is stateful and advances its generator on every call, so the
  per-step tables must be produced in the reference's call order (see
  `occupancy_tables`): the observation's `num_occupants` at reset, then per step
  the reward's occupancy of every zone followed by the observation's.  The
  draws of all occupants of one (zone, time) call are taken with one
  `rand(k)`, which yields the same stream as k successive `rand()` calls.
  """

  stateful = True
  per_zone = True

  def __init__(self, zone_assignment: int, earliest_expected_arrival_hour: int,
               latest_expected_arrival_hour: int, earliest_expected_departure_hour: int,
               latest_expected_departure_hour: int, time_step_sec: int,
               seed: Optional[int] = 17321, time_zone="UTC"):
    assert (earliest_expected_arrival_hour < latest_expected_arrival_hour
            < earliest_expected_departure_hour < latest_expected_departure_hour)
    self._zone_assignment = int(zone_assignment)
    self._arrival = (earliest_expected_arrival_hour, latest_expected_arrival_hour)
    self._departure = (earliest_expected_departure_hour, latest_expected_departure_hour)
    step = pd.Timedelta(time_step_sec, unit="second")
    # ZoneOccupant._get_event_probability :93-104: p = 1 / (window / step / 2)
    self._p_arrival = 1.0 / (pd.Timedelta(self._arrival[1] - self._arrival[0], unit="hour")
                             / step / 2.0)
    self._p_departure = 1.0 / (pd.Timedelta(self._departure[1] - self._departure[0],
                                            unit="hour") / step / 2.0)
    self._random_state = np.random.RandomState(seed)
    self._time_zone = resolve_timezone(time_zone)
    self._at_work = {}           # zone_id -> bool[zone_assignment]

  def average_zone_occupancy(self, zone_id: str, start_time: pd.Timestamp,
                             end_time: pd.Timestamp) -> float:
    """:198-238 with ZoneOccupant.peek :127-147 inlined over the zone's occupants."""
    state = self._at_work.get(zone_id)
    if state is None:
      state = self._at_work[zone_id] = np.zeros(self._zone_assignment, dtype=bool)
    local = start_time if start_time.tz is None else start_time.tz_convert(self._time_zone)
    day = pd.Timestamp(year=local.year, month=local.month, day=local.day)
    hour = local.hour
    if not is_work_day(day):
      state[:] = False
      return 0.0
    # AWAY occupants draw only inside [earliest, latest] arrival hours (:106-117);
    # WORK occupants draw once the earliest departure hour is reached (:119-125)
    arrive = (~state) if self._arrival[0] <= hour <= self._arrival[1] else np.zeros_like(state)
    depart = state if hour >= self._departure[0] else np.zeros_like(state)
    draws = arrive | depart
    k = int(draws.sum())
    if k:
      r = self._random_state.rand(k)
      p = np.where(arrive[draws], self._p_arrival, self._p_departure)
      flip = np.zeros_like(state)
      flip[draws] = r < p
      state ^= flip
    return float(state.sum())


class TableOccupancy:
  """Replays a host-generated [T, Z] table (e.g. produced by the reference's
  RandomizedArrivalDepartureOccupancy)."""

  def __init__(self, reward_table: np.ndarray, obs_table: np.ndarray):
    self.reward_table = np.asarray(reward_table, dtype=np.float64)
    self.obs_table = np.asarray(obs_table, dtype=np.int32)


def occupancy_tables(occupancy, zone_ids: Sequence[str], timestamps, time_step_sec: float,
                     per_zone: bool):
  """Returns (occ_reward [T, Zo], occ_obs [T], occ_obs_zone [T, Zo] or None).

  occ_reward[s, z] = average_zone_occupancy(zone z, t_s, t_s + dt)
      (simulator_flexible_floor_plan.py:206-210, evaluated at the post-step time)
  occ_obs[s] = int(sum_z average_zone_occupancy(z, t_s - 5 min, t_s))
      (simulator_building.py:305-315), summed over `zone_ids`
  occ_obs_zone[s, z] = the terms of that sum: the device adds up the zones each
      building actually has (batches of buildings with different zone counts)
  """
  if isinstance(occupancy, TableOccupancy):
    return occupancy.reward_table, occupancy.obs_table, None
  dt = pd.Timedelta(time_step_sec, unit="s")
  five = pd.Timedelta(5, unit="minute")
  zo = len(zone_ids) if per_zone else 1
  rew = np.zeros((len(timestamps), zo), dtype=np.float64)
  obs = np.zeros(len(timestamps), dtype=np.int32)
  obs_zone = np.zeros((len(timestamps), zo), dtype=np.float64)
  if getattr(occupancy, "stateful", False):
    # Call order of the reference for a model that advances a generator on every
    # call: _reset -> _get_observation -> num_occupants(t_0)
    # (environment.py:1165-1212, simulator_building.py:305-315); then each _step
    # (environment.py:1298-1306): _get_observation -> num_occupants(t_s) for every
    # zone, then _get_reward -> reward_info -> occupancy(z, t_s, t_s + dt) for every
    # zone (simulator_flexible_floor_plan.py:192-230).  Row 0 of the reward table
    # is never read, so it draws nothing.
    if not per_zone:
      raise ValueError("a stateful occupancy model needs per-zone tables")
    for s, ts in enumerate(timestamps):
      n = 0.0
      for z, zid in enumerate(zone_ids):
        obs_zone[s, z] = occupancy.average_zone_occupancy(zid, ts - five, ts)
        n += obs_zone[s, z]
      obs[s] = int(n)
      if s > 0:
        for z, zid in enumerate(zone_ids):
          rew[s, z] = occupancy.average_zone_occupancy(zid, ts, ts + dt)
    return rew, obs, obs_zone
  for s, ts in enumerate(timestamps):
    if per_zone:
      for z, zid in enumerate(zone_ids):
        rew[s, z] = occupancy.average_zone_occupancy(zid, ts, ts + dt)
        obs_zone[s, z] = occupancy.average_zone_occupancy(zid, ts - five, ts)
      n = 0.0
      for z in range(len(zone_ids)):
        n += obs_zone[s, z]
    else:
      rew[s, 0] = occupancy.average_zone_occupancy(zone_ids[0], ts, ts + dt)
      one = occupancy.average_zone_occupancy(zone_ids[0], ts - five, ts)
      obs_zone[s, 0] = one
      n = 0.0
      for _ in zone_ids:
        n += one
    obs[s] = int(n)
  return rew, obs, obs_zone


# ----------------------------------------------------------------------------
# energy cost
# ----------------------------------------------------------------------------

CARBON_EMISSION_BY_HOUR = (
    88.19666493, 87.79190866, 87.87607686, 87.83054163, 88.00279618,
    88.19648183, 89.70663283, 93.97947901, 98.85868291, 100.7853521,
    101.3866866, 101.7795612, 102.5919168, 103.4403736, 104.1380294,
    104.7359292, 102.0714466, 97.04226176, 93.57895651, 92.46355045,
    91.72914657, 90.69209747, 89.76552213, 88.99950995,
)
WEEKDAY_PRICE_BY_HOUR = (16.0,) * 6 + (18.0,) * 6 + (20.0,) * 7 + (16.0,) * 5
WEEKEND_PRICE_BY_HOUR = (16.0,) * 24
GAS_PRICE_BY_MONTH_SOURCE = (9.02, 8.35, 7.77, 7.26, 6.69, 6.86, 6.77, 6.76, 6.99, 7.19,
                             7.96, 8.98)
KWH_PER_KFT3_GAS = 293.07107
JOULES_PER_KWH = 3.6e6
GAS_CO2 = 53.12


class ElectricityEnergyCost:
  """Time-of-use price and hourly carbon intensity (electricity_energy_cost.py:127-224)."""

  def __init__(self, weekday_energy_prices=WEEKDAY_PRICE_BY_HOUR,
               weekend_energy_prices=WEEKEND_PRICE_BY_HOUR,
               carbon_emission_rates=CARBON_EMISSION_BY_HOUR):
    if len(weekday_energy_prices) != 24 or len(weekend_energy_prices) != 24:
      raise ValueError("Energy cost rates must have 24 entries.")
    if len(carbon_emission_rates) != 24:
      raise ValueError("Carbon emission rates must have 24 entries.")
    self._carbon = np.array(carbon_emission_rates) / 1.0e6 / 3600.0
    self._weekday = np.array(weekday_energy_prices) / 100.0 / 1000.0 / 3600.0
    self._weekend = np.array(weekend_energy_prices) / 100.0 / 1000.0 / 3600.0

  def price(self, start_time: pd.Timestamp) -> float:
    table = self._weekday if is_work_day(start_time) else self._weekend
    return float(table[start_time.hour])

  def carbon_rate(self, start_time: pd.Timestamp) -> float:
    return float(self._carbon[start_time.hour])

  def cost(self, start_time, end_time, energy_rate: float) -> float:
    return self.price(start_time) * np.abs(energy_rate) * (end_time - start_time).total_seconds()

  def carbon(self, start_time, end_time, energy_rate: float) -> float:
    return self.carbon_rate(start_time) * np.abs(energy_rate) * (
        end_time - start_time).total_seconds()


class NaturalGasEnergyCost:
  """Monthly gas price and constant carbon factor (natural_gas_energy_cost.py:48-138)."""

  def __init__(self, gas_price_by_month=GAS_PRICE_BY_MONTH_SOURCE):
    assert len(gas_price_by_month) == 12, "Gas price per month must have exactly 12 values."
    self._month_gas_price = np.array(gas_price_by_month) / KWH_PER_KFT3_GAS / JOULES_PER_KWH
    self.carbon_rate = GAS_CO2 / KWH_PER_KFT3_GAS / JOULES_PER_KWH

  def price(self, start_time: pd.Timestamp) -> float:
    return float(self._month_gas_price[start_time.month - 1])

  def cost(self, start_time, end_time, energy_rate: float) -> float:
    energy_rate = max(energy_rate, 0.0)
    return self.price(start_time) * (energy_rate * (end_time - start_time).total_seconds())

  def carbon(self, start_time, end_time, energy_rate: float) -> float:
    energy_rate = max(energy_rate, 0.0)
    return self.carbon_rate * (energy_rate * (end_time - start_time).total_seconds())


def energy_tables(electricity: ElectricityEnergyCost, gas: NaturalGasEnergyCost, timestamps):
  """Price / carbon tables at each step's RewardInfo.start_timestamp.  The reward
  function sees proto timestamps, i.e. UTC (conversion_utils.py:52-59)."""
  pe = np.zeros(len(timestamps))
  ce = np.zeros(len(timestamps))
  pg = np.zeros(len(timestamps))
  for s, ts in enumerate(timestamps):
    utc = pd.Timestamp(int(ts.timestamp()), unit="s", tz="UTC")
    pe[s] = electricity.price(utc)
    ce[s] = electricity.carbon_rate(utc)
    pg[s] = gas.price(utc)
  return pe, ce, pg
