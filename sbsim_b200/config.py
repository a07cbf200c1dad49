"""Configuration objects with the reference's class names and constructor arguments.

In the reference these classes ARE the simulation (they hold state and do the
arithmetic per building, per step, in Python).  Here they are plain parameter
holders: the arithmetic runs in libsbx for the whole batch.  Argument names,
defaults and validation errors follow the reference so that a gin / Python
config written for sbsim reads the same (paths relative to
/root/reference/smart_control/):

  AirHandler                          simulator/air_handler.py:51-135
  Boiler                              simulator/boiler.py:53-123
  FloorPlanBasedHvac                  simulator/hvac_floorplan_based.py:36-128
  SetpointEnergyCarbonRegretFunction  reward/setpoint_energy_carbon_regret.py:101-140
  BoundedActionNormalizer             utils/bounded_action_normalizer.py:28-126
  ActionConfig                        environment/environment.py:277-307
  StandardScoreObservationNormalizer  utils/observation_normalizer.py:31-66
  HistogramReducer                    utils/histogram_reducer.py:206-246
"""

from __future__ import annotations

import dataclasses
from typing import Dict, Mapping, Optional, Sequence, Tuple

import numpy as np

from sbsim_b200 import exogenous
from sbsim_b200 import specs

ACTION_TOLERANCE = 1e-5  # utils/bounded_action_normalizer.py:25


@dataclasses.dataclass
class AirHandler:
  recirculation: float
  heating_air_temp_setpoint: float
  cooling_air_temp_setpoint: float
  fan_differential_pressure: float
  fan_efficiency: float
  max_air_flow_rate: float = 8.67
  device_id: Optional[str] = None
  sim_weather_controller: Optional[object] = None

  def __post_init__(self):
    if self.cooling_air_temp_setpoint <= self.heating_air_temp_setpoint:
      raise ValueError("cooling_air_temp_setpoint must greater than"
                       " heating_air_temp_setpoint")


@dataclasses.dataclass
class Boiler:
  reheat_water_setpoint: float
  water_pump_differential_head: float
  water_pump_efficiency: float
  device_id: Optional[str] = None
  heating_rate: float = 0
  cooling_rate: float = 0
  convection_coefficient: float = 5.6
  tank_length: float = 2.0
  tank_radius: float = 0.5
  water_capacity: float = 1.5
  insulation_conductivity: float = 0.067
  insulation_thickness: float = 0.06


@dataclasses.dataclass
class FloorPlanBasedHvac:
  air_handler: AirHandler
  boiler: Boiler
  schedule: exogenous.SetpointSchedule
  vav_max_air_flow_rate: float
  vav_reheat_max_water_flow_rate: float


@dataclasses.dataclass
class Hvac:
  """The deprecated `Hvac` (simulator/hvac.py:35-123): one boiler, one air handler and one VAV
  per zone of the rectangular `Building`, zones addressed by (row, column).  Same devices and
  algebra as FloorPlanBasedHvac; what differs is the naming: device `vav_<i>_<j>` and zone id
  `zone_id_(i, j)` (conversion_utils.zone_coordinates_to_id).  Use with `legacy_building()`."""
  zone_coordinates: Sequence[Tuple[int, int]]
  air_handler: AirHandler
  boiler: Boiler
  schedule: exogenous.SetpointSchedule
  vav_max_air_flow_rate: float
  vav_reheat_max_water_flow_rate: float

  def __post_init__(self):
    self.zone_coordinates = [(int(i), int(j)) for i, j in self.zone_coordinates]
    if len(set(self.zone_coordinates)) != len(self.zone_coordinates):
      raise ValueError("zone_coordinates must be distinct")


@dataclasses.dataclass
class SetpointEnergyCarbonRegretFunction:
  max_productivity_personhour_usd: float
  min_productivity_personhour_usd: float
  max_electricity_rate: float
  max_natural_gas_rate: float
  productivity_midpoint_delta: float
  productivity_decay_stiffness: float
  electricity_energy_cost: exogenous.ElectricityEnergyCost
  natural_gas_energy_cost: exogenous.NaturalGasEnergyCost
  productivity_weight: float
  energy_cost_weight: float
  carbon_emission_weight: float

  def __post_init__(self):
    assert self.max_productivity_personhour_usd > self.min_productivity_personhour_usd


@dataclasses.dataclass
class SetpointEnergyCarbonRewardFunction:
  """reward/setpoint_energy_carbon_reward.py:103-125: productivity minus weighted energy
  cost and carbon cost, shifted and scaled (no regret normalisation, no rate caps)."""
  max_productivity_personhour_usd: float
  productivity_midpoint_delta: float
  productivity_decay_stiffness: float
  electricity_energy_cost: exogenous.ElectricityEnergyCost
  natural_gas_energy_cost: exogenous.NaturalGasEnergyCost
  energy_cost_weight: float
  carbon_cost_weight: float
  carbon_cost_factor: float
  reward_normalizer_shift: float = 0.0
  reward_normalizer_scale: float = 1.0


class BoundedActionNormalizer:
  """Maps the agent's [-1, 1] onto [min_native_value, max_native_value]."""

  def __init__(self, min_native_value: float, max_native_value: float,
               min_normalized_value: float = -1.0, max_normalized_value: float = 1.0):
    if (min_normalized_value, max_normalized_value) != (-1.0, 1.0):
      raise NotImplementedError("libsbx maps the reference's action_spec range [-1, 1]"
                                " (environment.py:647-653)")
    self._min_native_value = min_native_value
    self._max_native_value = max_native_value
    self._min_normalized_value = min_normalized_value
    self._max_normalized_value = max_normalized_value
    self._tolerance = ACTION_TOLERANCE

  def get_array_spec(self, name=None):
    return specs.BoundedArraySpec((), np.float32, minimum=self._min_normalized_value,
                                  maximum=self._max_normalized_value, name=name)

  def setpoint_value(self, agent_action) -> float:
    """Host restatement (used for validation and default actions only; the batch
    is mapped on the GPU)."""
    if np.ndim(agent_action) > 0:
      raise ValueError(f"agent_action expected to be scalar but received: {agent_action}")
    if (agent_action < self._min_normalized_value - self._tolerance
        or agent_action > self._max_normalized_value + self._tolerance):
      raise ValueError(f"agent_action: {agent_action} not within bounds"
                       f" [{self._min_normalized_value}, {self._max_normalized_value}]")
    ratio = (agent_action - self._min_normalized_value) / (
        self._max_normalized_value - self._min_normalized_value)
    return ratio * (self._max_native_value - self._min_native_value) + self._min_native_value

  def agent_value(self, setpoint_value: float) -> float:
    if setpoint_value > self._max_native_value or setpoint_value < self._min_native_value:
      raise ValueError(f"setpoint_value {setpoint_value} not within bounds"
                       f" [{self._min_native_value}, {self._max_native_value}]")
    return ((self._max_normalized_value - self._min_normalized_value)
            / (self._max_native_value - self._min_native_value)
            * (setpoint_value - self._min_native_value) + self._min_normalized_value)

  @property
  def setpoint_min(self) -> float:
    return self._min_native_value

  @property
  def setpoint_max(self) -> float:
    return self._max_native_value


class ActionConfig:
  """setpoint name -> action normalizer; only listed setpoints become actions."""

  def __init__(self, action_normalizers: Mapping[str, BoundedActionNormalizer]):
    self._action_normalizers = dict(action_normalizers)

  def get_action_normalizer(self, setpoint_name: str):
    return self._action_normalizers.get(setpoint_name)


class StandardScoreObservationNormalizer:
  """(value - sample_mean) / sqrt(sample_variance) by exact measurement name;
  unknown names pass through, non-positive variance gives 0.

  `normalization_constants` maps name -> (sample_mean, sample_variance) or any
  object with those two attributes (the reference passes ContinuousVariableInfo
  protos whose fields are `float`, hence the fp32 rounding applied here)."""

  def __init__(self, normalization_constants: Mapping[str, object]):
    self._constants: Dict[str, Tuple[float, float]] = {}
    for k, v in normalization_constants.items():
      if hasattr(v, "sample_mean"):
        mean, var = v.sample_mean, v.sample_variance
      else:
        mean, var = v
      self._constants[k] = (float(np.float32(mean)), float(np.float32(var)))

  def get(self, field_name: str) -> Tuple[float, float]:
    return self._constants.get(field_name, (0.0, 1.0))


class HistogramReducer:
  """Bins the per-VAV measurements into normalised histograms.

  Only the binning parameters matter on the step path (histogram_reducer.py:
  411-433); the reference's `reader` argument is used for its `expand()` device
  assignment, which the environment never calls."""

  def __init__(self, histogram_parameters_tuples: Sequence[Tuple[str, Sequence[float]]],
               reader=None, normalize_reduce: bool = True):
    if not normalize_reduce:
      raise NotImplementedError("the calibrated config uses normalize_reduce=True"
                                " (SAC_Demo.ipynb cell 4)")
    self.histogram_parameters = {p[0]: np.array(p[1], dtype=np.float64)
                                 for p in histogram_parameters_tuples}
