"""sbsim_b200: B200-native batched building-thermodynamics RL environment.

A drop-in for the per-timestep hot path of google/sbsim (Environment.step():
control-volume diffusion, HVAC update, observation pack, regret reward), batched
over independent buildings and executed by hand-written sm_100a CUDA kernels
behind the C ABI of include/sbx.h.  See DESIGN.md and INTEGRATION.md.
"""

from sbsim_b200._lib import (LIB_PATH, PATH_AUTO, PATH_RESIDENT, PATH_STREAMING,
                             SbxLibraryError)
from sbsim_b200.config import (ActionConfig, AirHandler, Boiler, BoundedActionNormalizer,
                               FloorPlanBasedHvac, HistogramReducer, Hvac,
                               SetpointEnergyCarbonRegretFunction,
                               SetpointEnergyCarbonRewardFunction,
                               StandardScoreObservationNormalizer)
from sbsim_b200.convection import StochasticConvectionSimulator
from sbsim_b200.environment import BatchedWeather, Environment, SimulatorBuilding
from sbsim_b200.exogenous import (ConstantOccupancy, ElectricityEnergyCost,
                                  NaturalGasEnergyCost, RandomizedArrivalDepartureOccupancy,
                                  ReplayWeatherController,
                                  SetpointSchedule, StepFunctionOccupancy, TableOccupancy,
                                  WeatherController)
from sbsim_b200.floorplan import (CompiledPlan, MaterialProperties, compile_plan,
                                  legacy_building, legacy_building_plan)

__all__ = [
    "ActionConfig", "AirHandler", "BatchedWeather", "Boiler", "BoundedActionNormalizer",
    "CompiledPlan", "ConstantOccupancy", "ElectricityEnergyCost", "Environment",
    "FloorPlanBasedHvac", "HistogramReducer", "Hvac", "LIB_PATH", "MaterialProperties",
    "NaturalGasEnergyCost", "PATH_AUTO", "PATH_RESIDENT", "PATH_STREAMING",
    "RandomizedArrivalDepartureOccupancy",
    "ReplayWeatherController", "SbxLibraryError", "SetpointEnergyCarbonRegretFunction",
    "SetpointEnergyCarbonRewardFunction",
    "SetpointSchedule", "SimulatorBuilding", "StandardScoreObservationNormalizer",
    "StepFunctionOccupancy", "StochasticConvectionSimulator", "TableOccupancy", "WeatherController", "compile_plan", "legacy_building", "legacy_building_plan",
]
