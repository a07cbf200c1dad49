"""ArraySpec / BoundedArraySpec / TimeStep.

When TF-Agents is importable its own classes are used, so a `BatchedEnvironment`
plugs into `tf_agents` drivers directly.  TF-Agents is not part of this image
([TF-Agents, not in tree]), so otherwise minimal stand-ins
with the same attribute names are provided.
"""

from __future__ import annotations

import collections

import numpy as np

try:  # pragma: no cover - exercised only where tf_agents is installed
  from tf_agents.specs.array_spec import ArraySpec, BoundedArraySpec  # type: ignore
  from tf_agents.trajectories.time_step import StepType, TimeStep  # type: ignore
  HAVE_TF_AGENTS = True
except Exception:  # pylint: disable=broad-except
  HAVE_TF_AGENTS = False

  class ArraySpec:
    def __init__(self, shape, dtype, name=None):
      self.shape = tuple(shape)
      self.dtype = np.dtype(dtype)
      self.name = name

    def __repr__(self):
      return f"ArraySpec(shape={self.shape}, dtype={self.dtype!r}, name={self.name!r})"

    def __eq__(self, other):
      return (isinstance(other, ArraySpec) and self.shape == other.shape
              and self.dtype == other.dtype)

  class BoundedArraySpec(ArraySpec):
    def __init__(self, shape, dtype, minimum=None, maximum=None, name=None):
      super().__init__(shape, dtype, name)
      self.minimum = np.asarray(minimum, dtype=self.dtype)
      self.maximum = np.asarray(maximum, dtype=self.dtype)

    def __repr__(self):
      return (f"BoundedArraySpec(shape={self.shape}, dtype={self.dtype!r}, "
              f"name={self.name!r}, minimum={self.minimum}, maximum={self.maximum})")

  class StepType:
    FIRST = np.asarray(0, dtype=np.int32)
    MID = np.asarray(1, dtype=np.int32)
    LAST = np.asarray(2, dtype=np.int32)

  class TimeStep(collections.namedtuple(
      "TimeStep", ["step_type", "reward", "discount", "observation"])):
    __slots__ = ()

    def is_first(self):
      return np.equal(self.step_type, StepType.FIRST)

    def is_mid(self):
      return np.equal(self.step_type, StepType.MID)

    def is_last(self):
      return np.equal(self.step_type, StepType.LAST)
