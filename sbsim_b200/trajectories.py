"""Batched TimeSteps -> per-env trajectories.

The reference's SAC loop (`smart_control/notebooks/SAC_Demo.ipynb` cells 11, 34, 46, 48)
drives ONE environment with `actor.Actor(env, policy, ..., observers=[replay observer])`
and evaluates with `compute_avg_return`.  `sbsim_b200.Environment` steps B buildings per
call (`batched == True`), so the pieces a per-env consumer needs are:

  from_transition   (time_step, action, next_time_step) -> Trajectory, the tuple layout of
                    tf_agents.trajectories.trajectory.from_transition [TF-Agents, not in tree]
  Unbatcher         an observer for batched trajectories that fans them out to one
                    observer per env (e.g. one Reverb / replay-buffer writer per building)
  EpisodeBuffer     an observer that accumulates per-env episodes as [T, ...] arrays
  run_steps         the collect loop of PyDriver for a batched env (auto-reset included)
  compute_avg_return  cell 11 of the notebook, over all B buildings at once

Host-side glue only; nothing here touches the GPU path.
"""

from __future__ import annotations

import collections
from typing import Callable, List, Optional, Sequence

import numpy as np

from sbsim_b200 import specs

try:  # pragma: no cover - exercised only where tf_agents is installed
  from tf_agents.trajectories.trajectory import Trajectory  # type: ignore
except Exception:  # pylint: disable=broad-except

  class Trajectory(collections.namedtuple(
      "Trajectory", ["step_type", "observation", "action", "policy_info", "next_step_type",
                     "reward", "discount"])):
    __slots__ = ()

    def is_first(self):
      return np.equal(self.step_type, specs.StepType.FIRST)

    def is_last(self):
      return np.equal(self.step_type, specs.StepType.LAST)

    def is_boundary(self):
      return np.equal(self.step_type, specs.StepType.LAST)


def from_transition(time_step, action, next_time_step, policy_info=()) -> Trajectory:
  """tf_agents.trajectories.trajectory.from_transition: reward and discount come from the
  NEXT time step, step types from both."""
  return Trajectory(step_type=time_step.step_type, observation=time_step.observation,
                    action=action, policy_info=policy_info,
                    next_step_type=next_time_step.step_type, reward=next_time_step.reward,
                    discount=next_time_step.discount)


def _index(x, b):
  if isinstance(x, tuple) and not x:
    return x
  return np.asarray(x)[b]


class Unbatcher:
  """Observer: splits a batched Trajectory into B single-env trajectories.

  `observers` is either one callable per env or a single callable taking (env_index,
  trajectory).  Arrays are copied, because the env's TimeStep buffers are recycled."""

  def __init__(self, observers, batch_size: Optional[int] = None):
    if callable(observers):
      if batch_size is None:
        raise ValueError("batch_size is required with a single (env_index, trajectory) observer")
      self._fn, self._per_env = observers, None
      self.batch_size = int(batch_size)
    else:
      self._fn, self._per_env = None, list(observers)
      self.batch_size = len(self._per_env)

  def __call__(self, traj: Trajectory):
    n = np.asarray(traj.reward).shape[0]
    if n != self.batch_size:
      raise ValueError(f"trajectory batch {n} != {self.batch_size}")
    for b in range(n):
      one = Trajectory(*(np.copy(_index(f, b)) if not (isinstance(f, tuple) and not f) else f
                         for f in traj))
      if self._per_env is not None:
        self._per_env[b](one)
      else:
        self._fn(b, one)


class EpisodeBuffer:
  """Observer: accumulates batched trajectories and hands out finished episodes.

  All envs of a handle share one clock, so they finish together; `episodes()` returns a
  list of B dicts of [T, ...] arrays (observation, action, reward, discount, step_type,
  next_step_type) once a boundary step has been seen."""

  def __init__(self):
    self._steps: List[Trajectory] = []
    self._done: List[List[Trajectory]] = []

  def __call__(self, traj: Trajectory):
    self._steps.append(Trajectory(*(np.copy(f) if not (isinstance(f, tuple) and not f) else f
                                    for f in traj)))
    if np.all(np.equal(traj.next_step_type, specs.StepType.LAST)):
      self._done.append(self._steps)
      self._steps = []

  @property
  def n_finished(self) -> int:
    return len(self._done)

  def episodes(self) -> List[dict]:
    out = []
    for steps in self._done:
      stacked = {k: np.stack([getattr(s, k) for s in steps])
                 for k in ("observation", "action", "reward", "discount", "step_type",
                           "next_step_type")}
      n_env = stacked["reward"].shape[1]
      out.extend({k: v[:, b] for k, v in stacked.items()} for b in range(n_env))
    self._done = []
    return out


def run_steps(env, policy: Callable, num_steps: int, observers: Sequence[Callable] = (),
              time_step=None):
  """PyDriver.run for a batched env: `policy(time_step) -> action [B, A]`; every observer
  sees every batched Trajectory, including the boundary one whose next step is LAST; the
  step after a LAST step is the env's auto-reset (environment.py:1252-1253) and yields a
  FIRST time step without a transition.  Returns the final time step."""
  ts = time_step if time_step is not None else env.reset()
  for _ in range(num_steps):
    if np.all(np.equal(ts.step_type, specs.StepType.LAST)):
      ts = env.step(np.zeros((env.batch_size,) + tuple(env.action_spec().shape),
                             dtype=np.float32))            # auto-reset
      continue
    action = np.asarray(policy(ts), dtype=np.float32)
    # the env recycles its buffers every other call: keep what the trajectory needs
    prev = specs.TimeStep(step_type=np.copy(ts.step_type), reward=np.copy(ts.reward),
                          discount=np.copy(ts.discount), observation=np.copy(ts.observation))
    ts = env.step(action)
    traj = from_transition(prev, action, ts)
    for ob in observers:
      ob(traj)
  return ts


def compute_avg_return(env, policy: Callable, num_episodes: int = 1) -> np.ndarray:
  """SAC_Demo.ipynb cell 11 for B buildings at once: undiscounted return per env, averaged
  over `num_episodes` episodes.  Returns float64 [B]."""
  total = np.zeros(env.batch_size, dtype=np.float64)
  for _ in range(num_episodes):
    ts = env.reset()
    ret = np.zeros(env.batch_size, dtype=np.float64)
    while not np.all(np.equal(ts.step_type, specs.StepType.LAST)):
      ts = env.step(np.asarray(policy(ts), dtype=np.float32))
      ret += ts.reward
    total += ret
  return total / num_episodes
