"""Host-side glue for per-env consumers of the batched env."""

import numpy as np
import pytest

from sbsim_b200 import specs, trajectories as T


class FakeBatchedEnv:
  """Minimal batched py-environment with the episode logic of environment.py:1228-1368
  (N MID steps, then LAST, then auto-reset) and recycled output buffers."""

  batched = True

  def __init__(self, batch_size=3, episode_steps=4, obs_dim=2):
    self.batch_size = batch_size
    self._n = episode_steps
    self._bufs = [np.zeros((batch_size, obs_dim), np.float32) for _ in range(2)]
    self._flip = 0
    self._count = 0
    self._ended = True

  def action_spec(self):
    return specs.BoundedArraySpec((2,), np.float32, -1.0, 1.0, name="action")

  def _ts(self, step_type, reward, discount):
    obs = self._bufs[self._flip]
    self._flip ^= 1
    obs[:] = self._count + np.arange(self.batch_size)[:, None]
    B = self.batch_size
    return specs.TimeStep(np.full(B, step_type, np.int32), np.full(B, reward, np.float32),
                          np.full(B, discount, np.float32), obs)

  def reset(self):
    self._count, self._ended = 0, False
    return self._ts(0, 0.0, 1.0)

  def step(self, action):
    assert np.asarray(action).shape == (self.batch_size, 2)
    if self._ended:
      return self.reset()
    ended = self._count >= self._n
    self._ended = ended
    if not ended:
      self._count += 1
    return self._ts(2 if ended else 1, float(self._count), 0.0 if ended else 0.9)


def _policy(ts):
  return np.tile(np.array([[0.25, -0.5]], np.float32), (ts.reward.shape[0], 1))


def test_run_steps_unbatcher_and_episode_buffer():
  env = FakeBatchedEnv()
  per_env = [[] for _ in range(env.batch_size)]
  buf = T.EpisodeBuffer()
  unb = T.Unbatcher([lst.append for lst in per_env])
  ts = T.run_steps(env, _policy, 12, observers=[unb, buf])
  # 5 transitions per episode (4 MID + LAST), 1 auto-reset call between episodes
  assert buf.n_finished == 2
  eps = buf.episodes()
  assert len(eps) == 2 * env.batch_size
  for e in eps:
    assert e["reward"].shape == (5,) and e["observation"].shape == (5, 2)
    assert list(e["step_type"]) == [0, 1, 1, 1, 1] and list(e["next_step_type"]) == [1, 1, 1, 1, 2]
    assert e["discount"][-1] == 0.0 and np.all(e["discount"][:-1] == np.float32(0.9))
  # observations were copied (the env recycles two buffers): env b, step t saw t + b
  assert [float(x) for x in eps[1]["observation"][:, 0]] == [1, 2, 3, 4, 5]
  assert all(len(lst) == 10 for lst in per_env)
  assert per_env[2][0].observation.shape == (2,) and float(per_env[2][3].reward) == 4.0
  assert ts is not None


def test_unbatcher_single_callable_and_errors():
  seen = []
  unb = T.Unbatcher(lambda b, tr: seen.append((b, float(tr.reward))), batch_size=3)
  env = FakeBatchedEnv()
  T.run_steps(env, _policy, 1, observers=[unb])
  assert seen == [(0, 1.0), (1, 1.0), (2, 1.0)]
  with pytest.raises(ValueError):
    T.Unbatcher(lambda b, tr: None)
  with pytest.raises(ValueError):
    T.Unbatcher([print, print])(T.from_transition(env.reset(), _policy(env.reset()), env.step(_policy(env.reset()))))


def test_compute_avg_return_matches_cell_11():
  env = FakeBatchedEnv(batch_size=2, episode_steps=3)
  # rewards 1, 2, 3 on the MID steps and 3 on the LAST one
  np.testing.assert_allclose(T.compute_avg_return(env, _policy, num_episodes=2), [9.0, 9.0])


@pytest.mark.gpu
def test_run_steps_on_the_cuda_environment():
  import scenarios as S
  sc = S.Scenario(floor_plan=S.small_plan(), num_days=10 * 300.0 / 86400.0 + 1e-6)
  env = S.make_env(sc, n_envs=3)
  try:
    n = env._num_timesteps_in_episode
    buf = T.EpisodeBuffer()
    T.run_steps(env, lambda ts: np.zeros((3, 2), np.float32), n + 2, observers=[buf])
    eps = buf.episodes()
    assert len(eps) == 3 and eps[0]["reward"].shape == (n + 1,)
    assert eps[0]["next_step_type"][-1] == 2 and np.all(np.isfinite(eps[0]["observation"]))
    np.testing.assert_array_equal(eps[0]["observation"], eps[1]["observation"])
  finally:
    env.close()
