"""World-size-2 gloo test of the multi-GPU host logic (sharding, seeds, all-gather)."""

import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from sbsim_b200 import distributed, workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_the_batch():
  for n, w in [(262144, 8), (10, 3), (7, 8), (65536, 2)]:
    spans = [distributed.shard_range(n, r, w) for r in range(w)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1
  with pytest.raises(ValueError):
    distributed.shard_range(8, 2, 2)


def test_two_rank_gloo_gather_and_seeding(tmp_path):
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  port = s.getsockname()[1]
  s.close()
  out = str(tmp_path / "gathered.npy")
  env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
  procs = [subprocess.Popen([sys.executable, "-m", "sbsim_b200.distributed", str(r), "2",
                             str(port), out], cwd=ROOT, env=env,
                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
           for r in range(2)]
  for p in procs:
    try:
      log, _ = p.communicate(timeout=180)
    except subprocess.TimeoutExpired:
      for q in procs:
        q.kill()
      pytest.fail("gloo self-test timed out")
    assert p.returncode == 0, log.decode()[-2000:]
  got = np.load(out)
  # rank 1's shard (seed 2025) differs from rank 0's (seed 2024) and is reproducible
  w0 = workloads.randomized(6, seed=2024, n_layouts=2).weather_low
  w1 = workloads.randomized(6, seed=2025, n_layouts=2).weather_low
  np.testing.assert_allclose(got[:6], w0.astype(np.float32))
  np.testing.assert_allclose(got[6:], w1.astype(np.float32))
  assert not np.allclose(w0, w1)
