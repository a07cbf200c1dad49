"""The kernels divide by a few per-building constants with Markstein's FMA
iteration on a correctly rounded reciprocal (sbx_device.cuh: div_rn).  This
checks the exact same sequence, emulated with NumPy (fp32 products are exact in
fp64, the residual FMA is exact by cancellation), against IEEE-754 fp32 division."""

import numpy as np


def _fma32(a, b, c):
  return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def div_rn(num, den):
  y = (1.0 / den.astype(np.float64)).astype(np.float32)
  q = (num * y).astype(np.float32)
  r = _fma32(-q, den, num)
  q = _fma32(r, y, q)
  r = _fma32(-q, den, num)
  return _fma32(r, y, q)


def test_markstein_division_matches_ieee():
  rng = np.random.default_rng(0)
  bad = 0
  for it in range(10):
    n = 1_000_000
    num = (rng.uniform(1, 2, n) * 2.0 ** rng.integers(-20, 40, n)).astype(np.float32)
    if it % 2 == 0:
      den = (rng.uniform(1, 2, n) * 2.0 ** rng.integers(-10, 30, n)).astype(np.float32)
    else:  # a handful of divisors, as in a building's coefficient table
      dd = rng.uniform(10, 1e6, 8).astype(np.float32)
      den = dd[rng.integers(0, 8, n)]
    bad += int((div_rn(num, den) != (num / den).astype(np.float32)).sum())
  assert bad == 0


def test_division_by_time_step():
  rng = np.random.default_rng(1)
  num = rng.uniform(0, 5e7, 2_000_000).astype(np.float32)
  den = np.full_like(num, 300.0)
  assert np.array_equal(div_rn(num, den), (num / den).astype(np.float32))
