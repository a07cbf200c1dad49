// Zone / grid means in the summation order of the reference (SBX_OPT_NUMPY_MEANS).
//
// The reference takes get_zone_average_temps as np.mean over the room's CVs in raster
// order (building.py:845-871) and the recirculation temperature as building.temp.mean()
// (simulator_flexible_floor_plan.py:134-150): float32 PAIRWISE sums as NumPy's add.reduce
// does them, divided once by the count.  The default zone sums of this library are exact
// integers (order-free, one rounding), a few ulp from that; with this option the sums follow
// NumPy's tree instead, so that the means -- and with them every thermostat decision and
// every later temperature field -- are bit-identical to the reference's.
//
// NumPy's tree over n values (pairwise sum, block size 128, 8 accumulators):
//   n <= 128 (a LEAF):  r[j] = a[j] (j < 8); r[j] += a[i + j] for i = 8, 16, .. below
//                       n - n % 8; ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7)); then the
//                       n % 8 tail values one by one.  (n < 8: the values one by one.)
//   n > 128:            n2 = n / 2 - (n / 2) % 8;  sum(a[:n2]) + sum(a[n2:]).
// The shape depends on n alone, so it is a per-plan table built once on the host
// (build_pairwise_tables in sbx_api.cu): the leaves of every zone's tree and of the grid's,
// and the inner nodes grouped by height.
//
//   k_pw_leaves    8 lanes per leaf, lane j = accumulator j (<= 15 dependent adds), three
//                  xor-shuffles for the fixed combine, the tail through shuffles.  Zone leaves
//                  gather through the plan's raster-ordered CV list; grid leaves are
//                  contiguous.  Each 8-lane load covers one 32-byte sector.
//   k_pw_combine   a warp (small trees) or a CTA per building: the inner nodes height by height, then
//                  sum / float(n) (one IEEE division) into pw_mean[b, 0..Z].
#pragma once
#include "sbx_device.cuh"

namespace sbx {

// kPwMetaInts, kPwGridFlag: sbx_device.cuh (k_resident_step<V, true> walks the same tables)
constexpr int kPwThreads = 256;

template <typename IDX>
__global__ void __launch_bounds__(kPwThreads) k_pw_leaves(const Params p) {
  const int capL = p.pw_capL;
  const long long g = ((long long)blockIdx.x * kPwThreads + threadIdx.x) >> 3;
  const int j = threadIdx.x & 7, lane = threadIdx.x & 31;
  const long long n_groups = (long long)(p.b_end - p.b_begin) * capL;
  const size_t n_cv = (size_t)p.H * p.W;
  bool on = g < n_groups;
  int b = 0, leaf = 0, plan = 0;
  uint2 lf = make_uint2(0u, 0u);
  if (on) {
    b = p.b_begin + (int)(g / capL);
    leaf = (int)(g - (long long)(b - p.b_begin) * capL);
    plan = p.n_plans == 1 ? 0 : b;
    lf = p.pw_leaf[(size_t)plan * capL + leaf];       // unused table entries have length 0
  }
  const int m = (int)lf.y;
  on = on && m > 0;
  const int main_n = m >= 8 ? (m & ~7) : 0, tail_n = m - main_n;   // tail_n < 8
  // All loads of the lane's chain are issued before the first add: the kernel is a latency
  // chain (leaf -> CV index -> temperature) with 16 short adds behind it.
  float v[16], tv = 0.f;
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = 0.f;
  if (on) {
    const float* t = p.tbuf[p.cur[b]] + (size_t)b * n_cv;
    const uint32_t start = lf.x & ~kPwGridFlag;
    if (lf.x & kPwGridFlag) {
      t += start;
#pragma unroll
      for (int k = 0; k < 16; ++k)
        if (j + 8 * k < main_n) v[k] = t[j + 8 * k];
      if (j < tail_n) tv = t[main_n + j];
    } else {
      const IDX* zl = reinterpret_cast<const IDX*>(p.pw_zlist) + (size_t)plan * n_cv + start;
      uint32_t ix[16], it = 0u;
#pragma unroll
      for (int k = 0; k < 16; ++k) ix[k] = j + 8 * k < main_n ? (uint32_t)zl[j + 8 * k] : 0u;
      if (j < tail_n) it = (uint32_t)zl[main_n + j];
#pragma unroll
      for (int k = 0; k < 16; ++k)
        if (j + 8 * k < main_n) v[k] = t[ix[k]];
      if (j < tail_n) tv = t[it];
    }
  }
  float res = v[0];                     // accumulator j: a[j], then a[j + 8], a[j + 16], ..
#pragma unroll
  for (int k = 1; k < 16; ++k)
    if (j + 8 * k < main_n) res = __fadd_rn(res, v[k]);
  // ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7)): IEEE addition commutes, so the xor
  // butterfly leaves exactly that value in every lane of the group
  res = __fadd_rn(res, __shfl_xor_sync(0xffffffffu, res, 1));
  res = __fadd_rn(res, __shfl_xor_sync(0xffffffffu, res, 2));
  res = __fadd_rn(res, __shfl_xor_sync(0xffffffffu, res, 4));
  // the n % 8 tail values one by one (lane k of the group loaded value k)
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    const float x = __shfl_sync(0xffffffffu, tv, (lane & ~7) + k);
    if (k < tail_n) res = __fadd_rn(res, x);
  }
  if (on && j == 0) p.pw_val[(size_t)b * p.pw_cap + leaf] = res;
}

// TEAM threads per building: a warp for small trees (no block barrier), the CTA for large ones
template <int TEAM>
__global__ void __launch_bounds__(kPwThreads) k_pw_combine(const Params p) {
  const int team = threadIdx.x / TEAM, tid = threadIdx.x % TEAM;
  const int b = p.b_begin + blockIdx.x * (kPwThreads / TEAM) + team;
  if (b >= p.b_end) return;             // uniform over the team
  const int plan = p.n_plans == 1 ? 0 : b, Z = p.Z;
  const int32_t* meta = p.pw_meta + (size_t)plan * kPwMetaInts;
  const int L = meta[0], n_levels = meta[2];
  float* val = p.pw_val + (size_t)b * p.pw_cap;
  const uint2* node = p.pw_node + (size_t)plan * p.pw_capI;
  for (int lv = 0; lv < n_levels; ++lv) {
    const int i1 = meta[3 + lv + 1];
    for (int i = meta[3 + lv] + tid; i < i1; i += TEAM) {
      const uint2 nd = node[i];
      val[L + i] = __fadd_rn(val[nd.x], val[nd.y]);
    }
    if constexpr (TEAM == 32) __syncwarp(); else __syncthreads();
  }
  const int2* root = p.pw_root + (size_t)plan * (Z + 1);
  for (int z = tid; z <= Z; z += TEAM) {
    const int2 r = root[z];           // {value index or -1, CVs}
    p.pw_mean[(size_t)b * (Z + 1) + z] = r.x >= 0 ? __fdiv_rn(val[r.x], (float)r.y) : 0.f;
  }
}

}  // namespace sbx
