// Kernels of libsbx (sm_100a).
//
//  k_resident_step   one CTA per building, the whole diffusion solve in ONE launch:
//                    TMA bulk load of the temperature grid + descriptor into
//                    shared memory, Jacobi sweeps to convergence entirely on
//                    chip, zone / whole-grid sums, TMA bulk store of the new
//                    grid.  HBM sees each temperature once in and once out.
//  k_sweep           streaming path (grids that do not fit an SM, e.g. the
//                    744x1004 calibrated plan): one Jacobi sweep per launch,
//                    128-bit coalesced row loads, rolling 3-row register window,
//                    horizontal neighbours by warp shuffle, per-building max|dT|.
//  k_zone_reduce     segmented zone sums + whole-grid sum (fp64 accumulators).
//  k_pre / k_post    HVAC prologue / observation+reward epilogue, warp per building.
#pragma once

#include "sbx_device.cuh"

namespace sbx {

constexpr int kResidentThreads = 512;
constexpr int kStreamThreads = 256;
constexpr int kStreamRowsPerWarp = 8;

template <int V>
struct Vec;
template <>
struct Vec<4> {
  using F = float4;
  using D = uint2;  // 4 x u16
};
template <>
struct Vec<1> {
  using F = float;
  using D = uint16_t;
};

template <int V>
__device__ __forceinline__ void load_f(const float* p, float (&o)[V]) {
  if constexpr (V == 4) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  } else {
    o[0] = *p;
  }
}
template <int V>
__device__ __forceinline__ void store_f(float* p, const float (&o)[V]) {
  if constexpr (V == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
  } else {
    *p = o[0];
  }
}
template <int V>
__device__ __forceinline__ void load_d(const uint16_t* p, uint32_t (&o)[V]) {
  if constexpr (V == 4) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    o[0] = v.x & 0xFFFFu; o[1] = v.x >> 16; o[2] = v.y & 0xFFFFu; o[3] = v.y >> 16;
  } else {
    o[0] = *p;
  }
}
template <int V>
__device__ __forceinline__ void fill(float (&o)[V], float v) {
#pragma unroll
  for (int e = 0; e < V; ++e) o[e] = v;
}

__device__ __forceinline__ double env_ambient(const Params& p, int b, int s) {
  return p.fd_only ? p.fd_ambient[b]
                   : p.ambient[(size_t)(p.n_weather == 1 ? 0 : b) * p.T_rows + s];
}
__device__ __forceinline__ double env_convection(const Params& p, int b) {
  return p.fd_only ? p.fd_convection[b] : p.convection[p.n_weather == 1 ? 0 : b];
}

// ---------------------------------------------------------------------------
// mbarrier / TMA bulk-copy helpers (cp.async.bulk -> SASS UBLKCP)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---------------------------------------------------------------------------
// resident path
// ---------------------------------------------------------------------------

// Per-building header the solve kernels consume, built by k_build_header and
// fetched with one TMA bulk copy: the 30-entry coefficient table for this
// building's (h, T_inf), the heat input per diffuser CV of each zone (slot Z = 0:
// "no heat input"), and {T_inf, n_fast}.
__host__ __device__ inline size_t header_q_slots(int Z) { return (size_t)((Z + 1 + 3) / 4) * 4; }
__host__ __device__ inline size_t header_bytes(int Z) {
  return sizeof(Combo) * kNumCombos + header_q_slots(Z) * 4 + 16;
}

struct ResidentLayout {
  size_t off_a, off_b, off_n3, off_desc, off_list, off_hdr, off_bins, off_wmax,
      off_bar, total;
};

__host__ __device__ inline ResidentLayout resident_layout(int n_cv, int Z, int V) {
  ResidentLayout L;
  auto al = [](size_t v) { return (v + 15) & ~(size_t)15; };
  size_t o = 0;
  L.off_a = o; o = al(o + (size_t)n_cv * 4);
  L.off_b = o; o = al(o + (size_t)n_cv * 4);
  L.off_n3 = o; o = al(o + (size_t)n_cv * 4);
  L.off_desc = o; o = al(o + (size_t)n_cv * 2);
  L.off_list = o; o = al(o + (size_t)(n_cv / V) * 2);
  L.off_hdr = o; o = al(o + header_bytes(Z));
  L.off_bins = o; o = al(o + (size_t)(Z + 1) * 8 * (kResidentThreads / 32 + 1));
  L.off_wmax = o; o = al(o + 32 * 4);
  L.off_bar = o; o = al(o + 16);
  L.total = o;
  return L;
}

// Accumulates V consecutive CVs into per-zone shared bins (fixed point) + grid total.
// (streaming path: one atomic per run of equal zone ids)
template <int V>
__device__ __forceinline__ void zone_accumulate(const float (&t)[V], const uint32_t (&d)[V],
                                                long long* bins, long long& total) {
  int z0 = desc_zone(d[0]);
  long long run = 0;
#pragma unroll
  for (int e = 0; e < V; ++e) {
    const int z = desc_zone(d[e]);
    const long long v = to_fix(t[e]);
    total += v;
    if (z != z0) {
      if (z0 != SBX_ZONE_NONE) fix_add(&bins[z0], run);
      z0 = z;
      run = 0;
    }
    run += v;
  }
  if (z0 != SBX_ZONE_NONE) fix_add(&bins[z0], run);
}

// Warp-cooperative version for 32 consecutive vectors (one per lane): zone ids
// form contiguous segments along a row, so a segmented shuffle reduction leaves
// one partial sum per segment and only segment heads touch the (warp-private)
// bins.  Vectors that straddle a zone boundary flush their tail directly.
template <int V>
__device__ __forceinline__ void zone_accumulate_warp(const float (&t)[V], const uint32_t (&d)[V],
                                                     bool valid, int lane, long long* wbins,
                                                     long long& total) {
  int z = SBX_ZONE_NONE;
  long long s = 0;
  if (valid) {
    // primary zone of the vector = its first CV that belongs to a room (walls and
    // exterior carry SBX_ZONE_NONE); CVs of a second room in the same vector are
    // flushed directly (needs two rooms within V cells: practically never)
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const int ze = desc_zone(d[e]);
      const long long v = to_fix(t[e]);
      total += v;
      if (ze == SBX_ZONE_NONE) continue;
      if (z == SBX_ZONE_NONE) z = ze;
      if (ze == z) s += v;
      else fix_add(&wbins[ze], v);
    }
  }
  const int z_prev = __shfl_up_sync(0xffffffffu, z, 1);
  const bool head = lane == 0 || z_prev != z;
  const unsigned heads = __ballot_sync(0xffffffffu, head);
  const int seg = __popc(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long s_up = __shfl_down_sync(0xffffffffu, s, o);
    const int seg_up = __shfl_down_sync(0xffffffffu, seg, o);
    if (lane + o < 32 && seg_up == seg) s += s_up;
  }
  if (head && z != SBX_ZONE_NONE) fix_add(&wbins[z], s);
}

// Packed descriptor: combo index (5 bits) | diffuser flag | zone.
__device__ __forceinline__ uint32_t repack_desc(uint32_t d) {
  const int cls = desc_class(d);
  const bool half_u = cls == SBX_CV_EDGE_LEFT || cls == SBX_CV_EDGE_RIGHT || cls >= SBX_CV_CORNER_TL;
  const bool half_v = cls == SBX_CV_EDGE_TOP || cls == SBX_CV_EDGE_BOTTOM || cls >= SBX_CV_CORNER_TL;
  return (uint32_t)combo_index(d) | (half_u ? kPackHalfU : 0u) | (d & SBX_DESC_DIFFUSER) |
         (half_v ? kPackHalfV : 0u) | (d & 0xFF00u);
}
constexpr uint32_t kFastDesc = SBX_CV_INTERIOR | (0u << SBX_DESC_MATERIAL_SHIFT);  // interior, air, no diffuser

// Warp per building: coefficient table + per-zone heat + scalars -> hdr[b].
__global__ void __launch_bounds__(128) k_build_header(const Params p) {
  const int wpb = blockDim.x / 32;
  const int b = blockIdx.x * wpb + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= p.B) return;
  const int Z = p.Z;
  const int plan = p.n_plans == 1 ? 0 : b;
  unsigned char* h8 = p.hdr + (size_t)b * header_bytes(Z);
  Combo* tab = reinterpret_cast<Combo*>(h8);
  float* qcv = reinterpret_cast<float*>(h8 + sizeof(Combo) * kNumCombos);
  float* scal = qcv + header_q_slots(Z);
  const float t_inf = (float)env_ambient(p, b, p.time_index);   // tf_simulator.py:785
  const float h = (float)env_convection(p, b);
  build_combo_table(tab, p, plan, b, h, t_inf, lane, 32);
  for (int zi = lane; zi < (int)header_q_slots(Z); zi += 32)
    qcv[zi] = zi < Z ? p.qcv[(size_t)b * Z + zi] : 0.f;
  if (lane == 0) {
    scal[0] = t_inf;
    scal[1] = h;
    scal[2] = __int_as_float(p.n_fast ? p.n_fast[plan] : 0);
    scal[3] = 0.f;
  }
}

// Once per uploaded plan: repack the descriptors and split the plan's vectors
// into a FAST list (every CV interior, air, no heat input: uniform register
// coefficients, all four neighbours in range by construction) and a SLOW list
// (generic, table-driven).  Both lists share one array: fast from the front, slow
// from the back.
//
// Order inside each list: by zone, then by (rank within the vector's index-mod-8
// class, index mod 8).
//   * by zone: a warp's 32 consecutive entries almost always belong to one room, so
//     the zone sums reduce with one hardware REDUX per warp instead of a segmented
//     scan plus atomics;
//   * residue interleave: every aligned group of 8 entries holds 8 different
//     residues mod 8, so the 8 lanes of a quarter-warp touch 8 different 16-byte
//     bank groups and the 128-bit shared-memory accesses of the sweep stay
//     conflict-free although the lists skip vectors.
// One CTA per plan; ranks are computed by counting smaller keys (O(n^2) on a few
// thousand items, once per upload), so the order is deterministic.
constexpr int kPrepThreads = 256;
template <int V>
__global__ void __launch_bounds__(kPrepThreads) k_prepare_plan(const Params p) {
  extern __shared__ __align__(16) unsigned char smem_prep[];
  uint32_t* key = reinterpret_cast<uint32_t*>(smem_prep);
  const int n_items_ = p.H * p.W / V;
  uint16_t* key1 = reinterpret_cast<uint16_t*>(key + n_items_);
  __shared__ int s_nfast;
  const int plan = blockIdx.x, tid = threadIdx.x;
  const int n_cv = p.H * p.W, n_items = n_cv / V;
  const uint16_t* raw = p.desc + (size_t)plan * n_cv;
  uint16_t* packed = p.desc_packed + (size_t)plan * n_cv;
  uint16_t* qlist = p.qlist + (size_t)plan * n_items;
  if (tid == 0) s_nfast = 0;
  for (int i = tid; i < n_cv; i += kPrepThreads) packed[i] = (uint16_t)repack_desc(raw[i]);
  __syncthreads();
  // key1 = kind | zone | residue
  int my_fast = 0;
  for (int it = tid; it < n_items; it += kPrepThreads) {
    bool fast = true;
    int zone = SBX_ZONE_NONE;
    for (int e = 0; e < V; ++e) {
      const uint32_t d = raw[it * V + e];
      fast = fast && ((d & 0x007Fu) == kFastDesc);
      if (zone == SBX_ZONE_NONE) zone = desc_zone(d);
    }
#ifndef SBX_LIST_ZONE_SORT
    zone = 0;
#endif
    key1[it] = (uint16_t)(((fast ? 0u : 1u) << 11) | ((uint32_t)zone << 3) | (uint32_t)(it & 7));
    my_fast += fast ? 1 : 0;
  }
  atomicAdd(&s_nfast, my_fast);
  __syncthreads();
  // m = rank inside the (kind, zone, residue) bucket; key2 = kind | zone | m | residue
  for (int it = tid; it < n_items; it += kPrepThreads) {
    const uint32_t k1 = key1[it];
    int m = 0;
    for (int j = (it & 7); j < it; j += 8) m += (key1[j] == k1) ? 1 : 0;   // same residue only
    key[it] = ((k1 >> 3) << 18) | ((uint32_t)m << 3) | (k1 & 7u);
  }
  __syncthreads();
  // final rank = number of smaller keys (keys are unique)
  const int n_fast = s_nfast;
  for (int it = tid; it < n_items; it += kPrepThreads) {
    const uint32_t k2 = key[it];
    int rank = 0;
    for (int j = 0; j < n_items; ++j) rank += (key[j] < k2) ? 1 : 0;
    if (rank < n_fast) {
      qlist[rank] = (uint16_t)it;
    } else {
      bool diff = false;                 // bit 15: the vector holds a diffuser CV
      for (int e = 0; e < V; ++e) diff = diff || (raw[it * V + e] & SBX_DESC_DIFFUSER);
      qlist[n_items - 1 - (rank - n_fast)] = (uint16_t)(it | (diff ? 0x8000 : 0));
    }
  }
  if (tid == 0) p.n_fast[plan] = n_fast;
}

// One Jacobi sweep over the CTA's building.  FIRST: `in` is T_prev itself, so
// n3 = (cm * T_prev) / dt (tf_simulator.py:743-749) is computed here and stored
// for the following sweeps instead of being loaded.
template <int V, bool FIRST>
__device__ __forceinline__ float resident_sweep(
    const float* __restrict__ in, float* __restrict__ out, float* __restrict__ n3p,
    const uint16_t* __restrict__ dsc, const uint16_t* __restrict__ qlist, const Combo* tab,
    const float* qcv, const FastCoef& fc, const AreaCoef& az, float cm_fast, float dt, float rdt, float t_inf,
    int n_fast, int n_items, int H, int W, int Z, unsigned wq_magic, int tid) {
  constexpr int NT = kResidentThreads;
  const int wq = W / V;
  float lmax = 0.f;
  for (int i = tid; i < n_fast; i += NT) {
    const int base = (int)qlist[i] * V;
    float c[V], up[V], dn[V], n3v[V], o[V];
    load_f<V>(in + base, c);
    load_f<V>(in + base - W, up);
    load_f<V>(in + base + W, dn);
    if constexpr (FIRST) {
#pragma unroll
      for (int e = 0; e < V; ++e) n3v[e] = div_rn(mul(cm_fast, c[e]), dt, rdt);
      store_f<V>(n3p + base, n3v);
    } else {
      load_f<V>(n3p + base, n3v);
    }
    const float left = in[base - 1];
    const float right = in[base + V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const float t_jm = e == 0 ? left : c[e - 1];
      const float t_jp = e == V - 1 ? right : c[e + 1];
      o[e] = cv_update_fast(fc, t_jp, t_jm, up[e], dn[e], n3v[e]);
      lmax = fmaxf(lmax, fabsf(__fsub_rn(o[e], c[e])));                       // :851-853
    }
    store_f<V>(out + base, o);
  }
  const int n_slow = n_items - n_fast;
  // slow vectors are dealt from the LAST thread downwards: the fast loop leaves the
  // low thread ids with one more iteration, so this evens out the work per warp
  // before the sweep barrier
  for (int i = NT - 1 - tid; i < n_slow; i += NT) {
    const int entry = (int)qlist[n_items - 1 - i];
    const int it = entry & 0x7FFF;
    const bool has_q = (entry & 0x8000) != 0;
    const int base = it * V;
    const int r = (int)__umulhi((unsigned)it, wq_magic);
    const int q = it - r * wq;
    float c[V], up[V], dn[V], n3v[V], o[V];
    uint32_t d[V];
    load_f<V>(in + base, c);
    load_d<V>(dsc + base, d);
    if constexpr (FIRST) {
#pragma unroll
      for (int e = 0; e < V; ++e) n3v[e] = div_rn(mul(tab[d[e] & 31u].cm, c[e]), dt, rdt);
      store_f<V>(n3p + base, n3v);
    } else {
      load_f<V>(n3p + base, n3v);
    }
    if (r > 0) load_f<V>(in + base - W, up); else fill<V>(up, t_inf);         // :642-644
    if (r < H - 1) load_f<V>(in + base + W, dn); else fill<V>(dn, t_inf);     // :646
    const float left = q > 0 ? in[base - 1] : t_inf;                          // :638-640
    const float right = q < wq - 1 ? in[base + V] : t_inf;                    // :636
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const float t_jm = e == 0 ? left : c[e - 1];
      const float t_jp = e == V - 1 ? right : c[e + 1];
      float qv = 0.f;
      if (has_q) qv = qcv[(d[e] & SBX_DESC_DIFFUSER) ? (int)(d[e] >> SBX_DESC_ZONE_SHIFT) : Z];
      o[e] = cv_update_packed(d[e], t_jp, t_jm, up[e], dn[e], n3v[e], qv, t_inf, az, tab);
      lmax = fmaxf(lmax, fabsf(__fsub_rn(o[e], c[e])));
    }
    store_f<V>(out + base, o);
  }
  return lmax;
}

// Optional per-phase cycle counters (build with -DSBX_PROFILE_PHASES; read back with
// SBX_F_PHASE_CYCLES).  Thread 0 of every CTA adds its phase durations.
#ifdef SBX_PROFILE_PHASES
#define SBX_PHASE(i)                                                                 \
  do {                                                                               \
    if (tid == 0) {                                                                  \
      const long long now__ = clock64();                                             \
      atomicAdd(&p.phase_cycles[i], (unsigned long long)(now__ - phase_t0__));       \
      phase_t0__ = now__;                                                            \
    }                                                                                \
  } while (0)
#else
#define SBX_PHASE(i) do {} while (0)
#endif

template <int V>
__global__ void __launch_bounds__(kResidentThreads, 2) k_resident_step(const Params p) {
  extern __shared__ __align__(128) unsigned char smem[];
#ifdef SBX_PROFILE_PHASES
  long long phase_t0__ = clock64();
#endif
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NT = kResidentThreads, NW = NT / 32;
  const int H = p.H, W = p.W, Z = p.Z;
  const int n_cv = H * W;
  const int n_items = n_cv / V;
  const int plan = p.n_plans == 1 ? 0 : b;
  const ResidentLayout L = resident_layout(n_cv, Z, V);
  float* bufA = reinterpret_cast<float*>(smem + L.off_a);
  float* bufB = reinterpret_cast<float*>(smem + L.off_b);
  float* n3p = reinterpret_cast<float*>(smem + L.off_n3);
  uint16_t* dsc = reinterpret_cast<uint16_t*>(smem + L.off_desc);
  uint16_t* qlist = reinterpret_cast<uint16_t*>(smem + L.off_list);
  Combo* tab = reinterpret_cast<Combo*>(smem + L.off_hdr);
  float* qcv = reinterpret_cast<float*>(smem + L.off_hdr + sizeof(Combo) * kNumCombos);
  const float* scal = qcv + header_q_slots(Z);
  long long* bins = reinterpret_cast<long long*>(smem + L.off_bins);
  float* wmax = reinterpret_cast<float*>(smem + L.off_wmax);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.off_bar);

  float* gT = p.tbuf[0] + (size_t)b * n_cv;
  const uint16_t* gD = p.desc_packed + (size_t)plan * n_cv;
  const uint16_t* gL = p.qlist + (size_t)plan * n_items;
  const uint32_t hdr_bytes = (uint32_t)header_bytes(Z);
  const unsigned char* gH = p.hdr + (size_t)b * hdr_bytes;
  const bool use_tma = (n_cv % 8) == 0 && (n_items % 8) == 0;

  // ---- stage 0: start the bulk loads (TMA, one mbarrier) ----------------------
  if (use_tma) {
    if (tid == 0) {
      mbar_init(bar, 1);
      mbar_expect_tx(bar, (uint32_t)(n_cv * 4 + n_cv * 2 + n_items * 2) + hdr_bytes);
      tma_load_1d(smem + L.off_hdr, gH, hdr_bytes, bar);
      tma_load_1d(bufA, gT, (uint32_t)(n_cv * 4), bar);
      tma_load_1d(dsc, gD, (uint32_t)(n_cv * 2), bar);
      tma_load_1d(qlist, gL, (uint32_t)(n_items * 2), bar);
    }
  } else {
    for (int i = tid; i < n_cv; i += NT) {
      bufA[i] = gT[i];
      dsc[i] = gD[i];
    }
    for (int i = tid; i < n_items; i += NT) qlist[i] = gL[i];
    for (int i = tid; i < (int)(hdr_bytes / 4); i += NT)
      reinterpret_cast<uint32_t*>(smem + L.off_hdr)[i] = reinterpret_cast<const uint32_t*>(gH)[i];
  }

  // ---- stage 1: zero the zone bins while the copies are in flight ---------------
  for (int i = tid; i < (Z + 1) * (NW + 1); i += NT) bins[i] = 0;
  __syncthreads();
  SBX_PHASE(0);   // launch .. TMA issued
  if (use_tma) mbar_wait(bar, 0);
  SBX_PHASE(1);   // waiting for the TMA loads
  const float t_inf = scal[0];
  const int n_fast = __float_as_int(scal[2]);
  FastCoef fc;
  {
    const Combo& c = tab[SBX_CV_INTERIOR * kNumMaterials + 0];
    fc.kq = c.k1; fc.vz = c.vz; fc.den = c.den; fc.rden = c.rden;
  }
  const float cm_fast = tab[SBX_CV_INTERIOR * kNumMaterials + 0].cm;
  AreaCoef az;
  az.full = tab[SBX_CV_INTERIOR * kNumMaterials].vz;
  az.half = tab[SBX_CV_CORNER_TL * kNumMaterials].vz;
  const float rdt = __frcp_rn(p.dt);
  const unsigned wq_magic = 0xFFFFFFFFu / (unsigned)(W / V) + 1u;   // it / wq for it < 2^16

  // ---- stage 2: Jacobi sweeps to convergence (simulator.py:348-364) ---------
  float* in = bufA;
  float* out = bufB;
  int k = 0;
  float md = 0.f, last_lmax = 0.f;
  const int limit = p.iteration_limit;
  while (k < limit) {
    ++k;
    float lmax;
    if (k == 1)
      lmax = resident_sweep<V, true>(in, out, n3p, dsc, qlist, tab, qcv, fc, az, cm_fast, p.dt, rdt,
                                     t_inf, n_fast, n_items, H, W, Z, wq_magic, tid);
    else
      lmax = resident_sweep<V, false>(in, out, n3p, dsc, qlist, tab, qcv, fc, az, cm_fast, p.dt, rdt,
                                      t_inf, n_fast, n_items, H, W, Z, wq_magic, tid);
    // max|dT| <= threshold  <=>  no thread saw a delta above it (simulator.py:362);
    // the barrier doubles as the ping-pong hazard fence.
    const int above = __syncthreads_or(lmax > p.threshold);
    last_lmax = lmax;
    float* tmp = in; in = out; out = tmp;
    if (k == 1) SBX_PHASE(2); else SBX_PHASE(3);   // first sweep (computes n3) / later sweeps
    if (!above) break;
  }
  // `in` now holds building.temp for the next step (simulator.py:369)

  // building.apply_convection() (simulator_flexible_floor_plan.py:156): the swaps of
  // the stochastic convection model, pre-composed on the host into a gather map
  if (p.conv_perm != nullptr && !p.fd_only) {
    const int32_t* perm = p.conv_perm + (size_t)b * n_cv;
    for (int i = tid; i < n_cv; i += NT) out[i] = in[perm[i]];
    __syncthreads();
    float* tmp = in; in = out; out = tmp;
  }

  // ---- stage 3: write back + zone / grid sums ---------------------------------
  if (use_tma) {
    if (tid == 0) tma_store_1d(gT, in, (uint32_t)(n_cv * 4));
  } else {
    for (int i = tid; i < n_cv; i += NT) gT[i] = in[i];
  }
  // block max of the last sweep (diagnostic SBX_F_MAX_DELTA)
  last_lmax = warp_max(last_lmax);
  if (lane == 0) wmax[warp] = last_lmax;
  long long total = 0;
  long long* wbins = bins + (size_t)(warp + 1) * (Z + 1);   // warp-private bins
  if (!p.fd_only) {
    // Zone sums: every lane converts its vector to integers and the warp reduces
    // them zone by zone with exact hardware integer REDUX (a warp's 32 vectors touch
    // a handful of rooms).  The running sum of zone z lives in a REGISTER of lane
    // z mod 32; shared memory is touched once per warp at the end.  No atomics in the
    // common case (a shared-memory 64-bit atomic is a CAS loop on this hardware).
    long long acc = 0;            // this lane's zone (lane, lane+32, ...: see flush below)
    int acc_zone = lane;
    for (int base_u = warp * 32; base_u < n_items; base_u += NT) {
      const int it = base_u + lane;
      int z1 = SBX_ZONE_NONE, z2 = SBX_ZONE_NONE, s1 = 0, s2 = 0;
      if (it < n_items) {
        float t[V];
        uint32_t d[V];
        load_f<V>(in + it * V, t);
        load_d<V>(dsc + it * V, d);
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const int v = to_fix32(t[e], t_inf);
          const int ze = desc_zone(d[e]);
          total += v;
          if (ze == SBX_ZONE_NONE) continue;
          if (z1 == SBX_ZONE_NONE || ze == z1) { z1 = ze; s1 += v; }
          else if (z2 == SBX_ZONE_NONE || ze == z2) { z2 = ze; s2 += v; }   // room | wall | room
          else fix_add(&bins[ze], (long long)v);      // 3 rooms in one vector: CTA-level row 0, atomic
        }
      }
      unsigned todo1 = __ballot_sync(0xffffffffu, z1 != SBX_ZONE_NONE);
      unsigned todo2 = __ballot_sync(0xffffffffu, z2 != SBX_ZONE_NONE);
      while (todo1 | todo2) {
        const int z0 = todo1 ? __shfl_sync(0xffffffffu, z1, __ffs(todo1) - 1)
                             : __shfl_sync(0xffffffffu, z2, __ffs(todo2) - 1);
        const bool m1 = (z1 == z0), m2 = (z2 == z0);
        const long long sum = warp_sum_i32((m1 ? s1 : 0) + (m2 ? s2 : 0));
        todo1 &= ~__ballot_sync(0xffffffffu, m1);
        todo2 &= ~__ballot_sync(0xffffffffu, m2);
        if (lane == (z0 & 31)) {
          if (z0 != acc_zone) {          // only with more than 32 zones
            wbins[acc_zone] += acc;      // warp-private, this lane owns zones = lane mod 32
            acc_zone = z0;
            acc = 0;
          }
          acc += sum;
        }
      }
    }
    __syncwarp();
    if (acc_zone < Z && acc != 0) wbins[acc_zone] += acc;
    __syncwarp();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    if (lane == 0) wbins[Z] = total;
  }
  __syncthreads();
  SBX_PHASE(4);   // convection gather (if any) + zone / grid sums
  if (tid == 0) {
    md = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) md = fmaxf(md, wmax[w]);
    p.n_sweeps[b] = k;
    p.max_delta[b] = md;
    if (!p.fd_only) atomicAdd(p.sweeps_total, (unsigned long long)k);
  }
  if (!p.fd_only) {
    // combine the warp-private bins (integer sums: exact, order-free), handed to
    // k_post through zone_sum[b, 0..Z]
    long long* zs = p.zone_sum + (size_t)b * (Z + 1);
    if (tid == 0) p.zone_ref[b] = t_inf;
    for (int i = tid; i <= Z; i += NT) {
      long long acc = 0;
#pragma unroll
      for (int w = 0; w <= NW; ++w) acc += bins[(size_t)w * (Z + 1) + i];
      zs[i] = acc;
    }
  }
  if (use_tma && tid == 0) tma_store_wait();
}


// ---------------------------------------------------------------------------
// Gauss-Seidel solver (the reference's legacy model, simulator.py:98-316): fp64,
// in place, raster order.  Cell (i, j) reads the ALREADY UPDATED (i-1, j), (i, j-1)
// and the old (i+1, j), (i, j+1), so the cells of one anti-diagonal i + j = d are
// independent: sweeping d = 0 .. H+W-2 with a barrier in between reproduces the
// sequential raster sweep exactly, operation for operation (Python floats are
// IEEE fp64, evaluated left to right as written in the reference).
// One CTA per building, grid resident in shared memory.
// ---------------------------------------------------------------------------
constexpr int kGsThreads = 128;

struct GsLayout {
  size_t off_t, off_prev, off_desc, off_q, off_bins, total;
};
__host__ __device__ inline GsLayout gs_layout(int n_cv, int Z) {
  GsLayout L;
  auto al = [](size_t v) { return (v + 15) & ~(size_t)15; };
  size_t o = 0;
  L.off_t = o; o = al(o + (size_t)n_cv * 8);
  L.off_prev = o; o = al(o + (size_t)n_cv * 8);
  L.off_q = o; o = al(o + (size_t)(Z + 1) * 8);
  L.off_bins = o; o = al(o + (size_t)(Z + 1) * 8);
  L.off_desc = o; o = al(o + (size_t)n_cv * 2);
  L.total = o;
  return L;
}

// which of (up, down, left, right) exist, by class (tf_simulator.py:208-243)
__device__ __forceinline__ unsigned gs_neighbor_mask(int cls) {
  switch (cls) {
    case SBX_CV_INTERIOR: return 0xF;      // up | down<<1 | left<<2 | right<<3
    case SBX_CV_EDGE_TOP: return 0xE;
    case SBX_CV_EDGE_BOTTOM: return 0xD;
    case SBX_CV_EDGE_LEFT: return 0xB;
    case SBX_CV_EDGE_RIGHT: return 0x7;
    case SBX_CV_CORNER_TL: return 0xA;     // down, right
    case SBX_CV_CORNER_BL: return 0x9;     // up, right
    case SBX_CV_CORNER_TR: return 0x6;     // down, left
    case SBX_CV_CORNER_BR: return 0x5;     // up, left
    default: return 0x0;
  }
}

__global__ void __launch_bounds__(kGsThreads) k_resident_gs(const Params p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int H = p.H, W = p.W, Z = p.Z, n_cv = H * W;
  const int plan = p.n_plans == 1 ? 0 : b;
  const GsLayout L = gs_layout(n_cv, Z);
  double* T = reinterpret_cast<double*>(smem + L.off_t);
  double* Tp = reinterpret_cast<double*>(smem + L.off_prev);
  double* qz = reinterpret_cast<double*>(smem + L.off_q);
  long long* bins = reinterpret_cast<long long*>(smem + L.off_bins);
  uint16_t* dsc = reinterpret_cast<uint16_t*>(smem + L.off_desc);
  double* gT = p.temp64 + (size_t)b * n_cv;
  const uint16_t* gD = p.desc + (size_t)plan * n_cv;
  for (int i = tid; i < n_cv; i += kGsThreads) {
    const double v = gT[i];
    T[i] = v;
    Tp[i] = v;
    dsc[i] = gD[i];
  }
  for (int i = tid; i <= Z; i += kGsThreads) {
    qz[i] = i < Z ? p.qcv64[(size_t)b * Z + i] : 0.0;
    bins[i] = 0;
  }
  const double t_inf = env_ambient(p, b, p.time_index);
  const double h = env_convection(p, b);
  const double dx = p.cv_size[plan];
  const double z = (double)p.z_double;
  const double dt = (double)p.dt_double;
  const double* mat = p.material + (size_t)plan * 9;
  __syncthreads();

  int k = 0;
  double md_block = 0.0;
  while (k < p.iteration_limit) {
    ++k;
    double lmax = 0.0;
    for (int d = 0; d <= H + W - 2; ++d) {
      const int i_lo = max(0, d - W + 1), i_hi = min(H - 1, d);
      for (int i = i_lo + tid; i <= i_hi; i += kGsThreads) {
        const int j = d - i;
        const int c = i * W + j;
        const uint32_t desc = dsc[c];
        const int cls = desc_class(desc);
        double nv;
        if (cls == SBX_CV_EXTERIOR) {
          nv = t_inf;                                            // simulator.py:256-258
        } else {
          const double* m = mat + desc_material(desc) * 3;
          const double kc = m[0], cap = m[1], rho = m[2];
          const double last = Tp[c];
          const unsigned mask = gs_neighbor_mask(cls);
          // neighbours in the reference's list order (building.py:808)
          const int nb[4] = {c - W, c + W, c - 1, c + 1};
          if (cls == SBX_CV_INTERIOR) {                          // :210-237
            const double alpha = kc / rho / cap;
            const double t0 = dx * dx / dt / alpha;
            double sum = 0.0;
#pragma unroll
            for (int n = 0; n < 4; ++n) sum += T[nb[n]];
            const double qv = (desc & SBX_DESC_DIFFUSER) ? qz[desc_zone(desc)] : 0.0;
            const double source = qv / kc / z;
            nv = (sum + source + t0 * last) / (4.0 + t0);
          } else if (cls >= SBX_CV_CORNER_TL) {                  // corner :117-142
            const double t0 = rho * (dx * dx) * cap / dt / 2.0;
            double sum = 0.0;
#pragma unroll
            for (int n = 0; n < 4; ++n) if (mask & (1u << n)) sum += T[nb[n]];
            const double transfer = kc * sum;
            const double conv = 2.0 * h * dx * t_inf;
            const double den = 2.0 * kc + 2.0 * h * dx + t0;
            nv = (transfer + conv + t0 * last) / den;
          } else {                                               // edge :163-195
            const double t0 = rho * (dx * dx) / 2 * cap / dt;
            double sum = 0.0;
#pragma unroll
            for (int n = 0; n < 4; ++n) {
              if (mask & (1u << n)) {
                // 0.5 for neighbours that are themselves edges / corners (:180-183)
                const double f = desc_class(dsc[nb[n]]) == SBX_CV_INTERIOR ? 1.0 : 0.5;
                sum += f * T[nb[n]];
              }
            }
            const double transfer = kc * sum;
            const double conv = h * dx * t_inf;
            const double den = 2.0 * kc + h * dx + t0;
            nv = (transfer + conv + t0 * last) / den;
          }
        }
        lmax = fmax(lmax, fabs(nv - T[c]));                      // :311-312
        T[c] = nv;
      }
      __syncthreads();
    }
    // max_delta <= convergence_threshold (simulator.py:362), fp64
    const int above = __syncthreads_or(lmax > p.threshold64);
    md_block = lmax;
    if (!above) break;
  }
  // write back (fp64 state + fp32 mirror), zone sums for k_post
  float* g32 = p.tbuf[0] + (size_t)b * n_cv;
  long long total = 0;
  for (int i = tid; i < n_cv; i += kGsThreads) {
    const double v = T[i];
    gT[i] = v;
    g32[i] = (float)v;
    const long long f = to_fix(v);
    total += f;
    const int zn = desc_zone(dsc[i]);
    if (zn != SBX_ZONE_NONE && !p.fd_only) fix_add(&bins[zn], f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    total += __shfl_xor_sync(0xffffffffu, total, o);
    md_block = fmax(md_block, __shfl_xor_sync(0xffffffffu, md_block, o));
  }
  if ((tid & 31) == 0) {
    fix_add(&bins[Z], total);
    atomicMax(&p.max_delta_bits[b], __float_as_uint((float)md_block));
  }
  __syncthreads();
  if (tid == 0) {
    p.n_sweeps[b] = k;
    p.max_delta[b] = __uint_as_float(p.max_delta_bits[b]);
    p.max_delta_bits[b] = 0u;
    if (!p.fd_only) atomicAdd(p.sweeps_total, (unsigned long long)k);
  }
  if (!p.fd_only) {
    long long* zs = p.zone_sum + (size_t)b * (Z + 1);
    if (tid == 0) p.zone_ref[b] = 0.f;
    for (int i = tid; i <= Z; i += kGsThreads) zs[i] = bins[i];
  }
}

// ---------------------------------------------------------------------------
// streaming path
// ---------------------------------------------------------------------------

struct StreamTiling {
  int tiles_x, tiles_y, tiles;
};
__host__ __device__ inline StreamTiling stream_tiling(int H, int W, int V) {
  StreamTiling t;
  const int tile_w = 32 * V;
  const int tile_h = (kStreamThreads / 32) * kStreamRowsPerWarp;
  t.tiles_x = (W + tile_w - 1) / tile_w;
  t.tiles_y = (H + tile_h - 1) / tile_h;
  t.tiles = t.tiles_x * t.tiles_y;
  return t;
}

__device__ __forceinline__ void sweep_buffers(int cur, int k, int& in, int& out) {
  const int s0 = (cur + 1) % 3, s1 = (cur + 2) % 3;
  in = k == 1 ? cur : ((k & 1) == 0 ? s0 : s1);
  out = (k & 1) ? s0 : s1;
}

// One Jacobi sweep of every still-active building.  Sweep index k is 1-based.
template <int V>
__global__ void __launch_bounds__(kStreamThreads) k_sweep(const Params p, const int k) {
  __shared__ Combo tab[kNumCombos];
  __shared__ float qcv[kMaxZones + 1];
  const StreamTiling tl = stream_tiling(p.H, p.W, V);
  const int b = blockIdx.x / tl.tiles;
  if (!p.active[b]) return;
  const int tile = blockIdx.x - b * tl.tiles;
  const int ty = tile / tl.tiles_x, tx = tile - ty * tl.tiles_x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = p.H, W = p.W, Z = p.Z;
  const size_t n_cv = (size_t)H * W;
  const int plan = p.n_plans == 1 ? 0 : b;
  const float t_inf = (float)env_ambient(p, b, p.time_index);
  const float h = (float)env_convection(p, b);
  build_combo_table(tab, p, plan, b, h, t_inf, tid, kStreamThreads);
  for (int i = tid; i < Z; i += kStreamThreads) qcv[i] = p.qcv[(size_t)b * Z + i];
  __syncthreads();

  const int cur = p.cur[b];
  int bi, bo;
  sweep_buffers(cur, k, bi, bo);
  const float* __restrict__ tin = p.tbuf[bi] + (size_t)b * n_cv;
  const float* __restrict__ tprev = p.tbuf[cur] + (size_t)b * n_cv;
  float* __restrict__ tout = p.tbuf[bo] + (size_t)b * n_cv;
  const uint16_t* __restrict__ dsc = p.desc + (size_t)plan * n_cv;

  const int c0 = (tx * 32 + lane) * V;
  const int r0 = (ty * (kStreamThreads / 32) + warp) * kStreamRowsPerWarp;
  const bool col_ok = c0 < W;
  float lmax = 0.f;
  if (r0 < H) {
    const int r1 = min(r0 + kStreamRowsPerWarp, H);
    float up[V], c[V], dn[V];
    fill<V>(up, t_inf);
    fill<V>(c, t_inf);
    if (col_ok) {
      if (r0 > 0) load_f<V>(tin + (size_t)(r0 - 1) * W + c0, up);
      load_f<V>(tin + (size_t)r0 * W + c0, c);
    }
    for (int r = r0; r < r1; ++r) {
      fill<V>(dn, t_inf);
      if (col_ok && r + 1 < H) load_f<V>(tin + (size_t)(r + 1) * W + c0, dn);
      // horizontal neighbours: shuffle inside the warp, global load at its ends
      float left = __shfl_up_sync(0xffffffffu, c[V - 1], 1);
      float right = __shfl_down_sync(0xffffffffu, c[0], 1);
      if (col_ok) {
        if (lane == 0) left = c0 > 0 ? tin[(size_t)r * W + c0 - 1] : t_inf;
        if (lane == 31 || c0 + V >= W) right = c0 + V < W ? tin[(size_t)r * W + c0 + V] : t_inf;
        float tp[V], o[V];
        uint32_t d[V];
        load_f<V>(tprev + (size_t)r * W + c0, tp);
        load_d<V>(dsc + (size_t)r * W + c0, d);
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const float t_jm = e == 0 ? left : c[e - 1];
          const float t_jp = e == V - 1 ? right : c[e + 1];
          const float n3 = fdiv(mul(cv_cm(d[e], tab), tp[e]), p.dt);
          const float qv = (d[e] & SBX_DESC_DIFFUSER) ? qcv[desc_zone(d[e])] : 0.f;
          o[e] = cv_update(d[e], t_jp, t_jm, up[e], dn[e], n3, qv, t_inf, tab);
          lmax = fmaxf(lmax, fabsf(__fsub_rn(o[e], c[e])));
        }
        store_f<V>(tout + (size_t)r * W + c0, o);
      }
#pragma unroll
      for (int e = 0; e < V; ++e) { up[e] = c[e]; c[e] = dn[e]; }
    }
  }
  lmax = warp_max(lmax);
  if (lane == 0 && lmax > 0.f) atomicMax(&p.max_delta_bits[b], __float_as_uint(lmax));
}

// Convergence bookkeeping after sweep k (simulator.py:348-369).
__global__ void k_check(const Params p, const int k) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  int still = 0;
  if (b < p.B && p.active[b]) {
    const float md = __uint_as_float(p.max_delta_bits[b]);
    p.max_delta_bits[b] = 0u;
    p.max_delta[b] = md;
    p.n_sweeps[b] = k;
    if (md <= p.threshold || k >= p.iteration_limit) {
      int bi, bo;
      sweep_buffers(p.cur[b], k, bi, bo);
      p.cur[b] = (uint8_t)bo;
      p.active[b] = 0;
      if (!p.fd_only) atomicAdd(p.sweeps_total, (unsigned long long)k);
    } else {
      still = 1;
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, still);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(p.n_active, __popc(m));
}

__global__ void k_activate(const Params p) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < p.B) {
    p.active[b] = 1;
    p.max_delta_bits[b] = 0u;
    p.n_sweeps[b] = 0;
  }
}

// Zone sums + grid total of building.temp (building.py:845-871, simulator.py:408).
template <int V>
__global__ void __launch_bounds__(kStreamThreads) k_zone_reduce(const Params p) {
  __shared__ long long bins[kMaxZones + 1];
  const StreamTiling tl = stream_tiling(p.H, p.W, V);
  const int b = blockIdx.x / tl.tiles;
  const int tile = blockIdx.x - b * tl.tiles;
  const int ty = tile / tl.tiles_x, tx = tile - ty * tl.tiles_x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = p.H, W = p.W, Z = p.Z;
  const size_t n_cv = (size_t)H * W;
  const int plan = p.n_plans == 1 ? 0 : b;
  for (int i = tid; i <= Z; i += kStreamThreads) bins[i] = 0;
  __syncthreads();
  const float* __restrict__ t = p.tbuf[p.cur[b]] + (size_t)b * n_cv;
  const uint16_t* __restrict__ dsc = p.desc + (size_t)plan * n_cv;
  const int c0 = (tx * 32 + lane) * V;
  const int r0 = (ty * (kStreamThreads / 32) + warp) * kStreamRowsPerWarp;
  long long total = 0;
  if (c0 < W && r0 < H) {
    const int r1 = min(r0 + kStreamRowsPerWarp, H);
    for (int r = r0; r < r1; ++r) {
      float tv[V];
      uint32_t d[V];
      load_f<V>(t + (size_t)r * W + c0, tv);
      load_d<V>(dsc + (size_t)r * W + c0, d);
      zone_accumulate<V>(tv, d, bins, total);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
  if (lane == 0) fix_add(&bins[Z], total);
  __syncthreads();
  long long* zs = p.zone_sum + (size_t)b * (Z + 1);
  if (tile == 0 && tid == 0) p.zone_ref[b] = 0.f;
  for (int i = tid; i <= Z; i += kStreamThreads)
    if (bins[i] != 0) fix_add(&zs[i], bins[i]);
}

struct CarryStore {
  double ahu_flow, boiler_flow, return_water;
  int32_t ahu_count, boiler_count;
};

// HVAC prologue, warp per building (streaming path).
__global__ void __launch_bounds__(128) k_pre(const Params p, CarryStore* carry) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int wpb = blockDim.x / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * wpb + warp;
  if (b >= p.B) return;
  const int Z = p.Z;
  double* scratch = reinterpret_cast<double*>(smem) + (size_t)warp * (3 * Z + 64 + Z);
  float* zpre = reinterpret_cast<float*>(scratch + 3 * Z + 64);
  const int plan = p.n_plans == 1 ? 0 : b;
  for (int zi = lane; zi < Z; zi += 32) zpre[zi] = p.zone_mean[(size_t)b * Z + zi];
  __syncwarp();
  PreOut o = hvac_pre(p, b, plan, lane, zpre, p.global_mean[b], scratch);
  if (lane == 0) {
    CarryStore c;
    c.ahu_flow = o.ahu_flow; c.boiler_flow = o.boiler_flow; c.return_water = o.return_water;
    c.ahu_count = o.ahu_count; c.boiler_count = o.boiler_count;
    carry[b] = c;
  }
}

// Means from zone sums, then observation + reward; warp per building.
__global__ void __launch_bounds__(128) k_post(const Params p, const CarryStore* carry,
                                              const int is_reset) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int wpb = blockDim.x / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * wpb + warp;
  if (b >= p.B) return;
  const int Z = p.Z;
  double* scratch = reinterpret_cast<double*>(smem) + (size_t)warp * (3 * Z + 64 + Z);
  float* zpre = reinterpret_cast<float*>(scratch + 3 * Z + 64);
  float* zpost = zpre + Z;
  const int plan = p.n_plans == 1 ? 0 : b;
  const int32_t* ncv = p.zone_ncv + (size_t)plan * Z;
  const long long* zs = p.zone_sum + (size_t)b * (Z + 1);
  const double ref = (double)p.zone_ref[b];
  for (int zi = lane; zi < Z; zi += 32) {
    const int n = ncv[zi];
    const float m = n > 0 ? (float)(ref + from_fix(zs[zi]) / (double)n) : 0.f;
    zpost[zi] = m;
    zpre[zi] = p.pre_zone_mean[(size_t)b * Z + zi];
    p.zone_mean[(size_t)b * Z + zi] = m;
    if (!is_reset) {
      p.qcv[(size_t)b * Z + zi] = p.qcv_next[(size_t)b * Z + zi];
      p.qcv64[(size_t)b * Z + zi] = p.qcv64_next[(size_t)b * Z + zi];
    }
  }
  const float gmean = (float)(ref + from_fix(zs[Z]) / (double)((size_t)p.H * p.W));
  if (lane == 0) p.global_mean[b] = gmean;
  __syncwarp();
  Carry cy = {0, 0, 0, 0, 0};
  if (!is_reset) {
    const CarryStore c = carry[b];
    cy.ahu_flow = c.ahu_flow; cy.boiler_flow = c.boiler_flow; cy.return_water = c.return_water;
    cy.ahu_count = c.ahu_count; cy.boiler_count = c.boiler_count;
  }
  hvac_post(p, b, plan, lane, is_reset != 0, zpre, zpost, gmean, cy, scratch);
}

// Environment._reset: building.reset + hvac.reset (building.py:784-792,
// hvac_floorplan_based.py:124-128).  Thermostat modes survive (vav.py:98).
__global__ void k_reset_state(const Params p) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  p.ahu_heat_sp[b] = p.ahu_init_heat;
  p.ahu_cool_sp[b] = p.ahu_init_cool;
  p.boiler_sp[b] = p.boiler_init_sp;
  p.boiler_tank[(size_t)b * 3 + 0] = p.boiler_init_sp;
  p.boiler_tank[(size_t)b * 3 + 1] = 0.0;
  p.boiler_tank[(size_t)b * 3 + 2] = 0.0;
  p.cur[b] = 0;
  p.n_sweeps[b] = 0;
  p.max_delta[b] = 0.f;
  for (int zi = 0; zi < p.Z; ++zi) {
    p.qcv[(size_t)b * p.Z + zi] = 0.f;          // input_q = zeros
    p.qcv_next[(size_t)b * p.Z + zi] = 0.f;
    p.qcv64[(size_t)b * p.Z + zi] = 0.0;
    p.qcv64_next[(size_t)b * p.Z + zi] = 0.0;
    p.pre_zone_mean[(size_t)b * p.Z + zi] = 0.f;
  }
}

__global__ void k_reset_temp(const Params p) {
  const size_t n_cv = (size_t)p.H * p.W;
  const size_t total = n_cv * p.B;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / n_cv, j = i - b * n_cv;
    float v;
    if (p.n_reset == 0) v = p.initial_temp[b];
    else v = p.reset_temps[(p.n_reset == 1 ? 0 : b) * n_cv + j];
    p.tbuf[0][i] = v;
    if (p.temp64) p.temp64[i] = (double)v;
  }
}

// Streaming path: convection gather into the next buffer of the rotation, then
// k_rotate_cur makes it current.
__global__ void k_permute(const Params p) {
  const size_t n_cv = (size_t)p.H * p.W;
  const size_t total = n_cv * p.B;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / n_cv;
    const int c = p.cur[b];
    p.tbuf[(c + 1) % 3][i] = p.tbuf[c][b * n_cv + p.conv_perm[i]];
  }
}
__global__ void k_rotate_cur(const Params p) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < p.B) p.cur[b] = (uint8_t)((p.cur[b] + 1) % 3);
}

// keep the fp64 (Gauss-Seidel) and fp32 copies of building.temp coherent after an upload
__global__ void k_mirror_temp(const Params p, const int from_f32) {
  const size_t total = (size_t)p.B * p.H * p.W;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  if (from_f32) p.temp64[i] = (double)p.tbuf[0][i];
  else p.tbuf[0][i] = (float)p.temp64[i];
}
__global__ void k_mirror_q(const Params p) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)p.B * p.Z) p.qcv64[i] = (double)p.qcv[i];
}

// dst[b] = tbuf[cur[b]][b]  (download of building.temp in the streaming path)
__global__ void k_gather_temp(const Params p, float* dst) {
  const size_t n_cv = (size_t)p.H * p.W;
  const size_t total = n_cv * p.B;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / n_cv;
    dst[i] = p.tbuf[p.cur[b]][i];
  }
}

}  // namespace sbx
