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

#ifndef SBX_RESIDENT_THREADS
#define SBX_RESIDENT_THREADS 512
#endif
constexpr int kResidentThreads = SBX_RESIDENT_THREADS;
// register-resident HEAD vectors per thread (see HeadVec)
#ifndef SBX_RESIDENT_HEADS
#define SBX_RESIDENT_HEADS (SBX_RESIDENT_THREADS <= 256 ? 4 : 2)
#endif
constexpr int kHeads = SBX_RESIDENT_HEADS;
constexpr int kStreamThreads = 256;
// Rows per warp of the streaming kernels: chosen per grid height so that a CTA column of
// 8 warps covers the height in few, evenly filled tiles (long strips amortise the 2 halo
// rows and the per-CTA set-up: 74 % of the HBM roofline at 47 rows against 65 % at 8).
constexpr int kStreamMaxRowsPerWarp = 64;
#ifndef SBX_SWEEP_PREFETCH
#define SBX_SWEEP_PREFETCH 4
#endif
constexpr int kSweepPrefetchRows = SBX_SWEEP_PREFETCH;
__host__ __device__ inline int stream_rows_per_warp(int H) {
  const int warps = 256 / 32;
  const int n_tiles = (H + warps * kStreamMaxRowsPerWarp - 1) / (warps * kStreamMaxRowsPerWarp);
  const int rows = (H + warps * n_tiles - 1) / (warps * n_tiles);
  // k_sweep is instantiated for these strip lengths (compile-time trip count)
  return rows <= 8 ? 8 : rows <= 16 ? 16 : rows <= 32 ? 32 : rows <= 48 ? 48 : 64;
}
#ifndef SBX_SWEEP_MIN_CTAS
#define SBX_SWEEP_MIN_CTAS 4      // k_sweep: <= 64 registers, 32 warps per SM
#endif
#ifndef SBX_SWEEP_UNROLL
#define SBX_SWEEP_UNROLL 2
#endif
#define SBX_STR_(x) #x
#define SBX_STR(x) SBX_STR_(x)

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
// the V packed descriptors of a vector as loaded (V == 4: one 64-bit word), expanded at use
template <int V> struct DescRaw { uint2 v; };
template <> struct DescRaw<1> { uint32_t v; };
template <int V>
__device__ __forceinline__ DescRaw<V> load_d_raw(const uint16_t* p) {
  DescRaw<V> r;
  if constexpr (V == 4) r.v = *reinterpret_cast<const uint2*>(p);
  else r.v = *p;
  return r;
}
template <int V>
__device__ __forceinline__ void expand_d(const DescRaw<V>& r, uint32_t (&o)[V]) {
  if constexpr (V == 4) {
    o[0] = r.v.x & 0xFFFFu; o[1] = r.v.x >> 16; o[2] = r.v.y & 0xFFFFu; o[3] = r.v.y >> 16;
  } else {
    o[0] = r.v;
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
// Tiled (tensor-map) copies of one building's [H, W] temperature plane.  The box is
// [1, H, P] with P > W: the out-of-range columns are zero-filled on load and clipped
// on store, which gives the pitched shared-memory rows (ResidentGeom) in ONE
// transaction per plane (SASS UTMALDG / UTMASTG).
__device__ __forceinline__ void tma_load_plane(void* smem_dst, const void* tmap, int b, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(0), "r"(0), "r"(b)
      : "memory");
}
__device__ __forceinline__ void tma_store_plane(const void* tmap, int b, const void* smem_src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tmap),
               "r"(smem_u32(smem_src)), "r"(0), "r"(0), "r"(b)
               : "memory");
}
// generic-proxy writes (made visible by the preceding CTA barrier) -> async proxy
__device__ __forceinline__ void tma_store_fence() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
// warms L2 with data a LATER CTA will fetch (no shared-memory destination, no completion)
__device__ __forceinline__ void l2_prefetch(const void* gmem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---------------------------------------------------------------------------
// device-RNG stochastic convection: pattern of one (building, step); see k_convect_reduce
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint2 philox2x32_10(uint2 ctr, uint32_t key) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi = __umulhi(0xD256D193u, ctr.x), lo = 0xD256D193u * ctr.x;
    ctr = make_uint2(hi ^ key ^ ctr.y, lo);
    key += 0x9E3779B9u;
  }
  return ctr;
}
__device__ __forceinline__ uint32_t mix32(uint32_t x) {     // per-pair draw: 2 multiply-xorshift rounds
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return x;
}
struct ConvectPattern {
  int vertical, m, phi;       // pair (s, s + m) along rows / columns when ((s - phi) mod 2m) < m
  uint32_t pair_key;          // seed of the per-pair Bernoulli draws
};
__device__ __forceinline__ ConvectPattern convect_pattern(const Params& p, int b) {
  const uint2 r = philox2x32_10(make_uint2((uint32_t)b, (uint32_t)p.time_index),
                                (uint32_t)p.conv_seed ^ (uint32_t)(p.conv_seed >> 32));
  ConvectPattern c;
  const int n_off = p.conv_distance >= 4 ? 4 : 2;          // offsets of squared length 1 (and 4)
  const int o = (int)(r.x % (uint32_t)n_off);
  c.vertical = o & 1;
  c.m = o < 2 ? 1 : 2;
  c.phi = (int)((r.x >> 8) % (uint32_t)(2 * c.m));
  c.pair_key = r.y;
  return c;
}

// Source CV of (r, c) under the pattern: its partner when the pair swaps, else itself.
// `dsc` is the plan's RAW descriptor plane in global memory (zone byte per CV).
__device__ __forceinline__ int convect_source(const Params& p, const ConvectPattern& cp,
                                              const uint16_t* __restrict__ dsc, int r, int c) {
  const int self = r * p.W + c;
  const int z = desc_zone(dsc[self]);
  if (z == SBX_ZONE_NONE) return self;                     // only room CVs move
  const int sc = cp.vertical ? r : c;
  const int tmod = (sc + 2 * cp.m - cp.phi) & (2 * cp.m - 1);
  const int sp = tmod < cp.m ? sc + cp.m : sc - cp.m;
  if (sp < 0 || sp >= (cp.vertical ? p.H : p.W)) return self;
  const int other = cp.vertical ? sp * p.W + c : r * p.W + sp;
  if (desc_zone(dsc[other]) != z) return self;             // partner in another room / a wall
  if (p.conv_p < 1.0) {
    const uint32_t thr = (uint32_t)(p.conv_p * 4294967296.0);
    if (mix32((uint32_t)(self < other ? self : other) ^ cp.pair_key) >= thr) return self;
  }
  return other;
}

// ---------------------------------------------------------------------------
// resident path
// ---------------------------------------------------------------------------

// Per-building header the solve kernels consume, built by k_build_header and
// fetched with one TMA bulk copy: the 30-entry coefficient table for this
// building's (h, T_inf), the heat input per diffuser CV of each zone (slot Z = 0:
// "no heat input"), and {T_inf, n_fast}.
__host__ __device__ inline size_t header_q_slots(int Z) { return (size_t)((Z + 1 + 3) / 4) * 4; }
// + 8 scalars {T_inf, h, n_fast, n_fast+n_med, n_fast+n_med+n_ext, n_reduce_chunks, 0, 0}
// + 3 x {k/dx, cm, den, 1/den}: interior-class coefficients per material (MEDIUM list)
// + 9 x {kq_a, kq_b, -den_a, -den_b | 1/den_a, 1/den_b, cm_a, cm_b}: the same coefficients for
//   every ordered PAIR of materials (packed arithmetic of k_resident_step2's MEDIUM vectors)
// + 16 ints: the plan's list sizes (Params::counts2) and those of the plan of building
//   b + v2_stride, the one the same persistent CTA of k_resident_step2 loads next
constexpr int kPairTabBytes = 9 * 32;
__host__ __device__ inline size_t header_pair_offset(int Z) {
  return sizeof(Combo) * kNumCombos + header_q_slots(Z) * 4 + 32 + 48;
}
__host__ __device__ inline size_t header_bytes(int Z) {
  return header_pair_offset(Z) + kPairTabBytes + 64;
}

// Shared-memory geometry of k_resident_step, computed once on the host and passed
// in Params (constant bank -> uniform registers; nothing is recomputed per thread).
//
// Rows are stored with a pitch of P CVs, P/V ODD: vector (r, q) lives in slot
// r * Pq + q, so the 16-byte bank group of a vector, slot mod 8, rotates from row to
// row.  The vector lists skip whole columns (walls); with the natural pitch
// (W/4 = 24 for the 64x96 benchmark grids) the skipped columns would empty some
// residue classes, the 8 lanes of a quarter-warp could not all hit different bank
// groups, and 128-bit shared-memory accesses replayed 1.5-1.9x (measured with ncu);
// with the rotated pitch every residue class is equally populated and the
// accesses are conflict-free.
// Capacity (entries) of a plan's zone-sum list: one entry per (vector, zone) pair --
// a vector straddling a wall contributes two or three -- plus each zone segment
// padded to a whole warp.  Plans that need more fall back to a generic loop.
__host__ __device__ inline int reduce_list_capacity(int n_items, int Z) {
  return ((n_items + n_items / 2 + 31) & ~31) + 32 * (Z + 1);
}

__host__ inline ResidentGeom resident_geom(int H, int W, int Z, int V) {
  ResidentGeom g;
  const int wq = W / V;
  g.Pq = (V == 4 && (wq % 2) == 0) ? wq + 1 : wq;
  g.P = g.Pq * V;
  g.plane_cv = H * g.P;
  g.pq_magic = 0xFFFFFFFFu / (unsigned)g.Pq + 1u;
  g.rl_cap = reduce_list_capacity(H * wq, Z);
  // SBX_OPT_NUMPY_MEANS: the same region takes the plan's raster-ordered zone CV list (u16 per CV)
  if (V == 4 && H * W <= 65536 && g.rl_cap < (H * W / 2 + 7) / 8 * 8) g.rl_cap = (H * W / 2 + 7) / 8 * 8;
  g.desc_stride = (g.plane_cv + 7) & ~7;
  g.list_stride = (H * wq + 7) & ~7;
  // one tensor-map box per plane (box dims <= 256); otherwise one bulk copy per row
  g.use_tmap = (V == 4 && g.P <= 256 && H <= 256) ? 1 : 0;
  auto al = [](size_t v) { return (int)((v + 15) & ~(size_t)15); };
  auto al128 = [](size_t v) { return (int)((v + 127) & ~(size_t)127); };
  size_t o = 0;
  g.off_a = (int)o; o = al128(o + (size_t)g.plane_cv * 4);
  g.off_b = (int)o; o = al128(o + (size_t)g.plane_cv * 4);
  g.off_n3 = (int)o; o = al(o + (size_t)g.plane_cv * 4);
  g.off_desc = (int)o; o = al(o + (size_t)g.desc_stride * 2);
  g.off_list = (int)o; o = al(o + (size_t)g.list_stride * 2);
  g.off_rlist = (int)o; o = al(o + (size_t)g.rl_cap * 4);
  g.off_hdr = (int)o; o = al(o + header_bytes(Z));
  g.off_bins = (int)o; o = al(o + (size_t)(Z + 1) * 8 * (kResidentThreads / 32 + 1));
  g.off_wmax = (int)o; o = al(o + 32 * 4);
  g.off_bar = (int)o; o = al(o + 16);
  g.total = (int)o;
  return g;
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
template <int G>
__device__ __forceinline__ void build_header(const Params& p, int b, int lane, unsigned amask) {
  const int Z = p.Z;
  const int plan = p.n_plans == 1 ? 0 : b;
  unsigned char* h8 = p.hdr + (size_t)b * header_bytes(Z);
  Combo* tab = reinterpret_cast<Combo*>(h8);
  float* qcv = reinterpret_cast<float*>(h8 + sizeof(Combo) * kNumCombos);
  float* scal = qcv + header_q_slots(Z);
  const float t_inf = (float)env_ambient(p, b, p.time_index);   // tf_simulator.py:785
  const float h = (float)env_convection(p, b);
  build_combo_table(tab, p, plan, b, h, t_inf, lane, G);
  for (int zi = lane; zi < (int)header_q_slots(Z); zi += G)
    qcv[zi] = zi < Z ? p.qcv[(size_t)b * Z + zi] : 0.f;
  if (lane == 0) {
    // vector-list sizes of k_resident_step (absent on the streaming path and under k_resident_step2)
    const bool v1 = p.n_fast != nullptr;
    const int n_f = v1 ? p.n_fast[plan * 4 + 0] : 0, n_m = v1 ? p.n_fast[plan * 4 + 1] : 0,
              n_e = v1 ? p.n_fast[plan * 4 + 2] : 0;
    scal[0] = t_inf;
    scal[1] = h;
    scal[2] = __int_as_float(n_f);
    scal[3] = __int_as_float(n_f + n_m);
    scal[4] = __int_as_float(n_f + n_m + n_e);
    scal[5] = __int_as_float(v1 ? p.rl_chunks[plan] : 0);
    scal[6] = scal[7] = 0.f;
  }
  __syncwarp(amask);
  if (lane < kNumMaterials) {
    const Combo& c = tab[SBX_CV_INTERIOR * kNumMaterials + lane];
    reinterpret_cast<float4*>(scal + 8)[lane] = make_float4(c.k1, c.cm, c.den, c.rden);
  }
  float4* pair = reinterpret_cast<float4*>(h8 + header_pair_offset(Z));
  for (int i = lane; i < kNumMaterials * kNumMaterials; i += G) {
    const Combo& a = tab[SBX_CV_INTERIOR * kNumMaterials + i / kNumMaterials];
    const Combo& c = tab[SBX_CV_INTERIOR * kNumMaterials + i % kNumMaterials];
    pair[2 * i] = make_float4(a.k1, c.k1, -a.den, -c.den);
    pair[2 * i + 1] = make_float4(a.rden, c.rden, a.cm, c.cm);
  }
  if (p.counts2 != nullptr && lane < 4) {
    const int bn = b + p.v2_stride;
    const int pl = lane < 2 ? plan : (p.n_plans == 1 ? 0 : (bn < p.B ? bn : plan));
    reinterpret_cast<int4*>(h8 + header_pair_offset(Z) + kPairTabBytes)[lane] =
        reinterpret_cast<const int4*>(p.counts2 + (size_t)pl * 8)[lane & 1];
  }
}
__global__ void __launch_bounds__(128) k_build_header(const Params p) {
  const int wpb = blockDim.x / 32;
  const int b = p.b_begin + blockIdx.x * wpb + (threadIdx.x >> 5);
  if (b >= p.b_end) return;
  build_header<32>(p, b, threadIdx.x & 31, 0xffffffffu);
}

// Once per uploaded plan: repack the descriptors and sort the plan's vectors
// (V consecutive CVs) into four lists, stored back to back in one array:
//   FAST    every CV interior class, air, no diffuser: uniform register
//           coefficients, all four neighbours in range by construction;
//   MEDIUM  every CV interior class (walls, diffusers): per-material coefficients,
//           neighbours in range, no convection terms;
//   EXT     every CV exterior space: T = T_inf;
//   SLOW    everything else (boundary classes): generic, table-driven.
// Bit 15 of a MEDIUM / SLOW entry says the vector holds a diffuser CV.
//
// Entries are shared-memory vector SLOTS (row * Pq + column, see ResidentGeom).
// Order inside each list: (rank within the vector's slot-mod-8 class, slot mod 8),
// i.e. every aligned group of 8 entries holds 8 different residues mod 8, so the
// 8 lanes of a quarter-warp touch 8 different 16-byte bank groups and the 128-bit
// shared-memory accesses of the sweep stay conflict-free although the lists skip
// vectors.  One CTA per plan; ranks are computed by counting smaller keys (O(n^2)
// on a few thousand items, once per upload), so the order is deterministic.
constexpr int kPrepThreads = 256;
enum { kListFast = 0, kListMedium = 1, kListExt = 2, kListSlow = 3 };
template <int V>
__global__ void __launch_bounds__(kPrepThreads) k_prepare_plan(const Params p) {
  extern __shared__ __align__(16) unsigned char smem_prep[];
  const int plan = blockIdx.x, tid = threadIdx.x;
  const int n_cv = p.H * p.W, n_items = n_cv / V;
  uint32_t* key = reinterpret_cast<uint32_t*>(smem_prep);
  uint16_t* key1 = reinterpret_cast<uint16_t*>(key + n_items);
  __shared__ int s_count[4];
  const ResidentGeom& g = p.geom;
  const int wq = p.W / V;
  const uint16_t* raw = p.desc + (size_t)plan * n_cv;
  uint16_t* packed = p.desc_packed + (size_t)plan * g.desc_stride;
  uint16_t* qlist = p.qlist + (size_t)plan * g.list_stride;
  if (tid < 4) s_count[tid] = 0;
  for (int i = tid; i < g.desc_stride; i += kPrepThreads) {     // padded rows (pitch P)
    const int r = i / g.P, j = i - r * g.P;
    packed[i] = (r < p.H && j < p.W) ? (uint16_t)repack_desc(raw[r * p.W + j]) : (uint16_t)0;
  }
  for (int i = n_items + tid; i < g.list_stride; i += kPrepThreads) qlist[i] = 0;
  __syncthreads();
  // key1 = list | residue
  for (int it = tid; it < n_items; it += kPrepThreads) {
    bool fast = true, interior = true, ext = true;
    for (int e = 0; e < V; ++e) {
      const uint32_t d = raw[it * V + e];
      fast = fast && ((d & 0x007Fu) == kFastDesc);
      interior = interior && (desc_class(d) == SBX_CV_INTERIOR);
      ext = ext && (desc_class(d) == SBX_CV_EXTERIOR);
    }
    const int kind = fast ? kListFast : interior ? kListMedium : ext ? kListExt : kListSlow;
    const int slot = (it / wq) * g.Pq + it % wq;       // where the vector lives in shared memory
    key1[it] = (uint16_t)((kind << 3) | (slot & 7));
    atomicAdd(&s_count[kind], 1);
  }
  __syncthreads();
  // m = rank inside the (list, residue) bucket; key2 = list | m | residue
  for (int it = tid; it < n_items; it += kPrepThreads) {
    const uint32_t k1 = key1[it];
    int m = 0;
    for (int j = 0; j < it; ++j) m += (key1[j] == k1) ? 1 : 0;
    key[it] = ((k1 >> 3) << 18) | ((uint32_t)m << 3) | (k1 & 7u);
  }
  __syncthreads();
  // position = number of smaller keys (keys are unique)
  for (int it = tid; it < n_items; it += kPrepThreads) {
    const uint32_t k2 = key[it];
    int rank = 0;
    for (int j = 0; j < n_items; ++j) rank += (key[j] < k2) ? 1 : 0;
    bool diff = false;
    for (int e = 0; e < V; ++e) diff = diff || (raw[it * V + e] & SBX_DESC_DIFFUSER);
    qlist[rank] = (uint16_t)(((it / wq) * g.Pq + it % wq) | (diff ? 0x8000 : 0));
  }
  if (tid < 4) p.n_fast[plan * 4 + tid] = s_count[tid];
}

// Once per uploaded plan: the ZONE-SUM list.  Entry = vector index | element mask
// << 16 | zone slot << 20 (slot Z = "no zone", needed for the grid mean).  Entries
// are grouped by zone slot and every group is padded to a multiple of 32 with null
// entries (mask 0), so a warp's 32 entries always belong to ONE zone: the zone sums
// then take one integer REDUX per warp and no per-element zone logic at all.
// A vector that straddles a wall appears once per zone it touches, with
// complementary masks, so every CV is counted exactly once.
template <int V>
__global__ void __launch_bounds__(kPrepThreads) k_prepare_reduce(const Params p) {
  extern __shared__ __align__(16) unsigned char smem_prep[];
  const int plan = blockIdx.x, tid = threadIdx.x;
  const int n_cv = p.H * p.W, n_items = n_cv / V, Z = p.Z;
  uint2* parts = reinterpret_cast<uint2*>(smem_prep);   // 4 x u16 slot per vector (0xFFFF = unused)
  int* cnt = reinterpret_cast<int*>(parts + n_items);   // [Z+1]
  int* start = cnt + (Z + 1);                           // [Z+1]
  __shared__ int s_total;
  const uint16_t* raw = p.desc + (size_t)plan * n_cv;
  uint32_t* rl = p.rlist + (size_t)plan * p.geom.rl_cap;
  const int wq = p.W / V, Pq = p.geom.Pq;
  for (int i = tid; i <= Z; i += kPrepThreads) cnt[i] = 0;
  __syncthreads();
  for (int it = tid; it < n_items; it += kPrepThreads) {
    uint32_t z[4] = {0xFFFFu, 0xFFFFu, 0xFFFFu, 0xFFFFu};
    int np = 0;
    for (int e = 0; e < V; ++e) {
      const int zs = desc_zone(raw[it * V + e]);
      const uint32_t slot = zs == SBX_ZONE_NONE ? (uint32_t)Z : (uint32_t)zs;
      bool found = false;
      for (int k = 0; k < np; ++k) found = found || z[k] == slot;
      if (!found) {
        z[np++] = slot;
        atomicAdd(&cnt[slot], 1);
      }
    }
    parts[it] = make_uint2(z[0] | (z[1] << 16), z[2] | (z[3] << 16));
  }
  __syncthreads();
  if (tid == 0) {
    int o = 0;
    for (int i = 0; i <= Z; ++i) {
      start[i] = o;
      o += (cnt[i] + 31) & ~31;
    }
    s_total = o;
  }
  __syncthreads();
  const int total = s_total;
  if (total > p.geom.rl_cap) {       // pathological plan: k_resident_step takes its generic loop
    if (tid == 0) p.rl_chunks[plan] = -1;
    return;
  }
  for (int i = tid; i <= Z; i += kPrepThreads)
    for (int k = start[i] + cnt[i]; k < start[i] + ((cnt[i] + 31) & ~31); ++k) rl[k] = (uint32_t)i << 20;
  for (int it = tid; it < n_items; it += kPrepThreads) {
    const uint2 pz = parts[it];
    for (int k = 0; k < 4; ++k) {
      const uint32_t slot = ((k < 2 ? pz.x : pz.y) >> (16 * (k & 1))) & 0xFFFFu;
      if (slot == 0xFFFFu) break;
      // position inside the zone's group: vectors in index order
      const uint32_t pat = slot * 0x00010001u;
      int rank = 0;
      for (int j = 0; j < it; ++j) {
        const uint2 q = parts[j];
        rank += ((__vcmpeq2(q.x, pat) | __vcmpeq2(q.y, pat)) != 0u) ? 1 : 0;
      }
      uint32_t mask = 0;
      for (int e = 0; e < V; ++e) {
        const int zs = desc_zone(raw[it * V + e]);
        if ((zs == SBX_ZONE_NONE ? (uint32_t)Z : (uint32_t)zs) == slot) mask |= 1u << e;
      }
      rl[start[slot] + rank] = (uint32_t)((it / wq) * Pq + it % wq) | (mask << 16) | (slot << 20);
    }
  }
  if (tid == 0) p.rl_chunks[plan] = total / 32;
}

// One FAST vector (4 interior air CVs, no heat input) on packed fp32 pairs (e0,e1),
// (e2,e3): the same operations in the same order as cv_update_fast.  `n3a/n3b` are the
// thermal-mass terms of the two pairs.  Writes the new temperatures to o4 and returns the
// running max |dT|.
__device__ __forceinline__ float fast_core4(const float4 c4, const float4 up4, const float4 dn4,
                                            const float left, const float right,
                                            const f32x2 n3a, const f32x2 n3b,
                                            const FastCoef2& f, const float kq1, float lmax,
                                            float4& o4) {
  const f32x2 c01 = pack2(c4.x, c4.y), c23 = pack2(c4.z, c4.w);
  // horizontal: kq*T(j+1) + kq*T(j-1); every product is used by two CVs
  float m0, m1, m2, m3;
  unpack2(mul2(f.kq, c01), m0, m1);
  unpack2(mul2(f.kq, c23), m2, m3);
  const float ml = mul(kq1, left), mr = mul(kq1, right);
  f32x2 n1a = pack2(add(m1, ml), add(m2, m0));
  f32x2 n1b = pack2(add(m3, m1), add(mr, m2));
  n1a = mul2(f.vz, n1a);
  n1b = mul2(f.vz, n1b);
  // vertical: kq*T(i+1) + kq*T(i-1), and n1 + n2: sums of packed PRODUCTS, taken with pair_add
  // (a plain add.rn.f32x2 would be contracted into FFMA2 and lose a rounding)
  const f32x2 va = pair_add(mul2(f.kq, pack2(dn4.x, dn4.y)), f.one, mul2(f.kq, pack2(up4.x, up4.y)));
  const f32x2 vb = pair_add(mul2(f.kq, pack2(dn4.z, dn4.w)), f.one, mul2(f.kq, pack2(up4.z, up4.w)));
  const f32x2 n2a = mul2(f.vz, va);
  const f32x2 n2b = mul2(f.vz, vb);
  const f32x2 suma = pair_add(n1a, f.one, n2a);
  const f32x2 sumb = pair_add(n1b, f.one, n2b);
  const f32x2 oa = div_rn2(add2(suma, n3a), f.nden, f.rden);
  const f32x2 ob = div_rn2(add2(sumb, n3b), f.nden, f.rden);
  unpack2(oa, o4.x, o4.y);
  unpack2(ob, o4.z, o4.w);
  float d0, d1, d2, d3;
  unpack2(sub2(oa, c01), d0, d1);
  unpack2(sub2(ob, c23), d2, d3);
  lmax = fmaxf(fmaxf(lmax, fabsf(d0)), fabsf(d1));
  return fmaxf(fmaxf(lmax, fabsf(d2)), fabsf(d3));
}

// resident kernel: operands from the shared-memory planes; n3 computed and stored in the
// first sweep, loaded afterwards
template <bool FIRST>
__device__ __forceinline__ float fast_vector4(const float* __restrict__ in, float* __restrict__ out,
                                              float* __restrict__ n3p, int base, int W,
                                              const FastCoef2& fc2, float kq1, float lmax) {
  const float4 c4 = *reinterpret_cast<const float4*>(in + base);
  const float4 up4 = *reinterpret_cast<const float4*>(in + base - W);
  const float4 dn4 = *reinterpret_cast<const float4*>(in + base + W);
  const float left = in[base - 1];
  const float right = in[base + 4];
  f32x2 n3a, n3b;
  if constexpr (FIRST) {
    n3a = div_rn2(mul2(fc2.cm, pack2(c4.x, c4.y)), fc2.ndt, fc2.rdt);
    n3b = div_rn2(mul2(fc2.cm, pack2(c4.z, c4.w)), fc2.ndt, fc2.rdt);
    float4 st4;
    unpack2(n3a, st4.x, st4.y);
    unpack2(n3b, st4.z, st4.w);
    *reinterpret_cast<float4*>(n3p + base) = st4;
  } else {
    const float4 n4 = *reinterpret_cast<const float4*>(n3p + base);
    n3a = pack2(n4.x, n4.y);
    n3b = pack2(n4.z, n4.w);
  }
  float4 o4;
  lmax = fast_core4(c4, up4, dn4, left, right, n3a, n3b, fc2, kq1, lmax, o4);
  *reinterpret_cast<float4*>(out + base) = o4;
  return lmax;
}

// A thread's HEAD vectors (its first two list entries, FAST by construction when the list
// holds two rounds of FAST vectors) are the same in every sweep: their own temperatures and
// thermal-mass terms stay in REGISTERS from sweep to sweep -- no shared-memory load of the
// centre or of n3, no n3 store -- and only the new temperatures go to shared memory for
// the neighbours.  The shared-memory pipe is ~70 % busy in this kernel.
struct HeadVec {
  float4 c;          // T_est of the vector (this sweep's input)
  f32x2 n3a, n3b;    // (cm * T_prev) / dt of its two pairs
  int base;          // CV index of the vector in the planes
};
template <bool FIRST>
__device__ __forceinline__ float fast_head4(const float* __restrict__ in, float* __restrict__ out,
                                            HeadVec& hv, int W, const FastCoef2& fc2, float kq1) {
  const int base = hv.base;
  if constexpr (FIRST) {
    hv.c = *reinterpret_cast<const float4*>(in + base);
    hv.n3a = div_rn2(mul2(fc2.cm, pack2(hv.c.x, hv.c.y)), fc2.ndt, fc2.rdt);
    hv.n3b = div_rn2(mul2(fc2.cm, pack2(hv.c.z, hv.c.w)), fc2.ndt, fc2.rdt);
  }
  const float4 up4 = *reinterpret_cast<const float4*>(in + base - W);
  const float4 dn4 = *reinterpret_cast<const float4*>(in + base + W);
  const float left = in[base - 1];
  const float right = in[base + 4];
  float4 o4;
  const float lmax = fast_core4(hv.c, up4, dn4, left, right, hv.n3a, hv.n3b, fc2, kq1, 0.f, o4);
  *reinterpret_cast<float4*>(out + base) = o4;
  hv.c = o4;         // next sweep's input
  return lmax;
}

// Per-material coefficients of an interior-class CV (header, MEDIUM list)
struct MedCoef {
  float kq, cm, den, rden;
};

// One Jacobi sweep over the CTA's building.  FIRST: `in` is T_prev itself, so
// n3 = (cm * T_prev) / dt (tf_simulator.py:743-749) is computed here and stored
// for the following sweeps instead of being loaded.
// The four lists are walked as ONE sequence strided by the CTA size, so every warp
// gets the same number of vectors (+-1) whatever the mix; a warp diverges only
// where its 32 entries straddle a list boundary (three places per building).
struct SweepCtx {
  const uint16_t* dsc;
  const uint16_t* qlist;
  const Combo* tab;
  const float4* mtab;
  const float* qcv;
  FastCoef fc;
  FastCoef2 fc2;
  AreaCoef az;
  float cm_fast, dt, rdt, t_inf;
  int n_fast, n_fm, n_fme, n_items, H, P, Pq, wq, Z;
  unsigned pq_magic;
};
template <int V, bool FIRST>
__device__ __forceinline__ float resident_sweep(const float* __restrict__ in, float* __restrict__ out,
                                                float* __restrict__ n3p, const SweepCtx& s, int tid,
                                                HeadVec (&heads)[kHeads]) {
  constexpr int NT = kResidentThreads;
  const int W = s.P, wq = s.wq;      // W: row pitch in CVs
  const float t_inf = s.t_inf;
  float lmax = 0.f;
  int u = tid;
  if constexpr (V == 4) {
    // Every thread's first two vectors are FAST when the list holds two rounds of them:
    // run them as ONE instruction stream (two independent dependency chains), the
    // sweeps being latency-bound at 8 warps per scheduler.
    // as many head rounds as the FAST list fills completely (uniform over the CTA)
    const int nh = min(kHeads, s.n_fast / NT);
    float lh[kHeads];
#pragma unroll
    for (int j = 0; j < kHeads; ++j) {
      lh[j] = 0.f;
      if (j < nh) {
        if constexpr (FIRST) heads[j].base = ((int)s.qlist[u + j * NT] & 0x7FFF) * V;
        lh[j] = fast_head4<FIRST>(in, out, heads[j], W, s.fc2, s.fc.kq);
      }
    }
#pragma unroll
    for (int j = 0; j < kHeads; ++j) lmax = fmaxf(lmax, lh[j]);
    u += nh * NT;
  }
  for (; u < s.n_items; u += NT) {
    const int entry = (int)s.qlist[u];
    const int it = entry & 0x7FFF;     // vector slot
    const int base = it * V;
    float c[V], o[V];
    load_f<V>(in + base, c);
    if (u < s.n_fast) {
      if constexpr (V == 4) {
        lmax = fast_vector4<FIRST>(in, out, n3p, base, W, s.fc2, s.fc.kq, lmax);
        continue;
      } else {
      float up[V], dn[V], n3v[V];
      load_f<V>(in + base - W, up);
      load_f<V>(in + base + W, dn);
      if constexpr (FIRST) {
#pragma unroll
        for (int e = 0; e < V; ++e) n3v[e] = div_rn(mul(s.cm_fast, c[e]), s.dt, s.rdt);
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
        o[e] = cv_update_fast(s.fc, t_jp, t_jm, up[e], dn[e], n3v[e]);
      }
      }
    } else if (u < s.n_fm) {
      // interior class, any material, maybe a diffuser: k1 = k2 = k3 = k4 = k/dx,
      // hh = hv = 0 (x + 0 == x), uz = vz = z*dx; neighbours always in range
      const bool has_q = (entry & 0x8000) != 0;
      float up[V], dn[V], n3v[V];
      uint32_t d[V];
      MedCoef mc[V];
      load_d<V>(s.dsc + base, d);
      load_f<V>(in + base - W, up);
      load_f<V>(in + base + W, dn);
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const float4 m4 = s.mtab[(d[e] & kPackIdxMask) - SBX_CV_INTERIOR * kNumMaterials];
        mc[e].kq = m4.x; mc[e].cm = m4.y; mc[e].den = m4.z; mc[e].rden = m4.w;
      }
      if constexpr (FIRST) {
#pragma unroll
        for (int e = 0; e < V; ++e) n3v[e] = div_rn(mul(mc[e].cm, c[e]), s.dt, s.rdt);
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
        float qv = 0.f;
        if (has_q) qv = s.qcv[(d[e] & SBX_DESC_DIFFUSER) ? (int)(d[e] >> SBX_DESC_ZONE_SHIFT) : s.Z];
        float n1 = add(mul(mc[e].kq, t_jp), mul(mc[e].kq, t_jm));
        n1 = mul(s.az.full, n1);
        float n2 = add(mul(mc[e].kq, dn[e]), mul(mc[e].kq, up[e]));
        n2 = mul(s.az.full, n2);
        const float num = add(add(add(n1, n2), n3v[e]), qv);
        o[e] = div_rn(num, mc[e].den, mc[e].rden);
      }
    } else if (u < s.n_fme) {
      // exterior space: T = T_inf (tf_simulator.py:847-849); its n3 is never read
#pragma unroll
      for (int e = 0; e < V; ++e) o[e] = t_inf;
    } else {
      const bool has_q = (entry & 0x8000) != 0;
      const int r = (int)__umulhi((unsigned)it, s.pq_magic);
      const int q = it - r * s.Pq;
      float up[V], dn[V], n3v[V];
      uint32_t d[V];
      load_d<V>(s.dsc + base, d);
      if constexpr (FIRST) {
#pragma unroll
        for (int e = 0; e < V; ++e) n3v[e] = div_rn(mul(s.tab[d[e] & kPackIdxMask].cm, c[e]), s.dt, s.rdt);
        store_f<V>(n3p + base, n3v);
      } else {
        load_f<V>(n3p + base, n3v);
      }
      if (r > 0) load_f<V>(in + base - W, up); else fill<V>(up, t_inf);         // :642-644
      if (r < s.H - 1) load_f<V>(in + base + W, dn); else fill<V>(dn, t_inf);   // :646
      const float left = q > 0 ? in[base - 1] : t_inf;                          // :638-640
      const float right = q < wq - 1 ? in[base + V] : t_inf;                    // :636
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const float t_jm = e == 0 ? left : c[e - 1];
        const float t_jp = e == V - 1 ? right : c[e + 1];
        float qv = 0.f;
        if (has_q) qv = s.qcv[(d[e] & SBX_DESC_DIFFUSER) ? (int)(d[e] >> SBX_DESC_ZONE_SHIFT) : s.Z];
        o[e] = cv_update_packed(d[e], t_jp, t_jm, up[e], dn[e], n3v[e], qv, t_inf, s.az, s.tab);
      }
    }
#pragma unroll
    for (int e = 0; e < V; ++e) lmax = fmaxf(lmax, fabsf(__fsub_rn(o[e], c[e])));   // :851-853
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

// PW: SBX_OPT_NUMPY_MEANS with the tables in reach (Params::pw_fused) -- the zone / grid means
// are taken here, from the field in shared memory, in NumPy's summation order (sbx_pairwise.cuh)
// instead of the integer zone sums.
template <int V, bool PW = false>
__global__ void __launch_bounds__(kResidentThreads, 2) k_resident_step(const Params p, const __grid_constant__ CUtensorMap tmap_t) {
  extern __shared__ __align__(128) unsigned char smem[];
#ifdef SBX_PROFILE_PHASES
  long long phase_t0__ = clock64();
  const long long cta_t0__ = phase_t0__;
#endif
  const int b = p.b_begin + blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NT = kResidentThreads, NW = NT / 32;
  const int H = p.H, W = p.W, Z = p.Z;
  const int n_cv = H * W;
  const int n_items = n_cv / V;
  const int plan = p.n_plans == 1 ? 0 : b;
  const ResidentGeom& L = p.geom;
  const int P = L.P;                                   // row pitch (CVs) in shared memory
  float* bufA = reinterpret_cast<float*>(smem + L.off_a);
  float* bufB = reinterpret_cast<float*>(smem + L.off_b);
  float* n3p = reinterpret_cast<float*>(smem + L.off_n3);
  uint16_t* dsc = reinterpret_cast<uint16_t*>(smem + L.off_desc);
  uint16_t* qlist = reinterpret_cast<uint16_t*>(smem + L.off_list);
  uint32_t* rlist = reinterpret_cast<uint32_t*>(smem + L.off_rlist);
  Combo* tab = reinterpret_cast<Combo*>(smem + L.off_hdr);
  float* qcv = reinterpret_cast<float*>(smem + L.off_hdr + sizeof(Combo) * kNumCombos);
  const float* scal = qcv + header_q_slots(Z);
  long long* bins = reinterpret_cast<long long*>(smem + L.off_bins);
  float* wmax = reinterpret_cast<float*>(smem + L.off_wmax);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.off_bar);

  float* gT = p.tbuf[0] + (size_t)b * n_cv;
  const uint16_t* gD = p.desc_packed + (size_t)plan * L.desc_stride;
  const uint16_t* gL = p.qlist + (size_t)plan * L.list_stride;
  const uint32_t hdr_bytes = (uint32_t)header_bytes(Z);
  const unsigned char* gH = p.hdr + (size_t)b * hdr_bytes;
  constexpr bool use_tma = (V == 4);     // rows are 16-byte multiples

  // ---- stage 0: start the bulk loads (TMA, one mbarrier) ----------------------
  // The temperature rows go to their pitched positions: one bulk copy per row,
  // issued by the lanes of warp 0.
  if constexpr (use_tma) {
    if (warp == 0) {
      if (lane == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
        mbar_expect_tx(bar, (uint32_t)((L.use_tmap ? L.plane_cv : n_cv) * 4 + L.desc_stride * 2 +
                                       L.list_stride * 2) + hdr_bytes);
        tma_load_1d(smem + L.off_hdr, gH, hdr_bytes, bar);
        if (L.use_tmap) tma_load_plane(bufA, &tmap_t, b, bar);
        tma_load_1d(dsc, gD, (uint32_t)(L.desc_stride * 2), bar);
        tma_load_1d(qlist, gL, (uint32_t)(L.list_stride * 2), bar);
      }
      __syncwarp();
      if (!L.use_tmap)
        for (int r = lane; r < H; r += 32) tma_load_1d(bufA + r * P, gT + r * W, (uint32_t)(W * 4), bar);
    }
  } else {
    for (int i = tid; i < n_cv; i += NT) {      // V == 1: P == W
      bufA[i] = gT[i];
      dsc[i] = gD[i];
    }
    for (int i = tid; i < n_items; i += NT) qlist[i] = gL[i];
    for (int i = tid; i < (int)(hdr_bytes / 4); i += NT)
      reinterpret_cast<uint32_t*>(smem + L.off_hdr)[i] = reinterpret_cast<const uint32_t*>(gH)[i];
  }

  // The CTA that will take over this SM slot works on building b + (CTAs in flight):
  // pull its inputs into L2 now, so its own bulk loads do not wait on HBM (+2 % measured).
  if (p.prefetch_dist > 0 && tid == 32 && b + p.prefetch_dist < p.b_end) {
    const int bn = b + p.prefetch_dist;
    l2_prefetch(p.tbuf[0] + (size_t)bn * n_cv, (uint32_t)(n_cv * 4) & ~15u);
    l2_prefetch(p.hdr + (size_t)bn * hdr_bytes, hdr_bytes & ~15u);
    if (p.n_plans != 1) {
      l2_prefetch(p.desc_packed + (size_t)bn * L.desc_stride, (uint32_t)(L.desc_stride * 2));
      l2_prefetch(p.qlist + (size_t)bn * L.list_stride, (uint32_t)(L.list_stride * 2));
    }
  }

  // ---- stage 1: zero the zone bins while the copies are in flight ---------------
  if constexpr (PW) {
    // no integer bins in this variant: their room takes the plan's tree header (sizes, level
    // offsets) and roots now, so that the means phase does not start with a global round trip
    int* s_meta = reinterpret_cast<int*>(bins);
    if (tid < kPwMetaInts) s_meta[tid] = p.pw_meta[(size_t)plan * kPwMetaInts + tid];
    int2* s_root = reinterpret_cast<int2*>(s_meta + kPwMetaInts);
    for (int i = tid; i <= Z; i += NT) s_root[i] = p.pw_root[(size_t)plan * (Z + 1) + i];
  } else {
    for (int i = tid; i < (Z + 1) * (NW + 1); i += NT) bins[i] = 0;
  }
  __syncthreads();
  SBX_PHASE(0);   // launch .. TMA issued
  if constexpr (use_tma) mbar_wait(bar, 0);
  SBX_PHASE(1);   // waiting for the TMA loads
  const float t_inf = scal[0];
  // the zone-sum list is only needed after the sweeps: fetch it behind them
  const int n_chunks = __float_as_int(scal[5]);
  if constexpr (PW) {
    // the summation tables of this plan are needed after the sweeps: the raster-ordered CV list
    // comes into the (otherwise unused) zone-sum list region behind them, the leaf and node
    // tables into L2
    if (!p.fd_only) {
      const int* meta = reinterpret_cast<const int*>(bins);
      if (tid == 0) {
        mbar_expect_tx(bar + 1, p.pw_zbytes);
        tma_load_1d(rlist, reinterpret_cast<const uint16_t*>(p.pw_zlist) + (size_t)plan * n_cv, p.pw_zbytes, bar + 1);
      }
      const char* lf = reinterpret_cast<const char*>(p.pw_leaf + (size_t)plan * p.pw_capL);
      const char* nd = reinterpret_cast<const char*>(p.pw_node + (size_t)plan * p.pw_capI);
      for (int o = tid * 128; o < meta[0] * 8; o += NT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(lf + o));
      for (int o = tid * 128; o < meta[1] * 8; o += NT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nd + o));
    }
  }
  if (!PW && !p.fd_only && n_chunks > 0) {
    const uint32_t* gR = p.rlist + (size_t)plan * L.rl_cap;
    if constexpr (use_tma) {
      if (tid == 0) {
        mbar_expect_tx(bar + 1, (uint32_t)n_chunks * 128u);
        tma_load_1d(rlist, gR, (uint32_t)n_chunks * 128u, bar + 1);
      }
    } else {
      for (int i = tid; i < n_chunks * 32; i += NT) rlist[i] = gR[i];
    }
  }
  SweepCtx sc;
  sc.dsc = dsc; sc.qlist = qlist; sc.tab = tab; sc.qcv = qcv;
  sc.mtab = reinterpret_cast<const float4*>(scal + 8);
  sc.n_fast = __float_as_int(scal[2]);
  sc.n_fm = __float_as_int(scal[3]);
  sc.n_fme = __float_as_int(scal[4]);
  {
    const Combo& c = tab[SBX_CV_INTERIOR * kNumMaterials + 0];
    sc.fc.kq = c.k1; sc.fc.vz = c.vz; sc.fc.den = c.den; sc.fc.rden = c.rden;
    sc.cm_fast = c.cm;
  }
  sc.az.full = tab[SBX_CV_INTERIOR * kNumMaterials].vz;
  sc.az.half = tab[SBX_CV_CORNER_TL * kNumMaterials].vz;
  sc.dt = p.dt; sc.rdt = __frcp_rn(p.dt); sc.t_inf = t_inf;
  sc.fc2.kq = pack2(sc.fc.kq, sc.fc.kq); sc.fc2.vz = pack2(sc.fc.vz, sc.fc.vz);
  sc.fc2.nden = pack2(-sc.fc.den, -sc.fc.den); sc.fc2.rden = pack2(sc.fc.rden, sc.fc.rden);
  sc.fc2.cm = pack2(sc.cm_fast, sc.cm_fast);
  sc.fc2.ndt = pack2(-sc.dt, -sc.dt); sc.fc2.rdt = pack2(sc.rdt, sc.rdt);
  sc.fc2.one = pack2(p.one, p.one);
  sc.n_items = n_items; sc.H = H; sc.P = P; sc.Pq = L.Pq; sc.wq = W / V; sc.Z = Z;
  sc.pq_magic = L.pq_magic;

  // ---- stage 2: Jacobi sweeps to convergence (simulator.py:348-364) ---------
  float* in = bufA;
  float* out = bufB;
  HeadVec heads[kHeads];
#pragma unroll
  for (int j = 0; j < kHeads; ++j) {
    heads[j].c = make_float4(0.f, 0.f, 0.f, 0.f);
    heads[j].n3a = heads[j].n3b = 0ull;
    heads[j].base = 0;
  }
  int k = 0;
  float md = 0.f, last_lmax = 0.f;
  const int limit = p.iteration_limit;
  while (k < limit) {
    ++k;
    float lmax;
    if (k == 1) lmax = resident_sweep<V, true>(in, out, n3p, sc, tid, heads);
    else lmax = resident_sweep<V, false>(in, out, n3p, sc, tid, heads);
    // max|dT| <= threshold  <=>  no thread saw a delta above it (simulator.py:362);
    // the barrier doubles as the ping-pong hazard fence.
#ifdef SBX_PROFILE_PHASES
    const long long t_arrive__ = clock64();
#endif
    const int above = __syncthreads_or(lmax > p.threshold);
#ifdef SBX_PROFILE_PHASES
    if (b == p.B / 2 && lane == 0 && k <= 3 && !p.fd_only) {   // one probe CTA: who waits for whom
      p.phase_cycles[8 + warp * 8 + 2 * (k - 1)] = (unsigned long long)(t_arrive__ - cta_t0__);
      p.phase_cycles[8 + warp * 8 + 2 * (k - 1) + 1] = (unsigned long long)(clock64() - cta_t0__);
    }
#endif
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
    for (int i = tid; i < n_cv; i += NT) {
      const int src = perm[i];
      const int r = i / W, rs = src / W;
      out[r * P + (i - r * W)] = in[rs * P + (src - rs * W)];
    }
    __syncthreads();
    float* tmp = in; in = out; out = tmp;
  }

  if (p.conv_p > 0.0 && p.conv_perm == nullptr && !p.fd_only) {     // device-RNG mode
    const ConvectPattern cp = convect_pattern(p, b);
    const uint16_t* raw = p.desc + (size_t)plan * n_cv;
    for (int i = tid; i < n_cv; i += NT) {
      const int r = i / W, c = i - r * W;
      const int src = convect_source(p, cp, raw, r, c);
      const int rs = src / W;
      out[r * P + c] = in[rs * P + (src - rs * W)];
    }
    __syncthreads();
    float* tmp = in; in = out; out = tmp;
  }

  // ---- stage 3: write back + zone / grid sums ---------------------------------
  if constexpr (use_tma) {
    if (warp == 0) {
      tma_store_fence();
      if (L.use_tmap) {
        if (lane == 0) tma_store_plane(&tmap_t, b, in);
      } else {
        for (int r = lane; r < H; r += 32) tma_store_1d(gT + r * W, in + r * P, (uint32_t)(W * 4));
      }
      tma_store_commit();
    }
  } else {
    for (int i = tid; i < n_cv; i += NT) gT[i] = in[i];
  }
  // block max of the last sweep (diagnostic SBX_F_MAX_DELTA)
  last_lmax = warp_max(last_lmax);
  if (lane == 0) wmax[warp] = last_lmax;
  long long* wbins = bins + (size_t)(warp + 1) * (Z + 1);   // warp-private bins
  SBX_PHASE(4);   // convection gather (if any), store issued
  if constexpr (PW) {
    if (!p.fd_only) {
      // NumPy's pairwise tree over every room's raster-ordered CVs and over the grid (see
      // sbx_pairwise.cuh: same leaves, same butterfly, same levels), from the field in `in`;
      // leaf sums, inner nodes and the node table live in the idle plane `out`.
      const int* meta = reinterpret_cast<const int*>(bins);
      const int Lf = meta[0], In = meta[1], n_levels = meta[2];
      float* vals = out;
      uint2* s_node = reinterpret_cast<uint2*>(out + ((Lf + In + 1) & ~1));
      uint2* s_leaf = s_node + In;
      const uint2* gNode = p.pw_node + (size_t)plan * p.pw_capI;
      const uint2* gLeaf = p.pw_leaf + (size_t)plan * p.pw_capL;
      for (int i = tid; i < In; i += NT) s_node[i] = gNode[i];
      for (int i = tid; i < Lf; i += NT) s_leaf[i] = gLeaf[i];
      const int* s_lvl = meta + 3;
      if constexpr (use_tma) mbar_wait(bar + 1, 0);
      __syncthreads();
      SBX_PHASE(5);   // PW: tables staged
      const uint16_t* zl0 = reinterpret_cast<const uint16_t*>(rlist);
      const unsigned wmagic = 0xFFFFFFFFu / (unsigned)W + 1u;     // i / W for i < 2^16
      const int pad = P - W;
      for (int g0 = warp * 4; g0 < Lf; g0 += NT / 8) {            // 8 lanes per leaf, warp-uniform trip count
        const int g = g0 + (lane >> 3), j = lane & 7;
        const uint2 lf = g < Lf ? s_leaf[g] : make_uint2(0u, 0u);
        const int m = (int)lf.y;
        const int main_n = m >= 8 ? (m & ~7) : 0, tail_n = m - main_n;
        const uint32_t start = lf.x & ~kPwGridFlag;
        float v[16], tv = 0.f;
        if (lf.x & kPwGridFlag) {
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const uint32_t i = start + (uint32_t)(j + 8 * k);
            v[k] = j + 8 * k < main_n ? in[i + __umulhi(i, wmagic) * pad] : 0.f;
          }
          if (j < tail_n) {
            const uint32_t i = start + (uint32_t)(main_n + j);
            tv = in[i + __umulhi(i, wmagic) * pad];
          }
        } else {
          const uint16_t* zl = zl0 + start;
          uint32_t ix[16], it = 0u;
#pragma unroll
          for (int k = 0; k < 16; ++k) ix[k] = j + 8 * k < main_n ? (uint32_t)zl[j + 8 * k] : 0u;
          if (j < tail_n) it = (uint32_t)zl[main_n + j];
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] = j + 8 * k < main_n ? in[ix[k] + __umulhi(ix[k], wmagic) * pad] : 0.f;
          if (j < tail_n) tv = in[it + __umulhi(it, wmagic) * pad];
        }
        float res = v[0];
#pragma unroll
        for (int k = 1; k < 16; ++k)
          if (j + 8 * k < main_n) res = __fadd_rn(res, v[k]);
        res = __fadd_rn(res, __shfl_xor_sync(0xffffffffu, res, 1));
        res = __fadd_rn(res, __shfl_xor_sync(0xffffffffu, res, 2));
        res = __fadd_rn(res, __shfl_xor_sync(0xffffffffu, res, 4));
#pragma unroll
        for (int k = 0; k < 7; ++k) {
          const float x = __shfl_sync(0xffffffffu, tv, (lane & ~7) + k);
          if (k < tail_n) res = __fadd_rn(res, x);
        }
        if (g < Lf && j == 0) vals[g] = res;
      }
      __syncthreads();
      SBX_PHASE(6);   // PW: leaf sums
      // inner nodes height by height: the CTA while a level has more than a warp's worth of
      // nodes (a level never has more nodes than the one below), then warp 0 alone -- no block
      // barrier for the upper levels, the roots and the means
      int lv = 0;
      for (; lv < n_levels && s_lvl[lv + 1] - s_lvl[lv] > 32; ++lv) {
        const int i1 = s_lvl[lv + 1];
        for (int i = s_lvl[lv] + tid; i < i1; i += NT) {
          const uint2 nd = s_node[i];
          vals[Lf + i] = __fadd_rn(vals[nd.x], vals[nd.y]);
        }
        __syncthreads();
      }
      if (warp == 0) {
        for (; lv < n_levels; ++lv) {
          const int i = s_lvl[lv] + lane;
          if (i < s_lvl[lv + 1]) {
            const uint2 nd = s_node[i];
            vals[Lf + i] = __fadd_rn(vals[nd.x], vals[nd.y]);
          }
          __syncwarp();
        }
        const int2* root = reinterpret_cast<const int2*>(meta + kPwMetaInts);
        for (int z = lane; z <= Z; z += 32) {
          const int2 r = root[z];
          p.pw_mean[(size_t)b * (Z + 1) + z] = r.x >= 0 ? __fdiv_rn(vals[r.x], (float)r.y) : 0.f;
        }
      }
      SBX_PHASE(7);   // PW: levels + means
    }
  } else if (!p.fd_only && n_chunks > 0) {
    // Zone sums from the zone-grouped list: a warp's 32 entries belong to one zone,
    // so the warp reduces its lanes' integers with exact hardware REDUX and the
    // running sum of zone z lives in a REGISTER of lane z mod 32; shared memory is
    // touched once per warp at the end.  No atomics (a shared-memory 64-bit atomic is
    // a CAS loop on this hardware), no per-element zone decoding.
    if constexpr (use_tma) mbar_wait(bar + 1, 0);
    // Each warp takes a CONTIGUOUS run of chunks: consecutive chunks mostly belong to one
    // zone, so the lanes accumulate 32-bit partial sums over the run and the warp pays the REDUX
    // pair once per zone change instead of once per chunk.  (T - ref) * 2^16 by one FFMA:
    // fma(T, 2^16, -ref * 2^16) is exact, like to_fix32.
    const float nref = -__fmul_rn(t_inf, kFixScaleF);
    const int per = (n_chunks + NW - 1) / NW;
    const int c_lo = min(n_chunks, warp * per), c_hi = min(n_chunks, c_lo + per);
    int acc = 0, cur_z = -1;          // a run holds at most `per` * 4 CVs per lane: no overflow
    for (int chunk = c_lo; chunk < c_hi; ++chunk) {
      const uint32_t en = rlist[chunk * 32 + lane];
      float t[V];
      load_f<V>(in + (en & 0xFFFFu) * V, t);
      int sl = 0;
#pragma unroll
      for (int e = 0; e < V; ++e)
        if (en & (0x10000u << e)) sl += __float2int_rn(__fmaf_rn(t[e], kFixScaleF, nref));
      const int z0 = (int)(__shfl_sync(0xffffffffu, en, 0) >> 20);
      if (z0 != cur_z) {               // warp-uniform
        if (cur_z >= 0) {
          const long long sum = warp_sum_i32(acc);
          if (lane == 0) wbins[cur_z] += sum;
        }
        cur_z = z0;
        acc = 0;
      }
      acc += sl;
    }
    if (cur_z >= 0) {
      const long long sum = warp_sum_i32(acc);
      if (lane == 0) wbins[cur_z] += sum;
    }
  } else if (!p.fd_only) {
    // generic loop (plans whose zone-sum list does not fit): per-CV atomics
    for (int i = tid; i < n_cv; i += NT) {
      const int r = i / W, idx = r * P + (i - r * W);
      const int zs = desc_zone(dsc[idx]);
      fix_add(&bins[zs == SBX_ZONE_NONE ? Z : zs], (long long)to_fix32(in[idx], t_inf));
    }
  }
  SBX_PHASE(5);   // zone / grid sums of this warp
  __syncthreads();
  SBX_PHASE(6);   // waiting for the other warps' sums
  if (tid == 0) {
    md = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) md = fmaxf(md, wmax[w]);
    p.n_sweeps[b] = k;
    p.max_delta[b] = md;
    if (!p.fd_only) atomicAdd(p.sweeps_total, (unsigned long long)k);
  }
  if (!PW && !p.fd_only) {
    // combine the warp-private bins (integer sums: exact, order-free), handed to
    // k_post through zone_sum[b, 0..Z]
    long long* zs = p.zone_sum + (size_t)b * (Z + 1);
    if (tid == 0) p.zone_ref[b] = t_inf;
    if (warp == 0) {
      long long grid = 0;            // slot Z collected the CVs outside every zone
      for (int i = lane; i <= Z; i += 32) {
        long long acc = 0;
#pragma unroll
        for (int w = 0; w <= NW; ++w) acc += bins[(size_t)w * (Z + 1) + i];
        if (i < Z) zs[i] = acc;
        grid += acc;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) grid += __shfl_xor_sync(0xffffffffu, grid, o);
      if (lane == 0) zs[Z] = grid;
    }
  }
  if (use_tma && warp == 0) tma_store_wait();
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

// IN_GLOBAL: grids that do not fit one SM's shared memory (sim_config_legacy.gin runs this solver
// on the 744x1004 plan) -- the same wavefront, one CTA per building, on the fp64 field in global
// memory (L2-resident: 6 MB per plane for that plan), in place; the previous-step plane lives
// in Params::temp64_prev.  A barrier per anti-diagonal makes it a parity path, not a fast one.
constexpr int kGsGlobalThreads = 1024;
template <bool IN_GLOBAL>
__global__ void __launch_bounds__(IN_GLOBAL ? kGsGlobalThreads : kGsThreads) k_resident_gs(const Params p) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int kGsThreads = IN_GLOBAL ? kGsGlobalThreads : sbx::kGsThreads;     // stride of every loop below
  const int b = blockIdx.x, tid = threadIdx.x;
  const int H = p.H, W = p.W, Z = p.Z, n_cv = H * W;
  const int plan = p.n_plans == 1 ? 0 : b;
  const GsLayout L = gs_layout(IN_GLOBAL ? 0 : n_cv, Z);
  double* gT = p.temp64 + (size_t)b * n_cv;
  const uint16_t* gD = p.desc + (size_t)plan * n_cv;
  // (volatile in global memory: every access goes to L2, where the CTA's own earlier stores are)
  volatile double* T = IN_GLOBAL ? gT : reinterpret_cast<double*>(smem + L.off_t);
  volatile double* Tp = IN_GLOBAL ? p.temp64_prev + (size_t)b * n_cv : reinterpret_cast<double*>(smem + L.off_prev);
  double* qz = reinterpret_cast<double*>(smem + L.off_q);
  long long* bins = reinterpret_cast<long long*>(smem + L.off_bins);
  uint16_t* dsc_s = reinterpret_cast<uint16_t*>(smem + L.off_desc);
  const uint16_t* dsc = IN_GLOBAL ? gD : dsc_s;
  for (int i = tid; i < n_cv; i += kGsThreads) {
    const double v = gT[i];
    if constexpr (!IN_GLOBAL) {
      T[i] = v;
      dsc_s[i] = gD[i];
    }
    Tp[i] = v;
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
  // building.apply_convection() (simulator_flexible_floor_plan.py:156; the legacy config runs
  // Gauss-Seidel WITH StochasticConvectionSimulator, sim_config_legacy.gin:39-41,101): the
  // host-composed gather map, through the previous-temperature plane that is no longer needed
  if (p.conv_perm != nullptr && !p.fd_only) {
    const int32_t* perm = p.conv_perm + (size_t)b * n_cv;
    for (int i = tid; i < n_cv; i += kGsThreads) Tp[i] = T[perm[i]];
    __syncthreads();
    for (int i = tid; i < n_cv; i += kGsThreads) T[i] = Tp[i];
    __syncthreads();
  }
  if (p.conv_p > 0.0 && p.conv_perm == nullptr && !p.fd_only) {     // device-RNG mode
    const ConvectPattern cp = convect_pattern(p, b);
    for (int i = tid; i < n_cv; i += kGsThreads) Tp[i] = T[convect_source(p, cp, gD, i / W, i % W)];
    __syncthreads();
    for (int i = tid; i < n_cv; i += kGsThreads) T[i] = Tp[i];
    __syncthreads();
  }
  // write back (fp64 state + fp32 mirror), zone sums for k_post
  float* g32 = p.tbuf[0] + (size_t)b * n_cv;
  long long total = 0;
  for (int i = tid; i < n_cv; i += kGsThreads) {
    const double v = T[i];
    if constexpr (!IN_GLOBAL) gT[i] = v;
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
  int tiles_x, tiles_y, tiles, rows_per_warp;
};
__host__ __device__ inline StreamTiling stream_tiling(int H, int W, int V) {
  StreamTiling t;
  const int tile_w = 32 * V;
  t.rows_per_warp = stream_rows_per_warp(H);
  const int tile_h = (kStreamThreads / 32) * t.rows_per_warp;
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

// Once per uploaded plan (streaming path): descriptors repacked as combo index |
// half-U | diffuser | half-V | zone (see repack_desc), natural pitch W.
__global__ void k_pack_stream(const Params p) {
  const size_t total = (size_t)p.n_plans * p.H * p.W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x)
    p.desc_spk[i] = (uint16_t)repack_desc(p.desc[i]);
}

// One Jacobi sweep of every still-active building.  Sweep index k is 1-based.
// Every class of CV runs the same instruction stream (cv_update_packed): warps that
// mix interior / wall / boundary / exterior CVs do not diverge.
// MODE 0: a sweep k >= 2 (k from the host, or from the device in the graph loop); MODE 1: the
// first sweep (T_est is T_prev: one load serves both) -- a compile-time split: the sweeps of the
// loop carry neither the first sweep's branches nor its registers (7.7 instead of 8.1 ms per
// step's sweeps on 4096 x 744x1004).  (MODE 2, a first sweep that also took the zone sums so that
// buildings converging with it could skip k_zone_reduce, was measured: the ~55 extra instructions
// per vector cost the first sweep 2.4 ms and saved 2.9 ms of reduction -- 11.1 -> 10.9 ms; removed.)
template <int V, int R, int MODE>
__global__ void __launch_bounds__(kStreamThreads, SBX_SWEEP_MIN_CTAS) k_sweep(const Params p, const int k_arg) {
  __shared__ __align__(16) Combo tab[kNumCombos];
  __shared__ float qcv[kMaxZones + 1];
  const int k = MODE != 0 ? 1 : (k_arg > 0 ? k_arg : *p.sweep_k);   // device-driven loop: the sweep index lives on the device
  const StreamTiling tl = stream_tiling(p.H, p.W, V);
  const int b = blockIdx.x / tl.tiles;
  if (!p.active[b]) return;
  const int tile = blockIdx.x - b * tl.tiles;
  const int ty = tile / tl.tiles_x, tx = tile - ty * tl.tiles_x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = p.H, W = p.W, Z = p.Z;
  const size_t n_cv = (size_t)H * W;
  const int plan = p.n_plans == 1 ? 0 : b;
  // coefficient table, heat inputs and T_inf of this building: built once per step by
  // k_pre / k_build_header (the solve header of the resident kernel), not once per CTA
  const unsigned char* gH = p.hdr + (size_t)b * header_bytes(Z);
  for (int i = tid; i < (int)(sizeof(Combo) * kNumCombos / 16); i += kStreamThreads)
    reinterpret_cast<float4*>(tab)[i] = reinterpret_cast<const float4*>(gH)[i];
  const float* hq = reinterpret_cast<const float*>(gH + sizeof(Combo) * kNumCombos);
  for (int i = tid; i < Z; i += kStreamThreads) qcv[i] = hq[i];
  const float t_inf = hq[header_q_slots(Z)];
  __syncthreads();
  AreaCoef az;
  az.full = tab[SBX_CV_INTERIOR * kNumMaterials].vz;
  az.half = tab[SBX_CV_CORNER_TL * kNumMaterials].vz;
  const float dt = p.dt, rdt = __frcp_rn(p.dt);

  const int cur = p.cur[b];
  int bi, bo;
  sweep_buffers(cur, k, bi, bo);
  constexpr bool first = MODE != 0;     // T_est is T_prev itself: one load serves both
  const float* __restrict__ tin = p.tbuf[bi] + (size_t)b * n_cv;
  const float* __restrict__ tprev = p.tbuf[cur] + (size_t)b * n_cv;
  float* __restrict__ tout = p.tbuf[bo] + (size_t)b * n_cv;
  const uint16_t* __restrict__ dsc = p.desc_spk + (size_t)plan * n_cv;

  const int c0 = (tx * 32 + lane) * V;
  const int r0 = (ty * (kStreamThreads / 32) + warp) * R;      // R == tl.rows_per_warp
  const bool col_ok = c0 < W;
  float lmax = 0.f;
  if (r0 < H) {
    const int r1 = min(r0 + R, H);
    float up[V], c[V], dn[V];
    fill<V>(up, t_inf);
    fill<V>(c, t_inf);
    int off = r0 * W + c0;              // < 2^31: one building's grid
    // The row's descriptors are fetched ONE ROW AHEAD, like the row below of T_est (sweeps of
    // 4096 x 744x1004 as shipped: 17.9 -> 17.7 ms per step, the first sweep alone 5 % faster).
    // Doing the same for T_prev costs four more registers at the 64-register cap and spills:
    // 18.3 ms with both, 18.2 ms with T_prev only (SBX_SWEEP_PF_TP=1).
    [[maybe_unused]] float tpn[V];
    DescRaw<V> dnx = {};
    fill<V>(tpn, 0.f);
#ifndef SBX_SWEEP_PF_TP
#define SBX_SWEEP_PF_TP 0
#endif
#ifndef SBX_SWEEP_PF_D
#define SBX_SWEEP_PF_D 1
#endif
    constexpr bool pf_tp = !first && SBX_SWEEP_PF_TP, pf_d = first || SBX_SWEEP_PF_D;
    if (col_ok) {
      if (r0 > 0) load_f<V>(tin + off - W, up);
      load_f<V>(tin + off, c);
      if constexpr (pf_tp) load_f<V>(tprev + off, tpn);
      if constexpr (pf_d) dnx = load_d_raw<V>(dsc + off);
    }
_Pragma(SBX_STR(unroll SBX_SWEEP_UNROLL))
    for (int r = r0; r < r1; ++r, off += W) {
      fill<V>(dn, t_inf);
      if (col_ok && r + 1 < H) load_f<V>(tin + off + W, dn);
      // Rows further down the strip: pull them into L2 while this row computes.  The sweep
      // is latency-bound at 32 warps per SM (long-scoreboard stalls); the hint costs no
      // registers and took it from 70 % to 80 % of the HBM roofline (distance 2 / 4 / 8 rows:
      // 14.7 / 14.5 / 15.3 ms per step's sweeps on 4096 x 744x1004).
      if (col_ok && r + kSweepPrefetchRows < r1) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(tin + off + (kSweepPrefetchRows + 1) * W));
        if (!first) asm volatile("prefetch.global.L2 [%0];" ::"l"(tprev + off + kSweepPrefetchRows * W));
      }
      // horizontal neighbours: shuffle inside the warp, global load at its ends
      float left = __shfl_up_sync(0xffffffffu, c[V - 1], 1);
      float right = __shfl_down_sync(0xffffffffu, c[0], 1);
      if (col_ok) {
        if (lane == 0) left = c0 > 0 ? tin[off - 1] : t_inf;
        if (lane == 31 || c0 + V >= W) right = c0 + V < W ? tin[off + V] : t_inf;
        float tp[V], o[V];
        uint32_t d[V];
#pragma unroll
        for (int e = 0; e < V; ++e) tp[e] = first ? c[e] : tpn[e];
        if constexpr (!first && !pf_tp) load_f<V>(tprev + off, tp);
        if constexpr (!pf_d) dnx = load_d_raw<V>(dsc + off);
        expand_d<V>(dnx, d);
        if (r + 1 < r1) {
          if constexpr (pf_tp) load_f<V>(tprev + off + W, tpn);
          if constexpr (pf_d) dnx = load_d_raw<V>(dsc + off + W);
        }
        uint32_t any_q = 0;
#pragma unroll
        for (int e = 0; e < V; ++e) any_q |= d[e];
        any_q &= SBX_DESC_DIFFUSER;
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const float t_jm = e == 0 ? left : c[e - 1];
          const float t_jp = e == V - 1 ? right : c[e + 1];
          const float n3 = div_rn(mul(tab[d[e] & kPackIdxMask].cm, tp[e]), dt, rdt);   // :743-749
          o[e] = cv_numerator_packed(d[e], t_jp, t_jm, up[e], dn[e], n3, az, tab);
        }
        // + input_q (:754): only diffuser CVs have any, and x + 0 == x exactly
        if (any_q) {
#pragma unroll
          for (int e = 0; e < V; ++e)
            if (d[e] & SBX_DESC_DIFFUSER) o[e] = add(o[e], qcv[d[e] >> SBX_DESC_ZONE_SHIFT]);
        }
#pragma unroll
        for (int e = 0; e < V; ++e) {
          o[e] = cv_divide_packed(d[e], o[e], t_inf, tab);
          lmax = fmaxf(lmax, fabsf(__fsub_rn(o[e], c[e])));
        }
        store_f<V>(tout + off, o);
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

// Device-driven sweep loop (CUDA graph WHILE node, see build_sweep_graph in sbx_api.cu): ONE
// block walks every building after a sweep, does the bookkeeping of k_check and tells the
// graph whether another sweep is needed.  No host round trip between sweeps.
__global__ void __launch_bounds__(1024) k_check_loop(const Params p, const cudaGraphConditionalHandle handle) {
  const int k = *p.sweep_k;
  int still = 0;
  for (int b = threadIdx.x; b < p.B; b += blockDim.x) {
    if (!p.active[b]) continue;
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
  const int n_still = __syncthreads_count(still);
  if (threadIdx.x == 0) {
    *p.sweep_k = k + 1;
    atomicAdd(p.sweep_launches, 1ull);
    cudaGraphSetConditional(handle, n_still > 0 ? 1u : 0u);
  }
}

__global__ void k_loop_begin(const Params p) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b == 0) *p.sweep_k = 1;
  if (b < p.B) {
    p.active[b] = 1;
    p.max_delta_bits[b] = 0u;
    p.n_sweeps[b] = 0;
  }
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
// A thread owns a column group (V CVs x 8 rows).  Its CVs belong to at most two zones
// in all but pathological plans -- the zone of the group's first column and the zone of
// its last column, with a wall (no zone) in between -- so it keeps TWO running integer
// sums in registers, branch-free.  The warp then merges equal zones with exact hardware
// REDUX on 16-bit pieces and ONE lane per zone touches the CTA's shared bins: the kernel
// is a plain bandwidth-bound read of the field.  (64-bit shared-memory atomics are CAS
// loops on this hardware; one per run of CVs made this kernel 10x slower.)
__device__ __forceinline__ void warp_merge_zone_sums(int zc, long long run, long long* bins, int lane) {
  unsigned todo = __ballot_sync(0xffffffffu, zc != SBX_ZONE_NONE && run != 0);
  while (todo) {
    const int leader = __ffs(todo) - 1;
    const int z = __shfl_sync(0xffffffffu, zc, leader);
    const bool mine = ((todo >> lane) & 1u) && zc == z;
    const long long v = mine ? run : 0;
    const unsigned lo = __reduce_add_sync(0xffffffffu, (unsigned)(v & 0xFFFF));
    const unsigned mid = __reduce_add_sync(0xffffffffu, (unsigned)((v >> 16) & 0xFFFF));
    const int hi = __reduce_add_sync(0xffffffffu, (int)(v >> 32));
    if (lane == leader) fix_add(&bins[z], ((long long)hi << 32) + ((long long)mid << 16) + (long long)lo);
    todo &= ~__ballot_sync(0xffffffffu, mine);
  }
}

// latency-bound read of the whole field: 6 CTAs per SM (40 registers, a few spilled words)
// and an L2 prefetch 8 rows down the strip: 3.2 ms instead of 4.3 ms on 4096 x 744x1004
#ifndef SBX_ZR_MIN_CTAS
#define SBX_ZR_MIN_CTAS 6
#endif
#ifndef SBX_ZR_PREFETCH
#define SBX_ZR_PREFETCH 8
#endif
template <int V>
__global__ void __launch_bounds__(kStreamThreads, SBX_ZR_MIN_CTAS) k_zone_reduce(const Params p) {
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
  const int r0 = (ty * (kStreamThreads / 32) + warp) * tl.rows_per_warp;
  long long total = 0;                       // every CV: the grid mean
  long long runL = 0, runR = 0;              // CVs of zone zL / zR seen by this thread
  int zL = SBX_ZONE_NONE, zR = SBX_ZONE_NONE;
  if (c0 < W && r0 < H) {
    const int nr = min(tl.rows_per_warp, H - r0);
    const float* tp = t + (size_t)r0 * W + c0;
    const uint16_t* dp = dsc + (size_t)r0 * W + c0;
#pragma unroll 4
    for (int i = 0; i < nr; ++i, tp += W, dp += W) {
      float tv[V];
      uint32_t d[V];
      load_f<V>(tp, tv);
      load_d<V>(dp, d);
      if (i + SBX_ZR_PREFETCH < nr) asm volatile("prefetch.global.L2 [%0];" ::"l"(tp + (size_t)SBX_ZR_PREFETCH * W));
      const int z_first = desc_zone(d[0]), z_last = desc_zone(d[V - 1]);
      // a horizontal wall crossed between rows: hand the finished sums over
      if (z_first != zL && z_first != SBX_ZONE_NONE) {
        if (zL != SBX_ZONE_NONE && runL != 0) fix_add(&bins[zL], runL);
        zL = z_first;
        runL = 0;
      }
      if (z_last != zR && z_last != SBX_ZONE_NONE && z_last != zL) {
        if (zR != SBX_ZONE_NONE && runR != 0) fix_add(&bins[zR], runR);
        zR = z_last;
        runR = 0;
      }
      int sl = 0, sr = 0, sa = 0;            // int32: V CVs x T * 2^16 < 2^31 for T < 8191 K
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const int z = desc_zone(d[e]);
        const int v = __float2int_rn(__fmul_rn(tv[e], kFixScaleF));
        sa += v;
        if (z == zL) sl += v;
        else if (z == zR) sr += v;
        else if (z != SBX_ZONE_NONE) fix_add(&bins[z], (long long)v);   // third zone in V CVs
      }
      total += sa;
      runL += sl;
      runR += sr;
    }
  }
  // a slot whose zone is still NONE only ever collected CVs outside every zone
  warp_merge_zone_sums(zL, runL, bins, lane);
  warp_merge_zone_sums(zR, runR, bins, lane);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
  if (lane == 0) fix_add(&bins[Z], total);
  __syncthreads();
  long long* zs = p.zone_sum + (size_t)b * (Z + 1);
  if (tile == 0 && tid == 0) p.zone_ref[b] = 0.f;
  for (int i = tid; i <= Z; i += kStreamThreads)
    if (bins[i] != 0) fix_add(&zs[i], bins[i]);
}

// ---------------------------------------------------------------------------
// Stochastic convection, DEVICE-RNG mode (streaming path), fused with the zone reduction.
//
// The reference model (stochastic_convection_simulator.py:62-145) draws, per room CV and with
// probability p, a partner within squared distance `distance`, shuffles the swap list and
// applies the swaps one after the other: inherently sequential, and driven by Python's
// Mersenne Twister.  The exact replay (sbsim_b200/convection.py, SBX_F_CONVECTION_PERM) keeps
// parity; it is a host loop over every room CV.  This mode keeps the model's invariants --
// temperatures only move inside a room, every move is within the same squared distance, a
// CV takes part in a swap with probability p, the grid's and every zone's sums are unchanged
// -- with a counter-based generator and a swap pattern that needs no sequential pass: per
// (building, step) one axis-aligned offset (0,1) / (1,0) / (0,2) / (2,0) (those with squared
// length <= distance) and a phase are drawn (Philox-2x32-10); CVs are paired along that
// offset in disjoint pairs, and a pair swaps when both CVs belong to the same room and its
// own Bernoulli(p) draw says so.  Over steps the offsets alternate at random, so heat mixes
// in both directions like it does under the reference's random partner choice.
//
// One pass reads the field once, takes the zone / grid sums (unchanged by an in-room swap)
// and writes the permuted field into the next buffer of the rotation: the traffic of
// k_zone_reduce plus one write of the field.
// ---------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(kStreamThreads, 4) k_convect_reduce(const Params p) {
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
  const int cur = p.cur[b];
  const float* __restrict__ t = p.tbuf[cur] + (size_t)b * n_cv;
  float* __restrict__ tout = p.tbuf[(cur + 1) % 3] + (size_t)b * n_cv;
  const uint16_t* __restrict__ dsc = p.desc + (size_t)plan * n_cv;
  const ConvectPattern cp = convect_pattern(p, b);
  const uint32_t thr = p.conv_p >= 1.0 ? 0xFFFFFFFFu : (uint32_t)(p.conv_p * 4294967296.0);
  const bool always = p.conv_p >= 1.0;
  const int c0 = (tx * 32 + lane) * V;
  const int r0 = (ty * (kStreamThreads / 32) + warp) * tl.rows_per_warp;
  long long total = 0, runL = 0, runR = 0;
  int zL = SBX_ZONE_NONE, zR = SBX_ZONE_NONE;
  const bool col_ok = c0 < W;           // lanes right of the grid still take part in the shuffles
  if (r0 < H) {                         // warp-uniform
    const int nr = min(tl.rows_per_warp, H - r0);
    for (int i = 0; i < nr; ++i) {
      const int r = r0 + i;
      const size_t off = (size_t)r * W + c0;
      float tv[V], o[V];
      uint32_t d[V];
#pragma unroll
      for (int e = 0; e < V; ++e) { tv[e] = 0.f; d[e] = (uint32_t)SBX_ZONE_NONE << SBX_DESC_ZONE_SHIFT; }
      if (col_ok) {
        load_f<V>(t + off, tv);
        load_d<V>(dsc + off, d);
        if (i + SBX_ZR_PREFETCH < nr) asm volatile("prefetch.global.L2 [%0];" ::"l"(t + off + (size_t)SBX_ZR_PREFETCH * W));
      }
      // ---- zone / grid sums of the INPUT field (== those of the output) ----
      const int z_first = desc_zone(d[0]), z_last = desc_zone(d[V - 1]);
      if (z_first != zL && z_first != SBX_ZONE_NONE) {
        if (zL != SBX_ZONE_NONE && runL != 0) fix_add(&bins[zL], runL);
        zL = z_first;
        runL = 0;
      }
      if (z_last != zR && z_last != SBX_ZONE_NONE && z_last != zL) {
        if (zR != SBX_ZONE_NONE && runR != 0) fix_add(&bins[zR], runR);
        zR = z_last;
        runR = 0;
      }
      int sl = 0, sr = 0, sa = 0;
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const int z = desc_zone(d[e]);
        const int v = __float2int_rn(__fmul_rn(tv[e], kFixScaleF));
        sa += v;
        if (z == zL) sl += v;
        else if (z == zR) sr += v;
        else if (z != SBX_ZONE_NONE) fix_add(&bins[z], (long long)v);
      }
      total += sa;
      runL += sl;
      runR += sr;
      // ---- the swap pattern ----
      if constexpr (V == 4) {
        // Zones of the four CVs and of their partners as BYTES of one word: "same room" is one
        // byte-wise compare for the vector.  Partner = the same columns of row r +- m (vertical
        // pattern: one more vector load of T and of the descriptor) or the columns c +- m of this
        // row (horizontal: a window of 8 columns around the vector, the outer 2 + 2 from the lane
        // neighbours by shuffle -- the lanes at the warp's ends read them from memory -- and a
        // per-(m, phase) selection with static indices: the phase is uniform over the building).
        const uint32_t zw = __byte_perm(d[0] | (d[1] << 16), d[2] | (d[3] << 16), 0x7531);
        float pt0 = tv[0], pt1 = tv[1], pt2 = tv[2], pt3 = tv[3];
        uint32_t pzw = 0xFFFFFFFFu;                       // no partner: zone NONE everywhere
        if (cp.vertical) {
          const int tmod = (r + 2 * cp.m - cp.phi) & (2 * cp.m - 1);
          const int rp = tmod < cp.m ? r + cp.m : r - cp.m;
          if (rp >= 0 && rp < H && col_ok) {
            const float4 pv = *reinterpret_cast<const float4*>(t + (size_t)rp * W + c0);
            const uint2 dp = *reinterpret_cast<const uint2*>(dsc + (size_t)rp * W + c0);
            pt0 = pv.x; pt1 = pv.y; pt2 = pv.z; pt3 = pv.w;
            pzw = __byte_perm(dp.x, dp.y, 0x7531);
          }
        } else {
          // window columns c0-2 .. c0+5: w0 w1 | tv | w6 w7, zones zlo = (w0 w1 tv0 tv1), zhi = (tv2 tv3 w6 w7)
          float w0 = __shfl_up_sync(0xffffffffu, tv[2], 1), w1 = __shfl_up_sync(0xffffffffu, tv[3], 1);
          float w6 = __shfl_down_sync(0xffffffffu, tv[0], 1), w7 = __shfl_down_sync(0xffffffffu, tv[1], 1);
          uint32_t zl = __shfl_up_sync(0xffffffffu, zw, 1) >> 16, zr = __shfl_down_sync(0xffffffffu, zw, 1) & 0xFFFFu;
          if (col_ok && (lane == 0 || lane == 31 || c0 + V >= W)) {     // the window leaves the warp's columns
            if (lane == 0) {
              const bool in = c0 >= 2;
              w0 = in ? t[off - 2] : 0.f; w1 = in ? t[off - 1] : 0.f;
              zl = in ? ((uint32_t)desc_zone(dsc[off - 2]) | ((uint32_t)desc_zone(dsc[off - 1]) << 8)) : 0xFFFFu;
            }
            if (lane == 31 || c0 + V >= W) {
              const bool in0 = c0 + V < W, in1 = c0 + V + 1 < W;
              w6 = in0 ? t[off + V] : 0.f; w7 = in1 ? t[off + V + 1] : 0.f;
              zr = (in0 ? (uint32_t)desc_zone(dsc[off + V]) : 0xFFu) | ((in1 ? (uint32_t)desc_zone(dsc[off + V + 1]) : 0xFFu) << 8);
            }
          }
          const uint32_t zlo = zl | (zw << 16), zhi = (zw >> 16) | (zr << 16);
          // partner window index of element e: 2 + e + m where ((e - phi) mod 2m) < m, else 2 + e - m
          switch (cp.m * 4 + cp.phi) {
            case 4 + 0: pt0 = tv[1]; pt1 = tv[0]; pt2 = tv[3]; pt3 = tv[2]; pzw = __byte_perm(zlo, zhi, 0x4523); break;
            case 4 + 1: pt0 = w1; pt1 = tv[2]; pt2 = tv[1]; pt3 = w6; pzw = __byte_perm(zlo, zhi, 0x6341); break;
            case 8 + 0: pt0 = tv[2]; pt1 = tv[3]; pt2 = tv[0]; pt3 = tv[1]; pzw = __byte_perm(zlo, zhi, 0x3254); break;
            case 8 + 1: pt0 = w0; pt1 = tv[3]; pt2 = w6; pt3 = tv[1]; pzw = __byte_perm(zlo, zhi, 0x3650); break;
            case 8 + 2: pt0 = w0; pt1 = w1; pt2 = w6; pt3 = w7; pzw = __byte_perm(zlo, zhi, 0x7610); break;
            default:    pt0 = tv[2]; pt1 = w1; pt2 = tv[0]; pt3 = w7; pzw = __byte_perm(zlo, zhi, 0x7214); break;   // m = 2, phase 3
          }
        }
        // swap where the CV is in a room and its partner is in the same room
        uint32_t sw = __vcmpeq4(zw, pzw) & ~__vcmpeq4(zw, 0xFFFFFFFFu);
        if (!always && sw != 0u) {
#pragma unroll
          for (int e = 0; e < V; ++e) {
            const int c = c0 + e;
            const int sc = cp.vertical ? r : c;
            const int tmod = (sc + 2 * cp.m - cp.phi) & (2 * cp.m - 1);
            const int sp = tmod < cp.m ? sc + cp.m : sc - cp.m;
            const size_t self = (size_t)r * W + c;
            const size_t poff = cp.vertical ? (size_t)sp * W + c : (size_t)r * W + sp;
            const uint32_t key = (uint32_t)(self < poff ? self : poff);
            if (mix32(key ^ cp.pair_key) >= thr) sw &= ~(0xFFu << (8 * e));
          }
        }
        o[0] = (sw & 0x000000FFu) ? pt0 : tv[0];
        o[1] = (sw & 0x0000FF00u) ? pt1 : tv[1];
        o[2] = (sw & 0x00FF0000u) ? pt2 : tv[2];
        o[3] = (sw & 0xFF000000u) ? pt3 : tv[3];
      } else {
        // candidate partner values / zones of the V CVs: the same columns of row r +- m
        // (vertical pattern: one more vector load of T and of the descriptor), or the columns
        // c +- m of this row (horizontal: this thread's vector and its lane neighbours' by
        // shuffle; the two lanes at the warp's ends read the adjacent tile from memory)
        float pt[V];
        int pz[V];
  #pragma unroll
        for (int e = 0; e < V; ++e) { pt[e] = tv[e]; pz[e] = -1; }
        if (cp.vertical) {
          const int tmod = (r + 2 * cp.m - cp.phi) & (2 * cp.m - 1);
          const int rp = tmod < cp.m ? r + cp.m : r - cp.m;
          if (rp >= 0 && rp < H && col_ok) {
            uint32_t dp[V];
            load_f<V>(t + (size_t)rp * W + c0, pt);
            load_d<V>(dsc + (size_t)rp * W + c0, dp);
  #pragma unroll
            for (int e = 0; e < V; ++e) pz[e] = desc_zone(dp[e]);
          }
        } else {
          // window of V + 4 columns around the vector: [c0 - 2, c0 + V + 2)
          float wt[V + 4];
          int wz[V + 4];
  #pragma unroll
          for (int e = 0; e < V; ++e) { wt[2 + e] = tv[e]; wz[2 + e] = desc_zone(d[e]); }
          if constexpr (V == 4) {
            wt[0] = __shfl_up_sync(0xffffffffu, tv[2], 1); wt[1] = __shfl_up_sync(0xffffffffu, tv[3], 1);
            wt[6] = __shfl_down_sync(0xffffffffu, tv[0], 1); wt[7] = __shfl_down_sync(0xffffffffu, tv[1], 1);
            const int zlo = wz[4] | (wz[5] << 8), zhi = wz[2] | (wz[3] << 8);
            const int zl = __shfl_up_sync(0xffffffffu, zlo, 1), zr = __shfl_down_sync(0xffffffffu, zhi, 1);
            wz[0] = zl & 0xFF; wz[1] = zl >> 8; wz[6] = zr & 0xFF; wz[7] = zr >> 8;
            if (col_ok && (lane == 0 || lane == 31 || c0 + V >= W)) {     // the window leaves the warp's columns
  #pragma unroll
              for (int k = 0; k < 2; ++k) {
                const int cl = c0 - 2 + k, cr = c0 + V + k;
                if (lane == 0) {
                  wz[k] = cl >= 0 ? desc_zone(dsc[(size_t)r * W + cl]) : -1;
                  wt[k] = cl >= 0 ? t[(size_t)r * W + cl] : 0.f;
                }
                if (lane == 31 || c0 + V >= W) {
                  wz[6 + k] = cr < W ? desc_zone(dsc[(size_t)r * W + cr]) : -1;
                  wt[6 + k] = cr < W ? t[(size_t)r * W + cr] : 0.f;
                }
              }
            }
          } else if (col_ok) {
  #pragma unroll
            for (int k = 0; k < 2; ++k) {                     // V == 1: plain loads
              const int cl = c0 - 2 + k, cr = c0 + V + k;
              wz[k] = cl >= 0 ? desc_zone(dsc[(size_t)r * W + cl]) : -1;
              wt[k] = cl >= 0 ? t[(size_t)r * W + cl] : 0.f;
              wz[V + 2 + k] = cr < W ? desc_zone(dsc[(size_t)r * W + cr]) : -1;
              wt[V + 2 + k] = cr < W ? t[(size_t)r * W + cr] : 0.f;
            }
          }
  #pragma unroll
          for (int e = 0; e < V; ++e) {
            const int tmod = (c0 + e + 2 * cp.m - cp.phi) & (2 * cp.m - 1);
            const int sh = tmod < cp.m ? cp.m : -cp.m;       // +-1 or +-2
            // select from the window without dynamic indexing
            const float a1 = sh > 0 ? wt[2 + e + 1] : wt[2 + e - 1];
            const float a2 = sh > 0 ? wt[2 + e + 2] : wt[2 + e - 2];
            const int b1 = sh > 0 ? wz[2 + e + 1] : wz[2 + e - 1];
            const int b2 = sh > 0 ? wz[2 + e + 2] : wz[2 + e - 2];
            pt[e] = cp.m == 1 ? a1 : a2;
            pz[e] = cp.m == 1 ? b1 : b2;
          }
        }
  #pragma unroll
        for (int e = 0; e < V; ++e) {
          o[e] = tv[e];
          const int z = desc_zone(d[e]);
          if (z == SBX_ZONE_NONE || pz[e] != z) continue;      // only room CVs move, inside their room
          if (!always) {
            const int c = c0 + e;
            const int sc = cp.vertical ? r : c;
            const int tmod = (sc + 2 * cp.m - cp.phi) & (2 * cp.m - 1);
            const int sp = tmod < cp.m ? sc + cp.m : sc - cp.m;
            const size_t self = (size_t)r * W + c;
            const size_t poff = cp.vertical ? (size_t)sp * W + c : (size_t)r * W + sp;
            const uint32_t key = (uint32_t)(self < poff ? self : poff);
            if (mix32(key ^ cp.pair_key) >= thr) continue;
          }
          o[e] = pt[e];
        }
      }
      if (col_ok) store_f<V>(tout + off, o);
    }
  }
  warp_merge_zone_sums(zL, runL, bins, lane);
  warp_merge_zone_sums(zR, runR, bins, lane);
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

// HVAC prologue; G lanes per building (see sbx_device.cuh), 128 threads per CTA.
#ifndef SBX_HVAC_MIN_CTAS
#define SBX_HVAC_MIN_CTAS 1
#endif
template <int G>
__global__ void __launch_bounds__(128, SBX_HVAC_MIN_CTAS) k_pre(const Params p, CarryStore* carry) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int group = threadIdx.x / G, lane = threadIdx.x % G;
  const int b = p.b_begin + blockIdx.x * (blockDim.x / G) + group;
  const unsigned amask = __ballot_sync(0xffffffffu, b < p.b_end);
  if (b >= p.b_end) return;
  const int Z = p.Z;
  double* scratch = reinterpret_cast<double*>(smem) + (size_t)group * (3 * Z + 64 + Z);
  float* zpre = reinterpret_cast<float*>(scratch + 3 * Z + 64);
  const int plan = p.n_plans == 1 ? 0 : b;
  if (p.build_hdr) build_header<G>(p, b, lane, amask);   // reads only last step's state
  for (int zi = lane; zi < Z; zi += G) zpre[zi] = p.zone_mean[(size_t)b * Z + zi];
  __syncwarp(amask);
  PreOut o = hvac_pre<G>(p, b, plan, lane, amask, zpre, p.global_mean[b], scratch);
  if (lane == 0) {
    CarryStore c;
    c.ahu_flow = o.ahu_flow; c.boiler_flow = o.boiler_flow; c.return_water = o.return_water;
    c.ahu_count = o.ahu_count; c.boiler_count = o.boiler_count;
    carry[b] = c;
  }
}

// Means from zone sums, then observation + reward; G lanes per building.
template <int G>
__global__ void __launch_bounds__(128, SBX_HVAC_MIN_CTAS) k_post(const Params p, const CarryStore* carry,
                                              const int is_reset) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int group = threadIdx.x / G, lane = threadIdx.x % G;
  const int b = p.b_begin + blockIdx.x * (blockDim.x / G) + group;
  const unsigned amask = __ballot_sync(0xffffffffu, b < p.b_end);
  if (b >= p.b_end) return;
  const int Z = p.Z;
  double* scratch = reinterpret_cast<double*>(smem) + (size_t)group * (3 * Z + 64 + Z);
  float* zpre = reinterpret_cast<float*>(scratch + 3 * Z + 64);
  float* zpost = zpre + Z;
  const int plan = p.n_plans == 1 ? 0 : b;
  const int32_t* ncv = p.zone_ncv + (size_t)plan * Z;
  const long long* zs = p.zone_sum + (size_t)b * (Z + 1);
  const double ref = (double)p.zone_ref[b];
  for (int zi = lane; zi < Z; zi += G) {
    const int n = ncv[zi];
    const float m = p.pw_on ? p.pw_mean[(size_t)b * (Z + 1) + zi]
                            : n > 0 ? (float)(ref + from_fix(zs[zi]) / (double)n) : 0.f;
    zpost[zi] = m;
    zpre[zi] = p.pre_zone_mean[(size_t)b * Z + zi];
    p.zone_mean[(size_t)b * Z + zi] = m;
    if (!is_reset) {
      p.qcv[(size_t)b * Z + zi] = p.qcv_next[(size_t)b * Z + zi];
      p.qcv64[(size_t)b * Z + zi] = p.qcv64_next[(size_t)b * Z + zi];
    }
  }
  const float gmean = p.pw_on ? p.pw_mean[(size_t)b * (Z + 1) + Z]
                              : (float)(ref + from_fix(zs[Z]) / (double)((size_t)p.H * p.W));
  if (lane == 0) p.global_mean[b] = gmean;
  __syncwarp(amask);
  Carry cy = {0, 0, 0, 0, 0};
  if (!is_reset) {
    const CarryStore c = carry[b];
    cy.ahu_flow = c.ahu_flow; cy.boiler_flow = c.boiler_flow; cy.return_water = c.return_water;
    cy.ahu_count = c.ahu_count; cy.boiler_count = c.boiler_count;
  }
  hvac_post<G>(p, b, plan, lane, amask, is_reset != 0, zpre, zpost, gmean, cy, scratch);
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
