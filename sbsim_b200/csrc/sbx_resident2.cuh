// k_resident_step2: the resident diffusion solve, second generation (grids whose width is
// a multiple of 4 and that fit one tensor-map box; everything else runs k_resident_step).
//
// Same arithmetic, same shared-memory planes and rotated pitch as k_resident_step; what
// changed is everything around the arithmetic (profiles/r01_ncu_full_randomized.txt showed
// only ~25 % of the executed instructions in the FAST update itself):
//
//  * PERSISTENT CTAs: grid = SMs x resident CTAs, each CTA walks buildings b, b + grid, ...
//    Pointer set-up, barrier initialisation and coefficient constants are paid once per
//    CTA; the next building's planes / lists / header are fetched by TMA into the two
//    planes the finished solve no longer needs WHILE the zone sums and the TMA store of the
//    current building run (three planes rotate roles: input, output, thermal-mass term).
//  * The per-CV descriptor plane is gone.  Vectors that need more than the uniform FAST
//    update carry a 16-byte RECORD built once per plan: slot, neighbour-in-range flags, four
//    combo bytes, four heat-slot bytes, two material-pair offsets.  No descriptor decode, no
//    row / column division, no per-CV predicated heat lookup in the sweeps; 12.8 kB less
//    shared memory and ~7 kB less HBM traffic per building.
//  * MEDIUM vectors (interior class, any material) run on packed fp32 pairs with per-pair
//    coefficients from a 9-entry material-pair table; vectors with and without heat input
//    are separate lists, so the heat lookup is a plain load where it exists at all.
//  * Pure exterior vectors are only written in sweeps 1 and 2 (afterwards both planes hold
//    T_inf there and their delta is exactly 0).
//  * Zone sums: single-zone vectors (almost all) have their own list section of bare u16
//    slots, fixed-point conversion by one FFMA + F2I per CV, per-lane accumulation over a
//    warp's contiguous run of chunks and one REDUX flush per zone change.
//
// Bit-exactness contract unchanged (tests/test_gpu_parity.py): every fp32 operation of
// tf_simulator.py:719-754 is one IEEE rounding, in the reference's order.
#pragma once

#include "sbx_kernels.cuh"

namespace sbx {

// CTA shape: kR2Compute threads sweep, ONE extra warp (the producer) issues every TMA load /
// store and writes the per-building results, so that no compute warp ever carries serial
// work into the next sweep barrier (with thread 0 of a 512-thread CTA doing it, the other 15
// warps waited ~6000 cycles per building at the first barrier of the next building).
#ifndef SBX_R2_PERSISTENT
#define SBX_R2_PERSISTENT 0
#endif
#ifndef SBX_R2_COMPUTE_THREADS
#define SBX_R2_COMPUTE_THREADS (SBX_R2_PERSISTENT ? 384 : 512)
#endif
#ifndef SBX_R2_HEADS
#define SBX_R2_HEADS 2
#endif
#ifndef SBX_R2_PAIR
#define SBX_R2_PAIR 0
#endif
constexpr int kR2Compute = SBX_R2_COMPUTE_THREADS;
constexpr int kR2Threads = kR2Compute + (SBX_R2_PERSISTENT ? 32 : 0);
constexpr int kR2Heads = SBX_R2_HEADS;
constexpr int kR2Kinds = 6;
enum { kR2Fast = 0, kR2Ext = 1, kR2MedNoQ = 2, kR2MedQ = 3, kR2SlowNoQ = 4, kR2SlowQ = 5 };
// counts2[plan*8 + i]
enum { kC2Fast = 0, kC2Ext = 1, kC2Heads = 2, kC2SchedWords = 3, kC2Spare = 4, kC2Rec = 5,
       kC2FullChunks = 6, kC2PartChunks = 7 };
// Per-warp WORK SCHEDULE of the sweeps, built once per plan.  The lists are cut into chunks
// of up to 32 entries of one kind; chunks are dealt to the compute warps by longest-
// processing-time-first on estimated costs, so that every warp reaches the sweep barrier
// with about the same number of instructions behind it (with a plain CTA-strided walk the
// warps that drew a chunk of boundary vectors arrived ~2x later than the others, and a
// sweep lasts as long as its slowest warp).  Word layout: [0 .. NW] offsets of each warp's
// descriptors, then descriptors = kind | (count - 1) << 3 | first entry << 8.
__host__ __device__ constexpr int r2_sched_cost(int kind) {   // FAST, EXT, MED, MED+q, SLOW, SLOW+q
  return kind == 0 ? 10 : kind == 1 ? 1 : kind == 2 ? 16 : kind == 3 ? 18 : kind == 4 ? 26 : 30;
}
__host__ __device__ inline int r2_sched_cap(int n_items) {
  return ((kR2Compute / 32 + 1) + n_items / 32 + 8 + 3) & ~3;
}
// record flags (rec.x >> 16)
constexpr uint32_t kRecUp = 1u, kRecDown = 2u, kRecLeft = 4u, kRecRight = 8u;

__host__ inline Resident2Geom resident2_geom(const ResidentGeom& g, int H, int W, int Z) {
  Resident2Geom r;
  const int n_items = H * (W / 4);
  r.rec_cap = (n_items / 2 + 7) & ~7;
  r.zfull_cap = ((n_items + 31) & ~31) + 32 * (Z + 1);
  r.zpart_cap = ((n_items / 2 + 31) & ~31) + 32 * (Z + 1);
  r.zchunk_cap = ((r.zfull_cap + r.zpart_cap) / 32 + 15) & ~15;
  auto al = [](size_t v) { return (int)((v + 15) & ~(size_t)15); };
  auto al128 = [](size_t v) { return (int)((v + 127) & ~(size_t)127); };
  size_t o = 0;
  r.off_p0 = (int)o; o = al128(o + (size_t)g.plane_cv * 4);
  r.off_p1 = (int)o; o = al128(o + (size_t)g.plane_cv * 4);
  r.off_p2 = (int)o; o = al128(o + (size_t)g.plane_cv * 4);
  r.off_flist = (int)o; o = al(o + (size_t)g.list_stride * 2);
  r.off_sched = (int)o; o = al(o + (size_t)r2_sched_cap(n_items) * 4);
  r.off_rec = (int)o; o = al(o + (size_t)r.rec_cap * 16);
  r.off_zfull = (int)o; o = al(o + (size_t)r.zfull_cap * 2);
  r.off_zpart = (int)o; o = al(o + (size_t)r.zpart_cap * 4);
  r.off_zchunk = (int)o; o = al(o + (size_t)r.zchunk_cap);
  r.off_hdr = (int)o; o = al(o + header_bytes(Z));
  r.off_bins = (int)o; o = al(o + (size_t)(Z + 1) * 8 * (kR2Compute / 32));   // one row per compute warp
  r.off_misc = (int)o; o = al(o + 32);
  r.off_bar = (int)o; o = al(o + 64);
  r.total = (int)o;
  return r;
}

// ---------------------------------------------------------------------------
// Once per uploaded plan.  One CTA per plan; positions by counting smaller keys
// (deterministic, O(n^2) on a few thousand vectors).
// ---------------------------------------------------------------------------
__host__ __device__ inline size_t prepare2_smem(int n_items, int Z) {
  return (size_t)n_items * 4 + (((size_t)n_items * 2 + 15) & ~(size_t)15) + (size_t)n_items * 8 +
         (size_t)(Z + 1) * 4 * 4 + (((size_t)n_items + 15) & ~(size_t)15) + 256;
}

__global__ void __launch_bounds__(kPrepThreads) k_prepare_plan2(const Params p) {
  extern __shared__ __align__(16) unsigned char smem_prep[];
  const int plan = blockIdx.x, tid = threadIdx.x;
  const int H = p.H, W = p.W, Z = p.Z, wq = W / 4, n_items = H * wq, n_cv = H * W;
  const ResidentGeom& g = p.geom;
  const Resident2Geom& g2 = p.g2;
  uint32_t* key = reinterpret_cast<uint32_t*>(smem_prep);
  uint16_t* key1 = reinterpret_cast<uint16_t*>(key + n_items);
  uint2* parts = reinterpret_cast<uint2*>(smem_prep + (size_t)n_items * 4 + (((size_t)n_items * 2 + 15) & ~(size_t)15));
  int* cnt_full = reinterpret_cast<int*>(parts + n_items);
  int* cnt_part = cnt_full + (Z + 1);
  int* start_full = cnt_part + (Z + 1);
  int* start_part = start_full + (Z + 1);
  uint8_t* single = reinterpret_cast<uint8_t*>(start_part + (Z + 1));   // 1: one zone, all four CVs counted
  uint8_t* warp_of = single + ((n_items + 15) & ~15);                   // schedule: chunk -> warp (<= 256 chunks)
  __shared__ int s_count[kR2Kinds], s_base[kR2Kinds + 1], s_tot[2];
  const uint16_t* raw = p.desc + (size_t)plan * n_cv;
  uint16_t* flist = p.flist2 + (size_t)plan * g.list_stride;
  uint4* rec = p.rec2 + (size_t)plan * g2.rec_cap;
  uint16_t* zfull = p.zfull + (size_t)plan * g2.zfull_cap;
  uint32_t* zpart = p.zpart + (size_t)plan * g2.zpart_cap;
  uint8_t* zchunk = p.zchunk + (size_t)plan * g2.zchunk_cap;
  int32_t* counts = p.counts2 + (size_t)plan * 8;
  uint32_t* sched = p.sched2 + (size_t)plan * r2_sched_cap(n_items);
  if (tid < kR2Kinds) s_count[tid] = 0;
  for (int i = tid; i <= Z; i += kPrepThreads) cnt_full[i] = cnt_part[i] = 0;
  for (int i = tid; i < g.list_stride; i += kPrepThreads) flist[i] = 0;
  __syncthreads();
  auto slot_of = [&](int it) { return (it / wq) * g.Pq + it % wq; };
  // zone-sum slot of a CV: its zone, Z for CVs outside every zone (walls: the grid mean needs
  // them), 0xFFFF for exterior space -- T == T_inf == the sums' reference there, so it
  // contributes exactly 0 and is left out of the lists
  auto zone_slot = [&](uint32_t d) {
    if (desc_class(d) == SBX_CV_EXTERIOR) return 0xFFFFu;
    const int zs = desc_zone(d);
    return zs == SBX_ZONE_NONE ? (uint32_t)Z : (uint32_t)zs;
  };
  // ---- kinds, (kind, residue) buckets, zone parts ----
  for (int it = tid; it < n_items; it += kPrepThreads) {
    bool fast = true, interior = true, ext = true, anyq = false, any_ext = false;
    uint32_t z[4] = {0xFFFFu, 0xFFFFu, 0xFFFFu, 0xFFFFu};
    int np = 0;
    for (int e = 0; e < 4; ++e) {
      const uint32_t d = raw[it * 4 + e];
      fast = fast && ((d & 0x007Fu) == kFastDesc);
      interior = interior && (desc_class(d) == SBX_CV_INTERIOR);
      ext = ext && (desc_class(d) == SBX_CV_EXTERIOR);
      anyq = anyq || (d & SBX_DESC_DIFFUSER);
      const uint32_t zs = zone_slot(d);
      bool found = zs == 0xFFFFu;
      any_ext = any_ext || zs == 0xFFFFu;
      for (int k = 0; k < np; ++k) found = found || z[k] == zs;
      if (!found) z[np++] = zs;
    }
    const int kind = fast ? kR2Fast : ext ? kR2Ext : interior ? (anyq ? kR2MedQ : kR2MedNoQ)
                                                             : (anyq ? kR2SlowQ : kR2SlowNoQ);
    // Records are additionally grouped by PATTERN class (which coefficient-table rows the
    // vector reads): a warp's 32 records then mostly read the same rows, which the shared-
    // memory pipe serves as broadcasts instead of bank-conflicting gathers.
    uint32_t cls = 0;
    if (kind >= kR2MedNoQ) {
      uint32_t hsh = 0;
      for (int e = 0; e < 4; ++e) hsh = hsh * 31u + (repack_desc(raw[it * 4 + e]) & 0xBFu);
      cls = (hsh ^ (hsh >> 7) ^ (hsh >> 13)) & 15u;
    }
    key1[it] = (uint16_t)((kind << 7) | (cls << 3) | (slot_of(it) & 7));
    atomicAdd(&s_count[kind], 1);
    parts[it] = make_uint2(z[0] | (z[1] << 16), z[2] | (z[3] << 16));
    const bool one = np == 1 && !any_ext;
    single[it] = one ? 1 : 0;
    if (one) atomicAdd(&cnt_full[z[0]], 1);
    else for (int k = 0; k < np; ++k) atomicAdd(&cnt_part[z[k]], 1);
  }
  __syncthreads();
  if (tid == 0) {
    int o = 0;
    for (int k = 0; k < kR2Kinds; ++k) { s_base[k] = o; o += s_count[k]; }
    s_base[kR2Kinds] = o;
    int of = 0, op = 0;
    for (int i = 0; i <= Z; ++i) {
      start_full[i] = of; of += (cnt_full[i] + 31) & ~31;
      start_part[i] = op; op += (cnt_part[i] + 31) & ~31;
    }
    s_tot[0] = of; s_tot[1] = op;
  }
  // rank inside the (kind, residue) bucket
  for (int it = tid; it < n_items; it += kPrepThreads) {
    const uint32_t k1 = key1[it];
    int m = 0;
    for (int j = 0; j < it; ++j) m += (key1[j] == k1) ? 1 : 0;
    key[it] = ((k1 >> 3) << 18) | ((uint32_t)m << 3) | (k1 & 7u);     // kind | class | rank | residue
  }
  __syncthreads();
  const int n_rec = s_count[kR2MedNoQ] + s_count[kR2MedQ] + s_count[kR2SlowNoQ] + s_count[kR2SlowQ];
  const int tot_full = s_tot[0], tot_part = s_tot[1];
  int n_sched_chunks = 0;
  for (int k = 0; k < kR2Kinds; ++k) n_sched_chunks += (s_count[k] + 31) / 32;
  const bool overflow = n_rec > g2.rec_cap || tot_full > g2.zfull_cap || tot_part > g2.zpart_cap ||
                        (tot_full + tot_part) / 32 > g2.zchunk_cap || Z > 254 || n_sched_chunks > 256 ||
                        (kR2Compute / 32 + 1) + n_sched_chunks > r2_sched_cap(n_items);
  if (tid == 0 && !overflow) {
    // ---- warp schedule: longest-processing-time-first over the chunks ----
    constexpr int NWc = kR2Compute / 32;
    int load[NWc];
    for (int w = 0; w < NWc; ++w) load[w] = 0;
    // chunk ids: kinds in list order FAST, EXT, MED, MED+q, SLOW, SLOW+q
    int cbase[kR2Kinds + 1];
    cbase[0] = 0;
    for (int k = 0; k < kR2Kinds; ++k) cbase[k + 1] = cbase[k] + (s_count[k] + 31) / 32;
    for (int c = 0; c < n_sched_chunks; ++c) warp_of[c] = 0xFF;
    // register-resident head vectors: the first kR2Heads chunks of every warp are full FAST chunks
    const int n_heads = (s_count[kR2Fast] / 32 >= kR2Heads * NWc) ? kR2Heads : 0;
    for (int w = 0; w < NWc; ++w)
      for (int j = 0; j < n_heads; ++j) {
        warp_of[w * n_heads + j] = (uint8_t)w;
        load[w] += r2_sched_cost(kR2Fast);
      }
    const int order[kR2Kinds] = {kR2SlowQ, kR2SlowNoQ, kR2MedQ, kR2MedNoQ, kR2Fast, kR2Ext};
    for (int oi = 0; oi < kR2Kinds; ++oi) {
      const int kind = order[oi];
      for (int c = cbase[kind]; c < cbase[kind + 1]; ++c) {
        if (warp_of[c] != 0xFF) continue;
        int best = 0;
        for (int w = 1; w < NWc; ++w) if (load[w] < load[best]) best = w;
        warp_of[c] = (uint8_t)best;
        load[best] += r2_sched_cost(kind);
      }
    }
    // entry index of a kind's first element: FAST / EXT index the vector list, the others the records
    int ebase[kR2Kinds];
    ebase[kR2Fast] = 0;
    ebase[kR2Ext] = s_count[kR2Fast];
    ebase[kR2MedNoQ] = 0;
    ebase[kR2MedQ] = s_count[kR2MedNoQ];
    ebase[kR2SlowNoQ] = ebase[kR2MedQ] + s_count[kR2MedQ];
    ebase[kR2SlowQ] = ebase[kR2SlowNoQ] + s_count[kR2SlowNoQ];
    // emission per warp: heads, other FAST chunks, records, EXT
    const int emit_order[kR2Kinds] = {kR2Fast, kR2SlowQ, kR2SlowNoQ, kR2MedQ, kR2MedNoQ, kR2Ext};
    int o = NWc + 1;
    for (int w = 0; w < NWc; ++w) {
      sched[w] = (uint32_t)o;
      for (int oi = 0; oi < kR2Kinds; ++oi) {
        const int kind = emit_order[oi];
        for (int c = cbase[kind]; c < cbase[kind + 1]; ++c) {
          if (warp_of[c] != w) continue;
          const int j = c - cbase[kind];
          const int n = min(32, s_count[kind] - 32 * j);
          sched[o++] = (uint32_t)kind | ((uint32_t)(n - 1) << 3) | ((uint32_t)(ebase[kind] + 32 * j) << 8);
        }
      }
    }
    sched[NWc] = (uint32_t)o;
    counts[kC2Heads] = n_heads;
    counts[kC2SchedWords] = (o + 3) & ~3;
  }
  if (tid == 0) {
    counts[kC2Fast] = s_count[kR2Fast];
    counts[kC2Ext] = s_count[kR2Ext];
    counts[kC2Spare] = 0;
    counts[kC2Rec] = n_rec;
    counts[kC2FullChunks] = tot_full / 32;
    counts[kC2PartChunks] = overflow ? -1 : tot_part / 32;
  }
  if (overflow) return;        // the handle falls back to k_resident_step
  // ---- vector lists and records ----
  for (int it = tid; it < n_items; it += kPrepThreads) {
    const uint32_t k2 = key[it];
    int rank = 0;
    for (int j = 0; j < n_items; ++j) rank += (key[j] < k2) ? 1 : 0;
    const int kind = (int)(k2 >> 22);
    const int slot = slot_of(it);
    if (kind == kR2Fast || kind == kR2Ext) {
      flist[rank] = (uint16_t)slot;          // EXT follows FAST: s_base[kR2Ext] == n_fast
    } else {
      const int r = it / wq, q = it - r * wq;
      uint32_t ci = 0, qi = 0, m[4];
      for (int e = 0; e < 4; ++e) {
        const uint32_t d = raw[it * 4 + e];
        ci |= (repack_desc(d) & 0xBFu) << (8 * e);                    // combo | half-U | half-V
        qi |= ((d & SBX_DESC_DIFFUSER) ? (uint32_t)desc_zone(d) : (uint32_t)Z) << (8 * e);
        m[e] = (uint32_t)desc_material(d);
      }
      const uint32_t flags = (r > 0 ? kRecUp : 0u) | (r < H - 1 ? kRecDown : 0u) |
                             (q > 0 ? kRecLeft : 0u) | (q < wq - 1 ? kRecRight : 0u);
      const uint32_t pa = (m[0] * 3 + m[1]) * 32, pb = (m[2] * 3 + m[3]) * 32;
      rec[rank - s_base[kR2MedNoQ]] = make_uint4((uint32_t)slot | (flags << 16), ci, qi, pa | (pb << 16));
    }
  }
  // ---- zone-sum list: padding, chunk zones, entries ----
  for (int i = tid; i <= Z; i += kPrepThreads) {
    for (int k = start_full[i] + cnt_full[i]; k < start_full[i] + ((cnt_full[i] + 31) & ~31); ++k) zfull[k] = 0xFFFFu;
    for (int k = start_part[i] + cnt_part[i]; k < start_part[i] + ((cnt_part[i] + 31) & ~31); ++k) zpart[k] = 0u;
    for (int c = start_full[i] / 32; c < (start_full[i] + ((cnt_full[i] + 31) & ~31)) / 32; ++c) zchunk[c] = (uint8_t)i;
    for (int c = start_part[i] / 32; c < (start_part[i] + ((cnt_part[i] + 31) & ~31)) / 32; ++c)
      zchunk[tot_full / 32 + c] = (uint8_t)i;
  }
  __syncthreads();             // every thread is done ranking with key[] before it is reused below
  for (int it = tid; it < n_items; it += kPrepThreads) {
    const uint2 pz = parts[it];
    if (single[it]) {
      // order inside the zone's group: (rank within the slot-mod-8 class, slot mod 8), like the
      // sweep lists, so that the 128-bit loads of a quarter-warp hit 8 different bank groups
      const uint32_t zs = pz.x & 0xFFFFu;
      const int res = slot_of(it) & 7;
      int m = 0;
      for (int j = 0; j < it; ++j)
        m += (single[j] && (parts[j].x & 0xFFFFu) == zs && (slot_of(j) & 7) == res) ? 1 : 0;
      key[it] = ((uint32_t)m << 3) | (uint32_t)res;      // key[] is free again (lists are written)
    } else {
      for (int k = 0; k < 4; ++k) {
        const uint32_t zs = ((k < 2 ? pz.x : pz.y) >> (16 * (k & 1))) & 0xFFFFu;
        if (zs == 0xFFFFu) break;
        const uint32_t pat = zs * 0x00010001u;
        int rank = 0;
        for (int j = 0; j < it; ++j) {
          const uint2 q = parts[j];
          rank += (!single[j] && (__vcmpeq2(q.x, pat) | __vcmpeq2(q.y, pat)) != 0u) ? 1 : 0;
        }
        uint32_t mask = 0;
        for (int e = 0; e < 4; ++e)
          if (zone_slot(raw[it * 4 + e]) == zs) mask |= 1u << e;
        zpart[start_part[zs] + rank] = (uint32_t)slot_of(it) | (mask << 16);
      }
    }
  }
  __syncthreads();
  for (int it = tid; it < n_items; it += kPrepThreads) {
    if (!single[it]) continue;
    const uint32_t zs = parts[it].x & 0xFFFFu, k3 = key[it];
    int rank = 0;
    for (int j = 0; j < n_items; ++j)
      rank += (single[j] && (parts[j].x & 0xFFFFu) == zs && key[j] < k3) ? 1 : 0;
    zfull[start_full[zs] + rank] = (uint16_t)slot_of(it);
  }
}

// ---------------------------------------------------------------------------
// sweeps
// ---------------------------------------------------------------------------
// The kernel runs at 64 registers with two register-resident head vectors per thread, so
// coefficients are (re-)read from the header in shared memory where they are used instead of
// living in registers over a building's sweeps.
struct Sweep2Ctx {
  unsigned char* smem;       // CTA's dynamic shared memory (offsets: Params::g2, constant bank)
  float t_inf;
};
#define R2_FLIST(s, p) (reinterpret_cast<const uint16_t*>((s).smem + (p).g2.off_flist))
#define R2_REC(s, p) (reinterpret_cast<const uint4*>((s).smem + (p).g2.off_rec))
#define R2_TAB(s, p) (reinterpret_cast<const Combo*>((s).smem + (p).g2.off_hdr))
#define R2_QCV(s, p) (reinterpret_cast<const float*>((s).smem + (p).g2.off_hdr + sizeof(Combo) * kNumCombos))
#define R2_PAIR(s, p) ((s).smem + (p).g2.off_hdr + header_pair_offset((p).Z))
#define R2_CNT(s, p) (reinterpret_cast<const int*>((s).smem + (p).g2.off_hdr + header_pair_offset((p).Z) + kPairTabBytes))

// MEDIUM vector: four interior-class CVs of any material (tf_simulator.py:719-754 with
// k1 = k2 = k3 = k4 = k/dx of the CV's OWN material, no convection terms), packed pairs.
template <bool FIRST, bool HASQ>
__device__ __forceinline__ float med_vector4(const float* __restrict__ in, float* __restrict__ out,
                                             float* __restrict__ n3p, const uint4 rc,
                                             const Sweep2Ctx& s, const Params& p, float lmax) {
  const int base = (int)(rc.x & 0xFFFFu) * 4, P = p.geom.P;
  const float4 c4 = *reinterpret_cast<const float4*>(in + base);
  const float4 up4 = *reinterpret_cast<const float4*>(in + base - P);
  const float4 dn4 = *reinterpret_cast<const float4*>(in + base + P);
  const float left = in[base - 1];
  const float right = in[base + 4];
  const unsigned char* pairtab = R2_PAIR(s, p);
  const float4 A0 = *reinterpret_cast<const float4*>(pairtab + (rc.w & 0xFFFFu));
  const float4 A1 = *reinterpret_cast<const float4*>(pairtab + (rc.w & 0xFFFFu) + 16);
  const float4 B0 = *reinterpret_cast<const float4*>(pairtab + (rc.w >> 16));
  const float4 B1 = *reinterpret_cast<const float4*>(pairtab + (rc.w >> 16) + 16);
  const f32x2 kqa = pack2(A0.x, A0.y), ndena = pack2(A0.z, A0.w), rdena = pack2(A1.x, A1.y);
  const f32x2 kqb = pack2(B0.x, B0.y), ndenb = pack2(B0.z, B0.w), rdenb = pack2(B1.x, B1.y);
  const f32x2 c01 = pack2(c4.x, c4.y), c23 = pack2(c4.z, c4.w);
  const float vz1 = R2_TAB(s, p)[SBX_CV_INTERIOR * kNumMaterials].vz;      // z * dx (:791-792)
  const f32x2 vz2 = pack2(vz1, vz1);
  f32x2 n3a, n3b;
  if constexpr (FIRST) {
    const float rdt = __frcp_rn(p.dt);
    const f32x2 ndt2 = pack2(-p.dt, -p.dt), rdt2 = pack2(rdt, rdt);
    n3a = div_rn2(mul2(pack2(A1.z, A1.w), c01), ndt2, rdt2);                // :743-749
    n3b = div_rn2(mul2(pack2(B1.z, B1.w), c23), ndt2, rdt2);
    float4 st4;
    unpack2(n3a, st4.x, st4.y);
    unpack2(n3b, st4.z, st4.w);
    *reinterpret_cast<float4*>(n3p + base) = st4;
  } else {
    const float4 n4 = *reinterpret_cast<const float4*>(n3p + base);
    n3a = pack2(n4.x, n4.y);
    n3b = pack2(n4.z, n4.w);
  }
  // horizontal: kq*T(j+1) + kq*T(j-1); sums of packed products are scalar adds (see fast_core4)
  float p0, p1, p2, p3, m0, m1, m2, m3;
  unpack2(mul2(kqa, pack2(c4.y, c4.z)), p0, p1);
  unpack2(mul2(kqa, pack2(left, c4.x)), m0, m1);
  unpack2(mul2(kqb, pack2(c4.w, right)), p2, p3);
  unpack2(mul2(kqb, pack2(c4.y, c4.z)), m2, m3);
  const f32x2 n1a = mul2(vz2, pack2(add(p0, m0), add(p1, m1)));
  const f32x2 n1b = mul2(vz2, pack2(add(p2, m2), add(p3, m3)));
  // vertical: kq*T(i+1) + kq*T(i-1)
  float a0, a1, a2, a3, b0, b1, b2, b3;
  unpack2(mul2(kqa, pack2(dn4.x, dn4.y)), a0, a1);
  unpack2(mul2(kqb, pack2(dn4.z, dn4.w)), a2, a3);
  unpack2(mul2(kqa, pack2(up4.x, up4.y)), b0, b1);
  unpack2(mul2(kqb, pack2(up4.z, up4.w)), b2, b3);
  const f32x2 n2a = mul2(vz2, pack2(add(a0, b0), add(a1, b1)));
  const f32x2 n2b = mul2(vz2, pack2(add(a2, b2), add(a3, b3)));
  float x0, x1, x2, x3, y0, y1, y2, y3;
  unpack2(n1a, x0, x1);
  unpack2(n1b, x2, x3);
  unpack2(n2a, y0, y1);
  unpack2(n2b, y2, y3);
  f32x2 numa = add2(pack2(add(x0, y0), add(x1, y1)), n3a);
  f32x2 numb = add2(pack2(add(x2, y2), add(x3, y3)), n3b);
  if constexpr (HASQ) {                                                     // + input_q (:754)
    const float* qcv = R2_QCV(s, p);
    const float q0 = qcv[rc.z & 0xFFu], q1 = qcv[(rc.z >> 8) & 0xFFu];
    const float q2 = qcv[(rc.z >> 16) & 0xFFu], q3 = qcv[rc.z >> 24];
    numa = add2(numa, pack2(q0, q1));
    numb = add2(numb, pack2(q2, q3));
  }
  const f32x2 oa = div_rn2(numa, ndena, rdena);
  const f32x2 ob = div_rn2(numb, ndenb, rdenb);
  float4 o4;
  unpack2(oa, o4.x, o4.y);
  unpack2(ob, o4.z, o4.w);
  *reinterpret_cast<float4*>(out + base) = o4;
  float d0, d1, d2, d3;
  unpack2(sub2(oa, c01), d0, d1);
  unpack2(sub2(ob, c23), d2, d3);
  lmax = fmaxf(fmaxf(lmax, fabsf(d0)), fabsf(d1));
  return fmaxf(fmaxf(lmax, fabsf(d2)), fabsf(d3));
}

// cv_update_packed without the heat term (x + 0 == x)
__device__ __forceinline__ float cv_update_packed_noq(uint32_t d, float t_jp, float t_jm, float t_im,
                                                      float t_ip, float n3, float t_inf,
                                                      const AreaCoef& az, const Combo* tab) {
  const int idx = (int)(d & kPackIdxMask);
  const float4* c4 = reinterpret_cast<const float4*>(tab + idx);
  const float4 k = c4[0], h = c4[1];
  const float vz = (d & kPackHalfV) ? az.half : az.full;
  const float uz = (d & kPackHalfU) ? az.half : az.full;
  float n1 = add(mul(k.x, t_jp), mul(k.y, t_jm));
  n1 = add(n1, h.x);
  n1 = mul(vz, n1);
  float n2 = add(mul(k.z, t_ip), mul(k.w, t_im));
  n2 = add(n2, h.y);
  n2 = mul(uz, n2);
  const float t = div_rn(add(add(n1, n2), n3), h.z, h.w);
  return idx < kNumMaterials ? t_inf : t;
}

// SLOW vector: boundary classes (possibly mixed with exterior / interior CVs), table-driven
template <bool FIRST, bool HASQ>
__device__ __forceinline__ float slow_vector4(const float* __restrict__ in, float* __restrict__ out,
                                              float* __restrict__ n3p, const uint4 rc,
                                              const Sweep2Ctx& s, const Params& p, float lmax) {
  const int base = (int)(rc.x & 0xFFFFu) * 4, P = p.geom.P;
  const uint32_t flags = rc.x >> 16;
  const float t_inf = s.t_inf;
  const Combo* tab = R2_TAB(s, p);
  AreaCoef az;
  az.full = tab[SBX_CV_INTERIOR * kNumMaterials].vz;
  az.half = tab[SBX_CV_CORNER_TL * kNumMaterials].vz;
  float c[4], up[4], dn[4], n3v[4], o[4];
  load_f<4>(in + base, c);
  if (flags & kRecUp) load_f<4>(in + base - P, up); else fill<4>(up, t_inf);          // :642-644
  if (flags & kRecDown) load_f<4>(in + base + P, dn); else fill<4>(dn, t_inf);        // :646
  const float left = (flags & kRecLeft) ? in[base - 1] : t_inf;                       // :638-640
  const float right = (flags & kRecRight) ? in[base + 4] : t_inf;                     // :636
  if constexpr (FIRST) {
    const float rdt = __frcp_rn(p.dt);
#pragma unroll
    for (int e = 0; e < 4; ++e)
      n3v[e] = div_rn(mul(tab[(rc.y >> (8 * e)) & kPackIdxMask].cm, c[e]), p.dt, rdt);
    store_f<4>(n3p + base, n3v);
  } else {
    load_f<4>(n3p + base, n3v);
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const uint32_t d = (rc.y >> (8 * e)) & 0xFFu;
    const float t_jm = e == 0 ? left : c[e - 1];
    const float t_jp = e == 3 ? right : c[e + 1];
    if constexpr (HASQ) {
      const float qv = R2_QCV(s, p)[(rc.z >> (8 * e)) & 0xFFu];
      o[e] = cv_update_packed(d, t_jp, t_jm, up[e], dn[e], n3v[e], qv, t_inf, az, tab);
    } else {
      o[e] = cv_update_packed_noq(d, t_jp, t_jm, up[e], dn[e], n3v[e], t_inf, az, tab);
    }
    lmax = fmaxf(lmax, fabsf(__fsub_rn(o[e], c[e])));                                  // :851-853
  }
  store_f<4>(out + base, o);
  return lmax;
}

#define R2_SCHED(s, p) (reinterpret_cast<const uint32_t*>((s).smem + (p).g2.off_sched))

template <bool FIRST>
__device__ __forceinline__ float resident2_sweep(const float* __restrict__ in, float* __restrict__ out,
                                                 float* __restrict__ n3p, const Sweep2Ctx& s,
                                                 const Params& p, int warp, int lane,
                                                 HeadVec (&heads)[kR2Heads], bool with_ext) {
  const int P = p.geom.P;
  const uint16_t* flist = R2_FLIST(s, p);
  const uint32_t* sched = R2_SCHED(s, p);
  int i = (int)sched[warp];
  const int i_end = (int)sched[warp + 1];
  float lmax = 0.f;
  {
    // FAST coefficients of (interior, air): from the header, live only in this block so that
    // the record paths below have the registers to themselves
    FastCoef2 fc2;
    float kq1;
    {
      const Combo& c = R2_TAB(s, p)[SBX_CV_INTERIOR * kNumMaterials + 0];
      kq1 = c.k1;
      fc2.kq = pack2(c.k1, c.k1); fc2.vz = pack2(c.vz, c.vz);
      fc2.nden = pack2(-c.den, -c.den); fc2.rden = pack2(c.rden, c.rden);
      fc2.cm = pack2(c.cm, c.cm);
      const float rdt = __frcp_rn(p.dt);
      fc2.ndt = pack2(-p.dt, -p.dt);
      fc2.rdt = pack2(rdt, rdt);
      fc2.one = pack2(p.one, p.one);
    }
    // register-resident HEAD vectors (see HeadVec): the warp's first chunks, full FAST ones
    const int nh = R2_CNT(s, p)[kC2Heads];
    float lh[kR2Heads];
#pragma unroll
    for (int j = 0; j < kR2Heads; ++j) {
      lh[j] = 0.f;
      if (j < nh) {
        if constexpr (FIRST) heads[j].base = (int)flist[(sched[i + j] >> 8) + lane] * 4;
        lh[j] = fast_head4<FIRST>(in, out, heads[j], P, fc2, kq1);
      }
    }
#pragma unroll
    for (int j = 0; j < kR2Heads; ++j) lmax = fmaxf(lmax, lh[j]);
    i += nh;
    // the warp's other FAST chunks, two at a time where both lanes have work: two independent
    // dependency chains in one instruction stream (the sweeps are latency-bound per warp)
#if SBX_R2_PAIR
    for (; i + 1 < i_end; i += 2) {
      const uint32_t d0 = sched[i], d1 = sched[i + 1];
      if ((d1 & 7u) != kR2Fast) break;
      const bool a0 = lane <= (int)((d0 >> 3) & 31u), a1 = lane <= (int)((d1 >> 3) & 31u);
      if (a0 && a1) {
        const int b0 = (int)flist[(d0 >> 8) + lane] * 4, b1 = (int)flist[(d1 >> 8) + lane] * 4;
        const float l0 = fast_vector4<FIRST>(in, out, n3p, b0, P, fc2, kq1, 0.f);
        const float l1 = fast_vector4<FIRST>(in, out, n3p, b1, P, fc2, kq1, 0.f);
        lmax = fmaxf(lmax, fmaxf(l0, l1));
      } else if (a0) {
        lmax = fast_vector4<FIRST>(in, out, n3p, (int)flist[(d0 >> 8) + lane] * 4, P, fc2, kq1, lmax);
      } else if (a1) {
        lmax = fast_vector4<FIRST>(in, out, n3p, (int)flist[(d1 >> 8) + lane] * 4, P, fc2, kq1, lmax);
      }
    }
#endif
    for (; i < i_end; ++i) {
      const uint32_t d = sched[i];
      if ((d & 7u) != kR2Fast) break;
      if (lane <= (int)((d >> 3) & 31u))
        lmax = fast_vector4<FIRST>(in, out, n3p, (int)flist[(d >> 8) + lane] * 4, P, fc2, kq1, lmax);
    }
  }
#if !SBX_R2_PERSISTENT
  // one-building-per-CTA kernel: the records arrive with the second TMA batch
  if constexpr (FIRST) mbar_wait(reinterpret_cast<uint64_t*>(s.smem + p.g2.off_bar) + 1, 0);
#endif
  // record chunks (kind is warp-uniform) and, in the first two sweeps, exterior chunks
  for (; i < i_end; ++i) {
    const uint32_t d = sched[i];
    const uint32_t kind = d & 7u;
    if (lane > (int)((d >> 3) & 31u)) continue;
    const int idx = (int)(d >> 8) + lane;
    if (kind == kR2Ext) {
      // exterior space: T = T_inf (tf_simulator.py:847-849).  After two sweeps both planes
      // hold T_inf there and the delta is exactly 0, so later sweeps skip these vectors.
      if (with_ext) {
        const float t_inf = s.t_inf;
        const int base = (int)flist[idx] * 4;
        const float4 c4 = *reinterpret_cast<const float4*>(in + base);
        *reinterpret_cast<float4*>(out + base) = make_float4(t_inf, t_inf, t_inf, t_inf);
        lmax = fmaxf(fmaxf(lmax, fabsf(__fsub_rn(t_inf, c4.x))), fabsf(__fsub_rn(t_inf, c4.y)));
        lmax = fmaxf(fmaxf(lmax, fabsf(__fsub_rn(t_inf, c4.z))), fabsf(__fsub_rn(t_inf, c4.w)));
      }
      continue;
    }
    const uint4 rc = R2_REC(s, p)[idx];
    if (kind == kR2MedNoQ) lmax = med_vector4<FIRST, false>(in, out, n3p, rc, s, p, lmax);
    else if (kind == kR2MedQ) lmax = med_vector4<FIRST, true>(in, out, n3p, rc, s, p, lmax);
    else if (kind == kR2SlowNoQ) lmax = slow_vector4<FIRST, false>(in, out, n3p, rc, s, p, lmax);
    else lmax = slow_vector4<FIRST, true>(in, out, n3p, rc, s, p, lmax);
  }
  return lmax;
}

// ---------------------------------------------------------------------------
// zone sums
// ---------------------------------------------------------------------------
// sum over the four CVs of round((T - ref) * 2^16): fma(T, 2^16, -ref * 2^16) is exact
// (see to_fix32) and costs one instruction per CV pair
__device__ __forceinline__ void fix4(const float4 t, const f32x2 scale2, const f32x2 nref2,
                                     int& i0, int& i1, int& i2, int& i3) {
  float f0, f1, f2, f3;
  unpack2(fma2(pack2(t.x, t.y), scale2, nref2), f0, f1);
  unpack2(fma2(pack2(t.z, t.w), scale2, nref2), f2, f3);
  i0 = __float2int_rn(f0); i1 = __float2int_rn(f1); i2 = __float2int_rn(f2); i3 = __float2int_rn(f3);
}
__device__ __forceinline__ long long warp_sum_i64(long long v) {
  const unsigned lo = __reduce_add_sync(0xffffffffu, (unsigned)(v & 0xFFFF));
  const unsigned mid = __reduce_add_sync(0xffffffffu, (unsigned)((v >> 16) & 0xFFFF));
  const int hi = __reduce_add_sync(0xffffffffu, (int)(v >> 32));
  return ((long long)hi << 32) + ((long long)mid << 16) + (long long)lo;
}

__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_load_plane_to(float* smem_dst, const void* tmap, int b, uint64_t* bar) {
  tma_load_plane(smem_dst, tmap, b, bar);
}

#ifdef SBX_PROFILE_PHASES
#define R2_PHASE(i)                                                                  \
  do {                                                                               \
    if (tid == 0) {                                                                  \
      const long long now__ = clock64();                                             \
      atomicAdd(&p.phase_cycles[i], (unsigned long long)(now__ - phase_t0__));       \
      phase_t0__ = now__;                                                            \
    }                                                                                \
  } while (0)
#else
#define R2_PHASE(i) do {} while (0)
#endif

// shared-memory scratch words (off_misc)
enum { kMiscMax = 0 /* +parity */, kMiscRot = 2 /* +parity: result plane | load plane << 2 | sweeps << 8 */ };
// mbarriers (off_bar)
enum { kBarFull = 0,   // producer's TMA: plane + header (+ lists) of a building      -> compute
       kBarZl = 1,     // producer's TMA: zone-sum list                               -> compute
       kBarSwept = 2,  // compute: sweeps over, result plane known                    -> producer
       kBarSummed = 3, // compute warps: zone sums in the bins                        -> producer
       kBarPlane = 4,  // producer: the TMA store has read its plane (reusable)       -> compute
       kBarBins = 5 }; // producer: bins read and zeroed                              -> compute

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// barrier + OR-reduction among the compute threads only (named barrier 1)
__device__ __forceinline__ int compute_barrier_or(int pred) {
  int r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.s32 q, %1, 0;\n\t"
      "barrier.cta.red.or.pred p, 1, %2, q;\n\t"
      "selp.s32 %0, 1, 0, p;\n\t}"
      : "=r"(r)
      : "r"(pred), "n"(kR2Compute)
      : "memory");
  return r;
}
__device__ __forceinline__ void compute_barrier() {
  asm volatile("barrier.cta.sync 1, %0;" ::"n"(kR2Compute) : "memory");
}

#if SBX_R2_PERSISTENT
__global__ void __launch_bounds__(kR2Threads, 2)
k_resident_step2(const Params p, const __grid_constant__ CUtensorMap tmap_t) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NT = kR2Compute, NW = NT / 32;
  const ResidentGeom& L = p.geom;
  const Resident2Geom& G = p.g2;
  const int Z = p.Z, n_cv = p.H * p.W;
  const int plane_bytes = G.off_p1 - G.off_p0;
  long long* bins = reinterpret_cast<long long*>(smem + G.off_bins);
  uint32_t* misc = reinterpret_cast<uint32_t*>(smem + G.off_misc);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + G.off_bar);
  const uint32_t hdr_bytes = (uint32_t)header_bytes(Z);
  const int stride = gridDim.x;
  int b = p.b_begin + blockIdx.x;
  if (b >= p.b_end) return;
  const bool shared_plan = p.n_plans == 1;
  const bool sums = !p.fd_only;
  Sweep2Ctx sc;
  sc.smem = smem;
  const int* hcnt = R2_CNT(sc, p);       // [0..7] this plan's list sizes, [8..15] the next building's plan
  auto plane = [&](int i) { return reinterpret_cast<float*>(smem + G.off_p0 + i * plane_bytes); };

  if (tid == 0) {
    mbar_init(bar + kBarFull, 1);
    mbar_init(bar + kBarZl, 1);
    mbar_init(bar + kBarSwept, 1);
    mbar_init(bar + kBarSummed, NW);
    mbar_init(bar + kBarPlane, 1);
    mbar_init(bar + kBarBins, 1);
    misc[kMiscMax] = misc[kMiscMax + 1] = 0u;
  }
  for (int i = tid; i < (Z + 1) * NW; i += kR2Threads) bins[i] = 0;
  __syncthreads();                       // barrier init + bins visible to everyone

  if (warp == NW) {
    // =========================== producer warp ===========================
    // temperature plane + header (+ vector list and records) of building bb -> kBarFull
    auto issue_main = [&](int bb, float* dst, bool with_lists, const int* c) {
      const int plan = shared_plan ? 0 : bb;
      const uint32_t fl_bytes = (uint32_t)(((c[kC2Fast] + c[kC2Ext] + 7) & ~7) * 2);
      const uint32_t rc_bytes = (uint32_t)c[kC2Rec] * 16u;
      const uint32_t sc_bytes = (uint32_t)c[kC2SchedWords] * 4u;
      mbar_expect_tx(bar + kBarFull, (uint32_t)(L.plane_cv * 4) + hdr_bytes +
                                         (with_lists ? fl_bytes + rc_bytes + sc_bytes : 0u));
      tma_load_1d(smem + G.off_hdr, p.hdr + (size_t)bb * hdr_bytes, hdr_bytes, bar + kBarFull);
      tma_load_plane(dst, &tmap_t, bb, bar + kBarFull);
      if (with_lists) {
        if (fl_bytes) tma_load_1d(smem + G.off_flist, p.flist2 + (size_t)plan * L.list_stride, fl_bytes, bar + kBarFull);
        if (rc_bytes) tma_load_1d(smem + G.off_rec, p.rec2 + (size_t)plan * G.rec_cap, rc_bytes, bar + kBarFull);
        tma_load_1d(smem + G.off_sched, p.sched2 + (size_t)plan * r2_sched_cap(p.H * (p.W / 4)), sc_bytes, bar + kBarFull);
      }
    };
    // zone-sum list of building bb -> kBarZl
    auto issue_zlist = [&](int bb, int n_full, int n_part) {
      const int plan = shared_plan ? 0 : bb;
      const uint32_t nf = (uint32_t)n_full, np = (uint32_t)n_part;
      const uint32_t zc_bytes = ((nf + np + 15u) & ~15u);
      mbar_expect_tx(bar + kBarZl, nf * 64u + np * 128u + zc_bytes);
      if (nf) tma_load_1d(smem + G.off_zfull, p.zfull + (size_t)plan * G.zfull_cap, nf * 64u, bar + kBarZl);
      if (np) tma_load_1d(smem + G.off_zpart, p.zpart + (size_t)plan * G.zpart_cap, np * 128u, bar + kBarZl);
      if (zc_bytes) tma_load_1d(smem + G.off_zchunk, p.zchunk + (size_t)plan * G.zchunk_cap, zc_bytes, bar + kBarZl);
    };
    unsigned long long sweeps_acc = 0;
    if (lane == 0) {
      // the first building's list sizes come from global memory (later ones ride in the header)
      int c0[8];
      const int4* cg = reinterpret_cast<const int4*>(p.counts2 + (size_t)(shared_plan ? 0 : b) * 8);
      *reinterpret_cast<int4*>(c0) = __ldg(cg);
      *reinterpret_cast<int4*>(c0 + 4) = __ldg(cg + 1);
      issue_main(b, plane(0), true, c0);
      if (sums) issue_zlist(b, c0[kC2FullChunks], c0[kC2PartChunks]);
    }
    for (int it = 0;; ++it) {
      const uint32_t ph = (uint32_t)(it & 1);
      const int b_next = b + stride;
      const bool has_next = b_next < p.b_end;
      mbar_wait(bar + kBarFull, ph);                       // this building's header is readable
      const float t_inf = R2_QCV(sc, p)[header_q_slots(Z)];
      int cn[8];
      *reinterpret_cast<int4*>(cn) = *reinterpret_cast<const int4*>(hcnt + 8);
      *reinterpret_cast<int4*>(cn + 4) = *reinterpret_cast<const int4*>(hcnt + 12);
      if (lane == 0 && has_next) {                         // warm L2 with what is loaded after the sweeps
        l2_prefetch(p.tbuf[0] + (size_t)b_next * n_cv, (uint32_t)(n_cv * 4) & ~15u);
        l2_prefetch(p.hdr + (size_t)b_next * hdr_bytes, hdr_bytes);
        if (!shared_plan) {
          const uint32_t fl_bytes = (uint32_t)(((cn[kC2Fast] + cn[kC2Ext] + 7) & ~7) * 2);
          if (fl_bytes) l2_prefetch(p.flist2 + (size_t)b_next * L.list_stride, fl_bytes);
          if (cn[kC2Rec]) l2_prefetch(p.rec2 + (size_t)b_next * G.rec_cap, (uint32_t)cn[kC2Rec] * 16u);
        }
      }
      mbar_wait(bar + kBarSwept, ph);                      // sweeps over
      const uint32_t rot = misc[kMiscRot + ph];
      const int k = (int)(rot >> 8);
      if (lane == 0) {
        tma_store_fence();                                 // generic writes -> async proxy
        tma_store_plane(&tmap_t, b, plane((int)(rot & 3u)));
        tma_store_commit();
        if (has_next) issue_main(b_next, plane((int)((rot >> 2) & 3u)), !shared_plan, cn);
        tma_store_wait_read();                             // the stored plane may be overwritten
        mbar_arrive(bar + kBarPlane);
      }
      if (sums) {
        mbar_wait(bar + kBarSummed, ph);                   // every compute warp's sums are in the bins
        long long* zs = p.zone_sum + (size_t)b * (Z + 1);
        long long grid = 0;                                // slot Z collected the CVs outside every zone
        for (int i = lane; i <= Z; i += 32) {
          long long v = 0;
#pragma unroll
          for (int w = 0; w < NW; ++w) {                   // warp-private rows: no atomics
            v += bins[w * (Z + 1) + i];
            bins[w * (Z + 1) + i] = 0;
          }
          if (i < Z) zs[i] = v;
          grid += v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) grid += __shfl_xor_sync(0xffffffffu, grid, o);
        __syncwarp();
        if (lane == 0) {
          zs[Z] = grid;
          p.zone_ref[b] = t_inf;
          mbar_arrive(bar + kBarBins);
          if (has_next && !shared_plan) issue_zlist(b_next, cn[kC2FullChunks], cn[kC2PartChunks]);
        }
      }
      if (lane == 0) {
        p.n_sweeps[b] = k;
        p.max_delta[b] = __uint_as_float(misc[kMiscMax + ph]);
        misc[kMiscMax + ph] = 0u;
        sweeps_acc += (unsigned long long)k;
      }
      if (!has_next) break;
      b = b_next;
    }
    if (lane == 0 && sums) atomicAdd(p.sweeps_total, sweeps_acc);
    return;
  }

  // =========================== compute warps ===========================
  const int limit = p.iteration_limit;
  const float thr = p.threshold;
  int i_in = 0, i_out = 1, i_n3 = 2;
#ifdef SBX_PROFILE_PHASES
  long long phase_t0__ = clock64();
#endif
  for (int it = 0;; ++it) {
    const uint32_t ph = (uint32_t)(it & 1);
    const int b_next = b + stride;
    const bool has_next = b_next < p.b_end;
    float* in = plane(i_in);
    float* out = plane(i_out);
    float* n3p = plane(i_n3);
    R2_PHASE(0);   // loop bookkeeping
    mbar_wait(bar + kBarFull, ph);
    if (it > 0) {
      // the previous result plane becomes this building's n3 plane: the TMA store must have
      // read it, and every warp must be through with its zone sums over it
      mbar_wait(bar + kBarPlane, ph ^ 1u);
      if (sums) mbar_wait(bar + kBarSummed, ph ^ 1u);
    }
    R2_PHASE(1);   // waiting for the TMA loads
    const float t_inf = R2_QCV(sc, p)[header_q_slots(Z)];     // scal[0]
    sc.t_inf = t_inf;
    const int n_full = hcnt[kC2FullChunks], n_chunks = n_full + hcnt[kC2PartChunks];

    // ---- Jacobi sweeps to convergence (simulator.py:348-364) ----
    HeadVec heads[kR2Heads];
#pragma unroll
    for (int j = 0; j < kR2Heads; ++j) {
      heads[j].c = make_float4(0.f, 0.f, 0.f, 0.f);
      heads[j].n3a = heads[j].n3b = 0ull;
      heads[j].base = 0;
    }
    int k = 0;
    float last_lmax = 0.f;
    while (true) {
      ++k;
      float lmax;
      if (k == 1) lmax = resident2_sweep<true>(in, out, n3p, sc, p, warp, lane, heads, true);
      else lmax = resident2_sweep<false>(in, out, n3p, sc, p, warp, lane, heads, k == 2);
#ifdef SBX_PROFILE_PHASES
      const long long t_arrive__ = clock64();
#endif
      const int above = compute_barrier_or(lmax > thr);    // also the ping-pong fence
#ifdef SBX_PROFILE_PHASES
      if (b == p.B / 2 && lane == 0 && k <= 3 && !p.fd_only) {   // one probe building: who waits for whom
        p.phase_cycles[8 + warp * 8 + 2 * (k - 1)] = (unsigned long long)t_arrive__;
        p.phase_cycles[8 + warp * 8 + 2 * (k - 1) + 1] = (unsigned long long)clock64();
      }
#endif
      last_lmax = lmax;
      { float* tmp = in; in = out; out = tmp; }
      { const int t = i_in; i_in = i_out; i_out = t; }
      if (k == 1) R2_PHASE(2); else R2_PHASE(3);           // first sweep (computes n3) / later sweeps
      if (!above || k >= limit) break;
    }
    // `in` holds building.temp of the next step (simulator.py:369); `out` and `n3p` are free

    if (p.conv_perm != nullptr && sums) {                // building.apply_convection()
      const int32_t* perm = p.conv_perm + (size_t)b * n_cv;
      const int W = p.W, P = L.P;
      for (int i = tid; i < n_cv; i += NT) {
        const int src = perm[i];
        const int r = i / W, rs = src / W;
        out[r * P + (i - r * W)] = in[rs * P + (src - rs * W)];
      }
      compute_barrier();
      { float* tmp = in; in = out; out = tmp; }
      { const int t = i_in; i_in = i_out; i_out = t; }
    }
    if (p.conv_p > 0.0 && p.conv_perm == nullptr && sums) {   // device-RNG convection (k_convect_reduce)
      const ConvectPattern cp = convect_pattern(p, b);
      const uint16_t* raw = p.desc + (size_t)(shared_plan ? 0 : b) * n_cv;
      const int W = p.W, P = L.P;
      for (int i = tid; i < n_cv; i += NT) {
        const int r = i / W, c = i - r * W;
        const int src = convect_source(p, cp, raw, r, c);
        const int rs = src / W;
        out[r * P + c] = in[rs * P + (src - rs * W)];
      }
      compute_barrier();
      { float* tmp = in; in = out; out = tmp; }
      { const int t = i_in; i_in = i_out; i_out = t; }
    }
    {
      const uint32_t wm = __reduce_max_sync(0xffffffffu, __float_as_uint(last_lmax));
      if (lane == 0) atomicMax(&misc[kMiscMax + ph], wm);
    }
    if (tid == 0) {
      // hand the result plane to the producer: it stores it and loads the next building
      // into `out`; both happen while the zone sums below run
      misc[kMiscRot + ph] = (uint32_t)i_in | ((uint32_t)i_out << 2) | ((uint32_t)k << 8);
      mbar_arrive(bar + kBarSwept);
    }
    R2_PHASE(4);   // convection gather (if any), hand-over

    // ---- zone / grid sums of the final field ----
    if (sums) {
      if (it == 0 || !shared_plan) mbar_wait(bar + kBarZl, shared_plan ? 0u : ph);
      if (it > 0) mbar_wait(bar + kBarBins, ph ^ 1u);
      const uint16_t* zfull = reinterpret_cast<const uint16_t*>(smem + G.off_zfull);
      const uint32_t* zpart = reinterpret_cast<const uint32_t*>(smem + G.off_zpart);
      const uint8_t* zchunk = reinterpret_cast<const uint8_t*>(smem + G.off_zchunk);
      const f32x2 scale2 = pack2(kFixScaleF, kFixScaleF);
    const float nref = -__fmul_rn(t_inf, kFixScaleF);
    const f32x2 nref2 = pack2(nref, nref);
      const int per = (n_chunks + NW - 1) / NW;
      const int c_lo = min(n_chunks, warp * per), c_hi = min(n_chunks, c_lo + per);
      long long* wbins = bins + warp * (Z + 1);            // this warp's private row
      long long acc = 0;
      int cur_z = -1;
      // entry -> (slot * 4, element mask); null entries have mask 0 and read slot 0
      auto entry = [&](int c, int& off, uint32_t& mask) {
        if (c < n_full) {
          const uint32_t e = zfull[c * 32 + lane];
          mask = e != 0xFFFFu ? 0xFu : 0u;
          off = e != 0xFFFFu ? (int)e * 4 : 0;
        } else {
          const uint32_t en = zpart[(c - n_full) * 32 + lane];
          mask = en >> 16;
          off = (int)(en & 0xFFFFu) * 4;
        }
      };
      auto masked_sum = [&](const float4 t, uint32_t mask) -> int {
        int i0, i1, i2, i3;
        fix4(t, scale2, nref2, i0, i1, i2, i3);
        return ((mask & 1u) ? i0 : 0) + ((mask & 2u) ? i1 : 0) + ((mask & 4u) ? i2 : 0) + ((mask & 8u) ? i3 : 0);
      };
      auto flush = [&]() {
        const long long tot = warp_sum_i64(acc);
        if (lane == 0) wbins[cur_z] += tot;
      };
      auto account = [&](int z, int sv) {
        if (z != cur_z) {                                  // warp-uniform
          if (cur_z >= 0) flush();
          cur_z = z;
          acc = 0;
        }
        acc += sv;
      };
      // four chunks per iteration: list entries, then temperatures, then conversions -- the
      // loads of a group are independent, so their latencies overlap
      int c = c_lo;
      for (; c + 3 < c_hi; c += 4) {
        int off[4];
        uint32_t mk[4];
        float4 t[4];
        int zz[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) zz[j] = (int)zchunk[c + j];
#pragma unroll
        for (int j = 0; j < 4; ++j) entry(c + j, off[j], mk[j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) t[j] = *reinterpret_cast<const float4*>(in + off[j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) account(zz[j], masked_sum(t[j], mk[j]));
      }
      for (; c + 1 < c_hi; c += 2) {
        int off[2];
        uint32_t mk[2];
        float4 t[2];
        const int z0 = (int)zchunk[c], z1 = (int)zchunk[c + 1];
        entry(c, off[0], mk[0]);
        entry(c + 1, off[1], mk[1]);
        t[0] = *reinterpret_cast<const float4*>(in + off[0]);
        t[1] = *reinterpret_cast<const float4*>(in + off[1]);
        account(z0, masked_sum(t[0], mk[0]));
        account(z1, masked_sum(t[1], mk[1]));
      }
      if (c < c_hi) {
        int off;
        uint32_t mk;
        entry(c, off, mk);
        account((int)zchunk[c], masked_sum(*reinterpret_cast<const float4*>(in + off), mk));
      }
      if (cur_z >= 0) flush();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar + kBarSummed);
    }
    R2_PHASE(5);   // zone / grid sums of warp 0
    if (!has_next) break;
    // roles for the next building: its plane was loaded into `out`
    { const int t = i_in; i_in = i_out; i_out = i_n3; i_n3 = t; }
    b = b_next;
  }
}

#else
// One building per CTA (the default): kR2Compute threads, all of them sweep.  Same lists,
// records, schedule and zone-sum code as the persistent variant; what it keeps from
// k_resident_step is the occupancy -- 32 warps per SM at 64 registers, which measured faster
// than 24 compute warps + 2 producer warps at 72 (0.955 vs 0.973 ms per 32768 buildings).
__global__ void __launch_bounds__(kR2Threads, 2)
k_resident_step2(const Params p, const __grid_constant__ CUtensorMap tmap_t) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NT = kR2Compute, NW = NT / 32;
  const ResidentGeom& L = p.geom;
  const Resident2Geom& G = p.g2;
  const int Z = p.Z, n_cv = p.H * p.W, n_items = p.H * (p.W / 4);
  const int b = p.b_begin + blockIdx.x;
  const bool shared_plan = p.n_plans == 1;
  const int plan = shared_plan ? 0 : b;
  const bool sums = !p.fd_only;
  float* in = reinterpret_cast<float*>(smem + G.off_p0);
  float* out = reinterpret_cast<float*>(smem + G.off_p1);
  float* n3p = reinterpret_cast<float*>(smem + G.off_p2);
  long long* bins = reinterpret_cast<long long*>(smem + G.off_bins);
  uint32_t* misc = reinterpret_cast<uint32_t*>(smem + G.off_misc);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + G.off_bar);
  const uint32_t hdr_bytes = (uint32_t)header_bytes(Z);
  const uint32_t sc_cap_bytes = (uint32_t)r2_sched_cap(n_items) * 4u;
  Sweep2Ctx sc;
  sc.smem = smem;
  const int* hcnt = R2_CNT(sc, p);       // [0..7] this plan's list sizes, [8..15] those of building b + prefetch distance

  // ---- first TMA batch: header, temperature plane, vector list, schedule (sizes known up front) ----
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
    misc[kMiscMax] = 0u;
    mbar_expect_tx(bar, (uint32_t)(L.plane_cv * 4) + hdr_bytes + (uint32_t)(L.list_stride * 2) + sc_cap_bytes);
    tma_load_1d(smem + G.off_hdr, p.hdr + (size_t)b * hdr_bytes, hdr_bytes, bar);
    tma_load_plane(in, &tmap_t, b, bar);
    tma_load_1d(smem + G.off_flist, p.flist2 + (size_t)plan * L.list_stride, (uint32_t)(L.list_stride * 2), bar);
    tma_load_1d(smem + G.off_sched, p.sched2 + (size_t)plan * r2_sched_cap(n_items), sc_cap_bytes, bar);
  }
  for (int i = tid; i < (Z + 1) * NW; i += NT) bins[i] = 0;
  __syncthreads();
  mbar_wait(bar, 0);
  const float t_inf = R2_QCV(sc, p)[header_q_slots(Z)];     // scal[0]
  sc.t_inf = t_inf;
  const int n_full = hcnt[kC2FullChunks], n_chunks = n_full + hcnt[kC2PartChunks];
  // ---- second batch: records and the zone-sum list, sized by the header's counts ----
  if (tid == 0) {
    const uint32_t rc_bytes = (uint32_t)hcnt[kC2Rec] * 16u;
    const uint32_t nf = (uint32_t)n_full, np = (uint32_t)(n_chunks - n_full);
    const uint32_t zc_bytes = sums ? ((nf + np + 15u) & ~15u) : 0u;
    mbar_expect_tx(bar + 1, rc_bytes + (sums ? nf * 64u + np * 128u : 0u) + zc_bytes);
    if (rc_bytes) tma_load_1d(smem + G.off_rec, p.rec2 + (size_t)plan * G.rec_cap, rc_bytes, bar + 1);
    if (sums) {
      if (nf) tma_load_1d(smem + G.off_zfull, p.zfull + (size_t)plan * G.zfull_cap, nf * 64u, bar + 1);
      if (np) tma_load_1d(smem + G.off_zpart, p.zpart + (size_t)plan * G.zpart_cap, np * 128u, bar + 1);
      if (zc_bytes) tma_load_1d(smem + G.off_zchunk, p.zchunk + (size_t)plan * G.zchunk_cap, zc_bytes, bar + 1);
    }
  }
  // The CTA that takes over this SM slot works on building b + (CTAs in flight): pull its
  // inputs into L2 now (its list sizes ride in this building's header)
  if (p.prefetch_dist > 0 && tid == 32 && b + p.prefetch_dist < p.b_end) {
    const int bn = b + p.prefetch_dist;
    l2_prefetch(p.tbuf[0] + (size_t)bn * n_cv, (uint32_t)(n_cv * 4) & ~15u);
    l2_prefetch(p.hdr + (size_t)bn * hdr_bytes, hdr_bytes);
    if (!shared_plan) {
      l2_prefetch(p.flist2 + (size_t)bn * L.list_stride, (uint32_t)(L.list_stride * 2));
      l2_prefetch(p.sched2 + (size_t)bn * r2_sched_cap(n_items), sc_cap_bytes);
      if (hcnt[8 + kC2Rec]) l2_prefetch(p.rec2 + (size_t)bn * G.rec_cap, (uint32_t)hcnt[8 + kC2Rec] * 16u);
      if (sums) {
        const uint32_t nf2 = (uint32_t)hcnt[8 + kC2FullChunks], np2 = (uint32_t)hcnt[8 + kC2PartChunks];
        if (nf2) l2_prefetch(p.zfull + (size_t)bn * G.zfull_cap, nf2 * 64u);
        if (np2) l2_prefetch(p.zpart + (size_t)bn * G.zpart_cap, np2 * 128u);
      }
    }
  }

  // ---- Jacobi sweeps to convergence (simulator.py:348-364) ----
  HeadVec heads[kR2Heads];
#pragma unroll
  for (int j = 0; j < kR2Heads; ++j) {
    heads[j].c = make_float4(0.f, 0.f, 0.f, 0.f);
    heads[j].n3a = heads[j].n3b = 0ull;
    heads[j].base = 0;
  }
  const int limit = p.iteration_limit;
  const float thr = p.threshold;
  int k = 0;
  float last_lmax = 0.f;
  while (true) {
    ++k;
    float lmax;
    if (k == 1) lmax = resident2_sweep<true>(in, out, n3p, sc, p, warp, lane, heads, true);
    else lmax = resident2_sweep<false>(in, out, n3p, sc, p, warp, lane, heads, k == 2);
    const int above = __syncthreads_or(lmax > thr);        // also the ping-pong fence
    last_lmax = lmax;
    float* tmp = in; in = out; out = tmp;
    if (!above || k >= limit) break;
  }
  // `in` holds building.temp of the next step (simulator.py:369)
  if (p.conv_perm != nullptr && sums) {                  // building.apply_convection()
    const int32_t* perm = p.conv_perm + (size_t)b * n_cv;
    const int W = p.W, P = L.P;
    for (int i = tid; i < n_cv; i += NT) {
      const int src = perm[i];
      const int r = i / W, rs = src / W;
      out[r * P + (i - r * W)] = in[rs * P + (src - rs * W)];
    }
    __syncthreads();
    float* tmp = in; in = out; out = tmp;
  }
  if (p.conv_p > 0.0 && p.conv_perm == nullptr && sums) {     // device-RNG convection (k_convect_reduce)
    const ConvectPattern cp = convect_pattern(p, b);
    const uint16_t* raw = p.desc + (size_t)plan * n_cv;
    const int W = p.W, P = L.P;
    for (int i = tid; i < n_cv; i += NT) {
      const int r = i / W, c = i - r * W;
      const int src = convect_source(p, cp, raw, r, c);
      const int rs = src / W;
      out[r * P + c] = in[rs * P + (src - rs * W)];
    }
    __syncthreads();
    float* tmp = in; in = out; out = tmp;
  }
  if (tid == 0) {
    tma_store_fence();                                   // generic writes -> async proxy
    tma_store_plane(&tmap_t, b, in);
    tma_store_commit();
  }
  {
    const uint32_t wm = __reduce_max_sync(0xffffffffu, __float_as_uint(last_lmax));
    if (lane == 0) atomicMax(&misc[kMiscMax], wm);
  }
  // ---- zone / grid sums of the final field ----
  if (sums) {
    // (the zone-sum list came with the second TMA batch, which sweep 1 waited for)
    const uint16_t* zfull = reinterpret_cast<const uint16_t*>(smem + G.off_zfull);
    const uint32_t* zpart = reinterpret_cast<const uint32_t*>(smem + G.off_zpart);
    const uint8_t* zchunk = reinterpret_cast<const uint8_t*>(smem + G.off_zchunk);
    const f32x2 scale2 = pack2(kFixScaleF, kFixScaleF);
    const float nref = -__fmul_rn(t_inf, kFixScaleF);
    const f32x2 nref2 = pack2(nref, nref);
    const int per = (n_chunks + NW - 1) / NW;
    const int c_lo = min(n_chunks, warp * per), c_hi = min(n_chunks, c_lo + per);
    long long* wbins = bins + warp * (Z + 1);            // this warp's private row
    long long acc = 0;
    int cur_z = -1;
    // entry -> (slot * 4, element mask); null entries have mask 0 and read slot 0
    auto entry = [&](int c, int& off, uint32_t& mask) {
      if (c < n_full) {
        const uint32_t e = zfull[c * 32 + lane];
        mask = e != 0xFFFFu ? 0xFu : 0u;
        off = e != 0xFFFFu ? (int)e * 4 : 0;
      } else {
        const uint32_t en = zpart[(c - n_full) * 32 + lane];
        mask = en >> 16;
        off = (int)(en & 0xFFFFu) * 4;
      }
    };
    auto masked_sum = [&](const float4 t, uint32_t mask) -> int {
      int i0, i1, i2, i3;
      fix4(t, scale2, nref2, i0, i1, i2, i3);
      return ((mask & 1u) ? i0 : 0) + ((mask & 2u) ? i1 : 0) + ((mask & 4u) ? i2 : 0) + ((mask & 8u) ? i3 : 0);
    };
    auto flush = [&]() {
      const long long tot = warp_sum_i64(acc);
      if (lane == 0) wbins[cur_z] += tot;
    };
    auto account = [&](int z, int sv) {
      if (z != cur_z) {                                  // warp-uniform
        if (cur_z >= 0) flush();
        cur_z = z;
        acc = 0;
      }
      acc += sv;
    };
    // four chunks per iteration: list entries, then temperatures, then conversions -- the
    // loads of a group are independent, so their latencies overlap
    int c = c_lo;
    for (; c + 3 < c_hi; c += 4) {
      int off[4];
      uint32_t mk[4];
      float4 t[4];
      int zz[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) zz[j] = (int)zchunk[c + j];
#pragma unroll
      for (int j = 0; j < 4; ++j) entry(c + j, off[j], mk[j]);
#pragma unroll
      for (int j = 0; j < 4; ++j) t[j] = *reinterpret_cast<const float4*>(in + off[j]);
#pragma unroll
      for (int j = 0; j < 4; ++j) account(zz[j], masked_sum(t[j], mk[j]));
    }
    for (; c + 1 < c_hi; c += 2) {
      int off[2];
      uint32_t mk[2];
      float4 t[2];
      const int z0 = (int)zchunk[c], z1 = (int)zchunk[c + 1];
      entry(c, off[0], mk[0]);
      entry(c + 1, off[1], mk[1]);
      t[0] = *reinterpret_cast<const float4*>(in + off[0]);
      t[1] = *reinterpret_cast<const float4*>(in + off[1]);
      account(z0, masked_sum(t[0], mk[0]));
      account(z1, masked_sum(t[1], mk[1]));
    }
    if (c < c_hi) {
      int off;
      uint32_t mk;
      entry(c, off, mk);
      account((int)zchunk[c], masked_sum(*reinterpret_cast<const float4*>(in + off), mk));
    }
    if (cur_z >= 0) flush();
  }
  __syncthreads();
  if (warp == 0) {
    if (sums) {
      long long* zs = p.zone_sum + (size_t)b * (Z + 1);
      long long grid = 0;                                // slot Z collected the CVs outside every zone
      for (int i = lane; i <= Z; i += 32) {
        long long v = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) v += bins[w * (Z + 1) + i];   // warp-private rows: no atomics
        if (i < Z) zs[i] = v;
        grid += v;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) grid += __shfl_xor_sync(0xffffffffu, grid, o);
      if (lane == 0) {
        zs[Z] = grid;
        p.zone_ref[b] = t_inf;
        atomicAdd(p.sweeps_total, (unsigned long long)k);
      }
    }
    if (lane == 0) {
      p.n_sweeps[b] = k;
      p.max_delta[b] = __uint_as_float(misc[kMiscMax]);
      tma_store_wait_read();
    }
  }
}

#endif  // SBX_R2_PERSISTENT

}  // namespace sbx
