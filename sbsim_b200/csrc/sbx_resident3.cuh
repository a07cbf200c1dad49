// k_resident_step3: the resident diffusion solve with the temperatures in REGISTERS.
//
// k_resident_step / k_resident_step2 keep the field in shared memory and walk lists of 4-CV
// vectors: every vector update loads its centre, its two vertical neighbours, two scalar
// horizontal neighbours and its thermal-mass term and stores its result -- 20-28 shared-memory
// wavefronts and ~68 issue slots per vector per sweep; both kernels sit at ~70 % of the
// shared-memory pipe AND ~70 % of the issue slots (profiles/r01_ncu_full_randomized.txt,
// profiles/r02_*), so neither fewer instructions (v2: -20 %) nor fewer wavefronts alone moved them.
//
// Here a thread owns one or two HALF-TILES of 2 x 4 control volumes for the whole solve:
//   * their temperatures and thermal-mass terms n3 = (cm * T_prev) / dt live in registers from
//     the first sweep to the last; no centre / n3 loads, no list entries, no address arithmetic
//     in the sweeps;
//   * per sweep a half-tile publishes its two rows and its left / right columns (three float4
//     in double-buffered exchange arrays) and reads the facing rows / columns of its four
//     neighbours (four float4): 7 x 128-bit shared-memory accesses per 8 CVs, all conflict-free
//     (the schedule interleaves the slots' 16-byte bank groups), none of them the 4-way
//     conflicted scalar neighbour loads of the vector kernels;
//   * sums that consume packed PRODUCTS are packed too: ptxas contracts add.rn.f32x2 of a
//     mul.rn.f32x2 result into FFMA2 (one rounding instead of two, even with --fmad=false), so the
//     sum is written fma.rn.f32x2(x, ONE, y) with ONE = 1.0f read from the kernel parameters -- a
//     value the compiler cannot see, hence nothing to simplify: RN(x * 1 + y) == RN(x + y);
//   * zone sums come straight from the registers (exact fixed point, integer REDUX per warp,
//     32-bit shared-memory atomics on split lo / hi words), no zone-sum list.
//
// Half-tiles are of three kinds, fixed per plan (k_prepare_plan3):
//   PURE  8 interior air CVs of one zone, no diffuser: uniform coefficients in registers;
//   MED   8 interior-class CVs of any material, diffusers allowed: per-PAIR coefficients from the
//         9-entry material-pair table of the solve header (two 128-bit loads per pair);
//   SLOW  everything else (boundary ring, exterior space): every horizontal pair of CVs carries
//         a pattern id into a per-building table of pair coefficients (the 30-entry
//         (class, material) table of the header interleaved for two CVs), so the general update
//         (tf_simulator.py:719-754 incl. convection terms, half cells and the exterior
//         override) also runs on packed pairs, branch-free.
// A sweep lasts as long as its slowest warp, so the schedule balances COST, not count: the
// half-tiles are sorted by cost; thread t takes the t-th most expensive one, and the cheapest
// n_b = (half-tiles - threads) -- PURE ones -- go as SECOND half-tiles to the threads whose
// first one is cheapest.  With 448 threads for the 768 half-tiles of a 64x96 grid the boundary
// threads carry one SLOW half-tile (~190 issue slots per sweep), the others MED + PURE or
// PURE + PURE (~200).  Kinds are contiguous in thread order: warps diverge only where two
// kinds meet.
//
// Arithmetic contract unchanged: every fp32 operation of tf_simulator.py:719-754 is one IEEE
// rounding in the reference's order; bit-exact against the oracle (tests/test_gpu_parity.py).
//
// Scope: grids with H % 2 == 0, W % 4 == 0 and at most 2 * kR3MaxThreads half-tiles, TF-Jacobi
// solver, no stochastic convection (those handles run k_resident_step).
#pragma once

#include "sbx_kernels.cuh"

namespace sbx {

#ifndef SBX_R3_MAX_THREADS
#define SBX_R3_MAX_THREADS 448
#endif
constexpr int kR3MaxThreads = SBX_R3_MAX_THREADS;
constexpr int kR3PatCap = 64;
constexpr int kR3PatStride = 7;      // float4 per pattern entry (28 words: distinct patterns -> distinct 16-byte bank groups)
constexpr int kR3MedStride = 48;     // bytes per material-pair entry (3 bank groups: the 9 entries do not collide)
enum { kR3Pure = 0, kR3Med = 1, kR3Slow = 2, kR3Idle = 3 };
constexpr uint32_t kEntSlotMask = 0xFFFu;
constexpr int kEntKindShift = 12, kEntFlagShift = 14, kEntZoneShift = 18;
constexpr uint32_t kEntHeat = 1u << 26;
constexpr uint32_t kNbUp = 1u, kNbDown = 2u, kNbLeft = 4u, kNbRight = 8u;
// counts3[plan*4 + i]
enum { kC3Pat = 2, kC3Capable = 3 };
// exchange arrays of one buffer
enum { kXRow0 = 0, kXRow1 = 1, kXCol = 2, kXArrays = 3 };

__host__ inline int resident3_threads(int H, int W) {
  const int n_half = (H / 2) * (W / 4);
  int nt = ((n_half * 7 + 11) / 12 + 31) & ~31;          // 7/12 of the half-tiles: 448 for 64x96
  if (const char* e = getenv("SBX_R3_THREADS")) nt = (atoi(e) + 31) & ~31;
  const int lo = ((n_half + 1) / 2 + 31) & ~31;
  if (nt < lo) nt = lo;
  if (nt > n_half) nt = (n_half + 31) & ~31;
  return nt;
}

__host__ inline bool resident3_supported(int H, int W) {
  if (H % 2 != 0 || W % 4 != 0 || H < 4 || W < 8) return false;
  const int hh = H / 2, tw = W / 4;
  const int Tq = (tw % 2 == 0) ? tw + 1 : tw;
  const int nt = resident3_threads(H, W);
  return nt <= kR3MaxThreads && nt >= kR3PatCap * 6 && 2 * nt >= hh * tw && hh * Tq <= (int)kEntSlotMask;
}

__host__ inline Resident3Geom resident3_geom(int H, int W, int Z, int plane_pitch) {
  Resident3Geom g;
  g.hh = H / 2;
  g.tw = W / 4;
  g.Tq = (g.tw % 2 == 0) ? g.tw + 1 : g.tw;
  g.nt = resident3_threads(H, W);
  g.tq_magic = 0xFFFFFFFFu / (unsigned)g.Tq + 1u;
  g.pat_cap = kR3PatCap;
  auto al = [](size_t v) { return (int)((v + 15) & ~(size_t)15); };
  auto al128 = [](size_t v) { return (int)((v + 127) & ~(size_t)127); };
  g.xarr = al128((size_t)g.hh * g.Tq * 16);
  size_t o = 0;
  g.off_x0 = (int)o; o = al128(o + (size_t)kXArrays * g.xarr);
  // the input plane of the TMA load (rows at the pitch of k_resident_step's tensor map) aliases
  // the second exchange buffer
  const size_t x1 = (size_t)kXArrays * g.xarr, plane = (size_t)H * plane_pitch * 4;
  g.off_x1 = (int)o; o = al128(o + (x1 > plane ? x1 : plane));
  g.plane_pitch = plane_pitch;
  g.off_hdr = (int)o; o = al(o + header_bytes(Z));
  g.off_ptab = (int)o; o = al(o + (size_t)kR3PatCap * kR3PatStride * 16);
  g.off_mtab = (int)o; o = al(o + (size_t)kNumMaterials * kNumMaterials * kR3MedStride);
  g.off_bins = (int)o; o = al(o + (size_t)(Z + 1) * 8);
  g.off_zparts = (int)o; o = al(o + (size_t)g.nt * 8);
  g.off_misc = (int)o; o = al(o + 32);
  g.off_bar = (int)o; o = al(o + 16);
  g.total = (int)o;
  return g;
}

// ---------------------------------------------------------------------------
// Once per uploaded plan: classify the half-tiles, number the pair patterns, build the
// cost-balanced thread schedule.  One CTA per plan; positions by counting smaller keys
// (deterministic).
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t r3_zone_slot(uint32_t d, int Z) {
  if (desc_class(d) == SBX_CV_EXTERIOR) return 0xFFu;       // T == T_inf == the sums' reference: contributes 0
  const int zs = desc_zone(d);
  return zs == SBX_ZONE_NONE ? (uint32_t)Z : (uint32_t)zs;
}

__global__ void __launch_bounds__(kPrepThreads) k_prepare_plan3(const Params p) {
  __shared__ uint32_t key[2 * kR3MaxThreads];
  __shared__ uint32_t key1[2 * kR3MaxThreads];
  __shared__ uint32_t bitmap[32];     // pair pattern a * 32 + b present (SLOW half-tiles)
  __shared__ uint32_t prefix[33];
  __shared__ int s_bad, s_cnt[3];      // half-tiles per kind rank: SLOW, MED, PURE
  const int plan = blockIdx.x, tid = threadIdx.x;
  const Resident3Geom& g = p.g3;
  const int W = p.W, Z = p.Z, n_half = g.hh * g.tw;
  const uint16_t* raw = p.desc + (size_t)plan * p.H * W;
  uint32_t* ent = p.ent3 + (size_t)plan * 2 * g.nt;
  uint4* rec = p.rec3 + (size_t)plan * 2 * g.nt;
  uint2* recz = p.recz3 + (size_t)plan * 2 * g.nt;
  uint16_t* pat = p.pat3 + (size_t)plan * g.pat_cap;
  if (tid < 32) bitmap[tid] = 0;
  if (tid == 0) s_bad = 0;
  __syncthreads();
  auto cv = [&](int it, int e) -> uint32_t {          // descriptor of CV e = 4 * row + col of half-tile it
    const int hr = it / g.tw, tc = it - hr * g.tw;
    return raw[(hr * 2 + (e >> 2)) * W + tc * 4 + (e & 3)];
  };
  auto kind_of = [&](int it, uint32_t& group) -> int {
    const int hr = it / g.tw, tc = it - hr * g.tw;
    bool pure = true, interior = true, heat = false;
    const uint32_t z0 = r3_zone_slot(cv(it, 0), Z);
    uint32_t hsh = 0;
    for (int e = 0; e < 8; ++e) {
      const uint32_t d = cv(it, e);
      pure = pure && ((d & 0x007Fu) == kFastDesc) && r3_zone_slot(d, Z) == z0;
      interior = interior && desc_class(d) == SBX_CV_INTERIOR;
      heat = heat || (d & SBX_DESC_DIFFUSER);
      hsh = hsh * 31u + (uint32_t)combo_index(d);
    }
    // interior-class CVs have all four neighbours inside the grid by construction
    const bool inner = hr > 0 && hr < g.hh - 1 && tc > 0 && tc < g.tw - 1;
    const int kind = (pure && inner) ? kR3Pure : (interior && inner) ? kR3Med : kR3Slow;
    // PURE: by zone (zone sums of a warp in one REDUX); others: half-tiles with heat input first
    // (the heat lookup is a per-thread branch), similar pattern sets next to each other
    group = kind == kR3Pure ? z0 : ((heat ? 0u : 256u) | ((hsh ^ (hsh >> 7) ^ (hsh >> 13)) & 255u));
    return kind;
  };
  // ---- kinds, groups, pattern bitmap ----
  for (int it = tid; it < n_half; it += kPrepThreads) {
    const int hr = it / g.tw, tc = it - hr * g.tw;
    uint32_t group;
    const int kind = kind_of(it, group);
    if (kind == kR3Slow)
      for (int e = 0; e < 8; e += 2) {
        const uint32_t k = (uint32_t)combo_index(cv(it, e)) * 32u + (uint32_t)combo_index(cv(it, e + 1));
        atomicOr(&bitmap[k >> 5], 1u << (k & 31u));
      }
    const uint32_t ord = kind == kR3Slow ? 0u : kind == kR3Med ? 1u : 2u;      // most expensive first
    const int slot = hr * g.Tq + tc;
    key1[it] = (ord << 12) | (group << 3) | (uint32_t)(slot & 7);
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t o = 0;
    for (int w = 0; w < 32; ++w) { prefix[w] = o; o += (uint32_t)__popc(bitmap[w]); }
    prefix[32] = o;
    if ((int)o > g.pat_cap || (int)o * 6 > g.nt) s_bad = 1;      // the kernel builds the table one item per thread
    // Thread blocks per kind, each padded to whole warps (a warp that mixes two kinds runs
    // both code paths): [SLOW | MED | PURE first half-tiles]; the remaining PURE half-tiles are
    // SECOND half-tiles of the last threads (the kernel keeps no record for second ones).
    int n[3] = {0, 0, 0};
    for (int it = 0; it < n_half; ++it) ++n[key1[it] >> 12];
    const int pad_s = (n[0] + 31) & ~31, pad_m = (n[1] + 31) & ~31;
    const int n_a = g.nt - pad_s - pad_m;              // PURE first half-tiles
    if (n_a < 0 || n[2] - n_a > g.nt) s_bad = 1;
    s_cnt[0] = n[0]; s_cnt[1] = n[1]; s_cnt[2] = n[2];
  }
  // rank inside the (kind, group, residue) bucket
  for (int it = tid; it < n_half; it += kPrepThreads) {
    const uint32_t k1 = key1[it];
    uint32_t m = 0;
    for (int j = 0; j < it; ++j) m += (key1[j] == k1) ? 1u : 0u;
    key[it] = ((k1 >> 3) << 14) | (m << 3) | (k1 & 7u);          // kind | group | rank | residue
  }
  __syncthreads();
  const int n_pat = (int)prefix[32];
  auto pat_id = [&](uint32_t k) -> uint32_t {
    return prefix[k >> 5] + (uint32_t)__popc(bitmap[k >> 5] & ((1u << (k & 31u)) - 1u));
  };
  if (!s_bad)
    for (int k = tid; k < 1024; k += kPrepThreads)
      if (bitmap[k >> 5] & (1u << (k & 31))) pat[pat_id((uint32_t)k)] = (uint16_t)((k >> 5) | ((k & 31) << 8));
  // ---- schedule and records ----
  for (int i = tid; i < 2 * g.nt; i += kPrepThreads) ent[i] = (uint32_t)kR3Idle << kEntKindShift;
  __syncthreads();
  if (s_bad) {
    if (tid == 0) p.counts3[(size_t)plan * 4 + kC3Capable] = 0;
    return;
  }
  for (int it = tid; it < n_half; it += kPrepThreads) {
    const uint32_t k2 = key[it];
    int pos = 0;
    for (int j = 0; j < n_half; ++j) pos += (key[j] < k2) ? 1 : 0;
    int idx;
    {
      const int n_s = s_cnt[0], n_m = s_cnt[1], n_p = s_cnt[2];
      const int pad_s = (n_s + 31) & ~31, pad_m = (n_m + 31) & ~31;
      const int n_a = g.nt - pad_s - pad_m, n_b = n_p - n_a;
      if (pos < n_s) idx = pos;
      else if (pos < n_s + n_m) idx = pad_s + (pos - n_s);
      else if (pos - n_s - n_m < n_a) idx = pad_s + pad_m + (pos - n_s - n_m);
      else idx = g.nt + (g.nt - n_b) + (pos - n_s - n_m - n_a);          // second half-tile of thread nt - n_b + j
    }
    const int hr = it / g.tw, tc = it - hr * g.tw;
    uint32_t group;
    const int kind = kind_of(it, group);
    const uint32_t flags = (hr > 0 ? kNbUp : 0u) | (hr < g.hh - 1 ? kNbDown : 0u) | (tc > 0 ? kNbLeft : 0u) |
                           (tc < g.tw - 1 ? kNbRight : 0u);
    uint32_t e = (uint32_t)(hr * g.Tq + tc) | ((uint32_t)kind << kEntKindShift) | (flags << kEntFlagShift);
    if (kind == kR3Pure) {
      e |= r3_zone_slot(cv(it, 0), Z) << kEntZoneShift;
    } else {
      uint32_t pats = 0u, qs[2] = {0u, 0u};
      uint32_t zs[4] = {0u, 0u, 0u, 0u}, zm[4] = {0u, 0u, 0u, 0u};
      int np = 0;
      bool heat = false, bad = false;
      for (int e2 = 0; e2 < 8; ++e2) {
        const uint32_t d = cv(it, e2);
        if ((e2 & 1) == 0) {
          const uint32_t d1 = cv(it, e2 + 1);
          uint32_t id;
          if (kind == kR3Med) id = (uint32_t)(desc_material(d) * kNumMaterials + desc_material(d1));
          else id = s_bad ? 0u : pat_id((uint32_t)combo_index(d) * 32u + (uint32_t)combo_index(d1));
          pats |= id << (8 * (e2 >> 1));
        }
        const bool dq = (d & SBX_DESC_DIFFUSER) != 0;
        heat = heat || dq;
        qs[e2 >> 2] |= (dq ? (uint32_t)desc_zone(d) : (uint32_t)Z) << (8 * (e2 & 3));
        const uint32_t z = r3_zone_slot(d, Z);
        if (z < (uint32_t)Z) {        // CVs outside every zone only enter the grid total
          int k = 0;
          while (k < np && zs[k] != z) ++k;
          if (k == np) {
            if (np == 4) { bad = true; continue; }
            zs[np++] = z;
          }
          zm[k] |= 1u << e2;
        }
      }
      if (bad) atomicExch(&s_bad, 1);
      if (heat) e |= kEntHeat;
      e |= (uint32_t)np << kEntZoneShift;
      rec[idx] = make_uint4(pats, qs[0], qs[1], 0u);
      recz[idx] = make_uint2((zs[0] | (zm[0] << 8)) | ((zs[1] | (zm[1] << 8)) << 16),
                             (zs[2] | (zm[2] << 8)) | ((zs[3] | (zm[3] << 8)) << 16));
    }
    ent[idx] = e;
  }
  __syncthreads();
  if (tid == 0) {
    int32_t* c = p.counts3 + (size_t)plan * 4;
    c[0] = 0;
    c[1] = 0;
    c[kC3Pat] = n_pat;
    c[kC3Capable] = s_bad ? 0 : 1;
  }
}

// ---------------------------------------------------------------------------
// sweeps
// ---------------------------------------------------------------------------
struct PureCoef {
  f32x2 kq, vz, nden, rden, one;
};
// One half-tile: temperatures (2 rows x 4), thermal-mass terms (4 pairs), static description
struct Half {
  float4 t0, t1;
  f32x2 n3[4];
  uint32_t ent;
};
// static description of a non-PURE half-tile (first half-tiles only: second ones are PURE)
struct HalfRec {
  uint32_t pats, q0, q1;     // pair patterns, heat slots of row 0 / row 1
};
// shared-memory context of a sweep
struct R3Ctx {
  const float4* Xr;           // exchange buffer read by this sweep
  float4* Xw;                 // ... written
  const float4* ptab;         // SLOW pair patterns
  const unsigned char* mtab;  // MED material pairs (header)
  const float* qcv;           // heat per diffuser CV of each zone
  int Tq, xq;
  float t_inf;
};

__device__ __forceinline__ float absmax2(float lmax, f32x2 d) {
  float a, b;
  unpack2(d, a, b);
  return fmaxf(fmaxf(lmax, fabsf(a)), fabsf(b));
}

__device__ __forceinline__ void publish_half(float4* X, const int xq, const int s, const float4 t0, const float4 t1) {
  X[kXRow0 * xq + s] = t0;
  X[kXRow1 * xq + s] = t1;
  X[kXCol * xq + s] = make_float4(t0.x, t1.x, t0.w, t1.w);        // left column | right column
}

// One row of a PURE half-tile.  pu / pc / pd: kq * T of the row above, of this row, of the row
// below (pairs of columns 0-1 and 2-3); pl / pr: kq * T of the left / right neighbour CV.
// Same operations, same order as cv_update_fast / fast_core4.
__device__ __forceinline__ float pure_row(float4& t, const f32x2 pua, const f32x2 pub, const f32x2 pca,
                                          const f32x2 pcb, const f32x2 pda, const f32x2 pdb, const float pl,
                                          const float pr, const f32x2 n3a, const f32x2 n3b, const PureCoef& c,
                                          float lmax) {
  float p0, p1, p2, p3;
  unpack2(pca, p0, p1);
  unpack2(pcb, p2, p3);
  // horizontal: kq*T(j+1) + kq*T(j-1) (scalar adds: the operand pairs are not register pairs)
  const f32x2 n1a = mul2(c.vz, pack2(add(p1, pl), add(p2, p0)));
  const f32x2 n1b = mul2(c.vz, pack2(add(p3, p1), add(pr, p2)));
  // vertical: kq*T(i+1) + kq*T(i-1)
  const f32x2 n2a = mul2(c.vz, fma2(pda, c.one, pua));
  const f32x2 n2b = mul2(c.vz, fma2(pdb, c.one, pub));
  const f32x2 oa = div_rn2(add2(fma2(n1a, c.one, n2a), n3a), c.nden, c.rden);
  const f32x2 ob = div_rn2(add2(fma2(n1b, c.one, n2b), n3b), c.nden, c.rden);
  lmax = absmax2(lmax, sub2(oa, pack2(t.x, t.y)));
  lmax = absmax2(lmax, sub2(ob, pack2(t.z, t.w)));
  unpack2(oa, t.x, t.y);
  unpack2(ob, t.z, t.w);
  return lmax;
}

__device__ __forceinline__ float pure_half(Half& h, const R3Ctx& x, const PureCoef& c, float lmax) {
  const int s = (int)(h.ent & kEntSlotMask), xq = x.xq;
  const float4 U = x.Xr[kXRow1 * xq + s - x.Tq];      // bottom row of the half-tile above
  const float4 D = x.Xr[kXRow0 * xq + s + x.Tq];      // top row of the one below
  const float4 CL = x.Xr[kXCol * xq + s - 1];         // .zw: right column of the left neighbour
  const float4 CR = x.Xr[kXCol * xq + s + 1];         // .xy: left column of the right neighbour
  float pl0, pl1, pr0, pr1;
  unpack2(mul2(c.kq, pack2(CL.z, CL.w)), pl0, pl1);
  unpack2(mul2(c.kq, pack2(CR.x, CR.y)), pr0, pr1);
  const f32x2 pua = mul2(c.kq, pack2(U.x, U.y)), pub = mul2(c.kq, pack2(U.z, U.w));
  const f32x2 p0a = mul2(c.kq, pack2(h.t0.x, h.t0.y)), p0b = mul2(c.kq, pack2(h.t0.z, h.t0.w));
  const f32x2 p1a = mul2(c.kq, pack2(h.t1.x, h.t1.y)), p1b = mul2(c.kq, pack2(h.t1.z, h.t1.w));
  const f32x2 pda = mul2(c.kq, pack2(D.x, D.y)), pdb = mul2(c.kq, pack2(D.z, D.w));
  lmax = pure_row(h.t0, pua, pub, p0a, p0b, p1a, p1b, pl0, pr0, h.n3[0], h.n3[1], c, lmax);
  lmax = pure_row(h.t1, p0a, p0b, p1a, p1b, pda, pdb, pl1, pr1, h.n3[2], h.n3[3], c, lmax);
  publish_half(x.Xw, xq, s, h.t0, h.t1);
  return lmax;
}

// MED pair: two interior-class CVs of any material (tf_simulator.py:719-754 with
// k1 = k2 = k3 = k4 = k/dx of the CV's OWN material, no convection terms: x + 0 == x).
// `e`: entry of the header's material-pair table {kq_a, kq_b, -den_a, -den_b | 1/den_a, 1/den_b, cm_a, cm_b}
__device__ __forceinline__ f32x2 med_pair(const unsigned char* __restrict__ e, const f32x2 tjp, const f32x2 tjm,
                                          const f32x2 tip, const f32x2 tim, const f32x2 n3, const f32x2 q,
                                          const bool heat, const PureCoef& c) {
  const float4 a0 = *reinterpret_cast<const float4*>(e);
  const float2 a1 = *reinterpret_cast<const float2*>(e + 16);
  const f32x2 kq = pack2(a0.x, a0.y);
  const f32x2 n1 = mul2(c.vz, fma2(mul2(kq, tjp), c.one, mul2(kq, tjm)));
  const f32x2 n2 = mul2(c.vz, fma2(mul2(kq, tip), c.one, mul2(kq, tim)));
  f32x2 num = add2(fma2(n1, c.one, n2), n3);
  if (heat) num = add2(num, q);                                                            // :754
  return div_rn2(num, pack2(a0.z, a0.w), pack2(a1.x, a1.y));
}

// SLOW pair (cv_update_packed on two CVs at once): coefficient pairs from the pattern table
// entry `e` (see the table build in the kernel).
//   tjp / tjm: T(i,j+1) / T(i,j-1) of the two CVs, tip / tim: T(i+1,j) / T(i-1,j)
__device__ __forceinline__ f32x2 slow_pair(const float4* __restrict__ e, const f32x2 tjp, const f32x2 tjm,
                                           const f32x2 tip, const f32x2 tim, const f32x2 n3, const f32x2 q,
                                           const bool heat, const f32x2 one, const uint32_t tinf_bits) {
  const float4 k13 = e[0], k24 = e[1], hhv = e[2], den = e[3], vu = e[4];
  const float2 em = *reinterpret_cast<const float2*>(&e[5].z);
  f32x2 n1 = fma2(mul2(pack2(k13.x, k13.y), tjp), one, mul2(pack2(k13.z, k13.w), tjm));   // :719-720, 731
  n1 = add2(n1, pack2(hhv.x, hhv.y));                                                      // :732-733
  n1 = mul2(pack2(vu.x, vu.y), n1);                                                        // :734
  f32x2 n2 = fma2(mul2(pack2(k24.x, k24.y), tip), one, mul2(pack2(k24.z, k24.w), tim));   // :721-722, 737
  n2 = add2(n2, pack2(hhv.z, hhv.w));                                                      // :738-739
  n2 = mul2(pack2(vu.z, vu.w), n2);                                                        // :740
  f32x2 num = add2(fma2(n1, one, n2), n3);                                                 // :752-753
  if (heat) num = add2(num, q);                                                            // :754 (x + 0 == x)
  const f32x2 o = div_rn2(num, pack2(den.x, den.y), pack2(den.z, den.w));                  // :843
  // exterior space: T = T_inf (:847-849); the masks are all ones there
  float o0, o1;
  unpack2(o, o0, o1);
  const uint32_t m0 = __float_as_uint(em.x), m1 = __float_as_uint(em.y);
  o0 = __uint_as_float((__float_as_uint(o0) & ~m0) | (tinf_bits & m0));
  o1 = __uint_as_float((__float_as_uint(o1) & ~m1) | (tinf_bits & m1));
  return pack2(o0, o1);
}

template <int KIND>
__device__ __forceinline__ float general_half(Half& h, const HalfRec& rec, const R3Ctx& x, const PureCoef& c,
                                              float lmax) {
  const int s = (int)(h.ent & kEntSlotMask), xq = x.xq;
  const uint32_t flags = (h.ent >> kEntFlagShift) & 15u;
  const bool heat = (h.ent & kEntHeat) != 0u;
  const float t_inf = x.t_inf;
  const float4 inf4 = make_float4(t_inf, t_inf, t_inf, t_inf);
  float4 U, D, CL, CR;
  if (KIND == kR3Med) {             // interior-class CVs: every neighbour exists
    U = x.Xr[kXRow1 * xq + s - x.Tq];
    D = x.Xr[kXRow0 * xq + s + x.Tq];
    CL = x.Xr[kXCol * xq + s - 1];
    CR = x.Xr[kXCol * xq + s + 1];
  } else {
    U = (flags & kNbUp) ? x.Xr[kXRow1 * xq + s - x.Tq] : inf4;          // tf_simulator.py:642-644
    D = (flags & kNbDown) ? x.Xr[kXRow0 * xq + s + x.Tq] : inf4;        // :646
    CL = (flags & kNbLeft) ? x.Xr[kXCol * xq + s - 1] : inf4;           // :638-640
    CR = (flags & kNbRight) ? x.Xr[kXCol * xq + s + 1] : inf4;          // :636
  }
  const uint32_t tinf_bits = __float_as_uint(t_inf);
  const float4 r0 = h.t0, r1 = h.t1;
  const uint32_t pats = rec.pats;
  f32x2 q[4] = {0ull, 0ull, 0ull, 0ull};
  if (heat) {
    const uint32_t w0 = rec.q0, w1 = rec.q1;
    q[0] = pack2(x.qcv[w0 & 0xFFu], x.qcv[(w0 >> 8) & 0xFFu]);
    q[1] = pack2(x.qcv[(w0 >> 16) & 0xFFu], x.qcv[w0 >> 24]);
    q[2] = pack2(x.qcv[w1 & 0xFFu], x.qcv[(w1 >> 8) & 0xFFu]);
    q[3] = pack2(x.qcv[(w1 >> 16) & 0xFFu], x.qcv[w1 >> 24]);
  }
  auto pair = [&](const int j, const f32x2 tjp, const f32x2 tjm, const f32x2 tip, const f32x2 tim) -> f32x2 {
    const uint32_t id = (pats >> (8 * j)) & 0xFFu;
    if (KIND == kR3Med) return med_pair(x.mtab + id * kR3MedStride, tjp, tjm, tip, tim, h.n3[j], q[j], heat, c);
    return slow_pair(x.ptab + id * kR3PatStride, tjp, tjm, tip, tim, h.n3[j], q[j], heat, c.one, tinf_bits);
  };
  const f32x2 mid0 = pack2(r0.y, r0.z), mid1 = pack2(r1.y, r1.z);
  const f32x2 o0a = pair(0, mid0, pack2(CL.z, r0.x), pack2(r1.x, r1.y), pack2(U.x, U.y));
  const f32x2 o0b = pair(1, pack2(r0.w, CR.x), mid0, pack2(r1.z, r1.w), pack2(U.z, U.w));
  const f32x2 o1a = pair(2, mid1, pack2(CL.w, r1.x), pack2(D.x, D.y), pack2(r0.x, r0.y));
  const f32x2 o1b = pair(3, pack2(r1.w, CR.y), mid1, pack2(D.z, D.w), pack2(r0.z, r0.w));
  lmax = absmax2(lmax, sub2(o0a, pack2(r0.x, r0.y)));               // :851-853
  lmax = absmax2(lmax, sub2(o0b, pack2(r0.z, r0.w)));
  lmax = absmax2(lmax, sub2(o1a, pack2(r1.x, r1.y)));
  lmax = absmax2(lmax, sub2(o1b, pack2(r1.z, r1.w)));
  unpack2(o0a, h.t0.x, h.t0.y);
  unpack2(o0b, h.t0.z, h.t0.w);
  unpack2(o1a, h.t1.x, h.t1.y);
  unpack2(o1b, h.t1.z, h.t1.w);
  publish_half(x.Xw, xq, s, h.t0, h.t1);
  return lmax;
}

__device__ __forceinline__ float sweep_half(Half& h, const HalfRec& rec, const R3Ctx& x, const PureCoef& c,
                                            float lmax) {
  const int kind = (int)((h.ent >> kEntKindShift) & 3u);
  if (kind == kR3Pure) return pure_half(h, x, c, lmax);
  if (kind == kR3Med) return general_half<kR3Med>(h, rec, x, c, lmax);
  if (kind == kR3Slow) return general_half<kR3Slow>(h, rec, x, c, lmax);
  return lmax;
}

// n3 = (cm * T_prev) / dt (tf_simulator.py:743-749), kept for every sweep of this step
__device__ __forceinline__ void init_half(Half& h, const uint32_t pats, const R3Ctx& x, const float cm_air,
                                          const float dt) {
  const int kind = (int)((h.ent >> kEntKindShift) & 3u);
  const float rdt = __frcp_rn(dt);
  const f32x2 ndt2 = pack2(-dt, -dt), rdt2 = pack2(rdt, rdt);
  f32x2 cm[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t id = (pats >> (8 * j)) & 0xFFu;
    if (kind == kR3Med) {
      const float2 v = *reinterpret_cast<const float2*>(x.mtab + id * kR3MedStride + 24);
      cm[j] = pack2(v.x, v.y);
    } else if (kind == kR3Slow) {
      const float2 v = *reinterpret_cast<const float2*>(&x.ptab[id * kR3PatStride + 5]);
      cm[j] = pack2(v.x, v.y);
    } else {
      cm[j] = pack2(cm_air, cm_air);
    }
  }
  h.n3[0] = div_rn2(mul2(cm[0], pack2(h.t0.x, h.t0.y)), ndt2, rdt2);
  h.n3[1] = div_rn2(mul2(cm[1], pack2(h.t0.z, h.t0.w)), ndt2, rdt2);
  h.n3[2] = div_rn2(mul2(cm[2], pack2(h.t1.x, h.t1.y)), ndt2, rdt2);
  h.n3[3] = div_rn2(mul2(cm[3], pack2(h.t1.z, h.t1.w)), ndt2, rdt2);
}

// exact segmented warp sum: lanes with the same `slot` add their `v`; one lane per slot
// adds the total to the split lo / hi bins (32-bit shared-memory atomics)
__device__ __forceinline__ void r3_zone_add(uint32_t* bins, int Z, const int slot, const int v, const bool valid,
                                            const int lane) {
  unsigned todo = __ballot_sync(0xffffffffu, valid);
  if (todo == 0xffffffffu && __all_sync(0xffffffffu, slot == __shfl_sync(0xffffffffu, slot, 0))) {
    const unsigned lo = __reduce_add_sync(0xffffffffu, (unsigned)v & 0xFFFFu);      // the common warp: one zone
    const int hi = __reduce_add_sync(0xffffffffu, v >> 16);
    if (lane == 0) {
      atomicAdd(&bins[slot], lo);
      atomicAdd(reinterpret_cast<int*>(&bins[Z + 1 + slot]), hi);
    }
    return;
  }
  while (todo) {
    const int leader = __ffs(todo) - 1;
    const int z = __shfl_sync(0xffffffffu, slot, leader);
    const bool mine = valid && slot == z;
    const unsigned m = __ballot_sync(0xffffffffu, mine);
    if (mine) {
      const unsigned lo = __reduce_add_sync(m, (unsigned)v & 0xFFFFu);
      const int hi = __reduce_add_sync(m, v >> 16);
      if (lane == leader) {
        atomicAdd(&bins[z], lo);
        atomicAdd(reinterpret_cast<int*>(&bins[Z + 1 + z]), hi);
      }
    }
    todo &= ~m;
  }
}

// round((T - ref) * 2^16) of the 8 CVs of a half-tile: fma(T, 2^16, -ref * 2^16) is exact (to_fix32)
__device__ __forceinline__ void fix_half(const Half& h, const f32x2 scale2, const f32x2 nref2, int (&f)[8]) {
  float a[8];
  unpack2(fma2(pack2(h.t0.x, h.t0.y), scale2, nref2), a[0], a[1]);
  unpack2(fma2(pack2(h.t0.z, h.t0.w), scale2, nref2), a[2], a[3]);
  unpack2(fma2(pack2(h.t1.x, h.t1.y), scale2, nref2), a[4], a[5]);
  unpack2(fma2(pack2(h.t1.z, h.t1.w), scale2, nref2), a[6], a[7]);
#pragma unroll
  for (int e = 0; e < 8; ++e) f[e] = __float2int_rn(a[e]);
}

// zone sums of one half-tile slot of the warp: PURE half-tiles by REDUX (one zone per warp in
// the common case), the others -- zones scattered over the warp -- by direct atomics
__device__ __forceinline__ void zone_sums_half(const Half& h, const int (&f)[8], const int total, uint32_t* bins,
                                               const int Z, const uint2* recz_g, const int lane) {
  const int kind = (int)((h.ent >> kEntKindShift) & 3u);
  const int zone = (int)((h.ent >> kEntZoneShift) & 0xFFu);
  r3_zone_add(bins, Z, zone, total, kind == kR3Pure, lane);
  if (kind == kR3Med || kind == kR3Slow) {
    const int np = zone;
    const uint2 pz = *recz_g;
    for (int k2 = 0; k2 < np; ++k2) {
      const uint32_t part = (k2 < 2 ? pz.x >> (16 * k2) : pz.y >> (16 * (k2 - 2))) & 0xFFFFu;
      int sv = 0;
#pragma unroll
      for (int e = 0; e < 8; ++e) sv += (part & (0x100u << e)) ? f[e] : 0;
      const int z = (int)(part & 0xFFu);
      atomicAdd(&bins[z], (unsigned)sv & 0xFFFFu);
      atomicAdd(reinterpret_cast<int*>(&bins[Z + 1 + z]), sv >> 16);
    }
  }
}

__global__ void __launch_bounds__(kR3MaxThreads, 2) k_resident_step3(const Params p,
                                                                        const __grid_constant__ CUtensorMap tmap_t) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31;
  const Resident3Geom& G = p.g3;
  const int b = p.b_begin + blockIdx.x;
  const int plan = p.n_plans == 1 ? 0 : b;
  const int Z = p.Z, W = p.W, n_cv = p.H * W;
  const int nt = G.nt;
  const Combo* tab = reinterpret_cast<const Combo*>(smem + G.off_hdr);
  const float* qcv = reinterpret_cast<const float*>(smem + G.off_hdr + sizeof(Combo) * kNumCombos);
  const float* scal = qcv + header_q_slots(Z);
  float4* ptab = reinterpret_cast<float4*>(smem + G.off_ptab);
  uint32_t* bins = reinterpret_cast<uint32_t*>(smem + G.off_bins);     // lo[Z+1], hi[Z+1]
  uint32_t* misc = reinterpret_cast<uint32_t*>(smem + G.off_misc);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + G.off_bar);
  float* gT = p.tbuf[0] + (size_t)b * n_cv;
  const uint32_t hdr_bytes = (uint32_t)header_bytes(Z);
  const bool sums = !p.fd_only;
#ifdef SBX_PROFILE_PHASES
  long long phase_t0__ = clock64();
  const long long cta_t0__ = phase_t0__;
#endif

  // ---- stage 0: bulk loads (plane + header); the schedule straight from global, every load
  // independent of the others (one L2 round trip) ----
  float* plane = reinterpret_cast<float*>(smem + G.off_x1);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_expect_tx(bar, (uint32_t)(p.H * G.plane_pitch * 4) + hdr_bytes);
    tma_load_plane(plane, &tmap_t, b, bar);
    tma_load_1d(smem + G.off_hdr, p.hdr + (size_t)b * hdr_bytes, hdr_bytes, bar);
    misc[0] = 0u;
  }
  const uint32_t* g_ent = p.ent3 + (size_t)plan * 2 * nt;
  const uint4* g_rec = p.rec3 + (size_t)plan * 2 * nt;
  Half ha, hb;
  ha.ent = __ldg(g_ent + tid);
  hb.ent = __ldg(g_ent + nt + tid);
  const int n_pat = __ldg(p.counts3 + (size_t)plan * 4 + kC3Pat);
  // this thread's item of the pattern table build (stage 1): the pattern's two combo indices
  const int pt_i = tid / 6, pt_j = tid - pt_i * 6;
  const uint32_t pt_pr = pt_i < G.pat_cap ? (uint32_t)__ldg(p.pat3 + (size_t)plan * G.pat_cap + pt_i) : 0u;
  const int kind_a = (int)((ha.ent >> kEntKindShift) & 3u), kind_b = (int)((hb.ent >> kEntKindShift) & 3u);
  // records of the non-PURE half-tiles: those threads come first, so the first nt/2 load theirs
  // without waiting for the schedule entry (a PURE thread in that range reads a stale record it
  // never uses)
  HalfRec ra;
  ra.pats = ra.q0 = ra.q1 = 0u;
  const bool eager = tid < nt / 2;
  uint4 r4 = make_uint4(0u, 0u, 0u, 0u);
  uint2 z2 = make_uint2(0u, 0u);
  if (eager) {
    r4 = __ldg(g_rec + tid);
    z2 = __ldg(p.recz3 + (size_t)plan * 2 * nt + tid);
  }
  const bool general_a = kind_a == kR3Med || kind_a == kR3Slow;
  if (general_a && !eager) {
    r4 = __ldg(g_rec + tid);
    z2 = __ldg(p.recz3 + (size_t)plan * 2 * nt + tid);
  }
  if (general_a) {
    ra.pats = r4.x; ra.q0 = r4.y; ra.q1 = r4.z;
    // zone parts: only needed after the sweeps -- parked in shared memory, not in registers
    reinterpret_cast<uint2*>(smem + G.off_zparts)[tid] = z2;
  }
  // The CTA that takes over this SM slot works on building b + (CTAs in flight): pull its
  // inputs into L2 now
  if (p.prefetch_dist > 0 && tid == 32 && b + p.prefetch_dist < p.b_end) {
    const int bn = b + p.prefetch_dist;
    l2_prefetch(p.tbuf[0] + (size_t)bn * n_cv, (uint32_t)(n_cv * 4));
    l2_prefetch(p.hdr + (size_t)bn * hdr_bytes, hdr_bytes & ~15u);
    if (p.n_plans != 1) {
      l2_prefetch(p.ent3 + (size_t)bn * 2 * nt, (uint32_t)(2 * nt * 4));
      l2_prefetch(p.rec3 + (size_t)bn * 2 * nt, (uint32_t)(nt / 2 * 16));     // the most expensive half-tiles come first
      l2_prefetch(p.pat3 + (size_t)bn * G.pat_cap, (uint32_t)(G.pat_cap * 2));
      l2_prefetch(p.counts3 + (size_t)bn * 4, 16u);
    }
  }
  if (tid < 2 * (Z + 1)) bins[tid] = 0u;                       // Z <= 254, nt >= 384
  if (tid + nt < 2 * (Z + 1)) bins[tid + nt] = 0u;
  __syncthreads();
  mbar_wait(bar, 0);
  SBX_PHASE(0);   // launch .. bulk loads and per-thread list entries have arrived
  const float t_inf = scal[0];

  // ---- stage 1: pair-pattern table of this building (coefficients change with h and T_inf) ----
  if (pt_i < n_pat) {        // kR3PatCap * 6 <= nt: one item per thread
    const int ia = (int)(pt_pr & 0xFFu), ib = (int)(pt_pr >> 8);
    const Combo& ca = tab[ia];
    const Combo& cb = tab[ib];
    float4 v;
    if (pt_j == 0) v = make_float4(ca.k1, cb.k1, ca.k3, cb.k3);
    else if (pt_j == 1) v = make_float4(ca.k2, cb.k2, ca.k4, cb.k4);
    else if (pt_j == 2) v = make_float4(ca.hh, cb.hh, ca.hv, cb.hv);
    else if (pt_j == 3) v = make_float4(-ca.den, -cb.den, ca.rden, cb.rden);
    else if (pt_j == 4) v = make_float4(ca.vz, cb.vz, ca.uz, cb.uz);
    else v = make_float4(ca.cm, cb.cm, __uint_as_float(ia < kNumMaterials ? 0xFFFFFFFFu : 0u),
                         __uint_as_float(ib < kNumMaterials ? 0xFFFFFFFFu : 0u));
    ptab[pt_i * kR3PatStride + pt_j] = v;
  }
  if (tid >= nt - 2 * kNumMaterials * kNumMaterials) {      // material-pair table at a conflict-free stride
    const int j = nt - 1 - tid;
    *reinterpret_cast<float4*>(smem + G.off_mtab + (j >> 1) * kR3MedStride + (j & 1) * 16) =
        *reinterpret_cast<const float4*>(smem + G.off_hdr + header_pair_offset(Z) + j * 16);
  }

  // ---- stage 2: the thread's half-tiles -> registers, their rims -> exchange buffer 0 ----
  R3Ctx x;
  x.Tq = G.Tq;
  x.xq = G.xarr / 16;
  x.ptab = ptab;
  x.mtab = smem + G.off_mtab;
  x.qcv = qcv;
  x.t_inf = t_inf;
  float4* X0 = reinterpret_cast<float4*>(smem + G.off_x0);
  float4* X1 = reinterpret_cast<float4*>(smem + G.off_x1);
  PureCoef pc;
  {
    // Uniform coefficients of (interior, air).  The warp-wide maximum of a value every lane
    // holds is that value -- and it comes back in a UNIFORM register (CREDUX), which the packed
    // arithmetic takes as a broadcast operand: no vector registers spent on constants.
    const float one = p.one;
    pc.one = pack2(one, one);
    const Combo& c = tab[SBX_CV_INTERIOR * kNumMaterials + 0];
    const float kq = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(c.k1)));
    const float vz = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(c.vz)));
    const float den = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(c.den)));
    const float rden = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(c.rden)));
    pc.kq = pack2(kq, kq); pc.vz = pack2(vz, vz);
    pc.nden = pack2(-den, -den); pc.rden = pack2(rden, rden);
  }
  auto half_offset = [&](const Half& h, const int pitch) -> int {
    const int s = (int)(h.ent & kEntSlotMask);
    const int hr = (int)__umulhi((unsigned)s, G.tq_magic);
    return (hr * 2) * pitch + (s - hr * G.Tq) * 4;
  };
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  ha.t0 = ha.t1 = hb.t0 = hb.t1 = zero4;
  if (kind_a != kR3Idle) {
    const float* q = plane + half_offset(ha, G.plane_pitch);
    ha.t0 = *reinterpret_cast<const float4*>(q);
    ha.t1 = *reinterpret_cast<const float4*>(q + G.plane_pitch);
  }
  if (kind_b != kR3Idle) {
    const float* q = plane + half_offset(hb, G.plane_pitch);
    hb.t0 = *reinterpret_cast<const float4*>(q);
    hb.t1 = *reinterpret_cast<const float4*>(q + G.plane_pitch);
  }
  if (kind_a != kR3Idle) publish_half(X0, x.xq, (int)(ha.ent & kEntSlotMask), ha.t0, ha.t1);
  if (kind_b != kR3Idle) publish_half(X0, x.xq, (int)(hb.ent & kEntSlotMask), hb.t0, hb.t1);
  // one barrier: every thread has its half-tiles (the plane, which aliases exchange buffer 1, is
  // free), the coefficient tables are complete and every rim is in exchange buffer 0
  __syncthreads();
  SBX_PHASE(1);   // pattern table, half-tiles -> registers, first rim exchange
  {
    const float cm_air = tab[SBX_CV_INTERIOR * kNumMaterials + 0].cm;
    init_half(ha, ra.pats, x, cm_air, p.dt);
    init_half(hb, 0u, x, cm_air, p.dt);
  }
  SBX_PHASE(2);   // thermal-mass terms

  // ---- stage 3: Jacobi sweeps to convergence (simulator.py:348-364) ----
  const int limit = p.iteration_limit;
  const float thr = p.threshold;
  x.Xr = X0;
  x.Xw = X1;
  int k = 0;
  float lmax = 0.f;
  while (k < limit) {
    ++k;
    lmax = sweep_half(ha, ra, x, pc, 0.f);
    if (kind_b == kR3Pure) lmax = pure_half(hb, x, pc, lmax);
    // max|dT| <= threshold  <=>  no thread saw a delta above it (simulator.py:362); the
    // barrier doubles as the fence between the two exchange buffers
#ifdef SBX_PROFILE_PHASES
    const long long t_arrive__ = clock64();
#endif
    const int above = __syncthreads_or(lmax > thr);
#ifdef SBX_PROFILE_PHASES
    if (b == p.B / 2 && lane == 0 && k <= 3 && !p.fd_only) {   // one probe CTA: who waits for whom
      p.phase_cycles[8 + (tid >> 5) * 8 + 2 * (k - 1)] = (unsigned long long)(t_arrive__ - cta_t0__);
      p.phase_cycles[8 + (tid >> 5) * 8 + 2 * (k - 1) + 1] = (unsigned long long)(clock64() - cta_t0__);
    }
#endif
    {
      float4* tmp = x.Xw;
      x.Xw = const_cast<float4*>(x.Xr);
      x.Xr = tmp;
    }
    if (k == 1) SBX_PHASE(3); else SBX_PHASE(4);   // first sweep / later sweeps
    if (!above) break;
  }

  // ---- stage 4: write back, straight from the registers ----
  if (kind_a != kR3Idle) {
    float* q = gT + half_offset(ha, W);
    *reinterpret_cast<float4*>(q) = ha.t0;
    *reinterpret_cast<float4*>(q + W) = ha.t1;
  }
  if (kind_b != kR3Idle) {
    float* q = gT + half_offset(hb, W);
    *reinterpret_cast<float4*>(q) = hb.t0;
    *reinterpret_cast<float4*>(q + W) = hb.t1;
  }
  {
    const uint32_t wm = __reduce_max_sync(0xffffffffu, __float_as_uint(lmax));
    if (lane == 0) atomicMax(&misc[0], wm);
  }
  // ---- stage 5: zone / grid sums from the registers ----
  if (sums) {
    const f32x2 scale2 = pack2(kFixScaleF, kFixScaleF);
    const float nref = -__fmul_rn(t_inf, kFixScaleF);
    const f32x2 nref2 = pack2(nref, nref);
    int fa[8], fb[8];
    fix_half(ha, scale2, nref2, fa);
    fix_half(hb, scale2, nref2, fb);      // (idle: temperatures 0 -> garbage, never used)
    int ta = 0, tb = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) { ta += fa[e]; tb += fb[e]; }
    if (kind_a == kR3Idle) ta = 0;
    if (kind_b == kR3Idle) tb = 0;
    // whole grid (slot Z): every CV of every half-tile (exterior space holds T_inf, the reference: 0)
    {
      const int total = ta + tb;
      const unsigned lo = __reduce_add_sync(0xffffffffu, (unsigned)total & 0xFFFFu);
      const int hi = __reduce_add_sync(0xffffffffu, total >> 16);
      if (lane == 0) {
        atomicAdd(&bins[Z], lo);
        atomicAdd(reinterpret_cast<int*>(&bins[Z + 1 + Z]), hi);
      }
    }
    zone_sums_half(ha, fa, ta, bins, Z, reinterpret_cast<const uint2*>(smem + G.off_zparts) + tid, lane);
    r3_zone_add(bins, Z, (int)((hb.ent >> kEntZoneShift) & 0xFFu), tb, kind_b == kR3Pure, lane);
  }
  __syncthreads();      // bins complete
  SBX_PHASE(5);   // write-back, zone sums
  if (tid < 32) {
    if (sums) {
      long long* zs = p.zone_sum + (size_t)b * (Z + 1);
      for (int i = lane; i <= Z; i += 32)                // slot Z: the whole grid
        zs[i] = (long long)(int)bins[Z + 1 + i] * 65536 + (long long)bins[i];
      if (lane == 0) {
        p.zone_ref[b] = t_inf;
        atomicAdd(p.sweeps_total, (unsigned long long)k);
      }
    }
    if (lane == 0) {
      p.n_sweeps[b] = k;
      p.max_delta[b] = __uint_as_float(misc[0]);
    }
  }
  SBX_PHASE(6);   // results out
}

}  // namespace sbx
