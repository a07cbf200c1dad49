// k_resident_step3: the resident diffusion solve with the temperatures in REGISTERS.
//
// k_resident_step / k_resident_step2 keep the field in shared memory and walk lists of 4-CV
// vectors: every vector update loads its centre, its two vertical neighbours, two scalar
// horizontal neighbours and its thermal-mass term and stores its result -- 20-28 shared-memory
// wavefronts and ~68 issue slots per vector per sweep; both kernels sit at ~70 % of the
// shared-memory pipe AND ~70 % of the issue slots (profiles/r01_ncu_full_randomized.txt,
// profiles/r02_*), so neither fewer instructions (v2: -20 %) nor fewer wavefronts alone moved them.
//
// Here one thread owns a 4x4 TILE of control volumes for the whole solve:
//   * its 16 temperatures and 16 thermal-mass terms n3 = (cm * T_prev) / dt live in registers
//     from the first sweep to the last; 12 of the 16 CVs find all four neighbours in the
//     thread's own registers;
//   * per sweep a tile exchanges only its rim with its four neighbour tiles: top row, bottom
//     row, left column, right column, each one float4 in a double-buffered exchange array --
//     4 x LDS.128 + 4 x STS.128 per 16 CVs (8 wavefronts per 4 CVs instead of 20-28), all
//     conflict-free (the thread -> tile list interleaves the slots' 16-byte bank groups);
//   * no per-vector list entry, address arithmetic or thermal-mass load in the sweeps;
//   * sums that consume packed PRODUCTS are packed too: ptxas contracts add.rn.f32x2 of a
//     mul.rn.f32x2 result into FFMA2 (one rounding instead of two, even with --fmad=false), so the
//     sum is written fma.rn.f32x2(x, ONE, y) with ONE = 1.0f read from the kernel parameters -- a
//     value the compiler cannot see, hence nothing to simplify: RN(x * 1 + y) == RN(x + y);
//   * zone sums come straight from the registers (exact fixed point, integer REDUX per warp,
//     32-bit shared-memory atomics on split lo / hi words), no zone-sum list.
//
// Tiles are of two kinds, fixed per plan (k_prepare_plan3):
//   PURE  16 interior air CVs of one zone, no diffuser: uniform coefficients in registers;
//   GEN   everything else (walls, diffusers, boundary ring, exterior space): every horizontal
//         PAIR of CVs carries a pattern id into a per-building table of pair coefficients
//         (the 30-entry (class, material) table of the solve header, interleaved for two CVs),
//         so the general update (tf_simulator.py:719-754 incl. convection terms, half cells and
//         the exterior override) also runs on packed pairs, branch-free.
// Threads are sorted PURE first, then GEN, so warps do not diverge except one per building.
//
// Arithmetic contract unchanged: every fp32 operation of tf_simulator.py:719-754 is one IEEE
// rounding in the reference's order; bit-exact against the oracle (tests/test_gpu_parity.py).
//
// Scope: grids with H % 4 == 0, W % 4 == 0 and at most kR3MaxThreads tiles (64x96: 384), TF-Jacobi
// solver, no stochastic convection (those handles run k_resident_step).
#pragma once

#include "sbx_kernels.cuh"

namespace sbx {

constexpr int kR3MaxThreads = 384;
constexpr int kR3PatCap = 64;
enum { kR3Pure = 0, kR3Gen = 1, kR3Idle = 3 };
constexpr uint32_t kEntSlotMask = 0x3FFu;
constexpr int kEntKindShift = 10, kEntFlagShift = 12, kEntZoneShift = 16;
constexpr uint32_t kEntHeat = 1u << 24;
constexpr uint32_t kNbUp = 1u, kNbDown = 2u, kNbLeft = 4u, kNbRight = 8u;
// counts3[plan*4 + i]
enum { kC3Pure = 0, kC3Gen = 1, kC3Pat = 2, kC3Capable = 3 };

__host__ inline bool resident3_supported(int H, int W) {
  if (H % 4 != 0 || W % 4 != 0 || H < 8 || W < 8) return false;
  const int th = H / 4, tw = W / 4;
  const int Tq = (tw % 2 == 0) ? tw + 1 : tw;
  return th * tw <= kR3MaxThreads && th * Tq <= (int)kEntSlotMask;
}

__host__ inline Resident3Geom resident3_geom(int H, int W, int Z) {
  Resident3Geom g;
  g.th = H / 4;
  g.tw = W / 4;
  g.Tq = (g.tw % 2 == 0) ? g.tw + 1 : g.tw;
  g.nt = (g.th * g.tw + 31) & ~31;
  g.tq_magic = 0xFFFFFFFFu / (unsigned)g.Tq + 1u;
  g.pat_cap = kR3PatCap;
  auto al = [](size_t v) { return (int)((v + 15) & ~(size_t)15); };
  auto al128 = [](size_t v) { return (int)((v + 127) & ~(size_t)127); };
  g.xarr = al128((size_t)g.th * g.Tq * 16);
  size_t o = 0;
  g.off_x0 = (int)o; o = al128(o + (size_t)4 * g.xarr);
  // the input / output plane of the TMA copies aliases the second exchange buffer
  const size_t x1 = (size_t)4 * g.xarr, plane = (size_t)H * W * 4;
  g.off_x1 = (int)o; o = al128(o + (x1 > plane ? x1 : plane));
  g.off_hdr = (int)o; o = al(o + header_bytes(Z));
  g.off_ptab = (int)o; o = al(o + (size_t)kR3PatCap * 96);
  g.off_bins = (int)o; o = al(o + (size_t)(Z + 1) * 8);
  g.off_misc = (int)o; o = al(o + 32);
  g.off_bar = (int)o; o = al(o + 16);
  g.total = (int)o;
  return g;
}

// ---------------------------------------------------------------------------
// Once per uploaded plan: classify the tiles, number the pair patterns, sort the tiles into the
// thread order.  One CTA per plan; positions by counting smaller keys (deterministic).
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t r3_zone_slot(uint32_t d, int Z) {
  if (desc_class(d) == SBX_CV_EXTERIOR) return 0xFFu;       // T == T_inf == the sums' reference: contributes 0
  const int zs = desc_zone(d);
  return zs == SBX_ZONE_NONE ? (uint32_t)Z : (uint32_t)zs;
}

__global__ void __launch_bounds__(kPrepThreads) k_prepare_plan3(const Params p) {
  __shared__ uint32_t key[kR3MaxThreads];
  __shared__ uint32_t key1[kR3MaxThreads];
  __shared__ uint32_t bitmap[32];     // pair pattern a * 32 + b present
  __shared__ uint32_t prefix[33];
  __shared__ int s_bad, s_npure, s_ngen;
  const int plan = blockIdx.x, tid = threadIdx.x;
  const Resident3Geom& g = p.g3;
  const int W = p.W, Z = p.Z, n_tiles = g.th * g.tw;
  const uint16_t* raw = p.desc + (size_t)plan * p.H * W;
  uint32_t* ent = p.ent3 + (size_t)plan * g.nt;
  uint4* ga = p.gen3a + (size_t)plan * g.nt;
  uint4* gb = p.gen3b + (size_t)plan * g.nt;
  uint4* gq = p.gen3q + (size_t)plan * g.nt;
  uint16_t* pat = p.pat3 + (size_t)plan * g.pat_cap;
  if (tid < 32) bitmap[tid] = 0;
  if (tid == 0) { s_bad = 0; s_npure = 0; s_ngen = 0; }
  __syncthreads();
  auto cv = [&](int it, int e) -> uint32_t {          // descriptor of CV e = 4 * row + col of tile it
    const int tr = it / g.tw, tc = it - tr * g.tw;
    return raw[(tr * 4 + (e >> 2)) * W + tc * 4 + (e & 3)];
  };
  // ---- kinds, groups, pattern bitmap ----
  for (int it = tid; it < n_tiles; it += kPrepThreads) {
    const int tr = it / g.tw, tc = it - tr * g.tw;
    bool pure = true;
    const uint32_t z0 = r3_zone_slot(cv(it, 0), Z);
    uint32_t hsh = 0;
    for (int e = 0; e < 16; ++e) {
      const uint32_t d = cv(it, e);
      pure = pure && ((d & 0x007Fu) == kFastDesc) && r3_zone_slot(d, Z) == z0;
      hsh = hsh * 31u + (uint32_t)combo_index(d);
    }
    pure = pure && tr > 0 && tr < g.th - 1 && tc > 0 && tc < g.tw - 1;
    bool heat = false;
    if (!pure)
      for (int e = 0; e < 16; e += 2) {
        const uint32_t k = (uint32_t)combo_index(cv(it, e)) * 32u + (uint32_t)combo_index(cv(it, e + 1));
        atomicOr(&bitmap[k >> 5], 1u << (k & 31u));
        heat = heat || ((cv(it, e) | cv(it, e + 1)) & SBX_DESC_DIFFUSER);
      }
    const int kind = pure ? kR3Pure : kR3Gen;
    // GEN tiles: those with heat input last (the heat lookup is a per-thread branch), similar
    // pattern sets next to each other (table reads of a warp then mostly broadcast)
    const uint32_t group = pure ? z0 : ((heat ? 16u : 0u) | ((hsh ^ (hsh >> 7) ^ (hsh >> 13)) & 15u));
    const int slot = tr * g.Tq + tc;
    key1[it] = ((uint32_t)kind << 11) | (group << 3) | (uint32_t)(slot & 7);
    atomicAdd(pure ? &s_npure : &s_ngen, 1);
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t o = 0;
    for (int w = 0; w < 32; ++w) { prefix[w] = o; o += (uint32_t)__popc(bitmap[w]); }
    prefix[32] = o;
    if ((int)o > g.pat_cap || (int)o * 6 > g.nt) s_bad = 1;      // the kernel builds the table one item per thread
  }
  // rank inside the (kind, group, residue) bucket
  for (int it = tid; it < n_tiles; it += kPrepThreads) {
    const uint32_t k1 = key1[it];
    uint32_t m = 0;
    for (int j = 0; j < it; ++j) m += (key1[j] == k1) ? 1u : 0u;
    key[it] = ((k1 >> 3) << 13) | (m << 3) | (k1 & 7u);          // kind | group | rank | residue
  }
  __syncthreads();
  const int n_pat = (int)prefix[32];
  auto pat_id = [&](uint32_t k) -> uint32_t {
    return prefix[k >> 5] + (uint32_t)__popc(bitmap[k >> 5] & ((1u << (k & 31u)) - 1u));
  };
  if (!s_bad)
    for (int k = tid; k < 1024; k += kPrepThreads)
      if (bitmap[k >> 5] & (1u << (k & 31))) pat[pat_id((uint32_t)k)] = (uint16_t)((k >> 5) | ((k & 31) << 8));
  // ---- thread order and records ----
  for (int it = tid; it < n_tiles; it += kPrepThreads) {
    const uint32_t k2 = key[it];
    int pos = 0;
    for (int j = 0; j < n_tiles; ++j) pos += (key[j] < k2) ? 1 : 0;
    const int tr = it / g.tw, tc = it - tr * g.tw;
    const int kind = (int)(key1[it] >> 11);
    const uint32_t flags = (tr > 0 ? kNbUp : 0u) | (tr < g.th - 1 ? kNbDown : 0u) | (tc > 0 ? kNbLeft : 0u) |
                           (tc < g.tw - 1 ? kNbRight : 0u);
    uint32_t e = (uint32_t)(tr * g.Tq + tc) | ((uint32_t)kind << kEntKindShift) | (flags << kEntFlagShift);
    if (kind == kR3Pure) {
      e |= r3_zone_slot(cv(it, 0), Z) << kEntZoneShift;
    } else {
      uint32_t pats[2] = {0u, 0u}, qs[4] = {0u, 0u, 0u, 0u};
      uint32_t zs[6] = {0u, 0u, 0u, 0u, 0u, 0u}, zm[6] = {0u, 0u, 0u, 0u, 0u, 0u};   // four rooms + walls meet in one tile
      int np = 0;
      bool heat = false, bad = false;
      for (int e2 = 0; e2 < 16; ++e2) {
        const uint32_t d = cv(it, e2);
        if ((e2 & 1) == 0) {
          const uint32_t k = (uint32_t)combo_index(d) * 32u + (uint32_t)combo_index(cv(it, e2 + 1));
          pats[e2 >> 3] |= (s_bad ? 0u : pat_id(k)) << (8 * ((e2 >> 1) & 3));
        }
        const bool dq = (d & SBX_DESC_DIFFUSER) != 0;
        heat = heat || dq;
        qs[e2 >> 2] |= (dq ? (uint32_t)desc_zone(d) : (uint32_t)Z) << (8 * (e2 & 3));
        const uint32_t z = r3_zone_slot(d, Z);
        if (z < (uint32_t)Z) {        // CVs outside every zone only enter the grid total
          int k = 0;
          while (k < np && zs[k] != z) ++k;
          if (k == np) {
            if (np == 6) { bad = true; continue; }
            zs[np++] = z;
          }
          zm[k] |= 1u << e2;
        }
      }
      if (bad) atomicExch(&s_bad, 1);
      if (heat) e |= kEntHeat;
      e |= (uint32_t)np << kEntZoneShift;
      ga[pos] = make_uint4(pats[0], pats[1], zs[0] | (zm[0] << 16), zs[1] | (zm[1] << 16));
      gb[pos] = make_uint4(zs[2] | (zm[2] << 16), zs[3] | (zm[3] << 16), zs[4] | (zm[4] << 16), zs[5] | (zm[5] << 16));
      gq[pos] = make_uint4(qs[0], qs[1], qs[2], qs[3]);
    }
    ent[pos] = e;
  }
  for (int i = n_tiles + tid; i < g.nt; i += kPrepThreads) ent[i] = (uint32_t)kR3Idle << kEntKindShift;
  __syncthreads();
  if (tid == 0) {
    int32_t* c = p.counts3 + (size_t)plan * 4;
    c[kC3Pure] = s_npure;
    c[kC3Gen] = s_ngen;
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

__device__ __forceinline__ float absmax2(float lmax, f32x2 d) {
  float a, b;
  unpack2(d, a, b);
  return fmaxf(fmaxf(lmax, fabsf(a)), fabsf(b));
}

// One row of a PURE tile.  pu / pc / pd: kq * T of the row above, of this row, of the row
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

__device__ __forceinline__ float pure_sweep(float4 (&T)[4], const f32x2 (&n3)[8], const float4* __restrict__ Xr,
                                            const int s, const int Tq, const int xq, const PureCoef& c) {
  const float4 U = Xr[xq + s - Tq];          // bottom row of the tile above
  const float4 D = Xr[s + Tq];               // top row of the tile below
  const float4 L = Xr[3 * xq + s - 1];       // right column of the tile to the left
  const float4 R = Xr[2 * xq + s + 1];       // left column of the tile to the right
  float pl[4], pr[4];
  unpack2(mul2(c.kq, pack2(L.x, L.y)), pl[0], pl[1]);
  unpack2(mul2(c.kq, pack2(L.z, L.w)), pl[2], pl[3]);
  unpack2(mul2(c.kq, pack2(R.x, R.y)), pr[0], pr[1]);
  unpack2(mul2(c.kq, pack2(R.z, R.w)), pr[2], pr[3]);
  f32x2 pua = mul2(c.kq, pack2(U.x, U.y)), pub = mul2(c.kq, pack2(U.z, U.w));
  f32x2 pca = mul2(c.kq, pack2(T[0].x, T[0].y)), pcb = mul2(c.kq, pack2(T[0].z, T[0].w));
  float lmax = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 dn = i < 3 ? T[i + 1] : D;
    const f32x2 pda = mul2(c.kq, pack2(dn.x, dn.y)), pdb = mul2(c.kq, pack2(dn.z, dn.w));
    lmax = pure_row(T[i], pua, pub, pca, pcb, pda, pdb, pl[i], pr[i], n3[2 * i], n3[2 * i + 1], c, lmax);
    pua = pca; pub = pcb;
    pca = pda; pcb = pdb;
  }
  return lmax;
}

// One horizontal pair of CVs of a GEN tile (cv_update_packed on two CVs at once): coefficient
// pairs from the pattern table entry `e` (6 x float4, see the table build in the kernel).
//   tjp / tjm: T(i,j+1) / T(i,j-1) of the two CVs, tip / tim: T(i+1,j) / T(i-1,j)
__device__ __forceinline__ f32x2 gen_pair(const float4* __restrict__ e, const f32x2 tjp, const f32x2 tjm,
                                          const f32x2 tip, const f32x2 tim, const f32x2 n3, const f32x2 q,
                                          const bool heat, const f32x2 one, const uint32_t tinf_bits) {
  const float4 k13 = e[0], k24 = e[1], hhv = e[2], den = e[3], vu = e[4], cme = e[5];
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
  const uint32_t m0 = __float_as_uint(cme.z), m1 = __float_as_uint(cme.w);
  o0 = __uint_as_float((__float_as_uint(o0) & ~m0) | (tinf_bits & m0));
  o1 = __uint_as_float((__float_as_uint(o1) & ~m1) | (tinf_bits & m1));
  return pack2(o0, o1);
}

__device__ __forceinline__ float gen_sweep(float4 (&T)[4], const f32x2 (&n3)[8], const float4* __restrict__ Xr,
                                           const int s, const int Tq, const int xq, const uint32_t flags,
                                           const uint32_t pat_lo, const uint32_t pat_hi, const bool heat,
                                           const uint4 gq, const float* __restrict__ qcv,
                                           const float4* __restrict__ ptab, const f32x2 one, const float t_inf) {
  const float4 inf4 = make_float4(t_inf, t_inf, t_inf, t_inf);
  const float4 U = (flags & kNbUp) ? Xr[xq + s - Tq] : inf4;         // tf_simulator.py:642-644
  const float4 D = (flags & kNbDown) ? Xr[s + Tq] : inf4;            // :646
  const float4 L = (flags & kNbLeft) ? Xr[3 * xq + s - 1] : inf4;    // :638-640
  const float4 R = (flags & kNbRight) ? Xr[2 * xq + s + 1] : inf4;   // :636
  const float lf[4] = {L.x, L.y, L.z, L.w}, rf[4] = {R.x, R.y, R.z, R.w};
  const uint32_t qw[4] = {gq.x, gq.y, gq.z, gq.w};
  const uint32_t tinf_bits = __float_as_uint(t_inf);
  float4 up = U;
  float lmax = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 cur = T[i];
    const float4 dn = i < 3 ? T[i + 1] : D;
    const uint32_t pw = i < 2 ? pat_lo : pat_hi;
    const uint32_t pa = (pw >> (16 * (i & 1))) & 0xFFu, pb = (pw >> (16 * (i & 1) + 8)) & 0xFFu;
    f32x2 qa = 0ull, qb = 0ull;
    if (heat) {
      const uint32_t w = qw[i];
      qa = pack2(qcv[w & 0xFFu], qcv[(w >> 8) & 0xFFu]);
      qb = pack2(qcv[(w >> 16) & 0xFFu], qcv[w >> 24]);
    }
    const f32x2 mid = pack2(cur.y, cur.z);
    const f32x2 oa = gen_pair(ptab + pa * 6, mid, pack2(lf[i], cur.x), pack2(dn.x, dn.y), pack2(up.x, up.y),
                              n3[2 * i], qa, heat, one, tinf_bits);
    const f32x2 ob = gen_pair(ptab + pb * 6, pack2(cur.w, rf[i]), mid, pack2(dn.z, dn.w), pack2(up.z, up.w),
                              n3[2 * i + 1], qb, heat, one, tinf_bits);
    lmax = absmax2(lmax, sub2(oa, pack2(cur.x, cur.y)));               // :851-853
    lmax = absmax2(lmax, sub2(ob, pack2(cur.z, cur.w)));
    unpack2(oa, T[i].x, T[i].y);
    unpack2(ob, T[i].z, T[i].w);
    up = cur;
  }
  return lmax;
}

// exact segmented warp sum: lanes with the same `slot` add their `v`; one lane per slot
// adds the total to the split lo / hi bins (32-bit shared-memory atomics)
__device__ __forceinline__ void r3_zone_add(uint32_t* bins, int Z, const int slot, const int v, const bool valid,
                                            const int lane) {
  unsigned todo = __ballot_sync(0xffffffffu, valid);
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

__global__ void __launch_bounds__(kR3MaxThreads, 2) k_resident_step3(const Params p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31;
  const Resident3Geom& G = p.g3;
  const int b = p.b_begin + blockIdx.x;
  const int plan = p.n_plans == 1 ? 0 : b;
  const int Z = p.Z, W = p.W, n_cv = p.H * W;
  const int Tq = G.Tq, xq = G.xarr / 16;
  float4* X0 = reinterpret_cast<float4*>(smem + G.off_x0);
  float4* X1 = reinterpret_cast<float4*>(smem + G.off_x1);
  float* plane = reinterpret_cast<float*>(smem + G.off_x1);
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

  // ---- stage 0: bulk loads (plane + header), static per-thread data straight from global ----
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_expect_tx(bar, (uint32_t)(n_cv * 4) + hdr_bytes);
    tma_load_1d(plane, gT, (uint32_t)(n_cv * 4), bar);
    tma_load_1d(smem + G.off_hdr, p.hdr + (size_t)b * hdr_bytes, hdr_bytes, bar);
    misc[0] = 0u;
  }
  const uint32_t ent = __ldg(p.ent3 + (size_t)plan * G.nt + tid);
  const int4 cnt = __ldg(reinterpret_cast<const int4*>(p.counts3) + plan);     // n_pure, n_gen, n_pat, capable
  const int n_pat = cnt.z;
  const int kind = (int)((ent >> kEntKindShift) & 3u);
  const bool heat = (ent & kEntHeat) != 0u;
  // this thread's item of the pattern table build (stage 1): the pattern's two combo indices
  const int pt_i = tid / 6, pt_j = tid - pt_i * 6;
  const uint32_t pt_pr = pt_i < G.pat_cap ? (uint32_t)__ldg(p.pat3 + (size_t)plan * G.pat_cap + pt_i) : 0u;
  uint4 ga = make_uint4(0u, 0u, 0u, 0u), gq = make_uint4(0u, 0u, 0u, 0u);
  if (kind == kR3Gen) {
    ga = __ldg(p.gen3a + (size_t)plan * G.nt + tid);
    if (heat) gq = __ldg(p.gen3q + (size_t)plan * G.nt + tid);
  }
  // The CTA that takes over this SM slot works on building b + (CTAs in flight): pull its
  // inputs into L2 now
  if (p.prefetch_dist > 0 && tid == 32 && b + p.prefetch_dist < p.b_end) {
    const int bn = b + p.prefetch_dist;
    l2_prefetch(p.tbuf[0] + (size_t)bn * n_cv, (uint32_t)(n_cv * 4));
    l2_prefetch(p.hdr + (size_t)bn * hdr_bytes, hdr_bytes & ~15u);
    if (p.n_plans != 1) {
      l2_prefetch(p.ent3 + (size_t)bn * G.nt, (uint32_t)(G.nt * 4));
      // (this plan's list sizes stand in for the next one's: a prefetch hint)
      if (cnt.y > 0) l2_prefetch(p.gen3a + (size_t)bn * G.nt + cnt.x, (uint32_t)(cnt.y * 16));
    }
  }
  for (int i = tid; i < 2 * (Z + 1); i += (int)blockDim.x) bins[i] = 0u;
  __syncthreads();
  mbar_wait(bar, 0);
  SBX_PHASE(0);   // launch .. bulk loads and per-thread list entries have arrived
  const float t_inf = scal[0];

  // ---- stage 1: pair-pattern table of this building (coefficients change with h and T_inf) ----
  if (pt_i < n_pat) {        // kR3PatCap * 6 <= kR3MaxThreads: one item per thread
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
    ptab[tid] = v;
  }

  // ---- stage 2: the thread's tile -> registers, rim -> exchange buffer 0 ----
  const int s = (int)(ent & kEntSlotMask);
  const int tr = (int)__umulhi((unsigned)s, G.tq_magic);
  const int tc = s - tr * Tq;
  const uint32_t flags = (ent >> kEntFlagShift) & 15u;
  float* tile = plane + (tr * 4) * W + tc * 4;
  float4 T[4];
  f32x2 n3[8];
  PureCoef pc;
  const float one = p.one;
  pc.one = pack2(one, one);
  {
    const Combo& c = tab[SBX_CV_INTERIOR * kNumMaterials + 0];
    pc.kq = pack2(c.k1, c.k1); pc.vz = pack2(c.vz, c.vz);
    pc.nden = pack2(-c.den, -c.den); pc.rden = pack2(c.rden, c.rden);
  }
  if (kind != kR3Idle) {
#pragma unroll
    for (int i = 0; i < 4; ++i) T[i] = *reinterpret_cast<const float4*>(tile + i * W);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) T[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();      // the pattern table is complete (and every thread has its tile: the plane is free)
  SBX_PHASE(1);   // pattern table, tile -> registers
  {
    // n3 = (cm * T_prev) / dt (tf_simulator.py:743-749), kept for every sweep of this step
    const float rdt = __frcp_rn(p.dt);
    const f32x2 ndt2 = pack2(-p.dt, -p.dt), rdt2 = pack2(rdt, rdt);
    if (kind == kR3Gen) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t pw = i < 2 ? ga.x : ga.y;
        const uint32_t pa = (pw >> (16 * (i & 1))) & 0xFFu, pb = (pw >> (16 * (i & 1) + 8)) & 0xFFu;
        const float4 ea = ptab[pa * 6 + 5], eb = ptab[pb * 6 + 5];
        n3[2 * i] = div_rn2(mul2(pack2(ea.x, ea.y), pack2(T[i].x, T[i].y)), ndt2, rdt2);
        n3[2 * i + 1] = div_rn2(mul2(pack2(eb.x, eb.y), pack2(T[i].z, T[i].w)), ndt2, rdt2);
      }
    } else {
      const float cmf = tab[SBX_CV_INTERIOR * kNumMaterials + 0].cm;
      const f32x2 cm2 = pack2(cmf, cmf);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        n3[2 * i] = div_rn2(mul2(cm2, pack2(T[i].x, T[i].y)), ndt2, rdt2);
        n3[2 * i + 1] = div_rn2(mul2(cm2, pack2(T[i].z, T[i].w)), ndt2, rdt2);
      }
    }
  }
  auto publish = [&](float4* X) {
    X[s] = T[0];
    X[xq + s] = T[3];
    X[2 * xq + s] = make_float4(T[0].x, T[1].x, T[2].x, T[3].x);
    X[3 * xq + s] = make_float4(T[0].w, T[1].w, T[2].w, T[3].w);
  };
  if (kind != kR3Idle) publish(X0);
  __syncthreads();
  SBX_PHASE(2);   // thermal-mass terms, first rim exchange

  // ---- stage 3: Jacobi sweeps to convergence (simulator.py:348-364) ----
  const int limit = p.iteration_limit;
  const float thr = p.threshold;
  float4* Xr = X0;
  float4* Xw = X1;
  int k = 0;
  float lmax = 0.f;
  while (k < limit) {
    ++k;
    lmax = 0.f;
    if (kind == kR3Pure) {
      lmax = pure_sweep(T, n3, Xr, s, Tq, xq, pc);
      publish(Xw);
    } else if (kind == kR3Gen) {
      lmax = gen_sweep(T, n3, Xr, s, Tq, xq, flags, ga.x, ga.y, heat, gq, qcv, ptab, pc.one, t_inf);
      publish(Xw);
    }
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
    float4* tmp = Xr; Xr = Xw; Xw = tmp;
    if (k == 1) SBX_PHASE(3); else SBX_PHASE(4);   // first sweep / later sweeps
    if (!above) break;
  }

  // ---- stage 4: write back (plane aliases exchange buffer 1: nobody reads it any more) ----
  if (kind != kR3Idle) {
#pragma unroll
    for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(tile + i * W) = T[i];
  }
  tma_store_fence();
  {
    const uint32_t wm = __reduce_max_sync(0xffffffffu, __float_as_uint(lmax));
    if (lane == 0) atomicMax(&misc[0], wm);
  }
  // ---- stage 5: zone / grid sums from the registers ----
  if (sums) {
    const f32x2 scale2 = pack2(kFixScaleF, kFixScaleF);
    const float nref = -__fmul_rn(t_inf, kFixScaleF);
    const f32x2 nref2 = pack2(nref, nref);
    int f[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float a0, a1, a2, a3;
      unpack2(fma2(pack2(T[i].x, T[i].y), scale2, nref2), a0, a1);      // (T - ref) * 2^16, exact (to_fix32)
      unpack2(fma2(pack2(T[i].z, T[i].w), scale2, nref2), a2, a3);
      f[4 * i] = __float2int_rn(a0); f[4 * i + 1] = __float2int_rn(a1);
      f[4 * i + 2] = __float2int_rn(a2); f[4 * i + 3] = __float2int_rn(a3);
    }
    int total = 0;
#pragma unroll
    for (int e = 0; e < 16; ++e) total += f[e];
    // whole grid (slot Z): every CV of every tile (exterior space holds T_inf, the reference: 0)
    const unsigned lo = __reduce_add_sync(0xffffffffu, (unsigned)total & 0xFFFFu);
    const int hi = __reduce_add_sync(0xffffffffu, total >> 16);
    const int zone = (int)((ent >> kEntZoneShift) & 0xFFu);
    const int zone0 = __shfl_sync(0xffffffffu, zone, 0);
    if (__all_sync(0xffffffffu, kind == kR3Pure && zone == zone0)) {
      // the common PURE warp: one zone, the sums are the ones just taken
      if (lane == 0) {
        atomicAdd(&bins[Z], lo);
        atomicAdd(reinterpret_cast<int*>(&bins[Z + 1 + Z]), hi);
        atomicAdd(&bins[zone0], lo);
        atomicAdd(reinterpret_cast<int*>(&bins[Z + 1 + zone0]), hi);
      }
    } else {
      if (lane == 0) {
        atomicAdd(&bins[Z], lo);
        atomicAdd(reinterpret_cast<int*>(&bins[Z + 1 + Z]), hi);
      }
      r3_zone_add(bins, Z, zone, total, kind == kR3Pure, lane);      // PURE lanes of a mixed warp
      if (kind == kR3Gen) {
        // GEN tiles: zones are scattered over the warp, so each thread adds its own parts
        const int np = zone;
        uint4 gb = make_uint4(0u, 0u, 0u, 0u);
        if (np > 2) gb = __ldg(p.gen3b + (size_t)plan * G.nt + tid);
        for (int k2 = 0; k2 < np; ++k2) {
          const uint32_t part = k2 == 0 ? ga.z : k2 == 1 ? ga.w : k2 == 2 ? gb.x : k2 == 3 ? gb.y : k2 == 4 ? gb.z : gb.w;
          int sv = 0;
#pragma unroll
          for (int e = 0; e < 16; ++e) sv += (part & (0x10000u << e)) ? f[e] : 0;
          const int z = (int)(part & 0xFFu);
          atomicAdd(&bins[z], (unsigned)sv & 0xFFFFu);
          atomicAdd(reinterpret_cast<int*>(&bins[Z + 1 + z]), sv >> 16);
        }
      }
    }
  }
  __syncthreads();      // plane complete (TMA store), bins complete
  SBX_PHASE(5);   // write-back to the plane, zone sums
  if (tid == 0) {
    tma_store_1d(gT, plane, (uint32_t)(n_cv * 4));
    tma_store_commit();
  }
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
      tma_store_wait();
    }
  }
  SBX_PHASE(6);   // results out, TMA store drained
}

}  // namespace sbx
