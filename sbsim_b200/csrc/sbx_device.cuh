// Device-side building blocks of libsbx: the per-CV Jacobi update and the
// per-building HVAC / observation / reward algebra.
//
// Arithmetic contract
//   * The diffusion update reproduces TFSimulator.update_temperature_estimates
//     (/root/reference/smart_control/simulator/tf_simulator.py:757-853) with the
//     SAME fp32 operation order, one IEEE rounding per op, no FMA contraction
//     (__fmul_rn/__fadd_rn/__fdiv_rn and -fmad=false), so the temperature field
//     is bit-identical to the reference's eager-TF arithmetic.  Terms that are
//     exactly +0 for a class (0 conductivity / 0 convection on a side) are
//     skipped only where x + 0 == x makes that exact.
//   * Device scalars (thermostat, VAV, AHU, boiler, reward, normalisation) are
//     evaluated in fp64 like the reference's Python floats, and rounded to fp32
//     at the points where the reference stores them in proto `float` fields.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sbx.h"

namespace sbx {

constexpr int kNumMaterials = 3;
constexpr int kNumCombos = SBX_CV_NUM_CLASSES * kNumMaterials;
constexpr int kMaxZones = 255;

// thermostat modes (thermostat.py:63-66)
constexpr int kOff = 0, kHeat = 1, kCool = 2, kPassiveCool = 3;

constexpr double kAirHeatCapacity = 1006.0;    // utils/constants.py:21
constexpr double kWaterHeatCapacity = 4180.0;  // utils/constants.py:22
constexpr double kWaterDensity = 1000.0;       // utils/constants.py:46
constexpr double kGravity = 9.8;               // utils/constants.py:47

// Coefficients of one (class, material) combination; tf_simulator.py:791-798,
// 669-690.  k1 pairs with T(i,j+1), k3 with T(i,j-1) (shift semantics of
// tf_simulator.py:459-472, 636-640), k2 with T(i+1,j), k4 with T(i-1,j).
struct __align__(16) Combo {
  float k1, k3, k2, k4;
  // T_inf*h of the ambient-facing horizontal side (left OR right: no class has
  // both, so (x + hl) + hr == x + hh exactly) and of the vertical side
  float hh, hv, den, rden;       // rden = RN(1/den)
  float vz, uz, cm, pad1;        // cm = ((((rho*U)*V)*c)*z)*c
};
// packed descriptor bits (k_prepare_plan): combo index | half-U | diffuser | half-V | zone
constexpr uint32_t kPackIdxMask = 31u, kPackHalfU = 0x20u, kPackHalfV = 0x80u;

struct ResidentGeom {
  int P, Pq;            // row pitch in CVs / in vectors
  int plane_cv;         // H * P
  unsigned pq_magic;    // slot / Pq == umulhi(slot, pq_magic) for slot < 2^16
  int rl_cap;           // zone-sum list capacity (entries)
  int use_tmap;         // temperature plane moved by one tensor-map TMA op (else one bulk copy per row)
  int desc_stride;      // global stride of a plan's packed descriptor plane (u16), 16 B multiple
  int list_stride;      // global stride of a plan's vector list (u16), 16 B multiple
  int off_a, off_b, off_n3, off_desc, off_list, off_rlist, off_hdr, off_bins, off_wmax, off_bar, total;
};

// Shared-memory layout and per-plan list capacities of k_resident_step2 (sbx_resident2.cuh)
struct Resident2Geom {
  int rec_cap;       // 16-byte records (non-FAST, non-EXT vectors) per plan
  int zfull_cap;     // u16 entries of the zone-sum list's single-zone section
  int zpart_cap;     // u32 entries of its masked section
  int zchunk_cap;    // u8 zone slot per 32-entry chunk
  int off_p0, off_p1, off_p2, off_flist, off_sched, off_rec, off_zfull, off_zpart, off_zchunk, off_hdr,
      off_bins, off_misc, off_bar, total;
};

// Shared-memory layout and per-plan list strides of k_resident_step3 (sbx_resident3.cuh):
// HALF-TILES of 2 x 4 control volumes, one or two per thread, temperatures in registers across
// the sweeps
struct Resident3Geom {
  int hh, tw;          // half-tiles per column / per row (H/2, W/4)
  int Tq;              // pitch of the exchange arrays in half-tiles (odd: the 16-byte bank group rotates row to row)
  int nt;              // threads per CTA
  unsigned tq_magic;   // slot / Tq == umulhi(slot, tq_magic)
  int pat_cap;         // pair patterns per plan
  int xarr;            // bytes of one exchange array (hh * Tq float4, 128 B multiple)
  int plane_pitch;     // row pitch (floats) of the TMA-loaded input plane (ResidentGeom::P)
  int off_x0, off_x1, off_hdr, off_ptab, off_mtab, off_bins, off_zparts, off_misc, off_bar, total;
};

// Everything a kernel needs; passed by value.
constexpr int kPwMetaInts = 40;            // per plan: {L, I, n_levels, level_start[0 .. n_levels]}
constexpr uint32_t kPwGridFlag = 0x80000000u;

struct Params {
  ResidentGeom geom;
  Resident2Geom g2;
  Resident3Geom g3;
  // sizes
  int B, H, W, Z, n_plans, n_weather, n_reset, n_occ_zones, T_rows;
  int obs_mode, D, n_actions;
  int action_target[SBX_MAX_ACTIONS];
  float action_min[SBX_MAX_ACTIONS], action_range[SBX_MAX_ACTIONS];
  int n_hist_bins[3];
  // physics
  float dt, z, threshold;
  int iteration_limit;
  // schedule
  double comfort_heat, comfort_cool, eco_heat, eco_cool;
  // devices
  double ahu_r, ahu_init_heat, ahu_init_cool, ahu_dp, ahu_eff, ahu_max_flow;
  double boiler_init_sp, boiler_head, boiler_eff, boiler_heat_rate, boiler_cool_rate;
  double boiler_diss_factor;  // L*2*pi / (ln(r2/r1)/k_ins + 1/(h*r2))  boiler.py:309-320
  double boiler_capacity;
  double vav_max_flow, vav_max_reheat;
  // reward
  double pmax, pmin, emax, gmax, delta, stiff, wu, wv, ww, gas_carbon;
  double carbon_cost_factor, reward_shift, reward_scale;
  int reward_kind;
  double discount, occ_norm;
  // static per plan
  const uint16_t* desc;      // [P,H,W]
  const double* material;    // [P,3,3]
  const double* cv_size;     // [P]
  const int32_t* zone_ncv;   // [P,Z]
  const int32_t* zone_ndiff; // [P,Z]
  const int32_t* obs_zone_order;  // [P,Z]
  uint16_t* desc_packed;     // [P,H,W] combo index | diffuser | zone (k_prepare_plan)
  uint16_t* desc_spk;        // [P,H,W] the same packing at the natural pitch W (streaming path)
  uint16_t* qlist;           // [P,H*W/V] fast vectors from the front, slow from the back
  int32_t* n_fast;           // [P,4] sizes of the FAST / MEDIUM / EXT / SLOW vector lists
  uint32_t* rlist;           // [P, rl_cap] zone-sum list (k_prepare_reduce)
  int32_t* rl_chunks;        // [P] warps' worth of entries in it, -1 = does not fit
  unsigned char* hdr;        // [B, header_bytes(Z)] per-building solve header (k_build_header)
  // k_resident_step2 (k_prepare_plan2): per-plan lists
  uint16_t* flist2;          // [P, list_stride] FAST slots then EXT slots
  uint32_t* sched2;          // [P, r2_sched_cap] per-warp work schedule of the sweeps
  uint4* rec2;               // [P, rec_cap] MEDIUM (no heat / heat) then SLOW (no heat / heat) records
  int32_t* counts2;          // [P, 8] n_fast, n_ext, n_heads, sched words, -, n_rec, n_full_chunks, n_part_chunks (-1: a list overflowed)
  uint16_t* zfull;           // [P, zfull_cap]
  uint32_t* zpart;           // [P, zpart_cap]
  uint8_t* zchunk;           // [P, zchunk_cap]
  // k_resident_step3 (k_prepare_plan3): per-plan half-tile schedule; index t = first half-tile of
  // thread t, nt + t = its second one
  uint32_t* ent3;            // [P, 2*nt] slot | kind << 12 | neighbour flags << 14 | zone (PURE) or parts << 18 | heat << 26
  uint4* rec3;               // [P, 2*nt] non-PURE half-tiles: {4 pair patterns, heat slots row 0, row 1, -}
  uint2* recz3;              // [P, 2*nt] their zone parts 0-3 (zone | CV mask << 8, two per word)
  uint16_t* pat3;            // [P, pat_cap] pair patterns of the boundary half-tiles: combo a | combo b << 8
  int32_t* counts3;          // [P, 4] -, -, n_pat, capable (0: a capacity was exceeded)
  float one;                 // 1.0f, opaque to the compiler (see pair_add in sbx_resident3.cuh)
  const float* reset_temps;  // [n_reset,H,W]
  const float* initial_temp; // [B]
  // tables
  const double* ambient;     // [Wn,T]
  const double* convection;  // [Wn]
  const uint8_t* comfort;    // [T]
  const uint8_t* comfort_soon;
  const double* occ_reward;  // [T,Zo]
  const int32_t* occ_obs;    // [T]
  const double* occ_obs_zone;  // [T,Zo]
  int occ_obs_per_env;       // occ_obs_zone was uploaded: count the building's own zones
  const double* price_e;
  const double* carbon_e;
  const double* price_g;
  const double* time_feat;   // [T,4]
  const double* obs_mean;    // [15]
  const double* obs_sd;      // [15] sqrt(variance), taken once on the host (IEEE sqrt: the same double)
  const double* obs_var;     // [15]
  const double* hist_bins;   // [3,SBX_MAX_HIST_BINS]
  // state
  float* tbuf[3];
  uint8_t* cur;              // [B] which tbuf holds building.temp
  float* zone_mean;          // [B,Z]
  float* global_mean;        // [B]
  float* qcv;                // [B,Z] heat per diffuser CV used by THIS step's solve
  float* qcv_next;           // [B,Z] produced by this step's VAV outputs
  double* qcv64;             // same two, fp64 (Gauss-Seidel solver reads fp64 input_q)
  double* qcv64_next;
  double* temp64;            // [B,H,W] fp64 building.temp, Gauss-Seidel solver only
  double* temp64_prev;       // [B,H,W] previous-step plane of the global-memory Gauss-Seidel wavefront (large grids)
  double threshold64, dt_double, z_double;
  uint8_t* therm_mode;       // [B,Z]
  double* ahu_heat_sp;
  double* ahu_cool_sp;
  double* boiler_sp;
  double* boiler_tank;       // [B,3]
  // transients / diagnostics
  float* pre_zone_mean;      // [B,Z]
  double* diag;              // [B,SBX_DIAG_N]
  double* q_zone;            // [B,Z]
  double* zone_supply;       // [B,Z]
  int32_t* n_sweeps;         // [B]
  uint32_t* max_delta_bits;  // [B] fp32 bits of the running max (non-negative => uint order)
  float* max_delta;          // [B] last completed sweep's max
  long long* zone_sum;       // [B,Z+1] integer sums of round((T - zone_ref[b]) * 2^16); slot Z = whole grid
  float* zone_ref;           // [B] reference temperature of those sums
  // SBX_OPT_NUMPY_MEANS (sbx_pairwise.cuh): means in NumPy's pairwise summation order
  int pw_on, pw_capL, pw_capI, pw_cap;
  int pw_fused;              // the tables fit the idle plane of k_resident_step: it takes the means itself
  uint32_t pw_zbytes;        // bytes of a plan's CV list that hold zone CVs (largest plan, 16-byte multiple)
  const void* pw_zlist;      // [P,H*W] u16 / u32: CVs of zone 0, zone 1, .. in raster order
  const uint2* pw_leaf;      // [P,capL] {start (| kPwGridFlag), length}
  const uint2* pw_node;      // [P,capI] {left, right} value indices, grouped by height
  const int32_t* pw_meta;    // [P,kPwMetaInts]
  const int2* pw_root;       // [P,Z+1] {value index of the tree's root or -1, CVs}
  float* pw_val;             // [B,pw_cap] leaf sums, then inner nodes
  float* pw_mean;            // [B,Z+1] zone means, slot Z = grid mean
  uint8_t* active;           // [B]
  int32_t* n_active;         // [1]
  unsigned long long* sweeps_total;  // [1]
  int32_t* sweep_k;          // [1] sweep index of the device-driven streaming loop (1-based)
  unsigned long long* sweep_launches;  // [1] k_sweep launches made by that loop since create
  unsigned long long* phase_cycles;  // [8] only with -DSBX_PROFILE_PHASES
  const int32_t* conv_perm;  // [B,H*W] or null: convection gather map for this step
  // device-RNG convection (sbx_set_device_convection): p == 0 -> off
  double conv_p;
  int conv_distance;
  unsigned long long conv_seed;
  // sbx_fd_step: solve only, ambient / convection given per env
  int fd_only;
  const double* fd_ambient;     // [B]
  const double* fd_convection;  // [B]
  // per call
  int b_begin, b_end;        // buildings [b_begin, b_end) are this launch's share of the batch
  int build_hdr;             // k_pre also builds the resident solve header of its buildings
  int v2_stride;             // k_resident_step2: CTAs of the launch (building b's CTA takes b + v2_stride next)
  int prefetch_dist;         // resident solve: CTA b prefetches building b + prefetch_dist into L2 (0 = off)
  int time_index;            // s: this step simulates [t_s, t_s + dt)
  int step_count;
  int episode_steps;
  int therm_seen, prev_comfort;
  const float* action;       // [B,A]
  float* obs;                // [B,D]
  float* reward;
  int32_t* step_type;
  float* discount_out;
};

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// Correctly rounded a / b for a divisor whose correctly rounded reciprocal
// y = RN(1/b) is known (Markstein's FMA iteration).  q0 = RN(a*y) is within
// 2 ulp of a/b; each step q <- RN(q + (a - b*q) * y) uses the EXACT residual
// (FMA) and shrinks the error by 2^-24, so q1 is a faithful rounding and q2 is
// the IEEE-754 round-to-nearest quotient (no overflow / underflow / zero divisor
// in this kernel's value ranges).  5 issue slots instead of the ~13 of the
// generic division sequence (MUFU.RCP + Newton + FCHK + slow-path branch).
// Verified against IEEE division: tests/test_division.py (CPU, 1e8 samples) and
// the bit-exact GPU parity tests.
__device__ __forceinline__ float div_rn(float a, float b, float y) {
  float q = __fmul_rn(a, y);
  float r = __fmaf_rn(-q, b, a);
  q = __fmaf_rn(r, y, q);
  r = __fmaf_rn(-q, b, a);
  return __fmaf_rn(r, y, q);
}

__device__ __forceinline__ int desc_class(uint32_t d) { return d & SBX_DESC_CLASS_MASK; }
__device__ __forceinline__ int desc_material(uint32_t d) { return (d >> SBX_DESC_MATERIAL_SHIFT) & 3; }
__device__ __forceinline__ int desc_zone(uint32_t d) { return (d >> SBX_DESC_ZONE_SHIFT) & 0xFF; }

// Builds the coefficient set of one (class, material); every line mirrors one
// TF op of tf_simulator.py (line numbers in comments).
__device__ inline Combo make_combo(int cls, const double* mat /*k,c,rho*/, double dx,
                                   float z, float dt, float h, float t_inf) {
  // get_cv_dimension_tensors :304-321 -- fp64 product stored to an fp32 array
  double hs = 1.0, vs = 1.0;
  bool sl = false, sr = false, st = false, sb = false;
  switch (cls) {
    case SBX_CV_EDGE_TOP: vs = 0.5; st = true; break;
    case SBX_CV_EDGE_BOTTOM: vs = 0.5; sb = true; break;
    case SBX_CV_EDGE_LEFT: hs = 0.5; sl = true; break;
    case SBX_CV_EDGE_RIGHT: hs = 0.5; sr = true; break;
    case SBX_CV_CORNER_TL: hs = vs = 0.5; st = sl = true; break;
    case SBX_CV_CORNER_TR: hs = vs = 0.5; st = sr = true; break;
    case SBX_CV_CORNER_BL: hs = vs = 0.5; sb = sl = true; break;
    case SBX_CV_CORNER_BR: hs = vs = 0.5; sb = sr = true; break;
    default: break;
  }
  const float u = (float)(dx * hs);
  const float v = (float)(dx * vs);
  const float kf = (float)mat[0];
  const float c = (float)mat[1];
  const float rho = (float)mat[2];
  const float kl = sl ? 0.f : kf, kr = sr ? 0.f : kf;   // :421-450
  const float kt = st ? 0.f : kf, kb = sb ? 0.f : kf;
  const float hl = sl ? h : 0.f, hr = sr ? h : 0.f;     // :354-392
  const float ht = st ? h : 0.f, hb = sb ? h : 0.f;
  Combo o;
  o.uz = mul(z, u);                                     // :791
  o.vz = mul(z, v);                                     // :792
  o.k1 = fdiv(kl, u);                                   // :795
  o.k3 = fdiv(kr, u);                                   // :796
  o.k2 = fdiv(kb, v);                                   // :797
  o.k4 = fdiv(kt, v);                                   // :798
  float d1 = add(o.k1, o.k3);                           // :669
  d1 = add(d1, hl);                                     // :670
  d1 = add(d1, hr);                                     // :671
  d1 = mul(o.vz, d1);                                   // :672
  float d2 = add(o.k2, o.k4);                           // :675
  d2 = add(d2, hb);                                     // :676
  d2 = add(d2, ht);                                     // :677
  d2 = mul(o.uz, d2);                                   // :678
  float d3 = mul(rho, u);                               // :681
  d3 = mul(d3, v);                                      // :682
  d3 = mul(d3, c);                                      // :683
  d3 = mul(z, d3);                                      // :684
  d3 = mul(d3, c);                                      // :685
  o.cm = d3;
  d3 = fdiv(d3, dt);                                    // :686
  o.den = add(add(d1, d2), d3);                         // :689-690
  // :725-728; exactly one of (hl, hr) and one of (ht, hb) can be non-zero
  o.hh = sl ? mul(t_inf, hl) : mul(t_inf, hr);
  o.hv = sb ? mul(t_inf, hb) : mul(t_inf, ht);
  o.rden = __frcp_rn(o.den);
  o.pad1 = 0.f;
  return o;
}

// Fills `tab[kNumCombos]` (shared memory) cooperatively.
__device__ inline void build_combo_table(Combo* tab, const Params& p, int plan, int b,
                                         float h, float t_inf, int tid, int nthreads) {
  const double* mat = p.material + (size_t)plan * 9;
  const double dx = p.cv_size[plan];
  for (int i = tid; i < kNumCombos; i += nthreads) {
    const int cls = i / kNumMaterials, m = i % kNumMaterials;
    tab[i] = make_combo(cls, mat + m * 3, dx, p.z, p.dt, h, t_inf);
  }
}

__device__ __forceinline__ int combo_index(uint32_t d) {
  return desc_class(d) * kNumMaterials + desc_material(d);
}

// One CV update (tf_simulator.py:719-754, 843, 847-849), branch-free: every
// class runs the same instruction stream with its own coefficients, so warps that
// mix interior / wall / boundary / exterior CVs do not diverge.  Two 128-bit
// shared-memory loads fetch {k1,k3,k2,k4} and {hh,hv,den,rden}; the face areas
// vz / uz take one of two per-building values (full or half cell) selected by a
// descriptor bit; exterior CVs (combo index < 3) are overridden with T_inf.
//   t_jp = T(i,j+1), t_jm = T(i,j-1), t_im = T(i-1,j), t_ip = T(i+1,j)
//   n3 = ((cm * T_prev) / dt)   precomputed (:743-749)
//   q  = input_q at this CV (0 unless it is a diffuser)
struct AreaCoef {
  float full, half;   // z * dx, z * (dx / 2)   (:791-792)
};
__device__ __forceinline__ float cv_update_packed(uint32_t d, float t_jp, float t_jm, float t_im,
                                                  float t_ip, float n3, float q, float t_inf,
                                                  const AreaCoef& az, const Combo* tab) {
  const int idx = (int)(d & kPackIdxMask);
  const float4* c4 = reinterpret_cast<const float4*>(tab + idx);
  const float4 k = c4[0], h = c4[1];
  const float vz = (d & kPackHalfV) ? az.half : az.full;
  const float uz = (d & kPackHalfU) ? az.half : az.full;
  float n1 = add(mul(k.x, t_jp), mul(k.y, t_jm));       // :719-720, 731
  n1 = add(n1, h.x);                                    // :732-733
  n1 = mul(vz, n1);                                     // :734
  float n2 = add(mul(k.z, t_ip), mul(k.w, t_im));       // :721-722, 737
  n2 = add(n2, h.y);                                    // :738-739
  n2 = mul(uz, n2);                                     // :740
  const float num = add(add(add(n1, n2), n3), q);       // :752-754
  const float t = div_rn(num, h.z, h.w);                // :843
  return idx < kNumMaterials ? t_inf : t;               // :847-849 (class EXTERIOR == 0)
}

// The same update in two halves (streaming path): the numerator without the heat input
// (n1 + n2) + n3, then the division and the exterior override.
__device__ __forceinline__ float cv_numerator_packed(uint32_t d, float t_jp, float t_jm, float t_im,
                                                     float t_ip, float n3, const AreaCoef& az,
                                                     const Combo* tab) {
  const int idx = (int)(d & kPackIdxMask);
  const float4* c4 = reinterpret_cast<const float4*>(tab + idx);
  const float4 k = c4[0];
  const float2 h = *reinterpret_cast<const float2*>(&c4[1]);
  const float vz = (d & kPackHalfV) ? az.half : az.full;
  const float uz = (d & kPackHalfU) ? az.half : az.full;
  float n1 = add(mul(k.x, t_jp), mul(k.y, t_jm));
  n1 = add(n1, h.x);
  n1 = mul(vz, n1);
  float n2 = add(mul(k.z, t_ip), mul(k.w, t_im));
  n2 = add(n2, h.y);
  n2 = mul(uz, n2);
  return add(add(n1, n2), n3);
}
__device__ __forceinline__ float cv_divide_packed(uint32_t d, float num, float t_inf, const Combo* tab) {
  const int idx = (int)(d & kPackIdxMask);
  const float2 dr = *reinterpret_cast<const float2*>(&tab[idx].den);
  const float t = div_rn(num, dr.x, dr.y);
  return idx < kNumMaterials ? t_inf : t;
}

// Interior CV of the dominant material with no heat input: coefficients are
// uniform over the building and live in registers (tf_simulator.py:719-754 with
// h = 0, k1 = k3 = k2 = k4 = k/dx, uz = vz).
struct FastCoef {
  float kq, vz, den, rden;
};
__device__ __forceinline__ float cv_update_fast(const FastCoef& f, float t_jp, float t_jm,
                                                float t_im, float t_ip, float n3) {
  float n1 = add(mul(f.kq, t_jp), mul(f.kq, t_jm));
  n1 = mul(f.vz, n1);        // (x + 0) + 0 == x: the convection terms vanish
  float n2 = add(mul(f.kq, t_ip), mul(f.kq, t_im));
  n2 = mul(f.vz, n2);
  const float num = add(add(n1, n2), n3);                // + q with q == 0 is exact
  return div_rn(num, f.den, f.rden);
}

// ---------------------------------------------------------------------------
// Packed fp32 pairs (sm_100 FMUL2 / FADD2 / FFMA2): one instruction issues two
// IEEE round-to-nearest fp32 operations, element by element exactly like the scalar
// instruction.  The resident kernel is issue-bound, so the FAST path (uniform
// coefficients) runs its arithmetic on pairs of neighbouring CVs.
// ---------------------------------------------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// div_rn on pairs; nb = -b (negation is exact, so fma(q, -b, a) == fma(-q, b, a))
__device__ __forceinline__ f32x2 div_rn2(f32x2 a, f32x2 nb, f32x2 y) {
  f32x2 q = mul2(a, y);
  f32x2 r = fma2(q, nb, a);
  q = fma2(r, y, q);
  r = fma2(q, nb, a);
  return fma2(r, y, q);
}
struct FastCoef2 {
  f32x2 kq, vz, nden, rden, cm, ndt, rdt;
  f32x2 one;     // (1.0f, 1.0f) from Params::one: pair_add below
};
// x + y on pairs where x (and possibly y) is a packed PRODUCT.  ptxas contracts
// add.rn.f32x2(mul.rn.f32x2(..), y) into FFMA2 (one rounding instead of two) even with
// --fmad=false; written as fma(x, ONE, y) with ONE = 1.0f read from the kernel parameters -- a value
// the compiler cannot see -- there is nothing to contract or simplify, and RN(x * 1 + y) == RN(x + y).
__device__ __forceinline__ f32x2 pair_add(f32x2 x, f32x2 one, f32x2 y) { return fma2(x, one, y); }

// Zone / grid sums are accumulated as integers: each temperature enters as
// round((T - T_ref) * 2^16), T_ref a per-building reference (the ambient
// temperature in the resident kernel, 0 elsewhere).  For fp32 temperatures >= 128 K
// the difference is exact (Sterbenz) and a multiple of 2^-16 K, so the conversion is
// EXACT; integer addition is associative, so the sums -- and therefore the zone
// means the thermostats and the reward see -- do not depend on the order in which
// lanes, warps or atomics happen to run: bitwise reproducible, and exact means.
constexpr float kFixScaleF = 65536.0f;   // 2^16
constexpr double kFixScale = 65536.0;
__device__ __forceinline__ int to_fix32(float t, float ref) {
  return __float2int_rn(__fmul_rn(__fsub_rn(t, ref), kFixScaleF));   // |T - ref| < 32768 K
}
__device__ __forceinline__ long long to_fix(float t) { return (long long)__float2ll_rn(__fmul_rn(t, kFixScaleF)); }
__device__ __forceinline__ long long to_fix(double t) { return __double2ll_rn(t * kFixScale); }
__device__ __forceinline__ double from_fix(long long s) { return (double)s / kFixScale; }
__device__ __forceinline__ void fix_add(long long* addr, long long v) {
  atomicAdd(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)v);
}
// exact warp-wide sum of int32 lane values (two hardware REDUX on 16-bit halves)
__device__ __forceinline__ long long warp_sum_i32(int s) {
  const unsigned lo = __reduce_add_sync(0xffffffffu, (unsigned)s & 0xFFFFu);
  const int hi = __reduce_add_sync(0xffffffffu, s >> 16);
  return (long long)hi * 65536 + (long long)lo;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------
// HVAC / observation / reward: executed by a GROUP of G lanes per building (G = 32:
// one warp per building; G = 8: four buildings per warp, for plans with few zones --
// these kernels are issue-bound and a warp instruction costs the same with 8 or 32
// useful lanes).  Lanes of the group stride over zones for the per-zone algebra,
// lane 0 of the group does the sequential device sums (the reference accumulates zone
// by zone in dict order, simulator_flexible_floor_plan.py:165-183).  `lane` is the lane
// index INSIDE the group, `amask` the warp's lanes that own a building (whole groups).
// `sh` is per-building scratch in shared memory: 3*Z + 64 doubles.
// ---------------------------------------------------------------------------

struct PreOut {
  double supply_air, ahu_flow, boiler_flow, return_water;
  int ahu_count, boiler_count;
  double ahu_heat_sp, ahu_cool_sp, boiler_sp;
};

__device__ inline double f32r(double v) { return (double)(float)v; }

// AirHandler.get_supply_air_temp air_handler.py:204-233
__device__ inline double mixed_air(const Params& p, double recirc, double amb) {
  return p.ahu_r * recirc + (1 - p.ahu_r) * amb;
}
__device__ inline double supply_air(double mixed, double heat_sp, double cool_sp) {
  if (mixed > cool_sp) return cool_sp;
  if (mixed < heat_sp) return heat_sp;
  return mixed;
}

// setup_step_sim + set_action + the VAV/AHU/boiler part of execute_step_sim.
// Everything here depends only on PRE-step temperatures (Q4/Q5 of SURVEY.md
// Appendix B), so it runs before the diffusion solve.
template <int G>
__device__ inline PreOut hvac_pre(const Params& p, int b, int plan, int lane, unsigned amask,
                                  const float* zone_mean /*[Z] pre-step*/, float global_mean,
                                  double* sh /*3*Z*/) {
  const int Z = p.Z;
  const int s = p.time_index;
  const bool comfort_now = p.comfort[s] != 0;
  const double w_heat = comfort_now ? p.comfort_heat : p.eco_heat;
  const double w_cool = comfort_now ? p.comfort_cool : p.eco_cool;
  const int32_t* ncv = p.zone_ncv + (size_t)plan * Z;
  const int32_t* ndiff = p.zone_ndiff + (size_t)plan * Z;

  // --- actions: BoundedActionNormalizer.setpoint_value in fp32 (NumPy weak
  // scalar promotion keeps the np.float32 action's dtype), then proto float.
  double heat_sp = p.ahu_heat_sp[b], cool_sp = p.ahu_cool_sp[b], boiler_sp = p.boiler_sp[b];
  // thermostats run BEFORE the action is applied (simulator_building.py:209)
  // but do not read any action-settable attribute, so order is immaterial here.
  for (int i = 0; i < p.n_actions; ++i) {
    const float a = p.action[(size_t)b * p.n_actions + i];
    const float ratio = fdiv(add(a, 1.0f), 2.0f);       // bounded_action_normalizer.py:93-94
    const float v = add(mul(ratio, p.action_range[i]), p.action_min[i]);  // :97-98
    const int tgt = p.action_target[i];
    if (tgt == SBX_ACT_BOILER_SETPOINT) boiler_sp = (double)v;
    else if (tgt == SBX_ACT_AHU_COOLING_SETPOINT) cool_sp = (double)v;
    else heat_sp = (double)v;
  }

  const double amb = p.ambient[(size_t)(p.n_weather == 1 ? 0 : b) * p.T_rows + s];
  const double mixed = mixed_air(p, (double)global_mean, amb);
  const double sup = supply_air(mixed, heat_sp, cool_sp);

  double* sh_flow = sh;
  double* sh_reheat = sh + Z;
  double* sh_zs = sh + 2 * Z;
  for (int zi = lane; zi < Z; zi += G) {
    double flow = 0, reheat = 0, zs = 0;
    if (ncv[zi] > 0) {
      const double zt = (double)zone_mean[zi];
      int mode = p.therm_mode[(size_t)b * Z + zi];
      // Thermostat.update thermostat.py:114-148
      const double mid = 0.5 * (w_cool - w_heat) + w_heat;
      auto default_control = [&](int m) {               // :76-112
        if (zt < w_heat) return kHeat;
        if (zt > w_cool) return kCool;
        if (zt < mid && m == kHeat) return kHeat;
        if (zt > mid && m == kCool) return kCool;
        return kOff;
      };
      if (comfort_now) mode = default_control(mode);
      else if (p.therm_seen && p.prev_comfort) mode = kPassiveCool;
      else if (mode == kPassiveCool && zt > w_heat) mode = kPassiveCool;
      else mode = default_control(mode);
      p.therm_mode[(size_t)b * Z + zi] = (uint8_t)mode;
      // Vav.update_settings vav.py:230-243
      const double damper = (mode == kHeat || mode == kCool) ? 1.0 : 0.1;
      const double valve = (mode == kHeat) ? 1.0 : 0.0;
      // Vav.output vav.py:168-217, 245-264 (input water = boiler SETPOINT)
      reheat = valve * p.vav_max_reheat;
      flow = damper * p.vav_max_flow;
      const double heat_diff = kAirHeatCapacity * flow - kWaterHeatCapacity * reheat;
      const double water_heat = boiler_sp * kWaterHeatCapacity * reheat;
      zs = (sup * heat_diff + water_heat) / flow / kAirHeatCapacity;
      const double q = flow * kAirHeatCapacity * (zs - zt);
      p.q_zone[(size_t)b * Z + zi] = q;
      p.zone_supply[(size_t)b * Z + zi] = zs;
      // apply_thermal_power_zone building.py:873-889: power * (1/n_diffusers)
      const int nd = ndiff[zi];
      const double qd = nd > 0 ? q * (1.0 / (double)nd) : 0.0;
      p.qcv_next[(size_t)b * Z + zi] = (float)qd;
      p.qcv64_next[(size_t)b * Z + zi] = qd;
      sh_zs[zi] = valve;  // stash valve; zs kept in zone_supply
    } else {
      p.qcv64_next[(size_t)b * Z + zi] = 0.0;
      p.qcv_next[(size_t)b * Z + zi] = 0.f;
      p.q_zone[(size_t)b * Z + zi] = 0.0;
      p.zone_supply[(size_t)b * Z + zi] = 0.0;
      sh_zs[zi] = 0.0;
    }
    sh_flow[zi] = flow;
    sh_reheat[zi] = reheat;
    p.pre_zone_mean[(size_t)b * Z + zi] = zone_mean[zi];
  }
  __syncwarp(amask);
  PreOut o;
  o.supply_air = sup;
  o.ahu_heat_sp = heat_sp;
  o.ahu_cool_sp = cool_sp;
  o.boiler_sp = boiler_sp;
  o.ahu_flow = 0; o.boiler_flow = 0; o.return_water = 0; o.ahu_count = 0; o.boiler_count = 0;
  if (lane == 0) {
    double af = 0, bf = 0, num = 0, den = 0;
    int ac = 0, bc = 0;
    for (int zi = 0; zi < Z; ++zi) {
      if (ncv[zi] <= 0) continue;
      if (sh_flow[zi] > 0) {                            // air_handler.py:254-268
        af += sh_flow[zi];
        if (af > p.ahu_max_flow) af = p.ahu_max_flow;
        ++ac;
      }
      if (sh_reheat[zi] > 0) { bf += sh_reheat[zi]; ++bc; }  // boiler.py:219-231
      const double valve = sh_zs[zi];                   // simulator.py:373-381
      num += valve * p.zone_supply[(size_t)b * Z + zi];
      den += valve;
    }
    o.ahu_flow = af; o.boiler_flow = bf; o.ahu_count = ac; o.boiler_count = bc;
    o.return_water = num / (den + 1e-6);
    p.ahu_heat_sp[b] = heat_sp;
    p.ahu_cool_sp[b] = cool_sp;
    p.boiler_sp[b] = boiler_sp;
    double* dg = p.diag + (size_t)b * SBX_DIAG_N;
    dg[SBX_DIAG_SUPPLY_AIR_TEMP] = sup;
    dg[SBX_DIAG_AHU_FLOW] = af;
    dg[SBX_DIAG_BOILER_FLOW] = bf;
    dg[SBX_DIAG_RETURN_WATER] = o.return_water;
  }
  // broadcast lane 0's sums
  o.ahu_flow = __shfl_sync(amask, o.ahu_flow, 0, G);
  o.boiler_flow = __shfl_sync(amask, o.boiler_flow, 0, G);
  o.return_water = __shfl_sync(amask, o.return_water, 0, G);
  o.ahu_count = __shfl_sync(amask, o.ahu_count, 0, G);
  o.boiler_count = __shfl_sync(amask, o.boiler_count, 0, G);
  __syncwarp(amask);
  return o;
}

// state carried from hvac_pre to hvac_post through global memory in the
// streaming path (the two run in different kernels there)
struct Carry {
  double ahu_flow, boiler_flow, return_water;
  int ahu_count, boiler_count;
};

// StandardScoreObservationNormalizer._normalize_one observation_normalizer.py:68-92
// on a proto-float value, result stored to a proto float again.
__device__ inline float normalize(const Params& p, int field, double native) {
  const double v = f32r(native);
  const double var = p.obs_var[field];
  if (var > 0.0) return (float)((v - p.obs_mean[field]) / p.obs_sd[field]);   // / np.sqrt(variance)
  return 0.f;
}

// Observation pack + reward.  `post_zone_mean` / `post_global_mean` are the
// fresh (post-solve) means; `is_reset` selects the restart TimeStep of
// Environment._reset (environment.py:1165-1212).
template <int G>
__device__ inline void hvac_post(const Params& p, int b, int plan, int lane, unsigned amask, bool is_reset,
                                 const float* pre_zone_mean, const float* post_zone_mean,
                                 float post_global_mean, const Carry& cy, double* sh /*3*Z*/) {
  const int Z = p.Z;
  const int s1 = is_reset ? 0 : p.time_index + 1;   // timestamp index of the observation
  const int32_t* ncv = p.zone_ncv + (size_t)plan * Z;
  const double amb1 = p.ambient[(size_t)(p.n_weather == 1 ? 0 : b) * p.T_rows + s1];
  const double heat_sp = p.ahu_heat_sp[b], cool_sp = p.ahu_cool_sp[b];
  const double boiler_sp = p.boiler_sp[b];
  float* obs = p.obs ? p.obs + (size_t)b * p.D : nullptr;

  // ---- observation (environment.py:873-985) ----
  // Boiler.supply_water_temperature_sensor is stateful (boiler.py:146-217) and is
  // read BEFORE the reward is computed (environment.py:1297-1306).
  double tank = p.boiler_tank[(size_t)b * 3 + 0];
  double tank_dt = p.boiler_tank[(size_t)b * 3 + 1];
  double last_dur = p.boiler_tank[(size_t)b * 3 + 2];
  if (lane == 0) {
    if (is_reset) {
      // after Boiler.reset(): temperature == setpoint, so _adjust_temperature is
      // the identity whatever the (possibly negative) duration is.
      tank = boiler_sp; tank_dt = 0.0; last_dur = 0.0;
    } else {
      last_dur = (double)p.dt;                          // obs_ts - action_ts
      if (p.boiler_cool_rate > 0.0 && p.boiler_heat_rate > 0.0) {
        const double begin = tank;
        if (boiler_sp > begin) tank = fmin(begin + p.boiler_heat_rate * last_dur / 60.0, boiler_sp);
        else if (boiler_sp < begin) tank = fmax(begin - p.boiler_cool_rate * last_dur / 60.0, boiler_sp);
        else tank = boiler_sp;
        tank_dt = tank - begin;
      } else {
        tank = boiler_sp;
      }
    }
    p.boiler_tank[(size_t)b * 3 + 0] = tank;
    p.boiler_tank[(size_t)b * 3 + 1] = tank_dt;
    p.boiler_tank[(size_t)b * 3 + 2] = last_dur;
    if (obs) {
      const double fan_pct = cy.ahu_flow / p.ahu_max_flow;    // air_handler.py:246-248
      obs[0] = normalize(p, 0, (double)cy.ahu_count);         // cooling_request_count
      obs[1] = normalize(p, 1, p.ahu_dp);                     // differential_pressure_setpoint
      obs[2] = normalize(p, 2, fan_pct);                      // discharge_fan_speed_percentage_command
      obs[3] = normalize(p, 3, (1.0 - p.ahu_r) * cy.ahu_flow);  // outside_air_flowrate_sensor
      obs[4] = normalize(p, 4, amb1);                         // outside_air_temperature_sensor
      obs[5] = normalize(p, 5, cool_sp);
      obs[6] = normalize(p, 6, cy.ahu_flow);                  // supply_air_flowrate_sensor
      obs[7] = normalize(p, 7, heat_sp);
      obs[8] = normalize(p, 8, fan_pct);                      // supply_fan_speed_percentage_command
      obs[9] = normalize(p, 9, (double)cy.boiler_count);      // heating_request_count
      obs[10] = normalize(p, 10, boiler_sp);                  // supply_water_setpoint
      obs[11] = normalize(p, 11, tank);                       // supply_water_temperature_sensor
    }
  }
  // per-VAV fields: damper, flow setpoint, cached zone temperature (vav.py:54-64;
  // the cached temperature is the PRE-step mean, 0 after reset -- SURVEY Q4/Q25)
  int n_hist_total = 0;
  if (obs) {
    if (p.obs_mode == SBX_OBS_RAW) {
      const int32_t* order = p.obs_zone_order + (size_t)plan * Z;
      for (int slot = lane; slot < Z; slot += G) {
        const int zi = order[slot];
        float v0 = 0.f, v1 = 0.f, v2 = 0.f;
        if (zi >= 0 && ncv[zi] > 0) {
          double damper = 0.1, ztemp = 0.0;
          if (!is_reset) {
            const int mode = p.therm_mode[(size_t)b * Z + zi];
            damper = (mode == kHeat || mode == kCool) ? 1.0 : 0.1;
            ztemp = (double)pre_zone_mean[zi];
          }
          v0 = normalize(p, 12, damper);
          v1 = normalize(p, 13, p.vav_max_flow);
          v2 = normalize(p, 14, ztemp);
        }
        obs[12 + slot * 3 + 0] = v0;
        obs[12 + slot * 3 + 1] = v1;
        obs[12 + slot * 3 + 2] = v2;
      }
      n_hist_total = 3 * Z;
    } else {
      // HistogramReducer (histogram_reducer.py:136-146, 411-433): clip to the bin
      // range, np.histogram with the last edge duplicated, divide by the count.
      // Lanes cooperate per measurement through shared counters.
      float* cnt = reinterpret_cast<float*>(sh + 3 * Z);  // 64 doubles = room for 3*SBX_MAX_HIST_BINS floats
      const int total_bins = p.n_hist_bins[0] + p.n_hist_bins[1] + p.n_hist_bins[2];
      for (int i = lane; i < total_bins; i += G) cnt[i] = 0.f;
      __syncwarp(amask);
      int base = 0;
      float nz = 0.f;
      for (int f = 0; f < 3; ++f) {
        const int nb = p.n_hist_bins[f];
        const double* bins = p.hist_bins + f * SBX_MAX_HIST_BINS;
        for (int zi = lane; zi < Z; zi += G) {
          if (ncv[zi] <= 0) continue;
          double native;
          if (f == 0) {
            int mode = p.therm_mode[(size_t)b * Z + zi];
            native = is_reset ? 0.1 : ((mode == kHeat || mode == kCool) ? 1.0 : 0.1);
          } else if (f == 1) {
            native = p.vav_max_flow;
          } else {
            native = is_reset ? 0.0 : (double)pre_zone_mean[zi];
          }
          double v = (double)normalize(p, 12 + f, native);
          v = fmin(fmax(v, bins[0]), bins[nb - 1]);
          int k = 0;  // number of edges <= v, minus 1
          for (int e = 1; e < nb; ++e) k += (bins[e] <= v) ? 1 : 0;
          atomicAdd(&cnt[base + k], 1.0f);
          if (f == 0) nz += 1.f;
        }
        base += nb;
      }
      __syncwarp(amask);
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) nz += __shfl_xor_sync(amask, nz, o);
      for (int i = lane; i < total_bins; i += G) obs[12 + i] = fdiv(cnt[i], nz);
      __syncwarp(amask);
      n_hist_total = total_bins;
    }
    if (lane == 0) {
      float* o = obs + 12 + n_hist_total;
      const double* tf = p.time_feat + (size_t)s1 * 4;
      o[0] = (float)tf[0]; o[1] = (float)tf[1]; o[2] = (float)tf[2]; o[3] = (float)tf[3];
      o[4] = (float)p.comfort[s1];
      o[5] = (float)p.comfort_soon[s1];
      double n_occ = (double)p.occ_obs[s1];
      if (p.occ_obs_per_env) {       // int(sum_z occupancy(z, t - 5 min, t)) simulator_building.py:305-315
        double acc = 0.0;
        for (int zi = 0; zi < Z; ++zi)
          if (ncv[zi] > 0) acc += p.occ_obs_zone[(size_t)s1 * p.n_occ_zones + (p.n_occ_zones == 1 ? 0 : zi)];
        n_occ = (double)(long long)acc;
      }
      o[6] = (float)((n_occ - p.occ_norm) / (p.occ_norm + 1));  // environment.py:952-956
    }
  }
  if (is_reset) {
    if (lane == 0) {
      if (p.reward) p.reward[b] = 0.f;
      if (p.step_type) p.step_type[b] = SBX_STEP_FIRST;
      if (p.discount_out) p.discount_out[b] = 1.f;
    }
    return;
  }

  // ---- reward (setpoint_energy_carbon_regret.py:142-291) ----
  const bool comfort1 = p.comfort[s1] != 0;
  const double w_heat = f32r(comfort1 ? p.comfort_heat : p.eco_heat);
  const double w_cool = f32r(comfort1 ? p.comfort_cool : p.eco_cool);
  const double dts = (double)p.dt;
  double prod = 0.0, occ_sum = 0.0;
  for (int zi = lane; zi < Z; zi += G) {
    if (ncv[zi] <= 0) continue;
    const double occ = f32r(p.occ_reward[(size_t)s1 * p.n_occ_zones + (p.n_occ_zones == 1 ? 0 : zi)]);
    const double t = (double)post_zone_mean[zi];
    const double x0low = w_heat - p.delta, x0high = w_cool + p.delta;
    double pr;                                          // base:83-123
    if (t < w_heat) pr = p.pmax / (1.0 + exp(-p.stiff * (t - x0low)));
    else if (t > w_cool) pr = p.pmax * (1.0 - 1.0 / (1.0 + exp(-p.stiff * (t - x0high))));
    else pr = p.pmax;
    prod += pr * occ * dts / 3600.0;
    occ_sum += occ;
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    prod += __shfl_xor_sync(amask, prod, o);
    occ_sum += __shfl_xor_sync(amask, occ_sum, o);
  }
  if (lane == 0) {
    // RewardInfo (simulator_flexible_floor_plan.py:238-283), each field -> proto float
    const double blower = f32r(cy.ahu_flow * p.ahu_dp / p.ahu_eff
                               + cy.ahu_flow * (1.0 - p.ahu_r) * p.ahu_dp / p.ahu_eff);
    const double mixed1 = mixed_air(p, (double)post_global_mean, amb1);
    const double sup1 = supply_air(mixed1, heat_sp, cool_sp);
    const double ac = f32r(cy.ahu_flow * kAirHeatCapacity * (sup1 - mixed1));
    const double rw = cy.return_water;
    const double sw = boiler_sp > rw ? boiler_sp : rw;  // boiler.py:244-250
    double gas = kWaterHeatCapacity * cy.boiler_flow * (sw - rw);
    gas += p.boiler_diss_factor * (sw - amb1);          // boiler.py:275-320
    if (last_dur > 0) gas += kWaterHeatCapacity * p.boiler_capacity * tank_dt / last_dur;
    gas = f32r(gas);
    const double pump = f32r(cy.boiler_flow * kWaterDensity * kGravity * p.boiler_head / p.boiler_eff);

    const double pe = p.price_e[s1], ce = p.carbon_e[s1], pg = p.price_g[s1];
    double regret = 0.0, n_cost = 0.0, n_carbon = 0.0, actual = prod;
    float value;
    if (p.reward_kind == SBX_REWARD_ENERGY_CARBON) {
      // SetpointEnergyCarbonRewardFunction (setpoint_energy_carbon_reward.py:127-190):
      // uncapped rates, absolute costs, carbon cost read back from a proto float
      const double e_rate = blower + fabs(ac) + pump;   // base:125-148
      const double cost_e = pe * fabs(e_rate) * dts;
      const double carb_e = ce * fabs(e_rate) * dts;
      const double g_pos = gas < 0.0 ? 0.0 : gas;
      const double cost_g = pg * (g_pos * dts);
      const double carb_g = p.gas_carbon * (g_pos * dts);
      const double carbon_cost = f32r((carb_e + carb_g) * p.carbon_cost_factor);   // :174-176
      const double raw = prod - p.wv * (cost_e + cost_g) - p.ww * carbon_cost;    // :178-183
      value = (float)((raw - p.reward_shift) / p.reward_scale);                   // :185-187
      n_cost = cost_e + cost_g;          // diagnostics: absolute USD / kg here
      n_carbon = carb_e + carb_g;
    } else {
      const double max_p = p.pmax * occ_sum * dts / 3600.0;
      const double min_p = p.pmin * occ_sum * dts / 3600.0;
      actual = fmax(prod, min_p);
      regret = occ_sum > 0.0 ? (actual - min_p) / (max_p - min_p) - 1.0 : 0.0;
      const double e_rate = fmin(blower + fabs(ac) + pump, p.emax);
      const double g_rate = fmin(gas, p.gmax);
      const double cost_e = pe * fabs(e_rate) * dts, cost_e_max = pe * fabs(p.emax) * dts;
      const double carb_e = ce * fabs(e_rate) * dts, carb_e_max = ce * fabs(p.emax) * dts;
      const double g_pos = g_rate < 0.0 ? 0.0 : g_rate;
      const double cost_g = pg * (g_pos * dts), cost_g_max = pg * (p.gmax * dts);
      const double carb_g = p.gas_carbon * (g_pos * dts), carb_g_max = p.gas_carbon * (p.gmax * dts);
      n_cost = (cost_e + cost_g) / (cost_e_max + cost_g_max);
      n_carbon = (carb_e + carb_g) / (carb_e_max + carb_g_max);
      const double raw = regret * p.wu - n_cost * p.wv - n_carbon * p.ww;
      value = (float)(raw / (p.wu + p.wv + p.ww));
    }
    const bool ended = p.step_count >= p.episode_steps;  // environment.py:1365-1368
    if (p.reward) p.reward[b] = value;
    if (p.step_type) p.step_type[b] = ended ? SBX_STEP_LAST : SBX_STEP_MID;
    if (p.discount_out) p.discount_out[b] = ended ? 0.f : (float)p.discount;
    double* dg = p.diag + (size_t)b * SBX_DIAG_N;
    dg[SBX_DIAG_BLOWER_W] = blower;
    dg[SBX_DIAG_AC_W] = ac;
    dg[SBX_DIAG_GAS_W] = gas;
    dg[SBX_DIAG_PUMP_W] = pump;
    dg[SBX_DIAG_REGRET] = regret;
    dg[SBX_DIAG_NORM_COST] = n_cost;
    dg[SBX_DIAG_NORM_CARBON] = n_carbon;
    dg[SBX_DIAG_TOTAL_OCC] = occ_sum;
    dg[SBX_DIAG_PRODUCTIVITY] = actual;
    dg[SBX_DIAG_COOLING_REQUESTS] = (double)cy.ahu_count;
    dg[SBX_DIAG_HEATING_REQUESTS] = (double)cy.boiler_count;
    dg[SBX_DIAG_TANK_TEMP] = tank;
  }
}

}  // namespace sbx
