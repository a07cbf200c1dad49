// libsbx.so: C ABI (include/sbx.h) over the sm_100a kernels in sbx_kernels.cuh.
//
// Host-side design
//   * The handle owns every device allocation (state, static plan data,
//     exogenous tables); callers pass only I/O buffers.
//   * All envs of a handle share one clock, so step_count / time_index /
//     episode_ended and the thermostats' "previous timestamp" (which the
//     reference never resets, vav.py:98) are host integers.
//   * A step is HVAC prologue (k_pre) -> diffusion solve -> epilogue (k_post).
//     Resident path: the solve is ONE launch, no host synchronisation anywhere
//     in the step.  Streaming path: one launch per Jacobi sweep with a
//     pinned-memory convergence poll, then a zone-reduction launch.
//   * sbx_step_host stages through pinned buffers so that host<->device copies
//     are asynchronous DMA on the handle's stream.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <new>
#include <string>
#include <vector>

#include "../../include/sbx.h"
#include "sbx_kernels.cuh"
#include "sbx_resident2.cuh"
#include "sbx_resident3.cuh"
#include "sbx_pairwise.cuh"
#include "sbx_sweep_tma.cuh"

using namespace sbx;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

}  // namespace

constexpr int SBX_MAX_CHUNKS = 8;

struct sbx_env {
  sbx_config cfg;
  int device = 0;
  int D = 0;
  int path = SBX_PATH_STREAMING;
  int V = 1;
  int n_sms = 0;
  int resident_ctas_per_sm = 0;
  size_t resident_smem = 0;
  CUtensorMap tmap_t;            // [B, H, W] temperature field, box [1, H, P] (k_resident_step)
  size_t gs_smem = 0;
  int gs_global = 0;             // Gauss-Seidel on a grid too large for shared memory: wavefront in global memory
  // k_resident_step2 (persistent, record-driven) when the grid allows it and no plan
  // overflows its list capacities; otherwise k_resident_step
  int v2_capable = 0, use_v2 = 0, v1_allocated = 0;
  // k_resident_step3 (one thread per 4x4 tile, field in registers): the default where the grid
  // allows it; handles with stochastic convection run k_resident_step for those steps
  int v3_capable = 0, use_v3 = 0, v1_ready = 0, v23_ready = 0;
  int resident3_ctas_per_sm = 0;
  int last_resident_kernel = 0;  // sbx_info.resident_kernel
  int resident2_ctas_per_sm = 0;
  Params P;
  // host-level episode state
  int step_count = 0, time_index = 0, episode_ended = 0, reset_called = 0;
  int therm_seen = 0, prev_comfort = 0;
  int plans_dirty = 1;           // packed descriptors / vector lists need (re)building
  int32_t* conv_perm = nullptr;  // lazily allocated [B,H*W]
  int conv_pending = 0;          // a permutation was uploaded for the next step
  uint8_t* h_comfort = nullptr;  // host copy of the comfort table (for prev_comfort)
  // device allocations
  DevBuf all[128];
  int n_all = 0;
  int64_t device_bytes = 0;
  int64_t launches = 0;
  int64_t env_steps = 0;
  CarryStore* carry = nullptr;
  float* gather = nullptr;       // [B,H,W] staging for download of T (streaming path)
  double* fd_ambient = nullptr;
  double* fd_convection = nullptr;
  double* d_obs_mean = nullptr;
  double* d_obs_var = nullptr;
  double* d_obs_sd = nullptr;
  double* d_hist_bins = nullptr;
  // I/O staging for the *_host calls
  float* d_action = nullptr;
  float* d_obs = nullptr;
  float* d_reward = nullptr;
  int32_t* d_step_type = nullptr;
  float* d_discount = nullptr;
  float* h_action = nullptr;     // pinned
  float* h_obs = nullptr;
  float* h_reward = nullptr;
  int32_t* h_step_type = nullptr;
  float* h_discount = nullptr;
  int32_t* h_n_active = nullptr; // pinned
  cudaStream_t stream = nullptr; // handle-owned stream for *_host calls and uploads
  // Resident path, large batches: the step is cut into `n_chunks` contiguous shares
  // of the batch.  The small HVAC kernels run on a high-priority stream, the solves
  // on a low-priority one, chained by events, so that the epilogue of share c fills
  // SM slots while share c+1 is still solving (see do_step).
  int n_chunks = 1;
  // streaming path: the sweep loop as a CUDA graph with a WHILE node ([0] step, [1] sbx_fd_step)
  cudaGraph_t loop_graph[2] = {nullptr, nullptr};
  cudaGraphExec_t loop_exec[2] = {nullptr, nullptr};
  int loop_graph_failed = 0;
  int64_t graph_launches = 0;
  unsigned long long loop_launches_seen = 0;   // sweep_launches at sbx_timing_begin
  cudaStream_t s_hi = nullptr, s_lo = nullptr, s_copy = nullptr;
  cudaEvent_t ev_share[SBX_MAX_CHUNKS] = {};
  unsigned share_seq = 0;
  int host_shares = 0;           // SBX_OPT_HOST_SHARES (0 = the library's choice)
  int sweep_tma = 0;             // streaming path: k_sweep_tma (TMA-staged tiles) instead of k_sweep
  SweepTensorMaps sweep_tm;
  int pw_dirty = 1;              // SBX_OPT_NUMPY_MEANS: the per-plan summation trees need (re)building
  int pw_by_solve = 0;           // this step's solve kernel wrote pw_mean itself
  size_t pw_alloc_leaf = 0, pw_alloc_node = 0, pw_alloc_val = 0;   // capacities allocated so far
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_pre[SBX_MAX_CHUNKS] = {}, ev_solve[SBX_MAX_CHUNKS] = {};
  // sbx_timing_begin / sbx_timing_end: CUDA events around every step and every solve launch
  int timing_on = 0;
  std::vector<cudaEvent_t> t_step, t_solve;   // pairs (begin, end)
  size_t t_step_used = 0, t_solve_used = 0;
  std::string err;
};

namespace {

int fail(sbx_handle h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf; else g_create_error = buf;
  return code;
}

#define CUDA_TRY(h, expr)                                                              \
  do {                                                                                 \
    cudaError_t e__ = (expr);                                                          \
    if (e__ != cudaSuccess)                                                            \
      return fail(h, SBX_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                  __FILE__, __LINE__);                                                 \
  } while (0)

template <typename T>
int dev_alloc(sbx_handle h, T** out, size_t count, bool zero = true) {
  void* p = nullptr;
  const size_t bytes = (count ? count : 1) * sizeof(T);
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess)
    return fail(h, SBX_E_NOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
  if (zero) {
    e = cudaMemset(p, 0, bytes);
    if (e != cudaSuccess) return fail(h, SBX_E_CUDA, "cudaMemset failed: %s", cudaGetErrorString(e));
  }
  if (h->n_all >= 128) return fail(h, SBX_E_INVALID, "too many allocations");
  h->all[h->n_all].p = p;
  h->all[h->n_all].bytes = bytes;
  ++h->n_all;
  h->device_bytes += (int64_t)bytes;
  *out = reinterpret_cast<T*>(p);
  return SBX_OK;
}

int obs_dim(const sbx_config& c) {
  int d = 9 + 3 + 7;
  if (c.obs_mode == SBX_OBS_RAW) d += 3 * c.n_zones;
  else d += c.n_hist_bins[0] + c.n_hist_bins[1] + c.n_hist_bins[2];
  return d;
}

int validate(const sbx_config& c) {
  if (c.abi_version != SBX_ABI_VERSION) return fail(nullptr, SBX_E_INVALID, "abi_version %d != %d", c.abi_version, SBX_ABI_VERSION);
  if (c.n_envs < 1 || c.height < 1 || c.width < 1) return fail(nullptr, SBX_E_INVALID, "n_envs/height/width must be >= 1");
  if (c.n_zones < 1 || c.n_zones > 254) return fail(nullptr, SBX_E_INVALID, "n_zones must be in [1, 254]");
  if (c.n_plans != 1 && c.n_plans != c.n_envs) return fail(nullptr, SBX_E_INVALID, "n_plans must be 1 or n_envs");
  if (c.n_weather != 1 && c.n_weather != c.n_envs) return fail(nullptr, SBX_E_INVALID, "n_weather must be 1 or n_envs");
  if (c.n_reset != 0 && c.n_reset != 1 && c.n_reset != c.n_envs) return fail(nullptr, SBX_E_INVALID, "n_reset must be 0, 1 or n_envs");
  if (c.n_occ_zones != 1 && c.n_occ_zones != c.n_zones) return fail(nullptr, SBX_E_INVALID, "n_occ_zones must be 1 or n_zones");
  if (c.episode_steps < 0 || c.n_table_steps < c.episode_steps + 2) return fail(nullptr, SBX_E_INVALID, "n_table_steps must be >= episode_steps + 2");
  if (c.n_actions < 0 || c.n_actions > SBX_MAX_ACTIONS) return fail(nullptr, SBX_E_INVALID, "n_actions must be in [0, %d]", SBX_MAX_ACTIONS);
  for (int i = 0; i < c.n_actions; ++i)
    if (c.action_target[i] < 0 || c.action_target[i] > 2) return fail(nullptr, SBX_E_INVALID, "bad action_target[%d]", i);
  if (c.obs_mode != SBX_OBS_RAW && c.obs_mode != SBX_OBS_HISTOGRAM) return fail(nullptr, SBX_E_INVALID, "bad obs_mode");
  if (c.obs_mode == SBX_OBS_HISTOGRAM)
    for (int f = 0; f < 3; ++f)
      if (c.n_hist_bins[f] < 1 || c.n_hist_bins[f] > SBX_MAX_HIST_BINS) return fail(nullptr, SBX_E_INVALID, "n_hist_bins[%d] out of range", f);
  if (c.iteration_limit < 1) return fail(nullptr, SBX_E_INVALID, "iteration_limit must be >= 1");
  if (!(c.time_step_sec > 0.0)) return fail(nullptr, SBX_E_INVALID, "time_step_sec must be > 0");
  if (c.solver != SBX_SOLVER_TF_JACOBI && c.solver != SBX_SOLVER_GAUSS_SEIDEL) return fail(nullptr, SBX_E_INVALID, "bad solver");
  if (c.discount_factor <= 0 || c.discount_factor > 1) return fail(nullptr, SBX_E_INVALID, "Discount factor must be in (0,1]");  // environment.py:446
  if (c.ahu_init_cooling_setpoint <= c.ahu_init_heating_setpoint) return fail(nullptr, SBX_E_INVALID, "cooling_air_temp_setpoint must greater than heating_air_temp_setpoint");  // air_handler.py:62-66
  if (c.reward_kind != SBX_REWARD_REGRET && c.reward_kind != SBX_REWARD_ENERGY_CARBON) return fail(nullptr, SBX_E_INVALID, "bad reward_kind");
  if (c.reward_kind == SBX_REWARD_REGRET && c.max_productivity_personhour_usd <= c.min_productivity_personhour_usd) return fail(nullptr, SBX_E_INVALID, "max productivity must exceed min productivity");
  if (c.reward_kind == SBX_REWARD_ENERGY_CARBON && c.reward_normalizer_scale == 0.0) return fail(nullptr, SBX_E_INVALID, "reward_normalizer_scale must not be 0");
  return SBX_OK;
}

void fill_params(sbx_handle h) {
  const sbx_config& c = h->cfg;
  Params& p = h->P;
  p.B = c.n_envs; p.H = c.height; p.W = c.width; p.Z = c.n_zones;
  p.n_plans = c.n_plans; p.n_weather = c.n_weather; p.n_reset = c.n_reset;
  p.n_occ_zones = c.n_occ_zones; p.T_rows = c.n_table_steps;
  p.obs_mode = c.obs_mode; p.D = h->D; p.n_actions = c.n_actions;
  for (int i = 0; i < SBX_MAX_ACTIONS; ++i) {
    p.action_target[i] = c.action_target[i];
    // `agent_ratio * output_range + min` with np.float32 agent_ratio: the Python
    // floats are cast to fp32 first (NEP 50 weak scalars).
    p.action_min[i] = (float)c.action_min[i];
    p.action_range[i] = (float)(c.action_max[i] - c.action_min[i]);
  }
  for (int f = 0; f < 3; ++f) p.n_hist_bins[f] = c.obs_mode == SBX_OBS_HISTOGRAM ? c.n_hist_bins[f] : 0;
  p.dt = (float)c.time_step_sec; p.z = (float)c.floor_height_m;
  p.threshold = (float)c.convergence_threshold;   // NumPy weak-scalar compare: fp32
  p.dt_double = c.time_step_sec; p.z_double = c.floor_height_m;
  p.threshold64 = c.convergence_threshold;
  p.iteration_limit = c.iteration_limit;
  p.comfort_heat = c.comfort_heat; p.comfort_cool = c.comfort_cool;
  p.eco_heat = c.eco_heat; p.eco_cool = c.eco_cool;
  p.ahu_r = c.ahu_recirculation; p.ahu_init_heat = c.ahu_init_heating_setpoint;
  p.ahu_init_cool = c.ahu_init_cooling_setpoint; p.ahu_dp = c.ahu_fan_differential_pressure;
  p.ahu_eff = c.ahu_fan_efficiency; p.ahu_max_flow = c.ahu_max_air_flow_rate;
  p.boiler_init_sp = c.boiler_init_setpoint; p.boiler_head = c.boiler_pump_head;
  p.boiler_eff = c.boiler_pump_efficiency; p.boiler_heat_rate = c.boiler_heating_rate;
  p.boiler_cool_rate = c.boiler_cooling_rate;
  {  // compute_thermal_dissipation_rate boiler.py:309-320
    const double r1 = c.boiler_tank_radius, r2 = r1 + c.boiler_insulation_thickness;
    const double conduction = log(r2 / r1) / c.boiler_insulation_conductivity;
    const double convection = 1.0 / c.boiler_convection_coefficient / r2;
    p.boiler_diss_factor = c.boiler_tank_length * 2.0 * M_PI / (conduction + convection);
  }
  p.boiler_capacity = c.boiler_water_capacity;
  p.vav_max_flow = c.vav_max_air_flow_rate; p.vav_max_reheat = c.vav_reheat_max_water_flow_rate;
  p.pmax = c.max_productivity_personhour_usd; p.pmin = c.min_productivity_personhour_usd;
  p.emax = c.max_electricity_rate; p.gmax = c.max_natural_gas_rate;
  p.delta = c.productivity_midpoint_delta; p.stiff = c.productivity_decay_stiffness;
  p.wu = c.productivity_weight; p.wv = c.energy_cost_weight; p.ww = c.carbon_emission_weight;
  p.gas_carbon = c.gas_carbon_rate;
  p.carbon_cost_factor = c.carbon_cost_factor; p.reward_shift = c.reward_normalizer_shift;
  p.reward_scale = c.reward_normalizer_scale; p.reward_kind = c.reward_kind;
  p.discount = c.discount_factor; p.occ_norm = c.occupancy_normalization_constant;
  p.episode_steps = c.episode_steps;
  p.fd_only = 0;
  p.one = 1.0f;
}

// launch helpers --------------------------------------------------------------

// lanes per building in k_pre / k_post: 4 (eight buildings per warp) for plans with few zones
int hvac_group(const sbx_handle h) {
  if (const char* e = getenv("SBX_HVAC_GROUP")) {     // tuning override: 4, 8 or 32
    const int g = atoi(e);
    if (g == 4 || g == 8 || g == 32) return g;
  }
  // measured at 12 zones (32768 buildings): 84 us per step with 4 lanes, 93 us with 8
  return h->cfg.n_zones <= 16 ? 4 : 32;
}
int pre_post_smem(const sbx_handle h, int groups) {
  return groups * (3 * h->cfg.n_zones + 64 + h->cfg.n_zones) * (int)sizeof(double);
}

int launch_check(sbx_handle h, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(h, SBX_E_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
  ++h->launches;
  return SBX_OK;
}

static bool step_has_convection(const Params& p) {
  return !p.fd_only && (p.conv_perm != nullptr || p.conv_p > 0.0);
}

int launch_zone_reduce(sbx_handle h, cudaStream_t st) {
  const Params& p = h->P;
  if (p.pw_on) return SBX_OK;          // k_post takes the means of k_pw_combine, not these sums
  CUDA_TRY(h, cudaMemsetAsync(p.zone_sum, 0, sizeof(long long) * (size_t)p.B * (p.Z + 1), st));
  const StreamTiling tl = stream_tiling(p.H, p.W, h->V);
  const unsigned grid = (unsigned)((size_t)tl.tiles * p.B);
  if (h->V == 4) k_zone_reduce<4><<<grid, kStreamThreads, 0, st>>>(p);
  else k_zone_reduce<1><<<grid, kStreamThreads, 0, st>>>(p);
  return launch_check(h, "k_zone_reduce");
}

// timing events (sbx_timing_begin .. sbx_timing_end)
int timing_record(sbx_handle h, std::vector<cudaEvent_t>& pool, size_t& used, cudaStream_t st) {
  if (!h->timing_on) return SBX_OK;
  if (used == pool.size()) {
    cudaEvent_t e;
    CUDA_TRY(h, cudaEventCreate(&e));
    pool.push_back(e);
  }
  CUDA_TRY(h, cudaEventRecord(pool[used++], st));
  return SBX_OK;
}

// CTAs of a k_resident_step2 launch over buildings [b_begin, b_end): one per resident slot
unsigned v2_grid(const sbx_handle h) {
  const unsigned nb = (unsigned)(h->P.b_end - h->P.b_begin);
#if SBX_R2_PERSISTENT
  const unsigned slots = (unsigned)(h->n_sms * (h->resident2_ctas_per_sm > 0 ? h->resident2_ctas_per_sm : 1));
  return nb < slots ? nb : slots;
#else
  return nb;                              // one building per CTA
#endif
}
// building whose list sizes ride in building b's header: the one the same SM slot takes next
int v2_next_distance(const sbx_handle h) {
#if SBX_R2_PERSISTENT
  return (int)v2_grid(h);
#else
  return h->P.prefetch_dist;
#endif
}

int launch_pre(sbx_handle h, cudaStream_t st) {
  h->P.v2_stride = v2_next_distance(h);  // build_header stores the next building's list sizes
  const Params& p = h->P;
  const int G = hvac_group(h), bpc = 128 / G;          // buildings per CTA
  const unsigned grid = (unsigned)((p.b_end - p.b_begin + bpc - 1) / bpc);
  if (G == 4) k_pre<4><<<grid, 128, pre_post_smem(h, bpc), st>>>(p, h->carry);
  else if (G == 8) k_pre<8><<<grid, 128, pre_post_smem(h, bpc), st>>>(p, h->carry);
  else k_pre<32><<<grid, 128, pre_post_smem(h, bpc), st>>>(p, h->carry);
  return launch_check(h, "k_pre");
}

// SBX_OPT_NUMPY_MEANS: NumPy's pairwise summation tree (see sbx_pairwise.cuh) of every zone
// and of the grid, per plan.  Built on the host from the plan's zone bytes -- a counting sort
// of the CVs by zone in raster order and a recursion whose shape depends on the counts alone.
struct PwTree {
  std::vector<uint2> leaf;                 // {start | flag, length}
  std::vector<uint2> node;                 // {left, right}: leaf index, or ~(node id) before renumbering
  std::vector<int> height;                 // of every node
};
// returns {reference, height}: reference = leaf index >= 0, or ~(node id) < 0
static std::pair<int, int> pw_build(PwTree& t, uint32_t start, uint32_t n, uint32_t flag) {
  if (n <= 128) {
    t.leaf.push_back(make_uint2(start | flag, n));
    return {(int)t.leaf.size() - 1, 0};
  }
  uint32_t n2 = n / 2;
  n2 -= n2 % 8;
  const std::pair<int, int> a = pw_build(t, start, n2, flag), b = pw_build(t, start + n2, n - n2, flag);
  t.node.push_back(make_uint2((uint32_t)a.first, (uint32_t)b.first));
  const int hgt = 1 + (a.second > b.second ? a.second : b.second);
  t.height.push_back(hgt);
  return {~((int)t.node.size() - 1), hgt};
}

int build_pairwise_tables(sbx_handle h, cudaStream_t st) {
  Params& p = h->P;
  const int Z = p.Z, P = p.n_plans;
  const size_t n_cv = (size_t)p.H * p.W;
  const bool idx16 = n_cv <= 65536;
  std::vector<uint16_t> desc((size_t)P * n_cv);
  CUDA_TRY(h, cudaMemcpyAsync(desc.data(), p.desc, desc.size() * sizeof(uint16_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(h, cudaStreamSynchronize(st));
  std::vector<unsigned char> zlist((size_t)P * n_cv * (idx16 ? 2 : 4), 0);
  std::vector<PwTree> trees((size_t)P);
  std::vector<int32_t> meta((size_t)P * kPwMetaInts, 0);
  std::vector<int2> root((size_t)P * (Z + 1));
  std::vector<std::vector<uint2>> ordered((size_t)P);
  size_t capL = 1, capI = 1, cap = 1, zone_cvs = 0;
  std::vector<uint32_t> cnt((size_t)Z + 1), pos((size_t)Z + 1);
  for (int pl = 0; pl < P; ++pl) {
    const uint16_t* d = desc.data() + (size_t)pl * n_cv;
    std::fill(cnt.begin(), cnt.end(), 0u);
    for (size_t i = 0; i < n_cv; ++i) {
      const int z = (d[i] >> SBX_DESC_ZONE_SHIFT) & 0xFF;
      if (z != SBX_ZONE_NONE && z < Z) ++cnt[(size_t)z];
    }
    uint32_t o = 0;
    for (int z = 0; z < Z; ++z) { pos[(size_t)z] = o; o += cnt[(size_t)z]; }
    zone_cvs = o > zone_cvs ? o : zone_cvs;
    unsigned char* zl = zlist.data() + (size_t)pl * n_cv * (idx16 ? 2 : 4);
    for (size_t i = 0; i < n_cv; ++i) {
      const int z = (d[i] >> SBX_DESC_ZONE_SHIFT) & 0xFF;
      if (z == SBX_ZONE_NONE || z >= Z) continue;
      const uint32_t k = pos[(size_t)z]++;
      if (idx16) reinterpret_cast<uint16_t*>(zl)[k] = (uint16_t)i;
      else reinterpret_cast<uint32_t*>(zl)[k] = (uint32_t)i;
    }
    PwTree& t = trees[(size_t)pl];
    std::vector<int> root_ref((size_t)Z + 1);
    o = 0;
    for (int z = 0; z < Z; ++z) {
      root_ref[(size_t)z] = cnt[(size_t)z] ? pw_build(t, o, cnt[(size_t)z], 0u).first : INT32_MIN;
      o += cnt[(size_t)z];
    }
    root_ref[(size_t)Z] = pw_build(t, 0u, (uint32_t)n_cv, kPwGridFlag).first;
    // inner nodes grouped by height (a node's children are lower): stable counting sort
    const int L = (int)t.leaf.size(), I = (int)t.node.size();
    int n_levels = 0;
    for (int v : t.height) n_levels = v > n_levels ? v : n_levels;
    if (n_levels + 4 > kPwMetaInts) return fail(h, SBX_E_INVALID, "summation tree of %d levels", n_levels);
    std::vector<int> lvl_start((size_t)n_levels + 1, 0), newpos((size_t)I);
    for (int v : t.height) ++lvl_start[(size_t)v];            // height v -> level v - 1, counted at slot v
    for (int l = 1; l <= n_levels; ++l) lvl_start[(size_t)l] += lvl_start[(size_t)l - 1];
    // lvl_start[l] = nodes of height <= l; level l - 1 spans [lvl_start[l - 1], lvl_start[l])
    std::vector<int> fill_at(lvl_start.begin(), lvl_start.end());
    for (int i = 0; i < I; ++i) newpos[(size_t)i] = fill_at[(size_t)t.height[(size_t)i] - 1]++;
    auto value_index = [&](int ref) { return ref >= 0 ? ref : L + newpos[(size_t)~ref]; };
    std::vector<uint2>& ord = ordered[(size_t)pl];
    ord.resize((size_t)I);
    for (int i = 0; i < I; ++i)
      ord[(size_t)newpos[(size_t)i]] = make_uint2((uint32_t)value_index((int)t.node[(size_t)i].x),
                                                  (uint32_t)value_index((int)t.node[(size_t)i].y));
    int32_t* m = meta.data() + (size_t)pl * kPwMetaInts;
    m[0] = L; m[1] = I; m[2] = n_levels;
    for (int l = 0; l <= n_levels; ++l) m[3 + l] = lvl_start[(size_t)l];
    for (int z = 0; z <= Z; ++z) {
      const int ref = root_ref[(size_t)z];
      root[(size_t)pl * (Z + 1) + z] = ref == INT32_MIN ? make_int2(-1, 0)
                                                        : make_int2(value_index(ref), z < Z ? (int)cnt[(size_t)z] : (int)n_cv);
    }
    capL = (size_t)L > capL ? (size_t)L : capL;
    capI = (size_t)I > capI ? (size_t)I : capI;
    cap = (size_t)(L + I) > cap ? (size_t)(L + I) : cap;
  }
  // device tables (allocated once; again only when a new plan set needs more room)
  if (!p.pw_zlist) {
    unsigned char* z = nullptr; int32_t* mt = nullptr; int2* rt = nullptr; float* mean = nullptr;
    if (int rc = dev_alloc<unsigned char>(h, &z, zlist.size() + 16)) return rc;   // + a bulk copy's rounding
    if (int rc = dev_alloc<int32_t>(h, &mt, meta.size())) return rc;
    if (int rc = dev_alloc<int2>(h, &rt, root.size())) return rc;
    if (int rc = dev_alloc<float>(h, &mean, (size_t)p.B * (Z + 1))) return rc;
    p.pw_zlist = z; p.pw_meta = mt; p.pw_root = rt; p.pw_mean = mean;
  }
  if (capL > h->pw_alloc_leaf) {
    uint2* lf = nullptr;
    if (int rc = dev_alloc<uint2>(h, &lf, (size_t)P * capL)) return rc;
    p.pw_leaf = lf; h->pw_alloc_leaf = capL;
  }
  if (capI > h->pw_alloc_node) {
    uint2* nd = nullptr;
    if (int rc = dev_alloc<uint2>(h, &nd, (size_t)P * capI)) return rc;
    p.pw_node = nd; h->pw_alloc_node = capI;
  }
  if (cap > h->pw_alloc_val) {
    float* v = nullptr;
    if (int rc = dev_alloc<float>(h, &v, (size_t)p.B * cap)) return rc;
    p.pw_val = v; h->pw_alloc_val = cap;
  }
  p.pw_capL = (int)h->pw_alloc_leaf; p.pw_capI = (int)h->pw_alloc_node; p.pw_cap = (int)h->pw_alloc_val;
  p.pw_zbytes = (uint32_t)((zone_cvs * (idx16 ? 2 : 4) + 15) & ~(size_t)15);
  // k_resident_step<4, true> keeps leaf sums, inner nodes, the node table and the level offsets in
  // its idle temperature plane and indexes CVs with 16 bits
  // (TMA: 16-byte source alignment per plan), and takes the CV list into its zone-sum list region
  p.pw_fused = (!getenv("SBX_PW_SEPARATE") &&      // developer switch: k_pw_leaves / k_pw_combine on the resident path too
                h->path == SBX_PATH_RESIDENT && h->V == 4 && idx16 && n_cv % 8 == 0 &&
                p.pw_zbytes > 0 && (size_t)p.pw_zbytes <= (size_t)p.geom.rl_cap * 4 &&
                ((cap + 1) & ~(size_t)1) + 2 * capI + 2 * capL <= (size_t)p.geom.plane_cv) ? 1 : 0;
  std::vector<uint2> leaf_all((size_t)P * p.pw_capL, make_uint2(0u, 0u)), node_all((size_t)P * p.pw_capI, make_uint2(0u, 0u));
  for (int pl = 0; pl < P; ++pl) {
    std::copy(trees[(size_t)pl].leaf.begin(), trees[(size_t)pl].leaf.end(), leaf_all.begin() + (size_t)pl * p.pw_capL);
    std::copy(ordered[(size_t)pl].begin(), ordered[(size_t)pl].end(), node_all.begin() + (size_t)pl * p.pw_capI);
  }
  CUDA_TRY(h, cudaMemcpyAsync(const_cast<void*>(p.pw_zlist), zlist.data(), zlist.size(), cudaMemcpyHostToDevice, st));
  CUDA_TRY(h, cudaMemcpyAsync(const_cast<int32_t*>(p.pw_meta), meta.data(), meta.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  CUDA_TRY(h, cudaMemcpyAsync(const_cast<int2*>(p.pw_root), root.data(), root.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
  CUDA_TRY(h, cudaMemcpyAsync(const_cast<uint2*>(p.pw_leaf), leaf_all.data(), leaf_all.size() * sizeof(uint2), cudaMemcpyHostToDevice, st));
  CUDA_TRY(h, cudaMemcpyAsync(const_cast<uint2*>(p.pw_node), node_all.data(), node_all.size() * sizeof(uint2), cudaMemcpyHostToDevice, st));
  CUDA_TRY(h, cudaStreamSynchronize(st));      // the host vectors go out of scope
  h->pw_dirty = 0;
  return SBX_OK;
}

// zone / grid means of buildings [b_begin, b_end) in NumPy's order, into pw_mean
int launch_pairwise_means(sbx_handle h, cudaStream_t st) {
  if (h->pw_dirty) if (int rc = build_pairwise_tables(h, st)) return rc;
  const Params& p = h->P;
  const long long lanes = (long long)(p.b_end - p.b_begin) * p.pw_capL * 8;
  const unsigned grid = (unsigned)((lanes + kPwThreads - 1) / kPwThreads);
  if ((size_t)p.H * p.W <= 65536) k_pw_leaves<uint16_t><<<grid, kPwThreads, 0, st>>>(p);
  else k_pw_leaves<uint32_t><<<grid, kPwThreads, 0, st>>>(p);
  if (int rc = launch_check(h, "k_pw_leaves")) return rc;
  const unsigned nb = (unsigned)(p.b_end - p.b_begin);
  if (p.pw_capI <= 2048) k_pw_combine<32><<<(nb + kPwThreads / 32 - 1) / (kPwThreads / 32), kPwThreads, 0, st>>>(p);
  else k_pw_combine<kPwThreads><<<nb, kPwThreads, 0, st>>>(p);
  return launch_check(h, "k_pw_combine");
}

int launch_post(sbx_handle h, cudaStream_t st, int is_reset) {
  if (h->P.pw_on && !(h->pw_by_solve && !is_reset)) if (int rc = launch_pairwise_means(h, st)) return rc;
  const Params& p = h->P;
  const int G = hvac_group(h), bpc = 128 / G;
  const unsigned grid = (unsigned)((p.b_end - p.b_begin + bpc - 1) / bpc);
  if (G == 4) k_post<4><<<grid, 128, pre_post_smem(h, bpc), st>>>(p, h->carry, is_reset);
  else if (G == 8) k_post<8><<<grid, 128, pre_post_smem(h, bpc), st>>>(p, h->carry, is_reset);
  else k_post<32><<<grid, 128, pre_post_smem(h, bpc), st>>>(p, h->carry, is_reset);
  return launch_check(h, "k_post");
}

template <int V, int MODE>
void* sweep_kernel_for(int rows_per_warp) {
  switch (rows_per_warp) {
    case 8: return (void*)k_sweep<V, 8, MODE>;
    case 16: return (void*)k_sweep<V, 16, MODE>;
    case 32: return (void*)k_sweep<V, 32, MODE>;
    case 48: return (void*)k_sweep<V, 48, MODE>;
    default: return (void*)k_sweep<V, 64, MODE>;
  }
}
void* sweep_tma_kernel(int mode) { return mode == 0 ? (void*)k_sweep_tma<0> : (void*)k_sweep_tma<1>; }
// mode 0: a sweep k >= 2; 1: the first sweep
void* sweep_kernel(int V, int rows_per_warp, int mode) {
  if (V == 4) return mode == 0 ? sweep_kernel_for<4, 0>(rows_per_warp) : sweep_kernel_for<4, 1>(rows_per_warp);
  return mode == 0 ? sweep_kernel_for<1, 0>(rows_per_warp) : sweep_kernel_for<1, 1>(rows_per_warp);
}

// The sweep loop of the streaming path as a graph: k_loop_begin -> first sweep -> k_check_loop
// (sets the handle) -> WHILE(handle) { k_sweep(k on the device) -> k_check_loop(sets the handle) }.
// Built once per handle and mode (0: a step, 1: solve only); the kernels read everything that changes from step to step from device memory
// (solve header, active flags, buffer rotation), so the baked-in Params stay valid.
int build_sweep_graph(sbx_handle h, int which) {
  Params p = h->P;                       // snapshot: fd_only is part of it
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaGraphCreate(&g, 0);
  if (e != cudaSuccess) return fail(h, SBX_E_CUDA, "cudaGraphCreate failed: %s", cudaGetErrorString(e));
  const StreamTiling tl = stream_tiling(p.H, p.W, h->V);
  cudaGraphConditionalHandle handle;
  e = cudaGraphConditionalHandleCreate(&handle, g, 1, cudaGraphCondAssignDefault);
  if (e != cudaSuccess) { cudaGraphDestroy(g); return fail(h, SBX_E_CUDA, "cudaGraphConditionalHandleCreate failed: %s", cudaGetErrorString(e)); }
  auto add_kernel = [&](cudaGraph_t where, cudaGraphNode_t* dep, void* func, dim3 grid, dim3 block, void** args,
                        cudaGraphNode_t* out, unsigned smem = 0) -> cudaError_t {
    cudaKernelNodeParams kp;
    memset(&kp, 0, sizeof(kp));
    kp.func = func; kp.gridDim = grid; kp.blockDim = block; kp.kernelParams = args; kp.sharedMemBytes = smem;
    return cudaGraphAddKernelNode(out, where, dep, dep ? 1 : 0, &kp);
  };
  const bool tma = h->sweep_tma != 0;
  const dim3 sweep_grid((unsigned)((size_t)tl.tiles * p.B));
  int k_first = 1, k_dev = 0;
  cudaGraphNode_t n_begin = nullptr, n_first = nullptr, n_check1 = nullptr, n_while = nullptr, n_sweep = nullptr, n_check = nullptr;
  void* a_begin[] = {&p};
  void* a_first[] = {&p, &k_first, &h->sweep_tm};      // the tensor maps: k_sweep_tma only
  void* a_check[] = {&p, &handle};
  void* a_sweep[] = {&p, &k_dev, &h->sweep_tm};
  e = add_kernel(g, nullptr, (void*)k_loop_begin, dim3((unsigned)((p.B + 255) / 256)), dim3(256), a_begin, &n_begin);
  if (e == cudaSuccess)
    e = add_kernel(g, &n_begin, tma ? sweep_tma_kernel(1) : sweep_kernel(h->V, tl.rows_per_warp, 1), sweep_grid,
                   dim3(kStreamThreads), a_first, &n_first, tma ? (unsigned)sweep_tma_smem(true) : 0u);
  if (e == cudaSuccess) e = add_kernel(g, &n_first, (void*)k_check_loop, dim3(1), dim3(1024), a_check, &n_check1);
  if (e != cudaSuccess) { cudaGraphDestroy(g); return fail(h, SBX_E_CUDA, "cudaGraphAddKernelNode failed: %s", cudaGetErrorString(e)); }
  cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
  cp.conditional.handle = handle;
  cp.conditional.type = cudaGraphCondTypeWhile;
  cp.conditional.size = 1;
  e = cudaGraphAddNode(&n_while, g, &n_check1, 1, &cp);
  if (e != cudaSuccess) { cudaGraphDestroy(g); return fail(h, SBX_E_CUDA, "conditional graph node failed: %s", cudaGetErrorString(e)); }
  cudaGraph_t body = cp.conditional.phGraph_out[0];
  e = add_kernel(body, nullptr, tma ? sweep_tma_kernel(0) : sweep_kernel(h->V, tl.rows_per_warp, 0), sweep_grid,
                 dim3(kStreamThreads), a_sweep, &n_sweep, tma ? (unsigned)sweep_tma_smem(false) : 0u);
  if (e == cudaSuccess) e = add_kernel(body, &n_sweep, (void*)k_check_loop, dim3(1), dim3(1024), a_check, &n_check);
  if (e != cudaSuccess) { cudaGraphDestroy(g); return fail(h, SBX_E_CUDA, "cudaGraphAddKernelNode (loop body) failed: %s", cudaGetErrorString(e)); }
  cudaGraphExec_t ex = nullptr;
  e = cudaGraphInstantiate(&ex, g, 0);
  if (e != cudaSuccess) { cudaGraphDestroy(g); return fail(h, SBX_E_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e)); }
  h->loop_graph[which] = g;
  h->loop_exec[which] = ex;
  return SBX_OK;
}

// Streaming Jacobi loop.  Default: the device-driven graph above (no host synchronisation,
// so sbx_step stays asynchronous on the caller's stream).  SBX_HOST_SWEEP_LOOP=1 or a driver
// without conditional graph nodes falls back to one launch per sweep + a pinned-memory poll.
int run_stream_sweeps(sbx_handle h, cudaStream_t st) {
  const Params& p = h->P;
  if (h->plans_dirty) {
    const size_t total = (size_t)p.n_plans * p.H * p.W;
    const unsigned g = (unsigned)((total + 256 * 8 - 1) / (256 * 8));
    k_pack_stream<<<g > 0 ? g : 1, 256, 0, st>>>(p);
    if (int rc = launch_check(h, "k_pack_stream")) return rc;
    h->plans_dirty = 0;
  }
  if (!h->loop_graph_failed && !getenv("SBX_HOST_SWEEP_LOOP")) {
    const int which = p.fd_only ? 1 : 0;
    if (!h->loop_exec[which] && build_sweep_graph(h, which) != SBX_OK) h->loop_graph_failed = 1;
    if (h->loop_exec[which]) {
      if (int rc = timing_record(h, h->t_solve, h->t_solve_used, st)) return rc;
      CUDA_TRY(h, cudaGraphLaunch(h->loop_exec[which], st));
      h->launches += 1;                  // k_loop_begin; the loop's k_sweep / k_check_loop pairs are counted on the device
      ++h->graph_launches;
      return timing_record(h, h->t_solve, h->t_solve_used, st);
    }
  }
  const unsigned gb = (unsigned)((p.B + 255) / 256);
  k_activate<<<gb, 256, 0, st>>>(p);
  if (int rc = launch_check(h, "k_activate")) return rc;
  const StreamTiling tl = stream_tiling(p.H, p.W, h->V);
  const unsigned grid = (unsigned)((size_t)tl.tiles * p.B);
  for (int k = 1; k <= p.iteration_limit; ++k) {
    if (int rc = timing_record(h, h->t_solve, h->t_solve_used, st)) return rc;
    {
      Params pk = p;
      int kk = k;
      void* args[] = {&pk, &kk, &h->sweep_tm};
      if (h->sweep_tma)
        CUDA_TRY(h, cudaLaunchKernel(sweep_tma_kernel(k > 1 ? 0 : 1), dim3(grid), dim3(kStreamThreads), args,
                                     sweep_tma_smem(k == 1), st));
      else
        CUDA_TRY(h, cudaLaunchKernel(sweep_kernel(h->V, tl.rows_per_warp, k > 1 ? 0 : 1),
                                     dim3(grid), dim3(kStreamThreads), args, 0, st));
    }
    if (int rc = launch_check(h, "k_sweep")) return rc;
    if (int rc = timing_record(h, h->t_solve, h->t_solve_used, st)) return rc;
    CUDA_TRY(h, cudaMemsetAsync(p.n_active, 0, sizeof(int32_t), st));
    k_check<<<gb, 256, 0, st>>>(p, k);
    if (int rc = launch_check(h, "k_check")) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(h->h_n_active, p.n_active, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(h, cudaStreamSynchronize(st));
    if (*h->h_n_active == 0) break;
  }
  return SBX_OK;
}

// Tensor map of the [B, H, W] fp32 temperature field with a [1, H, P] box (P > W:
// the extra columns are zero-filled on load, clipped on store).  The driver entry
// point is resolved through the runtime, so libsbx does not link libcuda.
int make_field_tensor_map(sbx_handle h, CUtensorMap* out, float* base, int B, int H, int W, int box_w, int box_h);
int make_plane_tensor_map(sbx_handle h, float* base, int B, int H, int W, int P) {
  return make_field_tensor_map(h, &h->tmap_t, base, B, H, W, P, H);
}
int make_field_tensor_map(sbx_handle h, CUtensorMap* out, float* base, int B, int H, int W, int box_w, int box_h) {
  const int P = box_w;
  typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
    return fail(h, SBX_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t gstride[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
  const cuuint32_t box[3] = {(cuuint32_t)P, (cuuint32_t)box_h, 1};
  const cuuint32_t estride[3] = {1, 1, 1};
  const CUresult r = reinterpret_cast<EncodeTiled>(fn)(
      out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, gdim, gstride, box, estride,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, SBX_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return SBX_OK;
}

// v1 structures (packed descriptor plane, vector list, zone-sum list) exist only when
// k_resident_step is the kernel in use
int ensure_v1_buffers(sbx_handle h) {
  if (h->v1_allocated) return SBX_OK;
  Params& p = h->P;
  const sbx_config& c = h->cfg;
  uint16_t* dp = nullptr; uint16_t* ql = nullptr; int32_t* nf = nullptr; uint32_t* rl = nullptr; int32_t* rc = nullptr;
  if (int rc_ = dev_alloc<uint16_t>(h, &dp, (size_t)c.n_plans * p.geom.desc_stride)) return rc_;
  if (int rc_ = dev_alloc<uint16_t>(h, &ql, (size_t)c.n_plans * p.geom.list_stride)) return rc_;
  if (int rc_ = dev_alloc<int32_t>(h, &nf, (size_t)c.n_plans * 4)) return rc_;
  if (int rc_ = dev_alloc<uint32_t>(h, &rl, (size_t)c.n_plans * p.geom.rl_cap)) return rc_;
  if (int rc_ = dev_alloc<int32_t>(h, &rc, (size_t)c.n_plans)) return rc_;
  p.desc_packed = dp; p.qlist = ql; p.n_fast = nf; p.rlist = rl; p.rl_chunks = rc;
  h->v1_allocated = 1;
  return SBX_OK;
}


int prepare_plans(sbx_handle h, cudaStream_t st) {
  Params& p = h->P;
  cudaError_t e;
  if (h->plans_dirty) { h->v1_ready = 0; h->v23_ready = 0; h->plans_dirty = 0; }
  if (!h->v23_ready) {
    h->use_v2 = 0;
    h->use_v3 = 0;
    if (h->v3_capable) {
      k_prepare_plan3<<<p.n_plans, kPrepThreads, 0, st>>>(p);
      if (int rc = launch_check(h, "k_prepare_plan3")) return rc;
      // any plan that exceeds a capacity (pair patterns, zone parts per tile) sends the handle
      // back to k_resident_step
      std::vector<int32_t> counts((size_t)p.n_plans * 4);
      CUDA_TRY(h, cudaMemcpyAsync(counts.data(), p.counts3, counts.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      CUDA_TRY(h, cudaStreamSynchronize(st));
      h->use_v3 = 1;
      for (int i = 0; i < p.n_plans; ++i)
        if (counts[(size_t)i * 4 + kC3Capable] == 0) { h->use_v3 = 0; break; }
    }
    if (!h->use_v3 && h->v2_capable) {
      const size_t sm = prepare2_smem(p.H * (p.W / 4), p.Z);
      e = cudaFuncSetAttribute(k_prepare_plan2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      if (e != cudaSuccess) return fail(h, SBX_E_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      k_prepare_plan2<<<p.n_plans, kPrepThreads, sm, st>>>(p);
      if (int rc = launch_check(h, "k_prepare_plan2")) return rc;
      // any plan whose lists do not fit sends the handle back to k_resident_step
      std::vector<int32_t> counts((size_t)p.n_plans * 8);
      CUDA_TRY(h, cudaMemcpyAsync(counts.data(), p.counts2, counts.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      CUDA_TRY(h, cudaStreamSynchronize(st));
      h->use_v2 = 1;
      for (int i = 0; i < p.n_plans; ++i)
        if (counts[(size_t)i * 8 + 7] < 0) { h->use_v2 = 0; break; }
    }
    h->v23_ready = 1;
  }
  const bool need_v1 = (!h->use_v2 && !h->use_v3) || (h->use_v3 && step_has_convection(p));
  if (need_v1 && !h->v1_ready) {
    if (int rc = ensure_v1_buffers(h)) return rc;
    const size_t smem = (sizeof(uint32_t) + sizeof(uint16_t)) * (size_t)(p.H * p.W / h->V);
    if (h->V == 4) e = cudaFuncSetAttribute(k_prepare_plan<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    else e = cudaFuncSetAttribute(k_prepare_plan<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(h, SBX_E_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    if (h->V == 4) k_prepare_plan<4><<<p.n_plans, kPrepThreads, smem, st>>>(p);
    else k_prepare_plan<1><<<p.n_plans, kPrepThreads, smem, st>>>(p);
    if (int rc = launch_check(h, "k_prepare_plan")) return rc;
    const size_t n_items = (size_t)(p.H * p.W / h->V);
    const size_t sm2 = sizeof(uint2) * n_items + sizeof(int) * 2 * (size_t)(p.Z + 1);
    if (h->V == 4) e = cudaFuncSetAttribute(k_prepare_reduce<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
    else e = cudaFuncSetAttribute(k_prepare_reduce<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
    if (e != cudaSuccess) return fail(h, SBX_E_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    if (h->V == 4) k_prepare_reduce<4><<<p.n_plans, kPrepThreads, sm2, st>>>(p);
    else k_prepare_reduce<1><<<p.n_plans, kPrepThreads, sm2, st>>>(p);
    if (int rc = launch_check(h, "k_prepare_reduce")) return rc;
    h->v1_ready = 1;
  }
  return SBX_OK;
}

// The solve of buildings [p.b_begin, p.b_end).  `with_header`: the per-building solve
// header has not been built by k_pre (sbx_fd_step, which has no HVAC prologue).
int run_resident(sbx_handle h, cudaStream_t st, bool with_header) {
  h->P.v2_stride = v2_next_distance(h);
  const Params& p = h->P;
  if (h->cfg.solver == SBX_SOLVER_GAUSS_SEIDEL) {
    if (int rc = timing_record(h, h->t_solve, h->t_solve_used, st)) return rc;
    if (h->gs_global) k_resident_gs<true><<<p.B, kGsGlobalThreads, h->gs_smem, st>>>(p);
    else k_resident_gs<false><<<p.B, kGsThreads, h->gs_smem, st>>>(p);
    if (int rc = launch_check(h, "k_resident_gs")) return rc;
    return timing_record(h, h->t_solve, h->t_solve_used, st);
  }
  if (with_header) {
    const int wpb = 4;
    k_build_header<<<(unsigned)((p.b_end - p.b_begin + wpb - 1) / wpb), wpb * 32, 0, st>>>(p);
    if (int rc = launch_check(h, "k_build_header")) return rc;
  }
  const unsigned grid = (unsigned)(p.b_end - p.b_begin);
  if (int rc = timing_record(h, h->t_solve, h->t_solve_used, st)) return rc;
  if (h->use_v3 && !step_has_convection(p)) {
    k_resident_step3<<<grid, (unsigned)p.g3.nt, p.g3.total, st>>>(p, h->tmap_t);
    h->last_resident_kernel = 3;
    if (int rc = launch_check(h, "k_resident_step3")) return rc;
    return timing_record(h, h->t_solve, h->t_solve_used, st);
  }
  if (h->use_v2) {
    // persistent CTAs: one per resident slot, each walks buildings b, b + grid, ...
    k_resident_step2<<<v2_grid(h), kR2Threads, p.g2.total, st>>>(p, h->tmap_t);
    h->last_resident_kernel = 2;
    if (int rc = launch_check(h, "k_resident_step2")) return rc;
    return timing_record(h, h->t_solve, h->t_solve_used, st);
  }
  if (p.pw_on && !p.fd_only && h->pw_dirty) if (int rc = build_pairwise_tables(h, st)) return rc;
  if (h->V == 4 && p.pw_on && p.pw_fused && !p.fd_only) {
    k_resident_step<4, true><<<grid, kResidentThreads, h->resident_smem, st>>>(p, h->tmap_t);
    h->pw_by_solve = 1;            // launch_post need not take the means again
  } else if (h->V == 4) k_resident_step<4><<<grid, kResidentThreads, h->resident_smem, st>>>(p, h->tmap_t);
  else k_resident_step<1><<<grid, kResidentThreads, h->resident_smem, st>>>(p, h->tmap_t);
  h->last_resident_kernel = 1;
  if (int rc = launch_check(h, "k_resident_step")) return rc;
  return timing_record(h, h->t_solve, h->t_solve_used, st);
}

int do_reset(sbx_handle h, float* obs, float* reward, int32_t* step_type, float* discount,
             cudaStream_t st) {
  Params& p = h->P;
  p.fd_only = 0;
  p.time_index = 0; p.step_count = 0;
  p.therm_seen = h->therm_seen; p.prev_comfort = h->prev_comfort;
  p.action = nullptr; p.obs = obs; p.reward = reward; p.step_type = step_type; p.discount_out = discount;
  p.b_begin = 0; p.b_end = p.B; p.build_hdr = 0;
  const unsigned gb = (unsigned)((p.B + 255) / 256);
  k_reset_state<<<gb, 256, 0, st>>>(p);
  if (int rc = launch_check(h, "k_reset_state")) return rc;
  const size_t total = (size_t)p.B * p.H * p.W;
  const unsigned gt = (unsigned)((total + 256 * 8 - 1) / (256 * 8));
  k_reset_temp<<<gt > 0 ? gt : 1, 256, 0, st>>>(p);
  if (int rc = launch_check(h, "k_reset_temp")) return rc;
  if (int rc = launch_zone_reduce(h, st)) return rc;
  if (int rc = launch_post(h, st, 1)) return rc;
  h->step_count = 0; h->time_index = 0; h->episode_ended = 0; h->reset_called = 1;
  return SBX_OK;
}

// `shares` > 1 (sbx_step_host): the batch is stepped as that many contiguous shares, one after
// the other on `st`, and `after_share(b0, b1)` is called when a share's kernels are queued --
// the host path starts that share's device->host copy on a second stream there, so that only
// the last share's copy is not hidden behind the following shares' kernels.
int do_step(sbx_handle h, const float* action, float* obs, float* reward, int32_t* step_type,
            float* discount, cudaStream_t st, int shares = 1,
            const std::function<int(int, int)>& after_share = nullptr) {
  if (!h->reset_called) return fail(h, SBX_E_STATE, "sbx_step before sbx_reset");
  if (h->episode_ended) return fail(h, SBX_E_STATE, "episode has ended; call sbx_reset (environment.py:1252)");
  if (h->cfg.n_actions > 0 && !action) return fail(h, SBX_E_INVALID, "action is NULL");
  if (h->time_index + 1 >= h->cfg.n_table_steps) return fail(h, SBX_E_STATE, "time index %d beyond the exogenous tables (%d rows)", h->time_index + 1, h->cfg.n_table_steps);
  Params& p = h->P;
  p.fd_only = 0;
  p.time_index = h->time_index; p.step_count = h->step_count;
  p.therm_seen = h->therm_seen; p.prev_comfort = h->prev_comfort;
  p.action = action; p.obs = obs; p.reward = reward; p.step_type = step_type; p.discount_out = discount;
  p.conv_perm = h->conv_pending ? h->conv_perm : nullptr;
  p.b_begin = 0; p.b_end = p.B; p.build_hdr = 0;
  const bool jacobi_resident = h->path == SBX_PATH_RESIDENT && h->cfg.solver == SBX_SOLVER_TF_JACOBI;
  if (jacobi_resident) if (int rc = prepare_plans(h, st)) return rc;
  if (int rc = timing_record(h, h->t_step, h->t_step_used, st)) return rc;
  if (jacobi_resident && h->n_chunks > 1) {
    // Pipelined step.  hi: pre(0) .. pre(n-1), post(0) .. post(n-1);  lo: solve(0) ..
    // solve(n-1);  solve(c) waits for pre(c), post(c) for solve(c).  The solves keep
    // every SM busy back to back; the HVAC kernels of the other shares run in the SM
    // slots that free up (high stream priority), instead of serialising before and
    // after one monolithic solve.
    const int n = h->n_chunks;
    auto range = [&](int c) {
      p.b_begin = (int)((int64_t)p.B * c / n);
      p.b_end = (int)((int64_t)p.B * (c + 1) / n);
    };
    CUDA_TRY(h, cudaEventRecord(h->ev_fork, st));
    CUDA_TRY(h, cudaStreamWaitEvent(h->s_hi, h->ev_fork, 0));
    p.build_hdr = 1;
    for (int c = 0; c < n; ++c) {
      range(c);
      if (int rc = launch_pre(h, h->s_hi)) return rc;
      CUDA_TRY(h, cudaEventRecord(h->ev_pre[c], h->s_hi));
    }
    for (int c = 0; c < n; ++c) {
      range(c);
      CUDA_TRY(h, cudaStreamWaitEvent(h->s_lo, h->ev_pre[c], 0));
      if (int rc = run_resident(h, h->s_lo, false)) return rc;
      CUDA_TRY(h, cudaEventRecord(h->ev_solve[c], h->s_lo));
    }
    for (int c = 0; c < n; ++c) {
      range(c);
      CUDA_TRY(h, cudaStreamWaitEvent(h->s_hi, h->ev_solve[c], 0));
      if (int rc = launch_post(h, h->s_hi, 0)) return rc;
    }
    CUDA_TRY(h, cudaEventRecord(h->ev_join, h->s_hi));
    CUDA_TRY(h, cudaStreamWaitEvent(st, h->ev_join, 0));
    p.b_begin = 0; p.b_end = p.B; p.build_hdr = 0;
    if (after_share) if (int rc = after_share(0, p.B)) return rc;
  } else if (jacobi_resident && shares > 1) {
    const int n = shares < p.B ? shares : p.B;
    // (uneven shares -- a smaller last share, whose copy is the exposed one -- measured the same:
    // 60 / 70 / 80 % first share: 26.1 / 26.4 / 26.6 M env-steps/s end to end against 26.3 M)
    for (int c = 0; c < n; ++c) {
      p.b_begin = (int)((int64_t)p.B * c / n);
      p.b_end = (int)((int64_t)p.B * (c + 1) / n);
      p.build_hdr = 1;
      if (int rc = launch_pre(h, st)) return rc;
      p.build_hdr = 0;
      if (int rc = run_resident(h, st, false)) return rc;
      if (int rc = launch_post(h, st, 0)) return rc;
      if (after_share) if (int rc = after_share(p.b_begin, p.b_end)) return rc;
    }
    p.b_begin = 0; p.b_end = p.B;
  } else {
    p.build_hdr = h->cfg.solver == SBX_SOLVER_TF_JACOBI ? 1 : 0;   // both Jacobi paths read the header
    if (int rc = launch_pre(h, st)) return rc;
    p.build_hdr = 0;
    if (h->path == SBX_PATH_RESIDENT) {
      if (int rc = run_resident(h, st, false)) return rc;      // solve + zone sums, one launch
    } else {
      if (int rc = run_stream_sweeps(h, st)) return rc;
      if (p.conv_perm) {
        const size_t total = (size_t)p.B * p.H * p.W;
        const unsigned g = (unsigned)((total + 256 * 8 - 1) / (256 * 8));
        k_permute<<<g > 0 ? g : 1, 256, 0, st>>>(p);
        if (int rc = launch_check(h, "k_permute")) return rc;
        k_rotate_cur<<<(unsigned)((p.B + 255) / 256), 256, 0, st>>>(p);
        if (int rc = launch_check(h, "k_rotate_cur")) return rc;
      }
      if (p.conv_p > 0.0 && !p.conv_perm) {
        // device-RNG convection fused with the zone reduction: one read of the field, the
        // sums, and the permuted field into the next buffer of the rotation
        CUDA_TRY(h, cudaMemsetAsync(p.zone_sum, 0, sizeof(long long) * (size_t)p.B * (p.Z + 1), st));
        const StreamTiling tl = stream_tiling(p.H, p.W, h->V);
        const unsigned grid = (unsigned)((size_t)tl.tiles * p.B);
        if (h->V == 4) k_convect_reduce<4><<<grid, kStreamThreads, 0, st>>>(p);
        else k_convect_reduce<1><<<grid, kStreamThreads, 0, st>>>(p);
        if (int rc = launch_check(h, "k_convect_reduce")) return rc;
        k_rotate_cur<<<(unsigned)((p.B + 255) / 256), 256, 0, st>>>(p);
        if (int rc = launch_check(h, "k_rotate_cur")) return rc;
      } else {
        if (int rc = launch_zone_reduce(h, st)) return rc;
      }
    }
    if (int rc = launch_post(h, st, 0)) return rc;
    if (after_share) if (int rc = after_share(0, p.B)) return rc;
  }
  if (int rc = timing_record(h, h->t_step, h->t_step_used, st)) return rc;
  p.conv_perm = nullptr;
  h->conv_pending = 0;
  h->pw_by_solve = 0;
  // host mirror of Thermostat._previous_timestamp (thermostat.py:147) and of the
  // episode counters (environment.py:1313, 1358-1359)
  h->therm_seen = 1;
  h->prev_comfort = h->h_comfort ? h->h_comfort[h->time_index] : 0;
  const bool ended = h->step_count >= h->cfg.episode_steps;
  h->episode_ended = ended ? 1 : 0;
  if (!ended) ++h->step_count;
  ++h->time_index;
  h->env_steps += h->cfg.n_envs;
  return SBX_OK;
}

struct FieldInfo {
  void* ptr;
  size_t bytes;
  bool writable;
};

}  // namespace

extern "C" {

const char* sbx_last_error(sbx_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int sbx_abi_info(int32_t* abi_version, size_t* config_bytes, size_t* info_bytes) {
  if (abi_version) *abi_version = SBX_ABI_VERSION;
  if (config_bytes) *config_bytes = sizeof(sbx_config);
  if (info_bytes) *info_bytes = sizeof(sbx_info);
  return SBX_OK;
}

int sbx_create(const sbx_config* cfg, int device, sbx_handle* out) {
  if (!cfg || !out) return fail(nullptr, SBX_E_INVALID, "null argument");
  *out = nullptr;
  if (int rc = validate(*cfg)) return rc;
  sbx_handle h = new (std::nothrow) sbx_env();
  if (!h) return fail(nullptr, SBX_E_NOMEM, "out of host memory");
  h->cfg = *cfg;
  h->device = device;
  h->D = obs_dim(*cfg);
  auto bail = [&](int rc) {
    g_create_error = h->err;
    sbx_destroy(h);
    return rc;
  };
  {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { fail(h, SBX_E_CUDA, "cudaSetDevice(%d) failed: %s", device, cudaGetErrorString(e)); return bail(SBX_E_CUDA); }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { fail(h, SBX_E_CUDA, "cudaGetDeviceProperties failed: %s", cudaGetErrorString(e)); return bail(SBX_E_CUDA); }
    h->n_sms = prop.multiProcessorCount;
    if (prop.major < 10) { fail(h, SBX_E_CUDA, "libsbx is built for sm_100a; device %d is sm_%d%d", device, prop.major, prop.minor); return bail(SBX_E_CUDA); }
  }
  const sbx_config& c = h->cfg;
  const size_t B = c.n_envs, Z = c.n_zones, N = (size_t)c.height * c.width, T = c.n_table_steps;
  h->V = (c.width % 4 == 0) ? 4 : 1;

  // path selection
  const ResidentGeom L = resident_geom(c.height, c.width, (int)Z, h->V);
  int max_optin = 0;
  cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  const bool fits = L.total <= max_optin && L.plane_cv / h->V <= 32767;
  if (c.kernel_path == SBX_PATH_RESIDENT && !fits && c.solver != SBX_SOLVER_GAUSS_SEIDEL) {
    fail(h, SBX_E_INVALID, "resident path needs %d B of shared memory per CTA; device allows %d", L.total, max_optin);
    return bail(SBX_E_INVALID);
  }
  h->path = (c.kernel_path == SBX_PATH_AUTO) ? (fits ? SBX_PATH_RESIDENT : SBX_PATH_STREAMING) : c.kernel_path;
  h->resident_smem = L.total;
  if (c.solver == SBX_SOLVER_GAUSS_SEIDEL) {
    // one CTA per building runs the raster sweep as an anti-diagonal wavefront: on the grid in
    // shared memory where it fits, else (sim_config_legacy.gin on the 744x1004 plan) on the fp64
    // field in global memory -- same arithmetic, a parity path
    GsLayout G = gs_layout((int)N, (int)Z);
    h->gs_global = (G.total > (size_t)max_optin || c.kernel_path == SBX_PATH_STREAMING) ? 1 : 0;
    if (h->gs_global) G = gs_layout(0, (int)Z);
    h->path = SBX_PATH_RESIDENT;
    h->gs_smem = G.total;
    cudaError_t e = h->gs_global
        ? cudaFuncSetAttribute(k_resident_gs<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.total)
        : cudaFuncSetAttribute(k_resident_gs<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.total);
    if (e != cudaSuccess) { fail(h, SBX_E_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return bail(SBX_E_CUDA); }
  }
  if (h->path == SBX_PATH_RESIDENT && c.solver != SBX_SOLVER_GAUSS_SEIDEL) {
    cudaError_t e;
    if (h->V == 4) {
      e = cudaFuncSetAttribute(k_resident_step<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(k_resident_step<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
    }
    else e = cudaFuncSetAttribute(k_resident_step<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
    if (e != cudaSuccess) { fail(h, SBX_E_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return bail(SBX_E_CUDA); }
    int nb = 0;
    if (h->V == 4) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_resident_step<4>, kResidentThreads, L.total);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_resident_step<1>, kResidentThreads, L.total);
    h->resident_ctas_per_sm = nb;
  }

  Params& p = h->P;
  memset(&p, 0, sizeof(p));
  p.geom = L;
  fill_params(h);
#define ALLOC(field, type, count)                                           \
  do {                                                                      \
    type* tmp__ = nullptr;                                                  \
    if (int rc = dev_alloc<type>(h, &tmp__, (count))) return bail(rc);      \
    field = tmp__;                                                          \
  } while (0)
  ALLOC(p.desc, uint16_t, (size_t)c.n_plans * N);
  ALLOC(p.material, double, (size_t)c.n_plans * 9);
  ALLOC(p.cv_size, double, (size_t)c.n_plans);
  ALLOC(p.zone_ncv, int32_t, (size_t)c.n_plans * Z);
  ALLOC(p.zone_ndiff, int32_t, (size_t)c.n_plans * Z);
  ALLOC(p.obs_zone_order, int32_t, (size_t)c.n_plans * Z);
  ALLOC(p.desc_spk, uint16_t, h->path == SBX_PATH_STREAMING ? (size_t)c.n_plans * N : 1);
  h->v2_capable = 0;
  // k_resident_step2 is opt-in (SBX_RESIDENT_V2=1): measured on B200 (32768 x 64x96) it ties
  // k_resident_step at 2.5 sweeps per step (0.955 ms both) and loses 3 % at 2.6 (1.010 vs
  // 0.981 ms): both are bound by the shared-memory pipe during the sweeps, see DESIGN.md 6
  if (h->path == SBX_PATH_RESIDENT && c.solver == SBX_SOLVER_TF_JACOBI && h->V == 4 && L.use_tmap &&
      getenv("SBX_RESIDENT_V2") && atoi(getenv("SBX_RESIDENT_V2")) != 0) {
    p.g2 = resident2_geom(L, c.height, c.width, (int)Z);
    if (p.g2.total <= max_optin) {
      cudaError_t e2 = cudaFuncSetAttribute(k_resident_step2, cudaFuncAttributeMaxDynamicSharedMemorySize, p.g2.total);
      int nb2 = 0;
      if (e2 == cudaSuccess) e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb2, k_resident_step2, kR2Threads, p.g2.total);
      if (e2 != cudaSuccess) { fail(h, SBX_E_CUDA, "k_resident_step2 set-up failed: %s", cudaGetErrorString(e2)); return bail(SBX_E_CUDA); }
      if (nb2 > 0) {
        h->v2_capable = 1;
        h->resident2_ctas_per_sm = nb2;
        ALLOC(p.flist2, uint16_t, (size_t)c.n_plans * L.list_stride);
        ALLOC(p.rec2, uint4, (size_t)c.n_plans * p.g2.rec_cap);
        ALLOC(p.sched2, uint32_t, (size_t)c.n_plans * r2_sched_cap(c.height * (c.width / 4)));
        ALLOC(p.counts2, int32_t, (size_t)c.n_plans * 8);
        ALLOC(p.zfull, uint16_t, (size_t)c.n_plans * p.g2.zfull_cap);
        ALLOC(p.zpart, uint32_t, (size_t)c.n_plans * p.g2.zpart_cap);
        ALLOC(p.zchunk, uint8_t, (size_t)c.n_plans * p.g2.zchunk_cap);
      }
    }
  }
  // k_resident_step3 is opt-in (SBX_RESIDENT_V3=1): measured on B200 (32768 x 64x96) it needs 1.06 ms
  // per launch against 0.96 ms of k_resident_step -- 24 % fewer instructions, but its phases are
  // short and separated by barriers, and at 72 registers the two half-tiles of a thread cannot be
  // interleaved, so it issues on 50 % of the cycles where k_resident_step reaches 77 %
  h->v3_capable = 0;
  if (h->path == SBX_PATH_RESIDENT && c.solver == SBX_SOLVER_TF_JACOBI && !h->v2_capable && Z < 255 &&
      L.use_tmap && resident3_supported(c.height, c.width) && getenv("SBX_RESIDENT_V3") && atoi(getenv("SBX_RESIDENT_V3")) != 0) {
    p.g3 = resident3_geom(c.height, c.width, (int)Z, L.P);
    if (p.g3.total <= max_optin) {
      cudaError_t e3 = cudaFuncSetAttribute(k_resident_step3, cudaFuncAttributeMaxDynamicSharedMemorySize, p.g3.total);
      int nb3 = 0;
      if (e3 == cudaSuccess) e3 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb3, k_resident_step3, p.g3.nt, p.g3.total);
      if (e3 != cudaSuccess) { fail(h, SBX_E_CUDA, "k_resident_step3 set-up failed: %s", cudaGetErrorString(e3)); return bail(SBX_E_CUDA); }
      if (nb3 > 0) {
        h->v3_capable = 1;
        h->resident3_ctas_per_sm = nb3;
        ALLOC(p.ent3, uint32_t, (size_t)c.n_plans * 2 * p.g3.nt);
        ALLOC(p.rec3, uint4, (size_t)c.n_plans * 2 * p.g3.nt);
        ALLOC(p.recz3, uint2, (size_t)c.n_plans * 2 * p.g3.nt);
        ALLOC(p.pat3, uint16_t, (size_t)c.n_plans * p.g3.pat_cap);
        ALLOC(p.counts3, int32_t, (size_t)c.n_plans * 4);
      }
    }
  }
  ALLOC(p.hdr, unsigned char, B * header_bytes((int)Z));
  ALLOC(p.reset_temps, float, (size_t)c.n_reset * N);
  ALLOC(p.initial_temp, float, B);
  ALLOC(p.ambient, double, (size_t)c.n_weather * T);
  ALLOC(p.convection, double, (size_t)c.n_weather);
  ALLOC(p.comfort, uint8_t, T);
  ALLOC(p.comfort_soon, uint8_t, T);
  ALLOC(p.occ_reward, double, T * c.n_occ_zones);
  ALLOC(p.occ_obs, int32_t, T);
  ALLOC(p.occ_obs_zone, double, T * c.n_occ_zones);
  ALLOC(p.price_e, double, T);
  ALLOC(p.carbon_e, double, T);
  ALLOC(p.price_g, double, T);
  ALLOC(p.time_feat, double, T * 4);
  ALLOC(h->d_obs_mean, double, SBX_N_DEVICE_FIELDS);
  ALLOC(h->d_obs_var, double, SBX_N_DEVICE_FIELDS);
  ALLOC(h->d_obs_sd, double, SBX_N_DEVICE_FIELDS);
  ALLOC(h->d_hist_bins, double, 3 * SBX_MAX_HIST_BINS);
  const int n_tbuf = h->path == SBX_PATH_RESIDENT ? 1 : 3;
  for (int i = 0; i < n_tbuf; ++i) ALLOC(p.tbuf[i], float, B * N);
  for (int i = n_tbuf; i < 3; ++i) p.tbuf[i] = p.tbuf[0];
  memset(&h->tmap_t, 0, sizeof(h->tmap_t));
  if (h->path == SBX_PATH_RESIDENT && p.geom.use_tmap) {
    if (int rc = make_plane_tensor_map(h, p.tbuf[0], (int)B, c.height, c.width, p.geom.P)) return bail(rc);
  }
  memset(&h->sweep_tm, 0, sizeof(h->sweep_tm));
  const char* tma_env = getenv("SBX_SWEEP_TMA");          // SBX_SWEEP_TMA=0: the plain-load k_sweep
  if (h->path == SBX_PATH_STREAMING && h->V == 4 && c.solver == SBX_SOLVER_TF_JACOBI && !(tma_env && tma_env[0] == '0')) {
    // TMA-staged sweep (sbx_sweep_tma.cuh): per rotation buffer one tensor map with the haloed
    // T_est box and one with the T_prev box
    for (int i = 0; i < 3; ++i) {
      if (int rc = make_field_tensor_map(h, &h->sweep_tm.in[i], p.tbuf[i], (int)B, c.height, c.width, kTmaBoxW, kTmaChunk + 2)) return bail(rc);
      if (int rc = make_field_tensor_map(h, &h->sweep_tm.prev[i], p.tbuf[i], (int)B, c.height, c.width, 128, kTmaChunk)) return bail(rc);
    }
    cudaError_t e = cudaFuncSetAttribute(k_sweep_tma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sweep_tma_smem(false));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sweep_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sweep_tma_smem(true));
    if (e != cudaSuccess) { fail(h, SBX_E_CUDA, "cudaFuncSetAttribute(k_sweep_tma) failed: %s", cudaGetErrorString(e)); return bail(SBX_E_CUDA); }
    h->sweep_tma = 1;
  }
  ALLOC(p.cur, uint8_t, B);
  ALLOC(p.zone_mean, float, B * Z);
  ALLOC(p.global_mean, float, B);
  ALLOC(p.qcv, float, B * Z);
  ALLOC(p.qcv_next, float, B * Z);
  ALLOC(p.qcv64, double, B * Z);
  ALLOC(p.qcv64_next, double, B * Z);
  if (c.solver == SBX_SOLVER_GAUSS_SEIDEL) ALLOC(p.temp64, double, B * N);
  if (h->gs_global) ALLOC(p.temp64_prev, double, B * N);
  ALLOC(p.therm_mode, uint8_t, B * Z);
  ALLOC(p.ahu_heat_sp, double, B);
  ALLOC(p.ahu_cool_sp, double, B);
  ALLOC(p.boiler_sp, double, B);
  ALLOC(p.boiler_tank, double, B * 3);
  ALLOC(p.pre_zone_mean, float, B * Z);
  ALLOC(p.diag, double, B * SBX_DIAG_N);
  ALLOC(p.q_zone, double, B * Z);
  ALLOC(p.zone_supply, double, B * Z);
  ALLOC(p.n_sweeps, int32_t, B);
  ALLOC(p.max_delta_bits, uint32_t, B);
  ALLOC(p.max_delta, float, B);
  ALLOC(p.zone_sum, long long, B * (Z + 1));
  ALLOC(p.zone_ref, float, B);
  ALLOC(p.active, uint8_t, B);
  ALLOC(p.n_active, int32_t, 1);
  ALLOC(p.sweeps_total, unsigned long long, 1);
  ALLOC(p.sweep_k, int32_t, 1);
  ALLOC(p.sweep_launches, unsigned long long, 1);
  ALLOC(p.phase_cycles, unsigned long long, SBX_PHASE_WORDS);
  ALLOC(h->carry, CarryStore, B);
  ALLOC(h->fd_ambient, double, B);
  ALLOC(h->fd_convection, double, B);
  ALLOC(h->d_action, float, B * (c.n_actions > 0 ? c.n_actions : 1));
  // one block, [obs | reward | step_type | discount]: one DMA moves a whole TimeStep
  ALLOC(h->d_obs, float, B * (h->D + 3));
  h->d_reward = h->d_obs + B * h->D;
  h->d_step_type = reinterpret_cast<int32_t*>(h->d_reward + B);
  h->d_discount = h->d_reward + 2 * B;
#undef ALLOC
  p.obs_mean = h->d_obs_mean; p.obs_var = h->d_obs_var; p.hist_bins = h->d_hist_bins;
  p.obs_sd = h->d_obs_sd;
  p.fd_ambient = h->fd_ambient; p.fd_convection = h->fd_convection;
  {
    cudaError_t e = cudaMemcpy(h->d_obs_mean, c.obs_mean, sizeof(c.obs_mean), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_obs_var, c.obs_variance, sizeof(c.obs_variance), cudaMemcpyHostToDevice);
    double sd[SBX_N_DEVICE_FIELDS];
    for (int i = 0; i < SBX_N_DEVICE_FIELDS; ++i) sd[i] = c.obs_variance[i] > 0.0 ? sqrt(c.obs_variance[i]) : 1.0;
    if (e == cudaSuccess) e = cudaMemcpy(h->d_obs_sd, sd, sizeof(sd), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_hist_bins, c.hist_bins, sizeof(c.hist_bins), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    int prio_lo = 0, prio_hi = 0;
    if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&h->s_hi, cudaStreamNonBlocking, prio_hi);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&h->s_lo, cudaStreamNonBlocking, prio_lo);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->s_copy, cudaStreamNonBlocking);
    for (int i = 0; i < SBX_MAX_CHUNKS && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&h->ev_share[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
    for (int i = 0; i < SBX_MAX_CHUNKS && e == cudaSuccess; ++i) {
      e = cudaEventCreateWithFlags(&h->ev_pre[i], cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_solve[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaMallocHost(&h->h_action, sizeof(float) * B * (c.n_actions > 0 ? c.n_actions : 1));
    if (e == cudaSuccess) e = cudaMallocHost(&h->h_obs, sizeof(float) * B * h->D);
    if (e == cudaSuccess) e = cudaMallocHost(&h->h_reward, sizeof(float) * B);
    if (e == cudaSuccess) e = cudaMallocHost(&h->h_step_type, sizeof(int32_t) * B);
    if (e == cudaSuccess) e = cudaMallocHost(&h->h_discount, sizeof(float) * B);
    if (e == cudaSuccess) e = cudaMallocHost(&h->h_n_active, sizeof(int32_t));
    if (e != cudaSuccess) { fail(h, SBX_E_CUDA, "host-side setup failed: %s", cudaGetErrorString(e)); return bail(SBX_E_CUDA); }
  }
  // One launch per kernel by default: measured on B200 (32768 x 64x96), pipelining the
  // step over 2 / 4 / 8 shares of the batch (SBX_OPT_PIPELINE_CHUNKS) is 3 / 7 / 14 % slower
  // than the monolithic step -- the cross-stream event hand-offs cost more than the
  // exposed HVAC kernels they hide.  The option stays for callers with other shapes.
  h->n_chunks = 1;
  if (h->path == SBX_PATH_RESIDENT && c.solver == SBX_SOLVER_TF_JACOBI) {
    const int per_sm = h->v3_capable ? h->resident3_ctas_per_sm
                                     : h->v2_capable ? h->resident2_ctas_per_sm : h->resident_ctas_per_sm;
    h->P.prefetch_dist = h->n_sms * (per_sm > 0 ? per_sm : 1);
  }
  h->h_comfort = new (std::nothrow) uint8_t[T]();
  if (!h->h_comfort) { fail(h, SBX_E_NOMEM, "out of host memory"); return bail(SBX_E_NOMEM); }
  *out = h;
  return SBX_OK;
}

int sbx_destroy(sbx_handle h) {
  if (!h) return SBX_OK;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (int i = 0; i < h->n_all; ++i) cudaFree(h->all[i].p);
  if (h->gather) cudaFree(h->gather);
  if (h->conv_perm) cudaFree(h->conv_perm);
  cudaFreeHost(h->h_action); cudaFreeHost(h->h_obs); cudaFreeHost(h->h_reward);
  cudaFreeHost(h->h_step_type); cudaFreeHost(h->h_discount); cudaFreeHost(h->h_n_active);
  for (int i = 0; i < 2; ++i) {
    if (h->loop_exec[i]) cudaGraphExecDestroy(h->loop_exec[i]);
    if (h->loop_graph[i]) cudaGraphDestroy(h->loop_graph[i]);
  }
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->s_hi) cudaStreamDestroy(h->s_hi);
  if (h->s_lo) cudaStreamDestroy(h->s_lo);
  if (h->s_copy) cudaStreamDestroy(h->s_copy);
  for (int i = 0; i < SBX_MAX_CHUNKS; ++i) if (h->ev_share[i]) cudaEventDestroy(h->ev_share[i]);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  for (int i = 0; i < SBX_MAX_CHUNKS; ++i) {
    if (h->ev_pre[i]) cudaEventDestroy(h->ev_pre[i]);
    if (h->ev_solve[i]) cudaEventDestroy(h->ev_solve[i]);
  }
  for (cudaEvent_t e : h->t_step) cudaEventDestroy(e);
  for (cudaEvent_t e : h->t_solve) cudaEventDestroy(e);
  delete[] h->h_comfort;
  delete h;
  return SBX_OK;
}

int sbx_get_info(sbx_handle h, sbx_info* out) {
  if (!h || !out) return fail(h, SBX_E_INVALID, "null argument");
  CUDA_TRY(h, cudaSetDevice(h->device));
  memset(out, 0, sizeof(*out));
  out->obs_dim = h->D;
  out->kernel_path = h->path;
  out->n_sms = h->n_sms;
  out->resident_ctas_per_sm = h->use_v3 ? h->resident3_ctas_per_sm : h->resident_ctas_per_sm;
  out->resident_kernel = h->last_resident_kernel;
  out->device_bytes = h->device_bytes;
  unsigned long long sw = 0, loops = 0;
  CUDA_TRY(h, cudaDeviceSynchronize());
  CUDA_TRY(h, cudaMemcpy(&sw, h->P.sweeps_total, sizeof(sw), cudaMemcpyDeviceToHost));
  CUDA_TRY(h, cudaMemcpy(&loops, h->P.sweep_launches, sizeof(loops), cudaMemcpyDeviceToHost));
  out->kernel_launches = h->launches + 2 * (int64_t)loops;   // device-driven loop: k_sweep + k_check_loop per iteration
  out->sweeps_total = (int64_t)sw;
  out->env_steps_total = h->env_steps;
  out->step_count = h->step_count;
  out->episode_ended = h->episode_ended;
  out->time_index = h->time_index;
  return SBX_OK;
}

static int field_info(sbx_handle h, int field, FieldInfo* fi) {
  const sbx_config& c = h->cfg;
  const Params& p = h->P;
  const size_t B = c.n_envs, Z = c.n_zones, N = (size_t)c.height * c.width, T = c.n_table_steps;
#define F(id, ptr_, count, type, w) case id: fi->ptr = (void*)(ptr_); fi->bytes = (size_t)(count) * sizeof(type); fi->writable = w; return SBX_OK
  switch (field) {
    F(SBX_F_PLAN_DESC, p.desc, (size_t)c.n_plans * N, uint16_t, true);
    F(SBX_F_PLAN_MATERIAL, p.material, (size_t)c.n_plans * 9, double, true);
    F(SBX_F_PLAN_CV_SIZE, p.cv_size, c.n_plans, double, true);
    F(SBX_F_ZONE_NCV, p.zone_ncv, (size_t)c.n_plans * Z, int32_t, true);
    F(SBX_F_ZONE_NDIFF, p.zone_ndiff, (size_t)c.n_plans * Z, int32_t, true);
    F(SBX_F_OBS_ZONE_ORDER, p.obs_zone_order, (size_t)c.n_plans * Z, int32_t, true);
    F(SBX_F_RESET_TEMPS, p.reset_temps, (size_t)c.n_reset * N, float, true);
    F(SBX_F_INITIAL_TEMP, p.initial_temp, B, float, true);
    F(SBX_F_AMBIENT, p.ambient, (size_t)c.n_weather * T, double, true);
    F(SBX_F_CONVECTION, p.convection, c.n_weather, double, true);
    F(SBX_F_COMFORT, p.comfort, T, uint8_t, true);
    F(SBX_F_COMFORT_SOON, p.comfort_soon, T, uint8_t, true);
    F(SBX_F_OCC_REWARD, p.occ_reward, T * c.n_occ_zones, double, true);
    F(SBX_F_OCC_OBS, p.occ_obs, T, int32_t, true);
    F(SBX_F_OCC_OBS_ZONE, p.occ_obs_zone, T * c.n_occ_zones, double, true);
    F(SBX_F_PRICE_ELEC, p.price_e, T, double, true);
    F(SBX_F_CARBON_ELEC, p.carbon_e, T, double, true);
    F(SBX_F_PRICE_GAS, p.price_g, T, double, true);
    F(SBX_F_TIME_FEATURES, p.time_feat, T * 4, double, true);
    F(SBX_F_TEMP, p.tbuf[0], B * N, float, true);
    F(SBX_F_ZONE_MEAN, p.zone_mean, B * Z, float, true);
    F(SBX_F_GLOBAL_MEAN, p.global_mean, B, float, true);
    F(SBX_F_Q_CV, p.qcv, B * Z, float, true);
    F(SBX_F_THERMOSTAT_MODE, p.therm_mode, B * Z, uint8_t, true);
    F(SBX_F_AHU_HEATING_SP, p.ahu_heat_sp, B, double, true);
    F(SBX_F_AHU_COOLING_SP, p.ahu_cool_sp, B, double, true);
    F(SBX_F_BOILER_SP, p.boiler_sp, B, double, true);
    F(SBX_F_BOILER_TANK, p.boiler_tank, B * 3, double, true);
    F(SBX_F_Q_CV64, p.qcv64, B * Z, double, true);
    case SBX_F_TEMP64:
      if (!p.temp64) return fail(h, SBX_E_INVALID, "SBX_F_TEMP64 exists only with the Gauss-Seidel solver");
      fi->ptr = (void*)p.temp64; fi->bytes = B * N * sizeof(double); fi->writable = true; return SBX_OK;
    F(SBX_F_N_SWEEPS, p.n_sweeps, B, int32_t, false);
    F(SBX_F_MAX_DELTA, p.max_delta, B, float, false);
    F(SBX_F_STEP_DIAG, p.diag, B * SBX_DIAG_N, double, false);
    F(SBX_F_Q_ZONE, p.q_zone, B * Z, double, false);
    F(SBX_F_ZONE_SUPPLY_TEMP, p.zone_supply, B * Z, double, false);
    F(SBX_F_PRE_ZONE_MEAN, p.pre_zone_mean, B * Z, float, false);
    F(SBX_F_PHASE_CYCLES, p.phase_cycles, SBX_PHASE_WORDS, unsigned long long, true);
    default: break;
  }
#undef F
  return fail(h, SBX_E_INVALID, "unknown field %d", field);
}

int sbx_upload(sbx_handle h, int field, const void* src, size_t nbytes) {
  if (!h || !src) return fail(h, SBX_E_INVALID, "null argument");
  CUDA_TRY(h, cudaSetDevice(h->device));
  if (field == SBX_F_THERMOSTAT_PREV) {
    if (nbytes != 2 * sizeof(int32_t)) return fail(h, SBX_E_INVALID, "field %d needs %zu bytes, got %zu", field, 2 * sizeof(int32_t), nbytes);
    const int32_t* v = (const int32_t*)src;
    h->therm_seen = v[0] != 0; h->prev_comfort = v[1] != 0;
    return SBX_OK;
  }
  if (field == SBX_F_CONVECTION_PERM) {
    const size_t need = sizeof(int32_t) * (size_t)h->cfg.n_envs * h->cfg.height * h->cfg.width;
    if (nbytes != need) return fail(h, SBX_E_INVALID, "field %d needs %zu bytes, got %zu", field, need, nbytes);
    if (!h->conv_perm) {
      cudaError_t e = cudaMalloc((void**)&h->conv_perm, need);
      if (e != cudaSuccess) return fail(h, SBX_E_NOMEM, "cudaMalloc(%zu) failed: %s", need, cudaGetErrorString(e));
      h->device_bytes += (int64_t)need;
    }
    CUDA_TRY(h, cudaDeviceSynchronize());
    CUDA_TRY(h, cudaMemcpy(h->conv_perm, src, need, cudaMemcpyHostToDevice));
    h->conv_pending = 1;
    return SBX_OK;
  }
  if (field == SBX_F_EPISODE) {
    if (nbytes != 4 * sizeof(int32_t)) return fail(h, SBX_E_INVALID, "field %d needs %zu bytes, got %zu", field, 4 * sizeof(int32_t), nbytes);
    const int32_t* v = (const int32_t*)src;
    if (v[1] < 0 || v[1] >= h->cfg.n_table_steps) return fail(h, SBX_E_INVALID, "time_index %d outside the tables", v[1]);
    h->step_count = v[0]; h->time_index = v[1]; h->episode_ended = v[2] != 0; h->reset_called = v[3] != 0;
    return SBX_OK;
  }
  FieldInfo fi;
  if (int rc = field_info(h, field, &fi)) return rc;
  if (!fi.writable) return fail(h, SBX_E_INVALID, "field %d is read-only", field);
  if (nbytes != fi.bytes) return fail(h, SBX_E_INVALID, "field %d needs %zu bytes, got %zu", field, fi.bytes, nbytes);
  CUDA_TRY(h, cudaDeviceSynchronize());
  CUDA_TRY(h, cudaMemcpy(fi.ptr, src, nbytes, cudaMemcpyHostToDevice));
  if (field == SBX_F_TEMP) CUDA_TRY(h, cudaMemset(h->P.cur, 0, h->cfg.n_envs));
  if (h->P.temp64 && (field == SBX_F_TEMP || field == SBX_F_TEMP64)) {
    const size_t total = (size_t)h->cfg.n_envs * h->cfg.height * h->cfg.width;
    const unsigned g = (unsigned)((total + 255) / 256);
    k_mirror_temp<<<g, 256>>>(h->P, field == SBX_F_TEMP ? 1 : 0);
    if (int rc = launch_check(h, "k_mirror_temp")) return rc;
    CUDA_TRY(h, cudaDeviceSynchronize());
  }
  if (field == SBX_F_Q_CV && h->P.qcv64) {   // keep the fp64 copy coherent for Gauss-Seidel
    const unsigned g = (unsigned)(((size_t)h->cfg.n_envs * h->cfg.n_zones + 255) / 256);
    k_mirror_q<<<g, 256>>>(h->P);
    if (int rc = launch_check(h, "k_mirror_q")) return rc;
    CUDA_TRY(h, cudaDeviceSynchronize());
  }
  if (field == SBX_F_COMFORT) memcpy(h->h_comfort, src, nbytes);
  if (field == SBX_F_OCC_OBS_ZONE) h->P.occ_obs_per_env = 1;
  if (field == SBX_F_OCC_OBS) h->P.occ_obs_per_env = 0;
  if (field == SBX_F_PLAN_DESC) { h->plans_dirty = 1; h->pw_dirty = 1; }
  return SBX_OK;
}

int sbx_download(sbx_handle h, int field, void* dst, size_t nbytes) {
  if (!h || !dst) return fail(h, SBX_E_INVALID, "null argument");
  CUDA_TRY(h, cudaSetDevice(h->device));
  if (field == SBX_F_THERMOSTAT_PREV) {
    if (nbytes != 2 * sizeof(int32_t)) return fail(h, SBX_E_INVALID, "field %d needs %zu bytes, got %zu", field, 2 * sizeof(int32_t), nbytes);
    int32_t* v = (int32_t*)dst;
    v[0] = h->therm_seen; v[1] = h->prev_comfort;
    return SBX_OK;
  }
  if (field == SBX_F_EPISODE) {
    if (nbytes != 4 * sizeof(int32_t)) return fail(h, SBX_E_INVALID, "field %d needs %zu bytes, got %zu", field, 4 * sizeof(int32_t), nbytes);
    int32_t* v = (int32_t*)dst;
    v[0] = h->step_count; v[1] = h->time_index; v[2] = h->episode_ended; v[3] = h->reset_called;
    return SBX_OK;
  }
  FieldInfo fi;
  if (int rc = field_info(h, field, &fi)) return rc;
  if (nbytes != fi.bytes) return fail(h, SBX_E_INVALID, "field %d needs %zu bytes, got %zu", field, fi.bytes, nbytes);
  CUDA_TRY(h, cudaDeviceSynchronize());
  if (field == SBX_F_TEMP && h->path != SBX_PATH_RESIDENT) {
    if (!h->gather) {
      cudaError_t e = cudaMalloc((void**)&h->gather, fi.bytes);
      if (e != cudaSuccess) return fail(h, SBX_E_NOMEM, "cudaMalloc(%zu) failed: %s", fi.bytes, cudaGetErrorString(e));
    }
    const size_t total = fi.bytes / sizeof(float);
    const unsigned g = (unsigned)((total + 256 * 8 - 1) / (256 * 8));
    k_gather_temp<<<g > 0 ? g : 1, 256>>>(h->P, h->gather);
    if (int rc = launch_check(h, "k_gather_temp")) return rc;
    CUDA_TRY(h, cudaMemcpy(dst, h->gather, nbytes, cudaMemcpyDeviceToHost));
    return SBX_OK;
  }
  CUDA_TRY(h, cudaMemcpy(dst, fi.ptr, nbytes, cudaMemcpyDeviceToHost));
  return SBX_OK;
}

int sbx_reset(sbx_handle h, float* obs, float* reward, int32_t* step_type, float* discount, void* stream) {
  if (!h) return fail(h, SBX_E_INVALID, "null handle");
  CUDA_TRY(h, cudaSetDevice(h->device));
  return do_reset(h, obs, reward, step_type, discount, (cudaStream_t)stream);
}

int sbx_step(sbx_handle h, const float* action, float* obs, float* reward, int32_t* step_type,
             float* discount, void* stream) {
  if (!h) return fail(h, SBX_E_INVALID, "null handle");
  CUDA_TRY(h, cudaSetDevice(h->device));
  return do_step(h, action, obs, reward, step_type, discount, (cudaStream_t)stream);
}

static bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

// device -> host for one output; straight into the caller's buffer when it is
// page-locked (sbx_host_alloc), otherwise through the handle's pinned stage.
static int d2h(sbx_handle h, void* user, void* stage, const void* dev, size_t bytes, bool* staged) {
  *staged = !is_pinned(user);
  CUDA_TRY(h, cudaMemcpyAsync(*staged ? stage : user, dev, bytes, cudaMemcpyDeviceToHost, h->stream));
  return SBX_OK;
}

static int copy_out(sbx_handle h, float* obs, float* reward, int32_t* step_type, float* discount) {
  const size_t B = h->cfg.n_envs;
  // caller buffers laid out like the device block and page-locked: a single copy
  if (obs && reward && step_type && discount && reward == obs + B * h->D &&
      (void*)step_type == (void*)(reward + B) && discount == reward + 2 * B && is_pinned(obs)) {
    CUDA_TRY(h, cudaMemcpyAsync(obs, h->d_obs, sizeof(float) * B * (h->D + 3), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return SBX_OK;
  }
  bool s_obs = false, s_rew = false, s_st = false, s_dis = false;
  if (obs) if (int rc = d2h(h, obs, h->h_obs, h->d_obs, sizeof(float) * B * h->D, &s_obs)) return rc;
  if (reward) if (int rc = d2h(h, reward, h->h_reward, h->d_reward, sizeof(float) * B, &s_rew)) return rc;
  if (step_type) if (int rc = d2h(h, step_type, h->h_step_type, h->d_step_type, sizeof(int32_t) * B, &s_st)) return rc;
  if (discount) if (int rc = d2h(h, discount, h->h_discount, h->d_discount, sizeof(float) * B, &s_dis)) return rc;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  if (s_obs) memcpy(obs, h->h_obs, sizeof(float) * B * h->D);
  if (s_rew) memcpy(reward, h->h_reward, sizeof(float) * B);
  if (s_st) memcpy(step_type, h->h_step_type, sizeof(int32_t) * B);
  if (s_dis) memcpy(discount, h->h_discount, sizeof(float) * B);
  return SBX_OK;
}

int sbx_reset_host(sbx_handle h, float* obs, float* reward, int32_t* step_type, float* discount) {
  if (!h) return fail(h, SBX_E_INVALID, "null handle");
  CUDA_TRY(h, cudaSetDevice(h->device));
  if (int rc = do_reset(h, h->d_obs, h->d_reward, h->d_step_type, h->d_discount, h->stream)) return rc;
  return copy_out(h, obs, reward, step_type, discount);
}

// device -> host copies of buildings [b0, b1) of every requested output, on `cs`; straight into
// the caller's buffers when they are page-locked, otherwise into the handle's pinned stage
static int d2h_share(sbx_handle h, float* obs, float* reward, int32_t* step_type, float* discount,
                     const bool pinned[4], int b0, int b1, cudaStream_t cs) {
  const size_t D = (size_t)h->D, n = (size_t)(b1 - b0);
  if (obs) CUDA_TRY(h, cudaMemcpyAsync((pinned[0] ? obs : h->h_obs) + b0 * D, h->d_obs + b0 * D, sizeof(float) * n * D, cudaMemcpyDeviceToHost, cs));
  if (reward) CUDA_TRY(h, cudaMemcpyAsync((pinned[1] ? reward : h->h_reward) + b0, h->d_reward + b0, sizeof(float) * n, cudaMemcpyDeviceToHost, cs));
  if (step_type) CUDA_TRY(h, cudaMemcpyAsync((pinned[2] ? step_type : h->h_step_type) + b0, h->d_step_type + b0, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, cs));
  if (discount) CUDA_TRY(h, cudaMemcpyAsync((pinned[3] ? discount : h->h_discount) + b0, h->d_discount + b0, sizeof(float) * n, cudaMemcpyDeviceToHost, cs));
  return SBX_OK;
}

int sbx_step_host(sbx_handle h, const float* action, float* obs, float* reward, int32_t* step_type,
                  float* discount) {
  if (!h) return fail(h, SBX_E_INVALID, "null handle");
  CUDA_TRY(h, cudaSetDevice(h->device));
  const size_t B = h->cfg.n_envs;
  const int A = h->cfg.n_actions;
  if (A > 0) {
    if (!action) return fail(h, SBX_E_INVALID, "action is NULL");
    // BoundedActionNormalizer.setpoint_value (bounded_action_normalizer.py:84-90): an agent
    // action outside [-1, 1] (tolerance 1e-5) is an error, NaN included; one pass over the batch
    {
      const float lim = 1.0f + 1e-5f;
      const size_t n = B * (size_t)A;
      int bad = 0;
      for (size_t i = 0; i < n; ++i) bad |= !(fabsf(action[i]) <= lim);
      if (bad) {
        size_t i = 0;
        while (fabsf(action[i]) <= lim) ++i;
        return fail(h, SBX_E_INVALID, "agent_action: %.9g not within bounds [-1.0, 1.0] (env %zu, action %zu)",
                    (double)action[i], i / (size_t)A, i % (size_t)A);
      }
    }
    const float* src = action;
    if (!is_pinned(action)) {
      memcpy(h->h_action, action, sizeof(float) * B * A);
      src = h->h_action;
    }
    CUDA_TRY(h, cudaMemcpyAsync(h->d_action, src, sizeof(float) * B * A, cudaMemcpyHostToDevice, h->stream));
  }
  // Large resident batches are stepped in two shares so that the device->host copy of the
  // first overlaps the kernels of the second (the copy of 32768 x 56 floats is 0.15 ms of a
  // 1.06 ms step when it runs after the last kernel).  Measured on B200, 32768 x 64x96,
  // end-to-end ms per step with 1 / 2 / 4 / 8 shares: 1.277 / 1.244 / 1.352 / 1.579 -- every
  // further share costs ~75 us (three more launches, their ramp and tail, cold L2 at the
  // share boundary) and hides less than that.
  const bool pinned[4] = {obs && is_pinned(obs), reward && is_pinned(reward),
                          step_type && is_pinned(step_type), discount && is_pinned(discount)};
  int shares = h->host_shares;
  if (shares <= 0) shares = (h->path == SBX_PATH_RESIDENT && B * (size_t)(h->D + 3) * 4 >= ((size_t)2 << 20)) ? 2 : 1;
  auto after_share = [&](int b0, int b1) -> int {
    cudaEvent_t ev = h->ev_share[h->share_seq++ % SBX_MAX_CHUNKS];
    CUDA_TRY(h, cudaEventRecord(ev, h->stream));
    CUDA_TRY(h, cudaStreamWaitEvent(h->s_copy, ev, 0));
    return d2h_share(h, obs, reward, step_type, discount, pinned, b0, b1, h->s_copy);
  };
  if (int rc = do_step(h, h->d_action, h->d_obs, h->d_reward, h->d_step_type, h->d_discount, h->stream,
                       shares, after_share)) return rc;
  CUDA_TRY(h, cudaStreamSynchronize(h->s_copy));      // the copies follow their share's kernels
  if (obs && !pinned[0]) memcpy(obs, h->h_obs, sizeof(float) * B * h->D);
  if (reward && !pinned[1]) memcpy(reward, h->h_reward, sizeof(float) * B);
  if (step_type && !pinned[2]) memcpy(step_type, h->h_step_type, sizeof(int32_t) * B);
  if (discount && !pinned[3]) memcpy(discount, h->h_discount, sizeof(float) * B);
  return SBX_OK;
}

int sbx_fd_step(sbx_handle h, const double* ambient, const double* convection) {
  if (!h || !ambient || !convection) return fail(h, SBX_E_INVALID, "null argument");
  CUDA_TRY(h, cudaSetDevice(h->device));
  const size_t B = h->cfg.n_envs;
  CUDA_TRY(h, cudaDeviceSynchronize());
  CUDA_TRY(h, cudaMemcpy(h->fd_ambient, ambient, sizeof(double) * B, cudaMemcpyHostToDevice));
  CUDA_TRY(h, cudaMemcpy(h->fd_convection, convection, sizeof(double) * B, cudaMemcpyHostToDevice));
  Params& p = h->P;
  p.fd_only = 1;
  p.time_index = 0;
  p.action = nullptr; p.obs = nullptr; p.reward = nullptr; p.step_type = nullptr; p.discount_out = nullptr;
  int rc;
  p.b_begin = 0; p.b_end = p.B; p.build_hdr = 0;
  if (h->path == SBX_PATH_RESIDENT) {
    rc = h->cfg.solver == SBX_SOLVER_TF_JACOBI ? prepare_plans(h, h->stream) : SBX_OK;
    if (!rc) rc = run_resident(h, h->stream, true);
  } else {
    rc = SBX_OK;
    {                                // the sweeps read the per-building solve header
      const int wpb = 4;
      k_build_header<<<(unsigned)((p.B + wpb - 1) / wpb), wpb * 32, 0, h->stream>>>(p);
      rc = launch_check(h, "k_build_header");
    }
    if (!rc) rc = run_stream_sweeps(h, h->stream);
  }
  p.fd_only = 0;
  if (rc) return rc;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return SBX_OK;
}

int sbx_host_alloc(size_t bytes, void** out) {
  if (!out) return fail(nullptr, SBX_E_INVALID, "null argument");
  cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
  if (e != cudaSuccess) return fail(nullptr, SBX_E_NOMEM, "cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e));
  return SBX_OK;
}

int sbx_host_free(void* p) {
  if (p) cudaFreeHost(p);
  return SBX_OK;
}

int sbx_set_device_convection(sbx_handle h, double p, int32_t distance, uint64_t seed) {
  if (!h) return fail(h, SBX_E_INVALID, "null handle");
  if (!(p >= 0.0) || p > 1.0) return fail(h, SBX_E_INVALID, "convection probability must be in [0, 1]");
  if (p > 0.0 && distance < 1) return fail(h, SBX_E_INVALID, "device-RNG convection needs distance >= 1 (the unbounded shuffle, distance = -1, is replay-only)");
  h->P.conv_p = distance == 0 ? 0.0 : p;       // p == 0 or distance == 0: no convection (stochastic_convection_simulator.py:70-72)
  h->P.conv_distance = distance;
  h->P.conv_seed = seed;
  return SBX_OK;
}

int sbx_set_option(sbx_handle h, int option, int64_t value) {
  if (!h) return fail(h, SBX_E_INVALID, "null handle");
  switch (option) {
    case SBX_OPT_PIPELINE_CHUNKS:
      if (value < 1 || value > SBX_MAX_CHUNKS || value > h->cfg.n_envs)
        return fail(h, SBX_E_INVALID, "SBX_OPT_PIPELINE_CHUNKS must be in [1, min(%d, n_envs)]", SBX_MAX_CHUNKS);
      if (h->path != SBX_PATH_RESIDENT || h->cfg.solver != SBX_SOLVER_TF_JACOBI) value = 1;
      h->n_chunks = (int)value;
      return SBX_OK;
    case SBX_OPT_HOST_SHARES:
      if (value < 0 || value > SBX_MAX_CHUNKS) return fail(h, SBX_E_INVALID, "SBX_OPT_HOST_SHARES must be in [0, %d]", SBX_MAX_CHUNKS);
      h->host_shares = (int)value;
      return SBX_OK;
    case SBX_OPT_L2_PREFETCH_DISTANCE:
      if (value < 0 || value > h->cfg.n_envs) return fail(h, SBX_E_INVALID, "SBX_OPT_L2_PREFETCH_DISTANCE must be in [0, n_envs]");
      h->P.prefetch_dist = (int)value;
      return SBX_OK;
    case SBX_OPT_NUMPY_MEANS:
      if (value != 0 && value != 1) return fail(h, SBX_E_INVALID, "SBX_OPT_NUMPY_MEANS must be 0 or 1");
      if (value && h->cfg.solver != SBX_SOLVER_TF_JACOBI)
        return fail(h, SBX_E_INVALID, "SBX_OPT_NUMPY_MEANS follows the fp32 field of the TF-Jacobi solver only");
      h->P.pw_on = (int)value;
      return SBX_OK;
    default:
      return fail(h, SBX_E_INVALID, "unknown option %d", option);
  }
}

int sbx_timing_begin(sbx_handle h) {
  if (!h) return fail(h, SBX_E_INVALID, "null handle");
  CUDA_TRY(h, cudaSetDevice(h->device));
  CUDA_TRY(h, cudaDeviceSynchronize());
  CUDA_TRY(h, cudaMemcpy(&h->loop_launches_seen, h->P.sweep_launches, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  h->t_step_used = h->t_solve_used = 0;
  h->timing_on = 1;
  return SBX_OK;
}

int sbx_timing_end(sbx_handle h, sbx_timing* out) {
  if (!h || !out) return fail(h, SBX_E_INVALID, "null argument");
  CUDA_TRY(h, cudaSetDevice(h->device));
  CUDA_TRY(h, cudaDeviceSynchronize());
  h->timing_on = 0;
  memset(out, 0, sizeof(*out));
  auto sum = [&](const std::vector<cudaEvent_t>& pool, size_t used, double* ms, int64_t* n) -> int {
    for (size_t i = 0; i + 1 < used; i += 2) {
      float t = 0.f;
      CUDA_TRY(h, cudaEventElapsedTime(&t, pool[i], pool[i + 1]));
      *ms += (double)t;
      ++*n;
    }
    return SBX_OK;
  };
  if (int rc = sum(h->t_step, h->t_step_used, &out->step_ms, &out->n_steps)) return rc;
  if (int rc = sum(h->t_solve, h->t_solve_used, &out->solve_ms, &out->n_solve_launches)) return rc;
  {
    // streaming path with the device-driven loop: one event pair brackets ALL sweeps of a
    // step; the number of k_sweep launches inside comes from the device counter
    unsigned long long loops = 0;
    CUDA_TRY(h, cudaMemcpy(&loops, h->P.sweep_launches, sizeof(loops), cudaMemcpyDeviceToHost));
    if (loops > h->loop_launches_seen) out->n_solve_launches = (int64_t)(loops - h->loop_launches_seen);
  }
  out->n_chunks = h->n_chunks;
  h->t_step_used = h->t_solve_used = 0;
  return SBX_OK;
}

int sbx_sync(sbx_handle h) {
  if (!h) return fail(h, SBX_E_INVALID, "null handle");
  CUDA_TRY(h, cudaSetDevice(h->device));
  CUDA_TRY(h, cudaDeviceSynchronize());
  return SBX_OK;
}

}  // extern "C"
