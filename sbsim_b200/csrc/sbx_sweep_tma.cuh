// k_sweep_tma: the streaming Jacobi sweep with TMA-staged tiles (the default for 4-CV vectors;
// SBX_SWEEP_TMA=0 selects the plain-load k_sweep).
//
// Same arithmetic and the same tile decomposition as k_sweep<4, R, MODE> (sbx_kernels.cuh): a
// CTA owns a band of 128 columns x (8 x rows_per_warp) rows of one building.  Here the band is
// streamed through shared memory in chunks of kTmaChunk rows by tensor-map copies
// (cp.async.bulk.tensor -> SASS UTMALDG): per chunk one [chunk + 2 halo rows, 128 + 8 halo
// columns] box of T_est and one [chunk, 128] box of T_prev, kTmaStages chunks in flight on
// mbarriers, issued by one thread.  Out-of-range rows / columns are zero-filled by the copy and
// replaced by T_inf exactly where k_sweep passes T_inf.  Descriptors (2 bytes per CV, natural pitch:
// their row stride is not a 16-byte multiple on the calibrated plan) and the output keep plain
// 64 / 128-bit accesses.
#pragma once
#include "sbx_kernels.cuh"

namespace sbx {

#ifndef SBX_TMA_CHUNK
#define SBX_TMA_CHUNK 16
#endif
#ifndef SBX_TMA_STAGES
#define SBX_TMA_STAGES 3
#endif
constexpr int kTmaChunk = SBX_TMA_CHUNK;      // rows per chunk (a multiple of 8): kTmaChunk / 8 vectors per thread
constexpr int kTmaStages = SBX_TMA_STAGES;
constexpr int kTmaRowsPerWarp = kTmaChunk / (kStreamThreads / 32);
constexpr int kTmaBoxW = 32 * 4 + 8;          // 4 halo columns on each side keep the box 16-byte aligned
constexpr int kTmaInBytes = ((kTmaChunk + 2) * kTmaBoxW * 4 + 127) & ~127;
constexpr int kTmaPrevBytes = kTmaChunk * 128 * 4;

__host__ __device__ inline size_t sweep_tma_smem(bool first) {
  return (size_t)kTmaStages * (kTmaInBytes + (first ? 0 : kTmaPrevBytes)) + 128;
}

__device__ __forceinline__ void tma_load_box(void* smem_dst, const void* tmap, int x, int y, int z, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
      : "memory");
}

// tm_in[i]: [B, H, W] fp32 of Params::tbuf[i] with box [1, kTmaChunk + 2, kTmaBoxW];
// tm_pv[i]: the same tensors with box [1, kTmaChunk, 128]
struct SweepTensorMaps {
  CUtensorMap in[3];
  CUtensorMap prev[3];
};

template <int MODE>
__global__ void __launch_bounds__(kStreamThreads, SBX_SWEEP_MIN_CTAS) k_sweep_tma(const Params p, const int k_arg,
                                                                                   const __grid_constant__ SweepTensorMaps tm) {
  constexpr int V = 4;
  constexpr bool first = MODE != 0;
  extern __shared__ __align__(128) unsigned char dyn[];
  __shared__ __align__(16) Combo tab[kNumCombos];
  __shared__ float qcv[kMaxZones + 1];
  __shared__ __align__(8) uint64_t full[kTmaStages];
  const int k = first ? 1 : (k_arg > 0 ? k_arg : *p.sweep_k);
  const StreamTiling tl = stream_tiling(p.H, p.W, V);
  const int b = blockIdx.x / tl.tiles;
  if (!p.active[b]) return;
  const int tile = blockIdx.x - b * tl.tiles;
  const int ty = tile / tl.tiles_x, tx = tile - ty * tl.tiles_x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = p.H, W = p.W, Z = p.Z;
  const size_t n_cv = (size_t)H * W;
  const int plan = p.n_plans == 1 ? 0 : b;
  const int cur = p.cur[b];
  int bi, bo;
  sweep_buffers(cur, k, bi, bo);
  const int tile_h = (kStreamThreads / 32) * tl.rows_per_warp;
  const int r_begin = ty * tile_h, r_end = min(r_begin + tile_h, H);
  const int n_chunks = (r_end - r_begin + kTmaChunk - 1) / kTmaChunk;
  const int x0 = tx * 128;
  unsigned char* s_in = dyn;
  unsigned char* s_pv = dyn + (size_t)kTmaStages * kTmaInBytes;
  constexpr uint32_t kStageBytes = (uint32_t)((kTmaChunk + 2) * kTmaBoxW * 4 + (first ? 0 : kTmaPrevBytes));
  auto issue = [&](int chunk) {      // one thread
    const int st = chunk % kTmaStages;
    const int r = r_begin + chunk * kTmaChunk;
    mbar_expect_tx(&full[st], kStageBytes);
    tma_load_box(s_in + (size_t)st * kTmaInBytes, &tm.in[bi], x0 - 4, r - 1, b, &full[st]);
    if constexpr (!first) tma_load_box(s_pv + (size_t)st * kTmaPrevBytes, &tm.prev[cur], x0, r, b, &full[st]);
  };
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kTmaStages; ++s) mbar_init(&full[s], 1);
  }
  const unsigned char* gH = p.hdr + (size_t)b * header_bytes(Z);
  for (int i = tid; i < (int)(sizeof(Combo) * kNumCombos / 16); i += kStreamThreads)
    reinterpret_cast<float4*>(tab)[i] = reinterpret_cast<const float4*>(gH)[i];
  const float* hq = reinterpret_cast<const float*>(gH + sizeof(Combo) * kNumCombos);
  for (int i = tid; i < Z; i += kStreamThreads) qcv[i] = hq[i];
  const float t_inf = hq[header_q_slots(Z)];
  __syncthreads();
  if (tid == 0)
    for (int c = 0; c < kTmaStages && c < n_chunks; ++c) issue(c);
  AreaCoef az;
  az.full = tab[SBX_CV_INTERIOR * kNumMaterials].vz;
  az.half = tab[SBX_CV_CORNER_TL * kNumMaterials].vz;
  const float dt = p.dt, rdt = __frcp_rn(p.dt);
  float* __restrict__ tout = p.tbuf[bo] + (size_t)b * n_cv;
  const uint16_t* __restrict__ dsc = p.desc_spk + (size_t)plan * n_cv;
  const int c0 = x0 + lane * V;
  const bool col_ok = c0 < W;
  const bool edge_l = lane == 0, edge_r = lane == 31 || c0 + V >= W;
  float lmax = 0.f;
  for (int chunk = 0; chunk < n_chunks; ++chunk) {
    const int st = chunk % kTmaStages;
    const float* sin_ = reinterpret_cast<const float*>(s_in + (size_t)st * kTmaInBytes);
    const float* spv = reinterpret_cast<const float*>(s_pv + (size_t)st * kTmaPrevBytes);
    const int rc = r_begin + chunk * kTmaChunk;
    // descriptors of this thread's rows: issued before the wait
    DescRaw<V> dr[kTmaRowsPerWarp] = {};
#pragma unroll
    for (int j = 0; j < kTmaRowsPerWarp; ++j) {
      const int r = rc + kTmaRowsPerWarp * warp + j;
      if (col_ok && r < r_end) dr[j] = load_d_raw<V>(dsc + (size_t)r * W + c0);
    }
    mbar_wait(&full[st], (uint32_t)((chunk / kTmaStages) & 1));
#pragma unroll
    for (int j = 0; j < kTmaRowsPerWarp; ++j) {
      const int lr = kTmaRowsPerWarp * warp + j, r = rc + lr;        // warp-uniform
      if (r >= r_end) break;
      const float* row = sin_ + (lr + 1) * kTmaBoxW + 4 + lane * V;     // this thread's vector in the centre row
      float up[V], c[V], dn[V];
      load_f<V>(row, c);
      if (r > 0) load_f<V>(row - kTmaBoxW, up); else fill<V>(up, t_inf);
      if (r + 1 < H) load_f<V>(row + kTmaBoxW, dn); else fill<V>(dn, t_inf);
      float left = __shfl_up_sync(0xffffffffu, c[V - 1], 1);
      float right = __shfl_down_sync(0xffffffffu, c[0], 1);
      if (col_ok) {
        if (edge_l) left = c0 > 0 ? row[-1] : t_inf;
        if (edge_r) right = c0 + V < W ? row[V] : t_inf;
        float tp[V], o[V];
        uint32_t d[V];
        expand_d<V>(dr[j], d);
        if constexpr (first) {
#pragma unroll
          for (int e = 0; e < V; ++e) tp[e] = c[e];
        } else {
          load_f<V>(spv + lr * 128 + lane * V, tp);
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
        store_f<V>(tout + (size_t)r * W + c0, o);
      }
    }
    __syncthreads();                 // every warp is done with this stage
    if (tid == 0 && chunk + kTmaStages < n_chunks) issue(chunk + kTmaStages);
  }
  lmax = warp_max(lmax);
  if (lane == 0 && lmax > 0.f) atomicMax(&p.max_delta_bits[b], __float_as_uint(lmax));
}

}  // namespace sbx
