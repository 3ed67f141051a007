// Merged backward for sm_100a: the two contractions of the BAGS head's backward in ONE persistent launch.
//
//   units [0, dw_units)            dW[C,K] += rowscale(c) * dz^T[C, rois] x[rois, K]   (split-K over rois, red.add)
//   units [dw_units, +dx_units)    dX[rois,K] = dz[rois, C] W'[C, K]                    (store)
//
// Both read dz in place (MN-major for dW, K-major for dX).  With 120 + 128 units on 148 CTAs most CTAs run a
// dW unit and then a dX unit: the dW epilogue (TMEM accumulator stage 0) overlaps the dX mainloop (stage 1),
// one kernel boundary and one prologue disappear, and all 148 SMs work instead of 120 / 128.
//
// Pipeline roles are those of bags_gemm_kernel (TMA producer warp, single-thread tcgen05.mma issuer, 8 epilogue
// warps); the per-unit operand layout / epilogue is selected at run time by the unit's kind.
// PDL: dW epilogues wait for bwd_prep (zeroed dW, bias-gradient partials) -- their mainloops do not; the
// producer waits before the first W' load of a dX unit.
#pragma once
#include "bags_gemm.cuh"

namespace bags {

struct BwdFusedParams {
  int C, Kf, Nr;                 // logits, features, rois
  int dw_m_tiles, dw_n_tiles, dw_splits, dw_kblocks;   // dW: M = C, N = Kf, contraction over rois
  int dx_m_tiles, dx_n_tiles, dx_kblocks;              // dX: M = rois, N = Kf, contraction over C
  int dw_units, dx_units;
  float* dW; long long lddw;
  void* dX; long long lddx;
  const float* gscale; int G; int gstart[kMaxGroups]; int glen[kMaxGroups];
  const float* colsum_in; int colsum_tiles; float* db;
  // In-kernel preparation (single-CTA kernel): the jobs bwd_prep_kernel would run (zero dW, W' = gout-scaled W,
  // bias-gradient partials) are done by the epilogue warps while the first mainloop runs; `sync` is a library-owned
  // pair of counters {jobs done, CTAs exited} that is zero between launches.  prep_jobs == 0: a separate
  // bwd_prep_kernel ran before this one (programmatic dependent launch).
  BwdPrepParams prep; int prep_jobs; unsigned int* sync;
  int dx_uniform_ok;   // gout given and tmap_w valid: dX units may read W and scale by gout[0] when gout is uniform
  int dw_store;        // split-K factor 1: a dW unit owns its tile -- plain stores, no zeroing, no red.add (the trunk's wide layers)
  int wait_dz;         // the producer waits (griddepcontrol.wait) before its first load: dz comes from the preceding kernel
  int dw_needs_prep;   // dW epilogues wait for preparation (zeroed dW / column-sum partials); 0: dW was zeroed by the
                       // forward's clear hook and the column sums come from the forward as well
  long long* timing;
};

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// wait until the preparation jobs of the grid are done.  With ticketed jobs (the default) the CTAs that are running finish
// all of them, so this wait cannot depend on CTAs that are not resident yet; the time bound (10 s, checked every 4096
// probes) only keeps a broken launch from hanging the device for ever.
__device__ __forceinline__ void wait_grid_jobs(const unsigned int* ctr, unsigned int target) {
  unsigned int spins = 0;
  long long t0 = 0;
  while (ld_acquire_gpu(ctr) < target) {
    __nanosleep(64);
    if ((++spins & 4095u) == 0u) {
      const long long now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 10000000000LL) { printf("bags: grid job counter timed out (block %d)\n", (int)blockIdx.x); __trap(); }
    }
  }
}

// MT = number of 128-row accumulator sub-tiles of one unit.  MT = 1: 128 x 256 units, two TMEM accumulator stages
// (a unit's epilogue overlaps the next unit's mainloop), 4 pipeline stages of 48 KB.  MT = 2: 256 x 256 units -- the
// B tile is shared by two MMAs, so a unit ingests 64 KB instead of 96 KB per 2 x (128 x 256 x 64) of work and the
// mainloop is no longer bound by the SM's 64 B/clk L2 port; all 512 TMEM columns hold one unit (no epilogue
// overlap), 3 pipeline stages of 64 KB.  Pays off when the problem is about one wave of 256 x 256 units.
template <bool TF32, int MT = 1>
struct BwdCfg {
  static_assert(MT == 1 || MT == 2, "one or two accumulator sub-tiles");
  static constexpr int SUB_M = 128;
  static constexpr int BLOCK_M = SUB_M * MT, BLOCK_N = 256, STAGES = (MT == 1) ? 4 : 3, ACC_STAGES = 2 / MT;
  static constexpr int ELT = TF32 ? 4 : 2;
  static constexpr int BLOCK_K = 128 / ELT, UMMA_K = 32 / ELT, K_STEPS = BLOCK_K / UMMA_K;
  static constexpr int SLAB = 128 / ELT;
  static constexpr int A_BYTES = BLOCK_M * 128, B_BYTES = BLOCK_N * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_WARPS = 8, EPI_BUF_BYTES = 32 * 128, EPI_STAGING_BYTES = EPI_WARPS * EPI_BUF_BYTES;
  static constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGING_BYTES + 1024 + 256;
  static_assert(SMEM_BYTES <= 232448, "exceeds shared memory");
};

// TICKET = true (default): preparation jobs are handed out through an atomic ticket, so whichever CTAs are resident
// finish them all -- no co-residency requirement (MPS / green-context SM limits, concurrent kernels on other streams).
// TICKET = false (BAGS_BWD_TICKET=0, opt-in): job j runs on CTA j % gridDim.x; only correct when all CTAs of the grid
// are resident together.
template <bool TF32, int MT, bool TICKET = false>
__global__ void __launch_bounds__(64 + 32 * 8, 1)
bags_bwd_fused_kernel(const __grid_constant__ CUtensorMap tmap_dzT,   // dz as MN-major A of dW  (box SLAB x BLOCK_K)
                      const __grid_constant__ CUtensorMap tmap_xT,    // x  as MN-major B of dW  (box SLAB x BLOCK_K)
                      const __grid_constant__ CUtensorMap tmap_dz,    // dz as K-major  A of dX  (box BLOCK_K x 128)
                      const __grid_constant__ CUtensorMap tmap_wT,    // W' as MN-major B of dX  (box SLAB x BLOCK_K)
                      const __grid_constant__ CUtensorMap tmap_w,     // W  (unscaled), same geometry: used instead of
                                                                      // W' when all gout[g] are equal
                      const BwdFusedParams p) {
  using Cfg = BwdCfg<TF32, MT>;
  constexpr int BLOCK_M = Cfg::BLOCK_M, BLOCK_N = Cfg::BLOCK_N, BLOCK_K = Cfg::BLOCK_K, STAGES = Cfg::STAGES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1 KB aligned, still a shared-space pointer
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  uint8_t* smem_epi = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + Cfg::EPI_STAGING_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + Cfg::ACC_STAGES;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * Cfg::ACC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { stamp(p.timing, 0); if (p.timing) p.timing[blockIdx.x * 8 + 7] = sm_id(); }
  pdl_trigger();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_dzT); tma_prefetch_desc(&tmap_xT); tma_prefetch_desc(&tmap_dz); tma_prefetch_desc(&tmap_wT);
    tma_prefetch_desc(&tmap_w);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
#pragma unroll
    for (int a = 0; a < Cfg::ACC_STAGES; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], Cfg::EPI_WARPS); }
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_holder, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  if (threadIdx.x == 0) stamp(p.timing, 1);

  const int num_units = p.dw_units + p.dx_units;
  // unit -> (kind, m0, n0, kb0, kb1, n_tile, split)
  struct Unit { bool is_dw; int m_tile, n_tile, split, kb0, kb1; };
  auto decode = [&](int u) -> Unit {
    Unit r;
    if (u < p.dw_units) {
      r.is_dw = true;
      const int per = p.dw_m_tiles * p.dw_n_tiles;
      r.split = u / per;
      const int t = u - r.split * per;
      r.m_tile = t / p.dw_n_tiles;
      r.n_tile = t - r.m_tile * p.dw_n_tiles;
      const int base = p.dw_kblocks / p.dw_splits, rem = p.dw_kblocks % p.dw_splits;
      r.kb0 = r.split * base + (r.split < rem ? r.split : rem);
      r.kb1 = r.kb0 + base + (r.split < rem ? 1 : 0);
    } else {
      r.is_dw = false;
      const int t = u - p.dw_units;
      r.split = 0;
      r.m_tile = t / p.dx_n_tiles;
      r.n_tile = t - r.m_tile * p.dx_n_tiles;
      r.kb0 = 0;
      r.kb1 = p.dx_kblocks;
    }
    return r;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool waited = false;
      const bool inkernel = p.prep_jobs > 0;
      if (p.wait_dz) pdl_wait();   // dz comes from the forward kernel (no preparation kernel in between)
      // equal per-bin upstream gradients (the usual case: every bin's loss has weight 1): dX = g0 * dz W, so the
      // scaled copy W' is neither made nor waited for
      float g0 = 1.f;
      const bool plain_w = p.dx_uniform_ok && gout_uniform(p.gscale, p.G, g0);
      for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
        const Unit un = decode(u);
        const int m0 = un.m_tile * BLOCK_M, n0 = un.n_tile * BLOCK_N;
        if (!un.is_dw && !waited && !plain_w) {   // W' is written by bwd_prep / by the preparation jobs of all CTAs
          if (inkernel) { wait_grid_jobs(p.sync, TICKET ? static_cast<unsigned int>(p.prep_jobs) : gridDim.x); asm volatile("fence.proxy.async;" ::: "memory"); }
          else pdl_wait();
          waited = true;
        }
        for (int kb = un.kb0; kb < un.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          const int k0 = kb * BLOCK_K;
          uint8_t* sa = smem_a + stage * Cfg::A_BYTES;
          uint8_t* sb = smem_b + stage * Cfg::B_BYTES;
          if (un.is_dw) {
#pragma unroll
            for (int i = 0; i < BLOCK_M / Cfg::SLAB; ++i)
              tma_load_2d(sa + i * (BLOCK_K * 128), &tmap_dzT, &full_bar[stage], m0 + i * Cfg::SLAB, k0);
#pragma unroll
            for (int i = 0; i < BLOCK_N / Cfg::SLAB; ++i)
              tma_load_2d(sb + i * (BLOCK_K * 128), &tmap_xT, &full_bar[stage], n0 + i * Cfg::SLAB, k0);
          } else {
            tma_load_2d(sa, &tmap_dz, &full_bar[stage], k0, m0);
            const CUtensorMap* tw = plain_w ? &tmap_w : &tmap_wT;
#pragma unroll
            for (int i = 0; i < BLOCK_N / Cfg::SLAB; ++i)
              tma_load_2d(sb + i * (BLOCK_K * 128), tw, &full_bar[stage], n0 + i * Cfg::SLAB, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== UMMA issuer (single thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc_dw = make_instr_desc(TF32 ? 2u : 1u, true, true, Cfg::SUB_M, BLOCK_N);
      constexpr uint32_t idesc_dx = make_instr_desc(TF32 ? 2u : 1u, false, true, Cfg::SUB_M, BLOCK_N);
      constexpr uint32_t SUB_A_BYTES = Cfg::SUB_M * 128;   // one 128-row sub-tile of A, in either operand layout
      constexpr uint64_t MN_LAYOUT = TF32 ? kSwizzle128B_Base32B : kSwizzle128B;
      constexpr uint32_t MN_LBO = BLOCK_K * 128, MN_SBO = TF32 ? 512 : 1024, MN_KSTEP = Cfg::UMMA_K * 128;
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int u = blockIdx.x; u < num_units; u += gridDim.x, ++local) {
        const Unit un = decode(u);
        const int acc = local % Cfg::ACC_STAGES;
        const uint32_t acc_phase = (local / Cfg::ACC_STAGES) & 1u;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * MT * BLOCK_N;
        for (int kb = un.kb0; kb < un.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (local == 0 && kb == un.kb0) stamp(p.timing, 2);
          const uint32_t sa = smem_u32(smem_a + stage * Cfg::A_BYTES);
          const uint32_t sb = smem_u32(smem_b + stage * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < Cfg::K_STEPS; ++k) {
            const uint64_t bdesc = make_smem_desc(sb + k * MN_KSTEP, MN_LBO, MN_SBO, MN_LAYOUT);
            const uint32_t accum = (kb > un.kb0 || k > 0) ? 1u : 0u;
#pragma unroll
            for (int h = 0; h < MT; ++h) {   // the sub-tiles share the B descriptor
              const uint32_t sah = sa + h * SUB_A_BYTES;
#ifdef BAGS_X_BWD_KMAJOR   // timing experiment only (wrong results): every operand read as a K-major tile
              {
                constexpr uint32_t idesc_kk = make_instr_desc(TF32 ? 2u : 1u, false, false, Cfg::SUB_M, BLOCK_N);
                const uint64_t ad = make_smem_desc(sah + k * 32, 16, 1024, kSwizzle128B);
                const uint64_t bd = make_smem_desc(sb + k * 32, 16, 1024, kSwizzle128B);
                if (TF32) umma_tf32(d_tmem + h * BLOCK_N, ad, bd, idesc_kk, accum);
                else      umma_bf16(d_tmem + h * BLOCK_N, ad, bd, idesc_kk, accum);
                continue;
              }
#endif
              if (un.is_dw) {
                const uint64_t adesc = make_smem_desc(sah + k * MN_KSTEP, MN_LBO, MN_SBO, MN_LAYOUT);
                if (TF32) umma_tf32(d_tmem + h * BLOCK_N, adesc, bdesc, idesc_dw, accum);
                else      umma_bf16(d_tmem + h * BLOCK_N, adesc, bdesc, idesc_dw, accum);
              } else {
                const uint64_t adesc = make_smem_desc(sah + k * 32, 16, 1024, kSwizzle128B);
                if (TF32) umma_tf32(d_tmem + h * BLOCK_N, adesc, bdesc, idesc_dx, accum);
                else      umma_bf16(d_tmem + h * BLOCK_N, adesc, bdesc, idesc_dx, accum);
              }
            }
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tfull_bar[acc]);
        if (local == 0) stamp(p.timing, 3);
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    uint8_t* buf = smem_epi + (warp - 2) * Cfg::EPI_BUF_BYTES;
    constexpr int HALF_N = BLOCK_N / 2;
    bool waited = false;
    int local = 0;
    const bool inkernel = p.prep_jobs > 0;
    float dx_scale = 1.f;
    if (p.dx_uniform_ok) {
      float g0;
      if (gout_uniform(p.gscale, p.G, g0)) dx_scale = g0;
    }
    if (inkernel) {
      // preparation jobs, spread over the grid; they run under the first mainloop (these warps are idle until then)
      const int tid = static_cast<int>(threadIdx.x) - 64;
      float (*s_part)[64] = reinterpret_cast<float (*)[64]>(smem_epi);
      pdl_wait();   // dW / W' may still be in use by whatever ran before the forward kernel; dz comes from it
      if constexpr (TICKET) {
        // Jobs are claimed through an atomic ticket by whichever CTAs are running, so the grid counter reaches its
        // target (= the number of jobs) without all CTAs having to be resident together: the CTAs that ARE resident
        // claim and finish every job before any of them waits.  One fence + one counted publication per CTA.
        __shared__ int s_job;
        unsigned int done = 0;
        for (;;) {
          if (tid == 0) s_job = static_cast<int>(atomicAdd(p.sync + 2, 1u));
          asm volatile("bar.sync 5, 256;" ::: "memory");
          const int j = s_job;
          if (j >= p.prep_jobs) break;
          bwd_prep_job<TF32, 5>(p.prep, j, tid, s_part);
          asm volatile("bar.sync 5, 256;" ::: "memory");     // s_job / s_part may be reused
          ++done;
        }
        if (done > 0) {
          asm volatile("fence.proxy.async;" ::: "memory");   // W' is read through TMA (async proxy) by other CTAs
          __threadfence();
          asm volatile("bar.sync 5, 256;" ::: "memory");
          if (tid == 0) atomicAdd(p.sync, done);
        }
      } else {
      for (int j = blockIdx.x; j < p.prep_jobs; j += gridDim.x) {
        bwd_prep_job<TF32, 5>(p.prep, j, tid, s_part);
        asm volatile("bar.sync 5, 256;" ::: "memory");   // s_part is reused by the next job
      }
      asm volatile("fence.proxy.async;" ::: "memory");   // W' is read through TMA (async proxy) by other CTAs
      __threadfence();
      asm volatile("bar.sync 5, 256;" ::: "memory");
      if (tid == 0) atomicAdd(p.sync, 1u);
      }
    }
    for (int u = blockIdx.x; u < num_units; u += gridDim.x, ++local) {
      const Unit un = decode(u);
      const int n0 = un.n_tile * BLOCK_N;
      const int acc = local % Cfg::ACC_STAGES;
      const uint32_t acc_phase = (local / Cfg::ACC_STAGES) & 1u;
      const int Mrows = un.is_dw ? p.C : p.Nr;
#pragma unroll 1
      for (int h = 0; h < MT; ++h) {
      const int m_warp = un.m_tile * BLOCK_M + h * Cfg::SUB_M + quarter * 32;
      const int m = m_warp + lane;

      float scale = dx_scale;
      if (un.is_dw) {
        if (!waited) {   // dW was zeroed / column-sum partials were made by bwd_prep or by every CTA's jobs
          if (inkernel && p.dw_needs_prep) { if (lane == 0) wait_grid_jobs(p.sync, TICKET ? static_cast<unsigned int>(p.prep_jobs) : gridDim.x); __syncwarp(); __threadfence(); }
          else pdl_wait();   // (the zeroed dW / the forward's column sums come from the preceding kernel)
          waited = true;
        }
        scale = 0.f;
        if (m < p.C) {
          if (p.gscale == nullptr) scale = 1.0f;
          else {
#pragma unroll
            for (int g = 0; g < kMaxGroups; ++g)
              if (g < p.G && m >= p.gstart[g] && m < p.gstart[g] + p.glen[g]) scale = __ldg(p.gscale + g);
          }
          if (p.db != nullptr && un.n_tile == 0 && un.split == 0 && half == 0) {
            float cs = 0.f;
            for (int tt = 0; tt < p.colsum_tiles; ++tt) cs += __ldcg(p.colsum_in + static_cast<long long>(tt) * p.C + m);
            p.db[m] = scale * cs;
          }
        }
      }

      if (h == 0) {
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        if (local == 0 && warp == 2 && lane == 0) stamp(p.timing, 4);
      }
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + (acc * MT + h) * BLOCK_N;
      const bool out_bf16 = (!un.is_dw) && !TF32;
      const int epi_cols = out_bf16 ? 64 : 32;
#pragma unroll 1
      for (int c = half * HALF_N; c < (half + 1) * HALF_N; c += epi_cols) {
        const int n = n0 + c;
        if (n >= p.Kf) break;
        uint32_t v[32];
        uint32_t v2[32];
        tmem_ld_32x32b_x32(t_row + c, v);
        if (out_bf16) tmem_ld_32x32b_x32(t_row + c + 32, v2);
        tmem_ld_wait();
        uint4* rowp = reinterpret_cast<uint4*>(buf + lane * 128);
        if (out_bf16) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t* src = (j < 4) ? (v + 8 * j) : (v2 + 8 * (j - 4));
            uint4 r;
            r.x = pack_bf16x2(__uint_as_float(src[0]) * scale, __uint_as_float(src[1]) * scale);
            r.y = pack_bf16x2(__uint_as_float(src[2]) * scale, __uint_as_float(src[3]) * scale);
            r.z = pack_bf16x2(__uint_as_float(src[4]) * scale, __uint_as_float(src[5]) * scale);
            r.w = pack_bf16x2(__uint_as_float(src[6]) * scale, __uint_as_float(src[7]) * scale);
            rowp[j ^ (lane & 7)] = r;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 r;
            r.x = __uint_as_float(v[4 * j + 0]) * scale; r.y = __uint_as_float(v[4 * j + 1]) * scale;
            r.z = __uint_as_float(v[4 * j + 2]) * scale; r.w = __uint_as_float(v[4 * j + 3]) * scale;
            rowp[j ^ (lane & 7)] = *reinterpret_cast<uint4*>(&r);
          }
        }
        __syncwarp();
        const int vec = out_bf16 ? 8 : 4;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 4 + (lane >> 3);
          const int ch = lane & 7;
          const uint4 val = *reinterpret_cast<const uint4*>(buf + r * 128 + ((ch ^ (r & 7)) << 4));
          const int gm = m_warp + r;
          const int gn = n + ch * vec;
          if (gm < Mrows && gn + vec <= p.Kf) {
            if (un.is_dw) {
              if (p.dw_store)
                *reinterpret_cast<uint4*>(p.dW + static_cast<long long>(gm) * p.lddw + gn) = val;
              else
                red_add_v4_f32(p.dW + static_cast<long long>(gm) * p.lddw + gn, __uint_as_float(val.x), __uint_as_float(val.y),
                               __uint_as_float(val.z), __uint_as_float(val.w));
            } else if (out_bf16) {
              *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.dX) + static_cast<long long>(gm) * p.lddx + gn) = val;
            } else {
              *reinterpret_cast<uint4*>(reinterpret_cast<float*>(p.dX) + static_cast<long long>(gm) * p.lddx + gn) = val;
            }
          }
        }
        __syncwarp();
      }
      }   // sub-tiles
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (local == 0 && warp == 2 && lane == 0) stamp(p.timing, 5);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  if (threadIdx.x == 0 && p.prep_jobs > 0) {
    // the last CTA to get here re-arms the counters for the next launch (nobody is waiting on them any more)
    if (atomicAdd(p.sync + 1, 1u) == gridDim.x - 1) { p.sync[0] = 0u; p.sync[1] = 0u; if (TICKET) p.sync[2] = 0u; __threadfence(); }
  }
  if (threadIdx.x == 0) stamp(p.timing, 6);
}


// ======================================================================================================
// CTA-pair variant (tcgen05 cta_group::2)
// ======================================================================================================
template <bool TF32>
struct BwdPairCfg {
  // CTA pair (cta_group::2): the MMA unit is 256 x 256; each CTA stages its own 128 A rows and HALF of the B
  // tile, so a stage is 32 KB (6 stages) and every SM ingests a third less operand data per FLOP.
#ifndef BAGS_PAIR_STAGES
#define BAGS_PAIR_STAGES 6
#endif
  static constexpr int BLOCK_M = 128, BLOCK_N = 256, STAGES = BAGS_PAIR_STAGES, ACC_STAGES = 2;
  static constexpr int HALF_B = BLOCK_N / 2;
  static constexpr int ELT = TF32 ? 4 : 2;
  static constexpr int BLOCK_K = 128 / ELT, UMMA_K = 32 / ELT, K_STEPS = BLOCK_K / UMMA_K;
  static constexpr int SLAB = 128 / ELT;
  static constexpr int A_BYTES = BLOCK_M * 128, B_BYTES = (BLOCK_N / 2) * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_WARPS = 8, EPI_BUF_BYTES = 32 * 128, EPI_STAGING_BYTES = EPI_WARPS * EPI_BUF_BYTES;
  static constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGING_BYTES + 1024 + 256;
  static_assert(SMEM_BYTES <= 232448, "exceeds shared memory");
};

template <bool TF32>
__global__ void __launch_bounds__(64 + 32 * 8, 1)
bags_bwd_pair_kernel(const __grid_constant__ CUtensorMap tmap_dzT,   // dz as MN-major A of dW  (box SLAB x BLOCK_K)
                      const __grid_constant__ CUtensorMap tmap_xT,    // x  as MN-major B of dW  (box SLAB x BLOCK_K)
                      const __grid_constant__ CUtensorMap tmap_dz,    // dz as K-major  A of dX  (box BLOCK_K x 128)
                      const __grid_constant__ CUtensorMap tmap_wT,    // W' as MN-major B of dX  (box SLAB x BLOCK_K)
                      const BwdFusedParams p) {
  using Cfg = BwdPairCfg<TF32>;
  constexpr int BLOCK_M = Cfg::BLOCK_M, BLOCK_N = Cfg::BLOCK_N, BLOCK_K = Cfg::BLOCK_K, STAGES = Cfg::STAGES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1 KB aligned, still a shared-space pointer
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  uint8_t* smem_epi = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + Cfg::EPI_STAGING_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + Cfg::ACC_STAGES;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * Cfg::ACC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = static_cast<int>(blockIdx.x & 1);   // position in the CTA pair (== %cluster_ctarank for 2x1x1 clusters)
  const bool leader = (rank == 0);
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  if (threadIdx.x == 0) { stamp(p.timing, 0); if (p.timing) p.timing[blockIdx.x * 8 + 7] = sm_id(); }
  pdl_trigger();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_dzT); tma_prefetch_desc(&tmap_xT); tma_prefetch_desc(&tmap_dz); tma_prefetch_desc(&tmap_wT);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
#pragma unroll
    for (int a = 0; a < Cfg::ACC_STAGES; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 2 * Cfg::EPI_WARPS); }
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc_2sm(tmem_holder, 512); tmem_relinquish_2sm(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");   // peer barriers are live
  const uint32_t tmem_base = *tmem_holder;
  if (threadIdx.x == 0) stamp(p.timing, 1);

  const int num_units = p.dw_units + p.dx_units;
  // unit -> (kind, m0, n0, kb0, kb1, n_tile, split)
  struct Unit { bool is_dw; int m_tile, n_tile, split, kb0, kb1; };
  auto decode = [&](int u) -> Unit {
    Unit r;
    if (u < p.dw_units) {
      r.is_dw = true;
      const int per = p.dw_m_tiles * p.dw_n_tiles;
      r.split = u / per;
      const int t = u - r.split * per;
      r.m_tile = t / p.dw_n_tiles;
      r.n_tile = t - r.m_tile * p.dw_n_tiles;
      const int base = p.dw_kblocks / p.dw_splits, rem = p.dw_kblocks % p.dw_splits;
      r.kb0 = r.split * base + (r.split < rem ? r.split : rem);
      r.kb1 = r.kb0 + base + (r.split < rem ? 1 : 0);
    } else {
      r.is_dw = false;
      const int t = u - p.dw_units;
      r.split = 0;
      r.m_tile = t / p.dx_n_tiles;
      r.n_tile = t - r.m_tile * p.dx_n_tiles;
      r.kb0 = 0;
      r.kb1 = p.dx_kblocks;
    }
    return r;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool waited = false;
      for (int u = pair_id; u < num_units; u += num_pairs) {
        const Unit un = decode(u);
        // this CTA's 128 rows of the 256-row unit, and its half of the 256 B columns
        const int m0 = (un.m_tile * 2 + rank) * BLOCK_M, n0 = un.n_tile * BLOCK_N + rank * Cfg::HALF_B;
        if (!un.is_dw && !waited) { pdl_wait(); waited = true; }   // W' is written by bwd_prep
        for (int kb = un.kb0; kb < un.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          // all four loads of the pair are counted on the LEADER's barrier
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
          const int k0 = kb * BLOCK_K;
          uint8_t* sa = smem_a + stage * Cfg::A_BYTES;
          uint8_t* sb = smem_b + stage * Cfg::B_BYTES;
          if (un.is_dw) {
#pragma unroll
            for (int i = 0; i < BLOCK_M / Cfg::SLAB; ++i)
              tma_load_2d_2sm(sa + i * (BLOCK_K * 128), &tmap_dzT, &full_bar[stage], m0 + i * Cfg::SLAB, k0);
#pragma unroll
            for (int i = 0; i < Cfg::HALF_B / Cfg::SLAB; ++i)
              tma_load_2d_2sm(sb + i * (BLOCK_K * 128), &tmap_xT, &full_bar[stage], n0 + i * Cfg::SLAB, k0);
          } else {
            tma_load_2d_2sm(sa, &tmap_dz, &full_bar[stage], k0, m0);
#pragma unroll
            for (int i = 0; i < Cfg::HALF_B / Cfg::SLAB; ++i)
              tma_load_2d_2sm(sb + i * (BLOCK_K * 128), &tmap_wT, &full_bar[stage], n0 + i * Cfg::SLAB, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== UMMA issuer (single thread) =====================
    if (lane == 0 && leader) {
      constexpr uint32_t idesc_dw = make_instr_desc(TF32 ? 2u : 1u, true, true, 2 * BLOCK_M, BLOCK_N);
      constexpr uint32_t idesc_dx = make_instr_desc(TF32 ? 2u : 1u, false, true, 2 * BLOCK_M, BLOCK_N);
      constexpr uint64_t MN_LAYOUT = TF32 ? kSwizzle128B_Base32B : kSwizzle128B;
      constexpr uint32_t MN_LBO = BLOCK_K * 128, MN_SBO = TF32 ? 512 : 1024, MN_KSTEP = Cfg::UMMA_K * 128;
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int u = pair_id; u < num_units; u += num_pairs, ++local) {
        const Unit un = decode(u);
        const int acc = local % Cfg::ACC_STAGES;
        const uint32_t acc_phase = (local / Cfg::ACC_STAGES) & 1u;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = un.kb0; kb < un.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (local == 0 && kb == un.kb0) stamp(p.timing, 2);
          const uint32_t sa = smem_u32(smem_a + stage * Cfg::A_BYTES);
          const uint32_t sb = smem_u32(smem_b + stage * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < Cfg::K_STEPS; ++k) {
            const uint64_t bdesc = make_smem_desc(sb + k * MN_KSTEP, MN_LBO, MN_SBO, MN_LAYOUT);
            const uint32_t accum = (kb > un.kb0 || k > 0) ? 1u : 0u;
#ifdef BAGS_X_PAIR_KMAJOR   // timing experiment only (wrong results): both operands read as K-major tiles
            {
              constexpr uint32_t idesc_kk = make_instr_desc(TF32 ? 2u : 1u, false, false, 2 * BLOCK_M, BLOCK_N);
              const uint64_t ad = make_smem_desc(sa + k * 32, 16, 1024, kSwizzle128B);
              const uint64_t bd = make_smem_desc(sb + k * 32, 16, 1024, kSwizzle128B);
              if (TF32) umma_tf32_2sm(d_tmem, ad, bd, idesc_kk, accum);
              else      umma_bf16_2sm(d_tmem, ad, bd, idesc_kk, accum);
              continue;
            }
#endif
            if (un.is_dw) {
              const uint64_t adesc = make_smem_desc(sa + k * MN_KSTEP, MN_LBO, MN_SBO, MN_LAYOUT);
              if (TF32) umma_tf32_2sm(d_tmem, adesc, bdesc, idesc_dw, accum);
              else      umma_bf16_2sm(d_tmem, adesc, bdesc, idesc_dw, accum);
            } else {
              const uint64_t adesc = make_smem_desc(sa + k * 32, 16, 1024, kSwizzle128B);
              if (TF32) umma_tf32_2sm(d_tmem, adesc, bdesc, idesc_dx, accum);
              else      umma_bf16_2sm(d_tmem, adesc, bdesc, idesc_dx, accum);
            }
          }
          umma_commit_2sm_mc(&empty_bar[stage], 0x3);   // frees the slot in both CTAs
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit_2sm_mc(&tfull_bar[acc], 0x3);       // both CTAs' epilogues may read their 128 rows
        if (local == 0) stamp(p.timing, 3);
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    uint8_t* buf = smem_epi + (warp - 2) * Cfg::EPI_BUF_BYTES;
    constexpr int HALF_N = BLOCK_N / 2;
    bool waited = false;
    int local = 0;
    for (int u = pair_id; u < num_units; u += num_pairs, ++local) {
      const Unit un = decode(u);
      const int m_warp = (un.m_tile * 2 + rank) * BLOCK_M + quarter * 32;
      const int m = m_warp + lane;
      const int n0 = un.n_tile * BLOCK_N;
      const int acc = local % Cfg::ACC_STAGES;
      const uint32_t acc_phase = (local / Cfg::ACC_STAGES) & 1u;
      const int Mrows = un.is_dw ? p.C : p.Nr;

      float scale = 1.0f;
      if (un.is_dw) {
        if (!waited) { pdl_wait(); waited = true; }   // dW was zeroed / column-sum partials were made by bwd_prep
        scale = 0.f;
        if (m < p.C) {
          if (p.gscale == nullptr) scale = 1.0f;
          else {
#pragma unroll
            for (int g = 0; g < kMaxGroups; ++g)
              if (g < p.G && m >= p.gstart[g] && m < p.gstart[g] + p.glen[g]) scale = __ldg(p.gscale + g);
          }
          if (p.db != nullptr && un.n_tile == 0 && un.split == 0 && half == 0) {
            float cs = 0.f;
            for (int tt = 0; tt < p.colsum_tiles; ++tt) cs += __ldg(p.colsum_in + static_cast<long long>(tt) * p.C + m);
            p.db[m] = scale * cs;
          }
        }
      }

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if (local == 0 && warp == 2 && lane == 0) stamp(p.timing, 4);
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N;
      const bool out_bf16 = (!un.is_dw) && !TF32;
      const int epi_cols = out_bf16 ? 64 : 32;
#pragma unroll 1
      for (int c = half * HALF_N; c < (half + 1) * HALF_N; c += epi_cols) {
        const int n = n0 + c;
        if (n >= p.Kf) break;
        uint32_t v[32];
        uint32_t v2[32];
        tmem_ld_32x32b_x32(t_row + c, v);
        if (out_bf16) tmem_ld_32x32b_x32(t_row + c + 32, v2);
        tmem_ld_wait();
        uint4* rowp = reinterpret_cast<uint4*>(buf + lane * 128);
        if (out_bf16) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t* src = (j < 4) ? (v + 8 * j) : (v2 + 8 * (j - 4));
            uint4 r;
            r.x = pack_bf16x2(__uint_as_float(src[0]), __uint_as_float(src[1]));
            r.y = pack_bf16x2(__uint_as_float(src[2]), __uint_as_float(src[3]));
            r.z = pack_bf16x2(__uint_as_float(src[4]), __uint_as_float(src[5]));
            r.w = pack_bf16x2(__uint_as_float(src[6]), __uint_as_float(src[7]));
            rowp[j ^ (lane & 7)] = r;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 r;
            r.x = __uint_as_float(v[4 * j + 0]) * scale; r.y = __uint_as_float(v[4 * j + 1]) * scale;
            r.z = __uint_as_float(v[4 * j + 2]) * scale; r.w = __uint_as_float(v[4 * j + 3]) * scale;
            rowp[j ^ (lane & 7)] = *reinterpret_cast<uint4*>(&r);
          }
        }
        __syncwarp();
        const int vec = out_bf16 ? 8 : 4;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 4 + (lane >> 3);
          const int ch = lane & 7;
          const uint4 val = *reinterpret_cast<const uint4*>(buf + r * 128 + ((ch ^ (r & 7)) << 4));
          const int gm = m_warp + r;
          const int gn = n + ch * vec;
          if (gm < Mrows && gn + vec <= p.Kf) {
            if (un.is_dw) {
              red_add_v4_f32(p.dW + static_cast<long long>(gm) * p.lddw + gn, __uint_as_float(val.x), __uint_as_float(val.y),
                             __uint_as_float(val.z), __uint_as_float(val.w));
            } else if (out_bf16) {
              *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.dX) + static_cast<long long>(gm) * p.lddx + gn) = val;
            } else {
              *reinterpret_cast<uint4*>(reinterpret_cast<float*>(p.dX) + static_cast<long long>(gm) * p.lddx + gn) = val;
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tempty_bar[acc]);
      if (local == 0 && warp == 2 && lane == 0) stamp(p.timing, 5);
    }
  }

  tc_fence_before();
  __syncthreads();
  // the peer may still be reading this CTA's shared memory / signalling its barriers
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
  if (threadIdx.x == 0) stamp(p.timing, 6);
}

}  // namespace bags
