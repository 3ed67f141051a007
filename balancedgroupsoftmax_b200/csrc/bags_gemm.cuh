// Persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[M,N] (+)= A[M,K] * B[N,K]^T       fp32 accumulation in TMEM
//
// A and B are each either "K-major" (contraction index contiguous in HBM, i.e.
// a row-major [rows,K] matrix) or "MN-major" (row-major [K,rows]: contraction
// index strided).  That covers the three contractions of the BAGS head without
// any transposed copies in HBM:
//
//   forward  z  = X  W^T  : A = X   [N_roi,K]  K-major , B = W   [C,K]      K-major
//   backward dX = dz W    : A = dz  [N_roi,C]  K-major , B = W   [C,K]      MN-major
//   backward dW = dz^T X  : A = dz  [N_roi,C]  MN-major, B = X   [N_roi,K]  MN-major
//
// Pipeline (one CTA per SM, 320 threads):
//   warp 0      : TMA producer   (cp.async.bulk.tensor, 128B swizzle, mbarrier tx)
//   warp 1      : UMMA issuer    (one thread; tcgen05.mma cta_group::1, M=128)
//   warps 2..9  : epilogue       (tcgen05.ld 32x32b -> registers -> swizzled smem transpose -> coalesced
//                                 16-byte st.global / red.global)
// smem ring of kStages {A tile, B tile}; TMEM holds kAccStages accumulators so
// the epilogue of tile i overlaps the mainloop of tile i+1.
#pragma once
#include "bags_ptx.cuh"

namespace bags {

enum EpiMode : int {
  EPI_STORE_F32 = 0,   // out fp32 = act(acc + bias[n])          act = ReLU when GemmParams::relu, else identity
  EPI_STORE_BF16 = 1,  // out bf16 = act(acc + bias[n])
  EPI_RED_F32 = 2,     // out fp32 += rowscale(m) * acc   (split-K via red.global)
};

constexpr int kMaxGroups = 8;

struct GemmParams {
  int M, N, K;         // logical problem
  int num_m_tiles, num_n_tiles, num_splits;
  int kblocks_total;   // ceil(K / BLOCK_K)
  void* out;           // [M, ldo]
  long long ldo;       // elements
  const float* bias;   // [N] or nullptr                    (EPI_STORE_F32 / EPI_STORE_BF16)
  int relu;            // 1: max(., 0) after the bias         (EPI_STORE_F32 / EPI_STORE_BF16; the shared-FC trunk,
                       //    convfc_bbox_head.py:138-143)
  // EPI_RED_F32: per-row scale = gscale[group(m)], groups = [gstart, gstart+glen)
  const float* gscale; // device [G] or nullptr (=> 1.0)
  int G;
  int gstart[kMaxGroups];
  int glen[kMaxGroups];
  const float* colsum_in;  // optional [colsum_tiles, M]: db_out[m] = rowscale(m) * sum_t colsum_in[t, m]
  int colsum_tiles;
  float* colsum_out;       // written by the (n_tile==0, split==0) unit
  long long* timing;       // debug timeline [grid][8] (ns, %globaltimer) or nullptr
  int dbg;                 // test hook: bit0 skip global stores, bit1 skip TMEM loads
  int pdl_wait_producer;   // griddepcontrol.wait before the first operand load (B comes from the previous kernel)
  int pdl_wait_epilogue;   // griddepcontrol.wait before the first output access (output prepared by the previous kernel)
};

template <int BLOCK_N, bool A_MN, bool B_MN, int EPI, bool TF32, int STAGES>
struct GemmCfg {
  static constexpr int BLOCK_M = 128;
  static constexpr int ELT = TF32 ? 4 : 2;
  static constexpr int BLOCK_K = 128 / ELT;          // 64 bf16 / 32 tf32 : one 128B swizzle row
  static constexpr int UMMA_K = 32 / ELT;            // 16 bf16 / 8 tf32
  static constexpr int K_STEPS = BLOCK_K / UMMA_K;   // 4
  static constexpr int N_MMA = (BLOCK_N > 256) ? 2 : 1;
  static constexpr int UMMA_N = BLOCK_N / N_MMA;
  static constexpr int ACC_STAGES = (2 * BLOCK_N <= 512) ? 2 : 1;
  static constexpr int TMEM_COLS_RAW = ACC_STAGES * BLOCK_N;
  static constexpr int TMEM_COLS = TMEM_COLS_RAW <= 32 ? 32 : TMEM_COLS_RAW <= 64 ? 64 : TMEM_COLS_RAW <= 128 ? 128 : TMEM_COLS_RAW <= 256 ? 256 : 512;
  static constexpr int A_BYTES = BLOCK_M * 128;      // 16 KB
  static constexpr int B_BYTES = BLOCK_N * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SLAB = 128 / ELT;             // MN elements per 128B (MN-major slab width)
  // TMA boxes per stage
  static constexpr int A_BOXES = A_MN ? BLOCK_M / SLAB : 1;
  static constexpr int A_BOX_BYTES = A_BYTES / A_BOXES;
  static constexpr int B_BOXES = B_MN ? BLOCK_N / SLAB : N_MMA;
  static constexpr int B_BOX_BYTES = B_BYTES / B_BOXES;
  // epilogue: 8 warps (two per TMEM lane quarter, each owning half of the tile's columns) so that every
  // SM sub-partition has two warps to interleave -- a lone warp per scheduler is issue-latency bound.
  // Each warp has one 32-row x 128-byte swizzled transpose buffer.
  static constexpr int EPI_WARPS = 8;
  static constexpr int EPI_BUF_BYTES = 32 * 128;
  static constexpr int EPI_STAGING_BYTES = EPI_WARPS * EPI_BUF_BYTES;   // 32 KB
  static constexpr int HALF_N = BLOCK_N / 2;
  static constexpr int EPI_COLS = (EPI == EPI_STORE_BF16) ? 64 : 32;  // output columns per 128-byte row
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB of shared memory");
  static_assert(HALF_N % EPI_COLS == 0, "half tile width must be a multiple of the epilogue chunk");
  static constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
  static_assert(UMMA_N % 16 == 0 && UMMA_N >= 16 && UMMA_N <= 256, "invalid UMMA N");
  static_assert(!B_MN || (BLOCK_N % SLAB == 0 && N_MMA == 1), "MN-major B needs slab-aligned single MMA");
  static_assert((UMMA_N * 128) % 1024 == 0, "B half offset must keep 1024B swizzle alignment");
  static_assert(TMEM_COLS_RAW <= 512, "accumulators exceed TMEM");
};

__device__ __forceinline__ float row_group_scale(const GemmParams& p, int m) {
  if (p.gscale == nullptr) return 1.0f;
  float s = 0.0f;
#pragma unroll
  for (int g = 0; g < kMaxGroups; ++g) {
    if (g < p.G && m >= p.gstart[g] && m < p.gstart[g] + p.glen[g]) s = __ldg(p.gscale + g);
  }
  return s;
}

template <int BLOCK_N, bool A_MN, bool B_MN, int EPI, bool TF32, int STAGES>
__global__ void __launch_bounds__(64 + 32 * 8, 1)
bags_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N, A_MN, B_MN, EPI, TF32, STAGES>;
  constexpr int BLOCK_M = Cfg::BLOCK_M;
  constexpr int BLOCK_K = Cfg::BLOCK_K;

  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms.
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1 KB aligned, still a shared-space pointer
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  uint8_t* smem_epi = smem + STAGES * Cfg::STAGE_BYTES;  // 1024-byte aligned (all tile sizes are multiples of 1024)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + Cfg::EPI_STAGING_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + Cfg::ACC_STAGES;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * Cfg::ACC_STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { stamp(p.timing, 0); if (p.timing) p.timing[blockIdx.x * 8 + 7] = sm_id(); }
  pdl_trigger();   // a dependent kernel may begin its own prologue

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
#pragma unroll
    for (int a = 0; a < Cfg::ACC_STAGES; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], Cfg::EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_holder, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  if (threadIdx.x == 0) stamp(p.timing, 1);   // setup done

  const int units_per_split = p.num_m_tiles * p.num_n_tiles;
  const int num_units = units_per_split * p.num_splits;
  // balanced k-block ranges per split
  auto split_range = [&](int split, int& kb0, int& kb1) {
    const int base = p.kblocks_total / p.num_splits, rem = p.kblocks_total % p.num_splits;
    kb0 = split * base + (split < rem ? split : rem);
    kb1 = kb0 + base + (split < rem ? 1 : 0);
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      if (p.pdl_wait_producer) pdl_wait();
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
        const int split = u / units_per_split;
        const int t = u - split * units_per_split;
        const int m_tile = t / p.num_n_tiles, n_tile = t - m_tile * p.num_n_tiles;
        const int m0 = m_tile * BLOCK_M, n0 = n_tile * BLOCK_N;
        int kb0, kb1;
        split_range(split, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          const int k0 = kb * BLOCK_K;
          uint8_t* sa = smem_a + stage * Cfg::A_BYTES;
          uint8_t* sb = smem_b + stage * Cfg::B_BYTES;
#pragma unroll
          for (int i = 0; i < Cfg::A_BOXES; ++i) {
            if (A_MN) tma_load_2d(sa + i * Cfg::A_BOX_BYTES, &tmap_a, &full_bar[stage], m0 + i * Cfg::SLAB, k0);
            else      tma_load_2d(sa + i * Cfg::A_BOX_BYTES, &tmap_a, &full_bar[stage], k0, m0);
          }
#pragma unroll
          for (int i = 0; i < Cfg::B_BOXES; ++i) {
            if (B_MN) tma_load_2d(sb + i * Cfg::B_BOX_BYTES, &tmap_b, &full_bar[stage], n0 + i * Cfg::SLAB, k0);
            else      tma_load_2d(sb + i * Cfg::B_BOX_BYTES, &tmap_b, &full_bar[stage], k0, n0 + i * Cfg::UMMA_N);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== UMMA issuer (single thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_instr_desc(TF32 ? 2u : 1u, A_MN, B_MN, BLOCK_M, Cfg::UMMA_N);
      // K-major (128B swizzle): rows are 128 B, 8-row atoms are 1024 B apart (SBO);
      //   a K step of UMMA_K elements advances the start address by 32 B.
      // MN-major (128B swizzle): 64(bf16)-wide MN slabs, each K row is 128 B, 8-row
      //   K groups 1024 B apart (SBO), slabs BLOCK_K*128 B apart (LBO);
      //   a K step advances UMMA_K rows = UMMA_K*128 B.
      // MN-major tf32 must use the 32B-atom swizzle: K atoms are 4 rows (512 B) instead of 8.
      constexpr uint64_t A_LAYOUT = (A_MN && TF32) ? kSwizzle128B_Base32B : kSwizzle128B;
      constexpr uint64_t B_LAYOUT = (B_MN && TF32) ? kSwizzle128B_Base32B : kSwizzle128B;
      constexpr uint32_t A_LBO = A_MN ? BLOCK_K * 128 : 16, A_SBO = (A_MN && TF32) ? 512 : 1024;
      constexpr uint32_t B_LBO = B_MN ? BLOCK_K * 128 : 16, B_SBO = (B_MN && TF32) ? 512 : 1024;
      constexpr uint32_t A_KSTEP = A_MN ? Cfg::UMMA_K * 128 : 32;
      constexpr uint32_t B_KSTEP = B_MN ? Cfg::UMMA_K * 128 : 32;
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int u = blockIdx.x; u < num_units; u += gridDim.x, ++local) {
        const int split = u / units_per_split;
        int kb0, kb1;
        split_range(split, kb0, kb1);
        const int acc = local % Cfg::ACC_STAGES;
        const uint32_t acc_phase = (local / Cfg::ACC_STAGES) & 1u;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (local == 0 && kb == kb0) stamp(p.timing, 2);   // first operands landed
          const uint32_t sa = smem_u32(smem_a + stage * Cfg::A_BYTES);
          const uint32_t sb = smem_u32(smem_b + stage * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < Cfg::K_STEPS; ++k) {
            const uint64_t adesc = make_smem_desc(sa + k * A_KSTEP, A_LBO, A_SBO, A_LAYOUT);
#pragma unroll
            for (int h = 0; h < Cfg::N_MMA; ++h) {
              const uint64_t bdesc = make_smem_desc(sb + h * Cfg::UMMA_N * 128 + k * B_KSTEP, B_LBO, B_SBO, B_LAYOUT);
              const uint32_t accum = (kb > kb0 || k > 0) ? 1u : 0u;
              if (TF32) umma_tf32(d_tmem + h * Cfg::UMMA_N, adesc, bdesc, idesc, accum);
              else      umma_bf16(d_tmem + h * Cfg::UMMA_N, adesc, bdesc, idesc, accum);
            }
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tfull_bar[acc]);      // accumulator complete -> epilogue
        if (local == 0) stamp(p.timing, 3);   // all MMAs of the first tile issued
      }
    }
  } else {
    // ===================== epilogue warps =====================
    // TMEM -> registers (thread = accumulator row) -> bias / row scale / bf16 pack -> 128B-swizzled smem
    // tile (conflict-free 16 B writes) -> read back transposed so that every warp-wide 16 B store /
    // reduction covers 4 complete 128-byte output rows (fully coalesced).
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;   // which half of the tile's columns
    uint8_t* buf = smem_epi + (warp - 2) * Cfg::EPI_BUF_BYTES;
    if (p.pdl_wait_epilogue) pdl_wait();
    constexpr int OUT_ELT = (EPI == EPI_STORE_BF16) ? 2 : 4;
    constexpr int VEC = 16 / OUT_ELT;  // output elements per 16-byte vector
    const bool vec_ok = ((p.ldo * OUT_ELT) % 16 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    int local = 0;
    for (int u = blockIdx.x; u < num_units; u += gridDim.x, ++local) {
      const int split = u / units_per_split;
      const int t = u - split * units_per_split;
      const int m_tile = t / p.num_n_tiles, n_tile = t - m_tile * p.num_n_tiles;
      const int m_warp = m_tile * BLOCK_M + quarter * 32;
      const int m = m_warp + lane;
      const int n0 = n_tile * BLOCK_N;
      const int acc = local % Cfg::ACC_STAGES;
      const uint32_t acc_phase = (local / Cfg::ACC_STAGES) & 1u;

      float scale = 1.0f;
      if (EPI == EPI_RED_F32) {
        scale = (m < p.M) ? row_group_scale(p, m) : 0.0f;
        if (p.colsum_out != nullptr && n_tile == 0 && split == 0 && half == 0 && m < p.M) {
          float cs = 0.f;
          for (int tt = 0; tt < p.colsum_tiles; ++tt) cs += __ldg(p.colsum_in + static_cast<long long>(tt) * p.M + m);
          p.colsum_out[m] = scale * cs;
        }
      }

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if (local == 0 && warp == 2 && lane == 0) stamp(p.timing, 4);   // first accumulator complete
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
      for (int c = half * Cfg::HALF_N; c < (half + 1) * Cfg::HALF_N; c += Cfg::EPI_COLS) {
        const int n = n0 + c;
        if (n >= p.N) break;  // warp-uniform: nothing of this chunk is inside the output
        uint32_t v[32];
        uint32_t v2[32];
        if (!(p.dbg & 2)) {
          tmem_ld_32x32b_x32(t_row + c, v);
          if (EPI == EPI_STORE_BF16) tmem_ld_32x32b_x32(t_row + c + 32, v2);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) { v[j] = c + j; v2[j] = lane + j; }
        }
        uint4* rowp = reinterpret_cast<uint4*>(buf + lane * 128);
        if (EPI == EPI_STORE_BF16) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {  // 16-byte chunk j = 8 bf16 = columns 8j..8j+7
            const uint32_t* src = (j < 4) ? (v + 8 * j) : (v2 + 8 * (j - 4));
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(src[e]);
            if (p.bias != nullptr) {   // (all rows of the warp read the same addresses: one broadcast transaction each)
#pragma unroll
              for (int e = 0; e < 8; ++e)
                if (n + 8 * j + e < p.N) f[e] += __ldg(p.bias + n + 8 * j + e);
            }
            if (p.relu) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
            }
            uint4 r;
            r.x = pack_bf16x2(f[0], f[1]);
            r.y = pack_bf16x2(f[2], f[3]);
            r.z = pack_bf16x2(f[4], f[5]);
            r.w = pack_bf16x2(f[6], f[7]);
            rowp[j ^ (lane & 7)] = r;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {  // 16-byte chunk j = 4 fp32 = columns 4j..4j+3
            float4 r;
            r.x = __uint_as_float(v[4 * j + 0]); r.y = __uint_as_float(v[4 * j + 1]);
            r.z = __uint_as_float(v[4 * j + 2]); r.w = __uint_as_float(v[4 * j + 3]);
            if (EPI == EPI_STORE_F32) {
              if (p.bias != nullptr && n + 4 * j + 3 < p.N) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n + 4 * j));
                r.x += b.x; r.y += b.y; r.z += b.z; r.w += b.w;
              } else if (p.bias != nullptr) {
                if (n + 4 * j + 0 < p.N) r.x += __ldg(p.bias + n + 4 * j + 0);
                if (n + 4 * j + 1 < p.N) r.y += __ldg(p.bias + n + 4 * j + 1);
                if (n + 4 * j + 2 < p.N) r.z += __ldg(p.bias + n + 4 * j + 2);
              }
              if (p.relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
            } else {
              r.x *= scale; r.y *= scale; r.z *= scale; r.w *= scale;
            }
            rowp[j ^ (lane & 7)] = *reinterpret_cast<uint4*>(&r);
          }
        }
        __syncwarp();
        // read back: iteration it covers tile rows 4*it .. 4*it+3, lane -> (row, 16-byte chunk)
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 4 + (lane >> 3);
          const int ch = lane & 7;
          const uint4 val = *reinterpret_cast<const uint4*>(buf + r * 128 + ((ch ^ (r & 7)) << 4));
          const int gm = m_warp + r;
          const int gn = n + ch * VEC;
          if (gm < p.M && gn < p.N && !(p.dbg & 1)) {
            uint8_t* gp = reinterpret_cast<uint8_t*>(p.out) + (static_cast<long long>(gm) * p.ldo + gn) * OUT_ELT;
            if (vec_ok && gn + VEC <= p.N) {
              if (EPI == EPI_RED_F32)
                red_add_v4_f32(reinterpret_cast<float*>(gp), __uint_as_float(val.x), __uint_as_float(val.y),
                               __uint_as_float(val.z), __uint_as_float(val.w));
              else
                *reinterpret_cast<uint4*>(gp) = val;
            } else {
              const uint32_t w4[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
              for (int e = 0; e < VEC; ++e) {
                if (gn + e < p.N) {
                  if (EPI == EPI_STORE_BF16) {
                    const uint32_t word = w4[e >> 1];
                    reinterpret_cast<unsigned short*>(gp)[e] = static_cast<unsigned short>((e & 1) ? (word >> 16) : (word & 0xffffu));
                  } else if (EPI == EPI_RED_F32) {
                    red_add_f32(reinterpret_cast<float*>(gp) + e, __uint_as_float(w4[e]));
                  } else {
                    reinterpret_cast<float*>(gp)[e] = __uint_as_float(w4[e]);
                  }
                }
              }
            }
          }
        }
        __syncwarp();   // buffer is rewritten by the next chunk
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
    if (warp == 2 && lane == 0) stamp(p.timing, 5);   // epilogue done
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
  if (threadIdx.x == 0) stamp(p.timing, 6);
}

}  // namespace bags
