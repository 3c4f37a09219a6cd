// tcgen05 / TMA bf16 GEMM for sm_100a — the contraction engine of the LoRA-SFT step.
//
// Replaces the cuBLAS hgemm calls hidden behind nn.Linear / peft lora.Linear on the reference's hot
// path (SURVEY §2.3 K3,K4,K7,K8,K10; call sites cmd/tuning/train.py:236-242,268-280,299).
//
// Structure (one CTA per SM, persistent over output tiles, 192 threads):
//   warp 0      TMA producer     : cp.async.bulk.tensor loads of A/B k-blocks into a STAGES-deep smem ring
//   warp 1      MMA issuer       : one thread issues tcgen05.mma (128 x BN x 16 per instruction), accumulator in TMEM
//   warps 2..5  epilogue         : tcgen05.ld TMEM -> registers -> (convert / +residual) -> global
// TMEM holds two BN-column accumulators so the epilogue of tile i overlaps the main loop of tile i+1.
// The contraction may be extended by a second (A2,B2) segment: that is how the rank-r LoRA update is
// accumulated into the same TMEM tile as the frozen base weight (one extra k-block, no extra pass).
#include "common.cuh"
#include "kernels.h"

#include <map>
#include <mutex>
#include <tuple>

namespace dtx {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 192;
constexpr int GROUP_M = 16;

struct GemmKParams {
  int M, N;
  int kb1, kb2;     // 64-wide k-blocks in segment 1 / segment 2
  int split_k;
  int kb_per_split;
  int m_tiles, n_tiles;
  void* C;
  long long ldc;
  const bf16* R;
  long long ldr;
  int group_m;  // pair kernel: 256-row tiles per rasterisation group
  void* aux;
  long long ld_aux;
  const float2* rope_cs;
  int rope_S, rope_cols, rope_inverse;
  const int32_t* rope_pos;  // optional per-row position (packed ragged batches); null: row % rope_S
  const int32_t* m_eff;  // pair kernel: device-side row count; 256-row tiles that start at or beyond it are skipped
};

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : ((BN == 128) ? 6 : 8);
  static constexpr int TMEM_COLS = 2 * BN;  // 512 / 256 / 128 : powers of two >= 32
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// store one 32-column chunk of an accumulator row (fp32 in registers) with the selected epilogue
template <int EPI>
__device__ __forceinline__ void epilogue_store_chunk(const GemmKParams& p, int row, int col0, int split, const uint32_t (&v)[32]) {
  const bool full = (col0 + 32 <= p.N);
  if (EPI == EPI_F32) {
    float* crow = reinterpret_cast<float*>(p.C) + static_cast<long long>(split) * p.M * p.ldc +
                  static_cast<long long>(row) * p.ldc + col0;
    if (full && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        reinterpret_cast<uint4*>(crow)[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)  // statically indexed (a dynamic index would push the whole fragment to local memory)
        if (col0 + j < p.N) crow[j] = __uint_as_float(v[j]);
    }
  } else {
    bf16* crow = reinterpret_cast<bf16*>(p.C) + static_cast<long long>(row) * p.ldc + col0;
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
    if (EPI == EPI_BF16_ADD) {
      const bf16* rrow = p.R + static_cast<long long>(row) * p.ldr + col0;
      if (full && ((reinterpret_cast<uintptr_t>(rrow) & 15) == 0)) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 r = reinterpret_cast<const uint4*>(rrow)[j];
          float2 a = unpack_bf16x2(r.x), b = unpack_bf16x2(r.y), cc = unpack_bf16x2(r.z), d = unpack_bf16x2(r.w);
          f[8 * j + 0] += a.x; f[8 * j + 1] += a.y; f[8 * j + 2] += b.x; f[8 * j + 3] += b.y;
          f[8 * j + 4] += cc.x; f[8 * j + 5] += cc.y; f[8 * j + 6] += d.x; f[8 * j + 7] += d.y;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < p.N) f[j] += __bfloat162float(rrow[j]);
      }
    }
    if (full && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 o;
        o.x = pack_bf16x2(f[8 * j + 0], f[8 * j + 1]);
        o.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
        o.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
        o.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
        reinterpret_cast<uint4*>(crow)[j] = o;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < p.N) crow[j] = __float2bfloat16_rn(f[j]);
    }
  }
}

template <int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2, const GemmKParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tfull_bar = empty_bar + Cfg::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB2);
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int kb_total = p.kb1 + p.kb2;
  const int tiles_mn = p.m_tiles * p.n_tiles;
  const int total_tiles = tiles_mn * p.split_k;

  auto decode = [&](int tile, int& m_blk, int& n_blk, int& split, int& kb_begin, int& kb_end) {
    split = tile / tiles_mn;
    int t = tile - split * tiles_mn;
    const int per_group = GROUP_M * p.n_tiles;
    int group = t / per_group;
    int first_m = group * GROUP_M;
    int gsz = min(GROUP_M, p.m_tiles - first_m);
    int r = t - group * per_group;
    m_blk = first_m + (r % gsz);
    n_blk = r / gsz;
    kb_begin = split * p.kb_per_split;
    kb_end = min(kb_total, kb_begin + p.kb_per_split);
  };

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    {
      const bool leader = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int m_blk, n_blk, split, kb_begin, kb_end;
        decode(tile, m_blk, n_blk, split, kb_begin, kb_end);
        const int m0 = m_blk * BM, n0 = n_blk * BN;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          uint8_t* a_dst = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* b_dst = a_dst + Cfg::A_BYTES;
          const bool seg2 = kb >= p.kb1;
          const int kk = (seg2 ? kb - p.kb1 : kb) * BK;
          const CUtensorMap* ma = seg2 ? &tmA2 : &tmA;
          const CUtensorMap* mb = seg2 ? &tmB2 : &tmB;
          if (leader) {
            if (!A_MN) {
              tma_load_2d(a_dst, ma, &full_bar[stage], kk, m0);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j) tma_load_2d(a_dst + j * 8192, ma, &full_bar[stage], m0 + 64 * j, kk);
            }
            if (!B_MN) {
              tma_load_2d(b_dst, mb, &full_bar[stage], kk, n0);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j) tma_load_2d(b_dst + j * 8192, mb, &full_bar[stage], n0 + 64 * j, kk);
            }
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    {
      const bool leader = elect_one();
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int m_blk, n_blk, split, kb_begin, kb_end;
        decode(tile, m_blk, n_blk, split, kb_begin, kb_end);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t b_base = a_base + Cfg::A_BYTES;
#pragma unroll
          for (int k16 = 0; k16 < BK / 16; ++k16) {
            const uint64_t da = A_MN ? umma_desc_mnmajor(a_base + k16 * 2048, 8192) : umma_desc_kmajor(a_base + k16 * 32);
            const uint64_t db = B_MN ? umma_desc_mnmajor(b_base + k16 * 2048, 8192) : umma_desc_kmajor(b_base + k16 * 32);
            if (leader) umma_bf16(d_tmem, da, db, idesc, (kb > kb_begin || k16 > 0) ? 1u : 0u);
          }
          if (leader) umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
        if (leader) umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ------------------------------ epilogue (warps 2..5) ------------------------------
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int m_blk, n_blk, split, kb_begin, kb_end;
      decode(tile, m_blk, n_blk, split, kb_begin, kb_end);
      const int m0 = m_blk * BM, n0 = n_blk * BN;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int col0 = n0 + c * 32;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t v[32];
        tmem_ld32(t_row + c * 32, v);
        tmem_ld_wait();
        if (!row_ok) continue;
        epilogue_store_chunk<EPI>(p, row, col0, split, v);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ================================================================================================
// CTA-pair kernel: 256 x 256 output tile per cluster of two CTAs (tcgen05 cta_group::2).
//
// The single-CTA kernel above tops out at ~67% tensor-pipe activity (ncu, profiles/r01_*): a 128x256x16 UMMA reads
// 12 KB of operands from shared memory per 128 cycles (96 B/clk) while TMA writes another 96 B/clk, against a 128 B/clk
// shared-memory port.  In pair mode each CTA stages only its half of B (128 rows) and the tensor cores exchange the
// halves, so per-SM shared-memory traffic drops to 64 + 64 B/clk and L2->SM traffic per FLOP by a third.
//   both CTAs : warp 0 = TMA producer (own A rows, own half of B; completion bytes land on the LEADER's barrier)
//   leader    : warp 1 lane 0 issues tcgen05.mma.cta_group::2 (M = 256) and multicast-commits to both CTAs' barriers
//   both CTAs : warps 2..5 drain their own 128 x 256 accumulator half from their own TMEM
// ================================================================================================
constexpr int G2_STAGES = 6;
constexpr int G2_A_BYTES = 128 * BK * 2;
constexpr int G2_B_BYTES = 128 * BK * 2;
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;
constexpr int G2_TMEM_COLS = 512;
constexpr int G2_SMEM_BYTES = G2_STAGES * G2_STAGE_BYTES + 1024 + 256;
// Epilogues that write more than they can afford to store row-by-row stage their output tile in shared memory and hand it to
// the TMA (SwiGLU backward: 1 KB read + 1 KB written per row and tile; one 16-byte access per thread touches 32 different
// 128-byte lines per instruction, and ~16 K such wavefronts per tile outlast the 16 K-cycle main loop of the next tile).
__host__ __device__ constexpr int g2_out_stage_bytes(int epi) { return epi == EPI_SWIGLU_BWD ? 32768 : 0; }
constexpr int G2_GROUP_M = 16;  // in 256-row tiles (swept in tools/sweep_group_m.py: 16 >= 8 on every step shape)

template <bool B_MN, int EPI, bool A_MN = false>  // A_MN: A is row-major [K, M] (weight-gradient GEMMs: dW = dY^T X)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
             const __grid_constant__ CUtensorMap tmC, const GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* out_stage = smem + G2_STAGES * G2_STAGE_BYTES;  // g2_out_stage_bytes(EPI) bytes, 1024-aligned
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + G2_STAGES * G2_STAGE_BYTES + g2_out_stage_bytes(EPI));
  uint64_t* empty_bar = full_bar + G2_STAGES;
  uint64_t* tfull_bar = empty_bar + G2_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB2);
    if (g2_out_stage_bytes(EPI)) tma_prefetch_desc(&tmC);
    for (int i = 0; i < G2_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);   // leader's producer arrive.expect_tx (both CTAs' bytes)
      mbar_init(&empty_bar[i], 1);  // one multicast tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 8);  // 4 epilogue warps of each CTA (only the leader's copy is waited on)
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_ptr_smem, G2_TMEM_COLS);
  tc_fence_before();
  __syncthreads();     // CTA-local ordering of the allocator's shared-memory write before the cluster barrier (racecheck still reports the
                       // cta_group::2 allocation instruction against itself: profiles/r02_sanitizer.txt)
  cluster_sync_all();  // both CTAs' barriers are initialised and their TMEM allocated before anyone signals across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int kb_total = p.kb1 + p.kb2;
  const int total_tiles = p.m_tiles * p.n_tiles;  // m_tiles counts 256-row tiles here
  // rows that exist on the device only (e.g. the number of unmasked tokens of this batch): every role skips the same tiles
  const int m_eff = p.m_eff ? min(p.M, __ldg(p.m_eff)) : p.M;

  auto decode = [&](int tile, int& m_blk, int& n_blk) {
    const int per_group = p.group_m * p.n_tiles;
    int group = tile / per_group;
    int first_m = group * p.group_m;
    int gsz = min(p.group_m, p.m_tiles - first_m);
    int r = tile - group * per_group;
    m_blk = first_m + (r % gsz);
    n_blk = r / gsz;
  };

  if (warp == 0) {
    {
      const bool leader = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs) {
        int m_blk, n_blk;
        decode(tile, m_blk, n_blk);
        if (m_blk * 256 >= m_eff) continue;
        const int m0 = m_blk * 256 + 128 * static_cast<int>(rank);
        const int n0 = n_blk * 256 + 128 * static_cast<int>(rank);
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          if (leader && rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * G2_STAGE_BYTES);
          const uint32_t leader_full = mapa_u32(smem_u32(&full_bar[stage]), 0);
          uint8_t* a_dst = smem + stage * G2_STAGE_BYTES;
          uint8_t* b_dst = a_dst + G2_A_BYTES;
          const bool seg2 = kb >= p.kb1;
          const int kk = (seg2 ? kb - p.kb1 : kb) * BK;
          const CUtensorMap* ma = seg2 ? &tmA2 : &tmA;
          const CUtensorMap* mb = seg2 ? &tmB2 : &tmB;
          if (leader) {
            if (!A_MN) {
              tma_load_2d_pair(a_dst, ma, leader_full, kk, m0);
            } else {  // two [64 k x 64 m] boxes
              tma_load_2d_pair(a_dst, ma, leader_full, m0, kk);
              tma_load_2d_pair(a_dst + 8192, ma, leader_full, m0 + 64, kk);
            }
            if (!B_MN) {
              tma_load_2d_pair(b_dst, mb, leader_full, kk, n0);
            } else {
              tma_load_2d_pair(b_dst, mb, leader_full, n0, kk);
              tma_load_2d_pair(b_dst + 8192, mb, leader_full, n0 + 64, kk);
            }
          }
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      const bool leader = elect_one();
      constexpr uint32_t idesc = umma_idesc_bf16(256, 256, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs) {
        {
          int m_blk, n_blk;
          decode(tile, m_blk, n_blk);
          if (m_blk * 256 >= m_eff) continue;
        }
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * 256);
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + stage * G2_STAGE_BYTES);
          const uint32_t b_base = a_base + G2_A_BYTES;
#pragma unroll
          for (int k16 = 0; k16 < BK / 16; ++k16) {
            const uint64_t da = A_MN ? umma_desc_mnmajor(a_base + k16 * 2048, 8192) : umma_desc_kmajor(a_base + k16 * 32);
            const uint64_t db = B_MN ? umma_desc_mnmajor(b_base + k16 * 2048, 8192) : umma_desc_kmajor(b_base + k16 * 32);
            if (leader) umma_bf16_pair(d_tmem, da, db, idesc, (kb > 0 || k16 > 0) ? 1u : 0u);
          }
          if (leader) umma_commit_pair(&empty_bar[stage], 3);
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1u; }
        }
        if (leader) umma_commit_pair(&tfull_bar[acc], 3);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < total_tiles; tile += npairs) {
      int m_blk, n_blk;
      decode(tile, m_blk, n_blk);
      if (m_blk * 256 >= m_eff) continue;
      const int m0 = m_blk * 256 + 128 * static_cast<int>(rank);
      const int n0 = n_blk * 256;
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      // The epilogue's global reads (residual / gate|up rows) are pulled into L2 while the tensor core is still working on
      // this tile: without it the epilogue is latency-bound on HBM and outlasts the next tile's main loop (r01: tensor pipe
      // 64% active on the SwiGLU-backward GEMM, 83% on the residual GEMMs, vs 90% with a store-only epilogue).
      if (row_ok) {
        if (EPI == EPI_BF16_ADD) {
          const bf16* rr = p.R + static_cast<long long>(row) * p.ldr + n0;
#pragma unroll
          for (int i = 0; i < 4; ++i) prefetch_l2(rr + i * 64);
        } else if (EPI == EPI_SWIGLU_BWD) {
          const bf16* gr = reinterpret_cast<const bf16*>(p.aux) + static_cast<long long>(row) * p.ld_aux + static_cast<long long>(n0 >> 7) * 256;
#pragma unroll
          for (int i = 0; i < 8; ++i) prefetch_l2(gr + i * 64);
        }
      }
      // SwiGLU backward: the gate|up fragments of the first two 32-feature chunks are requested before the wait for the
      // accumulator, later ones two chunks ahead of their use (profiles/r01_ncu_swiglu_bwd_gemm_staged.txt: with a one-chunk
      // lookahead issued after the wait, 31 % of the epilogue's samples sat on the first use of these loads).
      uint4 gq[4], uq[4], gn[4], un[4], gm[4], um[4];
      if (EPI == EPI_SWIGLU_BWD && row_ok) {
        const bf16* gbase0 = reinterpret_cast<const bf16*>(p.aux) + static_cast<long long>(row) * p.ld_aux;
        const long long c0 = static_cast<long long>(n0 >> 7) * 256 + (n0 & 127);
        if (n0 < p.N) {
#pragma unroll
          for (int j = 0; j < 4; ++j) { gq[j] = reinterpret_cast<const uint4*>(gbase0 + c0)[j]; uq[j] = reinterpret_cast<const uint4*>(gbase0 + c0 + 128)[j]; }
        }
        if (n0 + 32 < p.N) {
#pragma unroll
          for (int j = 0; j < 4; ++j) { gn[j] = reinterpret_cast<const uint4*>(gbase0 + c0 + 32)[j]; un[j] = reinterpret_cast<const uint4*>(gbase0 + c0 + 160)[j]; }
        }
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * 256);
      if (EPI == EPI_ROPE) {
        // two heads per tile (chunks 0-3 / 4-7); rotary pairs column i with i+64: chunk c with chunk c+2
        const bool rot = n0 < p.rope_cols;
        const float2* cs = p.rope_cs + static_cast<long long>(row_ok ? (p.rope_pos ? __ldg(p.rope_pos + row) : row % p.rope_S) : 0) * 64;
#pragma unroll 1
        for (int hc = 0; hc < 4; ++hc) {
          const int c = (hc >> 1) * 4 + (hc & 1);
          if (n0 + c * 32 >= p.N) break;
          uint32_t lo[32], hi[32];
          tmem_ld32(t_row + c * 32, lo);
          tmem_ld32(t_row + (c + 2) * 32, hi);
          tmem_ld_wait();
          if (!row_ok) continue;
          if (rot) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float2 t = __ldg(cs + (hc & 1) * 32 + j);
              const float sn = p.rope_inverse ? -t.y : t.y;
              const float x0 = __uint_as_float(lo[j]), x1 = __uint_as_float(hi[j]);
              lo[j] = __float_as_uint(x0 * t.x - x1 * sn);
              hi[j] = __float_as_uint(x1 * t.x + x0 * sn);
            }
          }
          epilogue_store_chunk<EPI_BF16>(p, row, n0 + c * 32, 0, lo);
          epilogue_store_chunk<EPI_BF16>(p, row, n0 + (c + 2) * 32, 0, hi);
        }
      } else if (EPI == EPI_SWIGLU_FWD) {
        // tile = [gate 128 | up 128] of the same 128 features: chunk c (gate) pairs with chunk c+4 (up)
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t g[32], u[32];
          tmem_ld32(t_row + c * 32, g);
          tmem_ld32(t_row + (c + 4) * 32, u);
          tmem_ld_wait();
          if (!row_ok) continue;
          epilogue_store_chunk<EPI_BF16>(p, row, n0 + c * 32, 0, g);
          epilogue_store_chunk<EPI_BF16>(p, row, n0 + (c + 4) * 32, 0, u);
          bf16* arow = reinterpret_cast<bf16*>(p.aux) + static_cast<long long>(row) * p.ld_aux + (n0 >> 1) + c * 32;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float a[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              // like the unfused path, gate/up are rounded to bf16 before the activation (fast reciprocal: ~1 ulp apart)
              const float gg = __bfloat162float(__float2bfloat16_rn(__uint_as_float(g[8 * j + e])));
              const float uu = __bfloat162float(__float2bfloat16_rn(__uint_as_float(u[8 * j + e])));
              a[e] = gg * __fdividef(1.f, 1.f + __expf(-gg)) * uu;
            }
            uint4 o;
            o.x = pack_bf16x2(a[0], a[1]); o.y = pack_bf16x2(a[2], a[3]);
            o.z = pack_bf16x2(a[4], a[5]); o.w = pack_bf16x2(a[6], a[7]);
            reinterpret_cast<uint4*>(arow)[j] = o;
          }
        }
      } else if (EPI == EPI_SWIGLU_BWD) {
        // acc = d(act) for features n0 + 32c + j; gate/up of feature f live at (f/128)*256 + f%128 (+128) of gu / d(gu).
        // The gate/up fragments of chunk c+1 are requested before chunk c is processed: the epilogue is latency-bound
        // on these HBM reads otherwise (r01: +420 us on the 1.0 ms GEMM without the prefetch).
        // Output goes through a staging tile [d gate 128 x 64 | d up 128 x 64] (128B-swizzled) and two TMA stores per 64 features.
        const bf16* gbase = reinterpret_cast<const bf16*>(p.aux) + static_cast<long long>(row_ok ? row : 0) * p.ld_aux;
        const uint32_t stg = smem_u32(out_stage);
        const int r = q * 32 + lane;
        const bool store_thread = (warp == 2 && lane == 0);
        auto gcol_of = [&](int c) { const int f0 = n0 + c * 32; return static_cast<long long>(f0 >> 7) * 256 + (f0 & 127); };
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
          const int f0 = n0 + c * 32;
          if (f0 >= p.N) break;
          const bool has_next2 = (c + 2 < 8) && (f0 + 64 < p.N);
          if (row_ok && has_next2) {
            const bf16* g2 = gbase + gcol_of(c + 2);
#pragma unroll
            for (int j = 0; j < 4; ++j) { gm[j] = reinterpret_cast<const uint4*>(g2)[j]; um[j] = reinterpret_cast<const uint4*>(g2 + 128)[j]; }
          }
          uint32_t v[32];
          tmem_ld32(t_row + c * 32, v);
          tmem_ld_wait();
          if ((c & 1) == 0) {  // the staging tile is free once the previous pair's bulk stores have read it
            if (store_thread) tma_store_wait_read0();
            named_bar_sync(1, 128);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t gw[4] = {gq[j].x, gq[j].y, gq[j].z, gq[j].w}, uw[4] = {uq[j].x, uq[j].y, uq[j].z, uq[j].w};
            uint32_t og[4], ou[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 g2 = unpack_bf16x2(gw[e]), u2 = unpack_bf16x2(uw[e]);
              // d(act) is rounded to bf16 first, exactly like the unfused path that stores it
              const float d0 = __bfloat162float(__float2bfloat16_rn(__uint_as_float(v[8 * j + 2 * e])));
              const float d1 = __bfloat162float(__float2bfloat16_rn(__uint_as_float(v[8 * j + 2 * e + 1])));
              const float s0 = __fdividef(1.f, 1.f + __expf(-g2.x)), s1 = __fdividef(1.f, 1.f + __expf(-g2.y));
              const float si0 = g2.x * s0, si1 = g2.y * s1;
              og[e] = pack_bf16x2(d0 * u2.x * (s0 + si0 * (1.f - s0)), d1 * u2.y * (s1 + si1 * (1.f - s1)));
              ou[e] = pack_bf16x2(d0 * si0, d1 * si1);
            }
            const uint32_t off = sw128_offset(r, (c & 1) * 4 + j);
            sts128(stg + off, og[0], og[1], og[2], og[3]);
            sts128(stg + 16384 + off, ou[0], ou[1], ou[2], ou[3]);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) { gq[j] = gn[j]; uq[j] = un[j]; gn[j] = gm[j]; un[j] = um[j]; }
          if ((c & 1) == 1) {  // 64 features staged: rows beyond M are clipped by the tensor map
            fence_proxy_async_smem();
            named_bar_sync(1, 128);
            if (store_thread) {
              const int gc = static_cast<int>(gcol_of(c - 1));
              tma_store_2d(&tmC, out_stage, gc, m0);
              tma_store_2d(&tmC, out_stage + 16384, gc + 128, m0);
              tma_store_commit();
            }
          }
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
          const int col0 = n0 + c * 32;
          if (col0 >= p.N) break;
          uint32_t v[32];
          tmem_ld32(t_row + c * 32, v);
          tmem_ld_wait();
          if (!row_ok) continue;
          epilogue_store_chunk<EPI>(p, row, col0, 0, v);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));  // the leader's MMA thread waits on it
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  if (g2_out_stage_bytes(EPI) && warp == 2 && lane == 0) tma_store_wait_read0();  // staging tile must outlive the bulk stores' reads
  tc_fence_before();
  cluster_sync_all();  // no CTA may exit (or free TMEM) while its peer can still signal / multicast into it
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, G2_TMEM_COLS);
  }
}

// ----------------------------------------------------------------------------------------------
// host side: tensor maps
// ----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

}  // namespace

// 2-D bf16 tensor map over a row-major [outer, inner] array with row stride ld (elements), 128B swizzle.
bool make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                       uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

int gemm_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

namespace {

// runs `fn` once per CUDA device (thread-safe): function attributes are per-device state
struct DeviceOnce {
  std::mutex mu;
  bool done[64] = {false};
  template <typename F>
  cudaError_t run(F fn) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(mu);
    if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
    e = fn();
    if (e == cudaSuccess && dev >= 0 && dev < 64) done[dev] = true;
    return e;
  }
};

template <int BN, bool A_MN, bool B_MN, int EPI>
cudaError_t launch(const GemmArgs& a, cudaStream_t s) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_kernel<BN, A_MN, B_MN, EPI>;
  static DeviceOnce attr;  // the attribute is per device: a second trainer on another GPU of the same process needs it too
  if (cudaError_t e = attr.run([&] { return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES); });
      e != cudaSuccess)
    return e;
  CUtensorMap tA, tB, tA2, tB2;
  bool ok = true;
  // K-major operand: global [rows, K], box {64 k, rows_per_tile}; MN-major: global [K, cols], box {64 cols, 64 k}
  ok &= A_MN ? make_tmap_2d_bf16(&tA, a.A, a.M, a.K, a.lda, 64, 64) : make_tmap_2d_bf16(&tA, a.A, a.K, a.M, a.lda, 64, BM);
  ok &= B_MN ? make_tmap_2d_bf16(&tB, a.B, a.N, a.K, a.ldb, 64, 64) : make_tmap_2d_bf16(&tB, a.B, a.K, a.N, a.ldb, 64, BN);
  if (a.K2 > 0) {
    ok &= A_MN ? make_tmap_2d_bf16(&tA2, a.A2, a.M, a.K2, a.lda2, 64, 64)
               : make_tmap_2d_bf16(&tA2, a.A2, a.K2, a.M, a.lda2, 64, BM);
    ok &= B_MN ? make_tmap_2d_bf16(&tB2, a.B2, a.N, a.K2, a.ldb2, 64, 64)
               : make_tmap_2d_bf16(&tB2, a.B2, a.K2, a.N, a.ldb2, 64, BN);
  } else {
    tA2 = tA;
    tB2 = tB;
  }
  if (!ok) return cudaErrorInvalidValue;

  GemmKParams p;
  p.M = a.M;
  p.N = a.N;
  p.kb1 = (a.K + BK - 1) / BK;
  p.kb2 = (a.K2 + BK - 1) / BK;
  p.split_k = a.split_k < 1 ? 1 : a.split_k;
  int kb_total = p.kb1 + p.kb2;
  if (p.split_k > kb_total) p.split_k = kb_total;
  p.kb_per_split = (kb_total + p.split_k - 1) / p.split_k;
  // every split must own at least one k-block (an empty split would publish a stale accumulator)
  while (p.split_k > 1 && (p.split_k - 1) * p.kb_per_split >= kb_total) --p.split_k;
  if (p.split_k != (a.split_k < 1 ? 1 : a.split_k)) return cudaErrorInvalidValue;
  p.m_tiles = (a.M + BM - 1) / BM;
  p.n_tiles = (a.N + BN - 1) / BN;
  p.C = a.C;
  p.ldc = a.ldc;
  p.R = a.R;
  p.ldr = a.ldr;
  p.group_m = 0;
  p.aux = nullptr;
  p.ld_aux = 0;
  p.rope_cs = nullptr;
  p.rope_S = p.rope_cols = p.rope_inverse = 0;
  p.rope_pos = nullptr;
  p.m_eff = nullptr;  // the single-CTA kernel computes every row (rows beyond the device-side count are dead, not wrong)
  int total = p.m_tiles * p.n_tiles * p.split_k;
  int grid = total < gemm_num_sms() ? total : gemm_num_sms();
  if (grid <= 0) return cudaSuccess;
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, s>>>(tA, tB, tA2, tB2, p);
  return cudaGetLastError();
}

template <int BN, bool A_MN, bool B_MN>
cudaError_t launch_epi(const GemmArgs& a, cudaStream_t s) {
  switch (a.epilogue) {
    case EPI_BF16: return launch<BN, A_MN, B_MN, EPI_BF16>(a, s);
    case EPI_F32: return launch<BN, A_MN, B_MN, EPI_F32>(a, s);
    case EPI_BF16_ADD: return launch<BN, A_MN, B_MN, EPI_BF16_ADD>(a, s);
  }
  return cudaErrorInvalidValue;
}

int g_use_pair_kernel = 1;
int g_pair_group_m = G2_GROUP_M;

template <bool B_MN, int EPI, bool A_MN = false>
cudaError_t launch2(const GemmArgs& a, cudaStream_t s) {
  auto kern = gemm2_kernel<B_MN, EPI, A_MN>;
  static DeviceOnce attr;
  if (cudaError_t e = attr.run([&] {
        return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BYTES + g2_out_stage_bytes(EPI));
      });
      e != cudaSuccess)
    return e;
  CUtensorMap tA, tB, tA2, tB2, tC;
  bool ok = true;
  if (EPI == EPI_SWIGLU_BWD)  // d(gate|up) [M, 2N] bf16, stored in [128 x 64] boxes
    ok &= make_tmap_2d_bf16(&tC, a.C, 2ull * static_cast<uint64_t>(a.N), a.M, a.ldc, 64, 128);
  ok &= A_MN ? make_tmap_2d_bf16(&tA, a.A, a.M, a.K, a.lda, 64, 64) : make_tmap_2d_bf16(&tA, a.A, a.K, a.M, a.lda, 64, 128);
  ok &= B_MN ? make_tmap_2d_bf16(&tB, a.B, a.N, a.K, a.ldb, 64, 64) : make_tmap_2d_bf16(&tB, a.B, a.K, a.N, a.ldb, 64, 128);
  if (A_MN && a.K2 > 0) return cudaErrorInvalidValue;
  if (a.K2 > 0) {
    ok &= make_tmap_2d_bf16(&tA2, a.A2, a.K2, a.M, a.lda2, 64, 128);
    ok &= B_MN ? make_tmap_2d_bf16(&tB2, a.B2, a.N, a.K2, a.ldb2, 64, 64) : make_tmap_2d_bf16(&tB2, a.B2, a.K2, a.N, a.ldb2, 64, 128);
  } else {
    tA2 = tA;
    tB2 = tB;
  }
  if (!ok) return cudaErrorInvalidValue;
  GemmKParams p;
  p.M = a.M;
  p.N = a.N;
  p.kb1 = (a.K + BK - 1) / BK;
  p.kb2 = (a.K2 + BK - 1) / BK;
  p.split_k = 1;
  p.kb_per_split = p.kb1 + p.kb2;
  p.m_tiles = (a.M + 255) / 256;
  p.n_tiles = (a.N + 255) / 256;
  p.C = a.C;
  p.ldc = a.ldc;
  p.R = a.R;
  p.ldr = a.ldr;
  p.group_m = g_pair_group_m;
  p.aux = a.aux;
  p.ld_aux = a.ld_aux;
  p.rope_cs = a.rope_cs;
  p.rope_S = a.rope_S;
  p.rope_cols = a.rope_cols;
  p.rope_inverse = a.rope_inverse;
  p.rope_pos = a.rope_pos;
  p.m_eff = a.m_eff;
  const int total = p.m_tiles * p.n_tiles;
  int pairs = gemm_num_sms() / 2;
  if (pairs > total) pairs = total;
  if (pairs <= 0) return cudaSuccess;
  if (EPI != EPI_SWIGLU_BWD) tC = tA;
  kern<<<2 * pairs, GEMM_THREADS, G2_SMEM_BYTES + g2_out_stage_bytes(EPI), s>>>(tA, tB, tA2, tB2, tC, p);
  return cudaGetLastError();
}

template <bool B_MN>
cudaError_t launch2_epi(const GemmArgs& a, cudaStream_t s) {
  switch (a.epilogue) {
    case EPI_BF16: return launch2<B_MN, EPI_BF16>(a, s);
    case EPI_F32: return launch2<B_MN, EPI_F32>(a, s);
    case EPI_BF16_ADD: return launch2<B_MN, EPI_BF16_ADD>(a, s);
    case EPI_ROPE: return launch2<B_MN, EPI_ROPE>(a, s);
    case EPI_SWIGLU_FWD: return launch2<B_MN, EPI_SWIGLU_FWD>(a, s);
    case EPI_SWIGLU_BWD: return launch2<B_MN, EPI_SWIGLU_BWD>(a, s);
  }
  return cudaErrorInvalidValue;
}

template <int BN>
cudaError_t launch_major(const GemmArgs& a, cudaStream_t s) {
  if (!a.a_mn_major && !a.b_mn_major) return launch_epi<BN, false, false>(a, s);
  if (!a.a_mn_major && a.b_mn_major) return launch_epi<BN, false, true>(a, s);
  if (a.a_mn_major && a.b_mn_major) return launch_epi<BN, true, true>(a, s);
  return cudaErrorInvalidValue;  // (MN-major A, K-major B) is not used on the training path
}

}  // namespace

void gemm_set_pair_kernel(int on) { g_use_pair_kernel = on; }
void gemm_set_pair_group_m(int g) { g_pair_group_m = g > 0 ? g : G2_GROUP_M; }

cudaError_t gemm_bf16(const GemmArgs& a, cudaStream_t s) {
  if (a.M <= 0 || a.N <= 0 || a.K <= 0) return cudaErrorInvalidValue;
  if (a.split_k > 1 && a.epilogue != EPI_F32) return cudaErrorInvalidValue;
  if (a.epilogue == EPI_BF16_ADD && a.R == nullptr) return cudaErrorInvalidValue;
  if (a.epilogue >= EPI_ROPE) {  // fused epilogues exist only in the CTA-pair kernel and need whole 256-column tiles
    if (a.a_mn_major || a.split_k > 1 || a.M <= 128 || !g_use_pair_kernel) return cudaErrorInvalidValue;
    if (a.epilogue == EPI_ROPE && (!a.rope_cs || a.rope_S <= 0 || (a.rope_cols & 255) || (a.N & 255))) return cudaErrorInvalidValue;
    if (a.epilogue == EPI_SWIGLU_FWD && (!a.aux || (a.N & 255))) return cudaErrorInvalidValue;
    if (a.epilogue == EPI_SWIGLU_BWD && (!a.aux || (a.N & 127))) return cudaErrorInvalidValue;
    return a.b_mn_major ? launch2_epi<true>(a, s) : launch2_epi<false>(a, s);
  }
  // TMA needs 16-byte aligned bases and row strides
  if ((a.lda & 7) || (a.ldb & 7) || (reinterpret_cast<uintptr_t>(a.A) & 15) || (reinterpret_cast<uintptr_t>(a.B) & 15))
    return cudaErrorInvalidValue;
  if (a.K2 > 0 && ((a.lda2 & 7) || (a.ldb2 & 7) || !a.A2 || !a.B2)) return cudaErrorInvalidValue;
  int bn = a.block_n;
  if (bn == 0) bn = (a.N <= 64) ? 64 : ((a.N <= 128) ? 128 : 256);
  if (bn == 256 && !a.a_mn_major && a.split_k <= 1 && g_use_pair_kernel && a.M > 128)
    return a.b_mn_major ? launch2_epi<true>(a, s) : launch2_epi<false>(a, s);
  // weight-gradient GEMMs of full-parameter SFT (A and B MN-major, no K-extension, wide in both dimensions): CTA-pair kernel too
  if (bn == 256 && a.a_mn_major && a.b_mn_major && a.split_k <= 1 && g_use_pair_kernel && a.M > 128 && a.N >= 256 && a.K2 == 0 &&
      (a.lda & 7) == 0 && (a.ldb & 7) == 0) {
    if (a.epilogue == EPI_BF16) return launch2<true, EPI_BF16, true>(a, s);
    if (a.epilogue == EPI_BF16_ADD) return launch2<true, EPI_BF16_ADD, true>(a, s);
  }
  switch (bn) {
    case 64: return launch_major<64>(a, s);
    case 128: return launch_major<128>(a, s);
    case 256: return launch_major<256>(a, s);
  }
  return cudaErrorInvalidValue;
}

}  // namespace dtx
