// Causal flash attention forward / backward for sm_100a, head_dim 128, on tcgen05 + TMA.
//
// Replaces HF 4.34 eager LlamaAttention (matmul -> mask add -> fp32 softmax -> matmul, SURVEY §2.3 K6,
// reached from cmd/tuning/train.py:299) and its autograd backward (K10).  The [B,H,S,S] score tensor
// never exists: scores live in TMEM, probabilities go TMEM -> registers -> swizzled smem -> tensor core.
//
// Data layout: packed qkv [B*S, 3*H*128] (per token: q heads | k heads | v heads), out [B*S, H*128],
// lse2 [B,H,S] = log2-domain log-sum-exp of the scaled scores (m + log2 l).
//
// Kernels (128 threads = 4 warps; thread r owns TMEM lane r = one row of the score tile):
//   attn_fwd_kernel   : CTA = 128 query rows, loops over 64-row KV blocks.  S = Q K^T (UMMA 128x64x16, K-major x K-major),
//                       online softmax in registers, P -> smem (K-major A operand), O_blk = P V (V is the MN-major B operand).
//   attn_dq_kernel    : CTA = 128 query rows.  S, dP = dO V^T, dS = P o (dP - delta) * scale, dQ += dS K (K as MN-major B),
//                       dQ accumulates in TMEM over the whole KV loop.
//   attn_dkv_kernel   : CTA = 128 KV rows, loops over 64-row Q blocks.  S^T = K Q^T, dP^T = V dO^T, dV += P^T dO, dK += dS^T Q.
// Two backward kernels instead of one with fp32 atomics on dQ: every reduction has a fixed order, so the
// step is bitwise reproducible (needed for the N-rank == 1-rank parity tests).
#include "common.cuh"
#include "kernels.h"

namespace dtx {

bool make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                       uint32_t box_outer);

namespace {

constexpr int HD = 128;      // head dim
constexpr int ATT_THREADS = 128;
constexpr float LOG2E = 1.4426950408889634f;

struct AttnKParams {
  int B, S, H;
  float scale_log2;  // softmax scale * log2(e)
  float scale;
  float* lse2;
  bf16* out;
  const float* delta;
  bf16* dqkv;
};

// issue `n16` UMMAs (K=16 each) walking both operands.  a/b step = byte advance per k16.
// K-major operands made of 64-wide subtiles: k16 -> subtile (k16/4) + 32 B * (k16%4).
__device__ __forceinline__ uint32_t kmaj_addr(uint32_t base, int k16, uint32_t subtile_bytes) {
  return base + (k16 >> 2) * subtile_bytes + (k16 & 3) * 32;
}

// ================================================================================================
// forward
// ================================================================================================
constexpr int FWD_SQ = 0;                 // 2 x [128 x 128B]
constexpr int FWD_SK = 32768;             // 2 x [64 x 128B]
constexpr int FWD_SV = FWD_SK + 16384;    // 2 x [64 x 128B]
constexpr int FWD_SP = FWD_SV + 16384;    // [128 x 128B]
constexpr int FWD_BAR = FWD_SP + 16384;   // barriers
constexpr int FWD_SMEM = FWD_BAR + 128 + 1024;

__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const AttnKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FWD_BAR);
  uint64_t *bar_q = bars, *bar_k = bars + 1, *bar_v = bars + 2, *bar_s = bars + 3, *bar_o = bars + 4;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 6);

  const int nqb = p.S / 128;
  const int qb = nqb - 1 - (blockIdx.x % nqb);  // heavy (late) query blocks first
  const int bh = blockIdx.x / nqb;
  const int h = bh % p.H, b = bh / p.H;
  const int q0 = qb * 128;
  const int row_base = b * p.S;  // token row of position 0
  const int n_kv = (q0 + 128) / 64;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int colQ = h * HD, colK = p.H * HD + h * HD, colV = 2 * p.H * HD + h * HD;

  if (tid == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t t_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  const uint32_t T_S = 0, T_O = 64;

  if (tid == 0) {
    mbar_arrive_expect_tx(bar_q, 32768);
    tma_load_2d(smem + FWD_SQ, &tmQ, bar_q, colQ, row_base + q0);
    tma_load_2d(smem + FWD_SQ + 16384, &tmQ, bar_q, colQ + 64, row_base + q0);
    mbar_arrive_expect_tx(bar_k, 16384);
    tma_load_2d(smem + FWD_SK, &tmKV, bar_k, colK, row_base);
    tma_load_2d(smem + FWD_SK + 8192, &tmKV, bar_k, colK + 64, row_base);
    mbar_arrive_expect_tx(bar_v, 16384);
    tma_load_2d(smem + FWD_SV, &tmKV, bar_v, colV, row_base);
    tma_load_2d(smem + FWD_SV + 8192, &tmKV, bar_v, colV + 64, row_base);
  }

  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);
  constexpr uint32_t idesc_o = umma_idesc_bf16(128, 128, 0, 1);
  const uint32_t sQ = smem_u32(smem + FWD_SQ), sK = smem_u32(smem + FWD_SK), sV = smem_u32(smem + FWD_SV),
                 sP = smem_u32(smem + FWD_SP);

  float m_run = -INFINITY, l_run = 0.f;
  float o[HD];
#pragma unroll
  for (int i = 0; i < HD; ++i) o[i] = 0.f;
  const int qrow = q0 + tid;

  for (int j = 0; j < n_kv; ++j) {
    const uint32_t ph = j & 1;
    const int kv0 = j * 64;
    if (tid == 0) {
      if (j == 0) mbar_wait(bar_q, 0);
      mbar_wait(bar_k, ph);
      tc_fence_after();
#pragma unroll
      for (int k16 = 0; k16 < 8; ++k16)
        umma_bf16(tmem + T_S, umma_desc_kmajor(kmaj_addr(sQ, k16, 16384)), umma_desc_kmajor(kmaj_addr(sK, k16, 8192)),
                  idesc_s, k16 > 0 ? 1u : 0u);
      umma_commit(bar_s);
    }
    mbar_wait(bar_s, ph);
    tc_fence_after();
    if (tid == 0 && j + 1 < n_kv) {  // K buffer is free: prefetch next K block
      mbar_arrive_expect_tx(bar_k, 16384);
      tma_load_2d(smem + FWD_SK, &tmKV, bar_k, colK, row_base + kv0 + 64);
      tma_load_2d(smem + FWD_SK + 8192, &tmKV, bar_k, colK + 64, row_base + kv0 + 64);
    }
    uint32_t sv[64];
    {
      uint32_t(&lo)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[0]);
      uint32_t(&hi)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[32]);
      tmem_ld32(t_lane + T_S, lo);
      tmem_ld32(t_lane + T_S + 32, hi);
      tmem_ld_wait();
    }
    const bool need_mask = (kv0 + 63 > q0);
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < 64; ++c) {
      float t = __uint_as_float(sv[c]) * p.scale_log2;
      if (need_mask && (kv0 + c > qrow)) t = -INFINITY;
      sv[c] = __float_as_uint(t);
      mx = fmaxf(mx, t);
    }
    const float m_new = fmaxf(m_run, mx);
    const float alpha = exp2f(m_run - m_new);
    float rs = 0.f;
    uint8_t* prow = smem + FWD_SP;
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
      float pv[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        pv[e] = exp2f(__uint_as_float(sv[c8 * 8 + e]) - m_new);
        rs += pv[e];
      }
      uint4 u;
      u.x = pack_bf16x2(pv[0], pv[1]); u.y = pack_bf16x2(pv[2], pv[3]);
      u.z = pack_bf16x2(pv[4], pv[5]); u.w = pack_bf16x2(pv[6], pv[7]);
      *reinterpret_cast<uint4*>(prow + sw128_offset(tid, c8)) = u;
    }
    l_run = l_run * alpha + rs;
    m_run = m_new;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      mbar_wait(bar_v, ph);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_bf16(tmem + T_O, umma_desc_kmajor(sP + kk * 32), umma_desc_mnmajor(sV + kk * 2048, 8192), idesc_o,
                  kk > 0 ? 1u : 0u);
      umma_commit(bar_o);
    }
    mbar_wait(bar_o, ph);
    tc_fence_after();
    if (tid == 0 && j + 1 < n_kv) {  // V buffer is free
      mbar_arrive_expect_tx(bar_v, 16384);
      tma_load_2d(smem + FWD_SV, &tmKV, bar_v, colV, row_base + kv0 + 64);
      tma_load_2d(smem + FWD_SV + 8192, &tmKV, bar_v, colV + 64, row_base + kv0 + 64);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld32(t_lane + T_O + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) o[c * 32 + e] = o[c * 32 + e] * alpha + __uint_as_float(v[e]);
    }
    tc_fence_before();
  }

  const float inv_l = 1.f / l_run;
  bf16* orow = p.out + static_cast<size_t>(row_base + qrow) * (p.H * HD) + h * HD;
#pragma unroll
  for (int c8 = 0; c8 < 16; ++c8) {
    uint4 u;
    u.x = pack_bf16x2(o[c8 * 8 + 0] * inv_l, o[c8 * 8 + 1] * inv_l);
    u.y = pack_bf16x2(o[c8 * 8 + 2] * inv_l, o[c8 * 8 + 3] * inv_l);
    u.z = pack_bf16x2(o[c8 * 8 + 4] * inv_l, o[c8 * 8 + 5] * inv_l);
    u.w = pack_bf16x2(o[c8 * 8 + 6] * inv_l, o[c8 * 8 + 7] * inv_l);
    reinterpret_cast<uint4*>(orow)[c8] = u;
  }
  if (p.lse2) p.lse2[(static_cast<size_t>(b) * p.H + h) * p.S + qrow] = m_run + log2f(l_run);

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// ================================================================================================
// backward: delta = rowsum(dO * O)
// ================================================================================================
__global__ void attn_delta_kernel(const bf16* __restrict__ out, const bf16* __restrict__ dout, float* __restrict__ delta,
                                  int B, int S, int H) {
  // one warp per (token, head): 128 elements -> 4 per lane
  const long long gw = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long total = static_cast<long long>(B) * S * H;
  if (gw >= total) return;
  const int lane = threadIdx.x & 31;
  const int h = static_cast<int>(gw % H);
  const long long tok = gw / H;
  const size_t off = static_cast<size_t>(tok) * H * HD + static_cast<size_t>(h) * HD + lane * 4;
  const uint2 a = *reinterpret_cast<const uint2*>(out + off);
  const uint2 d = *reinterpret_cast<const uint2*>(dout + off);
  const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), d0 = unpack_bf16x2(d.x), d1 = unpack_bf16x2(d.y);
  float s = a0.x * d0.x + a0.y * d0.y + a1.x * d1.x + a1.y * d1.y;
  s = warp_sum(s);
  if (lane == 0) {
    const int b = static_cast<int>(tok / S), pos = static_cast<int>(tok % S);
    delta[(static_cast<size_t>(b) * H + h) * S + pos] = s;
  }
}

// ================================================================================================
// backward: dQ
// ================================================================================================
constexpr int DQ_SQ = 0;                  // 2 x [128 x 128B]
constexpr int DQ_SDO = 32768;             // 2 x [128 x 128B]
constexpr int DQ_SK = 65536;              // 2 x [64 x 128B]
constexpr int DQ_SV = DQ_SK + 16384;      // 2 x [64 x 128B], reused for dS [128 x 128B]
constexpr int DQ_BAR = DQ_SV + 16384;
constexpr int DQ_SMEM = DQ_BAR + 128 + 1024;

__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
               const __grid_constant__ CUtensorMap tmDO, const AttnKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DQ_BAR);
  uint64_t *bar_q = bars, *bar_kv = bars + 1, *bar_s = bars + 2, *bar_o = bars + 3;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 6);

  const int nqb = p.S / 128;
  const int qb = nqb - 1 - (blockIdx.x % nqb);
  const int bh = blockIdx.x / nqb;
  const int h = bh % p.H, b = bh / p.H;
  const int q0 = qb * 128;
  const int row_base = b * p.S;
  const int n_kv = (q0 + 128) / 64;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int colQ = h * HD, colK = p.H * HD + h * HD, colV = 2 * p.H * HD + h * HD;

  if (tid == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmDO);
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t t_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  const uint32_t T_S = 0, T_DP = 64, T_DQ = 128;

  auto load_kv = [&](int kv0) {
    mbar_arrive_expect_tx(bar_kv, 32768);
    tma_load_2d(smem + DQ_SK, &tmKV, bar_kv, colK, row_base + kv0);
    tma_load_2d(smem + DQ_SK + 8192, &tmKV, bar_kv, colK + 64, row_base + kv0);
    tma_load_2d(smem + DQ_SV, &tmKV, bar_kv, colV, row_base + kv0);
    tma_load_2d(smem + DQ_SV + 8192, &tmKV, bar_kv, colV + 64, row_base + kv0);
  };
  if (tid == 0) {
    mbar_arrive_expect_tx(bar_q, 65536);
    tma_load_2d(smem + DQ_SQ, &tmQ, bar_q, colQ, row_base + q0);
    tma_load_2d(smem + DQ_SQ + 16384, &tmQ, bar_q, colQ + 64, row_base + q0);
    tma_load_2d(smem + DQ_SDO, &tmDO, bar_q, h * HD, row_base + q0);
    tma_load_2d(smem + DQ_SDO + 16384, &tmDO, bar_q, h * HD + 64, row_base + q0);
    load_kv(0);
  }
  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);
  constexpr uint32_t idesc_dq = umma_idesc_bf16(128, 128, 0, 1);
  const uint32_t sQ = smem_u32(smem + DQ_SQ), sDO = smem_u32(smem + DQ_SDO), sK = smem_u32(smem + DQ_SK),
                 sV = smem_u32(smem + DQ_SV);
  const int qrow = q0 + tid;
  const size_t stat_idx = (static_cast<size_t>(b) * p.H + h) * p.S + qrow;
  const float lse2 = p.lse2[stat_idx];
  const float delta = p.delta[stat_idx];

  for (int j = 0; j < n_kv; ++j) {
    const uint32_t ph = j & 1;
    const int kv0 = j * 64;
    if (tid == 0) {
      if (j == 0) mbar_wait(bar_q, 0);
      mbar_wait(bar_kv, ph);
      tc_fence_after();
#pragma unroll
      for (int k16 = 0; k16 < 8; ++k16)
        umma_bf16(tmem + T_S, umma_desc_kmajor(kmaj_addr(sQ, k16, 16384)), umma_desc_kmajor(kmaj_addr(sK, k16, 8192)),
                  idesc_s, k16 > 0 ? 1u : 0u);
#pragma unroll
      for (int k16 = 0; k16 < 8; ++k16)
        umma_bf16(tmem + T_DP, umma_desc_kmajor(kmaj_addr(sDO, k16, 16384)), umma_desc_kmajor(kmaj_addr(sV, k16, 8192)),
                  idesc_s, k16 > 0 ? 1u : 0u);
      umma_commit(bar_s);
    }
    mbar_wait(bar_s, ph);
    tc_fence_after();
    const bool need_mask = (kv0 + 63 > q0);
    uint8_t* dsrow = smem + DQ_SV;  // V is dead once dP is complete
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t sv[32], dv[32];
      tmem_ld32(t_lane + T_S + half * 32, sv);
      tmem_ld32(t_lane + T_DP + half * 32, dv);
      tmem_ld_wait();
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        float ds[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = half * 32 + c8 * 8 + e;
          float pr = exp2f(__uint_as_float(sv[c8 * 8 + e]) * p.scale_log2 - lse2);
          if (need_mask && (kv0 + c > qrow)) pr = 0.f;
          ds[e] = pr * (__uint_as_float(dv[c8 * 8 + e]) - delta) * p.scale;
        }
        uint4 u;
        u.x = pack_bf16x2(ds[0], ds[1]); u.y = pack_bf16x2(ds[2], ds[3]);
        u.z = pack_bf16x2(ds[4], ds[5]); u.w = pack_bf16x2(ds[6], ds[7]);
        *reinterpret_cast<uint4*>(dsrow + sw128_offset(tid, half * 4 + c8)) = u;
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_bf16(tmem + T_DQ, umma_desc_kmajor(sV + kk * 32), umma_desc_mnmajor(sK + kk * 2048, 8192), idesc_dq,
                  (j > 0 || kk > 0) ? 1u : 0u);
      umma_commit(bar_o);
      mbar_wait(bar_o, ph);  // K and V/dS buffers free again
      if (j + 1 < n_kv) load_kv(kv0 + 64);
    }
  }
  // all threads: wait for the last dQ MMA
  mbar_wait(bar_o, (n_kv - 1) & 1);
  tc_fence_after();
  bf16* drow = p.dqkv + static_cast<size_t>(row_base + qrow) * (3 * p.H * HD) + colQ;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t v[32];
    tmem_ld32(t_lane + T_DQ + c * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) {
      uint4 u;
      u.x = pack_bf16x2(__uint_as_float(v[c8 * 8 + 0]), __uint_as_float(v[c8 * 8 + 1]));
      u.y = pack_bf16x2(__uint_as_float(v[c8 * 8 + 2]), __uint_as_float(v[c8 * 8 + 3]));
      u.z = pack_bf16x2(__uint_as_float(v[c8 * 8 + 4]), __uint_as_float(v[c8 * 8 + 5]));
      u.w = pack_bf16x2(__uint_as_float(v[c8 * 8 + 6]), __uint_as_float(v[c8 * 8 + 7]));
      reinterpret_cast<uint4*>(drow)[c * 4 + c8] = u;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// ================================================================================================
// backward: dK, dV
// ================================================================================================
constexpr int DKV_SK = 0;                   // 2 x [128 x 128B]
constexpr int DKV_SV = 32768;               // 2 x [128 x 128B]
constexpr int DKV_SQ = 65536;               // 2 x [64 x 128B]
constexpr int DKV_SDO = DKV_SQ + 16384;     // 2 x [64 x 128B]
constexpr int DKV_SPT = DKV_SDO + 16384;    // [128 x 128B]  P^T
constexpr int DKV_SDST = DKV_SPT + 16384;   // [128 x 128B]  dS^T
constexpr int DKV_STAT = DKV_SDST + 16384;  // lse2[64], delta[64]
constexpr int DKV_BAR = DKV_STAT + 512;
constexpr int DKV_SMEM = DKV_BAR + 128 + 1024;

__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_dkv_kernel(const __grid_constant__ CUtensorMap tmKV128, const __grid_constant__ CUtensorMap tmQ64,
                const __grid_constant__ CUtensorMap tmDO64, const AttnKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DKV_BAR);
  uint64_t *bar_kv = bars, *bar_q = bars + 1, *bar_s = bars + 2, *bar_o = bars + 3;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 6);
  float* s_lse = reinterpret_cast<float*>(smem + DKV_STAT);
  float* s_delta = s_lse + 64;

  const int nkb = p.S / 128;
  const int kb = blockIdx.x % nkb;  // early KV blocks see the most query blocks: they come first
  const int bh = blockIdx.x / nkb;
  const int h = bh % p.H, b = bh / p.H;
  const int kv0 = kb * 128;
  const int row_base = b * p.S;
  const int i0 = kv0 / 64, n_q = p.S / 64;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int colQ = h * HD, colK = p.H * HD + h * HD, colV = 2 * p.H * HD + h * HD;

  if (tid == 0) {
    tma_prefetch_desc(&tmKV128);
    tma_prefetch_desc(&tmQ64);
    tma_prefetch_desc(&tmDO64);
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t t_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  const uint32_t T_ST = 0, T_DPT = 64, T_DV = 128, T_DK = 256;

  auto load_q = [&](int qs) {
    mbar_arrive_expect_tx(bar_q, 32768);
    tma_load_2d(smem + DKV_SQ, &tmQ64, bar_q, colQ, row_base + qs);
    tma_load_2d(smem + DKV_SQ + 8192, &tmQ64, bar_q, colQ + 64, row_base + qs);
    tma_load_2d(smem + DKV_SDO, &tmDO64, bar_q, h * HD, row_base + qs);
    tma_load_2d(smem + DKV_SDO + 8192, &tmDO64, bar_q, h * HD + 64, row_base + qs);
  };
  if (tid == 0) {
    mbar_arrive_expect_tx(bar_kv, 65536);
    tma_load_2d(smem + DKV_SK, &tmKV128, bar_kv, colK, row_base + kv0);
    tma_load_2d(smem + DKV_SK + 16384, &tmKV128, bar_kv, colK + 64, row_base + kv0);
    tma_load_2d(smem + DKV_SV, &tmKV128, bar_kv, colV, row_base + kv0);
    tma_load_2d(smem + DKV_SV + 16384, &tmKV128, bar_kv, colV + 64, row_base + kv0);
    load_q(i0 * 64);
  }
  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);
  constexpr uint32_t idesc_g = umma_idesc_bf16(128, 128, 0, 1);
  const uint32_t sK = smem_u32(smem + DKV_SK), sV = smem_u32(smem + DKV_SV), sQ = smem_u32(smem + DKV_SQ),
                 sDO = smem_u32(smem + DKV_SDO), sPT = smem_u32(smem + DKV_SPT), sDST = smem_u32(smem + DKV_SDST);
  const int kvrow = kv0 + tid;
  const float* g_lse = p.lse2 + (static_cast<size_t>(b) * p.H + h) * p.S;
  const float* g_delta = p.delta + (static_cast<size_t>(b) * p.H + h) * p.S;

  for (int i = i0; i < n_q; ++i) {
    const uint32_t ph = (i - i0) & 1;
    const int qs = i * 64;
    if (tid < 64) {
      s_lse[tid] = g_lse[qs + tid];
      s_delta[tid] = g_delta[qs + tid];
    }
    if (tid == 0) {
      if (i == i0) mbar_wait(bar_kv, 0);
      mbar_wait(bar_q, ph);
      tc_fence_after();
#pragma unroll
      for (int k16 = 0; k16 < 8; ++k16)
        umma_bf16(tmem + T_ST, umma_desc_kmajor(kmaj_addr(sK, k16, 16384)), umma_desc_kmajor(kmaj_addr(sQ, k16, 8192)),
                  idesc_s, k16 > 0 ? 1u : 0u);
#pragma unroll
      for (int k16 = 0; k16 < 8; ++k16)
        umma_bf16(tmem + T_DPT, umma_desc_kmajor(kmaj_addr(sV, k16, 16384)), umma_desc_kmajor(kmaj_addr(sDO, k16, 8192)),
                  idesc_s, k16 > 0 ? 1u : 0u);
      umma_commit(bar_s);
    }
    __syncthreads();  // s_lse / s_delta visible
    mbar_wait(bar_s, ph);
    tc_fence_after();
    const bool need_mask = (qs < kv0 + 127);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t sv[32], dv[32];
      tmem_ld32(t_lane + T_ST + half * 32, sv);
      tmem_ld32(t_lane + T_DPT + half * 32, dv);
      tmem_ld_wait();
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        float pt[8], ds[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = half * 32 + c8 * 8 + e;  // query column
          float pr = exp2f(__uint_as_float(sv[c8 * 8 + e]) * p.scale_log2 - s_lse[c]);
          if (need_mask && (kvrow > qs + c)) pr = 0.f;
          pt[e] = pr;
          ds[e] = pr * (__uint_as_float(dv[c8 * 8 + e]) - s_delta[c]) * p.scale;
        }
        uint4 u, w;
        u.x = pack_bf16x2(pt[0], pt[1]); u.y = pack_bf16x2(pt[2], pt[3]);
        u.z = pack_bf16x2(pt[4], pt[5]); u.w = pack_bf16x2(pt[6], pt[7]);
        w.x = pack_bf16x2(ds[0], ds[1]); w.y = pack_bf16x2(ds[2], ds[3]);
        w.z = pack_bf16x2(ds[4], ds[5]); w.w = pack_bf16x2(ds[6], ds[7]);
        const uint32_t off = sw128_offset(tid, half * 4 + c8);
        *reinterpret_cast<uint4*>(smem + DKV_SPT + off) = u;
        *reinterpret_cast<uint4*>(smem + DKV_SDST + off) = w;
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t first = (i == i0) ? 0u : 1u;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_bf16(tmem + T_DV, umma_desc_kmajor(sPT + kk * 32), umma_desc_mnmajor(sDO + kk * 2048, 8192), idesc_g,
                  (first || kk > 0) ? 1u : 0u);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_bf16(tmem + T_DK, umma_desc_kmajor(sDST + kk * 32), umma_desc_mnmajor(sQ + kk * 2048, 8192), idesc_g,
                  (first || kk > 0) ? 1u : 0u);
      umma_commit(bar_o);
      mbar_wait(bar_o, ph);  // Q/dO/P^T/dS^T buffers free again
      if (i + 1 < n_q) load_q(qs + 64);
    }
  }
  mbar_wait(bar_o, (n_q - 1 - i0) & 1);
  tc_fence_after();
  bf16* dkrow = p.dqkv + static_cast<size_t>(row_base + kvrow) * (3 * p.H * HD) + colK;
  bf16* dvrow = p.dqkv + static_cast<size_t>(row_base + kvrow) * (3 * p.H * HD) + colV;
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    bf16* drow = which ? dkrow : dvrow;
    const uint32_t tcol = which ? T_DK : T_DV;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld32(t_lane + tcol + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(v[c8 * 8 + 0]), __uint_as_float(v[c8 * 8 + 1]));
        u.y = pack_bf16x2(__uint_as_float(v[c8 * 8 + 2]), __uint_as_float(v[c8 * 8 + 3]));
        u.z = pack_bf16x2(__uint_as_float(v[c8 * 8 + 4]), __uint_as_float(v[c8 * 8 + 5]));
        u.w = pack_bf16x2(__uint_as_float(v[c8 * 8 + 6]), __uint_as_float(v[c8 * 8 + 7]));
        reinterpret_cast<uint4*>(drow)[c * 4 + c8] = u;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

cudaError_t set_smem(const void* fn, int bytes) {
  return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

}  // namespace

cudaError_t attn_fwd(const AttnArgs& a, cudaStream_t s) {
  if (a.S % 128 || a.B <= 0 || a.H <= 0) return cudaErrorInvalidValue;
  static bool init = false;
  if (!init) {
    cudaError_t e = set_smem(reinterpret_cast<const void*>(attn_fwd_kernel), FWD_SMEM);
    if (e != cudaSuccess) return e;
    init = true;
  }
  const uint64_t M = static_cast<uint64_t>(a.B) * a.S, W = 3ull * a.H * HD;
  CUtensorMap tmQ, tmKV;
  if (!make_tmap_2d_bf16(&tmQ, a.qkv, W, M, W, 64, 128) || !make_tmap_2d_bf16(&tmKV, a.qkv, W, M, W, 64, 64))
    return cudaErrorInvalidValue;
  AttnKParams p{};
  p.B = a.B; p.S = a.S; p.H = a.H;
  p.scale = a.scale;
  p.scale_log2 = a.scale * LOG2E;
  p.lse2 = a.lse;
  p.out = a.out;
  attn_fwd_kernel<<<a.B * a.H * (a.S / 128), ATT_THREADS, FWD_SMEM, s>>>(tmQ, tmKV, p);
  return cudaGetLastError();
}

cudaError_t attn_bwd(const AttnArgs& a, cudaStream_t s) {
  if (a.S % 128 || a.B <= 0 || a.H <= 0 || !a.delta || !a.lse || !a.dout || !a.dqkv) return cudaErrorInvalidValue;
  static bool init = false;
  if (!init) {
    cudaError_t e = set_smem(reinterpret_cast<const void*>(attn_dq_kernel), DQ_SMEM);
    if (e != cudaSuccess) return e;
    e = set_smem(reinterpret_cast<const void*>(attn_dkv_kernel), DKV_SMEM);
    if (e != cudaSuccess) return e;
    init = true;
  }
  const uint64_t M = static_cast<uint64_t>(a.B) * a.S, W = 3ull * a.H * HD, WO = static_cast<uint64_t>(a.H) * HD;
  CUtensorMap tmQ128, tmKV64, tmDO128, tmKV128, tmQ64, tmDO64;
  bool ok = make_tmap_2d_bf16(&tmQ128, a.qkv, W, M, W, 64, 128) && make_tmap_2d_bf16(&tmKV64, a.qkv, W, M, W, 64, 64) &&
            make_tmap_2d_bf16(&tmDO128, a.dout, WO, M, WO, 64, 128) && make_tmap_2d_bf16(&tmDO64, a.dout, WO, M, WO, 64, 64);
  tmKV128 = tmQ128;
  tmQ64 = tmKV64;
  if (!ok) return cudaErrorInvalidValue;
  AttnKParams p{};
  p.B = a.B; p.S = a.S; p.H = a.H;
  p.scale = a.scale;
  p.scale_log2 = a.scale * LOG2E;
  p.lse2 = a.lse;
  p.out = a.out;
  p.delta = a.delta;
  p.dqkv = a.dqkv;
  {
    const long long warps = static_cast<long long>(a.B) * a.S * a.H;
    const int block = 256;
    const long long grid = (warps * 32 + block - 1) / block;
    attn_delta_kernel<<<static_cast<unsigned>(grid), block, 0, s>>>(a.out, a.dout, a.delta, a.B, a.S, a.H);
  }
  attn_dq_kernel<<<a.B * a.H * (a.S / 128), ATT_THREADS, DQ_SMEM, s>>>(tmQ128, tmKV64, tmDO128, p);
  attn_dkv_kernel<<<a.B * a.H * (a.S / 128), ATT_THREADS, DKV_SMEM, s>>>(tmKV128, tmQ64, tmDO64, p);
  return cudaGetLastError();
}

}  // namespace dtx
