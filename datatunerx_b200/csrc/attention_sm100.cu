// Causal flash attention forward / backward for sm_100a, head_dim 128, on tcgen05 + TMA.
//
// Replaces HF 4.34 eager LlamaAttention (matmul -> mask add -> fp32 softmax -> matmul, SURVEY §2.3 K6,
// reached from cmd/tuning/train.py:299) and its autograd backward (K10).  The [B,H,S,S] score tensor
// never exists: scores live in TMEM, probabilities go TMEM -> registers -> swizzled smem -> tensor core.
//
// Data layout: packed qkv [B*S, (H + 2*Hkv)*128] (per token: q heads | k heads | v heads; Hkv <= H for grouped-query
// attention), out [B*S, H*128],
// lse2 [B,H,S] = log2-domain log-sum-exp of the scaled scores (m + log2 l).
//
// Kernels (one CTA per SM; 4 compute warps where thread r owns TMEM lane r = one row of the score tile, plus a 5th
// warp whose lane 0 issues every TMA load and every tcgen05.mma, so no compute warp ever executes serial issue code):
//   attn_fwd_kernel   : CTA = 128 query rows, loops over 64-row KV blocks.  S = Q K^T (UMMA 128x64x16, K-major x K-major),
//                       online softmax in registers, P -> smem (K-major A operand), O_blk = P V (V is the MN-major B operand).
//   attn_dq_kernel    : CTA = 128 query rows.  S, dP = dO V^T, dS = P o (dP - delta) * scale, dQ += dS K (K as MN-major B),
//                       dQ accumulates in TMEM over the whole KV loop.
//   attn_dkv_kernel   : CTA = 128 KV rows, loops over 64-row Q blocks.  S^T = K Q^T, dP^T = V dO^T, dV += P^T dO, dK += dS^T Q.
// All three are software-pipelined inside the CTA (r01 v2; v1 ran its phases back to back and left the tensor pipe 18-30%
// busy, profiles/r01_ncu_full_baseline.txt): the streamed operand blocks sit in a 3-deep TMA ring, the score MMAs of block
// j+1 are issued into a second TMEM buffer BEFORE the threads start the exp / dS arithmetic of block j, and the
// accumulating MMAs of block j run under the TMEM loads of block j+1.  mbarriers carry every hand-off (bar_kv/bar_q: TMA
// landed; bar_s: scores ready; bar_p: the 128 compute threads have written P / dS to smem; bar_o: accumulate MMA done).
// Two backward kernels instead of one with fp32 atomics on dQ: every reduction has a fixed order, so the
// step is bitwise reproducible (needed for the N-rank == 1-rank parity tests).
#include "common.cuh"
#include "kernels.h"

namespace dtx {

bool make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                       uint32_t box_outer);

namespace {

constexpr int HD = 128;      // head dim
constexpr int BWD_THREADS = 288;  // backward kernels: 8 compute warps (two per score row quadrant, splitting the columns) + issuer warp
constexpr int ATT_THREADS = 160;  // 4 compute warps (one TMEM lane / score row per thread) + 1 issuer warp (TMA + MMA)
constexpr float LOG2E = 1.4426950408889634f;

struct AttnKParams {
  int B, S, H;
  int Hkv, W;        // kv heads (grouped-query attention: q head h reads kv head h / (H / Hkv)); W = row stride of qkv
  int colK0, colV0;  // first column of the k / v sections inside a qkv row
  float scale_log2;  // softmax scale * log2(e)
  float scale;
  float* lse2;
  bf16* out;
  const float* delta;
  bf16* dqkv;
  const float2* rope_cs;  // [S][64] (cos, sin) or null
};

// issue `n16` UMMAs (K=16 each) walking both operands.  a/b step = byte advance per k16.
// K-major operands made of 64-wide subtiles: k16 -> subtile (k16/4) + 32 B * (k16%4).
__device__ __forceinline__ uint32_t kmaj_addr(uint32_t base, int k16, uint32_t subtile_bytes) {
  return base + (k16 >> 2) * subtile_bytes + (k16 & 3) * 32;
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ================================================================================================
// forward
// ================================================================================================
constexpr int FWD_SQ = 0;                    // 2 x [128 x 128B]
constexpr int FWD_SK = 32768;                // 3 x (2 x [64 x 128B])
constexpr int FWD_SV = FWD_SK + 3 * 16384;   // 3 x (2 x [64 x 128B])
constexpr int FWD_SP = FWD_SV + 3 * 16384;   // [128 x 128B]
constexpr int FWD_BAR = FWD_SP + 16384;
constexpr int FWD_SMEM = FWD_BAR + 256 + 1024;

__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const AttnKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FWD_BAR);
  uint64_t *bar_q = bars, *bar_kv = bars + 1 /*[3]*/, *bar_s = bars + 4 /*[2]*/, *bar_o = bars + 6, *bar_p = bars + 7;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);

  const int nqb = p.S / 128;
  const int qb = nqb - 1 - (blockIdx.x % nqb);  // heavy (late) query blocks first
  const int bh = blockIdx.x / nqb;
  const int h = bh % p.H, b = bh / p.H;
  const int q0 = qb * 128;
  const int row_base = b * p.S;
  const int n = (q0 + 128) / 64;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int hk = h / (p.H / p.Hkv);
  const int colQ = h * HD, colK = p.colK0 + hk * HD, colV = p.colV0 + hk * HD;

  if (tid == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    for (int i = 0; i < 7; ++i) mbar_init(&bars[i], 1);
    mbar_init(bar_p, 128);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t T_S = 0 /* +64*buf */, T_O = 128;

  if (warp == 4) {
    // ------------------------------------------ issuer ------------------------------------------
    {
      const bool leader = elect_one();
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 128, 0, 1);
      const uint32_t sQ = smem_u32(smem + FWD_SQ), sK = smem_u32(smem + FWD_SK), sV = smem_u32(smem + FWD_SV),
                     sP = smem_u32(smem + FWD_SP);
      auto load_kv = [&](int j) {  // KV block j -> ring slot j % 3
        const int slot = j % 3;
        if (!leader) return;
        mbar_arrive_expect_tx(&bar_kv[slot], 32768);
        tma_load_2d(smem + FWD_SK + slot * 16384, &tmKV, &bar_kv[slot], colK, row_base + j * 64);
        tma_load_2d(smem + FWD_SK + slot * 16384 + 8192, &tmKV, &bar_kv[slot], colK + 64, row_base + j * 64);
        tma_load_2d(smem + FWD_SV + slot * 16384, &tmKV, &bar_kv[slot], colV, row_base + j * 64);
        tma_load_2d(smem + FWD_SV + slot * 16384 + 8192, &tmKV, &bar_kv[slot], colV + 64, row_base + j * 64);
      };
      auto issue_s = [&](int j) {  // S(j) = Q K(j)^T into score buffer j & 1
        const int slot = j % 3;
        mbar_wait_backoff(&bar_kv[slot], (j / 3) & 1);
        tc_fence_after();
#pragma unroll
        for (int k16 = 0; k16 < 8; ++k16)
          if (leader) umma_bf16(tmem + T_S + (j & 1) * 64, umma_desc_kmajor(kmaj_addr(sQ, k16, 16384)),
                    umma_desc_kmajor(kmaj_addr(sK + slot * 16384, k16, 8192)), idesc_s, k16 > 0 ? 1u : 0u);
        if (leader) umma_commit(&bar_s[j & 1]);
      };
      if (leader) {
        mbar_arrive_expect_tx(bar_q, 32768);
        tma_load_2d(smem + FWD_SQ, &tmQ, bar_q, colQ, row_base + q0);
        tma_load_2d(smem + FWD_SQ + 16384, &tmQ, bar_q, colQ + 64, row_base + q0);
      }
      for (int j = 0; j < 3 && j < n; ++j) load_kv(j);
      mbar_wait_backoff(bar_q, 0);
      issue_s(0);
      if (n > 1) issue_s(1);
      for (int j = 0; j < n; ++j) {
        mbar_wait_backoff(bar_p, j & 1);  // P(j) is in smem; score buffer j&1 and the O tile have been consumed
        tc_fence_after();
        const uint32_t vb = sV + (j % 3) * 16384;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          if (leader) umma_bf16(tmem + T_O, umma_desc_kmajor(sP + kk * 32), umma_desc_mnmajor(vb + kk * 2048, 8192), idesc_o, kk > 0 ? 1u : 0u);
        if (leader) umma_commit(bar_o);
        if (j + 2 < n) issue_s(j + 2);
        if (j + 3 < n) {
          mbar_wait_backoff(bar_o, j & 1);  // P V(j) done: ring slot j % 3 is free
          load_kv(j + 3);
        }
      }
    }
  } else {
    // ------------------------------------------ softmax warps ------------------------------------------
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    const uint32_t sP_addr = smem_u32(smem + FWD_SP);
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;
    float o[HD];
#pragma unroll
    for (int i = 0; i < HD; ++i) o[i] = 0.f;
    const int qrow = q0 + tid;

    for (int j = 0; j < n; ++j) {
      const int kv0 = j * 64;
      mbar_wait(&bar_s[j & 1], (j >> 1) & 1);
      tc_fence_after();
      uint32_t sv[64];
      {
        uint32_t(&lo)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[0]);
        uint32_t(&hi)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[32]);
        tmem_ld32(t_lane + T_S + (j & 1) * 64, lo);
        tmem_ld32(t_lane + T_S + (j & 1) * 64 + 32, hi);
        tmem_ld_wait();
      }
      if (kv0 + 63 > q0) {  // diagonal blocks: causal mask
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if (kv0 + c > qrow) sv[c] = 0xff800000u;  // -inf
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int c = 0; c < 64; c += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(sv[c]));
        mx1 = fmaxf(mx1, __uint_as_float(sv[c + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(sv[c + 2]));
        mx3 = fmaxf(mx3, __uint_as_float(sv[c + 3]));
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * p.scale_log2;  // scale > 0: max commutes with the scaling
      const float m_new = fmaxf(m_run, mx);
      const float alpha = fast_exp2(m_run - m_new);
      uint32_t pk[32];
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int c = 0; c < 64; c += 2) {
        const float p0 = fast_exp2(fmaf(__uint_as_float(sv[c]), p.scale_log2, -m_new));
        const float p1 = fast_exp2(fmaf(__uint_as_float(sv[c + 1]), p.scale_log2, -m_new));
        rs0 += p0;
        rs1 += p1;
        pk[c >> 1] = pack_bf16x2(p0, p1);
      }
      l_run = l_run * alpha + (rs0 + rs1);
      m_run = m_new;
      if (j > 0) {  // fold the previous block's P V (finished under the work above) into the running output
        mbar_wait(bar_o, (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld32(t_lane + T_O + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) o[c * 32 + e] = fmaf(o[c * 32 + e], alpha_prev, __uint_as_float(v[e]));
        }
      }
      alpha_prev = alpha;
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8)
        sts128(sP_addr + sw128_offset(tid, c8), pk[4 * c8], pk[4 * c8 + 1], pk[4 * c8 + 2], pk[4 * c8 + 3]);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_p);
    }
    mbar_wait(bar_o, (n - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.f / l_run;
    bf16* orow = p.out + static_cast<size_t>(row_base + qrow) * (p.H * HD) + h * HD;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld32(t_lane + T_O + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = fmaf(o[c * 32 + c8 * 8 + e], alpha_prev, __uint_as_float(v[c8 * 8 + e])) * inv_l;
        uint4 u;
        u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
        u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
        reinterpret_cast<uint4*>(orow)[c * 4 + c8] = u;
      }
    }
    if (p.lse2) p.lse2[(static_cast<size_t>(b) * p.H + h) * p.S + qrow] = m_run + log2f(l_run);
  }

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
// backward: dQ — two 128-row query tiles per CTA (groups) ping-ponging on one tensor pipe
// ================================================================================================
// Group g owns query rows [256*pair + 128*g, +128): while its four warps turn (S, dP) into dS, the tensor core works for the
// other group (dQ accumulate, next S / dP).  Both groups read the same K/V ring, so each K/V block is staged once per 256
// query rows.  The single-tile version of this kernel left the tensor pipe 25% busy (profiles/r01_*attn*).
constexpr int DQ_SQ = 0;                    // 2 groups x (2 x [128 x 128B])
constexpr int DQ_SDO = 65536;               // 2 groups x (2 x [128 x 128B])
constexpr int DQ_SK = 131072;               // 2 slots x (2 x [64 x 128B])
constexpr int DQ_SV = DQ_SK + 2 * 16384;    // 2 slots x (2 x [64 x 128B])
constexpr int DQ_SDS = DQ_SV + 2 * 16384;   // 2 groups x [128 x 128B]
constexpr int DQ_BAR = DQ_SDS + 2 * 16384;
constexpr int DQ_SMEM = DQ_BAR + 256 + 1024;

__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
               const __grid_constant__ CUtensorMap tmDO, const AttnKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DQ_BAR);
  uint64_t *bar_q = bars, *bar_kv = bars + 1 /*[2]*/, *bar_s = bars + 3 /*[g]*/, *bar_o = bars + 5 /*[g]*/, *bar_p = bars + 7 /*[g]*/;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 10);

  const int npair = (p.S + 255) / 256;
  const int qp = npair - 1 - (blockIdx.x % npair);
  const int bh = blockIdx.x / npair;
  const int h = bh % p.H, b = bh / p.H;
  const int row_base = b * p.S;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int hk = h / (p.H / p.Hkv);
  const int colQ = h * HD, colK = p.colK0 + hk * HD, colV = p.colV0 + hk * HD;
  const int q00 = qp * 256, q01 = qp * 256 + 128;
  const int ng0 = (q00 + 128) / 64, ng1 = q01 < p.S ? (q01 + 128) / 64 : 0;  // KV blocks each group needs
  const int n = max(ng0, ng1);
  auto ngf = [&](int g) { return g ? ng1 : ng0; };

  if (tid == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmDO);
    for (int i = 0; i < 7; ++i) mbar_init(&bars[i], 1);
    mbar_init(&bar_p[0], 128);
    mbar_init(&bar_p[1], 128);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  // TMEM columns of group g: S 256g, dP 256g + 64, dQ 256g + 128

  if (warp == 8) {
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t idesc_dq = umma_idesc_bf16(128, 128, 0, 1);
    const uint32_t sQ = smem_u32(smem + DQ_SQ), sDO = smem_u32(smem + DQ_SDO), sK = smem_u32(smem + DQ_SK),
                   sV = smem_u32(smem + DQ_SV), sDS = smem_u32(smem + DQ_SDS);
    auto load_kv = [&](int j) {
      const int slot = j & 1;
      if (!leader) return;
      mbar_arrive_expect_tx(&bar_kv[slot], 32768);
      tma_load_2d(smem + DQ_SK + slot * 16384, &tmKV, &bar_kv[slot], colK, row_base + j * 64);
      tma_load_2d(smem + DQ_SK + slot * 16384 + 8192, &tmKV, &bar_kv[slot], colK + 64, row_base + j * 64);
      tma_load_2d(smem + DQ_SV + slot * 16384, &tmKV, &bar_kv[slot], colV, row_base + j * 64);
      tma_load_2d(smem + DQ_SV + slot * 16384 + 8192, &tmKV, &bar_kv[slot], colV + 64, row_base + j * 64);
    };
    auto issue_s = [&](int g, int j) {  // S_g(j) = Q_g K(j)^T and dP_g(j) = dO_g V(j)^T
      const int slot = j & 1;
      mbar_wait_backoff(&bar_kv[slot], (j >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int k16 = 0; k16 < 8; ++k16)
        if (leader) umma_bf16(tmem + g * 256, umma_desc_kmajor(kmaj_addr(sQ + g * 32768, k16, 16384)),
                              umma_desc_kmajor(kmaj_addr(sK + slot * 16384, k16, 8192)), idesc_s, k16 > 0 ? 1u : 0u);
#pragma unroll
      for (int k16 = 0; k16 < 8; ++k16)
        if (leader) umma_bf16(tmem + g * 256 + 64, umma_desc_kmajor(kmaj_addr(sDO + g * 32768, k16, 16384)),
                              umma_desc_kmajor(kmaj_addr(sV + slot * 16384, k16, 8192)), idesc_s, k16 > 0 ? 1u : 0u);
      if (leader) umma_commit(&bar_s[g]);
    };
    if (leader) {
      mbar_arrive_expect_tx(bar_q, ng1 > 0 ? 131072 : 65536);
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (ngf(g) == 0) continue;
        const int q0 = g ? q01 : q00;
        tma_load_2d(smem + DQ_SQ + g * 32768, &tmQ, bar_q, colQ, row_base + q0);
        tma_load_2d(smem + DQ_SQ + g * 32768 + 16384, &tmQ, bar_q, colQ + 64, row_base + q0);
        tma_load_2d(smem + DQ_SDO + g * 32768, &tmDO, bar_q, h * HD, row_base + q0);
        tma_load_2d(smem + DQ_SDO + g * 32768 + 16384, &tmDO, bar_q, h * HD + 64, row_base + q0);
      }
    }
    load_kv(0);
    if (n > 1) load_kv(1);
    mbar_wait_backoff(bar_q, 0);
    issue_s(0, 0);
    if (ng1 > 0) issue_s(1, 0);
    for (int j = 0; j < n; ++j) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (j >= ngf(g)) continue;
        mbar_wait_backoff(&bar_p[g], j & 1);  // dS_g(j) is in smem; group g's S / dP tiles have been consumed
        tc_fence_after();
        const uint32_t kb = sK + (j & 1) * 16384;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          if (leader) umma_bf16(tmem + g * 256 + 128, umma_desc_kmajor(sDS + g * 16384 + kk * 32), umma_desc_mnmajor(kb + kk * 2048, 8192),
                                idesc_dq, (j > 0 || kk > 0) ? 1u : 0u);
        if (leader) umma_commit(&bar_o[g]);
        if (j + 1 < ngf(g)) issue_s(g, j + 1);
      }
      if (j + 2 < n) {  // ring slot j & 1 is free once both groups' dQ MMAs of block j have completed
#pragma unroll
        for (int g = 0; g < 2; ++g)
          if (j < ngf(g)) mbar_wait_backoff(&bar_o[g], j & 1);
        load_kv(j + 2);
      }
    }
  } else {
    const int g = warp >> 2, w = warp & 3;
    const int r = w * 32 + (tid & 31);
    const int n_mine = ngf(g);
    const int q0 = g ? q01 : q00;
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(w * 32) << 16) + g * 256;
    const uint32_t T_S = 0, T_DP = 64, T_DQ = 128;
    const int qrow = q0 + r;
    const uint32_t sDS_addr = smem_u32(smem + DQ_SDS + g * 16384);
    uint64_t *my_s = bar_s + g, *my_o = bar_o + g, *my_p = bar_p + g;
    float lse2 = 0.f, delta = 0.f;
    if (n_mine > 0) {
      const size_t stat_idx = (static_cast<size_t>(b) * p.H + h) * p.S + qrow;
      lse2 = p.lse2[stat_idx];
      delta = p.delta[stat_idx];
    }
    for (int j = 0; j < n_mine; ++j) {
      const int kv0 = j * 64;
      mbar_wait(my_s, j & 1);
      tc_fence_after();
      const bool need_mask = (kv0 + 63 > q0);
      uint32_t pk[32];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t sv[32], dv[32];
        tmem_ld32(t_lane + T_S + half * 32, sv);
        tmem_ld32(t_lane + T_DP + half * 32, dv);
        tmem_ld_wait();
        if (need_mask) {
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            float p0 = fast_exp2(fmaf(__uint_as_float(sv[e]), p.scale_log2, -lse2));
            float p1 = fast_exp2(fmaf(__uint_as_float(sv[e + 1]), p.scale_log2, -lse2));
            if (kv0 + half * 32 + e > qrow) p0 = 0.f;
            if (kv0 + half * 32 + e + 1 > qrow) p1 = 0.f;
            pk[half * 16 + (e >> 1)] = pack_bf16x2(p0 * (__uint_as_float(dv[e]) - delta) * p.scale,
                                                   p1 * (__uint_as_float(dv[e + 1]) - delta) * p.scale);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float p0 = fast_exp2(fmaf(__uint_as_float(sv[e]), p.scale_log2, -lse2));
            const float p1 = fast_exp2(fmaf(__uint_as_float(sv[e + 1]), p.scale_log2, -lse2));
            pk[half * 16 + (e >> 1)] = pack_bf16x2(p0 * (__uint_as_float(dv[e]) - delta) * p.scale,
                                                   p1 * (__uint_as_float(dv[e + 1]) - delta) * p.scale);
          }
        }
      }
      if (j > 0) mbar_wait(my_o, (j - 1) & 1);  // dQ MMA of block j-1 done: the dS buffer is free
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8)
        sts128(sDS_addr + sw128_offset(r, c8), pk[4 * c8], pk[4 * c8 + 1], pk[4 * c8 + 2], pk[4 * c8 + 3]);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(my_p);
    }
    if (n_mine > 0) {
      mbar_wait(my_o, (n_mine - 1) & 1);
      tc_fence_after();
      bf16* drow = p.dqkv + static_cast<size_t>(row_base + qrow) * p.W + colQ;
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
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ================================================================================================
// backward: dK, dV
// ================================================================================================
constexpr int DKV_SK = 0;                      // 2 x [128 x 128B]
constexpr int DKV_SV = 32768;                  // 2 x [128 x 128B]
constexpr int DKV_SQ = 65536;                  // 3 x (2 x [64 x 128B])
constexpr int DKV_SDO = DKV_SQ + 3 * 16384;    // 3 x (2 x [64 x 128B])
constexpr int DKV_SPT = DKV_SDO + 3 * 16384;   // [128 x 128B]  P^T
constexpr int DKV_SDST = DKV_SPT + 16384;      // [128 x 128B]  dS^T
constexpr int DKV_STAT = DKV_SDST + 16384;     // 3 x (lse2[64], delta[64]) fp32, filled by bulk copies with the Q ring
constexpr int DKV_BAR = DKV_STAT + 3 * 512;
constexpr int DKV_SMEM = DKV_BAR + 256 + 1024;

__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_dkv_kernel(const __grid_constant__ CUtensorMap tmKV128, const __grid_constant__ CUtensorMap tmQ64,
                const __grid_constant__ CUtensorMap tmDO64, const AttnKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DKV_BAR);
  uint64_t *bar_kv = bars, *bar_q = bars + 1 /*[3]*/, *bar_s = bars + 4 /*[2]*/, *bar_o = bars + 6, *bar_p = bars + 7;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);

  const int nkb = p.S / 128;
  const int kb = blockIdx.x % nkb;  // early KV blocks see the most query blocks: they come first
  const int bh = blockIdx.x / nkb;
  const int hk = bh % p.Hkv, b = bh / p.Hkv;  // one CTA per (batch, KV head, 128-row KV block)
  const int grp = p.H / p.Hkv;                // query heads sharing this KV head (1 = multi-head attention)
  const int kv0 = kb * 128;
  const int row_base = b * p.S;
  const int i0 = kv0 / 64;
  const int nq = p.S / 64 - i0;               // query blocks per query head that see this KV block
  const int n = nq * grp;                     // streamed (head, query block) pairs; it -> head it / nq, block it % nq
  const int tid = threadIdx.x, warp = tid >> 5;
  const int colK = p.colK0 + hk * HD, colV = p.colV0 + hk * HD;
  const float* g_lse = p.lse2 + (static_cast<size_t>(b) * p.H + hk * grp) * p.S;
  const float* g_delta = p.delta + (static_cast<size_t>(b) * p.H + hk * grp) * p.S;

  if (tid == 0) {
    tma_prefetch_desc(&tmKV128);
    tma_prefetch_desc(&tmQ64);
    tma_prefetch_desc(&tmDO64);
    for (int i = 0; i < 7; ++i) mbar_init(&bars[i], 1);
    mbar_init(bar_p, 256);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t T_ST = 0 /* +64*buf */, T_DPT = 128 /* +64*buf */, T_DV = 256, T_DK = 384;

  if (warp == 8) {
    {
      const bool leader = elect_one();
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);
      constexpr uint32_t idesc_g = umma_idesc_bf16(128, 128, 0, 1);
      const uint32_t sK = smem_u32(smem + DKV_SK), sV = smem_u32(smem + DKV_SV), sQ = smem_u32(smem + DKV_SQ),
                     sDO = smem_u32(smem + DKV_SDO), sPT = smem_u32(smem + DKV_SPT), sDST = smem_u32(smem + DKV_SDST);
      auto load_q = [&](int ii) {  // (head, query block) pair ii: Q, dO, lse2, delta -> ring slot ii % 3
        const int slot = ii % 3, hq = ii / nq, qs = (i0 + ii % nq) * 64;
        const int colQ = (hk * grp + hq) * HD;
        if (!leader) return;
        mbar_arrive_expect_tx(&bar_q[slot], 32768 + 512);
        tma_load_2d(smem + DKV_SQ + slot * 16384, &tmQ64, &bar_q[slot], colQ, row_base + qs);
        tma_load_2d(smem + DKV_SQ + slot * 16384 + 8192, &tmQ64, &bar_q[slot], colQ + 64, row_base + qs);
        tma_load_2d(smem + DKV_SDO + slot * 16384, &tmDO64, &bar_q[slot], colQ, row_base + qs);
        tma_load_2d(smem + DKV_SDO + slot * 16384 + 8192, &tmDO64, &bar_q[slot], colQ + 64, row_base + qs);
        tma_load_1d(smem + DKV_STAT + slot * 512, g_lse + static_cast<size_t>(hq) * p.S + qs, 256, &bar_q[slot]);
        tma_load_1d(smem + DKV_STAT + slot * 512 + 256, g_delta + static_cast<size_t>(hq) * p.S + qs, 256, &bar_q[slot]);
      };
      auto issue_s = [&](int ii) {  // S^T = K Q^T, dP^T = V dO^T into buffer ii & 1
        const int slot = ii % 3;
        mbar_wait_backoff(&bar_q[slot], (ii / 3) & 1);
        tc_fence_after();
#pragma unroll
        for (int k16 = 0; k16 < 8; ++k16)
          if (leader) umma_bf16(tmem + T_ST + (ii & 1) * 64, umma_desc_kmajor(kmaj_addr(sK, k16, 16384)),
                    umma_desc_kmajor(kmaj_addr(sQ + slot * 16384, k16, 8192)), idesc_s, k16 > 0 ? 1u : 0u);
#pragma unroll
        for (int k16 = 0; k16 < 8; ++k16)
          if (leader) umma_bf16(tmem + T_DPT + (ii & 1) * 64, umma_desc_kmajor(kmaj_addr(sV, k16, 16384)),
                    umma_desc_kmajor(kmaj_addr(sDO + slot * 16384, k16, 8192)), idesc_s, k16 > 0 ? 1u : 0u);
        if (leader) umma_commit(&bar_s[ii & 1]);
      };
      if (leader) {
        mbar_arrive_expect_tx(bar_kv, 65536);
        tma_load_2d(smem + DKV_SK, &tmKV128, bar_kv, colK, row_base + kv0);
        tma_load_2d(smem + DKV_SK + 16384, &tmKV128, bar_kv, colK + 64, row_base + kv0);
        tma_load_2d(smem + DKV_SV, &tmKV128, bar_kv, colV, row_base + kv0);
        tma_load_2d(smem + DKV_SV + 16384, &tmKV128, bar_kv, colV + 64, row_base + kv0);
      }
      for (int ii = 0; ii < 3 && ii < n; ++ii) load_q(ii);
      mbar_wait_backoff(bar_kv, 0);
      issue_s(0);
      if (n > 1) issue_s(1);
      for (int ii = 0; ii < n; ++ii) {
        mbar_wait_backoff(bar_p, ii & 1);  // P^T / dS^T of block ii are in smem; score buffers ii&1 consumed
        tc_fence_after();
        const uint32_t first = (ii == 0) ? 0u : 1u;
        const uint32_t qb_ = sQ + (ii % 3) * 16384, dob = sDO + (ii % 3) * 16384;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          if (leader) umma_bf16(tmem + T_DV, umma_desc_kmajor(sPT + kk * 32), umma_desc_mnmajor(dob + kk * 2048, 8192), idesc_g,
                    (first || kk > 0) ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          if (leader) umma_bf16(tmem + T_DK, umma_desc_kmajor(sDST + kk * 32), umma_desc_mnmajor(qb_ + kk * 2048, 8192), idesc_g,
                    (first || kk > 0) ? 1u : 0u);
        if (leader) umma_commit(bar_o);
        if (ii + 2 < n) issue_s(ii + 2);
        if (ii + 3 < n) {
          mbar_wait_backoff(bar_o, ii & 1);  // dV/dK MMAs(ii) done: ring slot ii % 3 is free
          load_q(ii + 3);
        }
      }
    }
  } else {
    // 8 compute warps: warps w and w+4 share the TMEM lanes (kv rows) 32*(w&3)..+31 and split the 64 query columns
    const int rw = warp & 3, half = warp >> 2;
    const int r = rw * 32 + (tid & 31);
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(rw * 32) << 16);
    const int kvrow = kv0 + r;
    const uint32_t sPT_addr = smem_u32(smem + DKV_SPT), sDST_addr = smem_u32(smem + DKV_SDST);
    for (int ii = 0; ii < n; ++ii) {
      const int qs = (i0 + ii % nq) * 64;
      mbar_wait(&bar_q[ii % 3], (ii / 3) & 1);  // acquire the TMA-written row statistics of this ring slot
      mbar_wait(&bar_s[ii & 1], (ii >> 1) & 1);
      tc_fence_after();
      const uint32_t st = smem_u32(smem + DKV_STAT + (ii % 3) * 512) + half * 128;
      uint32_t sv[32], dv[32], ppk[16], dpk[16];
      tmem_ld32(t_lane + T_ST + (ii & 1) * 64 + half * 32, sv);
      tmem_ld32(t_lane + T_DPT + (ii & 1) * 64 + half * 32, dv);
      tmem_ld_wait();
      const bool need_mask = (qs < kv0 + 127);
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        const float4 l4 = lds128f(st + e * 4);
        const float4 d4 = lds128f(st + 256 + e * 4);
        const float ls[4] = {l4.x, l4.y, l4.z, l4.w}, ds4[4] = {d4.x, d4.y, d4.z, d4.w};
        float pr[4], dsv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          pr[u] = fast_exp2(fmaf(__uint_as_float(sv[e + u]), p.scale_log2, -ls[u]));
          if (need_mask && (kvrow > qs + half * 32 + e + u)) pr[u] = 0.f;
          dsv[u] = pr[u] * (__uint_as_float(dv[e + u]) - ds4[u]) * p.scale;
        }
        ppk[e >> 1] = pack_bf16x2(pr[0], pr[1]);
        ppk[(e >> 1) + 1] = pack_bf16x2(pr[2], pr[3]);
        dpk[e >> 1] = pack_bf16x2(dsv[0], dsv[1]);
        dpk[(e >> 1) + 1] = pack_bf16x2(dsv[2], dsv[3]);
      }
      if (ii > 0) mbar_wait(bar_o, (ii - 1) & 1);  // dV/dK MMAs of block ii-1 done: P^T/dS^T buffers are free
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        const uint32_t off = sw128_offset(r, half * 4 + c8);
        sts128(sPT_addr + off, ppk[4 * c8], ppk[4 * c8 + 1], ppk[4 * c8 + 2], ppk[4 * c8 + 3]);
        sts128(sDST_addr + off, dpk[4 * c8], dpk[4 * c8 + 1], dpk[4 * c8 + 2], dpk[4 * c8 + 3]);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_p);
    }
    mbar_wait(bar_o, (n - 1) & 1);
    tc_fence_after();
    // warps 0-3 write dV, warps 4-7 write dK (each thread one full 128-wide row)
    bf16* drow = p.dqkv + static_cast<size_t>(row_base + kvrow) * p.W + (half ? colK : colV);
    const uint32_t tcol = half ? T_DK : T_DV;
    const bool rot = false;  // (inverse rotary of dK here was measured slower than the separate HBM-bound kernel)
    const float2* cs = nullptr;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t lo[32], hi[32];
      tmem_ld32(t_lane + tcol + c * 32, lo);
      tmem_ld32(t_lane + tcol + 64 + c * 32, hi);
      tmem_ld_wait();
      if (rot) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float2 t = __ldg(cs + c * 32 + j);
          const float x0 = __uint_as_float(lo[j]), x1 = __uint_as_float(hi[j]);
          lo[j] = __float_as_uint(x0 * t.x + x1 * t.y);
          hi[j] = __float_as_uint(x1 * t.x - x0 * t.y);
        }
      }
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        uint4 u, w;
        u.x = pack_bf16x2(__uint_as_float(lo[c8 * 8 + 0]), __uint_as_float(lo[c8 * 8 + 1]));
        u.y = pack_bf16x2(__uint_as_float(lo[c8 * 8 + 2]), __uint_as_float(lo[c8 * 8 + 3]));
        u.z = pack_bf16x2(__uint_as_float(lo[c8 * 8 + 4]), __uint_as_float(lo[c8 * 8 + 5]));
        u.w = pack_bf16x2(__uint_as_float(lo[c8 * 8 + 6]), __uint_as_float(lo[c8 * 8 + 7]));
        w.x = pack_bf16x2(__uint_as_float(hi[c8 * 8 + 0]), __uint_as_float(hi[c8 * 8 + 1]));
        w.y = pack_bf16x2(__uint_as_float(hi[c8 * 8 + 2]), __uint_as_float(hi[c8 * 8 + 3]));
        w.z = pack_bf16x2(__uint_as_float(hi[c8 * 8 + 4]), __uint_as_float(hi[c8 * 8 + 5]));
        w.w = pack_bf16x2(__uint_as_float(hi[c8 * 8 + 6]), __uint_as_float(hi[c8 * 8 + 7]));
        reinterpret_cast<uint4*>(drow + c * 32)[c8] = u;
        reinterpret_cast<uint4*>(drow + 64 + c * 32)[c8] = w;
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
  const int Hkv = a.Hkv > 0 ? a.Hkv : a.H;
  if (a.H % Hkv) return cudaErrorInvalidValue;
  const uint64_t M = static_cast<uint64_t>(a.B) * a.S, W = static_cast<uint64_t>(a.H + 2 * Hkv) * HD;
  CUtensorMap tmQ, tmKV;
  if (!make_tmap_2d_bf16(&tmQ, a.qkv, W, M, W, 64, 128) || !make_tmap_2d_bf16(&tmKV, a.qkv, W, M, W, 64, 64))
    return cudaErrorInvalidValue;
  AttnKParams p{};
  p.B = a.B; p.S = a.S; p.H = a.H;
  p.Hkv = Hkv; p.W = static_cast<int>(W); p.colK0 = a.H * HD; p.colV0 = (a.H + Hkv) * HD;
  p.scale = a.scale;
  p.scale_log2 = a.scale * LOG2E;
  p.lse2 = a.lse;
  p.out = a.out;
  attn_fwd_kernel<<<a.B * a.H * (a.S / 128), ATT_THREADS, FWD_SMEM, s>>>(tmQ, tmKV, p);
  return cudaGetLastError();
}

cudaError_t attn_bwd(const AttnArgs& a, cudaStream_t s) {
  if (a.S % 128 || a.B <= 0 || a.H <= 0 || !a.delta || !a.lse || !a.dout || !a.dqkv) return cudaErrorInvalidValue;
  if (a.rope_cs) return cudaErrorInvalidValue;  // inverse rotary in the store epilogues was measured slower than the separate kernel
  static bool init = false;
  if (!init) {
    cudaError_t e = set_smem(reinterpret_cast<const void*>(attn_dq_kernel), DQ_SMEM);
    if (e != cudaSuccess) return e;
    e = set_smem(reinterpret_cast<const void*>(attn_dkv_kernel), DKV_SMEM);
    if (e != cudaSuccess) return e;
    init = true;
  }
  const int Hkv = a.Hkv > 0 ? a.Hkv : a.H;
  if (a.H % Hkv) return cudaErrorInvalidValue;
  const uint64_t M = static_cast<uint64_t>(a.B) * a.S, W = static_cast<uint64_t>(a.H + 2 * Hkv) * HD,
                 WO = static_cast<uint64_t>(a.H) * HD;
  CUtensorMap tmQ128, tmKV64, tmDO128, tmKV128, tmQ64, tmDO64;
  bool ok = make_tmap_2d_bf16(&tmQ128, a.qkv, W, M, W, 64, 128) && make_tmap_2d_bf16(&tmKV64, a.qkv, W, M, W, 64, 64) &&
            make_tmap_2d_bf16(&tmDO128, a.dout, WO, M, WO, 64, 128) && make_tmap_2d_bf16(&tmDO64, a.dout, WO, M, WO, 64, 64);
  tmKV128 = tmQ128;
  tmQ64 = tmKV64;
  if (!ok) return cudaErrorInvalidValue;
  AttnKParams p{};
  p.B = a.B; p.S = a.S; p.H = a.H;
  p.Hkv = Hkv; p.W = static_cast<int>(W); p.colK0 = a.H * HD; p.colV0 = (a.H + Hkv) * HD;
  p.scale = a.scale;
  p.scale_log2 = a.scale * LOG2E;
  p.lse2 = a.lse;
  p.out = a.out;
  p.delta = a.delta;
  p.dqkv = a.dqkv;
  p.rope_cs = a.rope_cs;
  {
    const long long warps = static_cast<long long>(a.B) * a.S * a.H;
    const int block = 256;
    const long long grid = (warps * 32 + block - 1) / block;
    attn_delta_kernel<<<static_cast<unsigned>(grid), block, 0, s>>>(a.out, a.dout, a.delta, a.B, a.S, a.H);
  }
  attn_dq_kernel<<<a.B * a.H * ((a.S + 255) / 256), BWD_THREADS, DQ_SMEM, s>>>(tmQ128, tmKV64, tmDO128, p);
  attn_dkv_kernel<<<a.B * Hkv * (a.S / 128), BWD_THREADS, DKV_SMEM, s>>>(tmKV128, tmQ64, tmDO64, p);
  return cudaGetLastError();
}

}  // namespace dtx
