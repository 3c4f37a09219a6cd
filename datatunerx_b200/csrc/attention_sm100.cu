// Causal flash attention forward / backward for sm_100a, head_dim 128, on tcgen05 + TMA.
//
// Replaces HF 4.34 eager LlamaAttention (matmul -> mask add -> fp32 softmax -> matmul, SURVEY §2.3 K6,
// reached from cmd/tuning/train.py:299) and its autograd backward (K10).  The [B,H,S,S] score tensor
// never exists: scores live in TMEM, probabilities go TMEM -> registers -> TMEM (bf16) -> tensor core.
//
// Data layout: packed qkv [B*S, (H + 2*Hkv)*128] (per token: q heads | k heads | v heads; Hkv <= H for grouped-query
// attention), out [B*S, H*128],
// lse2 [B,H,S] = log2-domain log-sum-exp of the scaled scores (m + log2 l).
//
// Kernels (one CTA per SM).  Roles are warp-specialised: compute warps where thread r owns TMEM lane r = one row of a score
// tile, MMA issuer warp(s) (warp-uniform code, an elected lane issues every tcgen05.mma) and ONE TMA loader warp:
//   attn_fwd2_kernel  : CTA = two 128-row query tiles sharing one K/V ring of 64-row blocks, ONE ISSUER WARP PER TILE.  S = Q K^T
//                       (UMMA 128x64x16) in two score buffers per tile, softmax in registers (a third of the exponentials on the
//                       FMA pipe), P is written bf16-packed over the score columns it came from and is the TENSOR-MEMORY A
//                       operand of O += P V; O accumulates in tensor memory (lazy rescale).
//   attn_dq1_kernel   : CTA = one 128-row query tile with Q and dO resident in tensor memory.  S, dP = dO V^T (double-buffered),
//                       dS = P o (dP - delta) * scale -> TMEM operand, dQ += dS K (K as MN-major B) accumulates in TMEM; delta =
//                       rowsum(dO o O) and the inverse rotary of dQ are computed here.
//   attn_dkv_kernel   : CTA = 128 KV rows, loops over 64-row Q blocks of every query head of its KV group.  S^T = K Q^T,
//                       dP^T = V dO^T (double-buffered), P^T / dS^T -> TMEM operands, dV += P^T dO, dK += dS^T Q.
// All three take optional true row lengths (tiles that hold only right padding are skipped and written as zeros), an
// optional sliding window (query i sees keys i - window .. i) and an optional table of first rows (PACKED ragged batches:
// sequence b owns rows row_start[b] .. row_start[b+1]) instead of b*S .. (b+1)*S; everything else in the kernels was
// relative to the sequence start already, a tile beyond a sequence's rows returns without touching memory).
// What shaped them (profiles/r01_ncu_attn_issue_bound.txt, r01_ubench_mma_latency.txt, r02_attn_phase_timing_*.txt):
//   * a tcgen05.mma issue blocks for the MMA's duration (the pipe's queue is ~1 deep): whatever else the issuing warp does -
//     mbarrier checks cost ~90 cycles each even when the phase is long complete - is time the tensor pipe idles.  The forward
//     has one issuer per tile; in the backward kernels the ring-slot checks moved to the compute warps, which have slack;
//   * a single warp can feed the tensor pipe only if its loop is tiny: the issuers are unrolled over the ring (slot, buffer,
//     parity are immediates) and add constants to precomputed descriptor low words; loads live in a separate warp;
//   * P / dS never touch shared memory (tcgen05.mma with A in TMEM): no st.shared + proxy fence, no operand re-read, and the
//     freed smem deepens the TMA rings;
//   * output tiles leave through a swizzled staging tile and TMA stores (row-per-thread 16-byte stores are LSU-bound);
//   * every hand-off is an mbarrier; a waiter may never fall two phases behind a parity wait, hence one P barrier per score
//     buffer and dedicated single-phase "all MMAs done" barriers.
// Two backward kernels instead of one with fp32 atomics on dQ: every reduction has a fixed order, so the
// step is bitwise reproducible (needed for the N-rank == 1-rank parity tests).
#include "common.cuh"
#include "kernels.h"

#include <mutex>
#include <type_traits>

namespace dtx {

bool make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                       uint32_t box_outer);

namespace {

// How the MMA issuer warps wait on mbarriers: their wake-up latency is on the critical path of the tensor pipe.
#ifndef DTX_ATTN_ISSUER_SPIN
#define DTX_ATTN_ISSUER_SPIN 1
#endif
#if DTX_ATTN_ISSUER_SPIN
#define ISSUER_WAIT(bar, parity) mbar_wait(bar, parity)
#else
#define ISSUER_WAIT(bar, parity) mbar_wait_backoff(bar, parity)
#endif

// clock64() phase instrumentation (make EXTRA_NVFLAGS=-DDTX_ATTN_TIMING, tools/attn_timing.py): compiled out by default
#ifdef DTX_ATTN_TIMING
#define TM_DECL(...) long long __VA_ARGS__
#define TM_SET(v) v = clock64()
#define TM_ACC(acc, since) acc += clock64() - (since)
#define TM_BLOCK(b) ((b) == 3 || (b) == (int)gridDim.x / 2 || (b) == (int)gridDim.x - 40)
#else
#define TM_DECL(...)
#define TM_SET(v)
#define TM_ACC(acc, since)
#endif

constexpr int HD = 128;      // head dim
constexpr int DKV_THREADS = 320;  // backward kernels: 8 compute warps + MMA issuer warp (8) + TMA loader warp (9)
constexpr float LOG2E = 1.4426950408889634f;
constexpr float FW2_RESCALE_T = 8.f;  // forward: log2 of the factor a row maximum may outgrow its exponent reference by before O is rescaled

struct AttnKParams {
  int B, S, H;
  int Hkv, W;        // kv heads (grouped-query attention: q head h reads kv head h / (H / Hkv)); W = row stride of qkv
  int colK0, colV0;  // first column of the k / v sections inside a qkv row
  float scale_log2;  // softmax scale * log2(e)
  float scale;
  float* lse2;
  bf16* out;
  const float* delta;
  float* delta_w;    // attn_dq1_kernel writes delta here (same buffer the dK/dV kernel then reads)
  bf16* dqkv;
  const float2* rope_cs;  // backward: transposed rotary table [64][rope_stride] (cos, sin), or null
  int rope_stride;
  const int32_t* seq_lens;  // [B] true row lengths (right padding beyond), or null = S
  int window;               // sliding-window attention: query i sees keys i - window .. i; 0 = plain causal
  const int32_t* row_start; // [B+1] packed ragged batch: sequence b owns rows row_start[b] .. row_start[b+1]) (multiples of 128); null: b*S ..
};

// first row of sequence b and the number of rows it owns (S, or its 128-rounded length in a packed batch)
__device__ __forceinline__ int seq_base(const AttnKParams& p, int b, int* cap) {
  if (p.row_start) {
    const int r0 = p.row_start[b];
    *cap = p.row_start[b + 1] - r0;
    return r0;
  }
  *cap = p.S;
  return b * p.S;
}

// true length of batch row b (tiles that start at or beyond it hold only padding)
__device__ __forceinline__ int row_len(const AttnKParams& p, int b) {
  return p.seq_lens ? min(max(p.seq_lens[b], 0), p.S) : p.S;
}
// zero a [128 x 128] bf16 tile of a row-major matrix (skipped padding tiles: downstream GEMMs contract over tokens, so the
// rows must hold finite values - and the gradient of a padded token is exactly zero)
__device__ __forceinline__ void zero_tile_128(bf16* base, long long ld, long long row0, int col0, int tid, int nthreads) {
  for (int i = tid; i < 128 * 16; i += nthreads)
    *reinterpret_cast<uint4*>(base + (row0 + (i >> 4)) * ld + col0 + (i & 15) * 8) = make_uint4(0u, 0u, 0u, 0u);
}

// descriptor low-word advance of the k16-th K=16 slice of a K-major operand made of 64-wide subtiles
__device__ __forceinline__ constexpr uint32_t kmaj_lo(int k16, uint32_t subtile_bytes) {
  return static_cast<uint32_t>(k16 >> 2) * (subtile_bytes >> 4) + static_cast<uint32_t>(k16 & 3) * 2u;
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x for a pair of values on the FMA pipe (no MUFU): round-to-nearest split x = n + f via the 1.5 * 2^23 trick, cubic
// minimax for 2^f on [-0.5, 0.5] (max. relative error 7.5e-5, far below bf16 resolution), 2^n added into the exponent field.
// x is clamped to >= -125 (the polynomial's value can sit just below 1, i.e. in exponent 126, and the exponent field must
// stay positive after adding n; 2^-125 ~ 0); callers guarantee x <= ~16.
// MEASURED AND SWITCHED OFF (exp2_pair's default FMA_EVERY = 0): the exp / dS phase of a block spends 512 of its ~670 cycles
// per scheduler on the ex2 unit (32 ex2 per thread), but moving every other pair here made the phase 780 cycles - the FMA
// pipe becomes the limiter (profiles/r01_attn_phase_timing.txt).  Kept for a cheaper-polynomial / smaller-fraction retry.
__device__ __forceinline__ float2 exp2_fma2(float2 x) {
  x.x = fmaxf(x.x, -125.f);
  x.y = fmaxf(x.y, -125.f);
  const float2 t = __fadd2_rn(x, make_float2(12582912.f, 12582912.f));
  const float2 n = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
  const float2 f = __ffma2_rn(n, make_float2(-1.f, -1.f), x);
  float2 q = __ffma2_rn(make_float2(0.055171649903059006f, 0.055171649903059006f), f, make_float2(0.2426111251115799f, 0.2426111251115799f));
  q = __ffma2_rn(q, f, make_float2(0.6932609677314758f, 0.6932609677314758f));
  q = __ffma2_rn(q, f, make_float2(0.9999280571937561f, 0.9999280571937561f));
  q.x = __int_as_float(__float_as_int(q.x) + (__float_as_int(t.x) << 23));
  q.y = __int_as_float(__float_as_int(q.y) + (__float_as_int(t.y) << 23));
  return q;
}
// exponentials of pair number `pair_idx` of a row: every FMA_EVERY-th pair on the FMA pipe (0 = none), the rest on the MUFU
template <int FMA_EVERY = 0>
__device__ __forceinline__ float2 exp2_pair(float2 x, int pair_idx) {
  if constexpr (FMA_EVERY > 0) {
    if ((pair_idx % FMA_EVERY) == FMA_EVERY - 1) return exp2_fma2(x);
  }
  return make_float2(fast_exp2(x.x), fast_exp2(x.y));
}

// ================================================================================================
// forward, two query tiles per CTA
// ================================================================================================
// Tiles t = 0, 1 own query rows [256*pair + 128*t, +128) and share one K/V ring; while the four softmax warps of one tile
// work on a score block the tensor core serves the other tile, and every scheduler has two softmax warps to interleave.
// The running output lives in TENSOR MEMORY (P V accumulates there) instead of 128 registers per thread: the row maximum
// used as the exponent reference is only moved when the true maximum has grown by more than 2^8 ("lazy rescale"), and
// only then is the output tile read back, scaled and rewritten.  exp2(s - m_ref) <= 256 keeps P and the row sum exact
// enough in bf16 / fp32, and out = O / l, lse = m_ref + log2 l do not depend on which reference was used.
// (r01: the one-tile kernel kept O in registers (255 registers, spills) and folded 128 FMAs per row per block - it was
// bound by its one softmax warp per scheduler, tensor pipe 25 % busy; a two-tile version of THAT design spilled and was
// 2.8x slower.)
// r02 (profiles/r02_attn_phase_timing.txt): with ONE issuer warp the kernel was bound by that warp, not by the tensor pipe or the
// softmax: a tcgen05.mma issue blocks for the MMA's duration (the pipe's queue is ~1 deep), so the issuer's own overhead per
// block - four mbarrier checks of ~90 cycles each even when already complete, commits, fences: ~750 of 1780 cycles - left
// the pipe idle.  Each tile now has its OWN issuer warp: while one sits in a barrier check the other feeds the pipe.  The
// tiles share nothing but the K/V ring (its slots are released when both issuers have committed).
constexpr int FW2_THREADS = 352;             // 8 softmax warps (4 per tile) + one MMA issuer warp per tile (8, 9) + TMA loader warp (10)
constexpr int FW2_NS = 4;
constexpr int FW2_SQ = 0;                    // 2 tiles x (2 x [128 x 128B])
constexpr int FW2_SK = 65536;                // FW2_NS x (2 x [64 x 128B])
constexpr int FW2_SV = FW2_SK + FW2_NS * 16384;
constexpr int FW2_BAR = FW2_SV + FW2_NS * 16384;
constexpr int FW2_SMEM = FW2_BAR + 256 + 1024;

// WIN: sliding-window instantiation (p.window > 0).  The plain causal instantiation carries none of the window code: in the
// backward kernels the extra compares in the (if-converted) mask path cost 100-260 cycles per block on EVERY block
// (profiles/r02_attn_window_regression.txt).
template <int EXP_FMA_EVERY, bool WIN>  // every N-th pair of exponentials on the FMA pipe instead of the MUFU (0 = none)
__global__ void __launch_bounds__(FW2_THREADS, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                 const __grid_constant__ CUtensorMap tmOut, const AttnKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FW2_BAR);
  uint64_t *bar_q = bars, *bar_kv = bars + 1 /*[4]*/, *bar_free = bars + 5 /*[4]*/, *bar_s = bars + 9 /*[t][2]*/, *bar_o = bars + 13 /*[t]*/,
           *bar_fin = bars + 15 /*[t]*/, *bar_p = bars + 17 /*[t][2]*/;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 21);
  static_assert(FW2_NS == 4, "the issue loop is unrolled over a 4-slot ring");

  const int npair = (p.S + 255) / 256;
  const int qp = npair - 1 - (blockIdx.x % npair);  // heavy (late) query pairs first
  const int bh = blockIdx.x / npair;
  const int h = bh % p.H, b = bh / p.H;
  int cap;
  const int row_base = seq_base(p, b, &cap);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int hk = h / (p.H / p.Hkv);
  const int colQ = h * HD, colK = p.colK0 + hk * HD, colV = p.colV0 + hk * HD;
  const int q00 = qp * 256, q01 = qp * 256 + 128;
  if (q00 >= cap) return;  // packed batch: this sequence has no such tile (the rows there belong to the next sequence)
  const int len = min(row_len(p, b), cap);
  // Sliding window: keys before q00 - window are invisible to every row of this CTA: the K/V ring starts at block jlo (both tiles
  // start there - the at most two leading blocks that only tile 0 can see are fully masked for tile 1).  Block counts below are
  // relative to jlo; a row may then meet blocks in which it sees nothing (handled by the -inf-safe softmax reference).
  const int window = WIN ? p.window : 0;
  const int jlo = WIN ? max(0, (q00 - window) / 64) : 0;
  const int n0 = q00 < len ? (q00 + 128) / 64 - jlo : 0, n1 = q01 < len ? (q01 + 128) / 64 - jlo : 0;  // KV blocks each tile needs (0: padding only)
  const int n = max(n0, n1);
  auto nt = [&](int t) { return t ? n1 : n0; };
  if (n0 == 0) {  // both tiles lie in this row's padding (n1 > 0 implies n0 > 0)
    zero_tile_128(p.out, static_cast<long long>(p.H) * HD, row_base + q00, h * HD, tid, FW2_THREADS);
    if (q01 < cap) zero_tile_128(p.out, static_cast<long long>(p.H) * HD, row_base + q01, h * HD, tid, FW2_THREADS);
    if (p.lse2 && tid < 256 && q00 + tid < cap) p.lse2[(static_cast<size_t>(b) * p.H + h) * p.S + q00 + tid] = 0.f;
    return;
  }

  if (tid == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmOut);
    for (int i = 0; i < 17; ++i) mbar_init(&bars[i], 1);
    for (int i = 0; i < FW2_NS; ++i) mbar_init(&bar_free[i], n1 > 0 ? 2 : 1);  // one commit / arrive per issuer warp with a live tile
    // "P_t(j) is in tensor memory": one barrier per score buffer, like bar_s.  With a single barrier per tile a tile whose
    // next score block is already there (double buffering) can complete TWO phases while the issuer is still serving the
    // other tile, and a parity wait that misses a phase never returns.
    for (int i = 0; i < 4; ++i) mbar_init(&bar_p[i], 128);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  // TMEM columns of tile t: score buffers 256t, 256t + 64; output 256t + 128

  if (warp == 10) {
    // ------------------------------------------ TMA loader ------------------------------------------
    if ((tid & 31) == 0) {
      mbar_arrive_expect_tx(bar_q, n1 > 0 ? 65536 : 32768);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (nt(t) == 0) continue;
        const int q0 = t ? q01 : q00;
        tma_load_2d(smem + FW2_SQ + t * 32768, &tmQ, bar_q, colQ, row_base + q0);
        tma_load_2d(smem + FW2_SQ + t * 32768 + 16384, &tmQ, bar_q, colQ + 64, row_base + q0);
      }
      for (int j = 0; j < n; ++j) {
        const int slot = j & 3;
        if (j >= FW2_NS) mbar_wait_backoff(&bar_free[slot], ((j >> 2) - 1) & 1);  // both tiles' P V(j - 4) have read the slot
        mbar_arrive_expect_tx(&bar_kv[slot], 32768);
        tma_load_2d(smem + FW2_SK + slot * 16384, &tmKV, &bar_kv[slot], colK, row_base + (jlo + j) * 64);
        tma_load_2d(smem + FW2_SK + slot * 16384 + 8192, &tmKV, &bar_kv[slot], colK + 64, row_base + (jlo + j) * 64);
        tma_load_2d(smem + FW2_SV + slot * 16384, &tmKV, &bar_kv[slot], colV, row_base + (jlo + j) * 64);
        tma_load_2d(smem + FW2_SV + slot * 16384 + 8192, &tmKV, &bar_kv[slot], colV + 64, row_base + (jlo + j) * 64);
      }
    }
  } else if (warp >= 8) {
    // ------------------------------------------ MMA issuer of tile t (lean: see attn_dkv_kernel) ------------------------------------------
    const int t = warp - 8;
    const int nm = nt(t);
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 128, 0, 1);
    const uint32_t loQ = umma_desc_lo(smem_u32(smem + FW2_SQ + t * 32768), 16), loK = umma_desc_lo(smem_u32(smem + FW2_SK), 16),
                   loVm = umma_desc_lo(smem_u32(smem + FW2_SV), 8192);
    const uint32_t tS = tmem + t * 256, tO = tmem + t * 256 + 128;
    uint64_t *my_s = bar_s + t * 2, *my_p = bar_p + t * 2;
    TM_DECL(ti_kv = 0, ti_s = 0, ti_p = 0, ti_pv = 0, ti_0, ti_a, ti_b);
    TM_SET(ti_0);
    auto issue_s = [&](const int slot, const int buf, const uint32_t parity) {  // S_t = Q_t K^T of the block in ring slot `slot` into buffer buf
      TM_SET(ti_a);
      ISSUER_WAIT(&bar_kv[slot], parity);
      tc_fence_after();
      TM_ACC(ti_kv, ti_a);
      TM_SET(ti_b);
      if (leader) {
#pragma unroll
        for (int k16 = 0; k16 < 8; ++k16)
          umma_bf16(tS + buf * 64, umma_desc_pack(loQ + kmaj_lo(k16, 16384)), umma_desc_pack(loK + slot * 1024 + kmaj_lo(k16, 8192)), idesc_s,
                    k16 > 0 ? 1u : 0u);
        umma_commit(&my_s[buf]);
      }
      TM_ACC(ti_s, ti_b);
    };
    if (nm > 0) {
      ISSUER_WAIT(bar_q, 0);
#ifdef DTX_ATTN_TIMING
      const long long ti_q = clock64() - ti_0;
#endif
      issue_s(0, 0, 0);
      if (nm > 1) issue_s(1, 1, 0);
      for (int base = 0; base < n; base += FW2_NS) {
        const uint32_t rp = (base >> 2) & 1;
#pragma unroll
        for (int u = 0; u < FW2_NS; ++u) {
          const int j = base + u;
          if (j < nm) {
            TM_SET(ti_a);
            ISSUER_WAIT(&my_p[u & 1], (j >> 1) & 1);  // P_t(j) sits bf16-packed in the first 32 columns of score buffer j&1
            tc_fence_after();
            TM_ACC(ti_p, ti_a);
            TM_SET(ti_b);
            if (leader) {
              const uint32_t acc0 = j > 0 ? 1u : 0u;
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_bf16_ts(tO, tS + (u & 1) * 64 + kk * 8, umma_desc_pack(loVm + u * 1024 + kk * 128), idesc_o, kk > 0 ? 1u : acc0);
              umma_commit(&bar_o[t]);
              if (j == nm - 1) umma_commit(&bar_fin[t]);
              umma_commit(&bar_free[u]);  // this tile's S(j) and P V(j) are the only MMAs of this warp that read ring slot u
            }
            TM_ACC(ti_pv, ti_b);
            if (j + 2 < nm) issue_s((u + 2) & 3, u & 1, (u + 2 >= FW2_NS) ? (rp ^ 1u) : rp);
          } else if (j < n) {
            if (leader) mbar_arrive(&bar_free[u]);  // the other tile's block only (it reaches two blocks further): nothing to read here
          }
        }
      }
#ifdef DTX_ATTN_TIMING
      if (leader && TM_BLOCK((int)blockIdx.x))
        printf("[fwd2 issuer] block %d tile %d n %d total %lld wait_q %lld | per block: wait_p %lld wait_kv %lld issue_s %lld issue_pv %lld\n",
               (int)blockIdx.x, t, nm, clock64() - ti_0, ti_q, ti_p / nm, ti_kv / nm, ti_s / nm, ti_pv / nm);
#endif
    }
  } else {
    // ------------------------------------------ softmax warps ------------------------------------------
    const int t = warp >> 2, w = warp & 3;
    const int r = w * 32 + (tid & 31);
    const int n_mine = nt(t);
    const int q0 = t ? q01 : q00;
    const int qrow = q0 + r;
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(w * 32) << 16) + t * 256;
    const uint32_t T_O = 128;
    uint64_t *my_s = bar_s + t * 2, *my_o = bar_o + t, *my_p = bar_p + t * 2;
    float m_ref = -INFINITY, l_run = 0.f;
    TM_DECL(tc_w = 0, tc_l = 0, tc_m = 0, tc_s = 0, tc_0, tc_a, tc_b, tc_c, tc_d);
    TM_SET(tc_0);

    for (int j = 0; j < n_mine; ++j) {
      const int kv0 = (jlo + j) * 64;
      TM_SET(tc_a);
      mbar_wait(&my_s[j & 1], (j >> 1) & 1);
      tc_fence_after();
      TM_SET(tc_b);
      uint32_t sv[64];
      {
        uint32_t(&lo)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[0]);
        uint32_t(&hi)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[32]);
        tmem_ld32(t_lane + (j & 1) * 64, lo);
        tmem_ld32(t_lane + (j & 1) * 64 + 32, hi);
        tmem_ld_wait();
      }
      TM_SET(tc_c);
      if (kv0 + 63 > q0) {  // diagonal blocks: causal mask
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if (kv0 + c > qrow) sv[c] = 0xff800000u;  // -inf
      }
      if (WIN && kv0 < q0 + 127 - window) {  // blocks that straddle the far edge of the window
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if (kv0 + c < qrow - window) sv[c] = 0xff800000u;
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
      if (j == 0) {
        m_ref = mx;  // plain causal: block 0 always holds column 0 <= qrow (finite); with a window the row may see nothing yet (-inf)
      } else {
        const bool grow = mx > m_ref + FW2_RESCALE_T;  // also the first visible key of a windowed row (m_ref = -inf): alpha = 0, O = 0 so far
        if (__any_sync(0xffffffffu, grow)) {  // rare: move the reference of the rows that need it and rescale their output
          const float m_new = grow ? mx : m_ref;
          // (-inf) - (-inf) of a windowed row that still sees nothing would be NaN
          const float alpha = (!WIN || grow) ? fast_exp2(m_ref - m_new) : 1.f;
          mbar_wait(my_o, (j - 1) & 1);  // P V(j-1) (and every earlier one) has landed in the output tile
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t v[32];
            tmem_ld32(t_lane + T_O + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * alpha);
            tmem_st32(t_lane + T_O + c * 32, v);
          }
          tmem_st_wait();
          l_run *= alpha;
          m_ref = m_new;
        }
      }
      uint32_t pk[32];
      float2 rs = make_float2(0.f, 0.f);
      // a row that has not met a visible key yet keeps m_ref = -inf: every score is -inf and must give exp2(-inf) = 0, not
      // exp2(-inf + inf) = NaN
      const float m_use = (WIN && m_ref == -INFINITY) ? 0.f : m_ref;
      const float2 sl2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(-m_use, -m_use);
#pragma unroll
      for (int c = 0; c < 64; c += 2) {  // packed fp32 pairs; every other pair of exponentials on the FMA pipe (exp2_fma2)
        const float2 pr = exp2_pair<EXP_FMA_EVERY>(__ffma2_rn(make_float2(__uint_as_float(sv[c]), __uint_as_float(sv[c + 1])), sl2, nm2), c >> 1);
        rs = __fadd2_rn(rs, pr);
        pk[c >> 1] = pack_bf16x2(pr.x, pr.y);
      }
      l_run += rs.x + rs.y;
      TM_SET(tc_d);
      tmem_st32(t_lane + (j & 1) * 64, pk);  // A operand of the P V MMA, read straight from tensor memory
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&my_p[j & 1]);
#ifdef DTX_ATTN_TIMING
      tc_w += tc_b - tc_a; tc_l += tc_c - tc_b; tc_m += tc_d - tc_c; tc_s += clock64() - tc_d;
#endif
    }
#ifdef DTX_ATTN_TIMING
    const long long tc_loop = clock64() - tc_0;
#endif
    if (n_mine > 0) {
      mbar_wait(&bar_fin[t], 0);
      tc_fence_after();
#ifdef DTX_ATTN_TIMING
      const long long tc_fin = clock64() - tc_0 - tc_loop;
#endif
      const float inv_l = 1.f / l_run;
      // output tile -> bf16 -> 128B-swizzled staging tile (this tile's Q buffer: every S MMA has completed) -> TMA store
      uint8_t* stage = smem + FW2_SQ + t * 32768;
      const uint32_t stage_addr = smem_u32(stage);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(t_lane + T_O + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8)
          sts128(stage_addr + (c >> 1) * 16384 + sw128_offset(r, (c & 1) * 4 + c8),
                 pack_bf16x2(__uint_as_float(v[c8 * 8 + 0]) * inv_l, __uint_as_float(v[c8 * 8 + 1]) * inv_l),
                 pack_bf16x2(__uint_as_float(v[c8 * 8 + 2]) * inv_l, __uint_as_float(v[c8 * 8 + 3]) * inv_l),
                 pack_bf16x2(__uint_as_float(v[c8 * 8 + 4]) * inv_l, __uint_as_float(v[c8 * 8 + 5]) * inv_l),
                 pack_bf16x2(__uint_as_float(v[c8 * 8 + 6]) * inv_l, __uint_as_float(v[c8 * 8 + 7]) * inv_l));
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + t, 128);
      if (r == 0) {
        tma_store_2d(&tmOut, stage, h * HD, row_base + q0);
        tma_store_2d(&tmOut, stage + 16384, h * HD + 64, row_base + q0);
        tma_store_commit();
        tma_store_wait_read0();
      }
      if (p.lse2) p.lse2[(static_cast<size_t>(b) * p.H + h) * p.S + qrow] = m_ref + log2f(l_run);
#ifdef DTX_ATTN_TIMING
      if (r == 0 && TM_BLOCK((int)blockIdx.x))
        printf("[fwd2 softmax] block %d tile %d n %d loop %lld (incl. first wait) wait_fin %lld epilogue %lld | per block: wait_s %lld ld %lld math %lld st+arrive %lld\n",
               (int)blockIdx.x, t, n_mine, tc_loop, tc_fin, clock64() - tc_0 - tc_loop - tc_fin, tc_w / n_mine, tc_l / n_mine, tc_m / n_mine, tc_s / n_mine);
#endif
    } else if (q0 < cap) {  // this tile holds only padding: zeros (see zero_tile_128)
      zero_tile_128(p.out, static_cast<long long>(p.H) * HD, row_base + q0, h * HD, r, 128);
      if (p.lse2) p.lse2[(static_cast<size_t>(b) * p.H + h) * p.S + qrow] = 0.f;
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
// backward: dQ, one query tile per CTA with Q and dO RESIDENT IN TENSOR MEMORY
// ================================================================================================
// The score MMAs of the two-group kernel read their 128-row A operand (Q / dO, 4 KB per 128x64x16 instruction) from shared
// memory for every 64-row K/V block: 48 cycles per instruction instead of 32 (profiles/r01_ubench_mma_latency.txt).  Here the
// tile's Q and dO rows are copied once into tensor memory (bf16-packed, 64 columns each) and every MMA of the kernel takes
// its A operand from there; shared memory only streams K and V.  TMEM: S 2x64 | dP 2x64 | dQ 128 | Q 64 | dO 64 = 512.
// Structure = attn_dkv_kernel: double-buffered scores, 8 compute warps (lane quadrant x column half), issuer + loader warps.
constexpr int DQ1_NS = 4;
constexpr int DQ1_SQ = 0;                       // 2 x [128 x 128B]  (TMA landing zone, later the dQ staging tile)
constexpr int DQ1_SDO = 32768;                  // 2 x [128 x 128B]
constexpr int DQ1_SO = 65536;                   // 2 x [128 x 128B]  forward output tile, only for delta = rowsum(dO o O)
constexpr int DQ1_SK = 98304;                   // DQ1_NS x (2 x [64 x 128B])
constexpr int DQ1_SV = DQ1_SK + DQ1_NS * 16384;
constexpr int DQ1_BAR = DQ1_SV + DQ1_NS * 16384;
constexpr int DQ1_XD = DQ1_BAR + 256;           // [2 halves][128] partial delta sums
constexpr int DQ1_SMEM = DQ1_XD + 1024 + 1024;

// WIN: see attn_fwd2_kernel.  EXP_FMA: every N-th pair of exponentials on the FMA pipe (A/B switch "attn_dq_exp_fma_every").
// The MMA issuer checks the K/V ring slot of block j + 2 itself: letting the compute warps do it before they report block j
// (as in the dK/dV kernel, whose compute warps have slack) was measured 4 % slower here - the block is not there yet when
// they look (st+arrive 45 -> 240 cycles), and the exp / dS phase is on this kernel's critical chain
// S(j) -> dS(j) -> dQ(j) -> S(j+2) (profiles/r02_attn_window_regression.txt).
template <bool WIN, int EXP_FMA>
__global__ void __launch_bounds__(DKV_THREADS, 1)
attn_dq1_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                const __grid_constant__ CUtensorMap tmDO, const __grid_constant__ CUtensorMap tmO,
                const __grid_constant__ CUtensorMap tmOut, const AttnKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DQ1_BAR);
  uint64_t *bar_q = bars, *bar_kv = bars + 1 /*[4]*/, *bar_free = bars + 5 /*[4]*/, *bar_s = bars + 9 /*[2]*/, *bar_fin = bars + 11,
           *bar_qt = bars + 12, *bar_p = bars + 13 /*[2]*/;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 15);
  static_assert(DQ1_NS == 4, "the issue loop is unrolled over a 4-slot ring");

  const int nqb = p.S / 128;
  const int qb = nqb - 1 - (blockIdx.x % nqb);  // heavy (late) query tiles first
  const int bh = blockIdx.x / nqb;
  const int h = bh % p.H, b = bh / p.H;
  const int q0 = qb * 128;
  int cap;
  const int row_base = seq_base(p, b, &cap);
  if (q0 >= cap) return;  // packed batch: not a tile of this sequence
  const int window = WIN ? p.window : 0;
  const int jlo = WIN ? max(0, (q0 - window) / 64) : 0;  // sliding window: first K/V block any row of the tile can see
  const int n = (q0 + 128) / 64 - jlo;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int hk = h / (p.H / p.Hkv);
  const int colQ = h * HD, colK = p.colK0 + hk * HD, colV = p.colV0 + hk * HD;
  if (q0 >= min(row_len(p, b), cap)) {  // the tile holds only padding: dQ = 0, delta = 0
    zero_tile_128(p.dqkv, p.W, row_base + q0, colQ, tid, DKV_THREADS);
    if (tid < 128) p.delta_w[(static_cast<size_t>(b) * p.H + h) * p.S + q0 + tid] = 0.f;
    return;
  }

  if (tid == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmO);
    tma_prefetch_desc(&tmOut);
    for (int i = 0; i < 12; ++i) mbar_init(&bars[i], 1);
    mbar_init(bar_qt, 256);
    mbar_init(&bar_p[0], 256);
    mbar_init(&bar_p[1], 256);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t T_S = 0 /* +64*buf */, T_DP = 128 /* +64*buf */, T_DQ = 256, T_QT = 384, T_DOT = 448;

  if (warp == 9) {
    // ------------------------------------------ TMA loader ------------------------------------------
    if ((tid & 31) == 0) {
      mbar_arrive_expect_tx(bar_q, 98304);
      tma_load_2d(smem + DQ1_SQ, &tmQ, bar_q, colQ, row_base + q0);
      tma_load_2d(smem + DQ1_SQ + 16384, &tmQ, bar_q, colQ + 64, row_base + q0);
      tma_load_2d(smem + DQ1_SDO, &tmDO, bar_q, h * HD, row_base + q0);
      tma_load_2d(smem + DQ1_SDO + 16384, &tmDO, bar_q, h * HD + 64, row_base + q0);
      tma_load_2d(smem + DQ1_SO, &tmO, bar_q, h * HD, row_base + q0);
      tma_load_2d(smem + DQ1_SO + 16384, &tmO, bar_q, h * HD + 64, row_base + q0);
      for (int j = 0; j < n; ++j) {
        const int slot = j & 3;
        if (j >= DQ1_NS) mbar_wait_backoff(&bar_free[slot], ((j >> 2) - 1) & 1);  // dQ MMA of block j - 4 has read the slot
        mbar_arrive_expect_tx(&bar_kv[slot], 32768);
        tma_load_2d(smem + DQ1_SK + slot * 16384, &tmKV, &bar_kv[slot], colK, row_base + (jlo + j) * 64);
        tma_load_2d(smem + DQ1_SK + slot * 16384 + 8192, &tmKV, &bar_kv[slot], colK + 64, row_base + (jlo + j) * 64);
        tma_load_2d(smem + DQ1_SV + slot * 16384, &tmKV, &bar_kv[slot], colV, row_base + (jlo + j) * 64);
        tma_load_2d(smem + DQ1_SV + slot * 16384 + 8192, &tmKV, &bar_kv[slot], colV + 64, row_base + (jlo + j) * 64);
      }
    }
  } else if (warp == 8) {
    // ------------------------------------------ MMA issuer (lean: see attn_dkv_kernel) ------------------------------------------
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t idesc_dq = umma_idesc_bf16(128, 128, 0, 1);
    const uint32_t loK = umma_desc_lo(smem_u32(smem + DQ1_SK), 16), loV = umma_desc_lo(smem_u32(smem + DQ1_SV), 16),
                   loKm = umma_desc_lo(smem_u32(smem + DQ1_SK), 8192);
#ifdef DTX_ATTN_TIMING
    long long ti_kv = 0, ti_mma = 0;
#endif
    // S = Q K^T, dP = dO V^T (A operands from TMEM)
    auto issue_s = [&](const int slot, const int buf, const uint32_t parity, const bool wait_slot) {
#ifdef DTX_ATTN_TIMING
      const long long tk0 = clock64();
#endif
      if (wait_slot) {
        ISSUER_WAIT(&bar_kv[slot], parity);
        tc_fence_after();
      }
#ifdef DTX_ATTN_TIMING
      const long long tk1 = clock64();
      ti_kv += tk1 - tk0;
#endif
      if (leader) {
#pragma unroll
        for (int k16 = 0; k16 < 8; ++k16)
          umma_bf16_ts(tmem + T_S + buf * 64, tmem + T_QT + k16 * 8, umma_desc_pack(loK + slot * 1024 + kmaj_lo(k16, 8192)), idesc_s,
                       k16 > 0 ? 1u : 0u);
#pragma unroll
        for (int k16 = 0; k16 < 8; ++k16)
          umma_bf16_ts(tmem + T_DP + buf * 64, tmem + T_DOT + k16 * 8, umma_desc_pack(loV + slot * 1024 + kmaj_lo(k16, 8192)), idesc_s,
                       k16 > 0 ? 1u : 0u);
        umma_commit(&bar_s[buf]);
      }
#ifdef DTX_ATTN_TIMING
      ti_mma += clock64() - tk1;
#endif
    };
#ifdef DTX_ATTN_TIMING
    long long ti_p = 0, ti_t0 = clock64();
#endif
    ISSUER_WAIT(bar_qt, 0);  // the compute warps have moved Q and dO into tensor memory
    tc_fence_after();
#ifdef DTX_ATTN_TIMING
    const long long ti_qt = clock64() - ti_t0;
#endif
    issue_s(0, 0, 0, true);
    if (n > 1) issue_s(1, 1, 0, true);
    for (int base = 0; base < n; base += DQ1_NS) {
      const uint32_t rp = (base >> 2) & 1;
#pragma unroll
      for (int u = 0; u < DQ1_NS; ++u) {
        const int j = base + u;
        if (j < n) {
#ifdef DTX_ATTN_TIMING
          const long long tw0 = clock64();
#endif
          ISSUER_WAIT(&bar_p[u & 1], (j >> 1) & 1);  // dS(j) sits bf16-packed in score buffer u&1 (columns 0-15 and 32-47)
          tc_fence_after();
#ifdef DTX_ATTN_TIMING
          ti_p += clock64() - tw0;
#endif
          if (leader) {
            const uint32_t acc0 = j > 0 ? 1u : 0u;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16_ts(tmem + T_DQ, tmem + T_S + (u & 1) * 64 + (kk >> 1) * 32 + (kk & 1) * 8,
                           umma_desc_pack(loKm + u * 1024 + kk * 128), idesc_dq, kk > 0 ? 1u : acc0);
            umma_commit(&bar_free[u]);
            if (j == n - 1) umma_commit(bar_fin);
          }
          if (j + 2 < n) issue_s((u + 2) & 3, u & 1, (u + 2 >= DQ1_NS) ? (rp ^ 1u) : rp, true);
        }
      }
    }
#ifdef DTX_ATTN_TIMING
    if (leader && TM_BLOCK((int)blockIdx.x))
      printf("[dq1 issuer] block %d n %d total %lld wait_qt %lld per block: wait_p %lld wait_kv %lld issue_s %lld\n", (int)blockIdx.x, n,
             clock64() - ti_t0, ti_qt, ti_p / n, ti_kv / n, ti_mma / n);
#endif
  } else {
    // ------------------------------------------ 8 compute warps ------------------------------------------
    const int rw = warp & 3, half = warp >> 2;
    const int r = rw * 32 + (tid & 31);
    const int qrow = q0 + r;
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(rw * 32) << 16);
    const size_t stat_idx = (static_cast<size_t>(b) * p.H + h) * p.S + qrow;
    const float lse2 = p.lse2[stat_idx];
    float delta;
    {  // Q / dO rows: swizzled smem -> registers -> tensor memory (thread (r, half) moves the 64 head-dim columns of its half);
       // delta = rowsum(dO o O) falls out of the same pass (the dO words are in registers) and is published for the dK/dV kernel
      mbar_wait(bar_q, 0);
      const uint32_t sq = smem_u32(smem + DQ1_SQ) + half * 16384, sdo = smem_u32(smem + DQ1_SDO) + half * 16384,
                     so = smem_u32(smem + DQ1_SO) + half * 16384;
      uint32_t v[32];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 x = lds128u(sq + sw128_offset(r, c));
        v[4 * c] = x.x; v[4 * c + 1] = x.y; v[4 * c + 2] = x.z; v[4 * c + 3] = x.w;
      }
      tmem_st32(t_lane + T_QT + half * 32, v);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 x = lds128u(sdo + sw128_offset(r, c));
        v[4 * c] = x.x; v[4 * c + 1] = x.y; v[4 * c + 2] = x.z; v[4 * c + 3] = x.w;
      }
      tmem_st32(t_lane + T_DOT + half * 32, v);
      float part = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 x = lds128u(so + sw128_offset(r, c));
        const uint32_t ow[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 o2 = unpack_bf16x2(ow[i]), d2 = unpack_bf16x2(v[4 * c + i]);
          part = fmaf(o2.x, d2.x, part);
          part = fmaf(o2.y, d2.y, part);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_qt);
      float* xd = reinterpret_cast<float*>(smem + DQ1_XD);
      xd[half * 128 + r] = part;
      named_bar_sync(2, 256);
      delta = part + xd[(1 - half) * 128 + r];
      if (half == 0) p.delta_w[stat_idx] = delta;
    }
#ifdef DTX_ATTN_TIMING
    long long tc_wait = 0, tc_ld = 0, tc_math = 0, tc_st = 0;
    const long long tc_t0 = clock64();
#endif
    for (int j = 0; j < n; ++j) {
      const int kv0 = (jlo + j) * 64;
#ifdef DTX_ATTN_TIMING
      const long long t_a = clock64();
#endif
      mbar_wait(&bar_s[j & 1], (j >> 1) & 1);
      tc_fence_after();
#ifdef DTX_ATTN_TIMING
      const long long t_b = clock64();
#endif
      uint32_t sv[32], dv[32], pk[16];
      tmem_ld32(t_lane + T_S + (j & 1) * 64 + half * 32, sv);
      tmem_ld32(t_lane + T_DP + (j & 1) * 64 + half * 32, dv);
      tmem_ld_wait();
#ifdef DTX_ATTN_TIMING
      const long long t_c = clock64();
#endif
      const bool need_mask = (kv0 + 63 > q0) || (WIN && kv0 < q0 + 127 - window);
      const int klo = WIN ? qrow - window : -(1 << 30);  // first visible key of this row
      // packed fp32 pairs (FFMA2 / FMUL2): dS = P o (dP * scale - delta * scale), P = 2^(S * scale_log2 - lse2)
      const float2 sl2 = make_float2(p.scale_log2, p.scale_log2), nl2 = make_float2(-lse2, -lse2), sc2 = make_float2(p.scale, p.scale),
                   nd2 = make_float2(-delta * p.scale, -delta * p.scale);
      // two copies of the loop, the mask branch outside: inside, the compiler predicates the compares into every block
      auto ds_math = [&](auto masked) {
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float2 x = __ffma2_rn(make_float2(__uint_as_float(sv[e]), __uint_as_float(sv[e + 1])), sl2, nl2);
          float2 pr = exp2_pair<EXP_FMA>(x, e >> 1);
          if constexpr (decltype(masked)::value) {
            const int k0 = kv0 + half * 32 + e;
            if (k0 > qrow || (WIN && k0 < klo)) pr.x = 0.f;
            if (k0 + 1 > qrow || (WIN && k0 + 1 < klo)) pr.y = 0.f;
          }
          const float2 dd = __ffma2_rn(make_float2(__uint_as_float(dv[e]), __uint_as_float(dv[e + 1])), sc2, nd2);
          const float2 ds = __fmul2_rn(pr, dd);
          pk[e >> 1] = pack_bf16x2(ds.x, ds.y);
        }
      };
      if (need_mask) ds_math(std::true_type{}); else ds_math(std::false_type{});
#ifdef DTX_ATTN_TIMING
      const long long t_d = clock64();
#endif
      tmem_st16(t_lane + T_S + (j & 1) * 64 + half * 32, pk);  // over this thread's own (already loaded) score columns
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&bar_p[j & 1]);
#ifdef DTX_ATTN_TIMING
      tc_wait += t_b - t_a; tc_ld += t_c - t_b; tc_math += t_d - t_c; tc_st += clock64() - t_d;
#endif
    }
#ifdef DTX_ATTN_TIMING
    if ((tid == 0 || tid == 128) && TM_BLOCK((int)blockIdx.x))
      printf("[dq1 compute] block %d tid %d n %d loop %lld per block: wait_s %lld ld %lld math %lld st+arrive %lld\n", (int)blockIdx.x, tid, n,
             clock64() - tc_t0, tc_wait / n, tc_ld / n, tc_math / n, tc_st / n);
#endif
    mbar_wait(bar_fin, 0);
    tc_fence_after();
    // dQ tile -> (inverse rotary) -> bf16 -> 128B-swizzled staging tile (the Q landing zone: its contents live in tensor
    // memory) -> TMA store.  Thread (r, half) owns the 32-column chunks `half` and `half + 2`: rotary partners i, i + 64.
    {
      uint32_t lo[32], hi[32];
      tmem_ld32(t_lane + T_DQ + half * 32, lo);
      tmem_ld32(t_lane + T_DQ + 64 + half * 32, hi);
      tmem_ld_wait();
      if (p.rope_cs) {
        const float2* cs = p.rope_cs + static_cast<size_t>(half * 32) * p.rope_stride + qrow;
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          const float2 tcs = __ldg(cs + static_cast<size_t>(jj) * p.rope_stride);
          const float x0 = __uint_as_float(lo[jj]), x1 = __uint_as_float(hi[jj]);
          lo[jj] = __float_as_uint(x0 * tcs.x + x1 * tcs.y);
          hi[jj] = __float_as_uint(x1 * tcs.x - x0 * tcs.y);
        }
      }
      const uint32_t stage_addr = smem_u32(smem + DQ1_SQ);
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        sts128(stage_addr + sw128_offset(r, half * 4 + c8),
               pack_bf16x2(__uint_as_float(lo[c8 * 8 + 0]), __uint_as_float(lo[c8 * 8 + 1])),
               pack_bf16x2(__uint_as_float(lo[c8 * 8 + 2]), __uint_as_float(lo[c8 * 8 + 3])),
               pack_bf16x2(__uint_as_float(lo[c8 * 8 + 4]), __uint_as_float(lo[c8 * 8 + 5])),
               pack_bf16x2(__uint_as_float(lo[c8 * 8 + 6]), __uint_as_float(lo[c8 * 8 + 7])));
        sts128(stage_addr + 16384 + sw128_offset(r, half * 4 + c8),
               pack_bf16x2(__uint_as_float(hi[c8 * 8 + 0]), __uint_as_float(hi[c8 * 8 + 1])),
               pack_bf16x2(__uint_as_float(hi[c8 * 8 + 2]), __uint_as_float(hi[c8 * 8 + 3])),
               pack_bf16x2(__uint_as_float(hi[c8 * 8 + 4]), __uint_as_float(hi[c8 * 8 + 5])),
               pack_bf16x2(__uint_as_float(hi[c8 * 8 + 6]), __uint_as_float(hi[c8 * 8 + 7])));
      }
    }
    fence_proxy_async_smem();
    named_bar_sync(1, 256);
    if (tid == 0) {
      tma_store_2d(&tmOut, smem + DQ1_SQ, colQ, row_base + q0);
      tma_store_2d(&tmOut, smem + DQ1_SQ + 16384, colQ + 64, row_base + q0);
      tma_store_commit();
      tma_store_wait_read0();
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
constexpr int DKV_NS = 4;                      // Q / dO ring depth
constexpr int DKV_SK = 0;                      // 2 x [128 x 128B]
constexpr int DKV_SV = 32768;                  // 2 x [128 x 128B]
constexpr int DKV_SQ = 65536;                  // DKV_NS x (2 x [64 x 128B])
constexpr int DKV_SDO = DKV_SQ + DKV_NS * 16384;    // DKV_NS x (2 x [64 x 128B])
constexpr int DKV_STAT = DKV_SDO + DKV_NS * 16384;  // DKV_NS x (lse2[64], delta[64]) fp32, filled by bulk copies with the Q ring
constexpr int DKV_BAR = DKV_STAT + DKV_NS * 512;
constexpr int DKV_SMEM = DKV_BAR + 256 + 1024;

// NCW compute warps (8 or 16): warp w owns TMEM lanes (kv rows) 32*(w&3)..+31 and the query columns of column group w>>2
// (64 / (NCW/4) columns per thread).  16 warps halve the latency of the exp / dS phase of a block (four warps per scheduler
// to interleave instead of two) - it sits on the critical path between the score MMAs and the accumulating MMAs.
template <int NCW, bool WIN>
__global__ void __launch_bounds__((NCW + 2) * 32, 1)
attn_dkv_kernel(const __grid_constant__ CUtensorMap tmKV128, const __grid_constant__ CUtensorMap tmQ64,
                const __grid_constant__ CUtensorMap tmDO64, const __grid_constant__ CUtensorMap tmOut, const AttnKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DKV_BAR);
  uint64_t *bar_kv = bars, *bar_q = bars + 1 /*[DKV_NS]*/, *bar_s = bars + 1 + DKV_NS /*[2]*/, *bar_free = bars + 3 + DKV_NS /*[DKV_NS]*/,
           *bar_p = bars + 3 + 2 * DKV_NS /*[2]: one per score buffer, so that no waiter can fall two phases behind*/,
           *bar_fin = bars + 5 + 2 * DKV_NS;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 6 + 2 * DKV_NS);
  static_assert(DKV_NS == 4, "the issue loop is unrolled over a 4-slot ring (slot and buffer indices are compile-time)");

  const int nkb = p.S / 128;
  const int kb = blockIdx.x % nkb;  // early KV blocks see the most query blocks: they come first
  const int bh = blockIdx.x / nkb;
  const int hk = bh % p.Hkv, b = bh / p.Hkv;  // one CTA per (batch, KV head, 128-row KV block)
  const int grp = p.H / p.Hkv;                // query heads sharing this KV head (1 = multi-head attention)
  const int kv0 = kb * 128;
  int cap;
  const int row_base = seq_base(p, b, &cap);
  if (kv0 >= cap) return;  // packed batch: not a tile of this sequence
  const int i0 = kv0 / 64;
  const int len = min(row_len(p, b), cap);
  // query blocks per query head that see this KV block and hold a real token; with a sliding window the last query row that
  // sees any key of the tile is kv0 + 127 + window
  const int window = WIN ? p.window : 0;
  const int q_end = WIN ? min((len + 63) / 64, (kv0 + 127 + window) / 64 + 1) : (len + 63) / 64;
  const int nq = q_end - i0;
  const int n = nq * grp;                     // streamed (head, query block) pairs; it -> head it / nq, block it % nq
  const int tid = threadIdx.x, warp = tid >> 5;
  const int colK = p.colK0 + hk * HD, colV = p.colV0 + hk * HD;
  const float* g_lse = p.lse2 + (static_cast<size_t>(b) * p.H + hk * grp) * p.S;
  const float* g_delta = p.delta + (static_cast<size_t>(b) * p.H + hk * grp) * p.S;
  if (kv0 >= len) {  // the KV tile holds only padding: dK = dV = 0
    zero_tile_128(p.dqkv, p.W, row_base + kv0, colK, tid, (NCW + 2) * 32);
    zero_tile_128(p.dqkv, p.W, row_base + kv0, colV, tid, (NCW + 2) * 32);
    return;
  }

  if (tid == 0) {
    tma_prefetch_desc(&tmKV128);
    tma_prefetch_desc(&tmQ64);
    tma_prefetch_desc(&tmDO64);
    tma_prefetch_desc(&tmOut);
    for (int i = 0; i < 3 + 2 * DKV_NS; ++i) mbar_init(&bars[i], 1);
    mbar_init(&bar_p[0], NCW * 32);
    mbar_init(&bar_p[1], NCW * 32);
    mbar_init(bar_fin, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t T_ST = 0 /* +64*buf */, T_DPT = 128 /* +64*buf */, T_DV = 256, T_DK = 384;

  constexpr int CPT = 64 / (NCW / 4);  // query columns per compute thread
  if (warp == NCW + 1) {
    // ------------------------------------------ TMA loader ------------------------------------------
    if ((tid & 31) == 0) {
      mbar_arrive_expect_tx(bar_kv, 65536);
      tma_load_2d(smem + DKV_SK, &tmKV128, bar_kv, colK, row_base + kv0);
      tma_load_2d(smem + DKV_SK + 16384, &tmKV128, bar_kv, colK + 64, row_base + kv0);
      tma_load_2d(smem + DKV_SV, &tmKV128, bar_kv, colV, row_base + kv0);
      tma_load_2d(smem + DKV_SV + 16384, &tmKV128, bar_kv, colV + 64, row_base + kv0);
      int hq = 0, qi = 0;  // (query head of the group, query block) of pair ii
      for (int ii = 0; ii < n; ++ii) {
        const int slot = ii & (DKV_NS - 1);
        if (ii >= DKV_NS) mbar_wait_backoff(&bar_free[slot], ((ii >> 2) - 1) & 1);  // dV/dK MMAs of pair ii - 4 have read the slot
        const int qs = (i0 + qi) * 64, colQ = (hk * grp + hq) * HD;
        mbar_arrive_expect_tx(&bar_q[slot], 32768 + 512);
        tma_load_2d(smem + DKV_SQ + slot * 16384, &tmQ64, &bar_q[slot], colQ, row_base + qs);
        tma_load_2d(smem + DKV_SQ + slot * 16384 + 8192, &tmQ64, &bar_q[slot], colQ + 64, row_base + qs);
        tma_load_2d(smem + DKV_SDO + slot * 16384, &tmDO64, &bar_q[slot], colQ, row_base + qs);
        tma_load_2d(smem + DKV_SDO + slot * 16384 + 8192, &tmDO64, &bar_q[slot], colQ + 64, row_base + qs);
        tma_load_1d(smem + DKV_STAT + slot * 512, g_lse + static_cast<size_t>(hq) * p.S + qs, 256, &bar_q[slot]);
        tma_load_1d(smem + DKV_STAT + slot * 512 + 256, g_delta + static_cast<size_t>(hq) * p.S + qs, 256, &bar_q[slot]);
        if (++qi == nq) { qi = 0; ++hq; }
      }
    }
  } else if (warp == NCW) {
    // ------------------------------------------ MMA issuer ------------------------------------------
    // Everything here is on the critical path of the tensor pipe (24 MMAs of 32-64 cycles per pair): the loop is unrolled
    // over the ring so that slot / buffer / parity are immediates, and operand descriptors are one add off precomputed
    // low words.  (r01: the first version rebuilt every 64-bit descriptor and ran ~490 instructions per pair on this single
    // warp - the tensor pipe idled 2/3 of the time waiting for it, profiles/r01_ncu_attn_issue_bound.txt.)
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t idesc_g = umma_idesc_bf16(128, 128, 0, 1);
    const uint32_t loK = umma_desc_lo(smem_u32(smem + DKV_SK), 16), loV = umma_desc_lo(smem_u32(smem + DKV_SV), 16),
                   loQ = umma_desc_lo(smem_u32(smem + DKV_SQ), 16), loDO = umma_desc_lo(smem_u32(smem + DKV_SDO), 16),
                   loQm = umma_desc_lo(smem_u32(smem + DKV_SQ), 8192), loDOm = umma_desc_lo(smem_u32(smem + DKV_SDO), 8192);
    TM_DECL(ti_q = 0, ti_s = 0, ti_p = 0, ti_acc = 0, ti_0, ti_a, ti_b);
    TM_SET(ti_0);
    // S^T = K Q^T, dP^T = V dO^T into buffer buf.  Only the two prologue calls wait for the ring slot themselves: inside the
    // loop the compute warps have already waited for the slot of pair ii + 2 before they arrive on bar_p(ii) (an mbarrier
    // check costs this warp ~90 cycles even when the phase is long complete, and every cycle it spends outside a
    // tcgen05.mma issue is a cycle the tensor pipe idles - profiles/r02_attn_phase_timing.txt).
    auto issue_s = [&](const int slot, const int buf, const uint32_t parity, const bool wait_slot) {
      TM_SET(ti_a);
      if (wait_slot) {
        ISSUER_WAIT(&bar_q[slot], parity);
        tc_fence_after();
      }
      TM_ACC(ti_q, ti_a);
      TM_SET(ti_b);
      if (leader) {
#pragma unroll
        for (int k16 = 0; k16 < 8; ++k16)
          umma_bf16(tmem + T_ST + buf * 64, umma_desc_pack(loK + kmaj_lo(k16, 16384)),
                    umma_desc_pack(loQ + slot * 1024 + kmaj_lo(k16, 8192)), idesc_s, k16 > 0 ? 1u : 0u);
#pragma unroll
        for (int k16 = 0; k16 < 8; ++k16)
          umma_bf16(tmem + T_DPT + buf * 64, umma_desc_pack(loV + kmaj_lo(k16, 16384)),
                    umma_desc_pack(loDO + slot * 1024 + kmaj_lo(k16, 8192)), idesc_s, k16 > 0 ? 1u : 0u);
        umma_commit(&bar_s[buf]);
      }
      TM_ACC(ti_s, ti_b);
    };
    ISSUER_WAIT(bar_kv, 0);
#ifdef DTX_ATTN_TIMING
    const long long ti_kv = clock64() - ti_0;
#endif
    issue_s(0, 0, 0, true);
    if (n > 1) issue_s(1, 1, 0, true);
    for (int base = 0; base < n; base += DKV_NS) {
      const uint32_t rp = (base >> 2) & 1;  // ring round parity of this group of four pairs
#pragma unroll
      for (int u = 0; u < DKV_NS; ++u) {
        const int ii = base + u;
        if (ii < n) {
          // P^T / dS^T of pair ii sit bf16-packed in the score buffers u&1 themselves (the CPT query columns of column group c
          // -> TMEM columns CPT*c .. CPT*c + CPT/2 - 1): the A operands of the accumulating MMAs come straight from tensor memory.
          TM_SET(ti_a);
          ISSUER_WAIT(&bar_p[u & 1], (ii >> 1) & 1);
          tc_fence_after();
          TM_ACC(ti_p, ti_a);
          TM_SET(ti_b);
          if (leader) {
            const uint32_t acc0 = ii > 0 ? 1u : 0u;
            const uint32_t aP = tmem + T_ST + (u & 1) * 64, aDS = tmem + T_DPT + (u & 1) * 64;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16_ts(tmem + T_DV, aP + ((16 * kk) / CPT) * CPT + ((16 * kk) % CPT) / 2, umma_desc_pack(loDOm + u * 1024 + kk * 128),
                           idesc_g, kk > 0 ? 1u : acc0);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16_ts(tmem + T_DK, aDS + ((16 * kk) / CPT) * CPT + ((16 * kk) % CPT) / 2, umma_desc_pack(loQm + u * 1024 + kk * 128),
                           idesc_g, kk > 0 ? 1u : acc0);
            umma_commit(&bar_free[u]);  // ring slot u may be refilled once these have read it
          }
          TM_ACC(ti_acc, ti_b);
          if (ii + 2 < n) issue_s((u + 2) & (DKV_NS - 1), u & 1, (u + 2 >= DKV_NS) ? (rp ^ 1u) : rp, false);
        }
      }
    }
    // the compute warps do not follow the MMA completions phase by phase (parity waits are only safe one phase behind):
    // a dedicated single-phase barrier reports that every MMA of this CTA has completed
    if (leader) umma_commit(bar_fin);
#ifdef DTX_ATTN_TIMING
    if (leader && TM_BLOCK((int)blockIdx.x))
      printf("[dkv issuer] block %d n %d total %lld wait_kv %lld | per pair: wait_p %lld wait_q %lld issue_s %lld issue_acc %lld\n", (int)blockIdx.x, n,
             clock64() - ti_0, ti_kv, ti_p / n, ti_q / n, ti_s / n, ti_acc / n);
#endif
  } else {
    const int rw = warp & 3, cq = warp >> 2;
    const int r = rw * 32 + (tid & 31);
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(rw * 32) << 16);
    const int kvrow = kv0 + r;
    int qi = 0;
    TM_DECL(tc_w = 0, tc_l = 0, tc_m = 0, tc_s = 0, tc_0, tc_a, tc_b, tc_c, tc_d);
    TM_SET(tc_0);
    for (int ii = 0; ii < n; ++ii) {
      const int qs = (i0 + qi) * 64;
      if (++qi == nq) qi = 0;
      TM_SET(tc_a);
      mbar_wait(&bar_q[ii % DKV_NS], (ii / DKV_NS) & 1);  // acquire the TMA-written row statistics of this ring slot
      mbar_wait(&bar_s[ii & 1], (ii >> 1) & 1);
      tc_fence_after();
      TM_SET(tc_b);
      const uint32_t st = smem_u32(smem + DKV_STAT + (ii % DKV_NS) * 512) + cq * (CPT * 4);
      uint32_t sv[CPT], dv[CPT], ppk[CPT / 2], dpk[CPT / 2];
      if constexpr (CPT == 32) {
        tmem_ld32(t_lane + T_ST + (ii & 1) * 64 + cq * CPT, sv);
        tmem_ld32(t_lane + T_DPT + (ii & 1) * 64 + cq * CPT, dv);
      } else {
        tmem_ld16(t_lane + T_ST + (ii & 1) * 64 + cq * CPT, sv);
        tmem_ld16(t_lane + T_DPT + (ii & 1) * 64 + cq * CPT, dv);
      }
      tmem_ld_wait();
      TM_SET(tc_c);
      const bool need_mask = (qs < kv0 + 127) || (WIN && qs + 63 > kv0 + window);
      const int qhi = WIN ? kvrow + window : (1 << 30);  // last query row that sees this key
      // packed fp32 pairs (FFMA2 / FMUL2)
      const float2 sl2 = make_float2(p.scale_log2, p.scale_log2), sc2 = make_float2(p.scale, p.scale), nsc2 = make_float2(-p.scale, -p.scale);
      // two copies of the loop, the mask branch outside (see attn_dq1_kernel)
      auto ds_math = [&](auto masked) {
#pragma unroll
        for (int e = 0; e < CPT; e += 4) {
          const float4 l4 = lds128f(st + e * 4);
          const float4 d4 = lds128f(st + 256 + e * 4);
#pragma unroll
          for (int u = 0; u < 4; u += 2) {
            const float2 nl = u ? make_float2(-l4.z, -l4.w) : make_float2(-l4.x, -l4.y);
            const float2 dl = u ? make_float2(d4.z, d4.w) : make_float2(d4.x, d4.y);
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(sv[e + u]), __uint_as_float(sv[e + u + 1])), sl2, nl);
            float2 pr = exp2_pair(x, (e + u) >> 1);
            if constexpr (decltype(masked)::value) {
              const int qa = qs + cq * CPT + e + u;  // query row of pr.x (pr.y: qa + 1); visible iff kvrow <= q <= kvrow + window
              if (kvrow > qa || (WIN && qa > qhi)) pr.x = 0.f;
              if (kvrow > qa + 1 || (WIN && qa + 1 > qhi)) pr.y = 0.f;
            }
            const float2 dd = __ffma2_rn(make_float2(__uint_as_float(dv[e + u]), __uint_as_float(dv[e + u + 1])), sc2, __fmul2_rn(dl, nsc2));
            const float2 ds = __fmul2_rn(pr, dd);
            ppk[(e + u) >> 1] = pack_bf16x2(pr.x, pr.y);
            dpk[(e + u) >> 1] = pack_bf16x2(ds.x, ds.y);
          }
        }
      };
      if (need_mask) ds_math(std::true_type{}); else ds_math(std::false_type{});
      TM_SET(tc_d);
      // overwrite this thread's own (already loaded) score columns with the packed operands
      if constexpr (CPT == 32) {
        tmem_st16(t_lane + T_ST + (ii & 1) * 64 + cq * CPT, ppk);
        tmem_st16(t_lane + T_DPT + (ii & 1) * 64 + cq * CPT, dpk);
      } else {
        tmem_st8(t_lane + T_ST + (ii & 1) * 64 + cq * CPT, ppk);
        tmem_st8(t_lane + T_DPT + (ii & 1) * 64 + cq * CPT, dpk);
      }
      tmem_st_wait();
      // the issuer starts S^T / dP^T of pair ii + 2 right after the accumulate MMAs of this pair without looking at the ring:
      // make sure that pair's Q / dO block has landed before telling it so (these warps have slack, the issuer has none)
      if (ii + 2 < n) mbar_wait(&bar_q[(ii + 2) % DKV_NS], ((ii + 2) / DKV_NS) & 1);
      tc_fence_before();
      mbar_arrive(&bar_p[ii & 1]);
#ifdef DTX_ATTN_TIMING
      tc_w += tc_b - tc_a; tc_l += tc_c - tc_b; tc_m += tc_d - tc_c; tc_s += clock64() - tc_d;
#endif
    }
#ifdef DTX_ATTN_TIMING
    const long long tc_loop = clock64() - tc_0;
#endif
    mbar_wait(bar_fin, 0);
    tc_fence_after();
#ifdef DTX_ATTN_TIMING
    const long long tc_fin = clock64() - tc_0 - tc_loop;
#endif
    // Hand dV and dK to the TMA: the eight 32-column chunks (dV 0-3, dK 4-7) are dealt to the column groups; each thread
    // converts its part of row r to bf16 into a 128B-swizzled staging tile (the Q/dO ring is free now), one thread per matrix
    // issues two bulk tensor stores.  (Direct 16-byte stores of one row per thread touch 32 different 128-byte lines per
    // instruction: 16 % of the compute warps' time, lg_throttle.)
    constexpr int CHUNKS = 8 / (NCW / 4);
    const int gc0 = cq * CHUNKS, m = gc0 >> 2;
    uint8_t* stage = smem + DKV_SQ + m * 32768;
    const uint32_t stage_addr = smem_u32(stage);
    if (NCW == 8 && m == 1 && p.rope_cs) {
      // dK with the inverse rotary: this thread holds the whole row; rotary partners are chunks c and c + 2
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t lo[32], hi[32];
        tmem_ld32(t_lane + T_DK + c * 32, lo);
        tmem_ld32(t_lane + T_DK + 64 + c * 32, hi);
        tmem_ld_wait();
        const float2* cs = p.rope_cs + static_cast<size_t>(c * 32) * p.rope_stride + kvrow;
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          const float2 tcs = __ldg(cs + static_cast<size_t>(jj) * p.rope_stride);
          const float x0 = __uint_as_float(lo[jj]), x1 = __uint_as_float(hi[jj]);
          lo[jj] = __float_as_uint(x0 * tcs.x + x1 * tcs.y);
          hi[jj] = __float_as_uint(x1 * tcs.x - x0 * tcs.y);
        }
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          sts128(stage_addr + sw128_offset(r, c * 4 + c8),
                 pack_bf16x2(__uint_as_float(lo[c8 * 8 + 0]), __uint_as_float(lo[c8 * 8 + 1])),
                 pack_bf16x2(__uint_as_float(lo[c8 * 8 + 2]), __uint_as_float(lo[c8 * 8 + 3])),
                 pack_bf16x2(__uint_as_float(lo[c8 * 8 + 4]), __uint_as_float(lo[c8 * 8 + 5])),
                 pack_bf16x2(__uint_as_float(lo[c8 * 8 + 6]), __uint_as_float(lo[c8 * 8 + 7])));
          sts128(stage_addr + 16384 + sw128_offset(r, c * 4 + c8),
                 pack_bf16x2(__uint_as_float(hi[c8 * 8 + 0]), __uint_as_float(hi[c8 * 8 + 1])),
                 pack_bf16x2(__uint_as_float(hi[c8 * 8 + 2]), __uint_as_float(hi[c8 * 8 + 3])),
                 pack_bf16x2(__uint_as_float(hi[c8 * 8 + 4]), __uint_as_float(hi[c8 * 8 + 5])),
                 pack_bf16x2(__uint_as_float(hi[c8 * 8 + 6]), __uint_as_float(hi[c8 * 8 + 7])));
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < CHUNKS; ++i) {
        const int c = (gc0 + i) & 3;
        uint32_t v[32];
        tmem_ld32(t_lane + (m ? T_DK : T_DV) + c * 32, v);
        tmem_ld_wait();
  #pragma unroll
        for (int c8 = 0; c8 < 4; ++c8)
          sts128(stage_addr + (c >> 1) * 16384 + sw128_offset(r, (c & 1) * 4 + c8),
                 pack_bf16x2(__uint_as_float(v[c8 * 8 + 0]), __uint_as_float(v[c8 * 8 + 1])),
                 pack_bf16x2(__uint_as_float(v[c8 * 8 + 2]), __uint_as_float(v[c8 * 8 + 3])),
                 pack_bf16x2(__uint_as_float(v[c8 * 8 + 4]), __uint_as_float(v[c8 * 8 + 5])),
                 pack_bf16x2(__uint_as_float(v[c8 * 8 + 6]), __uint_as_float(v[c8 * 8 + 7])));
      }
    }
    fence_proxy_async_smem();
    named_bar_sync(1 + m, NCW * 16);
    if (r == 0 && (gc0 & 3) == 0) {
      const int col = m ? colK : colV;
      tma_store_2d(&tmOut, stage, col, row_base + kv0);
      tma_store_2d(&tmOut, stage + 16384, col + 64, row_base + kv0);
      tma_store_commit();
      tma_store_wait_read0();  // the CTA's shared memory must stay valid until the bulk stores have read it
    }
#ifdef DTX_ATTN_TIMING
    if ((tid == 0 || tid == 128) && TM_BLOCK((int)blockIdx.x))
      printf("[dkv compute] block %d tid %d n %d loop %lld (incl. first wait) wait_fin %lld epilogue %lld | per pair: wait %lld ld %lld math %lld st+arrive %lld\n",
             (int)blockIdx.x, tid, n, tc_loop, tc_fin, clock64() - tc_0 - tc_loop - tc_fin, tc_w / n, tc_l / n, tc_m / n, tc_s / n);
#endif
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

int g_fwd_exp_fma_every = 3;  // measured (profiles/r02_attn_events.txt): 330 us (0) / 301 (4) / 292 (3) / 294 (2) per layer at the 7B shape
int g_dq_exp_fma_every = 0;  // dQ kernel: every N-th pair of exponentials on the FMA pipe (0 = none, 3, 4)
void attn_set_dq_exp_fma_every(int n) { g_dq_exp_fma_every = (n == 3 || n == 4) ? n : 0; }
void attn_set_fwd_exp_fma_every(int n) { g_fwd_exp_fma_every = (n == 2 || n == 3 || n == 4) ? n : 0; }  // 0 = all on the MUFU
int attn_bwd_launches() { return 2; }
bool attn_bwd_can_rope() { return true; }

namespace {
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device property of a function: set it once per device
cudaError_t attn_init_device() {
  static std::mutex mu;
  static bool done[64] = {false};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(mu);
  if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
  const void* fwd[] = {reinterpret_cast<const void*>(attn_fwd2_kernel<0, false>), reinterpret_cast<const void*>(attn_fwd2_kernel<2, false>),
                       reinterpret_cast<const void*>(attn_fwd2_kernel<3, false>), reinterpret_cast<const void*>(attn_fwd2_kernel<4, false>),
                       reinterpret_cast<const void*>(attn_fwd2_kernel<0, true>), reinterpret_cast<const void*>(attn_fwd2_kernel<2, true>),
                       reinterpret_cast<const void*>(attn_fwd2_kernel<3, true>), reinterpret_cast<const void*>(attn_fwd2_kernel<4, true>)};
  for (const void* f : fwd)
    if ((e = set_smem(f, FW2_SMEM)) != cudaSuccess) return e;
  const void* dq[] = {reinterpret_cast<const void*>(attn_dq1_kernel<false, 0>), reinterpret_cast<const void*>(attn_dq1_kernel<false, 3>),
                      reinterpret_cast<const void*>(attn_dq1_kernel<false, 4>), reinterpret_cast<const void*>(attn_dq1_kernel<true, 0>),
                      reinterpret_cast<const void*>(attn_dq1_kernel<true, 3>), reinterpret_cast<const void*>(attn_dq1_kernel<true, 4>)};
  for (const void* f : dq)
    if ((e = set_smem(f, DQ1_SMEM)) != cudaSuccess) return e;
  if ((e = set_smem(reinterpret_cast<const void*>(attn_dkv_kernel<8, false>), DKV_SMEM)) != cudaSuccess) return e;
  if ((e = set_smem(reinterpret_cast<const void*>(attn_dkv_kernel<8, true>), DKV_SMEM)) != cudaSuccess) return e;
  if (dev >= 0 && dev < 64) done[dev] = true;
  return cudaSuccess;
}
}  // namespace

cudaError_t attn_fwd(const AttnArgs& a, cudaStream_t s) {
  if (a.S % 128 || a.B <= 0 || a.H <= 0) return cudaErrorInvalidValue;
  if (a.window < 0) return cudaErrorInvalidValue;
  cudaError_t e = attn_init_device();
  if (e != cudaSuccess) return e;
  const int Hkv = a.Hkv > 0 ? a.Hkv : a.H;
  if (a.H % Hkv) return cudaErrorInvalidValue;
  if (a.row_start && (a.total_rows <= 0 || a.total_rows % 128)) return cudaErrorInvalidValue;
  const uint64_t M = a.row_start ? static_cast<uint64_t>(a.total_rows) : static_cast<uint64_t>(a.B) * a.S, W = static_cast<uint64_t>(a.H + 2 * Hkv) * HD;
  CUtensorMap tmQ, tmKV, tmOut;
  const uint64_t WO = static_cast<uint64_t>(a.H) * HD;
  if (!make_tmap_2d_bf16(&tmQ, a.qkv, W, M, W, 64, 128) || !make_tmap_2d_bf16(&tmKV, a.qkv, W, M, W, 64, 64) ||
      !make_tmap_2d_bf16(&tmOut, a.out, WO, M, WO, 64, 128))
    return cudaErrorInvalidValue;
  AttnKParams p{};
  p.B = a.B; p.S = a.S; p.H = a.H;
  p.Hkv = Hkv; p.W = static_cast<int>(W); p.colK0 = a.H * HD; p.colV0 = (a.H + Hkv) * HD;
  p.scale = a.scale;
  p.scale_log2 = a.scale * LOG2E;
  p.lse2 = a.lse;
  p.out = a.out;
  p.seq_lens = a.seq_lens;
  p.window = a.window;
  p.row_start = a.row_start;
  const int grid = a.B * a.H * ((a.S + 255) / 256);
#define DTX_FWD_LAUNCH(E)                                                                          \
  if (a.window > 0) attn_fwd2_kernel<E, true><<<grid, FW2_THREADS, FW2_SMEM, s>>>(tmQ, tmKV, tmOut, p); \
  else attn_fwd2_kernel<E, false><<<grid, FW2_THREADS, FW2_SMEM, s>>>(tmQ, tmKV, tmOut, p)
  switch (g_fwd_exp_fma_every) {
    case 2: DTX_FWD_LAUNCH(2); break;
    case 3: DTX_FWD_LAUNCH(3); break;
    case 4: DTX_FWD_LAUNCH(4); break;
    default: DTX_FWD_LAUNCH(0); break;
  }
#undef DTX_FWD_LAUNCH
  return cudaGetLastError();
}

cudaError_t attn_bwd(const AttnArgs& a, cudaStream_t s) {
  if (a.S % 128 || a.B <= 0 || a.H <= 0 || !a.delta || !a.lse || !a.dout || !a.dqkv) return cudaErrorInvalidValue;
  if (a.window < 0) return cudaErrorInvalidValue;
  cudaError_t e = attn_init_device();
  if (e != cudaSuccess) return e;
  const int Hkv = a.Hkv > 0 ? a.Hkv : a.H;
  if (a.H % Hkv) return cudaErrorInvalidValue;
  if (a.row_start && (a.total_rows <= 0 || a.total_rows % 128)) return cudaErrorInvalidValue;
  const uint64_t M = a.row_start ? static_cast<uint64_t>(a.total_rows) : static_cast<uint64_t>(a.B) * a.S, W = static_cast<uint64_t>(a.H + 2 * Hkv) * HD,
                 WO = static_cast<uint64_t>(a.H) * HD;
  CUtensorMap tmQ128, tmKV64, tmDO128, tmKV128, tmQ64, tmDO64, tmDqkv, tmO128;
  if (!make_tmap_2d_bf16(&tmO128, a.out, WO, M, WO, 64, 128)) return cudaErrorInvalidValue;
  bool ok = make_tmap_2d_bf16(&tmQ128, a.qkv, W, M, W, 64, 128) && make_tmap_2d_bf16(&tmKV64, a.qkv, W, M, W, 64, 64) &&
            make_tmap_2d_bf16(&tmDqkv, a.dqkv, W, M, W, 64, 128) &&
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
  p.delta_w = a.delta;
  p.dqkv = a.dqkv;
  p.rope_cs = a.rope_cs;
  p.rope_stride = a.rope_stride > 0 ? a.rope_stride : a.S;
  p.seq_lens = a.seq_lens;
  p.window = a.window;
  p.row_start = a.row_start;
  const int gq = a.B * a.H * (a.S / 128), gkv = a.B * Hkv * (a.S / 128);
  const bool win = a.window > 0;
#define DTX_DQ_LAUNCH(W, I) attn_dq1_kernel<W, I><<<gq, DKV_THREADS, DQ1_SMEM, s>>>(tmQ128, tmKV64, tmDO128, tmO128, tmDqkv, p)
  if (win) { if (g_dq_exp_fma_every == 3) DTX_DQ_LAUNCH(true, 3); else if (g_dq_exp_fma_every == 4) DTX_DQ_LAUNCH(true, 4); else DTX_DQ_LAUNCH(true, 0); }
  else { if (g_dq_exp_fma_every == 3) DTX_DQ_LAUNCH(false, 3); else if (g_dq_exp_fma_every == 4) DTX_DQ_LAUNCH(false, 4); else DTX_DQ_LAUNCH(false, 0); }
#undef DTX_DQ_LAUNCH
  if (win) attn_dkv_kernel<8, true><<<gkv, 10 * 32, DKV_SMEM, s>>>(tmKV128, tmQ64, tmDO64, tmDqkv, p);
  else attn_dkv_kernel<8, false><<<gkv, 10 * 32, DKV_SMEM, s>>>(tmKV128, tmQ64, tmDO64, tmDqkv, p);
  return cudaGetLastError();
}

}  // namespace dtx
