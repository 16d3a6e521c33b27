// Flash attention forward on tcgen05 for sm_100a (non-causal, unmasked, dim_head = 64).
// Replaces Attend.forward / F.scaled_dot_product_attention (attend.py:77-155) for the only configuration
// the denoiser uses: mask=None, causal=False, dropout=0.
//
// One CTA per (batch, head, 128-query tile); 192 threads:
//   warps 0-3  softmax: thread r owns query row r (TMEM lane r): tcgen05.ld S -> online softmax in fp32 ->
//              P (bf16) written to shared memory in the 128B-swizzled K-major layout UMMA expects ->
//              running output kept in registers, rescaled and accumulated from the per-tile P.V product
//   warp 4     TMA producer: Q once, then K_j / V_j tiles (128 keys x 64) through a 2-stage ring
//   warp 5     tcgen05.mma issuer: S_j = Q.K_j^T (both K-major), O_j = P_j.V_j (V is the MN-major operand:
//              its rows are keys = the reduction dimension, so no transpose of V is ever materialised)
// O_j is double-buffered in TMEM; S_{j+1} is issued as soon as the softmax warps have consumed S_j, ahead of
// P_j.V_j.  Two CTAs are co-resident per SM so one CTA's softmax overlaps the other's MMAs.
#include "ptx.cuh"
#include "host_common.h"
#include "../../include/ns2_b200.h"

#include <atomic>

namespace ns2 {

extern std::atomic<long long> g_launches;

namespace attn {
constexpr int BQ = 128;   // queries per CTA
constexpr int BKV = 128;  // keys per tile
constexpr int DH = 64;
constexpr int Q_BYTES = BQ * DH * 2;         // 16 KB
constexpr int KV_BYTES = BKV * DH * 2;       // 16 KB each for K and V
constexpr int P_BYTES = BQ * BKV * 2;        // 32 KB (two 64-key swizzle atoms of 16 KB)
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + Q_BYTES;                 // 2 stages
constexpr int OFF_V = OFF_K + 2 * KV_BYTES;            // 2 stages
constexpr int OFF_P = OFF_V + 2 * KV_BYTES;            // single buffer
constexpr int OFF_BAR = OFF_P + P_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256;              // 112.25 KB -> two CTAs per SM
constexpr int TMEM_COLS = 256;                         // two CTAs per SM share the 512 columns
constexpr int TM_S = 0;     // S at columns [0, 128)
constexpr int TM_O = 128;   // O accumulator at columns [128, 192)
}  // namespace attn

struct AttnDev {
  CUtensorMap tmQ, tmK, tmV;
  __nv_bfloat16* out;
  long long o_rs, o_bs;
  int q_len, kv_len;
  float scale_log2e;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Two-pass flash attention.  Pass A computes only the exact row maxima (S = Q.K^T, running max, nothing else);
// pass B recomputes S tile by tile, forms P = exp2(S*c - m) against the FINAL maximum and lets the tensor core
// accumulate O += P.V in TMEM across all key tiles.  No running rescale, no per-tile read-back of partial outputs,
// one exp pass: ~600 issue slots per 128x128 tile per warp instead of ~1050, at the price of issuing Q.K^T twice
// (tensor time stays below the MUFU time of the exponentials).
// Two CTAs are resident per SM (112 KB smem, 256 TMEM columns, <=170 registers each): while one CTA's softmax
// warps occupy the MUFU/FMA pipes, the other CTA's MMAs occupy the tensor core.
__global__ void __launch_bounds__(192, 2) attn_fwd_kernel(const __grid_constant__ AttnDev p) {
  using namespace attn;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;    // MMA -> softmax: S tile ready
  uint64_t* s_done = bars + 6;    // softmax -> MMA: S tile consumed (and, in pass B, P tile written); 128 arrivals
  uint64_t* pv_done = bars + 7;   // MMA -> softmax: P.V retired (P buffer reusable; after the last tile: O complete)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int T = (p.kv_len + BKV - 1) / BKV;
  // global tile counter g: pass A = [0, T), pass B = [T, 2T); key tile j = g mod T

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("ns2 attn: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
  }
  if (warp == 5 && lane == 0) {
    mbar_init(smem_u32(q_full), 1);
    mbar_init(smem_u32(s_full), 1);
    mbar_init(smem_u32(s_done), 128);
    mbar_init(smem_u32(pv_done), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&kv_full[i]), 1);
      mbar_init(smem_u32(&kv_empty[i]), 1);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_holder), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 4) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      mbar_arrive_expect_tx(smem_u32(q_full), Q_BYTES);
      tma_load_3d(smem_u32(smem + OFF_Q), &p.tmQ, smem_u32(q_full), head * DH, q0, b);
      for (int g = 0; g < 2 * T; ++g) {
        const int st = g & 1;
        const uint32_t ph = (g >> 1) & 1;
        const bool pass_b = g >= T;
        const int j = pass_b ? g - T : g;
        mbar_wait(smem_u32(&kv_empty[st]), ph ^ 1);
        const uint32_t fb = smem_u32(&kv_full[st]);
        mbar_arrive_expect_tx(fb, pass_b ? 2 * KV_BYTES : KV_BYTES);
        tma_load_3d(smem_u32(smem + OFF_K + st * KV_BYTES), &p.tmK, fb, head * DH, j * BKV, b);
        if (pass_b) tma_load_3d(smem_u32(smem + OFF_V + st * KV_BYTES), &p.tmV, fb, head * DH, j * BKV, b);
      }
    }
  } else if (warp == 5) {
    // ================================ MMA issuer ==================================
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_f16(BQ, BKV, 1, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_f16(BQ, DH, 1, 0, /*V is MN-major*/ 1);
      const uint64_t dq = umma_desc_sw128(smem_u32(smem + OFF_Q), 16, 1024);
      const uint32_t pbase = smem_u32(smem + OFF_P);
      auto issue_s = [&](int g) {
        const int st = g & 1;
        mbar_wait(smem_u32(&kv_full[st]), (g >> 1) & 1);
        tc_fence_after();
        const uint64_t dk = umma_desc_sw128(smem_u32(smem + OFF_K + st * KV_BYTES), 16, 1024);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          tc_mma_f16(tmem_base + TM_S, dq + 2 * k, dk + 2 * k, idesc_s, k > 0);
        tc_commit(smem_u32(s_full));
        if (g < T) tc_commit(smem_u32(&kv_empty[st]));  // pass A: the K tile is free once S is computed
      };
      mbar_wait(smem_u32(q_full), 0);
      issue_s(0);
      for (int g = 0; g < 2 * T; ++g) {
        // the softmax warps are done with S_g (pass B: and P_g is in shared memory)
        mbar_wait(smem_u32(s_done), g & 1);
        tc_fence_after();
        if (g + 1 < 2 * T) issue_s(g + 1);  // next S runs ahead of this tile's P.V
        if (g >= T) {
          const int st = g & 1;
          const uint32_t vbase = smem_u32(smem + OFF_V + st * KV_BYTES);
#pragma unroll
          for (int k = 0; k < BKV / 16; ++k) {
            // A = P: K-major, 64-key atoms of 16 KB, 32 bytes per 16-key step inside an atom
            const uint64_t dp = umma_desc_sw128(pbase + (k >> 2) * (BQ * 128) + (k & 3) * 32, 16, 1024);
            // B = V: MN-major (64 dh contiguous per key row of 128 B); 16 keys = 2048 bytes per step
            const uint64_t dv = umma_desc_sw128(vbase + k * 2048, 1024, 1024);
            tc_mma_f16(tmem_base + TM_O, dp, dv, idesc_o, (g > T) | (k > 0));  // O accumulates over all key tiles
          }
          tc_commit(smem_u32(pv_done));
          tc_commit(smem_u32(&kv_empty[st]));
        }
      }
    }
  } else {
    // ================================ softmax warps ===============================
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const float c = p.scale_log2e;
    uint8_t* prow = smem + OFF_P + row * 128;

    // ---- pass A: exact row maximum over all keys ----
    float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
    for (int g = 0; g < T; ++g) {
      mbar_wait(smem_u32(s_full), g & 1);
      tc_fence_after();
      const int valid = p.kv_len - g * BKV;  // columns >= valid are padding keys
#pragma unroll 1
      for (int cc = 0; cc < BKV / 64; ++cc) {
        uint32_t r0[32], r1[32];
        tmem_ld32(lane_addr + TM_S + cc * 64, r0);
        tmem_ld32(lane_addr + TM_S + cc * 64 + 32, r1);
        tmem_ld_wait();
        if (valid >= BKV) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            m0 = fmaxf(m0, __uint_as_float(r0[i]));
            m1 = fmaxf(m1, __uint_as_float(r0[i + 1]));
            m2 = fmaxf(m2, __uint_as_float(r1[i]));
            m3 = fmaxf(m3, __uint_as_float(r1[i + 1]));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (cc * 64 + i < valid) m0 = fmaxf(m0, __uint_as_float(r0[i]));
            if (cc * 64 + 32 + i < valid) m1 = fmaxf(m1, __uint_as_float(r1[i]));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(s_done));
    }
    const float mc = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * c;  // scaled maximum (log2 domain)

    // ---- pass B: probabilities against the final maximum; P.V accumulates in TMEM ----
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
    for (int g = T; g < 2 * T; ++g) {
      const int j = g - T;
      mbar_wait(smem_u32(s_full), g & 1);
      tc_fence_after();
      if (j > 0) mbar_wait(smem_u32(pv_done), (j - 1) & 1);  // previous P.V retired: the P buffer is free
      const int valid = p.kv_len - j * BKV;
      const bool full = valid >= BKV;
#pragma unroll 1
      for (int cc = 0; cc < BKV / 32; ++cc) {
        uint32_t r[32];
        tmem_ld32(lane_addr + TM_S + cc * 32, r);
        tmem_ld_wait();
        uint32_t pk[16];
        if (full) {
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(r[2 * i]), c, -mc));
            const float p1 = ex2_approx(fmaf(__uint_as_float(r[2 * i + 1]), c, -mc));
            const float p2 = ex2_approx(fmaf(__uint_as_float(r[2 * i + 2]), c, -mc));
            const float p3 = ex2_approx(fmaf(__uint_as_float(r[2 * i + 3]), c, -mc));
            l0 += p0; l1 += p1; l2 += p2; l3 += p3;
            pk[i] = pack_bf16x2(p0, p1);
            pk[i + 1] = pack_bf16x2(p2, p3);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c0 = cc * 32 + 2 * i;
            const float p0 = (c0 < valid) ? ex2_approx(fmaf(__uint_as_float(r[2 * i]), c, -mc)) : 0.f;
            const float p1 = (c0 + 1 < valid) ? ex2_approx(fmaf(__uint_as_float(r[2 * i + 1]), c, -mc)) : 0.f;
            l0 += p0; l1 += p1;
            pk[i] = pack_bf16x2(p0, p1);
          }
        }
        // 32 columns = 4 chunks of 16 bytes; chunk index within the 64-key atom is XOR-swizzled with row&7
        uint8_t* atom = prow + (cc >> 1) * (BQ * 128);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = (cc & 1) * 4 + q;
          *reinterpret_cast<uint4*>(atom + ((chunk ^ (row & 7)) << 4)) =
              make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
      }
      // publish P_j: generic-proxy writes -> async proxy, TMEM reads of S ordered before the next MMA
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(smem_u32(s_done));
    }
    // ---- epilogue: O / l ----
    mbar_wait(smem_u32(pv_done), (T - 1) & 1);
    tc_fence_after();
    const float inv = 1.0f / ((l0 + l1) + (l2 + l3));
    uint32_t o0[32], o1[32];
    tmem_ld32(lane_addr + TM_O, o0);
    tmem_ld32(lane_addr + TM_O + 32, o1);
    tmem_ld_wait();
    if (q0 + row < p.q_len) {
      __nv_bfloat16* op = p.out + static_cast<long long>(b) * p.o_bs +
                          static_cast<long long>(q0 + row) * p.o_rs + head * DH;
      uint4* o4 = reinterpret_cast<uint4*>(op);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 w;
        w.x = pack_bf16x2(__uint_as_float(o0[8 * i + 0]) * inv, __uint_as_float(o0[8 * i + 1]) * inv);
        w.y = pack_bf16x2(__uint_as_float(o0[8 * i + 2]) * inv, __uint_as_float(o0[8 * i + 3]) * inv);
        w.z = pack_bf16x2(__uint_as_float(o0[8 * i + 4]) * inv, __uint_as_float(o0[8 * i + 5]) * inv);
        w.w = pack_bf16x2(__uint_as_float(o0[8 * i + 6]) * inv, __uint_as_float(o0[8 * i + 7]) * inv);
        o4[i] = w;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 w;
        w.x = pack_bf16x2(__uint_as_float(o1[8 * i + 0]) * inv, __uint_as_float(o1[8 * i + 1]) * inv);
        w.y = pack_bf16x2(__uint_as_float(o1[8 * i + 2]) * inv, __uint_as_float(o1[8 * i + 3]) * inv);
        w.z = pack_bf16x2(__uint_as_float(o1[8 * i + 4]) * inv, __uint_as_float(o1[8 * i + 5]) * inv);
        w.w = pack_bf16x2(__uint_as_float(o1[8 * i + 6]) * inv, __uint_as_float(o1[8 * i + 7]) * inv);
        o4[4 + i] = w;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace ns2

extern "C" int ns2_attn_fwd(const ns2_attn_args* a, ns2_stream_t stream_) {
  using namespace ns2;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NS2_REQUIRE(a != nullptr && a->q && a->k && a->v && a->out, "attn_fwd: NULL pointer");
  NS2_REQUIRE(a->dim_head == 64, "attn_fwd: dim_head=%d, only 64 is supported", a->dim_head);
  NS2_REQUIRE(a->batches > 0 && a->heads > 0 && a->q_len > 0 && a->kv_len > 0, "attn_fwd: empty problem");
  NS2_REQUIRE(a->o_row_stride % 8 == 0 && a->o_batch_stride % 8 == 0 &&
                  (reinterpret_cast<uintptr_t>(a->out) & 15) == 0,
              "attn_fwd: out must be 16-byte aligned with strides multiple of 8");
  AttnDev dev;
  memset(&dev, 0, sizeof(dev));
  const uint32_t box[3] = {64, 128, 1};
  {
    const uint64_t dims[3] = {(uint64_t)a->heads * 64, (uint64_t)a->q_len, (uint64_t)a->batches};
    const uint64_t str[3] = {2, (uint64_t)a->q_row_stride * 2, (uint64_t)a->q_batch_stride * 2};
    int rc = make_tmap_16bit(&dev.tmQ, a->q, 3, dims, str, box);
    if (rc != kOk) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)a->heads * 64, (uint64_t)a->kv_len, (uint64_t)a->batches};
    const uint64_t strk[3] = {2, (uint64_t)a->k_row_stride * 2, (uint64_t)a->k_batch_stride * 2};
    const uint64_t strv[3] = {2, (uint64_t)a->v_row_stride * 2, (uint64_t)a->v_batch_stride * 2};
    int rc = make_tmap_16bit(&dev.tmK, a->k, 3, dims, strk, box);
    if (rc != kOk) return rc;
    rc = make_tmap_16bit(&dev.tmV, a->v, 3, dims, strv, box);
    if (rc != kOk) return rc;
  }
  dev.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  dev.o_rs = a->o_row_stride;
  dev.o_bs = a->o_batch_stride;
  dev.q_len = a->q_len;
  dev.kv_len = a->kv_len;
  dev.scale_log2e = a->scale * 1.4426950408889634f;

  static bool configured = false;
  if (!configured) {
    NS2_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        attn::SMEM_BYTES));
    configured = true;
  }
  dim3 grid((a->q_len + attn::BQ - 1) / attn::BQ, a->heads, a->batches);
  attn_fwd_kernel<<<grid, 192, attn::SMEM_BYTES, stream>>>(dev);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  NS2_CUDA_CHECK(cudaGetLastError());
  return kOk;
}
