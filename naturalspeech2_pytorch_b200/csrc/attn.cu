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
// 64-key tiles; S (TMEM), P (smem) and O_j (TMEM) are double-buffered so S_{j+1} and S_{j+2} are computed while the
// softmax warps process tile j.  Two CTAs are co-resident per SM.
#include "ptx.cuh"
#include "host_common.h"
#include "../../include/ns2_b200.h"

#include <atomic>

namespace ns2 {

extern std::atomic<long long> g_launches;

namespace attn {
constexpr int BQ = 128;   // queries per CTA
constexpr int BKV = 64;   // keys per tile
constexpr int DH = 64;
constexpr int KVS = 4;    // K/V ring depth
constexpr int Q_BYTES = BQ * DH * 2;         // 16 KB
constexpr int KV_BYTES = BKV * DH * 2;       // 8 KB each for K and V
constexpr int P_BYTES = BQ * BKV * 2;        // 16 KB: one 64-key swizzle atom
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + Q_BYTES;                 // KVS stages
constexpr int OFF_V = OFF_K + KVS * KV_BYTES;          // KVS stages
constexpr int OFF_P = OFF_V + KVS * KV_BYTES;          // 2 buffers
constexpr int OFF_BAR = OFF_P + 2 * P_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256;              // 112.25 KB -> two CTAs per SM
constexpr int TMEM_COLS = 256;                         // two CTAs per SM share the 512 columns
constexpr int TM_S = 0;     // S buffers at columns 0 and 64
constexpr int TM_O = 128;   // O_j buffers at columns 128 and 192
}  // namespace attn

struct AttnDev {
  CUtensorMap tmQ, tmK, tmV;
  __nv_bfloat16* out;
  long long o_rs, o_bs;
  int q_len, kv_len;
  float scale_log2e;
};

// bf16 pair from two non-negative finite floats with round-half-up done on the integer pipe (IADD + PRMT): the
// F2FP conversion shares the 16-lane XU pipe with ex2, which bounds the softmax throughput.
__device__ __forceinline__ uint32_t pack_bf16x2_pos(float lo, float hi) {
  return __byte_perm(__float_as_uint(lo) + 0x8000u, __float_as_uint(hi) + 0x8000u, 0x7632);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// One-pass online-softmax flash attention on 64-key tiles.  S and P are double-buffered (TMEM / smem) so that the
// tensor core computes S_{j+1}, S_{j+2} while the softmax warps work on tile j and never waits for them in steady
// state; the per-tile P.V product is folded into the register-resident output one tile late.
// Two CTAs are resident per SM (112 KB smem, 256 TMEM columns each).
__global__ void __launch_bounds__(192, 2) attn_fwd_kernel(const __grid_constant__ AttnDev p) {
  using namespace attn;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;            // [KVS]
  uint64_t* kv_empty = bars + 1 + KVS;     // [KVS]
  uint64_t* s_full = bars + 1 + 2 * KVS;   // [2]
  uint64_t* p_full = bars + 3 + 2 * KVS;   // [2] 128 arrivals: P_j written and S_j consumed
  uint64_t* o_full = bars + 5 + 2 * KVS;   // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 7 + 2 * KVS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int T = (p.kv_len + BKV - 1) / BKV;

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
    for (int i = 0; i < KVS; ++i) {
      mbar_init(smem_u32(&kv_full[i]), 1);
      mbar_init(smem_u32(&kv_empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&s_full[i]), 1);
      mbar_init(smem_u32(&p_full[i]), 128);
      mbar_init(smem_u32(&o_full[i]), 1);
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
      for (int j = 0; j < T; ++j) {
        const int st = j % KVS;
        const uint32_t ph = (j / KVS) & 1;
        mbar_wait(smem_u32(&kv_empty[st]), ph ^ 1);
        const uint32_t fb = smem_u32(&kv_full[st]);
        mbar_arrive_expect_tx(fb, 2 * KV_BYTES);
        tma_load_3d(smem_u32(smem + OFF_K + st * KV_BYTES), &p.tmK, fb, head * DH, j * BKV, b);
        tma_load_3d(smem_u32(smem + OFF_V + st * KV_BYTES), &p.tmV, fb, head * DH, j * BKV, b);
      }
    }
  } else if (warp == 5) {
    // ================================ MMA issuer ==================================
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_f16(BQ, BKV, 1, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_f16(BQ, DH, 1, 0, /*V is MN-major*/ 1);
      const uint64_t dq = umma_desc_sw128(smem_u32(smem + OFF_Q), 16, 1024);
      auto issue_s = [&](int j) {
        const int st = j % KVS;
        mbar_wait(smem_u32(&kv_full[st]), (j / KVS) & 1);
        tc_fence_after();
        const uint64_t dk = umma_desc_sw128(smem_u32(smem + OFF_K + st * KV_BYTES), 16, 1024);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          tc_mma_f16(tmem_base + TM_S + (j & 1) * BKV, dq + 2 * k, dk + 2 * k, idesc_s, k > 0);
        tc_commit(smem_u32(&s_full[j & 1]));
      };
      mbar_wait(smem_u32(q_full), 0);
      issue_s(0);
      if (T > 1) issue_s(1);
      for (int j = 0; j < T; ++j) {
        const int bsel = j & 1;
        const int st = j % KVS;
        // P_j is in smem and the softmax warps are done reading S_j
        mbar_wait(smem_u32(&p_full[bsel]), (j >> 1) & 1);
        tc_fence_after();
        const uint32_t pbase = smem_u32(smem + OFF_P + bsel * P_BYTES);
        const uint32_t vbase = smem_u32(smem + OFF_V + st * KV_BYTES);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          // A = P: K-major, one 64-key atom, 32 bytes per 16-key step
          const uint64_t dp = umma_desc_sw128(pbase + k * 32, 16, 1024);
          // B = V: MN-major (64 dh contiguous per key row of 128 B); 16 keys = 2048 bytes per step
          const uint64_t dv = umma_desc_sw128(vbase + k * 2048, 1024, 1024);
          tc_mma_f16(tmem_base + TM_O + bsel * DH, dp, dv, idesc_o, k > 0);
        }
        tc_commit(smem_u32(&o_full[bsel]));
        tc_commit(smem_u32(&kv_empty[st]));
        if (j + 2 < T) issue_s(j + 2);  // S buffer `bsel` is free again
      }
    }
  } else {
    // ================================ softmax warps ===============================
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const float c = p.scale_log2e;
    float o_acc[DH];
#pragma unroll
    for (int i = 0; i < DH; ++i) o_acc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, a_prev = 0.f;

    auto accumulate_o = [&](int jprev, float a) {
      const int bs = jprev & 1;
      mbar_wait(smem_u32(&o_full[bs]), (jprev >> 1) & 1);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      tmem_ld32(lane_addr + TM_O + bs * DH, r0);
      tmem_ld32(lane_addr + TM_O + bs * DH + 32, r1);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        o_acc[i] = fmaf(o_acc[i], a, __uint_as_float(r0[i]));
        o_acc[32 + i] = fmaf(o_acc[32 + i], a, __uint_as_float(r1[i]));
      }
    };

    for (int j = 0; j < T; ++j) {
      const int bsel = j & 1;
      mbar_wait(smem_u32(&s_full[bsel]), (j >> 1) & 1);
      tc_fence_after();
      const int valid = p.kv_len - j * BKV;  // columns >= valid are padding keys
      const bool full = valid >= BKV;
      const uint32_t s_addr = lane_addr + TM_S + bsel * BKV;
      // S_j is read from TMEM exactly once (TMEM reads run at ~64 B/clk per SM and were the bottleneck of the
      // two-pass version): 64 scores stay in registers for both the maximum and the exponentials
      uint32_t r0[32], r1[32];
      tmem_ld32(s_addr, r0);
      tmem_ld32(s_addr + 32, r1);
      tmem_ld_wait();
      float m_tile;
      {
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
        if (full) {
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
            if (i < valid) m0 = fmaxf(m0, __uint_as_float(r0[i]));
            if (32 + i < valid) m1 = fmaxf(m1, __uint_as_float(r1[i]));
          }
        }
        m_tile = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      }
      const float m_new = fmaxf(m_run, m_tile * c);
      const float a = ex2_approx(m_run - m_new);  // 0 on the first tile (m_run = -inf)
      // probabilities -> bf16 -> swizzled smem; row sum in fp32 (4 independent chains)
      // (P buffer `bsel` was last read by P.V of tile j-2, whose completion was awaited in iteration j-1)
      float l_tile;
      {
        uint8_t* prow = smem + OFF_P + bsel * P_BYTES + row * 128;
        float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
        auto chunk32 = [&](const uint32_t (&r)[32], int cc) {
          uint32_t pk[16];
          if (full) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              const float p0 = ex2_approx(fmaf(__uint_as_float(r[2 * i]), c, -m_new));
              const float p1 = ex2_approx(fmaf(__uint_as_float(r[2 * i + 1]), c, -m_new));
              const float p2 = ex2_approx(fmaf(__uint_as_float(r[2 * i + 2]), c, -m_new));
              const float p3 = ex2_approx(fmaf(__uint_as_float(r[2 * i + 3]), c, -m_new));
              l0 += p0; l1 += p1; l2 += p2; l3 += p3;
              pk[i] = pack_bf16x2_pos(p0, p1);
              pk[i + 1] = pack_bf16x2_pos(p2, p3);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int c0 = cc * 32 + 2 * i;
              const float p0 = (c0 < valid) ? ex2_approx(fmaf(__uint_as_float(r[2 * i]), c, -m_new)) : 0.f;
              const float p1 =
                  (c0 + 1 < valid) ? ex2_approx(fmaf(__uint_as_float(r[2 * i + 1]), c, -m_new)) : 0.f;
              l0 += p0; l1 += p1;
              pk[i] = pack_bf16x2_pos(p0, p1);
            }
          }
          // 32 columns = 4 chunks of 16 bytes of the 128-byte row; chunk index XOR-swizzled with row&7
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int chunk = cc * 4 + q;
            *reinterpret_cast<uint4*>(prow + ((chunk ^ (row & 7)) << 4)) =
                make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          }
        };
        chunk32(r0, 0);
        chunk32(r1, 1);
        l_tile = (l0 + l1) + (l2 + l3);
      }
      l_run = fmaf(l_run, a, l_tile);
      m_run = m_new;
      // publish P_j: generic-proxy writes -> async proxy, TMEM reads of S_j ordered before the next MMA
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(smem_u32(&p_full[bsel]));
      // fold in the previous tile's P.V while the tensor core works on this one
      if (j > 0) accumulate_o(j - 1, a_prev);
      a_prev = a;
    }
    accumulate_o(T - 1, a_prev);

    if (q0 + row < p.q_len) {
      const float inv = 1.0f / l_run;
      __nv_bfloat16* op = p.out + static_cast<long long>(b) * p.o_bs +
                          static_cast<long long>(q0 + row) * p.o_rs + head * DH;
      uint4* o4 = reinterpret_cast<uint4*>(op);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint4 w;
        w.x = pack_bf16x2(o_acc[8 * i + 0] * inv, o_acc[8 * i + 1] * inv);
        w.y = pack_bf16x2(o_acc[8 * i + 2] * inv, o_acc[8 * i + 3] * inv);
        w.z = pack_bf16x2(o_acc[8 * i + 4] * inv, o_acc[8 * i + 5] * inv);
        w.w = pack_bf16x2(o_acc[8 * i + 6] * inv, o_acc[8 * i + 7] * inv);
        o4[i] = w;
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
  const uint32_t box[3] = {64, attn::BQ, 1};
  const uint32_t box_kv[3] = {64, attn::BKV, 1};
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
    int rc = make_tmap_16bit(&dev.tmK, a->k, 3, dims, strk, box_kv);
    if (rc != kOk) return rc;
    rc = make_tmap_16bit(&dev.tmV, a->v, 3, dims, strv, box_kv);
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
