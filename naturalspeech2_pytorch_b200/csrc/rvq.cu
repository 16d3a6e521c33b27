// Residual vector quantisation (Encodec RVQ encode / decode) for sm_100a.
//
// Reference behaviour (third-party code behind audiolm_pytorch.EncodecWrapper, reached from ns2.py:1445,1611;
// restated in oracle/rvq_oracle.py from encodec's EuclideanCodebook.quantize + ResidualVectorQuantization):
//   for q in 0..Q-1:  idx = argmin_k ||r - C_q[k]||^2 (first minimum wins);  r -= C_q[idx]
//
// The distance contraction runs on tcgen05 tensor cores in fp16 (fp32 accumulate) as a *filter*:
//   s~_k = ||c_k||^2 - 2 r.c_k     with a rigorous bound |s~_k - s_k| <= E = 2 * 1.05 * 2^-10 * ||r|| * max_k ||c_k||
// Every code whose approximate score is within 2E of the approximate minimum is a candidate; candidates
// are re-scored exactly in fp64 from the fp32 operands, so the emitted index is the exact argmin
// (ties -> lowest index) — bit-exact against the fp64 oracle — while >99% of the flops stay on tensor cores.
//
// One CTA per 128 frames, 192 threads:
//   warps 0-3  "row" threads: thread t owns frame t; its fp32 residual (128 values) lives in shared memory
//              (transposed, so every access is bank-conflict free) for all Q stages; builds the fp16 A tile
//              in swizzled smem, scans the TMEM score tiles keeping the 4 best codes, re-scores near-ties
//              in fp64, subtracts the chosen fp32 codeword
//   warp 4     TMA producer: streams the fp16 codebooks (128 codes x 128 dims per chunk) through a 3-deep ring
//   warp 5     tcgen05.mma issuer: D[128 frames x 128 codes] per chunk, double-buffered in TMEM
#include "ptx.cuh"
#include "host_common.h"
#include "../../include/ns2_b200.h"

#include <atomic>

namespace ns2 {

extern std::atomic<long long> g_launches;

namespace rvq {
constexpr int D = 128;        // latent dimension
constexpr int BF = 128;       // frames per CTA
constexpr int BC = 128;       // codes per chunk
constexpr int A_BYTES = BF * D * 2;   // 32 KB: two 64-dim swizzle atoms
constexpr int B_BYTES = BC * D * 2;   // 32 KB per chunk
constexpr int RING = 3;
constexpr int OFF_A = 0;
constexpr int OFF_B = OFF_A + A_BYTES;
constexpr int OFF_R = OFF_B + RING * B_BYTES;        // fp32 residuals, transposed [dim][frame]: 64 KB
constexpr int OFF_CN2 = OFF_R + BF * D * 4;          // ||c||^2 of the current stage
constexpr int MAX_K = 2048;
constexpr int OFF_BAR = OFF_CN2 + MAX_K * 4;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
constexpr int TMEM_COLS = 256;
}  // namespace rvq

struct RvqDev {
  CUtensorMap tmB;            // fp16 codebooks viewed as (Q*K rows, 128 cols)
  const float* frames;        // (F, 128)
  const float* codebooks;     // (Q, K, 128) fp32
  const float* cn2;           // (Q, K)
  const float* meta;          // (Q, 2): max ||c||, 2^e_q
  long long* codes;           // (F, Q)
  unsigned long long* stats;  // optional: [0] rows*stages, [1] ambiguous, [2] overflow full scans
  long long num_frames;
  int Q, K;
};

// exact squared distance in fp64 between one frame's residual (column `rcol` of the transposed smem tile,
// stride BF floats) and one fp32 codeword.  Four independent accumulators (dims i mod 4) break the DFMA
// dependency chain; the summation order is fixed, so equal inputs always give bit-equal results.
__device__ __forceinline__ double exact_dist(const float* rcol, const float* __restrict__ c) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  const float4* c4 = reinterpret_cast<const float4*>(c);
#pragma unroll 8
  for (int i = 0; i < rvq::D / 4; ++i) {
    const float4 v = __ldg(c4 + i);
    const double d0 = static_cast<double>(rcol[(4 * i + 0) * rvq::BF]) - static_cast<double>(v.x);
    const double d1 = static_cast<double>(rcol[(4 * i + 1) * rvq::BF]) - static_cast<double>(v.y);
    const double d2 = static_cast<double>(rcol[(4 * i + 2) * rvq::BF]) - static_cast<double>(v.z);
    const double d3 = static_cast<double>(rcol[(4 * i + 3) * rvq::BF]) - static_cast<double>(v.w);
    a0 = fma(d0, d0, a0);
    a1 = fma(d1, d1, a1);
    a2 = fma(d2, d2, a2);
    a3 = fma(d3, d3, a3);
  }
  return (a0 + a1) + (a2 + a3);
}

__global__ void __launch_bounds__(192, 1) rvq_encode_kernel(const __grid_constant__ RvqDev p) {
  using namespace rvq;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  float* cn2_s = reinterpret_cast<float*>(smem + OFF_CN2);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* b_full = bars + 0;    // [RING]
  uint64_t* b_empty = bars + 3;   // [RING]
  uint64_t* a_full = bars + 6;    // row threads -> MMA: A tile of this stage written
  uint64_t* d_full = bars + 7;    // [2]
  uint64_t* d_empty = bars + 9;   // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long f0 = static_cast<long long>(blockIdx.x) * BF;
  const int chunks = p.K / BC;

  if (warp == 4 && lane == 0) tma_prefetch_desc(&p.tmB);
  if (warp == 5 && lane == 0) {
    for (int i = 0; i < RING; ++i) {
      mbar_init(smem_u32(&b_full[i]), 1);
      mbar_init(smem_u32(&b_empty[i]), 1);
    }
    mbar_init(smem_u32(a_full), 128);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&d_full[i]), 1);
      mbar_init(smem_u32(&d_empty[i]), 128);
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
      uint32_t it = 0;
      for (int q = 0; q < p.Q; ++q) {
        for (int c = 0; c < chunks; ++c, ++it) {
          const uint32_t st = it % RING, ph = (it / RING) & 1;
          mbar_wait(smem_u32(&b_empty[st]), ph ^ 1);
          const uint32_t fb = smem_u32(&b_full[st]);
          mbar_arrive_expect_tx(fb, B_BYTES);
          const uint32_t dst = smem_u32(smem + OFF_B + st * B_BYTES);
          const int row0 = q * p.K + c * BC;
          tma_load_2d(dst, &p.tmB, fb, 0, row0);
          tma_load_2d(dst + BC * 128, &p.tmB, fb, 64, row0);
        }
      }
    }
  } else if (warp == 5) {
    // ================================ MMA issuer ==================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BF, BC, /*fp16*/ 0, 0, 0);
      const uint32_t abase = smem_u32(smem + OFF_A);
      uint32_t it = 0;
      for (int q = 0; q < p.Q; ++q) {
        mbar_wait(smem_u32(a_full), q & 1);
        tc_fence_after();
        for (int c = 0; c < chunks; ++c, ++it) {
          const uint32_t st = it % RING, ph = (it / RING) & 1;
          const uint32_t buf = it & 1, dph = (it >> 1) & 1;
          mbar_wait(smem_u32(&b_full[st]), ph);
          mbar_wait(smem_u32(&d_empty[buf]), dph ^ 1);
          tc_fence_after();
          const uint32_t bbase = smem_u32(smem + OFF_B + st * B_BYTES);
#pragma unroll
          for (int k = 0; k < D / 16; ++k) {
            const uint32_t off = (k >> 2) * (128 * 128) + (k & 3) * 32;  // atom, 16-dim step inside it
            tc_mma_f16(tmem_base + buf * BC, umma_desc_sw128(abase + off, 16, 1024),
                       umma_desc_sw128(bbase + off, 16, 1024), idesc, k > 0);
          }
          tc_commit(smem_u32(&b_empty[st]));
          tc_commit(smem_u32(&d_full[buf]));
        }
      }
    }
  } else {
    // ================================ row threads =================================
    const int row = warp * 32 + lane;
    const long long f = f0 + row;
    const bool live = f < p.num_frames;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    float* rcol = reinterpret_cast<float*>(smem + OFF_R) + row;  // element i of this frame: rcol[i * BF]
    {
      const float4* src = reinterpret_cast<const float4*>(p.frames + (live ? f : 0) * D);
#pragma unroll 4
      for (int i = 0; i < D / 4; ++i) {
        const float4 v = live ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        rcol[(4 * i + 0) * BF] = v.x;
        rcol[(4 * i + 1) * BF] = v.y;
        rcol[(4 * i + 2) * BF] = v.z;
        rcol[(4 * i + 3) * BF] = v.w;
      }
    }
    unsigned long long n_ambig = 0, n_full = 0;
    uint32_t it = 0;
    for (int q = 0; q < p.Q; ++q) {
      // ---- stage prologue: ||c||^2 table, row scale and norm, fp16 A tile ----
      asm volatile("bar.sync 1, 128;" ::: "memory");  // everyone is done with the previous cn2 table
      for (int i = row; i < p.K; i += 128) cn2_s[i] = __ldg(p.cn2 + q * p.K + i);
      float amax = 0.f, ss = 0.f;
#pragma unroll 8
      for (int i = 0; i < D; ++i) {
        const float v = rcol[i * BF];
        amax = fmaxf(amax, fabsf(v));
        ss = fmaf(v, v, ss);
      }
      int ex = 0;
      if (amax > 0.f) (void)frexpf(amax, &ex);  // amax = m * 2^ex, m in [0.5, 1)
      const float xs = ldexpf(1.0f, -ex);      // exact power-of-two scale into fp16 range
      {
        uint8_t* arow = smem + OFF_A + row * 128;
#pragma unroll 2
        for (int ch = 0; ch < 16; ++ch) {  // 16 chunks of 8 halves
          uint32_t w[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const __half2 h = __floats2half2_rn(rcol[(ch * 8 + 2 * j) * BF] * xs,
                                                rcol[(ch * 8 + 2 * j + 1) * BF] * xs);
            w[j] = *reinterpret_cast<const uint32_t*>(&h);
          }
          uint8_t* atom = arow + (ch >> 3) * (BF * 128);
          *reinterpret_cast<uint4*>(atom + (((ch & 7) ^ (row & 7)) << 4)) =
              make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(smem_u32(a_full));
      asm volatile("bar.sync 1, 128;" ::: "memory");  // cn2 table visible to all row threads

      const float cmax = __ldg(p.meta + 2 * q), cscale = __ldg(p.meta + 2 * q + 1);
      const float dscale = -2.0f * ldexpf(cscale, ex);  // s~ = cn2 + dscale * dot'
      // rigorous filter margin (see header comment); 1.001 covers the fp32 rounding of ||r|| itself
      const float margin2 = 2.0f * (2.0f * 0.001026f /*1.05 * 2^-10*/ * sqrtf(ss) * 1.001f * cmax);

      // ---- scan all codes, keep the four best approximate scores ----
      float b0 = INFINITY, b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;
      int i0 = 0, i1 = 0, i2 = 0;
      for (int c = 0; c < chunks; ++c, ++it) {
        const uint32_t buf = it & 1, dph = (it >> 1) & 1;
        mbar_wait(smem_u32(&d_full[buf]), dph);
        tc_fence_after();
#pragma unroll 1
        for (int sub = 0; sub < BC / 32; ++sub) {
          uint32_t v[32];
          tmem_ld32(lane_addr + buf * BC + sub * 32, v);
          tmem_ld_wait();
          const int kbase = c * BC + sub * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float s = fmaf(dscale, __uint_as_float(v[i]), cn2_s[kbase + i]);
            if (s < b3) {  // rare after the first few codes: insertion into the sorted top-4
              const int k = kbase + i;
              if (s < b0) { b3 = b2; b2 = b1; i2 = i1; b1 = b0; i1 = i0; b0 = s; i0 = k; }
              else if (s < b1) { b3 = b2; b2 = b1; i2 = i1; b1 = s; i1 = k; }
              else if (s < b2) { b3 = b2; b2 = s; i2 = k; }
              else { b3 = s; }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(smem_u32(&d_empty[buf]));
      }

      // ---- exact decision ----
      const float* cbq = p.codebooks + static_cast<long long>(q) * p.K * D;
      int best = i0;
      const float lim = b0 + margin2;
      const bool overflow = b3 <= lim;  // more than three codes inside the uncertainty band
      if (b1 <= lim && !overflow) {
        ++n_ambig;
        double dbest = exact_dist(rcol, cbq + static_cast<long long>(i0) * D);
        const double d1 = exact_dist(rcol, cbq + static_cast<long long>(i1) * D);
        if (d1 < dbest || (d1 == dbest && i1 < best)) { dbest = d1; best = i1; }
        if (b2 <= lim) {
          const double d2 = exact_dist(rcol, cbq + static_cast<long long>(i2) * D);
          if (d2 < dbest || (d2 == dbest && i2 < best)) { dbest = d2; best = i2; }
        }
      }
      // overflow rows (rare): the whole warp scans the codebook exactly for that one row, 32 codes per lane
      unsigned omask = __ballot_sync(0xffffffffu, overflow);
      if (overflow) { ++n_ambig; ++n_full; }
      while (omask) {
        const int src = __ffs(omask) - 1;
        omask &= omask - 1;
        const float* rsrc = reinterpret_cast<const float*>(smem + OFF_R) + (warp * 32 + src);
        double dmin = INFINITY;
        int kmin = 0x7fffffff;
        for (int k = lane; k < p.K; k += 32) {
          const double dk = exact_dist(rsrc, cbq + static_cast<long long>(k) * D);
          if (dk < dmin) { dmin = dk; kmin = k; }  // k ascending per lane: first minimum kept
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double od = __shfl_xor_sync(0xffffffffu, dmin, o);
          const int ok = __shfl_xor_sync(0xffffffffu, kmin, o);
          if (od < dmin || (od == dmin && ok < kmin)) { dmin = od; kmin = ok; }
        }
        if (lane == src) best = kmin;
      }
      if (live) p.codes[f * p.Q + q] = best;
      // ---- residual update with the exact fp32 codeword (same op as the reference) ----
      {
        const float4* cw = reinterpret_cast<const float4*>(cbq + static_cast<long long>(best) * D);
#pragma unroll 4
        for (int i = 0; i < D / 4; ++i) {
          const float4 v = __ldg(cw + i);
          rcol[(4 * i + 0) * BF] -= v.x;
          rcol[(4 * i + 1) * BF] -= v.y;
          rcol[(4 * i + 2) * BF] -= v.z;
          rcol[(4 * i + 3) * BF] -= v.w;
        }
      }
    }
    if (p.stats != nullptr && live) {
      atomicAdd(p.stats + 0, static_cast<unsigned long long>(p.Q));
      if (n_ambig) atomicAdd(p.stats + 1, n_ambig);
      if (n_full) atomicAdd(p.stats + 2, n_full);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// prepare: fp16 copy scaled by a per-quantiser power of two, ||c||^2, max ||c||
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rvq_prepare_kernel(const float* __restrict__ cb, int K, int D,
                                                          __half* __restrict__ cb16,
                                                          float* __restrict__ cn2,
                                                          float* __restrict__ meta) {
  const int q = blockIdx.x;
  const float* c = cb + static_cast<long long>(q) * K * D;
  __shared__ float red[8];
  __shared__ float s_scale;
  float amax = 0.f;
  for (int i = threadIdx.x; i < K * D; i += blockDim.x) amax = fmaxf(amax, fabsf(c[i]));
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = amax;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = 0.f;
    for (int i = 0; i < 8; ++i) m = fmaxf(m, red[i]);
    int e = 0;
    if (m > 0.f) (void)frexpf(m, &e);
    s_scale = ldexpf(1.0f, e);  // codes are stored as c * 2^-e, |.| < 1
    meta[2 * q + 1] = s_scale;
  }
  __syncthreads();
  const float inv = 1.0f / s_scale;
  for (int i = threadIdx.x; i < K * D; i += blockDim.x)
    cb16[static_cast<long long>(q) * K * D + i] = __float2half_rn(c[i] * inv);
  // ||c||^2 in fp64, rounded once to fp32; one warp per code
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float nmax = 0.f;
  for (int k = warp; k < K; k += 8) {
    double s = 0.0;
    for (int i = lane; i < D; i += 32) {
      const double v = c[static_cast<long long>(k) * D + i];
      s += v * v;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) cn2[q * K + k] = static_cast<float>(s);
    nmax = fmaxf(nmax, static_cast<float>(sqrt(s)) * 1.0001f);
  }
  __syncthreads();
  if (lane == 0) red[warp] = nmax;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = 0.f;
    for (int i = 0; i < 8; ++i) m = fmaxf(m, red[i]);
    meta[2 * q] = m;
  }
}

// decode: emb = sum_q C_q[codes[:, q]] accumulated in order q = 0..Q-1 (one warp per frame, float4 per lane)
__global__ void __launch_bounds__(256) rvq_decode_kernel(const long long* __restrict__ codes,
                                                         long long F, int Q, int K,
                                                         const float* __restrict__ cb,
                                                         float* __restrict__ emb) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long f = static_cast<long long>(blockIdx.x) * 8 + warp;
  if (f >= F) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int q = 0; q < Q; ++q) {
    long long idx = codes[f * Q + q];
    idx = idx < 0 ? 0 : (idx >= K ? K - 1 : idx);
    const float4 v =
        __ldg(reinterpret_cast<const float4*>(cb + (static_cast<long long>(q) * K + idx) * 128) + lane);
    acc.x += v.x;
    acc.y += v.y;
    acc.z += v.z;
    acc.w += v.w;
  }
  reinterpret_cast<float4*>(emb + f * 128)[lane] = acc;
}

}  // namespace ns2

using namespace ns2;

extern "C" {

int ns2_rvq_prepare(const float* codebooks, int32_t q, int32_t k, int32_t d, void* cb_f16,
                    float* cb_norm2, float* cb_meta, ns2_stream_t stream) {
  NS2_REQUIRE(codebooks && cb_f16 && cb_norm2 && cb_meta, "rvq_prepare: NULL pointer");
  NS2_REQUIRE(q > 0 && k > 0 && d == 128, "rvq_prepare: d must be 128 (got %d)", d);
  rvq_prepare_kernel<<<q, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      codebooks, k, d, reinterpret_cast<__half*>(cb_f16), cb_norm2, cb_meta);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  NS2_CUDA_CHECK(cudaGetLastError());
  return kOk;
}

int ns2_rvq_encode(const float* frames, int64_t num_frames, int32_t d, const float* codebooks,
                   const void* cb_f16, const float* cb_norm2, const float* cb_meta, int32_t q,
                   int32_t k, int64_t* codes, int64_t* stats, ns2_stream_t stream) {
  NS2_REQUIRE(frames && codebooks && cb_f16 && cb_norm2 && cb_meta && codes, "rvq_encode: NULL pointer");
  NS2_REQUIRE(d == 128, "rvq_encode: d must be 128 (got %d)", d);
  NS2_REQUIRE(k % rvq::BC == 0 && k <= rvq::MAX_K && k > 0,
              "rvq_encode: codebook size %d must be a multiple of 128, <= %d", k, rvq::MAX_K);
  NS2_REQUIRE(q > 0 && num_frames > 0, "rvq_encode: empty problem");
  NS2_REQUIRE((reinterpret_cast<uintptr_t>(frames) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(codebooks) & 15) == 0,
              "rvq_encode: frames and codebooks must be 16-byte aligned");
  RvqDev dev;
  memset(&dev, 0, sizeof(dev));
  const uint64_t dims[2] = {128, (uint64_t)q * k};
  const uint64_t str[2] = {2, 256};
  const uint32_t box[2] = {64, rvq::BC};
  int rc = make_tmap_16bit(&dev.tmB, cb_f16, 2, dims, str, box);
  if (rc != kOk) return rc;
  dev.frames = frames;
  dev.codebooks = codebooks;
  dev.cn2 = cb_norm2;
  dev.meta = cb_meta;
  dev.codes = reinterpret_cast<long long*>(codes);
  dev.stats = reinterpret_cast<unsigned long long*>(stats);
  dev.num_frames = num_frames;
  dev.Q = q;
  dev.K = k;
  static bool configured = false;
  if (!configured) {
    NS2_CUDA_CHECK(cudaFuncSetAttribute(rvq_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        rvq::SMEM_BYTES));
    configured = true;
  }
  const long long grid = (num_frames + rvq::BF - 1) / rvq::BF;
  NS2_REQUIRE(grid <= 0x7fffffffLL, "rvq_encode: too many frames");
  rvq_encode_kernel<<<static_cast<unsigned>(grid), 192, rvq::SMEM_BYTES,
                      static_cast<cudaStream_t>(stream)>>>(dev);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  NS2_CUDA_CHECK(cudaGetLastError());
  return kOk;
}

int ns2_rvq_decode(const int64_t* codes, int64_t num_frames, int32_t q, int32_t k, int32_t d,
                   const float* codebooks, float* emb, ns2_stream_t stream) {
  NS2_REQUIRE(codes && codebooks && emb, "rvq_decode: NULL pointer");
  NS2_REQUIRE(d == 128 && q > 0 && k > 0 && num_frames > 0, "rvq_decode: d must be 128");
  const long long grid = (num_frames + 7) / 8;
  rvq_decode_kernel<<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(codes), num_frames, q, k, codebooks, emb);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  NS2_CUDA_CHECK(cudaGetLastError());
  return kOk;
}

}  // extern "C"
