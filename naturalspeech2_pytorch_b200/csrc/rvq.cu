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
// One CTA per 128 frames, 320 threads:
//   warps 0-7  scan threads: two threads per frame (TMEM lane), each scanning 64 of every 128 score columns with a
//              branch-free top-4 on packed (score|index) keys; fp32 residuals live in padded shared memory for all
//              Q stages; the same threads build the fp16 A tile, re-score near-ties in fp64 (per lane, or
//              warp-cooperatively for crowded 32-code blocks / bands) and subtract the chosen fp32 codeword
//   warp 8     TMA producer: streams the fp16 codebooks (128 codes x 128 dims per chunk) through a 3-deep ring
//   warp 9     tcgen05.mma issuer: D[128 frames x 128 codes] per chunk, double-buffered in TMEM
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
constexpr int RSTRIDE = D + 4;        // floats per residual row: +4 keeps per-thread float4 row reads conflict-free
constexpr int MAX_K = 2048;
constexpr int OFF_A = 0;
constexpr int OFF_B = OFF_A + A_BYTES;
constexpr int OFF_R = OFF_B + RING * B_BYTES;        // fp32 residuals, row-major padded: 67.6 KB
constexpr int OFF_CN2 = OFF_R + BF * RSTRIDE * 4;    // ||c||^2, double-buffered across stages (2 x 8 KB)
constexpr int OFF_KEYS = OFF_CN2 + 2 * MAX_K * 4;    // top-8 keys of each column half: [half][row][8] floats (8 KB)
constexpr int OFF_ROWP = OFF_KEYS + 2 * BF * 8 * 4;  // per-row {scale, dscale, E16, unused}
constexpr int OFF_SEL = OFF_ROWP + BF * 4 * 4;       // chosen code per row (this stage)
constexpr int OFF_CAND = OFF_SEL + BF * 4;           // per-warp candidate list of the row being re-scored
constexpr int OFF_BAR = OFF_CAND + 8 * 16 * 4;
constexpr int SMEM_BYTES = OFF_BAR + 256;
constexpr int TMEM_COLS = 256;
constexpr int SCAN_THREADS = 256;     // warps 0-7: quarter = warp & 3 (TMEM lanes), column half = warp >> 2
}  // namespace rvq

struct RvqDev {
  CUtensorMap tmB;            // fp16 codebooks viewed as (Q*K rows, 128 cols)
  const float* frames;        // (F, 128)
  const float* codebooks;     // (Q, K, 128) fp32
  const float* cn2;           // (Q, K)
  const float* meta;          // (Q, 2): max ||c||, 2^e_q
  long long* codes;           // (F, Q)
  unsigned long long* stats;  // optional: [0] lookups, [1] near-ties re-scored, [2] full scans, [3] sub-chunk scans
  long long num_frames;
  int Q, K;
};

// Exact squared distance in fp64 between one residual row (shared memory, contiguous) and one fp32 codeword.
// Four independent accumulators (dims i mod 4) break the DFMA dependency chain; the summation order is fixed, so
// equal inputs always give bit-equal results (ties between duplicate codewords resolve by index).
__device__ __forceinline__ double exact_dist(const float* rrow, const float* __restrict__ c) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  const float4* c4 = reinterpret_cast<const float4*>(c);
  const float4* r4 = reinterpret_cast<const float4*>(rrow);
#pragma unroll 8
  for (int i = 0; i < rvq::D / 4; ++i) {
    const float4 v = __ldg(c4 + i);
    const float4 r = r4[i];
    const double d0 = static_cast<double>(r.x) - static_cast<double>(v.x);
    const double d1 = static_cast<double>(r.y) - static_cast<double>(v.y);
    const double d2 = static_cast<double>(r.z) - static_cast<double>(v.z);
    const double d3 = static_cast<double>(r.w) - static_cast<double>(v.w);
    a0 = fma(d0, d0, a0);
    a1 = fma(d1, d1, a1);
    a2 = fma(d2, d2, a2);
    a3 = fma(d3, d3, a3);
  }
  return (a0 + a1) + (a2 + a3);
}

// Warp-cooperative variant: lane l owns dims [4l, 4l+4); one coalesced 512-byte load of the codeword (issued by the
// caller so that several are in flight), a fixed butterfly reduction (equal inputs give bit-equal results).
// Every lane returns the full distance.
__device__ __forceinline__ float4 coop_load(const float* __restrict__ c, int lane) {
  return __ldg(reinterpret_cast<const float4*>(c) + lane);
}
__device__ __forceinline__ double coop_reduce(const float4 r, const float4 v) {
  const double d0 = static_cast<double>(r.x) - static_cast<double>(v.x);
  const double d1 = static_cast<double>(r.y) - static_cast<double>(v.y);
  const double d2 = static_cast<double>(r.z) - static_cast<double>(v.z);
  const double d3 = static_cast<double>(r.w) - static_cast<double>(v.w);
  double a = fma(d0, d0, fma(d1, d1, fma(d2, d2, d3 * d3)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  return a;
}

__device__ __forceinline__ void scan_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// sorted insert of a key into the ascending 4-tuple (g0..g3): 7 min/max, no branches
__device__ __forceinline__ void insert4(float key, float& g0, float& g1, float& g2, float& g3) {
  float t = key, lo;
  lo = fminf(g0, t); t = fmaxf(g0, t); g0 = lo;
  lo = fminf(g1, t); t = fmaxf(g1, t); g1 = lo;
  lo = fminf(g2, t); t = fmaxf(g2, t); g2 = lo;
  g3 = fminf(g3, t);
}

// same for an ascending 8-tuple: 15 min/max
__device__ __forceinline__ void insert8(float key, float (&g)[8]) {
  float t = key;
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const float lo = fminf(g[i], t);
    t = fmaxf(g[i], t);
    g[i] = lo;
  }
  g[7] = fminf(g[7], t);
}

// (distance, index) lexicographic minimum across the warp
__device__ __forceinline__ void warp_argmin(double& d, int& k) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double od = __shfl_xor_sync(0xffffffffu, d, o);
    const int ok = __shfl_xor_sync(0xffffffffu, k, o);
    if (od < d || (od == d && ok < k)) { d = od; k = ok; }
  }
}

// Approximate scores are carried as "keys": the fp32 score with its low mantissa bits replaced by the code index,
// so that min/max on the keys sorts (score, index) pairs without branches or separate index registers.
// Low 11 bits = index (K <= 2048); the 2^-12 relative truncation is folded into the re-score margin.
__global__ void __launch_bounds__(320, 1) rvq_encode_kernel(const __grid_constant__ RvqDev p) {
  using namespace rvq;
  // declared 1024-byte aligned (checked below) and indexed directly so the compiler keeps every access in the
  // shared state space (LDS/STS instead of generic LD/ST)
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("ns2 rvq: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  float* R = reinterpret_cast<float*>(smem + OFF_R);
  float* cn2_s = reinterpret_cast<float*>(smem + OFF_CN2);
  float* keys_s = reinterpret_cast<float*>(smem + OFF_KEYS);
  float4* rowp_s = reinterpret_cast<float4*>(smem + OFF_ROWP);
  int* sel_s = reinterpret_cast<int*>(smem + OFF_SEL);
  int* cand_s = reinterpret_cast<int*>(smem + OFF_CAND);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* b_full = bars + 0;    // [RING]
  uint64_t* b_empty = bars + 3;   // [RING]
  uint64_t* a_full = bars + 6;    // scan threads -> MMA: A tile of this stage written
  uint64_t* d_full = bars + 7;    // [2]
  uint64_t* d_empty = bars + 9;   // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long f0 = static_cast<long long>(blockIdx.x) * BF;
  const int chunks = p.K / BC;

  if (warp == 8 && lane == 0) tma_prefetch_desc(&p.tmB);
  if (warp == 9 && lane == 0) {
    for (int i = 0; i < RING; ++i) {
      mbar_init(smem_u32(&b_full[i]), 1);
      mbar_init(smem_u32(&b_empty[i]), 1);
    }
    mbar_init(smem_u32(a_full), SCAN_THREADS);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&d_full[i]), 1);
      mbar_init(smem_u32(&d_empty[i]), SCAN_THREADS);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_holder), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 8) {
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
  } else if (warp == 9) {
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
    // ================================ scan threads ================================
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;       // frame owned (together with the thread of the other half)
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    float* rrow = R + row * RSTRIDE;

    // cooperative, coalesced load of the 128 frames: warp w fills rows [16w, 16w+16); stage-0 row parameters
    {
#pragma unroll 4
      for (int r = warp * 16; r < warp * 16 + 16; ++r) {
        const long long fr = f0 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (fr < p.num_frames) v = __ldg(reinterpret_cast<const float4*>(p.frames + fr * D) + lane);
        reinterpret_cast<float4*>(R + r * RSTRIDE)[lane] = v;
      }
      for (int i = threadIdx.x; i < p.K; i += SCAN_THREADS) cn2_s[i] = __ldg(p.cn2 + i);
    }
    unsigned long long n_ambig = 0, n_full = 0, n_sub = 0;
    uint32_t it = 0;
    for (int q = 0; q < p.Q; ++q) {
      scan_barrier();  // [B1] residuals and ||c||^2 of this stage are in place
      const float* cn2q = cn2_s + (q & 1) * MAX_K;
      // row scale (exact power of two into fp16 range), |r|^2 and the filter margin; both threads of a row compute
      // them redundantly from the same data in the same order (bit-identical), so no exchange is needed
      float4 rp;
      {
        const float cmax = __ldg(p.meta + 2 * q), cscale = __ldg(p.meta + 2 * q + 1);
        float amax = 0.f, ss = 0.f;
#pragma unroll 8
        for (int i = 0; i < D / 4; ++i) {
          const float4 v = reinterpret_cast<const float4*>(rrow)[i];
          amax = fmaxf(fmaxf(amax, fabsf(v.x)), fmaxf(fabsf(v.y), fmaxf(fabsf(v.z), fabsf(v.w))));
          ss = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, ss))));
        }
        int ex = 0;
        if (amax > 0.f) (void)frexpf(amax, &ex);          // amax = m * 2^ex, m in [0.5, 1)
        const float xs = ldexpf(1.0f, -ex);
        const float dscale = -2.0f * ldexpf(cscale, ex);  // s~ = cn2 + dscale * dot'
        // |s~_k - s_k| <= E16 = 2 * 1.05 * 2^-10 * ||r|| * max||c||  (fp16 operand rounding, fp32 accumulate)
        const float e16 = 2.0f * 0.001026f * sqrtf(ss) * 1.001f * cmax;
        rp = make_float4(xs, dscale, e16, ss);
        if (half == 0) rowp_s[row] = rp;   // read by the decision warps after [B2]
      }
      // ---- fp16 A tile: this thread converts dims [64*half, 64*half + 64) of its row into atom `half` ----
      {
        const float xs = rp.x;
        uint8_t* arow = smem + OFF_A + half * (BF * 128) + row * 128;
        const float4* src = reinterpret_cast<const float4*>(rrow + half * 64);
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {  // 8 chunks of 8 halves
          const float4 v0 = src[2 * ch], v1 = src[2 * ch + 1];
          const __half2 h0 = __floats2half2_rn(v0.x * xs, v0.y * xs), h1 = __floats2half2_rn(v0.z * xs, v0.w * xs);
          const __half2 h2 = __floats2half2_rn(v1.x * xs, v1.y * xs), h3 = __floats2half2_rn(v1.z * xs, v1.w * xs);
          *reinterpret_cast<uint4*>(arow + ((ch ^ (row & 7)) << 4)) =
              make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                         *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(smem_u32(a_full));
      // prefetch ||c||^2 of the next stage into registers (stored to the other buffer after the scan)
      float cn_next[MAX_K / SCAN_THREADS];
      if (q + 1 < p.Q) {
#pragma unroll
        for (int u = 0; u < MAX_K / SCAN_THREADS; ++u) {
          const int i = threadIdx.x + u * SCAN_THREADS;
          cn_next[u] = (i < p.K) ? __ldg(p.cn2 + (q + 1) * p.K + i) : 0.f;
        }
      }

      // ---- scan this thread's 64 columns of every 128-code chunk: branch-free top-8 on packed keys ----
      const float dscale = rp.y;
      float g[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] = INFINITY;
      for (int c = 0; c < chunks; ++c, ++it) {
        const uint32_t buf = it & 1, dph = (it >> 1) & 1;
        mbar_wait(smem_u32(&d_full[buf]), dph);
        tc_fence_after();
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
          uint32_t v[32];
          tmem_ld32(lane_addr + buf * BC + half * 64 + sub * 32, v);
          tmem_ld_wait();
          const int sub_id = c * 4 + half * 2 + sub;   // 32-code block index: code = sub_id * 32 + i
          const float* cn = cn2q + sub_id * 32;
          // local top-2 of the 32 scores, index i in the low 5 bits; two interleaved trackers for ILP
          float a0 = INFINITY, a1 = INFINITY, b0 = INFINITY, b1 = INFINITY;
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float s0 = fmaf(dscale, __uint_as_float(v[i]), cn[i]);
            const float s1 = fmaf(dscale, __uint_as_float(v[i + 1]), cn[i + 1]);
            const float k0 = __uint_as_float((__float_as_uint(s0) & 0xFFFFFFE0u) | static_cast<uint32_t>(i));
            const float k1 = __uint_as_float((__float_as_uint(s1) & 0xFFFFFFE0u) | static_cast<uint32_t>(i + 1));
            float t = fmaxf(a0, k0); a0 = fminf(a0, k0); a1 = fminf(a1, t);
            t = fmaxf(b0, k1); b0 = fminf(b0, k1); b1 = fminf(b1, t);
          }
          const float l0 = fminf(a0, b0);
          const float l1 = fminf(fmaxf(a0, b0), fminf(a1, b1));
          // widen the index field to 11 bits (block id above the 5 local bits) and merge into the global top-8
          const uint32_t blk = static_cast<uint32_t>(sub_id) << 5;
          insert8(__uint_as_float((__float_as_uint(l0) & 0xFFFFF81Fu) | blk), g);
          insert8(__uint_as_float((__float_as_uint(l1) & 0xFFFFF81Fu) | blk), g);
        }
        tc_fence_before();
        mbar_arrive(smem_u32(&d_empty[buf]));
      }
      {
        float4* kd = reinterpret_cast<float4*>(keys_s + (half * BF + row) * 8);
        kd[0] = make_float4(g[0], g[1], g[2], g[3]);
        kd[1] = make_float4(g[4], g[5], g[6], g[7]);
      }
      if (q + 1 < p.Q) {
        float* cnw = cn2_s + ((q + 1) & 1) * MAX_K;
#pragma unroll
        for (int u = 0; u < MAX_K / SCAN_THREADS; ++u) {
          const int i = threadIdx.x + u * SCAN_THREADS;
          if (i < p.K) cnw[i] = cn_next[u];
        }
      }
      scan_barrier();  // [B2] both halves' key lists are published

      // ---- exact decision: warp w owns rows [16w, 16w+16) ----
      // Candidates = every code whose key is within the error band of the best key.  Each column half keeps its own
      // sorted top-8; a band member can only be missing from the two lists if a list is entirely inside the band
      // (-> exact scan of the whole codebook) or if it was 3rd+ inside its 32-code block, in which case two better
      // band members share that block (-> that block is scanned exactly).
      const float* cbq = p.codebooks + static_cast<long long>(q) * p.K * D;
      {
        // lanes 0..15 classify one row each (lanes 16..31 mirror them): band sizes of the two sorted lists
        const int myrow = warp * 16 + (lane & 15);
        float ka[8], kb[8];
        {
          const float4* k0 = reinterpret_cast<const float4*>(keys_s + myrow * 8);
          const float4* k1 = reinterpret_cast<const float4*>(keys_s + (BF + myrow) * 8);
          const float4 x0 = k0[0], x1 = k0[1], y0 = k1[0], y1 = k1[1];
          ka[0] = x0.x; ka[1] = x0.y; ka[2] = x0.z; ka[3] = x0.w; ka[4] = x1.x; ka[5] = x1.y; ka[6] = x1.z; ka[7] = x1.w;
          kb[0] = y0.x; kb[1] = y0.y; kb[2] = y0.z; kb[3] = y0.w; kb[4] = y1.x; kb[5] = y1.y; kb[6] = y1.z; kb[7] = y1.w;
        }
        const float kmin = fminf(ka[0], kb[0]);
        const float e16 = rowp_s[myrow].z;
        const float etrunc = 0.000244140625f * 1.01f * (fabsf(kmin) + 2.0f * e16);  // 2^-12 key truncation
        const float lim = kmin + 2.0f * (e16 + etrunc);
        int na = 0, nb = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          na += (ka[i] <= lim) ? 1 : 0;   // lists are sorted: the band is a prefix
          nb += (kb[i] <= lim) ? 1 : 0;
        }
        if (lane < 16) sel_s[myrow] = __float_as_uint(kmin) & 0x7FF;
        unsigned todo = __ballot_sync(0xffffffffu, lane < 16 && na + nb > 1);
        int* cand = cand_s + warp * 16;
#pragma unroll 1
        while (todo) {
          const int src = __ffs(todo) - 1;
          todo &= todo - 1;
          const int r = warp * 16 + src;
          const int ca = __shfl_sync(0xffffffffu, na, src), cb = __shfl_sync(0xffffffffu, nb, src);
          // the row's owner publishes its candidate indices: list a in cand[0..8), list b in cand[8..16)
          if (lane == src) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              cand[u] = __float_as_uint(ka[u]) & 0x7FF;
              cand[8 + u] = __float_as_uint(kb[u]) & 0x7FF;
            }
          }
          __syncwarp();
          const float4 rv = reinterpret_cast<const float4*>(R + r * RSTRIDE)[lane];
          double dbest = INFINITY;
          int best = 0x7fffffff;
          ++n_ambig;
          // crowded 32-code block among the band members of one list? (blocks never span the two halves)
          int blk = -1;
          bool full = (ca >= 8) || (cb >= 8);
#pragma unroll 1
          for (int l = 0; l < 2; ++l) {
            const int cnt = l ? cb : ca;
#pragma unroll 1
            for (int i = 0; i + 1 < cnt; ++i)
#pragma unroll 1
              for (int j = i + 1; j < cnt; ++j)
                if ((cand[8 * l + i] >> 5) == (cand[8 * l + j] >> 5)) {
                  full = full || (blk >= 0 && blk != (cand[8 * l + i] >> 5));
                  blk = cand[8 * l + i] >> 5;
                }
          }
          if (full) {
            // a whole list inside the band, or two crowded blocks (astronomically rare): exact scan of the codebook
            ++n_full;
#pragma unroll 1
            for (int k0 = 0; k0 < p.K; k0 += 8) {
              float4 cv[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) cv[u] = coop_load(cbq + static_cast<long long>(k0 + u) * D, lane);
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const double dk = coop_reduce(rv, cv[u]);
                if (dk < dbest) { dbest = dk; best = k0 + u; }
              }
            }
          } else {
            // band members of both lists, four codewords in flight at a time
            const int total = ca + cb;
#pragma unroll 1
            for (int u0 = 0; u0 < total; u0 += 4) {
              float4 cv[4];
              int kx[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int t = u0 + u;
                kx[u] = (t < total) ? cand[t < ca ? t : 8 + (t - ca)] : cand[0];
                cv[u] = coop_load(cbq + static_cast<long long>(kx[u]) * D, lane);
              }
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const double dk = coop_reduce(rv, cv[u]);
                if (u0 + u < total && (dk < dbest || (dk == dbest && kx[u] < best))) { dbest = dk; best = kx[u]; }
              }
            }
            if (blk >= 0) {
              ++n_sub;
#pragma unroll 1
              for (int k0 = blk * 32; k0 < blk * 32 + 32; k0 += 8) {
                float4 c8[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) c8[u] = coop_load(cbq + static_cast<long long>(k0 + u) * D, lane);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                  const double dk = coop_reduce(rv, c8[u]);
                  if (dk < dbest || (dk == dbest && k0 + u < best)) { dbest = dk; best = k0 + u; }
                }
              }
            }
          }
          if (lane == 0) sel_s[r] = best;
          __syncwarp();
        }
      }
      __syncwarp();
      // ---- residual update with the exact fp32 codeword (same op as the reference) for this warp's own 16 rows;
      //      eight coalesced 512-byte codeword loads in flight ----
      {
#pragma unroll 1
        for (int u0 = 0; u0 < 16; u0 += 8) {
          float4 cw[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int r = warp * 16 + u0 + u;
            const int sel = sel_s[r];
            cw[u] = coop_load(cbq + static_cast<long long>(sel) * D, lane);
            if (lane == 0 && f0 + r < p.num_frames) p.codes[(f0 + r) * p.Q + q] = sel;
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            float4* dst = reinterpret_cast<float4*>(R + (warp * 16 + u0 + u) * RSTRIDE) + lane;
            float4 v = *dst;
            v.x -= cw[u].x; v.y -= cw[u].y; v.z -= cw[u].z; v.w -= cw[u].w;
            *dst = v;
          }
        }
      }
    }
    if (p.stats != nullptr) {
      if (half == 0 && f0 + row < p.num_frames) atomicAdd(p.stats + 0, static_cast<unsigned long long>(p.Q));
      if (lane == 0) {  // the cooperative decision counts per warp
        if (n_ambig) atomicAdd(p.stats + 1, n_ambig);
        if (n_full) atomicAdd(p.stats + 2, n_full);
        if (n_sub) atomicAdd(p.stats + 3, n_sub);
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
  NS2_CUDA_CHECK(set_max_smem_once(rvq_encode_kernel, rvq::SMEM_BYTES));
  const long long grid = (num_frames + rvq::BF - 1) / rvq::BF;
  NS2_REQUIRE(grid <= 0x7fffffffLL, "rvq_encode: too many frames");
  rvq_encode_kernel<<<static_cast<unsigned>(grid), 320, rvq::SMEM_BYTES,
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
