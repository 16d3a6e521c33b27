// Micro-benchmarks that calibrate the attention / epilogue designs on sm_100a (one CTA, one SM):
//   TMEM load / store throughput vs warps and vector width, MUFU ex2, FFMA2 polynomial exp2, F2FP packing,
//   FMNMX3.  Prints cycles per warp-instruction and bytes/clk/SM.  Build: tools/build_ubench.sh; run on a B200.
#include "../naturalspeech2_pytorch_b200/csrc/ptx.cuh"
#include <cstdio>
#include <vector>

using namespace ns2;

// mode 0: ld x32, wait after every load; 1: 4 loads (128 cols) then one wait; 2: st x32 (2 stores then wait)
template <int MODE>
__global__ void tmem_bw_kernel(long long* out, uint32_t* sink, int iters) {
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(smem_u32(&holder), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = holder + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
  if (MODE == 2) {  // initialise TMEM so loads read defined data
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if constexpr (MODE == 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(base + (((warp >> 2) * 128 + c * 32) % 512), v);
        tmem_ld_wait();
        acc += v[0] ^ v[31];
      }
    } else if constexpr (MODE == 1) {
      uint32_t v0[32], v1[32], v2[32], v3[32];
      const uint32_t a = base + (((warp >> 2) * 128) % 512);
      tmem_ld32(a, v0);
      tmem_ld32(a + 32, v1);
      tmem_ld32(a + 64, v2);
      tmem_ld32(a + 96, v3);
      tmem_ld_wait();
      acc += v0[0] ^ v1[31] ^ v2[5] ^ v3[7];
    } else {
      const uint32_t a = base + (((warp >> 2) * 128) % 512);
      r[0] = it;
      tmem_st32(a, r);
      tmem_st32(a + 32, r);
      tmem_st32(a + 64, r);
      tmem_st32(a + 96, r);
      tmem_st_wait();
    }
  }
  const long long t1 = clock64();
  sink[threadIdx.x] = acc;
  if ((threadIdx.x & 31) == 0) out[warp] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(holder, 512);
}

__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Cody-Waite exp2 on the FMA pipe for a pair (x <= 0 expected, clamped at -126): 2 elements per call
__device__ __forceinline__ void exp2_poly2(float x0, float x1, float& y0, float& y1) {
  // t = x + 1.5*2^23 (round to nearest), n = t - magic, f = x - n in [-0.5, 0.5]
  const float magic = 12582912.0f;
  unsigned long long X, T, N, F, P, M, C3, C2, C1, C0, NEG1;
  asm("mov.b64 %0, {%1,%2};" : "=l"(X) : "f"(x0), "f"(x1));
  asm("mov.b64 %0, {%1,%1};" : "=l"(M) : "f"(magic));
  asm("mov.b64 %0, {%1,%1};" : "=l"(NEG1) : "f"(-1.0f));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(T) : "l"(X), "l"(M));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(N) : "l"(M), "l"(NEG1), "l"(T));   // n = t - magic
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(F) : "l"(N), "l"(NEG1), "l"(X));   // f = x - n
  asm("mov.b64 %0, {%1,%1};" : "=l"(C3) : "f"(0.0555041f));
  asm("mov.b64 %0, {%1,%1};" : "=l"(C2) : "f"(0.2402265f));
  asm("mov.b64 %0, {%1,%1};" : "=l"(C1) : "f"(0.6931472f));
  asm("mov.b64 %0, {%1,%1};" : "=l"(C0) : "f"(1.0f));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(P) : "l"(C3), "l"(F), "l"(C2));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(P) : "l"(P), "l"(F), "l"(C1));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(P) : "l"(P), "l"(F), "l"(C0));
  float p0, p1, t0, t1;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(p0), "=f"(p1) : "l"(P));
  asm("mov.b64 {%0,%1}, %2;" : "=f"(t0), "=f"(t1) : "l"(T));
  y0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
  y1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
}

// MODE 0: 64 ex2 per iter; 1: 64 poly exp2 (32 pairs); 2: 32 ex2 + 32 poly; 3: 64 cvt.bf16x2 packs (128 floats);
// 4: 64 integer-trick packs; 5: 64 FMNMX3; 6: 64 FFMA2 (scale-sub)
template <int MODE>
__global__ void alu_kernel(long long* out, float* sink, int iters, float seed) {
  const int warp = threadIdx.x >> 5;
  float v[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = seed * (i + 1) - 3.0f - 0.01f * threadIdx.x;
  float acc = 0.f;
  uint32_t iacc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if constexpr (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 64; ++i) v[i] = ex2a(v[i]) - 2.0f;
    } else if constexpr (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        float a, b;
        exp2_poly2(v[i], v[i + 1], a, b);
        v[i] = a - 2.0f;
        v[i + 1] = b - 2.0f;
      }
    } else if constexpr (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        float a, b;
        exp2_poly2(v[i], v[i + 1], a, b);
        v[i] = a - 2.0f;
        v[i + 1] = b - 2.0f;
        v[i + 2] = ex2a(v[i + 2]) - 2.0f;
        v[i + 3] = ex2a(v[i + 3]) - 2.0f;
      }
    } else if constexpr (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        uint32_t pk;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(v[i + 1]), "f"(v[i]));
        iacc ^= pk;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(v[i]), "f"(v[i + 1]));
        iacc += pk;
      }
      v[it & 63] += 1.0f;
    } else if constexpr (MODE == 4) {
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        uint32_t pk = __byte_perm(__float_as_uint(v[i]) + 0x8000u, __float_as_uint(v[i + 1]) + 0x8000u, 0x7632);
        iacc ^= pk;
        pk = __byte_perm(__float_as_uint(v[i + 1]) + 0x8000u, __float_as_uint(v[i]) + 0x8000u, 0x7632);
        iacc += pk;
      }
      v[it & 63] += 1.0f;
    } else if constexpr (MODE == 5) {
      float m0 = acc, m1 = acc, m2 = acc, m3 = acc;
#pragma unroll
      for (int i = 0; i < 64; i += 8) {
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(m0) : "f"(v[i]), "f"(v[i + 1]));
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(m1) : "f"(v[i + 2]), "f"(v[i + 3]));
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(m2) : "f"(v[i + 4]), "f"(v[i + 5]));
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(m3) : "f"(v[i + 6]), "f"(v[i + 7]));
      }
      acc = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * 0.999f;
      // 32 FMNMX3 per iter
    } else {
      unsigned long long S, Bv;
      asm("mov.b64 %0, {%1,%1};" : "=l"(S) : "f"(0.999f));
      asm("mov.b64 %0, {%1,%1};" : "=l"(Bv) : "f"(-0.001f));
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        unsigned long long X;
        asm("mov.b64 %0, {%1,%2};" : "=l"(X) : "f"(v[i]), "f"(v[i + 1]));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(X) : "l"(X), "l"(S), "l"(Bv));
        asm("mov.b64 {%0,%1}, %2;" : "=f"(v[i]), "=f"(v[i + 1]) : "l"(X));
      }
    }
  }
  const long long t1 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) acc += v[i];
  sink[threadIdx.x] = acc + __uint_as_float(iacc);
  if ((threadIdx.x & 31) == 0) out[warp] = t1 - t0;
}

// the softmax inner loop of the attention kernel in isolation: 128 values per thread,
//   x = fma2(s, c, -m) -> ex2 -> sum (add2) -> pack to bf16 pairs
// PACK: 0 = cvt.rn.bf16x2 (F2FP), 1 = integer round + PRMT, 2 = no packing (sum only)
template <int PACK>
__global__ void softmax_loop_kernel(long long* out, float* sink, int iters, float seed) {
  const int warp = threadIdx.x >> 5;
  float s[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) s[i] = seed * (i + 1) - 0.01f * threadIdx.x;
  float acc = 0.f;
  uint32_t iacc = 0;
  unsigned long long c2, nm2;
  asm("mov.b64 %0, {%1,%1};" : "=l"(c2) : "f"(0.18f));
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    asm("mov.b64 %0, {%1,%1};" : "=l"(nm2) : "f"(-1.0f - 1e-3f * it));
    unsigned long long l0 = 0ull, l1 = 0ull;
#pragma unroll
    for (int i = 0; i < 128; i += 4) {
      unsigned long long x0, x1;
      asm("mov.b64 %0, {%1,%2};" : "=l"(x0) : "f"(s[i]), "f"(s[i + 1]));
      asm("mov.b64 %0, {%1,%2};" : "=l"(x1) : "f"(s[i + 2]), "f"(s[i + 3]));
      asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(x0) : "l"(x0), "l"(c2), "l"(nm2));
      asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(x1) : "l"(x1), "l"(c2), "l"(nm2));
      float a0, a1, a2, a3;
      asm("mov.b64 {%0,%1}, %2;" : "=f"(a0), "=f"(a1) : "l"(x0));
      asm("mov.b64 {%0,%1}, %2;" : "=f"(a2), "=f"(a3) : "l"(x1));
      a0 = ex2a(a0); a1 = ex2a(a1); a2 = ex2a(a2); a3 = ex2a(a3);
      unsigned long long e0, e1;
      asm("mov.b64 %0, {%1,%2};" : "=l"(e0) : "f"(a0), "f"(a1));
      asm("mov.b64 %0, {%1,%2};" : "=l"(e1) : "f"(a2), "f"(a3));
      asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(l0) : "l"(l0), "l"(e0));
      asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(l1) : "l"(l1), "l"(e1));
      if constexpr (PACK == 0) {
        uint32_t p0, p1;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p0) : "f"(a1), "f"(a0));
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p1) : "f"(a3), "f"(a2));
        iacc ^= p0 + p1;
      } else if constexpr (PACK == 1) {
        const uint32_t p0 = __byte_perm(__float_as_uint(a0) + 0x8000u, __float_as_uint(a1) + 0x8000u, 0x7632);
        const uint32_t p1 = __byte_perm(__float_as_uint(a2) + 0x8000u, __float_as_uint(a3) + 0x8000u, 0x7632);
        iacc ^= p0 + p1;
      }
    }
    float q0, q1, q2, q3;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(q0), "=f"(q1) : "l"(l0));
    asm("mov.b64 {%0,%1}, %2;" : "=f"(q2), "=f"(q3) : "l"(l1));
    acc += q0 + q1 + q2 + q3;
  }
  const long long t1 = clock64();
  sink[threadIdx.x] = acc + __uint_as_float(iacc);
  if ((threadIdx.x & 31) == 0) out[warp] = t1 - t0;
}

static long long maxof(const std::vector<long long>& v, int n) {
  long long m = 0;
  for (int i = 0; i < n; ++i) m = v[i] > m ? v[i] : m;
  return m;
}

int main() {
  long long* d_out;
  uint32_t* d_sink;
  cudaMalloc(&d_out, 64 * sizeof(long long));
  cudaMalloc(&d_sink, 4096 * sizeof(uint32_t));
  std::vector<long long> h(64);
  const int iters = 2000;
  const char* tn[3] = {"tmem ld x32 wait-each (4 per iter)", "tmem ld 4x x32 then wait", "tmem st 4x x32 then wait"};
  for (int mode = 0; mode < 3; ++mode) {
    for (int warps : {1, 4, 8, 12}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) tmem_bw_kernel<0><<<1, warps * 32>>>(d_out, d_sink, iters);
        if (mode == 1) tmem_bw_kernel<1><<<1, warps * 32>>>(d_out, d_sink, iters);
        if (mode == 2) tmem_bw_kernel<2><<<1, warps * 32>>>(d_out, d_sink, iters);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      }
      cudaMemcpy(h.data(), d_out, 64 * sizeof(long long), cudaMemcpyDeviceToHost);
      const double cyc = (double)maxof(h, warps);
      const double bytes = (double)iters * warps * 4 * 32 * 32 * 4;
      printf("%-40s warps=%2d  cycles/iter=%8.1f  B/clk/SM=%7.1f\n", tn[mode], warps, cyc / iters, bytes / cyc);
    }
  }
  const char* an[7] = {"64 MUFU ex2", "64 poly exp2 (FFMA2)", "32 ex2 + 32 poly", "64 cvt.bf16x2", "64 int-pack",
                       "32 FMNMX3", "32 FFMA2"};
  for (int mode = 0; mode < 7; ++mode) {
    for (int warps : {4, 8, 12}) {
      for (int rep = 0; rep < 2; ++rep) {
        float* fs = reinterpret_cast<float*>(d_sink);
        switch (mode) {
          case 0: alu_kernel<0><<<1, warps * 32>>>(d_out, fs, iters, 0.001f); break;
          case 1: alu_kernel<1><<<1, warps * 32>>>(d_out, fs, iters, 0.001f); break;
          case 2: alu_kernel<2><<<1, warps * 32>>>(d_out, fs, iters, 0.001f); break;
          case 3: alu_kernel<3><<<1, warps * 32>>>(d_out, fs, iters, 0.001f); break;
          case 4: alu_kernel<4><<<1, warps * 32>>>(d_out, fs, iters, 0.001f); break;
          case 5: alu_kernel<5><<<1, warps * 32>>>(d_out, fs, iters, 0.001f); break;
          default: alu_kernel<6><<<1, warps * 32>>>(d_out, fs, iters, 0.001f); break;
        }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      }
      cudaMemcpy(h.data(), d_out, 64 * sizeof(long long), cudaMemcpyDeviceToHost);
      const double cyc = (double)maxof(h, warps);
      printf("%-28s warps=%2d  cycles/iter=%8.1f  (per SMSP-warp-slot: %6.1f)\n", an[mode], warps, cyc / iters,
             cyc / iters / ((warps + 3) / 4));
    }
  }
  const char* sn[3] = {"softmax loop 128 elems (F2FP pack)", "softmax loop 128 elems (int pack)",
                       "softmax loop 128 elems (no pack)"};
  for (int mode = 0; mode < 3; ++mode) {
    for (int warps : {4, 8, 12}) {
      for (int rep = 0; rep < 2; ++rep) {
        float* fs = reinterpret_cast<float*>(d_sink);
        if (mode == 0) softmax_loop_kernel<0><<<1, warps * 32>>>(d_out, fs, iters, 0.01f);
        if (mode == 1) softmax_loop_kernel<1><<<1, warps * 32>>>(d_out, fs, iters, 0.01f);
        if (mode == 2) softmax_loop_kernel<2><<<1, warps * 32>>>(d_out, fs, iters, 0.01f);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      }
      cudaMemcpy(h.data(), d_out, 64 * sizeof(long long), cudaMemcpyDeviceToHost);
      const double cyc = (double)maxof(h, warps);
      printf("%-36s warps=%2d  cycles/iter=%8.1f  (per warp on an SMSP: %7.1f, per element-warp %5.2f)\n", sn[mode],
             warps, cyc / iters, cyc / iters / ((warps + 3) / 4), cyc / iters / ((warps + 3) / 4) / 128.0);
    }
  }
  return 0;
}
