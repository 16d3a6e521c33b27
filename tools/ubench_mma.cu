// tcgen05.mma throughput / issue-cost micro-benchmark on one SM (sm_100a): what the attention kernel's MMA shapes
// really cost.  One thread issues N MMAs + a commit and waits for the commit; optional background warps hammer TMEM
// loads or the MUFU to expose contention.  Build: tools/build_ubench.sh; run: ./tools/ubench_mma
#include "../naturalspeech2_pytorch_b200/csrc/ptx.cuh"
#include <cstdio>
#include <vector>

using namespace ns2;

struct Case {
  const char* name;
  int ts;       // 1: A from TMEM
  int n;        // MMA N
  int b_mn;     // B operand MN-major (V-like)
  int alt;      // alternate between two accumulators
};

// bg: 0 none, 1 = 8 warps looping on TMEM loads (x32 x4 + wait), 2 = 8 warps looping on MUFU ex2
// CONVERGED: the issuing warp runs the loop with all 32 lanes (warp-uniform operands) and elects one lane per
//            instruction, instead of a single-lane divergent region
template <bool CONVERGED>
__global__ void __launch_bounds__(320, 1) mma_kernel(long long* out, Case c, int count, int bg, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
    stop = 0;
  }
  if (warp == 0) tmem_alloc(smem_u32(&holder), 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = holder;
  if (warp == 8) {
    if (CONVERGED || lane == 0) {
      const uint32_t idesc = umma_idesc_f16(128, c.n, 1, 0, c.b_mn);
      const uint32_t a_smem = smem_u32(smem);
      const uint32_t b_smem = smem_u32(smem + 16384);
      const long long t0 = clock64();
      for (int i = 0; i < count; ++i) {
        const int k = i & 3;
        const uint32_t d = tm + ((c.alt && (i & 4)) ? 256 : 0);
        const uint64_t db = c.b_mn ? umma_desc_sw128(b_smem + k * 2048, 1024, 1024)
                                   : (umma_desc_sw128(b_smem, 16, 1024) + 2 * k);
        const uint64_t da = umma_desc_sw128(a_smem, 16, 1024) + 2 * k;
        if (!CONVERGED || elect_one()) {
          if (c.ts) tc_mma_f16_ts(d, tm + 384 + k * 8, db, idesc, 1);
          else tc_mma_f16(d, da, db, idesc, 1);
        }
        if (CONVERGED) __syncwarp();
      }
      const long long t1 = clock64();
      if (!CONVERGED || elect_one()) tc_commit(smem_u32(&bar));
      if (CONVERGED) __syncwarp();
      mbar_wait(smem_u32(&bar), 0);
      const long long t2 = clock64();
      if (lane == 0) {
        out[0] = t1 - t0;
        out[1] = t2 - t0;
        stop = 1;
      }
    }
  } else if (warp < 8 && bg != 0) {
    const uint32_t base = tm + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 128;
    float acc = 0.f;
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = 0.001f * (i + threadIdx.x);
    while (!stop) {
      if (bg == 1) {
        uint32_t r0[32], r1[32], r2[32], r3[32];
        tmem_ld32(base, r0);
        tmem_ld32(base + 32, r1);
        tmem_ld32(base + 64, r2);
        tmem_ld32(base + 96, r3);
        tmem_ld_wait();
        acc += __uint_as_float(r0[0] ^ r1[1] ^ r2[2] ^ r3[3]);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float y;
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v[i]));
          v[i] = y - 1.5f;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += v[i];
    sink[threadIdx.x] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tm, 512);
  }
}

int main() {
  long long* d_out;
  float* d_sink;
  cudaMalloc(&d_out, 16 * sizeof(long long));
  cudaMalloc(&d_sink, 4096 * sizeof(float));
  cudaFuncSetAttribute(mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const Case cases[] = {
      {"SS M128 N128 K16 (S = Q K^T)", 0, 128, 0, 0},
      {"SS M128 N256 K16", 0, 256, 0, 0},
      {"SS M128 N64  K16", 0, 64, 0, 0},
      {"SS M128 N64  K16 B MN-major", 0, 64, 1, 0},
      {"TS M128 N64  K16 B MN-major (O += P V)", 1, 64, 1, 0},
      {"TS M128 N64  K16 B K-major", 1, 64, 0, 0},
      {"TS M128 N128 K16 B K-major", 1, 128, 0, 0},
      {"SS M128 N128 K16 two accumulators", 0, 128, 0, 1},
  };
  const int count = 256;
  for (int conv = 0; conv < 2; ++conv)
  for (int bg = 0; bg < 3; ++bg) {
    printf("--- issuing warp: %s; background: %s\n", conv ? "converged + elect_one" : "single divergent lane",
           bg == 0 ? "none" : (bg == 1 ? "8 warps of TMEM loads" : "8 warps of MUFU ex2"));
    for (const Case& c : cases) {
      long long h[2] = {0, 0};
      for (int rep = 0; rep < 2; ++rep) {
        if (conv) mma_kernel<true><<<1, 320, 64 * 1024>>>(d_out, c, count, bg, d_sink);
        else mma_kernel<false><<<1, 320, 64 * 1024>>>(d_out, c, count, bg, d_sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      }
      cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
      printf("%-44s issue %6.1f clk/MMA   complete %6.1f clk/MMA\n", c.name, (double)h[0] / count,
             (double)h[1] / count);
    }
  }
  return 0;
}
