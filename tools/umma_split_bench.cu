// Round-2 feasibility probe (development aid, see tools/README.md): the WaveRNN layer GEMM as split-operand tensor-core
// MMAs.  D[128 utterances x 16 rows] = A[128 x K] . B[16 x K]^T with every fp32 operand split into three bf16 planes
// (v = v0 + v1 + v2) and the six products with i + j <= 2 issued as tcgen05.mma (kind::f16, fp32 accumulation in TMEM).
// Reports (a) the error against float64 next to the error of a plain fp32 fmaf loop and (b) cycles per MMA instruction
// at M = 128, N = 16, K = 16 -- the two unknowns of the plan in DESIGN.md section 3.1.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build_ab/umma_split_bench tools/umma_split_bench.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 16, K = 128, KS = K / 16, PLANES = 3;
constexpr int A_TILE = M * 16, B_TILE = N * 16;                       // bf16 elements of one k-step tile
constexpr int A_ELEMS = PLANES * KS * A_TILE, B_ELEMS = PLANES * KS * B_TILE;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: core matrix = 8 rows x 16 bytes, contiguous 128 B; LBO = stride between the two K halves of one
// MMA, SBO = stride between 8-row groups (cute/atom/mma_traits_sm100.hpp, "LayoutType::INTERLEAVE ((8,n),2):((1,SBO),LBO)")
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;                                            // descriptor version 1 (sm_100)
  return d;                                                          // base offset 0, layout type 0 = no swizzle
}

template <int NACC, int NN>
__global__ void __launch_bounds__(128, 1) umma_kernel(const uint16_t* __restrict__ gA, const uint16_t* __restrict__ gB,
                                                      float* __restrict__ D, int reps, int terms,
                                                      long long* cycles, int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint16_t* sA = reinterpret_cast<uint16_t*>(smem);
  uint16_t* sB = sA + A_ELEMS;
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < A_ELEMS / 8; i += blockDim.x) reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(gA)[i];
  for (int i = tid; i < B_ELEMS / 8; i += blockDim.x) reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(gB)[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // st.shared operands -> visible to the tensor core
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;

  // instruction descriptor: D = f32, A = B = bf16, both K-major, N = 16, M = 128 (cute/arch/mma_sm100_desc.hpp)
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    // small terms first: (2,0) (1,1) (0,2) (1,0) (0,1) (0,0); `terms` = 1 runs the plain bf16 product only
    t0 = clock64();
    // NACC > 1 (timing only): consecutive MMAs go round-robin to NACC independent 16-column accumulators, the way a real
    // kernel would interleave utterance tiles / GEMMs / product classes.  The 48 MMAs of one pass over K are fully
    // unrolled so that the descriptors are base + constant and the single issuing thread is not the bottleneck.
    const uint64_t dA0 = make_desc(a0, (M / 8) * 128, 128), dB0 = make_desc(b0, (NN / 8) * 128, 128);   // NN > N: timing only, B tiles alias
    for (int r = 0; r < reps; ++r) {
      const uint32_t roff = (uint32_t)(r * 48) & (NACC - 1);
#pragma unroll
      for (int s = 0; s < KS; ++s)
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          constexpr int kPi[6] = {2, 1, 0, 1, 0, 0}, kPj[6] = {0, 1, 2, 0, 1, 0};
          if (q < 6 - terms) continue;
          const uint64_t da = dA0 + (uint64_t)((2u * ((kPi[q] * KS + s) * A_TILE)) >> 4);
          const uint64_t db = NN == N ? dB0 + (uint64_t)((2u * ((kPj[q] * KS + s) * B_TILE)) >> 4) : dB0;
          const uint32_t slot = ((uint32_t)(s * 6 + q) + roff) & (NACC - 1);
          const uint32_t acc = (r > 0 || (s * 6 + q) >= NACC + (6 - terms)) ? 1u : 0u;
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem + slot * NN),
              "l"(da), "l"(db), "r"(idesc), "r"(acc));
        }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
  }
  // everybody waits for the MMAs (bounded: a descriptor mistake must not hang the GPU)
  {
    uint32_t done = 0;
    long long ts = clock64();
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
      if (!done && clock64() - ts > 2000000000LL) { if (tid == 0) *status = 1; break; }
    }
    if (tid == 0) t1 = clock64();
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[16];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  const int row = warp * 32 + lane;
#pragma unroll
  for (int c = 0; c < 16; ++c) D[(size_t)blockIdx.x * M * N + row * N + c] = __uint_as_float(v[c]);
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}


// ---- round-2 experiment: K-CHUNKED accumulation.  The tensor core's fp32 accumulator truncates on every
// MMA, and over the 32 k-steps of K = 512 that bias is what makes the 6-product result 7x less accurate than an fp32 FMA loop.
// Here the dominant hi x hi product is accumulated in TMEM only over CH consecutive k-steps (K = 16*CH) per accumulator; the
// 32/CH chunk accumulators and the one accumulator of the five small products are added in fp32 registers (round to nearest).
template <int CH>
__global__ void __launch_bounds__(128, 1) umma_chunk_kernel(const uint16_t* __restrict__ gA, const uint16_t* __restrict__ gB,
                                                            float* __restrict__ D, int reps, int* status) {
  constexpr int NCHUNK = 32 / CH;                      // reps * KS = 32 k-steps in total
  extern __shared__ __align__(1024) uint8_t smem[];
  uint16_t* sA = reinterpret_cast<uint16_t*>(smem);
  uint16_t* sB = sA + A_ELEMS;
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < A_ELEMS / 8; i += blockDim.x) reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(gA)[i];
  for (int i = tid; i < B_ELEMS / 8; i += blockDim.x) reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(gB)[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  if (tid == 0) {
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    const uint64_t dA0 = make_desc(a0, (M / 8) * 128, 128), dB0 = make_desc(b0, (N / 8) * 128, 128);
    bool small_started = false;
    for (int r = 0; r < reps; ++r)
      for (int s = 0; s < KS; ++s) {
        const int kstep = r * KS + s;
        for (int q = 0; q < 6; ++q) {
          const int kPi[6] = {2, 1, 0, 1, 0, 0}, kPj[6] = {0, 1, 2, 0, 1, 0};
          const uint64_t da = dA0 + (uint64_t)((2u * ((kPi[q] * KS + s) * A_TILE)) >> 4);
          const uint64_t db = dB0 + (uint64_t)((2u * ((kPj[q] * KS + s) * B_TILE)) >> 4);
          uint32_t slot, acc;
          if (q == 5) { slot = (uint32_t)(kstep / CH); acc = (kstep % CH) ? 1u : 0u; }           // hi x hi: one accumulator per chunk
          else { slot = (uint32_t)NCHUNK; acc = small_started ? 1u : 0u; small_started = true; }   // the five small products: one accumulator
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem + slot * N),
              "l"(da), "l"(db), "r"(idesc), "r"(acc));
        }
      }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
  }
  {
    uint32_t done = 0;
    long long ts = clock64();
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
      if (!done && clock64() - ts > 2000000000LL) { if (tid == 0) *status = 1; break; }
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  float sum[16], chunk[NCHUNK][16];
  for (int j = 0; j <= NCHUNK; ++j) {
    uint32_t v[16];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(j * N);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 16; ++c) {
      if (j < NCHUNK) chunk[j][c] = __uint_as_float(v[c]);
      else sum[c] = __uint_as_float(v[c]);
    }
  }
  // pairwise tree over the chunk accumulators, then the small-product accumulator
  for (int c = 0; c < 16; ++c) {
    for (int w = 1; w < NCHUNK; w <<= 1)
      for (int j = 0; j + w < NCHUNK; j += 2 * w) chunk[j][c] = __fadd_rn(chunk[j][c], chunk[j + w][c]);
    sum[c] = __fadd_rn(chunk[0][c], sum[c]);
  }
  const int row = warp * 32 + lane;
  for (int c = 0; c < 16; ++c) D[row * N + c] = sum[c];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

static uint16_t bf16_rne(float f) {
  uint32_t u; memcpy(&u, &f, 4);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static float bf16_to_f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }

int main() {
  std::vector<float> A(M * K), B(N * K);
  srand(1);
  auto rnd = []() { float s = 0; for (int i = 0; i < 6; ++i) s += rand() / (float)RAND_MAX - 0.5f; return s; };
  for (auto& x : A) x = rnd();
  for (auto& x : B) x = rnd() * 0.3f;
  std::vector<uint16_t> pA(A_ELEMS), pB(B_ELEMS);
  auto pack = [&](const std::vector<float>& src, int rows, std::vector<uint16_t>& dst) {
    for (int r = 0; r < rows; ++r)
      for (int k = 0; k < K; ++k) {
        float rest = src[r * K + k];
        for (int p = 0; p < PLANES; ++p) {
          uint16_t h = bf16_rne(rest);
          rest -= bf16_to_f(h);
          const int s = k / 16, ki = (k % 16) / 8, k8 = k % 8;
          const size_t tile = (size_t)(p * KS + s) * rows * 16;
          dst[tile + ((size_t)(ki * (rows / 8) + r / 8) * 8 + r % 8) * 8 + k8] = h;
        }
      }
  };
  pack(A, M, pA);
  pack(B, N, pB);

  const int nblk = 128;
  uint16_t *dA, *dB; float* dD; long long* dC; int* dS;
  cudaMalloc(&dA, A_ELEMS * 2); cudaMalloc(&dB, B_ELEMS * 2); cudaMalloc(&dD, sizeof(float) * M * N * nblk);
  cudaMalloc(&dC, 8 * nblk); cudaMalloc(&dS, 4); cudaMemset(dS, 0, 4);
  cudaMemcpy(dA, pA.data(), A_ELEMS * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, pB.data(), B_ELEMS * 2, cudaMemcpyHostToDevice);
  const size_t smem = (size_t)(A_ELEMS + B_ELEMS) * 2;
  cudaFuncSetAttribute(umma_kernel<1, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(umma_kernel<8, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(umma_kernel<32, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(umma_kernel<4, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(umma_kernel<2, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);

  auto run = [&](int reps, int terms, int blocks, std::vector<float>& out, double& cyc, int nacc = 1, int nn = N) {
    if (nn == 64) umma_kernel<4, 64><<<blocks, 128, smem>>>(dA, dB, dD, reps, terms, dC, dS);
    else if (nn == 256) umma_kernel<2, 256><<<blocks, 128, smem>>>(dA, dB, dD, reps, terms, dC, dS);
    else if (nacc == 1) umma_kernel<1, N><<<blocks, 128, smem>>>(dA, dB, dD, reps, terms, dC, dS);
    else if (nacc == 8) umma_kernel<8, N><<<blocks, 128, smem>>>(dA, dB, dD, reps, terms, dC, dS);
    else umma_kernel<32, N><<<blocks, 128, smem>>>(dA, dB, dD, reps, terms, dC, dS);
    cudaError_t e = cudaDeviceSynchronize();
    int st = 0; cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess || st) { printf("kernel failed: %s status=%d\n", cudaGetErrorString(e), st); exit(1); }
    out.resize(M * N); cudaMemcpy(out.data(), dD, sizeof(float) * M * N, cudaMemcpyDeviceToHost);
    std::vector<long long> c(blocks); cudaMemcpy(c.data(), dC, 8 * blocks, cudaMemcpyDeviceToHost);
    cyc = 0; for (auto x : c) cyc = x > cyc ? (double)x : cyc;
  };

  const int reps = 4;                                      // K = 4 x 128 = 512 accumulated into the same tile
  std::vector<double> ref(M * N); std::vector<float> f32(M * N);
  double scale = 0;
  for (int r = 0; r < M; ++r)
    for (int c = 0; c < N; ++c) {
      double s = 0; float f = 0.f;
      for (int rep = 0; rep < reps; ++rep)
        for (int k = 0; k < K; ++k) { s += (double)A[r * K + k] * (double)B[c * K + k]; f = fmaf(A[r * K + k], B[c * K + k], f); }
      ref[r * N + c] = s; f32[r * N + c] = f; scale = fmax(scale, fabs(s));
    }
  auto maxerr = [&](const std::vector<float>& x) { double e = 0; for (int i = 0; i < M * N; ++i) e = fmax(e, fabs(x[i] - ref[i])); return e / scale; };
  std::vector<float> out; double cyc;
  printf("K = %d accumulated, max|ref| = %.3f; errors relative to max|ref|, against float64\n", reps * K, scale);
  printf("  fp32 fmaf loop (CPU)            : %.3e\n", maxerr(f32));
  run(reps, 1, 1, out, cyc);
  printf("  tcgen05 bf16 x bf16 (1 product) : %.3e\n", maxerr(out));
  run(reps, 6, 1, out, cyc);
  printf("  tcgen05 bf16x3, 6 products      : %.3e   (%d MMAs in %.0f cycles)\n", maxerr(out), reps * KS * 6, cyc);
  {
    auto run_chunk = [&](int ch) {
      cudaMemset(dS, 0, 4);
      if (ch == 4) { cudaFuncSetAttribute(umma_chunk_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); umma_chunk_kernel<4><<<1, 128, smem>>>(dA, dB, dD, reps, dS); }
      else if (ch == 2) { cudaFuncSetAttribute(umma_chunk_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); umma_chunk_kernel<2><<<1, 128, smem>>>(dA, dB, dD, reps, dS); }
      else { cudaFuncSetAttribute(umma_chunk_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); umma_chunk_kernel<8><<<1, 128, smem>>>(dA, dB, dD, reps, dS); }
      cudaError_t e = cudaDeviceSynchronize();
      int st = 0; cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess || st) { printf("chunk kernel failed: %s status=%d\n", cudaGetErrorString(e), st); exit(1); }
      out.resize(M * N); cudaMemcpy(out.data(), dD, sizeof(float) * M * N, cudaMemcpyDeviceToHost);
      return maxerr(out);
    };
    // CPU models of the same split arithmetic with IEEE accumulation, to separate splitting error from accumulator truncation
    {
      std::vector<float> m(M * N);
      for (int r = 0; r < M; ++r)
        for (int c = 0; c < N; ++c) {
          float acc = 0.f;
          for (int rep = 0; rep < reps; ++rep)
            for (int k = 0; k < K; ++k) {
              float a = A[r * K + k], b = B[c * K + k], ap[3], bp[3];
              for (int p = 0; p < 3; ++p) { ap[p] = bf16_to_f(bf16_rne(a)); a -= ap[p]; bp[p] = bf16_to_f(bf16_rne(b)); b -= bp[p]; }
              double t = 0;
              for (int i = 0; i < 3; ++i) for (int j = 0; i + j <= 2; ++j) t += (double)ap[i] * bp[j];
              acc = (float)((double)acc + t);
            }
          m[r * N + c] = acc;
        }
      printf("  CPU model: 6 products, fp32 round-to-nearest accumulation over k : %.3e\n", maxerr(m));
    }
    printf("  tcgen05 bf16x3, hi x hi in chunks of K=128 (8 k-steps per accumulator, 4 accumulators) + registers : %.3e\n", run_chunk(8));
    printf("  tcgen05 bf16x3, hi x hi in chunks of K= 64 (4 k-steps per accumulator, 8 accumulators) + registers : %.3e\n", run_chunk(4));
    printf("  tcgen05 bf16x3, hi x hi in chunks of K= 32 (2 k-steps per accumulator, 16 accumulators) + registers: %.3e\n", run_chunk(2));
    printf("  (kill criterion of VERDICT r1 next #8: <= 6e-7 of max|out|)\n");
  }
  for (int blocks : {1, 128}) {
    const int r = 256;
    const double n = (double)r * KS * 6;
    for (int nacc : {1, 8, 32}) {
      run(r, 6, blocks, out, cyc, nacc);
      printf("  timing: %3d CTA(s), %5.0f MMAs (M128 N16 K16) round-robin over %2d accumulator(s): %.1f cycles per MMA, %.0f bf16 MAC/clk/SM\n",
             blocks, n, nacc, cyc / n, n * M * N * 16 / cyc);
    }
    for (int nn : {64, 256}) {                       // same A tiles, wider (aliased) B: is the cost per MMA fixed or per column?
      run(r, 6, blocks, out, cyc, 0, nn);
      printf("  timing: %3d CTA(s), %5.0f MMAs (M128 N%d K16): %.1f cycles per MMA, %.0f bf16 MAC/clk/SM\n", blocks, n, nn, cyc / n,
             n * M * nn * 16 / cyc);
    }
  }
  return 0;
}
