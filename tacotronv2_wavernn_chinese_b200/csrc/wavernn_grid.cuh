// WaveRNN per-sample recurrence, "grid" kernel: weight-stationary, persistent, cooperative.
//
// The 17.4 MB of fp32 step weights cannot live in one SM (228 KB), but they fit the chip: the grid is NCTA = R/4 = 128
// co-resident CTAs (one per SM, cooperative launch), CTA c permanently holds in SHARED MEMORY the rows of every layer
// that produce hidden units / fc rows [4c, 4c+4) and classes [8c, 8c+8) (136 KB), and all B utterances advance in
// lock step.  Per step each layer is a skinny GEMM  out[B, rows_c] = act[B, K] . W_c[rows_c, K]^T : the activations
// ([K][Bp] fp32, K-major, L2-resident, 2 KB per utterance per layer) are the only thing that moves; weights never do.
// Layers are separated by a grid-wide barrier (monotonic counter in L2, release/acquire).
//
// Replaces the hot loop of WaveRNN.generate, reference wavernn/models/fatchord_version.py:201-237, phase by phase:
//   P0  I        (:208-209)  x|m_t|a1 -> Iout                        4 rows  x 113
//   P1  GRU rnn1 (:210,:212) Iout,h1 -> h1', x1 = Iout + h1'         24 rows x 512
//   P2  GRU rnn2 (:213-216)  x1|a2,h2 -> h2', x2 = x1 + h2'          12 x 544 + 12 x 512
//   P3  fc1+relu (:217-218)  x2|a3 -> f1                              4 rows x 544
//   P4  fc2+relu (:220-221)  f1|a4 -> f2                              4 rows x 544
//   P5  fc3      (:223) + sampling (:232-235): every CTA owns 8 logits per utterance and joins a distributed
//       argmax of (logit - log q), q ~ Exp(1), through one 64-bit atomicMax per (CTA, utterance)  [Gumbel-max ==
//       Categorical(softmax(logits)).sample()]; the winner is read back by everybody at the next P0 (:235-237).
// Two thread mappings: "wide" (lanes = utterances, register tile U utterances x RT rows, k split across warps) for
// B >= 9, and "narrow" (lanes = k, warp-shuffle reductions) for B <= 8 where the step is pure latency.
#pragma once
#include "common.cuh"

namespace b200tts {

constexpr int kGridThreads = 256;
constexpr int kGridWarps = kGridThreads / 32;
constexpr int kUPC = 4;   // hidden units (and fc1/fc2 rows) per CTA
constexpr int kCPC = 8;   // classes (fc3 rows) per CTA

struct GridModel {        // layout of one CTA's weight blob (offsets in floats, every array 16-byte aligned)
  int ncta, R, F, AUX, FEAT, NC;
  int ldC;                // feat + aux   (cond columns of I, multiple of 4)
  int ldX;                // R + aux
  int ldF;                // F + aux
  int oI_w, oI_x, oI_b;                            // [4][ldC], [4], [4]
  int oih1, ohh1, oih2, ohh2;                      // [12][R], [12][R], [12][ldX], [12][R]   row = gate*4 + unit
  int ofc1, ofc2, ofc3;                            // [4][ldX], [4][ldF], [8][F]
  int obih1, obhh1, obih2, obhh2, obfc1, obfc2, obfc3;
  int blob;                                        // floats per CTA
  int ok;                                          // model fits this kernel
};

struct GridArgs {
  const float* wblob;          // [ncta][blob]
  float* Iout;                 // [R][Bp]
  float* h1;                   // [2][R][Bp]
  float* h2;                   // [2][R][Bp]
  float* x1;                   // [R][Bp]
  float* x2;                   // [R][Bp]
  float* f1;                   // [F][Bp]
  float* f2;                   // [F][Bp]
  unsigned long long* best;    // [2][Bp]   packed argmax per utterance, ping-pong by step parity
  unsigned int* barrier;       // monotonic arrival counter
  int* error;                  // set non-zero if a barrier wait timed out
  const float* mels_T;         // [S][FEAT][Bp]
  const float* aux_T;          // [T][4*AUX][Bp]
  int B, Bp, S, T, hop, steps;
  int rng_mode;
  unsigned long long seed, utt_offset;
  const float* q;              // [S][B][NC]
  const int16_t* teacher;      // [B][S]
  float* logits_out;           // [S][B][NC]
  int16_t* labels;             // [B][S]
};

// ---- grid-wide barrier --------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// All threads call.  `target` = (number of barriers passed so far + 1) * gridDim.x.  Returns false on timeout.
__device__ __forceinline__ bool grid_barrier(unsigned int* ctr, unsigned int target, int* error) {
  __shared__ int s_ok;
  __syncthreads();
  if (threadIdx.x == 0) {
    red_release_add_u32(ctr, 1u);                 // release: orders this CTA's prior global writes (cumulative via bar.sync)
    int ok = 1;
    long long t0 = clock64();
    while (ld_acquire_u32(ctr) < target) {
      if (clock64() - t0 > 4000000000LL) {        // ~2 s: a peer CTA is gone; bail out instead of hanging the GPU
        ok = 0;
        atomicExch(error, 1);
        break;
      }
    }
    s_ok = ok;
  }
  __syncthreads();
  return s_ok != 0;
}

// ---- activation loads (L2 only: these buffers are rewritten by other SMs every step) ----------------------------------
template <int U> struct ActLoad;
template <> struct ActLoad<1> {
  static __device__ __forceinline__ void ld(const float* p, float (&a)[1]) { a[0] = __ldcg(p); }
  static __device__ __forceinline__ void st(float* p, const float (&a)[1]) { p[0] = a[0]; }
};
template <> struct ActLoad<2> {
  static __device__ __forceinline__ void ld(const float* p, float (&a)[2]) {
    float2 v = __ldcg(reinterpret_cast<const float2*>(p)); a[0] = v.x; a[1] = v.y;
  }
  static __device__ __forceinline__ void st(float* p, const float (&a)[2]) { *reinterpret_cast<float2*>(p) = make_float2(a[0], a[1]); }
};
template <> struct ActLoad<4> {
  static __device__ __forceinline__ void ld(const float* p, float (&a)[4]) {
    float4 v = __ldcg(reinterpret_cast<const float4*>(p)); a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
  }
  static __device__ __forceinline__ void st(float* p, const float (&a)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(a[0], a[1], a[2], a[3]);
  }
};
template <> struct ActLoad<8> {
  static __device__ __forceinline__ void ld(const float* p, float (&a)[8]) {
    float4 v = __ldcg(reinterpret_cast<const float4*>(p)), w = __ldcg(reinterpret_cast<const float4*>(p) + 1);
    a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w; a[4] = w.x; a[5] = w.y; a[6] = w.z; a[7] = w.w;
  }
};

struct Seg {            // a run of activation rows: act[4*n4][Bp], matching 4*n4 consecutive weight columns
  const float* act;
  int n4;
};
struct Gemm {           // rows x (sum of segs) weight block in shared memory
  const float* W;
  int ldw;              // floats, multiple of 4
  Seg seg[2];
  int nseg;
};

// acc[r][u] += sum_{c4 in [lo,hi)} W[r][4*(col4+c4) .. +3] . act[4*c4 .. +3][u0 .. u0+U)
template <int U, int RT>
__device__ __forceinline__ void wide_accumulate(float (&acc)[RT][U], const float* __restrict__ W, int ldw, int col4,
                                                const float* __restrict__ act, int Bp, int u0, int lo, int hi) {
  const float4* W4 = reinterpret_cast<const float4*>(W) + col4;
  const int ldw4 = ldw >> 2;
#pragma unroll 2
  for (int c4 = lo; c4 < hi; ++c4) {
    float a[4][U];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) ActLoad<U>::ld(act + (size_t)(4 * c4 + kk) * Bp + u0, a[kk]);
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      float4 w = W4[r * ldw4 + c4];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        acc[r][u] = fmaf(w.x, a[0][u], acc[r][u]);
        acc[r][u] = fmaf(w.y, a[1][u], acc[r][u]);
        acc[r][u] = fmaf(w.z, a[2][u], acc[r][u]);
        acc[r][u] = fmaf(w.w, a[3][u], acc[r][u]);
      }
    }
  }
}

// Wide mapping: NG GEMMs of RT rows each; 8 warps = NG x UW (utterance warps) x KS (k slices).
// Partial sums land in part[((g*KS + ks)*RT + r)*BT + ul].
template <int U, int UW, int RT, int NG>
__device__ __forceinline__ void wide_partials(float* part, const Gemm& g0, const Gemm& g1, int tile_base, int Bp, int warp,
                                              int lane) {
  constexpr int KS = kGridWarps / (NG * UW);
  constexpr int BT = 32 * U * UW;
  static_assert(KS >= 1, "too many jobs for 8 warps");
  const int g = warp / (UW * KS), rem = warp % (UW * KS), uw = rem / KS, ks = rem % KS;
  const Gemm& G = (NG == 2 && g == 1) ? g1 : g0;
  const int ul = uw * 32 * U + lane * U;
  float acc[RT][U];
#pragma unroll
  for (int r = 0; r < RT; ++r)
#pragma unroll
    for (int u = 0; u < U; ++u) acc[r][u] = 0.f;
  int N4 = G.seg[0].n4 + (G.nseg > 1 ? G.seg[1].n4 : 0);
  const int lo = N4 * ks / KS, hi = N4 * (ks + 1) / KS;
  int col = 0;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    if (s < G.nseg) {
      int a = max(lo, col), b = min(hi, col + G.seg[s].n4);
      if (a < b) wide_accumulate<U, RT>(acc, G.W, G.ldw, col, G.seg[s].act, Bp, tile_base + ul, a - col, b - col);
      col += G.seg[s].n4;
    }
  }
  float* dst = part + (size_t)((g * KS + ks) * RT) * BT + ul;
#pragma unroll
  for (int r = 0; r < RT; ++r) ActLoad<U>::st(dst + r * BT, acc[r]);
}

// Narrow mapping (Bp == G <= 8): lanes stride over float4 columns, RT rows per warp pass, shuffle reduction.
// Warps [wbeg, wbeg+wcnt) take part.  Result in part[(gslot*NR + r)*G + u].
template <int G, int RT>
__device__ __forceinline__ void narrow_rows(float* part, int gslot, const Gemm& Gm, int nrows, int wbeg, int wcnt, int warp,
                                            int lane) {
  if (warp < wbeg || warp >= wbeg + wcnt) return;
  const int ldw4 = Gm.ldw >> 2;
  for (int r0 = (warp - wbeg) * RT; r0 < nrows; r0 += wcnt * RT) {
    float acc[RT][G];
#pragma unroll
    for (int r = 0; r < RT; ++r)
#pragma unroll
      for (int u = 0; u < G; ++u) acc[r][u] = 0.f;
    int col = 0;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      if (s < Gm.nseg) {
        const float4* W4 = reinterpret_cast<const float4*>(Gm.W) + (size_t)r0 * ldw4 + col;
        const float* act = Gm.seg[s].act;
        for (int c4 = lane; c4 < Gm.seg[s].n4; c4 += 32) {
          float a[4][G];
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) ActLoad<G>::ld(act + (size_t)(4 * c4 + kk) * G, a[kk]);
#pragma unroll
          for (int r = 0; r < RT; ++r) {
            float4 w = W4[r * ldw4 + c4];
#pragma unroll
            for (int u = 0; u < G; ++u) {
              acc[r][u] = fmaf(w.x, a[0][u], acc[r][u]);
              acc[r][u] = fmaf(w.y, a[1][u], acc[r][u]);
              acc[r][u] = fmaf(w.z, a[2][u], acc[r][u]);
              acc[r][u] = fmaf(w.w, a[3][u], acc[r][u]);
            }
          }
        }
        col += Gm.seg[s].n4;
      }
    }
#pragma unroll
    for (int r = 0; r < RT; ++r)
#pragma unroll
      for (int u = 0; u < G; ++u) {
        float v = warp_sum(acc[r][u]);
        if (lane == 0) part[(size_t)(gslot * nrows + r0 + r) * G + u] = v;
      }
  }
}

// Mapping traits.  U == 0 selects the narrow mapping with G = UW utterances.
template <int U, int UW> struct MapTraits {
  static constexpr bool kWide = true;
  static constexpr int BT = 32 * U * UW;                                   // utterances per tile
  static constexpr int KS1 = kGridWarps / UW;                              // k slices of a 1-GEMM phase
  static constexpr int KS2 = kGridWarps / (2 * UW);                        // ... of a 2-GEMM phase
};
template <int G> struct MapTraits<0, G> {
  static constexpr bool kWide = false;
  static constexpr int BT = G;
  static constexpr int KS1 = 1;
  static constexpr int KS2 = 1;
};

// sum over k slices of one output: gemm slot g, row r (of RT), local utterance ul
template <int KS, int RT, int BT>
__device__ __forceinline__ float part_sum(const float* part, int g, int r, int ul) {
  float v = 0.f;
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) v += part[(size_t)((g * KS + ks) * RT + r) * BT + ul];
  return v;
}

template <int U, int UW>
__global__ void __launch_bounds__(kGridThreads, 1) wavernn_grid_kernel(GridModel M, GridArgs A) {
  using MT = MapTraits<U, UW>;
  constexpr int BT = MT::BT;
  constexpr int KS1 = MT::KS1;
  constexpr int KS2 = MT::KS2;
  extern __shared__ __align__(16) float smem[];
  float* Wb = smem;                                   // this CTA's weights, resident for the whole kernel
  float* part = smem + M.blob;                        // partial sums of the current phase
  __shared__ float xs[BT > 32 ? BT : 32];             // fed-back sample of the tile's utterances

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = blockIdx.x, Bp = A.Bp;
  const int R = M.R, F = M.F, AUX = M.AUX;
  const float ncls_m1 = (float)(M.NC - 1);
  {
    const float4* src = reinterpret_cast<const float4*>(A.wblob + (size_t)c * M.blob);
    float4* dst = reinterpret_cast<float4*>(Wb);
    for (int i = tid; i < M.blob / 4; i += kGridThreads) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  const float* bI = Wb + M.oI_b;
  const float* wIx = Wb + M.oI_x;
  unsigned int nbar = 0;
  const unsigned int ncta = gridDim.x;
  const size_t RB = (size_t)R * Bp;

  for (int t = 0; t < A.steps; ++t) {
    const int cur = t & 1, fr = t / A.hop;
    const float* auxT = A.aux_T + (size_t)fr * 4 * AUX * Bp;
    const float* h1c = A.h1 + cur * RB;
    float* h1n = A.h1 + (cur ^ 1) * RB;
    const float* h2c = A.h2 + cur * RB;
    float* h2n = A.h2 + (cur ^ 1) * RB;

    // ================= P0: read back the previous step's winner, then the I layer =================
    for (int tb = 0; tb < Bp; tb += BT) {
      for (int ul = tid; ul < BT; ul += kGridThreads) {
        const int u = tb + ul;
        float x = 0.f;
        if (t > 0 && u < A.B) {
          unsigned long long pk = __ldcg(A.best + (size_t)((t - 1) & 1) * Bp + u);
          int label = (int)unpack_idx(pk);
          if (c == 0) A.labels[(size_t)u * A.S + (t - 1)] = (int16_t)label;
          int fb = A.teacher ? (int)A.teacher[(size_t)u * A.S + (t - 1)] : label;
          x = label_to_float(fb, ncls_m1);
        }
        xs[ul] = x;
      }
      Gemm g{};
      g.W = Wb + M.oI_w; g.ldw = M.ldC; g.nseg = 2;
      g.seg[0] = Seg{A.mels_T + (size_t)t * M.FEAT * Bp, M.FEAT / 4};
      g.seg[1] = Seg{auxT, AUX / 4};
      if constexpr (MT::kWide) wide_partials<U, UW, kUPC, 1>(part, g, g, tb, Bp, warp, lane);
      else narrow_rows<UW, 1>(part, 0, g, kUPC, 0, kGridWarps, warp, lane);
      __syncthreads();
      for (int idx = tid; idx < BT * kUPC; idx += kGridThreads) {
        const int ul = idx % BT, j = idx / BT;
        float v = part_sum<KS1, kUPC, BT>(part, 0, j, ul);
        v = fmaf(wIx[j], xs[ul], v) + bI[j];
        A.Iout[(size_t)(c * kUPC + j) * Bp + tb + ul] = v;
      }
      __syncthreads();
    }
    if (!grid_barrier(A.barrier, (++nbar) * ncta, A.error)) return;

    // ================= P1: GRU 1 =================
    for (int tb = 0; tb < Bp; tb += BT) {
      Gemm gi{}, gh{};
      gi.W = Wb + M.oih1; gi.ldw = R; gi.nseg = 1; gi.seg[0] = Seg{A.Iout, R / 4};
      gh.W = Wb + M.ohh1; gh.ldw = R; gh.nseg = 1; gh.seg[0] = Seg{h1c, R / 4};
      if constexpr (MT::kWide) wide_partials<U, UW, 3 * kUPC, 2>(part, gi, gh, tb, Bp, warp, lane);
      else {
        narrow_rows<UW, 3>(part, 0, gi, 3 * kUPC, 0, 4, warp, lane);
        narrow_rows<UW, 3>(part, 1, gh, 3 * kUPC, 4, 4, warp, lane);
      }
      __syncthreads();
      const float* bih = Wb + M.obih1; const float* bhh = Wb + M.obhh1;
      for (int idx = tid; idx < BT * kUPC; idx += kGridThreads) {
        const int ul = idx % BT, j = idx / BT;
        const size_t o = (size_t)(c * kUPC + j) * Bp + tb + ul;
        float gir = part_sum<KS2, 3 * kUPC, BT>(part, 0, j, ul) + bih[j];
        float giz = part_sum<KS2, 3 * kUPC, BT>(part, 0, kUPC + j, ul) + bih[kUPC + j];
        float gin = part_sum<KS2, 3 * kUPC, BT>(part, 0, 2 * kUPC + j, ul) + bih[2 * kUPC + j];
        float ghr = part_sum<KS2, 3 * kUPC, BT>(part, 1, j, ul) + bhh[j];
        float ghz = part_sum<KS2, 3 * kUPC, BT>(part, 1, kUPC + j, ul) + bhh[kUPC + j];
        float ghn = part_sum<KS2, 3 * kUPC, BT>(part, 1, 2 * kUPC + j, ul) + bhh[2 * kUPC + j];
        float r = sigmoidf_acc(gir + ghr), z = sigmoidf_acc(giz + ghz);
        float n = tanhf(gin + r * ghn);
        float h = (1.0f - z) * n + z * __ldcg(h1c + o);
        h1n[o] = h;
        A.x1[o] = __ldcg(A.Iout + o) + h;
      }
      __syncthreads();
    }
    if (!grid_barrier(A.barrier, (++nbar) * ncta, A.error)) return;

    // ================= P2: GRU 2 =================
    for (int tb = 0; tb < Bp; tb += BT) {
      Gemm gi{}, gh{};
      gi.W = Wb + M.oih2; gi.ldw = M.ldX; gi.nseg = 2; gi.seg[0] = Seg{A.x1, R / 4}; gi.seg[1] = Seg{auxT + (size_t)AUX * Bp, AUX / 4};
      gh.W = Wb + M.ohh2; gh.ldw = R; gh.nseg = 1; gh.seg[0] = Seg{h2c, R / 4};
      if constexpr (MT::kWide) wide_partials<U, UW, 3 * kUPC, 2>(part, gi, gh, tb, Bp, warp, lane);
      else {
        narrow_rows<UW, 3>(part, 0, gi, 3 * kUPC, 0, 4, warp, lane);
        narrow_rows<UW, 3>(part, 1, gh, 3 * kUPC, 4, 4, warp, lane);
      }
      __syncthreads();
      const float* bih = Wb + M.obih2; const float* bhh = Wb + M.obhh2;
      for (int idx = tid; idx < BT * kUPC; idx += kGridThreads) {
        const int ul = idx % BT, j = idx / BT;
        const size_t o = (size_t)(c * kUPC + j) * Bp + tb + ul;
        float gir = part_sum<KS2, 3 * kUPC, BT>(part, 0, j, ul) + bih[j];
        float giz = part_sum<KS2, 3 * kUPC, BT>(part, 0, kUPC + j, ul) + bih[kUPC + j];
        float gin = part_sum<KS2, 3 * kUPC, BT>(part, 0, 2 * kUPC + j, ul) + bih[2 * kUPC + j];
        float ghr = part_sum<KS2, 3 * kUPC, BT>(part, 1, j, ul) + bhh[j];
        float ghz = part_sum<KS2, 3 * kUPC, BT>(part, 1, kUPC + j, ul) + bhh[kUPC + j];
        float ghn = part_sum<KS2, 3 * kUPC, BT>(part, 1, 2 * kUPC + j, ul) + bhh[2 * kUPC + j];
        float r = sigmoidf_acc(gir + ghr), z = sigmoidf_acc(giz + ghz);
        float n = tanhf(gin + r * ghn);
        float h = (1.0f - z) * n + z * __ldcg(h2c + o);
        h2n[o] = h;
        A.x2[o] = __ldcg(A.x1 + o) + h;
      }
      __syncthreads();
    }
    if (!grid_barrier(A.barrier, (++nbar) * ncta, A.error)) return;

    // ================= P3: fc1 + relu  (CTA 0 also recycles the argmax slot the NEXT step will use) =================
    if (c == 0)
      for (int u = tid; u < Bp; u += kGridThreads) A.best[(size_t)((t + 1) & 1) * Bp + u] = 0ull;
    for (int tb = 0; tb < Bp; tb += BT) {
      Gemm g{};
      g.W = Wb + M.ofc1; g.ldw = M.ldX; g.nseg = 2; g.seg[0] = Seg{A.x2, R / 4}; g.seg[1] = Seg{auxT + (size_t)2 * AUX * Bp, AUX / 4};
      if constexpr (MT::kWide) wide_partials<U, UW, kUPC, 1>(part, g, g, tb, Bp, warp, lane);
      else narrow_rows<UW, 1>(part, 0, g, kUPC, 0, kGridWarps, warp, lane);
      __syncthreads();
      const float* b = Wb + M.obfc1;
      for (int idx = tid; idx < BT * kUPC; idx += kGridThreads) {
        const int ul = idx % BT, j = idx / BT;
        A.f1[(size_t)(c * kUPC + j) * Bp + tb + ul] = fmaxf(part_sum<KS1, kUPC, BT>(part, 0, j, ul) + b[j], 0.f);
      }
      __syncthreads();
    }
    if (!grid_barrier(A.barrier, (++nbar) * ncta, A.error)) return;

    // ================= P4: fc2 + relu =================
    for (int tb = 0; tb < Bp; tb += BT) {
      Gemm g{};
      g.W = Wb + M.ofc2; g.ldw = M.ldF; g.nseg = 2; g.seg[0] = Seg{A.f1, F / 4}; g.seg[1] = Seg{auxT + (size_t)3 * AUX * Bp, AUX / 4};
      if constexpr (MT::kWide) wide_partials<U, UW, kUPC, 1>(part, g, g, tb, Bp, warp, lane);
      else narrow_rows<UW, 1>(part, 0, g, kUPC, 0, kGridWarps, warp, lane);
      __syncthreads();
      const float* b = Wb + M.obfc2;
      for (int idx = tid; idx < BT * kUPC; idx += kGridThreads) {
        const int ul = idx % BT, j = idx / BT;
        A.f2[(size_t)(c * kUPC + j) * Bp + tb + ul] = fmaxf(part_sum<KS1, kUPC, BT>(part, 0, j, ul) + b[j], 0.f);
      }
      __syncthreads();
    }
    if (!grid_barrier(A.barrier, (++nbar) * ncta, A.error)) return;

    // ================= P5: fc3 + distributed Gumbel-max sampling =================
    for (int tb = 0; tb < Bp; tb += BT) {
      Gemm g{};
      g.W = Wb + M.ofc3; g.ldw = F; g.nseg = 1; g.seg[0] = Seg{A.f2, F / 4};
      if constexpr (MT::kWide) wide_partials<U, UW, kCPC, 1>(part, g, g, tb, Bp, warp, lane);
      else narrow_rows<UW, 1>(part, 0, g, kCPC, 0, kGridWarps, warp, lane);
      __syncthreads();
      const float* b = Wb + M.obfc3;
      for (int ul = tid; ul < BT; ul += kGridThreads) {
        const int u = tb + ul;
        if (u < A.B) {
          unsigned long long bestp = 0ull;
#pragma unroll
          for (int r4 = 0; r4 < kCPC / 4; ++r4) {
            float q[4];
            const int cls0 = c * kCPC + r4 * 4;
            if (A.rng_mode == 0) {
              philox_exp4(A.seed, A.utt_offset + (unsigned long long)u, (uint32_t)t, (uint32_t)(cls0 >> 2), q);
            } else {
              float4 qv = __ldg(reinterpret_cast<const float4*>(A.q + ((size_t)t * A.B + u) * M.NC + cls0));
              q[0] = qv.x; q[1] = qv.y; q[2] = qv.z; q[3] = qv.w;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int r = r4 * 4 + k;
              float l = part_sum<KS1, kCPC, BT>(part, 0, r, ul) + b[r];
              if (A.logits_out) A.logits_out[((size_t)t * A.B + u) * M.NC + cls0 + k] = l;
              unsigned long long p = pack_key(l - logf(q[k]), (uint32_t)(cls0 + k));
              bestp = p > bestp ? p : bestp;
            }
          }
          atomicMax(A.best + (size_t)(t & 1) * Bp + u, bestp);
        }
      }
      __syncthreads();
    }
    if (!grid_barrier(A.barrier, (++nbar) * ncta, A.error)) return;
  }
  // the last step's winner
  if (c == 0 && A.steps > 0) {
    const int t = A.steps;
    for (int u = tid; u < A.B; u += kGridThreads) {
      unsigned long long pk = __ldcg(A.best + (size_t)((t - 1) & 1) * Bp + u);
      A.labels[(size_t)u * A.S + (t - 1)] = (int16_t)unpack_idx(pk);
    }
  }
}

// conditioning in the K-major layout the grid kernel streams: mels_T[t][c][u], aux_T[fr][o][u]
__global__ void mel_fir_T_kernel(const float* __restrict__ mel /*[B][feat][T]*/, const float* __restrict__ fir, int B, int Bp,
                                 int T, int feat, int hop, int pad, int NT, float* __restrict__ mels_T /*[S][feat][Bp]*/) {
  const size_t S = (size_t)T * hop, total = S * feat * Bp;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    int u = (int)(e % Bp);
    size_t rest = e / Bp;
    int c = (int)(rest % feat);
    size_t n = rest / feat;
    float acc = 0.f;
    if (u < B) {
      size_t np = n + (size_t)pad * hop;
      int fr = (int)(np / hop), ph = (int)(np % hop);
      for (int j = 0; j < NT; ++j) {
        int f = fr + j - NT / 2 - pad;
        if (f >= 0 && f < T) acc = fmaf(fir[ph * NT + j], mel[((size_t)u * feat + c) * T + f], acc);
      }
    }
    mels_T[e] = acc;
  }
}

__global__ void aux_T_kernel(const float* __restrict__ aux_frames /*[B][T][O]*/, int B, int Bp, int T, int O,
                             float* __restrict__ aux_T /*[T][O][Bp]*/) {
  const size_t total = (size_t)T * O * Bp;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    int u = (int)(e % Bp);
    size_t rest = e / Bp;
    int o = (int)(rest % O);
    size_t f = rest / O;
    aux_T[e] = u < B ? aux_frames[((size_t)u * T + f) * O + o] : 0.f;
  }
}

}  // namespace b200tts
