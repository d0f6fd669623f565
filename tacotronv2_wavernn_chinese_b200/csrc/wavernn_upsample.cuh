// WaveRNN conditioning network on the GPU.
//
// Replaces UpsampleNetwork.forward (reference wavernn/models/fatchord_version.py:82-89):
//   * MelResNet (:31-48, ResBlock :13-28) evaluated at FRAME rate -- the reference stretches its output x hop
//     afterwards (:84), so aux is constant within a hop and is never materialised at sample rate here.
//   * the three Stretch2d + Conv2d stages (:73-80, :86-87) collapsed into one polyphase FIR of `hop` phases x NT
//     taps over the zero-padded frame sequence.  The composite taps are derived on the host (double precision)
//     by pushing an impulse through the stages; the stages' own zero padding only disturbs samples inside the
//     `indent` that :88 trims away (reach 341 samples < indent 550 for factors (5,5,11)).
#pragma once
#include "common.cuh"

namespace b200tts {

constexpr int kMaxTaps = 7;

struct ResnetParams {
  // all weights are transposed to [in][out] so that thread `c` (output channel) reads coalesced
  const float* conv_in_t;   // [feat*k][C]   (index (i*k + j)*C + c)
  const float* bn_scale;    // [(1 + 2*blocks)][C]   gamma / sqrt(var + eps)
  const float* bn_shift;    // [(1 + 2*blocks)][C]   beta - mean * scale
  const float* res_w_t;     // [blocks][2][C][C]     (in-major)
  const float* conv_out_t;  // [C][O]
  const float* conv_out_b;  // [O]
  int feat, k, C, O, blocks, pad;
};

// One CTA = FT frames of one utterance; thread c owns channel c for all FT frames.
template <int FT>
__global__ void melresnet_kernel(ResnetParams P, const float* __restrict__ mel /*[B][feat][T]*/,
                                                        int T, float* __restrict__ aux_frames /*[B][T][O]*/) {
  extern __shared__ float sm[];
  const int C = P.C;
  float* xin = sm;                              // [feat*k][FT]
  float* act = xin + P.feat * P.k * FT;         // [C][FT]
  float* tmp = act + C * FT;                    // [C][FT]
  const int b = blockIdx.y, f0 = blockIdx.x * FT, c = threadIdx.x;
  const float* melb = mel + (size_t)b * P.feat * T;
  // gather the k-frame windows; frame index into the padded sequence is f + j, i.e. unpadded f + j - pad
  for (int e = threadIdx.x; e < P.feat * P.k * FT; e += blockDim.x) {
    int ft = e % FT, ij = e / FT, j = ij % P.k, i = ij / P.k;
    int fr = f0 + ft + j - P.pad;
    xin[e] = (fr >= 0 && fr < T && f0 + ft < T) ? melb[(size_t)i * T + fr] : 0.f;
  }
  __syncthreads();
  float acc[FT];
  if (c < C) {
#pragma unroll
    for (int t = 0; t < FT; ++t) acc[t] = 0.f;
    for (int e = 0; e < P.feat * P.k; ++e) {
      float w = P.conv_in_t[(size_t)e * C + c];
#pragma unroll
      for (int t = 0; t < FT; ++t) acc[t] = fmaf(w, xin[e * FT + t], acc[t]);
    }
    float s = P.bn_scale[c], h = P.bn_shift[c];
#pragma unroll
    for (int t = 0; t < FT; ++t) act[c * FT + t] = fmaxf(fmaf(acc[t], s, h), 0.f);
  }
  __syncthreads();
  for (int blk = 0; blk < P.blocks; ++blk) {
    const float* w1 = P.res_w_t + (size_t)(blk * 2) * C * C;
    const float* w2 = w1 + (size_t)C * C;
    if (c < C) {
#pragma unroll
      for (int t = 0; t < FT; ++t) acc[t] = 0.f;
      for (int i = 0; i < C; ++i) {
        float w = w1[(size_t)i * C + c];
#pragma unroll
        for (int t = 0; t < FT; ++t) acc[t] = fmaf(w, act[i * FT + t], acc[t]);
      }
      float s = P.bn_scale[(1 + 2 * blk) * C + c], h = P.bn_shift[(1 + 2 * blk) * C + c];
#pragma unroll
      for (int t = 0; t < FT; ++t) tmp[c * FT + t] = fmaxf(fmaf(acc[t], s, h), 0.f);
    }
    __syncthreads();
    if (c < C) {
#pragma unroll
      for (int t = 0; t < FT; ++t) acc[t] = 0.f;
      for (int i = 0; i < C; ++i) {
        float w = w2[(size_t)i * C + c];
#pragma unroll
        for (int t = 0; t < FT; ++t) acc[t] = fmaf(w, tmp[i * FT + t], acc[t]);
      }
      float s = P.bn_scale[(2 + 2 * blk) * C + c], h = P.bn_shift[(2 + 2 * blk) * C + c];
      // act[c][*] is only re-read by its own thread here (residual); everybody else reads tmp -> no hazard
#pragma unroll
      for (int t = 0; t < FT; ++t) act[c * FT + t] = fmaf(acc[t], s, h) + act[c * FT + t];
    }
    __syncthreads();
  }
  for (int o = threadIdx.x; o < P.O; o += blockDim.x) {
#pragma unroll
    for (int t = 0; t < FT; ++t) acc[t] = 0.f;
    for (int i = 0; i < C; ++i) {
      float w = P.conv_out_t[(size_t)i * P.O + o];
#pragma unroll
      for (int t = 0; t < FT; ++t) acc[t] = fmaf(w, act[i * FT + t], acc[t]);
    }
    float bias = P.conv_out_b[o];
#pragma unroll
    for (int t = 0; t < FT; ++t)
      if (f0 + t < T) aux_frames[((size_t)b * T + f0 + t) * P.O + o] = acc[t] + bias;
  }
}

// mels_up[b][n][c] = sum_j fir[ph][j] * melpad[b][c][fr + j - NT/2],  n' = n + pad*hop, fr = n'/hop, ph = n'%hop
__global__ void mel_fir_kernel(const float* __restrict__ mel /*[B][feat][T]*/, const float* __restrict__ fir /*[hop][NT]*/,
                               int T, int feat, int hop, int pad, int NT, float* __restrict__ mels_up /*[B][T*hop][feat]*/) {
  const int b = blockIdx.y;
  const size_t S = (size_t)T * hop;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < S * feat; e += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(e % feat);
    size_t n = e / feat;
    size_t np = n + (size_t)pad * hop;
    int fr = (int)(np / hop), ph = (int)(np % hop);
    float acc = 0.f;
    for (int j = 0; j < NT; ++j) {
      int f = fr + j - NT / 2 - pad;   // unpadded frame index
      if (f >= 0 && f < T) acc = fmaf(fir[ph * NT + j], mel[((size_t)b * feat + c) * T + f], acc);
    }
    mels_up[((size_t)b * S + n) * feat + c] = acc;
  }
}

__global__ void aux_repeat_kernel(const float* __restrict__ aux_frames, int T, int hop, int O, float* __restrict__ aux_full) {
  const int b = blockIdx.y;
  const size_t S = (size_t)T * hop;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < S * O; e += (size_t)gridDim.x * blockDim.x) {
    int o = (int)(e % O);
    size_t n = e / O;
    aux_full[((size_t)b * S + n) * O + o] = aux_frames[((size_t)b * T + n / hop) * O + o];
  }
}

// generate() epilogue (fatchord_version.py:243-258): float64, decode_mu_law (dsp.py:98-103), truncate, 20-hop fade.
__global__ void finish_wave_kernel(const int16_t* __restrict__ labels /*[B][S]*/, int S, int wave_len, int fade_len,
                                   int ncls, int mu_law, double* __restrict__ wave /*[B][wave_len]*/) {
  const int b = blockIdx.y;
  const double mu = (double)(ncls - 1);
  const double step = -1.0 / (double)(fade_len - 1);   // np.linspace(1, 0, fade_len)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < wave_len; i += gridDim.x * blockDim.x) {
    float yf = label_to_float((int)labels[(size_t)b * S + i], (float)(ncls - 1));
    double y = (double)yf;
    if (mu_law) {
      double a = fabs(y);
      double sgn = (y > 0.0) - (y < 0.0);
      y = sgn / mu * (pow(1.0 + mu, a) - 1.0);
    }
    int k = i - (wave_len - fade_len);
    if (k >= 0) y *= (k == fade_len - 1) ? 0.0 : (1.0 + (double)k * step);
    wave[(size_t)b * wave_len + i] = y;
  }
}

}  // namespace b200tts
