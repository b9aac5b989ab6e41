// Fine-grained operators of the reverse-diffusion path (one CTA per token row, coalesced scalar
// accesses so that any pitch — including the reference's unaligned K+1 = 4097 — is accepted).
// These back the method-level drop-ins (q_posterior, log_sample_categorical, index<->log-one-hot);
// the production path is the fused step.
#pragma once

#include "d3pm_common.cuh"

namespace d3pm {

constexpr int kOpThreads = 256;

// ---------------------------------------------------------------- coefficient table
__device__ __forceinline__ double lae64(double a, double b) {  // log_add_exp (:32-34), -inf safe
  const double m = fmax(a, b);
  if (isinf(m) && m < 0) return m;
  return m + log(exp(a - m) + exp(b - m));
}

__global__ void coef_table_kernel(const float* __restrict__ sched, int T, int K, float* __restrict__ table) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const int P = T + 1;
  const int tp = (t + T) % P;  // t-1 with the modulo of q_pred (:203): t = 0 -> identity slot T
  const double Z = static_cast<double>(kLogTiny);
  const double la = sched[0 * P + t], lb = sched[1 * P + t], lc = sched[2 * P + t];
  const double lA = sched[4 * P + t], lB = sched[5 * P + t], lC = sched[6 * P + t];
  const double Ap = exp(static_cast<double>(sched[4 * P + tp])), Bp = exp(static_cast<double>(sched[5 * P + tp]));
  const double Cp = exp(static_cast<double>(sched[6 * P + tp])), omCp = exp(static_cast<double>(sched[7 * P + tp]));
  const double Wm = exp(-lC), Om = exp(lc);
  const double Wo = exp(-lae64(Z + lA, lB)), Ws = exp(-lae64(lA, lB));
  const double Oo = exp(lae64(Z + la, lb)), Os = exp(lae64(la, lb));
  float* row = table + static_cast<size_t>(t) * D3PM_COEF_STRIDE;
  for (int i = 0; i < D3PM_COEF_STRIDE; ++i) row[i] = 0.f;
  row[C_AM] = static_cast<float>(Wm * Ap * Om);
  row[C_BOM] = static_cast<float>(Bp * Om);
  row[C_WM] = static_cast<float>(Wm);
  row[C_C1] = static_cast<float>(1e-30 * omCp);
  row[C_CP] = static_cast<float>(Cp);
  row[C_AO] = static_cast<float>(Wo * Ap * Oo);
  row[C_AS] = static_cast<float>(Ws * Ap * Os);
  row[C_BOO] = static_cast<float>(Bp * Oo);
  row[C_BOS] = static_cast<float>(Bp * Os);
  row[C_WO] = static_cast<float>(Wo);
  row[C_WS] = static_cast<float>(Ws);
  row[C_PK1] = static_cast<float>(Cp * 1e-30);
}

// ---------------------------------------------------------------- small block reductions
template <int NW>
__device__ __forceinline__ float block_sum(float x, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  x = warp_sum(x);
  if (lane == 0) scratch[warp] = x;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < NW; ++w) s += scratch[w];
  return s;
}

__device__ __forceinline__ bool fetch_row_scalars(const int64_t* t, const int64_t* x_t, int64_t row, int N, int K,
                                                  int T, uint32_t* status, int& tt_out, uint32_t& j_out) {
  long long tt = t[row / N];
  long long jj = x_t[row];
  uint32_t st = 0;
  if (tt < 0 || tt >= T) {
    st |= D3PM_STATUS_BAD_T;
    tt = tt < 0 ? 0 : T - 1;
  }
  if (jj < 0 || jj > K) {
    st |= D3PM_STATUS_BAD_TOKEN;
    jj = K;
  }
  if (st != 0 && threadIdx.x == 0 && status != nullptr) atomicOr(status, st);
  tt_out = static_cast<int>(tt);
  j_out = static_cast<uint32_t>(jj);
  return jj == K;
}

// ---------------------------------------------------------------- q_posterior on arbitrary log p(x0)
__global__ void __launch_bounds__(kOpThreads) q_posterior_rows_kernel(
    const float* __restrict__ lxs, int64_t pitch_in, const int64_t* __restrict__ x_t, const int64_t* __restrict__ t,
    const float* __restrict__ table, float* __restrict__ post, int64_t pitch_out, int N, int K, int T,
    uint32_t* status) {
  __shared__ float scratch[kOpThreads / 32];
  const int64_t row = blockIdx.x;
  int tt;
  uint32_t j;
  const bool masked = fetch_row_scalars(t, x_t, row, N, K, T, status, tt, j);
  const RowCoef cf = load_row_coef(table, tt, masked);
  const float* __restrict__ in = lxs + row * pitch_in;
  float* __restrict__ out = post + row * pitch_out;

  float sum = 0.f;  // over the classes other than x_t (summing everything and subtracting p_j would cancel when p_j ~ 1)
  for (int k = threadIdx.x; k < K; k += kOpThreads) sum += (static_cast<uint32_t>(k) == j) ? 0.f : ex2(in[k] * kLog2e);
  sum = block_sum<kOpThreads / 32>(sum, scratch);
  const float pj = masked ? 0.f : ex2(in[j] * kLog2e);
  const float eL = masked ? fmaf(cf.W, sum, kTiny) : fmaf(cf.W, sum, fmaf(cf.WS, pj, kTiny));
  const float Bc = cf.BO * eL;
  const float Pj = fmaf(pj, cf.AS, cf.BOS * eL);
  for (int k = threadIdx.x; k < K; k += kOpThreads) {
    const float pk = ex2(in[k] * kLog2e);
    const float P = (static_cast<uint32_t>(k) == j) ? Pj : fmaf(pk, cf.A, Bc);
    out[k] = log_prob_clamped(P);
  }
  if (threadIdx.x == 0) out[K] = log_prob_clamped(fmaf(cf.PK1, eL, cf.PK0));
}

// ---------------------------------------------------------------- Gumbel-max over rows of C classes
template <int KIND>  // 0 gumbel given, 1 uniform given, 2 Philox
__global__ void __launch_bounds__(kOpThreads) gumbel_argmax_rows_kernel(
    const float* __restrict__ logits, int64_t pitch_logits, const float* __restrict__ noise, int64_t pitch_noise,
    int64_t* __restrict__ x, float* __restrict__ gap, int C, uint64_t seed, uint64_t offset, int64_t row_offset) {
  constexpr int NW = kOpThreads / 32;
  __shared__ unsigned long long skey[NW];
  __shared__ float sgap[NW];
  const int64_t row = blockIdx.x;
  const float* __restrict__ lg = logits + row * pitch_logits;
  const float* __restrict__ nz = KIND == 2 ? nullptr : noise + row * pitch_noise;
  const NoiseStream rng(seed, offset);
  const uint64_t grow = static_cast<uint64_t>(row_offset + row);
  unsigned long long best = 0ull;
  float second = -CUDART_INF_F;
  const int nq = (C + 3) >> 2;
  for (int q = threadIdx.x; q < nq; q += kOpThreads) {
    uint4 cw = make_uint4(0, 0, 0, 0), fw = cw;
    if (KIND == 2) cw = rng.coarse(NoiseStream::coarse_call_of_chunk(q), grow), fw = rng.fine(q >> 2, grow);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = 4 * q + e;
      if (k < C) {
        float g;
        if (KIND == 0) g = nz[k];
        else if (KIND == 1) g = gumbel_from_uniform(nz[k]);
        else g = gumbel_from_uniform(uniform_from_draw(
                 (NoiseStream::half_of(cw, (((q >> 7) & 1) << 2) | e) << 7) | NoiseStream::low7_of(fw, ((q & 3) << 2) | e)));
        const float sc = g + lg[k];
        const unsigned long long key = pack_key(sc, k);
        if (key > best) {
          if (best != 0ull) second = fmaxf(second, key_score(best));
          best = key;
        } else {
          second = fmaxf(second, sc);
        }
      }
    }
  }
  const unsigned long long win = group_max_u64<NW>(best, skey, CtaSync());
  if (threadIdx.x == 0) x[row] = key_class(win);
  if (gap != nullptr) {
    float cand = (best == win) ? second : (best != 0ull ? key_score(best) : -CUDART_INF_F);
    cand = group_max_f32<NW>(cand, sgap, CtaSync());
    if (threadIdx.x == 0) gap[row] = key_score(win) - cand;
  }
}

// The same draw (Philox noise) with the race thinned: softmax statistics of the row first, then a class is scored only
// if its 16 coarse noise bits h satisfy h / 2^16 <= c p_k (it cannot reach ln(sum exp) - ln c otherwise); the winner
// is accepted when it does reach that bound (probability 1 - e^-c), else every class is scored as in the kernel above.
// Identical tokens by construction (tested); ~60 instructions per class become ~12.  One CTA per row, a thread per
// coarse Philox call (the classes of float4 chunks q and q + 128).
constexpr float kOpThin = 8.0f;
__device__ __noinline__ unsigned long long gumbel_key_of(float logit, uint32_t k, uint64_t grow, uint64_t seed, uint64_t offset) {
  const NoiseStream rng(seed, offset);
  return pack_key(gumbel_from_uniform(uniform_from_draw(rng.draw(k, grow))) + logit, k);
}
__global__ void __launch_bounds__(kOpThreads) gumbel_argmax_thin_rows_kernel(const float* __restrict__ logits, int64_t pitch_logits,
                                                                             int64_t* __restrict__ x, int C, uint64_t seed,
                                                                             uint64_t offset, int64_t row_offset) {
  constexpr int NW = kOpThreads / 32;
  constexpr int kMaxCalls = 5;  // C <= 8193: 2049 chunks -> 1152 calls over 256 threads
  __shared__ unsigned long long skey[NW];
  __shared__ float sred[2][NW];
  const int64_t row = blockIdx.x;
  const float* __restrict__ lg = logits + row * pitch_logits;
  const NoiseStream rng(seed, offset);
  const uint64_t grow = static_cast<uint64_t>(row_offset + row);
  const int nq = (C + 3) >> 2;
  const int ncalls = ((nq + 255) >> 8) << 7;
  const int tid = threadIdx.x;
  // ---- the row: classes 4 chunk + e of chunks (lo, lo + 128) per call ----
  float v[kMaxCalls][8];
  float m = -CUDART_INF_F;
#pragma unroll
  for (int jc = 0; jc < kMaxCalls; ++jc) {
    const int call = tid + kOpThreads * jc;
    const int lo = ((call >> 7) << 8) | (call & 127);
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      const int k = 4 * (lo + 128 * (h >> 2)) + (h & 3);
      v[jc][h] = (call < ncalls && k < C) ? __ldg(lg + k) : -CUDART_INF_F;
      m = fmaxf(m, v[jc][h]);
    }
  }
  m = warp_max(m);
  if ((tid & 31) == 0) sred[0][tid >> 5] = m;
  __syncthreads();
  float M = sred[0][0];
#pragma unroll
  for (int w = 1; w < NW; ++w) M = fmaxf(M, sred[0][w]);
  unsigned long long best = 0ull;
  bool settled = false;
  if (M > -CUDART_INF_F && M < CUDART_INF_F) {
    const float M2 = __fmul_rn(M, kLog2e);
    float ssum = 0.f;
#pragma unroll
    for (int jc = 0; jc < kMaxCalls; ++jc)
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        v[jc][h] = ex2(fmaf(v[jc][h], kLog2e, -M2));  // softmax numerators (0 beyond C and for -inf entries)
        ssum += v[jc][h];
      }
    ssum = warp_sum(ssum);
    if ((tid & 31) == 0) sred[1][tid >> 5] = ssum;
    __syncthreads();
    float S = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) S += sred[1][w];
    const float scale = kOpThin / S * 65536.0f * 1.001f;  // h <= e_k * scale + 2
    const float accept = M + logf(S / kOpThin) + 0.02f;
#pragma unroll
    for (int jc = 0; jc < kMaxCalls; ++jc) {
      const int call = tid + kOpThreads * jc;
      if (call >= ncalls) continue;
      const uint4 cw = rng.coarse(static_cast<uint32_t>(call), grow);
      const uint32_t w4[4] = {cw.x, cw.y, cw.z, cw.w};
      const int lo = ((call >> 7) << 8) | (call & 127);
      uint32_t hits = 0;
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        const uint32_t half = (h & 1) ? (w4[h >> 1] >> 16) : (w4[h >> 1] & 0xffffu);
        hits |= (static_cast<float>(half) <= fmaf(v[jc][h], scale, 2.0f) ? 1u : 0u) << h;
      }
      while (hits != 0) {
        const int h = __ffs(hits) - 1;
        hits &= hits - 1;
        const int k = 4 * (lo + 128 * (h >> 2)) + (h & 3);
        if (k < C) {
          const unsigned long long key = gumbel_key_of(__ldg(lg + k), static_cast<uint32_t>(k), grow, seed, offset);
          best = key > best ? key : best;
        }
      }
    }
    best = group_max_u64<NW>(best, skey, CtaSync());
    settled = key_score(best) >= accept;
    __syncthreads();  // skey is reused below
  }
  if (!settled) {  // every class (probability ~e^-c, and rows without a finite maximum)
    best = 0ull;
    for (int k = tid; k < C; k += kOpThreads) {
      const unsigned long long key = gumbel_key_of(__ldg(lg + k), static_cast<uint32_t>(k), grow, seed, offset);
      best = key > best ? key : best;
    }
    best = group_max_u64<NW>(best, skey, CtaSync());
  }
  if (tid == 0) x[row] = key_class(best);
}

__global__ void __launch_bounds__(kOpThreads) philox_uniform_kernel(float* __restrict__ u, int K, int64_t pitch,
                                                                    uint64_t seed, uint64_t offset,
                                                                    int64_t row_offset) {
  const int64_t row = blockIdx.x;
  const NoiseStream rng(seed, offset);
  const uint64_t grow = static_cast<uint64_t>(row_offset + row);
  float* __restrict__ out = u + row * pitch;
  const int C = K + 1;
  for (int k = threadIdx.x; k < C; k += kOpThreads) out[k] = uniform_from_draw(rng.draw(k, grow));
}

// ---------------------------------------------------------------- index <-> log one-hot
__global__ void __launch_bounds__(kOpThreads) tokens_to_log_onehot_kernel(const int64_t* __restrict__ x,
                                                                          float* __restrict__ out, int64_t pitch,
                                                                          int C, uint32_t* status) {
  const int64_t row = blockIdx.x;
  const long long j = x[row];
  if ((j < 0 || j >= C) && threadIdx.x == 0 && status != nullptr) atomicOr(status, D3PM_STATUS_BAD_TOKEN);
  float* __restrict__ o = out + row * pitch;
  for (int k = threadIdx.x; k < C; k += kOpThreads) o[k] = (k == j) ? 0.0f : kLogTiny;
}

// class_stride == 1: one CTA per token
__global__ void __launch_bounds__(kOpThreads) argmax_rows_kernel(const float* __restrict__ x, int64_t batch_stride,
                                                                 int64_t token_stride, int64_t* __restrict__ idx,
                                                                 int C, int N) {
  constexpr int NW = kOpThreads / 32;
  __shared__ unsigned long long skey[NW];
  const int64_t row = blockIdx.x;
  const float* __restrict__ r = x + (row / N) * batch_stride + (row % N) * token_stride;
  unsigned long long best = 0ull;
  for (int k = threadIdx.x; k < C; k += kOpThreads) {
    const unsigned long long key = pack_key(r[k], k);
    best = key > best ? key : best;
  }
  best = group_max_u64<NW>(best, skey, CtaSync());
  if (threadIdx.x == 0) idx[row] = key_class(best);
}

// any strides: one thread per token, classes visited in order (coalesced when token_stride == 1)
__global__ void __launch_bounds__(kOpThreads) argmax_strided_kernel(const float* __restrict__ x,
                                                                    int64_t batch_stride, int64_t class_stride,
                                                                    int64_t token_stride, int64_t* __restrict__ idx,
                                                                    int C, int N) {
  const int n = blockIdx.x * kOpThreads + threadIdx.x;
  const int b = blockIdx.y;
  if (n >= N) return;
  const float* __restrict__ r = x + b * batch_stride + n * token_stride;
  float best = r[0];
  int arg = 0;
  for (int k = 1; k < C; ++k) {
    const float v = r[k * class_stride];
    if (v > best) best = v, arg = k;
  }
  idx[static_cast<int64_t>(b) * N + n] = arg;
}

// [B, C, N] contiguous -> [B*N][pitch] rows, 32x32 tiles through shared memory
__global__ void __launch_bounds__(256) to_token_major_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                             int64_t pitch, int C, int N) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float* __restrict__ s = src + static_cast<int64_t>(b) * C * N;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int c = c0 + ty + i, n = n0 + tx;
    if (c < C && n < N) tile[ty + i][tx] = s[static_cast<int64_t>(c) * N + n];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int n = n0 + ty + i, c = c0 + tx;
    if (c < C && n < N) dst[(static_cast<int64_t>(b) * N + n) * pitch + c] = tile[tx][ty + i];
  }
}

// ---------------------------------------------------------------- purity-prior reveal (p_sample :331-343)
// One CTA per video.  key_n = w_n / q_n with w_n the per-video-normalised purity of the [MASK] positions (0 for the
// already revealed ones) and q_n ~ Exp(1): the n_reveal largest keys are what torch.multinomial(w, n_reveal) returns
// (ATen draws exactly this race).  The keys are sorted with a shared-memory bitonic network as 64-bit words
// (key bits, ~position) so that equal keys -- only the zero weights -- resolve to the lowest position.
constexpr int kPurityThreads = 1024;
constexpr int kPurityMaxN = 8192;

__global__ void __launch_bounds__(kPurityThreads) purity_select_kernel(
    const int64_t* __restrict__ x_t, const int64_t* __restrict__ x_cand, const float* __restrict__ score,
    const float* __restrict__ expo, const int32_t* __restrict__ n_reveal, int64_t* __restrict__ x_out,
    int32_t* __restrict__ revealed, int N, int n2, int K, uint64_t seed, uint64_t offset, int64_t row_offset) {
  extern __shared__ unsigned long long skeys[];
  __shared__ float sred[kPurityThreads / 32];
  __shared__ int scount[kPurityThreads / 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int64_t base = static_cast<int64_t>(b) * N;
  const NoiseStream rng(seed, offset);

  // per-video maximum of the raw purity (:319)
  float mx = 0.f;
  if (score != nullptr)
    for (int n = tid; n < N; n += kPurityThreads) mx = fmaxf(mx, score[base + n]);
  mx = warp_max(mx);
  if ((tid & 31) == 0) sred[tid >> 5] = mx;
  __syncthreads();
  mx = sred[0];
  for (int w = 1; w < kPurityThreads / 32; ++w) mx = fmaxf(mx, sred[w]);
  const float denom = mx + 1e-10f;

  for (int n = tid; n < n2; n += kPurityThreads) {
    unsigned long long key = 0ull;  // padding sorts last
    if (n < N) {
      const bool is_mask = (x_t[base + n] == K);
      const float w = is_mask ? (score != nullptr ? score[base + n] / denom : 1.0f) : 0.0f;
      float q;
      if (expo != nullptr) {
        q = expo[base + n];
      } else {  // Exp(1) = -log(u), u from the stream's class-0 draw of this (global) row
        q = -logf(uniform_from_draw(rng.draw(0u, static_cast<uint64_t>(row_offset + base + n))));
      }
      const float r = w / q;  // >= 0, so its bit pattern orders like the value
      key = (static_cast<unsigned long long>(__float_as_uint(r)) << 32) | (0xffffffffu - static_cast<uint32_t>(n));
      key |= 1ull << 63;      // every real position outranks the padding
    }
    skeys[n] = key;
  }
  __syncthreads();
  // bitonic sort, descending
  for (int k = 2; k <= n2; k <<= 1)
    for (int jj = k >> 1; jj > 0; jj >>= 1) {
      for (int i = tid; i < n2; i += kPurityThreads) {
        const int l = i ^ jj;
        if (l > i) {
          const unsigned long long a = skeys[i], c = skeys[l];
          const bool desc = ((i & k) == 0);
          if (desc ? (a < c) : (a > c)) skeys[i] = c, skeys[l] = a;
        }
      }
      __syncthreads();
    }
  int want = n_reveal[b];
  want = want < 0 ? 0 : (want > N ? N : want);
  // x_out = x_t, then the selected positions take the candidate tokens
  int gained = 0;
  for (int n = tid; n < N; n += kPurityThreads) x_out[base + n] = x_t[base + n];
  __syncthreads();
  for (int i = tid; i < want; i += kPurityThreads) {
    const int n = static_cast<int>(0xffffffffu - static_cast<uint32_t>(skeys[i] & 0xffffffffu));
    const int64_t before = x_t[base + n], after = x_cand[base + n];
    x_out[base + n] = after;
    gained += (after != K ? 1 : 0) - (before != K ? 1 : 0);
  }
  for (int o = 16; o > 0; o >>= 1) gained += __shfl_xor_sync(0xffffffffu, gained, o);
  if ((tid & 31) == 0) scount[tid >> 5] = gained;
  __syncthreads();
  if (tid == 0 && revealed != nullptr) {
    int total = 0;
    for (int w = 0; w < kPurityThreads / 32; ++w) total += scount[w];
    revealed[b] = total;
  }
}

// ---------------------------------------------------------------- token -> video, first stage
// VQVAE.decode (videogpt_vq_vae.py:53-56) starts with  h = post_vq_conv(shift_dim(F.embedding(tokens, codebook), -1, 1)):
// a gather of E-dim codebook rows followed by a 1x1x1 convolution E -> C.  Both are per-token and the token takes K
// values, so they collapse into ONE table  lut[k][c] = sum_e codebook[k][e] w[c][e] + b[c]  (built once per weight version)
// and the stage becomes a gather that writes the channels-first [B, C, T*H*W] tensor the decoder's convolutions expect.
__global__ void __launch_bounds__(256) decode_lut_kernel(const float* __restrict__ codebook, const float* __restrict__ w,
                                                         const float* __restrict__ b, int E, int C, float* __restrict__ lut) {
  extern __shared__ float srow[];  // the code's embedding
  const int k = blockIdx.x;
  for (int e = threadIdx.x; e < E; e += blockDim.x) srow[e] = codebook[static_cast<size_t>(k) * E + e];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float* wc = w + static_cast<size_t>(c) * E;
    float acc = 0.f;
    for (int e = 0; e < E; ++e) acc = fmaf(srow[e], __ldg(wc + e), acc);
    lut[static_cast<size_t>(k) * C + c] = acc + (b != nullptr ? b[c] : 0.f);
  }
}

// out[b][c][n] = lut[tokens[b][n]][c].  One CTA per 32 consecutive positions of a video; 32 x 32 tiles are transposed
// through shared memory so that both the table reads (along c) and the stores (along n) are coalesced.
__global__ void __launch_bounds__(256) tokens_to_features_kernel(const int64_t* __restrict__ tokens, const float* __restrict__ lut,
                                                                 float* __restrict__ out, int N, int K, int C,
                                                                 uint32_t* __restrict__ status) {
  __shared__ float tile[32][33];
  __shared__ int stok[32];
  const int b = blockIdx.y, n0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  if (threadIdx.x < 32) {
    const int n = n0 + threadIdx.x;
    long long tk = n < N ? tokens[static_cast<size_t>(b) * N + n] : 0;
    if (tk < 0 || tk >= K) {  // [MASK] (= K) has no codebook row either
      if (status != nullptr) atomicOr(status, D3PM_STATUS_BAD_TOKEN);
      tk = -1;
    }
    stok[threadIdx.x] = static_cast<int>(tk);
  }
  __syncthreads();
  for (int c0 = 0; c0 < C; c0 += 32) {
#pragma unroll
    for (int r = ty; r < 32; r += 8) {  // row r = position n0 + r, column tx = channel c0 + tx
      const int tk = stok[r];
      tile[r][tx] = (tk >= 0 && c0 + tx < C) ? __ldg(lut + static_cast<size_t>(tk) * C + c0 + tx) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {  // row r = channel c0 + r, column tx = position n0 + tx
      if (c0 + r < C && n0 + tx < N) out[(static_cast<size_t>(b) * C + c0 + r) * N + n0 + tx] = tile[tx][r];
    }
    __syncthreads();
  }
}

}  // namespace d3pm
