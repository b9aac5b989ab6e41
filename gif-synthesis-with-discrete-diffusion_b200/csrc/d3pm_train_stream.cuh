// Training-side loss, production kernel: persistent CTAs, TMA-staged logit rows, loss AND gradient in
// one pass.  Same mathematics as train_rows_kernel (d3pm_train_rows.cuh, `_train_loss` diffusion_transformer.py:391-457);
// what changes is the schedule:
//   * one CTA per SM, four independent groups of 128 threads, a group owns every G-th token row;
//   * a two-stage shared-memory ring per group, filled with 1-D bulk TMA (16 KiB per row for K = 4096): two rows per
//     group are in flight while a third is processed from registers;
//   * the row lives in registers (32 classes per thread); every class is treated by the GENERIC formula and the (at
//     most two) special classes x_0 and x_t are corrected on scalars afterwards, so the inner loops carry no per-class
//     compares;
//   * the sum over classes of the gradient's softmax term is obtained in closed form from sums accumulated in the
//     forward sweep, so a row needs two group exchanges (two 128-thread named barriers), not four;
//   * the gradient row is written with streaming 128-bit stores straight from registers.
// HBM traffic is the algorithmic minimum of a fused forward + backward: 16 KiB read and 16 KiB written per token.
#pragma once

#include "d3pm_step_stream.cuh"
#include "d3pm_train_rows.cuh"

namespace d3pm {

__device__ __forceinline__ float rcp_fast(float x) {  // MUFU.RCP, 1 ulp: plenty for a gradient checked to 5e-5
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int NP>
struct __align__(128) TrainGroupSmem {
  float stage[2][1024 * NP];
  alignas(16) float red[2][8 * kGroupWarps];
  unsigned long long keys[2][kGroupWarps];
  unsigned long long full[2];
};

// MODE 0: forward only (per-token losses, arg-maxes); MODE 1: forward + gradient rows
template <int NP, bool WRITE_GRAD>
__global__ void __launch_bounds__(kStreamThreads, 1) train_stream_kernel(const TrainParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int K = 1024 * NP;
  constexpr int NC = 2 * NP;  // float4 chunks per thread
  constexpr uint32_t kRowBytes = K * sizeof(float);
  uint32_t tid;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
  const int g = tid / kGroupThreads, tg = tid % kGroupThreads;
  const int lane = tid & 31, warp = (tid >> 5) & (kGroupWarps - 1);
  TrainGroupSmem<NP>& S = reinterpret_cast<TrainGroupSmem<NP>*>(smem_raw)[g];
  const GroupSync sync{g + 1};
  const long long G = static_cast<long long>(gridDim.x) * kGroupsPerCta;
  const long long first_row = static_cast<long long>(g) * gridDim.x + blockIdx.x;
  const long long rows = p.rows;

  if (tg == 0) {
    mbar_init(&S.full[0], 1);
    mbar_init(&S.full[1], 1);
  }
  sync();
  auto issue_row = [&](long long row, int st) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&S.full[st], kRowBytes);
    tma_load_row(S.stage[st], p.logits + row * p.pitch, kRowBytes, &S.full[st]);
  };
  if (tg == 0) {
    if (first_row < rows) issue_row(first_row, 0);
    if (first_row + G < rows) issue_row(first_row + G, 1);
  }
  uint32_t phase[2] = {0, 0};
  uint32_t status_bits = 0;
  const bool want_arg = (p.x0_recon != nullptr);

  int it = 0;
  for (long long row = first_row; row < rows; row += G, ++it) {
    const int st = it & 1;
    const int b = static_cast<int>(row / p.N);
    long long tt = p.t[b], jj = p.x_t[row], x0l = p.x0[row];
    if (tt < 0 || tt >= p.T) status_bits |= D3PM_STATUS_BAD_T, tt = tt < 0 ? 0 : p.T - 1;
    if (jj < 0 || jj > K) status_bits |= D3PM_STATUS_BAD_TOKEN, jj = K;
    if (x0l < 0 || x0l >= K) status_bits |= D3PM_STATUS_BAD_TOKEN, x0l = 0;
    const bool masked = (jj == K), t0 = (tt == 0);
    const uint32_t j = static_cast<uint32_t>(jj), x0 = static_cast<uint32_t>(x0l);
    const RowCoef cf = load_row_coef(p.coef_table, static_cast<int>(tt), masked);
    const float am = WRITE_GRAD ? __ldg(p.w_main + b) : 0.f, aa = WRITE_GRAD ? __ldg(p.w_aux + b) : 0.f;

    mbar_wait(&S.full[st], phase[st]);
    phase[st] ^= 1u;
    const float* __restrict__ rowbuf = S.stage[st];
    float x[NC][4];
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const float4 a = lds4(rowbuf + 4 * (128 * i + tg));
      x[i][0] = a.x, x[i][1] = a.y, x[i][2] = a.z, x[i][3] = a.w;
    }
    const float c_x0 = rowbuf[x0], c_j = masked ? 0.f : rowbuf[j];

    // ---- exchange 1: (max, sum of exponentials relative to the thread-local max) and the arg-max of the logits ----
    float m = x[0][0], lo = x[0][0];
#pragma unroll
    for (int i = 0; i < NC; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) m = fmaxf(m, x[i][c]), lo = fminf(lo, x[i][c]);
    unsigned long long kbest = 0ull;
    uint32_t idx = 0;
    if (want_arg) {  // first class (lowest index) that attains the thread-local maximum
#pragma unroll
      for (int i = NC - 1; i >= 0; --i)
#pragma unroll
        for (int c = 3; c >= 0; --c) idx = (x[i][c] == m) ? 4u * (128u * i + tg) + c : idx;
      kbest = pack_key(m, idx);
    }
    const float e_top = ex2(fmaf(m, kLog2e, -to_log2_units(fmaxf(m, -3.0e38f))));  // numerator of the thread's best class
    m = fmaxf(m, -3.0e38f);
    const float m2 = to_log2_units(m);
    float e[NC][4];
    float sloc = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        e[i][c] = ex2(fmaf(x[i][c], kLog2e, -m2));
        sloc += e[i][c];
      }
    {
      const float mw = warp_max(m);
      const float mw2 = to_log2_units(mw);
      const float sw = warp_sum(sloc * ex2(m2 - mw2));
      const float lw = warp_min(lo);
      if (want_arg) kbest = warp_max_u64(kbest);
      if (lane == 0) {
        S.red[0][warp] = mw, S.red[0][kGroupWarps + warp] = sw, S.red[0][2 * kGroupWarps + warp] = lw;
        if (want_arg) S.keys[0][warp] = kbest;
      }
    }
    sync();  // everyone has drained the stage and published its partials
    if (tg == 0 && row + 2 * G < rows) issue_row(row + 2 * G, st);
    float M, Ssum, xmin;
    {
      const float4 mw = lds4(S.red[0]), sw = lds4(S.red[0] + kGroupWarps), lw = lds4(S.red[0] + 2 * kGroupWarps);
      xmin = fminf(fminf(lw.x, lw.y), fminf(lw.z, lw.w));
      M = fmaxf(fmaxf(mw.x, mw.y), fmaxf(mw.z, mw.w));
      const float M2g = to_log2_units(M);
      Ssum = fmaf(sw.x, ex2(to_log2_units(mw.x) - M2g),
                  fmaf(sw.y, ex2(to_log2_units(mw.y) - M2g), fmaf(sw.z, ex2(to_log2_units(mw.z) - M2g), sw.w * ex2(to_log2_units(mw.w) - M2g))));
      if (want_arg && tg == 0) {
        unsigned long long kb = S.keys[0][0];
#pragma unroll
        for (int w = 1; w < kGroupWarps; ++w) kb = S.keys[0][w] > kb ? S.keys[0][w] : kb;
        p.x0_recon[row] = key_class(kb);
      }
    }
    const float M2 = to_log2_units(M);
    const float lnS = ln_rel_sum(M, Ssum), rS = rcp_fast(Ssum);
    const float r = ex2(m2 - M2) * rS;  // thread-local numerators -> softmax

    // ---- scalars of the special classes and of the one-hot ("true") posterior (as train_rows_kernel) ----
    const float recon_x0 = fmaxf((c_x0 - M) - lnS, kClampLo);
    const float sm_x0 = ex2(fmaf(c_x0, kLog2e, -M2)) * rS, sm_j = masked ? 0.f : ex2(fmaf(c_j, kLog2e, -M2)) * rS;
    const float p_x0 = fminf(fmaxf(sm_x0, kPFloor), 1.0f);
    const float p_j = masked ? 0.f : fminf(fmaxf(sm_j, kPFloor), 1.0f);
    const float eL = masked ? cf.W + kTiny : fmaf(cf.W, 1.0f - p_j, fmaf(cf.WS, p_j, kTiny));
    const float Bc = cf.BO * eL;
    const bool j_is_x0 = (!masked && j == x0);
    const bool j_other = (!masked && !j_is_x0);
    const float P_j = masked ? 1.0f : fmaf(p_j, cf.AS, cf.BOS * eL);
    const float P_x0 = j_is_x0 ? P_j : fmaf(p_x0, cf.A, Bc);
    const float P_K = fmaf(cf.PK1, eL, cf.PK0);
    const float ptj = j_is_x0 ? 1.0f : kTiny;
    const float eLt = masked ? fmaf(cf.W, 1.0f, kTiny) : fmaf(cf.W, 1.0f - ptj, fmaf(cf.WS, ptj, kTiny));
    const float Tg_log = log_prob_clamped(fmaf(kTiny, cf.A, cf.BO * eLt));
    const float Tj_log = log_prob_clamped(masked ? 1.0f : fmaf(ptj, cf.AS, cf.BOS * eLt));
    const float Tx0_log = j_is_x0 ? Tj_log : log_prob_clamped(fmaf(1.0f, cf.A, cf.BO * eLt));
    const float TK_log = log_prob_clamped(fmaf(cf.PK1, eLt, cf.PK0));
    const float Tg = ex2(Tg_log * kLog2e), Tj = ex2(Tj_log * kLog2e), Tx0 = ex2(Tx0_log * kLog2e), TK = ex2(TK_log * kLog2e);
    const float M_x0 = log_prob_clamped(P_x0), M_j = log_prob_clamped(P_j), M_K = log_prob_clamped(P_K);
    const float wtok = masked ? p.mask_weight_masked : p.mask_weight_unmasked;

    // ---- forward sweep over the classes, all by the generic formula ----
    // sumM = sum M_k; with "inside" = the posterior clamp did not fire and "open" = the recon clamp did not fire:
    // sInv = sum_{inside} 1/P_k,  sPP = sum_{inside & open} p_k / P_k,  sP = sum_{open} p_k
    float sumM = 0.f, sInv = 0.f, sPP = 0.f, sP = 0.f;
    unsigned long long kpost = 0ull;
    float post_best = -CUDART_INF_F;
    uint32_t post_idx = 0;
    // "open" <=> log-softmax_k >= -70 <=> softmax_k >= exp(-70): decided on the softmax value, so the logits themselves
    // are dead after the first sweep (registers: numerators e and reciprocals inv only)
    float inv[NC][4];
    // Clamp-free fast path (row-uniform): when no log-softmax entry can reach -70 (smallest logit of the row) and every
    // generic P_k = p_k A + Bc lies in [exp(-70), 1] (Bc and A + Bc say so), none of the clamps of :236 / :283 can fire for
    // a generic class: the sweep needs no predicates, and the posterior arg-max is the logits' arg-max.
    const bool fast = ((xmin - M) - lnS >= kClampLo + 1.0e-3f) && (Bc >= 1.01f * kPFloor) && (cf.A + Bc <= 0.9999f) &&
                      (cf.A >= 0.f);
    if (fast) {
#pragma unroll
      for (int i = 0; i < NC; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float sm = e[i][c] * r;
          const float Pk = fmaf(sm, cf.A, Bc);
          sumM = fmaf(lg2(Pk), kLn2, sumM);
          if (WRITE_GRAD) {
            const float iv = rcp_fast(Pk);
            inv[i][c] = iv;
            sInv += iv;
            sPP = fmaf(sm, iv, sPP);
            sP += sm;
          }
        }
      post_best = log_prob_clamped(fmaf(fminf(e_top * r, 1.0f), cf.A, Bc));
      post_idx = idx;
    } else {
#pragma unroll
      for (int i = 0; i < NC; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float sm = e[i][c] * r;
          const float pk = fminf(fmaxf(sm, kPFloor), 1.0f);
          const float Pk = fmaf(pk, cf.A, Bc);
          const float lp = lg2(Pk) * kLn2;
          const float Mk = fminf(fmaxf(lp, kClampLo), 0.0f);
          sumM += Mk;
          if (WRITE_GRAD) {
            const bool inside = (lp >= kClampLo) && (lp <= 0.0f), open = sm >= kPFloor;
            const float iv = inside ? rcp_fast(Pk) : 0.f;
            inv[i][c] = iv;
            sInv += iv;
            sPP += open ? pk * iv : 0.f;
            sP += open ? pk : 0.f;
          }
          if (want_arg) {  // strict ">" keeps the lowest class of a tie inside the thread (classes ascend with i, c)
            const bool better = Mk > post_best;
            post_best = better ? Mk : post_best;
            post_idx = better ? 4u * (128u * i + tg) + c : post_idx;
          }
        }
    }
    if (want_arg) kpost = pack_key(post_best, post_idx);
    // the sweep scored x_t with the generic coefficients: the one thread that owns it redoes its 32 classes
    if (want_arg && !masked && tg == static_cast<int>((j >> 2) & 127u)) {
      kpost = 0ull;
#pragma unroll
      for (int i = 0; i < NC; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t k = 4u * (128u * i + tg) + c;
          const float pk = fminf(fmaxf(e[i][c] * r, kPFloor), 1.0f);
          const float Mk = (k == j) ? M_j : log_prob_clamped(fmaf(pk, cf.A, Bc));
          const unsigned long long key = pack_key(Mk, k);
          kpost = key > kpost ? key : kpost;
        }
    }
    // ---- exchange 2 ----
    {
      const float a0 = warp_sum(sumM), a1 = WRITE_GRAD ? warp_sum(sInv) : 0.f;
      const float a2 = WRITE_GRAD ? warp_sum(sPP) : 0.f, a3 = WRITE_GRAD ? warp_sum(sP) : 0.f;
      if (want_arg) kpost = warp_max_u64(kpost);
      if (lane == 0) {
        S.red[1][warp] = a0, S.red[1][kGroupWarps + warp] = a1;
        S.red[1][2 * kGroupWarps + warp] = a2, S.red[1][3 * kGroupWarps + warp] = a3;
        if (want_arg) S.keys[1][warp] = kpost;
      }
    }
    sync();
    {
      const float4 a0 = lds4(S.red[1]), a1 = lds4(S.red[1] + kGroupWarps);
      const float4 a2 = lds4(S.red[1] + 2 * kGroupWarps), a3 = lds4(S.red[1] + 3 * kGroupWarps);
      sumM = (a0.x + a0.y) + (a0.z + a0.w), sInv = (a1.x + a1.y) + (a1.z + a1.w);
      sPP = (a2.x + a2.y) + (a2.z + a2.w), sP = (a3.x + a3.y) + (a3.z + a3.w);
    }
    // remove what the generic formula contributed for the special classes (they are re-added with their own terms)
    auto generic_terms = [&](float sm_k, float pk, float& Mk, float& iv, float& pp, float& po) {
      const float Pk = fmaf(pk, cf.A, Bc);
      const float lp = lg2(Pk) * kLn2;
      Mk = fminf(fmaxf(lp, kClampLo), 0.0f);
      const bool inside = (lp >= kClampLo) && (lp <= 0.0f), open = sm_k >= kPFloor;
      iv = inside ? rcp_fast(Pk) : 0.f;
      pp = open ? pk * iv : 0.f;
      po = open ? pk : 0.f;
    };
    float gM_x0, gi_x0, gpp_x0, gpo_x0, gM_j = 0.f, gi_j = 0.f, gpp_j = 0.f, gpo_j = 0.f;
    generic_terms(sm_x0, p_x0, gM_x0, gi_x0, gpp_x0, gpo_x0);
    if (j_other) generic_terms(sm_j, p_j, gM_j, gi_j, gpp_j, gpo_j);
    const float sumM_gen = sumM - gM_x0 - gM_j;
    const int n_generic = K - 1 - (j_other ? 1 : 0);

    if (tg == 0) {
      float kl = Tg * fmaf(static_cast<float>(n_generic), Tg_log, -sumM_gen);
      kl += Tx0 * (Tx0_log - M_x0) + TK * (TK_log - M_K);
      if (j_other) kl += Tj * (Tj_log - M_j);
      const float nll = -M_x0, aux = -recon_x0;
      if (p.tok_main != nullptr) p.tok_main[row] = t0 ? nll : wtok * kl;
      if (p.tok_aux != nullptr) p.tok_aux[row] = t0 ? nll : wtok * aux;
      if (want_arg && p.xtm1_recon != nullptr) {
        unsigned long long kb = S.keys[1][0];
#pragma unroll
        for (int w = 1; w < kGroupWarps; ++w) kb = S.keys[1][w] > kb ? S.keys[1][w] : kb;
        const unsigned long long kK = pack_key(M_K, K);
        kb = kK > kb ? kK : kb;
        p.xtm1_recon[row] = key_class(kb);
      }
    }
    if (!WRITE_GRAD) continue;

    // ---- gradient (same formulas as train_rows_kernel) ----
    auto inside_of = [](float Pv) {
      const float lp = lg2(Pv) * kLn2;
      return lp >= kClampLo && lp <= 0.0f;
    };
    const float g_gen = t0 ? 0.f : -am * wtok * Tg;
    const float gP_x0 = inside_of(P_x0) ? (t0 ? -(am + aa) : -am * wtok * Tx0) * rcp_fast(P_x0) : 0.f;
    const float gP_j = (j_other && !t0 && inside_of(P_j)) ? (-am * wtok * Tj) * rcp_fast(P_j) : 0.f;
    const float gP_K = (!t0 && inside_of(P_K)) ? (-am * wtok * TK) * rcp_fast(P_K) : 0.f;
    const float sInv_gen = sInv - gi_x0 - gi_j;
    float Gs = g_gen * cf.BO * sInv_gen + gP_K * cf.PK1;
    if (j_is_x0) Gs += gP_x0 * cf.BOS;
    else Gs += gP_x0 * cf.BO + gP_j * cf.BOS;
    // h_k = p_k (gP_k A_k + W_k G) - aa w [k = x0, t > 0], zero where the recon clamp fired
    const bool open_x0 = sm_x0 >= kPFloor, open_j = j_other && (sm_j >= kPFloor);
    float h_x0 = p_x0 * fmaf(gP_x0, j_is_x0 ? cf.AS : cf.A, (j_is_x0 ? cf.WS : cf.W) * Gs);
    if (!t0) h_x0 -= aa * wtok;
    h_x0 = open_x0 ? h_x0 : 0.f;
    const float h_j = open_j ? p_j * fmaf(gP_j, cf.AS, cf.WS * Gs) : 0.f;
    const float gA = g_gen * cf.A, WG = cf.W * Gs;
    const float hsum = fmaf(gA, sPP - gpp_x0 - gpp_j, WG * (sP - gpo_x0 - gpo_j)) + h_x0 + h_j;
    float* __restrict__ rg = p.grad + row * p.pitch_grad;
    const uint32_t q_x0 = x0 >> 2, q_j = j_other ? (j >> 2) : 0xffffffffu;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const uint32_t q = 128u * i + tg;
      float o[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float sm = e[i][c] * r;
        if (fast) {
          o[c] = sm * (fmaf(gA, inv[i][c], WG) - hsum);
        } else {
          const float hk = (sm >= kPFloor) ? fmaf(gA, inv[i][c], WG) : 0.f;  // h_k / p_k, 0 where the recon clamp fired
          o[c] = fminf(sm, 1.0f) * hk - sm * hsum;
        }
      }
      if (q == q_x0) {
        const float v = fmaf(-sm_x0, hsum, h_x0);
#pragma unroll
        for (int c = 0; c < 4; ++c) o[c] = ((x0 & 3u) == static_cast<uint32_t>(c)) ? v : o[c];
      }
      if (q == q_j) {
        const float v = fmaf(-sm_j, hsum, h_j);
#pragma unroll
        for (int c = 0; c < 4; ++c) o[c] = ((j & 3u) == static_cast<uint32_t>(c)) ? v : o[c];
      }
      st_stream4(rg + 4 * q, make_float4(o[0], o[1], o[2], o[3]));
    }
  }
  if (status_bits != 0 && tg == 0 && p.status != nullptr) atomicOr(p.status, status_bits);
}

// grad[row][:] *= factor[video of row]; rows whose factor is exactly 1 are left untouched (no memory traffic), which is
// the common case: the gradient was written with the scale the caller announced (see d3pm_train_rows, backward == 2)
__global__ void __launch_bounds__(256) scale_rows_kernel(float* __restrict__ rows, int64_t pitch, const float* __restrict__ factor,
                                                         int N, int K) {
  const int64_t row = blockIdx.x;
  const float f = __ldg(factor + row / N);
  if (f == 1.0f) return;
  float* __restrict__ r = rows + row * pitch;
  for (int q = threadIdx.x; q < (K >> 2); q += 256) {
    float4 v = *reinterpret_cast<float4*>(r + 4 * q);
    v.x *= f, v.y *= f, v.z *= f, v.w *= f;
    *reinterpret_cast<float4*>(r + 4 * q) = v;
  }
}

inline bool train_stream_supports(const TrainParams& p) {
  return (p.K == 1024 || p.K == 2048 || p.K == 4096) && p.rows >= 2048 && p.pitch % 4 == 0;
}

template <int NP, bool WRITE_GRAD>
int launch_train_stream_t(const TrainParams& p, cudaStream_t s) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return D3PM_ERR_CUDA;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return D3PM_ERR_CUDA;
  const size_t smem = sizeof(TrainGroupSmem<NP>) * kGroupsPerCta;
  auto kern = train_stream_kernel<NP, WRITE_GRAD>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
    return D3PM_ERR_CUDA;
  kern<<<static_cast<unsigned>(sms), kStreamThreads, smem, s>>>(p);
  return D3PM_OK;
}

inline int launch_train_stream(const TrainParams& p, bool write_grad, cudaStream_t s) {
  switch (p.K) {
    case 1024: return write_grad ? launch_train_stream_t<1, true>(p, s) : launch_train_stream_t<1, false>(p, s);
    case 2048: return write_grad ? launch_train_stream_t<2, true>(p, s) : launch_train_stream_t<2, false>(p, s);
    case 4096: return write_grad ? launch_train_stream_t<4, true>(p, s) : launch_train_stream_t<4, false>(p, s);
    default: return D3PM_ERR_UNSUPPORTED;
  }
}

}  // namespace d3pm
