// Per-edge, per-head attention scores and their normalisation; device side of
// SpGraphTransAttentionLayer.forward (reference src/function_transformer_attention.py:128-214).
#pragma once
#include "common.cuh"

namespace grand {

// plain-old-data view of grand_attention that kernels take by value
struct AttnView {
  int ad;        // attention_dim
  int ad_eff;    // row stride of q / k (2*ad for EXP_KERNEL2)
  int heads;
  int dk;        // ad / heads
  int score_type;
  int norm_type;
  int norm_idx;
  float sqrt_dk;                  // (float)np.sqrt(d_k): the reference divides by it (:196)
  float inv_sqrt_dk_exact;        // 1/sqrt(d_k) when that is an exact power of two, else 0
  const float* kernel_scalars;    // device [4]; GAT: [2 d_k + 1]
  const float* edge_weight_csr;   // or null
};

// Numerically exact transcendental on purpose (parity bar 1e-5): expf, sqrtf, IEEE division -- no __expf / rsqrt.

// dot products over one head's d_k slice; q, k point at the head's first element
template <bool kVec4>
__device__ __forceinline__ void head_sums(const float* __restrict__ q, const float* __restrict__ k, int dk,
                                          float& sqk, float& sqq, float& skk) {
  sqk = 0.f; sqq = 0.f; skk = 0.f;
  if (kVec4) {
    for (int c = 0; c < dk; c += 4) {
      const float4 a = ldg4(q + c), b = ldg4(k + c);
      sqk = fmaf(a.x, b.x, sqk); sqk = fmaf(a.y, b.y, sqk); sqk = fmaf(a.z, b.z, sqk); sqk = fmaf(a.w, b.w, sqk);
      sqq = fmaf(a.x, a.x, sqq); sqq = fmaf(a.y, a.y, sqq); sqq = fmaf(a.z, a.z, sqq); sqq = fmaf(a.w, a.w, sqq);
      skk = fmaf(b.x, b.x, skk); skk = fmaf(b.y, b.y, skk); skk = fmaf(b.z, b.z, skk); skk = fmaf(b.w, b.w, skk);
    }
  } else {
    for (int c = 0; c < dk; ++c) {
      const float a = __ldg(q + c), b = __ldg(k + c);
      sqk = fmaf(a, b, sqk); sqq = fmaf(a, a, sqq); skk = fmaf(b, b, skk);
    }
  }
}

template <bool kVec4>
__device__ __forceinline__ float head_dot(const float* __restrict__ q, const float* __restrict__ k, int dk) {
  float s = 0.f;
  if (kVec4) {
    for (int c = 0; c < dk; c += 4) {
      const float4 a = ldg4(q + c), b = ldg4(k + c);
      s = fmaf(a.x, b.x, s); s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); s = fmaf(a.w, b.w, s);
    }
  } else {
    for (int c = 0; c < dk; ++c) s = fmaf(__ldg(q + c), __ldg(k + c), s);
  }
  return s;
}

template <bool kVec4>
__device__ __forceinline__ float head_sqdist(const float* __restrict__ q, const float* __restrict__ k, int dk) {
  float s = 0.f;
  if (kVec4) {
    for (int c = 0; c < dk; c += 4) {
      const float4 a = ldg4(q + c), b = ldg4(k + c);
      float t;
      t = a.x - b.x; s = fmaf(t, t, s); t = a.y - b.y; s = fmaf(t, t, s);
      t = a.z - b.z; s = fmaf(t, t, s); t = a.w - b.w; s = fmaf(t, t, s);
    }
  } else {
    for (int c = 0; c < dk; ++c) { const float t = __ldg(q + c) - __ldg(k + c); s = fmaf(t, t, s); }
  }
  return s;
}

__device__ __forceinline__ float cosine_from_sums(float sqk, float sqq, float skk) {
  // torch (>= 2.x) CosineSimilarity(eps): x.y / (max(|x|, eps) * max(|y|, eps)), eps = 1e-5 (:198, :205)
  const float nq = fmaxf(sqrtf(sqq), 1e-5f), nk = fmaxf(sqrtf(skk), 1e-5f);
  return sqk / (nq * nk);
}

// score of head h for the edge (query row q_i, key row k_j); q_i / k_j point at the node's full [ad_eff] row
template <bool kVec4>
__device__ __forceinline__ float edge_score(const AttnView& a, const float* __restrict__ qi,
                                            const float* __restrict__ kj, int h) {
  const int dk = a.dk;
  const float* q = qi + h * dk;
  const float* k = kj + h * dk;
  switch (a.score_type) {
    case GRAND_SCORE_SCALED_DOT:
      return head_dot<kVec4>(q, k, dk) / a.sqrt_dk;
    case GRAND_SCORE_COSINE: {
      float sqk, sqq, skk;
      head_sums<kVec4>(q, k, dk, sqk, sqq, skk);
      return cosine_from_sums(sqk, sqq, skk);
    }
    case GRAND_SCORE_PEARSON: {
      float mq = 0.f, mk = 0.f;
      for (int c = 0; c < dk; ++c) { mq += __ldg(q + c); mk += __ldg(k + c); }
      mq /= (float)dk; mk /= (float)dk;
      float sqk = 0.f, sqq = 0.f, skk = 0.f;
      for (int c = 0; c < dk; ++c) {
        const float u = __ldg(q + c) - mq, v = __ldg(k + c) - mk;
        sqk = fmaf(u, v, sqk); sqq = fmaf(u, u, sqq); skk = fmaf(v, v, skk);
      }
      return cosine_from_sums(sqk, sqq, skk);
    }
    case GRAND_SCORE_EXP_KERNEL: {
      const float ov = __ldg(a.kernel_scalars), ls = __ldg(a.kernel_scalars + 1);
      const float ss = head_sqdist<kVec4>(q, k, dk);
      return (ov * ov) * expf(-(ss / (2.f * (ls * ls))));
    }
    case GRAND_SCORE_EXP_KERNEL2: {
      // [Qx | Qp] layout: the positional half starts at column ad (reference :133-169)
      const float ovx = __ldg(a.kernel_scalars), lsx = __ldg(a.kernel_scalars + 1);
      const float ovp = __ldg(a.kernel_scalars + 2), lsp = __ldg(a.kernel_scalars + 3);
      const float ssx = head_sqdist<kVec4>(q, k, dk);
      const float ssp = head_sqdist<kVec4>(q + a.ad, k + a.ad, dk);
      return (ovx * ovx) * expf(-ssx / (2.f * (lsx * lsx))) * (ovp * ovp) * expf(-ssp / (2.f * (lsp * lsp)));
    }
    case GRAND_SCORE_GAT: {
      // sum(a * [h_i ; h_j]) over the 2 d_k stacked entries, then LeakyReLU (reference function_GAT_attention.py:111-113)
      float s = 0.f;
      for (int c = 0; c < dk; ++c) s = fmaf(__ldg(a.kernel_scalars + c), __ldg(q + c), s);
      for (int c = 0; c < dk; ++c) s = fmaf(__ldg(a.kernel_scalars + dk + c), __ldg(k + c), s);
      const float slope = __ldg(a.kernel_scalars + 2 * dk);
      return s > 0.f ? s : slope * s;
    }
  }
  return 0.f;
}

// squareplus numerator of reference src/utils.py:196-198 given o = s - global_max (always <= 0).
// (o + sqrt(o^2 + 4)) / 2 cancels catastrophically for o << 0: in fp32 the reference's own value carries a relative
// error of up to ~ulp(|o|)/sp(o) (3e-5 at o = -30) that depends on the last bit of o.  The algebraically identical
// 2 / (sqrt(o^2 + 4) - o) is accurate to fp32 rounding, i.e. it sits at the centre of the reference's error band,
// which is the closest any independent implementation can get to it.
__device__ __forceinline__ float squareplus_num(float o) {
  const float r = sqrtf(fmaf(o, o, 4.f));
  return o > 0.f ? 0.5f * (o + r) : 2.f / (r - o);
}

// numerator of the normaliser: exp(s - m_seg) (PyG softmax) or squareplus(s - gmax)
__device__ __forceinline__ float norm_num(int norm_type, float s, float m) {
  return norm_type == GRAND_NORM_SOFTMAX ? expf(s - m) : squareplus_num(s - m);
}

}  // namespace grand
