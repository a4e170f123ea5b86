// SURVEY.md section 8(f) "next" rows, the steps either side of the hot path:
//   f3  pooled-embedding producer -- MeanPool (lib/layers.py:6-30: masked temporal mean (B,C,T) -> (B,C)) and the
//       avg-pool collate (lib/embedding_dataset/collate_functions.py:131-172: per-track emb.mean(dim=0) over a
//       ragged number of frames), producing the fp32 [N, E] matrix the evaluator consumes, on the device;
//   f4  triplet mining (lib/losses.py:140-171): first positive / first negative of every anchor, replacing the
//       reference's Python loop over the batch (B tensor ops + .item() syncs per step).
// All HBM-bound, one pass over their input.
#pragma once
#include "loss_kernels.cuh"
#include "prep.cuh"

namespace wealy {

// MeanPool forward.  x [B, C, T] (T contiguous), mask [B, T] bytes (non-zero = VALID, may be null), out [B, C].
//   masked  : sum_t x mask / (sum_t mask + 1e-8)          (lib/layers.py:21-25)
//   no mask : mean over T                                 (lib/layers.py:27-28)
// One warp per (b, c) row.
template <typename T>
__global__ void __launch_bounds__(256) mean_pool_fwd_kernel(const T* __restrict__ x, const unsigned char* __restrict__ mask,
                                                            long long rows, int channels, int frames,
                                                            T* __restrict__ out) {
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = (int)(threadIdx.x & 31);
  if (row >= rows) return;
  const long long b = row / channels;
  const T* xr = x + row * frames;
  const unsigned char* mr = mask ? mask + b * frames : nullptr;
  float acc = 0.f, cnt = 0.f;
  for (int t = lane; t < frames; t += 32) {
    const float m = mr ? (mr[t] ? 1.f : 0.f) : 1.f;
    acc = fmaf(to_f32<T>(xr[t]), m, acc);
    cnt += m;
  }
  acc = warp_sum(acc);
  cnt = warp_sum(cnt);
  if (lane == 0) out[row] = from_f32<T>(mr ? acc / (cnt + 1e-8f) : acc / (float)frames);
}

// MeanPool backward: dx[b, c, t] = g[b, c] * mask[b, t] / (sum_t mask + 1e-8)   (or g / T without mask)
template <typename T>
__global__ void __launch_bounds__(256) mean_pool_bwd_kernel(const T* __restrict__ g, const unsigned char* __restrict__ mask,
                                                            long long rows, int channels, int frames,
                                                            T* __restrict__ dx) {
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = (int)(threadIdx.x & 31);
  if (row >= rows) return;
  const long long b = row / channels;
  const unsigned char* mr = mask ? mask + b * frames : nullptr;
  float cnt = 0.f;
  if (mr) {
    for (int t = lane; t < frames; t += 32) cnt += mr[t] ? 1.f : 0.f;
    cnt = warp_sum(cnt);
  }
  const float scale = to_f32<T>(g[row]) / (mr ? (cnt + 1e-8f) : (float)frames);
  T* dr = dx + row * frames;
  for (int t = lane; t < frames; t += 32) dr[t] = from_f32<T>((!mr || mr[t]) ? scale : 0.f);
}

// Avg-pool collate: ragged per-track embeddings concatenated along frames, x [sum_T, E] (E contiguous), track k owns
// rows [offsets[k], offsets[k+1]) -> out [K, E] fp32 = mean over its frames (a track with one frame is copied, as
// the SBERT branch of the collate does; an empty track yields zeros, its "missing embedding" branch).
// One block per track, threads stride over E (coalesced rows).
template <typename T>
__global__ void __launch_bounds__(256) segment_mean_kernel(const T* __restrict__ x, const long long* __restrict__ offsets,
                                                           int tracks, int dim, float* __restrict__ out) {
  const int k = blockIdx.x;
  if (k >= tracks) return;
  const long long r0 = offsets[k], r1 = offsets[k + 1];
  const float inv = r1 > r0 ? 1.f / (float)(r1 - r0) : 0.f;
  for (int e = threadIdx.x; e < dim; e += blockDim.x) {
    float acc = 0.f;
    for (long long r = r0; r < r1; ++r) acc += to_f32<T>(x[r * dim + e]);
    out[(long long)k * dim + e] = acc * inv;
  }
}

// Triplet mining: for anchor i, pos[i] = smallest j with label_j == label_i and idx_j != idx_i, neg[i] = smallest j with
// label_j != label_i (-1 when there is none); lib/losses.py:149-163 takes exactly these ("the first available").
// One warp per anchor: 32 candidates per step, ballot + find-first-set, early exit once both are found.
__global__ void __launch_bounds__(256) triplet_mine_kernel(const long long* __restrict__ label,
                                                           const long long* __restrict__ idx, int b,
                                                           long long* __restrict__ pos, long long* __restrict__ neg) {
  const int i = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
  const int lane = (int)(threadIdx.x & 31);
  if (i >= b) return;
  const long long li = label[i], ii = idx[i];
  int fp = -1, fn = -1;
  for (int j0 = 0; j0 < b && (fp < 0 || fn < 0); j0 += 32) {
    const int j = j0 + lane;
    const bool ok = j < b;
    const long long lj = ok ? label[j] : 0, ij = ok ? idx[j] : 0;
    const unsigned mp = __ballot_sync(0xffffffffu, ok && lj == li && ij != ii);
    const unsigned mn = __ballot_sync(0xffffffffu, ok && lj != li);
    if (fp < 0 && mp) fp = j0 + __ffs(mp) - 1;
    if (fn < 0 && mn) fn = j0 + __ffs(mn) - 1;
  }
  if (lane == 0) {
    pos[i] = fp;
    neg[i] = fn;
  }
}

// ---- f4: triplet margin loss on the mined triplets (lib/losses.py:76-137 -> torch.nn.TripletMarginLoss) ----
// d(x, y) = || x - y + eps ||_p (torch.pairwise_distance), l_i = max(margin + d(a,p) - d_neg, 0) with
// d_neg = d(a,n), or min(d(a,n), d(p,n)) under `swap`.  One warp per anchor; anchors without a positive or a
// negative are skipped exactly as lib/losses.py:156-157 does.
__device__ __forceinline__ float pnorm_term(float v, float p) {
  const float a = fabsf(v);
  return p == 2.f ? a * a : (p == 1.f ? a : __powf(a, p));
}
__device__ __forceinline__ float pnorm_root(float s, float p) { return p == 2.f ? sqrtf(s) : (p == 1.f ? s : powf(s, 1.f / p)); }
// d/dv of (sum |v|^p)^(1/p) given the distance dist
__device__ __forceinline__ float pnorm_grad(float v, float dist, float p) {
  if (dist <= 0.f) return 0.f;
  if (p == 2.f) return v / dist;
  const float sg = v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f);
  if (p == 1.f) return sg;
  return sg * powf(fabsf(v), p - 1.f) / powf(dist, p - 1.f);
}

struct TripletRow {  // what backward needs per anchor
  float d_ap, d_an, d_pn, loss;
};

template <typename T>
__global__ void __launch_bounds__(256) triplet_fwd_kernel(const T* __restrict__ z, long long ldz, int b, int d,
                                                          const long long* __restrict__ pos, const long long* __restrict__ neg,
                                                          float margin, float p, float eps, int swap,
                                                          TripletRow* __restrict__ rows, double* __restrict__ acc) {
  const int i = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
  const int lane = (int)(threadIdx.x & 31);
  if (i >= b) return;
  const long long jp = pos[i], jn = neg[i];
  TripletRow r = {0.f, 0.f, 0.f, -1.f};  // loss < 0 marks "no triplet"
  if (jp >= 0 && jn >= 0) {
    const T *za = z + (long long)i * ldz, *zp = z + jp * ldz, *zn = z + jn * ldz;
    float sap = 0.f, san = 0.f, spn = 0.f;
    for (int k = lane; k < d; k += 32) {
      const float a = to_f32<T>(za[k]), pp = to_f32<T>(zp[k]), nn = to_f32<T>(zn[k]);
      sap += pnorm_term(a - pp + eps, p);
      san += pnorm_term(a - nn + eps, p);
      spn += pnorm_term(pp - nn + eps, p);
    }
    r.d_ap = pnorm_root(warp_sum(sap), p);
    r.d_an = pnorm_root(warp_sum(san), p);
    r.d_pn = pnorm_root(warp_sum(spn), p);
    const float dneg = swap ? fminf(r.d_an, r.d_pn) : r.d_an;
    r.loss = fmaxf(margin + r.d_ap - dneg, 0.f);
    if (lane == 0) {
      atomicAdd(&acc[0], (double)r.loss);
      atomicAdd(&acc[1], 1.0);
    }
  }
  if (lane == 0) rows[i] = r;
}

// dz (fp32, zero-initialised) += g_i * d l_i / d z ; g_i = upstream / n_triplets ('mean') or upstream ('sum'),
// or upstream[i] (per-anchor weights, reduction 'none').
template <typename T>
__global__ void __launch_bounds__(256) triplet_bwd_kernel(const T* __restrict__ z, long long ldz, int b, int d,
                                                          const long long* __restrict__ pos, const long long* __restrict__ neg,
                                                          float p, float eps, int swap, const TripletRow* __restrict__ rows,
                                                          const float* __restrict__ upstream, int per_anchor, int mean,
                                                          const double* __restrict__ acc, float* __restrict__ dz) {
  const int i = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
  const int lane = (int)(threadIdx.x & 31);
  if (i >= b) return;
  const TripletRow r = rows[i];
  if (!(r.loss > 0.f)) return;  // inactive hinge or no triplet
  float g = per_anchor ? upstream[i] : upstream[0];
  if (mean) g /= (float)acc[1];
  const long long jp = pos[i], jn = neg[i];
  const T *za = z + (long long)i * ldz, *zp = z + jp * ldz, *zn = z + jn * ldz;
  float *ga = dz + (long long)i * d, *gp = dz + jp * d, *gn = dz + jn * d;
  const bool use_pn = swap && r.d_pn < r.d_an;
  for (int k = lane; k < d; k += 32) {
    const float a = to_f32<T>(za[k]), pp = to_f32<T>(zp[k]), nn = to_f32<T>(zn[k]);
    const float t_ap = g * pnorm_grad(a - pp + eps, r.d_ap, p);
    float da = t_ap, dp = -t_ap, dn = 0.f;
    if (use_pn) {
      const float t = g * pnorm_grad(pp - nn + eps, r.d_pn, p);
      dp -= t;
      dn += t;
    } else {
      const float t = g * pnorm_grad(a - nn + eps, r.d_an, p);
      da -= t;
      dn += t;
    }
    atomicAdd(&ga[k], da);
    atomicAdd(&gp[k], dp);
    atomicAdd(&gn[k], dn);
  }
}

}  // namespace wealy
