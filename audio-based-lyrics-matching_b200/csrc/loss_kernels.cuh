// Epilogues and small kernels of the batch similarity-matrix contrastive losses:
//   NT-Xent  (lib/losses.py:19-73)   and   CLEWS  (lib/losses.py:210-285).
//
// forward : prep (normalise + fp16 split) -> B x B sweep with a *statistics* epilogue (online
//           log-sum-exp / masked sums per anchor row, never storing S) -> finalize (loss + logdict
//           + per-row coefficients for the backward)
// backward: B x B sweep again (recompute S on the tensor cores) with a *W* epilogue that forms
//           W = dL/dS + (dL/dS)^T from the row and column coefficients and stores it as fp16
//           hi/lo planes (L2-resident at training batch sizes) -> dU = W * U on the same
//           contraction core -> row-wise normalisation Jacobian.
// Closed forms: SURVEY.md section 8(a5),(a6); checked against the reference's autograd.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "gemm_core.cuh"
#include "prep.cuh"

namespace wealy {

constexpr int kLossNtxent = 0;
constexpr int kLossClews = 1;
constexpr int kStatWidth = 8;  // floats per (part, row) partial record

struct LossParams {
  int kind;
  int b;                // (global) batch size = number of columns
  int row0, nb;         // data-parallel sharding: this launch owns anchors [row0, row0 + nb) (0, b on one GPU)
  const int2* lab_idx;  // [b padded to 256] {label, idx} (range-checked from int64)
  float c2;             // NT-Xent: log2(e) / tau
  float g2, b2;         // CLEWS: gamma * log2(e), b * log2(e)
  // statistics sweep
  float* partial;       // [parts][b][kStatWidth]
  int part_base;        // first partial slot of this launch (a sweep split into two launches: local block, then the rest)
  int col_tile_off;     // this launch's column 0 is column 256 * col_tile_off of the batch (local-block launch)
  // W sweep
  const float* rowstat; // [b padded to 256][4]
  __half* w_hi;         // [b][ldw] (ldw = b padded to the k-block)
  __half* w_lo;         // nullable
  long long ldw;
};

// int64 labels / ids -> packed int32 pairs (one 8-byte record per sample: the per-tile column slot is one bulk copy)
__global__ void pack_ids_kernel(const long long* __restrict__ label, const long long* __restrict__ idx, int2* __restrict__ out,
                                int n, int* bad) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const long long a = label[t], b = idx[t];
  if (a > 2147483647ll || a < -2147483648ll || b > 2147483647ll || b < -2147483648ll) atomicAdd(bad, 1);
  out[t] = make_int2((int)a, (int)b);
}

// The same for training-size batches, in ONE CTA, with the reference's single-label noise folded in (lib/losses.py:34-35:
// a batch that holds a single label gets its first max(2, n / 100) labels overwritten with -1, IN PLACE) and the
// workspace's statistics block zeroed -- what used to be five ATen launches and a memset in front of every forward.
__global__ void __launch_bounds__(1024) pack_ids_noise_kernel(long long* __restrict__ label, const long long* __restrict__ idx,
                                                              int2* __restrict__ out, int n, int noise, int* __restrict__ zero,
                                                              int zero_words, int* __restrict__ bad) {
  for (int t = threadIdx.x; t < zero_words; t += blockDim.x) zero[t] = 0;
  int differs = 0;
  if (noise) {
    const long long first = label[0];
    for (int t = threadIdx.x; t < n; t += blockDim.x) differs |= (label[t] != first);
  }
  const int single = noise && !__syncthreads_or(differs);   // (also orders the zeroing before the range-check atomics)
  if (!noise) __syncthreads();
  const int k = max(2, n / 100);
  int nbad = 0;
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    long long a = label[t];
    const long long b = idx[t];
    if (single && t < k) {
      a = -1;
      label[t] = -1;
    }
    nbad += (a > 2147483647ll || a < -2147483648ll || b > 2147483647ll || b < -2147483648ll);
    out[t] = make_int2((int)a, (int)b);
  }
  if (nbad) atomicAdd(bad, nbad);
}

// Per-tile column data of the loss epilogues, bulk-copied into shared memory by the TMA thread (gemm_core.cuh):
// the 256 columns' backward records (float4 rowstat) and {label, idx} pairs -- instead of three global loads per
// matrix element.
struct LossColSlots {
  static constexpr int kColSlots = 3;
  static constexpr int kLvlBytes = kTileN * 16;   // rowstat
  static constexpr int kInfoBytes = kTileN * 8;   // {label, idx}
  static constexpr int kColSlotBytes = kLvlBytes + kInfoBytes;
  static constexpr int kOffColSlots = 0;
  static constexpr int kCtaScratchBytes = kColSlots * kColSlotBytes;
  __device__ static __forceinline__ void col_bulk_src(const LossParams& p, int t, const void*& s0, const void*& s1) {
    s0 = p.rowstat + (size_t)(t + p.col_tile_off) * kTileN * 4;
    s1 = p.lab_idx + (size_t)(t + p.col_tile_off) * kTileN;
  }
  __device__ static __forceinline__ const float4* col_stat(const EpiCtx& c, int col0) {
    return reinterpret_cast<const float4*>(c.col_slot) + (col0 & (kTileN - 1));
  }
  __device__ static __forceinline__ const int2* col_ids(const EpiCtx& c, int col0) {
    return reinterpret_cast<const int2*>(c.col_slot + kLvlBytes) + (col0 & (kTileN - 1));
  }
};

__device__ __forceinline__ float neg_inf() { return __int_as_float(0xff800000); }

// ---------------------------------------------------------------------------------------------
// statistics epilogue
//   NT-Xent partial: {m2 (running max of logits in log2 units), A = sum exp, P = sum_pos exp}
//   CLEWS   partial: {npos, nneg, sum_pos d, sum_neg X, sum_all d, sum_neg d}
// ---------------------------------------------------------------------------------------------
struct LossStatsEpi : LossColSlots {
  using Params = LossParams;
  static constexpr int kWarpScratchBytes = 0;
  struct RowState {
    int lab, id;
    bool valid;
    float a0, a1, a2, a3, a4, a5;
  };

  __device__ static __forceinline__ void row_begin(const Params& p, RowState& st, int row, int, const GemmShape&, const EpiCtx&) {
    st.valid = row < p.nb;
    const int2 li = st.valid ? p.lab_idx[p.row0 + row] : make_int2(0, 0);
    st.lab = li.x;
    st.id = li.y;
    st.a0 = p.kind == kLossNtxent ? neg_inf() : 0.f;
    st.a1 = st.a2 = st.a3 = st.a4 = st.a5 = 0.f;
  }

  __device__ static __forceinline__ void chunk32(const Params& p, RowState& st, int row, int col0,
                                                 const uint32_t (&acc)[32], const GemmShape&, const EpiCtx& ctx) {
    const int gcol0 = col0 + p.col_tile_off * kTileN;  // column index in the (global) batch
    if (!st.valid || gcol0 >= p.b) return;
    const int2* cid = col_ids(ctx, col0);  // broadcast shared-memory reads
    if (p.kind == kLossNtxent) {
      float l[32];
      float cm = neg_inf();
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int j = gcol0 + e;
        const bool ok = (j < p.b) && (j != p.row0 + row);  // diagonal masked by POSITION (losses.py:52-53)
        l[e] = ok ? __uint_as_float(acc[e]) * p.c2 : neg_inf();
        cm = fmaxf(cm, l[e]);
      }
      if (cm > st.a0) {  // online max: rescale the running sums
        const float sc = exp2f(st.a0 - cm);  // a0 = -inf on first use -> 0
        st.a1 *= sc;
        st.a2 *= sc;
        st.a0 = cm;
      }
      if (st.a0 == neg_inf()) return;
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const float ex = exp2f(l[e] - st.a0);  // masked entries (incl. columns >= b): exp2(-inf) = 0
        const int2 cj = cid[e];
        const bool pos = (cj.x == st.lab) && (cj.y != st.id);
        st.a1 += ex;
        st.a2 += pos ? ex : 0.f;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int j = gcol0 + e;
        if (j < p.b) {
          const float s = __uint_as_float(acc[e]);
          const float d = 1.f - s;
          const int2 cj = cid[e];
          const bool same = cj.x == st.lab;
          const bool pos = same && (cj.y != st.id);
          const float x = exp2f(fmaf(-p.g2, d, p.b2));  // exp(b - gamma d)
          st.a0 += pos ? 1.f : 0.f;
          st.a1 += same ? 0.f : 1.f;
          st.a2 += pos ? d : 0.f;
          st.a3 += same ? 0.f : x;
          st.a4 += d;
          st.a5 += same ? 0.f : d;
        }
      }
    }
  }

  __device__ static __forceinline__ void tile_begin(const Params&, RowState&, const GemmShape&, const EpiCtx&, int) {}
  __device__ static __forceinline__ void tile_end(const Params&, RowState&, const GemmShape&, const EpiCtx&) {}
  __device__ static __forceinline__ void row_end(const Params& p, RowState& st, int row, int part, const GemmShape&, const EpiCtx&) {
    if (!st.valid) return;
    float* o = p.partial + ((long long)(p.part_base + part) * p.nb + row) * kStatWidth;
    reinterpret_cast<float4*>(o)[0] = make_float4(st.a0, st.a1, st.a2, st.a3);
    reinterpret_cast<float4*>(o)[1] = make_float4(st.a4, st.a5, 0.f, 0.f);
  }
};

// ---------------------------------------------------------------------------------------------
// W epilogue: W'_ij = scale * (dS_ij + dS_ji), stored as fp16 hi (+ lo) planes.
//   NT-Xent rowstat = {M2, a', b', -}:  dS_ij = E_ij (b'_i - pos a'_i),  E_ij = exp2(l2_ij - M2_i)   (scale = B)
//   CLEWS   rowstat = {ca', cu', -, -}: dS_ij = -pos ca'_i + X_ij neg cu'_i                          (scale = S, see finalize)
// ---------------------------------------------------------------------------------------------
struct LossWEpi : LossColSlots {
  using Params = LossParams;
  static constexpr int kWarpScratchBytes = 0;
  struct RowState {
    int lab, id;
    bool valid;
    float r0, r1, r2;
  };

  __device__ static __forceinline__ void row_begin(const Params& p, RowState& st, int row, int, const GemmShape&, const EpiCtx&) {
    st.valid = row < p.nb;
    const int2 li = st.valid ? p.lab_idx[p.row0 + row] : make_int2(0, 0);
    st.lab = li.x;
    st.id = li.y;
    const float4 r = st.valid ? reinterpret_cast<const float4*>(p.rowstat)[p.row0 + row] : make_float4(0.f, 0.f, 0.f, 0.f);
    st.r0 = r.x;
    st.r1 = r.y;
    st.r2 = r.z;
  }

  __device__ static __forceinline__ void chunk32(const Params& p, RowState& st, int row, int col0,
                                                 const uint32_t (&acc)[32], const GemmShape&, const EpiCtx& ctx) {
    if (!st.valid || col0 >= p.ldw) return;
    const float4* cst = col_stat(ctx, col0);  // broadcast shared-memory reads
    const int2* cid = col_ids(ctx, col0);
    __align__(16) __half hi[32];
    __align__(16) __half lo[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      const int j = col0 + e;
      float w = 0.f;
      if (j < p.b && j != p.row0 + row) {
        const float s = __uint_as_float(acc[e]);
        const float4 cj = cst[e];
        const int2 ci = cid[e];
        const bool same = ci.x == st.lab;
        const bool pos = same && (ci.y != st.id);
        if (p.kind == kLossNtxent) {
          const float l2 = s * p.c2;
          const float eij = exp2f(l2 - st.r0);
          const float eji = exp2f(l2 - cj.x);
          w = eij * (st.r2 - (pos ? st.r1 : 0.f)) + eji * (cj.z - (pos ? cj.y : 0.f));
        } else {
          const float x = exp2f(fmaf(-p.g2, 1.f - s, p.b2));
          w = (pos ? -(st.r0 + cj.x) : 0.f) + (same ? 0.f : x * (st.r1 + cj.y));
        }
      }
      hi[e] = __float2half_rn(w);
      lo[e] = __float2half_rn(w - __half2float(hi[e]));
    }
    // ldw is a multiple of 64 and col0 of 32 -> whole 64-byte segments, 16-byte aligned
    uint4* dh = reinterpret_cast<uint4*>(p.w_hi + (long long)row * p.ldw + col0);
    const uint4* sh = reinterpret_cast<const uint4*>(hi);
#pragma unroll
    for (int v = 0; v < 4; ++v) dh[v] = sh[v];
    if (p.w_lo) {
      uint4* dl = reinterpret_cast<uint4*>(p.w_lo + (long long)row * p.ldw + col0);
      const uint4* sl = reinterpret_cast<const uint4*>(lo);
#pragma unroll
      for (int v = 0; v < 4; ++v) dl[v] = sl[v];
    }
  }

  __device__ static __forceinline__ void tile_begin(const Params&, RowState&, const GemmShape&, const EpiCtx&, int) {}
  __device__ static __forceinline__ void tile_end(const Params&, RowState&, const GemmShape&, const EpiCtx&) {}
  __device__ static __forceinline__ void row_end(const Params&, RowState&, int, int, const GemmShape&, const EpiCtx&) {}
};

// ---------------------------------------------------------------------------------------------
// finalize, two small multi-block kernels:
//   loss_merge_kernel    one thread per anchor row: merge the partial records of the sweep, write the
//                        per-row quantities, block-reduce the batch sums into `acc` (double atomics)
//   loss_finish_kernel   turn the batch sums into the loss / logdict numbers (`out`, indices of
//                        include/wealy_b200.h) and the per-row coefficients of the W sweep (`rowstat`)
// scal[0] = factor the Jacobian kernel applies to dU (1/(B tau) for NT-Xent, 1/S for CLEWS).
// ---------------------------------------------------------------------------------------------
struct LossCfgDev {
  int kind;
  float temperature, gamma, b, eps, epsilon, uw;
  int numerically_friendly;
};

// acc layout (doubles): 0 loss sum | 1 sum align | 2 sum uniform | 3 npos | 4 nneg | 5 anchors with pos
//                       6 sum_pos d | 7 sum_all d | 8 sum_neg d ; acc_max (uint bits of floats): 0 max 1/npos, 1 max cu-term
constexpr int kAccCount = 16;

__device__ __forceinline__ void block_add(double v, double* target) {
  __shared__ double sh[32];
  v = warp_sum_d(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    double t = l < (int)(blockDim.x >> 5) ? sh[l] : 0.0;
    t = warp_sum_d(t);
    if (l == 0 && t != 0.0) atomicAdd(target, t);
  }
}

// b = anchors of this launch (rowstat already points at the first of them), bg = global batch size
__global__ void __launch_bounds__(256) loss_merge_kernel(LossCfgDev cfg, int b, int bg, int parts,
                                                         const float* __restrict__ partial, float* __restrict__ rowstat,
                                                         double* __restrict__ acc, unsigned int* __restrict__ acc_max) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < b;
  if (cfg.kind == kLossNtxent) {
    double term = 0.0;
    if (live) {
      float M = neg_inf();
      for (int p = 0; p < parts; ++p) M = fmaxf(M, partial[((long long)p * b + i) * kStatWidth]);
      float A = 0.f, P = 0.f;
      for (int p = 0; p < parts; ++p) {
        const float4 r = *reinterpret_cast<const float4*>(partial + ((long long)p * b + i) * kStatWidth);
        if (r.x > neg_inf()) {
          const float sc = exp2f(r.x - M);
          A = fmaf(r.y, sc, A);
          P = fmaf(r.z, sc, P);
        }
      }
      const float Ae = A + 1e-8f;  // losses.py:65-66
      const float rho = P / Ae + 1e-8f;
      term = -(double)logf(rho);
      reinterpret_cast<float4*>(rowstat)[i] = make_float4(M, 1.f / (rho * Ae), P / (rho * Ae * Ae), 0.f);
    }
    block_add(term, acc + 0);
  } else {
    double t_align = 0.0, t_uni = 0.0, t_np = 0.0, t_nn = 0.0, t_h = 0.0, t_pd = 0.0, t_ad = 0.0, t_nd = 0.0;
    float m_inv_np = 0.f, m_cu = 0.f;
    if (live) {
      float np = 0.f, nn = 0.f, pd = 0.f, nx = 0.f, ad = 0.f, nd = 0.f;
      for (int p = 0; p < parts; ++p) {
        const float* r = partial + ((long long)p * b + i) * kStatWidth;
        const float4 r0 = *reinterpret_cast<const float4*>(r);
        const float2 r1 = *reinterpret_cast<const float2*>(r + 4);
        np += r0.x; nn += r0.y; pd += r0.z; nx += r0.w; ad += r1.x; nd += r1.y;
      }
      const float align = pd / fmaxf(np, cfg.eps);  // _per_anchor_mean, losses.py:202-208
      const float uni = nx / fmaxf(nn, cfg.eps);
      const float lu = cfg.numerically_friendly ? log1pf(uni) : logf(uni + cfg.epsilon);
      if (np > 0.f) { t_align = align; t_h = 1.0; }
      t_uni = lu; t_np = np; t_nn = nn; t_pd = pd; t_ad = ad; t_nd = nd;
      // true coefficients: dS_ij = -pos ca_i + X_ij neg cu_i, ca_i = 1/(npos_i H) (H known in the finish kernel)
      const float outer = cfg.numerically_friendly ? 1.f / (1.f + uni) : 1.f / (uni + cfg.epsilon);
      const float cu = nn > 0.f ? (float)((double)cfg.uw * (double)cfg.gamma / ((double)bg * (double)nn)) * outer : 0.f;
      // X_ij <= e^b and X_ij <= sum_j X_ij neg_ij = nneg_i uni_i: rows with a vanishing uniformity term
      // (huge 1/(uni+eps) in the non-friendly branch) must not dictate the fp16 scale of W
      m_inv_np = np > 0.f ? 1.f / np : 0.f;
      m_cu = fabsf(cu) * fminf(expf(cfg.b), nn * uni);
      reinterpret_cast<float4*>(rowstat)[i] = make_float4(m_inv_np, cu, 0.f, 0.f);
    }
    block_add(t_align, acc + 1);
    block_add(t_uni, acc + 2);
    block_add(t_np, acc + 3);
    block_add(t_nn, acc + 4);
    block_add(t_h, acc + 5);
    block_add(t_pd, acc + 6);
    block_add(t_ad, acc + 7);
    block_add(t_nd, acc + 8);
    m_inv_np = warp_max(m_inv_np);
    m_cu = warp_max(m_cu);
    if ((threadIdx.x & 31) == 0) {
      atomicMax(acc_max + 0, __float_as_uint(m_inv_np));  // non-negative floats order like uints
      atomicMax(acc_max + 1, __float_as_uint(m_cu));
    }
  }
}

template <typename T>
__device__ __forceinline__ T from_f32(float v);

__device__ __forceinline__ void store_cast(void* dst, int dtype, int k, double v) {
  // dtype codes of include/wealy_b200.h: 0 f32, 1 f16, 2 bf16
  if (dtype == 0) reinterpret_cast<float*>(dst)[k] = (float)v;
  else if (dtype == 1) reinterpret_cast<__half*>(dst)[k] = __float2half_rn((float)v);
  else reinterpret_cast<__nv_bfloat16*>(dst)[k] = __float2bfloat16_rn((float)v);
}

// `out` receives all WEALY_OUT_COUNT doubles (no memset needed in front); out_cast (nullable) the same numbers in the
// batch's own element type, which is what the reference's modules return (lib/losses.py:65-72)
__global__ void __launch_bounds__(256) loss_finish_kernel(LossCfgDev cfg, int b, int d, const double* __restrict__ acc,
                                                          const unsigned int* __restrict__ acc_max,
                                                          const ZStats* __restrict__ zs, float* __restrict__ rowstat,
                                                          float* __restrict__ scal, double* __restrict__ out,
                                                          const int* __restrict__ bad, void* __restrict__ out_cast,
                                                          int cast_dtype) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const double B = (double)b;
  const bool writer = (i == 0);
  double o[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) o[k] = 0.0;
  if (cfg.kind == kLossNtxent) {
    if (writer) {
      o[0] = acc[0] / B;
      scal[0] = 1.f / ((float)b * cfg.temperature);
    }
  } else {
    const double H = acc[5] > 0.0 ? acc[5] : 1.0;
    // W is stored in fp16: scale it by S = 1 / max_i max(ca_i, cu_i max_j X_ij) so that |W'| <= 2 whatever the
    // batch looks like; the Jacobian kernel multiplies by 1/S
    const float bound = fmaxf((float)((double)__uint_as_float(acc_max[0]) / H), __uint_as_float(acc_max[1]));
    const float S = bound > 0.f ? 1.f / bound : 1.f;
    if (i < b) {
      float4 r = reinterpret_cast<const float4*>(rowstat)[i];
      r.x = (float)((double)r.x / H) * S;  // ca' = S / (npos_i H)
      r.y = r.y * S;                       // cu'
      reinterpret_cast<float4*>(rowstat)[i] = r;
    }
    if (writer) {
      const double l_align = acc[5] > 0.0 ? acc[1] / acc[5] : 0.0;  // losses.py:239
      const double l_uni = acc[2] / B;
      const double n2 = B * B;
      o[0] = l_align + (double)cfg.uw * l_uni;
      o[4] = l_align;
      o[5] = l_uni;
      o[6] = acc[3];
      o[7] = acc[4];
      o[8] = acc[5] / B;
      // tops.mmean(d, mask=pos_mask) averages over the COMPLEMENT of the mask (losses.py:267-268)
      o[9] = acc[3] > 0.0 ? (acc[7] - acc[6]) / fmax(n2 - acc[3], 1e-7) : 0.0;
      o[10] = acc[4] > 0.0 ? (acc[7] - acc[8]) / fmax(n2 - acc[4], 1e-7) : 0.0;
      scal[0] = 1.f / S;
    }
  }
  if (writer) {
    const double n = B * (double)d;
    double zsum = 0.0, zsq = 0.0;
    unsigned int zmax = 0u;
    for (int k = 0; k < kZSlots; ++k) {
      zsum += zs->sum[k];
      zsq += zs->sumsq[k];
      zmax = max(zmax, zs->maxabs_bits[k]);
    }
    o[1] = (double)__uint_as_float(zmax);
    o[2] = zsum / n;
    const double var = n > 1.0 ? (zsq - zsum * zsum / n) / (n - 1.0) : 0.0;
    o[3] = sqrt(var > 0.0 ? var : 0.0);
    // labels / ids that do not fit in 32 bits were truncated by pack_ids_kernel: positives would be wrong, so the
    // loss is poisoned (NaN) and the count reported -- no host sync on the training path
    o[11] = (double)bad[0];
    if (bad[0] != 0) o[0] = __longlong_as_double(0x7ff8000000000000ll);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      out[k] = o[k];
      if (out_cast) store_cast(out_cast, cast_dtype, k, o[k]);
    }
  }
}

// Data-parallel exchange 2 as ONE all-gather: every rank packs {batch sums, maxima, its anchors' records} into one
// record, the records of all ranks are gathered, and every rank folds them back into its workspace.
//   record layout (bytes): [0, 128) acc (16 doubles) | [128, 136) acc_max (2 x uint32) | [256, 256 + nb * 16) rowstat rows
constexpr int kDpRecordHeader = 256;
__global__ void loss_dp_pack_kernel(const double* __restrict__ acc, const unsigned int* __restrict__ acc_max,
                                    const float* __restrict__ rowstat_local, int nb, unsigned char* __restrict__ rec) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < kAccCount) reinterpret_cast<double*>(rec)[t] = acc[t];
  if (t < 2) reinterpret_cast<unsigned int*>(rec + 128)[t] = acc_max[t];
  if (t < nb) reinterpret_cast<float4*>(rec + kDpRecordHeader)[t] = reinterpret_cast<const float4*>(rowstat_local)[t];
}
__global__ void loss_dp_unpack_kernel(const unsigned char* __restrict__ recs, long long rec_bytes, int world, int nb,
                                      double* __restrict__ acc, unsigned int* __restrict__ acc_max,
                                      float* __restrict__ rowstat) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < kAccCount) {
    double v = 0.0;
    for (int r = 0; r < world; ++r) v += reinterpret_cast<const double*>(recs + r * rec_bytes)[t];
    acc[t] = v;
  }
  if (t < 2) {
    unsigned int v = 0u;
    for (int r = 0; r < world; ++r) v = max(v, reinterpret_cast<const unsigned int*>(recs + r * rec_bytes + 128)[t]);
    acc_max[t] = v;
  }
  if (t < world * nb) {
    const int r = t / nb, i = t - r * nb;
    reinterpret_cast<float4*>(rowstat)[t] = reinterpret_cast<const float4*>(recs + r * rec_bytes + kDpRecordHeader)[i];
  }
}

// [rows][ld_in] fp16 plane -> [cols][ld_out] (the K-major "B" operand of dU = W * U); columns
// rows <= r < ld_out are written as zeros (k padding).  32 x 32 shared-memory tiles, coalesced both ways.
__global__ void __launch_bounds__(256) transpose_plane_kernel(const __half* __restrict__ in0,
                                                              const __half* __restrict__ in1, int rows, int cols,
                                                              long long ld_in, __half* __restrict__ out0,
                                                              __half* __restrict__ out1, long long ld_out) {
  __shared__ __half tile[32][34];
  const __half* in = blockIdx.z == 0 ? in0 : in1;
  __half* out = blockIdx.z == 0 ? out0 : out1;
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int r = r0 + ty + k, c = c0 + tx;
    tile[ty + k][tx] = (r < rows && c < cols) ? in[(long long)r * ld_in + c] : __float2half_rn(0.f);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int c = c0 + ty + k, r = r0 + tx;
    if (c < cols && r < ld_out) out[(long long)c * ld_out + r] = tile[tx][ty + k];
  }
}

// ---------------------------------------------------------------------------------------------
// normalisation Jacobian (one warp per row):  u = z / div,  proj = u . dU
//   NT-Xent: dz = g * scal * (dU - u proj (r + 1e-6) / r) / (r + 1e-6)      (x/(|x|+eps), tensor_ops.py:169)
//   CLEWS  : dz = g * scal * (dU - u proj) / max(r, 1e-12)                  (F.normalize, losses.py:231)
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T>
__device__ __forceinline__ void store4(T* p, float4 v);
template <>
__device__ __forceinline__ void store4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <>
__device__ __forceinline__ void store4<__half>(__half* p, float4 v) {
  const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  uint2 o;
  o.x = *reinterpret_cast<const unsigned*>(&a);
  o.y = *reinterpret_cast<const unsigned*>(&b);
  *reinterpret_cast<uint2*>(p) = o;
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 o;
  o.x = *reinterpret_cast<const unsigned*>(&a);
  o.y = *reinterpret_cast<const unsigned*>(&b);
  *reinterpret_cast<uint2*>(p) = o;
}

template <typename T>
__global__ void __launch_bounds__(256) loss_jacobian_kernel(int kind, float eps, const T* __restrict__ z, long long ldz, int b, int d,
                                                            const float* __restrict__ norm,
                                                            const float* __restrict__ du, const float* __restrict__ scal,
                                                            const void* __restrict__ grad_out, T* __restrict__ dz,
                                                            long long ld_dz, int grad_dtype = 0) {
  const int row = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
  const int lane = (int)(threadIdx.x & 31);
  if (row >= b) return;
  const float r = norm[row];
  const float div = kind == kLossNtxent ? r + eps : fmaxf(r, eps);  // x / (|x| + eps)  or  x / max(|x|, eps)
  const float inv = 1.f / div;
  const T* zr = z + (long long)row * ldz;
  const float* g = du + (long long)row * d;
  T* o = dz + (long long)row * ld_dz;
  // HBM-bound row pass: 4 elements per lane and access when the rows are suitably aligned (the usual case)
  const bool vec = (d % 4 == 0) && (ldz % 4 == 0) && (ld_dz % 4 == 0) && ((reinterpret_cast<uintptr_t>(z) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(dz) & 15) == 0);
  float proj = 0.f;
  if (vec) {
    for (int k = lane * 4; k < d; k += 128) {
      const float4 zv = load4<T>(zr + k), gv = *reinterpret_cast<const float4*>(g + k);
      proj = fmaf(zv.x * inv, gv.x, proj);
      proj = fmaf(zv.y * inv, gv.y, proj);
      proj = fmaf(zv.z * inv, gv.z, proj);
      proj = fmaf(zv.w * inv, gv.w, proj);
    }
  } else {
    for (int k = lane; k < d; k += 32) proj = fmaf(to_f32<T>(zr[k]) * inv, g[k], proj);
  }
  proj = warp_sum(proj);
  // d(z/(r+eps))/dz has the extra (r+eps)/r on the radial term; r == 0 -> torch's norm subgradient is 0
  const float radial = kind == kLossNtxent ? (r > 0.f ? proj * div / r : 0.f) : proj;
  // upstream gradient of the scalar loss, in whatever type autograd hands it over (the loss carries z's dtype)
  float up = 1.f;
  if (grad_out) {
    up = grad_dtype == 0 ? reinterpret_cast<const float*>(grad_out)[0]
         : grad_dtype == 1 ? __half2float(reinterpret_cast<const __half*>(grad_out)[0])
                           : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(grad_out)[0]);
  }
  const float f = scal[0] * up * inv;
  if (vec) {
    for (int k = lane * 4; k < d; k += 128) {
      const float4 zv = load4<T>(zr + k), gv = *reinterpret_cast<const float4*>(g + k);
      store4<T>(o + k, make_float4(f * (gv.x - zv.x * inv * radial), f * (gv.y - zv.y * inv * radial),
                                   f * (gv.z - zv.z * inv * radial), f * (gv.w - zv.w * inv * radial)));
    }
  } else {
    for (int k = lane; k < d; k += 32) o[k] = from_f32<T>(f * (g[k] - to_f32<T>(zr[k]) * inv * radial));
  }
}

}  // namespace wealy
