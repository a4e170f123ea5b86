// a4: distance_tensor_redux (lib/tensor_ops.py:288-373) as ONE kernel.
//
// dist (b1, b2, s1, s2) -> (b1, b2): every (query track, candidate track) pair owns an s1 x s2 block of chunk-level
// distances (s <= 32: WEALY yields a handful of chunks per track) and reduces it with one of the reference's
// strategies.  One warp per pair; the block (and its mask) is staged in shared memory, row / column aggregates are
// computed one row or column per lane.  The reference composes every strategy from full-tensor passes (masked min /
// mean, topk, a Python loop of n rounds for the greedy pairing); here the 4-D tensor is read once.
//
// Semantics follow the reference to the letter, quirks included (see oracle/masked.py, pinned by tests/golden/redux.npz):
//   * mask != 0 means EXCLUDED; excluded entries count as +-inf_ = +-1e12 in min / max, as 0 * x in sums;
//   * means divide by max(count, eps);  "meanmin" weights a row's minimum by the row's number of included entries
//     (the reference broadcasts the row minima against the full mask);  a fully excluded row drops out of "minmean";
//   * "best-k": mean of those of the k smallest values that are below inf_;  "worst-k": the reference masks out
//     everything >= -inf_, i.e. everything -> 0 (NaN / inf propagate through 0 * x);
//   * "bpwr-n": greedy best pairs without replacement on the (already jittered) block, transposed first when
//     s2 < s1; n rounds, after each of the first n - 1 every row and column whose minimum is <= the round's minimum
//     is excluded; the result is the mean of the picked entries;
//   * the "s" prefix averages the strategy on the block and on its transpose.
#pragma once
#include "prep.cuh"
#include "loss_kernels.cuh"

namespace wealy {

enum ReduxOp : int { kRdxMin = 0, kRdxMax = 1, kRdxMean = 2, kRdxMinMean = 3, kRdxMeanMin = 4, kRdxBest = 5, kRdxWorst = 6, kRdxBpwr = 7 };

constexpr int kReduxMaxS = 32;  // chunks per track on either side

// One strategy on the block held in shared memory: x[i * ld_i + j * ld_j] (i < n1 rows, j < n2 columns), mk likewise.
template <typename A>
__device__ A redux_block(int op, int karg, const A* x, unsigned char* mk, int n1, int n2, int ld_i, int ld_j, A eps, A inf_, int lane,
                         A* agg /*[32]*/, A* work /*[1024]*/) {
  constexpr unsigned kFull = 0xffffffffu;
  const int n = n1 * n2;
  auto at = [&](int i, int j) { return x[i * ld_i + j * ld_j]; };
  auto ex = [&](int i, int j) { return mk[i * ld_i + j * ld_j] != 0; };
  auto wsum = [&](A v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
  };
  auto wmin = [&](A v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const A w = __shfl_xor_sync(kFull, v, o);
      v = (w < v || w != w) ? w : v;  // NaN wins, like torch.min
    }
    return v;
  };
  auto wmax = [&](A v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const A w = __shfl_xor_sync(kFull, v, o);
      v = (w > v || w != w) ? w : v;
    }
    return v;
  };
  if (op == kRdxMin || op == kRdxMax) {
    A v = op == kRdxMin ? inf_ : -inf_;
    bool any = false;
    for (int e = lane; e < n; e += 32) {
      const int i = e / n2, j = e - i * n2;
      const A w = ex(i, j) ? (op == kRdxMin ? inf_ : -inf_) : at(i, j);
      if (w != w) { v = w; any = true; }
      else if (!any) v = op == kRdxMin ? (w < v ? w : v) : (w > v ? w : v);
    }
    return op == kRdxMin ? wmin(v) : wmax(v);
  }
  if (op == kRdxMean) {
    A s = 0, c = 0;
    for (int e = lane; e < n; e += 32) {
      const int i = e / n2, j = e - i * n2;
      const bool m = ex(i, j);
      s += m ? (A)0 * at(i, j) : at(i, j);
      c += m ? (A)0 : (A)1;
    }
    s = wsum(s);
    c = wsum(c);
    return s / (c > eps ? c : eps);
  }
  if (op == kRdxMinMean || op == kRdxMeanMin) {
    // one row per lane: its masked mean / masked minimum and its number of included entries
    A r = 0, cnt = 0;
    if (lane < n1) {
      if (op == kRdxMinMean) {
        A s = 0;
        for (int j = 0; j < n2; ++j) {
          const bool m = ex(lane, j);
          s += m ? (A)0 * at(lane, j) : at(lane, j);
          cnt += m ? (A)0 : (A)1;
        }
        r = s / (cnt > eps ? cnt : eps);
      } else {
        r = inf_;
        bool nan = false;
        for (int j = 0; j < n2; ++j) {
          const bool m = ex(lane, j);
          const A w = m ? inf_ : at(lane, j);
          if (w != w) { r = w; nan = true; }
          else if (!nan && w < r) r = w;
          cnt += m ? (A)0 : (A)1;
        }
      }
    }
    if (op == kRdxMinMean) {
      // min over the rows that have an included entry (their mean stands at every included position)
      A v = (lane < n1 && cnt > 0) ? r : inf_;
      return wmin(v);
    }
    // mean of the row minima, every row weighted by its number of included entries
    A s = lane < n1 ? (cnt > 0 ? cnt * r : (A)0 * r) : (A)0;
    if (lane < n1 && cnt > 0 && cnt != (A)n2) {
      // (the reference adds included * r entry by entry: same value, kept as a product)
    }
    s = wsum(s);
    const A c = wsum(lane < n1 ? cnt : (A)0);
    return s / (c > eps ? c : eps);
  }
  if (op == kRdxBest || op == kRdxWorst) {
    // k smallest (largest) of the flattened block with excluded entries at +inf_ (-inf_): rank counting, n <= 1024
    const int k = karg < 1 ? 1 : (karg > n ? n : karg);
    for (int e = lane; e < n; e += 32) {
      const int i = e / n2, j = e - i * n2;
      work[e] = ex(i, j) ? (op == kRdxBest ? inf_ : -inf_) : at(i, j);
    }
    __syncwarp();
    A s = 0, c = 0;
    for (int e = lane; e < n; e += 32) {
      const A v = work[e];
      int rank = 0;
      for (int f = 0; f < n; ++f) {
        const A w = work[f];
        rank += op == kRdxBest ? ((w < v) || (w == v && f < e)) : ((w > v) || (w == v && f < e));
      }
      if (rank < k) {
        // mmean(sel, mask = sel >= ctt): best keeps what lies below +inf_, worst excludes everything (>= -inf_ always)
        const bool excluded = op == kRdxBest ? (v >= inf_) : (v >= -inf_);
        s += excluded ? (A)0 * v : v;
        c += excluded ? (A)0 : (A)1;
      }
    }
    __syncwarp();
    s = wsum(s);
    c = wsum(c);
    return s / (c > eps ? c : eps);
  }
  // ---- bpwr: greedy best pairs without replacement (the block is already jittered by the caller)
  {
    const int rounds = karg < 1 ? n1 : (karg > n1 ? n1 : karg);
    unsigned char* picked = reinterpret_cast<unsigned char*>(work);  // [n] bytes inside the work area
    for (int e = lane; e < n; e += 32) picked[e] = 0;
    __syncwarp();
    for (int it = 0; it < rounds; ++it) {
      A best = inf_;
      for (int e = lane; e < n; e += 32) {
        const int i = e / n2, j = e - i * n2;
        const A w = ex(i, j) ? inf_ : at(i, j);
        best = w < best ? w : best;
      }
      best = wmin(best);
      for (int e = lane; e < n; e += 32) {
        const int i = e / n2, j = e - i * n2;
        if (!ex(i, j) && at(i, j) <= best) picked[e] = 1;
      }
      if (it < rounds - 1) {
        // rows / columns whose minimum over the included entries is <= best drop out (one row and one column per lane)
        bool row_out = false, col_out = false;
        if (lane < n1) {
          A m = inf_;
          for (int j = 0; j < n2; ++j) { const A w = ex(lane, j) ? inf_ : at(lane, j); m = w < m ? w : m; }
          row_out = m <= best;
        }
        if (lane < n2) {
          A m = inf_;
          for (int i = 0; i < n1; ++i) { const A w = ex(i, lane) ? inf_ : at(i, lane); m = w < m ? w : m; }
          col_out = m <= best;
        }
        __syncwarp();
        const unsigned rows = __ballot_sync(kFull, row_out), cols = __ballot_sync(kFull, col_out);
        for (int e = lane; e < n; e += 32) {
          const int i = e / n2, j = e - i * n2;
          if (((rows >> i) & 1u) || ((cols >> j) & 1u)) mk[i * ld_i + j * ld_j] = 1;
        }
      }
      __syncwarp();
    }
    A s = 0, c = 0;
    for (int e = lane; e < n; e += 32) {
      const int i = e / n2, j = e - i * n2;
      s += picked[e] ? at(i, j) : (A)0 * at(i, j);
      c += picked[e] ? (A)1 : (A)0;
    }
    s = wsum(s);
    c = wsum(c);
    (void)agg;
    return s / (c > eps ? c : eps);
  }
}

template <typename A>
__host__ __device__ constexpr int redux_warps() { return sizeof(A) == 8 ? 2 : 4; }  // static shared memory: 2 x 1024 A + 2 x 1024 bytes per warp

template <typename T, typename A>
__global__ void __launch_bounds__(128) redux_pairs_kernel(const T* __restrict__ dist, const unsigned char* __restrict__ mask,
                                                          long long pairs, int s1, int s2, int op, int karg, int symmetric, A eps,
                                                          A inf_, T* __restrict__ out) {
  constexpr int kW = redux_warps<A>();
  __shared__ A xs[kW][kReduxMaxS * kReduxMaxS];
  __shared__ A ws[kW][kReduxMaxS * kReduxMaxS];
  __shared__ unsigned char ms[kW][kReduxMaxS * kReduxMaxS];
  __shared__ unsigned char ms0[kW][kReduxMaxS * kReduxMaxS];
  __shared__ A ag[kW][32];
  const int w = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
  const long long pair = blockIdx.x * (long long)kW + w;
  if (pair >= pairs) return;
  const int n = s1 * s2;
  const T* src = dist + pair * n;
  for (int e = lane; e < n; e += 32) {
    if constexpr (sizeof(A) == 8) xs[w][e] = (A)src[e]; else xs[w][e] = (A)to_f32<T>(src[e]);
    ms0[w][e] = mask ? (unsigned char)(mask[pair * n + e] != 0) : 0;
    ms[w][e] = ms0[w][e];
  }
  __syncwarp();
  // bpwr works on the orientation with the fewer rows (the reference transposes when s2 < s1)
  const bool bp = op == kRdxBpwr;
  auto run = [&](bool transposed) {
    const bool t = transposed != (bp && ((transposed ? s1 : s2) < (transposed ? s2 : s1)));
    const int n1 = t ? s2 : s1, n2 = t ? s1 : s2;
    return redux_block<A>(op, karg, xs[w], ms[w], n1, n2, t ? 1 : s2, t ? s2 : 1, eps, inf_, lane, ag[w], ws[w]);
  };
  A r = run(false);
  if (symmetric) {
    __syncwarp();
    for (int e = lane; e < n; e += 32) ms[w][e] = ms0[w][e];  // (bpwr edits its mask)
    __syncwarp();
    r = (A)0.5 * (r + run(true));
  }
  if (lane == 0) {
    if constexpr (sizeof(A) == 8) out[pair] = (T)r; else out[pair] = from_f32<T>((float)r);
  }
}

}  // namespace wealy
