// a3: masked reductions (lib/tensor_ops.py:182-258) as one HBM-bound pass: x and the exclusion mask
// are read once (coalesced), never materialising `included * x` or the +-inf filled copy the
// reference builds.  Mask polarity: non-zero == EXCLUDED (lib/tensor_ops.py:186,201,219).
#pragma once
#include "prep.cuh"
#include "loss_kernels.cuh"

namespace wealy {

enum MaskedOp : int { kMSum = 0, kMMean = 1, kMMin = 2, kMMax = 3 };

struct MaskedAcc {
  float sum, cnt, mn, mx;
};

template <typename T>
__device__ __forceinline__ void masked_visit(MaskedAcc& a, const T* __restrict__ x, const unsigned char* __restrict__ mask,
                                             long long i, float fill) {
  const float v = to_f32<T>(x[i]);
  const bool excluded = mask != nullptr && mask[i] != 0;
  // reference arithmetic: included * x (so an excluded inf/nan still poisons the sum, as upstream)
  a.sum += excluded ? 0.f * v : v;
  a.cnt += excluded ? 0.f : 1.f;
  const float w = excluded ? fill : v;
  a.mn = fminf(a.mn, w);
  a.mx = fmaxf(a.mx, w);
  if (w != w) { a.mn = w; a.mx = w; }  // torch.min / max propagate NaN
}

__device__ __forceinline__ float masked_result(const MaskedAcc& a, int op, float eps) {
  switch (op) {
    case kMSum: return a.sum;
    case kMMean: return a.sum / fmaxf(a.cnt, eps);  // den.clamp(min=eps), lib/tensor_ops.py:212
    case kMMin: return a.mn;
    default: return a.mx;
  }
}

// one warp per row (short rows: chunk-level reductions, per-anchor statistics)
template <typename T>
__global__ void __launch_bounds__(256) masked_reduce_warp_kernel(const T* __restrict__ x,
                                                                 const unsigned char* __restrict__ mask, long long rows,
                                                                 long long cols, int op, float fill, float eps,
                                                                 T* __restrict__ out) {
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = (int)(threadIdx.x & 31);
  if (row >= rows) return;
  MaskedAcc a{0.f, 0.f, __int_as_float(0x7f800000), __int_as_float(0xff800000)};
  const long long base = row * cols;
  for (long long c = lane; c < cols; c += 32) masked_visit<T>(a, x, mask, base + c, fill);
  a.sum = warp_sum(a.sum);
  a.cnt = warp_sum(a.cnt);
  const bool nan = __any_sync(0xffffffffu, a.mn != a.mn);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a.mn = fminf(a.mn, __shfl_xor_sync(0xffffffffu, a.mn, o));
    a.mx = fmaxf(a.mx, __shfl_xor_sync(0xffffffffu, a.mx, o));
  }
  if (nan) a.mn = a.mx = __int_as_float(0x7fc00000);
  if (lane == 0) out[row] = from_f32<T>(masked_result(a, op, eps));
}

// one block per row (long rows: whole-matrix statistics such as v_dpos / v_dneg, lib/losses.py:267-268)
template <typename T>
__global__ void __launch_bounds__(1024) masked_reduce_block_kernel(const T* __restrict__ x,
                                                                   const unsigned char* __restrict__ mask, long long rows,
                                                                   long long cols, int op, float fill, float eps,
                                                                   T* __restrict__ out) {
  __shared__ float s_sum[32], s_cnt[32], s_mn[32], s_mx[32];
  __shared__ int s_nan;
  const long long row = blockIdx.x;
  if (threadIdx.x == 0) s_nan = 0;
  __syncthreads();
  MaskedAcc a{0.f, 0.f, __int_as_float(0x7f800000), __int_as_float(0xff800000)};
  const long long base = row * cols;
  for (long long c = threadIdx.x; c < cols; c += blockDim.x) masked_visit<T>(a, x, mask, base + c, fill);
  if (a.mn != a.mn) atomicOr(&s_nan, 1);
  a.sum = warp_sum(a.sum);
  a.cnt = warp_sum(a.cnt);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a.mn = fminf(a.mn, __shfl_xor_sync(0xffffffffu, a.mn, o));
    a.mx = fmaxf(a.mx, __shfl_xor_sync(0xffffffffu, a.mx, o));
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { s_sum[w] = a.sum; s_cnt[w] = a.cnt; s_mn[w] = a.mn; s_mx[w] = a.mx; }
  __syncthreads();
  if (w == 0) {
    const int nw = blockDim.x >> 5;
    MaskedAcc b{l < nw ? s_sum[l] : 0.f, l < nw ? s_cnt[l] : 0.f, l < nw ? s_mn[l] : __int_as_float(0x7f800000),
                l < nw ? s_mx[l] : __int_as_float(0xff800000)};
    b.sum = warp_sum(b.sum);
    b.cnt = warp_sum(b.cnt);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      b.mn = fminf(b.mn, __shfl_xor_sync(0xffffffffu, b.mn, o));
      b.mx = fmaxf(b.mx, __shfl_xor_sync(0xffffffffu, b.mx, o));
    }
    if (s_nan) b.mn = b.mx = __int_as_float(0x7fc00000);
    if (l == 0) out[row] = from_f32<T>(masked_result(b, op, eps));
  }
}

}  // namespace wealy
