// float64 inputs (drop-in completeness: the reference returns the input dtype, lib/tensor_ops.py:152-176, SURVEY.md
// section 4 "Output dtype == input dtype").  Double precision has no tensor-core path worth the name on this part and
// float64 embeddings only appear in tests / debugging, so these are plain CUDA-core kernels: a shared-memory tiled
// DGEMM with the mode epilogue of StoreEpi, and a one-pass masked reduction with double accumulators.
#pragma once
#include "epilogues.cuh"

namespace wealy {

// L2 norm and squared norm of every row
__global__ void __launch_bounds__(256) row_norm_f64_kernel(const double* __restrict__ x, long long ld, int n, int d,
                                                           double* __restrict__ norm, double* __restrict__ sq) {
  const int row = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
  const int lane = (int)(threadIdx.x & 31);
  if (row >= n) return;
  const double* r = x + (long long)row * ld;
  double acc = 0.0;
  for (int k = lane; k < d; k += 32) acc = fma(r[k], r[k], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    sq[row] = acc;
    norm[row] = sqrt(acc);
  }
}

// out[i][j] = f(x_i . y_j): 64 x 64 output tile per CTA (16 x 16 threads, 4 x 4 outputs each), k-tiles of 16
__global__ void __launch_bounds__(256) sim_matrix_f64_kernel(const double* __restrict__ x, long long ldx, int n,
                                                             const double* __restrict__ y, long long ldy, int m, int d,
                                                             int mode, double eps, double post,
                                                             const double* __restrict__ xn, const double* __restrict__ xs,
                                                             const double* __restrict__ yn, const double* __restrict__ ys,
                                                             double* __restrict__ out, long long ld_out) {
  __shared__ double xs_t[16][65];
  __shared__ double ys_t[16][65];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  for (int k0 = 0; k0 < d; k0 += 16) {
    // 64 rows x 16 k of each operand: 1024 elements, 4 per thread; k fastest in global memory
    for (int e = threadIdx.x; e < 1024; e += 256) {
      const int rr = e >> 4, kk = e & 15;
      const int k = k0 + kk;
      xs_t[kk][rr] = (r0 + rr < n && k < d) ? x[(long long)(r0 + rr) * ldx + k] : 0.0;
      ys_t[kk][rr] = (c0 + rr < m && k < d) ? y[(long long)(c0 + rr) * ldy + k] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[u] = xs_t[kk][ty + 16 * u];
        b[u] = ys_t[kk][tx + 16 * u];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fma(a[u], b[v], acc[u][v]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int r = r0 + ty + 16 * u;
    if (r >= n) continue;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int c = c0 + tx + 16 * v;
      if (c >= m) continue;
      double s = acc[u][v], o;
      switch (mode) {
        case kSimCossim: o = s / ((xn[r] + eps) * (yn[c] + eps)); break;
        case kSimCos: o = 1.0 - s / ((xn[r] + eps) * (yn[c] + eps)); break;
        case kSimDotsim: o = s; break;
        case kSimDot: o = 1.0 - s; break;
        default: {
          double d2 = xs[r] - 2.0 * s + ys[c];
          d2 = d2 <= 0.0 ? 0.0 : d2;
          o = (mode == kSimSqeuc ? d2 : sqrt(d2)) * post;
        }
      }
      out[(long long)r * ld_out + c] = o;
    }
  }
}

// masked reduction over the columns of a [rows][cols] float64 matrix (mask non-zero = EXCLUDED), one warp per row;
// same arithmetic as the fp32 kernels of masked_kernels.cuh (included * x, +-fill for min / max, NaN propagation)
__global__ void __launch_bounds__(256) masked_reduce_f64_kernel(const double* __restrict__ x, const unsigned char* __restrict__ mask,
                                                                long long rows, long long cols, int op, double fill, double eps,
                                                                double* __restrict__ out) {
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = (int)(threadIdx.x & 31);
  if (row >= rows) return;
  double sum = 0.0, cnt = 0.0, mn = __longlong_as_double(0x7ff0000000000000ll), mx = __longlong_as_double(0xfff0000000000000ll);
  bool nan = false;
  const long long base = row * cols;
  for (long long c = lane; c < cols; c += 32) {
    const double v = x[base + c];
    const bool excluded = mask != nullptr && mask[base + c] != 0;
    sum += excluded ? 0.0 * v : v;
    cnt += excluded ? 0.0 : 1.0;
    const double w = excluded ? fill : v;
    mn = fmin(mn, w);
    mx = fmax(mx, w);
    nan |= (w != w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  nan = __any_sync(0xffffffffu, nan);
  if (lane == 0) {
    double r;
    switch (op) {
      case 0: r = sum; break;
      case 1: r = sum / fmax(cnt, eps); break;
      case 2: r = nan ? __longlong_as_double(0x7ff8000000000000ll) : mn; break;
      default: r = nan ? __longlong_as_double(0x7ff8000000000000ll) : mx;
    }
    out[row] = r;
  }
}

}  // namespace wealy
