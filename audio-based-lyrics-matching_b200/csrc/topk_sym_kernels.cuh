// Top-k inside the SYMMETRIC all-vs-all sweep (a7 with top-k output, e.g. BASELINE configs[4]: 50k x 2048, top-100).
//
// The rectangle sweep keeps a streaming top-k per (row, part) with a filter that tightens as the row's sweep
// advances; that needs every query to see ALL its candidates in its own row, i.e. the full N x N contraction.  The
// symmetric sweep contracts only the tiles above the diagonal and scores every element for its row query AND its
// column query, so a query's candidates arrive from many CTAs in no particular order.  What makes top-k possible
// there is a STATIC per-query lower bound beta_q of the k-th best similarity, known before the sweep starts:
//
//   1. sample:   S candidates, evenly spaced in clique-sorted order (S ~ N / 16, at least 4096);
//   2. pre-pass: every query against the sample with the rectangle kernel's streaming top-r (one fp16 pass: the
//                bound does not have to be exact), r = max(24, 3 k S / N), so that about 3 k candidates of the whole
//                corpus lie above the sample's r-th best (relative spread of that count: 1 / sqrt(r));
//   3. sweep:    EvalSymEpi<..., kTopk = true> appends every element above beta_row / beta_col to the row's / the
//                column's list (atomic cursor; lists hold 3 k + 8 sigma entries);
//   4. finalize: per query, select the k best of its list, order them, translate plane rows to the caller's indices.
//                A list that overflowed or ended short of k (adversarial data: thousands of exact ties around the
//                bound) is counted in `fail`; the host then re-runs the call on the rectangle kernel.
#pragma once
#include "epilogues.cuh"

namespace wealy {

// rows of the sample: sorted positions k * n / S, k = 0 .. S - 1 (distinct for S <= n) -> their plane rows copied into
// a dense [S][d_pad] operand (hi plane only: the pre-pass runs one fp16 pass); ids for the self test
__global__ void __launch_bounds__(256) sample_rows_kernel(const __half* __restrict__ hi, int d_pad, int n, int n_sample,
                                                          const int* __restrict__ s_i, __half* __restrict__ out,
                                                          int* __restrict__ out_i) {
  const int k = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
  const int lane = (int)(threadIdx.x & 31);
  if (k >= n_sample) return;
  const int srow = (int)(((long long)k * n) / n_sample);
  const uint4* src = reinterpret_cast<const uint4*>(hi + (long long)spread_plane_of(srow) * d_pad);
  uint4* dst = reinterpret_cast<uint4*>(out + (long long)k * d_pad);
  for (int v = lane; v < (d_pad >> 3); v += 32) dst[v] = __ldg(src + v);
  if (lane == 0) out_i[k] = s_i[srow];
}

// per-plane-row arrays the rectangle epilogue wants for the pre-pass: version id of the row (by sorted index),
// "no relevant item" limits (+inf: the pre-pass does no rank counting), zero counts / offsets
__global__ void plane_ids_kernel(const int* __restrict__ s_i, int n, int n_rows, int* __restrict__ qi_plane,
                                 float* __restrict__ lim_inf, int* __restrict__ zeros_i, long long* __restrict__ zeros_ll) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p > n_rows) return;
  zeros_ll[p] = 0;
  if (p == n_rows) return;
  const int srow = spread_sorted_of(p);
  qi_plane[p] = srow < n ? s_i[srow] : -1;
  lim_inf[p] = __int_as_float(0x7f800000);
  zeros_i[p] = 0;
}

// beta[p] = r-th best similarity of plane row p against the sample (from the pre-pass's candidate lists), also written
// into lvl[p].w where the sweep's per-tile column slots pick it up; padded rows keep +inf; a row with fewer than r
// candidates gets -inf (everything passes: its list overflows and the call falls back).  parts * cap <= 1024.
__global__ void __launch_bounds__(128) topk_beta_kernel(const float* __restrict__ cand_val, const int* __restrict__ cand_cnt,
                                                        int parts, int n_rows, int cap, int r, int n,
                                                        float* __restrict__ beta, float4* __restrict__ lvl, int n_lvl,
                                                        int* __restrict__ tk_cnt) {
  __shared__ float sv_all[4][1024];
  __shared__ int si_all[4][1024];
  const int p = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
  const int lane = (int)(threadIdx.x & 31);
  if (p >= n_lvl) return;
  float* sv = sv_all[threadIdx.x >> 5];
  int* si = si_all[threadIdx.x >> 5];
  if (lane == 0 && p < n_rows) tk_cnt[p] = 0;
  float b = __int_as_float(0x7f800000);
  if (p < n_rows && spread_sorted_of(p) < n) {
    int m = 0;
    for (int part = 0; part < parts; ++part) {
      const long long base = ((long long)part * n_rows + p) * cap;
      const int np = min(cand_cnt[(long long)part * n_rows + p], cap);
      for (int e = lane; e < np; e += 32) {
        sv[m + e] = cand_val[base + e];
        si[m + e] = e;
      }
      m += np;
    }
    __syncwarp();
    if (m > r) {
      b = warp_select_topk<32>(sv, si, m, r, lane);
    } else if (m == r) {
      float lo = __int_as_float(0x7f800000);
      for (int e = lane; e < m; e += 32) lo = fminf(lo, sv[e]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      b = lo;
    } else {
      b = __int_as_float(0xff800000);
    }
  }
  if (lane == 0) {
    if (p < n_rows) beta[p] = b;
    lvl[p].w = b;
  }
}

// finalize: one warp per plane row (query).  Its list holds every candidate above beta (plane rows); select the k best,
// order them (descending similarity, ties -> lower caller index, like a stable ascending-distance argsort) and write
// them at the caller's row.  Lists that overflowed or hold fewer than k entries are reported in `fail`.
__global__ void __launch_bounds__(128) topk_sym_finalize_kernel(float* __restrict__ tk_val, int* __restrict__ tk_idx,
                                                                const int* __restrict__ tk_cnt, int cap, int n_rows, int n,
                                                                int k, const int* __restrict__ perm,
                                                                long long* __restrict__ out_idx, float* __restrict__ out_sim,
                                                                int* __restrict__ fail) {
  const int p = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
  const int lane = (int)(threadIdx.x & 31);
  if (p >= n_rows) return;
  const int srow = spread_sorted_of(p);
  if (srow >= n) return;
  const long long q = perm[srow];
  const int m = tk_cnt[p];
  if (m > cap || m < k) {
    if (lane == 0) atomicAdd(fail, 1);
    return;
  }
  float* sv = tk_val + (long long)p * cap;
  int* si = tk_idx + (long long)p * cap;
  // plane row -> caller index first, so that ties are broken like the rectangle path does
  for (int e = lane; e < m; e += 32) si[e] = perm[spread_sorted_of(si[e])];
  if (m > k) warp_select_topk<32>(sv, si, m, k, lane);
  __syncwarp();
  for (int e = lane; e < k; e += 32) {
    const float ve = sv[e];
    const int ie = si[e];
    int r = 0;
    for (int f = 0; f < k; ++f) {
      const float vf = sv[f];
      r += (vf > ve) || (vf == ve && si[f] < ie);
    }
    out_idx[q * k + r] = ie;
    out_sim[q * k + r] = ve;
  }
}

}  // namespace wealy
