// Epilogue policies plugged into gemm_core.cuh.  Each thread owns ONE accumulator row (one query)
// and receives its similarities 32 consecutive candidates at a time, straight from TMEM.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "gemm_core.cuh"

namespace wealy {

// ---------------------------------------------------------------------------------------------
// StoreEpi: materialise the (n, m) matrix -- pairwise_distance_matrix (lib/tensor_ops.py:152-176)
// and pairwise_euclidean_distance_matrix (:131-149).  HBM-write bound; used for the drop-in
// `pairwise_distance_matrix` call and as the debug view of the fused kernels.
// ---------------------------------------------------------------------------------------------
enum SimMode : int {
  kSimCossim = 0,  // s                      (inputs pre-normalised by prep)
  kSimCos = 1,     // 1 - s
  kSimDotsim = 2,  // s * rs * cs            (power-of-two row scales folded back)
  kSimDot = 3,     // 1 - s * rs * cs
  kSimSqeuc = 4,   // max(|x|^2 - 2 x.y + |y|^2, 0) * post
  kSimEuc = 5,     // sqrt(max(..., 0)) * post
};

enum OutDtype : int { kOutF32 = 0, kOutF16 = 1, kOutBF16 = 2 };

struct StoreParams {
  void* out;
  long long ld;       // elements between consecutive output rows
  int mode;
  int out_dtype;
  float post;         // post scale (1/D for nsqeuc, D^-1/2 for neuc/nfro)
  const float* rscale;  // [n] row factors (RAW modes) or nullptr
  const float* cscale;  // [m]
  const float* rsq;     // [n] squared norms (euclidean modes)
  const float* csq;     // [m]
};

struct StoreEpi {
  using Params = StoreParams;
  static constexpr int kWarpScratchBytes = 0;
  struct RowState {
    float rs, rq;
    bool valid;
  };

  __device__ static __forceinline__ void row_begin(const Params& p, RowState& st, int row, int, const GemmShape& sh, uint8_t*) {
    st.valid = row < sh.m_rows;
    st.rs = (st.valid && p.rscale) ? p.rscale[row] : 1.f;
    st.rq = (st.valid && p.rsq) ? p.rsq[row] : 0.f;
  }

  __device__ static __forceinline__ float transform(const Params& p, const RowState& st, float s, int col) {
    switch (p.mode) {
      case kSimCossim: return s;
      case kSimCos: return 1.f - s;
      case kSimDotsim: return s * st.rs * __ldg(p.cscale + col);
      case kSimDot: return 1.f - s * st.rs * __ldg(p.cscale + col);
      default: {
        const float dot = s * st.rs * __ldg(p.cscale + col);
        float d2 = st.rq - 2.f * dot + __ldg(p.csq + col);
        d2 = d2 <= 0.f ? 0.f : d2;
        return (p.mode == kSimSqeuc ? d2 : sqrtf(d2)) * p.post;
      }
    }
  }

  __device__ static __forceinline__ void chunk32(const Params& p, RowState& st, int row, int col0,
                                                 const uint32_t (&acc)[32], const GemmShape& sh, uint8_t*) {
    if (!st.valid || col0 >= sh.n_cols) return;
    const int ncol = min(32, sh.n_cols - col0);
    float o[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) o[e] = e < ncol ? transform(p, st, __uint_as_float(acc[e]), col0 + e) : 0.f;
    const long long base = (long long)row * p.ld + col0;
    if (p.out_dtype == kOutF32) {
      float* dst = reinterpret_cast<float*>(p.out) + base;
      if (ncol == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
        for (int e = 0; e < 32; e += 4) *reinterpret_cast<float4*>(dst + e) = make_float4(o[e], o[e + 1], o[e + 2], o[e + 3]);
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (e < ncol) dst[e] = o[e];
      }
    } else if (p.out_dtype == kOutF16) {
      __half* dst = reinterpret_cast<__half*>(p.out) + base;
#pragma unroll
      for (int e = 0; e < 32; ++e)
        if (e < ncol) dst[e] = __float2half_rn(o[e]);
    } else {
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + base;
#pragma unroll
      for (int e = 0; e < 32; ++e)
        if (e < ncol) dst[e] = __float2bfloat16_rn(o[e]);
    }
  }

  __device__ static __forceinline__ void row_end(const Params&, RowState&, int, int, const GemmShape&, uint8_t*) {}
};

// ---------------------------------------------------------------------------------------------
// EvalEpi: fused self / same-clique masking + rank counting (+ streaming top-k candidates).
// The evaluator is absent from the reference; semantics follow
// lib/losses.py:40-42 (ids, not positions) and lib/audio_dataset/dataset.py:82-86.
//
// Per query q the plan provides its P_q relevant similarities sorted ascending (thr) and
// lim_q = thr[0].  A candidate j that is neither self (i_j == i_q) nor relevant (c_j == c_q)
// and has s_qj > lim_q is binned by the number of thresholds below it:
//     hist[q][k-1] += 1,   k = #{p : thr_p < s_qj}  (k >= 1)
// so that  #{negatives above thr_r} = sum_{k > r} hist[q][k-1]  -- ranks without any sort and
// without the N x N matrix ever leaving the SM.  The overwhelmingly common case (s <= lim_q)
// costs one max-reduction over the 32-column chunk and one compare.
// ---------------------------------------------------------------------------------------------
struct EvalParams {
  const float* lim;        // [nq] lowest relevant similarity of the query (+inf if none)
  const int* q_c;          // [nq] clique ids
  const int* q_i;          // [nq] version ids
  const int* c_c;          // [nc]
  const int* c_i;          // [nc]
  const float* thr;        // CSR: sorted ascending relevant similarities
  const long long* off;    // [nq] CSR offsets
  const int* cnt;          // [nq] P_q
  unsigned int* hist;      // CSR, same indexing as thr
  // streaming top-k (k == 0 disables)
  int topk;
  int cap;                 // candidate-buffer capacity per (row, part)
  float* cand_val;         // [parts][nq][cap]
  int* cand_idx;
  int* cand_cnt;           // [parts][nq]
  float* cand_tau;         // [parts][nq] running k-th best (lower bound filter)
  int nq_total;
};

struct EvalEpi {
  using Params = EvalParams;
  // Per-warp work queue: the elements that pass the per-row limit are scattered unevenly over the
  // 32 rows of a warp (a query with a poorly ranked relevant item passes almost everything, most
  // queries pass nothing).  Handling them in place would serialise the warp on its hottest lane,
  // so they are first compacted into a shared-memory queue and then processed 32 at a time with
  // every lane busy, whatever row they came from.
  static constexpr int kQueueCap = 256;
  static constexpr int kWarpScratchBytes = kQueueCap * 4 + kQueueCap * 2 + 32 * 4;  // values, (lane,col) tags, top-k counts
  struct RowState {
    float lim;      // min(lowest threshold, top-k filter): the only compare on the fast path
    float tlim;     // lowest threshold
    float tau;      // top-k filter (k-th best so far, -inf until k candidates are buffered)
    int qc, qi, cnt;
    long long off;
    long long cbase;
  };

  __device__ static __forceinline__ float* q_val(uint8_t* scratch) { return reinterpret_cast<float*>(scratch); }
  __device__ static __forceinline__ uint16_t* q_tag(uint8_t* scratch) {
    return reinterpret_cast<uint16_t*>(scratch + kQueueCap * 4);
  }
  __device__ static __forceinline__ int* n_cand(uint8_t* scratch) {
    return reinterpret_cast<int*>(scratch + kQueueCap * 6);
  }

  __device__ static __forceinline__ void row_begin(const Params& p, RowState& st, int row, int part,
                                                   const GemmShape& sh, uint8_t* scratch) {
    st.tlim = __int_as_float(0x7f800000);
    st.tau = __int_as_float(0x7f800000);
    st.qc = st.qi = st.cnt = 0;
    st.off = 0;
    st.cbase = ((long long)part * p.nq_total + row) * p.cap;
    if (row < sh.m_rows) {
      st.tlim = p.lim[row];
      st.qc = p.q_c[row];
      st.qi = p.q_i[row];
      st.cnt = p.cnt[row];
      st.off = p.off[row];
      if (p.topk > 0) st.tau = __int_as_float(0xff800000);
    }
    st.lim = fminf(st.tlim, st.tau);
    __syncwarp();
    n_cand(scratch)[ptx::lane_id()] = 0;
    __syncwarp();
  }

  // number of thresholds strictly below s (lower bound); thr[0 .. cnt) ascending
  __device__ static __forceinline__ int count_below(const float* __restrict__ thr, int cnt, float s) {
    int lo = 0, hi = cnt;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(thr + mid) < s) lo = mid + 1; else hi = mid;
    }
    return lo;
  }

  // Warp-cooperative compaction of one row's candidate buffer down to its k best; returns the
  // new k-th best value.  Rank-by-counting over <= cap entries (cap is a few hundred).
  __device__ static __noinline__ float compact_row(float* val, int* idx, int n, int k, int lane) {
    // every lane ranks the entries lane, lane+32, ...; entry e survives iff
    // rank(e) = #{f : val_f > val_e or (val_f == val_e and f < e)} < k ; survivors are scattered
    // to position rank (unique), staged in registers so reads complete before any write.
    constexpr int kMaxPerLane = 32;  // supports cap <= 1024
    __syncwarp();  // appends of the other lanes must be visible to the whole warp
    float mv[kMaxPerLane];
    int mi[kMaxPerLane], mr[kMaxPerLane];
    int cntl = 0;
    for (int e = lane; e < n; e += 32) {
      const float ve = val[e];
      int r = 0;
      for (int f = 0; f < n; ++f) {
        const float vf = val[f];
        r += (vf > ve) || (vf == ve && f < e);
      }
      mv[cntl] = ve;
      mi[cntl] = idx[e];
      mr[cntl] = r;
      ++cntl;
    }
    __syncwarp();
    float kth = __int_as_float(0x7f800000);
    for (int c = 0; c < cntl; ++c) {
      if (mr[c] < k) {
        val[mr[c]] = mv[c];
        idx[mr[c]] = mi[c];
        if (mr[c] == k - 1) kth = mv[c];
      }
    }
    __syncwarp();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kth = fminf(kth, __shfl_xor_sync(0xffffffffu, kth, o));
    return kth;
  }

  __device__ static __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    return v;
  }

  __device__ static __forceinline__ void chunk32(const Params& p, RowState& st, int row, int col0,
                                                 const uint32_t (&acc)[32], const GemmShape& sh, uint8_t* scratch) {
    constexpr unsigned kFull = 0xffffffffu;
    // ---- fast path: one compare per element, nothing else when no lane of the warp passes
    unsigned m = 0;
#pragma unroll
    for (int e = 0; e < 32; ++e) m |= (__uint_as_float(acc[e]) > st.lim) ? (1u << e) : 0u;
    if (!__any_sync(kFull, m != 0)) return;

    const int lane = (int)ptx::lane_id();
    float* qv = q_val(scratch);
    uint16_t* qt = q_tag(scratch);
    int* ncand = n_cand(scratch);
    // ids of the 32 candidates of this chunk, one per lane (coalesced), shared by every queued element
    const int mycol = col0 + lane;
    const int colok = mycol < sh.n_cols;
    const int cc = colok ? __ldg(p.c_c + mycol) : 0;
    const int ci = colok ? __ldg(p.c_i + mycol) : 0;

    const int total = __shfl_sync(kFull, warp_incl_scan(__popc(m), lane), 31);
    const int nb = total <= kQueueCap ? 1 : 4;  // 8 columns x 32 rows always fit
    for (int bi = 0; bi < nb; ++bi) {
      const unsigned mb = nb == 1 ? m : (m & (0xffu << (8 * bi)));
      const int mine = __popc(mb);
      const int incl = warp_incl_scan(mine, lane);
      const int btotal = __shfl_sync(kFull, incl, 31);
      if (btotal == 0) continue;
      int pos = incl - mine;
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        if (mb & (1u << e)) {
          qv[pos] = __uint_as_float(acc[e]);
          qt[pos] = (uint16_t)((lane << 5) | e);
          ++pos;
        }
      }
      __syncwarp();
      for (int r0 = 0; r0 < btotal; r0 += 32) {
        const int r = r0 + lane;
        const bool active = r < btotal;
        const float s = active ? qv[r] : 0.f;
        const int tag = active ? (int)qt[r] : 0;
        const int L = tag >> 5, e = tag & 31;  // owner lane (row) and column of the element
        const int qc = __shfl_sync(kFull, st.qc, L);
        const int qi = __shfl_sync(kFull, st.qi, L);
        const int pc = __shfl_sync(kFull, st.cnt, L);
        const float tl = __shfl_sync(kFull, st.tlim, L);
        const float tau = __shfl_sync(kFull, st.tau, L);
        const long long off = __shfl_sync(kFull, st.off, L);
        const int ccol = __shfl_sync(kFull, cc, e);
        const int cicol = __shfl_sync(kFull, ci, e);
        const int ok = __shfl_sync(kFull, colok, e);
        if (active && ok && cicol != qi) {  // i_j == i_q: self (or a version-id collision), never a candidate
          if (p.topk > 0 && s > tau) {
            const int slot = atomicAdd(&ncand[L], 1);
            const long long cb = st.cbase + (long long)(L - lane) * p.cap + slot;  // rows of a warp are consecutive
            p.cand_val[cb] = s;
            p.cand_idx[cb] = col0 + e;
          }
          if (s > tl && ccol != qc) {  // a negative above at least the lowest relevant item
            const int k = count_below(p.thr + off, pc, s);
            if (k > 0) atomicAdd(p.hist + off + (k - 1), 1u);
          }
        }
      }
      __syncwarp();
    }

    if (p.topk > 0) {
      // a row gains at most 32 candidates per chunk: compact (warp-cooperatively, one row at a time)
      // as soon as fewer than 32 free slots remain
      unsigned need = __ballot_sync(kFull, ncand[lane] > p.cap - 32);
      while (need) {
        const int src = __ffs(need) - 1;
        need &= need - 1;
        const long long cb = st.cbase + (long long)(src - lane) * p.cap;
        const int n = ncand[src];
        const float kth = compact_row(p.cand_val + cb, p.cand_idx + cb, n, p.topk, lane);
        if (lane == src) {
          ncand[lane] = p.topk;
          st.tau = kth;
          st.lim = fminf(st.tlim, st.tau);
        }
        __syncwarp();
      }
    }
    (void)row;
  }

  __device__ static __forceinline__ void row_end(const Params& p, RowState& st, int row, int part,
                                                 const GemmShape& sh, uint8_t* scratch) {
    __syncwarp();
    if (p.topk > 0 && row < sh.m_rows) p.cand_cnt[(long long)part * p.nq_total + row] = n_cand(scratch)[ptx::lane_id()];
    __syncwarp();
    (void)st;
  }
};

}  // namespace wealy
