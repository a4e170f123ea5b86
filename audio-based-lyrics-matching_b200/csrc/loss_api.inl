// a5 / a6: NT-Xent and CLEWS forward / backward (included by api.cu).
//
// Workspace layout (fixed by (b, d, passes) so that backward finds what forward left):
//   U planes (hi[,lo] [b][d_pad], norm, scale, sq) | U^T planes (hi[,lo] [d][b_pad]) | W planes (hi[,lo] [b][b_pad])
//   | dU [b][d] f32 | {label, idx} [b256] int2 | partial [parts_max][b][8] f32 | rowstat [b256][4] f32 | ZStats | scal | flags

constexpr int kLossZeroBytes = 1024 + 4 * 256;  // ZStats | scal | bad | acc | acc_max
static_assert(sizeof(ZStats) <= 1024, "ZStats block");

struct LossWs {
  Planes u;
  __half *ut_hi, *ut_lo;
  __half *w_hi, *w_lo;
  float* du;
  int2* lab_idx;
  float* partial;
  float* rowstat;
  ZStats* zs;
  float* scal;
  int* bad;
  double* acc;
  unsigned int* acc_max;
  int64_t b_pad;
  int parts_max;
  size_t bytes;
};

static void loss_ws_layout(LossWs& w, uint8_t* base, int64_t b, int64_t d, int passes, int64_t nb = -1) {
  if (nb < 0) nb = b;  // anchors owned by this rank (data-parallel: b is the global batch)
  uint8_t* cur = base;
  w.b_pad = pad_k(b);
  carve_planes(w.u, cur, b, d, passes);
  const size_t tplane = align_up((size_t)d * w.b_pad * 2, 1024);
  w.ut_hi = reinterpret_cast<__half*>(cur); cur += tplane;
  w.ut_lo = nullptr;
  if (passes == 3) { w.ut_lo = reinterpret_cast<__half*>(cur); cur += tplane; }
  const size_t wplane = align_up((size_t)nb * w.b_pad * 2, 1024);
  w.w_hi = reinterpret_cast<__half*>(cur); cur += wplane;
  w.w_lo = nullptr;
  if (passes == 3) { w.w_lo = reinterpret_cast<__half*>(cur); cur += wplane; }
  w.du = reinterpret_cast<float*>(cur); cur += align_up((size_t)nb * d * 4, 1024);
  const size_t b256 = (size_t)ceil_div(b, kTileN) * kTileN;  // per-tile bulk copies read whole 256-column groups
  w.lab_idx = reinterpret_cast<int2*>(cur); cur += align_up(b256 * 8, 256);
  // partial records: 2 per column tile (one per epilogue warp of a lane quadrant), for the whole-batch sweep plus --
  // when the forward is split into "own column block first, the rest behind the all-gather" -- the own block's sweep
  w.parts_max = (int)(ceil_div(b, kTileN) + ceil_div(nb, kTileN)) * 2;
  w.partial = reinterpret_cast<float*>(cur); cur += align_up((size_t)w.parts_max * nb * kStatWidth * 4, 1024);
  w.rowstat = reinterpret_cast<float*>(cur); cur += align_up(b256 * 16, 256);
  w.zs = reinterpret_cast<ZStats*>(cur); cur += 1024;  // (the blocks from here to acc_max are zeroed together: kLossZeroBytes)
  w.scal = reinterpret_cast<float*>(cur); cur += 256;
  w.bad = reinterpret_cast<int*>(cur); cur += 256;
  w.acc = reinterpret_cast<double*>(cur); cur += 256;
  w.acc_max = reinterpret_cast<unsigned int*>(cur); cur += 256;
  w.bytes = (size_t)(cur - base);
}

static size_t loss_ws_bytes(int64_t b, int64_t d, int passes, int64_t nb) {
  if (b <= 0 || d <= 0 || nb < 0 || nb > b) return 0;
  LossWs w;
  loss_ws_layout(w, reinterpret_cast<uint8_t*>(uintptr_t(1024)), b, d, passes, nb);
  return w.bytes + 1024;
}

extern "C" size_t wealy_loss_workspace_bytes(int64_t b, int64_t d, int passes) { return loss_ws_bytes(b, d, passes, b); }
extern "C" size_t wealy_loss_dp_workspace_bytes(int64_t b_global, int64_t d, int passes, int64_t nb) {
  return loss_ws_bytes(b_global, d, passes, nb);
}

static int loss_check(const wealy_loss_cfg* cfg, const void* z, int64_t b, int64_t d, void* ws, size_t ws_bytes,
                      int64_t row0, int64_t nb) {
  if (!cfg || !z || !ws) return fail(WEALY_ERR_BAD_ARG, "null pointer");
  if (cfg->kind != WEALY_LOSS_NTXENT && cfg->kind != WEALY_LOSS_CLEWS) return fail(WEALY_ERR_BAD_ARG, "unknown loss kind %d", cfg->kind);
  if (cfg->passes != 1 && cfg->passes != 3) return fail(WEALY_ERR_BAD_ARG, "passes must be 1 or 3");
  if (b <= 0 || d <= 0) return fail(WEALY_ERR_BAD_ARG, "bad shape b=%lld d=%lld", (long long)b, (long long)d);
  if (row0 < 0 || nb <= 0 || row0 + nb > b) return fail(WEALY_ERR_BAD_ARG, "bad shard [%lld, +%lld) of %lld", (long long)row0, (long long)nb, (long long)b);
  if (b >= (1 << 24)) return fail(WEALY_ERR_UNSUPPORTED, "batch too large");
  if (ws_bytes < loss_ws_bytes(b, d, cfg->passes, nb))
    return fail(WEALY_ERR_WORKSPACE, "workspace too small: %zu < %zu", ws_bytes, loss_ws_bytes(b, d, cfg->passes, nb));
  return WEALY_OK;
}

static void loss_params(LossParams& lp, const wealy_loss_cfg* cfg, const LossWs& w, int64_t b, int64_t row0, int64_t nb) {
  memset(&lp, 0, sizeof(lp));
  const float log2e = 1.4426950408889634f;
  lp.kind = cfg->kind;
  lp.b = (int)b;
  lp.row0 = (int)row0;
  lp.nb = (int)nb;
  lp.lab_idx = w.lab_idx;
  lp.c2 = log2e / cfg->temperature;
  lp.g2 = cfg->gamma * log2e;
  lp.b2 = cfg->b * log2e;
  lp.partial = w.partial;
  lp.rowstat = w.rowstat;
  lp.w_hi = w.w_hi;
  lp.w_lo = w.w_lo;
  lp.ldw = w.b_pad;
}

static LossCfgDev loss_cfg_dev(const wealy_loss_cfg* cfg) {
  LossCfgDev dc;
  dc.kind = cfg->kind;
  dc.temperature = cfg->temperature;
  dc.gamma = cfg->gamma;
  dc.b = cfg->b;
  dc.eps = cfg->eps;
  dc.epsilon = cfg->epsilon;
  dc.uw = cfg->uw;
  dc.numerically_friendly = cfg->numerically_friendly;
  return dc;
}

// The operand planes of the A side of a sharded launch: the anchors [row0, row0 + nb) of the global planes.
static Planes plane_rows(const Planes& p, int64_t row0, int64_t nb) {
  Planes q = p;
  q.hi = p.hi + row0 * p.d_pad;
  q.lo = p.lo ? p.lo + row0 * p.d_pad : nullptr;
  q.norm = p.norm + row0;
  q.scale = p.scale + row0;
  q.sq = p.sq + row0;
  q.rows = nb;
  return q;
}

// forward, part 1 (everything before the batch-wide sums are known): ids, prep of ALL rows, statistics sweep of the
// anchors [row0, row0 + nb) against all b columns, per-anchor merge -> rowstat[row0 .. row0 + nb), partial sums in acc.
// phase 0: all of it.  Split form for the data-parallel path (the other ranks' rows arrive by all-gather):
//   phase 1: ids + prep of the rank's OWN rows + sweep of its anchors against its own column block -- needs nothing
//            from the other ranks, runs while the all-gather is in flight;
//   phase 2: prep of the other rows + sweep against every other column block + merge.
// (own block not aligned to the 256-column tiles: phase 1 only packs the ids, phase 2 does everything.)
static int loss_forward_local(const wealy_loss_cfg* cfg, const void* z, int64_t b, int64_t ldz, int64_t d, int dtype,
                              int64_t* z_label, const int64_t* z_idx, int64_t row0, int64_t nb, void* workspace,
                              size_t workspace_bytes, cudaStream_t s, int phase = 0) {
  W_TRY(loss_check(cfg, z, b, d, workspace, workspace_bytes, row0, nb));
  if (!z_label || !z_idx) return fail(WEALY_ERR_BAD_ARG, "null pointer");
  if (phase < 0 || phase > 2) return fail(WEALY_ERR_BAD_ARG, "phase must be 0, 1 or 2");
  LossWs w;
  loss_ws_layout(w, reinterpret_cast<uint8_t*>(align_up((size_t)workspace, 1024)), b, d, cfg->passes, nb);
  const int T = 256;
  const bool ntx = cfg->kind == WEALY_LOSS_NTXENT;
  const bool split = phase != 0 && (row0 % kTileN) == 0 && ((nb % kTileN) == 0 || row0 + nb == b) && nb < b;
  const size_t esz = dtype == WEALY_F32 ? 4 : 2;
  auto prep_rows = [&](int64_t r0, int64_t nr) -> int {
    if (nr <= 0) return WEALY_OK;
    // NT-Xent: x/(|x|+1e-6) and statistics of the raw z; CLEWS: F.normalize (eps 1e-12), statistics of the normalised z
    return launch_prep(static_cast<const uint8_t*>(z) + (size_t)r0 * ldz * esz, ldz, nr, d, dtype, ntx ? kPrepL2AddEps : kPrepL2Clamp,
                       ntx ? 1e-6f : 1e-12f, plane_rows(w.u, r0, nr), nullptr, nullptr, 0, w.zs, ntx ? 0 : 1, s);
  };
  LossParams lp;
  loss_params(lp, cfg, w, b, row0, nb);
  const int halves = 2;  // epilogue warps per TMEM lane quadrant
  const int local_parts = (int)ceil_div(nb, kTileN) * halves;
  if (phase != 2) {
    if (b <= 65536) {
      // one CTA: zero {ZStats, scal, flags, batch accumulators}, optional single-label noise (in place), pack the ids
      pack_ids_noise_kernel<<<1, 1024, 0, s>>>((long long*)z_label, (const long long*)z_idx, w.lab_idx, (int)b,
                                               cfg->label_noise ? 1 : 0, reinterpret_cast<int*>(w.zs), kLossZeroBytes / 4, w.bad);
    } else {
      if (cfg->label_noise) return fail(WEALY_ERR_UNSUPPORTED, "label_noise is fused for batches of up to 65536 rows");
      CU_TRY(cudaMemsetAsync(w.zs, 0, kLossZeroBytes, s));
      pack_ids_kernel<<<(unsigned)ceil_div(b, T), T, 0, s>>>((const long long*)z_label, (const long long*)z_idx, w.lab_idx, (int)b,
                                                            w.bad);
    }
    CU_TRY(cudaGetLastError());
  }
  if (phase == 1) {
    if (!split) return WEALY_OK;
    W_TRY(prep_rows(row0, nb));
    GemmShape sh;
    fill_shape(sh, nb, nb, w.u.d_pad, 64, local_parts / halves);
    lp.col_tile_off = (int)(row0 / kTileN);
    lp.part_base = 0;
    if (sh.n_col_chunks * halves > local_parts) return fail(WEALY_ERR_UNSUPPORTED, "loss sweep: %d parts", sh.n_col_chunks * halves);
    // neutral records in the slots this launch does not write (it may use fewer column chunks than tiles)
    W_TRY(launch_gemm<LossStatsEpi>(cfg->passes, plane_rows(w.u, row0, nb), plane_rows(w.u, row0, nb), sh, lp, s));
    // remember how many slots were written: slot count is a pure function of the shape, recomputed in phase 2
    return WEALY_OK;
  }
  int parts_local_used = 0;
  if (phase == 2 && split) {
    W_TRY(prep_rows(0, row0));
    W_TRY(prep_rows(row0 + nb, b - row0 - nb));
    GemmShape shl;
    fill_shape(shl, nb, nb, w.u.d_pad, 64, local_parts / halves);
    parts_local_used = shl.n_col_chunks * halves;
  } else {
    W_TRY(prep_rows(0, b));
  }
  GemmShape sh;
  fill_shape(sh, nb, b, w.u.d_pad, 64, (w.parts_max - local_parts) / halves);
  if (parts_local_used > 0) {
    sh.skip0 = (int)(row0 / kTileN);
    sh.skip1 = (int)ceil_div(row0 + nb, kTileN);
  }
  lp.part_base = parts_local_used;
  const int parts = sh.n_col_chunks * halves;
  W_TRY(launch_gemm<LossStatsEpi>(cfg->passes, plane_rows(w.u, row0, nb), w.u, sh, lp, s));
  // (64-thread blocks: one thread per anchor walks its partial records; 256-thread blocks would put a 4096-anchor batch on 16 SMs)
  loss_merge_kernel<<<(unsigned)ceil_div(nb, 64), 64, 0, s>>>(loss_cfg_dev(cfg), (int)nb, (int)b, parts_local_used + parts, w.partial,
                                                                w.rowstat + row0 * 4, w.acc, w.acc_max);
  CU_TRY(cudaGetLastError());
  return WEALY_OK;
}

// forward, part 2: batch sums (acc, acc_max) and the per-anchor records (rowstat) of ALL b anchors are in the
// workspace (summed / gathered over the ranks by the caller when the batch is sharded) -> loss, logdict, coefficients
static int loss_forward_finish(const wealy_loss_cfg* cfg, int64_t b, int64_t d, int64_t nb, double* out, void* out_cast,
                               int cast_dtype, void* workspace, cudaStream_t s) {
  LossWs w;
  loss_ws_layout(w, reinterpret_cast<uint8_t*>(align_up((size_t)workspace, 1024)), b, d, cfg->passes, nb);
  if (out_cast && (cast_dtype < WEALY_F32 || cast_dtype > WEALY_BF16)) return fail(WEALY_ERR_BAD_ARG, "out_cast element type %d", cast_dtype);
  loss_finish_kernel<<<(unsigned)ceil_div(b, 256), 256, 0, s>>>(loss_cfg_dev(cfg), (int)b, (int)d, w.acc, w.acc_max, w.zs,
                                                                w.rowstat, w.scal, out, w.bad, out_cast, cast_dtype);
  CU_TRY(cudaGetLastError());
  return WEALY_OK;
}

// backward of the anchors [row0, row0 + nb): dz rows of this shard (complete: W is symmetrised, so the terms in
// which these rows act as columns of other ranks' anchors are included -- no reduce-scatter)
static int loss_backward_rows(const wealy_loss_cfg* cfg, const void* z, int64_t b, int64_t ldz, int64_t d, int dtype,
                              int64_t row0, int64_t nb, const void* grad_out, void* dz, int64_t ld_dz, void* workspace,
                              size_t workspace_bytes, cudaStream_t s) {
  W_TRY(loss_check(cfg, z, b, d, workspace, workspace_bytes, row0, nb));
  if (!dz) return fail(WEALY_ERR_BAD_ARG, "null pointer");
  const int gdt = cfg->grad_dtype;
  if (gdt < WEALY_F32 || gdt > WEALY_BF16) return fail(WEALY_ERR_BAD_ARG, "grad_dtype %d", gdt);
  LossWs w;
  loss_ws_layout(w, reinterpret_cast<uint8_t*>(align_up((size_t)workspace, 1024)), b, d, cfg->passes, nb);
  LossParams lp;
  loss_params(lp, cfg, w, b, row0, nb);
  // 1) W' = scale * (dL/dS + (dL/dS)^T) for this shard's rows: recompute S tile by tile, store fp16 hi/lo planes
  GemmShape sh;
  fill_shape(sh, nb, b, w.u.d_pad, 64, 1 << 20);
  W_TRY(launch_gemm<LossWEpi>(cfg->passes, plane_rows(w.u, row0, nb), w.u, sh, lp, s));
  // 2) dU = W' * U   (A = W' [nb][b_pad], B = U^T [d][b_pad], K = b_pad); the transposed planes are only
  //    needed here, so they are built now (tiled transpose, zero k-padding) rather than in the forward
  {
    dim3 grid((unsigned)ceil_div(w.b_pad, 32), (unsigned)ceil_div(d, 32), w.u.lo ? 2 : 1);
    transpose_plane_kernel<<<grid, 256, 0, s>>>(w.u.hi, w.u.lo, (int)b, (int)d, (long long)w.u.d_pad, w.ut_hi, w.ut_lo,
                                                (long long)w.b_pad);
    CU_TRY(cudaGetLastError());
  }
  Planes pw, put;
  pw.hi = w.w_hi; pw.lo = w.w_lo; pw.rows = nb; pw.d_pad = w.b_pad;
  put.hi = w.ut_hi; put.lo = w.ut_lo; put.rows = d; put.d_pad = w.b_pad;
  StoreParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.out = w.du;
  sp.ld = d;
  sp.mode = kSimCossim;
  sp.out_dtype = kOutF32;
  sp.post = 1.f;
  GemmShape sh2;
  fill_shape(sh2, nb, d, w.b_pad, 64, 1 << 20);
  W_TRY(launch_gemm<StoreEpi>(cfg->passes, pw, put, sh2, sp, s));
  // 3) normalisation Jacobian, upstream gradient, 1/(B tau) or 1/H
  const int T = 256;
  const unsigned blocks = (unsigned)ceil_div(nb * 32, T);
  const size_t esz = dtype == WEALY_F32 ? 4 : 2;
  const void* zrow = static_cast<const uint8_t*>(z) + (size_t)row0 * ldz * esz;
  const float jeps = cfg->kind == WEALY_LOSS_NTXENT ? 1e-6f : 1e-12f;  // the normalisations of the forward (prep)
  switch (dtype) {
    case WEALY_F32:
      loss_jacobian_kernel<float><<<blocks, T, 0, s>>>(cfg->kind, jeps, (const float*)zrow, (long long)ldz, (int)nb, (int)d,
                                                       w.u.norm + row0, w.du, w.scal, grad_out, (float*)dz, (long long)ld_dz, gdt);
      break;
    case WEALY_F16:
      loss_jacobian_kernel<__half><<<blocks, T, 0, s>>>(cfg->kind, jeps, (const __half*)zrow, (long long)ldz, (int)nb, (int)d,
                                                        w.u.norm + row0, w.du, w.scal, grad_out, (__half*)dz, (long long)ld_dz, gdt);
      break;
    case WEALY_BF16:
      loss_jacobian_kernel<__nv_bfloat16><<<blocks, T, 0, s>>>(cfg->kind, jeps, (const __nv_bfloat16*)zrow, (long long)ldz, (int)nb,
                                                               (int)d, w.u.norm + row0, w.du, w.scal, grad_out,
                                                               (__nv_bfloat16*)dz, (long long)ld_dz, gdt);
      break;
    default: return fail(WEALY_ERR_BAD_ARG, "unknown element type %d", dtype);
  }
  CU_TRY(cudaGetLastError());
  return WEALY_OK;
}

// ---- single GPU: the whole batch is one shard
extern "C" int wealy_loss_forward(const wealy_loss_cfg* cfg, const void* z, int64_t b, int64_t ldz, int64_t d, int dtype,
                                  int64_t* z_label, const int64_t* z_idx, double* out, void* out_cast, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  if (!out) return fail(WEALY_ERR_BAD_ARG, "null pointer");
  W_TRY(loss_forward_local(cfg, z, b, ldz, d, dtype, z_label, z_idx, 0, b, workspace, workspace_bytes, (cudaStream_t)stream));
  return loss_forward_finish(cfg, b, d, b, out, out_cast, dtype, workspace, (cudaStream_t)stream);
}

extern "C" int wealy_loss_backward(const wealy_loss_cfg* cfg, const void* z, int64_t b, int64_t ldz, int64_t d, int dtype,
                                   const void* grad_out, void* dz, int64_t ld_dz, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  return loss_backward_rows(cfg, z, b, ldz, d, dtype, 0, b, grad_out, dz, ld_dz, workspace, workspace_bytes,
                            (cudaStream_t)stream);
}

// ---- data parallel: z / labels / ids are the GLOBAL batch (all-gathered by the caller), this
// rank owns the anchors [row0, row0 + nb).  Between _local and _finish the caller sums `acc` (count_acc doubles, SUM)
// and `acc_max` (2 x uint32, MAX) over the ranks and all-gathers `rowstat` (4 floats per anchor, rank r's rows at
// [row0_r, row0_r + nb_r)): see wealy_loss_dp_buffers.
extern "C" int wealy_loss_dp_forward_local(const wealy_loss_cfg* cfg, const void* z, int64_t b_global, int64_t ldz, int64_t d,
                                           int dtype, const int64_t* z_label, const int64_t* z_idx, int64_t row0, int64_t nb,
                                           void* workspace, size_t workspace_bytes, void* stream) {
  if (cfg && cfg->label_noise) return fail(WEALY_ERR_BAD_ARG, "label noise acts on the GLOBAL batch: apply it before the call");
  return loss_forward_local(cfg, z, b_global, ldz, d, dtype, const_cast<int64_t*>(z_label), z_idx, row0, nb, workspace,
                            workspace_bytes, (cudaStream_t)stream);
}

// The same in two phases, so that the rank's own column block is swept while the all-gather of the other ranks' rows is
// still in flight: phase 1 reads only rows [row0, row0 + nb) of z, phase 2 everything else (call it once z is complete).
extern "C" int wealy_loss_dp_forward_phase(const wealy_loss_cfg* cfg, const void* z, int64_t b_global, int64_t ldz, int64_t d,
                                           int dtype, const int64_t* z_label, const int64_t* z_idx, int64_t row0, int64_t nb,
                                           int phase, void* workspace, size_t workspace_bytes, void* stream) {
  if (phase != 1 && phase != 2) return fail(WEALY_ERR_BAD_ARG, "phase must be 1 or 2");
  if (cfg && cfg->label_noise) return fail(WEALY_ERR_BAD_ARG, "label noise acts on the GLOBAL batch: apply it before the call");
  return loss_forward_local(cfg, z, b_global, ldz, d, dtype, const_cast<int64_t*>(z_label), z_idx, row0, nb, workspace,
                            workspace_bytes, (cudaStream_t)stream, phase);
}

extern "C" int wealy_loss_dp_buffers(const wealy_loss_cfg* cfg, void* workspace, size_t workspace_bytes, int64_t b_global,
                                     int64_t d, int64_t nb, double** acc, int64_t* count_acc, uint32_t** acc_max,
                                     float** rowstat) {
  if (!cfg || !workspace || !acc || !count_acc || !acc_max || !rowstat) return fail(WEALY_ERR_BAD_ARG, "null pointer");
  if (workspace_bytes < loss_ws_bytes(b_global, d, cfg->passes, nb)) return fail(WEALY_ERR_WORKSPACE, "workspace too small");
  LossWs w;
  loss_ws_layout(w, reinterpret_cast<uint8_t*>(align_up((size_t)workspace, 1024)), b_global, d, cfg->passes, nb);
  *acc = w.acc;
  *count_acc = kAccCount;
  *acc_max = w.acc_max;
  *rowstat = w.rowstat;
  return WEALY_OK;
}

// Exchange 2 as a single collective (equal shards: rank r owns anchors [r nb, (r + 1) nb)): pack this rank's record,
// all-gather the records (wealy_loss_dp_record_bytes(nb) bytes each, rank order), unpack them on every rank.
extern "C" size_t wealy_loss_dp_record_bytes(int64_t nb) { return nb < 0 ? 0 : (size_t)kDpRecordHeader + (size_t)nb * 16; }

extern "C" int wealy_loss_dp_pack(const wealy_loss_cfg* cfg, void* workspace, size_t workspace_bytes, int64_t b_global, int64_t d,
                                  int64_t row0, int64_t nb, void* record, void* stream) {
  if (!cfg || !workspace || !record) return fail(WEALY_ERR_BAD_ARG, "null pointer");
  if (nb <= 0 || row0 < 0 || row0 + nb > b_global) return fail(WEALY_ERR_BAD_ARG, "bad shard");
  if (workspace_bytes < loss_ws_bytes(b_global, d, cfg->passes, nb)) return fail(WEALY_ERR_WORKSPACE, "workspace too small");
  LossWs w;
  loss_ws_layout(w, reinterpret_cast<uint8_t*>(align_up((size_t)workspace, 1024)), b_global, d, cfg->passes, nb);
  const int n = (int)(nb > kAccCount ? nb : kAccCount);
  loss_dp_pack_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(w.acc, w.acc_max, w.rowstat + row0 * 4, (int)nb,
                                                                                  reinterpret_cast<unsigned char*>(record));
  CU_TRY(cudaGetLastError());
  return WEALY_OK;
}

extern "C" int wealy_loss_dp_unpack(const wealy_loss_cfg* cfg, void* workspace, size_t workspace_bytes, int64_t b_global, int64_t d,
                                    int64_t nb, const void* records, int world, void* stream) {
  if (!cfg || !workspace || !records) return fail(WEALY_ERR_BAD_ARG, "null pointer");
  if (world < 1 || nb <= 0 || (int64_t)world * nb != b_global) return fail(WEALY_ERR_BAD_ARG, "equal shards: world * nb must be the global batch");
  if (workspace_bytes < loss_ws_bytes(b_global, d, cfg->passes, nb)) return fail(WEALY_ERR_WORKSPACE, "workspace too small");
  LossWs w;
  loss_ws_layout(w, reinterpret_cast<uint8_t*>(align_up((size_t)workspace, 1024)), b_global, d, cfg->passes, nb);
  const int n = (int)(b_global > kAccCount ? b_global : kAccCount);
  loss_dp_unpack_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const unsigned char*>(records), (long long)wealy_loss_dp_record_bytes(nb), world, (int)nb, w.acc, w.acc_max,
      w.rowstat);
  CU_TRY(cudaGetLastError());
  return WEALY_OK;
}

extern "C" int wealy_loss_dp_forward_finish(const wealy_loss_cfg* cfg, int64_t b_global, int64_t d, int64_t nb, double* out,
                                            void* out_cast, int cast_dtype, void* workspace, size_t workspace_bytes,
                                            void* stream) {
  if (!cfg || !out || !workspace) return fail(WEALY_ERR_BAD_ARG, "null pointer");
  if (workspace_bytes < loss_ws_bytes(b_global, d, cfg->passes, nb)) return fail(WEALY_ERR_WORKSPACE, "workspace too small");
  return loss_forward_finish(cfg, b_global, d, nb, out, out_cast, cast_dtype, workspace, (cudaStream_t)stream);
}

extern "C" int wealy_loss_dp_backward(const wealy_loss_cfg* cfg, const void* z, int64_t b_global, int64_t ldz, int64_t d,
                                      int dtype, int64_t row0, int64_t nb, const void* grad_out, void* dz_rows,
                                      int64_t ld_dz, void* workspace, size_t workspace_bytes, void* stream) {
  return loss_backward_rows(cfg, z, b_global, ldz, d, dtype, row0, nb, grad_out, dz_rows, ld_dz, workspace, workspace_bytes,
                            (cudaStream_t)stream);
}
