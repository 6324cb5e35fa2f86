// tiled.cuh — spatially tiled Farneback: ONE frame pair split into row strips over the GPUs of a node
// (BASELINE.json config 5: 7680x4320 over 8 B200).  Included by farneback.cu.
//
// Every rank (one process and one ofb_handle per GPU) produces the rows of the final field it owns,
// rank r -> [r*rpr, min((r+1)*rpr, H)), rpr = ceil(H / world), and works on INDEPENDENT BANDS:
//   * what a rank needs from its neighbours at a level is a halo of the input flow (the blur window, m rows per
//     iteration) and of the expansions.  Instead of exchanging halos every iteration (round 1: 17 cross-GPU barriers
//     and 15 halo pulls per pair, 0.78 of 1.24 ms on 8 GPUs) a rank RECOMPUTES them: the band it computes at a
//     level is widened so that after `iterations` shrinking steps of m rows the rows the next finer level needs are
//     still valid (need_l, derived top-down from the rows the rank owns at level 0).  Every rank holds the two source
//     frames, so pyramid and PolyExp of the wider band are local.  Redundant work at 8K over 8 GPUs: +8 % at level
//     0, +24 % at level 1, more at the two small levels — about +15 % in all — for NO flow exchange at all;
//   * the inter-level upsample is fused into the first iteration of a level (UpsSrc), reading the rank's own coarse band;
//   * the only cross-rank access left is the bilinear gather of R1 where the displacement points more than
//     kTileGatherMargin rows outside the rank's band: a peer-pointer load over NVLink from the rank that owns the row.
//     It needs that rank's PolyExp of the level to be complete: ONE flag barrier in peer memory per level
//     (k_tile_barrier, enqueued on the stream — no host round trip, no NCCL on the data path).  The expansions of
//     every level have their own buffer, so no other ordering is needed, not even between consecutive pairs.
//
// With fewer GPUs than ranks (tests on one GPU) the ranks are emulated in ONE process: the stages of
// all ranks are launched in order on one device with a device synchronisation in between, and no
// barrier kernel (kernels that spin on each other must not share a GPU).
#pragma once

namespace ofb {

constexpr int kTileGatherMargin = 24;   // rows of R1 pulled beyond the blur halo (larger displacements read remotely)

struct PeerFlags {
  unsigned* p[kMaxTileRanks];
};

// Each of the first `world` threads publishes this rank's epoch to one peer and waits for that peer's.
__global__ void k_tile_barrier(volatile unsigned* my_flags, PeerFlags peers, int rank, int world, unsigned epoch,
                               int* err, long long timeout_cycles) {
  const int i = threadIdx.x;
  if (i >= world) return;
  __threadfence_system();                       // the stage's results are visible before the flag
  volatile unsigned* dst = peers.p[i] + rank;   // peer i's flags[rank]
  *dst = epoch;
  __threadfence_system();
  const long long t0 = clock64();
  while ((int)(my_flags[i] - epoch) < 0) {
    if (clock64() - t0 > timeout_cycles) {
      *err = 1;
      break;
    }
  }
  __threadfence_system();
}

struct TiledPlan {
  Level sched[kMaxLevels];
  int n_levels;
  PolyCoef pc;
  BlurCoef bc;
  int width, height;
  size_t pitch;
  const uint8_t* d_prev;
  const uint8_t* d_next;
  float* d_flow_out;
  const ofb_farneback_params* p;
  // per level (coarse -> fine): rows whose final flow of the level must be valid on this rank, rows of the level's
  // input flow / matrices, rows of the expansions; element offset of the level's expansions in the R buffers
  int need_lo[kMaxLevels], need_hi[kMaxLevels];
  int in_lo[kMaxLevels], in_hi[kMaxLevels];
  int r_lo[kMaxLevels], r_hi[kMaxLevels];
  size_t r_off[kMaxLevels];
  int res_idx[kMaxLevels];   // ping-pong buffer that holds the level's result
  // the three regular levels in one pass over the source (k_pyr_fast3): level index of the S = 8 level (-1: per level),
  // rows of that level to run, the level images' offsets in d_img (S = 2, 4, 8)
  int fused3_li, y3_lo, y3_hi;
  size_t img3_off[3];
  PyrFast3Coef fc3;
};

static inline void tile_rows(int hh, int world, int rank, int* rpr, int* yb, int* ye) {
  *rpr = (hh + world - 1) / world;
  *yb = std::min(rank * *rpr, hh);
  *ye = std::min(*yb + *rpr, hh);
}

// Expansions of level li of rank r: level 0 (the last of the schedule) lives in d_RA / d_RB, the coarser levels one
// after another in d_MA / d_MB (the generic path's buffers, unused in tiled mode).
static inline const float4* tiled_RA(const ofb_handle* h, const TiledPlan& pl, int li, int r) {
  return li == pl.n_levels - 1 ? (const float4*)h->tile.peer_RA[r] : (const float4*)h->tile.peer_MA[r] + pl.r_off[li];
}
static inline const float* tiled_RB(const ofb_handle* h, const TiledPlan& pl, int li, int r) {
  return li == pl.n_levels - 1 ? (const float*)h->tile.peer_RB[r] : (const float*)h->tile.peer_MB[r] + pl.r_off[li];
}

// One stage of one rank.  kind 0 = pyramid + PolyExp of the band of level li; kind 1 = iteration `it`; kind 2 = the
// pyramid part of kind 0 only, kind 3 = its PolyExp part only, on stream `px_stream` (the two-stream schedule of
// farneback_run_tiled).
static int tiled_stage(ofb_handle* h, const TiledPlan& pl, int li, int kind, int it, cudaStream_t px_stream = nullptr) {
  const int world = h->tile.world, rank = h->tile.rank;
  const Level& lv = pl.sched[li];
  const int w = lv.width, hh = lv.height;
  const bool last_level = li == pl.n_levels - 1;
  cudaStream_t st = h->stream;
  const ofb_farneback_params* p = pl.p;
  if (pl.need_hi[li] <= pl.need_lo[li]) return OFB_OK;      // (more ranks than rows: nothing to do at any level)
  float4* const RA = const_cast<float4*>(tiled_RA(h, pl, li, rank));
  float* const RB = const_cast<float*>(tiled_RB(h, pl, li, rank));

  if (kind == 0 || kind == 2 || kind == 3) {
    const bool do_pyr = kind != 3, do_px = kind != 2;
    cudaStream_t sp = (kind == 3 && px_stream) ? px_stream : st;
    const int yb = pl.r_lo[li], ye = pl.r_hi[li];
    // ---- pyramid + PolyExp of the band (local: every rank has the source frames)
    PyrCoef pyc;
    if (prepare_pyr(lv.ksize, lv.sigma, &pyc) != OFB_OK)
      return set_error(h, OFB_ERR_INVALID_ARG, "pyramid smoothing kernel too large (ksize=%d)", lv.ksize);
    FrameSrc src;
    src.a = pl.d_prev;
    src.b = pl.d_next;
    src.na = 1;
    src.pitch = pl.pitch;
    src.image_stride = 0;
    const int frames = 2;
    const bool fused_src = w == pl.width && hh == pl.height && pyc.r == 1 && pyc.k[0] == 0.5f && pyc.k[1] == 0.25f;
    const float* level_img = h->d_img;
    const bool in_fused3 = pl.fused3_li >= 0 && li >= pl.fused3_li && li < pl.n_levels - 1;
    if (in_fused3) {
      level_img = h->d_img + pl.img3_off[pl.n_levels - 2 - li];
      if (li == pl.fused3_li && do_pyr) {
        TB(OFB_STAGE_PYRAMID);
        constexpr int out3 = (PF_COLS - 2 * PF_HALO) / 8;
        const int chunks = (pl.width / 8 + out3 - 1) / out3, rows = pl.y3_hi - pl.y3_lo;
        const int segs = std::max(1, 2 * 3 * h->num_sms / (chunks * frames));
        const int seg_rows = std::max(2, (rows + segs - 1) / segs);
        dim3 g(chunks, (rows + seg_rows - 1) / seg_rows, frames);
        k_pyr_fast3<<<g, PF_THREADS, 0, st>>>(src, pl.width, pl.height, h->d_img + pl.img3_off[0], h->d_img + pl.img3_off[1],
                                              h->d_img + pl.img3_off[2], pl.fc3, seg_rows, pl.y3_lo, pl.y3_hi);
        OFB_LAUNCH_CHECK(h);
        TE();
      }
    } else if (!fused_src && do_pyr) {
      const int lb = std::max(yb - pl.pc.n, 0), le = std::min(ye + pl.pc.n, hh);   // level rows PolyExp reads
      const double sy = 1.0 / ((double)hh / pl.height);
      const int sb = std::max(lin_entry(lb, sy, pl.height).i0 - pyc.r - 1, 0);
      const int se = std::min(lin_entry(le - 1, sy, pl.height).i0 + pyc.r + 3, pl.height);
      float* hb = reinterpret_cast<float*>(h->d_VA);
      dim3 gh((w + 127) / 128, (se - sb + PYR_RPT - 1) / PYR_RPT, frames);
      dim3 bv(128, 2), gv((w + 127) / 128, (le - lb + 1) / 2, frames);
      TB(OFB_STAGE_PYRAMID);
#define OFB_PYR_LAUNCH(RT)                                                                                        \
  do {                                                                                                            \
    k_pyr_h<RT><<<gh, 128, 0, st>>>(src, pl.width, pl.height, hb, w, 1.0 / ((double)w / pl.width), pyc, sb, se); \
    OFB_LAUNCH_CHECK(h);                                                                                          \
    k_pyr_v<RT><<<gv, bv, 0, st>>>(hb, pl.height, h->d_img, w, hh, sy, pyc, lb, le);                              \
    OFB_LAUNCH_CHECK(h);                                                                                          \
  } while (0)
      if (pyc.r == 1) OFB_PYR_LAUNCH(1);
      else if (pyc.r == 4) OFB_PYR_LAUNCH(4);
      else if (pyc.r == 9) OFB_PYR_LAUNCH(9);
      else OFB_PYR_LAUNCH(0);
#undef OFB_PYR_LAUNCH
      TE();
    }
    if (!do_px) return OFB_OK;
    TB(OFB_STAGE_POLYEXP);
    {
      const int strips = (w + PX_TW - 1) / PX_TW;
      const int per = strips * frames;
      const int slots = 3 * h->num_sms * kPxWaves;
      int segs = std::max(1, slots / per);
      int seg_rows = std::max(16, ((ye - yb + segs - 1) / segs + PX_ROWS - 1) / PX_ROWS * PX_ROWS);
      segs = (ye - yb + seg_rows - 1) / seg_rows;
      dim3 g(strips * segs, frames);
      if (fused_src) {
        if (pl.pc.n == 5)
          k_polyexp_march<5, 1><<<g, PX_COLS, 0, sp>>>(nullptr, src, pyc.k[0], pyc.k[1], RA, RB, w, hh, seg_rows, strips,
                                                       pl.pc, yb, ye);
        else
          k_polyexp_march<0, 1><<<g, PX_COLS, 0, sp>>>(nullptr, src, pyc.k[0], pyc.k[1], RA, RB, w, hh, seg_rows, strips,
                                                       pl.pc, yb, ye);
      } else {
        if (pl.pc.n == 5)
          k_polyexp_march<5, 0><<<g, PX_COLS, 0, sp>>>(level_img, src, 0.f, 0.f, RA, RB, w, hh, seg_rows, strips, pl.pc, yb, ye);
        else
          k_polyexp_march<0, 0><<<g, PX_COLS, 0, sp>>>(level_img, src, 0.f, 0.f, RA, RB, w, hh, seg_rows, strips, pl.pc, yb, ye);
      }
      OFB_LAUNCH_CHECK(h);
    }
    TE();
    return OFB_OK;
  }

  // ---- iteration `it` of level li on the rows need +- (iterations - 1 - it) * m (all inputs local; R1 gathers beyond the
  // band go to the owner's buffer through the peer table)
  const int n_it = p->iterations, m = pl.bc.m;
  int yb = std::max(pl.need_lo[li] - (n_it - 1 - it) * m, 0);
  const int ye = std::min(pl.need_hi[li] + (n_it - 1 - it) * m, hh);
  const bool last_it = it == n_it - 1;
  // the fused upsample's interior form starts its chunks on odd matrix rows (even band start, odd radius): one more row
  // at the top where needed — local scratch rows (never the final field: a level's first iteration is not its last
  // unless iterations == 1, where the owned rows start at a multiple of the even rows-per-rank or this is skipped)
  if (it == 0 && li > 0 && !(last_level && last_it) && (yb & 1) && yb > 0) yb -= 1;
  // flow ping-pong as in the whole-frame driver: the first iteration of a level reads the coarser level's result (fused
  // upsample) and must not write the buffer that holds it
  const int prev_idx = li > 0 ? pl.res_idx[li - 1] : 1;
  const int first_out = prev_idx ^ 1;
  const int out_idx = first_out ^ (it & 1);
  const float2* fin = h->d_flow[out_idx ^ 1];
  float2* fout = (last_level && last_it) ? (float2*)pl.d_flow_out : h->d_flow[out_idx];
  UpsSrc ups;
  const UpsSrc* up = nullptr;
  if (it == 0) {
    if (li == 0) {
      TB(OFB_STAGE_FLOW_INIT);
      OFB_CUDA(h, cudaMemsetAsync(h->d_flow[out_idx ^ 1] + (size_t)pl.in_lo[li] * w, 0,
                                  (size_t)(pl.in_hi[li] - pl.in_lo[li]) * w * sizeof(float2), st));
      TE();
    } else {
      ups.prev = h->d_flow[prev_idx]; ups.pw = pl.sched[li - 1].width; ups.ph = pl.sched[li - 1].height;
      ups.tabx = h->d_lintab + h->tab_x_off[li]; ups.taby = h->d_lintab + h->tab_y_off[li];
      ups.mul = (float)(1.0 / p->pyr_scale);
      ups.exact2y = hh == 2 * ups.ph;
      up = &ups;
    }
  }
  PeerTab t;
  memset(&t, 0, sizeof(t));
  for (int r = 0; r < world; r++) {
    t.RA[r] = tiled_RA(h, pl, li, r);
    t.RB[r] = tiled_RB(h, pl, li, r);
  }
  t.rpr = (hh + world - 1) / world;
  t.world = world;
  t.r_lo = pl.r_lo[li];
  t.r_hi = pl.r_hi[li];
  const float reg = (float)(1e-3 / ((double)pl.bc.scale * (double)pl.bc.scale));
  cudaError_t e;
  // (own buffers for the prefetch addresses; the kernel takes remote rows from the peer table)
  const RSet rs1 = {RA, RB, RA + (size_t)w * hh, RB + (size_t)w * hh};
  TB(OFB_STAGE_ITERATION);
  if (pl.bc.m == 7) {
    // (the default schedule: two rows in flight, row-reuse gather, ring in tensor memory — as the whole-frame path)
    if (up) e = launch_iter_v<7, 256, 2, 2, 0, true, true, true, 4, true>(h, fin, fout, w, hh, 1, rs1, pl.bc.m, reg, st, up, yb, ye, &t, rank);
    else e = launch_iter_v<7, 256, 2, 2, 0, true, true, true, 4>(h, fin, fout, w, hh, 1, rs1, pl.bc.m, reg, st, nullptr, yb, ye, &t, rank);
  } else {
    if (up) e = launch_iter_v<0, 128, 4, 1, 0, true, false, false, 2, true>(h, fin, fout, w, hh, 1, rs1, pl.bc.m, reg, st, up, yb, ye, &t, rank);
    else e = launch_iter_v<0, 128, 4, 1, 0, true, false, false, 2>(h, fin, fout, w, hh, 1, rs1, pl.bc.m, reg, st, nullptr, yb, ye, &t, rank);
  }
  if (e != cudaSuccess) return set_error(h, OFB_ERR_CUDA, "tiled k_iter_v launch failed: %s", cudaGetErrorString(e));
  h->launches++;
  TE();
  return OFB_OK;
}

static int tiled_barrier(ofb_handle* h) {
  PeerFlags pf;
  for (int r = 0; r < kMaxTileRanks; r++) pf.p[r] = r < h->tile.world ? h->tile.peer_flags[r] : nullptr;
  h->tile.epoch++;
  // ~2 s at 1.9 GHz: a rank that died must not hang the others (and the GPU) forever.  The first barriers of a handle
  // get ~30 s: the peers may still be loading modules / setting function attributes.  A timeout sets d_err, which
  // ofb_synchronize / ofb_tiled_status report as an error.
  TB(OFB_STAGE_OTHER);
  k_tile_barrier<<<1, 32, 0, h->stream>>>(h->tile.d_flags, pf, h->tile.rank, h->tile.world, h->tile.epoch,
                                          h->tile.d_err, h->tile.epoch <= 2 ? 60000000000LL : 4000000000LL);
  OFB_LAUNCH_CHECK(h);
  TE();
  return OFB_OK;
}

static int tiled_make_plan(ofb_handle* h, TiledPlan* pl, const uint8_t* d_prev, const uint8_t* d_next, int width,
                           int height, size_t pitch, float* d_flow_out, const ofb_farneback_params* p) {
  if (build_schedule(width, height, p->pyr_scale, p->levels, pl->sched, &pl->n_levels) != OFB_OK)
    return set_error(h, OFB_ERR_INVALID_ARG, "too many pyramid levels");
  int st = ensure_lintabs(h, pl->sched, pl->n_levels, width, height, p->pyr_scale);
  if (st) return st;
  prepare_poly(p->poly_n, p->poly_sigma, &pl->pc);
  prepare_blur(p->winsize, false, &pl->bc);
  if (p->flags != 0) return set_error(h, OFB_ERR_INVALID_ARG, "tiled mode supports flags = 0 only");
  if (pl->bc.m < 2 || pl->bc.m > 19) return set_error(h, OFB_ERR_INVALID_ARG, "tiled mode needs winsize in [4, 39]");
  if (pl->pc.n > PX_MAXN) return set_error(h, OFB_ERR_INVALID_ARG, "tiled mode needs poly_n <= %d", PX_MAXN);
  if (p->iterations < 1) return set_error(h, OFB_ERR_INVALID_ARG, "tiled mode needs iterations >= 1");
  const int world = h->tile.world, rank = h->tile.rank, nl = pl->n_levels, halo = p->iterations * pl->bc.m;
  // bands, from the finest level (where the rank must deliver exactly the rows it owns) down to the coarsest:
  //   need_l : rows of the level's FINAL flow that must be valid here
  //   in_l   : rows of the level's input flow and of the matrices = need_l +- iterations * m
  //   coarser need = the rows the bilinear upsample of in_l touches (cv::resize coordinates, +-2 rows of slack)
  int rpr;
  tile_rows(height, world, rank, &rpr, &pl->need_lo[nl - 1], &pl->need_hi[nl - 1]);
  for (int li = nl - 1; li >= 0; li--) {
    const int hh = pl->sched[li].height;
    if (pl->need_hi[li] <= pl->need_lo[li]) {          // nothing owned: empty at every level
      pl->in_lo[li] = pl->in_hi[li] = pl->r_lo[li] = pl->r_hi[li] = 0;
      if (li > 0) pl->need_lo[li - 1] = pl->need_hi[li - 1] = 0;
      continue;
    }
    pl->in_lo[li] = std::max(pl->need_lo[li] - halo, 0);
    pl->in_hi[li] = std::min(pl->need_hi[li] + halo, hh);
    pl->r_lo[li] = std::max(pl->in_lo[li] - kTileGatherMargin, 0);
    pl->r_hi[li] = std::min(pl->in_hi[li] + kTileGatherMargin, hh);
    if (li > 0) {
      const int ph = pl->sched[li - 1].height;
      const double sy = 1.0 / ((double)hh / ph);
      pl->need_lo[li - 1] = std::max(lin_entry(pl->in_lo[li], sy, ph).i0 - 2, 0);
      pl->need_hi[li - 1] = std::min(lin_entry(pl->in_hi[li] - 1, sy, ph).i0 + 4, ph);
    }
  }
  // expansions of the coarser levels: one after another in d_MA / d_MB (two frames + spare rows each)
  size_t off = 0;
  for (int li = 0; li < nl - 1; li++) {
    pl->r_off[li] = off;
    off += 2 * (size_t)pl->sched[li].width * pl->sched[li].height + (size_t)kRowPad * pl->sched[li].width;
  }
  pl->r_off[nl - 1] = 0;
  if (off > (size_t)h->max_batch * h->max_w * h->max_h)
    return set_error(h, OFB_ERR_CAPACITY, "tiled mode: the coarser levels' expansions do not fit (pyr_scale too close to 1)");
  for (int li = 0; li < nl; li++) {
    const int prev_idx = li > 0 ? pl->res_idx[li - 1] : 1;
    pl->res_idx[li] = (prev_idx ^ 1) ^ ((p->iterations - 1) & 1);
  }
  pl->fused3_li = -1;
  if (OFB_EXP_PYR3 && pyr3_applicable(pl->sched, nl, width, height, pitch, d_prev, d_next, &pl->fc3) &&
      pl->need_hi[nl - 1] > pl->need_lo[nl - 1]) {
    pl->fused3_li = nl - 4;
    const size_t n1 = (size_t)(width / 2) * (height / 2);
    pl->img3_off[0] = 0;
    pl->img3_off[1] = 2 * n1;
    pl->img3_off[2] = 2 * n1 + 2 * (n1 / 4);
    int lo = height, hi = 0;
    for (int q = 0; q < 3; q++) {                        // S = 2, 4, 8: level rows PolyExp reads -> rows of the S = 8 level
      const int li = nl - 2 - q, hh = pl->sched[li].height, per = 4 >> q;
      const int lb = std::max(pl->r_lo[li] - pl->pc.n, 0), le = std::min(pl->r_hi[li] + pl->pc.n, hh);
      lo = std::min(lo, lb / per);
      hi = std::max(hi, (le + per - 1) / per);
    }
    pl->y3_lo = lo;
    pl->y3_hi = std::min(hi, height / 8);
  }
  pl->width = width;
  pl->height = height;
  pl->pitch = pitch;
  pl->d_prev = d_prev;
  pl->d_next = d_next;
  pl->d_flow_out = d_flow_out;
  pl->p = p;
  return OFB_OK;
}

int tiled_barrier_public(ofb_handle* h) { return tiled_barrier(h); }

// Real multi-GPU run of this rank: per level PolyExp of the band, ONE flag barrier (the peers' expansions of the level
// are complete: remote gathers may read them), then the iterations.
int farneback_run_tiled(ofb_handle* h, const uint8_t* d_prev, const uint8_t* d_next, int width, int height,
                        size_t pitch, float* d_flow_out, const ofb_farneback_params* p, int* row_begin,
                        int* row_end) {
  TiledPlan pl;
  int st = tiled_make_plan(h, &pl, d_prev, d_next, width, height, pitch, d_flow_out, p);
  if (st) return st;
  // Everybody has finished the previous pair: with the two-stream schedule below a rank writes the expansions of every
  // level right after its pyramid pass — a peer that is still iterating on the previous pair may be gathering from those
  // buffers.  (Taken on every call, whatever the schedule, so that the ranks' barrier epochs can never drift apart.)
  st = tiled_barrier(h);
  if (st) return st;
  // Two-stream schedule (as the whole-frame driver): all level images come from the one k_pyr_fast3 pass, every level
  // has expansion buffers of its own, and the coarse levels' iteration launches are a few CTAs marching a few rows — so
  // the expansions of ALL levels go to the expansion stream right behind the pyramid pass and run beside the coarse
  // iterations; the compute stream waits for a level's expansions, passes the level's flag barrier, iterates.
  const bool overlap = !h->no_overlap && !h->timing && h->s_px && pl.fused3_li == 0 && pl.n_levels == 4;
  if (overlap) {
    for (int li = 0; li < pl.n_levels; li++)
      if ((st = tiled_stage(h, pl, li, 2, 0))) return st;
    OFB_CUDA(h, cudaEventRecord(h->ev_fork, h->stream));
    OFB_CUDA(h, cudaStreamWaitEvent(h->s_px, h->ev_fork, 0));
    for (int li = 0; li < pl.n_levels; li++) {
      if ((st = tiled_stage(h, pl, li, 3, 0, h->s_px))) return st;
      OFB_CUDA(h, cudaEventRecord(h->ev_px[li], h->s_px));
    }
  }
  for (int li = 0; li < pl.n_levels; li++) {
    if (overlap) {
      OFB_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_px[li], 0));
    } else {
      st = tiled_stage(h, pl, li, 0, 0);
      if (st) return st;
    }
    st = tiled_barrier(h);
    if (st) return st;
    for (int it = 0; it < p->iterations; it++) {
      st = tiled_stage(h, pl, li, 1, it);
      if (st) return st;
    }
  }
  int rpr;
  tile_rows(height, h->tile.world, h->tile.rank, &rpr, row_begin, row_end);
  return OFB_OK;
}

// All ranks in one process on one device (tests): stage by stage, device-synchronised, no barrier kernel.
int farneback_run_tiled_emulated(ofb_handle* const* hs, int world, const uint8_t* d_prev, const uint8_t* d_next,
                                 int width, int height, size_t pitch, float* d_flow_out,
                                 const ofb_farneback_params* p) {
  std::vector<TiledPlan> pls(world);
  for (int r = 0; r < world; r++) {
    int st = tiled_make_plan(hs[r], &pls[r], d_prev, d_next, width, height, pitch, d_flow_out, p);
    if (st) return st;
  }
  auto sync_all = [&]() -> int {
    for (int r = 0; r < world; r++) OFB_CUDA(hs[r], cudaStreamSynchronize(hs[r]->stream));
    return OFB_OK;
  };
  int st = sync_all();
  if (st) return st;
  for (int li = 0; li < pls[0].n_levels; li++) {
    for (int r = 0; r < world; r++)
      if ((st = tiled_stage(hs[r], pls[r], li, 0, 0))) return st;
    if ((st = sync_all())) return st;                   // stands in for the per-level flag barrier
    for (int r = 0; r < world; r++)
      for (int it = 0; it < p->iterations; it++)
        if ((st = tiled_stage(hs[r], pls[r], li, 1, it))) return st;
    if ((st = sync_all())) return st;
  }
  return OFB_OK;
}

}  // namespace ofb
