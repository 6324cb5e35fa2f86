// tiled.cuh — spatially tiled Farneback: ONE frame pair split into row strips over the GPUs of a node
// (BASELINE.json config 5: 7680x4320 over 8 B200).  Included by farneback.cu.
//
// Every rank (one process and one ofb_handle per GPU) holds full-size level buffers but computes only
// the level rows it owns, rank r -> [r*rpr, min((r+1)*rpr, h_l)), rpr = ceil(h_l / world):
//   * pyramid + PolyExp of the own rows are local (each rank has the source frames; the level image
//     is produced for the own rows +- poly_n, nothing is exchanged);
//   * the fused iteration kernel needs, besides the own rows, R0 and the flow of the 2m halo rows of the
//     blur window and R1 wherever the displacement points.  It reads them where they live: through
//     NVLink peer pointers into the neighbours' buffers (k_iter_v<..., TILED>, um_issue2_tiled) — the
//     halo exchange is the kernel's own loads, overlapped tile by tile with the arithmetic, there is
//     no staging copy and no NCCL call on the data path;
//   * the inter-level flow upsample reads the coarse rows it needs the same way.
// What has to be ordered across GPUs is "all ranks finished stage s" before anyone reads a neighbour's
// rows in stage s+1: a flag barrier in peer memory (k_tile_barrier: every rank stores its epoch into
// every peer's flag array and spins until all peers' epochs arrived, with a timeout) enqueued on the
// stream between stages — 4 per level, no host round trip.
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

// Halo pull: copies the level rows [lo, yb) and [ye, hi) of a row-major buffer from their owners into this
// rank's buffer (same offsets), 16 bytes per thread, so that the compute kernel that follows finds the
// blur halo (and a margin for the displacement) locally.  One bulk NVLink transfer per stage instead of
// a remote round trip per halo row inside the marching producers (measured: the rank below a boundary
// ran its iteration kernels 44 % slower without it).
struct PullSrc {
  const uint4* p[kMaxTileRanks];
};
__global__ void __launch_bounds__(256) k_tile_pull(uint4* __restrict__ dst, PullSrc src, int rpr, int world,
                                                   int row_vec, int lo, int yb, int ye, int hi, size_t frame_vec) {
  // grid.y enumerates halo rows: first the rows above [lo, yb), then the rows below [ye, hi)
  int r = lo + blockIdx.y;
  if (r >= yb) r = ye + (r - yb);
  if (r >= hi) return;
  const int owner = min(r / rpr, world - 1);
  const size_t off = (size_t)blockIdx.z * frame_vec + (size_t)r * row_vec;
  const uint4* s = src.p[owner] + off;
  uint4* d = dst + off;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < row_vec; i += gridDim.x * blockDim.x) d[i] = s[i];
}

// Inter-level flow upsample of the own rows [y_begin, y_end); the coarse rows come from their owners.
__global__ void __launch_bounds__(256) k_upsample_flow_tiled(PeerTab prev, int pw, int ph, float2* __restrict__ out,
                                                             int w, int h, const LinTab* __restrict__ tabx,
                                                             const LinTab* __restrict__ taby, float mul, int y_begin,
                                                             int y_end) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int yb = y_begin + (blockIdx.y * blockDim.y + threadIdx.y) * UPS_ROWS;
  if (x >= w || yb >= y_end) return;
  const LinTab tx = tabx[x];
  const int x0 = tx.i0, x1 = min(x0 + 1, pw - 1);
  const float fx = tx.f;
#pragma unroll
  for (int j = 0; j < UPS_ROWS; j++) {
    const int y = yb + j;
    if (y >= y_end) break;
    const LinTab ty = taby[y];
    const int r0i = ty.i0, r1i = min(ty.i0 + 1, ph - 1);
    const float2* r0 = prev.flow[tile_owner(r0i, prev)] + (size_t)r0i * pw;
    const float2* r1 = prev.flow[tile_owner(r1i, prev)] + (size_t)r1i * pw;
    const float2 q00 = __ldg(r0 + x0), q01 = __ldg(r0 + x1), q10 = __ldg(r1 + x0), q11 = __ldg(r1 + x1);
    out[(size_t)y * w + x] = ups_blend(q00, q01, q10, q11, fx, ty.f, mul);
  }
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
  int cur_idx[kMaxLevels];   // ping-pong buffer that holds the level's initial flow (same on every rank)
  int res_idx[kMaxLevels];   // buffer that holds the level's result
};

static inline void tile_rows(int hh, int world, int rank, int* rpr, int* yb, int* ye) {
  *rpr = (hh + world - 1) / world;
  *yb = std::min(rank * *rpr, hh);
  *ye = std::min(*yb + *rpr, hh);
}

static int tiled_pull(ofb_handle* h, void* dst, void* const* peers, int w, int hh, int elem_bytes, int frames,
                      int halo, int* lo_out, int* hi_out);

// One stage of one rank.  kind 0 = flow init / upsample + pyramid + PolyExp of level li; kind 1 = iteration `it`.
static int tiled_stage(ofb_handle* h, const TiledPlan& pl, int li, int kind, int it) {
  const int world = h->tile.world, rank = h->tile.rank;
  const Level& lv = pl.sched[li];
  const int w = lv.width, hh = lv.height;
  const bool last_level = li == pl.n_levels - 1;
  int rpr, yb, ye;
  tile_rows(hh, world, rank, &rpr, &yb, &ye);
  cudaStream_t st = h->stream;
  const dim3 blk(32, 8);
  const int ci = pl.cur_idx[li];
  float2* cur = h->d_flow[ci];
  const ofb_farneback_params* p = pl.p;

  if (kind == 0) {
    // ---- initial flow of the level (own rows)
    TB(OFB_STAGE_FLOW_INIT);
    if (ye > yb) {
      if (li == 0) {
        OFB_CUDA(h, cudaMemsetAsync(cur + (size_t)yb * w, 0, (size_t)(ye - yb) * w * sizeof(float2), st));
      } else {
        // the previous level ended in buffer prev_idx on every rank
        const Level& pv = pl.sched[li - 1];
        const int prev_idx = pl.res_idx[li - 1];
        PeerTab t;
        memset(&t, 0, sizeof(t));
        for (int r = 0; r < world; r++) t.flow[r] = (const float2*)h->tile.peer_flow[prev_idx][r];
        t.rpr = (pv.height + world - 1) / world;
        t.world = world;
        dim3 g((w + blk.x - 1) / blk.x, ((ye - yb + UPS_ROWS - 1) / UPS_ROWS + blk.y - 1) / blk.y);
        k_upsample_flow_tiled<<<g, blk, 0, st>>>(t, pv.width, pv.height, cur, w, hh, h->d_lintab + h->tab_x_off[li],
                                                 h->d_lintab + h->tab_y_off[li], (float)(1.0 / p->pyr_scale), yb, ye);
        OFB_LAUNCH_CHECK(h);
      }
    }
    TE();
    if (ye <= yb) return OFB_OK;
    // ---- pyramid + PolyExp of the own rows (local)
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
    const bool fused_src = w == pl.width && hh == pl.height && pyc.r == 1;
    if (!fused_src) {
      const int lb = std::max(yb - pl.pc.n, 0), le = std::min(ye + pl.pc.n, hh);   // level rows PolyExp reads
      const double sy = 1.0 / ((double)hh / pl.height);
      const int sb = std::max(lin_entry(lb, sy, pl.height).i0 - pyc.r - 1, 0);
      const int se = std::min(lin_entry(le - 1, sy, pl.height).i0 + pyc.r + 3, pl.height);
      float* hb = reinterpret_cast<float*>(h->d_MA);
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
          k_polyexp_march<5, 1><<<g, PX_COLS, 0, st>>>(nullptr, src, pyc.k[0], pyc.k[1], h->d_RA, h->d_RB, w, hh, seg_rows,
                                                       strips, pl.pc, yb, ye);
        else
          k_polyexp_march<0, 1><<<g, PX_COLS, 0, st>>>(nullptr, src, pyc.k[0], pyc.k[1], h->d_RA, h->d_RB, w, hh, seg_rows,
                                                       strips, pl.pc, yb, ye);
      } else {
        if (pl.pc.n == 5)
          k_polyexp_march<5, 0><<<g, PX_COLS, 0, st>>>(h->d_img, src, 0.f, 0.f, h->d_RA, h->d_RB, w, hh, seg_rows, strips,
                                                       pl.pc, yb, ye);
        else
          k_polyexp_march<0, 0><<<g, PX_COLS, 0, st>>>(h->d_img, src, 0.f, 0.f, h->d_RA, h->d_RB, w, hh, seg_rows, strips,
                                                       pl.pc, yb, ye);
      }
      OFB_LAUNCH_CHECK(h);
    }
    TE();
    return OFB_OK;
  }

  // ---- iteration `it` of level li: own rows, neighbours' rows through the peer table
  if (ye <= yb) return OFB_OK;
  const bool last_it = it == p->iterations - 1;
  const int fin_idx = ci ^ (it & 1);
  const float2* fin = h->d_flow[fin_idx];
  float2* fout = (last_level && last_it) ? (float2*)pl.d_flow_out : h->d_flow[fin_idx ^ 1];
  (void)cur;
  PeerTab t;
  memset(&t, 0, sizeof(t));
  for (int r = 0; r < world; r++) {
    t.RA[r] = (const float4*)h->tile.peer_RA[r];
    t.RB[r] = (const float*)h->tile.peer_RB[r];
    t.flow[r] = (const float2*)h->tile.peer_flow[fin_idx][r];
  }
  t.rpr = rpr;
  t.world = world;
  // halo pulls (the barrier before this stage guarantees the neighbours' rows are final)
  TB(OFB_STAGE_OTHER);
  {
    int st2;
    const int halo_f = pl.bc.m + 2;                 // blur halo + the row the producers prefetch
    const int halo_r = pl.bc.m + kTileGatherMargin;  // blur halo + margin for the displacement of the gather
    if ((st2 = tiled_pull(h, h->d_flow[fin_idx], h->tile.peer_flow[fin_idx], w, hh, 8, 1, halo_f, &t.f_lo, &t.f_hi))) return st2;
    if (it == 0) {
      int a, b;
      if ((st2 = tiled_pull(h, h->d_RA, h->tile.peer_RA, w, hh, 16, 2, halo_r, &t.r_lo, &t.r_hi))) return st2;
      if ((st2 = tiled_pull(h, h->d_RB, h->tile.peer_RB, w, hh, 4, 2, halo_r, &a, &b))) return st2;
      t.r_lo = std::max(t.r_lo, a);                 // both R arrays must be local for a row to count as local
      t.r_hi = std::min(t.r_hi, b);
      h->tile.r_lo = t.r_lo;
      h->tile.r_hi = t.r_hi;
    } else {
      t.r_lo = h->tile.r_lo;
      t.r_hi = h->tile.r_hi;
    }
  }
  TE();
  const float reg = (float)(1e-3 / ((double)pl.bc.scale * (double)pl.bc.scale));
  cudaError_t e;
  // (the tiled kernel takes its R pointers per owner rank from the peer table; own buffers for the prefetch addresses)
  const RSet rs1 = {h->d_RA, h->d_RB, h->d_RA + (size_t)w * hh, h->d_RB + (size_t)w * hh};
  TB(OFB_STAGE_ITERATION);
  if (pl.bc.m == 7)
    e = launch_iter_v<7, 256, 2, 2, 3, true, false, false, 2>(h, fin, fout, w, hh, 1, rs1, pl.bc.m, reg, st, nullptr, yb, ye, &t, rank);
  else
    e = launch_iter_v<0, 128, 4, 1, 0, true, false, false, 2>(h, fin, fout, w, hh, 1, rs1, pl.bc.m, reg, st, nullptr, yb, ye, &t, rank);
  if (e != cudaSuccess) return set_error(h, OFB_ERR_CUDA, "tiled k_iter_v launch failed: %s", cudaGetErrorString(e));
  h->launches++;
  TE();
  return OFB_OK;
}

// Pull halo rows [yb - halo, yb) and [ye, ye + halo) of a buffer with `elem_bytes` per pixel and `frames`
// frames of hh x w pixels.  Rows must be a multiple of 16 bytes (else the halo stays remote: returns 0 rows).
static int tiled_pull(ofb_handle* h, void* dst, void* const* peers, int w, int hh, int elem_bytes, int frames,
                      int halo, int* lo_out, int* hi_out) {
  const int world = h->tile.world, rank = h->tile.rank;
  int rpr, yb, ye;
  tile_rows(hh, world, rank, &rpr, &yb, &ye);
  *lo_out = yb;
  *hi_out = ye;
  const size_t row_bytes = (size_t)w * elem_bytes;
  if (world == 1 || ye <= yb || (row_bytes & 15) || (((size_t)hh * row_bytes) & 15)) return OFB_OK;
  const int lo = std::max(yb - halo, 0), hi = std::min(ye + halo, hh);
  const int nrows = (yb - lo) + (hi - ye);
  if (nrows <= 0) return OFB_OK;
  PullSrc ps;
  for (int r = 0; r < kMaxTileRanks; r++) ps.p[r] = r < world ? (const uint4*)peers[r] : nullptr;
  const int row_vec = (int)(row_bytes / 16);
  dim3 g(std::min((row_vec + 255) / 256, 8), nrows, frames);
  k_tile_pull<<<g, 256, 0, h->stream>>>((uint4*)dst, ps, rpr, world, row_vec, lo, yb, ye, hi,
                                        (size_t)hh * row_bytes / 16);
  OFB_LAUNCH_CHECK(h);
  *lo_out = lo;
  *hi_out = hi;
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
  // ping-pong roles: a level's initial flow goes to the buffer that does NOT hold the previous result
  for (int li = 0; li < pl->n_levels; li++) {
    pl->cur_idx[li] = li == 0 ? 0 : (pl->res_idx[li - 1] ^ 1);
    pl->res_idx[li] = pl->cur_idx[li] ^ (p->iterations & 1);
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

// Real multi-GPU run of this rank: stages with the peer-memory flag barrier in between.
int farneback_run_tiled(ofb_handle* h, const uint8_t* d_prev, const uint8_t* d_next, int width, int height,
                        size_t pitch, float* d_flow_out, const ofb_farneback_params* p, int* row_begin,
                        int* row_end) {
  TiledPlan pl;
  int st = tiled_make_plan(h, &pl, d_prev, d_next, width, height, pitch, d_flow_out, p);
  if (st) return st;
  // entry barrier: nobody overwrites buffers a peer may still read from the previous call
  st = tiled_barrier(h);
  if (st) return st;
  for (int li = 0; li < pl.n_levels; li++) {
    st = tiled_stage(h, pl, li, 0, 0);
    if (st) return st;
    st = tiled_barrier(h);
    if (st) return st;
    for (int it = 0; it < p->iterations; it++) {
      st = tiled_stage(h, pl, li, 1, it);
      if (st) return st;
      if (!(li == pl.n_levels - 1 && it == p->iterations - 1)) {
        st = tiled_barrier(h);
        if (st) return st;
      }
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
    // (each rank's stage runs alone: the streams are synchronised after every rank, so per-rank stage
    //  timers are clean and no two ranks' kernels share the device)
    for (int r = 0; r < world; r++) {
      if ((st = tiled_stage(hs[r], pls[r], li, 0, 0))) return st;
      if ((st = sync_all())) return st;
    }
    for (int it = 0; it < p->iterations; it++) {
      for (int r = 0; r < world; r++) {
        if ((st = tiled_stage(hs[r], pls[r], li, 1, it))) return st;
        if ((st = sync_all())) return st;
      }
    }
  }
  return OFB_OK;
}

}  // namespace ofb
