// junction.cu — the reference's junction detector (ros2_ws/src/junction_point_detector/src/junction_detector.cpp:3-214,
// find_junctions_not_rotated + dampenIntensity) as the point source of the sparse path: gray -> 3x3 Gaussian -> adaptive
// Gaussian threshold -> contours (area, bounding box, tree order) -> box corners -> KD-tree clusters.  Restated and pinned
// in oracle/junction_np.py (pixel stages and contours against the cv2 wheel, clustering against the reference's own
// nanoflann header compiled into oracle/_ref).
//
// cv2.findContours is a serial border follower; here the same contours come out of two connected-component labellings:
// every border is the interface between one 8-connected foreground component F and one 4-connected background component H
// of the zero-padded binary image — the outer border of F, or the hole border around H.  A union-find labelling (label =
// the component's first pixel in raster order = the pixel where Suzuki's scan discovers the border) gives both; one
// thread per foreground pixel then adds, for each of its pixel edges towards the background ("cracks"), the shoelace term
// of the step to the next crack of the interface (8-connectivity at saddle points) and the pixel to the bounding box of
// the border the crack belongs to.  contourArea = |sum| / 2, boundingRect = the box, the parent of a border follows from
// which component encloses which, and cv2's output order is the pre-order of that tree with siblings in descending
// discovery order — rebuilt on the host for the (few) contours that pass the reference's size tests.  The final
// clustering (a few hundred points) runs on the host as nanoflann does it, approximate search included.
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"

namespace ofb {

namespace {

constexpr int kMaxDepth = 48;          // contour nesting levels carried to the host per passing contour
constexpr int kMaxPass = 65536;        // contours that may pass the size tests in one frame

struct PassRec {
  int x, y, w, h;                      // boundingRect
  int depth;                           // entries of path
  int path[kMaxDepth];                 // discovery keys from the top-level ancestor down to the contour itself
};

struct JunctionState {
  uint8_t *d_src = nullptr, *d_gray = nullptr, *d_blur = nullptr, *d_bin = nullptr;
  float* d_row = nullptr;
  int* d_lab = nullptr;
  long long* d_area = nullptr;
  int* d_box = nullptr;                // [4][n_pad]: min x, min y, max x, max y
  PassRec* d_pass = nullptr;
  int* d_cnt = nullptr;                // [0] passing contours, [1] depth overflow flag
  PassRec* h_pass = nullptr;           // pinned
  int* h_cnt = nullptr;                // pinned
  size_t src_cap = 0, px_cap = 0, pad_cap = 0;
};

// cv2.getGaussianKernel(11, 0, CV_32F) (sigma 2.0), taps 0..5 (symmetric)
__constant__ float c_g11[6] = {0x1.20c256p-7f, 0x1.bcb86ap-6f, 0x1.0ab50ap-4f, 0x1.f2464cp-4f, 0x1.6a7e1ep-3f, 0x1.9ac20ap-3f};

// dampenIntensity (junction_detector.cpp:3-28) fused with cvtColor(BGR2GRAY); channels == 1 copies
__global__ void __launch_bounds__(256) k_jd_gray(const uint8_t* __restrict__ src, size_t sp, int w, int h, int cn, int dampen,
                                                 double incline, double intercept, uint8_t* __restrict__ gray) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w) return;
  const uint8_t* p = src + (size_t)y * sp + (size_t)x * cn;
  if (cn == 1) { gray[(size_t)y * w + x] = p[0]; return; }
  int b = p[0], g = p[1], r = p[2];
  if (dampen) {
    double gain = __dadd_rn(__dmul_rn((double)(r - b), incline), intercept);
    gain = fmax(fmin(gain, 1.0), 0.0);
    b = (int)__dmul_rn((double)b, gain); g = (int)__dmul_rn((double)g, gain); r = (int)__dmul_rn((double)r, gain);
  }
  gray[(size_t)y * w + x] = (uint8_t)((b * 3735 + g * 19235 + r * 9798 + 16384) >> 15);
}

// GaussianBlur(gray, (3, 3), 0) on uint8: ([1 2 1] x [1 2 1] + 8) >> 4, BORDER_REFLECT_101
__global__ void __launch_bounds__(256) k_jd_blur3(const uint8_t* __restrict__ g, int w, int h, uint8_t* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w) return;
  const int xm = x > 0 ? x - 1 : (w > 1 ? 1 : 0), xp = x < w - 1 ? x + 1 : (w > 1 ? w - 2 : 0);
  const int ym = y > 0 ? y - 1 : (h > 1 ? 1 : 0), yp = y < h - 1 ? y + 1 : (h > 1 ? h - 2 : 0);
  const uint8_t *r0 = g + (size_t)ym * w, *r1 = g + (size_t)y * w, *r2 = g + (size_t)yp * w;
  const int s = (r0[xm] + 2 * r0[x] + r0[xp]) + 2 * (r1[xm] + 2 * r1[x] + r1[xp]) + (r2[xm] + 2 * r2[x] + r2[xp]);
  out[(size_t)y * w + x] = (uint8_t)((s + 8) >> 4);
}

// adaptiveThreshold's float Gaussian, row pass, in the wheel's operation order (oracle/junction_np.py::gauss11_f32):
// acc = k0 x0, then fma per tap; the last width % 4 columns: taps 1..8 multiply-then-add, taps 9 and 10 fused.
__global__ void __launch_bounds__(256) k_jd_gauss_row(const uint8_t* __restrict__ blur, int w, int h, float* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w) return;
  const uint8_t* r = blur + (size_t)y * w;
  float v[11];
#pragma unroll
  for (int i = 0; i < 11; i++) v[i] = (float)r[min(max(x + i - 5, 0), w - 1)];
  float acc = __fmul_rn(v[0], c_g11[0]);
  if (x < w - (w & 3)) {
#pragma unroll
    for (int i = 1; i < 11; i++) acc = __fmaf_rn(v[i], c_g11[i <= 5 ? i : 10 - i], acc);
  } else {
#pragma unroll
    for (int i = 1; i <= 8; i++) acc = __fadd_rn(acc, __fmul_rn(v[i], c_g11[i <= 5 ? i : 10 - i]));
    acc = __fmaf_rn(v[9], c_g11[1], acc);
    acc = __fmaf_rn(v[10], c_g11[0], acc);
  }
  out[(size_t)y * w + x] = acc;
}

// column pass (symmetric form; multiply-then-add in the last width % 8 columns), saturate_cast<uchar> of the mean
// (round half to even), THRESH_BINARY with delta 2: foreground where blur - mean > -2.  Writes the zero-padded binary
// image (pitch w + 2; the frame stays zero).
__global__ void __launch_bounds__(256) k_jd_gauss_col_thresh(const float* __restrict__ row, const uint8_t* __restrict__ blur, int w, int h,
                                                             uint8_t* __restrict__ bin) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w) return;
  float v[11];
#pragma unroll
  for (int i = 0; i < 11; i++) v[i] = row[(size_t)min(max(y + i - 5, 0), h - 1) * w + x];
  float acc = __fmul_rn(v[5], c_g11[5]);
  if (x < w - (w & 7)) {
#pragma unroll
    for (int i = 1; i <= 5; i++) acc = __fmaf_rn(__fadd_rn(v[5 + i], v[5 - i]), c_g11[5 - i], acc);
  } else {
#pragma unroll
    for (int i = 1; i <= 5; i++) acc = __fadd_rn(acc, __fmul_rn(__fadd_rn(v[5 + i], v[5 - i]), c_g11[5 - i]));
  }
  const int mean = min(max(__float2int_rn(acc), 0), 255);
  bin[(size_t)(y + 1) * (w + 2) + (x + 1)] = ((int)blur[(size_t)y * w + x] - mean > -2) ? 1 : 0;
}

// ---- connected components of the padded binary image: foreground 8-connected, background 4-connected ----
__device__ __forceinline__ int uf_find(const int* L, int a) {
  // (labels only ever decrease towards the component's first pixel; a stale read is still an ancestor)
  int p = *reinterpret_cast<const volatile int*>(L + a);
  while (p != a) { a = p; p = *reinterpret_cast<const volatile int*>(L + a); }
  return a;
}

__device__ __forceinline__ void uf_union(int* L, int a, int b) {
  while (true) {
    a = uf_find(L, a);
    b = uf_find(L, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }      // hang the larger root under the smaller: the root is the first pixel
    const int old = atomicMin(&L[a], b);
    if (old == a) return;
    a = old;
  }
}

// each pixel starts at the first pixel of its horizontal run within the warp's 32 columns (cuts the union work)
__global__ void __launch_bounds__(256) k_ccl_init(const uint8_t* __restrict__ bin, int W2, int H2, int* __restrict__ L) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  const bool in = x < W2;
  const int v = in ? bin[(size_t)y * W2 + x] : 2;
  const int left = __shfl_up_sync(0xffffffffu, v, 1);
  const bool start = (threadIdx.x & 31) == 0 || left != v;
  const unsigned starts = __ballot_sync(0xffffffffu, start);
  const unsigned lane = threadIdx.x & 31;
  const unsigned below = starts & (0xffffffffu >> (31 - lane));     // run starts at or before this lane
  const int run0 = 31 - __clz(below);
  if (in) L[(size_t)y * W2 + x] = y * W2 + (x - (int)lane + run0);
}

__global__ void __launch_bounds__(256) k_ccl_merge(const uint8_t* __restrict__ bin, int W2, int H2, int* __restrict__ L) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W2) return;
  const int p = y * W2 + x;
  const uint8_t v = bin[p];
  const bool west_same = x > 0 && bin[p - 1] == v;
  // runs were cut at the 32-column boundaries of k_ccl_init: join them there
  if (west_same && (x & 31) == 0) uf_union(L, p, p - 1);
  if (y == 0) return;
  const uint8_t* up = bin + p - W2;
  if (up[0] == v) {
    // north; not needed where the west neighbour is in the same run and already hangs under the same upper run
    if (!(west_same && (x & 31) != 0 && up[-1] == v)) uf_union(L, p, p - W2);
  } else if (v) {
    // foreground is 8-connected: the diagonals count where north does not already join them
    if (x > 0 && up[-1] && !bin[p - 1]) uf_union(L, p, p - W2 - 1);
    if (x < W2 - 1 && up[1]) uf_union(L, p, p - W2 + 1);
  }
}

__global__ void __launch_bounds__(256) k_ccl_flatten(int* __restrict__ L, int n) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) L[p] = uf_find(L, p);
}

// per foreground pixel: its cracks -> shoelace term and bounding box of the border (slot = root index of F for F's outer
// border, root index of H for the hole border around H)
__global__ void __launch_bounds__(256) k_jd_cracks(const uint8_t* __restrict__ bin, const int* __restrict__ L, int W2, int H2,
                                                   long long* __restrict__ area2, int* __restrict__ box, int n_pad) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W2 || x == 0 || y == 0 || y >= H2 - 1 || x >= W2 - 1) return;
  const int p = y * W2 + x;
  if (!bin[p]) return;
  const int f = L[p];
  const int enc = L[f - 1];                            // background component left of F's first pixel: the one around F
  const int dxs[4] = {0, 1, 0, -1}, dys[4] = {-1, 0, 1, 0};
#pragma unroll
  for (int d = 0; d < 4; d++) {
    const int q = p + dys[d] * W2 + dxs[d];
    if (bin[q]) continue;
    const int hroot = L[q];
    const int slot = hroot == enc ? f : hroot;
    const int tx = dxs[(d + 1) & 3], ty = dys[(d + 1) & 3];
    int nx, ny;
    if (bin[q + ty * W2 + tx]) { nx = x + dxs[d] + tx; ny = y + dys[d] + ty; }
    else if (bin[p + ty * W2 + tx]) { nx = x + tx; ny = y + ty; }
    else { nx = x; ny = y; }
    const long long term = (long long)x * ny - (long long)nx * y;
    if (term) atomicAdd(reinterpret_cast<unsigned long long*>(area2 + slot), (unsigned long long)term);
    atomicMin(box + slot, x);
    atomicMin(box + n_pad + slot, y);
    atomicMax(box + 2 * n_pad + slot, x);
    atomicMax(box + 3 * n_pad + slot, y);
  }
}

// junction_detector.cpp:76-105 per border; passing contours carry their ancestor keys for the host's ordering
__global__ void __launch_bounds__(256) k_jd_filter(const uint8_t* __restrict__ bin, const int* __restrict__ L, int W2, int n_pad,
                                                   const long long* __restrict__ area2, const int* __restrict__ box, double lo,
                                                   double hi, PassRec* __restrict__ pass, int* __restrict__ cnt) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_pad || L[s] != s || s == L[0]) return;    // roots only; the outermost background has no border of its own
  if (box[s] > box[2 * n_pad + s]) return;             // (defensive: a component without cracks cannot exist)
  long long a2 = area2[s];
  if (a2 < 0) a2 = -a2;
  const double area = (double)a2 * 0.5;
  if (!(lo < area && area < hi)) return;
  const int bx = box[s] - 1, by = box[n_pad + s] - 1, bw = box[2 * n_pad + s] - box[s] + 1, bh = box[3 * n_pad + s] - box[n_pad + s] + 1;
  const double width = bw, height = bh;
  if (!(area / (width * height) >= 0.4 && width / height >= 0.5 && width / height <= 2.0)) return;
  const int i = atomicAdd(cnt, 1);
  if (i >= kMaxPass) return;
  PassRec r;
  r.x = bx; r.y = by; r.w = bw; r.h = bh;
  // ancestors: outer border of F (slot f) -> hole border around the background component left of f; hole border around H
  // (slot b) -> outer border of the foreground component left of b
  int chain[kMaxDepth];
  int depth = 0, cur = s;
  const int root = L[0];
  while (true) {
    if (depth == kMaxDepth) { atomicExch(cnt + 1, 1); break; }
    chain[depth++] = cur;
    const int par = L[cur - 1];
    if (par == root) break;
    cur = par;
  }
  r.depth = depth;
  for (int k = 0; k < depth; k++) r.path[k] = chain[depth - 1 - k];
  pass[i] = r;
}

// ---- host: contour order, candidates, clustering ----
struct KdNode {
  int child1 = -1, child2 = -1;        // leaf: child1 == -1
  int left = 0, right = 0;             // leaf range
  int divfeat = 0;
  float divlow = 0, divhigh = 0;
};

// nanoflann's KDTreeSingleIndexAdaptor<L2, float, 2> as the reference uses it (junction_detector.cpp:129-147): middle
// split on the widest dimension clamped to the data, leaves of <= 7 points, radius search that opens the far child only
// while mindist * (1 + eps) <= radius^2 — the result depends on the tree, so the tree is built the same way.
struct KdTree {
  const std::vector<float>& px;        // x0 y0 x1 y1 ...
  std::vector<int> acc;                // nanoflann's vAcc_: tree order -> point index
  std::vector<float> perm;             // the points in tree order (kept in step with acc: the build and the leaf scans
                                       // then stream memory instead of chasing indices; same comparisons, same result)
  std::vector<KdNode> nodes;
  float root_lo[2], root_hi[2];
  int leaf_max;
  float at(int i, int d) const { return perm[2 * i + d]; }
  void swap_pts(int a, int b) {
    std::swap(acc[a], acc[b]);
    std::swap(perm[2 * a], perm[2 * b]);
    std::swap(perm[2 * a + 1], perm[2 * b + 1]);
  }

  KdTree(const std::vector<float>& pts, int leaf) : px(pts), leaf_max(leaf) {
    const int n = (int)pts.size() / 2;
    acc.resize(n);
    for (int i = 0; i < n; i++) acc[i] = i;
    perm = pts;
    nodes.reserve((size_t)n / 2 + 16);
    float lo[2] = {pts[0], pts[1]}, hi[2] = {pts[0], pts[1]};
    for (int i = 1; i < n; i++)
      for (int d = 0; d < 2; d++) { lo[d] = std::min(lo[d], pts[2 * i + d]); hi[d] = std::max(hi[d], pts[2 * i + d]); }
    divide(0, n, lo, hi);
    for (int d = 0; d < 2; d++) { root_lo[d] = lo[d]; root_hi[d] = hi[d]; }
  }

  int divide(int left, int right, float* lo, float* hi) {
    const int id = (int)nodes.size();
    nodes.emplace_back();
    if (right - left <= leaf_max) {
      nodes[id].left = left; nodes[id].right = right;
      const int first = std::min(left, (int)acc.size() - 1);     // (an empty leaf, possible with many duplicates, takes the next point's box as nanoflann does)
      for (int d = 0; d < 2; d++) {
        lo[d] = hi[d] = at(first, d);
        for (int k = left + 1; k < right; k++) { lo[d] = std::min(lo[d], at(k, d)); hi[d] = std::max(hi[d], at(k, d)); }
      }
      return id;
    }
    int idx, cutfeat;
    float cutval;
    middle_split(left, right - left, lo, hi, &idx, &cutfeat, &cutval);
    float llo[2] = {lo[0], lo[1]}, lhi[2] = {hi[0], hi[1]}, rlo[2] = {lo[0], lo[1]}, rhi[2] = {hi[0], hi[1]};
    lhi[cutfeat] = cutval;
    const int c1 = divide(left, left + idx, llo, lhi);
    rlo[cutfeat] = cutval;
    const int c2 = divide(left + idx, right, rlo, rhi);
    KdNode& nd = nodes[id];
    nd.child1 = c1; nd.child2 = c2; nd.divfeat = cutfeat;
    nd.divlow = lhi[cutfeat]; nd.divhigh = rlo[cutfeat];
    for (int d = 0; d < 2; d++) { lo[d] = std::min(llo[d], rlo[d]); hi[d] = std::max(lhi[d], rhi[d]); }
    return id;
  }

  void middle_split(int ind, int count, const float* lo, const float* hi, int* index, int* cutfeat, float* cutval) {
    const float EPS = 0.00001f;
    float max_span = hi[0] - lo[0];
    if (hi[1] - lo[1] > max_span) max_span = hi[1] - lo[1];
    float max_spread = -1, mn = 0, mx = 0;
    *cutfeat = 0;
    for (int d = 0; d < 2; d++) {
      const float span = hi[d] - lo[d];
      if (span > (1 - EPS) * max_span) {
        float a = at(ind, d), b = a;
        for (int k = 1; k < count; k++) { const float v = at(ind + k, d); if (v < a) a = v; if (v > b) b = v; }
        const float spread = b - a;
        if (spread > max_spread) { *cutfeat = d; max_spread = spread; mn = a; mx = b; }
      }
    }
    const float split = (lo[*cutfeat] + hi[*cutfeat]) / 2;
    *cutval = split < mn ? mn : (split > mx ? mx : split);
    int lim1, lim2;
    plane_split(ind, count, *cutfeat, *cutval, &lim1, &lim2);
    const int half = count / 2;
    *index = lim1 > half ? lim1 : (lim2 < half ? lim2 : half);
  }

  void plane_split(int ind, int count, int feat, float cutval, int* lim1, int* lim2) {
    int left = 0, right = count - 1;
    for (;;) {
      while (left <= right && at(ind + left, feat) < cutval) ++left;
      while (right && left <= right && at(ind + right, feat) >= cutval) --right;
      if (left > right || !right) break;
      swap_pts(ind + left, ind + right);
      ++left; --right;
    }
    *lim1 = left;
    right = count - 1;
    for (;;) {
      while (left <= right && at(ind + left, feat) <= cutval) ++left;
      while (right && left <= right && at(ind + right, feat) > cutval) --right;
      if (left > right || !right) break;
      swap_pts(ind + left, ind + right);
      ++left; --right;
    }
    *lim2 = left;
  }

  void search(int node, const float* q, float mindist, float* dists, float radius2, float eps_error, std::vector<int>* out) const {
    const KdNode& nd = nodes[node];
    if (nd.child1 < 0) {
      for (int i = nd.left; i < nd.right; i++) {
        const float d0 = q[0] - perm[2 * i], d1 = q[1] - perm[2 * i + 1];
        float dist = 0;
        dist += d0 * d0;
        dist += d1 * d1;
        if (dist < radius2) out->push_back(acc[i]);
      }
      return;
    }
    const float val = q[nd.divfeat];
    const float diff1 = val - nd.divlow, diff2 = val - nd.divhigh;
    int best, other;
    float cut;
    if (diff1 + diff2 < 0) { best = nd.child1; other = nd.child2; cut = (val - nd.divhigh) * (val - nd.divhigh); }
    else { best = nd.child2; other = nd.child1; cut = (val - nd.divlow) * (val - nd.divlow); }
    search(best, q, mindist, dists, radius2, eps_error, out);
    const float dst = dists[nd.divfeat];
    mindist = mindist + cut - dst;
    dists[nd.divfeat] = cut;
    if (mindist * eps_error <= radius2) search(other, q, mindist, dists, radius2, eps_error, out);
    dists[nd.divfeat] = dst;
  }

  void radius_search(const float* q, float radius2, float eps_error, std::vector<int>* out) const {
    out->clear();
    float dists[2] = {0, 0}, dist = 0;
    for (int d = 0; d < 2; d++) {
      if (q[d] < root_lo[d]) { dists[d] = (q[d] - root_lo[d]) * (q[d] - root_lo[d]); dist += dists[d]; }
      if (q[d] > root_hi[d]) { dists[d] = (q[d] - root_hi[d]) * (q[d] - root_hi[d]); dist += dists[d]; }
    }
    search(0, q, dist, dists, radius2, eps_error, out);
  }
};

// junction_detector.cpp:123-185
void cluster_junctions(const std::vector<float>& cand, int eps, std::vector<float>* centres) {
  centres->clear();
  const int n = (int)cand.size() / 2;
  if (n < 4) return;
  KdTree tree(cand, 7);
  const float radius = (float)eps;
  const float eps_error = 1 + 10.0f;
  std::vector<char> visited(n, 0);
  std::vector<int> nb;
  for (int i = 0; i < n; i++) {
    if (visited[i]) continue;
    tree.radius_search(&cand[2 * i], radius * radius, eps_error, &nb);
    if (nb.size() >= 3) {
      float x = 0, y = 0;
      for (int j : nb) { x += cand[2 * j]; y += cand[2 * j + 1]; }
      x /= nb.size();
      y /= nb.size();
      centres->push_back(x);
      centres->push_back(y);
      for (int j : nb) visited[j] = 1;
    }
  }
}

int reserve(ofb_handle* h, JunctionState* s, size_t src_bytes, size_t n_px, size_t n_pad) {
  if (src_bytes > s->src_cap) {
    if (s->d_src) cudaFree(s->d_src);
    s->d_src = nullptr; s->src_cap = 0;
    OFB_CUDA(h, cudaMalloc(&s->d_src, src_bytes));
    s->src_cap = src_bytes;
  }
  if (n_px > s->px_cap) {
    cudaFree(s->d_gray); cudaFree(s->d_blur); cudaFree(s->d_row);
    s->d_gray = s->d_blur = nullptr; s->d_row = nullptr; s->px_cap = 0;
    OFB_CUDA(h, cudaMalloc(&s->d_gray, n_px));
    OFB_CUDA(h, cudaMalloc(&s->d_blur, n_px));
    OFB_CUDA(h, cudaMalloc(&s->d_row, n_px * sizeof(float)));
    s->px_cap = n_px;
  }
  if (n_pad > s->pad_cap) {
    cudaFree(s->d_bin); cudaFree(s->d_lab); cudaFree(s->d_area); cudaFree(s->d_box);
    s->d_bin = nullptr; s->d_lab = nullptr; s->d_area = nullptr; s->d_box = nullptr; s->pad_cap = 0;
    OFB_CUDA(h, cudaMalloc(&s->d_bin, n_pad));
    OFB_CUDA(h, cudaMalloc(&s->d_lab, n_pad * sizeof(int)));
    OFB_CUDA(h, cudaMalloc(&s->d_area, n_pad * sizeof(long long)));
    OFB_CUDA(h, cudaMalloc(&s->d_box, 4 * n_pad * sizeof(int)));
    s->pad_cap = n_pad;
  }
  if (!s->d_pass) {
    OFB_CUDA(h, cudaMalloc(&s->d_pass, (size_t)kMaxPass * sizeof(PassRec)));
    OFB_CUDA(h, cudaMalloc(&s->d_cnt, 2 * sizeof(int)));
    OFB_CUDA(h, cudaHostAlloc(&s->h_pass, (size_t)kMaxPass * sizeof(PassRec), cudaHostAllocDefault));
    OFB_CUDA(h, cudaHostAlloc(&s->h_cnt, 2 * sizeof(int), cudaHostAllocDefault));
  }
  return OFB_OK;
}

// upload + the pixel stages up to the padded binary image
int threshold_stage(ofb_handle* h, JunctionState* s, const uint8_t* img, int w, int hh, size_t stride, int cn,
                    const ofb_junction_params* p) {
  const size_t row = (size_t)w * cn, n_px = (size_t)w * hh, n_pad = (size_t)(w + 2) * (hh + 2);
  cudaStream_t sm = h->stream;
  OFB_CUDA(h, cudaStreamSynchronize(sm));
  int st = reserve(h, s, row * hh, n_px, n_pad);
  if (st) return st;
  OFB_CUDA(h, cudaMemcpy2DAsync(s->d_src, row, img, stride, row, hh, cudaMemcpyHostToDevice, sm));
  if ((st = timing_begin(h, OFB_STAGE_OTHER))) return st;
  const dim3 g((w + 255) / 256, hh);
  double incline = 0, intercept = 0;
  if (p->dampen) { incline = 1.0 / (p->dampen_max - p->dampen_min); intercept = -p->dampen_min * incline; }
  k_jd_gray<<<g, 256, 0, sm>>>(s->d_src, row, w, hh, cn, p->dampen && cn == 3, incline, intercept, s->d_gray);
  OFB_LAUNCH_CHECK(h);
  k_jd_blur3<<<g, 256, 0, sm>>>(s->d_gray, w, hh, s->d_blur);
  OFB_LAUNCH_CHECK(h);
  k_jd_gauss_row<<<g, 256, 0, sm>>>(s->d_blur, w, hh, s->d_row);
  OFB_LAUNCH_CHECK(h);
  OFB_CUDA(h, cudaMemsetAsync(s->d_bin, 0, n_pad, sm));
  k_jd_gauss_col_thresh<<<g, 256, 0, sm>>>(s->d_row, s->d_blur, w, hh, s->d_bin);
  OFB_LAUNCH_CHECK(h);
  return timing_end(h);
}

}  // namespace

void junction_destroy(ofb_handle* h) {
  JunctionState* s = static_cast<JunctionState*>(h->junction);
  if (!s) return;
  cudaFree(s->d_src); cudaFree(s->d_gray); cudaFree(s->d_blur); cudaFree(s->d_row); cudaFree(s->d_bin); cudaFree(s->d_lab);
  cudaFree(s->d_area); cudaFree(s->d_box); cudaFree(s->d_pass); cudaFree(s->d_cnt);
  if (s->h_pass) cudaFreeHost(s->h_pass);
  if (s->h_cnt) cudaFreeHost(s->h_cnt);
  delete s;
  h->junction = nullptr;
}

}  // namespace ofb

using namespace ofb;

extern "C" {

static int check_args(ofb_handle* h, const uint8_t* img, int width, int height, size_t* stride, int channels,
                      const ofb_junction_params* p) {
  if (!img || !p) return set_error(h, OFB_ERR_INVALID_ARG, "NULL pointer");
  if (channels != 1 && channels != 3) return set_error(h, OFB_ERR_INVALID_ARG, "junction detector: 1 or 3 channels");
  if (width < 2 || height < 2 || (size_t)(width + 2) * (height + 2) > 0x7fffffffull)
    return set_error(h, OFB_ERR_INVALID_ARG, "bad size");
  if (*stride == 0) *stride = (size_t)width * channels;
  if (*stride < (size_t)width * channels) return set_error(h, OFB_ERR_INVALID_ARG, "stride smaller than a row");
  if (p->dampen && !(p->dampen_max != p->dampen_min)) return set_error(h, OFB_ERR_INVALID_ARG, "dampen thresholds are equal");
  return OFB_OK;
}

int ofb_junction_threshold(ofb_handle* h, const uint8_t* img, int width, int height, size_t stride_bytes, int channels,
                           const ofb_junction_params* p, uint8_t* thresh, size_t thresh_stride_bytes) {
  if (!h) return OFB_ERR_INVALID_ARG;
  if (!thresh) return set_error(h, OFB_ERR_INVALID_ARG, "NULL pointer");
  int st = check_args(h, img, width, height, &stride_bytes, channels, p);
  if (st) return st;
  if (thresh_stride_bytes == 0) thresh_stride_bytes = (size_t)width;
  if (thresh_stride_bytes < (size_t)width) return set_error(h, OFB_ERR_INVALID_ARG, "stride smaller than a row");
  OFB_CUDA(h, cudaSetDevice(h->device));
  if (!h->junction) h->junction = new JunctionState();
  JunctionState* s = static_cast<JunctionState*>(h->junction);
  if ((st = threshold_stage(h, s, img, width, height, stride_bytes, channels, p))) return st;
  // the padded 0/1 image -> the caller's 0/255 image
  std::vector<uint8_t> tmp((size_t)(width + 2) * (height + 2));
  OFB_CUDA(h, cudaMemcpyAsync(tmp.data(), s->d_bin, tmp.size(), cudaMemcpyDeviceToHost, h->stream));
  OFB_CUDA(h, cudaStreamSynchronize(h->stream));
  for (int y = 0; y < height; y++)
    for (int x = 0; x < width; x++) thresh[(size_t)y * thresh_stride_bytes + x] = tmp[(size_t)(y + 1) * (width + 2) + x + 1] ? 255 : 0;
  return OFB_OK;
}

int ofb_find_junctions(ofb_handle* h, const uint8_t* img, int width, int height, size_t stride_bytes, int channels,
                       const ofb_junction_params* p, float* junctions_xy, int capacity, int* n_out, float* candidates_xy,
                       int candidate_capacity, int* n_candidates) {
  if (!h) return OFB_ERR_INVALID_ARG;
  if (!n_out || (capacity > 0 && !junctions_xy)) return set_error(h, OFB_ERR_INVALID_ARG, "NULL pointer");
  int st = check_args(h, img, width, height, &stride_bytes, channels, p);
  if (st) return st;
  if (p->grid_area <= 0 || !(p->grid_area_threshold > 0) || p->eps <= 0)
    return set_error(h, OFB_ERR_INVALID_ARG, "junction detector: grid_area, grid_area_threshold and eps must be positive");
  OFB_CUDA(h, cudaSetDevice(h->device));
  if (!h->junction) h->junction = new JunctionState();
  JunctionState* s = static_cast<JunctionState*>(h->junction);
  if ((st = threshold_stage(h, s, img, width, height, stride_bytes, channels, p))) return st;
  const int W2 = width + 2, H2 = height + 2, n_pad = W2 * H2;
  cudaStream_t sm = h->stream;
  if ((st = timing_begin(h, OFB_STAGE_OTHER))) return st;
  const dim3 gp((W2 + 255) / 256, H2);
  k_ccl_init<<<gp, 256, 0, sm>>>(s->d_bin, W2, H2, s->d_lab);
  OFB_LAUNCH_CHECK(h);
  k_ccl_merge<<<gp, 256, 0, sm>>>(s->d_bin, W2, H2, s->d_lab);
  OFB_LAUNCH_CHECK(h);
  k_ccl_flatten<<<(n_pad + 255) / 256, 256, 0, sm>>>(s->d_lab, n_pad);
  OFB_LAUNCH_CHECK(h);
  OFB_CUDA(h, cudaMemsetAsync(s->d_area, 0, (size_t)n_pad * sizeof(long long), sm));
  OFB_CUDA(h, cudaMemsetAsync(s->d_box, 0x7f, (size_t)2 * n_pad * sizeof(int), sm));          // minima: 0x7f7f7f7f
  OFB_CUDA(h, cudaMemsetAsync(s->d_box + (size_t)2 * n_pad, 0x80, (size_t)2 * n_pad * sizeof(int), sm));   // maxima: 0x80808080 < 0
  OFB_CUDA(h, cudaMemsetAsync(s->d_cnt, 0, 2 * sizeof(int), sm));
  k_jd_cracks<<<gp, 256, 0, sm>>>(s->d_bin, s->d_lab, W2, H2, s->d_area, s->d_box, n_pad);
  OFB_LAUNCH_CHECK(h);
  // junction_detector.cpp:79: estimated_area * (1 / (2 * thr)) < area < estimated_area * (2 * thr), thr a float
  const float thr2 = 2 * p->grid_area_threshold;
  const double lo = (double)p->grid_area * (double)(1 / thr2), hi = (double)p->grid_area * (double)thr2;
  k_jd_filter<<<(n_pad + 255) / 256, 256, 0, sm>>>(s->d_bin, s->d_lab, W2, n_pad, s->d_area, s->d_box, lo, hi, s->d_pass, s->d_cnt);
  OFB_LAUNCH_CHECK(h);
  if ((st = timing_end(h))) return st;
  OFB_CUDA(h, cudaMemcpyAsync(s->h_cnt, s->d_cnt, 2 * sizeof(int), cudaMemcpyDeviceToHost, sm));
  OFB_CUDA(h, cudaStreamSynchronize(sm));
  const int n_pass = s->h_cnt[0];
  if (n_pass > kMaxPass) return set_error(h, OFB_ERR_CAPACITY, "junction detector: %d contours pass the size tests (limit %d)", n_pass, kMaxPass);
  if (s->h_cnt[1]) return set_error(h, OFB_ERR_CAPACITY, "junction detector: contours nested deeper than %d levels", kMaxDepth);
  if (n_pass) {
    OFB_CUDA(h, cudaMemcpyAsync(s->h_pass, s->d_pass, (size_t)n_pass * sizeof(PassRec), cudaMemcpyDeviceToHost, sm));
    OFB_CUDA(h, cudaStreamSynchronize(sm));
  }
  // cv2's contour order: pre-order of the border tree, siblings in descending discovery order
  std::vector<int> order(n_pass);
  for (int i = 0; i < n_pass; i++) order[i] = i;
  const PassRec* R = s->h_pass;
  std::sort(order.begin(), order.end(), [R](int a, int b) {
    const PassRec &A = R[a], &B = R[b];
    const int n = std::min(A.depth, B.depth);
    for (int k = 0; k < n; k++)
      if (A.path[k] != B.path[k]) return A.path[k] > B.path[k];
    return A.depth < B.depth;
  });
  std::vector<float> cand;
  cand.reserve((size_t)n_pass * 8);
  for (int i : order) {
    const PassRec& r = R[i];
    const float x0 = (float)(r.x - 1), y0 = (float)(r.y - 1), x1 = (float)(r.x + r.w + 1), y1 = (float)(r.y + r.h + 1);
    const float v[8] = {x0, y0, x1, y0, x1, y1, x0, y1};
    cand.insert(cand.end(), v, v + 8);
  }
  if (n_candidates) *n_candidates = (int)cand.size() / 2;
  if (candidates_xy) {
    if ((int)cand.size() / 2 > candidate_capacity) return set_error(h, OFB_ERR_CAPACITY, "%d candidates exceed the capacity %d", (int)cand.size() / 2, candidate_capacity);
    memcpy(candidates_xy, cand.data(), cand.size() * sizeof(float));
  }
  std::vector<float> centres;
  cluster_junctions(cand, p->eps, &centres);
  *n_out = (int)centres.size() / 2;
  if (*n_out > capacity) return set_error(h, OFB_ERR_CAPACITY, "%d junctions exceed the capacity %d", *n_out, capacity);
  if (*n_out) memcpy(junctions_xy, centres.data(), centres.size() * sizeof(float));
  return OFB_OK;
}

// The clustering step on its own (host only, no device work): junction candidates -> cluster centres.
int ofb_cluster_junctions(const float* candidates_xy, int n_candidates, int eps, float* junctions_xy, int capacity, int* n_out) {
  if ((n_candidates > 0 && !candidates_xy) || !n_out || eps <= 0 || n_candidates < 0) return OFB_ERR_INVALID_ARG;
  std::vector<float> cand(candidates_xy, candidates_xy + (size_t)2 * n_candidates), centres;
  cluster_junctions(cand, eps, &centres);
  *n_out = (int)centres.size() / 2;
  if (*n_out > capacity) return OFB_ERR_CAPACITY;
  if (*n_out && junctions_xy) memcpy(junctions_xy, centres.data(), centres.size() * sizeof(float));
  return OFB_OK;
}

}  // extern "C"
