// jpeg.cu — baseline-JPEG frame ingest: cv2.imdecode(buf, cv2.IMREAD_COLOR) of the compressed-image node
// (ros2_ws/src/optical_flow/optical_flow/opticalflow_comprerssed_node.py:43-46), bit-exact with the wheel's
// libjpeg-turbo 3.1.2 defaults (JDCT_ISLOW, fancy upsampling, JCS_EXT_BGR).  Restated in oracle/jpeg_np.py.
//
// Split of the work: the entropy-coded segment is a serial bit stream (every symbol's position depends on all symbols
// before it), so the HOST walks it — marker parsing and Huffman decoding into quantised coefficient blocks written
// straight into pinned staging — and everything that is per-sample arithmetic runs on the DEVICE:
//   k_jpeg_idct   dequantisation + the 13-bit fixed-point LL&M inverse DCT (jidctint.c::jpeg_idct_islow), one thread
//                 per block column / row, workspace in shared memory, planes of 8-bit samples out;
//   k_jpeg_color  chroma up-sampling with the triangle filter (jdsample.c: h2v1 / h2v2 / h1v2 "fancy", edge rules of
//                 the library) fused with YCbCr -> BGR (jdcolor.c, 16-bit fixed point) and, when asked for, the gray
//                 conversion the nodes apply next (cv2.cvtColor BGR2GRAY, 15-bit fixed point).
// The decoded frame never exists on the host unless the caller asks for it.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace ofb {

namespace {

constexpr int kMaxComp = 3;

struct HuffTab {
  // canonical code tables (T.81 F.2.2.3) + a kLook-bit look-ahead table: entry = (length << 8) | symbol, 0 = longer code
  static constexpr int kLook = 10;
  uint16_t look[1 << kLook];
  int32_t maxcode[18];   // largest code of each length, -1 if none; [17] = sentinel
  int32_t valoff[17];    // index of the first symbol of a length minus its first code
  uint8_t vals[256];
  uint8_t counts[16];
  int n_vals = 0;
  bool present = false;
};

struct Comp {
  int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
  int blocks_x = 0, blocks_y = 0;     // whole blocks of the padded plane
  int dw = 0, dh = 0;                 // real samples (ceil(width * h / hmax), ceil(height * v / vmax))
  size_t coef_off = 0;                // first coefficient of the component in the staging array (int16 units)
};

struct Frame {
  int width = 0, height = 0, nc = 0, hmax = 1, vmax = 1, mcux = 0, mcuy = 0, dri = 0;
  Comp comp[kMaxComp];
  uint16_t qt[4][64];                 // natural order
  bool qt_present[4] = {false, false, false, false};
  HuffTab dc[4], ac[4];
  const uint8_t* data = nullptr;      // entropy-coded segment
  size_t n_data = 0;
  size_t n_coef = 0;                  // int16 coefficients of all components
};

const uint8_t kZigzag[64 + 16] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,
                                  6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31,
                                  39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
                                  63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};   // corrupt runs land on 63

bool build_huff(const uint8_t* counts, const uint8_t* vals, int n, HuffTab* t) {
  memset(t->look, 0, sizeof(t->look));
  memset(t->vals, 0, sizeof(t->vals));
  memcpy(t->vals, vals, n);
  memcpy(t->counts, counts, 16);
  t->n_vals = n;
  int code = 0, k = 0;
  for (int len = 1; len <= 16; len++) {
    t->valoff[len] = k - code;
    if (counts[len - 1]) {
      for (int i = 0; i < counts[len - 1]; i++, code++, k++) {
        if (len <= HuffTab::kLook) {
          const int lo = code << (HuffTab::kLook - len);
          for (int j = 0; j < (1 << (HuffTab::kLook - len)); j++) t->look[lo + j] = (uint16_t)((len << 8) | vals[k]);
        }
      }
      t->maxcode[len] = code - 1;
    } else {
      t->maxcode[len] = -1;
    }
    if (code > (1 << len)) return false;
    code <<= 1;
  }
  t->maxcode[17] = 0x7fffffff;
  t->present = true;
  return true;
}

// Markers up to the first SOS.  Returns NULL or the reason the stream is not served.
const char* parse_jpeg(const uint8_t* b, size_t n, Frame* f, bool header_only) {
  if (n < 4 || b[0] != 0xFF || b[1] != 0xD8) return "not a JPEG stream";
  size_t pos = 2;
  bool have_sof = false;
  while (true) {
    while (pos < n && b[pos] != 0xFF) pos++;
    while (pos < n && b[pos] == 0xFF) pos++;
    if (pos >= n) return "no scan in the JPEG stream";
    const int m = b[pos++];
    if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
    if (m == 0xD9) return "no scan in the JPEG stream";
    if (pos + 2 > n) return "truncated JPEG stream";
    const size_t len = ((size_t)b[pos] << 8) | b[pos + 1];
    if (len < 2 || pos + len > n) return "truncated JPEG stream";
    const uint8_t* s = b + pos + 2;
    const size_t sl = len - 2;
    if (m == 0xDB) {
      size_t i = 0;
      while (i < sl) {
        const int pq = s[i] >> 4, tq = s[i] & 15;
        i++;
        if (tq > 3 || i + (pq ? 128 : 64) > sl) return "bad quantisation table";
        for (int k = 0; k < 64; k++) {
          f->qt[tq][kZigzag[k]] = pq ? (uint16_t)((s[i + 2 * k] << 8) | s[i + 2 * k + 1]) : s[i + k];
        }
        f->qt_present[tq] = true;
        i += pq ? 128 : 64;
      }
    } else if (m == 0xC4) {
      size_t i = 0;
      while (i < sl) {
        if (i + 17 > sl) return "bad Huffman table";
        const int tc = s[i] >> 4, th = s[i] & 15;
        int cnt = 0;
        for (int k = 0; k < 16; k++) cnt += s[i + 1 + k];
        if (tc > 1 || th > 3 || cnt > 256 || i + 17 + cnt > sl) return "bad Huffman table";
        if (!build_huff(s + i + 1, s + i + 17, cnt, tc ? &f->ac[th] : &f->dc[th])) return "bad Huffman table";
        i += 17 + cnt;
      }
    } else if (m == 0xC0 || m == 0xC1) {
      if (sl < 6) return "bad frame header";
      if (s[0] != 8) return "only 8-bit JPEG is decoded on the device";
      f->height = (s[1] << 8) | s[2];
      f->width = (s[3] << 8) | s[4];
      f->nc = s[5];
      if (f->nc != 1 && f->nc != 3) return "only gray-scale and YCbCr JPEG (1 or 3 components) are decoded on the device";
      if (sl < (size_t)6 + 3 * f->nc || f->width < 1 || f->height < 1) return "bad frame header";
      for (int c = 0; c < f->nc; c++) {
        Comp& k = f->comp[c];
        k.id = s[6 + 3 * c];
        k.h = s[7 + 3 * c] >> 4;
        k.v = s[7 + 3 * c] & 15;
        k.tq = s[8 + 3 * c];
        if (k.tq > 3 || k.h < 1 || k.v < 1 || k.h > 2 || k.v > 2) return "sampling factors beyond 2 are not decoded on the device";
      }
      have_sof = true;
      if (header_only) return nullptr;
    } else if (m >= 0xC2 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
      return "only baseline (SOF0/SOF1, Huffman, sequential) JPEG is decoded on the device";
    } else if (m == 0xDD) {
      if (sl < 2) return "bad restart interval";
      f->dri = (s[0] << 8) | s[1];
    } else if (m == 0xDA) {
      if (!have_sof) return "scan before the frame header";
      if (sl < 1 || s[0] != f->nc || sl < (size_t)1 + 2 * f->nc + 3) return "non-interleaved scans are not decoded on the device";
      for (int c = 0; c < f->nc; c++) {
        Comp* k = nullptr;
        for (int j = 0; j < f->nc; j++)
          if (f->comp[j].id == s[1 + 2 * c]) k = &f->comp[j];
        if (!k || k != &f->comp[c]) return "scan component order differs from the frame header";
        k->td = s[2 + 2 * c] >> 4;
        k->ta = s[2 + 2 * c] & 15;
        if (k->td > 3 || k->ta > 3 || !f->dc[k->td].present || !f->ac[k->ta].present || !f->qt_present[k->tq])
          return "scan refers to a missing table";
      }
      f->data = b + pos + len;
      f->n_data = n - (pos + len);
      break;
    }
    pos += len;
  }
  if (f->nc == 1) f->comp[0].h = f->comp[0].v = 1;       // a one-component scan is never interleaved
  f->hmax = f->vmax = 1;
  for (int c = 0; c < f->nc; c++) { f->hmax = std::max(f->hmax, f->comp[c].h); f->vmax = std::max(f->vmax, f->comp[c].v); }
  if (f->nc == 3) {
    // luma at full resolution, both chroma planes with the same factors: 4:4:4, 4:2:2, 4:2:0, 4:4:0
    if (f->comp[0].h != f->hmax || f->comp[0].v != f->vmax || f->comp[1].h != f->comp[2].h || f->comp[1].v != f->comp[2].v ||
        f->comp[1].h != 1 || f->comp[1].v != 1)
      return "chroma sampling other than 4:4:4 / 4:2:2 / 4:2:0 / 4:4:0 is not decoded on the device";
    if (f->hmax == 2 && (f->width + 1) / 2 <= 2) return "frame too narrow";   // (the library switches to box up-sampling there)
  }
  f->mcux = (f->width + 8 * f->hmax - 1) / (8 * f->hmax);
  f->mcuy = (f->height + 8 * f->vmax - 1) / (8 * f->vmax);
  size_t off = 0;
  for (int c = 0; c < f->nc; c++) {
    Comp& k = f->comp[c];
    k.blocks_x = f->mcux * k.h;
    k.blocks_y = f->mcuy * k.v;
    k.dw = (f->width * k.h + f->hmax - 1) / f->hmax;
    k.dh = (f->height * k.v + f->vmax - 1) / f->vmax;
    k.coef_off = off;
    off += (size_t)k.blocks_x * k.blocks_y * 64;
  }
  f->n_coef = off;
  return nullptr;
}

// The entropy-coded segment with the byte stuffing removed (FF00 -> FF) so that the bit reader loads whole words:
// restart markers are dropped and their byte positions recorded, any other marker ends the data, and zeros follow it
// (libjpeg feeds zeros past the end too: a truncated stream decodes to grey instead of failing).
struct Unstuffed {
  std::vector<uint8_t> buf;
  std::vector<size_t> rst;       // byte offset (in buf) of the data behind the i-th restart marker
};

void unstuff(const uint8_t* p, size_t n, Unstuffed* u) {
  u->buf.resize(n + 64);
  u->rst.clear();
  uint8_t* o = u->buf.data();
  const uint8_t* end = p + n;
  while (p < end) {
    const uint8_t* q = static_cast<const uint8_t*>(memchr(p, 0xFF, end - p));
    if (!q) q = end;
    memcpy(o, p, q - p);
    o += q - p;
    p = q;
    if (p >= end) break;
    if (p + 1 >= end) break;
    const uint8_t m = p[1];
    if (m == 0) { *o++ = 0xFF; p += 2; }
    else if (m >= 0xD0 && m <= 0xD7) { u->rst.push_back((size_t)(o - u->buf.data())); p += 2; }
    else if (m == 0xFF) { p++; }                         // fill byte
    else break;                                          // EOI or another marker
  }
  const size_t used = o - u->buf.data();
  memset(o, 0, u->buf.size() - used);
  u->buf.resize(used + 64);                              // (no reallocation: shrinking)
}

struct BitReader {
  const uint8_t* p;
  const uint8_t* end;            // 32 readable bytes before the end of the padded buffer
  uint64_t acc = 0;
  int n = 0;
  inline void fill() {           // >= 32 valid bits afterwards
    if (n < 32) {
      uint32_t v = 0;
      if (p < end) { memcpy(&v, p, 4); p += 4; }
      acc |= (uint64_t)__builtin_bswap32(v) << (32 - n);
      n += 32;
    }
  }
  inline uint32_t peek(int k) const { return (uint32_t)(acc >> (64 - k)); }
  inline void skip(int k) { acc <<= k; n -= k; }
};

inline int extend(uint32_t v, int s) { return (int)v - ((v >> (s - 1)) ? 0 : (1 << s) - 1); }

inline int huff_decode(BitReader& br, const HuffTab& t) {
  const uint32_t e = t.look[br.peek(HuffTab::kLook)];
  if (e) { br.skip(e >> 8); return e & 255; }
  int len = HuffTab::kLook + 1;
  int32_t code = (int32_t)br.peek(len);
  while (len <= 16 && code > t.maxcode[len]) { len++; code = (int32_t)br.peek(len); }
  if (len > 16) { br.skip(16); return 0; }          // corrupt code: libjpeg warns and returns 0
  br.skip(len);
  return t.vals[(code + t.valoff[len]) & 255];
}

// AC look-ahead with the value bits folded in: for the next kLook bits, entry = value << 16 | run << 8 | bits consumed
// when the code AND its value bits fit (the common case: short codes of small coefficients); 0 otherwise.
struct AcFast { int32_t e[1 << HuffTab::kLook]; };

void build_ac_fast(const HuffTab& t, AcFast* f) {
  for (int i = 0; i < (1 << HuffTab::kLook); i++) {
    f->e[i] = 0;
    const uint32_t e = t.look[i];
    if (!e) continue;
    const int len = e >> 8, rs = e & 255, r = rs >> 4, s = rs & 15;
    if (s == 0 || len + s > HuffTab::kLook) continue;
    const uint32_t bits = ((uint32_t)i >> (HuffTab::kLook - len - s)) & ((1u << s) - 1);
    f->e[i] = (int32_t)(((uint32_t)(extend(bits, s) & 0xffff) << 16) | (uint32_t)(r << 8) | (uint32_t)(len + s));
  }
}

// All MCUs of the scan into `coef` (zeroed by the caller): component planes of blocks, 64 int16 each in natural order.
void decode_scan(const Frame& f, int16_t* coef) {
  static thread_local Unstuffed u;
  unstuff(f.data, f.n_data, &u);
  static thread_local AcFast fast[4];
  for (int c = 0; c < f.nc; c++) build_ac_fast(f.ac[f.comp[c].ta], &fast[f.comp[c].ta]);
  const uint8_t* base = u.buf.data();
  BitReader br{base, base + u.buf.size() - 32};
  int pred[kMaxComp] = {0, 0, 0};
  int until_restart = f.dri;
  size_t n_rst = 0;
  for (int my = 0; my < f.mcuy; my++)
    for (int mx = 0; mx < f.mcux; mx++) {
      if (f.dri) {
        if (until_restart == 0) {
          // behind the next restart marker (a stream that has run out of markers continues in zeros)
          br.p = n_rst < u.rst.size() ? base + u.rst[n_rst] : br.end;
          n_rst++;
          br.acc = 0; br.n = 0;
          pred[0] = pred[1] = pred[2] = 0;
          until_restart = f.dri;
        }
        until_restart--;
      }
      for (int c = 0; c < f.nc; c++) {
        const Comp& k = f.comp[c];
        const HuffTab& dct = f.dc[k.td];
        const HuffTab& act = f.ac[k.ta];
        const int32_t* fa = fast[k.ta].e;
        for (int by = 0; by < k.v; by++)
          for (int bx = 0; bx < k.h; bx++) {
            int16_t* blk = coef + k.coef_off + ((size_t)(my * k.v + by) * k.blocks_x + (mx * k.h + bx)) * 64;
            br.fill();
            int s = huff_decode(br, dct) & 15;
            if (s) { br.fill(); pred[c] += extend(br.peek(s), s); br.skip(s); }
            blk[0] = (int16_t)pred[c];
            for (int i = 1; i < 64;) {
              br.fill();                                 // >= 32 bits: a code (<= 16) and its value bits (<= 15)
              const int32_t e = fa[br.peek(HuffTab::kLook)];
              if (e) {
                i += (e >> 8) & 15;
                br.skip(e & 31);
                blk[kZigzag[i]] = (int16_t)(e >> 16);   // i <= 78: the table is padded
                i++;
                continue;
              }
              const int rs = huff_decode(br, act);
              const int r = rs >> 4;
              s = rs & 15;
              if (s == 0) {
                if (r != 15) break;
                i += 16;
                continue;
              }
              i += r;
              blk[kZigzag[i]] = (int16_t)extend(br.peek(s), s);
              br.skip(s);
              i++;
            }
          }
      }
    }
}

// ---- device ----
struct IdctPlane {
  const int16_t* coef;     // [blocks][64]
  uint8_t* plane;          // [blocks_y * 8][pitch]
  int blocks_x, n_blocks;
  size_t pitch;
  int first_block;         // index of the component's first block in the launch
  const int16_t* dc;       // device entropy decoding: DC values in the component's scan order (NULL: coefficient 0 holds the DC)
  int h, v, mcux;          // blocks per MCU of the component, MCUs per row (for dc)
  uint16_t q[64];
};
struct IdctArgs {
  IdctPlane p[kMaxComp];
  int nc, total_blocks;
};

// jidctint.c constants (CONST_BITS = 13)
#define J_0_298 2446
#define J_0_390 3196
#define J_0_541 4433
#define J_0_765 6270
#define J_0_899 7373
#define J_1_175 9633
#define J_1_501 12299
#define J_1_847 15137
#define J_1_961 16069
#define J_2_053 16819
#define J_2_562 20995
#define J_3_072 25172

template <int SHIFT>
__device__ __forceinline__ void idct8(const int (&x)[8], int (&o)[8]) {
  int z1 = (x[2] + x[6]) * J_0_541;
  const int tmp2 = z1 - x[6] * J_1_847, tmp3 = z1 + x[2] * J_0_765;
  const int tmp0 = (x[0] + x[4]) << 13, tmp1 = (x[0] - x[4]) << 13;
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  int t0 = x[7], t1 = x[5], t2 = x[3], t3 = x[1];
  z1 = t0 + t3;
  int z2 = t1 + t2, z3 = t0 + t2, z4 = t1 + t3;
  const int z5 = (z3 + z4) * J_1_175;
  t0 *= J_0_298; t1 *= J_2_053; t2 *= J_3_072; t3 *= J_1_501;
  z1 *= -J_0_899; z2 *= -J_2_562;
  z3 = z3 * -J_1_961 + z5;
  z4 = z4 * -J_0_390 + z5;
  t0 += z1 + z3; t1 += z2 + z4; t2 += z2 + z3; t3 += z1 + z4;
  constexpr int R = 1 << (SHIFT - 1);
  o[0] = (tmp10 + t3 + R) >> SHIFT; o[7] = (tmp10 - t3 + R) >> SHIFT;
  o[1] = (tmp11 + t2 + R) >> SHIFT; o[6] = (tmp11 - t2 + R) >> SHIFT;
  o[2] = (tmp12 + t1 + R) >> SHIFT; o[5] = (tmp12 - t1 + R) >> SHIFT;
  o[3] = (tmp13 + t0 + R) >> SHIFT; o[4] = (tmp13 - t0 + R) >> SHIFT;
}

// 256 threads = 32 blocks; thread (b, t): column t of block b in pass 1, row t in pass 2.
__global__ void __launch_bounds__(256) k_jpeg_idct(const __grid_constant__ IdctArgs a) {
  __shared__ int ws[32][8][9];
  const int lb = threadIdx.x >> 3, t = threadIdx.x & 7;
  const int gb = blockIdx.x * 32 + lb;
  const bool live = gb < a.total_blocks;
  int c = 0;
  if (live) {
    if (a.nc > 1 && gb >= a.p[1].first_block) c = 1;
    if (a.nc > 2 && gb >= a.p[2].first_block) c = 2;
  }
  const IdctPlane& P = a.p[c];
  const int b = gb - P.first_block;
  if (live) {
    const int16_t* cf = P.coef + (size_t)b * 64;
    int x[8], o[8];
#pragma unroll
    for (int r = 0; r < 8; r++) x[r] = (int)(int16_t)((int)cf[r * 8 + t] * (int)P.q[r * 8 + t]);   // 16-bit product, as the library's vector code
    if (t == 0 && P.dc) {
      const int by = b / P.blocks_x, bx = b - by * P.blocks_x;
      const int mcu = (by / P.v) * P.mcux + bx / P.h, j = (by % P.v) * P.h + bx % P.h;
      x[0] = (int)(int16_t)((int)P.dc[(size_t)mcu * (P.h * P.v) + j] * (int)P.q[0]);
    }
    idct8<11>(x, o);
#pragma unroll
    for (int r = 0; r < 8; r++) ws[lb][r][t] = min(max(o[r], -32768), 32767);
  }
  __syncwarp();
  if (live) {
    int x[8], o[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = ws[lb][t][k];
    idct8<18>(x, o);
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      lo |= (uint32_t)(min(max(o[k], -128), 127) + 128) << (8 * k);
      hi |= (uint32_t)(min(max(o[k + 4], -128), 127) + 128) << (8 * k);
    }
    const int by = b / P.blocks_x, bx = b - by * P.blocks_x;
    *reinterpret_cast<uint2*>(P.plane + (size_t)(by * 8 + t) * P.pitch + (size_t)bx * 8) = make_uint2(lo, hi);
  }
}

struct ColorArgs {
  const uint8_t *Y, *Cb, *Cr;
  size_t y_pitch, c_pitch;
  int w, h, cw, ch;        // frame size; real chroma samples
  int fh, fv;              // chroma up-sampling factors (1 or 2)
  int gray_only_source;    // 1: one-component JPEG (B = G = R = Y)
  uint8_t* bgr;            // may be NULL
  size_t bgr_pitch;
  uint8_t* gray;           // may be NULL
  size_t gray_pitch;
};

__device__ __forceinline__ int chroma_at(const uint8_t* __restrict__ C, size_t pitch, int x, int y, int cw, int ch, int fh, int fv) {
  if (fh == 1 && fv == 1) return C[(size_t)y * pitch + x];
  if (fv == 1) {                                         // h2v1_fancy_upsample
    const int cx = x >> 1;
    const uint8_t* r = C + (size_t)y * pitch;
    const int cur = r[cx];
    if (x & 1) return cx == cw - 1 ? cur : (3 * cur + r[cx + 1] + 2) >> 2;
    return cx == 0 ? cur : (3 * cur + r[cx - 1] + 1) >> 2;
  }
  const int cy = y >> 1;
  const int oy = (y & 1) ? min(cy + 1, ch - 1) : max(cy - 1, 0);
  const uint8_t* r0 = C + (size_t)cy * pitch;
  const uint8_t* r1 = C + (size_t)oy * pitch;
  if (fh == 1) return (3 * r0[x] + r1[x] + ((y & 1) ? 2 : 1)) >> 2;   // h1v2_fancy_upsample
  const int cx = x >> 1;                                 // h2v2_fancy_upsample
  const int cs = 3 * r0[cx] + r1[cx];
  if (x & 1) return cx == cw - 1 ? (4 * cs + 7) >> 4 : (3 * cs + 3 * r0[cx + 1] + r1[cx + 1] + 7) >> 4;
  return cx == 0 ? (4 * cs + 8) >> 4 : (3 * cs + 3 * r0[cx - 1] + r1[cx - 1] + 8) >> 4;
}

// One thread = 4 consecutive pixels of a row: 12 bytes of BGR (three 32-bit stores on a 4-byte-aligned row) and 4 of gray.
__global__ void __launch_bounds__(256) k_jpeg_color(const __grid_constant__ ColorArgs a) {
  const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
  if (x4 >= a.w) return;
  const int n = min(4, a.w - x4);
  uint8_t px[12], g[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int x = min(x4 + j, a.w - 1);
    const int yy = a.Y[(size_t)y * a.y_pitch + x];
    int b = yy, gg = yy, r = yy;
    if (!a.gray_only_source) {
      const int cb = chroma_at(a.Cb, a.c_pitch, x, y, a.cw, a.ch, a.fh, a.fv) - 128;
      const int cr = chroma_at(a.Cr, a.c_pitch, x, y, a.cw, a.ch, a.fh, a.fv) - 128;
      r = min(max(yy + ((91881 * cr + 32768) >> 16), 0), 255);
      b = min(max(yy + ((116130 * cb + 32768) >> 16), 0), 255);
      gg = min(max(yy + ((-22554 * cb + 32768 - 46802 * cr) >> 16), 0), 255);
    }
    px[3 * j] = (uint8_t)b; px[3 * j + 1] = (uint8_t)gg; px[3 * j + 2] = (uint8_t)r;
    g[j] = (uint8_t)((b * 3735 + gg * 19235 + r * 9798 + 16384) >> 15);
  }
  if (a.bgr) {
    uint8_t* d = a.bgr + (size_t)y * a.bgr_pitch + (size_t)x4 * 3;
    if (n == 4 && ((reinterpret_cast<uintptr_t>(d) & 3) == 0)) {
      uint32_t* d32 = reinterpret_cast<uint32_t*>(d);
      d32[0] = px[0] | (px[1] << 8) | (px[2] << 16) | ((uint32_t)px[3] << 24);
      d32[1] = px[4] | (px[5] << 8) | (px[6] << 16) | ((uint32_t)px[7] << 24);
      d32[2] = px[8] | (px[9] << 8) | (px[10] << 16) | ((uint32_t)px[11] << 24);
    } else {
      for (int i = 0; i < 3 * n; i++) d[i] = px[i];
    }
  }
  if (a.gray) {
    uint8_t* d = a.gray + (size_t)y * a.gray_pitch + x4;
    if (n == 4 && ((reinterpret_cast<uintptr_t>(d) & 3) == 0)) *reinterpret_cast<uint32_t*>(d) = g[0] | (g[1] << 8) | (g[2] << 16) | ((uint32_t)g[3] << 24);
    else for (int j = 0; j < n; j++) d[j] = g[j];
  }
}

// ---- entropy decoding on the device (streams without restart markers) ----
// A Huffman stream has no entry points, but a JPEG decoder started at an arbitrary bit in an arbitrary state falls
// into step with the true decode after a few symbols, and from then on stays in step (self-synchronisation; Klein &
// Wiseman 2003, Weissenberger & Schmidt 2018).  The bit stream is cut into subsequences of kSubBits bits, one thread
// each.  The decoder state at a bit position is (bit position, block index within the MCU, zig-zag index within the
// block).  k_huff_sync finds every subsequence's TRUE entry state as a fixed point: a thread decodes its subsequence
// from its predecessor's current exit state and publishes its own exit state; subsequence 0 starts from the true
// state, so once nothing changes any more every entry state is the true one.  Rounds inside a CTA of 256 consecutive
// subsequences cost no launch (shared memory + barrier); the dependency across CTAs is carried by re-launching until a
// launch changes nothing (typically three launches).  The same pass counts the blocks each subsequence completes; an
// exclusive scan of the counts gives each subsequence the number of its first block, and k_huff_write decodes once more
// into the coefficient planes (DC differences first; k_dc_prefix turns them into DC values per component).
constexpr int kSubBits = 512;
constexpr int kSyncThreads = 256;

struct HuffLayout {
  const uint32_t* bits;        // unstuffed stream as big-endian bytes, zero padded
  uint32_t total_bits;
  int n_sub;
  int bpm;                     // blocks per MCU
  int blk_comp[6];             // component of block b of an MCU
  int blk_j[6];                // index of the block among its component's blocks in the MCU
  int blk_dc[6], blk_ac[6];    // Huffman tables (0, 1)
  uint32_t blk_tabs;           // the same, 4 bits per block: bit 0 DC table, bit 1 AC table
  int comp_h[kMaxComp], comp_v[kMaxComp], comp_blocks_x[kMaxComp];
  unsigned long long comp_off[kMaxComp];   // int16 offset of the component's first coefficient
  unsigned long long dc_off[kMaxComp];     // offset of the component's DC differences in the compact array
  int mcux, total_blocks;
  const uint16_t* lut;         // [4][65536]: DC 0, DC 1, AC 0, AC 1; entry = length << 8 | symbol (0: no such code)
  // subsequences: bit range, and for the first subsequence of a restart interval (whose entry state is known: block 0,
  // DC next, at the interval's first bit) the number of the interval's first block; -1 otherwise.  sub_headidx = index of
  // the interval's first subsequence.
  const uint32_t* sub_start;
  const uint32_t* sub_end;
  const int* sub_head_block;
  const int* sub_headidx;
  int dri_blocks;              // blocks of one component-independent restart interval in MCUs (0: no restart intervals)
};

__device__ __forceinline__ uint32_t be32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

// the 32 bits that start at bit position p
__device__ __forceinline__ uint32_t peek32(const uint32_t* __restrict__ bits, uint32_t p) {
  const uint32_t i = p >> 5, o = p & 31;
  return __funnelshift_l(be32(__ldg(bits + i + 1)), be32(__ldg(bits + i)), o);
}

__device__ __forceinline__ unsigned long long pack_state(uint32_t p, int blk, int z) {
  return ((unsigned long long)p << 16) | (unsigned long long)(blk << 8) | (unsigned long long)z;
}

__constant__ uint8_t c_zigzag[80];

constexpr int kPrimBits = 12;          // primary look-up (shared memory); longer codes go to the full table in global memory

// Decodes from `state` until the bit position reaches end_bit; returns the exit state, counts completed blocks.
// WRITE: coefficients go to their blocks, numbered from first_block on; DC differences to dc_diff (component, scan order).
// The stream is read through a three-word register window (the word after next is always in flight); a code and its
// value bits (<= 31 bits) come out of one 32-bit view.
template <bool WRITE>
__device__ __forceinline__ unsigned long long huff_run(const HuffLayout& L, const uint16_t* __restrict__ prim, unsigned long long state,
                                                       uint32_t end_bit, int* n_blocks, int first_block, int block_limit,
                                                       int16_t* __restrict__ coef, int16_t* __restrict__ dc_diff) {
  uint32_t p = (uint32_t)(state >> 16);
  int blk = (int)(state >> 8) & 0xff, z = (int)state & 0xff;
  int done = 0;
  int16_t* dst = nullptr;
  int16_t* dcp = nullptr;
  auto open_block = [&](int g) {
    dst = nullptr;
    if (g >= block_limit) return;      // past the interval's (or the scan's) last block: padding bits, or a corrupt stream
    const int mcu = g / L.bpm, b = g - mcu * L.bpm;
    const int c = L.blk_comp[b], j = L.blk_j[b], ch = L.comp_h[c], cv = L.comp_v[c];
    const int my = mcu / L.mcux, mx = mcu - my * L.mcux;
    const int jy = ch == 1 ? j : j >> 1, jx = ch == 1 ? 0 : j & 1;
    dst = coef + L.comp_off[c] + ((size_t)(my * cv + jy) * L.comp_blocks_x[c] + (mx * ch + jx)) * 64;
    dcp = dc_diff + L.dc_off[c] + (size_t)mcu * (ch * cv) + j;
  };
  if (WRITE) open_block(first_block);
  uint32_t idx = p >> 5;
  uint32_t w0 = be32(__ldg(L.bits + idx)), w1 = be32(__ldg(L.bits + idx + 1)), w2 = be32(__ldg(L.bits + idx + 2));
  const uint32_t tabs = L.blk_tabs;          // per block of the MCU: bit 0 DC table, bit 1 AC table (4 bits per block)
  const int bpm = L.bpm;
  // the loop body is written without data-dependent branches (the 32 lanes of a warp are at 32 unrelated places of
  // the stream: DC / AC / end-of-block / zero-run decisions as selects keep the warp converged)
  while (p < end_bit) {
    const uint32_t w = __funnelshift_l(w1, w0, p & 31);
    const bool dc = z == 0;
    const uint32_t tb = tabs >> (blk << 2);
    const int t = dc ? (int)(tb & 1) : 2 + (int)((tb >> 1) & 1);
    uint32_t e = prim[(t << kPrimBits) + (w >> (32 - kPrimBits))];
    if (!e) e = __ldg(L.lut + (t << 16) + (w >> 16));          // code longer than kPrimBits (rare)
    const int len = e ? (int)(e >> 8) : 16, rs = e & 255;
    const int s = rs & 15, r = rs >> 4;
    const uint32_t v = (uint32_t)((unsigned long long)(w << len) >> (32 - s));
    const int val = (int)v - (((v >> max(s - 1, 0)) & 1) ? 0 : (1 << s) - 1);
    p += len + s;
    const int zw = z + r;                                       // AC: where the coefficient goes
    const int zn = dc ? 1 : (s == 0 ? (r == 15 ? z + 16 : 64) : zw + 1);
    if (WRITE && dst) {
      if (dc) *dcp = (int16_t)val;
      else if (s) dst[c_zigzag[min(zw, 79)]] = (int16_t)val;
    }
    const bool complete = zn >= 64;
    z = complete ? 0 : zn;
    blk = complete ? (blk + 1 == bpm ? 0 : blk + 1) : blk;
    done += complete ? 1 : 0;
    if (WRITE && complete) open_block(first_block + done);
    if ((p >> 5) != idx) {
      idx++;
      w0 = w1; w1 = w2;
      w2 = be32(__ldg(L.bits + idx + 2));
    }
  }
  *n_blocks = done;
  return pack_state(p, blk, z);
}

// the first kPrimBits of every table into shared memory: entry of the full table if the code is that short, else 0
__device__ __forceinline__ void load_primary(const HuffLayout& L, uint16_t* prim) {
  for (int k = threadIdx.x; k < (4 << kPrimBits); k += blockDim.x) {
    const int t = k >> kPrimBits, i = k & ((1 << kPrimBits) - 1);
    const uint16_t e = L.lut[(t << 16) + (i << (16 - kPrimBits))];
    prim[k] = (e >> 8) <= kPrimBits ? e : 0;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kSyncThreads) k_huff_sync(const __grid_constant__ HuffLayout L, unsigned long long* __restrict__ exit_state,
                                                            int* __restrict__ counts, int first_launch, int* __restrict__ changed) {
  __shared__ unsigned long long sh_exit[kSyncThreads];
  __shared__ uint16_t prim[4 << kPrimBits];
  load_primary(L, prim);
  const int tid = threadIdx.x, i = blockIdx.x * kSyncThreads + tid;
  const bool live = i < L.n_sub;
  const uint32_t end_bit = live ? L.sub_end[i] : 0;
  // entry state: the true start for the first subsequence of a restart interval (or of the scan); otherwise, in the
  // first launch, a guess (block 0, DC next) at the subsequence's first bit, afterwards the predecessor's exit state of
  // the previous launch
  const bool head = live && L.sub_head_block[i] >= 0;
  unsigned long long entry = pack_state(live ? L.sub_start[i] : 0, 0, 0);
  if (live && !head && !first_launch) entry = exit_state[i - 1];
  unsigned long long last_entry = ~0ull, my_exit = 0;
  int cnt = 0;
  while (true) {
    bool redo = false;
    if (live && entry != last_entry) {
      my_exit = huff_run<false>(L, prim, entry, end_bit, &cnt, 0, 0, nullptr, nullptr);
      last_entry = entry;
      redo = true;
    }
    sh_exit[tid] = my_exit;
    if (!__syncthreads_or(redo)) break;
    if (tid > 0 && live && !head) entry = sh_exit[tid - 1];
    __syncthreads();
  }
  if (live) {
    if (first_launch || exit_state[i] != my_exit) { exit_state[i] = my_exit; *changed = 1; }
    counts[i] = cnt;
  }
}

// exclusive scan of the block counts (one CTA; n_sub is a few thousand)
__global__ void __launch_bounds__(1024) k_huff_scan(const int* __restrict__ counts, int* __restrict__ first_block, int n) {
  __shared__ int sh[1024];
  const int tid = threadIdx.x, per = (n + 1023) / 1024;
  const int lo = min(tid * per, n), hi = min(lo + per, n);
  int s = 0;
  for (int k = lo; k < hi; k++) s += counts[k];
  sh[tid] = s;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    const int a = tid >= d ? sh[tid - d] : 0;
    __syncthreads();
    sh[tid] += a;
    __syncthreads();
  }
  int run = sh[tid] - s;
  for (int k = lo; k < hi; k++) { first_block[k] = run; run += counts[k]; }
}

// blocks before a subsequence = the interval's first block + the blocks completed since the interval's first subsequence
__global__ void __launch_bounds__(256) k_huff_rebase(const __grid_constant__ HuffLayout L, const int* __restrict__ excl, int* __restrict__ first_block) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= L.n_sub) return;
  const int hd = L.sub_headidx[i];
  first_block[i] = L.sub_head_block[hd] + excl[i] - excl[hd];
}

__global__ void __launch_bounds__(kSyncThreads) k_huff_write(const __grid_constant__ HuffLayout L, const unsigned long long* __restrict__ exit_state,
                                                             const int* __restrict__ first_block, int16_t* __restrict__ coef,
                                                             int16_t* __restrict__ dc_diff) {
  __shared__ uint16_t prim[4 << kPrimBits];
  load_primary(L, prim);
  const int i = blockIdx.x * kSyncThreads + threadIdx.x;
  if (i >= L.n_sub) return;
  const unsigned long long entry = L.sub_head_block[i] >= 0 ? pack_state(L.sub_start[i], 0, 0) : exit_state[i - 1];
  int cnt;
  const int head_block = L.sub_head_block[L.sub_headidx[i]];
  const int limit = L.dri_blocks ? min(head_block + L.dri_blocks, L.total_blocks) : L.total_blocks;
  huff_run<true>(L, prim, entry, L.sub_end[i], &cnt, first_block[i], limit, coef, dc_diff);
}

// DC differences -> DC values, in place in the compact array (scan order of the component): one CTA per component.
// k_jpeg_idct takes a block's DC from there.
__global__ void __launch_bounds__(1024) k_dc_prefix(const __grid_constant__ HuffLayout L, int16_t* __restrict__ dc, int n_mcu) {
  __shared__ int sh[1024];
  const int c = blockIdx.x, tid = threadIdx.x;
  const int n = n_mcu * L.comp_h[c] * L.comp_v[c], per = (n + 1023) / 1024;
  const int lo = min(tid * per, n), hi = min(lo + per, n);
  int16_t* d = dc + L.dc_off[c];
  int s = 0;
  for (int k = lo; k < hi; k++) s += d[k];
  sh[tid] = s;
  __syncthreads();
  for (int dd = 1; dd < 1024; dd <<= 1) {
    const int a = tid >= dd ? sh[tid - dd] : 0;
    __syncthreads();
    sh[tid] += a;
    __syncthreads();
  }
  int run = sh[tid] - s;
  for (int k = lo; k < hi; k++) { run += d[k]; d[k] = (int16_t)run; }
}

// the same with restart intervals (the prediction restarts with every interval): one warp per (component, interval)
__global__ void __launch_bounds__(256) k_dc_prefix_intervals(const __grid_constant__ HuffLayout L, int16_t* __restrict__ dc, int n_mcu,
                                                             int dri, int n_int, int nc) {
  const int wid = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= n_int * nc) return;
  const int c = wid / n_int, k = wid - c * n_int;
  const int hv = L.comp_h[c] * L.comp_v[c];
  const int lo = k * dri * hv, hi = min((k + 1) * dri, n_mcu) * hv;
  int16_t* d = dc + L.dc_off[c];
  int carry = 0;
  for (int base = lo; base < hi; base += 32) {
    const int idx = base + lane;
    int v = idx < hi ? d[idx] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    v += carry;
    if (idx < hi) d[idx] = (int16_t)v;
    carry = __shfl_sync(0xffffffffu, v, 31);
  }
}

struct JpegState {
  int16_t* h_coef = nullptr;    // pinned
  int16_t* d_coef = nullptr;
  uint8_t* d_planes = nullptr;
  size_t coef_cap = 0, plane_cap = 0;
  // device entropy decoding
  uint8_t* h_bits = nullptr;    // pinned, unstuffed stream
  uint32_t* d_bits = nullptr;
  size_t bits_cap = 0;
  uint16_t* h_lut = nullptr;    // pinned [4][65536]
  uint16_t* d_lut = nullptr;
  uint8_t lut_key[4][16 + 256]; // the DHT contents the device tables were built from
  bool lut_valid = false;
  unsigned long long* d_exit = nullptr;
  int *d_counts = nullptr, *d_first = nullptr, *d_changed = nullptr, *h_changed = nullptr;
  int16_t* d_dc = nullptr;      // DC differences, compact
  uint32_t* d_subtab = nullptr; // [4][sub_cap]: start bit, end bit, head block, head index
  uint32_t* h_subtab = nullptr; // pinned
  int* d_excl = nullptr;
  size_t sub_cap = 0, dc_cap = 0;
  bool zigzag_up = false;
};

}  // namespace

void jpeg_destroy(ofb_handle* h) {
  JpegState* s = static_cast<JpegState*>(h->jpeg);
  if (!s) return;
  if (s->h_coef) cudaFreeHost(s->h_coef);
  if (s->d_coef) cudaFree(s->d_coef);
  if (s->d_planes) cudaFree(s->d_planes);
  if (s->h_bits) cudaFreeHost(s->h_bits);
  if (s->h_lut) cudaFreeHost(s->h_lut);
  if (s->h_changed) cudaFreeHost(s->h_changed);
  cudaFree(s->d_bits); cudaFree(s->d_lut); cudaFree(s->d_exit); cudaFree(s->d_counts); cudaFree(s->d_first); cudaFree(s->d_changed); cudaFree(s->d_dc); cudaFree(s->d_subtab); cudaFree(s->d_excl);
  if (s->h_subtab) cudaFreeHost(s->h_subtab);
  delete s;
  h->jpeg = nullptr;
}

// Copies the entropy-coded segment without its byte stuffing into `out` (capacity n + 64), zero padded; returns the bytes.
static size_t unstuff_into(const uint8_t* p, size_t n, uint8_t* out, std::vector<size_t>* rst) {
  uint8_t* o = out;
  const uint8_t* end = p + n;
  while (p < end) {
    const uint8_t* q = static_cast<const uint8_t*>(memchr(p, 0xFF, end - p));
    if (!q) q = end;
    memcpy(o, p, q - p);
    o += q - p;
    p = q;
    if (p + 1 >= end) break;
    if (p[1] == 0) { *o++ = 0xFF; p += 2; }
    else if (p[1] == 0xFF) p++;
    else if (p[1] >= 0xD0 && p[1] <= 0xD7) { rst->push_back((size_t)(o - out)); p += 2; }   // the next interval starts here
    else break;                                          // any other marker ends the data
  }
  const size_t used = o - out;
  memset(o, 0, 64);
  return used;
}

// Huffman decoding of the scan on the device: coefficient planes in s->d_coef, DC values in s->d_dc.
static int entropy_decode_device(ofb_handle* h, JpegState* s, const Frame& f) {
  cudaStream_t sm = h->stream;
  if (f.n_data + 64 > s->bits_cap) {
    if (s->h_bits) cudaFreeHost(s->h_bits);
    cudaFree(s->d_bits);
    s->h_bits = nullptr; s->d_bits = nullptr; s->bits_cap = 0;
    const size_t cap = (f.n_data + 64 + 4095) & ~(size_t)4095;
    OFB_CUDA(h, cudaHostAlloc(&s->h_bits, cap, cudaHostAllocDefault));
    OFB_CUDA(h, cudaMalloc(&s->d_bits, cap));
    s->bits_cap = cap;
  }
  if (!s->d_lut) {
    OFB_CUDA(h, cudaHostAlloc(&s->h_lut, 4 * 65536 * sizeof(uint16_t), cudaHostAllocDefault));
    OFB_CUDA(h, cudaMalloc(&s->d_lut, 4 * 65536 * sizeof(uint16_t)));
    OFB_CUDA(h, cudaMalloc(&s->d_changed, 16 * sizeof(int)));
    OFB_CUDA(h, cudaHostAlloc(&s->h_changed, sizeof(int), cudaHostAllocDefault));
  }
  if (!s->zigzag_up) {
    OFB_CUDA(h, cudaMemcpyToSymbolAsync(c_zigzag, kZigzag, 80, 0, cudaMemcpyHostToDevice, sm));
    s->zigzag_up = true;
  }
  static thread_local std::vector<size_t> rst;
  rst.clear();
  const size_t used = unstuff_into(f.data, f.n_data, s->h_bits, &rst);
  const size_t up_bytes = (used + 32 + 3) & ~(size_t)3;
  OFB_CUDA(h, cudaMemcpyAsync(s->d_bits, s->h_bits, up_bytes, cudaMemcpyHostToDevice, sm));
  // full 16-bit code tables, rebuilt only when the stream's DHT segments change (cameras send the same ones every frame)
  const HuffTab* tabs[4] = {&f.dc[0], &f.dc[1], &f.ac[0], &f.ac[1]};
  bool same = s->lut_valid;
  for (int t = 0; t < 4 && same; t++) {
    uint8_t key[16 + 256] = {0};
    if (tabs[t]->present) { memcpy(key, tabs[t]->counts, 16); memcpy(key + 16, tabs[t]->vals, 256); }
    same = memcmp(key, s->lut_key[t], sizeof(key)) == 0;
  }
  if (!same) {
    for (int t = 0; t < 4; t++) {
      uint16_t* lut = s->h_lut + (size_t)t * 65536;
      memset(lut, 0, 65536 * sizeof(uint16_t));
      memset(s->lut_key[t], 0, sizeof(s->lut_key[t]));
      if (!tabs[t]->present) continue;
      memcpy(s->lut_key[t], tabs[t]->counts, 16);
      memcpy(s->lut_key[t] + 16, tabs[t]->vals, 256);
      int code = 0, k = 0;
      for (int len = 1; len <= 16; len++) {
        for (int i = 0; i < tabs[t]->counts[len - 1]; i++, code++, k++) {
          const uint16_t e = (uint16_t)((len << 8) | tabs[t]->vals[k]);
          const int lo = code << (16 - len);
          for (int j = 0; j < (1 << (16 - len)); j++) lut[lo + j] = e;
        }
        code <<= 1;
      }
    }
    OFB_CUDA(h, cudaMemcpyAsync(s->d_lut, s->h_lut, 4 * 65536 * sizeof(uint16_t), cudaMemcpyHostToDevice, sm));
    s->lut_valid = true;
  }
  HuffLayout L = {};
  L.bits = s->d_bits;
  L.total_bits = (uint32_t)(used * 8);
  // restart intervals: [start byte, end byte) of each; without DRI the whole scan is one interval
  const int n_mcu = f.mcux * f.mcuy;
  const int n_int = f.dri ? (n_mcu + f.dri - 1) / f.dri : 1;
  static thread_local std::vector<size_t> ibeg;
  ibeg.assign(1, 0);
  for (size_t k = 0; k < rst.size() && (int)ibeg.size() < n_int; k++) ibeg.push_back(rst[k]);
  while ((int)ibeg.size() < n_int) ibeg.push_back(used);            // missing markers: empty intervals (blocks stay zero)
  ibeg.push_back(used);
  size_t n_sub = 0;
  for (int k = 0; k < n_int; k++) n_sub += std::max<size_t>(1, ((ibeg[k + 1] - ibeg[k]) * 8 + kSubBits - 1) / kSubBits);
  L.n_sub = (int)n_sub;
  int b = 0;
  for (int c = 0; c < f.nc; c++) {
    const Comp& k = f.comp[c];
    L.comp_h[c] = k.h; L.comp_v[c] = k.v; L.comp_blocks_x[c] = k.blocks_x; L.comp_off[c] = k.coef_off;
    for (int j = 0; j < k.h * k.v; j++, b++) {
      if (b >= 6) return set_error(h, OFB_ERR_UNSUPPORTED, "JPEG: more than 6 blocks per MCU");
      L.blk_comp[b] = c; L.blk_j[b] = j; L.blk_dc[b] = k.td; L.blk_ac[b] = k.ta;
      L.blk_tabs |= (uint32_t)(k.td | (k.ta << 1)) << (4 * b);
    }
  }
  L.bpm = b;
  L.dri_blocks = f.dri * b;
  size_t n_dc = 0;
  for (int c = 0; c < f.nc; c++) { L.dc_off[c] = n_dc; n_dc += (size_t)f.mcux * f.mcuy * f.comp[c].h * f.comp[c].v; }
  if (n_dc > s->dc_cap) {
    cudaFree(s->d_dc);
    s->d_dc = nullptr; s->dc_cap = 0;
    OFB_CUDA(h, cudaMalloc(&s->d_dc, n_dc * sizeof(int16_t)));
    s->dc_cap = n_dc;
  }
  L.mcux = f.mcux;
  L.total_blocks = f.mcux * f.mcuy * b;
  L.lut = s->d_lut;
  if ((size_t)L.n_sub > s->sub_cap) {
    cudaFree(s->d_exit); cudaFree(s->d_counts); cudaFree(s->d_first);
    s->d_exit = nullptr; s->d_counts = nullptr; s->d_first = nullptr; s->sub_cap = 0;
    const size_t cap = ((size_t)L.n_sub + 1023) & ~(size_t)1023;
    OFB_CUDA(h, cudaMalloc(&s->d_exit, cap * sizeof(unsigned long long)));
    OFB_CUDA(h, cudaMalloc(&s->d_counts, cap * sizeof(int)));
    OFB_CUDA(h, cudaMalloc(&s->d_first, cap * sizeof(int)));
    cudaFree(s->d_subtab); cudaFree(s->d_excl);
    if (s->h_subtab) cudaFreeHost(s->h_subtab);
    s->d_subtab = nullptr; s->d_excl = nullptr; s->h_subtab = nullptr;
    OFB_CUDA(h, cudaMalloc(&s->d_subtab, 4 * cap * sizeof(uint32_t)));
    OFB_CUDA(h, cudaMalloc(&s->d_excl, cap * sizeof(int)));
    OFB_CUDA(h, cudaHostAlloc(&s->h_subtab, 4 * cap * sizeof(uint32_t), cudaHostAllocDefault));
    s->sub_cap = cap;
  }
  {
    uint32_t* t0 = s->h_subtab;
    uint32_t* t1 = t0 + s->sub_cap;
    int* t2 = reinterpret_cast<int*>(t1 + s->sub_cap);
    int* t3 = t2 + s->sub_cap;
    size_t i = 0;
    for (int k = 0; k < n_int; k++) {
      const uint32_t b0 = (uint32_t)(ibeg[k] * 8), b1 = (uint32_t)(ibeg[k + 1] * 8);
      const size_t m = std::max<size_t>(1, ((size_t)(b1 - b0) + kSubBits - 1) / kSubBits), head = i;
      for (size_t q = 0; q < m; q++, i++) {
        t0[i] = b0 + (uint32_t)(q * kSubBits);
        t1[i] = std::min(b0 + (uint32_t)((q + 1) * kSubBits), b1);
        t2[i] = q == 0 ? k * (f.dri ? f.dri : 0) * L.bpm : -1;
        t3[i] = (int)head;
      }
    }
    OFB_CUDA(h, cudaMemcpyAsync(s->d_subtab, s->h_subtab, 4 * s->sub_cap * sizeof(uint32_t), cudaMemcpyHostToDevice, sm));
    L.sub_start = s->d_subtab;
    L.sub_end = s->d_subtab + s->sub_cap;
    L.sub_head_block = reinterpret_cast<const int*>(s->d_subtab + 2 * s->sub_cap);
    L.sub_headidx = reinterpret_cast<const int*>(s->d_subtab + 3 * s->sub_cap);
  }
  int st;
  if ((st = timing_begin(h, OFB_STAGE_OTHER))) return st;
  OFB_CUDA(h, cudaMemsetAsync(s->d_coef, 0, f.n_coef * sizeof(int16_t), sm));
  OFB_CUDA(h, cudaMemsetAsync(s->d_dc, 0, n_dc * sizeof(int16_t), sm));
  OFB_CUDA(h, cudaMemsetAsync(s->d_changed, 0, 16 * sizeof(int), sm));
  const int ctas = (L.n_sub + kSyncThreads - 1) / kSyncThreads;
  // a CTA settles its own 256 subsequences in one launch; the exit state of a CTA reaches the next CTA in the next
  // launch, from where on that CTA is right too: two launches make every state true in all but pathological streams,
  // the third (or a later one) proves it by changing nothing
  int round = 0;
  for (; round < 3; round++) {
    k_huff_sync<<<ctas, kSyncThreads, 0, sm>>>(L, s->d_exit, s->d_counts, round == 0, s->d_changed + round);
    OFB_LAUNCH_CHECK(h);
  }
  while (true) {
    OFB_CUDA(h, cudaMemcpyAsync(s->h_changed, s->d_changed + (round - 1) % 16, sizeof(int), cudaMemcpyDeviceToHost, sm));
    OFB_CUDA(h, cudaStreamSynchronize(sm));
    if (!*s->h_changed) break;
    if (round > ctas + 3) return set_error(h, OFB_ERR_CUDA, "JPEG: entropy decoder did not settle");
    OFB_CUDA(h, cudaMemsetAsync(s->d_changed + round % 16, 0, sizeof(int), sm));
    k_huff_sync<<<ctas, kSyncThreads, 0, sm>>>(L, s->d_exit, s->d_counts, 0, s->d_changed + round % 16);
    OFB_LAUNCH_CHECK(h);
    round++;
  }
  k_huff_scan<<<1, 1024, 0, sm>>>(s->d_counts, s->d_excl, L.n_sub);
  OFB_LAUNCH_CHECK(h);
  k_huff_rebase<<<(L.n_sub + 255) / 256, 256, 0, sm>>>(L, s->d_excl, s->d_first);
  OFB_LAUNCH_CHECK(h);
  k_huff_write<<<ctas, kSyncThreads, 0, sm>>>(L, s->d_exit, s->d_first, s->d_coef, s->d_dc);
  OFB_LAUNCH_CHECK(h);
  if (f.dri) k_dc_prefix_intervals<<<(n_int * f.nc * 32 + 255) / 256, 256, 0, sm>>>(L, s->d_dc, n_mcu, f.dri, n_int, f.nc);
  else k_dc_prefix<<<f.nc, 1024, 0, sm>>>(L, s->d_dc, n_mcu);
  OFB_LAUNCH_CHECK(h);
  return timing_end(h);
}

// Decodes `jpeg` into device frames on the handle's stream: d_bgr ([height][width][3], pitch bgr_pitch) and/or d_gray.
// Returns after the launches are enqueued; the pinned coefficient staging is reused by the next call, which
// synchronises first.
static int jpeg_decode_device(ofb_handle* h, const Frame& f, uint8_t* d_bgr, size_t bgr_pitch, uint8_t* d_gray, size_t gray_pitch,
                              bool luma_only = false) {
  if (!h->jpeg) h->jpeg = new JpegState();
  JpegState* s = static_cast<JpegState*>(h->jpeg);
  cudaStream_t sm = h->stream;
  size_t plane_bytes = 0, plane_off[kMaxComp], pitch[kMaxComp];
  for (int c = 0; c < f.nc; c++) {
    pitch[c] = ((size_t)f.comp[c].blocks_x * 8 + 15) & ~(size_t)15;
    plane_off[c] = plane_bytes;
    plane_bytes += pitch[c] * f.comp[c].blocks_y * 8;
  }
  OFB_CUDA(h, cudaStreamSynchronize(sm));                 // the previous frame's coefficients may still be in flight
  if (f.n_coef > s->coef_cap) {
    if (s->h_coef) cudaFreeHost(s->h_coef);
    if (s->d_coef) cudaFree(s->d_coef);
    s->h_coef = nullptr; s->d_coef = nullptr; s->coef_cap = 0;
    OFB_CUDA(h, cudaHostAlloc(&s->h_coef, f.n_coef * sizeof(int16_t), cudaHostAllocDefault));
    OFB_CUDA(h, cudaMalloc(&s->d_coef, f.n_coef * sizeof(int16_t)));
    s->coef_cap = f.n_coef;
  }
  if (plane_bytes > s->plane_cap) {
    if (s->d_planes) cudaFree(s->d_planes);
    s->d_planes = nullptr; s->plane_cap = 0;
    OFB_CUDA(h, cudaMalloc(&s->d_planes, plane_bytes));
    s->plane_cap = plane_bytes;
  }
  bool on_device = f.n_data < (1u << 27) && !h->jpeg_host_entropy;
  for (int c = 0; c < f.nc; c++) on_device = on_device && f.comp[c].td <= 1 && f.comp[c].ta <= 1;
  if (on_device) {
    int st = entropy_decode_device(h, s, f);
    if (st) return st;
  } else {
    // (unusual table numbering, or asked for: the host walks the stream)
    memset(s->h_coef, 0, f.n_coef * sizeof(int16_t));
    decode_scan(f, s->h_coef);
    OFB_CUDA(h, cudaMemcpyAsync(s->d_coef, s->h_coef, f.n_coef * sizeof(int16_t), cudaMemcpyHostToDevice, sm));
  }
  IdctArgs ia = {};
  ia.nc = f.nc;
  int first = 0;
  size_t dc_off = 0;
  for (int c = 0; c < f.nc; c++) {
    const Comp& k = f.comp[c];
    IdctPlane& P = ia.p[c];
    P.coef = s->d_coef + k.coef_off;
    P.plane = s->d_planes + plane_off[c];
    P.blocks_x = k.blocks_x;
    P.n_blocks = k.blocks_x * k.blocks_y;
    P.pitch = pitch[c];
    P.first_block = first;
    P.dc = on_device ? s->d_dc + dc_off : nullptr;
    P.h = k.h; P.v = k.v; P.mcux = f.mcux;
    dc_off += (size_t)f.mcux * f.mcuy * k.h * k.v;
    memcpy(P.q, f.qt[k.tq], sizeof(P.q));
    first += P.n_blocks;
  }
  ia.total_blocks = first;
  int st;
  if ((st = timing_begin(h, OFB_STAGE_OTHER))) return st;
  k_jpeg_idct<<<(first + 31) / 32, 256, 0, sm>>>(ia);
  OFB_LAUNCH_CHECK(h);
  ColorArgs ca = {};
  ca.Y = s->d_planes + plane_off[0];
  ca.y_pitch = pitch[0];
  ca.w = f.width; ca.h = f.height;
  ca.gray_only_source = f.nc == 1 || luma_only;   // (B = G = R = Y: the gray weights add up to 2^15, so gray == Y)
  if (f.nc == 3 && !luma_only) {
    ca.Cb = s->d_planes + plane_off[1];
    ca.Cr = s->d_planes + plane_off[2];
    ca.c_pitch = pitch[1];
    ca.cw = f.comp[1].dw; ca.ch = f.comp[1].dh;
    ca.fh = f.hmax / f.comp[1].h; ca.fv = f.vmax / f.comp[1].v;
  }
  ca.bgr = d_bgr; ca.bgr_pitch = bgr_pitch;
  ca.gray = d_gray; ca.gray_pitch = gray_pitch;
  k_jpeg_color<<<dim3(((f.width + 3) / 4 + 255) / 256, f.height), 256, 0, sm>>>(ca);
  OFB_LAUNCH_CHECK(h);
  if ((st = timing_end(h))) return st;
  return OFB_OK;
}

}  // namespace ofb

using namespace ofb;

extern "C" {

int ofb_jpeg_info(const uint8_t* jpeg, size_t n_bytes, int* width, int* height, int* components) {
  if (!jpeg) return OFB_ERR_INVALID_ARG;
  Frame* f = new Frame();
  const char* why = parse_jpeg(jpeg, n_bytes, f, true);
  if (!why) {
    if (width) *width = f->width;
    if (height) *height = f->height;
    if (components) *components = f->nc;
  }
  delete f;
  return why ? OFB_ERR_UNSUPPORTED : OFB_OK;
}

int ofb_jpeg_set_host_entropy(ofb_handle* h, int on) {
  if (!h) return OFB_ERR_INVALID_ARG;
  h->jpeg_host_entropy = on != 0;
  return OFB_OK;
}

int ofb_jpeg_entropy_decode(const uint8_t* jpeg, size_t n_bytes, int16_t* coef, size_t coef_capacity, size_t* n_coef) {
  if (!jpeg) return OFB_ERR_INVALID_ARG;
  Frame* f = new Frame();
  const char* why = parse_jpeg(jpeg, n_bytes, f, false);
  int st = why ? OFB_ERR_UNSUPPORTED : OFB_OK;
  if (!st) {
    if (n_coef) *n_coef = f->n_coef;
    if (coef) {
      if (coef_capacity < f->n_coef) st = OFB_ERR_CAPACITY;
      else { memset(coef, 0, f->n_coef * sizeof(int16_t)); decode_scan(*f, coef); }
    }
  }
  delete f;
  return st;
}

int ofb_jpeg_decode(ofb_handle* h, const uint8_t* jpeg, size_t n_bytes, uint8_t* bgr, size_t bgr_stride_bytes, uint8_t* gray,
                    size_t gray_stride_bytes) {
  if (!h) return OFB_ERR_INVALID_ARG;
  if (!jpeg || (!bgr && !gray)) return set_error(h, OFB_ERR_INVALID_ARG, "NULL pointer");
  Frame* fp = new Frame();
  const char* why = parse_jpeg(jpeg, n_bytes, fp, false);
  if (why) { delete fp; return set_error(h, OFB_ERR_UNSUPPORTED, "JPEG: %s", why); }
  const Frame& f = *fp;
  const size_t row3 = (size_t)f.width * 3, gpitch = ((size_t)f.width + 15) & ~(size_t)15;
  if (bgr_stride_bytes == 0) bgr_stride_bytes = row3;
  if (gray_stride_bytes == 0) gray_stride_bytes = (size_t)f.width;
  int st = OFB_OK;
  if ((bgr && bgr_stride_bytes < row3) || (gray && gray_stride_bytes < (size_t)f.width))
    st = set_error(h, OFB_ERR_INVALID_ARG, "stride smaller than a row");
  if (!st && cudaSetDevice(h->device) != cudaSuccess) st = set_error(h, OFB_ERR_CUDA, "cudaSetDevice failed");
  if (!st) st = ingest_reserve(h, bgr ? row3 * f.height : 0, gray ? gpitch * f.height : 0);
  ofb_handle::Ingest& g = h->ingest;
  if (!st) st = jpeg_decode_device(h, f, bgr ? g.d_a : nullptr, row3, gray ? g.d_b : nullptr, gpitch);
  const int w = f.width, hh = f.height;
  delete fp;
  if (st) return st;
  if (bgr) OFB_CUDA(h, cudaMemcpy2DAsync(bgr, bgr_stride_bytes, g.d_a, row3, row3, hh, cudaMemcpyDeviceToHost, h->stream));
  if (gray) OFB_CUDA(h, cudaMemcpy2DAsync(gray, gray_stride_bytes, g.d_b, gpitch, w, hh, cudaMemcpyDeviceToHost, h->stream));
  OFB_CUDA(h, cudaStreamSynchronize(h->stream));
  return OFB_OK;
}

int ofb_jpeg_decode_luma(ofb_handle* h, const uint8_t* jpeg, size_t n_bytes, uint8_t* gray, size_t gray_stride_bytes) {
  if (!h) return OFB_ERR_INVALID_ARG;
  if (!jpeg || !gray) return set_error(h, OFB_ERR_INVALID_ARG, "NULL pointer");
  Frame* fp = new Frame();
  const char* why = parse_jpeg(jpeg, n_bytes, fp, false);
  if (why) { delete fp; return set_error(h, OFB_ERR_UNSUPPORTED, "JPEG: %s", why); }
  const int w = fp->width, hh = fp->height;
  const size_t gpitch = ((size_t)w + 15) & ~(size_t)15;
  if (gray_stride_bytes == 0) gray_stride_bytes = (size_t)w;
  int st = gray_stride_bytes < (size_t)w ? set_error(h, OFB_ERR_INVALID_ARG, "stride smaller than a row") : OFB_OK;
  if (!st && cudaSetDevice(h->device) != cudaSuccess) st = set_error(h, OFB_ERR_CUDA, "cudaSetDevice failed");
  if (!st) st = ingest_reserve(h, 0, gpitch * hh);
  if (!st) st = jpeg_decode_device(h, *fp, nullptr, 0, h->ingest.d_b, gpitch, true);
  delete fp;
  if (st) return st;
  OFB_CUDA(h, cudaMemcpy2DAsync(gray, gray_stride_bytes, h->ingest.d_b, gpitch, w, hh, cudaMemcpyDeviceToHost, h->stream));
  OFB_CUDA(h, cudaStreamSynchronize(h->stream));
  return OFB_OK;
}

int ofb_ingest_jpeg_gray(ofb_handle* h, const uint8_t* jpeg, size_t n_bytes, uint8_t* dst, int dst_width, int dst_height,
                         size_t dst_stride_bytes) {
  if (!h) return OFB_ERR_INVALID_ARG;
  if (!jpeg || !dst) return set_error(h, OFB_ERR_INVALID_ARG, "NULL pointer");
  if (dst_width < 1 || dst_height < 1) return set_error(h, OFB_ERR_INVALID_ARG, "bad size");
  if (dst_stride_bytes == 0) dst_stride_bytes = (size_t)dst_width;
  if (dst_stride_bytes < (size_t)dst_width) return set_error(h, OFB_ERR_INVALID_ARG, "stride smaller than a row");
  Frame* fp = new Frame();
  const char* why = parse_jpeg(jpeg, n_bytes, fp, false);
  if (why) { delete fp; return set_error(h, OFB_ERR_UNSUPPORTED, "JPEG: %s", why); }
  const Frame& f = *fp;
  const bool same = f.width == dst_width && f.height == dst_height;
  const size_t srow = (size_t)f.width * 3, drow = (size_t)dst_width * 3, gpitch = ((size_t)dst_width + 15) & ~(size_t)15;
  int st = OFB_OK;
  if (cudaSetDevice(h->device) != cudaSuccess) st = set_error(h, OFB_ERR_CUDA, "cudaSetDevice failed");
  if (!st) st = ingest_reserve(h, same ? 0 : srow * f.height, (same ? 0 : drow * dst_height) + gpitch * dst_height);
  ofb_handle::Ingest& g = h->ingest;
  uint8_t* gray = g.d_b;
  if (!st) {
    if (same) {
      st = jpeg_decode_device(h, f, nullptr, 0, gray, gpitch);         // gray straight out of the colour kernel
    } else {
      // the nodes resize the colour frame first (lfn3_sub_node.py:152-153), then convert
      st = jpeg_decode_device(h, f, g.d_a, srow, nullptr, 0);
      uint8_t* small = g.d_b + gpitch * dst_height;
      if (!st) st = resize_u8_device(h, g.d_a, srow, f.width, f.height, 3, small, drow, dst_width, dst_height);
      if (!st) st = cvt_gray_device(h, small, drow, gray, gpitch, dst_width, dst_height, 0);
    }
  }
  delete fp;
  if (st) return st;
  OFB_CUDA(h, cudaMemcpy2DAsync(dst, dst_stride_bytes, gray, gpitch, dst_width, dst_height, cudaMemcpyDeviceToHost, h->stream));
  OFB_CUDA(h, cudaStreamSynchronize(h->stream));
  return OFB_OK;
}

}  // extern "C"
