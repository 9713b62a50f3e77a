// pwb_hostio.cpp -- native DL_POLY HISTORY indexer / decoder (host side of the trajectory path).
//
// The reference decodes one frame at a time in Python: mmap slice -> str.split per line -> np.array
// (trajectory.py:217-249, :647-689, :712-766; 1.7 ms per 168-atom frame).  Once the kernels run at
// ~8e5 frames/s that text decode is the end-to-end limit, so it is done here: frames are parsed in
// parallel by host threads straight into the [F, N, 3] float64 batch the GPU path consumes.
// Numbers are converted with Clinger's exact fast path (mantissa < 2^53, |exp10| <= 22, one correctly
// rounded multiply or divide), falling back to strtod -- bit-identical to Python's float().
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <atomic>
#include <thread>
#include <vector>
#if defined(__SSE2__)
#include <immintrin.h>
#endif

namespace {

const double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                           1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

inline const char* skip_line(const char* p, const char* end) {
  const char* q = (const char*)memchr(p, '\n', (size_t)(end - p));
  return q ? q + 1 : end;
}

// eight ASCII digits packed in a little-endian word: test and convert without a loop
inline bool all_digits8(uint64_t w) {
  return (((w & 0xF0F0F0F0F0F0F0F0ull) | (((w + 0x0606060606060606ull) & 0xF0F0F0F0F0F0F0F0ull) >> 4)) == 0x3333333333333333ull);
}
inline uint32_t parse_digits8(uint64_t w) {
  w = (w & 0x0F0F0F0F0F0F0F0Full) * 2561 >> 8;
  w = (w & 0x00FF00FF00FF00FFull) * 6553601 >> 16;
  return (uint32_t)((w & 0x0000FFFF0000FFFFull) * 42949672960001ull >> 32);
}

inline bool is_digit(char c) { return (unsigned char)(c - '0') < 10; }

// Fast path for DL_POLY's own records, "[-]d.ddddE[+-]dd" (%12.4E): five significant digits and a two-digit
// exponent -> mantissa < 2^53 and |exp10| <= 22 + 4, i.e. Clinger's exact case, one correctly rounded multiply or
// divide.  Returns nullptr if the token does not have exactly this shape (the generic parser takes over).
inline const char* parse_e4(const char* p, const char* end, double* out) {
  if (end - p < 12) return nullptr;
  const bool neg = *p == '-';
  if (neg || *p == '+') p++;
  if (!(is_digit(p[0]) && p[1] == '.' && is_digit(p[2]) && is_digit(p[3]) && is_digit(p[4]) && is_digit(p[5]) &&
        (p[6] == 'E' || p[6] == 'e') && (p[7] == '-' || p[7] == '+') && is_digit(p[8]) && is_digit(p[9]))) return nullptr;
  if (!(p[10] == ' ' || p[10] == '\n' || p[10] == '\r' || p[10] == '\t')) return nullptr;
  const uint64_t mant = (uint64_t)(p[0] - '0') * 10000 + (p[2] - '0') * 1000 + (p[3] - '0') * 100 + (p[4] - '0') * 10 + (p[5] - '0');
  int e = (p[8] - '0') * 10 + (p[9] - '0');
  e = (p[7] == '-' ? -e : e) - 4;
  if (e < -22 || e > 22) return nullptr;
  double v = (double)mant;
  v = e < 0 ? v / kPow10[-e] : v * kPow10[e];
  *out = neg ? -v : v;
  return p + 10;
}

// The same record read from its END, for the template decoder below: e points one past the last exponent digit.  In
// a fixed-width file the numbers are right aligned, so every token ends where the template frame has it whatever the
// sign.  The frame's SHAPE has been checked byte by byte against the template before (frame_shape_ok: digits where
// digits belong, '.', 'E', blanks and line ends in place, sign bytes within [' ', '-'] and ['+', '-']); what is left
// here is the exact test of the two sign bytes, the exponent range and the conversion -- branch free, `bad`
// accumulates and the caller tests it once per frame.
// value = (mantissa * kMul10[k]) / kDiv10[k], k = exp10 + 26: one of the two factors is 1, so there is exactly one
// rounding (Clinger's exact case) and no branch on the sign of the exponent
#define P10_ONES_26 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1
const double kDiv10[64] = {1e26, 1e25, 1e24, 1e23, 1e22, 1e21, 1e20, 1e19, 1e18, 1e17, 1e16, 1e15, 1e14, 1e13, 1e12, 1e11,
                           1e10, 1e9, 1e8, 1e7, 1e6, 1e5, 1e4, 1e3, 1e2, 1e1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                           1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
const double kMul10[64] = {P10_ONES_26, 1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15,
                           1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
inline unsigned parse_e4_end(const char* e, double* out) {
  uint64_t w;
  uint16_t ex;
  memcpy(&w, e - 10, 8);                                        // "d.ddddE+"
  memcpy(&ex, e - 2, 2);                                        // "dd"
  const unsigned sg = (unsigned char)e[-11];                    // in [0x20, 0x2D]: ' ', '+', '-' are valid
  const uint64_t x = w ^ 0x2B45303030302E30ull;                 // digits -> 0..9, '.', 'E', '+' -> 0, ',' -> 7, '-' -> 6
  const unsigned s7 = (unsigned)(x >> 56);
  unsigned bad = (s7 & 1u) | (~(0x2801u >> ((sg - 0x20u) & 15u)) & 1u);
  const unsigned exd = ex ^ 0x3030u;
  const uint64_t t = (x >> 16) & 0x0F0F0F0Full;                 // d2 d3 d4 d5
  const uint64_t u = t * 10 + (t >> 8);                         // byte 0: d2 d3, byte 2: d4 d5
  const uint64_t mant = (x & 0xF) * 10000 + (u & 0xFF) * 100 + ((u >> 16) & 0xFF);
  const int ea = (int)((exd & 0xF) * 10 + (exd >> 8));
  const int neg_e = -(int)(s7 >> 2);                            // s7 is 0 or 6 -> 0 or -1
  const int e10 = ((ea ^ neg_e) - neg_e) - 4;                   // conditional negate without a branch
  bad |= (unsigned)(e10 + 22) > 44u;
  const unsigned k = (unsigned)(e10 + 26) & 63u;
  const double v = ((double)(int64_t)mant * kMul10[k]) / kDiv10[k];
  uint64_t bits;
  memcpy(&bits, &v, 8);
  bits |= (uint64_t)(sg == '-') << 63;                          // (v >= 0 here; "-0.0000E+00" -> -0.0 like float())
  memcpy(out, &bits, 8);
  return bad;
}

// lo[i] <= p[i] <= hi[i] for every byte of a frame (unsigned): 16 or 32 bytes per step
inline bool frame_shape_ok_scalar(const unsigned char* p, const unsigned char* lo, const unsigned char* hi, int64_t i, int64_t n) {
  unsigned bad = 0;
  for (; i < n; i++) bad |= (unsigned)(p[i] < lo[i]) | (unsigned)(p[i] > hi[i]);
  return bad == 0;
}
#if defined(__SSE2__)
#if defined(__GNUC__) && defined(__x86_64__)
__attribute__((target("avx2"))) bool frame_shape_ok_avx2(const unsigned char* p, const unsigned char* lo,
                                                         const unsigned char* hi, int64_t n) {
  __m256i acc = _mm256_set1_epi8((char)0xFF);
  int64_t i = 0;
  for (; i + 32 <= n; i += 32) {
    const __m256i v = _mm256_loadu_si256((const __m256i*)(p + i));
    const __m256i l = _mm256_loadu_si256((const __m256i*)(lo + i)), h = _mm256_loadu_si256((const __m256i*)(hi + i));
    acc = _mm256_and_si256(acc, _mm256_and_si256(_mm256_cmpeq_epi8(_mm256_max_epu8(v, l), v),
                                                 _mm256_cmpeq_epi8(_mm256_min_epu8(v, h), v)));
  }
  return _mm256_movemask_epi8(acc) == -1 && frame_shape_ok_scalar(p, lo, hi, i, n);
}
#define PWB_HAVE_AVX2_DISPATCH 1
#endif
inline bool frame_shape_ok_sse2(const unsigned char* p, const unsigned char* lo, const unsigned char* hi, int64_t n) {
  __m128i acc = _mm_set1_epi8((char)0xFF);
  int64_t i = 0;
  for (; i + 16 <= n; i += 16) {
    const __m128i v = _mm_loadu_si128((const __m128i*)(p + i));
    const __m128i l = _mm_loadu_si128((const __m128i*)(lo + i)), h = _mm_loadu_si128((const __m128i*)(hi + i));
    acc = _mm_and_si128(acc, _mm_and_si128(_mm_cmpeq_epi8(_mm_max_epu8(v, l), v), _mm_cmpeq_epi8(_mm_min_epu8(v, h), v)));
  }
  return _mm_movemask_epi8(acc) == 0xFFFF && frame_shape_ok_scalar(p, lo, hi, i, n);
}
#endif
inline bool frame_shape_ok(const unsigned char* p, const unsigned char* lo, const unsigned char* hi, int64_t n) {
#if defined(PWB_HAVE_AVX2_DISPATCH)
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (avx2) return frame_shape_ok_avx2(p, lo, hi, n);
#endif
#if defined(__SSE2__)
  return frame_shape_ok_sse2(p, lo, hi, n);
#else
  return frame_shape_ok_scalar(p, lo, hi, 0, n);
#endif
}

// number of line feeds in [p, p + n)
inline int64_t count_newlines(const char* p, int64_t n) {
  int64_t nl = 0, i = 0;
#if defined(__SSE2__)
  // compare 16 bytes at a time, keep per-byte counts (0xFF = -1 subtracted) and fold them with PSADBW before they
  // can overflow: two instructions per 16 bytes, no popcount (the host compiler may not assume POPCNT)
  const __m128i lf = _mm_set1_epi8('\n'), zero = _mm_setzero_si128();
  while (i + 16 <= n) {
    __m128i acc = zero;
    int64_t blocks = (n - i) / 16;
    if (blocks > 255) blocks = 255;
    for (int64_t b = 0; b < blocks; b++, i += 16)
      acc = _mm_sub_epi8(acc, _mm_cmpeq_epi8(_mm_loadu_si128((const __m128i*)(p + i)), lf));
    const __m128i sum = _mm_sad_epu8(acc, zero);
    nl += _mm_cvtsi128_si64(sum) + _mm_cvtsi128_si64(_mm_unpackhi_epi64(sum, sum));
  }
#endif
  for (; i < n; i++) nl += p[i] == '\n';
  return nl;
}

// parse one floating point token starting at p (leading blanks skipped); returns pointer past it
inline const char* parse_double(const char* p, const char* end, double* out, bool* ok) {
  while (p < end && (*p == ' ' || *p == '\t')) p++;
  if (const char* q = parse_e4(p, end, out)) return q;
  const char* start = p;
  bool neg = false;
  if (p < end && (*p == '-' || *p == '+')) { neg = *p == '-'; p++; }
  uint64_t mant = 0;
  int digits = 0, exp10 = 0;
  bool any = false, exact = true;
  while (p < end && *p >= '0' && *p <= '9') {
    if (digits < 19) { mant = mant * 10 + (uint64_t)(*p - '0'); if (mant || digits) digits += (mant != 0); }
    else { exp10++; if (*p != '0') exact = false; }
    p++; any = true;
  }
  if (p < end && *p == '.') {
    p++;
    // eight fraction digits at a time (SWAR) once the mantissa is non-zero, i.e. every digit is significant
    while (digits > 0 && digits <= 11 && end - p >= 8) {
      uint64_t w;
      memcpy(&w, p, 8);
      if (!all_digits8(w)) break;
      mant = mant * 100000000ull + parse_digits8(w);
      digits += 8; exp10 -= 8; p += 8; any = true;
    }
    while (p < end && *p >= '0' && *p <= '9') {
      if (digits < 19) { mant = mant * 10 + (uint64_t)(*p - '0'); digits += (mant != 0); exp10--; }
      else if (*p != '0') exact = false;
      p++; any = true;
    }
  }
  if (!any) { *ok = false; return p; }
  if (p < end && (*p == 'e' || *p == 'E' || *p == 'd' || *p == 'D')) {
    const char* q = p + 1;
    bool eneg = false;
    if (q < end && (*q == '-' || *q == '+')) { eneg = *q == '-'; q++; }
    if (q < end && *q >= '0' && *q <= '9') {
      int e = 0;
      while (q < end && *q >= '0' && *q <= '9') { if (e < 10000) e = e * 10 + (*q - '0'); q++; }
      exp10 += eneg ? -e : e;
      p = q;
    }
  }
  double v;
  if (exact && mant < (1ull << 53) && exp10 >= -22 && exp10 <= 22) {
    v = (double)mant;
    v = exp10 < 0 ? v / kPow10[-exp10] : v * kPow10[exp10];
  } else {
    char tmp[64];
    size_t n = (size_t)(p - start) < sizeof(tmp) - 1 ? (size_t)(p - start) : sizeof(tmp) - 1;
    memcpy(tmp, start, n);
    tmp[n] = 0;
    for (size_t i = 0; i < n; i++) if (tmp[i] == 'd' || tmp[i] == 'D') tmp[i] = 'e';
    v = strtod(tmp, nullptr);
    neg = false;   // strtod saw the sign
  }
  *out = neg ? -v : v;
  return p;
}

// run one_frame(f) for f in [0, n) on a pool of host threads; 0 or 1 + index of the first failing frame
template <class F>
static int64_t decode_frames(int64_t n, int threads, F one_frame) {
  std::atomic<int64_t> next(0), bad(0);
  auto work = [&]() {
    while (true) {
      const int64_t f = next.fetch_add(1);
      if (f >= n) break;
      if (!one_frame(f)) { int64_t z = 0; bad.compare_exchange_strong(z, f + 1); }
    }
  };
  int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if ((int64_t)nt > n) nt = (int)(n > 0 ? n : 1);
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; t++) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
  return bad.load();
}

}  // namespace

extern "C" {

// Byte offsets of every line that starts with "timestep" (DL_POLY frame records).  Returns the
// number found (may exceed cap: call again with a larger array).
int64_t pwb_history_index(const char* buf, int64_t len, int64_t* starts, int64_t cap) {
  // Fixed-width files (every DL_POLY writer): all frames have the byte length of the first one.  If a "timestep"
  // record really starts at every multiple of that length -- and the file ends on one -- the map is arithmetic and
  // the 1.4 GB of a 100 000-frame file need not be scanned.  Anything else falls through to the full scan.
  {
    const char* end = buf + len;
    const char* p = buf;
    const char* first = nullptr;
    const char* second = nullptr;
    while (p < end && p - buf < (int64_t)1 << 20) {            // header is two short lines
      if ((size_t)(end - p) >= 8 && memcmp(p, "timestep", 8) == 0) { first = p; break; }
      p = skip_line(p, end);
    }
    if (first) {
      p = skip_line(first, end);
      while (p < end) {
        if ((size_t)(end - p) >= 8 && memcmp(p, "timestep", 8) == 0) { second = p; break; }
        p = skip_line(p, end);
      }
    }
    if (first && second) {
      const int64_t flen = second - first, body = len - (first - buf);
      if (flen > 8 && body % flen == 0) {
        const int64_t n = body / flen;
        // (one touch per frame = one page fault per frame on a fresh mapping: spread over the host threads)
        std::atomic<int> irregular(0);
        int vt = (int)std::thread::hardware_concurrency();
        vt = vt < 1 ? 1 : (vt > 32 ? 32 : vt);
        if (n < 4096) vt = 1;
        auto verify = [&](int t) {
          for (int64_t k = n * t / vt; k < n * (t + 1) / vt && !irregular.load(std::memory_order_relaxed); k++) {
            const char* q = first + k * flen;
            if (!(memcmp(q, "timestep", 8) == 0 && q[-1] == '\n')) irregular.store(1);
          }
        };
        {
          std::vector<std::thread> pool;
          for (int t = 1; t < vt; t++) pool.emplace_back(verify, t);
          verify(0);
          for (auto& th : pool) th.join();
        }
        const bool regular = irregular.load() == 0;
        // a frame must not hide a second "timestep" line: the atom count of the record fixes the line count
        if (regular) {
          for (int64_t k = 0; k < n; k++) if (k < cap) starts[k] = (first - buf) + k * flen;
          return n;
        }
      }
    }
  }
  // the file is cut into one range per host thread; a thread owns the lines that START inside its range
  int nt = (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if (nt > 64) nt = 64;
  if (len < (int64_t)(1 << 22)) nt = 1;
  std::vector<std::vector<int64_t>> found((size_t)nt);
  auto scan = [&](int t) {
    const char* end = buf + len;
    const char* lo = buf + len * t / nt;
    const char* hi = buf + len * (t + 1) / nt;
    const char* p = lo;
    if (t > 0 && p[-1] != '\n') p = skip_line(p, end);         // first line start at or after lo
    while (p < hi) {
      if ((size_t)(end - p) >= 8 && memcmp(p, "timestep", 8) == 0) found[(size_t)t].push_back((int64_t)(p - buf));
      p = skip_line(p, end);
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; t++) pool.emplace_back(scan, t);
  scan(0);
  for (auto& th : pool) th.join();
  int64_t n = 0;
  for (auto& v : found)
    for (int64_t s : v) { if (n < cap) starts[n] = s; n++; }
  return n;
}

// Decode `n` frames (byte ranges [starts[i], ends[i])) of `natoms` atoms into xyz[n][natoms][3].
// keytrj: 0 coordinates, 1 +velocities, 2 +forces (extra lines skipped); imcon 1..3: three lattice
// lines follow the timestep record and are written to lattice[n][3][3] TRANSPOSED like the reference
// (trajectory.py:723-726).  Returns 0, or 1 + index of the first malformed frame.
int64_t pwb_history_decode(const char* buf, int64_t len, const int64_t* starts, const int64_t* ends, int64_t n,
                           int natoms, int keytrj, int imcon, double* xyz, double* lattice, int threads) {
  // generic decoder of one frame: walks the lines; optionally records where every coordinate token starts and
  // where every line of the frame ends (relative to the frame start) -- the TEMPLATE of a fixed-width file
  auto generic = [&](int64_t f, std::vector<int32_t>* tok, std::vector<int32_t>* eol) -> bool {
    const char* p0 = buf + starts[f];
    const char* p = p0;
    const char* end = buf + (ends[f] <= len ? ends[f] : len);
    bool ok = true;
    auto line_done = [&](const char* q) { if (eol) eol->push_back((int32_t)(q - p0)); };
    p = skip_line(p, end); line_done(p);                     // "timestep ..." record
    if (imcon >= 1 && imcon <= 3) {
      double m[9];
      for (int r = 0; r < 3 && ok; r++) {
        for (int c = 0; c < 3 && ok; c++) p = parse_double(p, end, &m[3 * r + c], &ok);
        p = skip_line(p, end); line_done(p);
      }
      if (lattice)
        for (int r = 0; r < 3; r++)
          for (int c = 0; c < 3; c++) lattice[f * 9 + 3 * r + c] = m[3 * c + r];
    }
    double* o = xyz + (size_t)f * natoms * 3;
    for (int a = 0; a < natoms && ok; a++) {
      p = skip_line(p, end); line_done(p);                   // atom key / index / mass / charge
      for (int c = 0; c < 3 && ok; c++) {
        while (p < end && (*p == ' ' || *p == '\t')) p++;
        if (tok) tok->push_back((int32_t)(p - p0));
        p = parse_double(p, end, &o[3 * a + c], &ok);
      }
      p = skip_line(p, end); line_done(p);
      for (int k = 0; k < keytrj; k++) { p = skip_line(p, end); line_done(p); }
      if (p > end) ok = false;
    }
    return ok;
  };
  // Template of the first frame, if the file looks fixed width (all requested frames have its byte length and its
  // coordinate tokens have the %12.4E shape): every other frame is decoded by reading the 3*natoms tokens at the
  // template's END offsets (right-aligned numbers end in the same column whatever their sign) -- no line scanning.
  // A frame qualifies only if every byte lies in the range the template allows at its position: a newline at every
  // line end and nowhere else, " [ +-]d.ddddE[+-]dd" + the template's own terminator around every token end, anything
  // but a line break in between (frame_shape_ok; the sign bytes are tested exactly in parse_e4_end) -- any other frame
  // takes the generic path.
  std::vector<int32_t> tok, eol;
  std::vector<unsigned char> shape_lo, shape_hi;
  bool templ = false;
  const bool has_lat = imcon >= 1 && imcon <= 3;           // lattice rows are always parsed generically
  if (n > 1 && ends[0] - starts[0] < (int64_t)1 << 30) {
    std::vector<double> scratch;
    templ = generic(0, &tok, &eol) && (int)tok.size() == 3 * natoms;
    if (templ) {
      const char* p0 = buf + starts[0];
      double v;
      for (size_t i = 0; i < tok.size() && templ; i++) {
        const char* q = parse_e4(p0 + tok[i], buf + ends[0], &v);
        templ = q != nullptr && tok[i] >= 2;
        if (templ) tok[i] = (int32_t)(q - p0);              // from here on: where the token ENDS
      }
      templ = templ && !eol.empty() && eol.back() == ends[0] - starts[0];
      if (templ) {
        const int64_t L = ends[0] - starts[0];
        shape_lo.assign((size_t)L + 32, 0x0B);                // default: anything above '\n' (no stray line breaks)
        shape_hi.assign((size_t)L + 32, 0xFF);
        auto lit = [&](int64_t i, unsigned char c) { shape_lo[(size_t)i] = shape_hi[(size_t)i] = c; };
        auto range = [&](int64_t i, unsigned char a, unsigned char b) { shape_lo[(size_t)i] = a; shape_hi[(size_t)i] = b; };
        for (int32_t q : eol) lit(q - 1, '\n');
        for (size_t i = 0; i < tok.size() && templ; i++) {
          const int64_t e = tok[i];
          const char term = p0[e], sg = p0[e - 11];
          templ = e >= 12 && e < L && p0[e - 12] == ' ' && (sg == ' ' || sg == '-' || sg == '+') && p0[e - 4] == 'E';
          if (!templ) break;
          lit(e - 12, ' ');
          range(e - 11, 0x20, 0x2D);
          range(e - 10, '0', '9');
          lit(e - 9, '.');
          for (int k = 5; k <= 8; k++) range(e - k, '0', '9');
          lit(e - 4, 'E');
          range(e - 3, '+', '-');
          range(e - 2, '0', '9');
          range(e - 1, '0', '9');
          lit(e, (unsigned char)term);                        // ' ' between tokens, the line end after the last
        }
      }
    }
  }
  const int64_t flen = templ ? ends[0] - starts[0] : 0;
  std::atomic<int64_t> next(0), bad(0);
  auto work = [&]() {
    while (true) {
      const int64_t f = next.fetch_add(1);
      if (f >= n) break;
      bool done = false;
      if (templ && f > 0 && ends[f] - starts[f] == flen && ends[f] <= len) {
        const char* p0 = buf + starts[f];
        const char* end = p0 + flen;
        double* o = xyz + (size_t)f * natoms * 3;
        bool ok = true;
        if (has_lat) {
          const char* p = skip_line(p0, end);
          double m[9];
          for (int r = 0; r < 3 && ok; r++) {
            for (int c = 0; c < 3 && ok; c++) p = parse_double(p, end, &m[3 * r + c], &ok);
            p = skip_line(p, end);
          }
          if (ok && lattice)
            for (int r = 0; r < 3; r++)
              for (int c = 0; c < 3; c++) lattice[f * 9 + 3 * r + c] = m[3 * c + r];
        }
        ok = ok && frame_shape_ok((const unsigned char*)p0, shape_lo.data(), shape_hi.data(), flen);
        if (ok) {
          unsigned badtok = 0;
          const int32_t* te = tok.data();
          const size_t nt3 = tok.size();
          for (size_t i = 0; i < nt3; i++) badtok |= parse_e4_end(p0 + te[i], &o[i]);
          ok = badtok == 0;
        }
        done = ok;
      }
      if (!done && !generic(f, nullptr, nullptr)) { int64_t z = 0; bad.compare_exchange_strong(z, f + 1); }
    }
  };
  int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if ((int64_t)nt > n) nt = (int)(n > 0 ? n : 1);
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; t++) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
  return bad.load();
}


// ---- multi-frame XYZ (trajectory.py:792-915): natoms / comment / natoms lines "El x y z" -------------------
// Frame map: byte range and atom count of every frame (blank lines between frames are skipped).  Returns the
// number of frames found (may exceed cap: call again with larger arrays), or -(1 + byte offset) of a line
// that should hold an atom count and does not.
int64_t pwb_xyz_index(const char* buf, int64_t len, int64_t* starts, int64_t* ends, int32_t* natoms, int64_t cap) {
  const char* end = buf + len;
  const char* p = buf;
  int64_t n = 0;
  while (p < end) {
    const char* q = p;
    while (q < end && (*q == ' ' || *q == '\t' || *q == '\r')) q++;
    if (q >= end) break;
    if (*q == '\n') { p = q + 1; continue; }                   // blank line
    int64_t na = 0; int digits = 0;
    while (q < end && *q >= '0' && *q <= '9') { na = na * 10 + (*q - '0'); q++; digits++; if (na > 2000000000) break; }
    if (!digits || (q < end && *q != ' ' && *q != '\t' && *q != '\r' && *q != '\n')) return -(1 + (int64_t)(p - buf));
    const char* e = skip_line(p, end);                          // count line
    for (int64_t i = 0; i < na + 1 && e < end; i++) e = skip_line(e, end);   // comment + atoms
    if (n < cap) { starts[n] = (int64_t)(p - buf); ends[n] = (int64_t)(e - buf); natoms[n] = (int32_t)na; }
    n++;
    p = e;
  }
  return n;
}

// Decode n XYZ frames of `natoms` atoms each into xyz[n][natoms][3].  Returns 0, or 1 + index of the first
// malformed frame.
int64_t pwb_xyz_decode(const char* buf, int64_t len, const int64_t* starts, const int64_t* ends, int64_t n, int natoms,
                       double* xyz, int threads) {
  return decode_frames(n, threads, [&](int64_t f) {
    const char* p = buf + starts[f];
    const char* end = buf + (ends[f] <= len ? ends[f] : len);
    p = skip_line(p, end);                                      // atom count
    p = skip_line(p, end);                                      // comment
    double* o = xyz + (size_t)f * natoms * 3;
    bool ok = true;
    for (int a = 0; a < natoms && ok; a++) {
      if (p >= end) return false;
      while (p < end && (*p == ' ' || *p == '\t')) p++;
      while (p < end && *p != ' ' && *p != '\t' && *p != '\n') p++;          // element / atom id token
      for (int c = 0; c < 3 && ok; c++) p = parse_double(p, end, &o[3 * a + c], &ok);
      p = skip_line(p, end);
    }
    return ok;
  });
}

// Decode n PDB models (trajectory.py:1013-1045; fixed columns 31-38, 39-46, 47-54 of ATOM / HETATM records)
// of `natoms` atoms each into xyz[n][natoms][3].  Returns 0, or 1 + index of the first malformed frame
// (wrong atom count, short record, unparsable field).
int64_t pwb_pdb_decode(const char* buf, int64_t len, const int64_t* starts, const int64_t* ends, int64_t n, int natoms,
                       double* xyz, int threads) {
  return decode_frames(n, threads, [&](int64_t f) {
    const char* p = buf + starts[f];
    const char* end = buf + (ends[f] <= len ? ends[f] : len);
    double* o = xyz + (size_t)f * natoms * 3;
    int a = 0;
    bool ok = true;
    while (p < end && ok) {
      const char* e = skip_line(p, end);
      if (e - p >= 6 && (memcmp(p, "ATOM  ", 6) == 0 || memcmp(p, "HETATM", 6) == 0)) {
        if (a >= natoms || e - p < 54) return false;
        for (int c = 0; c < 3 && ok; c++) {
          const char* q = parse_double(p + 30 + 8 * c, p + 38 + 8 * c, &o[3 * a + c], &ok);
          while (ok && q < p + 38 + 8 * c) { if (*q != ' ') ok = false; q++; }   // trailing junk inside the field
        }
        a++;
      }
      p = e;
    }
    return ok && a == natoms;
  });
}

// Measurement utility: line feeds in [buf, buf + len) counted by `threads` host threads (SSE2 compare + popcount,
// ~1 instruction per 5 bytes) -- the rate at which this host can merely STREAM a mapped trajectory file, i.e. the
// ceiling of any text decoder on it (bench.py reports it beside the decoder's own rate).
int64_t pwb_stream_probe(const char* buf, int64_t len, int threads) {
  int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  std::vector<int64_t> part((size_t)nt, 0);
  auto work = [&](int t) {
    const int64_t lo = len * t / nt, hi = len * (t + 1) / nt;
    part[(size_t)t] = count_newlines(buf + lo, hi - lo);
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; t++) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
  int64_t n = 0;
  for (int64_t v : part) n += v;
  return n;
}

// Write a DL_POLY HISTORY file (keytrj 0; imcon 0, or 3 with `lattice`) in the record layout the readers above
// and the reference's DLPOLY class (trajectory.py:647-766) parse: fixed-width records, coordinates as %20.10E so a
// float64 survives the round trip to 11 significant digits (compact != 0: DL_POLY's own %12.4E records, the
// format of the reference's examples/data/input/HISTORY_singlemol_short).  Every frame has the same byte length, so host threads
// format disjoint frame ranges and write them at their own offsets.  keys: natoms x 8 characters (blank padded);
// lattice: NULL, [9] (lattice_per_frame = 0) or [n_frames][9], row-major, written transposed like the reference
// reads it (trajectory.py:723-726).  Returns the file size in bytes, or -1.
int64_t pwb_history_write(const char* path, const char* title, const char* keys, const double* masses,
                          const double* xyz, int64_t n_frames, int natoms, const double* lattice,
                          int lattice_per_frame, int compact, int threads) {
  const int imcon = lattice ? 3 : 0;
  const int cw = compact ? 12 : 20;                             // DL_POLY's own %12.4E records, or %20.10E
  char head[512];
  const int head_len = snprintf(head, sizeof head, "%.200s\n%10d%10d%10d\n", title, 0, imcon, natoms);
  const int64_t frame_len = 61 + (lattice ? 3 * 61 : 0) + (int64_t)natoms * (43 + 3 * cw + 1);
  const int fd = open(path, O_CREAT | O_TRUNC | O_WRONLY, 0644);
  if (fd < 0) return -1;
  if (write(fd, head, (size_t)head_len) != head_len) { close(fd); return -1; }
  std::atomic<int64_t> next(0);
  std::atomic<int> failed(0);
  const int64_t block = 64;                                     // frames per work item
  auto work = [&]() {
    std::vector<char> buf((size_t)(block * frame_len + 8));
    while (!failed.load()) {
      const int64_t f0 = next.fetch_add(block);
      if (f0 >= n_frames) break;
      const int64_t f1 = f0 + block < n_frames ? f0 + block : n_frames;
      char* p = buf.data();
      for (int64_t f = f0; f < f1; f++) {
        p += sprintf(p, "timestep%10lld%10d%10d%10d%12.6f\n", (long long)(f + 1), natoms, 0, imcon, 0.0007);
        if (lattice) {
          const double* m = lattice + (lattice_per_frame ? f * 9 : 0);
          for (int r = 0; r < 3; r++) p += sprintf(p, "%20.10f%20.10f%20.10f\n", m[r], m[3 + r], m[6 + r]);
        }
        const double* x = xyz + (size_t)f * natoms * 3;
        for (int a = 0; a < natoms; a++) {
          memcpy(p, keys + 8 * a, 8); p += 8;
          p += sprintf(p, "%10d%12.6f%12.6f\n", a + 1, masses ? masses[a] : 1.0, 0.0);
          p += compact ? sprintf(p, "%12.4E%12.4E%12.4E\n", x[3 * a], x[3 * a + 1], x[3 * a + 2])
                       : sprintf(p, "%20.10E%20.10E%20.10E\n", x[3 * a], x[3 * a + 1], x[3 * a + 2]);
        }
      }
      const int64_t want = (f1 - f0) * frame_len;
      if (p - buf.data() != want ||        // a value wider than its field would shift every later frame
          pwrite(fd, buf.data(), (size_t)want, (off_t)(head_len + f0 * frame_len)) != want) failed.store(1);
    }
  };
  int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; t++) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
  close(fd);
  return failed.load() ? -1 : head_len + n_frames * frame_len;
}

}  // extern "C"
