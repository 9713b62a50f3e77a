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
#include <emmintrin.h>
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

// number of line feeds in [p, p + n)
inline int64_t count_newlines(const char* p, int64_t n) {
  int64_t nl = 0, i = 0;
#if defined(__SSE2__)
  const __m128i lf = _mm_set1_epi8('\n');
  for (; i + 16 <= n; i += 16)
    nl += __builtin_popcount((unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128((const __m128i*)(p + i)), lf)));
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
  // template's offsets -- no line scanning.  A frame qualifies only if a newline sits at every line end of the
  // template, nowhere else between them... (checked per token: the byte before a token is blank and parse_e4
  // accepts exactly 10-11 characters followed by a blank or newline) -- anything else takes the generic path.
  std::vector<int32_t> tok, eol;
  bool templ = false;
  const bool has_lat = imcon >= 1 && imcon <= 3;           // lattice rows are always parsed generically
  if (n > 1 && ends[0] - starts[0] < (int64_t)1 << 30) {
    std::vector<double> scratch;
    templ = generic(0, &tok, &eol) && (int)tok.size() == 3 * natoms;
    if (templ) {
      const char* p0 = buf + starts[0];
      double v;
      for (size_t i = 0; i < tok.size() && templ; i++) templ = parse_e4(p0 + tok[i], buf + ends[0], &v) != nullptr;
      templ = templ && !eol.empty() && eol.back() == ends[0] - starts[0];
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
        for (size_t i = 0; i < eol.size() && ok; i++) ok = p0[eol[i] - 1] == '\n';
        for (size_t i = 0; i < tok.size() && ok; i++) {
          const char* t = p0 + tok[i];
          ok = (t[-1] == ' ' || t[-1] == '\n') && parse_e4(t, end, &o[i]) != nullptr;
        }
        // the tokens and line ends sit where the template has them; what lies between them must not hold another
        // line break (a shorter record followed by junk): count the newlines of the frame
        if (ok) {
          ok = count_newlines(p0, flen) == (int64_t)eol.size();
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
