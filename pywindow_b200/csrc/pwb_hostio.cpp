// pwb_hostio.cpp -- native DL_POLY HISTORY indexer / decoder (host side of the trajectory path).
//
// The reference decodes one frame at a time in Python: mmap slice -> str.split per line -> np.array
// (trajectory.py:217-249, :647-689, :712-766; 1.7 ms per 168-atom frame).  Once the kernels run at
// ~8e5 frames/s that text decode is the end-to-end limit, so it is done here: frames are parsed in
// parallel by host threads straight into the [F, N, 3] float64 batch the GPU path consumes.
// Numbers are converted with Clinger's exact fast path (mantissa < 2^53, |exp10| <= 22, one correctly
// rounded multiply or divide), falling back to strtod -- bit-identical to Python's float().
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

namespace {

const double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                           1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

inline const char* skip_line(const char* p, const char* end) {
  const char* q = (const char*)memchr(p, '\n', (size_t)(end - p));
  return q ? q + 1 : end;
}

// parse one floating point token starting at p (leading blanks skipped); returns pointer past it
inline const char* parse_double(const char* p, const char* end, double* out, bool* ok) {
  while (p < end && (*p == ' ' || *p == '\t')) p++;
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

}  // namespace

extern "C" {

// Byte offsets of every line that starts with "timestep" (DL_POLY frame records).  Returns the
// number found (may exceed cap: call again with a larger array).
int64_t pwb_history_index(const char* buf, int64_t len, int64_t* starts, int64_t cap) {
  int64_t n = 0;
  const char* end = buf + len;
  const char* p = buf;
  while (p < end) {
    if ((size_t)(end - p) >= 8 && memcmp(p, "timestep", 8) == 0) {
      if (n < cap) starts[n] = (int64_t)(p - buf);
      n++;
    }
    p = skip_line(p, end);
  }
  return n;
}

// Decode `n` frames (byte ranges [starts[i], ends[i])) of `natoms` atoms into xyz[n][natoms][3].
// keytrj: 0 coordinates, 1 +velocities, 2 +forces (extra lines skipped); imcon 1..3: three lattice
// lines follow the timestep record and are written to lattice[n][3][3] TRANSPOSED like the reference
// (trajectory.py:723-726).  Returns 0, or 1 + index of the first malformed frame.
int64_t pwb_history_decode(const char* buf, int64_t len, const int64_t* starts, const int64_t* ends, int64_t n,
                           int natoms, int keytrj, int imcon, double* xyz, double* lattice, int threads) {
  std::atomic<int64_t> next(0), bad(0);
  auto work = [&]() {
    while (true) {
      const int64_t f = next.fetch_add(1);
      if (f >= n) break;
      const char* p = buf + starts[f];
      const char* end = buf + (ends[f] <= len ? ends[f] : len);
      bool ok = true;
      p = skip_line(p, end);                                   // "timestep ..." record
      if (imcon >= 1 && imcon <= 3) {
        double m[9];
        for (int r = 0; r < 3 && ok; r++) {
          for (int c = 0; c < 3 && ok; c++) p = parse_double(p, end, &m[3 * r + c], &ok);
          p = skip_line(p, end);
        }
        if (lattice)
          for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) lattice[f * 9 + 3 * r + c] = m[3 * c + r];
      }
      double* o = xyz + (size_t)f * natoms * 3;
      for (int a = 0; a < natoms && ok; a++) {
        p = skip_line(p, end);                                 // atom key / index / mass / charge
        for (int c = 0; c < 3 && ok; c++) p = parse_double(p, end, &o[3 * a + c], &ok);
        p = skip_line(p, end);
        for (int k = 0; k < keytrj; k++) p = skip_line(p, end);
        if (p > end) ok = false;
      }
      if (!ok) { int64_t z = 0; bad.compare_exchange_strong(z, f + 1); }
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

}  // extern "C"
