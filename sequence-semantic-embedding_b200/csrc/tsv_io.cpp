// targetEncodingIndex.tsv fast writer / reader.  Host-only C++ (no CUDA): the index file is the
// interchange format between sse_index (writer, reference sse_index.py:93-97) and sse_evaluator / demo / webserver
// (readers, reference sse_evaluator.py:79-88).  One row per target:
//     targetId \t raw target text \t E comma-separated float32 decimals \n
// where each decimal is numpy's str(np.float32): the SHORTEST digit string that round-trips to the same float32,
// printed positionally for 1e-4 <= |x| < 1e16 (always with a fractional part, "1.0") and in scientific notation
// otherwise ("1e-05", "1.5e+16", at least two exponent digits); "nan", "inf", "-inf", "-0.0" as numpy prints them.
// std::to_chars(float, scientific) yields exactly that shortest digit string (Ryu in libstdc++); only the layout
// around the digits is numpy's.  The reader restates `line.strip().split('\t')`, skips rows that do not have three
// fields, and parses the decimals with std::from_chars (correctly rounded => the writer's float32 comes back
// bit-exactly).  Both directions are split over host threads by row ranges.
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sse_b200.h"

namespace {

thread_local std::string g_tsv_error;

// digits of |x| (shortest round trip) and decimal exponent: x = 0.d1d2...dn * 10^(e10+1) i.e. d1.d2...dn * 10^e10
inline int shortest_digits(float ax, char* digits, int* e10) {
  char buf[32];
  auto r = std::to_chars(buf, buf + sizeof(buf), ax, std::chars_format::scientific);
  // buf = d[.ddd]e[+-]XX
  int nd = 0;
  char* p = buf;
  while (p < r.ptr && *p != 'e') {
    if (*p != '.') digits[nd++] = *p;
    ++p;
  }
  ++p;   // 'e'
  int sign = 1;
  if (*p == '-') { sign = -1; ++p; } else if (*p == '+') ++p;
  int e = 0;
  while (p < r.ptr) e = e * 10 + (*p++ - '0');
  *e10 = sign * e;
  while (nd > 1 && digits[nd - 1] == '0') --nd;   // "1.0e0" style never comes out of to_chars, but keep the invariant
  return nd;
}

// numpy prints a float32 scalar positionally only while every printed integer digit is significant: 1e-4 <= |x| < 1e6
// ... measured against numpy 2.3 (tests/test_tsv_io.py sweeps all magnitudes); float64 would switch at 1e16.
constexpr float NP_F32_SCI_HI = 1e6f;

// writes str(np.float32(x)) at out, returns the number of bytes
inline int format_np_float32(float x, char* out) {
  char* o = out;
  if (std::isnan(x)) { memcpy(o, "nan", 3); return 3; }
  if (std::signbit(x)) *o++ = '-';
  const float ax = std::fabs(x);
  if (std::isinf(ax)) { memcpy(o, "inf", 3); return (int)(o - out) + 3; }
  if (ax == 0.f) { memcpy(o, "0.0", 3); return (int)(o - out) + 3; }
  char d[16];
  int e10;
  const int nd = shortest_digits(ax, d, &e10);
  if ((double)ax >= 1e-4 && ax < NP_F32_SCI_HI) {   // the lower bound is compared in double, as numpy does: float32(1e-4) < 1e-4
    // positional, trim='0': at least one digit on each side of the point
    if (e10 >= 0) {
      for (int i = 0; i <= e10; ++i) *o++ = i < nd ? d[i] : '0';
      *o++ = '.';
      if (nd > e10 + 1) { for (int i = e10 + 1; i < nd; ++i) *o++ = d[i]; }
      else *o++ = '0';
    } else {
      *o++ = '0'; *o++ = '.';
      for (int i = 0; i < -e10 - 1; ++i) *o++ = '0';
      for (int i = 0; i < nd; ++i) *o++ = d[i];
    }
  } else {
    // scientific, trim='.': "1e-05", "1.5e-05"; exponent sign always, at least two digits
    *o++ = d[0];
    if (nd > 1) { *o++ = '.'; for (int i = 1; i < nd; ++i) *o++ = d[i]; }
    *o++ = 'e';
    *o++ = e10 < 0 ? '-' : '+';
    int ae = e10 < 0 ? -e10 : e10;
    if (ae >= 100) { *o++ = (char)('0' + ae / 100); ae %= 100; }
    *o++ = (char)('0' + ae / 10);
    *o++ = (char)('0' + ae % 10);
  }
  return (int)(o - out);
}

constexpr int MAX_FLOAT_CHARS = 16;   // "-1.2345678e-38," is 15

template <class F>
void parallel_ranges(int64_t n, int threads, F&& fn) {
  if (threads < 1) threads = 1;
  if ((int64_t)threads > n) threads = (int)(n > 0 ? n : 1);
  if (threads == 1) { fn(0, (int64_t)0, n); return; }
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t) {
    const int64_t lo = n * t / threads, hi = n * (t + 1) / threads;
    pool.emplace_back([&fn, t, lo, hi] { fn(t, lo, hi); });
  }
  for (auto& th : pool) th.join();
}

inline bool is_py_space(unsigned char c) {   // str.strip() on the ASCII range
  return c == ' ' || (c >= 9 && c <= 13) || c == 0x1c || c == 0x1d || c == 0x1e || c == 0x1f;
}

}  // namespace

extern "C" {

// CRC-32C (Castagnoli), the checksum of TensorFlow's table blocks and tensor-bundle entries (tf_bundle.py): table-driven,
// 8 bytes per step (slicing-by-8), no ISA extensions required.
uint32_t sse_crc32c(const void* data, size_t n, uint32_t seed) {
  static uint32_t T[8][256];
  static bool ready = false;
  if (!ready) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      T[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int t = 1; t < 8; ++t) T[t][i] = (T[t - 1][i] >> 8) ^ T[0][T[t - 1][i] & 0xFF];
    ready = true;
  }
  const unsigned char* p = static_cast<const unsigned char*>(data);
  uint32_t crc = ~seed;
  while (n >= 8) {
    uint32_t lo, hi;
    memcpy(&lo, p, 4); memcpy(&hi, p + 4, 4);
    lo ^= crc;
    crc = T[7][lo & 0xFF] ^ T[6][(lo >> 8) & 0xFF] ^ T[5][(lo >> 16) & 0xFF] ^ T[4][lo >> 24] ^
          T[3][hi & 0xFF] ^ T[2][(hi >> 8) & 0xFF] ^ T[1][(hi >> 16) & 0xFF] ^ T[0][hi >> 24];
    p += 8; n -= 8;
  }
  while (n--) crc = T[0][(crc ^ *p++) & 0xFF] ^ (crc >> 8);
  return ~crc;
}

const char* sse_tsv_last_error(void) { return g_tsv_error.c_str(); }

int sse_tsv_format_f32(const float* values, int64_t n, char* out, size_t cap, int64_t* ends) {
  if (!values || !out || n < 0) { g_tsv_error = "sse_tsv_format_f32: bad argument"; return SSE_EINVAL; }
  size_t used = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (used + MAX_FLOAT_CHARS > cap) { g_tsv_error = "sse_tsv_format_f32: output buffer too small"; return SSE_EINVAL; }
    used += (size_t)format_np_float32(values[i], out + used);
    if (ends) ends[i] = (int64_t)used;
  }
  return SSE_OK;
}

int sse_tsv_write_index(const char* path, const char* const* ids, const char* const* texts, const float* rows, int64_t n_rows, int E,
                        int append, int threads) {
  if (!path || !ids || !texts || !rows || n_rows < 0 || E < 1) { g_tsv_error = "sse_tsv_write_index: bad argument"; return SSE_EINVAL; }
  FILE* f = fopen(path, append ? "ab" : "wb");
  if (!f) { g_tsv_error = std::string("sse_tsv_write_index: cannot open ") + path; return SSE_EINVAL; }
  if (threads < 1) threads = (int)std::thread::hardware_concurrency();
  // slabs of 16k rows: the per-thread text buffers are allocated once and reused (no repeated first-touch page
  // faults), and formatting of slab s+1 does not start before slab s is on its way to the page cache
  const int64_t slab = 16384;
  const int nt_max = (int)std::max<int64_t>(1, std::min<int64_t>(threads, std::min<int64_t>(slab, n_rows)));
  std::vector<std::string> parts((size_t)nt_max);
  for (int64_t r0 = 0; r0 < n_rows; r0 += slab) {
    const int64_t nr = std::min<int64_t>(slab, n_rows - r0);
    const int nt = (int)std::min<int64_t>(nt_max, nr);
    parallel_ranges(nr, nt, [&](int t, int64_t lo, int64_t hi) {
      std::string& s = parts[(size_t)t];
      s.clear();
      char tmp[MAX_FLOAT_CHARS + 2];
      for (int64_t r = r0 + lo; r < r0 + hi; ++r) {
        s.append(ids[r]); s.push_back('\t'); s.append(texts[r]); s.push_back('\t');
        const float* v = rows + (size_t)r * E;
        for (int e = 0; e < E; ++e) {
          int len = format_np_float32(v[e], tmp);
          if (e + 1 < E) tmp[len++] = ',';
          s.append(tmp, (size_t)len);
        }
        s.push_back('\n');
      }
    });
    for (int t = 0; t < nt; ++t)
      if (fwrite(parts[(size_t)t].data(), 1, parts[(size_t)t].size(), f) != parts[(size_t)t].size()) {
        fclose(f);
        g_tsv_error = "sse_tsv_write_index: short write";
        return SSE_EINVAL;
      }
  }
  fclose(f);
  return SSE_OK;
}

// spans: [max_rows, 4] = (id_begin, id_end, text_begin, text_end) byte offsets into buf for every ACCEPTED row.
int sse_tsv_parse_index(const char* buf, size_t len, int E, int64_t max_rows, float* out, int64_t* spans, int64_t* n_rows,
                        int64_t* n_skipped, int threads) {
  if (!buf || !out || !spans || !n_rows || E < 1 || max_rows < 0) { g_tsv_error = "sse_tsv_parse_index: bad argument"; return SSE_EINVAL; }
  if (threads < 1) threads = (int)std::thread::hardware_concurrency();
  // pass 1: line starts (readlines() semantics: split at '\n'; a trailing fragment without newline is a line too)
  std::vector<size_t> starts;
  starts.reserve(len / 512 + 16);
  size_t p = 0;
  while (p < len) {
    starts.push_back(p);
    const void* nl = memchr(buf + p, '\n', len - p);
    p = nl ? (size_t)((const char*)nl - buf) + 1 : len;
  }
  const int64_t n_lines = (int64_t)starts.size();
  starts.push_back(len);
  // pass 2 (parallel): validate + parse every line into a per-line slot; rejected lines are compacted afterwards
  std::vector<uint8_t> ok((size_t)n_lines, 0);
  std::vector<int64_t> sp((size_t)n_lines * 4, 0);
  if (n_lines > max_rows) { g_tsv_error = "sse_tsv_parse_index: more lines than max_rows"; return SSE_EINVAL; }
  std::vector<int> bad_counts((size_t)std::max(1, threads), 0);
  std::string first_error;
  std::vector<std::string> errs((size_t)std::max(1, threads));
  parallel_ranges(n_lines, threads, [&](int t, int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; ++i) {
      size_t b = starts[(size_t)i], e = starts[(size_t)i + 1];
      while (b < e && is_py_space((unsigned char)buf[b])) ++b;          // line.strip()
      while (e > b && is_py_space((unsigned char)buf[e - 1])) --e;
      const char* t1 = (const char*)memchr(buf + b, '\t', e - b);
      const char* t2 = t1 ? (const char*)memchr(t1 + 1, '\t', (size_t)(buf + e - (t1 + 1))) : nullptr;
      const char* t3 = t2 ? (const char*)memchr(t2 + 1, '\t', (size_t)(buf + e - (t2 + 1))) : nullptr;
      if (!t1 || !t2 || t3) { ++bad_counts[(size_t)t]; continue; }     // len(info) != 3
      const char* q = t2 + 1;
      const char* qe = buf + e;
      float* dst = out + (size_t)i * E;
      int cnt = 0;
      bool fail = false;
      while (q <= qe) {
        const char* c = (const char*)memchr(q, ',', (size_t)(qe - q));
        const char* fe = c ? c : qe;
        const char* fb = q;
        while (fb < fe && is_py_space((unsigned char)*fb)) ++fb;       // float() tolerates surrounding blanks
        const char* ft = fe;
        while (ft > fb && is_py_space((unsigned char)ft[-1])) --ft;
        if (fb < ft && *fb == '+') ++fb;                              // from_chars rejects a leading '+'
        float v = 0.f;
        auto r = std::from_chars(fb, ft, v);
        if (r.ec == std::errc::result_out_of_range) {                 // float() saturates: inf / 0 / denormal
          char tmpf[64];
          const size_t L = (size_t)(ft - fb) < sizeof(tmpf) - 1 ? (size_t)(ft - fb) : sizeof(tmpf) - 1;
          memcpy(tmpf, fb, L); tmpf[L] = 0;
          v = strtof(tmpf, nullptr);
        } else if (r.ec != std::errc() || r.ptr != ft) { fail = true; break; }
        if (cnt < E) dst[cnt] = v;
        ++cnt;
        if (!c) break;
        q = c + 1;
      }
      if (fail || cnt != E) {
        if (errs[(size_t)t].empty()) errs[(size_t)t] = "sse_tsv_parse_index: line " + std::to_string(i + 1) + (fail ? ": not a float" : ": wrong number of values");
        continue;
      }
      ok[(size_t)i] = 1;
      int64_t* s4 = &sp[(size_t)i * 4];
      s4[0] = (int64_t)b; s4[1] = (int64_t)(t1 - buf); s4[2] = (int64_t)(t1 + 1 - buf); s4[3] = (int64_t)(t2 - buf);
    }
  });
  for (auto& e : errs)
    if (!e.empty()) { g_tsv_error = e; return SSE_EINVAL; }   // a row with three fields but a broken vector is an error, as float()/np.array would raise
  // compact accepted rows in file order
  int64_t w = 0;
  for (int64_t i = 0; i < n_lines; ++i) {
    if (!ok[(size_t)i]) continue;
    if (w != i) memmove(out + (size_t)w * E, out + (size_t)i * E, (size_t)E * sizeof(float));
    memcpy(spans + (size_t)w * 4, &sp[(size_t)i * 4], 4 * sizeof(int64_t));
    ++w;
  }
  *n_rows = w;
  if (n_skipped) { int64_t b = 0; for (int c : bad_counts) b += c; *n_skipped = b; }
  return SSE_OK;
}

}  // extern "C"
