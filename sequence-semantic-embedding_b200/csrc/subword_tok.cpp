// Batched subword tokenizer + padder.  Host-only C++ (no CUDA): turns raw (already lower-cased)
// utf-8 sentences into the left-padded int32 [B,T] token rows the encoders consume.  Restates the ENCODE half of the
// reference's Tensor2Tensor-style pipeline for a vocabulary.txt the reference wrote:
//   * tokenizer.encode (tokenizer.py:68-90): split at alphanumeric / non-alphanumeric boundaries (Unicode general
//     category L* or N*), a lone space between two tokens is dropped unless it starts the text;
//   * SubwordTextEncoder._escape_token (text_encoder.py:334-356): '\\' -> "\\\\", '_' -> "\\u", characters outside
//     the vocabulary alphabet (and '\n') -> "\\<code point>;", then a trailing '_';
//   * _escaped_token_to_subtoken_ids (text_encoder.py:491-532): greedy longest match against the subtoken set,
//     longest candidate first (lengths in characters, not bytes);
//   * the row rule of data_utils.py:149-155 / sse_index.py:79-85: [PAD]*(T-len-1) + ids + [EOS], or
//     [PAD] + ids[:T-2] + [EOS] when the sentence is too long.
// The python implementation (text_encoder.py, pinned against the real reference encoder on the golden corpus) stays
// the single-sentence path; this is the bulk path (index build, corpus preparation), split over host threads by rows.
#include <cstdint>
#include <cstring>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/sse_b200.h"
#include "unicode_alnum.inc"

namespace {

thread_local std::string g_tok_error;

struct Bitmap {
  std::vector<uint64_t> w;
  Bitmap() : w(0x110000 / 64, 0) {}
  void set(uint32_t c) { w[c >> 6] |= 1ull << (c & 63); }
  bool get(uint32_t c) const { return c < 0x110000 && ((w[c >> 6] >> (c & 63)) & 1); }
};

// lenient utf-8 decoder: malformed bytes decode as U+FFFD one byte at a time (python would have rejected the string
// before it got here; the entry points read their files as utf-8)
inline uint32_t decode(const unsigned char* s, size_t n, size_t* adv) {
  const unsigned char c = s[0];
  if (c < 0x80) { *adv = 1; return c; }
  int len = (c >= 0xF0 && c < 0xF8) ? 4 : (c >= 0xE0) ? 3 : (c >= 0xC0) ? 2 : 0;
  if (len == 0 || (size_t)len > n) { *adv = 1; return 0xFFFD; }
  uint32_t cp = c & (0xFF >> (len + 1));
  for (int i = 1; i < len; ++i) {
    if ((s[i] & 0xC0) != 0x80) { *adv = 1; return 0xFFFD; }
    cp = (cp << 6) | (s[i] & 0x3F);
  }
  *adv = (size_t)len;
  return cp;
}

}  // namespace

struct sse_tokenizer {
  std::vector<std::string> strings;                         // subtoken id -> utf-8
  std::unordered_map<std::string_view, int32_t> ids;        // views into `strings`
  Bitmap alnum, alphabet;
  int maxlen = 0;                                           // longest subtoken, in characters

  // appends the subtoken ids of one token (utf-8 bytes [b,e)) to out; false if some piece is not in the vocabulary
  bool encode_token(const unsigned char* b, const unsigned char* e, std::string& esc, std::vector<uint32_t>& off,
                    std::vector<int32_t>& out) const {
    esc.clear();
    off.clear();
    auto push_ascii = [&](char c) { off.push_back((uint32_t)esc.size()); esc.push_back(c); };
    for (const unsigned char* p = b; p < e;) {
      size_t adv;
      const uint32_t cp = decode(p, (size_t)(e - p), &adv);
      if (cp == '\\') { push_ascii('\\'); push_ascii('\\'); }
      else if (cp == '_') { push_ascii('\\'); push_ascii('u'); }
      else if (alphabet.get(cp) && cp != '\n') { off.push_back((uint32_t)esc.size()); esc.append(reinterpret_cast<const char*>(p), adv); }
      else {
        char tmp[16];
        const int n = snprintf(tmp, sizeof(tmp), "\\%u;", cp);
        for (int i = 0; i < n; ++i) push_ascii(tmp[i]);
      }
      p += adv;
    }
    push_ascii('_');
    off.push_back((uint32_t)esc.size());
    const int n = (int)off.size() - 1;                      // characters in the escaped token
    int start = 0;
    while (start < n) {
      int end = std::min(n, start + maxlen);
      bool found = false;
      for (; end > start; --end) {
        auto it = ids.find(std::string_view(esc.data() + off[(size_t)start], off[(size_t)end] - off[(size_t)start]));
        if (it != ids.end()) { out.push_back(it->second); start = end; found = true; break; }
      }
      if (!found) return false;
    }
    return true;
  }

  bool encode_text(const char* text, std::string& esc, std::vector<uint32_t>& off, std::vector<int32_t>& out) const {
    out.clear();
    const unsigned char* s = reinterpret_cast<const unsigned char*>(text);
    const size_t n = strlen(text);
    if (n == 0) return true;
    size_t adv;
    size_t start = 0, pos = 0;
    bool prev = alnum.get(decode(s, n, &adv));
    pos = adv;
    while (pos < n) {
      const bool cur = alnum.get(decode(s + pos, n - pos, &adv));
      if (cur != prev) {
        // a single space between two tokens is dropped, except at the very beginning of the text
        if (!(pos - start == 1 && s[start] == ' ' && start != 0))
          if (!encode_token(s + start, s + pos, esc, off, out)) return false;
        start = pos;
        prev = cur;
      }
      pos += adv;
    }
    return encode_token(s + start, s + n, esc, off, out);
  }
};

extern "C" {

const char* sse_tok_last_error(void) { return g_tok_error.c_str(); }

int sse_tok_create(const char* const* subtokens_utf8, int n_subtokens, sse_tokenizer** out) {
  if (!subtokens_utf8 || n_subtokens < 2 || !out) { g_tok_error = "sse_tok_create: bad argument"; return SSE_EINVAL; }
  auto* t = new sse_tokenizer();
  for (auto& r : kAlnumRanges)
    for (uint32_t c = r[0]; c <= r[1]; ++c) t->alnum.set(c);
  t->strings.reserve((size_t)n_subtokens);
  for (int i = 0; i < n_subtokens; ++i) t->strings.emplace_back(subtokens_utf8[i] ? subtokens_utf8[i] : "");
  for (int i = 0; i < n_subtokens; ++i) {
    const std::string& s = t->strings[(size_t)i];
    if (s.empty()) continue;
    t->ids[std::string_view(s)] = i;                        // later duplicates win, as the python dict comprehension does
    int chars = 0;
    const unsigned char* p = reinterpret_cast<const unsigned char*>(s.data());
    for (size_t k = 0; k < s.size();) {
      size_t adv;
      t->alphabet.set(decode(p + k, s.size() - k, &adv));
      k += adv;
      ++chars;
    }
    if (chars > t->maxlen) t->maxlen = chars;
  }
  for (const char* c = "\\_u;0123456789"; *c; ++c) t->alphabet.set((uint32_t)(unsigned char)*c);
  *out = t;
  return SSE_OK;
}

int sse_tok_destroy(sse_tokenizer* t) {
  delete t;
  return SSE_OK;
}

int sse_tok_vocab_size(const sse_tokenizer* t) { return t ? (int)t->strings.size() : 0; }

// rows: [n, T] int32, padded by the reference's row rule (T > 2); lengths[i] = number of subtokens BEFORE padding /
// truncation (so the caller can issue the reference's "too long" warning).  T == 0: no rows are written, only lengths.
int sse_tok_encode_batch(const sse_tokenizer* t, const char* const* texts_utf8, int64_t n, int T, int32_t* rows, int32_t* lengths,
                         int threads) {
  if (!t || !texts_utf8 || n < 0 || (T != 0 && (T < 3 || !rows))) { g_tok_error = "sse_tok_encode_batch: bad argument"; return SSE_EINVAL; }
  if (threads < 1) threads = (int)std::thread::hardware_concurrency();
  if ((int64_t)threads > n) threads = (int)(n > 0 ? n : 1);
  std::vector<int64_t> bad((size_t)threads, -1);
  auto work = [&](int tid, int64_t lo, int64_t hi) {
    std::string esc;
    std::vector<uint32_t> off;
    std::vector<int32_t> ids;
    for (int64_t i = lo; i < hi; ++i) {
      if (!texts_utf8[i] || !t->encode_text(texts_utf8[i], esc, off, ids)) { if (bad[(size_t)tid] < 0) bad[(size_t)tid] = i; ids.clear(); }
      const int len = (int)ids.size();
      if (lengths) lengths[i] = len;
      if (T == 0) continue;
      int32_t* r = rows + (size_t)i * T;
      if (len > T - 2) {
        r[0] = 0;
        memcpy(r + 1, ids.data(), (size_t)(T - 2) * sizeof(int32_t));
        r[T - 1] = 1;
      } else {
        const int lead = T - len - 1;
        for (int k = 0; k < lead; ++k) r[k] = 0;
        if (len) memcpy(r + lead, ids.data(), (size_t)len * sizeof(int32_t));
        r[T - 1] = 1;
      }
    }
  };
  if (threads == 1) work(0, 0, n);
  else {
    std::vector<std::thread> pool;
    for (int k = 0; k < threads; ++k) pool.emplace_back(work, k, n * k / threads, n * (k + 1) / threads);
    for (auto& th : pool) th.join();
  }
  for (int64_t b : bad)
    if (b >= 0) { g_tok_error = "sse_tok_encode_batch: text " + std::to_string(b) + ": token substring not found in subtoken vocabulary"; return SSE_EINVAL; }
  return SSE_OK;
}

}  // extern "C"
