// Host side of the scorer's file formats (no GPU code): vocabulary hash, n-best text -> packed int32 ids written
// straight into the caller's (pinned) staging buffers, scores -> "lmwt.nn" text.  Replaces the per-hypothesis Python of
// compute_sentence_scores_bayes_jianwei.py:20-120 (load_nbest, read_vocab, get_input_and_target) and :283-303
// (write_scores), which at ~2 M tokens/s on one core was 4x slower than the GPU work it feeds.
//
// Text semantics kept (Python's, as the reference runs them):
//   * a line is stripped of surrounding whitespace; the key is everything before the FIRST ' ', the hypothesis the
//     rest (no space: the hypothesis is ' ', i.e. empty); the utterance is the key up to its LAST '-';
//   * the hypothesis is split on runs of whitespace (str.split()); input = <s> + words, target = words + <s>;
//     a word missing from the vocabulary becomes <unk>;
//   * utterances are numbered in order of first appearance, hypotheses inside an utterance in file order from 1.
// Whitespace here is the ASCII set of str.split() (space, \t-\r, \x1c-\x1f); a file that contains a Unicode-only
// separator (NBSP, U+2028, ...) is reported through `flags` so the caller can take its own slow path.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "blm_host.h"

namespace {

// ASCII whitespace of str.split() / str.strip(): space, \t \n \v \f \r, \x1c-\x1f
struct SpaceTable {
  bool t[256];
  constexpr SpaceTable() : t() {
    for (int c = 0; c < 256; ++c) t[c] = c == ' ' || (c >= 9 && c <= 13) || (c >= 0x1c && c <= 0x1f);
  }
};
constexpr SpaceTable kSpace;
inline bool is_space(unsigned char c) { return kSpace.t[c]; }

inline uint64_t hash_bytes(const char* s, int64_t n) {
  uint64_t h = 0xcbf29ce484222325ull;   // FNV-1a, finished with a multiply-xorshift mix
  for (int64_t i = 0; i < n; ++i) h = (h ^ static_cast<unsigned char>(s[i])) * 0x100000001b3ull;
  h ^= h >> 32;
  h *= 0x9e3779b97f4a7c15ull;
  return h ^ (h >> 29);
}

// open-addressing map from byte strings (copied into an arena) to dense ids in insertion order
struct ByteMap {
  std::vector<int32_t> slot;        // -1 = empty, else entry index
  std::vector<int64_t> begin;       // arena offsets per entry
  std::vector<int32_t> len;
  std::vector<uint64_t> hash;
  std::string arena;
  uint64_t mask = 0;

  explicit ByteMap(size_t expect = 1024) { rehash(expect * 2 + 16); }

  void rehash(size_t want) {
    size_t cap = 16;
    while (cap < want) cap <<= 1;
    slot.assign(cap, -1);
    mask = cap - 1;
    for (size_t e = 0; e < begin.size(); ++e) {
      uint64_t i = hash[e] & mask;
      while (slot[i] >= 0) i = (i + 1) & mask;
      slot[i] = static_cast<int32_t>(e);
    }
  }

  int32_t find(const char* s, int64_t n, uint64_t h) const {
    uint64_t i = h & mask;
    for (;;) {
      const int32_t e = slot[i];
      if (e < 0) return -1;
      if (hash[e] == h && len[e] == n && memcmp(arena.data() + begin[e], s, static_cast<size_t>(n)) == 0) return e;
      i = (i + 1) & mask;
    }
  }
  int32_t find(const char* s, int64_t n) const { return find(s, n, hash_bytes(s, n)); }

  // id of s, inserting it if new
  int32_t intern(const char* s, int64_t n) {
    const uint64_t h = hash_bytes(s, n);
    const int32_t e = find(s, n, h);
    if (e >= 0) return e;
    if ((begin.size() + 1) * 2 > slot.size()) rehash(slot.size() * 2);
    const int32_t id = static_cast<int32_t>(begin.size());
    begin.push_back(static_cast<int64_t>(arena.size()));
    len.push_back(static_cast<int32_t>(n));
    hash.push_back(h);
    arena.append(s, static_cast<size_t>(n));
    uint64_t i = h & mask;
    while (slot[i] >= 0) i = (i + 1) & mask;
    slot[i] = id;
    return id;
  }
};

// Unicode code points str.split() / str.strip() treat as whitespace beyond ASCII, in UTF-8
inline bool unicode_space_at(const unsigned char* p, const unsigned char* end) {
  if (p[0] == 0xC2 && p + 1 < end) return p[1] == 0x85 || p[1] == 0xA0;
  if (p[0] == 0xE1 && p + 2 < end) return p[1] == 0x9A && p[2] == 0x80;
  if (p[0] == 0xE2 && p + 2 < end) {
    if (p[1] == 0x80) return (p[2] >= 0x80 && p[2] <= 0x8A) || p[2] == 0xA8 || p[2] == 0xA9 || p[2] == 0xAF;
    return p[1] == 0x81 && p[2] == 0x9F;
  }
  if (p[0] == 0xE3 && p + 2 < end) return p[1] == 0x80 && p[2] == 0x80;
  return false;
}

struct LineView {
  const char* key;      // stripped line start
  int64_t key_len;      // bytes before the first ' '
  const char* hyp;      // after that space (may equal end)
  const char* end;      // stripped line end
};

inline LineView view_line(const char* b, const char* e) {
  while (b < e && is_space(static_cast<unsigned char>(*b))) ++b;
  while (e > b && is_space(static_cast<unsigned char>(e[-1]))) --e;
  const char* sp = static_cast<const char*>(memchr(b, ' ', static_cast<size_t>(e - b)));
  LineView v;
  v.key = b;
  v.end = e;
  if (sp) {
    v.key_len = sp - b;
    v.hyp = sp + 1;
  } else {
    v.key_len = e - b;
    v.hyp = e;
  }
  return v;
}

// words of str.split(): a word starts wherever a non-space byte follows a space byte (or the start)
inline int64_t count_words(const char* b, const char* e) {
  int64_t n = 0;
  bool prev_space = true;
  for (; b < e; ++b) {
    const bool sp = is_space(static_cast<unsigned char>(*b));
    n += static_cast<int64_t>(prev_space & !sp);
    prev_space = sp;
  }
  return n;
}

// "%.4f" of a float: v * 10000 is exact in double (24-bit significand x a 14-bit integer), so rounding that product
// to the nearest-even integer is the correctly rounded decimal printf produces; huge / non-finite values go to printf
inline int format_score(char* out, float v) {
  const double y = static_cast<double>(v) * 10000.0;
  if (!(y > -9.0e17 && y < 9.0e17)) return snprintf(out, 48, "%.4f", static_cast<double>(v));
  long long r = static_cast<long long>(__builtin_nearbyint(y));   // round-half-even in the default rounding mode
  char tmp[32];
  int n = 0;
  const bool neg = (r < 0) || (r == 0 && __builtin_signbit(y));
  unsigned long long a = r < 0 ? 0ull - static_cast<unsigned long long>(r) : static_cast<unsigned long long>(r);
  for (int i = 0; i < 4; ++i) {
    tmp[n++] = static_cast<char>('0' + a % 10);
    a /= 10;
  }
  tmp[n++] = '.';
  do {
    tmp[n++] = static_cast<char>('0' + a % 10);
    a /= 10;
  } while (a);
  if (neg) tmp[n++] = '-';
  for (int i = 0; i < n; ++i) out[i] = tmp[n - 1 - i];
  return n;
}

}  // namespace

struct blm_vocab {
  ByteMap map;
  int32_t bos = -1, unk = -1;
};

extern "C" {

blm_vocab* blm_vocab_from_text(const char* text, int64_t nbytes) {
  if (!text || nbytes < 0) {
    blm::set_error("blm_vocab_from_text: null text");
    return nullptr;
  }
  blm_vocab* v = new blm_vocab();
  const char* p = text;
  const char* end = text + nbytes;
  int64_t line_no = 0;
  while (p < end) {
    const char* nl = static_cast<const char*>(memchr(p, '\n', static_cast<size_t>(end - p)));
    const char* le = nl ? nl : end;
    ++line_no;
    // fields of line.split(): exactly two (score.py:79 asserts it)
    const char* w0 = nullptr;
    int64_t w0n = 0, fields = 0;
    const char* q = p;
    while (q < le) {
      while (q < le && is_space(static_cast<unsigned char>(*q))) ++q;
      if (q == le) break;
      const char* s = q;
      while (q < le && !is_space(static_cast<unsigned char>(*q))) ++q;
      if (fields == 0) {
        w0 = s;
        w0n = q - s;
      }
      ++fields;
    }
    if (fields != 2) {
      blm::set_error("vocabulary line %lld: expected 'word index', found %lld field(s)", (long long)line_no, (long long)fields);
      delete v;
      return nullptr;
    }
    v->map.intern(w0, w0n);
    p = nl ? nl + 1 : end;
  }
  v->bos = v->map.find("<s>", 3);
  v->unk = v->map.find("<unk>", 5);
  return v;
}

int64_t blm_vocab_size(const blm_vocab* v) { return v ? static_cast<int64_t>(v->map.begin.size()) : 0; }

int32_t blm_vocab_id(const blm_vocab* v, const char* word, int64_t len) {
  return (v && word && len >= 0) ? v->map.find(word, len) : -1;
}

void blm_vocab_free(blm_vocab* v) { delete v; }

/* pass 1 over the n-best text: line starts (byte offsets, n_lines + 1 entries incl. the end) and scored positions per
 * line (words + 1) are written when the arrays are given; returns counts.  flags bit 0: a Unicode-only whitespace
 * character occurs (take the slow path); lines that are empty after stripping still count (the reference keeps them
 * as an empty hypothesis with an empty key).                                                                      */
int blm_nbest_scan(const char* text, int64_t nbytes, int64_t cap_lines, int64_t* line_begin, int32_t* line_tokens,
                   int64_t* n_lines, int64_t* n_tokens, int32_t* flags, int32_t n_threads) {
  BLM_REQUIRE(text && nbytes >= 0 && n_lines && n_tokens, BLM_ERR_ARG, "bad nbest_scan arguments");
  const char* end = text + nbytes;
  // byte ranges that start right after a newline: one per thread
  int nt = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(n_threads, nbytes / (64 << 10) + 1)));
  std::vector<const char*> cut(static_cast<size_t>(nt) + 1);
  cut[0] = text;
  for (int t = 1; t < nt; ++t) {
    const char* q = text + nbytes * t / nt;
    if (q < cut[static_cast<size_t>(t) - 1]) q = cut[static_cast<size_t>(t) - 1];
    const char* nl = static_cast<const char*>(memchr(q, '\n', static_cast<size_t>(end - q)));
    cut[static_cast<size_t>(t)] = nl ? nl + 1 : end;
  }
  cut[static_cast<size_t>(nt)] = end;
  std::vector<int64_t> lines(static_cast<size_t>(nt), 0), tokens(static_cast<size_t>(nt), 0);
  std::vector<int> flagged(static_cast<size_t>(nt), 0);
  // pass A: lines per range, and whether Python would have split the text differently -- it reads the file with
  // universal newlines (a lone '\r' ends a line) and splits on Unicode whitespace; both are looked for only in ranges
  // that hold a '\r' or a non-ASCII byte at all
  auto pass_a = [&](int t) {
    const char* p = cut[static_cast<size_t>(t)];
    const char* e = cut[static_cast<size_t>(t) + 1];
    const int64_t nb = e - p;
    bool suspicious = nb > 0 && memchr(p, '\r', static_cast<size_t>(nb)) != nullptr;
    if (!suspicious) {
      uint64_t acc = 0;
      int64_t i = 0;
      for (; i + 8 <= nb; i += 8) {
        uint64_t w;
        memcpy(&w, p + i, 8);
        acc |= w;
      }
      for (; i < nb; ++i) acc |= static_cast<unsigned char>(p[i]);
      suspicious = (acc & 0x8080808080808080ull) != 0;
    }
    if (suspicious) {
      const unsigned char* uend = reinterpret_cast<const unsigned char*>(end);
      for (const unsigned char* u = reinterpret_cast<const unsigned char*>(p); u < reinterpret_cast<const unsigned char*>(e); ++u)
        if ((*u == '\r' && (u + 1 == uend || u[1] != '\n')) || (*u >= 0xC2 && *u <= 0xE3 && unicode_space_at(u, uend))) {
          flagged[static_cast<size_t>(t)] = 1;
          break;
        }
    }
    int64_t n = 0;
    while (p < e) {
      const char* nl = static_cast<const char*>(memchr(p, '\n', static_cast<size_t>(e - p)));
      ++n;
      p = nl ? nl + 1 : e;
    }
    lines[static_cast<size_t>(t)] = n;
  };
  // pass B: line starts and scored positions (words + 1) per line, written at the range's line offset
  std::vector<int64_t> first(static_cast<size_t>(nt) + 1, 0);
  auto pass_b = [&](int t) {
    const char* p = cut[static_cast<size_t>(t)];
    const char* e = cut[static_cast<size_t>(t) + 1];
    int64_t i = first[static_cast<size_t>(t)], tk = 0;
    while (p < e) {
      const char* nl = static_cast<const char*>(memchr(p, '\n', static_cast<size_t>(e - p)));
      const char* le = nl ? nl : e;
      const LineView v = view_line(p, le);
      const int64_t c = count_words(v.hyp, v.end) + 1;
      if (line_begin) line_begin[i] = p - text;
      if (line_tokens) line_tokens[i] = static_cast<int32_t>(c);
      tk += c;
      ++i;
      p = nl ? nl + 1 : e;
    }
    tokens[static_cast<size_t>(t)] = tk;
  };
  auto run = [&](auto&& fn) {
    if (nt == 1) {
      fn(0);
      return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back(fn, t);
    for (auto& x : th) x.join();
  };
  run(pass_a);
  for (int t = 0; t < nt; ++t) first[static_cast<size_t>(t) + 1] = first[static_cast<size_t>(t)] + lines[static_cast<size_t>(t)];
  const int64_t total = first[static_cast<size_t>(nt)];
  BLM_REQUIRE(!(line_begin || line_tokens) || total <= cap_lines, BLM_ERR_SHAPE, "line arrays too small: %lld lines, room for %lld",
              (long long)total, (long long)cap_lines);
  run(pass_b);
  if (line_begin) line_begin[total] = nbytes;
  int64_t tk = 0;
  int32_t fl = 0;
  for (int t = 0; t < nt; ++t) {
    tk += tokens[static_cast<size_t>(t)];
    fl |= flagged[static_cast<size_t>(t)];
  }
  *n_lines = total;
  *n_tokens = tk;
  if (flags) *flags = fl;
  return BLM_OK;
}

/* pass 2 over lines [l0, l1): ids of line i go to tok / tgt / pos [offs[i] - offs[l0] ...), where offs is the prefix sum
 * of line_tokens (n_lines + 1 entries).  Threads split the range; returns BLM_ERR_ARG if a word is out of vocabulary
 * and the vocabulary has no <unk> (the reference raises KeyError there).                                          */
int blm_nbest_tokenize(const blm_vocab* vocab, const char* text, const int64_t* line_begin, const int64_t* offs, int64_t l0,
                       int64_t l1, int32_t* tok, int32_t* tgt, int32_t* pos, int32_t n_threads) {
  BLM_REQUIRE(vocab && text && line_begin && offs && tok && tgt && l0 >= 0 && l1 >= l0, BLM_ERR_ARG,
              "bad nbest_tokenize arguments");
  BLM_REQUIRE(vocab->bos >= 0, BLM_ERR_ARG, "the vocabulary has no <s>");
  const int64_t base = offs[l0];
  std::vector<int> failed(static_cast<size_t>(std::max(1, n_threads)), 0);
  auto work = [&](int64_t a, int64_t b, int tid) {
    for (int64_t i = a; i < b; ++i) {
      const char* lb = text + line_begin[i];
      const char* le = text + line_begin[i + 1];
      if (le > lb && le[-1] == '\n') --le;
      const LineView v = view_line(lb, le);
      int64_t o = offs[i] - base;
      int32_t k = 0;
      tok[o] = vocab->bos;
      if (pos) pos[o] = 0;
      const char* q = v.hyp;
      while (q < v.end) {
        while (q < v.end && is_space(static_cast<unsigned char>(*q))) ++q;
        if (q == v.end) break;
        const char* s = q;
        while (q < v.end && !is_space(static_cast<unsigned char>(*q))) ++q;
        int32_t id = vocab->map.find(s, q - s);
        if (id < 0) {
          id = vocab->unk;
          if (id < 0) {
            failed[static_cast<size_t>(tid)] = 1;
            id = 0;
          }
        }
        tgt[o + k] = id;
        ++k;
        tok[o + k] = id;
        if (pos) pos[o + k] = k;
      }
      tgt[o + k] = vocab->bos;
    }
  };
  const int64_t n = l1 - l0;
  int nt = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(n_threads, n / 256 + 1)));
  if (nt == 1) {
    work(l0, l1, 0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) {
      // split by tokens, not lines
      const int64_t ta = base + (offs[l1] - base) * t / nt, tb = base + (offs[l1] - base) * (t + 1) / nt;
      const int64_t a = std::lower_bound(offs + l0, offs + l1, ta) - offs;
      const int64_t b = (t + 1 == nt) ? l1 : std::lower_bound(offs + l0, offs + l1, tb) - offs;
      th.emplace_back(work, a, b, t);
    }
    for (auto& x : th) x.join();
  }
  for (int f : failed)
    BLM_REQUIRE(!f, BLM_ERR_ARG, "a word is out of vocabulary and the vocabulary has no <unk>");
  return BLM_OK;
}

/* utterance of every line (dense ids in order of first appearance = the reference's dict order), the 1-based index
 * of the line inside its utterance, and where the utterance key sits in the text (for the output).              */
int blm_nbest_group(const char* text, const int64_t* line_begin, int64_t n_lines, int32_t* utt_of_line, int32_t* idx_in_utt,
                    int64_t* key_begin, int32_t* key_len, int64_t* n_utts) {
  BLM_REQUIRE(text && line_begin && utt_of_line && idx_in_utt && key_begin && key_len && n_utts, BLM_ERR_ARG,
              "bad nbest_group arguments");
  ByteMap groups(static_cast<size_t>(n_lines / 8 + 16));
  std::vector<int32_t> count;
  for (int64_t i = 0; i < n_lines; ++i) {
    const char* lb = text + line_begin[i];
    const char* le = text + line_begin[i + 1];
    if (le > lb && le[-1] == '\n') --le;
    const LineView v = view_line(lb, le);
    // key.rsplit('-', 1)[0]: up to the last '-', or the whole key when there is none
    int64_t kl = v.key_len;
    for (int64_t j = v.key_len - 1; j >= 0; --j)
      if (v.key[j] == '-') {
        kl = j;
        break;
      }
    const int32_t g = groups.intern(v.key, kl);
    if (g == static_cast<int32_t>(count.size())) count.push_back(0);
    utt_of_line[i] = g;
    idx_in_utt[i] = ++count[static_cast<size_t>(g)];
    key_begin[i] = v.key - text;
    key_len[i] = static_cast<int32_t>(kl);
  }
  *n_utts = static_cast<int64_t>(count.size());
  return BLM_OK;
}

/* "<utt>-<idx> %.4f\n" for the lines in `order` (line indices, n of them); returns the bytes written, or the bytes
 * needed (> cap) without writing past cap.                                                                        */
int64_t blm_scores_format(const char* text, const int64_t* key_begin, const int32_t* key_len, const int32_t* idx_in_utt,
                          const int64_t* order, int64_t n, const float* scores, char* out, int64_t cap) {
  int64_t w = 0;
  char num[96];
  for (int64_t r = 0; r < n; ++r) {
    const int64_t i = order ? order[r] : r;
    int m = 0;
    num[m++] = '-';
    {
      char d[16];
      int nd = 0;
      unsigned v = static_cast<unsigned>(idx_in_utt[i] < 0 ? 0 : idx_in_utt[i]);
      do {
        d[nd++] = static_cast<char>('0' + v % 10);
        v /= 10;
      } while (v);
      while (nd) num[m++] = d[--nd];
    }
    num[m++] = ' ';
    m += format_score(num + m, scores[i]);
    num[m++] = '\n';
    const int64_t need = key_len[i] + m;
    if (out && w + need <= cap) {
      memcpy(out + w, text + key_begin[i], static_cast<size_t>(key_len[i]));
      memcpy(out + w + key_len[i], num, static_cast<size_t>(m));
    }
    w += need;
  }
  return w;
}

}  // extern "C"
