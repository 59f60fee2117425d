// host_io.cpp -- native SPLIT parser and SEGMENT formatter (host side of libfreddie_b200.so).
//
// Replaces, for the CLI, the reference's regex-based read_split (freddie_segment.py:121-171),
// read_sequence (:174-185) and the row formatting of run_segment (:715-731).  The grammar accepted is
// the one the reference's regexes accept (:17-38); a line that the reference would fail to match, or
// any of its asserts (:136-140, :158-164, :181, :666-668), aborts the batch with a message naming it.
// Tints are parsed in parallel (one file pair per task) and concatenated into one packed batch.
#include <errno.h>
#include <fcntl.h>
#include <immintrin.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/freddie_b200.h"

#include <chrono>

namespace {

// FRS_HOST_PROFILE=1: thread-seconds per parser phase on stderr after every frs_parse_tints
struct HostProf {
  std::atomic<long long> ns[6];
  bool on;
  HostProf() : on(getenv("FRS_HOST_PROFILE") != nullptr) { for (auto& x : ns) x = 0; }
};
HostProf g_prof;
struct ProfScope {
  int k;
  std::chrono::steady_clock::time_point t0;
  explicit ProfScope(int k_) : k(k_) { if (g_prof.on) t0 = std::chrono::steady_clock::now(); }
  ~ProfScope() {
    if (g_prof.on) g_prof.ns[k] += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
  }
};
enum { PROF_LOAD = 0, PROF_SPLIT = 1, PROF_READS_LINES = 2, PROF_PLANES = 3, PROF_CONCAT = 4 };

struct ReadMeta {
  int64_t rid;
  int64_t tint;       // the row's own tint column (printed back)
  uint32_t name_off, name_len;
  uint32_t chr_off, chr_len;
  char strand;
};

struct TintData {
  std::string chr;
  int64_t id = 0;
  int64_t read_count = 0;
  std::vector<int32_t> isl_s, isl_e;       // islands
  std::vector<int32_t> isl_off;            // tint-local sample offsets, size n_isl+1
  std::vector<int32_t> rep_iv_off{0}, rep_w, rep_fs, rep_fe;
  std::vector<int32_t> read_rep, read_len, read_iv_off{0};
  std::vector<uint8_t> read_strand;
  std::vector<int64_t> read_seq_off{0};
  std::vector<int32_t> riv_ts, riv_te, riv_qs, riv_qe, riv_cig_off{0};
  std::vector<uint32_t> cigar, seq_a, seq_t;
  std::vector<ReadMeta> meta;
  std::string text;  // names and chr strings of the rows
  std::string error;
};

// A whole input file, read-only.  Small files (the typical tint) are read into a buffer the thread
// keeps and reuses: no page faults, no process-wide mmap lock with thousands of files in flight on all
// threads.  Large files (giant tints) are mapped from the page cache instead of copied.  The parsers
// only look at [p, p + n) -- nothing relies on a terminator.
struct FileBuf {
  char* p = nullptr;
  size_t n = 0;
  bool mapped = false;
  static constexpr size_t SMALL = 8u << 20;
  ~FileBuf() {
    if (mapped) munmap(p, n);
  }
  static std::vector<char>& scratch(int which) {
    static thread_local std::vector<char> buf[2];
    return buf[which];
  }
  // `which`: 0 = split file, 1 = reads file (both may be alive in one thread at the same time)
  bool load(const char* path, std::string& err, int which) {
    ProfScope ps(PROF_LOAD);
    int fd = open(path, O_RDONLY);
    if (fd < 0) { err = std::string("FileNotFoundError: ") + path + ": " + strerror(errno); return false; }
    struct stat st;
    if (fstat(fd, &st) != 0) { err = std::string("OSError: ") + path + ": " + strerror(errno); close(fd); return false; }
    n = (size_t)st.st_size;
    if (n > SMALL) {
      void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
      if (m != MAP_FAILED) {
        p = (char*)m;
        mapped = true;
        madvise(p, n, MADV_SEQUENTIAL);
        close(fd);
        return true;
      }
    }
    std::vector<char>& buf = scratch(which);
    if (buf.size() < n + 1) buf.resize(n + 1);
    p = buf.data();
    size_t got = 0;
    while (got < n) {
      ssize_t r = read(fd, p + got, n - got);
      if (r < 0 && errno == EINTR) continue;
      if (r <= 0) break;
      got += (size_t)r;
    }
    n = got;
    close(fd);
    return true;
  }
};

// isA / isT bit-planes of one read: bit k of word w = (seq[32 w + k] == 'A' / 'T'), exact upper-case
// compare like the reference's  base == 'A'  (freddie_segment.py:352-367).
static void planes_scalar(const char* b, uint32_t L, uint32_t* pa, uint32_t* pt) {
  const uint32_t nw = (L + 31) / 32;
  for (uint32_t w = 0; w < nw; ++w) {
    const uint32_t n = std::min<uint32_t>(32, L - w * 32);
    const char* c = b + (size_t)w * 32;
    uint32_t ma = 0, mt = 0;
    for (uint32_t k = 0; k < n; ++k) {
      ma |= (uint32_t)(c[k] == 'A') << k;
      mt |= (uint32_t)(c[k] == 'T') << k;
    }
    pa[w] = ma;
    pt[w] = mt;
  }
}
// 32 bases per step: byte compare + movemask IS the plane word
__attribute__((target("avx2"))) static void planes_avx2(const char* b, uint32_t L, uint32_t* pa, uint32_t* pt) {
  const __m256i vA = _mm256_set1_epi8('A'), vT = _mm256_set1_epi8('T');
  const uint32_t full = L / 32;
  for (uint32_t w = 0; w < full; ++w) {
    const __m256i v = _mm256_loadu_si256((const __m256i*)(b + (size_t)w * 32));
    pa[w] = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, vA));
    pt[w] = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, vT));
  }
  if (L % 32) planes_scalar(b + (size_t)full * 32, L % 32, pa + full, pt + full);
}
static void seq_planes(const char* b, uint32_t L, uint32_t* pa, uint32_t* pt) {
  static const bool have_avx2 = __builtin_cpu_supports("avx2");
  if (have_avx2) planes_avx2(b, L, pa, pt);
  else planes_scalar(b, L, pa, pt);
}

inline bool is_digit(char c) { return c >= '0' && c <= '9'; }
inline bool chr_first(unsigned char c) {
  return (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || strchr("!#$%&+./:;?@^_|~-", c) != nullptr;
}
inline bool chr_rest(unsigned char c) { return chr_first(c) || c == '*' || c == '='; }
inline bool name_char(unsigned char c) { return (c >= '!' && c <= '?') || (c >= 'A' && c <= '~'); }

// [0-9]+ ; returns false if no digit or overflow
inline bool parse_uint(const char*& s, const char* e, int64_t& v) {
  if (s >= e || !is_digit(*s)) return false;
  int64_t x = 0;
  while (s < e && is_digit(*s)) {
    if (x > (INT64_MAX - 9) / 10) return false;
    x = x * 10 + (*s - '0');
    ++s;
  }
  v = x;
  return true;
}
inline bool parse_i32(const char*& s, const char* e, int32_t& v) {
  int64_t x;
  if (!parse_uint(s, e, x) || x > INT32_MAX) return false;
  v = (int32_t)x;
  return true;
}
inline bool expect(const char*& s, const char* e, char c) {
  if (s < e && *s == c) { ++s; return true; }
  return false;
}
inline bool parse_chr(const char*& s, const char* e) {
  if (s >= e || !chr_first((unsigned char)*s)) return false;
  ++s;
  while (s < e && chr_rest((unsigned char)*s)) ++s;
  return true;
}

struct KeyHash {
  size_t operator()(const std::vector<int32_t>& k) const {
    uint64_t h = 1469598103934665603ull;
    for (int32_t v : k) { h ^= (uint32_t)v; h *= 1099511628211ull; }
    return (size_t)h;
  }
};

std::string line_snip(const char* s, const char* e) {
  size_t n = (size_t)(e - s);
  if (n > 60) n = 60;
  return std::string(s, n);
}

bool parse_split_file(const char* path, TintData& T) {
  FileBuf fb;
  if (!fb.load(path, T.error, 0)) return false;
  ProfScope ps(PROF_SPLIT);
  const char* s = fb.p;
  const char* end = fb.p + fb.n;
  bool have_header = false;
  // read-rep dedupe (:165-170): open-addressing table of rep ids; a rep's key is compared against the
  // target intervals of its first read (no per-rep key allocation)
  std::vector<int32_t> rep_slot;       // -1 = empty, else rep id
  std::vector<int32_t> rep_first_iv;   // rep -> first interval index (into riv_ts / riv_te) of its first read
  size_t rep_mask = 0;
  auto rep_table_init = [&](size_t n_reads) {
    size_t cap = 64;
    while (cap < 2 * n_reads + 2) cap <<= 1;
    rep_slot.assign(cap, -1);
    rep_mask = cap - 1;
  };
  auto rep_table_grow = [&]() {
    std::vector<int32_t> old;
    old.swap(rep_slot);
    rep_slot.assign(old.size() * 2, -1);
    rep_mask = rep_slot.size() - 1;
    for (int32_t r : old) {
      if (r < 0) continue;
      uint64_t h = 1469598103934665603ull;
      for (int32_t k = T.rep_iv_off[(size_t)r], f = rep_first_iv[(size_t)r]; k < T.rep_iv_off[(size_t)r + 1]; ++k, ++f) {
        h ^= (uint32_t)T.riv_ts[(size_t)f]; h *= 1099511628211ull;
        h ^= (uint32_t)T.riv_te[(size_t)f]; h *= 1099511628211ull;
      }
      size_t p = (size_t)h & rep_mask;
      while (rep_slot[p] >= 0) p = (p + 1) & rep_mask;
      rep_slot[p] = r;
    }
  };
  std::vector<int32_t> key;
  while (s < end) {
    const char* nl = (const char*)memchr(s, '\n', (size_t)(end - s));
    if (!nl) { T.error = std::string("AttributeError: line without newline does not match (") + path + "): " + line_snip(s, end); return false; }
    const char* e = nl;  // line content is [s, e)
    if (*s == '#') {
      // #<chr>\t<id>\t<s>-<e>(,<s>-<e>)*\t<count>
      const char* q = s + 1;
      const char* c0 = q;
      int64_t id, cnt;
      if (!parse_chr(q, e)) goto bad_header;
      {
        std::string chr(c0, (size_t)(q - c0));
        if (!expect(q, e, '\t') || !parse_uint(q, e, id) || !expect(q, e, '\t')) goto bad_header;
        std::vector<int32_t> is, ie;
        for (;;) {
          int32_t a, b;
          if (!parse_i32(q, e, a) || !expect(q, e, '-') || !parse_i32(q, e, b)) goto bad_header;
          is.push_back(a);
          ie.push_back(b);
          if (q < e && *q == ',') { ++q; continue; }
          break;
        }
        if (!expect(q, e, '\t') || !parse_uint(q, e, cnt) || q != e) goto bad_header;
        if (have_header) {
          if (id == T.id) T.error = "AssertionError: Transcriptional interval with id " + std::to_string(id) + " is repeated!";
          else T.error = std::string("AssertionError: assert len(tints) == 1 (freddie_segment.py:699): ") + path;
          return false;
        }
        for (size_t i = 0; i + 1 < is.size(); ++i)
          if (!(ie[i] < is[i + 1])) { T.error = "AssertionError: tint intervals overlap or are unordered (freddie_segment.py:138)"; return false; }
        for (size_t i = 0; i < is.size(); ++i)
          if (!(is[i] < ie[i])) { T.error = "AssertionError: empty tint interval (freddie_segment.py:140)"; return false; }
        T.chr = chr;
        T.id = id;
        T.read_count = cnt;
        T.isl_s = is;
        T.isl_e = ie;
        rep_table_init((size_t)std::max<int64_t>(cnt, 0) < (size_t)(fb.n / 16 + 16) ? (size_t)std::max<int64_t>(cnt, 0) : fb.n / 16 + 16);
        T.isl_off.assign(1, 0);
        int64_t off = 0;
        for (size_t i = 0; i < is.size(); ++i) {
          off += (int64_t)ie[i] - is[i] + 1;
          if (off > INT32_MAX) { T.error = "tint spans more than 2^31 samples"; return false; }
          T.isl_off.push_back((int32_t)off);
        }
        have_header = true;
      }
      s = nl + 1;
      continue;
    bad_header:
      T.error = std::string("AttributeError: tint header does not match tint_prog (freddie_segment.py:22): ") + line_snip(s, e);
      return false;
    } else {
      // <rid>\t<name>\t<chr>\t<strand>\t<tint>\t<iv>(\t<iv>)*
      const char* q = s;
      ReadMeta m;
      if (!parse_uint(q, e, m.rid) || !expect(q, e, '\t')) goto bad_read;
      {
        const char* n0 = q;
        while (q < e && name_char((unsigned char)*q)) ++q;
        size_t nlen = (size_t)(q - n0);
        if (nlen < 1 || nlen > 254 || !expect(q, e, '\t')) goto bad_read;
        const char* c0 = q;
        if (!parse_chr(q, e)) goto bad_read;
        size_t clen = (size_t)(q - c0);
        if (!expect(q, e, '\t')) goto bad_read;
        if (q >= e || (*q != '+' && *q != '-')) goto bad_read;
        m.strand = *q++;
        if (!expect(q, e, '\t') || !parse_uint(q, e, m.tint) || !expect(q, e, '\t')) goto bad_read;
        if (!have_header || m.tint != T.id) { T.error = "KeyError: read row refers to tint " + std::to_string(m.tint) + " (freddie_segment.py:162)"; return false; }
        m.name_off = (uint32_t)T.text.size();
        m.name_len = (uint32_t)nlen;
        T.text.append(n0, nlen);
        m.chr_off = (uint32_t)T.text.size();
        m.chr_len = (uint32_t)clen;
        T.text.append(c0, clen);
        key.clear();
        size_t iv0 = T.riv_ts.size();
        for (;;) {
          int32_t ts, te, qs, qe;
          if (!parse_i32(q, e, ts) || !expect(q, e, '-') || !parse_i32(q, e, te) || !expect(q, e, ':') ||
              !parse_i32(q, e, qs) || !expect(q, e, '-') || !parse_i32(q, e, qe) || !expect(q, e, ':'))
            goto bad_read;
          int nops = 0;
          while (q < e && is_digit(*q)) {
            int64_t c;
            if (!parse_uint(q, e, c) || q >= e) goto bad_read;
            uint32_t op;
            switch (*q) {
              case 'M': case 'X': case '=': op = 0; break;
              case 'I': op = 1; break;
              case 'D': op = 2; break;
              case 'N': case 'S': case 'H': case 'P': op = 3; break;
              default: goto bad_read;
            }
            ++q;
            if (c >= (1ll << 28)) { T.error = "CIGAR operation longer than 2^28"; return false; }
            T.cigar.push_back(((uint32_t)c << 4) | op);
            ++nops;
          }
          if (nops == 0) goto bad_read;
          T.riv_ts.push_back(ts);
          T.riv_te.push_back(te);
          T.riv_qs.push_back(qs);
          T.riv_qe.push_back(qe);
          T.riv_cig_off.push_back((int32_t)T.cigar.size());
          key.push_back(ts);
          key.push_back(te);
          if (q < e && *q == '\t') { ++q; continue; }
          break;
        }
        if (q != e) goto bad_read;
        size_t iv1 = T.riv_ts.size();
        for (size_t k = iv0; k + 1 < iv1; ++k)
          if (!(T.riv_te[k] <= T.riv_ts[k + 1] && T.riv_qe[k] <= T.riv_qs[k + 1])) {
            T.error = "AssertionError: read intervals out of order (freddie_segment.py:158)";
            return false;
          }
        for (size_t k = iv0; k < iv1; ++k)
          if (!(T.riv_ts[k] < T.riv_te[k] && T.riv_qs[k] < T.riv_qe[k])) {
            T.error = "AssertionError: empty read interval (freddie_segment.py:160)";
            return false;
          }
        T.read_iv_off.push_back((int32_t)iv1);
        T.read_strand.push_back(m.strand == '+' ? 0 : 1);
        T.meta.push_back(m);
        // read rep (first-seen order, :165-170)
        uint64_t kh = 1469598103934665603ull;
        for (int32_t v : key) { kh ^= (uint32_t)v; kh *= 1099511628211ull; }
        if ((T.rep_w.size() + 1) * 2 > rep_slot.size()) rep_table_grow();
        size_t pos = (size_t)kh & rep_mask;
        int32_t rep = -1;
        const size_t n_key_iv = key.size() / 2;
        while (rep_slot[pos] >= 0) {
          const int32_t r = rep_slot[pos];
          if ((size_t)(T.rep_iv_off[(size_t)r + 1] - T.rep_iv_off[(size_t)r]) == n_key_iv) {
            const int32_t f = rep_first_iv[(size_t)r];
            bool same = true;
            for (size_t k = 0; k < n_key_iv && same; ++k)
              same = T.riv_ts[(size_t)f + k] == key[2 * k] && T.riv_te[(size_t)f + k] == key[2 * k + 1];
            if (same) { rep = r; break; }
          }
          pos = (pos + 1) & rep_mask;
        }
        if (rep < 0) {
          rep = (int32_t)T.rep_w.size();
          rep_slot[pos] = rep;
          rep_first_iv.push_back((int32_t)iv0);
          T.rep_w.push_back(0);
          for (size_t k = 0; k < key.size(); k += 2) {
            int32_t ts = key[k], te = key[k + 1];
            // island of ts: last island with start <= ts
            size_t a = (size_t)(std::upper_bound(T.isl_s.begin(), T.isl_s.end(), ts) - T.isl_s.begin());
            if (a == 0 || ts > T.isl_e[a - 1]) { T.error = "KeyError: " + std::to_string(ts) + " (freddie_segment.py:666)"; return false; }
            --a;
            if (te > T.isl_e[a]) {
              size_t b = (size_t)(std::upper_bound(T.isl_s.begin(), T.isl_s.end(), te) - T.isl_s.begin());
              if (b == 0 || te > T.isl_e[b - 1]) T.error = "KeyError: " + std::to_string(te) + " (freddie_segment.py:667)";
              else T.error = "AssertionError: assert Y_idx_s == Y_idx_e (freddie_segment.py:668)";
              return false;
            }
            T.rep_fs.push_back(T.isl_off[a] + (ts - T.isl_s[a]));
            T.rep_fe.push_back(T.isl_off[a] + (te - T.isl_s[a]));
          }
          T.rep_iv_off.push_back((int32_t)T.rep_fs.size());
        }
        T.rep_w[rep] += 1;
        T.read_rep.push_back(rep);
      }
      s = nl + 1;
      continue;
    bad_read:
      T.error = std::string("AttributeError: read row does not match read_prog (freddie_segment.py:28): ") + line_snip(s, e);
      return false;
    }
  }
  if (!have_header) { T.error = std::string("AssertionError: assert len(tints) == 1 (freddie_segment.py:699): ") + path; return false; }
  if ((int64_t)T.meta.size() != T.read_count) { T.error = "AssertionError: assert len(tint['reads']) == tint['read_count'] (freddie_segment.py:164)"; return false; }
  return true;
}

// read_sequence (:174-185): cols 0 and 3 of every line; later duplicates of a rid win
bool parse_reads_file(const char* path, TintData& T) {
  FileBuf fb;
  if (!fb.load(path, T.error, 1)) return false;
  std::unordered_map<int64_t, std::pair<const char*, uint32_t>> seqs;
  seqs.reserve(T.meta.size() * 2);
  const char* s = fb.p;
  const char* end = fb.p + fb.n;
  {
  ProfScope ps_lines(PROF_READS_LINES);
  while (s < end) {
    const char* nl = (const char*)memchr(s, '\n', (size_t)(end - s));
    const char* e = nl ? nl : end;
    const char* next = nl ? nl + 1 : end;
    // rstrip(): trailing whitespace
    while (e > s && (e[-1] == ' ' || e[-1] == '\t' || e[-1] == '\r' || e[-1] == '\n' || e[-1] == '\v' || e[-1] == '\f')) --e;
    // int(line[0]) tolerates surrounding blanks and a sign; split rows only hold digits
    const char* q = s;
    int64_t rid;
    if (!parse_uint(q, e, rid) || (q < e && *q != '\t')) { T.error = std::string("ValueError: invalid read id in ") + path; return false; }
    int tabs = 0;
    const char* f3 = nullptr;
    for (const char* c = s; c < e; ++c)
      if (*c == '\t') { if (++tabs == 3) { f3 = c + 1; break; } }
    if (!f3) { T.error = std::string("IndexError: list index out of range (freddie_segment.py:179): ") + path; return false; }
    const char* f3e = (const char*)memchr(f3, '\t', (size_t)(e - f3));
    if (!f3e) f3e = e;
    seqs[rid] = std::make_pair(f3, (uint32_t)(f3e - f3));
    s = next;
  }
  }
  if (seqs.size() != T.meta.size()) { T.error = "AssertionError: assert len(rid_to_seq) == len(tint['reads']) (freddie_segment.py:181)"; return false; }
  // bit-planes
  ProfScope ps_planes(PROF_PLANES);
  size_t words = 0;
  T.read_len.reserve(T.meta.size());
  for (const ReadMeta& m : T.meta) {
    auto it = seqs.find(m.rid);
    if (it == seqs.end()) { T.error = "KeyError: " + std::to_string(m.rid) + " (freddie_segment.py:183)"; return false; }
    uint32_t L = it->second.second;
    if (L > (uint32_t)INT32_MAX) { T.error = "read longer than 2^31"; return false; }
    T.read_len.push_back((int32_t)L);
    words += (L + 31) / 32;
    T.read_seq_off.push_back((int64_t)words);
  }
  T.seq_a.resize(words);
  T.seq_t.resize(words);
  size_t w0 = 0;
  for (const ReadMeta& m : T.meta) {
    const auto& sq = seqs[m.rid];
    seq_planes(sq.first, sq.second, T.seq_a.data() + w0, T.seq_t.data() + w0);
    w0 += (sq.second + 31) / 32;
  }
  return true;
}

template <typename F>
void parallel_for(int n, int n_threads, F f) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > n) n_threads = n;
  if (n_threads <= 1) {
    for (int i = 0; i < n; ++i) f(i);
    return;
  }
  std::atomic<int> next(0);
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; ++t)
    th.emplace_back([&]() {
      for (;;) {
        int i = next.fetch_add(1);
        if (i >= n) break;
        f(i);
      }
    });
  for (auto& x : th) x.join();
}

template <typename T>
void append_shift(std::vector<T>& dst, const std::vector<T>& src, T shift, size_t skip = 0) {
  size_t o = dst.size();
  dst.resize(o + src.size() - skip);
  for (size_t i = skip; i < src.size(); ++i) dst[o + i - skip] = src[i] + shift;
}

}  // namespace

// uninitialised storage for the large concatenated arrays (a std::vector would zero-fill ~400 MB
// single-threaded before the parallel copy overwrites every element)
template <typename T>
struct RawBuf {
  T* p = nullptr;
  size_t n = 0;
  ~RawBuf() { free(p); }
  void resize(size_t m) { free(p); p = (T*)malloc((m ? m : 1) * sizeof(T)); n = m; }
  T* data() { return p; }
  const T* data() const { return p; }
  size_t size() const { return n; }
};

struct frs_parsed {
  std::vector<TintData> tints;
  // concatenated batch
  std::vector<int32_t> tint_island_off, tint_rep_off, tint_read_off, island_start, island_sample_off, rep_iv_off,
      rep_weight, rep_iv_fs, rep_iv_fe, read_rep, read_len, read_iv_off, riv_ts, riv_te, riv_qs, riv_qe, riv_cig_off;
  std::vector<uint8_t> read_strand;
  std::vector<int64_t> read_seq_off;
  RawBuf<uint32_t> cigar, seq_a, seq_t;
};

extern "C" {

int frs_parse_tints(const char* const* split_paths, const char* const* reads_paths, int n, int n_threads,
                    frs_parsed** out, char* err, size_t err_cap) {
  if (err && err_cap) err[0] = 0;
  if (!out || n <= 0) {
    if (err) snprintf(err, err_cap, "frs_parse_tints: bad arguments");
    return FRS_ERR_ARG;
  }
  frs_parsed* P = new frs_parsed();
  P->tints.resize((size_t)n);
  std::atomic<int> failed(-1);
  parallel_for(n, n_threads, [&](int i) {
    TintData& T = P->tints[(size_t)i];
    if (!parse_split_file(split_paths[i], T) || !parse_reads_file(reads_paths[i], T)) {
      int exp = -1;
      failed.compare_exchange_strong(exp, i);
    }
  });
  // deterministic error: the first failing tint in input order
  for (int i = 0; i < n; ++i)
    if (!P->tints[(size_t)i].error.empty()) {
      if (err) snprintf(err, err_cap, "%s [%s]", P->tints[(size_t)i].error.c_str(), split_paths[i]);
      bool io = P->tints[(size_t)i].error.rfind("FileNotFoundError", 0) == 0;
      delete P;
      return io ? FRS_ERR_IO : FRS_ERR_ARG;
    }
  // concatenate
  ProfScope* ps_cat = new ProfScope(PROF_CONCAT);
  int64_t smp = 0, reps = 0, rivs = 0, rep_ivs = 0, cig = 0, reads = 0, isl = 0, words = 0;
  for (TintData& T : P->tints) {
    smp += T.isl_off.back();
    reps += (int64_t)T.rep_w.size();
    rep_ivs += (int64_t)T.rep_fs.size();
    rivs += (int64_t)T.riv_ts.size();
    cig += (int64_t)T.cigar.size();
    reads += (int64_t)T.meta.size();
    isl += (int64_t)T.isl_s.size();
    words += (int64_t)T.seq_a.size();
  }
  if (smp > INT32_MAX || rep_ivs > INT32_MAX || rivs > INT32_MAX || cig > INT32_MAX) {
    if (err) snprintf(err, err_cap, "batch too large for 32-bit offsets: split it");
    delete P;
    return FRS_ERR_LIMIT;
  }
  // every tint's slice of every array is known from the prefix sums: size once, copy in parallel
  const size_t NT = P->tints.size();
  struct Base { int64_t smp, reps, rep_ivs, rivs, cig, reads, isl, words; };
  std::vector<Base> base(NT + 1);
  {
    Base a{0, 0, 0, 0, 0, 0, 0, 0};
    for (size_t t = 0; t < NT; ++t) {
      const TintData& T = P->tints[t];
      base[t] = a;
      a.smp += T.isl_off.back();
      a.reps += (int64_t)T.rep_w.size();
      a.rep_ivs += (int64_t)T.rep_fs.size();
      a.rivs += (int64_t)T.riv_ts.size();
      a.cig += (int64_t)T.cigar.size();
      a.reads += (int64_t)T.meta.size();
      a.isl += (int64_t)T.isl_s.size();
      a.words += (int64_t)T.seq_a.size();
    }
    base[NT] = a;
  }
  P->tint_island_off.resize(NT + 1);
  P->tint_rep_off.resize(NT + 1);
  P->tint_read_off.resize(NT + 1);
  P->island_start.resize((size_t)isl);
  P->island_sample_off.resize((size_t)isl + 1);
  P->rep_iv_off.resize((size_t)reps + 1);
  P->rep_weight.resize((size_t)reps);
  P->rep_iv_fs.resize((size_t)rep_ivs);
  P->rep_iv_fe.resize((size_t)rep_ivs);
  P->read_rep.resize((size_t)reads);
  P->read_strand.resize((size_t)reads);
  P->read_len.resize((size_t)reads);
  P->read_iv_off.resize((size_t)reads + 1);
  P->read_seq_off.resize((size_t)reads + 1);
  P->riv_ts.resize((size_t)rivs);
  P->riv_te.resize((size_t)rivs);
  P->riv_qs.resize((size_t)rivs);
  P->riv_qe.resize((size_t)rivs);
  P->riv_cig_off.resize((size_t)rivs + 1);
  P->cigar.resize((size_t)cig);
  P->seq_a.resize((size_t)words);
  P->seq_t.resize((size_t)words);
  P->tint_island_off[NT] = (int32_t)isl;
  P->tint_rep_off[NT] = (int32_t)reps;
  P->tint_read_off[NT] = (int32_t)reads;
  P->island_sample_off[0] = 0;
  P->rep_iv_off[0] = 0;
  P->read_iv_off[0] = 0;
  P->riv_cig_off[0] = 0;
  P->read_seq_off[0] = 0;
  parallel_for((int)NT, n_threads, [&](int ti) {
    TintData& T = P->tints[(size_t)ti];
    const Base& b = base[(size_t)ti];
    auto put = [](auto& dst, int64_t at, const auto& src, size_t skip, auto shift) {
      for (size_t i = skip; i < src.size(); ++i) dst[(size_t)at + i - skip] = (decltype(shift))(src[i] + shift);
    };
    P->tint_island_off[(size_t)ti] = (int32_t)b.isl;
    P->tint_rep_off[(size_t)ti] = (int32_t)b.reps;
    P->tint_read_off[(size_t)ti] = (int32_t)b.reads;
    put(P->island_start, b.isl, T.isl_s, 0, (int32_t)0);
    put(P->island_sample_off, b.isl + 1, T.isl_off, 1, (int32_t)b.smp);
    put(P->rep_iv_off, b.reps + 1, T.rep_iv_off, 1, (int32_t)b.rep_ivs);
    put(P->rep_weight, b.reps, T.rep_w, 0, (int32_t)0);
    put(P->rep_iv_fs, b.rep_ivs, T.rep_fs, 0, (int32_t)b.smp);
    put(P->rep_iv_fe, b.rep_ivs, T.rep_fe, 0, (int32_t)b.smp);
    put(P->read_rep, b.reads, T.read_rep, 0, (int32_t)b.reps);
    put(P->read_strand, b.reads, T.read_strand, 0, (uint8_t)0);
    put(P->read_len, b.reads, T.read_len, 0, (int32_t)0);
    put(P->read_iv_off, b.reads + 1, T.read_iv_off, 1, (int32_t)b.rivs);
    put(P->read_seq_off, b.reads + 1, T.read_seq_off, 1, (int64_t)b.words);
    put(P->riv_ts, b.rivs, T.riv_ts, 0, (int32_t)0);
    put(P->riv_te, b.rivs, T.riv_te, 0, (int32_t)0);
    put(P->riv_qs, b.rivs, T.riv_qs, 0, (int32_t)0);
    put(P->riv_qe, b.rivs, T.riv_qe, 0, (int32_t)0);
    put(P->riv_cig_off, b.rivs + 1, T.riv_cig_off, 1, (int32_t)b.cig);
    if (!T.cigar.empty()) memcpy(P->cigar.data() + b.cig, T.cigar.data(), T.cigar.size() * 4);
    if (!T.seq_a.empty()) {
      memcpy(P->seq_a.data() + b.words, T.seq_a.data(), T.seq_a.size() * 4);
      memcpy(P->seq_t.data() + b.words, T.seq_t.data(), T.seq_t.size() * 4);
    }
    // per-tint copies of the big arrays are no longer needed
    std::vector<uint32_t>().swap(T.seq_a);
    std::vector<uint32_t>().swap(T.seq_t);
    std::vector<uint32_t>().swap(T.cigar);
  });
  delete ps_cat;
  if (g_prof.on) {
    fprintf(stderr, "[frs host profile] thread-seconds: load %.3f  split rows %.3f  reads lines %.3f  bit-planes %.3f  concat %.3f\n",
            g_prof.ns[0] * 1e-9, (g_prof.ns[1]) * 1e-9, g_prof.ns[2] * 1e-9, g_prof.ns[3] * 1e-9, g_prof.ns[4] * 1e-9);
    for (auto& x : g_prof.ns) x = 0;
  }
  *out = P;
  return 0;
}

int frs_parsed_batch(const frs_parsed* P, frs_batch* b) {
  if (!P || !b) return FRS_ERR_ARG;
  memset(b, 0, sizeof *b);
  b->n_tints = (int32_t)P->tints.size();
  b->n_islands = (int32_t)P->island_start.size();
  b->n_reps = (int32_t)P->rep_weight.size();
  b->n_rep_ivs = (int32_t)P->rep_iv_fs.size();
  b->n_reads = (int32_t)P->read_rep.size();
  b->n_read_ivs = (int32_t)P->riv_ts.size();
  b->n_cigar_ops = (int32_t)P->cigar.size();
  b->n_samples = P->island_sample_off.back();
  b->n_seq_words = (int64_t)P->seq_a.size();
  b->tint_island_off = P->tint_island_off.data();
  b->tint_rep_off = P->tint_rep_off.data();
  b->tint_read_off = P->tint_read_off.data();
  b->island_start = P->island_start.data();
  b->island_sample_off = P->island_sample_off.data();
  b->rep_iv_off = P->rep_iv_off.data();
  b->rep_weight = P->rep_weight.data();
  b->rep_iv_fs = P->rep_iv_fs.data();
  b->rep_iv_fe = P->rep_iv_fe.data();
  b->read_rep = P->read_rep.data();
  b->read_strand = P->read_strand.data();
  b->read_len = P->read_len.data();
  b->read_iv_off = P->read_iv_off.data();
  b->read_seq_off = P->read_seq_off.data();
  b->riv_ts = P->riv_ts.data();
  b->riv_te = P->riv_te.data();
  b->riv_qs = P->riv_qs.data();
  b->riv_qe = P->riv_qe.data();
  b->riv_cig_off = P->riv_cig_off.data();
  b->cigar = P->cigar.data();
  b->seq_is_a = P->seq_a.data();
  b->seq_is_t = P->seq_t.data();
  return 0;
}

void frs_parsed_free(frs_parsed* P) { delete P; }

// run_segment output (:715-731): "#chr\tid\tpos,pos,...\n" then one row per read in file order;
// gap strings sorted as python strings (:472), each followed by a comma.
int frs_format_tints(const frs_parsed* P, const frs_result* R, const char* const* out_paths,
                     const char* const* log_paths, int n_threads, char* err, size_t err_cap) {
  if (err && err_cap) err[0] = 0;
  if (!P || !R || !out_paths) return FRS_ERR_ARG;
  const int n = (int)P->tints.size();
  std::vector<std::string> errors((size_t)n);
  parallel_for(n, n_threads, [&](int t) {
    const TintData& T = P->tints[(size_t)t];
    const int32_t f0 = R->tint_final_off[t], f1 = R->tint_final_off[t + 1];
    const int64_t S = f1 - f0 - 1;
    const int64_t d0 = R->tint_digit_off[t];
    const int32_t rep0 = P->tint_rep_off[(size_t)t];
    const int32_t r0 = P->tint_read_off[(size_t)t];
    std::string o;
    o.reserve((size_t)(T.meta.size() * (size_t)(S + 96) + (size_t)(f1 - f0) * 11 + 64));
    char num[32];
    o.push_back('#');
    o += T.chr;
    o.push_back('\t');
    o.append(num, (size_t)snprintf(num, sizeof num, "%lld", (long long)T.id));
    o.push_back('\t');
    for (int32_t f = f0; f < f1; ++f) {
      if (f > f0) o.push_back(',');
      o.append(num, (size_t)snprintf(num, sizeof num, "%d", R->final_pos[f]));
    }
    o.push_back('\n');
    std::vector<std::string> gaps;
    for (size_t k = 0; k < T.meta.size(); ++k) {
      const ReadMeta& m = T.meta[k];
      const int64_t i = (int64_t)r0 + (int64_t)k;
      o.append(num, (size_t)snprintf(num, sizeof num, "%lld", (long long)m.rid));
      o.push_back('\t');
      o.append(T.text, m.name_off, m.name_len);
      o.push_back('\t');
      o.append(T.text, m.chr_off, m.chr_len);
      o.push_back('\t');
      o.push_back(m.strand);
      o.push_back('\t');
      o.append(num, (size_t)snprintf(num, sizeof num, "%lld", (long long)m.tint));
      o.push_back('\t');
      const int32_t rep = P->read_rep[(size_t)i] - rep0;
      o.append((const char*)R->digits + d0 + (int64_t)rep * S, (size_t)S);
      o.push_back('\t');
      const int32_t* h = R->read_head + i * 8;
      if (h[FRS_HEAD_FLAGS] & 1) {
        gaps.clear();
        char g[64];
        int sk = (h[FRS_HEAD_FLAGS] >> 8) & 3, ek = (h[FRS_HEAD_FLAGS] >> 16) & 3;
        if (sk) { snprintf(g, sizeof g, "S%c_%d:%d", sk == 1 ? 'A' : 'T', h[FRS_HEAD_S_LEN], h[FRS_HEAD_S_GAP]); gaps.emplace_back(g); }
        snprintf(g, sizeof g, "SSC:%d", h[FRS_HEAD_SSC]);
        gaps.emplace_back(g);
        if (ek) { snprintf(g, sizeof g, "E%c_%d:%d", ek == 1 ? 'A' : 'T', h[FRS_HEAD_E_LEN], h[FRS_HEAD_E_GAP]); gaps.emplace_back(g); }
        snprintf(g, sizeof g, "ESC:%d", h[FRS_HEAD_ESC]);
        gaps.emplace_back(g);
        for (int32_t q = R->read_gap_off[i]; q < R->read_gap_off[i + 1]; ++q) {
          const int32_t* r = R->gap_rec + (int64_t)q * 3;
          snprintf(g, sizeof g, "%d-%d:%d", r[0], r[1], r[2]);
          gaps.emplace_back(g);
        }
        std::sort(gaps.begin(), gaps.end());
        gaps.erase(std::unique(gaps.begin(), gaps.end()), gaps.end());  // read['gaps'] is a set (:371)
        for (const std::string& s : gaps) { o += s; o.push_back(','); }
      }
      o.push_back('\n');
    }
    FILE* f = fopen(out_paths[t], "wb");
    if (!f) { errors[(size_t)t] = std::string("cannot open ") + out_paths[t] + ": " + strerror(errno); return; }
    if (fwrite(o.data(), 1, o.size(), f) != o.size()) errors[(size_t)t] = std::string("short write to ") + out_paths[t];
    fclose(f);
    if (log_paths && log_paths[t]) {
      FILE* l = fopen(log_paths[t], "wb");  // the reference leaves an empty .log per tint (:695,:734)
      if (l) fclose(l);
      else errors[(size_t)t] = std::string("cannot open ") + log_paths[t];
    }
  });
  for (int t = 0; t < n; ++t)
    if (!errors[(size_t)t].empty()) {
      if (err) snprintf(err, err_cap, "%s", errors[(size_t)t].c_str());
      return FRS_ERR_IO;
    }
  return 0;
}

}  // extern "C"
