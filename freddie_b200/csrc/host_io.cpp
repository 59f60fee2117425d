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
#include <memory>
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
  bool load(const char* path, std::string& err, int which, bool always_map = false) {
    ProfScope ps(PROF_LOAD);
    int fd = open(path, O_RDONLY);
    if (fd < 0) { err = std::string("FileNotFoundError: ") + path + ": " + strerror(errno); return false; }
    struct stat st;
    if (fstat(fd, &st) != 0) { err = std::string("OSError: ") + path + ": " + strerror(errno); close(fd); return false; }
    n = (size_t)st.st_size;
    if (n > SMALL || (always_map && n > 0)) {
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

std::string line_snip(const char* s, const char* e) {
  size_t n = (size_t)(e - s);
  if (n > 60) n = 60;
  return std::string(s, n);
}

// ---- tint header:  #<chr>\t<id>\t<s>-<e>(,<s>-<e>)*\t<count>   (tint_prog, :22) ----
bool parse_header_line(const char* s, const char* e, const char* path, bool have_header, TintData& T) {
  const char* q = s + 1;
  const char* c0 = q;
  int64_t id, cnt;
  std::vector<int32_t> is, ie;
  if (!parse_chr(q, e)) goto bad_header;
  {
    std::string chr(c0, (size_t)(q - c0));
    if (!expect(q, e, '\t') || !parse_uint(q, e, id) || !expect(q, e, '\t')) goto bad_header;
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
    T.isl_off.assign(1, 0);
    int64_t off = 0;
    for (size_t i = 0; i < is.size(); ++i) {
      off += (int64_t)ie[i] - is[i] + 1;
      if (off > INT32_MAX) { T.error = "tint spans more than 2^31 samples"; return false; }
      T.isl_off.push_back((int32_t)off);
    }
  }
  return true;
bad_header:
  T.error = std::string("AttributeError: tint header does not match tint_prog (freddie_segment.py:22): ") + line_snip(s, e);
  return false;
}

// ---- read row:  <rid>\t<name>\t<chr>\t<strand>\t<tint>\t<iv>(\t<iv>)*   (read_prog, :28) ----
// Appends the row to C (meta, text, intervals, CIGARs); the read-rep dedupe is a separate step
// (RepTable::add_read) so that the rows of a giant tint can be parsed by several threads.
bool parse_read_row(const char* s, const char* e, bool have_header, int64_t tint_id, TintData& C) {
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
    if (!have_header || m.tint != tint_id) { C.error = "KeyError: read row refers to tint " + std::to_string(m.tint) + " (freddie_segment.py:162)"; return false; }
    m.name_off = (uint32_t)C.text.size();
    m.name_len = (uint32_t)nlen;
    C.text.append(n0, nlen);
    m.chr_off = (uint32_t)C.text.size();
    m.chr_len = (uint32_t)clen;
    C.text.append(c0, clen);
    size_t iv0 = C.riv_ts.size();
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
        if (c >= (1ll << 28)) { C.error = "CIGAR operation longer than 2^28"; return false; }
        C.cigar.push_back(((uint32_t)c << 4) | op);
        ++nops;
      }
      if (nops == 0) goto bad_read;
      C.riv_ts.push_back(ts);
      C.riv_te.push_back(te);
      C.riv_qs.push_back(qs);
      C.riv_qe.push_back(qe);
      C.riv_cig_off.push_back((int32_t)C.cigar.size());
      if (q < e && *q == '\t') { ++q; continue; }
      break;
    }
    if (q != e) goto bad_read;
    size_t iv1 = C.riv_ts.size();
    for (size_t k = iv0; k + 1 < iv1; ++k)
      if (!(C.riv_te[k] <= C.riv_ts[k + 1] && C.riv_qe[k] <= C.riv_qs[k + 1])) {
        C.error = "AssertionError: read intervals out of order (freddie_segment.py:158)";
        return false;
      }
    for (size_t k = iv0; k < iv1; ++k)
      if (!(C.riv_ts[k] < C.riv_te[k] && C.riv_qs[k] < C.riv_qe[k])) {
        C.error = "AssertionError: empty read interval (freddie_segment.py:160)";
        return false;
      }
    C.read_iv_off.push_back((int32_t)iv1);
    C.read_strand.push_back(m.strand == '+' ? 0 : 1);
    C.meta.push_back(m);
  }
  return true;
bad_read:
  C.error = std::string("AttributeError: read row does not match read_prog (freddie_segment.py:28): ") + line_snip(s, e);
  return false;
}

// read-rep dedupe (:165-170): reps are the distinct tuples of target intervals in first-seen order.
// Open-addressing table of rep ids; a rep's key is compared against the target intervals of its first
// read (no per-rep key allocation).
struct RepTable {
  std::vector<int32_t> slot;       // -1 = empty, else rep id
  std::vector<int32_t> first_iv;   // rep -> first interval index (into riv_ts / riv_te) of its first read
  size_t mask = 0;
  static uint64_t hash_ivs(const TintData& T, size_t f, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t k = 0; k < n; ++k) {
      h ^= (uint32_t)T.riv_ts[f + k]; h *= 1099511628211ull;
      h ^= (uint32_t)T.riv_te[f + k]; h *= 1099511628211ull;
    }
    return h;
  }
  void init(size_t n_reads) {
    size_t cap = 64;
    while (cap < 2 * n_reads + 2) cap <<= 1;
    slot.assign(cap, -1);
    mask = cap - 1;
  }
  void grow(const TintData& T) {
    std::vector<int32_t> old;
    old.swap(slot);
    slot.assign(old.size() * 2, -1);
    mask = slot.size() - 1;
    for (int32_t r : old) {
      if (r < 0) continue;
      size_t p = (size_t)hash_ivs(T, (size_t)first_iv[(size_t)r], (size_t)(T.rep_iv_off[(size_t)r + 1] - T.rep_iv_off[(size_t)r])) & mask;
      while (slot[p] >= 0) p = (p + 1) & mask;
      slot[p] = r;
    }
  }
  // read `k` of T (its intervals are in place): find or create its rep
  bool add_read(TintData& T, size_t k) {
    if (slot.empty()) init(64);
    const size_t iv0 = (size_t)T.read_iv_off[k], n_iv = (size_t)T.read_iv_off[k + 1] - iv0;
    if ((T.rep_w.size() + 1) * 2 > slot.size()) grow(T);
    size_t pos = (size_t)hash_ivs(T, iv0, n_iv) & mask;
    int32_t rep = -1;
    while (slot[pos] >= 0) {
      const int32_t r = slot[pos];
      if ((size_t)(T.rep_iv_off[(size_t)r + 1] - T.rep_iv_off[(size_t)r]) == n_iv) {
        const size_t f = (size_t)first_iv[(size_t)r];
        bool same = true;
        for (size_t i = 0; i < n_iv && same; ++i)
          same = T.riv_ts[f + i] == T.riv_ts[iv0 + i] && T.riv_te[f + i] == T.riv_te[iv0 + i];
        if (same) { rep = r; break; }
      }
      pos = (pos + 1) & mask;
    }
    if (rep < 0) {
      rep = (int32_t)T.rep_w.size();
      slot[pos] = rep;
      first_iv.push_back((int32_t)iv0);
      T.rep_w.push_back(0);
      for (size_t i = 0; i < n_iv; ++i) {
        const int32_t ts = T.riv_ts[iv0 + i], te = T.riv_te[iv0 + i];
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
    T.rep_w[(size_t)rep] += 1;
    T.read_rep.push_back(rep);
    return true;
  }
};

template <typename F>
void parallel_for(int n, int n_threads, F f);

size_t format_big_rows() {  // tints with more reads are formatted by all threads together
  const char* e = getenv("FRS_FORMAT_BIG_ROWS");  // tests force the chunked path with 0
  return e ? (size_t)strtoull(e, nullptr, 10) : (size_t)20000;
}

size_t big_file_bytes() {  // files above this are parsed by all threads together (giant tints)
  const char* e = getenv("FRS_PARSE_BIG_BYTES");  // read per batch: tests force the chunked path with 1
  return e ? (size_t)strtoull(e, nullptr, 10) : (size_t)(32u << 20);
}

// One split file.  inner_threads > 1 (giant tint): the rows after the header are cut into chunks at
// line ends, parsed in parallel into private TintData blocks, appended in order, then deduped in row
// order -- same arrays, same first error as the sequential walk.
bool parse_split_file(const char* path, TintData& T, int inner_threads) {
  FileBuf fb;
  if (!fb.load(path, T.error, 0)) return false;
  ProfScope ps(PROF_SPLIT);
  const char* s = fb.p;
  const char* end = fb.p + fb.n;
  bool have_header = false;
  RepTable reps;
  auto no_newline = [&](const char* at) {
    T.error = std::string("AttributeError: line without newline does not match (") + path + "): " + line_snip(at, end);
    return false;
  };
  // ---- parallel path ----
  if (inner_threads > 1 && fb.n > 0 && *s == '#') {
    const char* nl = (const char*)memchr(s, '\n', (size_t)(end - s));
    if (!nl) return no_newline(s);
    if (!parse_header_line(s, nl, path, false, T)) return false;
    have_header = true;
    s = nl + 1;
    const int K = inner_threads * 4;
    std::vector<const char*> cut((size_t)K + 1, end);
    cut[0] = s;
    for (int k = 1; k < K; ++k) {
      const char* c = s + (size_t)((end - s) / K) * (size_t)k;
      if (c < cut[(size_t)k - 1]) c = cut[(size_t)k - 1];
      const char* n2 = c < end ? (const char*)memchr(c, '\n', (size_t)(end - c)) : nullptr;
      cut[(size_t)k] = n2 ? n2 + 1 : end;
    }
    std::vector<TintData> part((size_t)K);
    std::vector<int> bad_kind((size_t)K, 0);  // 1 = row error (part.error), 2 = header line inside, 3 = no newline
    std::vector<const char*> bad_at((size_t)K, nullptr);
    parallel_for(K, inner_threads, [&](int k) {
      TintData& C = part[(size_t)k];
      const char* p = cut[(size_t)k];
      const char* pe = cut[(size_t)k + 1];
      while (p < pe) {
        const char* n2 = (const char*)memchr(p, '\n', (size_t)(pe - p));
        if (!n2) { bad_kind[(size_t)k] = 3; bad_at[(size_t)k] = p; return; }
        if (*p == '#') { bad_kind[(size_t)k] = 2; bad_at[(size_t)k] = p; return; }
        if (!parse_read_row(p, n2, true, T.id, C)) { bad_kind[(size_t)k] = 1; return; }
        p = n2 + 1;
      }
    });
    // rows up to the first failing chunk (inclusive: its complete rows precede the failing one) are put
    // in place in parallel, then deduped in row order; the first error in row order wins
    int k_end = K;
    for (int k = 0; k < K; ++k)
      if (bad_kind[(size_t)k]) { k_end = k + 1; break; }
    struct Off { size_t reads, text, ivs, cig; };
    std::vector<Off> off((size_t)k_end + 1);
    {
      Off a{0, 0, 0, 0};
      for (int k = 0; k < k_end; ++k) {
        off[(size_t)k] = a;
        const TintData& C = part[(size_t)k];
        a.reads += C.meta.size();
        a.text += C.text.size();
        a.ivs += C.read_iv_off.back();  // complete rows only (a failing row may have left a partial tail)
        a.cig += C.meta.empty() ? 0 : (size_t)C.riv_cig_off[(size_t)C.read_iv_off.back()];
      }
      off[(size_t)k_end] = a;
      if (a.ivs > (size_t)INT32_MAX || a.cig > (size_t)INT32_MAX || a.text > (size_t)UINT32_MAX) {
        T.error = "tint too large for 32-bit offsets";
        return false;
      }
      T.meta.resize(a.reads);
      T.text.resize(a.text);
      T.riv_ts.resize(a.ivs);
      T.riv_te.resize(a.ivs);
      T.riv_qs.resize(a.ivs);
      T.riv_qe.resize(a.ivs);
      T.riv_cig_off.resize(a.ivs + 1);
      T.cigar.resize(a.cig);
      T.read_iv_off.resize(a.reads + 1);
      T.read_strand.resize(a.reads);
    }
    parallel_for(k_end, inner_threads, [&](int k) {
      TintData& C = part[(size_t)k];
      const Off& o = off[(size_t)k];
      const size_t n_r = C.meta.size(), n_iv = (size_t)C.read_iv_off.back();
      const size_t n_cig = n_r ? (size_t)C.riv_cig_off[n_iv] : 0;
      memcpy(&T.text[o.text], C.text.data(), C.text.size());
      for (size_t r = 0; r < n_r; ++r) {
        ReadMeta m = C.meta[r];
        m.name_off += (uint32_t)o.text;
        m.chr_off += (uint32_t)o.text;
        T.meta[o.reads + r] = m;
        T.read_iv_off[o.reads + r + 1] = C.read_iv_off[r + 1] + (int32_t)o.ivs;
        T.read_strand[o.reads + r] = C.read_strand[r];
      }
      if (n_iv) {
        memcpy(&T.riv_ts[o.ivs], C.riv_ts.data(), n_iv * 4);
        memcpy(&T.riv_te[o.ivs], C.riv_te.data(), n_iv * 4);
        memcpy(&T.riv_qs[o.ivs], C.riv_qs.data(), n_iv * 4);
        memcpy(&T.riv_qe[o.ivs], C.riv_qe.data(), n_iv * 4);
        for (size_t i = 1; i <= n_iv; ++i) T.riv_cig_off[o.ivs + i] = C.riv_cig_off[i] + (int32_t)o.cig;
      }
      if (n_cig) memcpy(&T.cigar[o.cig], C.cigar.data(), n_cig * 4);
      if (!bad_kind[(size_t)k]) C = TintData();
    });
    reps.init(T.meta.size());
    for (size_t r = 0; r < T.meta.size(); ++r)
      if (!reps.add_read(T, r)) return false;
    if (k_end > 0 && bad_kind[(size_t)k_end - 1]) {
      const int k = k_end - 1;
      if (bad_kind[(size_t)k] == 1) { T.error = part[(size_t)k].error; return false; }
      if (bad_kind[(size_t)k] == 2) {
        const char* at = bad_at[(size_t)k];
        const char* n2 = (const char*)memchr(at, '\n', (size_t)(end - at));
        parse_header_line(at, n2 ? n2 : end, path, true, T);  // sets the "repeated" / "len(tints) == 1" error
        if (T.error.empty()) T.error = std::string("AssertionError: assert len(tints) == 1 (freddie_segment.py:699): ") + path;
        return false;
      }
      return no_newline(bad_at[(size_t)k]);
    }
  } else {
    // ---- sequential walk ----
    while (s < end) {
      const char* nl = (const char*)memchr(s, '\n', (size_t)(end - s));
      if (!nl) return no_newline(s);
      if (*s == '#') {
        if (!parse_header_line(s, nl, path, have_header, T)) return false;
        have_header = true;
        reps.init((size_t)std::min<int64_t>(std::max<int64_t>(T.read_count, 0), (int64_t)(fb.n / 16 + 16)));
      } else {
        if (!parse_read_row(s, nl, have_header, T.id, T)) return false;
        if (!reps.add_read(T, T.meta.size() - 1)) return false;
      }
      s = nl + 1;
    }
  }
  if (!have_header) { T.error = std::string("AssertionError: assert len(tints) == 1 (freddie_segment.py:699): ") + path; return false; }
  if ((int64_t)T.meta.size() != T.read_count) { T.error = "AssertionError: assert len(tint['reads']) == tint['read_count'] (freddie_segment.py:164)"; return false; }
  return true;
}

// read_sequence (:174-185): cols 0 and 3 of every line; later duplicates of a rid win
bool parse_reads_file(const char* path, TintData& T, int inner_threads) {
  FileBuf fb;
  if (!fb.load(path, T.error, 1)) return false;
  std::unordered_map<int64_t, std::pair<const char*, uint32_t>> seqs;
  seqs.reserve(T.meta.size() * 2);
  const char* s = fb.p;
  const char* end = fb.p + fb.n;
  {
    ProfScope ps_lines(PROF_READS_LINES);
    // one line -> (rid, sequence); returns 0 ok, 1 ValueError, 2 IndexError
    struct Row { int64_t rid; const char* seq; uint32_t len; };
    auto parse_line = [](const char* ls, const char* e, Row& row) -> int {
      // rstrip(): trailing whitespace
      while (e > ls && (e[-1] == ' ' || e[-1] == '\t' || e[-1] == '\r' || e[-1] == '\n' || e[-1] == '\v' || e[-1] == '\f')) --e;
      // int(line[0]) tolerates surrounding blanks and a sign; split rows only hold digits
      const char* q = ls;
      if (!parse_uint(q, e, row.rid) || (q < e && *q != '\t')) return 1;
      int tabs = 0;
      const char* f3 = nullptr;
      for (const char* c = ls; c < e; ++c)
        if (*c == '\t') { if (++tabs == 3) { f3 = c + 1; break; } }
      if (!f3) return 2;
      const char* f3e = (const char*)memchr(f3, '\t', (size_t)(e - f3));
      if (!f3e) f3e = e;
      row.seq = f3;
      row.len = (uint32_t)(f3e - f3);
      return 0;
    };
    auto fail_line = [&](int code) {
      if (code == 1) T.error = std::string("ValueError: invalid read id in ") + path;
      else T.error = std::string("IndexError: list index out of range (freddie_segment.py:179): ") + path;
      return false;
    };
    if (inner_threads > 1) {
      // giant tint: line chunks in parallel, rows entered in file order (later duplicates of a rid win)
      const int K = inner_threads * 4;
      std::vector<const char*> cut((size_t)K + 1, end);
      cut[0] = s;
      for (int k = 1; k < K; ++k) {
        const char* c = s + (size_t)((end - s) / K) * (size_t)k;
        if (c < cut[(size_t)k - 1]) c = cut[(size_t)k - 1];
        const char* n2 = c < end ? (const char*)memchr(c, '\n', (size_t)(end - c)) : nullptr;
        cut[(size_t)k] = n2 ? n2 + 1 : end;
      }
      std::vector<std::vector<Row>> rows((size_t)K);
      std::vector<int> bad((size_t)K, 0);
      parallel_for(K, inner_threads, [&](int k) {
        const char* p = cut[(size_t)k];
        const char* pe = cut[(size_t)k + 1];
        while (p < pe) {
          const char* n2 = (const char*)memchr(p, '\n', (size_t)(pe - p));
          Row row;
          const int rc = parse_line(p, n2 ? n2 : pe, row);
          if (rc) { bad[(size_t)k] = rc; return; }
          rows[(size_t)k].push_back(row);
          p = n2 ? n2 + 1 : pe;
        }
      });
      for (int k = 0; k < K; ++k) {
        for (const Row& r : rows[(size_t)k]) seqs[r.rid] = std::make_pair(r.seq, r.len);
        if (bad[(size_t)k]) return fail_line(bad[(size_t)k]);
      }
    } else {
      while (s < end) {
        const char* nl = (const char*)memchr(s, '\n', (size_t)(end - s));
        Row row;
        const int rc = parse_line(s, nl ? nl : end, row);
        if (rc) return fail_line(rc);
        seqs[row.rid] = std::make_pair(row.seq, row.len);
        s = nl ? nl + 1 : end;
      }
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
  const size_t n_reads = T.meta.size();
  const int blocks = inner_threads > 1 ? inner_threads * 8 : 1;
  parallel_for(blocks, inner_threads, [&](int bk) {
    const size_t k0 = n_reads * (size_t)bk / (size_t)blocks, k1 = n_reads * ((size_t)bk + 1) / (size_t)blocks;
    for (size_t k = k0; k < k1; ++k) {
      const auto& sq = seqs.find(T.meta[k].rid)->second;
      const size_t w0 = (size_t)T.read_seq_off[k];
      seq_planes(sq.first, sq.second, T.seq_a.data() + w0, T.seq_t.data() + w0);
    }
  });
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
  std::vector<T> own;  // adopt(): the storage of a tint's own vector (single-tint batches: no copy)
  bool borrowed = false;
  ~RawBuf() { if (!borrowed) free(p); }
  void resize(size_t m) {
    if (!borrowed) free(p);
    borrowed = false;
    p = (T*)malloc((m ? m : 1) * sizeof(T));
    n = m;
  }
  void borrow(T* ptr, size_t m) {  // storage owned by somebody who outlives this object (a file mapping)
    if (!borrowed) free(p);
    p = ptr;
    n = m;
    borrowed = true;
  }
  void adopt(std::vector<T>&& v) {
    if (!borrowed) free(p);
    own = std::move(v);
    p = own.data();
    n = own.size();
    borrowed = true;
  }
  T* data() { return p; }
  const T* data() const { return p; }
  size_t size() const { return n; }
};

struct frs_parsed {
  FileBuf* mapping = nullptr;  // frs_packed_read: the large arrays point into the mapped file
  ~frs_parsed() { delete mapping; }
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
  // giant tints (files above big_file_bytes()) one after the other with all threads inside the file,
  // then the many small tints one per task
  std::vector<int> small, big;
  for (int i = 0; i < n; ++i) {
    struct stat st1, st2;
    const size_t b1 = stat(split_paths[i], &st1) == 0 ? (size_t)st1.st_size : 0;
    const size_t b2 = stat(reads_paths[i], &st2) == 0 ? (size_t)st2.st_size : 0;
    ((n_threads > 1 && std::max(b1, b2) > big_file_bytes()) ? big : small).push_back(i);
  }
  auto one = [&](int i, int inner) {
    TintData& T = P->tints[(size_t)i];
    if (!parse_split_file(split_paths[i], T, inner) || !parse_reads_file(reads_paths[i], T, inner)) {
      int exp = -1;
      failed.compare_exchange_strong(exp, i);
    }
  };
  for (int i : big) one(i, n_threads);
  parallel_for((int)small.size(), n_threads, [&](int k) { one(small[(size_t)k], 1); });
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
  const size_t NT = P->tints.size();
  if (NT == 1) {
    // a giant tint is a batch of its own: its arrays ARE the batch (all shifts are zero) -- move them
    TintData& T = P->tints[0];
    P->tint_island_off = {0, (int32_t)isl};
    P->tint_rep_off = {0, (int32_t)reps};
    P->tint_read_off = {0, (int32_t)reads};
    P->island_start = std::move(T.isl_s);
    P->island_sample_off = std::move(T.isl_off);
    P->rep_iv_off = std::move(T.rep_iv_off);
    P->rep_weight = std::move(T.rep_w);
    P->rep_iv_fs = std::move(T.rep_fs);
    P->rep_iv_fe = std::move(T.rep_fe);
    P->read_rep = std::move(T.read_rep);
    P->read_strand = std::move(T.read_strand);
    P->read_len = std::move(T.read_len);
    P->read_iv_off = std::move(T.read_iv_off);
    P->read_seq_off = std::move(T.read_seq_off);
    P->riv_ts = std::move(T.riv_ts);
    P->riv_te = std::move(T.riv_te);
    P->riv_qs = std::move(T.riv_qs);
    P->riv_qe = std::move(T.riv_qe);
    P->riv_cig_off = std::move(T.riv_cig_off);
    P->cigar.adopt(std::move(T.cigar));
    P->seq_a.adopt(std::move(T.seq_a));
    P->seq_t.adopt(std::move(T.seq_t));
    delete ps_cat;
    ps_cat = nullptr;
    if (g_prof.on) {
      fprintf(stderr, "[frs host profile] thread-seconds: load %.3f  split rows %.3f  reads lines %.3f  bit-planes %.3f  concat %.3f (moved)\n",
              g_prof.ns[0] * 1e-9, (g_prof.ns[1]) * 1e-9, g_prof.ns[2] * 1e-9, g_prof.ns[3] * 1e-9, g_prof.ns[4] * 1e-9);
      for (auto& x : g_prof.ns) x = 0;
    }
    *out = P;
    return 0;
  }
  // every tint's slice of every array is known from the prefix sums: size once, copy in parallel
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

// ---------------------------------------------------------------------------------------------
// Packed side-channel: a parsed batch as ONE binary file, so that a pipeline can hand
// tints to this stage without the TSV round trip (text stays the default).  Layout, little-endian:
//   "FRSBATC1" | u64 n_sections | n_sections x (u64 offset, u64 bytes) | sections, 64-byte aligned
// Sections 0..21 are the arrays of frs_batch in declaration order; then per tint: id (i64), chr offsets
// (u64[T+1]) + chr text; per read: rid (i64), the row's own tint column (i64), name offsets (u64[N+1]) +
// names, chr offsets (u64[N+1]) + chr text -- everything frs_format_tints prints back.
// ---------------------------------------------------------------------------------------------
namespace {
const char PACK_MAGIC[8] = {'F', 'R', 'S', 'B', 'A', 'T', 'C', '1'};
enum { PK_BATCH0 = 0, PK_TINT_ID = 22, PK_TINT_CHR_OFF, PK_TINT_CHR, PK_READ_RID, PK_READ_TINT, PK_NAME_OFF, PK_NAMES,
       PK_RCHR_OFF, PK_RCHR, PK_SECTIONS };
struct Sec { const void* p; uint64_t bytes; };
int pack_fail(char* err, size_t cap, int code, const std::string& m) {
  if (err && cap) snprintf(err, cap, "%s", m.c_str());
  return code;
}
}  // namespace

int frs_packed_write(const frs_parsed* P, const char* path, char* err, size_t err_cap) {
  if (err && err_cap) err[0] = 0;
  if (!P || !path) return pack_fail(err, err_cap, FRS_ERR_ARG, "frs_packed_write: NULL argument");
  const size_t T = P->tints.size(), N = P->read_rep.size();
  std::vector<int64_t> tint_id(T), rid(N), rtint(N);
  std::vector<uint64_t> tchr_off(T + 1, 0), name_off(N + 1, 0), rchr_off(N + 1, 0);
  std::string tchr, names, rchr;
  size_t r = 0;
  for (size_t t = 0; t < T; ++t) {
    const TintData& D = P->tints[t];
    tint_id[t] = D.id;
    tchr += D.chr;
    tchr_off[t + 1] = tchr.size();
    for (const ReadMeta& m : D.meta) {
      rid[r] = m.rid;
      rtint[r] = m.tint;
      names.append(D.text, m.name_off, m.name_len);
      name_off[r + 1] = names.size();
      rchr.append(D.text, m.chr_off, m.chr_len);
      rchr_off[r + 1] = rchr.size();
      ++r;
    }
  }
  if (r != N) return pack_fail(err, err_cap, FRS_ERR_STATE, "frs_packed_write: read tables out of step");
  Sec sec[PK_SECTIONS] = {
      {P->tint_island_off.data(), P->tint_island_off.size() * 4}, {P->tint_rep_off.data(), P->tint_rep_off.size() * 4},
      {P->tint_read_off.data(), P->tint_read_off.size() * 4}, {P->island_start.data(), P->island_start.size() * 4},
      {P->island_sample_off.data(), P->island_sample_off.size() * 4}, {P->rep_iv_off.data(), P->rep_iv_off.size() * 4},
      {P->rep_weight.data(), P->rep_weight.size() * 4}, {P->rep_iv_fs.data(), P->rep_iv_fs.size() * 4},
      {P->rep_iv_fe.data(), P->rep_iv_fe.size() * 4}, {P->read_rep.data(), P->read_rep.size() * 4},
      {P->read_strand.data(), P->read_strand.size()}, {P->read_len.data(), P->read_len.size() * 4},
      {P->read_iv_off.data(), P->read_iv_off.size() * 4}, {P->read_seq_off.data(), P->read_seq_off.size() * 8},
      {P->riv_ts.data(), P->riv_ts.size() * 4}, {P->riv_te.data(), P->riv_te.size() * 4},
      {P->riv_qs.data(), P->riv_qs.size() * 4}, {P->riv_qe.data(), P->riv_qe.size() * 4},
      {P->riv_cig_off.data(), P->riv_cig_off.size() * 4}, {P->cigar.data(), P->cigar.size() * 4},
      {P->seq_a.data(), P->seq_a.size() * 4}, {P->seq_t.data(), P->seq_t.size() * 4},
      {tint_id.data(), T * 8}, {tchr_off.data(), (T + 1) * 8}, {tchr.data(), tchr.size()},
      {rid.data(), N * 8}, {rtint.data(), N * 8}, {name_off.data(), (N + 1) * 8}, {names.data(), names.size()},
      {rchr_off.data(), (N + 1) * 8}, {rchr.data(), rchr.size()},
  };
  FILE* f = fopen(path, "wb");
  if (!f) return pack_fail(err, err_cap, FRS_ERR_IO, std::string("cannot open ") + path + ": " + strerror(errno));
  std::vector<uint64_t> table(2 * PK_SECTIONS);
  uint64_t at = 8 + 8 + 16 * (uint64_t)PK_SECTIONS;
  for (int k = 0; k < PK_SECTIONS; ++k) {
    at = (at + 63) & ~(uint64_t)63;
    table[2 * (size_t)k] = at;
    table[2 * (size_t)k + 1] = sec[k].bytes;
    at += sec[k].bytes;
  }
  const uint64_t ns = PK_SECTIONS;
  bool ok = fwrite(PACK_MAGIC, 1, 8, f) == 8 && fwrite(&ns, 8, 1, f) == 1 && fwrite(table.data(), 8, table.size(), f) == table.size();
  uint64_t pos = 8 + 8 + 16 * (uint64_t)PK_SECTIONS;
  static const char zeros[64] = {0};
  for (int k = 0; k < PK_SECTIONS && ok; ++k) {
    const uint64_t pad = table[2 * (size_t)k] - pos;
    ok = (pad == 0 || fwrite(zeros, 1, (size_t)pad, f) == pad) &&
         (sec[k].bytes == 0 || fwrite(sec[k].p, 1, (size_t)sec[k].bytes, f) == sec[k].bytes);
    pos = table[2 * (size_t)k] + sec[k].bytes;
  }
  ok = (fclose(f) == 0) && ok;
  if (!ok) return pack_fail(err, err_cap, FRS_ERR_IO, std::string("short write to ") + path);
  return 0;
}

// The SEGMENT twin ("FRSSEGM1"): what frs_format_tints prints, as arrays -- for a consumer that would
// otherwise re-parse segment_*.tsv with regexes (freddie_cluster.py:119-172).  Sections: tint id, tint chr
// offsets + text, tint_read_off, tint_rep_off, tint_final_off, final_pos, tint_digit_off, digits (ASCII,
// rep-major), read_rep, read rid, read tint column, name offsets + names, chr offsets + chrs, read_strand,
// read_head (8 x i32, FRS_HEAD_*), read_gap_off, gap_rec ((l1, f2, size) triples).
int frs_packed_write_segment(const frs_parsed* P, const frs_result* R, const char* path, char* err, size_t err_cap) {
  if (err && err_cap) err[0] = 0;
  if (!P || !R || !path) return pack_fail(err, err_cap, FRS_ERR_ARG, "frs_packed_write_segment: NULL argument");
  static const char MAGIC[8] = {'F', 'R', 'S', 'S', 'E', 'G', 'M', '1'};
  const size_t T = P->tints.size(), N = P->read_rep.size();
  std::vector<int64_t> tint_id(T), rid(N), rtint(N);
  std::vector<uint64_t> tchr_off(T + 1, 0), name_off(N + 1, 0), rchr_off(N + 1, 0);
  std::string tchr, names, rchr;
  size_t r = 0;
  for (size_t t = 0; t < T; ++t) {
    const TintData& D = P->tints[t];
    tint_id[t] = D.id;
    tchr += D.chr;
    tchr_off[t + 1] = tchr.size();
    for (const ReadMeta& m : D.meta) {
      rid[r] = m.rid;
      rtint[r] = m.tint;
      names.append(D.text, m.name_off, m.name_len);
      name_off[r + 1] = names.size();
      rchr.append(D.text, m.chr_off, m.chr_len);
      rchr_off[r + 1] = rchr.size();
      ++r;
    }
  }
  if (r != N) return pack_fail(err, err_cap, FRS_ERR_STATE, "frs_packed_write_segment: read tables out of step");
  const size_t NF = (size_t)R->tint_final_off[T], ND = (size_t)R->tint_digit_off[T], NG = (size_t)R->read_gap_off[N];
  const Sec sec[] = {
      {tint_id.data(), T * 8}, {tchr_off.data(), (T + 1) * 8}, {tchr.data(), tchr.size()},
      {P->tint_read_off.data(), (T + 1) * 4}, {P->tint_rep_off.data(), (T + 1) * 4},
      {R->tint_final_off, (T + 1) * 4}, {R->final_pos, NF * 4}, {R->tint_digit_off, (T + 1) * 8}, {R->digits, ND},
      {P->read_rep.data(), N * 4}, {rid.data(), N * 8}, {rtint.data(), N * 8},
      {name_off.data(), (N + 1) * 8}, {names.data(), names.size()}, {rchr_off.data(), (N + 1) * 8}, {rchr.data(), rchr.size()},
      {P->read_strand.data(), N}, {R->read_head, N * 32}, {R->read_gap_off, (N + 1) * 4}, {R->gap_rec, NG * 12},
  };
  const int NS = (int)(sizeof sec / sizeof sec[0]);
  FILE* f = fopen(path, "wb");
  if (!f) return pack_fail(err, err_cap, FRS_ERR_IO, std::string("cannot open ") + path + ": " + strerror(errno));
  std::vector<uint64_t> table(2 * (size_t)NS);
  uint64_t at = 16 + 16 * (uint64_t)NS;
  for (int k = 0; k < NS; ++k) {
    at = (at + 63) & ~(uint64_t)63;
    table[2 * (size_t)k] = at;
    table[2 * (size_t)k + 1] = sec[k].bytes;
    at += sec[k].bytes;
  }
  const uint64_t ns = (uint64_t)NS;
  bool ok = fwrite(MAGIC, 1, 8, f) == 8 && fwrite(&ns, 8, 1, f) == 1 && fwrite(table.data(), 8, table.size(), f) == table.size();
  uint64_t pos = 16 + 16 * (uint64_t)NS;
  static const char zeros[64] = {0};
  for (int k = 0; k < NS && ok; ++k) {
    const uint64_t pad = table[2 * (size_t)k] - pos;
    ok = (pad == 0 || fwrite(zeros, 1, (size_t)pad, f) == pad) &&
         (sec[k].bytes == 0 || fwrite(sec[k].p, 1, (size_t)sec[k].bytes, f) == sec[k].bytes);
    pos = table[2 * (size_t)k] + sec[k].bytes;
  }
  ok = (fclose(f) == 0) && ok;
  if (!ok) return pack_fail(err, err_cap, FRS_ERR_IO, std::string("short write to ") + path);
  return 0;
}

int frs_packed_read(const char* path, frs_parsed** out, char* err, size_t err_cap) {
  if (err && err_cap) err[0] = 0;
  if (!path || !out) return pack_fail(err, err_cap, FRS_ERR_ARG, "frs_packed_read: NULL argument");
  std::unique_ptr<FileBuf> fbp(new FileBuf());
  FileBuf& fb = *fbp;
  std::string e;
  if (!fb.load(path, e, 0, true)) return pack_fail(err, err_cap, FRS_ERR_IO, e);
  const std::string bad = std::string("not a packed batch (FRSBATC1): ") + path;
  if (fb.n < 16 || memcmp(fb.p, PACK_MAGIC, 8) != 0) return pack_fail(err, err_cap, FRS_ERR_ARG, bad);
  uint64_t ns;
  memcpy(&ns, fb.p + 8, 8);
  if (ns != PK_SECTIONS || fb.n < 16 + 16 * (size_t)PK_SECTIONS) return pack_fail(err, err_cap, FRS_ERR_ARG, bad);
  uint64_t table[2 * PK_SECTIONS];
  memcpy(table, fb.p + 16, sizeof table);
  for (int k = 0; k < PK_SECTIONS; ++k)
    if (table[2 * k] > fb.n || table[2 * k + 1] > fb.n - table[2 * k]) return pack_fail(err, err_cap, FRS_ERR_ARG, bad + " (section out of range)");
  auto bytes = [&](int k) { return (size_t)table[2 * k + 1]; };
  auto at = [&](int k) { return (const char*)fb.p + table[2 * k]; };
  // element counts and their consistency
  if (bytes(0) < 8 || bytes(0) % 4) return pack_fail(err, err_cap, FRS_ERR_ARG, bad);
  const size_t T = bytes(0) / 4 - 1;
  const size_t NI = bytes(3) / 4, NR = bytes(6) / 4, NRI = bytes(7) / 4, N = bytes(9) / 4, NV = bytes(14) / 4, NC = bytes(19) / 4,
               NW = bytes(20) / 4;
  const size_t want[PK_SECTIONS] = {(T + 1) * 4, (T + 1) * 4, (T + 1) * 4, NI * 4, (NI + 1) * 4, (NR + 1) * 4, NR * 4, NRI * 4, NRI * 4,
                                    N * 4, N, N * 4, (N + 1) * 4, (N + 1) * 8, NV * 4, NV * 4, NV * 4, NV * 4, (NV + 1) * 4, NC * 4,
                                    NW * 4, NW * 4, T * 8, (T + 1) * 8, bytes(PK_TINT_CHR), N * 8, N * 8, (N + 1) * 8,
                                    bytes(PK_NAMES), (N + 1) * 8, bytes(PK_RCHR)};
  for (int k = 0; k < PK_SECTIONS; ++k)
    if (bytes(k) != want[k]) return pack_fail(err, err_cap, FRS_ERR_ARG, bad + " (inconsistent section sizes)");
  frs_parsed* P = new frs_parsed();
  auto fill = [&](auto& v, int k) {
    using E = typename std::remove_reference<decltype(v)>::type::value_type;
    v.resize(bytes(k) / sizeof(E));
    if (bytes(k)) memcpy(v.data(), at(k), bytes(k));
  };
  fill(P->tint_island_off, 0); fill(P->tint_rep_off, 1); fill(P->tint_read_off, 2); fill(P->island_start, 3);
  fill(P->island_sample_off, 4); fill(P->rep_iv_off, 5); fill(P->rep_weight, 6); fill(P->rep_iv_fs, 7); fill(P->rep_iv_fe, 8);
  fill(P->read_rep, 9); fill(P->read_strand, 10); fill(P->read_len, 11); fill(P->read_iv_off, 12); fill(P->read_seq_off, 13);
  fill(P->riv_ts, 14); fill(P->riv_te, 15); fill(P->riv_qs, 16); fill(P->riv_qe, 17); fill(P->riv_cig_off, 18);
  if (fb.mapped) {  // the three large arrays stay in the page cache: no copy
    P->cigar.borrow((uint32_t*)at(19), NC);
    P->seq_a.borrow((uint32_t*)at(20), NW);
    P->seq_t.borrow((uint32_t*)at(21), NW);
  } else {
    P->cigar.resize(NC); if (NC) memcpy(P->cigar.data(), at(19), NC * 4);
    P->seq_a.resize(NW); if (NW) memcpy(P->seq_a.data(), at(20), NW * 4);
    P->seq_t.resize(NW); if (NW) memcpy(P->seq_t.data(), at(21), NW * 4);
  }
  // offset tables must be monotone and end at the counts (frs_upload checks the batch arrays again)
  const uint64_t* tco = (const uint64_t*)at(PK_TINT_CHR_OFF);
  const uint64_t* no = (const uint64_t*)at(PK_NAME_OFF);
  const uint64_t* co = (const uint64_t*)at(PK_RCHR_OFF);
  bool mono = tco[0] == 0 && no[0] == 0 && co[0] == 0 && tco[T] == bytes(PK_TINT_CHR) && no[N] == bytes(PK_NAMES) &&
              co[N] == bytes(PK_RCHR) && P->tint_read_off[0] == 0 && (size_t)P->tint_read_off[T] == N;
  for (size_t t = 0; t < T && mono; ++t) mono = tco[t] <= tco[t + 1] && P->tint_read_off[t] <= P->tint_read_off[t + 1];
  for (size_t i = 0; i < N && mono; ++i) mono = no[i] <= no[i + 1] && co[i] <= co[i + 1] && no[i + 1] - no[i] <= 254;
  if (!mono) { delete P; return pack_fail(err, err_cap, FRS_ERR_ARG, bad + " (offset tables)"); }
  const int64_t* tid = (const int64_t*)at(PK_TINT_ID);
  const int64_t* rid = (const int64_t*)at(PK_READ_RID);
  const int64_t* rti = (const int64_t*)at(PK_READ_TINT);
  P->tints.resize(T);
  for (size_t t = 0; t < T; ++t) {
    TintData& D = P->tints[t];
    D.id = tid[t];
    D.chr.assign(at(PK_TINT_CHR) + tco[t], (size_t)(tco[t + 1] - tco[t]));
    const size_t r0 = (size_t)P->tint_read_off[t], r1 = (size_t)P->tint_read_off[t + 1];
    D.read_count = (int64_t)(r1 - r0);
    D.meta.resize(r1 - r0);
    D.text.reserve((size_t)(no[r1] - no[r0] + co[r1] - co[r0]));
    for (size_t i = r0; i < r1; ++i) {
      ReadMeta& m = D.meta[i - r0];
      m.rid = rid[i];
      m.tint = rti[i];
      m.strand = P->read_strand[i] ? '-' : '+';
      m.name_off = (uint32_t)D.text.size();
      m.name_len = (uint32_t)(no[i + 1] - no[i]);
      D.text.append(at(PK_NAMES) + no[i], m.name_len);
      m.chr_off = (uint32_t)D.text.size();
      m.chr_len = (uint32_t)(co[i + 1] - co[i]);
      D.text.append(at(PK_RCHR) + co[i], m.chr_len);
    }
  }
  if (fb.mapped) P->mapping = fbp.release();
  *out = P;
  return 0;
}

// run_segment output (:715-731): "#chr\tid\tpos,pos,...\n" then one row per read in file order;
// gap strings sorted as python strings (:472), each followed by a comma.
int frs_format_tints(const frs_parsed* P, const frs_result* R, const char* const* out_paths,
                     const char* const* log_paths, int n_threads, char* err, size_t err_cap) {
  if (err && err_cap) err[0] = 0;
  if (!P || !R || !out_paths) return FRS_ERR_ARG;
  const int n = (int)P->tints.size();
  std::vector<std::string> errors((size_t)n);
  // rows [k0, k1) of tint t, appended to o
  auto format_rows = [&](int t, size_t k0, size_t k1, std::string& o) {
    const TintData& T = P->tints[(size_t)t];
    const int32_t f0 = R->tint_final_off[t], f1 = R->tint_final_off[t + 1];
    const int64_t S = f1 - f0 - 1;
    const int64_t d0 = R->tint_digit_off[t];
    const int32_t rep0 = P->tint_rep_off[(size_t)t];
    const int32_t r0 = P->tint_read_off[(size_t)t];
    char num[32];
    std::vector<std::string> gaps;
    for (size_t k = k0; k < k1; ++k) {
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
  };
  // one tint: header + rows (in `inner` chunks formatted by `inner` threads for giant tints) -> file
  auto format_tint = [&](int t, int inner) {
    const TintData& T = P->tints[(size_t)t];
    const int32_t f0 = R->tint_final_off[t], f1 = R->tint_final_off[t + 1];
    const int64_t S = f1 - f0 - 1;
    const size_t n_rows = T.meta.size();
    const int K = inner > 1 ? inner * 2 : 1;
    std::vector<std::string> part((size_t)K);
    parallel_for(K, inner, [&](int c) {
      const size_t k0 = n_rows * (size_t)c / (size_t)K, k1 = n_rows * ((size_t)c + 1) / (size_t)K;
      std::string& o = part[(size_t)c];
      o.reserve((size_t)((k1 - k0) * (size_t)(S + 96) + (c == 0 ? (size_t)(f1 - f0) * 11 + 64 : 0)));
      if (c == 0) {
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
      }
      format_rows(t, k0, k1, o);
    });
    FILE* f = fopen(out_paths[t], "wb");
    if (!f) { errors[(size_t)t] = std::string("cannot open ") + out_paths[t] + ": " + strerror(errno); return; }
    for (const std::string& o : part)
      if (fwrite(o.data(), 1, o.size(), f) != o.size()) { errors[(size_t)t] = std::string("short write to ") + out_paths[t]; break; }
    fclose(f);
    if (log_paths && log_paths[t]) {
      FILE* l = fopen(log_paths[t], "wb");  // the reference leaves an empty .log per tint (:695,:734)
      if (l) fclose(l);
      else errors[(size_t)t] = std::string("cannot open ") + log_paths[t];
    }
  };
  // giant tints one after the other with all threads inside, the rest one tint per task
  std::vector<int> small, big;
  for (int t = 0; t < n; ++t)
    ((n_threads > 1 && P->tints[(size_t)t].meta.size() > format_big_rows()) ? big : small).push_back(t);
  for (int t : big) format_tint(t, n_threads);
  parallel_for((int)small.size(), n_threads, [&](int k) { format_tint(small[(size_t)k], 1); });
  for (int t = 0; t < n; ++t)
    if (!errors[(size_t)t].empty()) {
      if (err) snprintf(err, err_cap, "%s", errors[(size_t)t].c_str());
      return FRS_ERR_IO;
    }
  return 0;
}

}  // extern "C"
