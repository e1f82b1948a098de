// Host-side I/O of the inference path, multi-threaded, behind the C ABI (include/m6anet_b200.h):
//   m6a_ingest_parts    data.json byte ranges -> flat normalised float32 buffers
//                       (reference utils/data_utils.py:152-248 _load_data/load_data/__getitem__/get_norm_factor,
//                        :395-427 replicate concatenation; file format utils/dataprep_utils.py:473-485)
//   m6a_write_site_csv / m6a_write_indiv_csv   the reference row formats (utils/inference_utils.py:59-67)
// Numbers are parsed with std::from_chars (correctly rounded, == Python float()), normalised in double and
// rounded once to float32, so the buffers equal the reference's NanopolishDS output bit for bit.
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <charconv>
#include <cmath>
#include <string>
#include <thread>
#include <vector>

#include "../../include/m6anet_b200.h"

namespace {

inline int base_code(char c) {
  switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default: return -1;
  }
}

// 5-mer at s -> base-4 code in [0, 1024), -1 if a letter is not ACGT
inline int kmer5_code(const char* s) {
  int code = 0;
  for (int i = 0; i < 5; ++i) {
    const int b = base_code(s[i]);
    if (b < 0) return -1;
    code = code * 4 + b;
  }
  return code;
}

struct IngestCtx {
  const char* const* paths;
  const m6a_part_t* parts;
  int64_t n_parts;
  int n_flank;
  const double* norm_mean;
  const double* norm_std;
  const int32_t* kmer_id;
  float* feats;
  int64_t* read_ids;
  int32_t* kmer_idx;
  // optional keys of the sites (m6a_ingest_parts_keyed): the line of every part must carry its site's transcript id and
  // position (the reference indexes json.loads(line)[tx_id][str(tx_pos)] and raises KeyError on a stale index)
  const char* tx_buf = nullptr;
  const int64_t* tx_off = nullptr;
  const int64_t* tx_pos = nullptr;
  std::vector<int32_t> part_ids;     // five-mer ids seen by every part (replicates of a site must agree)
  std::vector<int> fds;
  std::atomic<int64_t> next{0};
  std::atomic<int64_t> bad{-1};
  std::atomic<int> status{M6A_OK};
};

void fail(IngestCtx& c, int64_t part, int status) {
  int expected = M6A_OK;
  if (c.status.compare_exchange_strong(expected, status)) c.bad.store(part);
}

// ---- strict scanner for one site line --------------------------------------------------------------------------
// The line must be exactly what `m6anet dataprep` writes (reference utils/dataprep_utils.py:473-480) and what the
// reference reads back with json.loads (utils/data_utils.py:182-189):
//     {"<tx>":{"<pos>":{"<KMER>":[[v,...,v],[v,...,v],...]}}}   + optional trailing whitespace
// Anything json.loads would reject is rejected here too (M6A_EPARSE) -- a truncated byte range, a missing bracket, a
// malformed number -- so a corrupted data.json can not be scored silently.
inline const char* skip_ws(const char* r, const char* e) {
  while (r < e && (*r == ' ' || *r == '\t' || *r == '\n' || *r == '\r')) ++r;
  return r;
}
// expects (after whitespace) the character ch; returns the position after it or nullptr
inline const char* expect(const char* r, const char* e, char ch) {
  r = skip_ws(r, e);
  return (r < e && *r == ch) ? r + 1 : nullptr;
}
// expects a string without escapes; returns the position after the closing quote, [*s0, *s1) = contents.
// Non-ASCII bytes must form valid UTF-8 (json.loads decodes the line as UTF-8 and refuses anything else).
inline const char* expect_string(const char* r, const char* e, const char** s0, const char** s1) {
  r = expect(r, e, '"');
  if (!r) return nullptr;
  *s0 = r;
  while (r < e && *r != '"') {
    const unsigned char c = static_cast<unsigned char>(*r);
    if (c == '\\' || c < 0x20) return nullptr;          // escapes never occur in ids / k-mers
    if (c < 0x80) {
      ++r;
      continue;
    }
    int extra;                                            // continuation bytes of the sequence
    unsigned cp;
    if (c >= 0xC2 && c <= 0xDF) { extra = 1; cp = c & 0x1Fu; }
    else if (c >= 0xE0 && c <= 0xEF) { extra = 2; cp = c & 0x0Fu; }
    else if (c >= 0xF0 && c <= 0xF4) { extra = 3; cp = c & 0x07u; }
    else return nullptr;
    if (e - r <= extra) return nullptr;
    for (int i = 1; i <= extra; ++i) {
      const unsigned char t = static_cast<unsigned char>(r[i]);
      if ((t & 0xC0u) != 0x80u) return nullptr;
      cp = (cp << 6) | (t & 0x3Fu);
    }
    if ((extra == 2 && cp < 0x800u) || (extra == 3 && (cp < 0x10000u || cp > 0x10FFFFu))) return nullptr;   // overlong / range
    // (surrogate code points U+D800..DFFF are let through: json.loads decodes with 'surrogatepass')
    r += extra + 1;
  }
  if (r >= e) return nullptr;
  *s1 = r;
  return r + 1;
}
// One JSON number (plus the NaN / Infinity / -Infinity literals Python's json writes and reads); nullptr if malformed.
// Grammar (RFC 8259): -?(0|[1-9][0-9]*)(\.[0-9]+)?([eE][+-]?[0-9]+)?  -- validated while the digits are accumulated.
// Fast path (W. Clinger): no exponent, at most 19 digits, decimal mantissa w <= 2^53 and at most 22 fraction digits =>
// w and 10^k are exact doubles and the IEEE division w / 10^k is the correctly rounded value, i.e. what Python's float()
// returns.  Everything else goes through std::from_chars (correctly rounded as well).
// The buffer is NUL-terminated at e, so single look-aheads never leave it.
inline bool is_digit(char ch) { return static_cast<unsigned>(static_cast<unsigned char>(ch) - '0') <= 9u; }
constexpr double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                               1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
inline const char* parse_number(const char* r, const char* e, double* v) {
  const bool neg = (*r == '-');
  const char* d = r + neg;
  if (!is_digit(*d)) {
    if (e - r >= 3 && !strncmp(r, "NaN", 3)) { *v = NAN; return r + 3; }
    if (e - d >= 8 && !strncmp(d, "Infinity", 8)) { *v = neg ? -INFINITY : INFINITY; return d + 8; }
    return nullptr;
  }
  if (*d == '0' && is_digit(d[1])) return nullptr;      // leading zero
  uint64_t w = 0;
  const char* q = d;
  for (; is_digit(*q); ++q) w = w * 10u + static_cast<unsigned>(*q - '0');   // may wrap beyond 19 digits: checked below
  const char* frac0 = q;
  if (*q == '.') {
    ++q;
    frac0 = q;
    if (!is_digit(*q)) return nullptr;                   // "1." / "1.e5"
    for (; is_digit(*q); ++q) w = w * 10u + static_cast<unsigned>(*q - '0');
  }
  const int n_frac = static_cast<int>(q - frac0);
  const bool has_exp = (*q == 'e' || *q == 'E');
  if (!has_exp && (q - d) <= 19 && w <= (1ull << 53) && n_frac <= 22) {   // (q - d) <= 19 characters: w did not wrap
    const double x = static_cast<double>(w) / kPow10[n_frac];
    *v = neg ? -x : x;
    return q;
  }
  if (has_exp) {                                          // [eE][+-]?[0-9]+
    const char* x = q + 1;
    if (*x == '+' || *x == '-') ++x;
    if (!is_digit(*x)) return nullptr;
  }
  const auto res = std::from_chars(r, e, *v);
  // out of range: json.loads gives +-inf / 0.0 there; dataprep never writes such values, so refuse rather than guess
  if (res.ec != std::errc()) return nullptr;
  return res.ptr;
}
// whitespace is legal JSON between tokens but dataprep writes none: test before calling
inline const char* ws(const char* r, const char* e) {
  return static_cast<unsigned char>(*r) <= ' ' ? skip_ws(r, e) : r;
}

bool parse_part(IngestCtx& c, int64_t pi, std::vector<char>& buf) {
  const m6a_part_t& p = c.parts[pi];
  const int64_t len = p.end - p.start;
  if (len <= 0 || p.file < 0 || p.n_rows < 0) return false;
  buf.resize(static_cast<size_t>(len) + 1);
  int64_t got = 0;
  while (got < len) {
    const ssize_t r = pread(c.fds[p.file], buf.data() + got, static_cast<size_t>(len - got), p.start + got);
    if (r < 0) {
      fail(c, pi, M6A_EIO);
      return true;  // status already set
    }
    if (r == 0) return false;   // the byte range runs past the end of the file: data.info does not match data.json
    got += r;
  }
  buf[len] = 0;
  const char* e = buf.data() + len;
  // {"tx":{"pos":{"KMER":[
  const char *t0, *t1, *q0 = nullptr, *q1 = nullptr, *k0, *k1;
  const char* r = expect(buf.data(), e, '{');
  if (r) r = expect_string(r, e, &t0, &t1);
  if (r) r = expect(r, e, ':');
  if (r) r = expect(r, e, '{');
  if (r) r = expect_string(r, e, &q0, &q1);
  if (r && c.tx_buf != nullptr) {                       // keys of the line == keys of the data.info row
    const int64_t a = c.tx_off[p.site], b = c.tx_off[p.site + 1];
    if (t1 - t0 != b - a || memcmp(t0, c.tx_buf + a, static_cast<size_t>(b - a)) != 0) return false;
    char num[24];
    const int nl = snprintf(num, sizeof num, "%lld", static_cast<long long>(c.tx_pos[p.site]));
    if (q1 - q0 != nl || memcmp(q0, num, static_cast<size_t>(nl)) != 0) return false;
  }
  if (r) r = expect(r, e, ':');
  if (r) r = expect(r, e, '{');
  if (r) r = expect_string(r, e, &k0, &k1);
  if (r) r = expect(r, e, ':');
  if (r) r = expect(r, e, '[');
  if (!r) return false;
  const int klen = static_cast<int>(k1 - k0);
  if (klen < 5 || ((klen - 5) & 1)) return false;
  const int T = (klen - 5) / 2, n = c.n_flank;       // flanks in the file / wanted
  if (n > T || T > 5) return false;
  const int n_pos = 2 * n + 1, n_sig = 3 * n_pos, row_w = 3 * (2 * T + 1) + 1;
  // five-mers of the centred (5+2n)-mer, normalisation vectors, ids
  double mean[33], stdv[33];
  int32_t ids[11];
  for (int j = 0; j < n_pos; ++j) {
    const int code = kmer5_code(k0 + (T - n) + j);
    if (code < 0 || c.kmer_id[code] < 0) return false;
    ids[j] = c.kmer_id[code];
    for (int i = 0; i < 3; ++i) {
      mean[3 * j + i] = c.norm_mean[3 * code + i];
      stdv[3 * j + i] = c.norm_std[3 * code + i];
      if (std::isnan(mean[3 * j + i]) || std::isnan(stdv[3 * j + i])) return false;   // k-mer missing from norm factors
    }
  }
  const int col0 = (T - n) * 3;   // selected signal columns are contiguous: [(T-n)*3, (T+n+1)*3)
  // rows: [v,...,v] separated by ',' and closed by ']'
  double vals[40];
  int64_t row = 0;
  r = skip_ws(r, e);
  if (r < e && *r == ']') {
    ++r;                                  // empty list
  } else {
    for (;;) {
      r = expect(r, e, '[');
      if (!r) return false;
      int nv = 0;
      for (;;) {
        r = ws(r, e);
        if (nv >= 40 || r >= e) return false;
        r = parse_number(r, e, &vals[nv]);
        if (!r) return false;
        ++nv;
        r = ws(r, e);
        if (*r == ',') { ++r; continue; }      // (*e == 0: running off the end fails the three tests)
        if (*r == ']') { ++r; break; }
        return false;
      }
      if (nv != row_w || row >= p.n_rows) return false;
      if (!(std::fabs(vals[row_w - 1]) < 9.2e18)) return false;      // read id must fit int64 (NaN / inf / 1e300 do not)
      float* out = c.feats + (p.row_off + row) * n_sig;
      for (int k = 0; k < n_sig; ++k) out[k] = static_cast<float>((vals[col0 + k] - mean[k]) / stdv[k]);
      c.read_ids[p.row_off + row] = static_cast<int64_t>(vals[row_w - 1]);
      ++row;
      r = skip_ws(r, e);
      if (r >= e) return false;
      if (*r == ',') { ++r; continue; }
      if (*r == ']') { ++r; break; }
      return false;
    }
  }
  // }}} and nothing but whitespace up to the end of the byte range
  for (int k = 0; k < 3 && r; ++k) r = expect(r, e, '}');
  if (!r || skip_ws(r, e) != e) return false;
  if (row != p.n_rows) return false;
  if (p.first_of_site) {
    int32_t* kid = c.kmer_idx + p.site * n_pos;
    for (int j = 0; j < n_pos; ++j) kid[j] = ids[j];
  }
  if (!c.part_ids.empty())
    for (int j = 0; j < n_pos; ++j) c.part_ids[static_cast<size_t>(pi) * 11 + j] = ids[j];
  return true;
}

void ingest_worker(IngestCtx* c) {
  std::vector<char> buf;
  for (;;) {
    const int64_t i = c->next.fetch_add(1);
    if (i >= c->n_parts || c->status.load() != M6A_OK) return;
    if (!parse_part(*c, i, buf)) fail(*c, i, M6A_EPARSE);
  }
}

int n_workers(int32_t n_threads, int64_t items) {
  int n = n_threads > 0 ? n_threads : static_cast<int>(std::thread::hardware_concurrency());
  if (n < 1) n = 1;
  if (n > 256) n = 256;
  if (static_cast<int64_t>(n) > items) n = static_cast<int>(items > 0 ? items : 1);
  return n;
}

bool write_all(int fd, const char* p, size_t n) {
  while (n > 0) {
    const ssize_t w = write(fd, p, n);
    if (w <= 0) return false;
    p += w;
    n -= static_cast<size_t>(w);
  }
  return true;
}

// Growable output buffer without zero-fill (std::string::resize would touch every byte twice).
struct OutBuf {
  char* p = nullptr;
  size_t len = 0, cap = 0;
  OutBuf() = default;
  OutBuf(const OutBuf&) = delete;
  OutBuf& operator=(const OutBuf&) = delete;
  ~OutBuf() { free(p); }
  bool ensure(size_t extra) {          // room for `extra` more bytes
    if (len + extra <= cap) return true;
    size_t want = cap ? cap * 2 : (1u << 16);
    while (want < len + extra) want *= 2;
    char* q = static_cast<char*>(realloc(p, want));
    if (!q) return false;
    p = q;
    cap = want;
    return true;
  }
  void append(const char* src, size_t n) {
    memcpy(p + len, src, n);
    len += n;
  }
};

// Sites are cut into chunks (a few per worker); workers format chunks into private buffers, the calling thread writes
// them to fd in site order as they complete, so formatting overlaps the write() copies.  format_range(a, b, out) returns
// false when it ran out of memory.
template <class F>
int format_parallel(int fd, int64_t n_sites, int32_t n_threads, F&& format_range) {
  const int nw = n_workers(n_threads, n_sites);
  const int64_t n_chunks = std::min<int64_t>(n_sites, static_cast<int64_t>(nw) * 4);
  std::vector<OutBuf> out(static_cast<size_t>(n_chunks));
  std::vector<int> state(static_cast<size_t>(n_chunks), 0);      // 0 pending, 1 done, -1 failed (guarded by mu)
  std::mutex mu;
  std::condition_variable cv;
  std::atomic<int64_t> next{0};
  std::atomic<bool> stop{false};
  std::vector<std::thread> th;
  for (int t = 0; t < nw; ++t)
    th.emplace_back([&] {
      for (;;) {
        const int64_t c = next.fetch_add(1);
        if (c >= n_chunks || stop.load()) return;
        const int64_t a = n_sites * c / n_chunks, b = n_sites * (c + 1) / n_chunks;
        const bool ok = format_range(a, b, out[static_cast<size_t>(c)]);
        {
          std::lock_guard<std::mutex> g(mu);
          state[static_cast<size_t>(c)] = ok ? 1 : -1;
        }
        cv.notify_all();
      }
    });
  int rc = M6A_OK;
  for (int64_t c = 0; c < n_chunks && rc == M6A_OK; ++c) {
    {
      std::unique_lock<std::mutex> g(mu);
      cv.wait(g, [&] { return state[static_cast<size_t>(c)] != 0; });
      if (state[static_cast<size_t>(c)] < 0) rc = M6A_ENOMEM;
    }
    OutBuf& o = out[static_cast<size_t>(c)];
    if (rc == M6A_OK && !write_all(fd, o.p, o.len)) rc = M6A_EIO;
    free(o.p);                                                     // release as soon as written
    o.p = nullptr;
    o.len = o.cap = 0;
  }
  stop.store(true);
  for (auto& x : th) x.join();
  return rc;
}

}  // namespace

extern "C" int m6a_ingest_parts_keyed(const char* const* paths, int32_t n_files, const m6a_part_t* parts, int64_t n_parts,
                                      int32_t n_flank, const double* norm_mean, const double* norm_std, const int32_t* kmer_id,
                                      const char* tx_buf, const int64_t* tx_off, const int64_t* tx_pos, float* feats,
                                      int64_t* read_ids, int32_t* kmer_idx, int32_t n_threads, int64_t* bad_part) {
  if (bad_part) *bad_part = -1;
  if (n_parts < 0 || n_files < 0 || n_flank < 0 || n_flank > 5) return M6A_EINVAL;
  if (n_parts == 0) return M6A_OK;
  if (!paths || !parts || !norm_mean || !norm_std || !kmer_id || !feats || !read_ids || !kmer_idx) return M6A_EINVAL;
  if ((tx_buf != nullptr) != (tx_off != nullptr) || (tx_buf != nullptr) != (tx_pos != nullptr)) return M6A_EINVAL;
  IngestCtx c;
  c.paths = paths;
  c.parts = parts;
  c.n_parts = n_parts;
  c.n_flank = n_flank;
  c.norm_mean = norm_mean;
  c.norm_std = norm_std;
  c.kmer_id = kmer_id;
  c.feats = feats;
  c.read_ids = read_ids;
  c.kmer_idx = kmer_idx;
  c.tx_buf = tx_buf;
  c.tx_off = tx_off;
  c.tx_pos = tx_pos;
  bool multi = false;
  for (int64_t i = 0; i < n_parts && !multi; ++i) multi = parts[i].first_of_site == 0;
  if (multi) c.part_ids.assign(static_cast<size_t>(n_parts) * 11, -1);
  for (int f = 0; f < n_files; ++f) {
    const int fd = open(paths[f], O_RDONLY);
    if (fd < 0) {
      for (int g : c.fds) close(g);
      return M6A_EIO;
    }
    c.fds.push_back(fd);
  }
  for (int64_t i = 0; i < n_parts; ++i)
    if (parts[i].file >= n_files) {
      for (int g : c.fds) close(g);
      return M6A_EINVAL;
    }
  const int nw = n_workers(n_threads, n_parts);
  std::vector<std::thread> th;
  for (int t = 0; t < nw; ++t) th.emplace_back(ingest_worker, &c);
  for (auto& x : th) x.join();
  for (int g : c.fds) close(g);
  // every replicate of a site must describe the same 7-mer (the reference asserts it, utils/data_utils.py:418)
  if (multi && c.status.load() == M6A_OK) {
    const int n_pos = 2 * n_flank + 1;
    for (int64_t i = 0; i < n_parts; ++i) {
      if (parts[i].first_of_site) continue;
      const int32_t* kid = kmer_idx + parts[i].site * n_pos;
      for (int j = 0; j < n_pos; ++j)
        if (c.part_ids[static_cast<size_t>(i) * 11 + j] != kid[j]) {
          fail(c, i, M6A_EPARSE);
          break;
        }
      if (c.status.load() != M6A_OK) break;
    }
  }
  if (bad_part) *bad_part = c.bad.load();
  return c.status.load();
}

extern "C" int m6a_ingest_parts(const char* const* paths, int32_t n_files, const m6a_part_t* parts, int64_t n_parts,
                                int32_t n_flank, const double* norm_mean, const double* norm_std, const int32_t* kmer_id,
                                float* feats, int64_t* read_ids, int32_t* kmer_idx, int32_t n_threads, int64_t* bad_part) {
  return m6a_ingest_parts_keyed(paths, n_files, parts, n_parts, n_flank, norm_mean, norm_std, kmer_id, nullptr, nullptr, nullptr,
                                feats, read_ids, kmer_idx, n_threads, bad_part);
}

// ---- data.info reader: csv with a header naming transcript_id,transcript_position,start,end,n_reads (any order,
// extra columns ignored) -- reference utils/data_utils.py:118-129 (pd.read_csv) ---------------------------------------
namespace {
struct InfoFile {
  std::vector<char> text;
  int col[5] = {-1, -1, -1, -1, -1};   // column index of transcript_id, transcript_position, start, end, n_reads
  size_t body = 0;                     // offset of the first data row
};

int load_info(const char* path, InfoFile& f) {
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return M6A_EIO;
  const off_t size = lseek(fd, 0, SEEK_END);
  if (size < 0) { close(fd); return M6A_EIO; }
  f.text.resize(static_cast<size_t>(size));
  size_t got = 0;
  while (got < f.text.size()) {
    const ssize_t r = pread(fd, f.text.data() + got, f.text.size() - got, static_cast<off_t>(got));
    if (r <= 0) { close(fd); return M6A_EIO; }
    got += static_cast<size_t>(r);
  }
  close(fd);
  if (f.text.empty()) return M6A_EPARSE;      // not even a header
  // header
  const char* p = f.text.data();
  const char* e = p + f.text.size();
  const char* eol = static_cast<const char*>(memchr(p, '\n', e - p));
  if (!eol) eol = e;
  static const char* names[5] = {"transcript_id", "transcript_position", "start", "end", "n_reads"};
  int c = 0;
  const char* q = p;
  while (q <= eol) {
    const char* comma = static_cast<const char*>(memchr(q, ',', eol - q));
    const char* stop = comma ? comma : eol;
    size_t len = static_cast<size_t>(stop - q);
    if (len && q[len - 1] == '\r') --len;
    for (int k = 0; k < 5; ++k)
      if (strlen(names[k]) == len && !strncmp(q, names[k], len)) f.col[k] = c;
    ++c;
    if (!comma) break;
    q = comma + 1;
  }
  for (int k = 0; k < 5; ++k)
    if (f.col[k] < 0) return M6A_EPARSE;
  f.body = static_cast<size_t>(eol - p) + (eol < e ? 1 : 0);
  return M6A_OK;
}
}  // namespace

// Counts the data rows and the total bytes of the transcript ids (to size the buffers of m6a_info_read).
extern "C" int m6a_info_count(const char* path, int64_t* n_rows, int64_t* tx_bytes) {
  if (!path || !n_rows || !tx_bytes) return M6A_EINVAL;
  InfoFile f;
  const int rc = load_info(path, f);
  if (rc != M6A_OK) return rc;
  int64_t rows = 0, bytes = 0;
  const char* p = f.text.data() + f.body;
  const char* e = f.text.data() + f.text.size();
  while (p < e) {
    const char* eol = static_cast<const char*>(memchr(p, '\n', e - p));
    if (!eol) eol = e;
    if (eol > p && !(eol - p == 1 && *p == '\r')) {
      // length of field col[0]
      const char* q = p;
      for (int c = 0; c < f.col[0]; ++c) {
        q = static_cast<const char*>(memchr(q, ',', eol - q));
        if (!q) return M6A_EPARSE;
        ++q;
      }
      const char* stop = static_cast<const char*>(memchr(q, ',', eol - q));
      if (!stop) stop = eol;
      bytes += stop - q;
      ++rows;
    }
    p = eol + 1;
  }
  *n_rows = rows;
  *tx_bytes = bytes;
  return M6A_OK;
}

// Fills tx_buf/tx_off (concatenated transcript ids, CSR offsets [n_rows+1]) and the four integer columns.
extern "C" int m6a_info_read(const char* path, int64_t n_rows, int64_t tx_bytes, char* tx_buf, int64_t* tx_off,
                             int64_t* tx_pos, int64_t* start, int64_t* end, int64_t* n_reads) {
  if (!path || n_rows < 0 || !tx_off || (n_rows > 0 && (!tx_buf || !tx_pos || !start || !end || !n_reads))) return M6A_EINVAL;
  InfoFile f;
  const int rc = load_info(path, f);
  if (rc != M6A_OK) return rc;
  int64_t* dst[5] = {nullptr, tx_pos, start, end, n_reads};
  int64_t row = 0, used = 0;
  tx_off[0] = 0;
  const char* p = f.text.data() + f.body;
  const char* e = f.text.data() + f.text.size();
  while (p < e) {
    const char* eol = static_cast<const char*>(memchr(p, '\n', e - p));
    if (!eol) eol = e;
    if (eol > p && !(eol - p == 1 && *p == '\r')) {
      if (row >= n_rows) return M6A_EPARSE;
      const char* q = p;
      int c = 0, seen = 0;
      while (q <= eol) {
        const char* comma = static_cast<const char*>(memchr(q, ',', eol - q));
        const char* stop = comma ? comma : eol;
        for (int k = 0; k < 5; ++k) {
          if (f.col[k] != c) continue;
          ++seen;
          if (k == 0) {
            const int64_t len = stop - q;
            if (used + len > tx_bytes) return M6A_EPARSE;
            memcpy(tx_buf + used, q, static_cast<size_t>(len));
            used += len;
            tx_off[row + 1] = used;
          } else {
            int64_t v = 0;
            const char* stop2 = (stop > q && stop[-1] == '\r') ? stop - 1 : stop;
            auto res = std::from_chars(q, stop2, v);
            if (res.ec != std::errc()) {         // pandas may have written a float ("12.0"): accept integral doubles
              double dv;
              auto r2 = std::from_chars(q, stop2, dv);
              if (r2.ec != std::errc()) return M6A_EPARSE;
              v = static_cast<int64_t>(dv);
            }
            dst[k][row] = v;
          }
        }
        ++c;
        if (!comma) break;
        q = comma + 1;
      }
      if (seen != 5) return M6A_EPARSE;
      ++row;
    }
    p = eol + 1;
  }
  return row == n_rows ? M6A_OK : M6A_EPARSE;
}

// ---- exact fast formatting --------------------------------------------------------------------------------------
// "%.16f" of a double in [0, 1] -- every probability and ratio of the two CSV files -- exactly as printf / Python's '%'
// print it (round-half-even on the exact binary value), without the multi-precision machinery of printf:
// p = m * 2^-k with m < 2^53, so p * 10^16 = m * 10^16 / 2^k needs 107 bits: one 128-bit product, one shift, one
// comparison of the remainder with the half.  Returns the number of characters written (18), or 0 if p is outside
// [0, 1] or not finite (the caller then falls back to snprintf).
static const char kDigitPairs[201] =
    "00010203040506070809101112131415161718192021222324252627282930313233343536373839404142434445464748495051525354555657585960"
    "616263646566676869707172737475767778798081828384858687888990919293949596979899";

inline int format_prob16(double p, char* out) {
  uint64_t bits;
  memcpy(&bits, &p, sizeof bits);
  if (bits > 0x3FF0000000000000ull) return 0;        // negative (sign bit), > 1, inf or NaN
  const int be = static_cast<int>(bits >> 52);
  uint64_t m = bits & ((1ull << 52) - 1);
  int k;                                               // p = m * 2^-k, k >= 52
  if (be == 0) {
    k = 1074;
  } else {
    m |= 1ull << 52;
    k = 1075 - be;
  }
  uint64_t q = 0;                                      // round(p * 10^16), <= 10^16
  if (k < 128) {
    const unsigned __int128 num = static_cast<unsigned __int128>(m) * 10000000000000000ull;   // < 2^107
    q = static_cast<uint64_t>(num >> k);
    const unsigned __int128 rem = num & ((static_cast<unsigned __int128>(1) << k) - 1);
    const unsigned __int128 half = static_cast<unsigned __int128>(1) << (k - 1);
    if (rem > half || (rem == half && (q & 1u))) ++q;
  }                                                    // k >= 128: num < 2^107 < half => 0
  const uint64_t ip = q / 10000000000000000ull;        // 0 or 1
  uint64_t frac = q - ip * 10000000000000000ull;
  out[0] = static_cast<char>('0' + ip);
  out[1] = '.';
  for (int i = 7; i >= 0; --i) {                       // 16 digits, two at a time from the right
    const uint64_t d = frac % 100;
    frac /= 100;
    out[2 + 2 * i] = kDigitPairs[2 * d];
    out[3 + 2 * i] = kDigitPairs[2 * d + 1];
  }
  return 18;
}

inline int format_prob16_or_printf(double p, char* out) {     // out holds >= 400 bytes ("%.16f" of DBL_MAX is 326 long)
  const int n = format_prob16(p, out);
  return n ? n : snprintf(out, 400, "%.16f", p);
}

inline int format_i64(long long v, char* out) {              // "%lld"
  char tmp[24];
  unsigned long long u = v < 0 ? 0ull - static_cast<unsigned long long>(v) : static_cast<unsigned long long>(v);
  int n = 0;
  do {
    tmp[n++] = static_cast<char>('0' + u % 10);
    u /= 10;
  } while (u);
  int w = 0;
  if (v < 0) out[w++] = '-';
  while (n) out[w++] = tmp[--n];
  return w;
}

// '%s,%d,%s,%.16f,%s,%.16f\n' % (tx_id, tx_pos, n_reads, site_prob, kmer, mod_ratio)   utils/inference_utils.py:59-60
extern "C" int m6a_write_site_csv(int32_t fd, int64_t n_sites, const char* tx_buf, const int64_t* tx_off,
                                  const int64_t* tx_pos, const int64_t* read_off, const float* site_prob,
                                  const int32_t* mod_count, const char* kmer5, int32_t n_threads) {
  if (n_sites < 0) return M6A_EINVAL;
  if (n_sites == 0) return M6A_OK;
  if (!tx_buf || !tx_off || !tx_pos || !read_off || !site_prob || !mod_count || !kmer5) return M6A_EINVAL;
  return format_parallel(fd, n_sites, n_threads, [&](int64_t a, int64_t b, OutBuf& out) {
    char line[1024];
    for (int64_t s = a; s < b; ++s) {
      const int64_t n = read_off[s + 1] - read_off[s];
      const double ratio = static_cast<double>(mod_count[s]) / static_cast<double>(n > 0 ? n : 1);
      const size_t tl = static_cast<size_t>(tx_off[s + 1] - tx_off[s]);
      if (!out.ensure(tl + sizeof line)) return false;
      out.append(tx_buf + tx_off[s], tl);
      char* w = line;
      *w++ = ',';
      w += format_i64(tx_pos[s], w);
      *w++ = ',';
      w += format_i64(n, w);
      *w++ = ',';
      w += format_prob16_or_printf(static_cast<double>(site_prob[s]), w);     // "nan" for a site without reads
      *w++ = ',';
      const size_t kl = strnlen(kmer5 + 5 * s, 5);                             // "%.5s"
      memcpy(w, kmer5 + 5 * s, kl);
      w += kl;
      *w++ = ',';
      w += format_prob16_or_printf(ratio, w);
      *w++ = '\n';
      out.append(line, static_cast<size_t>(w - line));
    }
    return true;
  });
}

// '%s,%d,%s,%.16f\n' % (tx_id, tx_pos, read_id, read_prob)   utils/inference_utils.py:63-64
// read_rep == NULL: read_index is the integer id; else "{id}_{rep}" (utils/data_utils.py:421-423)
extern "C" int m6a_write_indiv_csv(int32_t fd, int64_t n_sites, const char* tx_buf, const int64_t* tx_off,
                                   const int64_t* tx_pos, const int64_t* read_off, const int64_t* read_ids,
                                   const int32_t* read_rep, const float* read_prob, int32_t n_threads) {
  if (n_sites < 0) return M6A_EINVAL;
  if (n_sites == 0) return M6A_OK;
  if (!tx_buf || !tx_off || !tx_pos || !read_off || !read_ids || !read_prob) return M6A_EINVAL;
  return format_parallel(fd, n_sites, n_threads, [&](int64_t a, int64_t b, OutBuf& out) {
    // rows are formatted in place behind the site prefix "tx,pos,": a row adds at most 20 + 1 + 11 + 1 + 400 + 1 bytes
    std::string prefix;
    for (int64_t s = a; s < b; ++s) {
      char pos[32];
      int pl = 0;
      pos[pl++] = ',';
      pl += format_i64(tx_pos[s], pos + pl);
      pos[pl++] = ',';
      prefix.assign(tx_buf + tx_off[s], static_cast<size_t>(tx_off[s + 1] - tx_off[s]));
      prefix.append(pos, static_cast<size_t>(pl));
      const size_t row_max = prefix.size() + 440;
      for (int64_t r = read_off[s]; r < read_off[s + 1]; ++r) {
        if (out.cap - out.len < row_max && !out.ensure(row_max)) return false;
        char* w = out.p + out.len;
        memcpy(w, prefix.data(), prefix.size());
        w += prefix.size();
        w += format_i64(read_ids[r], w);
        if (read_rep) {
          *w++ = '_';
          w += format_i64(read_rep[r], w);
        }
        *w++ = ',';
        w += format_prob16_or_printf(static_cast<double>(read_prob[r]), w);     // printf spelling only for NaN / p outside [0, 1]
        *w++ = '\n';
        out.len = static_cast<size_t>(w - out.p);
      }
    }
    return true;
  });
}
