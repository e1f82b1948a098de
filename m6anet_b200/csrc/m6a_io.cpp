// Host-side I/O of the inference path, multi-threaded, behind the C ABI (include/m6anet_b200.h):
//   m6a_ingest_parts    data.json byte ranges -> flat normalised float32 buffers
//                       (reference utils/data_utils.py:152-248 _load_data/load_data/__getitem__/get_norm_factor,
//                        :395-427 replicate concatenation; file format utils/dataprep_utils.py:473-485)
//   m6a_write_site_csv / m6a_write_indiv_csv   the reference row formats (utils/inference_utils.py:59-67)
// Numbers are parsed with std::from_chars (correctly rounded, == Python float()), normalised in double and
// rounded once to float32, so the buffers equal the reference's NanopolishDS output bit for bit.
#include <fcntl.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include <atomic>
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
// expects a string without escapes; returns the position after the closing quote, [*s0, *s1) = contents
inline const char* expect_string(const char* r, const char* e, const char** s0, const char** s1) {
  r = expect(r, e, '"');
  if (!r) return nullptr;
  *s0 = r;
  for (; r < e && *r != '"'; ++r)
    if (*r == '\\' || static_cast<unsigned char>(*r) < 0x20) return nullptr;   // escapes never occur in ids / k-mers
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
  const char *t0, *t1, *k0, *k1;
  const char* r = expect(buf.data(), e, '{');
  if (r) r = expect_string(r, e, &t0, &t1);
  if (r) r = expect(r, e, ':');
  if (r) r = expect(r, e, '{');
  if (r) r = expect_string(r, e, &t0, &t1);
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

template <class F>
int format_parallel(int fd, int64_t n_sites, int32_t n_threads, F&& format_range) {
  const int nw = n_workers(n_threads, n_sites);
  std::vector<std::string> out(nw);
  std::vector<std::thread> th;
  for (int t = 0; t < nw; ++t) {
    const int64_t a = n_sites * t / nw, b = n_sites * (t + 1) / nw;
    th.emplace_back([&, t, a, b] { format_range(a, b, out[t]); });
  }
  for (auto& x : th) x.join();
  for (int t = 0; t < nw; ++t)
    if (!write_all(fd, out[t].data(), out[t].size())) return M6A_EIO;
  return M6A_OK;
}

}  // namespace

extern "C" int m6a_ingest_parts(const char* const* paths, int32_t n_files, const m6a_part_t* parts, int64_t n_parts,
                                int32_t n_flank, const double* norm_mean, const double* norm_std, const int32_t* kmer_id,
                                float* feats, int64_t* read_ids, int32_t* kmer_idx, int32_t n_threads, int64_t* bad_part) {
  if (bad_part) *bad_part = -1;
  if (n_parts < 0 || n_files < 0 || n_flank < 0 || n_flank > 5) return M6A_EINVAL;
  if (n_parts == 0) return M6A_OK;
  if (!paths || !parts || !norm_mean || !norm_std || !kmer_id || !feats || !read_ids || !kmer_idx) return M6A_EINVAL;
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
  if (bad_part) *bad_part = c.bad.load();
  return c.status.load();
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

// '%s,%d,%s,%.16f,%s,%.16f\n' % (tx_id, tx_pos, n_read, site_prob, kmer, mod_ratio)   utils/inference_utils.py:59-60
extern "C" int m6a_write_site_csv(int32_t fd, int64_t n_sites, const char* tx_buf, const int64_t* tx_off,
                                  const int64_t* tx_pos, const int64_t* read_off, const float* site_prob,
                                  const int32_t* mod_count, const char* kmer5, int32_t n_threads) {
  if (n_sites < 0) return M6A_EINVAL;
  if (n_sites == 0) return M6A_OK;
  if (!tx_buf || !tx_off || !tx_pos || !read_off || !site_prob || !mod_count || !kmer5) return M6A_EINVAL;
  return format_parallel(fd, n_sites, n_threads, [&](int64_t a, int64_t b, std::string& out) {
    char line[160];
    out.reserve(static_cast<size_t>(b - a) * 96);
    for (int64_t s = a; s < b; ++s) {
      const int64_t n = read_off[s + 1] - read_off[s];
      const double ratio = static_cast<double>(mod_count[s]) / static_cast<double>(n > 0 ? n : 1);
      out.append(tx_buf + tx_off[s], static_cast<size_t>(tx_off[s + 1] - tx_off[s]));
      const int m = snprintf(line, sizeof line, ",%lld,%lld,%.16f,%.5s,%.16f\n", static_cast<long long>(tx_pos[s]),
                             static_cast<long long>(n), static_cast<double>(site_prob[s]), kmer5 + 5 * s, ratio);
      out.append(line, static_cast<size_t>(m));
    }
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
  return format_parallel(fd, n_sites, n_threads, [&](int64_t a, int64_t b, std::string& out) {
    char line[128];
    out.reserve(static_cast<size_t>(read_off[b] - read_off[a]) * 64);
    for (int64_t s = a; s < b; ++s) {
      const char* tx = tx_buf + tx_off[s];
      const size_t tl = static_cast<size_t>(tx_off[s + 1] - tx_off[s]);
      char pos[32];
      const int pl = snprintf(pos, sizeof pos, ",%lld,", static_cast<long long>(tx_pos[s]));
      for (int64_t r = read_off[s]; r < read_off[s + 1]; ++r) {
        out.append(tx, tl);
        out.append(pos, static_cast<size_t>(pl));
        int m;
        if (read_rep)
          m = snprintf(line, sizeof line, "%lld_%d,%.16f\n", static_cast<long long>(read_ids[r]), read_rep[r],
                       static_cast<double>(read_prob[r]));
        else
          m = snprintf(line, sizeof line, "%lld,%.16f\n", static_cast<long long>(read_ids[r]), static_cast<double>(read_prob[r]));
        out.append(line, static_cast<size_t>(m));
      }
    }
  });
}
