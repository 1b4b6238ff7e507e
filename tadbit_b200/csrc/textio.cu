// Host-side text output of the extraction side of N3: the cell lines of the .abc files that
// write_matrix of the reference's BAM parser produces
// (_pytadbit/parsers/hic_bam_parser.py:1101-1169: '{}\t{}\n'.format('{}\t{}\t'.format(a, b), value)).
// The values arrive as arrays from tb_region_cells; formatting millions of lines in Python costs far
// more than extracting the cells, so the lines are formatted here.  A float is written as Python's
// repr writes it (Objects/floatobject.c float_repr -> PyOS_double_to_string(x, 'r', 0, ADD_DOT_0)):
// the shortest digit string that reads back as the same double, in fixed notation when the decimal
// point falls at -4 < decpt <= 16, else as d.ddde+XX with at least two exponent digits; "nan", "inf",
// "-inf".  std::to_chars gives the same shortest digits (tests/test_textio.py compares the two on
// random doubles of every magnitude).  No GPU is involved.
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace tb {

namespace {

inline char *put_int(char *p, long long v) {
  auto r = std::to_chars(p, p + 24, v);
  return r.ptr;
}

// repr(float) of Python 3
inline char *put_repr(char *p, double x) {
  if (std::isnan(x)) { memcpy(p, "nan", 3); return p + 3; }
  if (std::isinf(x)) {
    if (x < 0) *p++ = '-';
    memcpy(p, "inf", 3);
    return p + 3;
  }
  char tmp[40];
  auto r = std::to_chars(tmp, tmp + sizeof(tmp), x, std::chars_format::scientific);
  const char *s = tmp, *end = r.ptr;
  if (*s == '-') *p++ = *s++;
  const char *e = s;
  while (e < end && *e != 'e') ++e;
  int exp10 = 0;
  {
    const char *q = e + 1;
    const bool neg = *q == '-';
    if (*q == '-' || *q == '+') ++q;
    for (; q < end; ++q) exp10 = exp10 * 10 + (*q - '0');
    if (neg) exp10 = -exp10;
  }
  const int decpt = exp10 + 1;
  if (decpt <= -4 || decpt > 16) {               // exponent notation: to_chars already writes d.ddde+XX
    memcpy(p, s, end - s);
    return p + (end - s);
  }
  char digits[24];
  int nd = 0;
  for (const char *q = s; q < e; ++q)
    if (*q != '.') digits[nd++] = *q;
  if (decpt <= 0) {
    *p++ = '0'; *p++ = '.';
    for (int k = 0; k < -decpt; ++k) *p++ = '0';
    memcpy(p, digits, nd);
    return p + nd;
  }
  if (decpt >= nd) {
    memcpy(p, digits, nd);
    p += nd;
    for (int k = nd; k < decpt; ++k) *p++ = '0';
    *p++ = '.'; *p++ = '0';
    return p;
  }
  memcpy(p, digits, decpt);
  p += decpt;
  *p++ = '.';
  memcpy(p, digits + decpt, nd - decpt);
  return p + (nd - decpt);
}

}  // namespace

void write_abc_lines(const char *path, int64_t n, const int32_t *a, const int32_t *b, const int32_t *ival,
                     const double *fval) {
  TB_REQUIRE(path && n >= 0 && (n == 0 || (a && b && (ival || fval))), "write_abc_lines: bad arguments");
  FILE *f = fopen(path, "ab");
  TB_REQUIRE(f, "write_abc_lines: cannot open %s", path);
  constexpr int64_t kBatch = 1 << 15;
  std::vector<char> buf((size_t)kBatch * 96);
  bool ok = true;
  for (int64_t k0 = 0; k0 < n && ok; k0 += kBatch) {
    const int64_t k1 = k0 + kBatch < n ? k0 + kBatch : n;
    char *p = buf.data();
    for (int64_t k = k0; k < k1; ++k) {
      p = put_int(p, a[k]);
      *p++ = '\t';
      p = put_int(p, b[k]);
      *p++ = '\t'; *p++ = '\t';
      p = ival ? put_int(p, ival[k]) : put_repr(p, fval[k]);
      *p++ = '\n';
    }
    ok = fwrite(buf.data(), 1, (size_t)(p - buf.data()), f) == (size_t)(p - buf.data());
  }
  ok = (fclose(f) == 0) && ok;
  TB_REQUIRE(ok, "write_abc_lines: write to %s failed", path);
}

// Lines "a<sep>b<sep>value<tail>" of Experiment.write_interaction_pairs
// (_pytadbit/experiment.py:1171-1186: '%s\t%s\t%s\n' for tsv, '%s,%s,%s\n' / '%s,%s,%s,\n' for
// json), appended to `path`; value = ival[k] or fval[k] as Python's '%s' prints a float (= repr).
void write_pair_lines(const char *path, int64_t n, const int64_t *a, const int64_t *b, const int64_t *ival,
                      const double *fval, const char *sep, const char *tail) {
  TB_REQUIRE(path && sep && tail && n >= 0 && (n == 0 || (a && b && (ival || fval))),
             "write_pair_lines: bad arguments");
  const size_t ls = strlen(sep), lt = strlen(tail);
  TB_REQUIRE(ls <= 8 && lt <= 8, "write_pair_lines: separator or line end longer than 8 bytes");
  FILE *f = fopen(path, "ab");
  TB_REQUIRE(f, "write_pair_lines: cannot open %s", path);
  constexpr int64_t kBatch = 1 << 15;
  std::vector<char> buf((size_t)kBatch * 128);
  bool ok = true;
  for (int64_t k0 = 0; k0 < n && ok; k0 += kBatch) {
    const int64_t k1 = k0 + kBatch < n ? k0 + kBatch : n;
    char *p = buf.data();
    for (int64_t k = k0; k < k1; ++k) {
      p = put_int(p, a[k]);
      memcpy(p, sep, ls); p += ls;
      p = put_int(p, b[k]);
      memcpy(p, sep, ls); p += ls;
      p = ival ? put_int(p, ival[k]) : put_repr(p, fval[k]);
      memcpy(p, tail, lt); p += lt;
    }
    ok = fwrite(buf.data(), 1, (size_t)(p - buf.data()), f) == (size_t)(p - buf.data());
  }
  ok = (fclose(f) == 0) && ok;
  TB_REQUIRE(ok, "write_pair_lines: write to %s failed", path);
}

}  // namespace tb
