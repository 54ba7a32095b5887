// modelio.cu — host-side writer of the reference's model text format (no device code; built into libmrscore.so with the kernels).
//
// MusicRecommender.scala:489-496 writeModelOnFile:  model foreach { case (user, (song, rank)) => out.print(s"$user\t$song\t$rank\n") }
// i.e. one line per emitted (user, song) pair, the score printed by java.lang.Double.toString.  importModelFromFile (MR:505-512)
// and distributed.scala's importModel (DIST:42-49) read it back with `.toDouble`, so the digits must round-trip; to diff a native file
// against one written by the JVM they must also be laid out by Java's rules:
//   * the shortest decimal digit string that uniquely distinguishes the double (what std::to_chars produces; JDK >= 19 prints exactly
//     these digits, older JDKs print a longer string for a few well-known values — SURVEY.md §8f N1),
//   * 1e-3 <= |x| < 1e7: plain decimal, at least one digit after the point ("0.4999999999999999", "123.0"),
//   * otherwise "computerised scientific notation" d.dddE[-]n ("1.0E7", "1.234E-5"), zero is "0.0".
// At configs[2] (2000 / 100 users) a model has 4.4 M lines; the JVM formats them one string interpolation at a time.
#include "../../include/mrscore.h"

#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

// Double.toString(x) into out (>= 32 bytes); returns the length.
int java_double_to_string(double x, char* out) {
  if (std::isnan(x)) { memcpy(out, "NaN", 3); return 3; }
  char* p = out;
  if (std::signbit(x)) { *p++ = '-'; x = -x; }
  if (std::isinf(x)) { memcpy(p, "Infinity", 8); return static_cast<int>(p - out) + 8; }
  if (x == 0.0) { memcpy(p, "0.0", 3); return static_cast<int>(p - out) + 3; }
  char sci[40];
  const auto res = std::to_chars(sci, sci + sizeof sci, x, std::chars_format::scientific);   // d[.ddd]e[+-]XX, shortest round trip
  char digits[24]; int nd = 0; const char* q = sci;
  for (; q < res.ptr && *q != 'e'; ++q) if (*q != '.') digits[nd++] = *q;
  int e10 = 0;
  { ++q; const bool neg = *q == '-'; if (*q == '-' || *q == '+') ++q; for (; q < res.ptr; ++q) e10 = e10 * 10 + (*q - '0'); if (neg) e10 = -e10; }
  while (nd > 1 && digits[nd - 1] == '0') --nd;
  if (e10 >= -3 && e10 < 7) {
    if (e10 >= 0) {
      for (int i = 0; i <= e10; ++i) *p++ = i < nd ? digits[i] : '0';
      *p++ = '.';
      if (nd > e10 + 1) { for (int i = e10 + 1; i < nd; ++i) *p++ = digits[i]; } else *p++ = '0';
    } else {
      *p++ = '0'; *p++ = '.';
      for (int i = 0; i < -e10 - 1; ++i) *p++ = '0';
      for (int i = 0; i < nd; ++i) *p++ = digits[i];
    }
  } else {
    *p++ = digits[0]; *p++ = '.';
    if (nd > 1) { for (int i = 1; i < nd; ++i) *p++ = digits[i]; } else *p++ = '0';
    *p++ = 'E';
    p = std::to_chars(p, p + 8, e10).ptr;
  }
  return static_cast<int>(p - out);
}

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

int mr_format_double(double x, char* out32) {
  if (!out32) return -1;
  const int n = java_double_to_string(x, out32);
  out32[n] = 0;
  return n;
}

int mr_write_model(const char* path, const double* scores_UxS, int n_users, int n_songs, const char* user_chars, const int64_t* user_off,
                   const char* song_chars, const int64_t* song_off, int append, int64_t* rows_written) {
  if (!path || !scores_UxS || n_users < 0 || n_songs < 0 || !user_chars || !user_off || !song_chars || !song_off) return MR_ERR_BAD_ARG;
  FILE* f = fopen(path, append ? "ab" : "wb");
  if (!f) return MR_ERR_BAD_ARG;
  std::vector<char> buf;
  buf.reserve(1 << 22);
  int64_t rows = 0;
  bool ok = true;
  for (int u = 0; u < n_users && ok; ++u) {
    const char* un = user_chars + user_off[u]; const size_t ul = static_cast<size_t>(user_off[u + 1] - user_off[u]);
    const double* row = scores_UxS + static_cast<int64_t>(u) * n_songs;
    for (int s = 0; s < n_songs; ++s) {
      const double x = row[s];
      if (x != x) continue;                                   // NaN = pair the reference does not emit (MR:109)
      const size_t sl = static_cast<size_t>(song_off[s + 1] - song_off[s]);
      const size_t at = buf.size();
      buf.resize(at + ul + sl + 36);
      char* p = buf.data() + at;
      memcpy(p, un, ul); p += ul; *p++ = '\t';
      memcpy(p, song_chars + song_off[s], sl); p += sl; *p++ = '\t';
      p += java_double_to_string(x, p); *p++ = '\n';
      buf.resize(static_cast<size_t>(p - buf.data()));
      ++rows;
    }
    if (buf.size() > (3u << 20) || u + 1 == n_users) {
      ok = fwrite(buf.data(), 1, buf.size(), f) == buf.size();
      buf.clear();
    }
  }
  ok = (fclose(f) == 0) && ok;
  if (rows_written) *rows_written = rows;
  return ok ? MR_OK : MR_ERR_BAD_ARG;
}

}  // extern "C"
#pragma GCC visibility pop
