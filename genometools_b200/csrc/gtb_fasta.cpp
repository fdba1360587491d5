// gtb_fasta.cpp -- FASTA files -> the index files of a GtEncseq (.esq .ssp .des .sds .md5) for the
// DNA alphabet, on all host cores; the step in front of the sorter.
//
// The reference does this with two passes of one thread that fetch every character through a virtual
// call (gt_encseq_new_from_files, /root/reference/src/core/encseq.c:7503-7714: first pass
// gt_inputfiles2sequencekeyvalues :5421-5673 with encseq_charproc.gen, second pass
// files2encodedsequence :4529-4647 with the fill function of the chosen representation; the reader is
// gt_sequence_buffer_fasta_advance, src/core/sequence_buffer_fasta.c:41-165): 1.5-1.9 s for 64 Mbp,
// most of the wall time of the drop-in binary once the sort takes 50 ms.  Here the files are mapped,
// cut into chunks anywhere but inside a description (make_chunks), and every pass is a parallel loop
// over chunks or over blocks of the symbol array:
//
//   count    symbols + separators per chunk, descriptions located, characters checked
//   emit     one code per symbol into a byte array (0..3, 254 wildcard, 255 separator), separator
//            positions, wildcard runs, distribution of the original characters
//   pack     32 symbols per 64-bit word (GtTwobitencoding), special positions filled the way the
//            chosen representation fills them
//   md5      one sequence per task, read from the text, on threads of its own from the end of `count` on
//
// and the tables of the representation (wildcard ranges in pages, separator positions in pages) are
// built from the run lists.  The files are written field by field in the order of the reference's
// map specifications, every field padded to 8 bytes (gt_mapspec_write, src/core/mapspec.c:366-466).
//
// What is not covered returns GTB_FASTA_UNSUPPORTED with the reason, BEFORE anything is written, and
// the caller runs the reference's encoder instead (which then also words the error messages): other
// alphabets, non-regular files (.gz and .bz2 files are inflated with zlib / libbz2, by one thread), files that do not
// begin with '>', characters outside the
// alphabet, empty sequences, a description that ends with the file or holds a NUL, 2^32-2 symbols or
// more.  Byte identity with the reference's files is tested in tests/test_fasta_encseq.py.
#include <algorithm>
#include <atomic>
#include <cctype>
#include <cerrno>
#include <chrono>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/gtb200.h"

namespace {

constexpr uint8_t CODE_WILDCARD = 254, CODE_SEPARATOR = 255, CODE_UNDEF = 253;   // src/core/chardef.h:34-46
// GtEncseqAccessType, src/core/encseq_access_type.h:24-34
enum Sat : uint64_t { SAT_DIRECT = 0, SAT_BYTECOMPRESS, SAT_EQUALLENGTH, SAT_BITACCESS, SAT_UCHAR, SAT_USHORT,
                      SAT_UINT32, SAT_UNDEFINED };
constexpr uint64_t ENCSEQ_VERSION = 3;                                           // src/core/encseq.h:40

struct Unsupported { std::string why; };
struct IoError { std::string why; };

std::string format(const char *fmt, ...)
{
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  return buf;
}

double now()
{
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---------------------------------------------------------------- md5 (RFC 1321), one object per sequence
struct Md5 {
  uint32_t a = 0x67452301u, b = 0xefcdab89u, c = 0x98badcfeu, d = 0x10325476u;
  uint64_t total = 0;
  uint8_t tail[64];
  unsigned ntail = 0;

  static uint32_t rol(uint32_t x, int s) { return (x << s) | (x >> (32 - s)); }

  void block(const uint8_t *p)
  {
    uint32_t w[16];
    memcpy(w, p, 64);                                  // little-endian host
    uint32_t A = a, B = b, C = c, D = d;
#define GTB_MD5_F(x, y, z) ((z) ^ ((x) & ((y) ^ (z))))
#define GTB_MD5_G(x, y, z) ((y) ^ ((z) & ((x) ^ (y))))
#define GTB_MD5_H(x, y, z) ((x) ^ (y) ^ (z))
#define GTB_MD5_I(x, y, z) ((y) ^ ((x) | ~(z)))
#define GTB_MD5_STEP(f, a_, b_, c_, d_, k, s, t) a_ = b_ + rol(a_ + f(b_, c_, d_) + w[k] + t, s)
    GTB_MD5_STEP(GTB_MD5_F, A, B, C, D, 0, 7, 0xd76aa478u);  GTB_MD5_STEP(GTB_MD5_F, D, A, B, C, 1, 12, 0xe8c7b756u);
    GTB_MD5_STEP(GTB_MD5_F, C, D, A, B, 2, 17, 0x242070dbu); GTB_MD5_STEP(GTB_MD5_F, B, C, D, A, 3, 22, 0xc1bdceeeu);
    GTB_MD5_STEP(GTB_MD5_F, A, B, C, D, 4, 7, 0xf57c0fafu);  GTB_MD5_STEP(GTB_MD5_F, D, A, B, C, 5, 12, 0x4787c62au);
    GTB_MD5_STEP(GTB_MD5_F, C, D, A, B, 6, 17, 0xa8304613u); GTB_MD5_STEP(GTB_MD5_F, B, C, D, A, 7, 22, 0xfd469501u);
    GTB_MD5_STEP(GTB_MD5_F, A, B, C, D, 8, 7, 0x698098d8u);  GTB_MD5_STEP(GTB_MD5_F, D, A, B, C, 9, 12, 0x8b44f7afu);
    GTB_MD5_STEP(GTB_MD5_F, C, D, A, B, 10, 17, 0xffff5bb1u); GTB_MD5_STEP(GTB_MD5_F, B, C, D, A, 11, 22, 0x895cd7beu);
    GTB_MD5_STEP(GTB_MD5_F, A, B, C, D, 12, 7, 0x6b901122u); GTB_MD5_STEP(GTB_MD5_F, D, A, B, C, 13, 12, 0xfd987193u);
    GTB_MD5_STEP(GTB_MD5_F, C, D, A, B, 14, 17, 0xa679438eu); GTB_MD5_STEP(GTB_MD5_F, B, C, D, A, 15, 22, 0x49b40821u);
    GTB_MD5_STEP(GTB_MD5_G, A, B, C, D, 1, 5, 0xf61e2562u);  GTB_MD5_STEP(GTB_MD5_G, D, A, B, C, 6, 9, 0xc040b340u);
    GTB_MD5_STEP(GTB_MD5_G, C, D, A, B, 11, 14, 0x265e5a51u); GTB_MD5_STEP(GTB_MD5_G, B, C, D, A, 0, 20, 0xe9b6c7aau);
    GTB_MD5_STEP(GTB_MD5_G, A, B, C, D, 5, 5, 0xd62f105du);  GTB_MD5_STEP(GTB_MD5_G, D, A, B, C, 10, 9, 0x02441453u);
    GTB_MD5_STEP(GTB_MD5_G, C, D, A, B, 15, 14, 0xd8a1e681u); GTB_MD5_STEP(GTB_MD5_G, B, C, D, A, 4, 20, 0xe7d3fbc8u);
    GTB_MD5_STEP(GTB_MD5_G, A, B, C, D, 9, 5, 0x21e1cde6u);  GTB_MD5_STEP(GTB_MD5_G, D, A, B, C, 14, 9, 0xc33707d6u);
    GTB_MD5_STEP(GTB_MD5_G, C, D, A, B, 3, 14, 0xf4d50d87u); GTB_MD5_STEP(GTB_MD5_G, B, C, D, A, 8, 20, 0x455a14edu);
    GTB_MD5_STEP(GTB_MD5_G, A, B, C, D, 13, 5, 0xa9e3e905u); GTB_MD5_STEP(GTB_MD5_G, D, A, B, C, 2, 9, 0xfcefa3f8u);
    GTB_MD5_STEP(GTB_MD5_G, C, D, A, B, 7, 14, 0x676f02d9u); GTB_MD5_STEP(GTB_MD5_G, B, C, D, A, 12, 20, 0x8d2a4c8au);
    GTB_MD5_STEP(GTB_MD5_H, A, B, C, D, 5, 4, 0xfffa3942u);  GTB_MD5_STEP(GTB_MD5_H, D, A, B, C, 8, 11, 0x8771f681u);
    GTB_MD5_STEP(GTB_MD5_H, C, D, A, B, 11, 16, 0x6d9d6122u); GTB_MD5_STEP(GTB_MD5_H, B, C, D, A, 14, 23, 0xfde5380cu);
    GTB_MD5_STEP(GTB_MD5_H, A, B, C, D, 1, 4, 0xa4beea44u);  GTB_MD5_STEP(GTB_MD5_H, D, A, B, C, 4, 11, 0x4bdecfa9u);
    GTB_MD5_STEP(GTB_MD5_H, C, D, A, B, 7, 16, 0xf6bb4b60u); GTB_MD5_STEP(GTB_MD5_H, B, C, D, A, 10, 23, 0xbebfbc70u);
    GTB_MD5_STEP(GTB_MD5_H, A, B, C, D, 13, 4, 0x289b7ec6u); GTB_MD5_STEP(GTB_MD5_H, D, A, B, C, 0, 11, 0xeaa127fau);
    GTB_MD5_STEP(GTB_MD5_H, C, D, A, B, 3, 16, 0xd4ef3085u); GTB_MD5_STEP(GTB_MD5_H, B, C, D, A, 6, 23, 0x04881d05u);
    GTB_MD5_STEP(GTB_MD5_H, A, B, C, D, 9, 4, 0xd9d4d039u);  GTB_MD5_STEP(GTB_MD5_H, D, A, B, C, 12, 11, 0xe6db99e5u);
    GTB_MD5_STEP(GTB_MD5_H, C, D, A, B, 15, 16, 0x1fa27cf8u); GTB_MD5_STEP(GTB_MD5_H, B, C, D, A, 2, 23, 0xc4ac5665u);
    GTB_MD5_STEP(GTB_MD5_I, A, B, C, D, 0, 6, 0xf4292244u);  GTB_MD5_STEP(GTB_MD5_I, D, A, B, C, 7, 10, 0x432aff97u);
    GTB_MD5_STEP(GTB_MD5_I, C, D, A, B, 14, 15, 0xab9423a7u); GTB_MD5_STEP(GTB_MD5_I, B, C, D, A, 5, 21, 0xfc93a039u);
    GTB_MD5_STEP(GTB_MD5_I, A, B, C, D, 12, 6, 0x655b59c3u); GTB_MD5_STEP(GTB_MD5_I, D, A, B, C, 3, 10, 0x8f0ccc92u);
    GTB_MD5_STEP(GTB_MD5_I, C, D, A, B, 10, 15, 0xffeff47du); GTB_MD5_STEP(GTB_MD5_I, B, C, D, A, 1, 21, 0x85845dd1u);
    GTB_MD5_STEP(GTB_MD5_I, A, B, C, D, 8, 6, 0x6fa87e4fu);  GTB_MD5_STEP(GTB_MD5_I, D, A, B, C, 15, 10, 0xfe2ce6e0u);
    GTB_MD5_STEP(GTB_MD5_I, C, D, A, B, 6, 15, 0xa3014314u); GTB_MD5_STEP(GTB_MD5_I, B, C, D, A, 13, 21, 0x4e0811a1u);
    GTB_MD5_STEP(GTB_MD5_I, A, B, C, D, 4, 6, 0xf7537e82u);  GTB_MD5_STEP(GTB_MD5_I, D, A, B, C, 11, 10, 0xbd3af235u);
    GTB_MD5_STEP(GTB_MD5_I, C, D, A, B, 2, 15, 0x2ad7d2bbu); GTB_MD5_STEP(GTB_MD5_I, B, C, D, A, 9, 21, 0xeb86d391u);
#undef GTB_MD5_STEP
#undef GTB_MD5_F
#undef GTB_MD5_G
#undef GTB_MD5_H
#undef GTB_MD5_I
    a += A; b += B; c += C; d += D;
  }

  void update(const uint8_t *p, size_t n)
  {
    total += n;
    if (ntail) {
      const size_t take = std::min<size_t>(64 - ntail, n);
      memcpy(tail + ntail, p, take);
      ntail += (unsigned) take; p += take; n -= take;
      if (ntail < 64) return;
      block(tail);
      ntail = 0;
    }
    for (; n >= 64; p += 64, n -= 64) block(p);
    if (n) { memcpy(tail, p, n); ntail = (unsigned) n; }
  }

  void hex(char out[33])
  {
    const uint64_t bits = total * 8;
    const uint8_t one = 0x80, zero[64] = {0};
    update(&one, 1);
    update(zero, (ntail <= 56) ? 56 - ntail : 120 - ntail);
    uint8_t len[8];
    memcpy(len, &bits, 8);
    update(len, 8);
    const uint32_t v[4] = {a, b, c, d};
    const uint8_t *raw = reinterpret_cast<const uint8_t *>(v);
    static const char digits[] = "0123456789abcdef";
    for (int i = 0; i < 16; i++) { out[2 * i] = digits[raw[i] >> 4]; out[2 * i + 1] = digits[raw[i] & 15]; }
    out[32] = '\0';
  }
};

// ---------------------------------------------------------------- inputs
struct Input {
  std::string name;
  const uint8_t *p = nullptr;
  size_t len = 0;
  bool inflated = false;         // p is a malloc'ed buffer with the text of a .gz / .bz2 file
};

struct Mapped {
  std::vector<Input> files;
  ~Mapped()
  {
    for (auto &f : files) {
      if (f.inflated) free(const_cast<uint8_t *>(f.p));
      else if (f.p != nullptr && f.len > 0) munmap(const_cast<uint8_t *>(f.p), f.len);
    }
  }
};

// zlib and libbz2, found at run time (the library gains no dependency; without them such files are declined).
// Both through their stdio-like interface, as the reference reads them (gt_xgzread, src/core/xzlib.c:62-71;
// gt_xbzread, src/core/xbzlib.c:70-78).
struct Inflater {
  void *(*open)(const char *, const char *) = nullptr;
  int (*read)(void *, void *, unsigned) = nullptr;
  void (*close_void)(void *) = nullptr;
  int (*close_int)(void *) = nullptr;
  int (*buffer)(void *, unsigned) = nullptr;
  bool ok = false;
  Inflater(const char *lib, const char *fopen_, const char *fread_, const char *fclose_, bool close_returns_int,
           const char *fbuffer)
  {
    void *h = dlopen(lib, RTLD_NOW | RTLD_LOCAL);
    if (h == nullptr) return;
    open = reinterpret_cast<void *(*)(const char *, const char *)>(dlsym(h, fopen_));
    read = reinterpret_cast<int (*)(void *, void *, unsigned)>(dlsym(h, fread_));
    if (close_returns_int) close_int = reinterpret_cast<int (*)(void *)>(dlsym(h, fclose_));
    else close_void = reinterpret_cast<void (*)(void *)>(dlsym(h, fclose_));
    if (fbuffer != nullptr) buffer = reinterpret_cast<int (*)(void *, unsigned)>(dlsym(h, fbuffer));
    ok = open != nullptr && read != nullptr && (close_int != nullptr || close_void != nullptr);
  }
  void close(void *f) const { if (close_int) close_int(f); else close_void(f); }
};

// the text of a .gz / .bz2 file (gt_file_mode_determine, src/core/file.c: the suffix selects the library; gzread also
// passes a file through that is not compressed at all); one thread -- such a stream has no second entry point --
// but everything after it is parallel again
void inflate_input(Input &f, bool bzip2)
{
  static const Inflater gz("libz.so.1", "gzopen", "gzread", "gzclose", true, "gzbuffer");
  static const Inflater bz("libbz2.so.1.0", "BZ2_bzopen", "BZ2_bzread", "BZ2_bzclose", false, nullptr);
  const Inflater &z = bzip2 ? bz : gz;
  if (!z.ok) throw Unsupported{format("file \"%s\" is compressed and the library to read it is not there", f.name.c_str())};
  struct stat st;
  if (stat(f.name.c_str(), &st) != 0 || !S_ISREG(st.st_mode))
    throw Unsupported{format("cannot read file \"%s\"", f.name.c_str())};
  void *in = z.open(f.name.c_str(), "rb");
  if (in == nullptr) throw Unsupported{format("cannot open file \"%s\"", f.name.c_str())};
  if (z.buffer != nullptr) z.buffer(in, 1u << 20);
  size_t cap = (size_t) st.st_size * 4 + (size_t(1) << 16), len = 0;
  uint8_t *buf = static_cast<uint8_t *>(malloc(cap));
  for (;;) {
    if (buf == nullptr) { z.close(in); throw IoError{"out of memory (inflated text)"}; }
    const size_t room = std::min<size_t>(cap - len, size_t(1) << 30);
    const int got = z.read(in, buf + len, (unsigned) room);
    if (got < 0) { z.close(in); free(buf); throw Unsupported{format("file \"%s\" cannot be inflated", f.name.c_str())}; }
    if (got == 0) break;
    len += (size_t) got;
    if (len == cap) {
      cap *= 2;
      uint8_t *bigger = static_cast<uint8_t *>(realloc(buf, cap));
      if (bigger == nullptr) free(buf);
      buf = bigger;
    }
  }
  z.close(in);
  if (len == 0) { free(buf); throw Unsupported{format("file \"%s\" is empty", f.name.c_str())}; }
  f.p = buf;
  f.len = len;
  f.inflated = true;
}

bool has_suffix(const std::string &s, const char *suf)
{
  const size_t n = strlen(suf);
  return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

void map_inputs(const gtb_fasta_request *rq, Mapped &m)
{
  if (rq->numoffiles == 0) throw Unsupported{"no sequence files"};
  m.files.resize(rq->numoffiles);
  for (uint64_t i = 0; i < rq->numoffiles; i++) {
    Input &f = m.files[i];
    f.name = rq->filenames[i];
    // gt_file_mode_determine, src/core/file.c: the suffix selects the decompressor
    if (has_suffix(f.name, ".gz") || has_suffix(f.name, ".bz2")) {
      inflate_input(f, has_suffix(f.name, ".bz2"));
      if (f.p[0] != '>') throw Unsupported{format("file \"%s\" does not begin with '>'", f.name.c_str())};
      continue;
    }
    const int fd = open(f.name.c_str(), O_RDONLY);
    if (fd < 0) throw Unsupported{format("cannot open file \"%s\": %s", f.name.c_str(), strerror(errno))};
    struct stat st;
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) {
      close(fd);
      throw Unsupported{format("\"%s\" is not a regular file", f.name.c_str())};
    }
    f.len = (size_t) st.st_size;
    if (f.len == 0) { close(fd); throw Unsupported{format("file \"%s\" is empty", f.name.c_str())}; }
    const char *how = getenv("GTB200_FASTA_MAP");               // experiments: populate | seq | (plain)
    const bool populate = how != nullptr && strcmp(how, "populate") == 0;
    void *p = mmap(nullptr, f.len, PROT_READ, MAP_PRIVATE | (populate ? MAP_POPULATE : 0), fd, 0);
    close(fd);
    if (p == MAP_FAILED) { f.len = 0; throw IoError{format("cannot map file \"%s\": %s", f.name.c_str(), strerror(errno))}; }
    if (how != nullptr && strcmp(how, "seq") == 0) madvise(p, f.len, MADV_SEQUENTIAL);
    f.p = static_cast<const uint8_t *>(p);
    // gt_sequence_buffer_new_guess_type looks at the first file only (src/core/sequence_buffer.c:63-103);
    // a later file that begins with sequence characters would continue the last sequence of the file before
    if (f.p[0] != '>') throw Unsupported{format("file \"%s\" does not begin with '>'", f.name.c_str())};
  }
}

// ---------------------------------------------------------------- chunks
struct Header { size_t desc_begin, desc_end; uint64_t sep_pos; };   // description = [desc_begin, desc_end), '\n' at desc_end
struct Run { uint64_t start, len; };

struct Chunk {
  unsigned file = 0;
  size_t begin = 0, end = 0;
  uint64_t emitted = 0;          // symbols + one per header (a separator; the first header overall emits none)
  uint64_t nheaders = 0;
  uint64_t out = 0;              // position of the first symbol this chunk emits
  std::vector<Header> headers;
  std::vector<Run> wild;         // maximal within the chunk
  uint64_t orig[256];
  std::string bad;               // why the input is not covered
};

// classes as bit flags: a stretch of text is scanned without branches, the flags of its characters OR-ed
enum Cls : uint8_t { C_SYMBOL = 0, C_SPACE = 1, C_HEADER = 2, C_ILLEGAL = 4 };
constexpr uint16_t E_SYMBOL = 0x100, E_WILD = 0x200;            // Tables::emit: code | flags

struct Tables {
  uint8_t cls[256];
  uint8_t code[256];
  uint16_t emit[256];
};

Tables make_tables(const uint8_t *symbolmap)
{
  Tables t;
  for (int c = 0; c < 256; c++) {
    t.code[c] = symbolmap[c];
    if (isspace(c)) t.cls[c] = C_SPACE;                 // sequence_buffer_fasta.c:124
    else if (c == '>') t.cls[c] = C_HEADER;             // :126
    else if (symbolmap[c] == CODE_UNDEF || symbolmap[c] == CODE_SEPARATOR || c >= 128 || c == 0)
      t.cls[c] = C_ILLEGAL;                             // process_char, sequence_buffer_inline.h:34-43
    else t.cls[c] = C_SYMBOL;
    t.emit[c] = symbolmap[c];
    if (t.cls[c] == C_SYMBOL) t.emit[c] |= E_SYMBOL | (symbolmap[c] == CODE_WILDCARD ? E_WILD : 0);
  }
  return t;
}

// first pass over a chunk: how much it emits, where its descriptions are, which characters occur.
// Line by line: the text up to the next newline is scanned without branches; only a line that holds a
// '>' or a character outside the alphabet is looked at again, character by character.
void count_chunk(const Input &f, const Tables &t, Chunk &c)
{
  const uint8_t *p = f.p;
  size_t i = c.begin;
  uint64_t nsym = 0;
  uint64_t hist[4][256];
  memset(hist, 0, sizeof hist);
  while (i < c.end) {
    const void *nl = memchr(p + i, '\n', c.end - i);
    const size_t e = nl ? (size_t) (static_cast<const uint8_t *>(nl) - p) : c.end;
    unsigned flags = 0;
    uint64_t cnt = 0;
    size_t j = i;
    for (; j + 4 <= e; j += 4) {
      const uint8_t c0 = p[j], c1 = p[j + 1], c2 = p[j + 2], c3 = p[j + 3];
      const unsigned k0 = t.cls[c0], k1 = t.cls[c1], k2 = t.cls[c2], k3 = t.cls[c3];
      flags |= k0 | k1 | k2 | k3;
      cnt += (k0 == 0) + (k1 == 0) + (k2 == 0) + (k3 == 0);
      hist[0][c0]++; hist[1][c1]++; hist[2][c2]++; hist[3][c3]++;
    }
    for (; j < e; j++) {
      const unsigned k = t.cls[p[j]];
      flags |= k;
      cnt += (k == 0);
      hist[0][p[j]]++;
    }
    if (!(flags & (C_HEADER | C_ILLEGAL))) { nsym += cnt; i = e + 1; continue; }
    // the line again, up to its '>' (the description then ends with the line) or its illegal character
    for (j = i; j < e; j++) {
      const unsigned k = t.cls[p[j]];
      if (k == C_SYMBOL) nsym++;
      else if (k == C_HEADER) break;
      else if (k == C_ILLEGAL) {
        c.bad = format("illegal character '%c' in file \"%s\"", p[j], f.name.c_str());
        return;
      }
    }
    for (size_t q = j; q < e; q++) hist[0][p[q]]--;              // the description is not sequence text
    if (nl == nullptr) { c.bad = format("file \"%s\" ends inside a description", f.name.c_str()); return; }
    if (memchr(p + j + 1, '\0', e - (j + 1)) != nullptr) {
      c.bad = format("a description of file \"%s\" holds a NUL character", f.name.c_str());
      return;
    }
    c.headers.push_back(Header{j + 1, e, 0});
    i = e + 1;
  }
  for (int k = 0; k < 256; k++) c.orig[k] = hist[0][k] + hist[1][k] + hist[2][k] + hist[3][k];
  c.nheaders = c.headers.size();
  c.emitted = nsym + c.nheaders;
}

// second pass: the codes, the separator positions, the wildcard runs
void emit_chunk(const Input &f, const Tables &t, Chunk &c, uint8_t *codes, bool first_header_overall_here)
{
  const uint8_t *p = f.p;
  size_t i = c.begin, h = 0;
  uint64_t o = c.out, wstart = 0, wlen = 0;
  auto close_run = [&]() { if (wlen) { c.wild.push_back(Run{wstart, wlen}); wlen = 0; } };
  auto runs_of = [&](uint64_t from, uint64_t to) {              // wildcard runs among the codes just written
    for (uint64_t q = from; q < to; q++) {
      if (codes[q] == CODE_WILDCARD) { if (!wlen) wstart = q; wlen++; }
      else close_run();
    }
  };
  while (i < c.end) {
    const void *nl = memchr(p + i, '\n', c.end - i);
    const size_t e = nl ? (size_t) (static_cast<const uint8_t *>(nl) - p) : c.end;
    size_t stop = e;                                            // the text of this line ends here
    const bool header_next = h < c.headers.size() && c.headers[h].desc_end == e;
    if (header_next) stop = c.headers[h].desc_begin - 1;        // at its '>'
    const uint64_t o0 = o;
    unsigned any = 0;
    for (size_t j = i; j < stop; j++) {
      const uint16_t v = t.emit[p[j]];
      if (v & E_SYMBOL) codes[o++] = (uint8_t) v;               // (nothing is stored for the rest: the next
      any |= v;                                                 //  position may belong to the next chunk)
    }
    if (any & E_WILD) runs_of(o0, o);
    else if (o > o0) close_run();
    if (header_next) {
      Header &hd = c.headers[h];
      if (first_header_overall_here && h == 0) hd.sep_pos = UINT64_MAX;
      else { close_run(); hd.sep_pos = o; codes[o++] = CODE_SEPARATOR; }
      h++;
    }
    i = e + 1;
  }
  close_run();
}

// tasks 0..ntasks-1 dealt to nthreads threads; what a task throws (out of memory) is thrown again here,
// after every thread has been joined
template <class F> void parallel_for(unsigned nthreads, size_t ntasks, F &&fn)
{
  if (ntasks == 0) return;
  std::atomic<size_t> next{0};
  std::atomic<bool> failed{false};
  std::exception_ptr first;
  std::mutex first_mutex;
  auto worker = [&]() {
    try {
      for (size_t i; !failed.load(std::memory_order_relaxed) && (i = next.fetch_add(1)) < ntasks;) fn(i);
    } catch (...) {
      std::lock_guard<std::mutex> lock(first_mutex);
      if (!first) first = std::current_exception();
      failed.store(true);
    }
  };
  const unsigned nt = (unsigned) std::min<size_t>(nthreads, ntasks);
  std::vector<std::thread> th;
  for (unsigned k = 1; k < nt; k++) {
    try { th.emplace_back(worker); } catch (...) { break; }      // fewer threads, same work
  }
  worker();
  for (auto &x : th) x.join();
  if (first) std::rethrow_exception(first);
}

// The files are cut every `target` bytes, wherever that falls -- a FASTA file may hold its sequence on one line --
// except inside a description: a position is inside one iff a '>' stands between the last newline in front of it
// and itself.  Per piece, in parallel: where its last newline is and whether a '>' follows it (or stands anywhere
// in a piece without newline); then one walk over the pieces, and a cut that fell into a description moves behind
// the newline that ends it.  Every chunk therefore begins outside a description, and a description never leaves
// the chunk its '>' is in.
std::vector<Chunk> make_chunks(const std::vector<Input> &files, size_t target, unsigned nthreads)
{
  struct Piece { unsigned file; size_t begin, end; bool has_nl, gt_after; };
  std::vector<Piece> pieces;
  for (unsigned fi = 0; fi < files.size(); fi++)
    for (size_t b = 0; b < files[fi].len; b += target)
      pieces.push_back(Piece{fi, b, std::min(files[fi].len, b + target), false, false});
  parallel_for(nthreads, pieces.size(), [&](size_t i) {
    Piece &pc = pieces[i];
    const uint8_t *p = files[pc.file].p;
    const void *nl = memrchr(p + pc.begin, '\n', pc.end - pc.begin);
    const size_t from = nl ? (size_t) (static_cast<const uint8_t *>(nl) - p) + 1 : pc.begin;
    pc.has_nl = nl != nullptr;
    pc.gt_after = from < pc.end && memchr(p + from, '>', pc.end - from) != nullptr;
  });
  std::vector<Chunk> chunks;
  size_t k = 0;
  for (unsigned fi = 0; fi < files.size(); fi++) {
    const Input &f = files[fi];
    size_t last_cut = 0;
    bool in_desc = false;                                        // at the begin of the piece looked at
    for (; k < pieces.size() && pieces[k].file == fi; k++) {
      const Piece &pc = pieces[k];
      size_t cut = pc.begin;
      if (cut > 0 && in_desc) {
        const void *nl = memchr(f.p + cut, '\n', f.len - cut);
        cut = nl ? (size_t) (static_cast<const uint8_t *>(nl) - f.p) + 1 : f.len;
      }
      if (cut > last_cut && cut < f.len) {
        Chunk c;
        c.file = fi; c.begin = last_cut; c.end = cut;
        chunks.push_back(std::move(c));
        last_cut = cut;
      }
      in_desc = pc.has_nl ? pc.gt_after : (in_desc || pc.gt_after);
    }
    Chunk c;
    c.file = fi; c.begin = last_cut; c.end = f.len;
    chunks.push_back(std::move(c));
  }
  return chunks;
}

// the two big arrays (one byte per symbol, the packed words): anonymous mappings that ask for huge pages --
// with 4 KB pages the first touch of 64 MB is 16 000 page faults, taken by all threads at once on one
// address-space lock (and on the lock the thread that creates the CUDA context holds most of the time)
struct BigBuffer {
  void *p = nullptr;
  size_t bytes = 0;
  bool mapped = false;
  explicit BigBuffer(size_t want)
  {
    const size_t huge = size_t(2) << 20;
    bytes = (want + huge - 1) / huge * huge;
    void *q = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (q != MAP_FAILED) {
      p = q;
      mapped = true;
#ifdef MADV_HUGEPAGE
      madvise(p, bytes, MADV_HUGEPAGE);
#endif
    } else {
      p = calloc(want, 1);
      if (p == nullptr) throw IoError{"out of memory"};
    }
  }
  ~BigBuffer() { if (mapped) munmap(p, bytes); else free(p); }
  BigBuffer(const BigBuffer &) = delete;
  BigBuffer &operator=(const BigBuffer &) = delete;
};

// ---------------------------------------------------------------- the representation
// GtSpecialcharinfo, src/core/chardef.h:90-115
struct SpecialCharInfo {
  uint64_t specialcharacters, specialranges, realspecialranges, lengthofspecialprefix, lengthofspecialsuffix,
      wildcards, wildcardranges, realwildcardranges, lengthofwildcardprefix, lengthofwildcardsuffix,
      lengthoflongestnonspecial, exceptioncharacters, exceptionranges, realexceptionranges;
};
static_assert(sizeof(SpecialCharInfo) == 112, "14 GtUword");

const uint64_t MAXRANGE[3] = {0xffu, 0xffffu, 0xffffffffu};    // initSWtable, encseq.c:1738-1767
const unsigned WIDTH[3] = {1, 2, 4};

uint64_t units_of_twobit(uint64_t n)                            // gt_unitsoftwobitencoding, src/core/intbits.h:194-205
{
  return n < 32 ? 2 : 2 + (n - 1) / 32;
}

uint64_t ints_for_bits(uint64_t nbits)                          // GT_NUMOFINTSFORBITS, intbits.h:98-101
{
  return (nbits >> 6) == 0 ? 1 : 1 + ((nbits - 1) >> 6);
}

// SIZEOFSWTABLE, encseq.c:923-950
uint64_t size_of_swtable(int kind, bool withrangelength, uint64_t n, uint64_t items)
{
  if (items == 0) return 0;
  return (withrangelength ? 2 : 1) * (uint64_t) WIDTH[kind] * items + 8 * (n / MAXRANGE[kind] + 1);
}

// how many table entries the runs need when an entry covers at most maxrange+1 positions
// (currentspecialrangevalue, encseq.c:5061-5074)
void ranges_tab(const std::vector<Run> &runs, uint64_t tab[3])
{
  tab[0] = tab[1] = 0;
  tab[2] = runs.size();
  for (const Run &r : runs)
    for (int k = 0; k < 2; k++) {
      const uint64_t page = MAXRANGE[k] + 1;
      tab[k] += r.len <= page ? 1 : (r.len + page - 1) / page;
    }
}

struct Writer {
  FILE *fp = nullptr;
  std::string path;
  uint64_t off = 0;
  Writer(const std::string &indexname, const char *suffix) : path(indexname + suffix)
  {
    fp = fopen(path.c_str(), "wb");
    if (fp == nullptr) throw IoError{format("cannot open file \"%s\" for writing: %s", path.c_str(), strerror(errno))};
    setvbuf(fp, nullptr, _IOFBF, 1 << 20);
  }
  ~Writer() { if (fp) fclose(fp); }
  void raw(const void *p, size_t bytes)
  {
    if (bytes && fwrite(p, 1, bytes, fp) != bytes) throw IoError{format("cannot write to \"%s\": %s", path.c_str(), strerror(errno))};
    off += bytes;
  }
  // one entry of a map specification: nothing for zero units, else the units and the padding to 8 bytes
  void field(const void *p, size_t unit, uint64_t units)
  {
    static const uint8_t zeros[8] = {0};
    if (units == 0) return;
    raw(p, unit * units);
    if (off % 8) raw(zeros, 8 - off % 8);
  }
  void word(uint64_t v) { field(&v, 8, 1); }
  void finish()
  {
    FILE *f = fp;
    fp = nullptr;
    if (fclose(f) != 0) throw IoError{format("cannot close \"%s\": %s", path.c_str(), strerror(errno))};
  }
};

// positions / (lengths) / endidxinpage of a table with pages of maxrange+1 positions
// (fillSWtable, src/core/accspecialrange.gen:29-262; ssptaboutinfo_*, encseq.c:1841-1910):
// an entry holds its start modulo the page size, a run longer than a page's worth continues in a new
// entry, endidxinpage[p] = entries that start at or before the last position of page p
template <class T> struct SwTable {
  std::vector<T> positions, lengths;
  std::vector<uint64_t> endidx;
};

template <class T>
SwTable<T> build_swtable(const std::vector<Run> &runs, int kind, uint64_t n, uint64_t items, bool withlengths)
{
  SwTable<T> t;
  const uint64_t page = MAXRANGE[kind] + 1, npages = n / MAXRANGE[kind] + 1;
  t.positions.reserve(items);
  if (withlengths) t.lengths.reserve(items);
  t.endidx.assign(npages, 0);
  for (const Run &r : runs) {
    uint64_t s = r.start, left = r.len;
    while (left) {
      const uint64_t take = withlengths ? std::min(left, page) : 1;
      t.positions.push_back((T) (s & MAXRANGE[kind]));
      if (withlengths) t.lengths.push_back((T) (take - 1));
      t.endidx[s / page]++;
      s += take;
      left -= take;
    }
  }
  for (uint64_t p = 1; p < npages; p++) t.endidx[p] += t.endidx[p - 1];
  if (t.positions.size() != items) throw IoError{"internal: table entries do not add up"};
  return t;
}

template <class T> void write_swtable(Writer &w, const SwTable<T> &t, bool withlengths)
{
  // addswtabletomapspectable, encseq.c:829-897
  if (t.positions.empty()) return;
  w.field(t.positions.data(), sizeof(T), t.positions.size());
  if (withlengths) w.field(t.lengths.data(), sizeof(T), t.lengths.size());
  w.field(t.endidx.data(), 8, t.endidx.size());
}

void write_table_of_kind(Writer &w, const std::vector<Run> &runs, int kind, uint64_t n, uint64_t items, bool withlengths)
{
  if (items == 0) return;
  if (kind == 0) write_swtable(w, build_swtable<uint8_t>(runs, 0, n, items, withlengths), withlengths);
  else if (kind == 1) write_swtable(w, build_swtable<uint16_t>(runs, 1, n, items, withlengths), withlengths);
  else write_swtable(w, build_swtable<uint32_t>(runs, 2, n, items, withlengths), withlengths);
}

const char *sat_name(uint64_t sat)                              // src/core/encseq_access_type.c:28-37
{
  static const char *names[] = {"direct", "bytecompress", "eqlen", "bit", "uchar", "ushort", "uint32", "undefined"};
  return names[sat <= SAT_UNDEFINED ? sat : SAT_UNDEFINED];
}

int encode(const gtb_fasta_request *rq, gtb_fasta_summary *sum)
{
  const double t0 = now();
  // the two alphabets the reference knows by name (alphabet_to_key_values, encseq.c:1080-1111: alphatype 0 and 1;
  // any other is stored with its definition, which is not rebuilt here)
  const bool dna = rq->alphatype == 0 && rq->numofchars == 4;
  const bool protein = rq->alphatype == 1 && rq->numofchars > 4 && rq->numofchars <= 64 && rq->bits_per_symbol >= 1 &&
                       rq->bits_per_symbol <= 8 && (1u << rq->bits_per_symbol) >= rq->numofchars + 2;
  if (!dna && !protein) throw Unsupported{"neither the DNA nor the protein alphabet"};
  const unsigned K = rq->numofchars;
  unsigned nthreads = rq->threads > 0 ? (unsigned) rq->threads : std::thread::hardware_concurrency();
  if (nthreads == 0) nthreads = 1;
  if (rq->threads <= 0 && nthreads > 32) nthreads = 32;
  if (nthreads > 256) nthreads = 256;
  const Tables tables = make_tables(rq->symbolmap);
  const std::string indexname = rq->indexname;

  // the directory of the index must take new files: if not, the reference's encoder is the one to say so
  {
    const size_t slash = indexname.rfind('/');
    const std::string dir = slash == std::string::npos ? "." : (slash == 0 ? "/" : indexname.substr(0, slash));
    if (access(dir.c_str(), W_OK | X_OK) != 0)
      throw Unsupported{format("cannot create files in \"%s\": %s", dir.c_str(), strerror(errno))};
  }
  Mapped mapped;
  map_inputs(rq, mapped);
  const std::vector<Input> &files = mapped.files;
  size_t total_bytes = 0;
  for (const Input &f : files) total_bytes += f.len;
  size_t target = std::max<size_t>(size_t(1) << 16, std::min<size_t>(size_t(8) << 20, total_bytes / (8 * (size_t) nthreads) + 1));
  if (const char *e = getenv("GTB200_FASTA_CHUNK"))              // tests: chunk borders everywhere
    if (atol(e) > 0) target = (size_t) atol(e);
  std::vector<Chunk> chunks = make_chunks(files, target, nthreads);

  // ---- count
  parallel_for(nthreads, chunks.size(), [&](size_t i) { count_chunk(files[chunks[i].file], tables, chunks[i]); });
  for (const Chunk &c : chunks)
    if (!c.bad.empty()) throw Unsupported{c.bad};
  uint64_t emitted = 0, numofsequences = 0;
  for (Chunk &c : chunks) {
    c.out = emitted == 0 ? 0 : emitted - 1;     // the first header overall (file 0, offset 0) emits no separator
    emitted += c.emitted;
    numofsequences += c.nheaders;
  }
  const uint64_t n = emitted - 1;
  if (n == 0) throw Unsupported{"no symbols"};
  if (n + 1 >= 0xffffffffull) throw Unsupported{"2^32-2 symbols or more"};
  const double t_count = now();

  // ---- md5 of every sequence: upper case of the decoded symbols (encseq_charproc.gen).  One sequence per task,
  //      read from the text itself (the descriptions found above bound it), so that it can start now, on threads
  //      of its own beside everything that follows: a genome of one sequence is ONE task and takes as long as
  //      all other passes together
  struct Span { unsigned file; size_t begin, end; };
  std::vector<Span> spans;
  std::vector<char> md5tab;
  std::thread md5_thread;
  double md5_seconds = 0;
  bool md5_failed = false;
  struct Join { std::thread &t; ~Join() { if (t.joinable()) t.join(); } } join_md5{md5_thread};
  if (rq->out_md5) {
    spans.reserve(numofsequences);
    for (const Chunk &c : chunks)
      for (const Header &h : c.headers) {
        if (!spans.empty() && spans.back().file == c.file) spans.back().end = h.desc_begin - 1;   // at its '>'
        spans.push_back(Span{c.file, h.desc_end + 1, files[c.file].len});
      }
    md5tab.assign(33 * numofsequences, '\0');
    auto all_md5 = [&]() {
      try {
      const double t = now();
      uint8_t up[256];                                          // 0: not a symbol (white space)
      for (int c = 0; c < 256; c++)
        up[c] = tables.cls[c] == C_SYMBOL ? (uint8_t) toupper((unsigned char) rq->decode[tables.code[c]]) : 0;
      parallel_for(nthreads, spans.size(), [&](size_t s) {
        const uint8_t *p = files[spans[s].file].p;
        Md5 m;
        uint8_t buf[(1 << 14) + 8];
        for (size_t i = spans[s].begin; i < spans[s].end;) {
          const size_t stop = std::min(spans[s].end, i + (size_t(1) << 14));
          size_t k = 0;
          for (; i < stop; i++) {
            const uint8_t u = up[p[i]];
            buf[k] = u;
            k += (u != 0);
          }
          m.update(buf, k);
        }
        m.hex(&md5tab[33 * s]);
      });
      md5_seconds = now() - t;
      } catch (...) { md5_failed = true; }
    };
    try { md5_thread = std::thread(all_md5); } catch (...) { all_md5(); }
  }

  // ---- emit
  BigBuffer codes_buffer(n + 64);
  uint8_t *codes = static_cast<uint8_t *>(codes_buffer.p);
  parallel_for(nthreads, chunks.size(), [&](size_t i) { emit_chunk(files[chunks[i].file], tables, chunks[i], codes, i == 0); });
  const double t_emit = now();

  // ---- the lists: separator positions, descriptions, wildcard runs (joined over chunk borders), special runs
  std::vector<uint64_t> seppos;
  seppos.reserve(numofsequences);
  std::vector<Run> wild;
  uint64_t orig[256] = {0};
  std::vector<uint64_t> file_symbols(files.size(), 0), file_headers(files.size(), 0);
  for (const Chunk &c : chunks) {
    for (const Header &h : c.headers)
      if (h.sep_pos != UINT64_MAX) seppos.push_back(h.sep_pos);
    for (const Run &r : c.wild) {
      if (!wild.empty() && wild.back().start + wild.back().len == r.start) wild.back().len += r.len;
      else wild.push_back(r);
    }
    for (int k = 0; k < 256; k++) orig[k] += c.orig[k];
    file_symbols[c.file] += c.emitted - c.nheaders;
    file_headers[c.file] += c.nheaders;
  }
  // sequence lengths; an empty sequence is an error of the reference ("contains an empty sequence",
  // encseq_charproc.gen) or, as the last one, a case not worth a second implementation
  uint64_t minseqlen = UINT64_MAX, maxseqlen = 0, first_len = 0;
  bool all_equal = true;
  {
    uint64_t start = 0;
    for (uint64_t s = 0; s <= seppos.size(); s++) {
      const uint64_t end = s < seppos.size() ? seppos[s] : n;
      const uint64_t len = end - start;
      if (len == 0) throw Unsupported{"an empty sequence"};
      if (s == 0) first_len = len;
      else if (len != first_len) all_equal = false;
      minseqlen = std::min(minseqlen, len);
      maxseqlen = std::max(maxseqlen, len);
      start = end + 1;
    }
  }
  // special runs = wildcard runs and separators, neighbours joined
  std::vector<Run> special;
  special.reserve(wild.size() + seppos.size());
  {
    size_t a = 0, b = 0;
    while (a < wild.size() || b < seppos.size()) {
      Run r;
      if (b >= seppos.size() || (a < wild.size() && wild[a].start < seppos[b])) r = wild[a++];
      else r = Run{seppos[b++], 1};
      if (!special.empty() && special.back().start + special.back().len == r.start) special.back().len += r.len;
      else special.push_back(r);
    }
  }
  SpecialCharInfo sci;
  memset(&sci, 0, sizeof sci);
  for (const Run &r : wild) sci.wildcards += r.len;
  sci.specialcharacters = sci.wildcards + seppos.size();
  sci.realspecialranges = special.size();
  sci.realwildcardranges = wild.size();
  if (!special.empty() && special.front().start == 0) sci.lengthofspecialprefix = special.front().len;
  if (!special.empty() && special.back().start + special.back().len == n) sci.lengthofspecialsuffix = special.back().len;
  if (!wild.empty() && wild.front().start == 0) sci.lengthofwildcardprefix = wild.front().len;
  if (!wild.empty() && wild.back().start + wild.back().len == n) sci.lengthofwildcardsuffix = wild.back().len;
  {
    uint64_t prev_end = 0;
    for (const Run &r : special) {
      sci.lengthoflongestnonspecial = std::max(sci.lengthoflongestnonspecial, r.start - prev_end);
      prev_end = r.start + r.len;
    }
    sci.lengthoflongestnonspecial = std::max(sci.lengthoflongestnonspecial, n - prev_end);
  }
  // equallength: every sequence as long as the first and nothing special but the separators
  // (gt_inputfiles2sequencekeyvalues, encseq.c:5574-5587,5640-5655)
  const bool equallength = all_equal && sci.wildcards == 0;

  // character distribution; the distinct original characters per code (determine_original_subdist,
  // encseq.c:5270-5359: printable characters 1..127)
  uint64_t chardist[64] = {0};
  uint64_t numofallchars = 0, perclass[256] = {0};
  for (int c = 1; c < 128; c++) {
    if (orig[c] == 0 || tables.cls[c] != C_SYMBOL) continue;    // (the histogram also saw the white space)
    const uint8_t code = tables.code[c];
    if (code < K) chardist[code] += orig[c];
    perclass[code]++;
    numofallchars++;
  }
  uint64_t maxsub = 0;
  for (unsigned k = 0; k < K; k++) maxsub = std::max(maxsub, perclass[k]);
  maxsub = std::max(maxsub, perclass[CODE_WILDCARD]);
  unsigned lpc = 0;                                             // determineleastprobablecharacter, encseq.c:4468-4485
  for (unsigned k = 1; k < K; k++)
    if (chardist[k] < chardist[lpc]) lpc = k;

  // ---- which representation (doupdatesumranges, encseq.c:5215-5256; determinesmallestrep,
  //      src/core/encseq_access_type.c:96-129): sizes differ only in what follows the header
  uint64_t specialtab[3], wildtab[3];
  ranges_tab(special, specialtab);
  ranges_tab(wild, wildtab);
  const uint64_t twobit_bytes = units_of_twobit(n) * 8;
  {
    uint64_t smallest = 0;
    for (int k = 0; k < 3; k++) {
      const uint64_t size = twobit_bytes + size_of_swtable(k, true, n, wildtab[k]);
      if (k == 0 || size < smallest) {
        smallest = size;
        sci.specialranges = specialtab[k];
        sci.wildcardranges = wildtab[k];
      }
    }
  }
  uint64_t sat = SAT_BITACCESS, table_items = wildtab[0];
  int table_kind = -1;
  // -sat: the representation is given (getsatforcevalue, encseq.c:797-814: a table type also decides which entry
  // counts the header carries; gt_encseq_access_type_determine, encseq_access_type.c:164-247: what the reference
  // refuses -- eqlen without equal lengths, bytecompress for DNA, a 2-bit type for protein -- is left to it)
  const std::string forced = rq->sat != nullptr ? rq->sat : "";
  if (!forced.empty()) {
    static const char *table_names[3] = {"uchar", "ushort", "uint32"};
    int k = -1;
    for (int i = 0; i < 3; i++)
      if (forced == table_names[i]) k = i;
    if (forced == "direct") sat = SAT_DIRECT;
    else if (!dna && forced == "bytecompress") sat = SAT_BYTECOMPRESS;
    else if (dna && forced == "bit") sat = SAT_BITACCESS;
    else if (dna && forced == "eqlen" && equallength) sat = SAT_EQUALLENGTH;
    else if (dna && k >= 0) {
      sat = SAT_UCHAR + (uint64_t) k;
      table_kind = k;
      table_items = wildtab[k];
      sci.specialranges = specialtab[k];
      sci.wildcardranges = wildtab[k];
    } else throw Unsupported{format("-sat %s for this input", forced.c_str())};
  } else if (!dna) sat = SAT_BYTECOMPRESS;                      // gt_encseq_access_type_determine, encseq_access_type.c:152-163
  else if (equallength) sat = SAT_EQUALLENGTH;
  else {
    uint64_t cmin = twobit_bytes + ((wildtab[0] > 0 || numofsequences > 1) ? 8 * ints_for_bits(n + 64) : 0);
    for (int k = 0; k < 3; k++) {
      const uint64_t size = twobit_bytes + size_of_swtable(k, true, n, wildtab[k]);
      if (size < cmin) { cmin = size; sat = SAT_UCHAR + (uint64_t) k; table_kind = k; table_items = wildtab[k]; }
    }
  }
  // the separator table (determineoptimalsssptablerep, encseq.c:1714-1736; files2encodedsequence :4609-4619)
  int sep_kind = -1;
  if (numofsequences > 1 && sat != SAT_EQUALLENGTH && (rq->out_ssp || table_kind >= 0)) {
    uint64_t smallest = size_of_swtable(0, false, n, numofsequences - 1);
    sep_kind = 0;
    for (int k = 1; k < 3; k++) {
      const uint64_t size = size_of_swtable(k, false, n, numofsequences - 1);
      if (size < smallest) { smallest = size; sep_kind = k; }
    }
  }
  const double t_lists = now();

  // ---- pack (fillSWtable / fillViaequallength / fillViabitaccess: a special position holds the least
  //      probable character, with bit access 0 for a wildcard and 1 for a separator)
  const bool twobit = dna && sat != SAT_DIRECT, bitstring = !dna && sat != SAT_DIRECT;
  const uint64_t units = twobit ? units_of_twobit(n) : 0, full = n / 32;
  // protein: a bit string, bits_per_symbol bits per symbol from the top of byte 0 on, wildcard = K, separator = K+1
  // (fillViabytecompress, encseq.c:2324-2440; gt_bsStoreUInt32, src/core/bitpackstringop32.c)
  const unsigned bps = rq->bits_per_symbol;
  const uint64_t packed_bytes = bitstring ? (n * bps + 7) / 8 : 0;
  BigBuffer words_buffer(twobit ? units * 8 : packed_bytes + 16);
  uint64_t *words = static_cast<uint64_t *>(words_buffer.p);   // zeros: the words behind the last symbol stay 0
  uint8_t *packed = static_cast<uint8_t *>(words_buffer.p);
  uint8_t fill[256];
  for (int c = 0; c < 256; c++) fill[c] = (uint8_t) (c < 4 ? c : lpc);
  if (sat == SAT_BITACCESS) { fill[CODE_WILDCARD] = 0; fill[CODE_SEPARATOR] = 1; }
  if (bitstring) {
    for (int c = 0; c < 256; c++) fill[c] = (uint8_t) c;
    fill[CODE_WILDCARD] = (uint8_t) K;
    fill[CODE_SEPARATOR] = (uint8_t) (K + 1);
    const uint64_t groups = n / 8, block = 1 << 16;             // 8 symbols = bps bytes
    parallel_for(nthreads, (size_t) ((groups + block - 1) / block), [&](size_t task) {
      const uint64_t g0 = task * block, g1 = std::min(groups, g0 + block);
      for (uint64_t g = g0; g < g1; g++) {
        uint64_t v = 0;
        for (int j = 0; j < 8; j++) v = (v << bps) | fill[codes[8 * g + j]];
        uint8_t *out = packed + g * bps;
        for (unsigned b = 0; b < bps; b++) out[b] = (uint8_t) (v >> (8 * (bps - 1 - b)));
      }
    });
    if (n % 8) {
      uint64_t v = 0;
      for (uint64_t j = 8 * groups; j < n; j++) v = (v << bps) | fill[codes[j]];
      v <<= bps * (8 - n % 8);
      uint8_t *out = packed + groups * bps;
      const uint64_t left = packed_bytes - groups * bps;
      for (uint64_t b = 0; b < left; b++) out[b] = (uint8_t) (v >> (8 * (bps - 1 - b)));
    }
  } else if (twobit) {
    const uint64_t block = 1 << 15;                              // words per task
    parallel_for(nthreads, (size_t) ((full + block - 1) / block), [&](size_t task) {
      const uint64_t w0 = task * block, w1 = std::min(full, w0 + block);
      for (uint64_t w = w0; w < w1; w++) {
        const uint8_t *c = codes + 32 * w;
        uint64_t v = 0;
        for (int q = 0; q < 4; q++) {                            // 8 codes at a time
          uint64_t x;
          memcpy(&x, c + 8 * q, 8);
          if (x & 0xfcfcfcfcfcfcfcfcull) {                       // a special among them
            uint64_t y = 0;
            for (int j = 0; j < 8; j++) y = (y << 2) | fill[c[8 * q + j]];
            v = (v << 16) | y;
            continue;
          }
          x = __builtin_bswap64(x);                              // first code in the top byte
          x = (x | (x >> 6)) & 0x000f000f000f000full;
          x = (x | (x >> 12)) & 0x000000ff000000ffull;
          x = (x | (x >> 24)) & 0xffffull;
          v = (v << 16) | x;
        }
        words[w] = v;
      }
    });
    if (n % 32) {
      uint64_t v = 0;
      for (uint64_t j = 32 * full; j < n; j++) v = (v << 2) | fill[codes[j]];
      words[full] = v << (2 * (32 - n % 32));
    }
  }
  const double t_pack = now();

  // ---- the files
  {
    Writer w(indexname, ".esq");                                // gt_encseq_assign_header_mapspec, encseq.c:1288-1307
    const uint8_t is64bit = 1;
    w.field(&is64bit, 1, 1);
    w.word(ENCSEQ_VERSION);
    w.word(sat);
    w.word(n);
    w.word(numofsequences);
    w.word(files.size());
    std::string names;                                          // every name with its NUL
    for (const Input &f : files) { names += f.name; names.push_back('\0'); }
    w.word(names.size());
    w.field(&sci, sizeof sci, 1);
    w.word(minseqlen);
    w.word(maxseqlen);
    w.word(rq->alphatype);                                      // 0 DNA, 1 protein (alphabet_to_key_values, encseq.c:1080-1111)
    w.word(0);                                                  // lengthofalphadef
    w.field(names.data(), 1, names.size());
    const uint8_t maxsubalphasize = (uint8_t) maxsub;
    w.field(&maxsubalphasize, 1, 1);
    w.word(numofallchars);
    // GtFilelengthvalues: bytes of the file; symbols + separators it contributed, the separator in
    // front of a later file's first sequence not counted (sequence_buffer_fasta.c:52-146)
    std::vector<uint64_t> flv(2 * files.size());
    for (size_t i = 0; i < files.size(); i++) {
      flv[2 * i] = files[i].len;
      flv[2 * i + 1] = file_symbols[i] + file_headers[i] - 1;
    }
    w.field(flv.data(), 16, files.size());
    w.field(chardist, 8, K);
    if (sat == SAT_DIRECT) w.field(codes, 1, n);                // gt_encseq_assign_sequence_mapspec, encseq.c:1346-1402
    else if (bitstring) w.field(packed, 1, packed_bytes);       // (direct: the codes as they are, fillViadirectaccess :2162-2270)
    else w.field(words, 8, units);
    if (sat == SAT_BITACCESS && (wildtab[0] > 0 || numofsequences > 1)) {
      const uint64_t nw = ints_for_bits(n + 64);
      std::vector<uint64_t> bits(nw, 0);
      auto setbit = [&](uint64_t i) { bits[i >> 6] |= (uint64_t(1) << 63) >> (i & 63); };
      for (const Run &r : special)
        for (uint64_t i = 0; i < r.len; i++) setbit(r.start + i);
      for (uint64_t i = n; i < n + 64; i++) setbit(i);
      w.field(bits.data(), 8, nw);
    } else if (table_kind >= 0)
      write_table_of_kind(w, wild, table_kind, n, table_items, true);
    w.finish();
  }
  if (sep_kind >= 0) {
    Writer w(indexname, ".ssp");
    std::vector<Run> seps(seppos.size());
    for (size_t i = 0; i < seppos.size(); i++) seps[i] = Run{seppos[i], 1};
    write_table_of_kind(w, seps, sep_kind, n, numofsequences - 1, false);
    w.finish();
  }
  if (rq->out_des) {
    // descriptions without '\r', each followed by '\n'; behind the last one the length of the longest
    // and a word of ones; .sds: where each description but the last ends (encseq_charproc.gen,
    // encseq.c:5612-5623; gt_desc_buffer_*, src/core/desc_buffer.c)
    // (the text of every chunk's descriptions is put together in parallel -- a read set has millions of them --
    //  and the chunks' texts are then written one after the other)
    Writer des(indexname, ".des");
    struct DesPart { std::string text; std::vector<uint32_t> lengths; };
    std::vector<DesPart> parts(chunks.size());
    const bool clip = rq->clip_desc != 0;
    parallel_for(nthreads, chunks.size(), [&](size_t k) {
      const Chunk &c = chunks[k];
      const uint8_t *p = files[c.file].p;
      DesPart &part = parts[k];
      part.lengths.reserve(c.headers.size());
      for (const Header &h : c.headers) {
        const size_t before = part.text.size();
        bool clipped = false;
        for (size_t i = h.desc_begin; i < h.desc_end; i++) {
          const uint8_t ch = p[i];
          if (ch == '\r') continue;
          if (clip) {
            if (clipped) continue;
            if (isspace(ch)) { clipped = true; continue; }
          }
          part.text.push_back((char) ch);
        }
        part.lengths.push_back((uint32_t) (part.text.size() - before));
        part.text.push_back('\n');
      }
    });
    std::vector<uint64_t> sds;
    sds.reserve(numofsequences);
    uint64_t longest = 0, seq = 0, base = 0;
    for (const DesPart &part : parts) {
      uint64_t off = base;
      for (const uint32_t len : part.lengths) {
        longest = std::max<uint64_t>(longest, len);
        off += len;
        if (++seq < numofsequences) sds.push_back(off);       // where the description ends, in front of its '\n'
        off += 1;
      }
      des.raw(part.text.data(), part.text.size());
      base += part.text.size();
    }
    const uint64_t fin = ~uint64_t(0);
    des.raw(&longest, 8);
    des.raw(&fin, 8);
    des.finish();
    if (rq->out_sds) {
      Writer w(indexname, ".sds");
      w.raw(sds.data(), 8 * sds.size());
      w.finish();
    }
  } else if (rq->out_sds) {
    Writer w(indexname, ".sds");                                // opened by the reference, never written
    w.finish();
  }
  if (rq->out_md5) {
    if (md5_thread.joinable()) md5_thread.join();
    if (md5_failed) throw IoError{"out of memory (md5 of the sequences)"};
    Writer w(indexname, ".md5");
    w.raw(md5tab.data(), md5tab.size());
    w.finish();
  }
  const double t_write = now();

  if (sum != nullptr) {
    memset(sum, 0, sizeof *sum);
    sum->totallength = n;
    sum->numofsequences = numofsequences;
    sum->numoffiles = files.size();
    sum->specialcharacters = sci.specialcharacters;
    sum->specialranges = sci.specialranges;
    sum->realspecialranges = sci.realspecialranges;
    sum->wildcards = sci.wildcards;
    sum->wildcardranges = sci.wildcardranges;
    sum->realwildcardranges = sci.realwildcardranges;
    sum->sat = sat;
    sum->satsep = sep_kind >= 0 ? SAT_UCHAR + (uint64_t) sep_kind : (uint64_t) SAT_UNDEFINED;
    for (unsigned k = 0; k < K && k < 32; k++) sum->characterdistribution[k] = chardist[k];
    snprintf(sum->satname, sizeof sum->satname, "%s", sat_name(sat));
    sum->threads = nthreads;
    sum->input_bytes = total_bytes;
    sum->seconds_count = t_count - t0;
    sum->seconds_emit = t_emit - t_count;
    sum->seconds_lists = t_lists - t_emit;
    sum->seconds_pack = t_pack - t_lists;
    sum->seconds_md5 = md5_seconds;                   /* beside emit, pack and write */
    sum->seconds_write = t_write - t_pack;
    sum->seconds_total = t_write - t0;
  }
  return GTB_FASTA_OK;
}

}  // namespace

extern "C" int gtb_fasta_encode(const gtb_fasta_request *rq, gtb_fasta_summary *summary, char *msg, size_t msglen)
{
  auto say = [&](const std::string &s) { if (msg != nullptr && msglen > 0) snprintf(msg, msglen, "%s", s.c_str()); };
  if (msg != nullptr && msglen > 0) msg[0] = '\0';
  if (rq == nullptr || rq->filenames == nullptr || rq->indexname == nullptr || rq->symbolmap == nullptr ||
      rq->decode == nullptr) {
    say("gtb_fasta_encode: incomplete request");
    return GTB_FASTA_ERROR;
  }
  try {
    return encode(rq, summary);
  } catch (const Unsupported &u) {
    say(u.why);
    return GTB_FASTA_UNSUPPORTED;
  } catch (const IoError &e) {
    say(e.why);
    return GTB_FASTA_ERROR;
  } catch (const std::exception &e) {
    say(std::string("gtb_fasta_encode: ") + e.what());
    return GTB_FASTA_ERROR;
  } catch (...) {
    say("gtb_fasta_encode: unknown error");
    return GTB_FASTA_ERROR;
  }
}
