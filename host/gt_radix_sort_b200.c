/*
  host/gt_radix_sort_b200.c -- the in-place record sorts of src/core/radix_sort.h on the GPU.

  Defines  gt_radixsort_inplace_ulong / _GtUwordPair / _Gtuint64keyPair / _flba
  (src/core/radix_sort.h:91,107,125,138: the sorters behind `gt dev sortbench -impl radixinplace |
  radixkeypair | radixflba`, `gt encseq2spm` and the seed lists of src/match/diagbandseed.c:1067,2476) on
  top of gtb_radixsort_u64 / _u64pair / _u64keypair of
  libgtb200.so, and is linked AHEAD of src/core/radix_sort.o (whose other functions -- the
  GtRadixsortinfo workspace API -- the archive still needs; see host/Makefile).  Same contracts: sorted in
  place, ascending; GtUwordPair by component a (equal keys: order unspecified in the reference, input order
  here), Gtuint64keyPair by (a, b); flba = records of `unitsize` bytes in the order of memcmp, carried as
  one big-endian 64-bit key (up to 8 bytes) or a key pair (up to 16; longer records are refused).  A failure of the device path is a programming error for these void
  functions, reported like the reference reports its own (exit code GT_EXIT_PROGRAMMING_ERROR).
  Written from scratch; no reference code is copied.
*/
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "core/error_api.h"
#include "core/radix_sort.h"
#include "gtb200.h"

static void b200_radix_fail(const char *what, const char *msg)
{
  fprintf(stderr, "libgtb200: %s: %s\n", what, msg);
  exit(GT_EXIT_PROGRAMMING_ERROR);
}

void gt_radixsort_inplace_ulong(GtUword *source, GtUword len)
{
  char msg[256];
  if (gtb_radixsort_u64(0, (uint64_t *) source, (uint64_t) len, msg, sizeof msg) != 0)
    b200_radix_fail("gt_radixsort_inplace_ulong", msg);
}

void gt_radixsort_inplace_GtUwordPair(GtUwordPair *source, GtUword len)
{
  char msg[256];
  if (gtb_radixsort_u64pair(0, (uint64_t *) source, (uint64_t) len, msg, sizeof msg) != 0)
    b200_radix_fail("gt_radixsort_inplace_GtUwordPair", msg);
}

void gt_radixsort_inplace_Gtuint64keyPair(Gtuint64keyPair *source, GtUword len)
{
  char msg[256];
  if (gtb_radixsort_u64keypair(0, (uint64_t *) source, (uint64_t) len, msg, sizeof msg) != 0)
    b200_radix_fail("gt_radixsort_inplace_Gtuint64keyPair", msg);
}

/* fixed-length byte arrays (src/core/radix_sort.c:769-778, radixsort-ip-flba.inc: most significant byte
   first): equal records are the same bytes, so the sorted array is unique -- the records travel as
   big-endian integers, left-aligned, and come back the same way */
void gt_radixsort_inplace_flba(uint8_t *source, GtUword len, size_t unitsize)
{
  char msg[256];
  const size_t words = unitsize <= 8 ? 1 : 2;
  uint64_t *keys;
  GtUword i;
  size_t b;

  if (len < 2 || unitsize == 0) return;
  if (unitsize > 16) {
    snprintf(msg, sizeof msg, "records of %lu bytes (at most 16 are carried)", (unsigned long) unitsize);
    b200_radix_fail("gt_radixsort_inplace_flba", msg);
  }
  keys = malloc(sizeof *keys * words * (size_t) len);
  if (keys == NULL) b200_radix_fail("gt_radixsort_inplace_flba", "out of memory");
  for (i = 0; i < len; i++) {
    const uint8_t *rec = source + (size_t) i * unitsize;
    uint64_t hi = 0, lo = 0;
    for (b = 0; b < unitsize && b < 8; b++) hi |= (uint64_t) rec[b] << (56 - 8 * b);
    for (b = 8; b < unitsize; b++) lo |= (uint64_t) rec[b] << (56 - 8 * (b - 8));
    keys[words * i] = hi;
    if (words == 2) keys[2 * i + 1] = lo;
  }
  if ((words == 1 ? gtb_radixsort_u64(0, keys, (uint64_t) len, msg, sizeof msg)
                  : gtb_radixsort_u64keypair(0, keys, (uint64_t) len, msg, sizeof msg)) != 0)
    b200_radix_fail("gt_radixsort_inplace_flba", msg);
  for (i = 0; i < len; i++) {
    uint8_t *rec = source + (size_t) i * unitsize;
    const uint64_t hi = keys[words * i], lo = words == 2 ? keys[2 * i + 1] : 0;
    for (b = 0; b < unitsize && b < 8; b++) rec[b] = (uint8_t) (hi >> (56 - 8 * b));
    for (b = 8; b < unitsize; b++) rec[b] = (uint8_t) (lo >> (56 - 8 * (b - 8)));
  }
  free(keys);
}
