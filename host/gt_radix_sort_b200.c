/*
  host/gt_radix_sort_b200.c -- the in-place record sorts of src/core/radix_sort.h on the GPU.

  Defines  gt_radixsort_inplace_ulong / _GtUwordPair / _Gtuint64keyPair
  (src/core/radix_sort.h:91,107,125: the sorters behind `gt dev sortbench -impl radixinplace |
  radixkeypair` and `gt encseq2spm`) on top of gtb_radixsort_u64 / _u64pair / _u64keypair of
  libgtb200.so, and is linked AHEAD of src/core/radix_sort.o (whose other functions -- the
  GtRadixsortinfo workspace API -- the archive still needs; see host/Makefile).  Same contracts: sorted in
  place, ascending; GtUwordPair by component a (equal keys: order unspecified in the reference, input order
  here), Gtuint64keyPair by (a, b).  A failure of the device path is a programming error for these void
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
