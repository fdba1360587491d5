/*
  oracle/ref_driver.c -- TEST INFRASTRUCTURE, not product code.

  A small main() that links against objects compiled (by oracle/Makefile)
  directly from the UNMODIFIED reference sources under /root/reference and
  exposes the two reference entry points the parity tests need:

    gtref suffixerator <args>   -> gt_parseargsandcallsuffixerator
                                   (/root/reference/src/match/sfx-run.c:719)
    gtref sfxmap <args>         -> gt_sfxmap (brute-force ESA verifier,
                                   /root/reference/src/tools/gt_sfxmap.c)
    gtref radixsort ulong|ulongpair|keypair|flba<bytes> <in> <out>
                                -> gt_radixsort_inplace_ulong / _GtUwordPair /
                                   _Gtuint64keyPair / _flba (/root/reference/src/core/radix_sort.h:91,107,125,138)
                                   on a file of raw uint64 values (flba: records of <bytes> bytes);
                                   prints the seconds of the call

  It replaces src/gt.c + the toolbox so that only the files on the
  suffixerator path have to be compiled.  No reference source is copied.
*/
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include "core/init_api.h"
#include "core/error_api.h"
#include "core/tool_api.h"
#include "core/ma_api.h"
#include "core/radix_sort.h"
#include "core/timer_api.h"
#include "core/thread_api.h"
#include "match/sfx-run.h"
#include <sys/time.h>

static int ref_radixsort(int argc, char **argv)
{
  FILE *fp;
  long bytes;
  size_t n;
  uint64_t *buf;
  struct timeval t0, t1;

  if (argc != 5) { fprintf(stderr, "usage: radixsort ulong|ulongpair|keypair|flba<bytes> <in> <out>\n"); return 1; }
  fp = fopen(argv[3], "rb");
  if (fp == NULL) { perror(argv[3]); return 1; }
  fseek(fp, 0, SEEK_END); bytes = ftell(fp); fseek(fp, 0, SEEK_SET);
  n = (size_t) bytes / sizeof (uint64_t);
  buf = gt_malloc(bytes > 0 ? (size_t) bytes + 8 : 8);
  if (fread(buf, 1, (size_t) bytes, fp) != (size_t) bytes) { fclose(fp); return 1; }
  fclose(fp);
  gettimeofday(&t0, NULL);
  if (strcmp(argv[2], "ulong") == 0) gt_radixsort_inplace_ulong((GtUword*) buf, (GtUword) n);
  else if (strcmp(argv[2], "ulongpair") == 0) gt_radixsort_inplace_GtUwordPair((GtUwordPair*) buf, (GtUword) (n / 2));
  else if (strcmp(argv[2], "keypair") == 0) gt_radixsort_inplace_Gtuint64keyPair((Gtuint64keyPair*) buf, (GtUword) (n / 2));
  else if (strncmp(argv[2], "flba", 4) == 0 && atoi(argv[2] + 4) > 0)   /* records of that many bytes */
    gt_radixsort_inplace_flba((uint8_t*) buf, (GtUword) ((size_t) bytes / (size_t) atoi(argv[2] + 4)),
                              (size_t) atoi(argv[2] + 4));
  else { fprintf(stderr, "unknown kind %s\n", argv[2]); return 1; }
  gettimeofday(&t1, NULL);
  printf("%.6f\n", (double) (t1.tv_sec - t0.tv_sec) + 1e-6 * (double) (t1.tv_usec - t0.tv_usec));
  fp = fopen(argv[4], "wb");
  if (fp == NULL) { perror(argv[4]); return 1; }
  fwrite(buf, 1, (size_t) bytes, fp);
  fclose(fp);
  gt_free(buf);
  return 0;
}

GtTool* gt_sfxmap(void);  /* /root/reference/src/tools/gt_sfxmap.c:1530 */

int main(int argc, char **argv)
{
  GtError *err;
  int rval = 1;
  unsigned int jobs = 1;

  /* `gtref -j N <tool> ...`: the global option of `gt` (/root/reference/src/gtr.c:181) that sets
     gt_jobs, the number of sorter threads (the SA-only multi-core figure of bench.py) */
  if (argc >= 4 && strcmp(argv[1], "-j") == 0) {
    int j = atoi(argv[2]);
    jobs = j > 0 ? (unsigned int) j : 1u;
    argv[2] = argv[0];
    argv += 2; argc -= 2;
  }
  if (argc < 2) {
    fprintf(stderr, "usage: %s [-j N] suffixerator|sfxmap [options]\n", argv[0]);
    return 2;
  }
  gt_lib_init();
  gt_jobs = jobs;
  err = gt_error_new();
  gt_error_set_progname(err, argv[0]);
  if (strcmp(argv[1], "radixsort") == 0)
    rval = ref_radixsort(argc, argv);
  else if (strcmp(argv[1], "suffixerator") == 0)
    rval = gt_parseargsandcallsuffixerator(true, argc - 1,
                                           (const char**) argv + 1, err);
  else if (strcmp(argv[1], "packedindex_mkindex") == 0)   /* gt packedindex mkindex, src/tools/gt_packedindex.c:33-36 */
    rval = gt_parseargsandcallsuffixerator(false, argc - 1,
                                           (const char**) argv + 1, err);
#ifdef GTREF_WITH_SFXMAP
  else if (strcmp(argv[1], "sfxmap") == 0) {
    GtTool *tool = gt_sfxmap();
    rval = gt_tool_run(tool, argc - 1, (const char**) argv + 1, err);
    gt_tool_delete(tool);
  }
#endif
  else
    fprintf(stderr, "unknown subcommand %s\n", argv[1]);
  if (gt_error_is_set(err))
    fprintf(stderr, "%s: error: %s\n", argv[0], gt_error_get(err));
  gt_error_delete(err);
  if (gt_lib_clean())
    return 3;
  return rval ? 1 : 0;
}
