/*
  host/gt_sfxiterator_b200.c -- the Sfxiterator of GenomeTools 1.5.11 on top of libgtb200.so.

  SURVEY.md section 8b names this interface (src/match/sfx-suffixer.h:33-72) as the seam of the sort
  core.  This object defines the interface's functions

      gt_Sfxiterator_new, gt_Sfxiterator_new_withadditionalvalues, gt_Sfxiterator_next,
      gt_Sfxiterator_longest, gt_Sfxiterator_bcktab2file, gt_Sfxiterator_delete,
      gt_Sfxiterator_postsortfromstream

  and is linked into `gt` AHEAD of src/match/sfx-suffixer.o, so that every consumer that pulls the
  suffix array through the iterator -- `gt packedindex mkindex` (src/match/eis-suffixerator-interface.c:
  253,399), the maximal-match search on the fly (src/match/esa-mmsearch.c:608,635) -- gets it from the
  GPU while its own code stays untouched.  The consumers see what they see today: one
  GtSuffixsortspace (the reference's own type, created and read through its own functions,
  src/match/sfx-suffixgetset.h) with all suffixes that start at a regular symbol, sorted; then the
  special positions in ascending order and the final entry n; then NULL.  gt_Sfxiterator_next hands out
  the whole non-special table as one part whatever `numofparts` says (parts bound the memory of the
  CPU sorter; the output of the reference does not depend on them, testsuite/gt_suffixerator_include.rb:
  64-68).

  Outside the path, failing loudly: an lcp side channel (voidoutlcpinfo -- only the reference's own
  suffixerator tool passes one, and that tool is replaced by gt_suffixerator_b200.c), difference covers,
  -spmopt, sort depth limits, uint32 tables, compressed output, postsortfromstream.
  Written from scratch; no reference code is copied.
*/
#include <limits.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "core/encseq.h"
#include "core/error_api.h"
#include "core/fa.h"
#include "core/logger.h"
#include "core/ma_api.h"
#include "core/readmode.h"
#include "core/timer_api.h"
#include "core/xansi_api.h"
#include "match/sfx-suffixer.h"
#include "match/sfx-suffixgetset.h"
#include "gtb200.h"
#include "b200_encseq.h"

struct Sfxiterator {
  GtSuffixsortspace *regular,     /* entries 0 .. n-S-1 of the suffix table                    */
                    *special;     /* the special positions ascending, then n                   */
  GtUword nonspecials, specials_plus_one, longest;
  int stage;                      /* 0: nothing handed out yet, 1: regular part out, 2: all out */
  FILE *outfpbcktab;
  uint32_t *leftborder, *csc, *dist;
  uint64_t nall, nspec, ndist;
  GtLogger *logger;
};

static int b200_strategy_unsupported(const Sfxstrategy *st, void *voidoutlcpinfo, GtError *err)
{
  const char *what = NULL;
  if (voidoutlcpinfo != NULL) what = "an lcp side channel (GtOutlcpinfo)";
  else if (st != NULL && st->differencecover > 0) what = "a difference cover";
  else if (st != NULL && st->spmopt_minlength > 0) what = "-spmopt";
  else if (st != NULL && st->userdefinedsortmaxdepth > 0) what = "a sort depth limit";
  else if (st != NULL && st->suftabuint) what = "uint32 suffix tables";
  else if (st != NULL && st->compressedoutput) what = "compressed output";
  else if (st != NULL && st->onlybucketinsertion) what = "-onlybucketinsertion";
  if (what != NULL) {
    gt_error_set(err, "%s is not supported by the B200 Sfxiterator (no silent fallback); "
                      "use the CPU build of gt for it", what);
    return -1;
  }
  return 0;
}

Sfxiterator *gt_Sfxiterator_new_withadditionalvalues(const GtEncseq *encseq, GtReadmode readmode,
                                                     unsigned int prefixlength, unsigned int numofparts,
                                                     GtUword maximumspace, void *voidoutlcpinfo,
                                                     FILE *outfpbcktab, const Sfxstrategy *sfxstrategy,
                                                     GtTimer *sfxprogress, bool withprogressbar,
                                                     GtLogger *logger, GtError *err)
{
  Sfxiterator *sfi;
  const GtUword n = gt_encseq_total_length(encseq),
                S = gt_encseq_specialcharacters(encseq);
  const unsigned int numofchars = gt_encseq_alphabetnumofchars(encseq);
  int devices[64], ndev, i, rc = 0;
  gtb_group *g;
  gtb_stats stats;
  char msg[512];
  const uint64_t *table;

  (void) numofparts; (void) maximumspace; (void) withprogressbar;
  gt_error_check(err);
  if (b200_strategy_unsupported(sfxstrategy, voidoutlcpinfo, err) != 0) return NULL;
  if (gt_encseq_is_mirrored(encseq)) {
    gt_error_set(err, "mirrored sequences are not supported by the B200 Sfxiterator (no silent fallback)");
    return NULL;
  }
  if ((uint64_t) n + 1 >= (uint64_t) UINT32_MAX) {
    gt_error_set(err, "sequences of total length >= 2^32-2 are not supported by the B200 Sfxiterator");
    return NULL;
  }
  if (prefixlength == 0) prefixlength = 1;
  if (sfxprogress != NULL) gt_timer_show_progress(sfxprogress, "sorting the suffixes on the GPU", stdout);
  ndev = b200_gpu_count();
  for (i = 0; i < ndev; i++) devices[i] = i;
  g = gtb_group_new(devices, ndev, msg, sizeof msg);
  if (g == NULL) { gt_error_set(err, "libgtb200: %s", msg); return NULL; }
  sfi = gt_calloc(1, sizeof *sfi);
  sfi->logger = logger;
  sfi->outfpbcktab = outfpbcktab;
  sfi->nonspecials = n - S;
  sfi->specials_plus_one = S + 1;
  rc = gtb_group_set_readmode(g, (unsigned) readmode);
  if (rc == 0) rc = b200_group_set_encseq(g, encseq, false);
  if (rc == 0) rc = gtb_group_run(g, prefixlength, GTB_WANT_SUF | (outfpbcktab != NULL ? GTB_WANT_BCK : 0u));
  if (rc == 0 && gtb_group_num_entries(g) != (uint64_t) n + 1) {
    snprintf(msg, sizeof msg, "internal: %llu suffix-table entries for a sequence of length %llu",
             (unsigned long long) gtb_group_num_entries(g), (unsigned long long) n);
    rc = -2;
  }
  if (rc == 0) {
    /* the reference's own container for a piece of the suffix table: GtUword entries (useuint = false) */
    sfi->regular = gt_suffixsortspace_new(n + 1, n, false, logger);
    gt_suffixsortspace_nooffsets(sfi->regular);
    table = gt_suffixsortspace_getptr_ulong(sfi->regular, 0);
    if (outfpbcktab != NULL) {
      gtb_bck_sizes(numofchars, prefixlength, &sfi->nall, &sfi->nspec, &sfi->ndist);
      sfi->leftborder = gt_malloc(sizeof *sfi->leftborder * (sfi->nall + 1));
      sfi->csc = gt_malloc(sizeof *sfi->csc * (sfi->nspec + 1));
      sfi->dist = gt_malloc(sizeof *sfi->dist * (sfi->ndist + 1));
    }
    /* every GPU copies its shard to its offset of the one table */
    rc = gtb_group_copy_results(g, (uint64_t *) table, NULL, NULL, sfi->leftborder, sfi->csc, sfi->dist);
    if (rc == 0) rc = gtb_group_get_stats(g, &stats);
  }
  if (rc == 0) {
    GtUword k;
    sfi->longest = (GtUword) stats.longest;
    sfi->special = gt_suffixsortspace_new(S + 1, n, false, logger);
    gt_suffixsortspace_nooffsets(sfi->special);
    memcpy((GtUword *) gt_suffixsortspace_getptr_ulong(sfi->special, 0),
           gt_suffixsortspace_getptr_ulong(sfi->regular, 0) + sfi->nonspecials, sizeof (GtUword) * (S + 1));
    (void) k;
    gt_logger_log(logger, "B200 Sfxiterator: %d GPU(s), device time %.3f ms, %u kernel launches",
                  ndev, stats.ms_total, stats.kernel_launches);
  }
  if (rc == -1) snprintf(msg, sizeof msg, "%s", gtb_group_error(g));
  gtb_group_delete(g);
  if (rc != 0) {
    gt_error_set(err, "libgtb200: %s", msg);
    (void) gt_Sfxiterator_delete(sfi, NULL);
    return NULL;
  }
  return sfi;
}

Sfxiterator *gt_Sfxiterator_new(const GtEncseq *encseq, GtReadmode readmode, unsigned int prefixlength,
                                unsigned int numofparts, GtUword maximumspace,
                                const Sfxstrategy *sfxstrategy, GtTimer *sfxprogress,
                                bool withprogressbar, GtLogger *logger, GtError *err)
{
  return gt_Sfxiterator_new_withadditionalvalues(encseq, readmode, prefixlength, numofparts, maximumspace,
                                                 NULL, NULL, sfxstrategy, sfxprogress, withprogressbar,
                                                 logger, err);
}

const GtSuffixsortspace *gt_Sfxiterator_next(GtUword *numberofsuffixes, bool *specialsuffixes,
                                             Sfxiterator *sfi)
{
  if (sfi->stage == 0) {
    sfi->stage = 1;
    if (sfi->nonspecials > 0) {
      *numberofsuffixes = sfi->nonspecials;
      if (specialsuffixes != NULL) *specialsuffixes = false;
      return sfi->regular;
    }
  }
  if (sfi->stage == 1) {
    sfi->stage = 2;
    *numberofsuffixes = sfi->specials_plus_one;
    if (specialsuffixes != NULL) *specialsuffixes = true;
    return sfi->special;
  }
  return NULL;
}

GtUword gt_Sfxiterator_longest(const Sfxiterator *sfi)
{
  return sfi->longest;
}

/* writes the bucket table and closes the file, as the reference does (sfx-suffixer.c:2206-2214) */
int gt_Sfxiterator_bcktab2file(FILE *fp, Sfxiterator *sfi, GtError *err)
{
  gt_error_check(err);
  if (sfi->leftborder == NULL) {
    gt_error_set(err, "the B200 Sfxiterator was created without a bucket-table file");
    if (fp != NULL) gt_fa_fclose(fp);
    return -1;
  }
  b200_append_table(fp, sfi->leftborder, sfi->nall + 1);
  b200_append_table(fp, sfi->csc, sfi->nspec);
  b200_append_table(fp, sfi->dist, sfi->ndist);
  gt_fa_fclose(fp);
  if (fp == sfi->outfpbcktab) sfi->outfpbcktab = NULL;
  return 0;
}

int gt_Sfxiterator_postsortfromstream(Sfxiterator *sfi, const GtStr *indexname, GtError *err)
{
  (void) sfi; (void) indexname;
  gt_error_set(err, "gt_Sfxiterator_postsortfromstream (difference-cover post-sorting) is not supported by "
                    "the B200 Sfxiterator (no silent fallback)");
  return -1;
}

int gt_Sfxiterator_delete(Sfxiterator *sfi, GtError *err)
{
  int had_err = 0;
  if (sfi == NULL) return 0;
  if (sfi->outfpbcktab != NULL && sfi->leftborder != NULL)     /* not flushed by the caller: as parts > 1 */
    had_err = gt_Sfxiterator_bcktab2file(sfi->outfpbcktab, sfi, err);
  if (sfi->regular != NULL) gt_suffixsortspace_delete(sfi->regular, false);
  if (sfi->special != NULL) gt_suffixsortspace_delete(sfi->special, false);
  gt_free(sfi->leftborder); gt_free(sfi->csc); gt_free(sfi->dist);
  gt_free(sfi);
  return had_err;
}
