/*
  host/b200_encseq.h -- hand a GtEncseq to the sorter: the 2-bit words and the special ranges as the
  reference's own sorter reads them (gt_encseq_twobitencoding_export, src/core/encseq.c:6687;
  gt_specialrangeiterator_*, src/core/encseq.h:127-133), or the extracted symbols for every other
  representation (gt_encseq_extract_encoded).  Shared by the two drop-in objects
  (gt_suffixerator_b200.c, gt_sfxiterator_b200.c).
*/
#ifndef B200_ENCSEQ_H
#define B200_ENCSEQ_H
#include <stdint.h>
#include "core/encseq.h"
#include "core/ma_api.h"
#include "core/range_api.h"
#include "core/thread_api.h"
#include "gtb200.h"

/* GPUs of a run: GTB200_GPUS, else `gt -j N` (gt_jobs, src/gtr.c:181), never more than the box has */
/* how many GPUs the caller asked for: `gt -j N` (gt_jobs, src/gtr.c:181) or GTB200_GPUS; does not ask the
   driver (that is the expensive first call into CUDA) */
static inline int b200_gpus_wanted(void)
{
  const char *e = getenv("GTB200_GPUS");
  int want = e != NULL ? atoi(e) : (int) gt_jobs;
  return want < 1 ? 1 : want;
}

static inline int b200_gpu_count(void)
{
  int want = b200_gpus_wanted(), have = gtb_device_count();
  if (have >= 1 && want > have) want = have;
  return want;
}

static inline int b200_group_set_encseq(gtb_group *g, const GtEncseq *encseq, bool with_separators)
{
  const GtUword n = gt_encseq_total_length(encseq);
  int rc;
  if (gt_encseq_has_twobitencoding(encseq)) {
    const GtTwobitencoding *tbe = gt_encseq_twobitencoding_export(encseq);
    GtUword nranges = 0, fill = 0;
    gtb_range *ranges = NULL;
    if (gt_encseq_has_specialranges(encseq)) {
      GtSpecialrangeiterator *sri = gt_specialrangeiterator_new(encseq, true);
      GtRange range;
      GtUword alloc = gt_encseq_realspecialranges(encseq) + 16;
      ranges = gt_malloc(sizeof *ranges * alloc);
      while (gt_specialrangeiterator_next(sri, &range)) {
        /* the iterator may split one run into several pieces: merge them */
        if (fill > 0 && ranges[fill-1].end == (uint64_t) range.start) {
          ranges[fill-1].end = range.end;
        } else {
          if (fill == alloc) { alloc *= 2; ranges = gt_realloc(ranges, sizeof *ranges * alloc); }
          ranges[fill].start = range.start; ranges[fill].end = range.end; fill++;
        }
      }
      gt_specialrangeiterator_delete(sri);
      nranges = fill;
    }
    rc = gtb_group_set_input_2bit(g, (const uint64_t *) tbe, gt_unitsoftwobitencoding(n), n, ranges, nranges);
    gt_free(ranges);
    if (rc == 0 && with_separators) {
      /* which special positions are separators: one before every sequence but the first */
      const GtUword nseq = gt_encseq_num_of_sequences(encseq);
      uint64_t *sep = gt_malloc(sizeof *sep * (nseq + 1));
      GtUword i;
      for (i = 1; i < nseq; i++) sep[i-1] = (uint64_t) gt_encseq_seqstartpos(encseq, i) - 1;
      rc = gtb_group_set_separators(g, sep, nseq - 1);
      gt_free(sep);
    }
  } else {
    GtUchar *symbols = gt_malloc(n + 1);
    if (n > 0) gt_encseq_extract_encoded(encseq, symbols, 0, n - 1);
    rc = gtb_group_set_input_bytes(g, symbols, n, gt_encseq_alphabetnumofchars(encseq));
    gt_free(symbols);
  }
  return rc;
}

/* one table of a .bck file: gt_mapspec_write pads every table to 8 bytes (src/core/mapspec.c:350-365) */
static inline void b200_append_table(FILE *fp, const uint32_t *tab, uint64_t n)
{
  static const char zeros[8] = {0};
  if (n > 0) gt_xfwrite(tab, sizeof *tab, (size_t) n, fp);
  if ((n * sizeof *tab) % 8 != 0) gt_xfwrite(zeros, 1, 8 - (n * sizeof *tab) % 8, fp);
}
#endif
