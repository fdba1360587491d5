/*
  host/gt_suffixerator_b200.c -- the drop-in `gt suffixerator` for GenomeTools 1.5.11.

  This object defines   int gt_suffixerator(int argc, const char **argv, GtError *err)
  and is linked into `gt` IN PLACE OF obj/src/tools/gt_suffixerator.o (the toolbox
  binds the tool by symbol, src/gtt.c:188).  Everything around the sort core stays the
  reference's own code, called through its public functions:

    * option parsing / -help:  gt_suffixeratoroptions        (src/match/sfx-opt.c:225)
    * FASTA -> GtEncseq:       gt_encseq_encoder_encode, gt_encseq_loader_load
                               (as src/match/sfx-run.c:496-531 does); for DNA in plain FASTA files
                               the index files are written by gtb_fasta_encode instead (all host
                               cores, byte-identical files) and only LOADED by the reference
    * prefix length policy:    gt_recommendedprefixlength, gt_whatisthemaximalprefixlength,
                               gt_checkprefixlength          (src/match/sfx-apfxlen.c)
    * project file:            gt_outprjfile                 (src/match/sfx-outprj.c:84)

  What it replaces is suffixeratorwithoutput() (src/match/sfx-run.c:212-317), i.e. the
  Sfxiterator + GtOutlcpinfo pair: the packed sequence is exported from the GtEncseq
  (gt_encseq_twobitencoding_export + gt_specialrangeiterator_*, or
  gt_encseq_extract_encoded for non-2-bit alphabets), handed to libgtb200.so through the
  C-ABI of include/gtb200.h, and the tables that come back are written in the reference's
  file formats (.suf .lcp .llv .bck).  Written from scratch; no reference code is copied.

  Several GPUs: `gt -j N suffixerator ...` (the global option that sets gt_jobs, src/gtr.c:181;
  environment GTB200_GPUS overrides it) puts one bucket-code range on each of N GPUs of the
  box; `-parts p` cuts p ranges per GPU.  All ranges run inside this process (gtb_group): the
  packed sequence is replicated to every GPU, the ranges reach each other's HBM through peer
  access, and every GPU copies its shard straight to its offset in the one host table -- the
  concatenated shards are the global suffix array (src/match/sfx-partssuf.c:172-347).

  Options outside the accelerated path fail loudly (no silent CPU fallback).
  Build: see host/Makefile (needs the reference tree for headers and libgenometools.a).
*/
#include <limits.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/time.h>
#include <sys/types.h>
#include <unistd.h>
#include <stdio.h>
#include "core/alphabet.h"
#include "core/chardef.h"
#include "core/encseq.h"
#include "core/error_api.h"
#include "core/fa.h"
#include "core/logger.h"
#include "core/ma_api.h"
#include "core/range_api.h"
#include "core/readmode.h"
#include "core/str_api.h"
#include "core/thread_api.h"
#include "core/xansi_api.h"
#include "match/bcktab.h"
#include "match/sfx-apfxlen.h"
#include "match/sfx-opt.h"
#include "match/sfx-outprj.h"
#include "match/sfx-strategy.h"
#include "gtb200.h"
#include "b200_encseq.h"

static int b200_unsupported(const Suffixeratoroptions *so, GtError *err)
{
  Sfxstrategy st = gt_index_options_sfxstrategy_value(so->idxopts);
  const char *what = NULL;

  /* a mirrored GtEncseq reports 2n+1 symbols while gt_encseq_twobitencoding_export holds n */
  if (gt_encseq_options_mirrored_value(so->loadopts)) what = "-mirrored";
  else if (gt_index_options_outkystab_value(so->idxopts)) what = "-kys";
  else if (gt_index_options_lcpdist_value(so->idxopts)) what = "-lcpdist";
  else if (gt_index_options_maximumspace_value(so->idxopts) > 0) what = "-memlimit";
  else if (so->genomediff) what = "-genomediff";
  else if (st.differencecover > 0) what = "-dc";
  else if (st.spmopt_minlength > 0) what = "-spmopt";
  else if (st.userdefinedsortmaxdepth > 0) what = "-sortmaxdepth";
  else if (st.suftabuint) what = "-suftabuint";
  else if (st.compressedoutput) what = "-compressedoutput";
  else if (st.onlybucketinsertion) what = "-onlybucketinsertion";
  if (what != NULL) {
    gt_error_set(err, "option %s is not supported by the B200 suffixerator path "
                      "(no silent fallback); use the CPU build of gt for it", what);
    return -1;
  }
  return 0;
}

static double b200_now(void)
{
  struct timeval tv;
  gettimeofday(&tv, NULL);
  return (double) tv.tv_sec + 1e-6 * (double) tv.tv_usec;
}

/* The first call into the CUDA runtime initialises the driver (0.5-1.8 s in a fresh process on the bench
   boxes), creating the sorter objects builds the contexts (0.2-0.8 s): both happen on a thread of their own
   that is started before the FASTA files are encoded -- the main thread asks the driver nothing, not even
   how many GPUs there are, until it has joined that thread. */
typedef struct {
  gtb_group *group;
  int want_devices;              /* gt -j N / GTB200_GPUS */
  unsigned int parts;            /* -parts */
  int ndevices, nranges;         /* what the thread settled on */
  char msg[512];
} B200Init;

static void *b200_release_thread(void *p)
{
  (void) p;
  (void) gtb_release_devices();
  return NULL;
}

static void *b200_init_thread(void *p)
{
  B200Init *init = p;
  int devices[64], i, have = gtb_device_count();
  unsigned int parts = init->parts < 1 ? 1u : init->parts;
  init->ndevices = init->want_devices < 1 ? 1 : init->want_devices;
  if (have >= 1 && init->ndevices > have) init->ndevices = have;
  if ((unsigned int) init->ndevices * parts > 64u) parts = 64u / (unsigned int) init->ndevices;
  init->nranges = init->ndevices * (int) parts;
  for (i = 0; i < init->nranges; i++) devices[i] = i % init->ndevices;   /* range i on GPU i mod N */
  init->group = gtb_group_new(devices, init->nranges, init->msg, sizeof init->msg);
  return NULL;
}

/* FASTA -> .esq/.ssp/.des/.sds/.md5 with gtb_fasta_encode (all host cores; include/gtb200.h) where
   it covers the request: DNA or protein (given with -dna / -protein or guessed by the reference from the
   first file), no -plain / -lossless / -smap.  Returns 0 when the files are written, 1 when the reference's
   encoder has to run (also for every input the library declines: gt_encseq_encoder_encode then words
   the error messages), -1 on an I/O error.  GTB200_ENCODER=reference switches it off. */
static int b200_fast_encode(Suffixeratoroptions *so, GtLogger *logger, GtError *err)
{
  GtEncseqOptions *o = so->encopts;
  const char *which = getenv("GTB200_ENCODER");
  GtAlphabet *alpha = NULL;
  gtb_fasta_request rq;
  gtb_fasta_summary sum;
  const char **names;
  char decode[256], msg[1024];
  GtUword i, nfiles = gt_str_array_size(so->db);
  int rc;

  if (which != NULL && strcmp(which, "reference") == 0) return 1;
  if (nfiles == 0 || gt_encseq_options_plain_value(o) || gt_encseq_options_plain_value(so->loadopts) ||
      gt_encseq_options_lossless_value(o) || gt_str_length(gt_encseq_options_smap_value(o)) > 0)
    return 1;
  if (gt_encseq_options_dna_value(o)) alpha = gt_alphabet_new_dna();
  else if (gt_encseq_options_protein_value(o)) alpha = gt_alphabet_new_protein();
  else {                                 /* gt_encseq_new_from_files, src/core/encseq.c:7560-7569 */
    alpha = gt_alphabet_new_from_sequence(so->db, err);
    if (alpha == NULL) { gt_error_unset(err); return 1; }
  }
  if (!gt_alphabet_is_dna(alpha) && !gt_alphabet_is_protein(alpha)) { gt_alphabet_delete(alpha); return 1; }
  memset(decode, 0, sizeof decode);
  for (i = 0; i < (GtUword) gt_alphabet_num_of_chars(alpha); i++) decode[i] = gt_alphabet_decode(alpha, (GtUchar) i);
  decode[WILDCARD] = gt_alphabet_decode(alpha, (GtUchar) WILDCARD);
  names = gt_malloc(sizeof *names * nfiles);
  for (i = 0; i < nfiles; i++) names[i] = gt_str_array_get(so->db, i);
  memset(&rq, 0, sizeof rq);
  rq.filenames = names;
  rq.numoffiles = nfiles;
  rq.indexname = gt_str_get(so->indexname);
  rq.symbolmap = gt_alphabet_symbolmap(alpha);
  rq.decode = decode;
  rq.numofchars = gt_alphabet_num_of_chars(alpha);
  rq.alphatype = gt_alphabet_is_dna(alpha) ? 0u : 1u;
  rq.bits_per_symbol = gt_alphabet_bits_per_symbol(alpha);
  rq.out_des = gt_encseq_options_des_value(o);
  rq.out_sds = gt_encseq_options_sds_value(o);
  rq.out_ssp = gt_encseq_options_ssp_value(o);
  rq.out_md5 = gt_encseq_options_md5_value(o);
  rq.clip_desc = gt_encseq_options_clip_desc_value(o);
  rq.sat = gt_str_get(gt_encseq_options_sat_value(o));
  rq.threads = 0;
  rc = gtb_fasta_encode(&rq, &sum, msg, sizeof msg);
  gt_free(names);
  gt_alphabet_delete(alpha);
  if (rc == GTB_FASTA_UNSUPPORTED) {
    gt_logger_log(logger, "B200 encoder does not cover this input (%s): the reference's encoder runs", msg);
    return 1;
  }
  if (rc != GTB_FASTA_OK) { gt_error_set(err, "libgtb200: %s", msg); return -1; }
  gt_logger_log(logger, "B200 encoder: %llu symbols in %llu sequence(s), representation %s, %u threads, %.3f s "
                        "(count %.3f, emit %.3f, pack %.3f, write %.3f; md5 %.3f beside emit, pack and write)",
                (unsigned long long) sum.totallength, (unsigned long long) sum.numofsequences, sum.satname,
                sum.threads, sum.seconds_total, sum.seconds_count, sum.seconds_emit, sum.seconds_pack,
                sum.seconds_write, sum.seconds_md5);
  return 0;
}

/* the big result tables: anonymous mappings that ask for huge pages (the first touch of 512 MB in 4 KB pages
   is 131 000 page faults under the copy threads of the library) */
static void *b200_big_alloc(size_t bytes)
{
  const size_t huge = (size_t) 2 << 20;
  size_t size = (bytes + huge - 1) / huge * huge;
  void *p = mmap(NULL, size, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (p == MAP_FAILED) {
    fprintf(stderr, "cannot map %lu bytes of memory\n", (unsigned long) size);
    exit(EXIT_FAILURE);                      /* what gt_malloc does when memory runs out */
  }
#ifdef MADV_HUGEPAGE
  (void) madvise(p, size, MADV_HUGEPAGE);
#endif
  return p;
}

static void b200_big_free(void *p, size_t bytes)
{
  const size_t huge = (size_t) 2 << 20;
  if (p != NULL) (void) munmap(p, (bytes + huge - 1) / huge * huge);
}

/* The sizes of .suf and .lcp are known as soon as the sequences are encoded, long before their contents:
   while the driver builds the CUDA context and the GPU sorts, the host has nothing to do.  A thread uses that
   time to write the two files once with zeros, so that the page cache holds their pages when the tables
   arrive; the real write then only copies into pages that exist (most of what a buffered write costs is
   getting the pages).  The thread is stopped and joined before the real write of a file starts; the file is
   then opened without truncation.  GTB200_PREFILL=0 switches it off. */
typedef struct {
  char path[2][4096];
  size_t bytes[2];
  volatile int stop;
  int done[2];                 /* the file exists and has been written up to bytes[i] (or the stop) */
  pthread_t tid;
  bool started;
} B200Prefill;

static void *b200_prefill_thread(void *p)
{
  B200Prefill *pf = p;
  const size_t chunk = (size_t) 4 << 20;
  char *zeros = calloc(1, chunk);
  int i;
  if (zeros == NULL) return NULL;
  for (i = 0; i < 2 && !pf->stop; i++) {
    size_t off = 0;
    int fd;
    if (pf->bytes[i] == 0) continue;
    fd = open(pf->path[i], O_WRONLY | O_CREAT | O_TRUNC, 0666);
    if (fd < 0) continue;
    pf->done[i] = 1;
    while (off < pf->bytes[i] && !pf->stop) {
      size_t want = pf->bytes[i] - off < chunk ? pf->bytes[i] - off : chunk;
      ssize_t got = write(fd, zeros, want);
      if (got <= 0) break;
      off += (size_t) got;
    }
    close(fd);
  }
  free(zeros);
  return NULL;
}

static void b200_prefill_finish(B200Prefill *pf)
{
  if (!pf->started) return;
  pf->stop = 1;
  pthread_join(pf->tid, NULL);
  pf->started = false;
}

static int b200_write(const char *indexname, const char *suffix, const void *data,
                      size_t size, size_t nmemb, size_t pad_to, bool prefilled, GtError *err)
{
  static const char zeros[8] = {0};
  const size_t bytes = size * nmemb;
  FILE *fp = NULL;
  if (prefilled) {             /* keep the pages: no truncation; the length is set below */
    fp = gt_fa_fopen_with_suffix(indexname, suffix, "rb+", err);
    if (fp != NULL && ftruncate(fileno(fp), (off_t) bytes) != 0) { gt_fa_xfclose(fp); fp = NULL; }
    if (fp == NULL) gt_error_unset(err);
  }
  if (fp == NULL) fp = gt_fa_fopen_with_suffix(indexname, suffix, "wb", err);
  if (fp == NULL) return -1;
  if (nmemb > 0) gt_xfwrite(data, size, nmemb, fp);
  if (pad_to > 0 && bytes % pad_to != 0)
    gt_xfwrite(zeros, 1, pad_to - bytes % pad_to, fp);
  gt_fa_xfclose(fp);
  return 0;
}

int gt_suffixerator(int argc, const char **argv, GtError *err)
{
  Suffixeratoroptions so;
  GtEncseq *encseq = NULL;
  GtLogger *logger = NULL;
  B200Init init;
  pthread_t init_tid;
  bool init_started = false, init_ever = false, encoded = false, release_started = false;
  const char *load_from = NULL;
  pthread_t release_tid;
  int retval, had_err = 0;
  double t_start = b200_now(), t_encoded = 0, t_uploaded = 0, t_sorted = 0, t_copied = 0;

  gt_error_check(err);
  retval = gt_suffixeratoroptions(&so, true, argc, argv, err);
  if (retval != 0) {                     /* > 0: -help/-version were served */
    gt_sfxoptions_delete(&so);
    return retval < 0 ? -1 : 0;
  }
  logger = gt_logger_new(so.beverbose, GT_LOGGER_DEFLT_PREFIX, stdout);
  had_err = b200_unsupported(&so, err);
  if (!had_err && (gt_index_options_outsuftab_value(so.idxopts) || gt_index_options_outlcptab_value(so.idxopts) ||
                   gt_index_options_outbcktab_value(so.idxopts) || gt_index_options_outbwttab_value(so.idxopts))) {
    /* one code range per GPU, times -parts (at most 64 ranges; they are all resident at once: unlike
       the reference's parts loop, src/match/sfx-suffixer.c:1791-1838, -parts does not bound memory) */
    memset(&init, 0, sizeof init);
    init.want_devices = b200_gpus_wanted();
    init.parts = gt_index_options_numofparts_value(so.idxopts);
    init_ever = true;
    init_started = pthread_create(&init_tid, NULL, b200_init_thread, &init) == 0;
    if (!init_started) b200_init_thread(&init);
  }

  /* -ii: the sequences are encoded already, the index files are only loaded (src/match/sfx-run.c:454-492) */
  if (!had_err && so.inputindex != NULL && gt_str_length(so.inputindex) > 0) {
    encoded = true;
    load_from = gt_str_get(so.inputindex);
  }
  if (!had_err && !encoded) {
    int fast = b200_fast_encode(&so, logger, err);
    if (fast < 0) had_err = -1;
    else encoded = (fast == 0);
  }
  if (!had_err) {                        /* encode (unless done above) + load, exactly the reference's calls */
    GtEncseqEncoder *ee = gt_encseq_encoder_new_from_options(so.encopts, err);
    if (ee == NULL) had_err = -1;
    /* '-plain' implies no description support (src/match/sfx-run.c:498-502) */
    if (!had_err && gt_encseq_options_plain_value(so.loadopts)) {
      gt_encseq_encoder_do_not_create_des_tab(ee);
      gt_encseq_encoder_do_not_create_sds_tab(ee);
    }
    if (!had_err && !encoded) {
      gt_encseq_encoder_set_logger(ee, logger);
      if (gt_encseq_encoder_encode(ee, so.db, gt_str_get(so.indexname), err) != 0)
        had_err = -1;
    }
    gt_encseq_encoder_delete(ee);
    if (!had_err) {
      GtEncseqLoader *el = gt_encseq_loader_new_from_options(so.loadopts, err);
      gt_encseq_loader_disable_autosupport(el);
      gt_encseq_loader_do_not_require_des_tab(el);
      gt_encseq_loader_do_not_require_sds_tab(el);
      gt_encseq_loader_do_not_require_ssp_tab(el);
      encseq = gt_encseq_loader_load(el, load_from != NULL ? load_from : gt_str_get(so.indexname), err);
      gt_encseq_loader_delete(el);
      if (encseq == NULL) had_err = -1;
      else if (gt_encseq_is_mirrored(encseq)) {
        gt_error_set(err, "option -mirrored is not supported by the B200 suffixerator path "
                          "(no silent fallback); use the CPU build of gt for it");
        had_err = -1;
      }
    }
  }

  t_encoded = b200_now();
  if (!had_err) {
    const bool want_suf = gt_index_options_outsuftab_value(so.idxopts),
               want_lcp = gt_index_options_outlcptab_value(so.idxopts),
               want_bck = gt_index_options_outbcktab_value(so.idxopts),
               want_bwt = gt_index_options_outbwttab_value(so.idxopts);
    const bool want_any = want_suf || want_lcp || want_bck || want_bwt;
    const GtUword n = gt_encseq_total_length(encseq);
    const GtReadmode readmode = gt_index_options_readmode_value(so.idxopts);
    const unsigned int numofchars = gt_encseq_alphabetnumofchars(encseq);
    unsigned int prefixlength = gt_index_options_prefixlength_value(so.idxopts);
    uint64_t *suftab = NULL, *llv = NULL, nllv = 0, nall = 0, nspec = 0, ndist = 0;
    uint8_t *lcptab = NULL, *bwttab = NULL;
    uint32_t *leftborder = NULL, *csc = NULL, *dist = NULL;
    gtb_stats stats;
    B200Prefill prefill;
    char msg[512];

    memset(&stats, 0, sizeof stats);
    memset(&prefill, 0, sizeof prefill);
    /* the two -dir checks of gt_runsuffixerator, src/match/sfx-run.c:541-549,586-593 */
    if ((readmode == GT_READMODE_COMPL || readmode == GT_READMODE_REVCOMPL) &&
        !gt_alphabet_is_dna(gt_encseq_alphabet(encseq))) {
      gt_error_set(err, "option -%s only can be used for DNA alphabets",
                   readmode == GT_READMODE_COMPL ? "cpl" : "rcl");
      had_err = -1;
    }
    if (!had_err && !want_any && readmode != GT_READMODE_FORWARD) {
      gt_error_set(err, "option '-dir %s' only makes sense in combination with at least one of the "
                        "options -suf, -lcp, or -bwt", gt_readmode_show(readmode));
      had_err = -1;
    }
    if (!had_err && want_any) {
      /* detpfxlen, src/match/sfx-run.c:319-367 */
      unsigned int rec = gt_recommendedprefixlength(numofchars, n,
                                                    GT_RECOMMENDED_MULTIPLIER_DEFAULT, true);
      if (prefixlength == GT_PREFIXLENGTH_AUTOMATIC) {
        prefixlength = rec;
        gt_logger_log(logger, "automatically determined prefixlength=%u", prefixlength);
      } else {
        unsigned int maxpl = gt_whatisthemaximalprefixlength(numofchars, n, 0, true);
        if (gt_checkprefixlength(maxpl, prefixlength, err) != 0) had_err = -1;
      }
    }
    if (!had_err && (uint64_t) n + 1 >= (uint64_t) UINT32_MAX) {
      gt_error_set(err, "sequences of total length >= 2^32-2 are not supported by the B200 "
                        "suffixerator path");
      had_err = -1;
    }
    if (!had_err && want_any && (want_suf || want_lcp) &&
        !(getenv("GTB200_PREFILL") != NULL && getenv("GTB200_PREFILL")[0] == '0')) {
      memset(&prefill, 0, sizeof prefill);
      if (want_suf) {
        snprintf(prefill.path[0], sizeof prefill.path[0], "%s.suf", gt_str_get(so.indexname));
        prefill.bytes[0] = sizeof (uint64_t) * ((size_t) n + 1);
      }
      if (want_lcp) {
        snprintf(prefill.path[1], sizeof prefill.path[1], "%s.lcp", gt_str_get(so.indexname));
        prefill.bytes[1] = (size_t) n + 1;
      }
      prefill.started = pthread_create(&prefill.tid, NULL, b200_prefill_thread, &prefill) == 0;
    }
    if (!had_err && want_any) {
      int rc = 0;
      gtb_group *g = NULL;
      if (init_started) {                       /* created while the sequences were being encoded */
        pthread_join(init_tid, NULL);
        init_started = false;
      }
      g = init.group;
      init.group = NULL;
      if (g == NULL) { snprintf(msg, sizeof msg, "%s", init.msg); rc = -1; }
      else gt_logger_log(logger, "B200: %d GPU(s), %d bucket-code range(s) (gt -j %u, -parts %u)", init.ndevices,
                         init.nranges, gt_jobs, gt_index_options_numofparts_value(so.idxopts));
      /* GtReadmode values are the library's: fwd 0, rev 1, cpl 2, rcl 3 (src/core/readmode.h) */
      if (rc == 0) rc = gtb_group_set_readmode(g, (unsigned) readmode);
      if (rc == 0) rc = b200_group_set_encseq(g, encseq, want_bwt);
      t_uploaded = b200_now();
      if (rc == 0)
        rc = gtb_group_run(g, prefixlength, (want_suf ? GTB_WANT_SUF : 0u) | (want_lcp ? GTB_WANT_LCP : 0u) |
                                            (want_bck ? GTB_WANT_BCK : 0u));
      t_sorted = b200_now();
      if (rc == 0 && gtb_group_num_entries(g) != (uint64_t) n + 1) {
        snprintf(msg, sizeof msg, "internal: the code ranges hold %llu entries, expected %llu",
                 (unsigned long long) gtb_group_num_entries(g), (unsigned long long) n + 1);
        rc = -2;
      }
      if (rc == 0) {
        /* the result tables, sized from what the run produced */
        if (want_suf) suftab = b200_big_alloc(sizeof *suftab * (n + 1));
        if (want_lcp) {
          nllv = gtb_group_num_llv(g);
          lcptab = b200_big_alloc(sizeof *lcptab * (n + 1));
          llv = gt_malloc(sizeof *llv * 2 * (nllv + 1));
        }
        if (want_bck) {
          gtb_bck_sizes(numofchars, prefixlength, &nall, &nspec, &ndist);
          leftborder = gt_malloc(sizeof *leftborder * (nall + 1));
          csc = gt_malloc(sizeof *csc * (nspec + 1));
          dist = gt_malloc(sizeof *dist * (ndist + 1));
        }
        /* the gather: every range copies its shard to its offset of the one table, all tables in one call */
        rc = gtb_group_copy_results(g, suftab, lcptab, nllv > 0 ? llv : NULL, leftborder, csc, dist);
        if (rc == 0 && want_bwt) {
          bwttab = b200_big_alloc(sizeof *bwttab * (n + 1));
          rc = gtb_group_copy_bwttab(g, bwttab);
        }
        if (rc == 0) rc = gtb_group_get_stats(g, &stats);
      }
      if (rc == -1 && g != NULL) snprintf(msg, sizeof msg, "%s", gtb_group_error(g));
      gtb_group_delete(g);
      if (rc != 0) { gt_error_set(err, "libgtb200: %s", msg); had_err = -1; }
    }
    t_copied = b200_now();
    /* nothing is left on the GPUs: their contexts are destroyed beside the writing of the files
       (gtb_release_devices, include/gtb200.h) instead of behind the end of main */
    if (want_any && !had_err)
      release_started = pthread_create(&release_tid, NULL, b200_release_thread, NULL) == 0;
    /* the reference's files */
    b200_prefill_finish(&prefill);
    if (had_err) {                         /* no tables: the zero-filled files go */
      if (prefill.done[0]) (void) unlink(prefill.path[0]);
      if (prefill.done[1]) (void) unlink(prefill.path[1]);
    }
    if (!had_err && want_suf)
      had_err = b200_write(gt_str_get(so.indexname), ".suf", suftab, sizeof *suftab, n + 1, 0, prefill.done[0], err);
    if (!had_err && want_bwt)
      had_err = b200_write(gt_str_get(so.indexname), ".bwt", bwttab, 1, n + 1, 0, false, err);
    if (!had_err && want_lcp) {
      had_err = b200_write(gt_str_get(so.indexname), ".lcp", lcptab, 1, n + 1, 0, prefill.done[1], err);
      if (!had_err)
        had_err = b200_write(gt_str_get(so.indexname), ".llv", llv, sizeof *llv, 2 * nllv, 0, false, err);
    }
    if (!had_err && want_bck) {
      FILE *fp = gt_fa_fopen_with_suffix(gt_str_get(so.indexname), ".bck", "wb", err);
      if (fp == NULL) had_err = -1;
      else {
        b200_append_table(fp, leftborder, nall + 1);
        b200_append_table(fp, csc, nspec);
        b200_append_table(fp, dist, ndist);
        gt_fa_xfclose(fp);
      }
    }
    if (!had_err) {
      Definedunsignedlong longest;
      longest.defined = want_any;
      longest.valueunsignedlong = (GtUword) stats.longest;
      if (gt_outprjfile(gt_str_get(so.indexname), readmode, encseq,
                        want_any ? n + 1 : 0, prefixlength,
                        want_lcp ? (GtUword) stats.numoflargelcpvalues : 0,
                        want_lcp ? stats.lcptabsum / (double) (n + 1) : 0.0,
                        want_lcp ? (GtUword) stats.maxbranchdepth : 0,
                        &longest, err) != 0)
        had_err = -1;
    }
    gt_logger_log(logger, "B200 device time %.3f ms, %u kernel launches, %u radix passes",
                  stats.ms_total, stats.kernel_launches, stats.radix_passes);
    if (t_copied > 0)
      gt_logger_log(logger, "wall seconds: encode+load %.3f, export+upload (incl. waiting for the CUDA context) "
                            "%.3f, sort %.3f, copy to host %.3f, write files %.3f",
                    t_encoded - t_start, t_uploaded - t_encoded, t_sorted - t_uploaded, t_copied - t_sorted,
                    b200_now() - t_copied);
    b200_big_free(suftab, sizeof *suftab * (n + 1));
    b200_big_free(lcptab, sizeof *lcptab * (n + 1));
    b200_big_free(bwttab, sizeof *bwttab * (n + 1));
    gt_free(llv);
    gt_free(leftborder); gt_free(csc); gt_free(dist);
  }
  if (release_started) pthread_join(release_tid, NULL);
  if (init_started) pthread_join(init_tid, NULL);
  if (init_ever && init.group != NULL) gtb_group_delete(init.group);   /* an error came first: never used */
  gt_encseq_delete(encseq);
  gt_logger_delete(logger);
  gt_sfxoptions_delete(&so);
  return had_err ? -1 : 0;
}
