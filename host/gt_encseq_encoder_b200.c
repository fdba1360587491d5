/*
  host/gt_encseq_encoder_b200.c -- gt_encseq_encoder_encode (src/core/encseq_api.h:347, the one function every
  tool of GenomeTools that encodes sequences calls: `gt encseq encode`, `gt suffixerator`, `gt packedindex
  mkindex`, `gt tallymer`, ...) on top of gtb_fasta_encode of libgtb200.so.

  The function is defined in src/core/encseq.c (:8479-8505) next to the rest of the GtEncseq class, so it cannot
  be replaced by leaving an object out.  host/Makefile therefore renames the reference's definition in a COPY of
  the compiled object (objcopy --redefine-sym gt_encseq_encoder_encode=gt_encseq_encoder_encode_reference; no
  source is touched) and links this object in front: every caller binds to the function below, which

    * asks the encoder object what was requested (its public getters),
    * hands DNA / protein FASTA input to gtb_fasta_encode (all host cores, byte-identical index files),
    * and calls the reference's own function for everything else -- other alphabets, -lossless, -plain,
      inputs the library declines (it declines before it writes anything).

  Two settings of the encoder have no getter (lossless support, header-less .esq).  They are read from the
  object itself: its ten flags are declared below in the order of src/core/encseq.c:8141-8151, and before they
  are trusted the eight flags that DO have getters are compared with what the getters return -- an encoder
  object laid out differently fails that comparison and goes to the reference's function.  In a GenomeTools
  build a maintainer would instead put the call of gtb_fasta_encode at the top of gt_encseq_encoder_encode
  itself, where the fields are in scope.
  Written from scratch; no reference code is copied.
*/
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "core/alphabet.h"
#include "core/chardef.h"
#include "core/encseq.h"
#include "core/error_api.h"
#include "core/ma_api.h"
#include "core/str_api.h"
#include "core/str_array_api.h"
#include "gtb200.h"

int gt_encseq_encoder_encode_reference(GtEncseqEncoder *ee, GtStrArray *seqfiles, const char *indexname,
                                       GtError *err);

typedef struct {              /* the first members of struct GtEncseqEncoder, src/core/encseq.c:8141-8151 */
  bool destab, ssptab, sdstab, oistab, md5tab, isdna, isprotein, isplain, esq_no_header, clip_desc;
} B200EncoderFlags;

static bool b200_flags_are_readable(GtEncseqEncoder *ee, const B200EncoderFlags *f)
{
  return f->destab == gt_encseq_encoder_des_tab_requested(ee) &&
         f->ssptab == gt_encseq_encoder_ssp_tab_requested(ee) &&
         f->sdstab == gt_encseq_encoder_sds_tab_requested(ee) &&
         f->md5tab == gt_encseq_encoder_md5_tab_requested(ee) &&
         f->isdna == gt_encseq_encoder_is_input_dna(ee) &&
         f->isprotein == gt_encseq_encoder_is_input_protein(ee) &&
         f->isplain == gt_encseq_encoder_is_input_preencoded(ee) &&
         f->clip_desc == gt_encseq_encoder_are_descs_clipped(ee);
}

int gt_encseq_encoder_encode(GtEncseqEncoder *ee, GtStrArray *seqfiles, const char *indexname, GtError *err)
{
  const B200EncoderFlags *f = (const B200EncoderFlags *) ee;
  const char *which = getenv("GTB200_ENCODER");
  GtAlphabet *alpha = NULL;
  bool covered;

  covered = !(which != NULL && strcmp(which, "reference") == 0) &&
            gt_str_array_size(seqfiles) > 0 &&
            b200_flags_are_readable(ee, f) && !f->oistab && !f->esq_no_header && !f->isplain &&
            strlen(gt_encseq_encoder_symbolmap_file(ee)) == 0;
  if (covered) {
    if (f->isdna) alpha = gt_alphabet_new_dna();
    else if (f->isprotein) alpha = gt_alphabet_new_protein();
    else {                                   /* gt_encseq_new_from_files, src/core/encseq.c:7560-7569 */
      alpha = gt_alphabet_new_from_sequence(seqfiles, err);
      if (alpha == NULL) gt_error_unset(err);
    }
    covered = alpha != NULL && (gt_alphabet_is_dna(alpha) || gt_alphabet_is_protein(alpha));
  }
  if (covered) {
    gtb_fasta_request rq;
    gtb_fasta_summary sum;
    char decode[256], msg[1024];
    GtUword i, nfiles = gt_str_array_size(seqfiles);
    const char **names = gt_malloc(sizeof *names * nfiles);
    int rc;
    memset(decode, 0, sizeof decode);
    for (i = 0; i < (GtUword) gt_alphabet_num_of_chars(alpha); i++) decode[i] = gt_alphabet_decode(alpha, (GtUchar) i);
    decode[WILDCARD] = gt_alphabet_decode(alpha, (GtUchar) WILDCARD);
    for (i = 0; i < nfiles; i++) names[i] = gt_str_array_get(seqfiles, i);
    memset(&rq, 0, sizeof rq);
    rq.filenames = names;
    rq.numoffiles = nfiles;
    rq.indexname = indexname;
    rq.symbolmap = gt_alphabet_symbolmap(alpha);
    rq.decode = decode;
    rq.numofchars = gt_alphabet_num_of_chars(alpha);
    rq.alphatype = gt_alphabet_is_dna(alpha) ? 0u : 1u;
    rq.bits_per_symbol = gt_alphabet_bits_per_symbol(alpha);
    rq.out_des = f->destab;
    rq.out_sds = f->sdstab;
    rq.out_ssp = f->ssptab;
    rq.out_md5 = f->md5tab;
    rq.clip_desc = f->clip_desc;
    rq.sat = gt_str_get(gt_encseq_encoder_representation(ee));
    rc = gtb_fasta_encode(&rq, &sum, msg, sizeof msg);
    gt_free(names);
    gt_alphabet_delete(alpha);
    if (getenv("GTB200_TRACE_ENCODER") != NULL) {       /* tests: which encoder ran */
      if (rc == GTB_FASTA_OK)
        fprintf(stderr, "B200 encoder: %llu symbols, %s, %.3f s\n", (unsigned long long) sum.totallength, sum.satname,
                sum.seconds_total);
      else fprintf(stderr, "B200 encoder declines: %s\n", msg);
    }
    if (rc == GTB_FASTA_OK) return 0;
    if (rc != GTB_FASTA_UNSUPPORTED) { gt_error_set(err, "libgtb200: %s", msg); return -1; }
  } else gt_alphabet_delete(alpha);
  return gt_encseq_encoder_encode_reference(ee, seqfiles, indexname, err);
}
