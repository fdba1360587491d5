/*
  oracle/esa_oracle.c -- TEST INFRASTRUCTURE ONLY (see esa_oracle.h).

  Plain-C restatement of what GenomeTools' suffixerator computes.  It follows
  the reference's *specification* functions, not its sorter:

  * total order + lcp: gt_encseq_check_comparetwosuffixes
      (/root/reference/src/core/encseq.c:7371-7460) with specialsareequal=false:
      regular symbols compare by alphabet rank; a special (wildcard, separator
      or the end of the text) is larger than every regular symbol, never equal
      to another special, and two specials compare by text position.
  * suftab layout: sorted suffixes that start with a regular symbol, then the
      special positions ascending, then n
      (/root/reference/src/match/sfx-suffixgetset.c:586-690).
  * lcptab: lcp[j] = number of equal leading regular symbols of suf[j-1], suf[j];
      0 for j = 0 and for the special tail
      (/root/reference/src/match/sfx-lcpvalues.c:371-471).
  * bucket table: leftborder / countspecialcodes / distpfxidx as defined by
      gt_updateleftborderforkmer / ...forspecialkmer
      (/root/reference/src/match/sfx-suffixer.c:1070-1102),
      gt_bcktab_updatespecials (/root/reference/src/match/bcktab.c:876-901) and
      gt_bcktab_leftborderpartialsums (bcktab.c:1274-1304).
  * stats for the .prj file: sfx-lcpvalues.c:371-433 (maxbranchdepth, large
      values, lcptabsum over the non-special part of every bucket),
      sfx-suffixgetset.c:228-254 (longest).

  Complexity is O(n log n * lcp); meant for inputs the tests sort in seconds.
*/
#include <stdlib.h>
#include <string.h>
#include "esa_oracle.h"

static const uint8_t *g_sym;   /* qsort context (single-threaded checker) */
static const uint64_t *g_run;  /* g_run[i] = #regular symbols from i to the next special */

static int is_special(uint8_t c) { return c >= ESA_ORACLE_WILDCARD; }

static uint64_t common_regular_prefix(uint64_t a, uint64_t b)
{
  uint64_t lim = g_run[a] < g_run[b] ? g_run[a] : g_run[b], l = 0;
  const uint8_t *pa = g_sym + a, *pb = g_sym + b;
  while (l < lim && pa[l] == pb[l]) l++;
  return l;
}

static int suffix_cmp(const void *va, const void *vb)
{
  uint64_t a = *(const uint64_t*) va, b = *(const uint64_t*) vb;
  uint64_t ra = g_run[a], rb = g_run[b], lim = ra < rb ? ra : rb;
  int c;
  if (a == b) return 0;
  c = memcmp(g_sym + a, g_sym + b, (size_t) lim);
  if (c != 0) return c;
  /* equal up to the first special met by either suffix */
  if (ra != rb) return ra < rb ? 1 : -1;   /* special > regular symbol        */
  return a < b ? -1 : 1;                   /* two specials: by text position  */
}

static uint64_t ipow(uint64_t b, unsigned e)
{
  uint64_t r = 1; while (e--) r *= b; return r;
}

void esa_oracle_bck_sizes(unsigned K, unsigned pl, uint64_t *nall,
                          uint64_t *nspecial, uint64_t *ndist)
{
  uint64_t d = 0; unsigned i;
  *nall = pl ? ipow(K, pl) : 0;
  *nspecial = pl ? ipow(K, pl - 1) : 0;
  for (i = 1; i + 2 <= pl; i++) d += ipow(K, i);
  *ndist = d;
}

int esa_oracle_build(const uint8_t *sym, uint64_t n, unsigned K, unsigned pl,
                     uint64_t *suf, uint64_t *lcp, uint64_t *leftborder,
                     uint64_t *countspecialcodes, uint64_t *distpfxidx,
                     esa_oracle_stats *st)
{
  uint64_t *run, i, j, nreg = 0, S = 0, nall, nspec, ndist;
  if (sym == NULL && n > 0) return -1;
  if (K < 1 || K > 253) return -1;
  run = malloc(sizeof *run * (n + 1));
  if (run == NULL) return -1;
  run[n] = 0;
  for (i = n; i-- > 0; ) {
    if (is_special(sym[i])) { run[i] = 0; S++; }
    else { if (sym[i] >= K) { free(run); return -1; } run[i] = run[i+1] + 1; }
  }
  /* suftab */
  for (i = 0; i < n; i++) if (run[i] > 0) suf[nreg++] = i;
  g_sym = sym; g_run = run;
  qsort(suf, (size_t) nreg, sizeof *suf, suffix_cmp);
  for (i = 0, j = nreg; i < n; i++) if (run[i] == 0) suf[j++] = i;
  suf[n] = n;
  /* lcptab */
  memset(lcp, 0, sizeof *lcp * (n + 1));
  for (j = 1; j < nreg; j++) lcp[j] = common_regular_prefix(suf[j-1], suf[j]);
  /* stats */
  memset(st, 0, sizeof *st);
  st->totallength = n; st->specialcharacters = S;
  for (j = 0; j <= n; j++) {
    if (suf[j] == 0) st->longest = j;
    if (lcp[j] > st->maxbranchdepth) st->maxbranchdepth = lcp[j];
    if (lcp[j] >= 255) st->numoflargelcpvalues++;
    if (j < nreg && run[suf[j]] >= pl) st->lcptabsum += (double) lcp[j];
  }
  /* bucket table */
  esa_oracle_bck_sizes(K, pl, &nall, &nspec, &ndist);
  st->numofallcodes = nall; st->numofspecialcodes = nspec;
  st->numofdistpfxidx = ndist;
  if (pl > 0 && leftborder != NULL) {
    memset(leftborder, 0, sizeof *leftborder * (nall + 1));
    if (countspecialcodes) memset(countspecialcodes, 0, sizeof *countspecialcodes * nspec);
    if (distpfxidx && ndist) memset(distpfxidx, 0, sizeof *distpfxidx * ndist);
    for (i = 0; i < n; i++) {
      uint64_t code = 0, lead = 0; unsigned k, u;
      if (run[i] == 0) continue;
      u = run[i] < pl ? (unsigned) run[i] : pl;
      for (k = 0; k < pl; k++) {
        code = code * K + (k < u ? sym[i+k] : K - 1);   /* fill with largest symbol */
        if (k < u) lead = lead * K + sym[i+k];
      }
      leftborder[code + 1]++;                /* counts, shifted by one */
      if (u < pl) {
        if (countspecialcodes) countspecialcodes[code / K]++;
        if (u + 1 < pl && distpfxidx) {      /* prefixindex < pl-1, bcktab.c:884 */
          uint64_t off = 0; unsigned l;
          for (l = 1; l < u; l++) off += ipow(K, l);
          distpfxidx[off + lead]++;
        }
      }
    }
    for (i = 0; i < nall; i++) leftborder[i+1] += leftborder[i];  /* bucket starts */
  }
  free(run);
  return 0;
}
