/*
  oracle/esa_oracle.h -- TEST INFRASTRUCTURE ONLY.

  CPU restatement of the result `gt suffixerator -suf -lcp -bck` defines
  ("rule R").  Only tests/, __graft_entry__.smoke() and
  bench.py's cpu_baseline / --impl reference legs may load this library, and
  only as the checker.  The product (genometools_b200/, libgtb200.so) never
  links, imports or calls anything in oracle/.

  Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement
  byte-for-byte against outputs of the unmodified reference (oracle/_ref/gtref,
  built by oracle/Makefile from /root/reference) committed under tests/golden/.
*/
#ifndef ESA_ORACLE_H
#define ESA_ORACLE_H
#include <stdint.h>

#define ESA_ORACLE_WILDCARD  254u   /* same byte values as GtUchar WILDCARD / */
#define ESA_ORACLE_SEPARATOR 255u   /* SEPARATOR, src/core/chardef.h          */

typedef struct {
  uint64_t totallength;        /* n                                             */
  uint64_t specialcharacters;  /* S                                             */
  uint64_t numofallcodes;      /* K^pl                                          */
  uint64_t numofspecialcodes;  /* K^(pl-1)                                      */
  uint64_t numofdistpfxidx;    /* sum_{i=1}^{pl-2} K^i                          */
  uint64_t longest;            /* SA index of suffix 0                          */
  uint64_t numoflargelcpvalues;/* #{j : lcp[j] >= 255}                          */
  uint64_t maxbranchdepth;     /* max lcp                                       */
  double   lcptabsum;          /* sum of lcp[j] over suffixes with >= pl regular
                                  leading symbols (sfx-lcpvalues.c:414)         */
} esa_oracle_stats;

/* sizes of the three .bck tables for (numofchars, prefixlength) */
void esa_oracle_bck_sizes(unsigned numofchars, unsigned prefixlength,
                          uint64_t *numofallcodes, uint64_t *numofspecialcodes,
                          uint64_t *numofdistpfxidx);

/*
  symbols[0..n-1]: 0..numofchars-1 regular, ESA_ORACLE_WILDCARD / _SEPARATOR special.
  Outputs (caller allocated):
    suf[n+1], lcp[n+1] (exact values, not clamped),
    leftborder[K^pl+1], countspecialcodes[K^(pl-1)], distpfxidx[sum K^i]
    (any of the three bck pointers may be NULL when prefixlength == 0).
  Returns 0, or -1 on bad arguments / allocation failure.
*/
int esa_oracle_build(const uint8_t *symbols, uint64_t n, unsigned numofchars,
                     unsigned prefixlength,
                     uint64_t *suf, uint64_t *lcp,
                     uint64_t *leftborder, uint64_t *countspecialcodes,
                     uint64_t *distpfxidx, esa_oracle_stats *stats);
#endif
