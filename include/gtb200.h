/*
  gtb200.h -- C-ABI of libgtb200.so, the B200-native enhanced-suffix-array
  constructor that replaces the sort core behind `gt suffixerator`.

  The reference (GenomeTools 1.5.11) has no FFI for this path; the seam is the
  Sfxiterator interface (src/match/sfx-suffixer.h:33-72) as driven by
  suffixeratorwithoutput() (src/match/sfx-run.c:212-317).  Every entry point
  below names the reference interface it stands in for.  Plain pointers and
  sizes only; no CUDA or torch types cross this boundary.

  Input conventions
  -----------------
  * 2-bit path: `twobitenc` is exactly what gt_encseq_twobitencoding_export()
    (src/core/encseq.c:6687) returns: uint64 words, 32 bases per word, base i in
    word i/32 at bits 62-2*(i%32) (src/core/intbits.h:78-83); positions that
    hold a wildcard/separator contain an arbitrary filler base.  `specials` are
    the maximal special runs as half-open ranges in ascending order, i.e. the
    sequence gt_specialrangeiterator_next() (src/core/encseq.h:127-133) yields
    in forward direction.
  * byte path (protein and every other non-2-bit alphabet): one byte per
    position as gt_encseq_extract_encoded() (src/core/encseq.h) delivers it:
    0..numofchars-1 regular, 254 = WILDCARD, 255 = SEPARATOR
    (src/core/chardef.h).

  Result conventions (identical to the reference's files)
  ---------------------------------------------------------------------------
  * suftab: n+1 entries: sorted suffixes starting at a regular symbol, then the
    special positions ascending, then n.            (.suf, uint64 each)
  * lcptab: n+1 bytes min(lcp,255); entries >= 255 additionally listed as
    {uint64 index; uint64 lcp} pairs in ascending index order (.lcp / .llv).
  * bucket table: leftborder[K^pl+1], countspecialcodes[K^(pl-1)],
    distpfxidx[sum_{i=1}^{pl-2} K^i], uint32 each (.bck; the caller pads every
    table to 8 bytes as gt_mapspec_write does, src/core/mapspec.c:350-365).

  All functions return 0 on success and -1 on error with a message in the
  handle (gtb_esa_error) or in errbuf.  There is no CPU fallback: if no CUDA
  device is usable every call fails.
*/
#ifndef GTB200_H
#define GTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GTB200_ABI_VERSION 3

typedef struct gtb_esa gtb_esa;              /* stands in for Sfxiterator   */

typedef struct { uint64_t start, end; } gtb_range;

/* what gt_Sfxiterator_longest, gt_Outlcpinfo_numoflargelcpvalues /
   _maxbranchdepth / _lcptabsum (sfx-run.c:300,676-680) report */
typedef struct {
  uint64_t totallength;          /* n                                           */
  uint64_t specialcharacters;    /* S                                           */
  uint64_t nonspecials;          /* suffixes sorted by this handle (shard)      */
  uint64_t sa_offset;            /* global SA index of the shard's first entry  */
  uint64_t longest;              /* SA index of suffix 0 (UINT64_MAX: not in shard) */
  uint64_t numoflargelcpvalues;  /* #{lcp >= 255}                               */
  uint64_t maxbranchdepth;       /* max lcp                                     */
  double   lcptabsum;            /* sum lcp over suffixes with >= pl regular
                                    leading symbols (averagelcp = sum/(n+1))    */
  uint32_t prefixlength;
  uint32_t numofchars;
  uint64_t unresolved_after_first_sort; /* elements that needed prefix doubling */
  uint32_t doubling_rounds;
  uint32_t radix_passes;         /* onesweep launches in the last run           */
  uint64_t radix_pairs_moved;    /* sum over passes of elements moved           */
  uint32_t kernel_launches;      /* all kernels launched by the last run        */
  /* device times of the last run in milliseconds (CUDA events on the handle's
     stream) */
  float ms_total, ms_upload, ms_count, ms_hist, ms_radix, ms_analyze,
        ms_doubling, ms_lcp, ms_tail;
  /* the first-level sort alone (the passes over all suffixes of the handle: the dominant kernel at
     its dominant size; ms_radix / radix_passes also count the small passes of the refinement rounds) */
  float    ms_radix_first;
  uint32_t radix_passes_first;
  uint64_t radix_pairs_first;    /* sum over those passes of elements moved */
} gtb_stats;

/* flags for gtb_esa_run (mirror the -suf -lcp -bck switches, index_options.c) */
#define GTB_WANT_SUF 1u
#define GTB_WANT_LCP 2u
#define GTB_WANT_BCK 4u
/* reuse the bucket table a preceding gtb_esa_count / gtb_esa_run on the same input
   and prefix length left in HBM (the -parts loop); without it every run recounts */
#define GTB_REUSE_COUNTS 8u

int  gtb_abi_version(void);
int  gtb_device_count(void);
/* Destroys the CUDA context of every device this process has used through the library and has no
   live handle on any more (cudaDeviceReset); returns how many.  For a tool that has its tables on
   the host and only files left to write: destroying a context takes the driver 0.3-1 s, which
   otherwise follows the end of main -- called on a thread of its own it passes beside the writes
   (host/gt_suffixerator_b200.c).  No counterpart in the reference. */
int  gtb_release_devices(void);

/* gt_Sfxiterator_new_withadditionalvalues (sfx-suffixer.c:1363): create the
   sorter object on CUDA device `device`. NULL on error (message in errbuf). */
gtb_esa *gtb_esa_new(int device, char *errbuf, size_t errlen);
/* gt_Sfxiterator_delete (sfx-suffixer.c:537-598) */
void gtb_esa_delete(gtb_esa *h);
const char *gtb_esa_error(const gtb_esa *h);

/* replicate the packed sequence to HBM (north_star: "keeps GtEncseq as the
   2-bit packed input, replicates it to HBM").  Host buffers; copies are done
   inside the call.  nwords = number of uint64 words readable at twobitenc
   (>= ceil(n/32)). */
int gtb_esa_set_input_2bit(gtb_esa *h, const uint64_t *twobitenc, uint64_t nwords,
                           uint64_t totallength, const gtb_range *specials,
                           uint64_t nspecialranges);
int gtb_esa_set_input_bytes(gtb_esa *h, const uint8_t *symbols,
                            uint64_t totallength, unsigned numofchars);

/* -dir fwd|rev|cpl|rcl (GtReadmode, src/core/readmode.h; the first argument of
   gt_Sfxiterator_new_withadditionalvalues, sfx-suffixer.h:39): the direction the sequence is
   read in.  0 forward, 1 reverse, 2 complement, 3 reverse complement; position i of the
   sorted text is then text[n-1-i] (1, 3), complemented (2, 3) -- what
   gt_encseq_get_encoded_char(encseq, i, readmode) returns (encseq.c:6094-6140).  Call it
   BEFORE gtb_esa_set_input_*: the inputs are still given in forward coordinates (as the
   GtEncseq exports them) and rewritten in read direction while they are uploaded;
   gtb_esa_set_separators takes forward coordinates as well.  Complement modes on the byte
   path fail like the reference ("only can be used for DNA alphabets", sfx-run.c:541-549). */
int gtb_esa_set_readmode(gtb_esa *h, unsigned readmode);

/* let `h` use the sequence `src` already holds in HBM (same device; `src` must outlive
   the use): several code ranges processed on one GPU share one copy of the input */
int gtb_esa_share_input(gtb_esa *h, const gtb_esa *src);

/* restrict the handle to the bucket codes [mincode, maxcode] (inclusive), the
   role of one part of gt_suftabparts_new (sfx-partssuf.c:172-347, filter at
   sfx-suffixer.c:375-376). sa_offset = leftborder[mincode] of the global
   table. The handle that owns the last code also emits the special tail.
   Default: all codes. */
int gtb_esa_set_code_range(gtb_esa *h, uint64_t mincode, uint64_t maxcode,
                           uint64_t sa_offset, int emit_special_tail);

/* gt_Sfxiterator_next over all parts + the GtOutlcpinfo side channel
   (sfx-lcpvalues.c:591-795): count codes, sort, prefix-double, lcp. Results
   stay in HBM until copied. */
int gtb_esa_run(gtb_esa *h, unsigned prefixlength, unsigned flags);

/* ---- staged execution: prefix doubling across code ranges (multi-GPU, -parts) ----
   A tie group lies inside one bucket, hence inside one code range, but a doubling
   round needs the rank of the suffix h positions further, which another range may
   own.  gtb_esa_run = sort_begin; while (unresolved) round_local; sort_end.
   With several ranges the caller runs the rounds in lock step on all handles and
   moves the requested positions / returned ranks between them (NCCL all-to-all,
   or device copies when the handles share a GPU):

     sort_begin on every range; if any range has unresolved suffixes:
       ensure_ranks on every range
       repeat until no range has unresolved suffixes:
         round_prepare  -> positions whose rank another range owns, grouped by range
         (exchange positions)   rank_lookup on the owner   (exchange ranks back)
         round_finish
     sort_end on every range

   The reference needs nothing of the kind because every CPU sorter compares text
   (gt_encseq_compare_viatwobitencoding, src/core/encseq.c:6719). */
int gtb_esa_sort_begin(gtb_esa *h, unsigned prefixlength, unsigned flags);
uint64_t gtb_esa_unresolved(const gtb_esa *h);
int gtb_esa_ensure_ranks(gtb_esa *h);
int gtb_esa_round_local(gtb_esa *h);
/* range_first_keys[r] = gtb_code_first_key(mincode of range r); counts_out[r] (host)
   receives how many positions of dev_send_positions (device, uint32) belong to range r;
   they are stored grouped by range in ascending range order. */
int gtb_esa_round_prepare(gtb_esa *h, const uint64_t *range_first_keys, int nranges, int my_range,
                          uint32_t *dev_send_positions, uint64_t send_capacity, uint64_t *counts_out);
int gtb_esa_rank_lookup(gtb_esa *h, const uint32_t *dev_positions, uint64_t count, uint32_t *dev_ranks);
/* dev_answers[i] = rank of dev_send_positions[i] */
int gtb_esa_round_finish(gtb_esa *h, const uint32_t *dev_answers);
int gtb_esa_sort_end(gtb_esa *h);
uint64_t gtb_code_first_key(unsigned numofchars, unsigned prefixlength, uint64_t code);

/* only the counting phase (updateleftborder_getencseqkmers_twobitencoding,
   sfx-suffixer.c:1132-1138 + gt_bcktab_leftborderpartialsums): fills the bucket
   table so that the caller can derive balanced code ranges
   (gt_suftabparts_new) before sorting. */
int gtb_esa_count(gtb_esa *h, unsigned prefixlength);

/* The same, split for several GPUs (north_star: "NCCL ... used only for the count
   allreduce and the result gather"): every device counts the k-mers of the text positions
   [first_pos, end_pos) only and leaves RAW counts in its three tables; the caller sums the
   tables of all devices in place (ncclAllReduce on the pointers of gtb_esa_dev_bcktab;
   uint32 sums, sizes from gtb_bck_sizes, leftborder has nall+1 entries) and then calls
   gtb_esa_count_finish (gt_bcktab_leftborderpartialsums, bcktab.c:1274-1304) on each. */
int gtb_esa_count_partial(gtb_esa *h, unsigned prefixlength, uint64_t first_pos, uint64_t end_pos);
int gtb_esa_count_finish(gtb_esa *h);
int gtb_esa_dev_bcktab(const gtb_esa *h, uint32_t **leftborder, uint32_t **countspecialcodes,
                       uint32_t **distpfxidx);

/* Code ranges for several GPUs WITHOUT a fine-grained counting pass (one atomic per suffix
   into a table larger than L2 is the slowest way to learn where to cut): the first
   plc <= prefixlength symbols of the filled keys (at most 4096 coarse codes) are counted
   in shared memory over the text positions [first_pos, end_pos); the caller sums the tiny
   table of all ranks in place (*dev_counts, *ncounts uint32 entries) and every rank cuts the
   same ranges with gtb_esa_coarse_split -- fine-grained mincode/maxcode, global offset and
   width of each part, whole coarse buckets per part.  A rank then takes its part with
   gtb_esa_set_code_range_known; its run derives the bucket-table entries of its own codes
   from its sorted keys (all other entries 0: the caller sums the tables of the ranks,
   gtb_esa_dev_bcktab, when the whole table is wanted). */
int gtb_esa_coarse_partial(gtb_esa *h, unsigned prefixlength, uint64_t first_pos, uint64_t end_pos,
                           uint32_t **dev_counts, uint64_t *ncounts);
int gtb_esa_coarse_split(gtb_esa *h, unsigned numofparts, uint64_t *out4, unsigned *nparts);
int gtb_esa_set_code_range_known(gtb_esa *h, uint64_t mincode, uint64_t maxcode, uint64_t sa_offset,
                                 uint64_t width, int emit_special_tail);

/* Sharding the text SCAN as well (every rank holds the whole packed sequence, but key
   generation is the expensive part of the first pass): rank r turns the text positions
   [first_pos, end_pos) of its slice into (filled key, position) pairs and groups them,
   stably, by the code range that owns the key (range_first_keys[g] = gtb_code_first_key of
   range g's first code) -- one onesweep pass whose "digit" is the owner.  dev_keys /
   dev_positions are caller-owned DEVICE buffers of `capacity` pairs; counts_out[g] = pairs
   for range g (they lie in that order).  The caller moves the groups to their owners
   (ncclAllToAll, or peer copies) so that every owner holds its pairs in text order -- slices
   in rank order -- and starts the sort there with gtb_esa_sort_begin_pairs instead of
   gtb_esa_sort_begin (same staged protocol afterwards).  The owner needs its code range:
   gtb_esa_set_code_range after gtb_esa_count* (+ GTB_REUSE_COUNTS), or
   gtb_esa_set_code_range_known after the coarse counts. */
int gtb_esa_slice_partition(gtb_esa *h, unsigned prefixlength, uint64_t first_pos, uint64_t end_pos,
                            const uint64_t *range_first_keys, int nranges, uint64_t *dev_keys,
                            uint32_t *dev_positions, uint64_t capacity, uint64_t *counts_out);
int gtb_esa_sort_begin_pairs(gtb_esa *h, unsigned prefixlength, unsigned flags, const uint64_t *dev_keys,
                             const uint32_t *dev_positions, uint64_t count);
/* the same with a third of the exchange volume: gtb_esa_slice_partition with dev_keys = NULL
   keeps the positions only (4 bytes per suffix cross the links); the owner regenerates the
   keys of its positions once (they arrive ascending per slice) */
int gtb_esa_sort_begin_positions(gtb_esa *h, unsigned prefixlength, unsigned flags,
                                 const uint32_t *dev_positions, uint64_t count);

/* gt_suftabparts_new (sfx-partssuf.c:172-347) on the bucket table in HBM: cut the codes
   into at most `numofparts` (<= 64) contiguous ranges of about equal suffix counts.
   out4[4*p .. 4*p+3] = mincode, maxcode, sa_offset, width of part p; *nparts = parts made
   (empty ones are dropped). out4 must hold 4*numofparts entries. */
int gtb_esa_split_ranges(gtb_esa *h, unsigned numofparts, uint64_t *out4, unsigned *nparts);

int gtb_esa_get_stats(const gtb_esa *h, gtb_stats *st);

/* filled keys of the first / last sorted suffix of this handle's code range, and
   the seam fix-up: lcp between the previous range's last suffix and this
   range's first one (computelocallcpvalue, sfx-lcpvalues.c:91-111); patches
   lcptab[0] and the stats of this handle. */
int gtb_esa_boundary_keys(const gtb_esa *h, uint64_t *first_key, uint64_t *last_key);
int gtb_esa_fix_seam(gtb_esa *h, uint64_t prev_last_key);

/* number of suftab / lcptab entries this handle produced (shard part incl.
   its share of the special tail) */
uint64_t gtb_esa_num_entries(const gtb_esa *h);
uint64_t gtb_esa_num_llv(const gtb_esa *h);

/* gt_suffixsortspace_to_file (sfx-suffixgetset.c:462-477): copy suftab entries
   [first, first+count) of this handle to host memory, widened to the file's
   uint64, or as the device-native uint32. */
int gtb_esa_copy_suftab_u64(gtb_esa *h, uint64_t *dst, uint64_t first, uint64_t count);
int gtb_esa_copy_suftab_u32(gtb_esa *h, uint32_t *dst, uint64_t first, uint64_t count);
/* outlcpvalues (sfx-lcpvalues.c:371-433) */
int gtb_esa_copy_lcptab(gtb_esa *h, uint8_t *dst, uint64_t first, uint64_t count);
/* both tables of the same entries in one call (either pointer may be NULL): with pinned buffers the lcp
   bytes cross the bus while host threads still widen the suffix table */
int gtb_esa_copy_tables(gtb_esa *h, uint64_t *suftab, uint8_t *lcptab, uint64_t first, uint64_t count);
/* all results of this handle in one call -- what suffixeratorwithoutput() (sfx-run.c:212-317) writes:
   with pinned buffers lcptab, llv (2*gtb_esa_num_llv() uint64) and the bucket table cross the bus on a
   second stream while the suffix table is copied and widened.  Any pointer may be NULL. */
int gtb_esa_copy_results(gtb_esa *h, uint64_t *suftab, uint8_t *lcptab, uint64_t *llv,
                         uint32_t *leftborder, uint32_t *countspecialcodes, uint32_t *distpfxidx);
/* gtb_esa_run and gtb_esa_copy_results in ONE call with the copy overlapped: the first-level order is
   final for all but the tied suffixes (1.5 % of a human-sized genome), so the suffix table starts to cross
   PCIe right after the first-level sort, on a stream of its own, while the analysis, the refinement rounds
   and the lcp kernels still run; the entries that left too early are patched on the host afterwards
   (their (index, position) pairs follow in one small copy).  What suffixeratorwithoutput() (sfx-run.c:212-317)
   does with the iterator, as one overlapped operation.  llv holds 2*llv_capacity uint64 (*nllv = pairs
   written; error if more are needed); any output pointer may be NULL. */
int gtb_esa_run_to_host(gtb_esa *h, unsigned prefixlength, unsigned flags, uint64_t *suftab, uint8_t *lcptab,
                        uint64_t *llv, uint64_t llv_capacity, uint64_t *nllv, uint32_t *leftborder,
                        uint32_t *countspecialcodes, uint32_t *distpfxidx);
/* -bwt (bwttab2file, src/match/sfx-run.c:173-210): one encoded symbol per suffix-table
   entry, the symbol before the suffix (0..numofchars-1, 254 wildcard, 255 separator;
   UNDEFBWTCHAR = 254 for the suffix that starts at 0, chardef.h:65).  The 2-bit input does
   not say which special positions are separators: pass their positions (ascending; for
   sequence i >= 1 that is gt_encseq_seqstartpos(encseq, i) - 1) before the run, otherwise
   every special is reported as a wildcard. */
int gtb_esa_set_separators(gtb_esa *h, const uint64_t *positions, uint64_t count);
int gtb_esa_copy_bwttab(gtb_esa *h, uint8_t *dst, uint64_t first, uint64_t count);
/* Largelcpvalue pairs (lcpoverflow.h:25-29): dst holds 2*gtb_esa_num_llv()
   uint64 {index, value}; index is global (sa_offset added). */
int gtb_esa_copy_llv(gtb_esa *h, uint64_t *dst);
/* gt_Sfxiterator_bcktab2file (sfx-suffixer.c:2206): the three tables of the
   whole text (not only the shard); any pointer may be NULL. Sizes from
   gtb_bck_sizes. */
int gtb_esa_copy_bcktab(gtb_esa *h, uint32_t *leftborder,
                        uint32_t *countspecialcodes, uint32_t *distpfxidx);
void gtb_bck_sizes(unsigned numofchars, unsigned prefixlength,
                   uint64_t *numofallcodes, uint64_t *numofspecialcodes,
                   uint64_t *numofdistpfxidx);

/* Order-dependent 64-bit checksums of the results AS THEY LIE IN HBM (no copy to the host):
     H = sum_i fin((i + 1) * C1 xor fin(v_i + C2)) mod 2^64,   i = global index of entry v_i
   (genometools_b200/mixhash.py states the same function over the reference's files).  The sum
   composes across code ranges: add the values of all shards.  out3[0]: suftab entries (index =
   global suffix-array index, value = position -- what .suf holds as uint64), out3[1]: lcptab bytes,
   out3[2]: the flat uint64 sequence of this shard's .llv pairs, the first pair being pair number
   llv_pairs_before of the whole file (= sum of gtb_esa_num_llv over the preceding ranges).
   gtb_esa_hash_bcktab: the uint32 words of the .bck file (gt_bcktab_flush_to_file, bcktab.c:519-577,
   every table padded to 8 bytes) from the three tables of this handle -- call it on a handle that
   holds the WHOLE table (single range, or after the tables of the ranges were summed).
   This is how a 3.1 Gbp run at 1/2/4/8 GPUs is compared with the reference's files without moving
   28 GB: tests/golden/config_md5.json holds the same checksums of the unmodified reference's output. */
int gtb_esa_hash_results(gtb_esa *h, uint64_t llv_pairs_before, uint64_t out3[3]);
int gtb_esa_hash_bcktab(gtb_esa *h, uint64_t *out);

/* ---- one job sharded over several GPUs (north_star: "contiguous .bck code
   ranges balanced by bucket counts are assigned to the 8 GPUs of one box, each holding the full
   replicated encseq") ------------------------------------------------------------------------
   Every code range -- one handle on one GPU -- executes gtb_esa_run_sharded with its rank: as a
   thread of one process (gtb_group below: what the C host does for `gt -j N`, src/gtr.c:181) or as a
   process of its own (bench.py under torchrun, separate_processes = 1: the buffers peers touch are
   virtual-memory-API allocations passed as file descriptors and mapped with 2 MB pages, csrc/gtb_vmm.cuh).  The ranges meet at a handful of all-gathers of small host structs, the one collective the
   caller provides (in-process: a shared buffer; across processes: NCCL).  The heavy exchanges go
   through peer memory inside the kernels: the partition pass of the position-sharded text scan
   stores every position straight into the owning range's HBM over NVLink, and a doubling round
   reads rank(p + h) of a foreign suffix from the owner's rank map in place.  Partitioning as
   gt_suftabparts_new (src/match/sfx-partssuf.c:172-347) on coarse buckets; seam lcp as
   computelocallcpvalue (src/match/sfx-lcpvalues.c:91-111).  Ranks beyond the number of parts that
   could be cut (tiny or one-bucket inputs) end with an empty result.  prefixlength >= 1.
   allgather(ctx, mine, bytes, all): all[r * bytes ..] = block of rank r, 0 on success. */
typedef int (*gtb_allgather_fn)(void *ctx, const void *mine, size_t bytes, void *all);
int gtb_esa_run_sharded(gtb_esa *h, unsigned prefixlength, unsigned flags, int rank, int world,
                        gtb_allgather_fn allgather, void *ctx, int separate_processes);
/* .llv pairs of the ranges before this one (the shard's place in the .llv file) */
uint64_t gtb_esa_llv_before(const gtb_esa *h);

/* The ranges of one job inside ONE process: nranges handles, handle i on CUDA device devices[i] (a
   device may be named several times: -parts on one GPU, sfx-suffixer.c:1791-1838; all ranges are
   resident at once).  The packed sequence is uploaded once per device.  gtb_group_run = the threads
   of the process run gtb_esa_run_sharded; gtb_group_copy_results is the result gather: every range
   copies its shard straight to its offset in the caller's tables, all GPUs at once (the bucket
   table is summed over the ranges through peer memory first). */
typedef struct gtb_group gtb_group;
gtb_group *gtb_group_new(const int *devices, int nranges, char *errbuf, size_t errlen);
void gtb_group_delete(gtb_group *g);
const char *gtb_group_error(const gtb_group *g);
int gtb_group_size(const gtb_group *g);
gtb_esa *gtb_group_range(gtb_group *g, int i);
int gtb_group_set_readmode(gtb_group *g, unsigned readmode);
int gtb_group_set_input_2bit(gtb_group *g, const uint64_t *twobitenc, uint64_t nwords, uint64_t totallength,
                             const gtb_range *specials, uint64_t nspecialranges);
int gtb_group_set_input_bytes(gtb_group *g, const uint8_t *symbols, uint64_t totallength, unsigned numofchars);
int gtb_group_set_separators(gtb_group *g, const uint64_t *positions, uint64_t count);
int gtb_group_run(gtb_group *g, unsigned prefixlength, unsigned flags);
int gtb_group_get_stats(const gtb_group *g, gtb_stats *st);   /* the job: sums / maxima over the ranges */
uint64_t gtb_group_num_entries(const gtb_group *g);
uint64_t gtb_group_num_llv(const gtb_group *g);
int gtb_group_copy_results(gtb_group *g, uint64_t *suftab, uint8_t *lcptab, uint64_t *llv,
                           uint32_t *leftborder, uint32_t *countspecialcodes, uint32_t *distpfxidx);
int gtb_group_copy_bwttab(gtb_group *g, uint8_t *dst);
/* out4 = checksums of suftab, lcptab, llv, bucket table of the whole job (gtb_esa_hash_results) */
int gtb_group_hash_results(gtb_group *g, uint64_t out4[4]);

/* the CUDA stream (cudaStream_t) all work of this handle is launched on, so that a caller
   can record its own timing events on it */
void *gtb_esa_stream(const gtb_esa *h);

/* device pointers of the results, for callers that exchange shards with NCCL
   (the pointers stay owned by the handle) */
const uint32_t *gtb_esa_dev_suftab(const gtb_esa *h);
const uint8_t  *gtb_esa_dev_lcptab(const gtb_esa *h);
const uint32_t *gtb_esa_dev_leftborder(const gtb_esa *h);

/* One-shot convenience with host buffers: what suffixeratorwithoutput()
   (sfx-run.c:212-317) gets from the iterator. Output pointers may be NULL.
   llv must hold 2*llv_capacity uint64; *nllv receives the count (error if it
   exceeds the capacity). */
int gtb_esa_build_2bit(int device, const uint64_t *twobitenc, uint64_t nwords,
                       uint64_t totallength, const gtb_range *specials,
                       uint64_t nspecialranges, unsigned prefixlength,
                       uint64_t *suftab, uint8_t *lcptab,
                       uint64_t *llv, uint64_t llv_capacity, uint64_t *nllv,
                       uint32_t *leftborder, uint32_t *countspecialcodes,
                       uint32_t *distpfxidx, gtb_stats *stats,
                       char *errbuf, size_t errlen);
int gtb_esa_build_bytes(int device, const uint8_t *symbols, uint64_t totallength,
                        unsigned numofchars, unsigned prefixlength,
                        uint64_t *suftab, uint8_t *lcptab,
                        uint64_t *llv, uint64_t llv_capacity, uint64_t *nllv,
                        uint32_t *leftborder, uint32_t *countspecialcodes,
                        uint32_t *distpfxidx, gtb_stats *stats,
                        char *errbuf, size_t errlen);

/* stand-alone access to the hand-written onesweep LSD radix sort (the engine of
   every sorting step; also the replacement for gt_radixsort_inplace_ulong /
   _GtUwordPair, src/core/radix_sort.h:91,107 -- SURVEY section 8f).  Host
   buffers, sorted in place by the bits [begin_bit, end_bit) of the key. */
int gtb_radixsort_pairs_u64_u32(int device, uint64_t *keys, uint32_t *values,
                                uint64_t count, unsigned begin_bit,
                                unsigned end_bit, char *errbuf, size_t errlen);

/* The record sorts of src/core/radix_sort.h on the same engine (host buffers, sorted in place;
   fewer than 2^32-1 records):
     gtb_radixsort_u64        gt_radixsort_inplace_ulong (radix_sort.h:91): plain 64-bit keys
     gtb_radixsort_u64pair    gt_radixsort_inplace_GtUwordPair (:107): records {a, b}, a is the key;
                              equal keys keep their input order (the reference's in-place MSD sort
                              leaves that order unspecified)
     gtb_radixsort_u64keypair gt_radixsort_inplace_Gtuint64keyPair (:125): records {a, b}, both
                              components are keys (a first) */
int gtb_radixsort_u64(int device, uint64_t *keys, uint64_t count, char *errbuf, size_t errlen);
int gtb_radixsort_u64pair(int device, uint64_t *pairs, uint64_t count, char *errbuf, size_t errlen);
int gtb_radixsort_u64keypair(int device, uint64_t *pairs, uint64_t count, char *errbuf, size_t errlen);

/* ---- FASTA -> GtEncseq index files -------------------------------
   gtb_fasta_encode writes <indexname>.esq and, as requested, .ssp .des .sds .md5 -- the files
   gt_encseq_encoder_encode (src/core/encseq.c:8479-8505 -> gt_encseq_new_from_files :7503-7714)
   writes for DNA or protein sequences in FASTA files, byte for byte -- with all host cores instead of the
   reference's one (genometools_b200/csrc/gtb_fasta.cpp; host code, no GPU involved: 64 MB of text
   are not worth a PCIe round trip, and the index files are written by the host anyway).  The caller
   hands over the alphabet as the reference holds it, so that the tables are not restated here:
     symbolmap  gt_alphabet_symbolmap (src/core/alphabet.h): 256 codes, 0..numofchars-1, 254 wildcard,
                253 undefined
     decode     gt_alphabet_decode of the codes 0..numofchars-1 and 254 (256 entries, the rest unused)
   Returns GTB_FASTA_OK; GTB_FASTA_UNSUPPORTED (msg says why; NOTHING was written) when the input is
   outside what this encoder covers -- the caller then runs gt_encseq_encoder_encode, which also words
   the reference's error messages: an alphabet read from a file, non-regular files (.gz / .bz2 files are inflated with
   the zlib / libbz2 found at run time, by one thread), a file that does not
   begin with '>', a -sat the reference refuses, a character outside the alphabet, an empty sequence, a description cut off by the end
   of the file or holding a NUL, 2^32-2 symbols or more; GTB_FASTA_ERROR for I/O errors. */
#define GTB_FASTA_OK 0
#define GTB_FASTA_UNSUPPORTED 1
#define GTB_FASTA_ERROR (-1)

typedef struct {
  const char *const *filenames;   /* as given to the tool: stored verbatim in the .esq header */
  uint64_t numoffiles;
  const char *indexname;
  const uint8_t *symbolmap;       /* 256 entries */
  const char *decode;             /* 256 entries */
  unsigned numofchars;            /* gt_alphabet_num_of_chars: 4 (DNA) or 20 (protein) */
  unsigned alphatype;             /* 0: the DNA alphabet, 1: the protein alphabet (gt_alphabet_is_dna / _is_protein);
                                     alphabets read from a file are unsupported */
  unsigned bits_per_symbol;       /* gt_alphabet_bits_per_symbol: protein 5 (unused for DNA) */
  int out_des, out_sds, out_ssp, out_md5;   /* -des -sds -ssp -md5 of the encseq options */
  int clip_desc;                  /* -clipdesc: descriptions end at their first white space */
  const char *sat;                /* -sat: NULL or "" = the smallest representation (the default), else one of
                                     direct bytecompress eqlen bit uchar ushort uint32 (what the reference refuses
                                     for the input is unsupported here: it then words the error) */
  int threads;                    /* 0: all cores, at most 32 */
} gtb_fasta_request;

typedef struct {
  uint64_t totallength, numofsequences, numoffiles;
  uint64_t specialcharacters, specialranges, realspecialranges;
  uint64_t wildcards, wildcardranges, realwildcardranges;
  uint64_t sat, satsep;           /* GtEncseqAccessType (src/core/encseq_access_type.h:24-34); 7 = none */
  uint64_t characterdistribution[32];
  uint64_t input_bytes;
  char satname[16];
  unsigned threads;
  double seconds_count, seconds_emit, seconds_lists, seconds_pack, seconds_md5, seconds_write,
         seconds_total;
} gtb_fasta_summary;

int gtb_fasta_encode(const gtb_fasta_request *request, gtb_fasta_summary *summary,
                     char *msg, size_t msglen);

#ifdef __cplusplus
}
#endif
#endif
