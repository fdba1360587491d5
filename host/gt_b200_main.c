/* host/gt_b200_main.c -- a minimal `gt`-like front end that exposes only the replaced
   tool:  gt_b200 suffixerator <options>.  In a real installation the object
   gt_suffixerator_b200.o is linked into the full `gt` binary instead. */
#include <stdio.h>
#include <string.h>
#include "core/init_api.h"
#include "core/error.h"
#include "core/error_api.h"
#include "core/thread_api.h"
#include <stdlib.h>
#include <sys/time.h>
#include <unistd.h>

int gt_suffixerator(int argc, const char **argv, GtError *err);
#ifdef B200_REFERENCE_TOOLS
/* the binary that carries gt_sfxiterator_b200.o: the reference's OWN tools, unchanged, on top of the
   B200 Sfxiterator -- `packedindex mkindex` (src/tools/gt_packedindex.c:33-36) and, as a cross-check of
   the iterator, the reference's suffixerator driver (src/match/sfx-run.c:719) */
#include "match/sfx-run.h"
#endif

/* GTB200_TRACE_WALL=1: wall-clock stamps (seconds since the epoch) at the start of main, around the tool and
   at the end of main, so that a caller that stamps before it starts the process and after it has waited for
   it can tell the time in front of main (loading) and behind it (the exit handlers) from the tool's own */
static void b200_stamp(const char *what)
{
  struct timeval tv;
  const char *e = getenv("GTB200_TRACE_WALL");
  if (e == NULL || e[0] != '1') return;
  gettimeofday(&tv, NULL);
  fprintf(stderr, "wallstamp %s %.6f\n", what, (double) tv.tv_sec + 1e-6 * (double) tv.tv_usec);
}

int main(int argc, char **argv)
{
  GtError *err;
  int rval;
  unsigned int jobs = 1;
  b200_stamp("main");
  /* `gt -j N <tool>`: the global option of gt (src/gtr.c:181); here N = number of GPUs */
  if (argc >= 4 && strcmp(argv[1], "-j") == 0) {
    int j = atoi(argv[2]);
    jobs = j > 0 ? (unsigned int) j : 1u;
    argv[2] = argv[0];
    argv += 2; argc -= 2;
  }
#ifdef B200_REFERENCE_TOOLS
  if (argc < 2 || (strcmp(argv[1], "suffixerator") != 0 && strcmp(argv[1], "packedindex_mkindex") != 0)) {
    fprintf(stderr, "usage: %s [-j N] suffixerator|packedindex_mkindex [options]\n", argv[0]);
    return 2;
  }
#else
  if (argc < 2 || strcmp(argv[1], "suffixerator") != 0) {
    fprintf(stderr, "usage: %s [-j N] suffixerator [options]\n", argv[0]);
    return 2;
  }
#endif
  gt_lib_init();
  b200_stamp("tool");
  gt_jobs = jobs;
  err = gt_error_new();
  gt_error_set_progname(err, argv[0]);
#ifdef B200_REFERENCE_TOOLS
  rval = gt_parseargsandcallsuffixerator(strcmp(argv[1], "suffixerator") == 0, argc - 1, (const char **) argv + 1, err);
#else
  rval = gt_suffixerator(argc - 1, (const char **) argv + 1, err);
#endif
  b200_stamp("tool_done");
  if (gt_error_is_set(err))
    fprintf(stderr, "%s: error: %s\n", argv[0], gt_error_get(err));
  gt_error_delete(err);
  if (gt_lib_clean()) return 3;
  b200_stamp("main_done");
  /* GTB200_QUICK_EXIT=1: leave without the exit handlers (the CUDA runtime unloads its modules and
     destroys the context there); every file of the tool is closed by now */
  if (getenv("GTB200_QUICK_EXIT") != NULL && getenv("GTB200_QUICK_EXIT")[0] == '1') {
    fflush(NULL);
    _exit(rval ? 1 : 0);
  }
  return rval ? 1 : 0;
}
