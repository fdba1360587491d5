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

int gt_suffixerator(int argc, const char **argv, GtError *err);

int main(int argc, char **argv)
{
  GtError *err;
  int rval;
  unsigned int jobs = 1;
  /* `gt -j N <tool>`: the global option of gt (src/gtr.c:181); here N = number of GPUs */
  if (argc >= 4 && strcmp(argv[1], "-j") == 0) {
    int j = atoi(argv[2]);
    jobs = j > 0 ? (unsigned int) j : 1u;
    argv[2] = argv[0];
    argv += 2; argc -= 2;
  }
  if (argc < 2 || strcmp(argv[1], "suffixerator") != 0) {
    fprintf(stderr, "usage: %s [-j N] suffixerator [options]\n", argv[0]);
    return 2;
  }
  gt_lib_init();
  gt_jobs = jobs;
  err = gt_error_new();
  gt_error_set_progname(err, argv[0]);
  rval = gt_suffixerator(argc - 1, (const char **) argv + 1, err);
  if (gt_error_is_set(err))
    fprintf(stderr, "%s: error: %s\n", argv[0], gt_error_get(err));
  gt_error_delete(err);
  if (gt_lib_clean()) return 3;
  return rval ? 1 : 0;
}
