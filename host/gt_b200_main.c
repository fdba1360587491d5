/* host/gt_b200_main.c -- a minimal `gt`-like front end that exposes only the replaced
   tool:  gt_b200 suffixerator <options>.  In a real installation the object
   gt_suffixerator_b200.o is linked into the full `gt` binary instead. */
#include <stdio.h>
#include <string.h>
#include "core/init_api.h"
#include "core/error.h"
#include "core/error_api.h"

int gt_suffixerator(int argc, const char **argv, GtError *err);

int main(int argc, char **argv)
{
  GtError *err;
  int rval;
  if (argc < 2 || strcmp(argv[1], "suffixerator") != 0) {
    fprintf(stderr, "usage: %s suffixerator [options]\n", argv[0]);
    return 2;
  }
  gt_lib_init();
  err = gt_error_new();
  gt_error_set_progname(err, argv[0]);
  rval = gt_suffixerator(argc - 1, (const char **) argv + 1, err);
  if (gt_error_is_set(err))
    fprintf(stderr, "%s: error: %s\n", argv[0], gt_error_get(err));
  gt_error_delete(err);
  if (gt_lib_clean()) return 3;
  return rval ? 1 : 0;
}
