/*
  oracle/ref_driver.c -- TEST INFRASTRUCTURE, not product code.

  A small main() that links against objects compiled (by oracle/Makefile)
  directly from the UNMODIFIED reference sources under /root/reference and
  exposes the two reference entry points the parity tests need:

    gtref suffixerator <args>   -> gt_parseargsandcallsuffixerator
                                   (/root/reference/src/match/sfx-run.c:719)
    gtref sfxmap <args>         -> gt_sfxmap (brute-force ESA verifier,
                                   /root/reference/src/tools/gt_sfxmap.c)

  It replaces src/gt.c + the toolbox so that only the files on the
  suffixerator path have to be compiled.  No reference source is copied.
*/
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include "core/init_api.h"
#include "core/error_api.h"
#include "core/tool_api.h"
#include "match/sfx-run.h"

GtTool* gt_sfxmap(void);  /* /root/reference/src/tools/gt_sfxmap.c:1530 */

int main(int argc, char **argv)
{
  GtError *err;
  int rval = 1;

  if (argc < 2) {
    fprintf(stderr, "usage: %s suffixerator|sfxmap [options]\n", argv[0]);
    return 2;
  }
  gt_lib_init();
  err = gt_error_new();
  gt_error_set_progname(err, argv[0]);
  if (strcmp(argv[1], "suffixerator") == 0)
    rval = gt_parseargsandcallsuffixerator(true, argc - 1,
                                           (const char**) argv + 1, err);
#ifdef GTREF_WITH_SFXMAP
  else if (strcmp(argv[1], "sfxmap") == 0) {
    GtTool *tool = gt_sfxmap();
    rval = gt_tool_run(tool, argc - 1, (const char**) argv + 1, err);
    gt_tool_delete(tool);
  }
#endif
  else
    fprintf(stderr, "unknown subcommand %s\n", argv[1]);
  if (gt_error_is_set(err))
    fprintf(stderr, "%s: error: %s\n", argv[0], gt_error_get(err));
  gt_error_delete(err);
  if (gt_lib_clean())
    return 3;
  return rval ? 1 : 0;
}
