/* nb_helper.h -- the resident helper of the drop-in executables (opt-in: NECAT_B200_HELPER=1).
 *
 * The reference's tools are short CPU processes; here every oc2pmov / oc2cns process first creates a CUDA context,
 * which on a B200 box costs 1 - 3 s -- several times the work itself (0.1 - 0.5 s per call on the E. coli-size set).
 * With NECAT_B200_HELPER=1 the executables hand their command line to a helper process (necat_b200/bin/nb_helperd,
 * started on demand, one per user / device / set of NB_* tunables) that keeps the context, its cached device blocks
 * and pinned buffers alive between the calls of one necat.pl run and leaves by itself after NECAT_B200_HELPER_IDLE
 * seconds (default 30) without a call.  Files, exit status and what is printed are those of the stand-alone run: the
 * helper runs the same code, in the caller's directory, writing to the caller's stdout / stderr.
 * Nothing changes for a caller that does not set the variable. */
#ifndef NB_HELPER_H
#define NB_HELPER_H

#include "nb_host.h"

enum { NBH_TOOL_PMOV = 0, NBH_TOOL_CNS = 1 };

/* client side (stand-alone executables): 1 = the helper ran the job, *status is its exit status; 0 = run it here */
int nbh_try_helper(int tool, int argc, char** argv, int* status);

/* the context of this process: created per call stand-alone, kept across calls inside the helper */
extern int nbh_in_helper;
int  nbh_ctx_get(int device, nb_ctx** out);
void nbh_ctx_put(nb_ctx* ctx);
/* the same in the background: the context is created (1 - 3 s) on a thread while the caller reads its volumes;
 * nbh_ctx_wait joins it.  One at a time. */
void nbh_ctx_begin(int device);
int  nbh_ctx_wait(nb_ctx** out);

int oc2pmov_main(int argc, char** argv);
int oc2cns_main(int argc, char** argv);

#endif
