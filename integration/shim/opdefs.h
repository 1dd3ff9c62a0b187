/* Stand-in for the reference's GENERATED opdefs.h (ops.lisp:477 make-ops-file: one OP_<NAME> index per operator).
 * backends/cuda.c identifies operators by op->name, so no index is needed here. */
#ifndef MMB_SHIM_OPDEFS_H
#define MMB_SHIM_OPDEFS_H
#define NUM_OPS 0
#endif
