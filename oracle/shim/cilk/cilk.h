/* oracle/shim/cilk/cilk.h -- TEST INFRASTRUCTURE ONLY.
 * No Cilk Plus / OpenCilk compiler exists in this image, so the reference's
 * -DCILK build is compiled as its *serial elision*: spawn/sync vanish and
 * cilk_for is a plain for loop.  Any number timed on this build is a 1-core
 * number and is labelled as such wherever it is reported. */
#ifndef DC_ORACLE_CILK_STUB_H
#define DC_ORACLE_CILK_STUB_H
#define cilk_spawn
#define cilk_sync
#define cilk_for for
#endif
