/* pb2_kat — per-function device known-answer hooks.  TEST INFRASTRUCTURE, not part of the drop-in boundary: the entry point
 * lives in its own library (libpb2_kat.so, built from pupiloptixlab_b200/csrc/test_hooks/kat.cu next to libpb2.so) so that the
 * product library carries no test code.  It runs the DEVICE restatement of one function of framework/render/material,
 * framework/render/emitter, framework/optix/util.h or framework/cuda/random.h over n inputs; `what` selects the function and
 * the array layouts are documented at the top of kat.cu. */
#ifndef PB2_KAT_H
#define PB2_KAT_H
#include "pb2.h"
#ifdef __cplusplus
extern "C" {
#endif
int pb2_kat(const char *what, const void *in0, const void *in1, const void *in2, uint64_t n, void *out);
const char *pb2_kat_last_error(void);
#ifdef __cplusplus
}
#endif
#endif
