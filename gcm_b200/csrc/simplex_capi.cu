// libgcm_b200.so -- the tetrahedral (simplex) half of the C ABI, a translation unit of its own so that it
// builds in parallel with the cubic one.
#include <algorithm>
#include <cstring>
#include <memory>

#include "capi_internal.cuh"
#include "thread_fns.h"

using namespace gcmb;

#include "simplex_capi.inc"
