// Placeholder tables used when the beam-field kernels are not compiled in.
#include "kargs.h"
namespace xlb {
#ifndef XLB_HAVE_BEAMFIELDS
const Variant *fast_bf_variants(int *n) { *n = 0; return nullptr; }
const Variant *strict_bf_variants(int *n) { *n = 0; return nullptr; }
#endif
}  // namespace xlb
