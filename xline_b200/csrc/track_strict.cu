// Strict variants: compiled with -fmad=false; evaluates every map in the reference's
// operation order (XLB_STRICT 1) so results are IEEE-identical to the NumPy path wherever
// only + - * / sqrt are involved.  Parity instrument, not the performance path.
#define XLB_STRICT 1
#ifndef XLB_BEAMFIELDS
#define XLB_BEAMFIELDS 0
#endif
#if XLB_BEAMFIELDS
#define XLB_NS strict_bf
#else
#define XLB_NS strict_lean
#endif
#include "track_impl.cuh"
#include "variants.inc"

namespace xlb {
using namespace XLB_NS;
XLB_DEF_TRACE_VARIANT()
XLB_DEF_VARIANT(1, 256, 1)
XLB_DEF_VARIANT(2, 256, 1)

#if XLB_BEAMFIELDS
#define XLB_TABLE strict_bf_table
#define XLB_TABLE_FN strict_bf_variants
#define XLB_SUFFIX "/beamfields"
#else
#define XLB_TABLE strict_table
#define XLB_TABLE_FN strict_variants
#define XLB_SUFFIX "/lean"
#endif
static const Variant XLB_TABLE[] = {
    XLB_VARIANT_ENTRY("strict/ppt1" XLB_SUFFIX, 1, 256, 1),
    XLB_VARIANT_ENTRY("strict/ppt2" XLB_SUFFIX, 2, 256, 1),
    XLB_TRACE_ENTRY("strict/trace"),
};
const Variant *XLB_TABLE_FN(int *n) {
  *n = static_cast<int>(sizeof(XLB_TABLE) / sizeof(XLB_TABLE[0]));
  return XLB_TABLE;
}
}  // namespace xlb
