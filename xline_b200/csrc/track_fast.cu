// Fast variants: FMA contraction on, pack-time constant folding (XLB_STRICT 0).
#define XLB_STRICT 0
#ifndef XLB_BEAMFIELDS
#define XLB_BEAMFIELDS 0
#endif
// XLB_BEAMFIELDS: 0 = thin lenses only, 1 = + BeamBeam4D and space charge, 2 = + BeamBeam6D.
// The 6D lens is compiled into its own set of kernels: its register appetite makes ptxas
// spill in the dispatch loop of every kernel that contains it, whether or not a lattice has
// such a lens (ncu on the PS Booster: 6 % of all executed instructions were LDL/STL).
#if XLB_BEAMFIELDS == 2
#define XLB_NS fast_bf6
#elif XLB_BEAMFIELDS
#define XLB_NS fast_bf
#else
#define XLB_NS fast_lean
#endif
#include "track_impl.cuh"
#include "variants.inc"

namespace xlb {
using namespace XLB_NS;
XLB_DEF_TRACE_VARIANT()
#if XLB_BEAMFIELDS
// Same launch bounds as the lean kernels: the thin-lens records dominate even a beam-beam
// lattice (74 lenses among 5 500 records on C3), so the register budget is set by them and
// the rarely executed beam-field code is allowed to spill.
XLB_DEF_VARIANT(1, 256, 2)
XLB_DEF_VARIANT(2, 256, 2)
XLB_DEF_VARIANT(3, 128, 3)
#if XLB_BEAMFIELDS == 2
#define XLB_BF_SUFFIX "/beamfields6d"
#define XLB_BF_TABLE_FN fast_bf6_variants
#else
#define XLB_BF_SUFFIX "/beamfields"
#define XLB_BF_TABLE_FN fast_bf_variants
#endif
static const Variant fast_bf_table[] = {
    XLB_VARIANT_ENTRY("fast/ppt1/t256" XLB_BF_SUFFIX, 1, 256, 2),
    XLB_VARIANT_ENTRY("fast/ppt2/t256" XLB_BF_SUFFIX, 2, 256, 2),
    XLB_VARIANT_ENTRY("fast/ppt3/t128" XLB_BF_SUFFIX, 3, 128, 3),
    XLB_TRACE_ENTRY("fast/trace" XLB_BF_SUFFIX),
};
const Variant *XLB_BF_TABLE_FN(int *n) {
  *n = static_cast<int>(sizeof(fast_bf_table) / sizeof(fast_bf_table[0]));
  return fast_bf_table;
}
#if XLB_BEAMFIELDS == 2
void fast_bb6d_launch(const KArgs &a, const unsigned long long *rec, int blocks, int threads, void *stream) {
  bf::bb6d_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      a, reinterpret_cast<const double2 *>(rec));
}
#endif
#else
// resident CTAs per SM the launch bounds ask for (experiment builds override them)
#ifndef XLB_MINB_2_128
#define XLB_MINB_2_128 3
#endif
#ifndef XLB_MINB_3_128
#define XLB_MINB_3_128 3
#endif
XLB_DEF_VARIANT(1, 128, 5)
XLB_DEF_VARIANT(1, 256, 3)
XLB_DEF_VARIANT(1, 512, 2)
XLB_DEF_VARIANT(2, 128, XLB_MINB_2_128)
XLB_DEF_VARIANT(2, 256, 2)
XLB_DEF_VARIANT(3, 128, XLB_MINB_3_128)
XLB_DEF_VARIANT(4, 128, 2)
#if XLB_EXP_T160
XLB_DEF_VARIANT(3, 160, 3)
#endif

#if XLB_BEAMFIELDS
#define XLB_TABLE fast_bf_table
#define XLB_TABLE_FN fast_bf_variants
#define XLB_SUFFIX "/beamfields"
#else
#define XLB_TABLE fast_table
#define XLB_TABLE_FN fast_variants
#define XLB_SUFFIX "/lean"
#endif
static const Variant XLB_TABLE[] = {
    XLB_VARIANT_ENTRY("fast/ppt1/t128" XLB_SUFFIX, 1, 128, 5),
    XLB_VARIANT_ENTRY("fast/ppt1/t256" XLB_SUFFIX, 1, 256, 3),
    XLB_VARIANT_ENTRY("fast/ppt1/t512" XLB_SUFFIX, 1, 512, 2),
    XLB_VARIANT_ENTRY("fast/ppt2/t128" XLB_SUFFIX, 2, 128, XLB_MINB_2_128),
    XLB_VARIANT_ENTRY("fast/ppt2/t256" XLB_SUFFIX, 2, 256, 2),
    XLB_VARIANT_ENTRY("fast/ppt3/t128" XLB_SUFFIX, 3, 128, XLB_MINB_3_128),
    XLB_VARIANT_ENTRY("fast/ppt4/t128" XLB_SUFFIX, 4, 128, 2),
#if XLB_EXP_T160
    XLB_VARIANT_ENTRY("fast/ppt3/t160" XLB_SUFFIX, 3, 160, 3),
#endif
    XLB_TRACE_ENTRY("fast/trace"),
};
const Variant *XLB_TABLE_FN(int *n) {
  *n = static_cast<int>(sizeof(XLB_TABLE) / sizeof(XLB_TABLE[0]));
  return XLB_TABLE;
}
#endif
}  // namespace xlb
