/*
 * xline_b200 -- C ABI of the B200-native `Line.track` particle push.
 *
 * This is the drop-in boundary for the one hot path this library replaces: the
 * reference's element-by-element push
 *
 *     Line.track(p):  for el in self.elements: el.track(p)     xline/line.py:89-95
 *
 * over the element maps in xline/elements.py and xline/be_beamfields/*.py, acting on a
 * particle container with the attributes of xpart's Pyparticles (xline/particles.py:1-6).
 * The reference is pure Python and has no FFI; the entry points below are what a
 * ctypes/cffi binding added to the reference's `Line.track` would call (the stub is in
 * INTEGRATION.md).  Plain pointers and sizes only -- no torch types cross this boundary.
 *
 * Conventions
 *   - every function returns 0 on success, a negative XLB_E* code on failure;
 *     xlb_last_error() returns a thread-local, NUL-terminated description.
 *   - "device" entry points take CUDA device pointers and a cudaStream_t (as void*),
 *     enqueue work on that stream and (unless stated) do not synchronise the device;
 *     "host" entry points take host pointers, perform the H2D/D2H copies themselves and
 *     return when the results are in the host buffers.
 *   - particle state is structure-of-arrays, fp64 / int64, length n, caller-owned.
 *     Lost particles are NOT removed from the arrays (the reference compacts them,
 *     xline/elements.py:420): they keep the coordinates they had at the aperture with
 *     state = 0, at_element = index of the aperture in the Line, at_turn = turn of loss.
 */
#ifndef XLINE_B200_H
#define XLINE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XLB_ABI_VERSION 4

/* error codes */
#define XLB_OK 0
#define XLB_EINVAL (-1)    /* bad argument (null pointer, negative size, bad lattice)  */
#define XLB_ECUDA (-2)     /* CUDA runtime error (message in xlb_last_error)            */
#define XLB_ELATTICE (-3)  /* packed lattice failed validation                          */
#define XLB_ENOGPU (-4)    /* no CUDA device / not an sm_100 class device               */

/* ---------------------------------------------------------------------------------
 * Packed lattice ("flat device lattice buffer": element type tags + fp64 parameters).
 *
 * A lattice is an array of 8-byte words, cut into `n_chunks` chunks of `chunk_words`
 * words (chunk_words even, chunk = unit of one TMA bulk copy into shared memory).  A
 * chunk is a sequence of records; a record starts on an even word (16 B aligned):
 *
 *   word 0  header  = tag (8 bits) | aux << 8 (8 bits) | size << 16 (14 bits, record length
 *                     in 16-byte pairs) | XLB_HDR_* flags (bits 30, 31) |
 *                     (uint64)element_index << 32
 *   word 1  first fp64 parameter (or padding)
 *   word 2..        further parameters, record padded to an even number of words
 *
 * Records never straddle a chunk; every chunk ends with an XLB_T_END_CHUNK record and
 * the last chunk ends with XLB_T_END_TURN instead.  Parameter layout per tag is
 * documented next to the tag.
 *
 * Path length.  `s` advances by the drift lengths only (elements.py:56,72), by the same
 * amount for every particle that stays in the beam.  The fast kernels therefore do no
 * arithmetic on it in the element maps: word 1 of an END_TURN record is the length of the
 * pass it closes (the whole lattice, or one segment of a segmented lattice), added once per
 * pass; and every record at which a particle can be lost (the three LIMIT_* tags and the
 * block records) carries `s_here`, the drift lengths from the start of the pass up to the
 * record, added for the particles it removes.  The strict kernels ignore both and keep the
 * reference's per-particle sequential sum.  `element_index` is the position in the reference's
 * `Line.elements` list (what `at_element` reports).  Two encodings exist, selected by
 * XLB_F_STRICT: "fast" (constants pre-folded for FMA-friendly evaluation) and "strict"
 * (raw reference parameters, evaluated in the reference's operation order).
 * --------------------------------------------------------------------------------- */
enum xlb_tag {
  XLB_T_END_TURN = 0,       /* [hdr,length of the pass that ends here]                */
  XLB_T_END_CHUNK = 1,      /* no parameters                                          */
  XLB_T_DRIFT = 2,          /* [hdr,length]                      elements.py:48-56    */
  XLB_T_DRIFT_EXACT = 3,    /* [hdr,length]                      elements.py:64-72    */
  XLB_T_MULTIPOLE = 4,      /* aux=order; [hdr,0] then (kn_i,ks_i) i=order..0
                               fast: kn_i = knl[i]/i!            elements.py:120-156  */
  XLB_T_MULTIPOLE_CURVED = 5, /* aux=order; [hdr,hxl][hyl,length][inv_length,0] then
                               pairs as above                    elements.py:139-154  */
  XLB_T_CAVITY = 6,         /* [hdr,voltage][k=2*pi*f/c, lag_rad] elements.py:239-245 */
  XLB_T_RFMULTIPOLE = 7,    /* aux=order; [hdr,voltage][k,lag_rad] then per order
                               [knl_i,ksl_i][pn_i_rad,ps_i_rad]  elements.py:182-227  */
  XLB_T_XYSHIFT = 8,        /* [hdr,dx][dy,0]                    elements.py:274-276  */
  XLB_T_SROTATION = 9,      /* [hdr,cos][sin,0]                  elements.py:379-390  */
  XLB_T_DIPOLE_EDGE = 10,   /* [hdr,r21][r43,0]                  elements.py:538-548  */
  /* The three LIMIT_* records may carry the drift that follows the aperture (pack-time
     peephole Limit* -> Drift): aux bit 4 (XLB_AUX_DRIFT), bit 5 = it is a DriftExact; the drift
     length follows the record's own pairs as [length,0].                                     */
  XLB_T_LIMIT_RECT = 11,    /* aux bit 0: symmetric box (fast encoding only);
                               [hdr,min_x][max_x,min_y][max_y,s_here] elements.py:401-420 */
  XLB_T_LIMIT_ELLIPSE = 12, /* [hdr,a*a][b*b,1/(a*a)][1/(b*b),s_here] elements.py:429-442 */
  XLB_T_LIMIT_RECT_ELLIPSE = 13, /* [hdr,max_x][max_y,a*a][b*b,1/(a*a)][1/(b*b),s_here]
                                                                 elements.py:453-474  */
  XLB_T_MONITOR = 14,       /* [hdr,0][i64 start,i64 skip][i64 num_stores,i64 min_id]
                               [i64 max_id,i64 rolling][i64 data_offset,0]
                                                                 elements.py:485-527  */
  XLB_T_SAWTOOTH_CAVITY = 15, /* as CAVITY                       elements.py:257-263  */
  XLB_T_BEAMBEAM4D = 16,    /* see xline_b200/lattice.py         beambeam.py:45-82    */
  XLB_T_SPACECHARGE = 17,   /* aux bits 0-3 = profile kind (0 coasting,1 q-Gaussian,2 linear
                               interp,3 cubic spline); aux bit 4/5: the kick is followed by
                               a Drift / DriftExact of length word 1  spacecharge.py  */
  XLB_T_BEAMBEAM6D = 18,    /* [hdr,(double)n_slices] ...        BB6D.py:15-155       */
  XLB_T__COUNT = 19,
  /* Fused thin multipole -> [aperture] -> [drift] records (pack-time peephole).  The tag
     has bit 7 set and describes the block: bits 0-1 aperture kind (0 none, 1 symmetric
     rect, 2 rect, 3 ellipse), bit 2 curved multipole, bit 3 drift present, bit 4 that drift
     is a DriftExact.  aux=order;
     [hdr,drift_length][i64 aperture_element_index,s_here] (kn_i,ks_i) i=order..0
     [hxl,hyl][length,1/length] if curved (header bit XLB_HDR_HX_ONLY: hyl == 0, fast encoding)
     [min_x,max_x][min_y,max_y] (rect) or [a*a,b*b][1/(a*a),1/(b*b)] (ellipse)            */
  XLB_T_THIN_BLOCK = 0x80,
  /* Merged block (fast encoding only): two co-located thin multipoles K1, K2 evaluated as ONE
     multipole with summed coefficients -- [K1][A1][K2][A2][drift] of the Line, where thin kicks
     leave x, y untouched so both aperture tests are unaffected.  Tag = 0xA0 | the thin-block
     bits (A2 kind, K2 curved, drift, exact).  aux = merged order;
     [hdr,drift_length][i64 a1_index | a2_index << 32, i64 k1_order | has_a1 << 8]
     merged (kn_i,ks_i) i=order..0  [hxl,hyl][length,1/length][knl0,ksl0 of K2] if curved
     [a*a,b*b][1/(a*a),1/(b*b)] if has_a1 (A1 is an ellipse)  A2 limits as above
     K1's own (kn_i,ks_i) i=k1_order..0 (re-evaluated only for particles lost at A1)
     [s_here,0]                                                                            */
  XLB_T_MERGED_BLOCK = 0xA0,
  /* Dipole edge followed by a drift: tag = 0xC0 | drift << 3 | exact << 4;
     [hdr,drift_length][r21,r43]                              elements.py:538-548, 48-72   */
  XLB_T_EDGE_BLOCK = 0xC0
};

/* Header bit 31, fast encoding, curved block records only: the multipole's hyl is exactly 0 (a
   horizontal bend).  The kernel then leaves the hyl terms of elements.py:139-154 out -- exact
   zeros for every finite y -- which saves 6 of the 14 FP64 instructions of the curved kick.
   Optional: a packer that never sets it gets the general formula.                          */
#define XLB_HDR_HX_ONLY 0x80000000u
/* aux bits of the LIMIT_* and SPACECHARGE records: a drift fused into the record              */
#define XLB_AUX_DRIFT 0x10
#define XLB_AUX_DRIFT_EXACT 0x20
/* Header bit 30, merged block records: the block has an aperture A1 between its two kicks (the
   same fact as bit 8 of the record's second i64; the header copy is warp-uniform in the kernel).
   Mandatory on merged blocks with an A1.                                                    */
#define XLB_HDR_HAS_A1 0x40000000u

typedef struct xlb_lattice {
  const uint64_t *words; /* n_chunks * chunk_words words                                */
  int64_t n_words;
  int32_t chunk_words;   /* even, 16 <= chunk_words*8 <= 96 KiB                         */
  int32_t n_chunks;
  int32_t n_elements;    /* length of the reference Line (size of loss tallies)         */
  uint32_t flags;        /* XLB_F_* below                                               */
  /* Segmented lattice (fast encoding of lattices with BeamBeam6D lenses): the chunks form
     n_segments consecutive groups, each terminated by END_TURN, that are executed one after
     the other every turn -- XLB_SEG_MAIN groups by the tracking kernel, XLB_SEG_BB6D groups
     (one chunk holding one BEAMBEAM6D record) by the stand-alone 6D-lens kernel.  The last
     group is a MAIN group (possibly empty); it is the one that counts the turn.  `segments`
     is a HOST array of n_segments triples {first_chunk, n_chunks, kind}, also when `words`
     is device memory.  n_segments == 0: the whole lattice is one MAIN group.            */
  int32_t n_segments;
  const int32_t *segments;
} xlb_lattice_t;

#define XLB_SEG_MAIN 0
#define XLB_SEG_BB6D 1

#define XLB_F_STRICT 1u      /* lattice encoded for / kernel evaluates in the reference's
                                operation order without FMA contraction                 */
#define XLB_F_BEAMFIELDS 2u  /* the MAIN groups contain BEAMBEAM4D/6D or SPACECHARGE records */
#define XLB_F_BB6D 4u        /* lattice contains BEAMBEAM6D records: in MAIN groups (then
                                BEAMFIELDS is set too and the kernels that carry the 6D
                                lens are used) or in XLB_SEG_BB6D groups                */
#define XLB_F_LOW_ORDER 8u   /* every MULTIPOLE / block record has order (aux) <= 3: the
                                kernels with a straight-line Horner evaluation may be used
                                (checked by xlb_lattice_validate; optional, never required) */

/* Particle set, mirrors the attributes the reference's elements read and write
 * (SURVEY.md §8a row a2).  All arrays have length n.  chi and charge_ratio may be NULL
 * (treated as 1.0).  chi == NULL is the fast case: one species, the reference's default -- the
 * kernels compiled without a chi register are used (4 particles per thread instead of 3);
 * xlb_track_host passes NULL on by itself when the host chi column is all ones.
 * s, particle_id, at_element, at_turn may NOT be NULL. */
typedef struct xlb_particles {
  int64_t n;
  double *x, *px, *y, *py, *zeta, *delta, *rpp, *rvv, *s;
  const double *chi, *charge_ratio;
  int64_t *state, *at_element, *at_turn;
  const int64_t *particle_id;
  double q0, mass0, p0c, beta0, gamma0, energy0; /* reference particle                  */
} xlb_particles_t;

typedef struct xlb_track_options {
  int32_t num_turns;         /* >= 0                                                     */
  int32_t particles_per_thread; /* 0 = default (4 for thin-lens lattices with chi == NULL, 3
                                   with a chi column or strict, 2 for beam-field lattices;
                                   fewer when the beam would not fill the GPU); 1..4       */
  int32_t threads_per_block;    /* 0 = default (128; 256 for beam-field lattices at 1-2
                                   particles/thread); multiple of 32, <= the variant's
                                   launch bound                                            */
  int32_t turns_per_launch;  /* > 0: survivors are re-compacted (warp-ballot stream
                                compaction) between launches of this many turns; 0 =
                                automatic (one launch up to 150 turns, else launches of 100
                                turns); < 0 = all turns in one launch                      */
  int64_t *loss_tally;       /* optional [n_elements] int64 counters, incremented per
                                element where a particle was lost (same memory space as
                                the particle arrays)                                     */
  double *monitor_data;      /* optional BeamMonitor storage (fp64 words), see
                                XLB_T_MONITOR; same memory space as the particles        */
  int64_t monitor_words;     /* capacity of monitor_data in fp64 words                   */
  double compact_threshold;  /* re-compact when lost/active exceeds this (default 1/128) */
  int32_t turns_per_item;    /* granularity of the device-side work queue: a launch of more
                                turns than this, over more particle blocks than there are SMs,
                                runs persistent CTAs that pull (particle block, turn segment)
                                items.  0 = automatic (one turn, or as many as stream 16
                                lattice chunks), < 0 = off                                */
  int32_t flags;             /* XLB_OPT_* below                                          */
  double *trace;             /* optional element-by-element trace, [n_elements][6][trace_particles]
                                fp64 (x px y py zeta delta after every element for the first
                                trace_particles particle slots; the device form of
                                Line.track_elem_by_elem, xline/line.py:97-108).  Needs
                                num_turns == 1 and a lattice packed without record fusing     */
  int64_t trace_particles;
  int64_t element_index_offset; /* added to the element index recorded in at_element (not to the
                                loss-tally index): a caller that pushes the elements of a Line
                                one call at a time -- el.track(p) in a loop, the reference's
                                track_elem_by_elem (xline/line.py:97-108) -- passes the position
                                of the element in that Line                                 */
} xlb_track_options_t;

#define XLB_OPT_NO_TURN_COUNT 1 /* do not advance at_turn at the end of the lattice: the call is
                                   one element (or a part) of a turn, like the reference's
                                   el.track(p), which never touches the turn counter          */

/* Statistics of the last xlb_track_* call on this thread. */
typedef struct xlb_track_stats {
  int64_t n_alive_in, n_alive_out;
  int32_t kernel_launches;   /* tracking-kernel launches                                 */
  int32_t compactions;       /* compaction-kernel launches                               */
  int32_t regs_per_thread, smem_bytes, blocks, threads; /* last tracking launch          */
  float kernel_ms;           /* CUDA-event time of the tracking kernels (host entry point
                                and xlb_track_device_timed only)                         */
} xlb_track_stats_t;

int xlb_abi_version(void);
const char *xlb_last_error(void);

/* Validates a packed lattice held in HOST memory (record structure, tags, chunking). */
int xlb_lattice_validate(const xlb_lattice_t *host_lattice);

/* The hot path, device pointers.  Replaces the loop of xline/line.py:89-95 repeated
 * opts->num_turns times.  lattice->words, all particle arrays, opts->loss_tally and
 * opts->monitor_data are DEVICE pointers.  `stream` is a cudaStream_t.  Synchronises the
 * stream only when turns_per_launch splits the job (to read the survivor count). */
int xlb_track_device(const xlb_lattice_t *lattice, xlb_particles_t *particles,
                     const xlb_track_options_t *opts, void *stream);

/* Same, but brackets the tracking kernels with CUDA events on `stream`, synchronises and
 * reports kernel_ms through xlb_get_stats(). */
int xlb_track_device_timed(const xlb_lattice_t *lattice, xlb_particles_t *particles,
                           const xlb_track_options_t *opts, void *stream);

/* The hot path, HOST pointers (what a NumPy-holding caller such as the reference's own
 * Line.track would bind): copies lattice + particle columns to the current device, tracks,
 * copies the particle columns (and tallies / monitor data) back.  Blocking. */
int xlb_track_host(const xlb_lattice_t *host_lattice, xlb_particles_t *host_particles,
                   const xlb_track_options_t *opts);

int xlb_get_stats(xlb_track_stats_t *out);

/* Stream compaction of survivors: writes the indices i with state[i]==1, in increasing
 * order, to idx_out (device, capacity n) and their count to *n_out (device).  Warp-ballot
 * prefix scan; deterministic.  Used internally between launches and exposed for the host
 * container's remove_lost_particles(). */
int xlb_compact_alive_device(const int64_t *state, int64_t n, int32_t *idx_out,
                             int32_t *n_out, void *stream);

/* Register-resident independent-DFMA-chain microbenchmark: the measured FP64 pipe peak of
 * the current device in FLOP/s (2 flops per DFMA), best of `repeats`.  The roofline
 * denominator of bench.py (MEASURED_PEAKS.json has no FP64 entry). */
int xlb_measure_fp64_peak(int repeats, double *flops_out, double *ms_out);

/* Single-warp DFMA timing: cycles per DFMA with 1, 2, 4 and 8 independent dependent chains
 * (cycles_per_dfma[0..3]); the first is the dependent-issue latency of the FP64 pipe. */
int xlb_measure_dfma_latency(double *cycles_per_dfma, int n);

/* Device self-test of the strict kernels' division sequences (csrc/track_impl.cuh: reciprocal
 * from the record or the constant table + FMA remainder corrections in place of the IEEE
 * division subroutine; the reference divides at xline/elements.py:130-134, 143-144, 436):
 * every quotient is compared bit for bit with the device's own a / b.  mode 0: integer divisors
 * 1..255 (the Horner step), mode 1: arbitrary positive divisors (`length`, a*a, b*b).  Dividends:
 * random sign and mantissa, exponents uniform in 2^+-exponent_span, every 16th next to a
 * rounding midpoint of the quotient.  *mismatches must come back 0. */
int xlb_selftest_exact_division(const double *divisors, int n_divisors, int mode,
                                int samples_per_thread, uint64_t seed, int exponent_span,
                                uint64_t *mismatches, uint64_t *samples);

/* Number of tracking-kernel variants compiled in, and a description of variant i
 * ("fast/ppt2/lean", ...) with its register count -- build introspection for tests. */
int xlb_kernel_variant_count(void);
int xlb_kernel_variant_info(int i, char *name, int name_len, int *regs, int *max_threads);

#ifdef __cplusplus
}
#endif
#endif /* XLINE_B200_H */
