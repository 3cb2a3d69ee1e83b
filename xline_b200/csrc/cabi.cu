// C ABI (include/xline_b200.h) and host-side orchestration of the tracking kernels:
// variant selection, shared-memory ring sizing, turn segmentation with survivor
// compaction between launches, the host-buffer entry point, and the FP64-peak probe.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/xline_b200.h"
#include "kargs.h"

namespace xlb {

static thread_local std::string g_err;
static thread_local xlb_track_stats_t g_stats;

static int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
#define XLB_CUDA(call)                                                                     \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      return fail(XLB_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__));         \
    }                                                                                      \
  } while (0)

// ------------------------------------------------------------------ survivor compaction
// Order-preserving stream compaction of {i : state[i]==1} built from warp ballots:
// pass 1 counts per block, pass 2 scans the block totals, pass 3 re-derives the ballots
// and scatters.  Deterministic (no atomics on the output position).
constexpr int CP_THREADS = 512;
constexpr int CP_ITEMS = 8;  // elements per thread

__global__ void __launch_bounds__(CP_THREADS) compact_count(const long long *state, long long n,
                                                            int *block_sums) {
  __shared__ int warp_cnt[CP_THREADS / 32];
  const long long base = static_cast<long long>(blockIdx.x) * (CP_THREADS * CP_ITEMS);
  int cnt = 0;
#pragma unroll
  for (int it = 0; it < CP_ITEMS; ++it) {
    const long long i = base + static_cast<long long>(it) * CP_THREADS + threadIdx.x;
    const bool keep = (i < n) && (state[i] == 1);
    cnt += __popc(__ballot_sync(0xffffffffu, keep));
  }
  if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < CP_THREADS / 32; ++w) t += warp_cnt[w];
    block_sums[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(1024) compact_scan(int *block_sums, int nblocks, int *n_out) {
  // exclusive scan of block_sums by one CTA (nblocks <= a few 1e4), warp-shuffle based
  __shared__ int warp_tot[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int start = 0; start < nblocks; start += 1024) {
    const int i = start + threadIdx.x;
    const int v = (i < nblocks) ? block_sums[i] : 0;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if ((threadIdx.x & 31) >= d) incl += t;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = warp_tot[threadIdx.x];
      int wi = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, wi, d);
        if (threadIdx.x >= d) wi += t;
      }
      warp_tot[threadIdx.x] = wi - w;  // exclusive warp offsets
    }
    __syncthreads();
    const int excl = carry + warp_tot[threadIdx.x >> 5] + incl - v;
    if (i < nblocks) block_sums[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_out = carry;
}

__global__ void __launch_bounds__(CP_THREADS) compact_scatter(const long long *state, long long n,
                                                              const int *block_offs, int *idx_out) {
  __shared__ int warp_cnt[CP_ITEMS][CP_THREADS / 32];
  const long long base = static_cast<long long>(blockIdx.x) * (CP_THREADS * CP_ITEMS);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned ballots[CP_ITEMS];
#pragma unroll
  for (int it = 0; it < CP_ITEMS; ++it) {
    const long long i = base + static_cast<long long>(it) * CP_THREADS + threadIdx.x;
    const bool keep = (i < n) && (state[i] == 1);
    ballots[it] = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_cnt[it][warp] = __popc(ballots[it]);
  }
  __syncthreads();
  // output order = increasing i: item-major, then warp, then lane
  int off = block_offs[blockIdx.x];
#pragma unroll
  for (int it = 0; it < CP_ITEMS; ++it) {
    int before = 0;
    for (int w = 0; w < CP_THREADS / 32; ++w) {
      const int c = warp_cnt[it][w];
      if (w < warp) before += c;
    }
    int tot = 0;
    for (int w = 0; w < CP_THREADS / 32; ++w) tot += warp_cnt[it][w];
    const long long i = base + static_cast<long long>(it) * CP_THREADS + threadIdx.x;
    if ((ballots[it] >> lane) & 1u) {
      const int pos = off + before + __popc(ballots[it] & ((1u << lane) - 1u));
      idx_out[pos] = static_cast<int>(i);
    }
    off += tot;
  }
}

// ------------------------------------------------------------------ per-device scratch
struct Scratch {
  int device = -1;
  int *idx = nullptr;         // survivor list
  int *block_sums = nullptr;  // compaction scratch
  long long cap = 0;          // capacity of idx (entries)
  int nblocks_cap = 0;
  unsigned int *queue = nullptr;   // work queue of the persistent kernel: two counters + ring of ready items
  size_t queue_cap = 0;
  unsigned int *n_lost = nullptr;  // device counter
  int *n_active = nullptr;         // device survivor count
  unsigned int *h_pinned = nullptr;  // [0]=n_lost, [1]=n_active (pinned host)
  // host entry point arena
  unsigned char *arena = nullptr;
  size_t arena_bytes = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // One call at a time per device: the queue, the survivor list, the counters and the pinned
  // read-back words above are shared by every xlb_track_* call on this GPU.  Host threads
  // serialise on `call_mu` (recursive: the host entry point calls the device one); work that a
  // previous call left running on ANOTHER stream is ordered before the next call's by `done`,
  // recorded at the end of every call and waited for at the start of the next.
  std::recursive_mutex call_mu;
  cudaEvent_t done = nullptr;
  bool have_done = false;
  // launch-length schedule (see track_device_impl): loss rate per turn seen by the last
  // segmented call, and the particle set it belonged to
  const double *sched_x = nullptr;
  long long sched_n = 0;
  double sched_rate = -1.0;
};

// Launch length for a loss rate: lanes of lost particles idle until the next launch boundary,
// on average rate * length / 2 of the lanes over a launch -- keep that near 3 %.
static int launch_length_for_rate(double rate_per_turn, int longest) {
  if (!(rate_per_turn > 0)) return longest;
  const double len = 0.06 / rate_per_turn;
  if (len >= longest) return longest;
  return len < 1.0 ? 1 : static_cast<int>(len);
}
static std::mutex g_mu;
static std::vector<Scratch *> g_scratch;

static int get_scratch(Scratch **out) {
  int dev = 0;
  XLB_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_mu);
  for (Scratch *s : g_scratch)
    if (s->device == dev) {
      *out = s;
      return XLB_OK;
    }
  Scratch *s = new Scratch();
  s->device = dev;
  XLB_CUDA(cudaMalloc(&s->n_lost, sizeof(unsigned int)));
  XLB_CUDA(cudaMalloc(&s->n_active, sizeof(int)));
  XLB_CUDA(cudaMemset(s->n_lost, 0, sizeof(unsigned int)));
  XLB_CUDA(cudaMallocHost(&s->h_pinned, 2 * sizeof(unsigned int)));
  XLB_CUDA(cudaEventCreate(&s->ev0));
  XLB_CUDA(cudaEventCreate(&s->ev1));
  XLB_CUDA(cudaEventCreateWithFlags(&s->done, cudaEventDisableTiming));
  g_scratch.push_back(s);
  *out = s;
  return XLB_OK;
}

static int ensure_queue(Scratch *s, size_t words) {
  if (words > s->queue_cap) {
    if (s->queue) cudaFree(s->queue);
    s->queue = nullptr;
    s->queue_cap = 0;
    XLB_CUDA(cudaMalloc(&s->queue, words * sizeof(unsigned int)));
    s->queue_cap = words;
  }
  return XLB_OK;
}

static int ensure_compaction_scratch(Scratch *s, long long n) {
  const int nblocks = static_cast<int>((n + CP_THREADS * CP_ITEMS - 1) / (CP_THREADS * CP_ITEMS));
  if (n > s->cap) {
    if (s->idx) cudaFree(s->idx);
    s->idx = nullptr;
    XLB_CUDA(cudaMalloc(&s->idx, static_cast<size_t>(n) * sizeof(int)));
    s->cap = n;
  }
  if (nblocks > s->nblocks_cap) {
    if (s->block_sums) cudaFree(s->block_sums);
    s->block_sums = nullptr;
    XLB_CUDA(cudaMalloc(&s->block_sums, static_cast<size_t>(nblocks) * sizeof(int)));
    s->nblocks_cap = nblocks;
  }
  return XLB_OK;
}

static int compact_alive(const long long *state, long long n, int *idx_out, int *block_sums,
                         int *n_out, cudaStream_t st) {
  if (n <= 0) {
    XLB_CUDA(cudaMemsetAsync(n_out, 0, sizeof(int), st));
    return XLB_OK;
  }
  const int nblocks = static_cast<int>((n + CP_THREADS * CP_ITEMS - 1) / (CP_THREADS * CP_ITEMS));
  compact_count<<<nblocks, CP_THREADS, 0, st>>>(state, n, block_sums);
  compact_scan<<<1, 1024, 0, st>>>(block_sums, nblocks, n_out);
  compact_scatter<<<nblocks, CP_THREADS, 0, st>>>(state, n, block_sums, idx_out);
  XLB_CUDA(cudaGetLastError());
  return XLB_OK;
}

// ------------------------------------------------------------------ variant selection
// All kernel families (one variant table per translation unit), most specialised first.
typedef const Variant *(*TableFn)(int *);
static const TableFn kTables[] = {fast_lean_nc_lo_variants, fast_lean_nc_variants, fast_lean_variants,
                                  fast_bf_nc_lo_variants,   fast_bf_variants,      fast_bf6_variants,
                                  strict_variants,          strict_bf_variants};
static const int kNumTables = static_cast<int>(sizeof(kTables) / sizeof(kTables[0]));

// What a call asks of a kernel family.
struct Need {
  bool strict, beamfields, bb6d, chi, trace;
  int max_order;  // highest multipole order the lattice may hold (255 = unknown)
};
static bool admits(const Variant &v, const Need &q) {
  if ((v.strict != 0) != q.strict) return false;
  if (q.strict) return (v.beamfields != 0) == q.beamfields;
  const int bf = q.beamfields ? (q.bb6d ? 2 : 1) : 0;
  if (v.beamfields != bf) return false;
  if (v.nochi && q.chi) return false;
  if (v.maxorder && q.max_order > v.maxorder) return false;
  return true;
}
// The first (most specialised) admissible family that has the requested particles-per-thread;
// failing that, the closest particles-per-thread of the most specialised admissible family.
// Within a family: the variant with the smallest launch-bounds ceiling that still admits
// `threads` (tighter ceilings allow more registers).
static const Variant *pick_variant(const Need &q, int ppt, int threads) {
  const Variant *fallback = nullptr;
  for (int t = 0; t < kNumTables; ++t) {
    int n = 0;
    const Variant *tab = kTables[t](&n);
    const Variant *best = nullptr;
    for (int i = 0; i < n; ++i) {
      const Variant &v = tab[i];
      if (!admits(v, q)) continue;
      if (q.trace) {
        if (v.trace) return &v;
        continue;
      }
      if (v.trace) continue;
      if (!best) { best = &v; continue; }
      const int dv = std::abs(v.ppt - ppt), db = std::abs(best->ppt - ppt);
      if (dv < db) { best = &v; continue; }
      if (dv > db) continue;
      const bool vfit = v.threads >= threads, bfit = best->threads >= threads;
      if (vfit && (!bfit || v.threads < best->threads)) best = &v;
      if (!vfit && !bfit && v.threads > best->threads) best = &v;
    }
    if (!best) continue;
    if (best->ppt == ppt && best->threads >= threads) return best;
    if (!fallback) fallback = best;
  }
  return fallback;
}

static int check_lattice_header(const xlb_lattice_t *lat) {
  if (!lat || !lat->words) return fail(XLB_EINVAL, "lattice is null");
  if (lat->chunk_words < 4 || (lat->chunk_words & 1))
    return fail(XLB_ELATTICE, "chunk_words must be even and >= 4");
  if (static_cast<long long>(lat->chunk_words) * 8 > 96 * 1024)
    return fail(XLB_ELATTICE, "chunk larger than 96 KiB");
  if (lat->n_chunks < 1) return fail(XLB_ELATTICE, "n_chunks < 1");
  if (lat->n_words != static_cast<int64_t>(lat->chunk_words) * lat->n_chunks)
    return fail(XLB_ELATTICE, "n_words != chunk_words * n_chunks");
  if (reinterpret_cast<uintptr_t>(lat->words) & 15)
    return fail(XLB_ELATTICE, "lattice words must be 16-byte aligned");
  if (lat->n_segments < 0 || (lat->n_segments > 0 && !lat->segments))
    return fail(XLB_ELATTICE, "bad segment table");
  int next = 0;
  for (int k = 0; k < lat->n_segments; ++k) {
    const int32_t *sg = lat->segments + 3 * k;
    if (sg[0] != next || sg[1] < 1) return fail(XLB_ELATTICE, "segments must tile the chunks in order");
    if (sg[2] != XLB_SEG_MAIN && sg[2] != XLB_SEG_BB6D) return fail(XLB_ELATTICE, "unknown segment kind");
    if (sg[2] == XLB_SEG_BB6D && (sg[1] != 1 || !(lat->flags & XLB_F_BB6D) || (lat->flags & XLB_F_STRICT)))
      return fail(XLB_ELATTICE, "a 6D-lens segment is one chunk of a fast lattice flagged XLB_F_BB6D");
    next += sg[1];
  }
  if (lat->n_segments > 0) {
    if (next != lat->n_chunks) return fail(XLB_ELATTICE, "segments do not cover all chunks");
    if (lat->segments[3 * (lat->n_segments - 1) + 2] != XLB_SEG_MAIN)
      return fail(XLB_ELATTICE, "the last segment must be a MAIN segment");
  }
  return XLB_OK;
}

static int check_particles(const xlb_particles_t *p) {
  if (!p) return fail(XLB_EINVAL, "particles is null");
  if (p->n < 0) return fail(XLB_EINVAL, "negative particle count");
  if (p->n > 2147483647LL) return fail(XLB_EINVAL, "more than 2^31-1 particles per call");
  if (p->n == 0) return XLB_OK;
  if (!p->x || !p->px || !p->y || !p->py || !p->zeta || !p->delta || !p->rpp || !p->rvv ||
      !p->s || !p->state || !p->at_element || !p->at_turn || !p->particle_id)
    return fail(XLB_EINVAL, "a required particle column is null");
  if (!(p->beta0 > 0) || !(p->energy0 > 0) || !(p->p0c > 0))
    return fail(XLB_EINVAL, "reference particle (p0c, beta0, energy0) must be positive");
  return XLB_OK;
}

static int track_device_impl(const xlb_lattice_t *lat, xlb_particles_t *p,
                             const xlb_track_options_t *o, cudaStream_t st, bool timed) {
  memset(&g_stats, 0, sizeof(g_stats));
  int rc;
  if ((rc = check_lattice_header(lat)) != XLB_OK) return rc;
  if ((rc = check_particles(p)) != XLB_OK) return rc;
  if (!o) return fail(XLB_EINVAL, "options is null");
  if (o->num_turns < 0) return fail(XLB_EINVAL, "num_turns < 0");
  if (p->n == 0 || o->num_turns == 0) return XLB_OK;

  const bool strict = (lat->flags & XLB_F_STRICT) != 0;
  const bool beamfields = (lat->flags & XLB_F_BEAMFIELDS) != 0;
  const bool split = lat->n_segments > 1;  // 6D lenses run as kernels of their own
  const bool bb6d = (lat->flags & XLB_F_BB6D) != 0 && !split;
  // Defaults from the B200 sweeps (scripts/probe_variants.py, scripts/probe_strict.py).  Thin-lens
  // lattices of one species (chi == NULL): 4 particles per thread in 128-thread CTAs, 3 CTAs/SM at
  // 168 registers = 1 536 particles in flight per SM (C2 +5 %, C4 +23 % over 3 x 128 x 3); with a
  // chi column 3 x 128 (162 registers, no spills).  Strict C2: 1.08e7 particle-turns/s with 3 x 128
  // against 9.1e6 (2 x 128), 7.3e6 (1 x 256), 6.7e6 (2 x 256).  Beam-field lattices: 2 x 256 unless
  // the caller knows the lenses are sparse (xline_b200/line.py decides from the record counts).
  Need need;
  need.strict = strict;
  need.beamfields = beamfields;
  need.bb6d = bb6d;
  need.chi = p->chi != nullptr;
  need.trace = o->trace != nullptr;
  need.max_order = (lat->flags & XLB_F_LOW_ORDER) ? 3 : 255;
  const bool trace = need.trace;
  if (trace && (o->num_turns != 1 || o->trace_particles < 1))
    return fail(XLB_EINVAL, "element-by-element trace needs num_turns == 1 and trace_particles >= 1");
  if (split && (trace || strict))
    return fail(XLB_EINVAL, "segmented lattices are for the fast kernels without trace");
  int dev = 0, sms = 0;
  XLB_CUDA(cudaGetDevice(&dev));
  XLB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t chunk_bytes = static_cast<size_t>(lat->chunk_words) * 8;
  const size_t smem = XLB_STAGES * chunk_bytes + (2 * XLB_STAGES + 2) * sizeof(unsigned long long);
  // The launch shape is chosen per launch from the number of particles still in the beam: a small
  // (or scraped-down) beam is spread over all SMs first -- fewer particles per thread as long as
  // the CTAs of the preferred shape would leave CTA slots empty (strong scaling, C1, the heavy-loss
  // C2 beam).  All variants give the same bits, so the choice is invisible in the results.
  struct Shape {
    const Variant *v;
    int threads, resident, regs;
  };
  auto choose = [&](long long n_now, Shape *out) -> int {
    int ppt_req = o->particles_per_thread;
    if (ppt_req <= 0) {
      if (beamfields) ppt_req = 2;
      else if (strict || need.chi) ppt_req = 3;
      else ppt_req = 4;
      if (!beamfields && o->threads_per_block <= 0) {
        // (below 2 only for beams that cannot give every SM one 128-thread CTA: one particle per
        // thread doubles the warps but also the per-record instructions issued, and loses on the
        // C2 lattice from 30 000 particles on -- profiles/r2_sweep_n.json; C1's 10 000 particles
        // on 79 CTAs instead of 40: 3.3e9 against 2.3e9 particle-turns/s)
        // (three per thread wins nowhere in the closing sweep, profiles/r2c_sweep_n.json: 200 k
        // particles 2.45e7 with two, 2.38e7 with three, 2.11e7 with four)
        if (ppt_req > 2 && n_now < static_cast<long long>(ppt_req) * 128 * 3 * sms) ppt_req = 2;
        if (ppt_req == 2 && n_now < 128LL * sms) ppt_req = 1;
      }
    }
    const int threads_req =
        o->threads_per_block > 0 ? o->threads_per_block : (beamfields && ppt_req < 3 ? 256 : 128);
    const Variant *v = pick_variant(need, ppt_req, threads_req);
    if (!v) return fail(XLB_EINVAL, "no kernel variant compiled for this lattice");
    const int threads =
        trace ? v->threads : (o->threads_per_block > 0 ? o->threads_per_block : std::min(v->threads, threads_req));
    if (threads % 32 || threads > v->threads)
      return fail(XLB_EINVAL, "threads_per_block must be a multiple of 32 and <= the variant's limit");
    XLB_CUDA(cudaFuncSetAttribute(v->func, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    cudaFuncAttributes fa;
    XLB_CUDA(cudaFuncGetAttributes(&fa, v->func));
    int occ = 1;
    XLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, v->func, threads, smem));
    out->v = v;
    out->threads = threads;
    out->resident = std::max(1, occ) * sms;  // CTAs the device holds at once
    out->regs = fa.numRegs;
    return XLB_OK;
  };
  Shape shape;
  if ((rc = choose(p->n, &shape)) != XLB_OK) return rc;  // also validates the caller's request

  Scratch *s = nullptr;
  if ((rc = get_scratch(&s)) != XLB_OK) return rc;
  std::lock_guard<std::recursive_mutex> call_lock(s->call_mu);
  if (s->have_done) XLB_CUDA(cudaStreamWaitEvent(st, s->done, 0));
  struct DoneMark {  // whatever way the call ends, later calls wait for what it enqueued
    Scratch *s;
    cudaStream_t st;
    ~DoneMark() {
      if (cudaEventRecord(s->done, st) == cudaSuccess) s->have_done = true;
    }
  } done_mark{s, st};

  KArgs a;
  memset(&a, 0, sizeof(a));
  a.lat = lat->words;
  a.chunk_words = lat->chunk_words;
  a.n_chunks = lat->n_chunks;
  a.x = p->x; a.px = p->px; a.y = p->y; a.py = p->py; a.zeta = p->zeta; a.delta = p->delta;
  a.rpp = p->rpp; a.rvv = p->rvv; a.s = p->s; a.chi = p->chi; a.qr = p->charge_ratio;
  a.state = reinterpret_cast<long long *>(p->state);
  a.at_element = reinterpret_cast<long long *>(p->at_element);
  a.at_turn = reinterpret_cast<long long *>(p->at_turn);
  a.pid = reinterpret_cast<const long long *>(p->particle_id);
  a.q0 = p->q0; a.p0c = p->p0c; a.beta0 = p->beta0; a.energy0 = p->energy0;
  a.sc_common = p->q0 * p->q0 * (1.0 - p->beta0 * p->beta0) / (p->p0c * p->beta0);
  a.loss_tally = reinterpret_cast<long long *>(o->loss_tally);
  a.mon = o->monitor_data;
  a.mon_words = o->monitor_words;
  a.n_lost = s->n_lost;
  a.trace = o->trace;
  a.trace_n = o->trace_particles;
  a.elem_off = static_cast<int>(o->element_index_offset);
  const int count_turns = (o->flags & XLB_OPT_NO_TURN_COUNT) ? 0 : 1;

  // Work-item length.  An item costs a load and a store of its block's particles and a ticket
  // (microseconds); a shorter item leaves a shorter tail at the end of a launch (the last, partly
  // filled round of items) and rotates the blocks faster.  Automatic: one turn when a turn streams
  // 16 chunks or more (LHC: 121, C2 at 250 k particles +3 %, at 1 M +1 % over five turns), else
  // as many turns as make 16 chunks (PS Booster: 6).
  const int tpi = (o->turns_per_item == 0) ? std::max(1, (16 + lat->n_chunks - 1) / lat->n_chunks)
                                           : o->turns_per_item;
  // 0 = automatic: long jobs are cut into launches of 100 turns so that survivors get
  // re-compacted now and then; < 0 = one launch whatever the length
  // (the kernel counts the chunks of a launch in 32 bits: a launch never exceeds 2^31 / n_chunks turns)
  const int seg_cap = std::max(1, static_cast<int>(0x7fffffffLL / std::max(1, lat->n_chunks)));
  const int seg = std::min(seg_cap, (o->turns_per_launch > 0) ? o->turns_per_launch
                  : (o->turns_per_launch == 0 && o->num_turns > 150 ? 100 : o->num_turns));
  const double thr = (o->compact_threshold > 0) ? o->compact_threshold : (1.0 / 128.0);
  long long n_active = p->n;
  const int *idx = nullptr;
  long long lost_since_compact = 0;
  const bool segmented = seg < o->num_turns;
  if (segmented) XLB_CUDA(cudaMemsetAsync(s->n_lost, 0, sizeof(unsigned int), st));
  // Launch-length ramp: a beam that is scraped hard in its first turns (80 % of the
  // C2 beam within a few turns) would keep those lanes idle for a whole launch.  A segmented job
  // on a particle set not seen before starts with a launch of ONE turn; after every launch the
  // loss rate per turn just measured sets the next length (at most twice the previous one, at
  // most `seg`), and survivors are re-compacted as before.  A call that continues with the
  // particle set of the previous call starts from the rate that call ended with.  The schedule
  // never changes results (a launch boundary is invisible to the particles).
  const bool ramp = segmented && p->n >= 32768;
  int cur_len = seg;
  if (ramp) {
    const bool continuing = (s->sched_x == p->x && s->sched_n == p->n && s->sched_rate >= 0.0);
    cur_len = continuing ? launch_length_for_rate(s->sched_rate, seg) : 1;
  }
  double last_rate = -1.0;

  if (timed) XLB_CUDA(cudaEventRecord(s->ev0, st));
  if (o->turns_per_launch > 0 || seg < o->num_turns) {
    // particles lost in earlier calls still sit in the caller's arrays (state 0): start from
    // the compacted survivor list so they do not occupy lanes
    if ((rc = ensure_compaction_scratch(s, p->n)) != XLB_OK) return rc;
    if ((rc = compact_alive(a.state, p->n, s->idx, s->block_sums, s->n_active, st)) != XLB_OK)
      return rc;
    XLB_CUDA(cudaMemcpyAsync(&s->h_pinned[1], s->n_active, sizeof(int), cudaMemcpyDeviceToHost, st));
    XLB_CUDA(cudaStreamSynchronize(st));
    g_stats.compactions += 1;
    if (static_cast<long long>(s->h_pinned[1]) < p->n) {
      n_active = static_cast<int>(s->h_pinned[1]);
      idx = s->idx;
    }
  }
  float total_ms = 0.f;
  long long lost_before_launch = 0;
  for (int done = 0; done < o->num_turns && n_active > 0;) {
    const int turns = std::min(ramp ? cur_len : seg, o->num_turns - done);
    a.num_turns = turns;
    a.n = n_active;
    a.idx = idx;
    if ((rc = choose(n_active, &shape)) != XLB_OK) return rc;
    const Variant *v = shape.v;
    const int threads = shape.threads, resident = shape.resident;
    const long long per_block = static_cast<long long>(threads) * v->ppt;
    const int blocks = static_cast<int>((n_active + per_block - 1) / per_block);
    int grid = blocks;
    a.queue = nullptr;
    a.n_blocks = static_cast<unsigned int>(blocks);
    a.n_items = static_cast<unsigned int>(blocks);
    a.turns_per_item = turns;
    a.count_turns = count_turns;
    if (split) {
      // turn by turn, segment by segment: tracking kernel up to the next 6D lens, the lens
      // kernel, and so on; the closing MAIN segment counts the turn
      a.num_turns = 1;
      a.turns_per_item = 1;
      const int lens_threads = 128;
      const int lens_blocks = static_cast<int>((n_active + lens_threads - 1) / lens_threads);
      for (int t = 0; t < turns; ++t) {
        for (int k = 0; k < lat->n_segments; ++k) {
          const int32_t *sg = lat->segments + 3 * k;
          const uint64_t *first = lat->words + static_cast<size_t>(sg[0]) * lat->chunk_words;
          if (sg[2] == XLB_SEG_BB6D) {
            fast_bb6d_launch(a, reinterpret_cast<const unsigned long long *>(first), lens_blocks,
                             lens_threads, st);
          } else {
            a.lat = first;
            a.n_chunks = sg[1];
            a.count_turns = (k == lat->n_segments - 1) ? count_turns : 0;
            v->launch(a, blocks, threads, smem, st);
          }
          g_stats.kernel_launches += 1;
        }
      }
      XLB_CUDA(cudaGetLastError());
      g_stats.blocks = blocks;
      g_stats.threads = threads;
    } else {
    if (tpi > 0 && turns > tpi && blocks > sms) {
      // persistent CTAs + a device-side ring of ready (block, turn segment) items (track_impl.cuh).
      // Also when every CTA could be resident at once (sms < blocks <= resident): SMs would hold
      // different numbers of CTAs, a CTA on a fuller SM runs slower, and with one item per CTA
      // the launch lasts as long as its slowest CTA.
      const long long segs = (turns + tpi - 1) / tpi;
      if (static_cast<long long>(blocks) * segs < (1LL << 26)) {  // ring of at most 512 MiB
        // two counters + one 64-bit ready entry per item beyond the first segment of every block
        const size_t qwords = 2 + 2 * static_cast<size_t>(blocks) * static_cast<size_t>(segs - 1);
        if ((rc = ensure_queue(s, qwords)) != XLB_OK) return rc;
        XLB_CUDA(cudaMemsetAsync(s->queue, 0, qwords * sizeof(unsigned int), st));
        a.queue = s->queue;
        a.n_items = static_cast<unsigned int>(blocks * segs);
        a.turns_per_item = tpi;
        // Fewer CTAs than blocks (unless the blocks load every SM equally): the blocks beyond the
        // grid wait in the ring, whose FIFO
        // order then rotates the blocks over the CTAs, and all of them advance at the average speed.  (With every block resident on a CTA of its own a finishing CTA finds only its
        // own block in the ring -- a static assignment, as slow as the CTAs of the fullest SMs:
        // C2 at 125 k particles 1.9e7 against 2.2e7.)  Preferably the same number of CTAs on every
        // SM; when that would idle more than 15 % of the CTAs, a backlog of 1/16 of the blocks.
        const int q = std::min(blocks, resident), rem = q % sms;
        grid = (rem * 100 <= q * 15) ? q - rem : q - std::max(1, q / 16);
      }
    }
    v->launch(a, grid, threads, smem, st);
    XLB_CUDA(cudaGetLastError());
    g_stats.kernel_launches += 1;
    g_stats.blocks = blocks;
    g_stats.threads = threads;
    }
    done += turns;
    if (done >= o->num_turns) break;
    // survivors: read the loss counter; re-compact when enough lanes went idle
    XLB_CUDA(cudaMemcpyAsync(&s->h_pinned[0], s->n_lost, sizeof(unsigned int),
                             cudaMemcpyDeviceToHost, st));
    XLB_CUDA(cudaStreamSynchronize(st));
    const long long lost_total = s->h_pinned[0];
    if (ramp) {
      last_rate = static_cast<double>(lost_total - lost_before_launch) /
                  (static_cast<double>(n_active) * static_cast<double>(turns));
      cur_len = std::min(launch_length_for_rate(last_rate, seg), 2 * cur_len);
    }
    lost_before_launch = lost_total;
    const long long newly = lost_total - lost_since_compact;
    if (newly > 0 && static_cast<double>(newly) >= thr * static_cast<double>(n_active)) {
      if ((rc = ensure_compaction_scratch(s, p->n)) != XLB_OK) return rc;
      if ((rc = compact_alive(a.state, p->n, s->idx, s->block_sums, s->n_active, st)) != XLB_OK)
        return rc;
      XLB_CUDA(cudaMemcpyAsync(&s->h_pinned[1], s->n_active, sizeof(int), cudaMemcpyDeviceToHost, st));
      XLB_CUDA(cudaStreamSynchronize(st));
      n_active = static_cast<int>(s->h_pinned[1]);
      idx = s->idx;
      lost_since_compact = lost_total;
      g_stats.compactions += 1;
    }
  }
  if (timed) {
    XLB_CUDA(cudaEventRecord(s->ev1, st));
    XLB_CUDA(cudaEventSynchronize(s->ev1));
    XLB_CUDA(cudaEventElapsedTime(&total_ms, s->ev0, s->ev1));
  }
  if (ramp) {
    s->sched_x = p->x;
    s->sched_n = p->n;
    if (last_rate >= 0.0) s->sched_rate = last_rate;
    else if (!(s->sched_rate >= 0.0)) s->sched_rate = 0.0;
  }
  g_stats.kernel_ms = total_ms;
  g_stats.regs_per_thread = shape.regs;
  g_stats.smem_bytes = static_cast<int>(smem);
  g_stats.n_alive_in = p->n;
  g_stats.n_alive_out = n_active;
  return XLB_OK;
}

// ------------------------------------------------------------------ FP64 peak probe
template <int CHAINS>
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double a, double b) {
  double acc[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) acc[c] = threadIdx.x * 1e-3 + c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = fma(acc[c], a, b);
  }
  double t = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) t += acc[c];
  if (t == 123.456) out[0] = t;  // keep the chains alive
}

// One warp, CHAINS independent dependent-DFMA chains: cycles per DFMA (clock64).  CHAINS = 1
// gives the dependent-issue latency, larger CHAINS the single-warp throughput.
template <int CHAINS>
__global__ void dfma_latency_kernel(double *out, long long *cycles, int iters, double a, double b) {
  double acc[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) acc[c] = threadIdx.x * 1e-3 + c;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = fma(acc[c], a, b);
  }
  const long long t1 = clock64();
  double t = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) t += acc[c];
  if (threadIdx.x == 0) *cycles = t1 - t0;
  if (t == 123.456) out[0] = t;
}

template <int CHAINS>
static int run_latency(double *d, long long *dc, double *cyc_per_dfma) {
  const int iters = 1 << 14;
  dfma_latency_kernel<CHAINS><<<1, 32>>>(d, dc, iters, 1.0000001, 1e-9);
  dfma_latency_kernel<CHAINS><<<1, 32>>>(d, dc, iters, 1.0000001, 1e-9);
  long long h = 0;
  XLB_CUDA(cudaMemcpy(&h, dc, sizeof(h), cudaMemcpyDeviceToHost));
  *cyc_per_dfma = static_cast<double>(h) / (static_cast<double>(iters) * CHAINS);
  return XLB_OK;
}

}  // namespace xlb

using namespace xlb;

extern "C" {

int xlb_abi_version(void) { return XLB_ABI_VERSION; }
const char *xlb_last_error(void) { return g_err.c_str(); }

int xlb_get_stats(xlb_track_stats_t *out) {
  if (!out) return fail(XLB_EINVAL, "out is null");
  *out = g_stats;
  return XLB_OK;
}

int xlb_lattice_validate(const xlb_lattice_t *lat) {
  int rc = check_lattice_header(lat);
  if (rc != XLB_OK) return rc;
  bool saw_end_turn = false;
  // last chunk of every segment (of the lattice when it is not segmented) and the chunks
  // that belong to 6D-lens segments
  std::vector<char> seg_last(lat->n_chunks, 0), lens_chunk(lat->n_chunks, 0);
  seg_last[lat->n_chunks - 1] = 1;
  for (int k = 0; k < lat->n_segments; ++k) {
    const int32_t *sg = lat->segments + 3 * k;
    seg_last[sg[0] + sg[1] - 1] = 1;
    if (sg[2] == XLB_SEG_BB6D) lens_chunk[sg[0]] = 1;
  }
  for (int c = 0; c < lat->n_chunks; ++c) {
    const uint64_t *w = lat->words + static_cast<size_t>(c) * lat->chunk_words;
    int pos = 0;
    bool closed = false;
    while (pos < lat->chunk_words) {
      const uint64_t hdr = w[pos];
      const int tag = static_cast<int>(hdr & 0xff);
      const int aux = static_cast<int>((hdr >> 8) & 0xff);
      const int pairs = static_cast<int>((hdr >> 16) & 0x3fff);
      const bool hdr_a1 = (hdr & XLB_HDR_HAS_A1) != 0;
      const bool hx_only = ((hdr >> 31) & 1) != 0;  // XLB_HDR_HX_ONLY
      int want = -1;  // expected record length in pairs, -1 = variable
      switch (tag) {
        case XLB_T_END_TURN:
          if (!seg_last[c]) return fail(XLB_ELATTICE, "END_TURN before the last chunk");
          saw_end_turn = true;
          closed = true;
          break;
        case XLB_T_END_CHUNK:
          if (seg_last[c]) return fail(XLB_ELATTICE, "last chunk must end with END_TURN");
          closed = true;
          break;
        case XLB_T_DRIFT:
        case XLB_T_DRIFT_EXACT: want = 1; break;
        case XLB_T_MULTIPOLE:
        case XLB_T_MULTIPOLE_CURVED:
          want = (tag == XLB_T_MULTIPOLE ? 1 : 3) + aux + 1;
          if ((lat->flags & XLB_F_LOW_ORDER) && aux > 3)
            return fail(XLB_ELATTICE, "XLB_F_LOW_ORDER set but a multipole has order > 3");
          break;
        case XLB_T_CAVITY:
        case XLB_T_SAWTOOTH_CAVITY:
        case XLB_T_XYSHIFT:
        case XLB_T_SROTATION:
        case XLB_T_DIPOLE_EDGE: want = 2; break;
        case XLB_T_RFMULTIPOLE: want = 2 + 2 * (aux + 1); break;
        case XLB_T_LIMIT_RECT:
        case XLB_T_LIMIT_ELLIPSE: want = 3 + ((aux & XLB_AUX_DRIFT) ? 1 : 0); break;
        case XLB_T_LIMIT_RECT_ELLIPSE: want = 4 + ((aux & XLB_AUX_DRIFT) ? 1 : 0); break;
        case XLB_T_MONITOR: want = 5; break;
        case XLB_T_BEAMBEAM4D:
        case XLB_T_SPACECHARGE:
        case XLB_T_BEAMBEAM6D:
          if (lens_chunk[c]) {
            if (tag != XLB_T_BEAMBEAM6D || pos != 0)
              return fail(XLB_ELATTICE, "a 6D-lens segment holds exactly one BEAMBEAM6D record");
          } else {
            if (!(lat->flags & XLB_F_BEAMFIELDS))
              return fail(XLB_ELATTICE, "beam-field record without XLB_F_BEAMFIELDS");
            if (tag == XLB_T_BEAMBEAM6D && lat->n_segments > 1)
              return fail(XLB_ELATTICE, "BEAMBEAM6D record in a MAIN segment of a segmented lattice");
          }
          if (tag == XLB_T_BEAMBEAM6D && !(lat->flags & XLB_F_BB6D))
            return fail(XLB_ELATTICE, "BeamBeam6D record without XLB_F_BB6D");
          if (pairs < 6) return fail(XLB_ELATTICE, "bad beam-field record length");
          break;
        default: {
          if ((tag & 0xc0) == XLB_T_EDGE_BLOCK) {
            want = 2;
            break;
          }
          if ((tag & 0xc0) == XLB_T_THIN_BLOCK && (lat->flags & XLB_F_LOW_ORDER) && aux > 3)
            return fail(XLB_ELATTICE, "XLB_F_LOW_ORDER set but a block record has order > 3");
          if ((tag & 0xe0) == XLB_T_THIN_BLOCK) {
            want = 2 + aux + 1 + ((tag & 4) ? 2 : 0) + ((tag & 3) ? 2 : 0);
            break;
          }
          if ((tag & 0xe0) == XLB_T_MERGED_BLOCK) {
            if (lat->flags & XLB_F_STRICT) return fail(XLB_ELATTICE, "merged block in a strict lattice");
            const int64_t f = static_cast<int64_t>(w[pos + 3]);
            if ((((f >> 8) & 1) != 0) != hdr_a1)
              return fail(XLB_ELATTICE, "merged block: XLB_HDR_HAS_A1 disagrees with the record");
            want = 2 + aux + 1 + ((tag & 4) ? 3 : 0) + (((f >> 8) & 1) ? 2 : 0) + ((tag & 3) ? 2 : 0) +
                   static_cast<int>(f & 0xff) + 1 + 1;  // ... K1's pairs, [path length, 0]
            break;
          }
          char buf[96];
          snprintf(buf, sizeof buf, "unknown tag %d in chunk %d at word %d", tag, c, pos);
          return fail(XLB_ELATTICE, buf);
        }
      }
      if (hdr_a1 && (tag & 0xe0) != XLB_T_MERGED_BLOCK)
        return fail(XLB_ELATTICE, "XLB_HDR_HAS_A1 on a record that is not a merged block");
      if (hx_only) {  // only on curved block records of the fast encoding, and only when hyl == 0
        const bool block = (tag & 0xc0) == XLB_T_THIN_BLOCK && (tag & 4);
        if (!block || (lat->flags & XLB_F_STRICT))
          return fail(XLB_ELATTICE, "XLB_HDR_HX_ONLY on a record that is not a curved block of the fast encoding");
        double hyl;
        memcpy(&hyl, &w[pos + 2 * (2 + aux + 1) + 1], sizeof hyl);
        if (hyl != 0.0) return fail(XLB_ELATTICE, "XLB_HDR_HX_ONLY set but hyl != 0");
      }
      if (lens_chunk[c] && tag != XLB_T_BEAMBEAM6D && tag != XLB_T_END_TURN)
        return fail(XLB_ELATTICE, "a 6D-lens segment holds exactly one BEAMBEAM6D record");
      if (!closed && (pairs < 1 || (want >= 0 && pairs != want))) {
        char buf[112];
        snprintf(buf, sizeof buf, "record of tag %d in chunk %d at word %d has size %d, expected %d",
                 tag, c, pos, pairs, want);
        return fail(XLB_ELATTICE, buf);
      }
      if (!closed && pos + 2 * pairs + 2 > lat->chunk_words)
        return fail(XLB_ELATTICE, "record runs past the end of its chunk");
      if (closed) break;
      pos += 2 * pairs;
    }
    if (!closed) return fail(XLB_ELATTICE, "chunk not terminated by END_CHUNK/END_TURN");
  }
  if (!saw_end_turn) return fail(XLB_ELATTICE, "no END_TURN record");
  return XLB_OK;
}

int xlb_track_device(const xlb_lattice_t *lattice, xlb_particles_t *particles,
                     const xlb_track_options_t *opts, void *stream) {
  return track_device_impl(lattice, particles, opts, static_cast<cudaStream_t>(stream), false);
}

int xlb_track_device_timed(const xlb_lattice_t *lattice, xlb_particles_t *particles,
                           const xlb_track_options_t *opts, void *stream) {
  return track_device_impl(lattice, particles, opts, static_cast<cudaStream_t>(stream), true);
}

int xlb_compact_alive_device(const int64_t *state, int64_t n, int32_t *idx_out, int32_t *n_out,
                             void *stream) {
  if (n < 0 || (n > 0 && (!state || !idx_out)) || !n_out) return fail(XLB_EINVAL, "bad argument");
  if (n > 2147483647LL) return fail(XLB_EINVAL, "n too large");
  Scratch *s = nullptr;
  int rc = get_scratch(&s);
  if (rc != XLB_OK) return rc;
  std::lock_guard<std::recursive_mutex> call_lock(s->call_mu);
  cudaStream_t cst = static_cast<cudaStream_t>(stream);
  if (s->have_done) XLB_CUDA(cudaStreamWaitEvent(cst, s->done, 0));
  struct DoneMark {
    Scratch *s;
    cudaStream_t st;
    ~DoneMark() {
      if (cudaEventRecord(s->done, st) == cudaSuccess) s->have_done = true;
    }
  } done_mark{s, cst};
  if ((rc = ensure_compaction_scratch(s, 1)) != XLB_OK) return rc;
  const int nblocks = static_cast<int>((n + CP_THREADS * CP_ITEMS - 1) / (CP_THREADS * CP_ITEMS));
  if (nblocks > s->nblocks_cap) {
    if (s->block_sums) cudaFree(s->block_sums);
    s->block_sums = nullptr;
    XLB_CUDA(cudaMalloc(&s->block_sums, static_cast<size_t>(nblocks) * sizeof(int)));
    s->nblocks_cap = nblocks;
  }
  return compact_alive(reinterpret_cast<const long long *>(state), n, idx_out, s->block_sums, n_out,
                       static_cast<cudaStream_t>(stream));
}

int xlb_track_host(const xlb_lattice_t *hl, xlb_particles_t *hp, const xlb_track_options_t *o) {
  int rc;
  if ((rc = check_lattice_header(hl)) != XLB_OK) return rc;
  if ((rc = check_particles(hp)) != XLB_OK) return rc;
  if (!o) return fail(XLB_EINVAL, "options is null");
  if ((rc = xlb_lattice_validate(hl)) != XLB_OK) return rc;
  if (o->trace) return fail(XLB_EINVAL, "element-by-element trace is a device-entry-point feature");
  const long long n = hp->n;
  if (n == 0 || o->num_turns == 0) return XLB_OK;
  Scratch *s = nullptr;
  if ((rc = get_scratch(&s)) != XLB_OK) return rc;
  std::lock_guard<std::recursive_mutex> call_lock(s->call_mu);
  if (!s->stream) XLB_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  cudaStream_t st = s->stream;

  // device arena: lattice | 11 fp64 columns | 4 int64 columns | tally | monitor
  auto al = [](size_t v) { return (v + 255) & ~static_cast<size_t>(255); };
  const size_t col = al(static_cast<size_t>(n) * 8);
  const size_t lat_bytes = al(static_cast<size_t>(hl->n_words) * 8);
  const size_t tally_bytes = o->loss_tally ? al(static_cast<size_t>(hl->n_elements) * 8) : 0;
  const size_t mon_bytes = o->monitor_data ? al(static_cast<size_t>(o->monitor_words) * 8) : 0;
  const size_t need = lat_bytes + 15 * col + tally_bytes + mon_bytes;
  if (need > s->arena_bytes) {
    if (s->arena) cudaFree(s->arena);
    s->arena = nullptr;
    s->arena_bytes = 0;
    XLB_CUDA(cudaMalloc(&s->arena, need));
    s->arena_bytes = need;
  }
  unsigned char *cur = s->arena;
  auto take = [&](size_t bytes) { unsigned char *r = cur; cur += bytes; return r; };
  xlb_lattice_t dl = *hl;
  dl.words = reinterpret_cast<const uint64_t *>(take(lat_bytes));
  XLB_CUDA(cudaMemcpyAsync(const_cast<uint64_t *>(dl.words), hl->words,
                           static_cast<size_t>(hl->n_words) * 8, cudaMemcpyHostToDevice, st));
  xlb_particles_t dp = *hp;
  double **dcols[] = {&dp.x, &dp.px, &dp.y, &dp.py, &dp.zeta, &dp.delta, &dp.rpp, &dp.rvv, &dp.s};
  double *const hcols[] = {hp->x, hp->px, hp->y, hp->py, hp->zeta, hp->delta, hp->rpp, hp->rvv, hp->s};
  int64_t **dicols[] = {&dp.state, &dp.at_element, &dp.at_turn};
  int64_t *const hicols[] = {hp->state, hp->at_element, hp->at_turn};
  // a chi column of ones is the one-species case: not uploaded, and the kernels without a chi
  // register serve the call (see xlb_particles_t)
  bool chi_trivial = true;
  if (hp->chi)
    for (long long i = 0; i < n; ++i)
      if (hp->chi[i] != 1.0) { chi_trivial = false; break; }
  if (chi_trivial) dp.chi = nullptr;

  // Small single-launch calls on PINNED host buffers (C1: 10 000 particles x 100 turns): the
  // kernel reads and writes the caller's arrays in place through their device aliases -- the
  // particle state crosses the bus once each way inside the kernel's own entry and exit, instead
  // of 26 separate DMA transfers of 80 kB whose set-up latencies dominate a 0.3 ms call.
  const int seg_len = (o->turns_per_launch > 0) ? o->turns_per_launch
                      : (o->turns_per_launch == 0 && o->num_turns > 150 ? 100 : o->num_turns);
  bool in_place = n <= 16384 && seg_len >= o->num_turns;
  if (in_place) {
    auto alias = [&](const void *h, void **d) {
      cudaPointerAttributes at;
      if (cudaPointerGetAttributes(&at, h) != cudaSuccess) {
        cudaGetLastError();
        return false;
      }
      if (at.type != cudaMemoryTypeHost || !at.devicePointer) return false;
      *d = at.devicePointer;
      return true;
    };
    xlb_particles_t ap = dp;
    bool ok = true;
    double **acols[] = {&ap.x, &ap.px, &ap.y, &ap.py, &ap.zeta, &ap.delta, &ap.rpp, &ap.rvv, &ap.s};
    for (int c = 0; c < 9 && ok; ++c) ok = alias(hcols[c], reinterpret_cast<void **>(acols[c]));
    int64_t **aicols[] = {&ap.state, &ap.at_element, &ap.at_turn};
    for (int c = 0; c < 3 && ok; ++c) ok = alias(hicols[c], reinterpret_cast<void **>(aicols[c]));
    if (ok) ok = alias(hp->particle_id, reinterpret_cast<void **>(const_cast<int64_t **>(&ap.particle_id)));
    if (ok && ap.chi) ok = alias(hp->chi, reinterpret_cast<void **>(const_cast<double **>(&ap.chi)));
    if (ok && ap.charge_ratio)
      ok = alias(hp->charge_ratio, reinterpret_cast<void **>(const_cast<double **>(&ap.charge_ratio)));
    in_place = ok;
    if (ok) dp = ap;
  }
  if (!in_place)
    for (int c = 0; c < 9; ++c) {
      *dcols[c] = reinterpret_cast<double *>(take(col));
      XLB_CUDA(cudaMemcpyAsync(*dcols[c], hcols[c], static_cast<size_t>(n) * 8, cudaMemcpyHostToDevice, st));
    }
  if (!in_place && hp->chi && !chi_trivial) {
    double *d = reinterpret_cast<double *>(take(col));
    XLB_CUDA(cudaMemcpyAsync(d, hp->chi, static_cast<size_t>(n) * 8, cudaMemcpyHostToDevice, st));
    dp.chi = d;
  }
  if (!in_place && hp->charge_ratio) {
    double *d = reinterpret_cast<double *>(take(col));
    XLB_CUDA(cudaMemcpyAsync(d, hp->charge_ratio, static_cast<size_t>(n) * 8, cudaMemcpyHostToDevice, st));
    dp.charge_ratio = d;
  }
  if (!in_place)
    for (int c = 0; c < 3; ++c) {
      *dicols[c] = reinterpret_cast<int64_t *>(take(col));
      XLB_CUDA(cudaMemcpyAsync(*dicols[c], hicols[c], static_cast<size_t>(n) * 8, cudaMemcpyHostToDevice, st));
    }
  if (!in_place) {
    int64_t *d = reinterpret_cast<int64_t *>(take(col));
    XLB_CUDA(cudaMemcpyAsync(d, hp->particle_id, static_cast<size_t>(n) * 8, cudaMemcpyHostToDevice, st));
    dp.particle_id = d;
  }
  xlb_track_options_t od = *o;
  if (o->loss_tally) {
    od.loss_tally = reinterpret_cast<int64_t *>(take(tally_bytes));
    XLB_CUDA(cudaMemcpyAsync(od.loss_tally, o->loss_tally, static_cast<size_t>(hl->n_elements) * 8,
                             cudaMemcpyHostToDevice, st));
  }
  if (o->monitor_data) {
    od.monitor_data = reinterpret_cast<double *>(take(mon_bytes));
    XLB_CUDA(cudaMemcpyAsync(od.monitor_data, o->monitor_data, static_cast<size_t>(o->monitor_words) * 8,
                             cudaMemcpyHostToDevice, st));
  }
  if ((rc = track_device_impl(&dl, &dp, &od, st, !in_place)) != XLB_OK) return rc;
  if (!in_place) {
    for (int c = 0; c < 9; ++c)
      XLB_CUDA(cudaMemcpyAsync(hcols[c], *dcols[c], static_cast<size_t>(n) * 8, cudaMemcpyDeviceToHost, st));
    for (int c = 0; c < 3; ++c)
      XLB_CUDA(cudaMemcpyAsync(hicols[c], *dicols[c], static_cast<size_t>(n) * 8, cudaMemcpyDeviceToHost, st));
  }
  if (o->loss_tally)
    XLB_CUDA(cudaMemcpyAsync(o->loss_tally, od.loss_tally, static_cast<size_t>(hl->n_elements) * 8,
                             cudaMemcpyDeviceToHost, st));
  if (o->monitor_data)
    XLB_CUDA(cudaMemcpyAsync(o->monitor_data, od.monitor_data, static_cast<size_t>(o->monitor_words) * 8,
                             cudaMemcpyDeviceToHost, st));
  XLB_CUDA(cudaStreamSynchronize(st));
  return XLB_OK;
}

int xlb_measure_fp64_peak(int repeats, double *flops_out, double *ms_out) {
  if (!flops_out) return fail(XLB_EINVAL, "flops_out is null");
  int dev = 0;
  XLB_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  XLB_CUDA(cudaGetDeviceProperties(&prop, dev));
  constexpr int CHAINS = 8;
  const int iters = 1 << 15;
  const int blocks = prop.multiProcessorCount * 8;
  double *d = nullptr;
  XLB_CUDA(cudaMalloc(&d, 64));
  cudaEvent_t e0, e1;
  XLB_CUDA(cudaEventCreate(&e0));
  XLB_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  if (repeats < 1) repeats = 1;
  for (int r = 0; r < repeats + 2; ++r) {
    XLB_CUDA(cudaEventRecord(e0));
    dfma_peak_kernel<CHAINS><<<blocks, 256>>>(d, iters, 1.0000001, 1e-9);
    XLB_CUDA(cudaEventRecord(e1));
    XLB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    XLB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (r >= 2 && ms < best) best = ms;  // first two are warm-up
  }
  XLB_CUDA(cudaGetLastError());
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  const double flops = 2.0 * CHAINS * static_cast<double>(iters) * 256.0 * blocks;
  *flops_out = flops / (best * 1e-3);
  if (ms_out) *ms_out = best;
  return XLB_OK;
}

int xlb_measure_dfma_latency(double *cycles_per_dfma, int n) {
  if (!cycles_per_dfma || n < 4) return fail(XLB_EINVAL, "need room for 4 results");
  double *d = nullptr;
  long long *dc = nullptr;
  XLB_CUDA(cudaMalloc(&d, 64));
  XLB_CUDA(cudaMalloc(&dc, 8));
  int rc = run_latency<1>(d, dc, &cycles_per_dfma[0]);
  if (rc == XLB_OK) rc = run_latency<2>(d, dc, &cycles_per_dfma[1]);
  if (rc == XLB_OK) rc = run_latency<4>(d, dc, &cycles_per_dfma[2]);
  if (rc == XLB_OK) rc = run_latency<8>(d, dc, &cycles_per_dfma[3]);
  cudaFree(d);
  cudaFree(dc);
  return rc;
}

int xlb_selftest_exact_division(const double *divisors, int n_divisors, int mode, int samples_per_thread,
                                uint64_t seed, int exponent_span, uint64_t *mismatches, uint64_t *samples) {
  if (!divisors || n_divisors < 1 || !mismatches || samples_per_thread < 1 || exponent_span < 0 ||
      exponent_span > 900 || (mode != 0 && mode != 1))
    return fail(XLB_EINVAL, "bad argument");
  for (int i = 0; i < n_divisors; ++i) {
    const double b = divisors[i];
    if (mode == 0 && !(b >= 1 && b <= 255 && b == static_cast<double>(static_cast<int>(b))))
      return fail(XLB_EINVAL, "mode 0 takes integer divisors 1..255");
    if (!(b > 0) || b > 1e150 || b < 1e-150) return fail(XLB_EINVAL, "divisor out of range");
  }
  double *d = nullptr;
  unsigned long long *dm = nullptr;
  XLB_CUDA(cudaMalloc(&d, sizeof(double) * n_divisors));
  XLB_CUDA(cudaMalloc(&dm, sizeof(unsigned long long)));
  XLB_CUDA(cudaMemcpy(d, divisors, sizeof(double) * n_divisors, cudaMemcpyHostToDevice));
  XLB_CUDA(cudaMemset(dm, 0, sizeof(unsigned long long)));
  const int rc = strict_selftest_division(d, n_divisors, mode, samples_per_thread, seed, exponent_span, dm, nullptr);
  if (rc != 0) return fail(XLB_ECUDA, cudaGetErrorString(static_cast<cudaError_t>(rc)));
  unsigned long long h = 0;
  XLB_CUDA(cudaMemcpy(&h, dm, sizeof(h), cudaMemcpyDeviceToHost));
  cudaFree(d);
  cudaFree(dm);
  *mismatches = h;
  if (samples) *samples = 148ull * 4 * 256 * static_cast<unsigned long long>(samples_per_thread) * n_divisors;
  return XLB_OK;
}

int xlb_kernel_variant_count(void) {
  int total = 0;
  for (int t = 0; t < kNumTables; ++t) {
    int n = 0;
    kTables[t](&n);
    total += n;
  }
  return total;
}

int xlb_kernel_variant_info(int i, char *name, int name_len, int *regs, int *max_threads) {
  for (int t = 0; t < kNumTables; ++t) {
    int n = 0;
    const Variant *tab = kTables[t](&n);
    if (i < n) {
      const Variant &v = tab[i];
      if (name && name_len > 0) {
        strncpy(name, v.name, name_len - 1);
        name[name_len - 1] = 0;
      }
      cudaFuncAttributes fa;
      cudaError_t e = cudaFuncGetAttributes(&fa, v.func);
      if (regs) *regs = (e == cudaSuccess) ? fa.numRegs : -1;
      if (max_threads) *max_threads = v.threads;
      if (e != cudaSuccess) cudaGetLastError();
      return XLB_OK;
    }
    i -= n;
  }
  return fail(XLB_EINVAL, "variant index out of range");
}

}  // extern "C"
