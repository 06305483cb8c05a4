// C-ABI entry points of the fading hot path: planning, device-resident launch, host-buffer pipeline.
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <cmath>
#include <mutex>
#include <vector>

#include "fading_fused.cuh"
#include "fading_poly64.cuh"
#include "fading_siso.cuh"
#include "fading_tma.cuh"

namespace hb {

// tensor-core spatial product of the large-array path (spatial_gemm.cu)
int launch_spatial_gemm(const double2* S, const float2* z, float2* y, int B, int nrx, int ntx, int T, cudaStream_t st,
                        int z_pitch = 0);

// ---- problem validation / delay groups -------------------------------------------------------------
static int build_delay_table(const hb_fading_problem* p, DelayTable* dt) {
  if (!p) {
    set_error("problem pointer is NULL");
    return HB_ERR_INVALID;
  }
  if (p->batch < 0 || p->num_tx < 1 || p->num_rx < 0 || p->num_samples < 0 || p->max_delay < 0 ||
      p->num_sinusoids < 0) {
    set_error("invalid problem shape (B=%d Ntx=%d Nrx=%d T=%d D=%d N=%d)", p->batch, p->num_tx, p->num_rx,
              p->num_samples, p->max_delay, p->num_sinusoids);
    return HB_ERR_INVALID;
  }
  if (p->num_taps < 1 || p->num_taps > HB_MAX_TAPS) {
    set_error("number of taps %d outside [1, %d]", p->num_taps, HB_MAX_TAPS);
    return p->num_taps < 1 ? HB_ERR_INVALID : HB_ERR_UNSUPPORTED;
  }
  if (!p->tap_delay) {
    set_error("tap_delay is NULL");
    return HB_ERR_INVALID;
  }
  if (p->precision != HB_F32 && p->precision != HB_F64) {
    set_error("unknown precision %d", p->precision);
    return HB_ERR_INVALID;
  }
  if (p->sos_mode < HB_SOS_AUTO || p->sos_mode > HB_SOS_POLY_SISO) {
    set_error("unknown sos_mode %d", p->sos_mode);
    return HB_ERR_INVALID;
  }
  memset(dt, 0, sizeof(*dt));
  dt->num_taps = p->num_taps;
  int g = -1;
  for (int l = 0; l < p->num_taps; ++l) {
    const int d = p->tap_delay[l];
    if (d < 0 || d > p->max_delay) {
      set_error("tap %d: delay %d outside [0, max_delay=%d]", l, d, p->max_delay);
      return HB_ERR_INVALID;
    }
    if (l > 0 && d < p->tap_delay[l - 1]) {
      set_error("tap delays must be ascending (tap %d)", l);
      return HB_ERR_INVALID;
    }
    dt->tap_delay[l] = d;
    if (g < 0 || d != dt->group_delay[g]) {
      ++g;
      dt->group_delay[g] = d;
      dt->group_start[g] = (uint16_t)l;
    }
  }
  dt->num_groups = g + 1;
  dt->group_start[dt->num_groups] = (uint16_t)p->num_taps;
  return HB_OK;
}

static int pick_ntx_template(int n) { return n <= 1 ? 1 : (n <= 2 ? 2 : (n <= 4 ? 4 : 8)); }

struct Plan {
  int mode, tile, P, ntiles, Dpad, ntx_tpl, taps_per_chunk;
  int large_array;  // TMA variant only: z = tap delay lines per transmit antenna, then y = S z on the tensor cores
  int f64poly;      // HB_F64 on the Taylor path (fading_poly64.cuh): FP64 moments + FP64 gather kernel
  int fused;        // large arrays up to 64 x 64: u = S x on the tensor cores with the delay lines on its accumulator (ONE kernel)
  int variant, poly_tile, npoly, threads, large_halo, lin;  // POLY: kernel variant, Taylor window, windows per link, CTA size
  int unstaged;     // DIRECT: x read from global memory, FP64 evaluation (delay spreads no staged tile can hold)
  size_t smem;
  double bound;
  WindowPlan wp;
  TmaPlan tp;  // HB_VARIANT_TMA
};

constexpr size_t kTmaSmemBudget = 226 * 1024;  // dynamic shared memory of the one persistent TMA CTA per SM
constexpr size_t kSmemSoftLimit = 72 * 1024;   // keeps >= 3 CTAs per SM
constexpr size_t kSmemHardLimit = 200 * 1024;  // below the 227 KB per-CTA maximum
constexpr double kPolyTarget = 5e-8;           // truncation bound, relative to the RMS tap gain

static double poly_bound(double u_half, int P, int K) {
  if (u_half <= 0.0) return 0.0;
  double t = 1.0;
  for (int p = 1; p <= P; ++p) t *= u_half / (double)p;
  const double tail = u_half < (double)(P + 1) ? 1.0 / (1.0 - u_half / (double)(P + 1)) : 1e30;
  return sqrt((double)K) * t * tail;
}

static size_t poly_smem(int ntx_tpl, int tile, int Dpad, int G, int P, int nrx) {
  return sizeof(float2) * ((size_t)ntx_tpl * (tile + Dpad) + (size_t)G * P + (size_t)nrx * ntx_tpl);
}

static int window_R(int ntx_tpl) { return ntx_tpl <= 4 ? 8 : 4; }
static int window_split_of(int) { return 1; }  // threads per output row (window_split<NTX>() in fading_window.cuh)

// Sliding-window variant: eligible when the delay walk fits the masks and reads no more shared memory than the
// gather kernel would ((dmax + R) / R loads per output and antenna against G).
static bool window_eligible(const DelayTable& dt, int ntx_tpl) {
  const int R = window_R(ntx_tpl);
  const int dmax = dt.group_delay[dt.num_groups - 1];
  return dmax <= kWindowMaxDelay && dmax + R <= dt.num_groups * R;
}

// Fill pl->wp / tile / threads / smem for the window kernel given pl->poly_tile and pl->P.
static void plan_window(const hb_fading_problem* p, const DelayTable& dt, Plan* pl) {
  const int R = window_R(pl->ntx_tpl);
  const int Tout = p->num_samples + p->max_delay;
  WindowPlan& wp = pl->wp;
  memset(&wp, 0, sizeof(wp));
  wp.num_groups = dt.num_groups;
  const int dmax = dt.group_delay[dt.num_groups - 1];
  wp.nblk = dmax / R + 1;
  for (int g = 0; g < dt.num_groups; ++g) {
    const int d = dt.group_delay[g];
    wp.mask[d / R] |= 1u << (d % R);
    for (int e = std::max(0, d - R + 1); e <= d; ++e) wp.mask[e / R] |= 0x100u << (e % R);
  }
  for (int c = 0; c < wp.nblk; ++c)
    if (wp.mask[c + 1] & 0x100u) wp.mask[c] |= 0x100u << R;
  // CTA tile: the largest of {128, 64, 32} threads x R outputs that divides the Taylor window and is not
  // (much) longer than the frame
  const int split = window_split_of(pl->ntx_tpl);
  int threads = kWindowThreads;
  while (threads > 32 && (pl->poly_tile % (threads / split * R) != 0 || (threads / 2 / split) * R >= Tout)) threads /= 2;
#ifdef HB_ATTRIBUTION
  if (const char* ev = getenv("HB_WINDOW_THREADS")) {  // experiments only
    const int t = atoi(ev);
    if ((t == 32 || t == 64 || t == 128) && pl->poly_tile % (t / split * R) == 0) threads = t;
  }
#endif
  pl->threads = threads;
  pl->tile = threads / split * R;
  pl->large_halo = wp.nblk > kWindowHaloSmall;
  const int PL = kWindowThreads / split + (pl->large_halo ? kWindowHaloLarge : kWindowHaloSmall);
  wp.poly_tile = pl->poly_tile;
  pl->npoly = std::max(1, (Tout + pl->poly_tile - 1) / pl->poly_tile);
  wp.npoly = pl->npoly;
  pl->smem = (size_t)pl->ntx_tpl * R * PL * 8 +
             sizeof(float2) * ((size_t)(dt.num_groups + 1) * pl->P + (size_t)p->num_rx * pl->ntx_tpl);
  pl->Dpad = R * wp.nblk;
  // linear extension of the tap gains over a thread's R outputs: neglected curvature, same normalization as
  // poly_bound (relative to the RMS tap gain)
  const double eps_w = 0.5 * (R - 1) * p->omega_max;
  pl->lin = pl->P >= 3 && pl->P <= 4 && sqrt((double)(p->num_sinusoids + 1)) * eps_w * eps_w * 0.5 <= kPolyTarget;
  if (pl->lin) pl->bound += sqrt((double)(p->num_sinusoids + 1)) * eps_w * eps_w * 0.5;

}

// Persistent TMA-pipelined window kernel (fading_tma.cuh): complex64 frames whose length is a multiple of 16 samples
// (the frame is described to the copy engine as rows of 16 samples), delays below 128 samples, antenna chunks of at
// most 4 (R = 8 outputs per thread), frames of at least two 1024-output tiles.  Sparse delay sets are fine: the walk
// skips empty blocks of 8 delays and loads only the pairs a present delay reads (never more than the gather kernel).
static bool tma_shape_ok(const hb_fading_problem* p, const DelayTable& dt, int ntx_tpl) {
  const int Tout = p->num_samples + p->max_delay;
  const int dmax = dt.group_delay[dt.num_groups - 1];
  const bool zmode = p->num_tx >= 16 && p->num_rx >= 16;
  return !p->io_complex128 && ntx_tpl <= 4 && p->num_samples % 16 == 0 && p->num_samples >= 16 &&
         Tout >= 2 * kTmaTile && dmax / kTmaR + 1 <= kTmaMaxBlocks && (p->num_rx <= 64 || zmode) &&
         (long long)p->num_samples * 8 * p->num_tx < (1ll << 40);
}

static void plan_tma(const hb_fading_problem* p, const DelayTable& dt, Plan* pl) {
  const int Tout = p->num_samples + p->max_delay;
  TmaPlan& tp = pl->tp;
  memset(&tp, 0, sizeof(tp));
  const int dmax = dt.group_delay[dt.num_groups - 1];
  tp.num_groups = dt.num_groups;
  tp.nblk = dmax / kTmaR + 1;
  tp.hrows = (tp.nblk + 1) / 2;
  tp.rows = kTmaTile / 16 + tp.hrows;
  tp.poly_tile = pl->poly_tile;
  tp.npoly = std::max(1, (Tout + pl->poly_tile - 1) / pl->poly_tile);
  tp.ntiles = std::max(1, (Tout + kTmaTile - 1) / kTmaTile);
  tp.coef_stride = (dt.num_groups * pl->P + 1) & ~1;
  const bool zmode = p->num_tx >= 16 && p->num_rx >= 16;  // large arrays: no spatial matrix in the kernel
  tp.s_stride = zmode ? 2 : (p->num_rx * pl->ntx_tpl + 1) & ~1;
  tp.nchunks = (p->num_tx + pl->ntx_tpl - 1) / pl->ntx_tpl;
  tp.stage_bytes = (uint32_t)pl->ntx_tpl * kTmaPlaneBytes;  // one 1024-aligned plane of <= 72 rows per antenna
  tp.coef_bytes = (uint32_t)tp.coef_stride * 8u;
  tp.s_bytes = (uint32_t)tp.s_stride * 8u;
  tp.s_off = (uint32_t)align_up((size_t)(dt.num_groups + 1) * pl->P * 8, 16);
  tp.aux_bytes = (uint32_t)align_up(tp.s_off + tp.s_bytes + (size_t)pl->ntx_tpl * 8, 16);  // + one row: the epilogue prefetches S
  for (int g = 0; g < dt.num_groups; ++g) tp.mask[dt.group_delay[g] / kTmaR] |= 1u << (dt.group_delay[g] % kTmaR);
  auto present = [&](int d) { return d >= 0 && d <= dmax && ((tp.mask[d / kTmaR] >> (d % kTmaR)) & 1u); };
  for (int d = 1; d <= dmax; d += 2) {  // pair (x[m0-d-1], x[m0-d]) entering at odd d
    bool need = false;
    for (int e = d; e <= d + kTmaR; ++e) need = need || present(e);
    if (need) tp.mask[d / kTmaR] |= 0x100u << (d % kTmaR);
  }
  pl->variant = HB_VARIANT_TMA;
  pl->threads = kTmaThreads;
  pl->tile = kTmaTile;
  pl->large_halo = 0;
  pl->npoly = tp.npoly;
  pl->Dpad = 16 * tp.hrows;
  tp.slot_bytes = (uint32_t)align_up((size_t)tp.stage_bytes + tp.aux_bytes, 1024);
  // ring depth: as many slots as fit the 227 KB of one SM (one CTA per SM), at most kTmaMaxSlots
  const size_t ctl = 20 * kTmaMaxSlots + 32;  // barriers, release counters, tile indices, ticket, done flags, retire ptr, lock
  tp.num_slots = (int)std::min<size_t>(kTmaMaxSlots, (kTmaSmemBudget - ctl) / tp.slot_bytes);
  pl->smem = (size_t)tp.num_slots * tp.slot_bytes + ctl;
  const double eps_w = 0.5 * (kTmaR - 1) * p->omega_max;
  const double curv = sqrt((double)(p->num_sinusoids + 1)) * eps_w * eps_w * 0.5;
  pl->lin = pl->P >= 3 && pl->P <= 4 && curv <= kPolyTarget;
  if (pl->lin) pl->bound += curv;
}

static int make_plan_as_asked(const hb_fading_problem* p, const DelayTable& dt, Plan* pl, bool allow_tma) {
  const int Tout = p->num_samples + p->max_delay;
  const int K = p->num_sinusoids + 1;
  pl->Dpad = (p->max_delay + 1) & ~1;
  pl->ntx_tpl = pick_ntx_template(std::min(p->num_tx, 8));
  pl->bound = 0.0;
  const bool f64 = p->precision == HB_F64;
  const int tile_cap = std::max(kThreads, ((Tout + kThreads - 1) / kThreads) * kThreads);

  pl->f64poly = 0;
  pl->unstaged = 0;
  if (f64 && (p->sos_mode == HB_SOS_AUTO || p->sos_mode == HB_SOS_POLY)) {
    // float64 parity mode on the Taylor path (fading_poly64.cuh) when the truncation bound can be held below 1e-14 and
    // (AUTO) the phases of the frame are small enough for the reference's own argument rounding not to matter
    static const int kOrders64[] = {4, 6, 8};
    static const int kTiles64[] = {2048, 1024, 512, 256};
    const bool phase_ok = p->sos_mode == HB_SOS_POLY || p->omega_max * (double)Tout <= kPoly64MaxPhase;
    double best_cost = 1e300;
    for (int P : kOrders64) {
      for (int tile0 : kTiles64) {
        int tile = tile0;
        while (tile > kThreads && tile / 2 >= Tout) tile /= 2;
        const size_t smem = poly64_smem(pl->ntx_tpl, tile, pl->Dpad, dt.num_groups, P, p->num_rx);
        if (smem > (tile > kThreads ? kSmemSoftLimit : kSmemHardLimit)) continue;
        const double bnd = poly_bound(0.5 * p->omega_max * tile, P, K);
        if (bnd > kPolyTarget64) continue;
        const double cost = dt.num_groups * (2.0 * (P - 1) + 4.0 * pl->ntx_tpl) + 200.0 * p->num_taps * K / (double)tile;
        if (phase_ok && cost < best_cost) {
          best_cost = cost;
          pl->mode = HB_SOS_POLY;
          pl->f64poly = 1;
          pl->P = P;
          pl->tile = tile;
          pl->poly_tile = tile;
          pl->bound = bnd;
          pl->smem = smem;
        }
      }
    }
    if (pl->f64poly) {
      pl->variant = HB_VARIANT_GATHER;
      pl->large_array = 0;
      pl->fused = 0;
      pl->threads = kThreads;
      pl->large_halo = 0;
      pl->lin = 0;
      pl->taps_per_chunk = 0;
      pl->ntiles = std::max(1, (Tout + pl->tile - 1) / pl->tile);
      pl->npoly = pl->ntiles;
      return HB_OK;
    }
    if (p->sos_mode == HB_SOS_POLY) {
      set_error("HB_F64 with HB_SOS_POLY: omega_max=%g rad/sample cannot meet the %g bound of the float64 Taylor path",
                p->omega_max, kPolyTarget64);
      return HB_ERR_UNSUPPORTED;
    }
  }
  bool poly = !f64 && p->sos_mode != HB_SOS_DIRECT;
  if (f64 && p->sos_mode == HB_SOS_POLY_FUSED) {
    set_error("HB_F64 parity mode only supports direct evaluation");
    return HB_ERR_UNSUPPORTED;
  }
  pl->variant = HB_VARIANT_GATHER;
  pl->large_array = 0;
  pl->fused = 0;
  pl->poly_tile = 0;
  pl->npoly = 0;
  pl->threads = kThreads;
  pl->large_halo = 0;
  pl->lin = 0;
  // one antenna per side: the time-packed kernel (fading_siso.cuh) on request.  Measured on C5 (profiles/r02_siso.md):
  // 0.56 ms against 0.535 ms for the window kernel and 0.71 ms for the TMA kernel, so AUTO / POLY take the WINDOW kernel
  // for 1 x 1 links (the persistent TMA kernel's 12 warps cannot amortize the walk's bookkeeping over one antenna).
  const bool one_by_one = p->num_tx == 1 && p->num_rx == 1;
  const bool siso = poly && one_by_one && p->sos_mode == HB_SOS_POLY_SISO;
  const bool window = poly && p->sos_mode != HB_SOS_POLY_GATHER && p->sos_mode != HB_SOS_POLY_SISO &&
                      window_eligible(dt, pl->ntx_tpl);
  // large arrays (16 x 16 and more) run the persistent kernel in z mode, chunks of 4 antennas, then the tensor-core GEMM
  const int tpl_tma = (p->num_tx >= 16 && p->num_rx >= 16) ? 4 : pl->ntx_tpl;
  const bool tma_shape = poly && allow_tma && p->sos_mode != HB_SOS_POLY_GATHER && p->sos_mode != HB_SOS_POLY_WINDOW &&
                         p->sos_mode != HB_SOS_POLY_SISO && tma_shape_ok(p, dt, tpl_tma) &&
                         !(one_by_one && p->sos_mode != HB_SOS_POLY_TMA && window_eligible(dt, pl->ntx_tpl));
  if (f64 && p->sos_mode == HB_SOS_POLY_GATHER) {
    set_error("HB_F64 parity mode: direct evaluation or the float64 Taylor path (HB_SOS_AUTO / HB_SOS_POLY)");
    return HB_ERR_UNSUPPORTED;
  }
  if (poly) {
    static const int kOrders[] = {1, 2, 3, 4, 6, 8};
    static const int kTiles[] = {2048, 1024, 512, 256};
    double best_cost = 1e300;
    int best_P = 0, best_tile = 0;
    double best_bound = 0;
    for (int P : kOrders) {
      for (int tile0 : kTiles) {
        // window variant: CTA tiles are 32/64/128 threads x R outputs and must divide the Taylor window, so the
        // window stays a power of two (not longer than the padded frame)
        int tile = std::min(tile0, tile_cap);
        if (window || tma_shape || siso) {
          const int floor_tile = siso ? 512 : kThreads;  // the single-antenna kernel's smallest tile is 512 outputs
          if (tile0 < floor_tile) continue;
          tile = tile0;
          while (tile > floor_tile && tile / 2 >= Tout) tile /= 2;
        }
        if (!window && !tma_shape && poly_smem(pl->ntx_tpl, tile, pl->Dpad, dt.num_groups, P, p->num_rx) > kSmemSoftLimit &&
            tile > kThreads)
          continue;
        const double bnd = poly_bound(0.5 * p->omega_max * tile, P, K);
        if (bnd > kPolyTarget) continue;
        const double cost = dt.num_groups * (2.0 * (P - 1) + 6.0 * pl->ntx_tpl) +
                            60.0 * p->num_taps * K / (double)tile;
        if (cost < best_cost) {
          best_cost = cost;
          best_P = P;
          best_tile = tile;
          best_bound = bnd;
        }
      }
    }
    if (best_P == 0) {
      if (p->sos_mode == HB_SOS_POLY || p->sos_mode == HB_SOS_POLY_GATHER) {
        set_error("HB_SOS_POLY requested but omega_max=%g rad/sample cannot meet the %g bound", p->omega_max,
                  kPolyTarget);
        return HB_ERR_UNSUPPORTED;
      }
      poly = false;
    } else {
      pl->mode = HB_SOS_POLY;
      pl->P = best_P;
      pl->tile = best_tile;
      pl->bound = best_bound;
      pl->poly_tile = best_tile;
      pl->npoly = std::max(1, (Tout + best_tile - 1) / best_tile);
      pl->smem = poly_smem(pl->ntx_tpl, pl->tile, pl->Dpad, dt.num_groups, pl->P, p->num_rx);
      pl->taps_per_chunk = 0;
      const double bound0 = pl->bound;
      // HB_SOS_POLY_FUSED, 16..64 antennas per side, delays within the on-chip history: spatial GEMM first, delay lines
      // on its accumulator, ONE kernel (fading_fused.cuh).  Not what AUTO picks: on B200 the shared-memory pipe (MMA operand
      // reads + history-ring reads) makes it 8 % slower than the two-kernel path (profiles/r02_c4.md).
      const int dmax_f = dt.group_delay[dt.num_groups - 1];
      int siso_tile = std::min(2048, pl->poly_tile);
      while (siso_tile > 512 && siso_tile / 2 >= Tout) siso_tile /= 2;
      const size_t siso_smem = siso_smem_bytes(siso_tile, pl->Dpad, dt.num_groups, pl->P);
      if (siso && pl->poly_tile >= 512 && siso_smem <= kSmemHardLimit) {
        pl->variant = HB_VARIANT_SISO;
        pl->tile = siso_tile;
        pl->threads = kSisoThreads;
        pl->smem = siso_smem;
      } else if (p->sos_mode == HB_SOS_POLY_SISO) {
        set_error("HB_SOS_POLY_SISO needs one antenna per side and a Taylor window of at least 512 samples");
        return HB_ERR_UNSUPPORTED;
      } else if (p->sos_mode == HB_SOS_POLY_FUSED && !p->io_complex128 && p->num_tx >= 16 && p->num_rx >= 16 &&
          p->num_tx <= kGemmMaxAnt && p->num_rx <= kGemmMaxAnt && dmax_f <= kFusedMaxDelay && dt.num_groups <= kFusedMaxGroups &&
          pl->P <= 4 && pl->poly_tile % kGemmTileSamples == 0) {
        pl->fused = 1;
        pl->variant = HB_VARIANT_FUSED;
        pl->tile = kGemmTileSamples;
        pl->threads = kFusedThreads;
        pl->smem = kFusedSmemBytes;
        pl->ntx_tpl = kGemmMaxAnt;
        pl->npoly = std::max(1, (Tout + pl->poly_tile - 1) / pl->poly_tile);
      } else if (tma_shape && pl->poly_tile % kTmaTile == 0) {
        const int tpl0 = pl->ntx_tpl;
        pl->ntx_tpl = tpl_tma;
        plan_tma(p, dt, pl);
        if (pl->tp.num_slots < 4) {  // ring too shallow to hide the loads: keep the other kernels
          pl->bound = bound0;
          pl->variant = HB_VARIANT_GATHER;
          pl->ntx_tpl = tpl0;
          pl->tile = pl->poly_tile;
          pl->npoly = std::max(1, (Tout + pl->poly_tile - 1) / pl->poly_tile);
          pl->threads = kThreads;
          pl->lin = 0;
          pl->Dpad = (p->max_delay + 1) & ~1;
          pl->smem = poly_smem(pl->ntx_tpl, pl->tile, pl->Dpad, dt.num_groups, pl->P, p->num_rx);
        } else {
          // 16 x 16 antennas and more: 8 Nrx Ntx flop per sample are tensor-core work
          pl->large_array = p->num_tx >= 16 && p->num_rx >= 16;
        }
      }
      if (pl->variant == HB_VARIANT_GATHER && window) {
        pl->variant = HB_VARIANT_WINDOW;
        plan_window(p, dt, pl);
      }
    }
  }
  if (!poly) {
    pl->mode = HB_SOS_DIRECT;
    pl->P = 0;
    pl->tile = kThreads;
    const size_t csz = f64 ? sizeof(double2) : sizeof(float2);
    const size_t spsz = f64 ? sizeof(double2) : sizeof(uint2);
    const size_t rsz = f64 ? sizeof(double) : sizeof(float);
    int tpc = (int)std::max<size_t>(1, kDirectParamBytes / (K * spsz));
    tpc = std::min(tpc, p->num_taps);
    pl->taps_per_chunk = tpc;
    pl->smem = csz * ((size_t)pl->ntx_tpl * (pl->tile + pl->Dpad) + (size_t)p->num_rx * pl->ntx_tpl) +
               (size_t)tpc * K * spsz + (size_t)tpc * 2 * rsz;
  }
  // Very long delay spreads: shrink the antenna chunk before giving up.
  while (pl->smem > kSmemHardLimit && pl->ntx_tpl > 1 && pl->variant == HB_VARIANT_GATHER) {
    const int old = pl->ntx_tpl;
    pl->ntx_tpl = old / 2;
    const size_t per_ant = (f64 && pl->mode == HB_SOS_DIRECT ? sizeof(double2) : sizeof(float2)) *
                           ((size_t)(pl->tile + pl->Dpad) + p->num_rx);
    pl->smem -= per_ant * (old - pl->ntx_tpl);
  }
  if (pl->mode == HB_SOS_DIRECT && pl->smem > kSmemHardLimit) {
    // No staged tile can hold this delay spread: per-sample FP64 evaluation with x read from global memory.  Slow (no
    // reuse of x in shared memory) but complete -- the reference handles such links and there is no CPU path behind us.
    pl->unstaged = 1;
    pl->ntx_tpl = pick_ntx_template(std::min(p->num_tx, 8));
    int tpc = (int)std::max<size_t>(1, kDirectParamBytes / (K * sizeof(double2)));
    tpc = std::min(tpc, p->num_taps);
    pl->taps_per_chunk = tpc;
    pl->smem = sizeof(double2) * (size_t)p->num_rx * pl->ntx_tpl + (size_t)tpc * K * sizeof(double2) + (size_t)tpc * 2 * sizeof(double);
  }
  if (pl->variant != HB_VARIANT_TMA && pl->variant != HB_VARIANT_FUSED && pl->smem > kSmemHardLimit) {
    set_error("delay spread of %d samples needs %zu bytes of shared memory per CTA (limit %zu)", p->max_delay,
              pl->smem, kSmemHardLimit);
    return HB_ERR_UNSUPPORTED;
  }
  pl->ntiles = std::max(1, (Tout + pl->tile - 1) / pl->tile);
  return HB_OK;
}

// AUTO / POLY requests whose preferred kernel does not fit (the window kernel keeps an [Nrx x Ntx] spatial matrix and all
// delay groups' coefficients next to its x tile: thousands of receive antennas or very many groups overflow shared
// memory) fall back to the gather kernel, which chunks the antennas, and then to per-sample evaluation, before the
// problem is refused -- the drop-in has no CPU path to hand such a link to.
static int make_plan(const hb_fading_problem* p, const DelayTable& dt, Plan* pl, bool allow_tma = true) {
  int rc = make_plan_as_asked(p, dt, pl, allow_tma);
  if (rc != HB_ERR_UNSUPPORTED || (p->sos_mode != HB_SOS_AUTO && p->sos_mode != HB_SOS_POLY)) return rc;
  hb_fading_problem q = *p;
  if (p->precision == HB_F32) {
    q.sos_mode = HB_SOS_POLY_GATHER;
    Plan alt;
    if (make_plan_as_asked(&q, dt, &alt, false) == HB_OK) {
      *pl = alt;
      return HB_OK;
    }
  }
  if (p->sos_mode == HB_SOS_AUTO) {
    q.sos_mode = HB_SOS_DIRECT;
    Plan alt;
    if (make_plan_as_asked(&q, dt, &alt, false) == HB_OK) {
      *pl = alt;
      return HB_OK;
    }
  }
  return make_plan_as_asked(p, dt, pl, allow_tma);  // restores the first refusal's message
}

static void fill_info(const Plan& pl, const DelayTable& dt, const hb_fading_problem* p,
                      hb_fading_plan_info* info) {
  if (!info) return;
  const int chunks = (p->num_tx + pl.ntx_tpl - 1) / pl.ntx_tpl;
  info->mode = pl.mode;
  info->tile = pl.tile;
  info->poly_order = pl.P;
  info->num_groups = dt.num_groups;
  info->num_tiles = pl.ntiles;
  info->launches = pl.fused ? 2
                            : (pl.mode == HB_SOS_POLY ? 1 : 0) + (pl.large_array ? 1 : chunks) +
                                  (pl.large_array ? ((p->num_rx + 63) / 64) * ((p->num_tx + 63) / 64) : 0);
  info->error_bound = pl.bound;
  info->variant = pl.mode == HB_SOS_POLY ? pl.variant : 0;
  info->poly_tile = pl.mode == HB_SOS_POLY ? pl.poly_tile : pl.tile;
}

template <int P>
static int launch_coef(const FadingArgs& a, const DelayTable& dt, cudaStream_t st) {
  ProfileScope prof(KIND_SOS_COEF, st);
  if (a.ntiles >= 2) {
    // several Taylor windows per link: the rotating form (one sincos per pair and warp, then FP64 rotations).  Windows per
    // warp: as many as keep ~32 warps per SM in flight, at least 8 so that the start-up sincos is amortized.
    const long long bg = (long long)a.B * dt.num_groups;
    int wchunk = a.ntiles;
    const long long want = 32ll * device_sm_count();
    if (bg < want) wchunk = (int)std::max<long long>(std::min<long long>(8, a.ntiles), a.ntiles * bg / want);
    const int nchunk = (a.ntiles + wchunk - 1) / wchunk;
    const long long items = bg * nchunk;
    sos_poly_coef_rot_kernel<P><<<(unsigned)((items + 3) / 4), 128, 0, st>>>(a, dt, wchunk, nchunk);
    HB_CUDA(cudaGetLastError());
    return HB_OK;
  }
  FadingArgs ak = a;
  ak.coef_flat = dt.num_groups < 4;  // few delay groups: one warp per (link, window, group), see the kernel
  const size_t blocks = ak.coef_flat ? ((size_t)a.ntiles * a.B * dt.num_groups + 3) / 4 : (size_t)a.ntiles * a.B;
  sos_poly_coef_kernel<P><<<(unsigned)blocks, 128, 0, st>>>(ak, dt);
  HB_CUDA(cudaGetLastError());
  return HB_OK;
}

static int launch_coef_any(int P, const FadingArgs& a, const DelayTable& dt, cudaStream_t st) {
  switch (P) {
    case 1: return launch_coef<1>(a, dt, st);
    case 2: return launch_coef<2>(a, dt, st);
    case 3: return launch_coef<3>(a, dt, st);
    case 4: return launch_coef<4>(a, dt, st);
    case 6: return launch_coef<6>(a, dt, st);
    case 8: return launch_coef<8>(a, dt, st);
  }
  set_error("polynomial order %d outside the compiled set", P);
  return HB_ERR_UNSUPPORTED;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int encode_fn(EncodeTiledFn* out) {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  });
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return HB_ERR_CUDA;
  }
  *out = fn;
  return HB_OK;
}

// x[B, Ntx, T] complex64 as (32 floats, T / 16 rows, Ntx, B); box = one tile + halo of every antenna of a chunk.
static int make_x_map(const void* x, int B, int ntx, int T, int rows, int ntx_box, CUtensorMap* map) {
  EncodeTiledFn fn;
  if (int e = encode_fn(&fn)) return e;
  const cuuint64_t dims[4] = {32, (cuuint64_t)(T / 16), (cuuint64_t)ntx, (cuuint64_t)B};
  const cuuint64_t strides[3] = {128, (cuuint64_t)T * 8, (cuuint64_t)T * 8 * (cuuint64_t)ntx};
  const cuuint32_t box[4] = {32, (cuuint32_t)rows, (cuuint32_t)ntx_box, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(x), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (T=%d Ntx=%d B=%d rows=%d)", (int)r, T, ntx, B, rows);
    return HB_ERR_CUDA;
  }
  return HB_OK;
}

static int launch_chunk_tma(const Plan& pl, const FadingArgs& a, const TmaPlan& tp, const CUtensorMap& map, int grid,
                            cudaStream_t st) {
  ProfileScope prof(KIND_TDL_POLY, st);
  switch (pl.ntx_tpl) {
    case 1: return launch_tdl_tma<1>(pl.P, pl.lin != 0, a.z_mode != 0, a, tp, map, grid, pl.smem, st);
    case 2: return launch_tdl_tma<2>(pl.P, pl.lin != 0, a.z_mode != 0, a, tp, map, grid, pl.smem, st);
    default: return launch_tdl_tma<4>(pl.P, pl.lin != 0, a.z_mode != 0, a, tp, map, grid, pl.smem, st);
  }
}

static int launch_chunk(const Plan& pl, bool f64, bool io128, const FadingArgs& a, const DelayTable& dt,
                        cudaStream_t st) {
  ProfileScope prof(pl.mode == HB_SOS_POLY ? KIND_TDL_POLY : KIND_TDL_DIRECT, st);
  if (pl.mode == HB_SOS_POLY && pl.f64poly) return launch_tdl_poly64(pl.ntx_tpl, pl.P, io128, a, dt, pl.smem, st);
  if (pl.mode == HB_SOS_POLY && pl.variant == HB_VARIANT_SISO)
    return launch_tdl_siso(pl.P, pl.tile, io128, a, dt, pl.poly_tile, pl.npoly, pl.smem, st);
  if (pl.mode == HB_SOS_POLY && pl.variant == HB_VARIANT_WINDOW) {
    switch (pl.ntx_tpl) {
      case 1: return launch_tdl_window<1>(pl.P, io128, pl.large_halo != 0, pl.lin != 0, a, pl.wp, pl.threads, pl.smem, st);
      case 2: return launch_tdl_window<2>(pl.P, io128, pl.large_halo != 0, pl.lin != 0, a, pl.wp, pl.threads, pl.smem, st);
      case 4: return launch_tdl_window<4>(pl.P, io128, pl.large_halo != 0, pl.lin != 0, a, pl.wp, pl.threads, pl.smem, st);
      default: return launch_tdl_window<8>(pl.P, io128, pl.large_halo != 0, pl.lin != 0, a, pl.wp, pl.threads, pl.smem, st);
    }
  }
  if (pl.mode == HB_SOS_POLY) {
    switch (pl.ntx_tpl) {
      case 1: return launch_tdl_poly<1>(pl.P, io128, a, dt, pl.smem, st);
      case 2: return launch_tdl_poly<2>(pl.P, io128, a, dt, pl.smem, st);
      case 4: return launch_tdl_poly<4>(pl.P, io128, a, dt, pl.smem, st);
      default: return launch_tdl_poly<8>(pl.P, io128, a, dt, pl.smem, st);
    }
  }
  switch (pl.ntx_tpl) {
    case 1: return launch_tdl_direct<1>(f64, io128, pl.unstaged != 0, a, dt, pl.taps_per_chunk, pl.smem, st);
    case 2: return launch_tdl_direct<2>(f64, io128, pl.unstaged != 0, a, dt, pl.taps_per_chunk, pl.smem, st);
    case 4: return launch_tdl_direct<4>(f64, io128, pl.unstaged != 0, a, dt, pl.taps_per_chunk, pl.smem, st);
    default: return launch_tdl_direct<8>(f64, io128, pl.unstaged != 0, a, dt, pl.taps_per_chunk, pl.smem, st);
  }
}

// Enqueue one batched propagation on `st`; all pointers are device pointers.
static int propagate_device(const hb_fading_problem* p, const DelayTable& dt, const Plan& pl, const void* x,
                            void* y, cudaStream_t st) {
  const int Tout = p->num_samples + p->max_delay;
  if (p->batch == 0 || Tout == 0 || p->num_rx == 0) return HB_OK;
  FadingArgs a;
  memset(&a, 0, sizeof(a));
  a.x = x;
  a.y = y;
  a.omega = p->omega;
  a.phi = p->phi;
  a.amp = p->amp;
  a.spatial = reinterpret_cast<const double2*>(p->spatial);
  a.B = p->batch;
  a.ntx = p->num_tx;
  a.nrx = p->num_rx;
  a.T = p->num_samples;
  a.D = p->max_delay;
  a.L = p->num_taps;
  a.K = p->num_sinusoids + 1;
  a.tile = pl.tile;
  a.ntiles = pl.ntiles;
  a.Dpad = pl.Dpad;
#ifdef HB_ATTRIBUTION
  if (const char* ev = getenv("HB_DBG")) a.dbg = atoi(ev);
#endif
  if ((size_t)a.ntiles * a.B * std::max(1, dt.num_groups) > 0x7fffffffull * 4ull || (size_t)a.ntiles * a.B > 0x7fffffffull) {
    set_error("grid of %zu CTAs exceeds the launch limit; split the batch", (size_t)a.ntiles * a.B);
    return HB_ERR_UNSUPPORTED;
  }
  float2* coef = nullptr;
  const bool use_tma = pl.mode == HB_SOS_POLY && pl.variant == HB_VARIANT_TMA;
  a.coef_stride = (dt.num_groups * pl.P + 1) & ~1;
  if (pl.mode == HB_SOS_POLY) {
    // one allocation: coefficients, then (TMA variant) the chunked FP32 spatial matrices
    const size_t coef_bytes = align_up((pl.f64poly ? sizeof(double2) : sizeof(float2)) * (size_t)a.B * pl.npoly * a.coef_stride, 256);
    const size_t s_bytes = use_tma ? align_up(sizeof(float2) * (size_t)a.B * pl.tp.nchunks * pl.tp.s_stride, 256) : 0;
    const size_t c_bytes = use_tma ? sizeof(unsigned int) * (size_t)pl.tp.nchunks : 0;
    HB_CUDA(cudaMallocAsync((void**)&coef, coef_bytes + s_bytes + c_bytes, st));
    a.coef = coef;
    a.spatial32 = nullptr;
    if (use_tma) {
      a.spatial32 = reinterpret_cast<float2*>(reinterpret_cast<char*>(coef) + coef_bytes);
      a.s32_tpl = pl.large_array ? 0 : pl.ntx_tpl;  // z mode: the slot is copied but never read, K1 leaves it alone
      a.s32_stride = pl.tp.s_stride;
      a.tile_counters = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(coef) + coef_bytes + s_bytes);
      a.num_counters = pl.tp.nchunks;
    }
    FadingArgs ac = a;  // K1 runs over the Taylor windows, which may span several CTA tiles
    ac.tile = pl.poly_tile;
    ac.ntiles = pl.npoly;
    int ek;
    if (pl.f64poly) {
      ProfileScope prof(KIND_SOS_COEF, st);
      ek = launch_coef64(pl.P, ac, dt, st);
    } else {
      ek = launch_coef_any(pl.P, ac, dt, st);
    }
    if (int e = ek) {
      cudaFreeAsync(coef, st);
      return e;
    }
  }
  if (pl.fused) {
    FusedArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.S = a.spatial;
    fa.x = reinterpret_cast<const float2*>(x);
    fa.y = reinterpret_cast<float2*>(y);
    fa.coef = coef;
    fa.B = a.B;
    fa.T = a.T;
    fa.D = a.D;
    fa.ntx = a.ntx;
    fa.nrx = a.nrx;
    fa.poly_tile = pl.poly_tile;
    fa.npoly = pl.npoly;
    fa.coef_stride = a.coef_stride;
    fa.num_groups = dt.num_groups;
    for (int g = 0; g < dt.num_groups; ++g) fa.group_delay[g] = dt.group_delay[g];
    const int rcf = launch_fused_gemm_tdl(pl.P, fa, st);
    cudaFreeAsync(coef, st);
    return rcf;
  }
  int rc = HB_OK;
  CUtensorMap xmap;
  TmaPlan tp;
  int tma_grid = 0;
  float2* zbuf = nullptr;
  if (use_tma && pl.large_array) {
    // even row pitch: every z row starts 16-byte aligned, so the kernel stores pairs of samples (Tout itself is odd for
    // C4: 16 384 + 115)
    a.ypitch = (Tout + 1) & ~1;
    const cudaError_t ce = cudaMallocAsync((void**)&zbuf, sizeof(float2) * (size_t)a.B * a.ntx * a.ypitch, st);
    if (ce != cudaSuccess) {
      if (coef) cudaFreeAsync(coef, st);
      return cuda_fail(ce, "cudaMallocAsync(z workspace of the large-array path)");
    }
    a.y = zbuf;
    a.z_mode = 1;
  }
  if (use_tma) {
    tp = pl.tp;
    tp.total_tiles = a.B * tp.ntiles * (pl.large_array ? tp.nchunks : 1);  // z mode: every chunk in ONE launch
    tma_grid = std::min(tp.total_tiles, persistent_sm_count());
    rc = make_x_map(x, a.B, a.ntx, a.T, tp.rows, 1, &xmap);
  }
  for (int tx0 = 0; tx0 < (a.z_mode ? 1 : p->num_tx) && rc == HB_OK; tx0 += pl.ntx_tpl) {
    a.tx0 = tx0;
    a.ntx_chunk = std::min(pl.ntx_tpl, p->num_tx - tx0);
    a.accumulate = tx0 > 0 && !a.z_mode;
    if (use_tma) {
      tp.chunk = tx0 / pl.ntx_tpl;
      tp.tile_counter = a.tile_counters + tp.chunk;
      rc = launch_chunk_tma(pl, a, tp, xmap, tma_grid, st);
    } else {
      rc = launch_chunk(pl, p->precision == HB_F64, p->io_complex128 != 0, a, dt, st);
    }
  }
  if (zbuf) {
    if (rc == HB_OK)
      rc = launch_spatial_gemm(a.spatial, zbuf, reinterpret_cast<float2*>(y), a.B, a.nrx, a.ntx, Tout, st, a.ypitch);
    cudaFreeAsync(zbuf, st);
  }
  if (coef) cudaFreeAsync(coef, st);
  return rc;
}


}  // namespace hb

using namespace hb;

extern "C" {

int hb_fading_plan(const hb_fading_problem* p, hb_fading_plan_info* info) {
  DelayTable dt;
  if (int e = build_delay_table(p, &dt)) return e;
  Plan pl;
  if (int e = make_plan(p, dt, &pl)) return e;
  fill_info(pl, dt, p, info);
  return HB_OK;
}

int hb_fading_propagate(const hb_fading_problem* p, const void* x, void* y, void* stream,
                        hb_fading_plan_info* info) {
  DelayTable dt;
  if (int e = build_delay_table(p, &dt)) return e;
  Plan pl;
  if (int e = make_plan(p, dt, &pl, (reinterpret_cast<uintptr_t>(x) & 15) == 0)) return e;
  fill_info(pl, dt, p, info);
  if (int e = require_device()) return e;
  if (p->batch > 0 && (!x || !y || !p->omega || !p->phi || !p->amp || !p->spatial)) {
    set_error("NULL device pointer in fading problem");
    return HB_ERR_INVALID;
  }
  return propagate_device(p, dt, pl, x, y, (cudaStream_t)stream);
}

int hb_fading_propagate_host(const hb_fading_problem* p, const void* x, void* y, int32_t chunk_links,
                             hb_fading_plan_info* info) {
  DelayTable dt;
  if (int e = build_delay_table(p, &dt)) return e;
  Plan pl;
  if (int e = make_plan(p, dt, &pl)) return e;
  fill_info(pl, dt, p, info);
  if (int e = require_device()) return e;
  const int Tout = p->num_samples + p->max_delay;
  if (p->batch == 0 || Tout == 0 || p->num_rx == 0) return HB_OK;
  if (!x || !y || !p->omega || !p->phi || !p->amp || !p->spatial) {
    set_error("NULL host pointer in fading problem");
    return HB_ERR_INVALID;
  }
  const size_t esz = p->io_complex128 ? 16 : 8;
  const int K = p->num_sinusoids + 1;
  const size_t x_link = esz * (size_t)p->num_tx * p->num_samples;
  const size_t y_link = esz * (size_t)p->num_rx * Tout;
  const size_t om_link = sizeof(double) * (size_t)p->num_taps * K;
  const size_t am_link = sizeof(double) * (size_t)p->num_taps * 2;
  const size_t s_link = 16 * (size_t)p->num_rx * p->num_tx;
  int chunk = chunk_links;
  if (chunk <= 0) {
    const size_t target = 48u << 20;  // ~48 MB of input per chunk keeps the copy engines busy
    chunk = (int)std::max<size_t>(1, target / std::max<size_t>(1, x_link + y_link));
    // at least kSlots chunks so that copies and kernels overlap -- unless the whole batch is a few megabytes, where
    // the extra launches cost more than the overlap hides (the batched drop runner's rounds of short frames)
    if ((x_link + y_link) * (size_t)p->batch > (8u << 20))
      chunk = std::min(chunk, std::max(1, (p->batch + kSlots - 1) / kSlots));
  }
  chunk = std::min(chunk, p->batch);
  // per-slot device layout: x | y | omega | phi | amp | spatial
  const size_t off_x = 0;
  const size_t off_y = align_up(off_x + x_link * chunk, 256);
  const size_t off_om = align_up(off_y + y_link * chunk, 256);
  const size_t off_ph = align_up(off_om + om_link * chunk, 256);
  const size_t off_am = align_up(off_ph + om_link * chunk, 256);
  const size_t off_s = align_up(off_am + am_link * chunk, 256);
  const size_t total = align_up(off_s + s_link * chunk, 256);

  std::lock_guard<std::mutex> lock(g_pipe.mu);
  if (int e = pipe_prepare(total)) return e;
  int rc = HB_OK;
  int ci = 0;
  for (int b0 = 0; b0 < p->batch && rc == HB_OK; b0 += chunk, ++ci) {
    const int nb = std::min(chunk, p->batch - b0);
    const int s = ci % kSlots;
    cudaStream_t st = g_pipe.st[s];
    char* base = (char*)g_pipe.buf[s];
    // stream order on slot s guarantees the previous D2H of this slot finished before we overwrite it
    hb_fading_problem q = *p;
    q.batch = nb;
    q.omega = (const double*)(base + off_om);
    q.phi = (const double*)(base + off_ph);
    q.amp = (const double*)(base + off_am);
    q.spatial = base + off_s;
#define HB_TRY(call)                          \
  do {                                        \
    cudaError_t _e = (call);                  \
    if (_e != cudaSuccess) {                  \
      rc = cuda_fail(_e, #call);              \
      break;                                  \
    }                                         \
  } while (0)
    do {
      HB_TRY(cudaMemcpyAsync(base + off_om, (const char*)p->omega + om_link * b0, om_link * nb,
                             cudaMemcpyHostToDevice, st));
      HB_TRY(cudaMemcpyAsync(base + off_ph, (const char*)p->phi + om_link * b0, om_link * nb,
                             cudaMemcpyHostToDevice, st));
      HB_TRY(cudaMemcpyAsync(base + off_am, (const char*)p->amp + am_link * b0, am_link * nb,
                             cudaMemcpyHostToDevice, st));
      HB_TRY(cudaMemcpyAsync(base + off_s, (const char*)p->spatial + s_link * b0, s_link * nb,
                             cudaMemcpyHostToDevice, st));
      HB_TRY(cudaMemcpyAsync(base + off_x, (const char*)x + x_link * b0, x_link * nb, cudaMemcpyHostToDevice,
                             st));
      rc = propagate_device(&q, dt, pl, base + off_x, base + off_y, st);
      if (rc != HB_OK) break;
      HB_TRY(cudaMemcpyAsync((char*)y + y_link * b0, base + off_y, y_link * nb, cudaMemcpyDeviceToHost, st));
    } while (0);
#undef HB_TRY
  }
  for (int s = 0; s < kSlots; ++s) {
    cudaError_t e = cudaStreamSynchronize(g_pipe.st[s]);
    if (e != cudaSuccess && rc == HB_OK) rc = cuda_fail(e, "cudaStreamSynchronize(pipeline)");
  }
  return rc;
}

int hb_fading_sinc_taps(const double* delay_samples, int32_t num_taps, int32_t half_width, double kaiser_beta,
                        int32_t capacity, int32_t* out_delay, double* out_weight, int32_t* out_source, int32_t* num_out) {
  if (!delay_samples || !out_delay || !out_weight || !out_source || !num_out || num_taps < 0 || half_width < 1 ||
      half_width > 64 || !(kaiser_beta >= 0.0)) {
    set_error("invalid windowed-sinc expansion request (taps=%d half_width=%d beta=%g)", num_taps, half_width, kaiser_beta);
    return HB_ERR_INVALID;
  }
  struct Tap {
    int32_t delay, source;
    double weight;
  };
  std::vector<Tap> taps;
  const double i0b = std::cyl_bessel_i(0.0, kaiser_beta);
  for (int l = 0; l < num_taps; ++l) {
    const double tau = delay_samples[l];
    if (!(tau >= 0.0) || tau > 1e9) {
      set_error("tap %d: delay of %g samples is not a finite non-negative number", l, tau);
      return HB_ERR_INVALID;
    }
    const long long fl = (long long)floor(tau);
    for (long long j = fl - half_width + 1; j <= fl + half_width; ++j) {
      if (j < 0) continue;  // non-causal precursor: dropped (documented truncation)
      const double u = (double)j - tau;
      if (fabs(u) >= (double)half_width) continue;
      const double sinc = u == floor(u) ? (u == 0.0 ? 1.0 : 0.0) : sin(M_PI * u) / (M_PI * u);  // exact zeros at integers
      const double r = u / (double)half_width;
      const double g = sinc * std::cyl_bessel_i(0.0, kaiser_beta * sqrt(1.0 - r * r)) / i0b;
      if (g == 0.0) continue;
      taps.push_back({(int32_t)j, (int32_t)l, g});
    }
  }
  std::stable_sort(taps.begin(), taps.end(), [](const Tap& a, const Tap& b) { return a.delay < b.delay; });
  *num_out = (int32_t)taps.size();
  if ((int)taps.size() > capacity) {
    set_error("windowed-sinc expansion needs %zu taps, capacity is %d (HB_MAX_TAPS = %d per launch)", taps.size(), capacity,
              HB_MAX_TAPS);
    return HB_ERR_UNSUPPORTED;
  }
  for (size_t t = 0; t < taps.size(); ++t) {
    out_delay[t] = taps[t].delay;
    out_weight[t] = taps[t].weight;
    out_source[t] = taps[t].source;
  }
  return HB_OK;
}

int hb_fading_state(const hb_fading_problem* p, void* h, int32_t* group_delay_out, void* stream) {
  DelayTable dt;
  if (int e = build_delay_table(p, &dt)) return e;
  if (group_delay_out)
    for (int g = 0; g < dt.num_groups; ++g) group_delay_out[g] = dt.group_delay[g];
  if (p->batch == 0 || p->num_samples == 0) return HB_OK;
  if (int e = require_device()) return e;
  if (!h || !p->omega || !p->phi || !p->amp) {
    set_error("NULL device pointer in fading state request");
    return HB_ERR_INVALID;
  }
  FadingArgs a;
  memset(&a, 0, sizeof(a));
  a.y = h;
  a.omega = p->omega;
  a.phi = p->phi;
  a.amp = p->amp;
  a.B = p->batch;
  a.T = p->num_samples;
  a.L = p->num_taps;
  a.K = p->num_sinusoids + 1;
  a.ntiles = (p->num_samples + kThreads - 1) / kThreads;
  const size_t blocks = (size_t)a.ntiles * dt.num_groups * a.B;
  if (blocks > 0x7fffffffull) {
    set_error("state grid of %zu CTAs exceeds the launch limit; split the batch", blocks);
    return HB_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  ProfileScope prof(KIND_SOS_STATE, st);
  const bool f64 = p->precision == HB_F64, io128 = p->io_complex128 != 0;
  if (f64 && io128) sos_state_kernel<double, double2><<<(unsigned)blocks, kThreads, 0, st>>>(a, dt);
  else if (f64) sos_state_kernel<double, float2><<<(unsigned)blocks, kThreads, 0, st>>>(a, dt);
  else if (io128) sos_state_kernel<float, double2><<<(unsigned)blocks, kThreads, 0, st>>>(a, dt);
  else sos_state_kernel<float, float2><<<(unsigned)blocks, kThreads, 0, st>>>(a, dt);
  HB_CUDA(cudaGetLastError());
  return HB_OK;
}

}  // extern "C"
