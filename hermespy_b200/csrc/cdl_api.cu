// C-ABI entry points of the cluster-delay-line hot path.
#include <math.h>

#include <algorithm>
#include <vector>

#include "cdl_kernels.cuh"
#include "cdl_umma.cuh"
#include "cdl_umma_bf16.cuh"

namespace hb {

struct CdlPlan {
  int mode, tile, P, ntiles, Dpad, nrx_tpl, R, threads, max_group_terms;
  int variant;  // hb_cdl_variant actually used (POLY only)
  int ptile, nwin;  // Taylor window (multiple of tile) and windows per frame
  size_t smem;
  double bound;
};

constexpr double kCdlPolyTarget = 5e-8;

// Terms of one link grouped by delay index (stable: the summation order inside a group is the term order).
static int group_terms(const std::vector<int>& delay, int max_delay, CdlTable* tb) {
  const int Rt = (int)delay.size();
  for (int t = 0; t < Rt; ++t) {
    if (delay[t] < 0 || delay[t] > max_delay || delay[t] > 65535) {
      set_error("ray term %d: delay index %d outside [0, max_delay=%d]", t, delay[t], max_delay);
      return HB_ERR_INVALID;
    }
    tb->term_delay[t] = (uint16_t)delay[t];
  }
  std::vector<int> order(Rt);
  for (int t = 0; t < Rt; ++t) order[t] = t;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return delay[a] < delay[b]; });
  int g = -1;
  for (int c = 0; c < Rt; ++c) {
    const int t = order[c];
    if (g < 0 || delay[t] != tb->group_delay[g]) {
      if (g + 1 >= kCdlMaxGroups) {
        set_error("more than %d distinct delay indices", kCdlMaxGroups);
        return HB_ERR_UNSUPPORTED;
      }
      ++g;
      tb->group_delay[g] = delay[t];
      tb->group_start[g] = (uint16_t)c;
    }
    tb->term_order[c] = (uint16_t)t;
  }
  tb->num_groups = g + 1;
  tb->group_start[tb->num_groups] = (uint16_t)Rt;
  return HB_OK;
}

static int max_group_terms_of(const CdlTable& tb) {
  int m = 1;
  for (int g = 0; g < tb.num_groups; ++g) m = std::max(m, (int)tb.group_start[g + 1] - (int)tb.group_start[g]);
  return m;
}

// Host-side tables of a problem: the launch-uniform one (kernel parameter space) and, for heterogeneous batches
// (hb_cdl_problem.link_term_delay), one table per link padded to the batch's largest group count.
struct CdlTables {
  CdlTable tb;
  std::vector<CdlTable> links;  // empty: uniform batch
  std::vector<double> los_amp;  // [B] with `links`
  int max_group_terms = 1;
};

static int build_cdl_table(const hb_cdl_problem* p, CdlTables* out) {
  CdlTable* tb = &out->tb;
  if (!p) {
    set_error("problem pointer is NULL");
    return HB_ERR_INVALID;
  }
  if (p->batch < 0 || p->num_tx < 1 || p->num_rx < 1 || p->num_samples < 0 || p->max_delay < 0 || p->num_terms < 0) {
    set_error("invalid CDL problem shape (B=%d Ntx=%d Nrx=%d T=%d D=%d terms=%d)", p->batch, p->num_tx, p->num_rx,
              p->num_samples, p->max_delay, p->num_terms);
    return HB_ERR_INVALID;
  }
  const int Rt = p->num_terms + (p->line_of_sight ? 1 : 0);
  if (Rt < 1 || Rt > kCdlMaxTerms) {
    set_error("number of ray terms %d outside [1, %d]", Rt, kCdlMaxTerms);
    return Rt < 1 ? HB_ERR_INVALID : HB_ERR_UNSUPPORTED;
  }
  const bool per_link = p->link_term_delay != nullptr && p->batch > 0;
  if (p->num_terms > 0 && !p->term_delay && !per_link) {
    set_error("term_delay is NULL");
    return HB_ERR_INVALID;
  }
  if (p->precision != HB_F32 && p->precision != HB_F64) {
    set_error("unknown precision %d", p->precision);
    return HB_ERR_INVALID;
  }
  if (!(p->sampling_rate > 0.0) || p->carrier_frequency < 0.0) {
    set_error("sampling rate must be positive and carrier frequency non-negative");
    return HB_ERR_INVALID;
  }
  if (p->element_mode < HB_ELEMENTS_IDEAL || p->element_mode > HB_ELEMENTS_PER_ELEMENT) {
    set_error("unknown element_mode %d", p->element_mode);
    return HB_ERR_INVALID;
  }
  memset(tb, 0, sizeof(*tb));
  out->links.clear();
  out->los_amp.clear();
  std::vector<int> delay(Rt);
  if (!per_link) {
    for (int t = 0; t < p->num_terms; ++t) delay[t] = p->term_delay[t];
    if (p->line_of_sight) delay[Rt - 1] = p->los_delay;
    if (int e = group_terms(delay, p->max_delay, tb)) return e;
    tb->num_terms = Rt;
    tb->has_los = p->line_of_sight ? 1 : 0;
    out->max_group_terms = max_group_terms_of(*tb);
    return HB_OK;
  }
  out->links.resize((size_t)p->batch);
  out->los_amp.assign((size_t)p->batch, p->los_amplitude);
  int gmax = 0;
  out->max_group_terms = 1;
  for (int b = 0; b < p->batch; ++b) {
    CdlTable* lt = &out->links[(size_t)b];
    memset(lt, 0, sizeof(*lt));
    for (int t = 0; t < p->num_terms; ++t) delay[t] = p->link_term_delay[(size_t)b * p->num_terms + t];
    if (p->line_of_sight) {
      delay[Rt - 1] = p->link_los_delay ? p->link_los_delay[b] : p->los_delay;
      if (p->link_los_amplitude) out->los_amp[(size_t)b] = p->link_los_amplitude[b];
    }
    if (int e = group_terms(delay, p->max_delay, lt)) return e;
    lt->num_terms = Rt;
    lt->has_los = p->line_of_sight ? 1 : 0;
    gmax = std::max(gmax, lt->num_groups);
    out->max_group_terms = std::max(out->max_group_terms, max_group_terms_of(*lt));
  }
  for (CdlTable& lt : out->links) {  // empty groups up to the batch maximum: zero moments, no terms
    for (int g = lt.num_groups + 1; g <= gmax; ++g) lt.group_start[g] = (uint16_t)Rt;
    lt.num_groups = gmax;
  }
  *tb = out->links[0];  // shapes (terms, groups, LOS slot) are what the kernels read from the launch table
  return HB_OK;
}

static size_t cdl_smem(int tile, int Dpad, int G, int P, int nrx_tpl) {
  return sizeof(float2) * ((size_t)kCdlTxChunk * (tile + Dpad) + (size_t)G * P * nrx_tpl * kCdlTxChunk);
}

// Truncation bound of a P-term Taylor expansion over windows of `tile` samples.
static double cdl_poly_bound(const hb_cdl_problem* p, int tile, int P, int max_group_terms) {
  const double w_max = 2.0 * M_PI * p->max_speed * p->carrier_frequency / kSpeedOfLight / p->sampling_rate;
  const double u_half = 0.5 * w_max * tile;
  if (u_half <= 0.0) return 0.0;
  double t = 1.0;
  for (int k = 1; k <= P; ++k) t *= u_half / (double)k;
  const double tail = u_half < P + 1 ? 1.0 / (1.0 - u_half / (P + 1)) : 1e30;
  return sqrt((double)max_group_terms) * t * tail;
}
// smallest P <= 4 meeting the target (0: none)
static int cdl_poly_order(const hb_cdl_problem* p, int tile, int max_group_terms, double* bound_out) {
  for (int cand = 1; cand <= 4; ++cand) {
    const double bound = cdl_poly_bound(p, tile, cand, max_group_terms);
    if (bound <= kCdlPolyTarget) {
      *bound_out = bound;
      return cand;
    }
  }
  return 0;
}

static int make_cdl_plan(const hb_cdl_problem* p, const CdlTables& tabs, CdlPlan* pl) {
  const CdlTable& tb = tabs.tb;
  const int Tout = p->num_samples + p->max_delay;
  pl->nrx_tpl = p->num_rx <= 1 ? 1 : (p->num_rx <= 2 ? 2 : (p->num_rx <= 4 ? 4 : 8));
  pl->threads = 128;
  pl->bound = 0.0;
  pl->variant = HB_CDL_VARIANT_GATHER;
  pl->max_group_terms = tabs.max_group_terms;
  if (p->variant < HB_CDL_VARIANT_AUTO || p->variant > HB_CDL_VARIANT_UMMA_BF16) {
    set_error("unknown CDL variant %d", p->variant);
    return HB_ERR_INVALID;
  }
  bool poly = false;
  if (p->precision == HB_F32 && p->variant != HB_CDL_VARIANT_GATHER) {
    // tensor-core kernel (cdl_umma.cuh): the Taylor window is 128 MT samples, MT fixed by the accumulator width, so the
    // order is searched with the window it implies.  AUTO takes it from 8 transmit antennas (below, K is too short to pay
    // for the operand staging).
    const bool want = p->variant == HB_CDL_VARIANT_UMMA || p->variant == HB_CDL_VARIANT_UMMA_BF16 || p->num_tx >= 8;
    const int Dpad = (p->max_delay + 7) & ~7;
    // Joint choice of the Taylor order P and the window (1, 2, 4 or 8 K6 tiles): the K6 cost does not depend on either
    // (N is padded to 16 columns), the moment kernel's is proportional to windows x P.
    int best_cost = 1 << 30;
    const bool bf16 = p->variant == HB_CDL_VARIANT_UMMA_BF16;  // BF16x3 form: 256-sample tiles, 8 antennas per K stage
    for (int cand = 1; cand <= 4 && want; ++cand) {
      if (bf16 && !cb_eligible(pl->nrx_tpl, cand)) continue;
      const int tile = bf16 ? kCbTile : cu_tile(pl->nrx_tpl, cand);
      const size_t smem = bf16 ? cb_smem_bytes(pl->nrx_tpl, cand, Dpad, tb.num_groups)
                               : cu_smem_bytes(pl->nrx_tpl, cand, Dpad, tb.num_groups);
      if (smem > 226 * 1024 || tile + Dpad >= 16384) continue;  // both operand images of a K stage, two slots
      for (int tw = 8; tw >= 1; tw >>= 1) {
        const double bound = cdl_poly_bound(p, tile * tw, cand, pl->max_group_terms);
        if (bound > kCdlPolyTarget) continue;
        const int nwin = std::max(1, (Tout + tile * tw - 1) / (tile * tw));
        if (nwin * cand < best_cost) {
          best_cost = nwin * cand;
          pl->mode = HB_SOS_POLY;
          pl->variant = bf16 ? HB_CDL_VARIANT_UMMA_BF16 : HB_CDL_VARIANT_UMMA;
          pl->P = cand;
          pl->tile = tile;
          pl->ptile = tile * tw;
          pl->Dpad = Dpad;
          pl->smem = smem;
          pl->R = 0;
          pl->bound = bound;
          poly = true;
        }
        break;  // smaller windows of this order only cost more
      }
    }
    if (!poly && (p->variant == HB_CDL_VARIANT_UMMA || p->variant == HB_CDL_VARIANT_UMMA_BF16)) {
      set_error("the tensor-core CDL kernel does not take this problem (Doppler too fast for four Taylor terms, or the "
                "operand images of %d delay groups and a %d-sample delay halo exceed the shared memory of one SM)",
                tb.num_groups, Dpad);
      return HB_ERR_UNSUPPORTED;
    }
  }
  if (!poly && p->precision == HB_F32) {
    // FP32-pipe kernel: tile = 128 R outputs; fast links (whose four-term Taylor bound fails on the 512-sample tile) get
    // shorter tiles before the planner gives the problem to the per-ray path
    pl->Dpad = (p->max_delay + 1) & ~1;
    for (int R = pl->nrx_tpl <= 4 ? 4 : 2; R >= 1 && !poly; R >>= 1) {
      const int tile = pl->threads * R;
      double bound = 0.0;
      const int P = cdl_poly_order(p, tile, pl->max_group_terms, &bound);
      const size_t smem = P ? cdl_smem(tile, pl->Dpad, tb.num_groups, P, pl->nrx_tpl) : 0;
      if (P && smem <= 200 * 1024) {
        pl->mode = HB_SOS_POLY;
        pl->R = R;
        pl->tile = tile;
        pl->smem = smem;
        pl->P = P;
        pl->bound = bound;
        poly = true;
      }
    }  // none: Doppler too fast for four Taylor terms on 128 samples, or delay spread too long for shared memory: per-ray path
  }
  if (!poly) {
    pl->mode = HB_SOS_DIRECT;
    pl->P = 0;
    pl->tile = 128;
    pl->Dpad = (p->max_delay + 1) & ~1;
    pl->smem = 0;
  }
  pl->ntiles = std::max(1, (Tout + pl->tile - 1) / pl->tile);
  if ((pl->variant != HB_CDL_VARIANT_UMMA && pl->variant != HB_CDL_VARIANT_UMMA_BF16) || pl->mode != HB_SOS_POLY) pl->ptile = pl->tile;
  pl->nwin = std::max(1, (Tout + pl->ptile - 1) / pl->ptile);
  return HB_OK;
}

static void fill_cdl_info(const CdlPlan& pl, const CdlTable& tb, const hb_cdl_problem* p, hb_cdl_plan_info* info) {
  if (!info) return;
  info->mode = pl.mode;
  info->tile = pl.tile;
  info->poly_order = pl.P;
  info->num_groups = tb.num_groups;
  info->num_tiles = pl.ntiles;
  info->launches = pl.mode == HB_SOS_POLY ? 2 + (p->num_rx + pl.nrx_tpl - 1) / pl.nrx_tpl : 2;
  info->error_bound = pl.bound;
  info->variant = pl.mode == HB_SOS_POLY ? pl.variant : 0;
  info->poly_tile = pl.ptile;
}

template <int P>
static int launch_moments(const CdlArgs& a, const CdlTable& tb, cudaStream_t st) {
  ProfileScope prof(KIND_CDL_RAYS, st);
  const size_t smem = sizeof(float2) * (size_t)kMomTerms * kMomWin * P + sizeof(double2) * (size_t)kMomTerms * a.rank * (a.nrx + a.ntx);
  if (smem <= 48 * 1024) {  // one CTA per (link, delay group), all Taylor windows at once
    if (a.link_tab) cdl_moment_all_kernel<P, true><<<(unsigned)((size_t)a.B * tb.num_groups), 128, smem, st>>>(a, tb);
    else cdl_moment_all_kernel<P, false><<<(unsigned)((size_t)a.B * tb.num_groups), 128, smem, st>>>(a, tb);
  } else {  // arrays beyond ~80 elements per side: the per-window kernel reads the steering phases from global memory
    const size_t blocks = (size_t)a.B * a.nwin * tb.num_groups;
    cdl_moment_kernel<P><<<(unsigned)blocks, 128, 0, st>>>(a, tb);
  }
  HB_CUDA(cudaGetLastError());
  return HB_OK;
}

template <int NRX, int P, int R, typename IO>
static int launch_poly_one(const CdlArgs& a, const CdlTable& tb, size_t smem, cudaStream_t st) {
  auto kern = cdl_poly_kernel<NRX, P, R, 128, IO>;
  if (smem > 48 * 1024) HB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfileScope prof(KIND_CDL_PROPAGATE, st);
  kern<<<(unsigned)((size_t)a.B * a.ntiles), 128, smem, st>>>(a, tb);
  HB_CUDA(cudaGetLastError());
  return HB_OK;
}

template <int NRX, int R, typename IO>
static int launch_poly_p(int P, const CdlArgs& a, const CdlTable& tb, size_t smem, cudaStream_t st) {
  switch (P) {
    case 1: return launch_poly_one<NRX, 1, R, IO>(a, tb, smem, st);
    case 2: return launch_poly_one<NRX, 2, R, IO>(a, tb, smem, st);
    case 3: return launch_poly_one<NRX, 3, R, IO>(a, tb, smem, st);
    default: return launch_poly_one<NRX, 4, R, IO>(a, tb, smem, st);
  }
}

// R = outputs per thread = tile / 128: 4 (2 for 8 receive antennas) by default, halved for fast links whose Taylor bound needs
// shorter windows (make_cdl_plan)
template <int NRX, typename IO>
static int launch_poly_r(int R, int P, const CdlArgs& a, const CdlTable& tb, size_t smem, cudaStream_t st) {
  if constexpr (NRX <= 4) {
    if (R == 4) return launch_poly_p<NRX, 4, IO>(P, a, tb, smem, st);
  }
  if (R == 2) return launch_poly_p<NRX, 2, IO>(P, a, tb, smem, st);
  return launch_poly_p<NRX, 1, IO>(P, a, tb, smem, st);
}

template <typename IO>
static int launch_poly_nrx(int nrx_tpl, int R, int P, const CdlArgs& a, const CdlTable& tb, size_t smem, cudaStream_t st) {
  switch (nrx_tpl) {
    case 1: return launch_poly_r<1, IO>(R, P, a, tb, smem, st);
    case 2: return launch_poly_r<2, IO>(R, P, a, tb, smem, st);
    case 4: return launch_poly_r<4, IO>(R, P, a, tb, smem, st);
    default: return launch_poly_r<8, IO>(R, P, a, tb, smem, st);
  }
}

struct CdlWorkspace {
  double2* alpha = nullptr;
  double* w = nullptr;
  double2* u = nullptr;
  double2* v = nullptr;
  float2* moments = nullptr;
};

static void free_ws(CdlWorkspace& ws, cudaStream_t st) {
  if (ws.alpha) cudaFreeAsync(ws.alpha, st);
  if (ws.w) cudaFreeAsync(ws.w, st);
  if (ws.u) cudaFreeAsync(ws.u, st);
  if (ws.v) cudaFreeAsync(ws.v, st);
  if (ws.moments) cudaFreeAsync(ws.moments, st);
  ws = CdlWorkspace();
}

static void fill_args(const hb_cdl_problem* p, const CdlTable& tb, const CdlPlan& pl, CdlArgs* a) {
  memset(a, 0, sizeof(*a));
  a->angles = p->angles;
  a->jones = reinterpret_cast<const double2*>(p->jones);
  a->amp = p->amplitude;
  a->tx_pose = p->tx_pose;
  a->rx_pose = p->rx_pose;
  a->rel_velocity = p->rel_velocity;
  a->tx_topology = p->tx_topology;
  a->rx_topology = p->rx_topology;
  a->tx_elements = p->tx_elements;
  a->rx_elements = p->rx_elements;
  a->element_mode = p->element_mode;
  a->rank = p->element_mode == HB_ELEMENTS_PER_ELEMENT ? 2 : 1;
  a->wavelength_factor = p->carrier_frequency / kSpeedOfLight;
  a->fs = p->sampling_rate;
  a->los_amp = p->los_amplitude;
  a->B = p->batch;
  a->ntx = p->num_tx;
  a->nrx = p->num_rx;
  a->T = p->num_samples;
  a->D = p->max_delay;
  a->Rn = p->num_terms;
  a->Rt = tb.num_terms;
  a->tile = pl.tile;
  a->ntiles = pl.ntiles;
  a->Dpad = pl.Dpad;
  a->P = pl.P;
  a->ptile = pl.ptile;
  a->nwin = pl.nwin;
}

static int alloc_rays(CdlArgs* a, CdlWorkspace* ws, cudaStream_t st) {
  const size_t bt = (size_t)a->B * a->Rt;
  HB_CUDA(cudaMallocAsync((void**)&ws->alpha, sizeof(double2) * bt, st));
  HB_CUDA(cudaMallocAsync((void**)&ws->w, sizeof(double) * bt, st));
  HB_CUDA(cudaMallocAsync((void**)&ws->u, sizeof(double2) * bt * a->nrx * a->rank, st));
  HB_CUDA(cudaMallocAsync((void**)&ws->v, sizeof(double2) * bt * a->ntx * a->rank, st));
  a->alpha = ws->alpha;
  a->w = ws->w;
  a->u = ws->u;
  a->v = ws->v;
  return HB_OK;
}

static int launch_rays(const CdlArgs& a, const CdlTable& tb, cudaStream_t st) {
  ProfileScope prof(KIND_CDL_RAYS, st);
  const size_t n = (size_t)a.B * a.Rt;
  cdl_ray_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(a, tb);
  HB_CUDA(cudaGetLastError());
  return HB_OK;
}

// d_link_tab / d_link_los: device copies of the per-link tables of THIS batch slice (nullptr: uniform batch).
static int cdl_propagate_device(const hb_cdl_problem* p, const CdlTable& tb, const CdlPlan& pl, const void* x, void* y,
                                cudaStream_t st, const CdlTable* d_link_tab = nullptr, const double* d_link_los = nullptr) {
  const int Tout = p->num_samples + p->max_delay;
  if (p->batch == 0 || Tout == 0) return HB_OK;
  CdlArgs a;
  fill_args(p, tb, pl, &a);
  a.x = x;
  a.y = y;
  a.link_tab = d_link_tab;
  a.link_los_amp = d_link_los;
  if ((size_t)a.B * a.ntiles * std::max(1, tb.num_groups) > 0x7fffffffull) {
    set_error("CDL grid exceeds the launch limit; split the batch");
    return HB_ERR_UNSUPPORTED;
  }
  CdlWorkspace ws;
  int rc = alloc_rays(&a, &ws, st);
  if (rc == HB_OK) rc = launch_rays(a, tb, st);
  const bool io128 = p->io_complex128 != 0;
  if (rc == HB_OK && pl.mode == HB_SOS_POLY) {
    const size_t mbytes = sizeof(float2) * (size_t)a.B * a.nwin * tb.num_groups * pl.P * a.nrx * a.ntx;
    cudaError_t e = cudaMallocAsync((void**)&ws.moments, mbytes, st);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaMallocAsync(moments)");
    a.moments = ws.moments;
    if (rc == HB_OK) {
      switch (pl.P) {
        case 1: rc = launch_moments<1>(a, tb, st); break;
        case 2: rc = launch_moments<2>(a, tb, st); break;
        case 3: rc = launch_moments<3>(a, tb, st); break;
        default: rc = launch_moments<4>(a, tb, st); break;
      }
    }
    for (int rx0 = 0; rx0 < p->num_rx && rc == HB_OK; rx0 += pl.nrx_tpl) {
      a.rx0 = rx0;
      a.nrx_chunk = std::min(pl.nrx_tpl, p->num_rx - rx0);
      if (pl.variant == HB_CDL_VARIANT_UMMA_BF16)
        rc = launch_cdl_umma_bf16(pl.nrx_tpl, pl.P, io128, a, tb, pl.smem, st);
      else if (pl.variant == HB_CDL_VARIANT_UMMA)
        rc = io128 ? launch_cdl_umma_io<double2>(pl.nrx_tpl, pl.P, a, tb, pl.smem, st)
                   : launch_cdl_umma_io<float2>(pl.nrx_tpl, pl.P, a, tb, pl.smem, st);
      else
        rc = io128 ? launch_poly_nrx<double2>(pl.nrx_tpl, pl.R, pl.P, a, tb, pl.smem, st)
                   : launch_poly_nrx<float2>(pl.nrx_tpl, pl.R, pl.P, a, tb, pl.smem, st);
    }
  } else if (rc == HB_OK) {
    ProfileScope prof(KIND_CDL_PROPAGATE, st);
    const size_t blocks = (size_t)a.B * ((Tout + 127) / 128);
    if (io128) cdl_direct_f64_kernel<double2><<<(unsigned)blocks, 128, 0, st>>>(a, tb);
    else cdl_direct_f64_kernel<float2><<<(unsigned)blocks, 128, 0, st>>>(a, tb);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = cuda_fail(e, "cdl_direct_f64_kernel");
  }
  free_ws(ws, st);
  return rc;
}

}  // namespace hb

using namespace hb;

extern "C" {

int hb_cdl_plan(const hb_cdl_problem* p, hb_cdl_plan_info* info) {
  CdlTables tabs;
  if (int e = build_cdl_table(p, &tabs)) return e;
  CdlPlan pl;
  if (int e = make_cdl_plan(p, tabs, &pl)) return e;
  fill_cdl_info(pl, tabs.tb, p, info);
  return HB_OK;
}

int hb_cdl_propagate(const hb_cdl_problem* p, const void* x, void* y, void* stream, hb_cdl_plan_info* info) {
  CdlTables tabs;
  if (int e = build_cdl_table(p, &tabs)) return e;
  const CdlTable& tb = tabs.tb;
  CdlPlan pl;
  if (int e = make_cdl_plan(p, tabs, &pl)) return e;
  fill_cdl_info(pl, tb, p, info);
  if (int e = require_device()) return e;
  if (p->batch > 0 && (!x || !y || !p->tx_pose || !p->rx_pose || !p->rel_velocity || !p->tx_topology ||
                       !p->rx_topology || (p->num_terms > 0 && (!p->angles || !p->jones || !p->amplitude)) ||
                       (p->element_mode != HB_ELEMENTS_IDEAL && (!p->tx_elements || !p->rx_elements)))) {
    set_error("NULL device pointer in CDL problem");
    return HB_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (tabs.links.empty()) return cdl_propagate_device(p, tb, pl, x, y, st);
  // heterogeneous batch: the per-link tables go to the device first (pageable source: the copy is staged before the call
  // returns, so the vectors may die with this frame)
  CdlTable* d_tab = nullptr;
  double* d_los = nullptr;
  HB_CUDA(cudaMallocAsync((void**)&d_tab, sizeof(CdlTable) * tabs.links.size(), st));
  cudaError_t e = cudaMallocAsync((void**)&d_los, sizeof(double) * tabs.los_amp.size(), st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_tab, tabs.links.data(), sizeof(CdlTable) * tabs.links.size(), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_los, tabs.los_amp.data(), sizeof(double) * tabs.los_amp.size(), cudaMemcpyHostToDevice, st);
  int rc = e == cudaSuccess ? cdl_propagate_device(p, tb, pl, x, y, st, d_tab, d_los) : cuda_fail(e, "per-link CDL tables");
  cudaFreeAsync(d_tab, st);
  if (d_los) cudaFreeAsync(d_los, st);
  return rc;
}

int hb_cdl_propagate_host(const hb_cdl_problem* p, const void* x, void* y, int32_t chunk_links,
                          hb_cdl_plan_info* info) {
  CdlTables tabs;
  if (int e = build_cdl_table(p, &tabs)) return e;
  const CdlTable& tb = tabs.tb;
  CdlPlan pl;
  if (int e = make_cdl_plan(p, tabs, &pl)) return e;
  fill_cdl_info(pl, tb, p, info);
  if (int e = require_device()) return e;
  const bool per_link = !tabs.links.empty();
  const int Tout = p->num_samples + p->max_delay;
  if (p->batch == 0 || Tout == 0) return HB_OK;
  if (!x || !y || !p->tx_pose || !p->rx_pose || !p->rel_velocity || !p->tx_topology || !p->rx_topology ||
      (p->num_terms > 0 && (!p->angles || !p->jones || !p->amplitude)) ||
      (p->element_mode != HB_ELEMENTS_IDEAL && (!p->tx_elements || !p->rx_elements))) {
    set_error("NULL host pointer in CDL problem");
    return HB_ERR_INVALID;
  }
  const size_t esz = p->io_complex128 ? 16 : 8;
  const size_t x_link = esz * (size_t)p->num_tx * p->num_samples;
  const size_t y_link = esz * (size_t)p->num_rx * Tout;
  const size_t ang_link = sizeof(double) * 4 * (size_t)p->num_terms;
  const size_t jon_link = 16 * 4 * (size_t)p->num_terms;
  const size_t amp_link = sizeof(double) * (size_t)p->num_terms;
  const size_t topo_tx = sizeof(double) * 3 * (size_t)p->num_tx, topo_rx = sizeof(double) * 3 * (size_t)p->num_rx;
  const size_t el_row = sizeof(double) * HB_ELEMENT_STRIDE;
  const size_t el_tx = p->element_mode == HB_ELEMENTS_IDEAL ? 0 : el_row * (p->element_mode == HB_ELEMENTS_UNIFORM ? 1 : (size_t)p->num_tx);
  const size_t el_rx = p->element_mode == HB_ELEMENTS_IDEAL ? 0 : el_row * (p->element_mode == HB_ELEMENTS_UNIFORM ? 1 : (size_t)p->num_rx);
  int chunk = chunk_links;
  if (chunk <= 0) {
    chunk = (int)std::max<size_t>(1, (48u << 20) / std::max<size_t>(1, x_link + y_link));
    if ((x_link + y_link) * (size_t)p->batch > (8u << 20))  // small batches: one chunk, one set of launches
      chunk = std::min(chunk, std::max(1, (p->batch + kSlots - 1) / kSlots));
  }
  chunk = std::min(chunk, p->batch);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  const size_t off_x = take(x_link * chunk), off_y = take(y_link * chunk), off_ang = take(ang_link * chunk),
               off_jon = take(jon_link * chunk), off_amp = take(amp_link * chunk), off_tp = take(96 * (size_t)chunk),
               off_rp = take(96 * (size_t)chunk), off_rv = take(24 * (size_t)chunk), off_tt = take(topo_tx),
               off_rt = take(topo_rx), off_et = take(el_tx), off_er = take(el_rx),
               off_lt = take(per_link ? sizeof(CdlTable) * (size_t)chunk : 0), off_la = take(per_link ? 8 * (size_t)chunk : 0);
  const size_t total = off;

  std::lock_guard<std::mutex> lock(g_pipe.mu);
  if (int e = pipe_prepare(total)) return e;
  int rc = HB_OK;
  int ci = 0;
  for (int b0 = 0; b0 < p->batch && rc == HB_OK; b0 += chunk, ++ci) {
    const int nb = std::min(chunk, p->batch - b0);
    const int s = ci % kSlots;
    cudaStream_t st = g_pipe.st[s];
    char* base = (char*)g_pipe.buf[s];
    hb_cdl_problem q = *p;
    q.batch = nb;
    q.angles = (const double*)(base + off_ang);
    q.jones = base + off_jon;
    q.amplitude = (const double*)(base + off_amp);
    q.tx_pose = (const double*)(base + off_tp);
    q.rx_pose = (const double*)(base + off_rp);
    q.rel_velocity = (const double*)(base + off_rv);
    q.tx_topology = (const double*)(base + off_tt);
    q.rx_topology = (const double*)(base + off_rt);
    q.tx_elements = (const double*)(base + off_et);
    q.rx_elements = (const double*)(base + off_er);
    struct Cp {
      size_t dst;
      const void* src;
      size_t bytes;
    } cps[] = {
        {off_ang, (const char*)p->angles + ang_link * b0, ang_link * nb},
        {off_jon, (const char*)p->jones + jon_link * b0, jon_link * nb},
        {off_amp, (const char*)p->amplitude + amp_link * b0, amp_link * nb},
        {off_tp, (const char*)p->tx_pose + 96 * (size_t)b0, 96 * (size_t)nb},
        {off_rp, (const char*)p->rx_pose + 96 * (size_t)b0, 96 * (size_t)nb},
        {off_rv, (const char*)p->rel_velocity + 24 * (size_t)b0, 24 * (size_t)nb},
        {off_tt, p->tx_topology, topo_tx},
        {off_rt, p->rx_topology, topo_rx},
        {off_et, p->tx_elements, el_tx},
        {off_er, p->rx_elements, el_rx},
        {off_lt, per_link ? (const void*)(tabs.links.data() + b0) : nullptr, per_link ? sizeof(CdlTable) * (size_t)nb : 0},
        {off_la, per_link ? (const void*)(tabs.los_amp.data() + b0) : nullptr, per_link ? 8 * (size_t)nb : 0},
        {off_x, (const char*)x + x_link * b0, x_link * nb},
    };
    for (const Cp& c : cps) {
      if (c.bytes == 0) continue;
      cudaError_t e = cudaMemcpyAsync(base + c.dst, c.src, c.bytes, cudaMemcpyHostToDevice, st);
      if (e != cudaSuccess) {
        rc = cuda_fail(e, "cudaMemcpyAsync(H2D, cdl)");
        break;
      }
    }
    if (rc != HB_OK) break;
    rc = cdl_propagate_device(&q, tb, pl, base + off_x, base + off_y, st,
                              per_link ? (const CdlTable*)(base + off_lt) : nullptr, per_link ? (const double*)(base + off_la) : nullptr);
    if (rc != HB_OK) break;
    cudaError_t e = cudaMemcpyAsync((char*)y + y_link * b0, base + off_y, y_link * nb, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpyAsync(D2H, cdl)");
  }
  for (int s = 0; s < kSlots; ++s) {
    cudaError_t e = cudaStreamSynchronize(g_pipe.st[s]);
    if (e != cudaSuccess && rc == HB_OK) rc = cuda_fail(e, "cudaStreamSynchronize(pipeline)");
  }
  return rc;
}

int hb_cdl_state(const hb_cdl_problem* p, void* h, int32_t* group_delay_out, int32_t* num_groups_out, void* stream) {
  CdlTables tabs;
  if (int e = build_cdl_table(p, &tabs)) return e;
  if (!tabs.links.empty()) {
    set_error("hb_cdl_state takes batches of one delay structure (link_term_delay must be NULL): its output is indexed by delay group");
    return HB_ERR_UNSUPPORTED;
  }
  const CdlTable& tb = tabs.tb;
  if (num_groups_out) *num_groups_out = tb.num_groups;
  if (group_delay_out)
    for (int g = 0; g < tb.num_groups; ++g) group_delay_out[g] = tb.group_delay[g];
  if (p->batch == 0 || p->num_samples == 0 || !h) return HB_OK;
  if (int e = require_device()) return e;
  CdlPlan pl;
  memset(&pl, 0, sizeof(pl));
  pl.tile = 128;
  pl.ntiles = (p->num_samples + 127) / 128;
  CdlArgs a;
  fill_args(p, tb, pl, &a);
  a.y = h;
  cudaStream_t st = (cudaStream_t)stream;
  CdlWorkspace ws;
  int rc = alloc_rays(&a, &ws, st);
  if (rc == HB_OK) rc = launch_rays(a, tb, st);
  if (rc == HB_OK) {
    const size_t blocks = (size_t)a.B * tb.num_groups * a.nrx * a.ntx * ((a.T + 127) / 128);
    if (blocks > 0x7fffffffull) {
      set_error("CDL state grid exceeds the launch limit; split the batch");
      rc = HB_ERR_UNSUPPORTED;
    } else {
      ProfileScope prof(KIND_SOS_STATE, st);
      if (p->io_complex128) cdl_state_kernel<double2><<<(unsigned)blocks, 128, 0, st>>>(a, tb);
      else cdl_state_kernel<float2><<<(unsigned)blocks, 128, 0, st>>>(a, tb);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) rc = cuda_fail(e, "cdl_state_kernel");
    }
  }
  free_ws(ws, st);
  return rc;
}

}  // extern "C"
