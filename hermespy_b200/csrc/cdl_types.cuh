// Launch-uniform tables and the argument block shared by the cluster-delay-line kernels (cdl_kernels.cuh, cdl_umma.cuh).
#pragma once
#include "hb_common.cuh"

namespace hb {

constexpr int kCdlMaxTerms = HB_CDL_MAX_TERMS;
constexpr int kCdlMaxGroups = HB_CDL_MAX_GROUPS;
constexpr int kCdlTxChunk = 8;  // transmit antennas staged per pass of K6
constexpr double kSpeedOfLight = 299792458.0;

// Launch-uniform term tables, passed by value in kernel parameter space.
struct CdlTable {
  int32_t num_terms;   // ray terms incl. the optional line-of-sight term (always the last one)
  int32_t num_groups;  // distinct delay indices
  int32_t has_los;     // 1: term num_terms-1 is the line-of-sight term
  int32_t group_delay[kCdlMaxGroups];
  uint16_t group_start[kCdlMaxGroups + 1];  // into term_order
  uint16_t term_order[kCdlMaxTerms];        // terms sorted by delay group
  uint16_t term_delay[kCdlMaxTerms];        // delay index per term (original order)
};

struct CdlArgs {
  const void* x;
  void* y;
  // inputs of K5
  const double* angles;       // [B, Rn, 4] aoa, zoa, aod, zod
  const double2* jones;       // [B, Rn, 4]
  const double* amp;          // [B, Rn]
  const double* tx_pose;      // [B, 12]
  const double* rx_pose;      // [B, 12]
  const double* rel_velocity; // [B, 3]
  const double* tx_topology;  // [Ntx, 3]
  const double* rx_topology;  // [Nrx, 3]
  const double* tx_elements;  // [Ntx or 1, HB_ELEMENT_STRIDE] element models (element_mode != IDEAL)
  const double* rx_elements;  // [Nrx or 1, HB_ELEMENT_STRIDE]
  int element_mode;           // hb_element_mode
  int rank;                   // 1: H_t = alpha u v^T;  2: H_t = sum_c u[:, c] v[:, c]^T (per-element patterns, alpha = 1)
  // ray coefficients (K5 out)
  double2* alpha;  // [B, Rt]
  double* w;       // [B, Rt] rad / sample
  double2* u;      // [B, Rt, Nrx, rank]   receive steering phases (x element polarization when rank == 2)
  double2* v;      // [B, Rt, Ntx, rank]   transmit steering phases (x amp J F_tx when rank == 2)
  float2* moments; // [B, nwin, G, P, Nrx, Ntx]
  double wavelength_factor;  // fc / c0
  double fs;
  double los_amp;
  int B, ntx, nrx, T, D, Rn, Rt;
  int tile, ntiles, Dpad, P;
  int ptile, nwin;  // Taylor window length (a multiple of tile) and windows per frame: moments are [B, nwin, G, P, Nrx, Ntx]
  int rx0, nrx_chunk;
  // heterogeneous batches (hb_cdl_problem.link_term_delay): one table per link in device memory, padded to the launch
  // table's num_terms / num_groups (empty groups: group_start[g] == group_start[g + 1]); NULL = the launch-uniform table
  const CdlTable* link_tab;
  const double* link_los_amp;  // [B], NULL = los_amp
};

// Table lookups: the per-link table of link b when the batch carries one, else the launch-uniform table in parameter space.
__device__ __forceinline__ int cdl_group_delay(const CdlArgs& a, const CdlTable& tb, int b, int g) {
  return a.link_tab ? a.link_tab[b].group_delay[g] : tb.group_delay[g];
}
__device__ __forceinline__ int cdl_group_start(const CdlArgs& a, const CdlTable& tb, int b, int g) {
  return a.link_tab ? (int)a.link_tab[b].group_start[g] : (int)tb.group_start[g];
}
__device__ __forceinline__ int cdl_term_order(const CdlArgs& a, const CdlTable& tb, int b, int c) {
  return a.link_tab ? (int)a.link_tab[b].term_order[c] : (int)tb.term_order[c];
}
__device__ __forceinline__ int cdl_term_delay(const CdlArgs& a, const CdlTable& tb, int b, int t) {
  return a.link_tab ? (int)a.link_tab[b].term_delay[t] : (int)tb.term_delay[t];
}

}  // namespace hb
