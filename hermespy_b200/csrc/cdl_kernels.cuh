// 3GPP cluster-delay-line hot path kernels (sm_100a).
//
//   K5  cdl_ray_kernel        per (link, ray term), FP64: polarization scalar alpha, Doppler rate w, and the
//                             unit-modulus steering phases u[Nrx], v[Ntx] of the array responses
//                             (cluster_delay_lines.py:409-523, core/antennas.py:138-210, 883-1000)
//   K5b cdl_moment_kernel     per (link, tile, delay group): Taylor moments of the time-varying MIMO matrix
//                             H_g(m) = sum_{t in g} alpha_t e^{j w_t (m - k_g)} u_t v_t^T   (FP32 output)
//   K6  cdl_poly_kernel       y[:, m] = sum_p r^p sum_g M[g,p] x[:, m - k_g]; x tile + delay halo and the moment
//                             matrices staged in shared memory per chunk of 8 transmit antennas
//                             (cluster_delay_lines.py:526-558)
//   K6' cdl_direct_f64_kernel FP64 parity mode: every ray term evaluated per sample exactly like the reference
//   cdl_state_kernel          H_g(n) for the channel state (cluster_delay_lines.py:561-592)
//
// Structure exploited: for arrays of identical, identically oriented elements the per-ray matrix of the reference,
// a_rx J a_tx^T, is rank one: H_t = alpha_t u_t v_t^T with u, v pure phases.  Terms sharing a delay index are summed
// into one matrix per delay group before touching the signal, which removes the ray count from the per-sample cost.
#pragma once
#include <type_traits>

#include "cdl_types.cuh"

namespace hb {

struct Vec3 {
  double x, y, z;
};
__device__ __forceinline__ Vec3 mat_t_vec(const double* R, Vec3 a) {  // R^T a, R row-major
  return {R[0] * a.x + R[3] * a.y + R[6] * a.z, R[1] * a.x + R[4] * a.y + R[7] * a.z,
          R[2] * a.x + R[5] * a.y + R[8] * a.z};
}
__device__ __forceinline__ Vec3 mat_vec(const double* R, Vec3 a) {
  return {R[0] * a.x + R[1] * a.y + R[2] * a.z, R[3] * a.x + R[4] * a.y + R[5] * a.z,
          R[6] * a.x + R[7] * a.y + R[8] * a.z};
}
__device__ __forceinline__ double dot3(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// theta / phi unit vectors of TR 38.901 eq. 7.1-13/14 for a unit direction d (azimuth = atan2(y, x),
// zenith = acos(z), as hermespy/core/transformation.py:66-84), formed without inverse trigonometry.
__device__ __forceinline__ void sph_basis(Vec3 d, Vec3* th, Vec3* ph) {
  const double rho = sqrt(d.x * d.x + d.y * d.y);
  const double ca = rho > 0.0 ? d.x / rho : 1.0, sa = rho > 0.0 ? d.y / rho : 0.0;
  const double cz = fmin(1.0, fmax(-1.0, d.z)), sz = sqrt(fmax(0.0, 1.0 - cz * cz));
  *th = {cz * ca, cz * sa, -sz};
  *ph = {-sa, ca, 0.0};
}

// Local field pattern (F_theta, F_phi) of the reference's element models towards the LOCAL unit direction l
// (core/antennas.py:435-436 ideal, :509-510 linear, :556-560 patch, :610-614 dipole; the reference hands the zenith angle
// to the parameter it calls "elevation").
__device__ __forceinline__ void local_pattern(int kind, double param, Vec3 l, double* f0, double* f1) {
  switch (kind) {
    case HB_ELEMENT_LINEAR: {
      double s, c;
      sincos(param, &s, &c);
      *f0 = c;
      *f1 = s;
      return;
    }
    case HB_ELEMENT_PATCH: {
      const double az = atan2(l.y, l.x);
      const double cz = fmin(1.0, fmax(-1.0, l.z));
      const double va = 0.1 + 0.9 * exp(-1.315 * az * az);
      *f0 = fmax(0.1, va * cz * cz);
      *f1 = 0.0;
      return;
    }
    case HB_ELEMENT_DIPOLE: {
      const double cz = fmin(1.0, fmax(-1.0, l.z));
      const double ze = acos(cz);  // as the reference: the pattern is evaluated on arccos(z)
      *f0 = ze == 0.0 ? 0.0 : cos(1.5707963267948966 * cos(ze)) / sin(ze);
      *f1 = 0.0;
      return;
    }
    default:
      *f0 = 0.70710678118654752440;
      *f1 = 0.70710678118654752440;
  }
}

// Polarization of an element whose orientation is R (element frame -> global) towards global unit direction g:
// the local pattern rotated into the global theta / phi basis by the 2x2 matrix of TR 38.901 eq. 7.1-12
// (core/antennas.py:138-210).
__device__ __forceinline__ void element_polarization(const double* R, Vec3 g, int kind, double param, double* f_theta,
                                                     double* f_phi) {
  Vec3 thg, phg, thl, phl;
  sph_basis(g, &thg, &phg);
  const Vec3 l = mat_t_vec(R, g);
  sph_basis(l, &thl, &phl);
  const Vec3 thlt = mat_vec(R, thl), phlt = mat_vec(R, phl);
  double f0, f1;
  local_pattern(kind, param, l, &f0, &f1);
  *f_theta = dot3(thg, thlt) * f0 + dot3(thg, phlt) * f1;
  *f_phi = dot3(phg, thlt) * f0 + dot3(phg, phlt) * f1;
}

// C = A B for row-major 3x3 matrices
__device__ __forceinline__ void mat_mul3(const double* A, const double* B, double* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// exp(-2j pi * turns) with the integer part of `turns` removed in FP64 first.
__device__ __forceinline__ double2 phasor_neg_turns(double turns) {
  const double f = turns - rint(turns);
  double s, c;
  sincospi(-2.0 * f, &s, &c);
  return make_double2(c, s);
}

// ------------------------------------------------------------------------------------------------
// K5: one thread per (link, term).
__global__ void __launch_bounds__(128) cdl_ray_kernel(const CdlArgs a, const __grid_constant__ CdlTable tb) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)a.B * a.Rt) return;
  const int b = (int)(gid / a.Rt), t = (int)(gid - (long long)b * a.Rt);
  const double* tp = a.tx_pose + (size_t)b * 12;
  const double* rp = a.rx_pose + (size_t)b * 12;
  const Vec3 ttx = {tp[9], tp[10], tp[11]}, trx = {rp[9], rp[10], rp[11]};
  const Vec3 rv = {a.rel_velocity[b * 3 + 0], a.rel_velocity[b * 3 + 1], a.rel_velocity[b * 3 + 2]};
  const bool los = tb.has_los && t == a.Rt - 1;

  Vec3 tgt_rx, tgt_tx, wave;  // positions handed to the array responses, and the Doppler wave vector
  double2 j00, j01, j10, j11;
  double2 scale;
  if (!los) {
    const double* ang = a.angles + ((size_t)b * a.Rn + t) * 4;
    double sa, ca, sz, cz;
    sincos(ang[0], &sa, &ca);
    sincos(ang[1], &sz, &cz);
    tgt_rx = {sz * ca, sz * sa, cz};  // unit vector used as a global *position* 
    wave = tgt_rx;
    sincos(ang[2], &sa, &ca);
    sincos(ang[3], &sz, &cz);
    tgt_tx = {sz * ca, sz * sa, cz};
    const double2* J = a.jones + ((size_t)b * a.Rn + t) * 4;
    j00 = J[0];
    j01 = J[1];
    j10 = J[2];
    j11 = J[3];
    scale = make_double2(a.amp[(size_t)b * a.Rn + t], 0.0);
  } else {
    tgt_rx = ttx;  // receiver array looks at the transmitter position and vice versa
    tgt_tx = trx;
    const Vec3 dv = {trx.x - ttx.x, trx.y - ttx.y, trx.z - ttx.z};
    const double dist = sqrt(dot3(dv, dv));
    wave = {dv.x / dist, dv.y / dist, dv.z / dist};
    j00 = make_double2(1.0, 0.0);  // [[1, 0], [-1, 0]] (sic, cluster_delay_lines.py:515)
    j01 = make_double2(0.0, 0.0);
    j10 = make_double2(-1.0, 0.0);
    j11 = make_double2(0.0, 0.0);
    const double2 ph = phasor_neg_turns(dist * a.wavelength_factor);
    const double los_amp = a.link_los_amp ? a.link_los_amp[b] : a.los_amp;
    scale = make_double2(los_amp * ph.x, los_amp * ph.y);
  }

  // polarization towards normalize(target - array position)
  Vec3 grx = {tgt_rx.x - trx.x, tgt_rx.y - trx.y, tgt_rx.z - trx.z};
  Vec3 gtx = {tgt_tx.x - ttx.x, tgt_tx.y - ttx.y, tgt_tx.z - ttx.z};
  const Vec3 lrx = mat_t_vec(rp, grx), ltx = mat_t_vec(tp, gtx);  // targets in array coordinates
  double n = sqrt(dot3(grx, grx));
  grx = {grx.x / n, grx.y / n, grx.z / n};
  n = sqrt(dot3(gtx, gtx));
  gtx = {gtx.x / n, gtx.y / n, gtx.z / n};
  a.w[(size_t)b * a.Rt + t] = kTwoPi * dot3(wave, rv) * a.wavelength_factor / a.fs;
  double2* u = a.u + ((size_t)b * a.Rt + t) * a.nrx * a.rank;
  double2* v = a.v + ((size_t)b * a.Rt + t) * a.ntx * a.rank;

  if (a.rank == 1) {
    // identical, identically oriented elements: one polarization pair per array, H_t = alpha u v^T
    double frt, frp, ftt, ftp;
    if (a.element_mode == HB_ELEMENTS_IDEAL) {
      element_polarization(rp, grx, HB_ELEMENT_IDEAL, 0.0, &frt, &frp);
      element_polarization(tp, gtx, HB_ELEMENT_IDEAL, 0.0, &ftt, &ftp);
    } else {
      double Rr[9], Rt_[9];
      mat_mul3(rp, a.rx_elements, Rr);
      mat_mul3(tp, a.tx_elements, Rt_);
      element_polarization(Rr, grx, (int)a.rx_elements[9], a.rx_elements[10], &frt, &frp);
      element_polarization(Rt_, gtx, (int)a.tx_elements[9], a.tx_elements[10], &ftt, &ftp);
    }
    // F_rx^T J F_tx
    double2 jt0 = make_double2(j00.x * ftt + j01.x * ftp, j00.y * ftt + j01.y * ftp);
    double2 jt1 = make_double2(j10.x * ftt + j11.x * ftp, j10.y * ftt + j11.y * ftp);
    double2 pol = make_double2(frt * jt0.x + frp * jt1.x, frt * jt0.y + frp * jt1.y);
    a.alpha[(size_t)b * a.Rt + t] = cmul(pol, scale);
    for (int i = 0; i < a.nrx; ++i) {
      const double dx = a.rx_topology[i * 3] - lrx.x, dy = a.rx_topology[i * 3 + 1] - lrx.y,
                   dz = a.rx_topology[i * 3 + 2] - lrx.z;
      u[i] = phasor_neg_turns(sqrt(dx * dx + dy * dy + dz * dz) * a.wavelength_factor);
    }
    for (int j = 0; j < a.ntx; ++j) {
      const double dx = a.tx_topology[j * 3] - ltx.x, dy = a.tx_topology[j * 3 + 1] - ltx.y,
                   dz = a.tx_topology[j * 3 + 2] - ltx.z;
      v[j] = phasor_neg_turns(sqrt(dx * dx + dy * dy + dz * dz) * a.wavelength_factor);
    }
    return;
  }
  // per-element patterns / orientations: a_rx[i, :] = u_i F_rx,i and (scale J F_tx,j) v_j are stored per polarization
  // component; H_t[i, j] = sum_c u[i, c] v[j, c]  (a_rx J a_tx^T, cluster_delay_lines.py:481-486)
  a.alpha[(size_t)b * a.Rt + t] = make_double2(1.0, 0.0);
  for (int i = 0; i < a.nrx; ++i) {
    const double* el = a.rx_elements + (size_t)i * HB_ELEMENT_STRIDE;
    double Rm[9], ft, fp;
    mat_mul3(rp, el, Rm);
    element_polarization(Rm, grx, (int)el[9], el[10], &ft, &fp);
    const double dx = a.rx_topology[i * 3] - lrx.x, dy = a.rx_topology[i * 3 + 1] - lrx.y,
                 dz = a.rx_topology[i * 3 + 2] - lrx.z;
    const double2 ph = phasor_neg_turns(sqrt(dx * dx + dy * dy + dz * dz) * a.wavelength_factor);
    u[2 * i] = make_double2(ph.x * ft, ph.y * ft);
    u[2 * i + 1] = make_double2(ph.x * fp, ph.y * fp);
  }
  for (int j = 0; j < a.ntx; ++j) {
    const double* el = a.tx_elements + (size_t)j * HB_ELEMENT_STRIDE;
    double Rm[9], ft, fp;
    mat_mul3(tp, el, Rm);
    element_polarization(Rm, gtx, (int)el[9], el[10], &ft, &fp);
    const double dx = a.tx_topology[j * 3] - ltx.x, dy = a.tx_topology[j * 3 + 1] - ltx.y,
                 dz = a.tx_topology[j * 3 + 2] - ltx.z;
    const double2 ph = phasor_neg_turns(sqrt(dx * dx + dy * dy + dz * dz) * a.wavelength_factor);
    const double2 jf0 = cmul(make_double2(j00.x * ft + j01.x * fp, j00.y * ft + j01.y * fp), scale);
    const double2 jf1 = cmul(make_double2(j10.x * ft + j11.x * fp, j10.y * ft + j11.y * fp), scale);
    v[2 * j] = cmul(jf0, ph);
    v[2 * j + 1] = cmul(jf1, ph);
  }
}

// Entry (i, j) of the unit-amplitude ray matrix of term t: u_i v_j (rank one) or sum_c u[i, c] v[j, c] (rank two).
__device__ __forceinline__ double2 ray_entry(const CdlArgs& a, int b, int t, int i, int j) {
  const double2* u = a.u + (((size_t)b * a.Rt + t) * a.nrx + i) * a.rank;
  const double2* v = a.v + (((size_t)b * a.Rt + t) * a.ntx + j) * a.rank;
  double2 r = cmul(u[0], v[0]);
  if (a.rank == 2) {
    const double2 r1 = cmul(u[1], v[1]);
    r.x += r1.x;
    r.y += r1.y;
  }
  return r;
}

// ------------------------------------------------------------------------------------------------
// K5b: moments of the group matrices about the centre of the Taylor window (a.ptile samples).  grid = B * nwin * G,
// 128 threads over (i, j).
//   M[g][p][i][j] = sum_{t in g} alpha_t e^{j w_t (centre - k_g)} (j w_t tile)^p / p! * u_t[i] v_t[j]
template <int P>
__global__ void __launch_bounds__(128) cdl_moment_kernel(const CdlArgs a, const __grid_constant__ CdlTable tb) {
  __shared__ float2 beta[64];
  __shared__ float ut[64];
  const int G = tb.num_groups;
  const int bq = blockIdx.x / G, g = blockIdx.x - bq * G;
  const int b = bq / a.nwin, q = bq - b * a.nwin;
  const int t0 = cdl_group_start(a, tb, b, g), t1 = cdl_group_start(a, tb, b, g + 1);
  const double shift = (double)q * a.ptile + 0.5 * a.ptile - (double)cdl_group_delay(a, tb, b, g);
  const int nij = a.nrx * a.ntx;
  float2* out = a.moments + ((((size_t)b * a.nwin + q) * G + g) * P) * nij;
  for (int ij0 = 0; ij0 < nij; ij0 += blockDim.x) {
    const int ij = ij0 + threadIdx.x;
    const int i = ij / a.ntx, j = ij - i * a.ntx;
    float accr[P], acci[P];
#pragma unroll
    for (int p = 0; p < P; ++p) accr[p] = acci[p] = 0.f;
    for (int c0 = t0; c0 < t1; c0 += 64) {
      __syncthreads();
      if (threadIdx.x < 64 && c0 + threadIdx.x < t1) {
        const int t = cdl_term_order(a, tb, b, c0 + threadIdx.x);
        const double w = a.w[(size_t)b * a.Rt + t];
        double turns = w * shift * kInvTwoPi;
        turns -= rint(turns);
        double s, c;
        sincospi(2.0 * turns, &s, &c);
        const double2 al = a.alpha[(size_t)b * a.Rt + t];
        beta[threadIdx.x] = make_float2((float)(al.x * c - al.y * s), (float)(al.x * s + al.y * c));
        ut[threadIdx.x] = (float)(w * (double)a.ptile);
      }
      __syncthreads();
      if (ij < nij) {
        const int nc = min(64, t1 - c0);
        for (int k = 0; k < nc; ++k) {
          const int t = cdl_term_order(a, tb, b, c0 + k);
          const double2 uv = ray_entry(a, b, t, i, j);
          const float uvr = (float)uv.x, uvi = (float)uv.y;
          float tr = beta[k].x * uvr - beta[k].y * uvi, ti = beta[k].x * uvi + beta[k].y * uvr;
          accr[0] += tr;
          acci[0] += ti;
#pragma unroll
          for (int p = 1; p < P; ++p) {
            const float f = ut[k] * (1.0f / (float)p);
            const float nr = -ti * f, ni = tr * f;
            tr = nr;
            ti = ni;
            accr[p] += tr;
            acci[p] += ti;
          }
        }
      }
    }
    if (ij < nij) {
#pragma unroll
      for (int p = 0; p < P; ++p) out[(size_t)p * nij + ij] = make_float2(accr[p], acci[p]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K5b, second form: one CTA per (link, delay group) forms the moments of ALL Taylor windows of the frame.
//   gamma[t][q][p] = alpha_t e^{j w_t (centre_q - k_g)} (j w_t tile)^p / p!     (phase reduced in FP64, stored FP32)
//   M[q][g][p][i][j] = sum_{t in g} gamma[t][q][p] * (u_t[i] v_t[j])           (FP32 FMA pipe)
// The ray product u_i v_j is formed once per term (FP64, rounded once to FP32) instead of once per (term, window).
// Windows are a.ptile samples long (a multiple of the K6 tile).  grid = B * G, 128 threads over (i, j); dynamic shared memory: gamma[kMomTerms][kMomWin][P] | us | vs.
constexpr int kMomTerms = 32;  // ray terms staged per pass
constexpr int kMomWin = 4;     // Taylor windows accumulated per pass

// HET: the batch carries per-link delay tables (a.link_tab); the uniform instantiation reads parameter space only.
template <int P, bool HET>
__global__ void __launch_bounds__(128) cdl_moment_all_kernel(const CdlArgs a, const __grid_constant__ CdlTable tb) {
  extern __shared__ __align__(16) unsigned char mom_smem[];
  float2* gam = reinterpret_cast<float2*>(mom_smem);                 // [kMomTerms][kMomWin][P]
  double2* us = reinterpret_cast<double2*>(gam + kMomTerms * kMomWin * P);  // [kMomTerms][nrx * rank]
  double2* vs = us + kMomTerms * a.nrx * a.rank;                            // [kMomTerms][ntx * rank]
  const int G = tb.num_groups;
  const int b = blockIdx.x / G, g = blockIdx.x - b * G;
  const int t0 = HET ? (int)a.link_tab[b].group_start[g] : (int)tb.group_start[g],
            t1 = HET ? (int)a.link_tab[b].group_start[g + 1] : (int)tb.group_start[g + 1];
  const uint16_t* lorder = HET ? a.link_tab[b].term_order : nullptr;
  auto term_at = [&](int c) { return HET ? (int)lorder[c] : (int)tb.term_order[c]; };
  const int gdelay = HET ? a.link_tab[b].group_delay[g] : tb.group_delay[g];
  const int nij = a.nrx * a.ntx;
  const int nu = a.nrx * a.rank, nv = a.ntx * a.rank;
  const int tid = threadIdx.x;
  for (int q0 = 0; q0 < a.nwin; q0 += kMomWin) {
    const int nq = min(kMomWin, a.nwin - q0);
    for (int ij0 = 0; ij0 < nij; ij0 += 128) {
      const int ij = ij0 + tid;
      const int i = ij / a.ntx, j = ij - i * a.ntx;
      float2 acc[kMomWin][P];
#pragma unroll
      for (int q = 0; q < kMomWin; ++q)
#pragma unroll
        for (int p = 0; p < P; ++p) acc[q][p] = make_float2(0.f, 0.f);
      for (int c0 = t0; c0 < t1; c0 += kMomTerms) {
        const int nc = min(kMomTerms, t1 - c0);
        __syncthreads();
        for (int e = tid; e < nc * nq; e += 128) {
          const int k = e / nq, q = e - k * nq;
          const int t = term_at(c0 + k);
          const double w = a.w[(size_t)b * a.Rt + t];
          const double shift = (double)(q0 + q) * a.ptile + 0.5 * a.ptile - (double)gdelay;
          double turns = w * shift * kInvTwoPi;
          turns -= rint(turns);
          double sn, cs;
          sincospi(2.0 * turns, &sn, &cs);
          const double2 al = a.alpha[(size_t)b * a.Rt + t];
          float tr = (float)(al.x * cs - al.y * sn), ti = (float)(al.x * sn + al.y * cs);
          const float ut = (float)(w * (double)a.ptile);
          float2* dst = gam + (k * kMomWin + q) * P;
          dst[0] = make_float2(tr, ti);
#pragma unroll
          for (int p = 1; p < P; ++p) {
            const float f = ut * (1.0f / (float)p);
            const float nr = -ti * f, ni = tr * f;
            tr = nr;
            ti = ni;
            dst[p] = make_float2(tr, ti);
          }
        }
        for (int e = tid; e < nc * nu; e += 128) {
          const int k = e / nu, c = e - k * nu;
          us[k * nu + c] = a.u[((size_t)b * a.Rt + term_at(c0 + k)) * nu + c];
        }
        for (int e = tid; e < nc * nv; e += 128) {
          const int k = e / nv, c = e - k * nv;
          vs[k * nv + c] = a.v[((size_t)b * a.Rt + term_at(c0 + k)) * nv + c];
        }
        __syncthreads();
        if (ij < nij) {
          auto accumulate = [&](auto nq_tag) {  // fully unrolled over the windows of this pass
            constexpr int NQ = decltype(nq_tag)::value;
            for (int k = 0; k < nc; ++k) {
              // the ray product in FP64, rounded once (an FP32 product of FP32-rounded phases costs the small-array f32 path
              // a factor two in accuracy: the reference's own 6-decimal unit tests notice)
              double2 uvd = cmul(us[k * nu + a.rank * i], vs[k * nv + a.rank * j]);
              if (a.rank == 2) {
                const double2 r1 = cmul(us[k * nu + 2 * i + 1], vs[k * nv + 2 * j + 1]);
                uvd.x += r1.x;
                uvd.y += r1.y;
              }
              const float2 uv = make_float2((float)uvd.x, (float)uvd.y);
              const float2* gk = gam + k * kMomWin * P;
#pragma unroll
              for (int q = 0; q < NQ; ++q)
#pragma unroll
                for (int p = 0; p < P; ++p) cmac<float>(acc[q][p], gk[q * P + p], uv);
            }
          };
          switch (nq) {  // CTA-uniform
            case 1: accumulate(std::integral_constant<int, 1>{}); break;
            case 2: accumulate(std::integral_constant<int, 2>{}); break;
            case 3: accumulate(std::integral_constant<int, 3>{}); break;
            default: accumulate(std::integral_constant<int, 4>{}); break;
          }
        }
      }
      if (ij < nij) {
#pragma unroll
        for (int q = 0; q < kMomWin; ++q) {
          if (q < nq) {
            float2* out = a.moments + ((((size_t)b * a.nwin + q0 + q) * G + g) * P) * nij + ij;
#pragma unroll
            for (int p = 0; p < P; ++p) out[(size_t)p * nij] = acc[q][p];
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K6: grid = B * ntiles, THREADS threads, R consecutive-stride samples per thread, NRX receive rows per launch.
//   smem: xs[kCdlTxChunk][W] | ms[G][P][NRX][kCdlTxChunk]
template <int NRX, int P, int R, int THREADS, typename IO>
__global__ void __launch_bounds__(THREADS) cdl_poly_kernel(const CdlArgs a, const __grid_constant__ CdlTable tb) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x / a.ntiles, q = blockIdx.x - b * a.ntiles, tid = threadIdx.x;
  const int W = a.tile + a.Dpad;
  const int G = tb.num_groups;
  float2* xs = reinterpret_cast<float2*>(smem_raw);
  float2* ms = xs + kCdlTxChunk * W;
  const int Tout = a.T + a.D;
  const int nij = a.nrx * a.ntx;
  const IO* xb = reinterpret_cast<const IO*>(a.x) + (size_t)b * a.ntx * a.T;
  const float2* mb = a.moments + (((size_t)b * a.ntiles + q) * G) * P * nij;
  const int n0 = q * a.tile - a.Dpad;

  float2 acc[R][P][NRX];
#pragma unroll
  for (int u = 0; u < R; ++u)
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
      for (int i = 0; i < NRX; ++i) acc[u][p][i] = make_float2(0.f, 0.f);

  for (int j0 = 0; j0 < a.ntx; j0 += kCdlTxChunk) {
    const int nj = min(kCdlTxChunk, a.ntx - j0);
    __syncthreads();
    for (int jj = 0; jj < nj; ++jj) {
      const IO* row = xb + (size_t)(j0 + jj) * a.T;
#pragma unroll 4
      for (int c = tid; c < W; c += THREADS) {
        const int n = n0 + c;
        float2 v = make_float2(0.f, 0.f);
        if (n >= 0 && n < a.T) v = to_c32(ldg_stream(row + n));
        xs[jj * W + c] = v;
      }
    }
    // ms[((g * P + p) * NRX + i) * chunk + jj] <- M[g][p][rx0 + i][j0 + jj]
    for (int c = tid; c < G * P * NRX * kCdlTxChunk; c += THREADS) {
      const int jj = c % kCdlTxChunk;
      const int i = (c / kCdlTxChunk) % NRX;
      const int gp = c / (kCdlTxChunk * NRX);
      float2 v = make_float2(0.f, 0.f);
      if (jj < nj && i < a.nrx_chunk) v = mb[(size_t)gp * nij + (size_t)(a.rx0 + i) * a.ntx + j0 + jj];
      ms[c] = v;
    }
    __syncthreads();

    for (int g = 0; g < G; ++g) {
      const int off = tid + a.Dpad - cdl_group_delay(a, tb, b, g);
      const float2* mg = ms + (size_t)g * P * NRX * kCdlTxChunk;
      for (int jj = 0; jj < nj; ++jj) {
        float2 xv[R];
#pragma unroll
        for (int u = 0; u < R; ++u) xv[u] = xs[jj * W + off + u * THREADS];
#pragma unroll
        for (int p = 0; p < P; ++p) {
#pragma unroll
          for (int i = 0; i < NRX; ++i) {
            const float2 m = mg[(p * NRX + i) * kCdlTxChunk + jj];
#pragma unroll
            for (int u = 0; u < R; ++u) cmac<float>(acc[u][p][i], m, xv[u]);
          }
        }
      }
    }
  }

  const float inv_tile = 1.0f / (float)a.tile, half = 0.5f * (float)a.tile;
#pragma unroll
  for (int u = 0; u < R; ++u) {
    const int il = u * THREADS + tid;
    const int m = q * a.tile + il;
    if (il < a.tile && m < Tout) {
      const float r = ((float)il - half) * inv_tile;
#pragma unroll
      for (int i = 0; i < NRX; ++i) {
        if (i < a.nrx_chunk) {
          float2 v = acc[u][P - 1][i];
#pragma unroll
          for (int p = P - 2; p >= 0; --p) {
            v.x = fmaf(v.x, r, acc[u][p][i].x);
            v.y = fmaf(v.y, r, acc[u][p][i].y);
          }
          IO* dst = reinterpret_cast<IO*>(a.y) + ((size_t)b * a.nrx + a.rx0 + i) * Tout + m;
          stg_stream(dst, IoConv<IO>::make(v.x, v.y));
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K6' FP64 parity mode: y[b, :, m] = sum_t alpha_t e^{j w_t (m - k_t)} u_t (v_t^T x[:, m - k_t]).
// One thread per (link, output sample); x read through L1/L2 (coalesced along time).
template <typename IO>
__global__ void __launch_bounds__(128) cdl_direct_f64_kernel(const CdlArgs a, const __grid_constant__ CdlTable tb) {
  const int Tout = a.T + a.D;
  const int nblk = (Tout + 127) / 128;
  const int b = blockIdx.x / nblk;
  const int m = (blockIdx.x - b * nblk) * 128 + threadIdx.x;
  if (m >= Tout) return;
  const IO* xb = reinterpret_cast<const IO*>(a.x) + (size_t)b * a.ntx * a.T;
  IO* yb = reinterpret_cast<IO*>(a.y) + (size_t)b * a.nrx * Tout + m;
  // accumulate receive rows in chunks of 8 to bound registers
  for (int i0 = 0; i0 < a.nrx; i0 += 8) {
    double2 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_double2(0.0, 0.0);
    for (int t = 0; t < a.Rt; ++t) {
      const int n = m - cdl_term_delay(a, tb, b, t);
      if (n < 0 || n >= a.T) continue;
      const double2* v = a.v + ((size_t)b * a.Rt + t) * a.ntx * a.rank;
      double2 s0 = make_double2(0.0, 0.0), s1 = make_double2(0.0, 0.0);
      if (a.rank == 1) {
        for (int j = 0; j < a.ntx; ++j) cmac<double>(s0, v[j], to_c64(xb[(size_t)j * a.T + n]));
      } else {
        for (int j = 0; j < a.ntx; ++j) {
          const double2 xv = to_c64(xb[(size_t)j * a.T + n]);
          cmac<double>(s0, v[2 * j], xv);
          cmac<double>(s1, v[2 * j + 1], xv);
        }
      }
      double sn, cs;
      sincos(a.w[(size_t)b * a.Rt + t] * (double)n, &sn, &cs);
      const double2 e = cmul(a.alpha[(size_t)b * a.Rt + t], make_double2(cs, sn));
      const double2 g0 = cmul(e, s0), g1 = cmul(e, s1);
      const double2* u = a.u + ((size_t)b * a.Rt + t) * a.nrx * a.rank;
      if (a.rank == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (i0 + i < a.nrx) cmac<double>(acc[i], u[i0 + i], g0);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (i0 + i < a.nrx) {
            cmac<double>(acc[i], u[2 * (i0 + i)], g0);
            cmac<double>(acc[i], u[2 * (i0 + i) + 1], g1);
          }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i0 + i < a.nrx) yb[(size_t)(i0 + i) * Tout] = IoConv<IO>::make(acc[i].x, acc[i].y);
  }
}

// ------------------------------------------------------------------------------------------------
// Channel state: H[b, g, i, j, n] = sum_{t in g} alpha_t e^{j w_t n} u_t[i] v_t[j]   (FP64 evaluation)
template <typename IO>
__global__ void __launch_bounds__(128) cdl_state_kernel(const CdlArgs a, const __grid_constant__ CdlTable tb) {
  const int nblk = (a.T + 127) / 128;
  const int nij = a.nrx * a.ntx;
  long long blk = blockIdx.x;
  const int nb = (int)(blk % nblk);
  blk /= nblk;
  const int ij = (int)(blk % nij);
  blk /= nij;
  const int g = (int)(blk % tb.num_groups);
  const int b = (int)(blk / tb.num_groups);
  const int n = nb * 128 + threadIdx.x;
  if (n >= a.T) return;
  const int i = ij / a.ntx, j = ij - i * a.ntx;
  double2 h = make_double2(0.0, 0.0);
  for (int c = cdl_group_start(a, tb, b, g); c < cdl_group_start(a, tb, b, g + 1); ++c) {
    const int t = cdl_term_order(a, tb, b, c);
    double sn, cs;
    sincos(a.w[(size_t)b * a.Rt + t] * (double)n, &sn, &cs);
    const double2 uv = ray_entry(a, b, t, i, j);
    cmac<double>(h, cmul(a.alpha[(size_t)b * a.Rt + t], make_double2(cs, sn)), uv);
  }
  IO* out = reinterpret_cast<IO*>(a.y) + ((((size_t)b * tb.num_groups + g) * nij + ij) * a.T) + n;
  *out = IoConv<IO>::make(h.x, h.y);
}

}  // namespace hb
