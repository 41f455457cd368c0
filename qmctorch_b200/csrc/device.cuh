// Device-side building blocks shared by the fused and the operator-level kernels:
// shared-memory table staging, shell evaluation (radial x cartesian harmonic),
// Jastrow/potential terms per electron, small dense determinants/inverses.
#pragma once
// QMCB_SPEC: this header is also compiled by NVRTC as part of the structure-specialised kernel
// (spec_kernel.cuh); every function that reads the system description is a template over the
// description type, so that the specialised build can hand in compile-time constants.
#ifndef QMCB_SPEC
#include <cuda_runtime.h>

#include <cstdint>

#include "fused_args.h"
#include "plan.h"
#define QMCB_UNROLL
#else
#define QMCB_UNROLL _Pragma("unroll")
#endif

#define QMCB_EPS 1e-16
#ifndef QMCB_ETAB_REP
#define QMCB_ETAB_REP 1
#endif

#ifndef QMCB_SPEC
// Views into the staged tables.  Only the two base pointers are kept live; every table pointer is
// re-derived from the offsets in the kernel-parameter struct (constant bank) where it is used, which
// keeps ~30 registers free in the hot loop.
struct Tab {
  const double *sd;
  const int *si;
  const DevSys *S;
#define QMCB_TAB_D(name, off) __device__ __forceinline__ const double *name() const { return sd + S->off; }
#define QMCB_TAB_I(name, off) __device__ __forceinline__ const int *name() const { return si + S->off; }
  QMCB_TAB_D(atoms, o_atoms) QMCB_TAB_D(mow, o_mow) QMCB_TAB_D(ci, o_ci) QMCB_TAB_D(etab, o_etab)
  QMCB_TAB_I(ash, o_ash) QMCB_TAB_I(used, o_used) QMCB_TAB_I(ucu, o_ucu) QMCB_TAB_I(ucd, o_ucd)
  QMCB_TAB_I(ciu, o_ciu) QMCB_TAB_I(cid, o_cid)
  __device__ __forceinline__ const double2 *stream() const {
    return reinterpret_cast<const double2 *>(sd + S->o_stream);
  }
};
// Copies both blobs into shared memory (all threads), returns the first free double slot.
__device__ __forceinline__ double *stage_tables(const DevSys &S, double *smem, Tab &T) {
  double *sd = smem;
  int *si = reinterpret_cast<int *>(smem + S.ndbl);
  for (int i = threadIdx.x; i < S.ndbl; i += blockDim.x) sd[i] = S.dblob[i];
  for (int i = threadIdx.x; i < S.nint; i += blockDim.x) si[i] = S.iblob[i];
  T.sd = sd; T.si = si; T.S = &S;
  return smem + S.ndbl + S.nint / 2;
}

__host__ __device__ inline int table_doubles(const DevSys &S) { return S.ndbl + S.nint / 2; }

#endif  // QMCB_SPEC

__device__ __forceinline__ double ipow(double x, int k) {
  switch (k) {
    case 0: return 1.0;
    case 1: return x;
    case 2: return x * x;
    default: {
      double r = x * x;
      for (int i = 2; i < k; ++i) r *= x;
      return r;
    }
  }
}

// 1/x and 1/sqrt(x) from the hardware seeds (MUFU.RCP64H / MUFU.RSQ64H, relative error e ~ 2^-22)
// plus ONE third-order correction: with the exact residual e, 1/x = y (1 + e + e^2 + O(e^3)) and
// 1/sqrt(x) = y (1 + e + 3/2 e^2 + O(e^3)); e^3 ~ 2^-66 is below the rounding of the result.
// 3 and 6 FP64 instructions (two Newton steps: 4 and 7), ~1 ulp, no slow-path branches.  x must be a
// normal positive number (distances, 1 + w r); NaN propagates, x = 0 gives inf/NaN like the exact
// operations would.
__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);
  return fma(y, fma(e, e, e), y);
}
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double h = 0.5 * x;
  const double e = fma(-h * y, y, 0.5);        // x y^2 = 1 - 2 e
  return fma(y * e, fma(1.5, e, 1.0), y);
}

// r^m for m >= -2 given r and 1/r
__device__ __forceinline__ double rpow(double r, double rinv, int m) {
  if (m >= 0) return ipow(r, m);
  return m == -1 ? rinv : rinv * rinv;
}

// ---------------------------------------------------------------------------------------
// exp(x), a few ulp, NaN propagates, no branches:
//   x = (128 m + j) ln2/128 + r, |r| <= ln2/256;  exp(x) = 2^m * T[j] * (1 + q(r)),
//   T[j] = 2^(j/128) from a 128-entry shared-memory table, q = r + r^2 (1/2 + r/6 + r^2/24)
//   (truncation r^5/120 <= 1.2e-15); the reduction uses ONE constant -ln2/128 (its rounding error,
//   4e-19 |128 m + j|, is below 1e-15 for |x| < 14); 2^m is applied by an integer add on the
//   exponent field.  8 FP64 instructions.  Arguments are clamped to [-708, 708] (3e-308 instead of
//   a denormal).  The constants come from the kernel-parameter struct (constant bank 0): S.expc.
// ---------------------------------------------------------------------------------------
template <class SYS>
__device__ __forceinline__ double exp_core(const SYS &S, const double *etab, double x) {
  const double t = fma(x, S.expc[0], S.expc[1]);
  const int ki = __double2loint(t);
#ifndef QMCB_EXP_I2F
#define QMCB_EXP_I2F 1
#endif
  // kd = the integer just read from the low word of t.  Converted back with I2F.F64.S32 it costs a slot of the
  // conversion (XU) pipe instead of a DADD on the FP64 pipe, which is the pipe these kernels are bound by
  // (QMCB_EXP_I2F=0: kd = t - 1.5 * 2^52, the classical form).
  const double kd = QMCB_EXP_I2F ? (double)ki : t - S.expc[1];
  const double r = fma(kd, S.expc[2], x);
  double p;
  if (QMCB_ETAB_LOG2 >= 10) {
    p = fma(r, S.expc[5], 0.5);
  } else {
    p = fma(r, S.expc[4], S.expc[5]);
    p = fma(p, r, 0.5);
  }
  const double q = fma(r * r, p, r);
  // QMCB_ETAB_REP > 1: the table is replicated REP times, entry j of replica c at [j * REP + c], and the
  // caller passes etab + (lane % REP): lanes of a half-warp that look up different entries then hit
  // different bank pairs (the random-index lookup of the unreplicated table costs ~5 wavefronts)
  const double tj = etab[(ki & (QMCB_ETAB - 1)) * QMCB_ETAB_REP];
  const double y = fma(tj, q, tj);
  return __hiloint2double(__double2hiint(y) + ((ki >> QMCB_ETAB_LOG2) << 20), __double2loint(y));
}

// arguments <= 0 (plus NaN, which passes): the lower clamp is one integer min on the high word -
// more negative doubles have larger high words - instead of DSETP + 2 FSEL
template <class SYS>
__device__ __forceinline__ double exp_neg(const SYS &S, const double *etab, double x) {
  const unsigned hi = min((unsigned)__double2hiint(x), 0xC0862000u);
  return exp_core(S, etab, __hiloint2double((int)hi, __double2loint(x)));
}

// arguments of either sign
template <class SYS>
__device__ __forceinline__ double exp_clamped(const SYS &S, const double *etab, double x) {
  x = x < -708.0 ? -708.0 : x;
  x = x > 708.0 ? 708.0 : x;
  return exp_core(S, etab, x);
}

// ---------------------------------------------------------------------------------------
// Shell evaluation.  For every shell the radial part is contracted first:
//   R(r) = S0,  grad R = S1 * (x,y,z),  lap R = S2        (spherically symmetric)
// then each cartesian component Y = x^kx y^ky z^kz (degree L) gives
//   ao = R Y,  grad ao = S1 Y (x,y,z) + R grad Y,  lap ao = (S2 + 2 L S1) Y + R lap Y
// (Euler: (x,y,z).grad Y = L Y).  Restates atomic_orbitals.py:578-669,
// radial_functions.py:6-406, spherical_harmonics.py:102-199.
// The basis is walked as a packed program of 16-byte records (plan.cu): no index tables,
// one LDS.128 per primitive / component group.
// Sink::emit(ao_index, v[NCH]) consumes the values: v[0]=ao, v[1..3]=grad, v[4]=lap.
// RT = 0: gto_pure (compile-time fast path), RT = 1: radial type read from S at run time.
// ---------------------------------------------------------------------------------------
// One primitive c exp(-a r^2):  S0 += c e,  S1 += a c e,  T2 += a^2 c e;  the shell is finished by
// grad R = -2 S1 (x,y,z),  lap R = -6 S1 + 4 r^2 T2  (gto_pure_finish): 6 FP64 instructions + exp.
template <int NCH, class SYS>
__device__ __forceinline__ void gto_pure_prim(const SYS &S, const double *etab, double a, double c, double r2, double &S0, double &S1,
                                              double &T2) {
  const double ce = c * exp_neg(S, etab, -a * r2);
  S0 += ce;
  if (NCH > 1) {
    const double t = a * ce;
    S1 += t;
    if (NCH > 4) T2 = fma(a, t, T2);
  }
}
template <int NCH>
__device__ __forceinline__ void gto_pure_finish(double r2, double &S1, double &S2, double T2) {
  if (NCH > 1) {
    if (NCH > 4) S2 = fma(4.0 * r2, T2, -6.0 * S1);
    S1 *= -2.0;
  }
}

template <int NCH, int RT, class SYS>
__device__ __forceinline__ const double2 *radial_sums(const SYS &S, const double *etab, const double2 *rec,
                                                      int nprim, double r2,
                                                      double r, double rinv, double &S0, double &S1,
                                                      double &S2) {
  if (RT == 0 && nprim == 1) {
    // single-primitive shell (most polarisation / diffuse shells): no accumulators, no finish
    const double2 p0 = rec[0];
    const double ce = p0.y * exp_neg(S, etab, -p0.x * r2);
    S0 = ce; S1 = 0.0; S2 = 0.0;
    if (NCH > 1) {
      const double t = p0.x * ce;
      S1 = -2.0 * t;
      if (NCH > 4) S2 = t * fma(4.0 * p0.x, r2, -6.0);
    }
    return rec + 1;
  }
  S0 = 0.0; S1 = 0.0; S2 = 0.0;
  if (RT == 0) {
    int i = 0;
    double T0 = 0.0, T1 = 0.0, T2 = 0.0, U2 = 0.0;      // second accumulator set: two independent chains
    for (; i + 2 <= nprim; i += 2) {
      const double2 p0 = rec[0], p1 = rec[1];
      rec += 2;
      gto_pure_prim<NCH>(S, etab, p0.x, p0.y, r2, S0, S1, U2);
      gto_pure_prim<NCH>(S, etab, p1.x, p1.y, r2, T0, T1, T2);
    }
    if (i < nprim) {
      const double2 p0 = rec[0];
      rec += 1;
      gto_pure_prim<NCH>(S, etab, p0.x, p0.y, r2, S0, S1, U2);
    }
    S0 += T0; S1 += T1; T2 += U2;
    gto_pure_finish<NCH>(r2, S1, S2, T2);
    return rec;
  }
  if (S.radial_type == QMCB_GTO_PURE) {
    double T2 = 0.0;
    for (int i = 0; i < nprim; ++i, ++rec) gto_pure_prim<NCH>(S, etab, rec->x, rec->y, r2, S0, S1, T2);
    gto_pure_finish<NCH>(r2, S1, S2, T2);
  } else if (S.radial_type == QMCB_STO_PURE) {
    for (int i = 0; i < nprim; ++i, ++rec) {
      const double a = rec->x;
      const double ce = rec->y * exp_neg(S, etab, -a * r);
      S0 += ce;
      if (NCH > 1) {
        const double t = a * ce;
        S1 -= t * rinv;
        if (NCH > 4) S2 += t * (a - 2.0 * rinv);
      }
    }
  } else {
    const bool gto = S.radial_type == QMCB_GTO;
    for (int i = 0; i < nprim; ++i, rec += 2) {
      const double a = rec[0].x;
      const int n = (int)rec[1].x;
      const double ce = rec[0].y * exp_neg(S, etab, gto ? -a * r2 : -a * r);
      const double rn = ipow(r, n);
      S0 += ce * rn;
      if (NCH > 1) {
        const double nrnm2 = n == 0 ? 0.0 : n * rpow(r, rinv, n - 2);
        if (gto) {
          S1 += ce * (nrnm2 - 2.0 * a * rn);
          if (NCH > 4) S2 += ce * (nrnm2 * (n + 1) - 4.0 * a * n * rn + a * rn * (4.0 * a * r2 - 6.0));
        } else {
          S1 += ce * (nrnm2 - a * rn * rinv);
          if (NCH > 4) S2 += ce * (nrnm2 * (n + 1) - 2.0 * a * nrnm2 * r + a * rn * (a - 2.0 * rinv));
        }
      }
    }
  }
  return rec;
}

// Cartesian component x^kx y^ky z^kz (degree L) of a shell: v[0] = ao, v[1..3] = grad, v[4] = lap
template <int NCH>
__device__ __forceinline__ void generic_component(int kk, double sc, double x, double y, double z, double S0,
                                                  double S1, double S2, double (&v)[NCH]) {
  const int kx = kk & 255, ky = (kk >> 8) & 255, kz = (kk >> 16) & 255;
  const int L = kx + ky + kz;
  const double px = ipow(x, kx), py = ipow(y, ky), pz = ipow(z, kz);
  const double Y = px * py * pz;
  const double R = S0 * sc;
  v[0] = R * Y;
  if (NCH > 1) {
    const double dYx = kx ? kx * ipow(x, kx - 1) * py * pz : 0.0;
    const double dYy = ky ? ky * px * ipow(y, ky - 1) * pz : 0.0;
    const double dYz = kz ? kz * px * py * ipow(z, kz - 1) : 0.0;
    const double t = S1 * sc * Y;
    v[1] = t * x + R * dYx;
    v[2] = t * y + R * dYy;
    v[3] = t * z + R * dYz;
    if (NCH > 4) {
      double lapY = 0.0;
      if (kx > 1) lapY += kx * (kx - 1) * ipow(x, kx - 2) * py * pz;
      if (ky > 1) lapY += ky * (ky - 1) * px * ipow(y, ky - 2) * pz;
      if (kz > 1) lapY += kz * (kz - 1) * px * py * ipow(z, kz - 2);
      v[4] = (S2 + 2.0 * L * S1) * sc * Y + R * lapY;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Folded kinetic channel (local energy only).  The Jacobi kinetic operator needs, per electron e,
//   B_kin[e][m] = -1/2 sum_a (lap ao_a + 2 g . grad ao_a + l ao_a) W[a][m],   g = grad_e ln J, l = lap_e J / J
// (slater_jastrow.py:449-482).  The reference projects lap ao, the three grad ao and ao separately
// (5 contractions) and combines afterwards; the combination is linear in the AO channels, so it is
// formed PER AO before the projection:  K_a = lap ao_a + 2 g . grad ao_a + l ao_a  - two channels
// (ao, K) are contracted instead of five.  With ao = sc R(r) Y, Y of degree L, u = (x,y,z):
//   K = sc Y (S2 + 2 L S1 + S1 (2 g . u) + l S0) + sc S0 (2 g . grad Y + lap Y)
// so one value per SHELL, Wf = S2 + S1 (2 g . u) + l S0, serves all its components.
// ---------------------------------------------------------------------------------------
struct FoldJ { double g2x, g2y, g2z, lp; };   // 2 grad_e ln J, lap_e J / J

__device__ __forceinline__ void generic_component_fold(int kk, double sc, double x, double y, double z, double S0,
                                                       double S1, double Wf, const FoldJ &f, double (&v)[2]) {
  const int kx = kk & 255, ky = (kk >> 8) & 255, kz = (kk >> 16) & 255;
  const int L = kx + ky + kz;
  const double px = ipow(x, kx), py = ipow(y, ky), pz = ipow(z, kz);
  const double Y = px * py * pz;
  const double R = S0 * sc;
  v[0] = R * Y;
  double t = 0.0;    // 2 g . grad Y + lap Y
  if (kx) t = fma(f.g2x, kx * ipow(x, kx - 1) * py * pz, t);
  if (ky) t = fma(f.g2y, ky * px * ipow(y, ky - 1) * pz, t);
  if (kz) t = fma(f.g2z, kz * px * py * ipow(z, kz - 1), t);
  if (kx > 1) t += kx * (kx - 1) * ipow(x, kx - 2) * py * pz;
  if (ky > 1) t += ky * (ky - 1) * px * ipow(y, ky - 2) * pz;
  if (kz > 1) t += kz * (kz - 1) * px * py * ipow(z, kz - 2);
  v[1] = fma(fma(2.0 * L, S1, Wf) * sc, Y, R * t);
}

// NCH = 1: ao;  4: ao + gradient;  5: + Laplacian;  FOLD (NCH == 2): ao and the folded kinetic channel
template <int NCH, int RT, bool FOLD = false, class Sink, class SYS, class TAB>
__device__ __forceinline__ void eval_aos(const SYS &S, const TAB &T, double ex, double ey, double ez,
                                         Sink &sink, const FoldJ fj = FoldJ{0.0, 0.0, 0.0, 0.0}) {
  static_assert(FOLD == (NCH == 2), "two channels <=> folded kinetic channel");
  constexpr int RD = FOLD ? 5 : NCH;        // radial derivative level
  const double2 *rec = T.stream();
  const double *et = T.etab();          // hoisted: one live pointer instead of a re-derivation per exp
  const double *at = T.atoms();
  const int *ash = T.ash();
  for (int A = 0; A < S.natom; ++A) {
    const double x = ex - at[4 * A], y = ey - at[4 * A + 1], z = ez - at[4 * A + 2];
    const double r2 = x * x + y * y + z * z;
    double r = 0.0, rinv = 0.0;
    if (RT != 0 && S.radial_type != QMCB_GTO_PURE) { rinv = fast_rsqrt(r2); r = r2 * rinv; }
    const double gd = FOLD ? fma(fj.g2x, x, fma(fj.g2y, y, fj.g2z * z)) : 0.0;
    const int ns = ash[A + 1] - ash[A];
    for (int s = 0; s < ns; ++s) {
      const double hdr = rec->x;
      ++rec;
      const int nprim = __double2loint(hdr), ngrp = __double2hiint(hdr);
      double S0, S1, S2;
      rec = radial_sums<RD, RT>(S, et, rec, nprim, r2, r, rinv, S0, S1, S2);
      const double Wf = FOLD ? fma(S1, gd, fma(fj.lp, S0, S2)) : 0.0;
      for (int g = 0; g < ngrp; ++g, ++rec) {
        const double2 gr = *rec;
        const int kk = __double2loint(gr.x), ao = __double2hiint(gr.x);
        const double sc = gr.y;
        double v[NCH];
        if (kk == 0) {                       // s
          v[0] = S0 * sc;
          if (FOLD) {
            v[NCH - 1] = Wf * sc;
          } else if (NCH > 1) {
            const double t = S1 * sc;
            v[1] = t * x; v[2] = t * y; v[3] = t * z;
            if (NCH > 4) v[NCH - 1] = S2 * sc;
          }
          sink.emit(ao, v);
        } else if (kk == (1 << 24)) {        // px, py, pz on consecutive AOs
          const double R = S0 * sc;
          if (FOLD) {
            const double Wp = fma(2.0, S1, Wf) * sc;
            v[0] = R * x; v[NCH - 1] = fma(Wp, x, R * fj.g2x);
            sink.emit(ao, v);
            v[0] = R * y; v[NCH - 1] = fma(Wp, y, R * fj.g2y);
            sink.emit(ao + 1, v);
            v[0] = R * z; v[NCH - 1] = fma(Wp, z, R * fj.g2z);
            sink.emit(ao + 2, v);
          } else if (NCH > 1) {
            const double t = S1 * sc, lf = NCH > 4 ? fma(2.0, S1, S2) * sc : 0.0;
            const double tx = t * x, ty = t * y, tz = t * z;
            v[0] = R * x; v[1] = fma(tx, x, R); v[2] = tx * y; v[3] = tx * z;
            if (NCH > 4) v[NCH - 1] = lf * x;
            sink.emit(ao, v);
            v[0] = R * y; v[1] = ty * x; v[2] = fma(ty, y, R); v[3] = ty * z;
            if (NCH > 4) v[NCH - 1] = lf * y;
            sink.emit(ao + 1, v);
            v[0] = R * z; v[1] = tz * x; v[2] = tz * y; v[3] = fma(tz, z, R);
            if (NCH > 4) v[NCH - 1] = lf * z;
            sink.emit(ao + 2, v);
          } else {
            v[0] = R * x; sink.emit(ao, v);
            v[0] = R * y; sink.emit(ao + 1, v);
            v[0] = R * z; sink.emit(ao + 2, v);
          }
        } else {
          if constexpr (FOLD) generic_component_fold(kk, sc, x, y, z, S0, S1, Wf, fj, v);
          else generic_component<NCH>(kk, sc, x, y, z, S0, S1, S2, v);
          sink.emit(ao, v);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Per-electron Jastrow / potential terms.  sp = this walker's positions [3*nelec] (shared).
// Output: g[3] = grad_e ln J, lap = (lap_e J)/J, ks = this electron's share of ln J
// (pairs j>e, all nuclei), ven, vee (pairs j>e).
// e-e: jastrow_factor_electron_electron.py:124-260, kernels/pade_jastrow_kernel.py:34-153,
//      distance/electron_electron_distance.py:47-190 (Gram-form r_ij reproduced exactly);
// e-n: jastrow_factor_electron_nuclei.py:60-161, kernels/pade_jastrow_kernel.py:36-116,
//      distance/electron_nuclei_distance.py:53-162;  product rule: combine_jastrow.py:116-195;
// potentials: wf_base.py:49-95.
// ---------------------------------------------------------------------------------------
// Gram-form distances exactly as the reference forms them (electron_electron_distance.py:177-190,
// electron_nuclei_distance.py:153-162)
__device__ __forceinline__ double gram_norm(double x, double y, double z) {
  return __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
}
template <class SYS>
__device__ __forceinline__ double gram_d2_ee(const SYS &S, double xi, double yi, double zi, double ni,
                                             double xj, double yj, double zj, double nj) {
  double dot;
  if (S.gram_fma) dot = __fma_rn(zi, zj, __fma_rn(yi, yj, __dmul_rn(xi, xj)));
  else dot = __dadd_rn(__dadd_rn(__dmul_rn(xi, xj), __dmul_rn(yi, yj)), __dmul_rn(zi, zj));
  return __dsub_rn(__dadd_rn(ni, nj), __dmul_rn(2.0, dot));
}
__device__ __forceinline__ double gram_d2_en(double xi, double yi, double zi, double ni, double xa, double ya,
                                             double za, double na) {
  const double dot = __fma_rn(zi, za, __fma_rn(yi, ya, __dmul_rn(xi, xa)));
  return __dsub_rn(__dadd_rn(ni, na), __dmul_rn(2.0, dot));
}

// Three-body Boys-Handy term for electron e (exponents 1):
//   ln J = sum_A sum_{i<j} sum_mu c f(r_iA) f(r_jA) g(r_ij),  f = a r/(1+b r),  g = a' r/(1+b' r)
// adds: ks (pairs j>e), grad_e ln J, lap_e ln J.  The reference differentiates this by autograd
// (jastrow_factor_electron_electron_nuclei.py:253-300,385-439); the closed forms use
// |grad r| = 1, lap r = 2/r:  lap_e = F''FG + F'FG 2/r_eA + FFG'' + FFG' 2/r_ej + 2 F'FG' cos(eA,ej).
template <bool DERIV, class SYS, class TAB>
__device__ __forceinline__ void een_terms(const SYS &S, const TAB &T, const double *sp, int e, double &gx,
                                          double &gy, double &gz, double &h, double &ks) {
  const int nt = S.een_nterm;
  const double xi = sp[3 * e], yi = sp[3 * e + 1], zi = sp[3 * e + 2];
  const double ni = gram_norm(xi, yi, zi);
  for (int j = DERIV ? 0 : e + 1; j < S.nelec; ++j) {
    if (j == e) continue;
    const double xj = sp[3 * j], yj = sp[3 * j + 1], zj = sp[3 * j + 2];
    const double nj = gram_norm(xj, yj, zj);
    const double d2 = gram_d2_ee(S, xi, yi, zi, ni, xj, yj, zj, nj);
    const double irej = fast_rsqrt(d2), rej = d2 * irej;
    const double ux = (xi - xj) * irej, uy = (yi - yj) * irej, uz = (zi - zj) * irej;
    for (int A = 0; A < S.natom; ++A) {
      const double xa = T.atoms()[4 * A], ya = T.atoms()[4 * A + 1], za = T.atoms()[4 * A + 2];
      const double na = gram_norm(xa, ya, za);
      const double dE2 = gram_d2_en(xi, yi, zi, ni, xa, ya, za, na);
      const double dJ2 = gram_d2_en(xj, yj, zj, nj, xa, ya, za, na);
      const double irE = fast_rsqrt(dE2), rE = dE2 * irE;
      const double rJ = dJ2 * fast_rsqrt(dJ2);
      const double vx = (xi - xa) * irE, vy = (yi - ya) * irE, vz = (zi - za) * irE;
      const double cos2 = 2.0 * (ux * vx + uy * vy + uz * vz);
      double s0 = 0, s1 = 0, s3 = 0, sl = 0;   // sum c FFG ; c F'FG ; c FFG' ; laplacian terms
      for (int m = 0; m < nt; ++m) {
        const double a = S.een_a[m], b = S.een_b[m], a2 = S.een_a2[m], b2 = S.een_b2[m], c = S.een_c[m];
        const double dE = fast_rcp(fma(b, rE, 1.0)), dJ = fast_rcp(fma(b, rJ, 1.0)), dG = fast_rcp(fma(b2, rej, 1.0));
        const double adE = a * dE, adG = a2 * dG;
        const double FE = adE * rE, G = adG * rej;
        const double cFJ = c * (a * rJ * dJ);
        s0 = fma(cFJ * FE, G, s0);
        if (DERIV) {
          const double FE1 = adE * dE, G1 = adG * dG;
          const double FE2 = -2.0 * b * FE1 * dE, G2 = -2.0 * b2 * G1 * dG;
          s1 = fma(cFJ * FE1, G, s1);
          s3 = fma(cFJ * FE, G1, s3);
          sl = fma(cFJ, fma(FE2, G, fma(FE, G2, FE1 * G1 * cos2)), sl);
        }
      }
      if (j > e) ks += s0;
      if (DERIV) {
        gx += s1 * vx + s3 * ux; gy += s1 * vy + s3 * uy; gz += s1 * vz + s3 * uz;
        h += sl + 2.0 * (s1 * irE + s3 * irej);
      }
    }
  }
}

struct ElecTerms {
  double gx, gy, gz, lap, ks, ven, vee;
};

// electron-nucleus part of electron_terms: V_en and the e-n Pade Jastrow terms of one electron
template <bool DERIV, bool POT, class SYS, class TAB>
__device__ __forceinline__ void nuclei_terms(const SYS &S, const TAB &T, double xi, double yi, double zi, double ni,
                                             double &gx, double &gy, double &gz, double &h, double &ks, double &ven) {
  double gnx = 0, gny = 0, gnz = 0;
  for (int A = 0; A < S.natom; ++A) {
    const double xa = T.atoms()[4 * A], ya = T.atoms()[4 * A + 1], za = T.atoms()[4 * A + 2];
    const double dx = xi - xa, dy = yi - ya, dz = zi - za;
    const double s2 = dx * dx + dy * dy + dz * dz;
    if (POT) ven -= T.atoms()[4 * A + 3] * fast_rsqrt(s2);
    if (S.use_jen) {
      const double wn = S.jen_w;
      const double na = __dadd_rn(__dadd_rn(__dmul_rn(xa, xa), __dmul_rn(ya, ya)), __dmul_rn(za, za));
      const double dot = __fma_rn(zi, za, __fma_rn(yi, ya, __dmul_rn(xi, xa)));
      const double d2n = __dsub_rn(__dadd_rn(ni, na), __dmul_rn(2.0, dot));
      const double r = d2n > 0.0 ? d2n * fast_rsqrt(d2n) : 0.0;   // electron on a nucleus: r = 0, kept finite by eps below
      const double den = fast_rcp(1.0 + wn * r);
      ks += r * den;
      if (DERIV) {
        const double invr = fast_rcp(r + QMCB_EPS);
        const double invr3 = fast_rcp(r * r * r + QMCB_EPS);
        const double kp = den * den * invr;
        gnx += kp * dx; gny += kp * dy; gnz += kp * dz;
        const double sdr2 = s2 * invr * invr;   // sum_c dr_c^2
        const double sd2r = 2.0 * s2 * invr3;   // sum_c d2r_c
        const double den2 = den * den;
        h += den * sd2r - 2.0 * wn * den2 * sdr2 - wn * r * den2 * sd2r +
             2.0 * wn * wn * r * den2 * den * sdr2;
      }
    }
  }
  gx += gnx; gy += gny; gz += gnz;
}

// POT_EN = false: the caller adds the electron-nucleus potential itself (specialised kernels take
// 1/r_eA from the basis-function loop, which forms the same distance anyway)
// EEN = false: the three-body term is left to the caller (warp tiles: een_table_*);  FINAL = false: o.lap
// returns the sum of the second-derivative terms WITHOUT |grad ln J|^2 (the caller adds further terms first)
template <bool DERIV, bool POT, bool POT_EN = POT, bool EEN = true, bool FINAL = true, class SYS, class TAB>
__device__ __forceinline__ void electron_terms(const SYS &S, const TAB &T, const double *sp, int e,
                                               ElecTerms &o) {
  const double xi = sp[3 * e], yi = sp[3 * e + 1], zi = sp[3 * e + 2];
  double gx = 0, gy = 0, gz = 0, h = 0, ks = 0, ven = 0, vee = 0;
  const double ni = __dadd_rn(__dadd_rn(__dmul_rn(xi, xi), __dmul_rn(yi, yi)), __dmul_rn(zi, zi));
  const bool up_i = e < S.nup;
  const double w = S.jee_w;
  for (int j = DERIV ? 0 : e + 1; j < S.nelec; ++j) {
    if (j == e) continue;
    const double xj = sp[3 * j], yj = sp[3 * j + 1], zj = sp[3 * j + 2];
    const double dx = xi - xj, dy = yi - yj, dz = zi - zj;
    const double s2 = dx * dx + dy * dy + dz * dz;
    if (POT && j > e) vee += fast_rsqrt(s2);
    if (S.use_jee) {
      const double nj = __dadd_rn(__dadd_rn(__dmul_rn(xj, xj), __dmul_rn(yj, yj)), __dmul_rn(zj, zj));
      double dot;
      if (S.gram_fma) dot = __fma_rn(zi, zj, __fma_rn(yi, yj, __dmul_rn(xi, xj)));
      else dot = __dadd_rn(__dadd_rn(__dmul_rn(xi, xj), __dmul_rn(yi, yj)), __dmul_rn(zi, zj));
      const double d2 = __dsub_rn(__dadd_rn(ni, nj), __dmul_rn(2.0, dot));
      const double rinv = fast_rsqrt(d2);
      const double r = d2 * rinv;
      const double w0 = (up_i == (j < S.nup)) ? 0.25 : 0.5;
      const double den = fast_rcp(1.0 + w * r);
      if (j > e) ks += w0 * r * den;
      if (DERIV) {
        const double kp = w0 * den * den * rinv;
        gx += kp * dx; gy += kp * dy; gz += kp * dz;
        h += 2.0 * kp * den * (s2 * rinv * rinv);
      }
    }
  }
  nuclei_terms<DERIV, POT_EN>(S, T, xi, yi, zi, ni, gx, gy, gz, h, ks, ven);
  if (EEN && S.een_nterm > 0) een_terms<DERIV>(S, T, sp, e, gx, gy, gz, h, ks);
  o.gx = gx; o.gy = gy; o.gz = gz;
  o.lap = FINAL ? h + gx * gx + gy * gy + gz * gz : h;
  o.ks = ks; o.ven = ven; o.vee = vee;
}

// Pair-once variant for shared tiles (derivative modes, no three-body term): the Ne electrons of a
// walker sit on Ne consecutive lanes of one warp (first lane: base).  In round k = 1..Ne/2 lane e
// evaluates the pair (e, e+k mod Ne) and receives the contribution of the pair (e-k mod Ne, e) from
// the lane that evaluated it (four double shuffles), so every pair is evaluated ONCE instead of once
// per electron; for even Ne the last round is shared out between the two halves.  Fixed order:
// deterministic.  All 32 lanes must call this (inactive lanes pass act = false).
template <bool POT, bool POT_EN = POT, bool FINAL = true, class SYS, class TAB>
__device__ __forceinline__ void electron_terms_paired(const SYS &S, const TAB &T, const double *sp, int e, int base,
                                                      bool act, ElecTerms &o) {
  const int Ne = S.nelec;
  const double xi = act ? sp[3 * e] : 0.0, yi = act ? sp[3 * e + 1] : 0.0, zi = act ? sp[3 * e + 2] : 0.0;
  double gx = 0, gy = 0, gz = 0, h = 0, ks = 0, ven = 0, vee = 0;
  const double ni = gram_norm(xi, yi, zi);
  const bool up_i = e < S.nup;
  const double w = S.jee_w;
  const int half = Ne >> 1;
  for (int k = 1; k <= half; ++k) {
    const bool halfround = 2 * k == Ne;
    int j = e + k; if (j >= Ne) j -= Ne;
    int src = e - k; if (src < 0) src += Ne;
    double px = 0, py = 0, pz = 0, hp = 0;
    if (act && (!halfround || e < half)) {
      const double xj = sp[3 * j], yj = sp[3 * j + 1], zj = sp[3 * j + 2];
      const double dx = xi - xj, dy = yi - yj, dz = zi - zj;
      const double s2 = dx * dx + dy * dy + dz * dz;
      if (POT) vee += fast_rsqrt(s2);
      const double d2 = gram_d2_ee(S, xi, yi, zi, ni, xj, yj, zj, gram_norm(xj, yj, zj));
      const double rinv = fast_rsqrt(d2);
      const double r = d2 * rinv;
      const double w0 = (up_i == (j < S.nup)) ? 0.25 : 0.5;
      const double den = fast_rcp(1.0 + w * r);
      ks += w0 * r * den;
      const double kp = w0 * den * den * rinv;
      px = kp * dx; py = kp * dy; pz = kp * dz;
      hp = 2.0 * kp * den * (s2 * rinv * rinv);
      gx += px; gy += py; gz += pz; h += hp;
    }
    const int sl = (base + src) & 31;
    const double qx = __shfl_sync(0xffffffffu, px, sl), qy = __shfl_sync(0xffffffffu, py, sl);
    const double qz = __shfl_sync(0xffffffffu, pz, sl), qh = __shfl_sync(0xffffffffu, hp, sl);
    if (act && (!halfround || e >= half)) { gx -= qx; gy -= qy; gz -= qz; h += qh; }
  }
  if (act) nuclei_terms<true, POT_EN>(S, T, xi, yi, zi, ni, gx, gy, gz, h, ks, ven);
  o.gx = gx; o.gy = gy; o.gz = gz;
  o.lap = FINAL ? h + gx * gx + gy * gy + gz * gz : h;
  o.ks = ks; o.ven = ven; o.vee = vee;
}

// ---------------------------------------------------------------------------------------
// Three-body Boys-Handy term through per-(electron, atom, term) factor tables (warp tiles, spec_tile.cuh).
// een_terms above recomputes f, f', f'' of BOTH electrons for every (partner, atom, term) - three
// reciprocals and two reciprocal square roots per innermost iteration, 85 % of the time of BASELINE
// config 4.  Here every lane first tabulates its own electron (een_table_fill):
//     tab[A * (1 + 3 NT)]            = 1 / r_eA
//     tab[A * (1 + 3 NT) + 1 + 3 m ..] = F, F', F''  with F = a r d, F' = a d^2, F'' = -2 a b d^3, d = 1 / (1 + b r)
// and the pair loop (een_table_terms) takes its own factors and the partner's F from the tables (shared
// memory, lane-major with an odd stride): one reciprocal per (pair, term) and none per atom.
// Same formulas and summation order as een_terms.
// ---------------------------------------------------------------------------------------
template <class SYS>
__host__ __device__ constexpr int een_table_doubles() { return (SYS::natom * (1 + 3 * SYS::een_nterm)) | 1; }

template <class SYS, class TAB>
__device__ __forceinline__ void een_table_fill(const SYS &S, const TAB &T, const double *sp, int e, double *tab) {
  constexpr int NT = SYS::een_nterm, NA = SYS::natom;
  const double xi = sp[3 * e], yi = sp[3 * e + 1], zi = sp[3 * e + 2];
  const double ni = gram_norm(xi, yi, zi);
  QMCB_UNROLL
  for (int A = 0; A < NA; ++A) {
    const double xa = T.atoms()[4 * A], ya = T.atoms()[4 * A + 1], za = T.atoms()[4 * A + 2];
    const double d2 = gram_d2_en(xi, yi, zi, ni, xa, ya, za, gram_norm(xa, ya, za));
    const double ir = fast_rsqrt(d2), r = d2 * ir;
    double *q = tab + A * (1 + 3 * NT);
    q[0] = ir;
    QMCB_UNROLL
    for (int m = 0; m < NT; ++m) {
      const double b = S.een_b[m];
      const double d = fast_rcp(fma(b, r, 1.0));
      const double ad = S.een_a[m] * d;
      const double F1 = ad * d;
      q[1 + 3 * m] = ad * r;
      q[2 + 3 * m] = F1;
      q[3 + 3 * m] = -2.0 * b * F1 * d;
    }
  }
}

// tabs: the walker's tables (lane of electron 0), stride = doubles between consecutive electrons
template <bool DERIV, class SYS, class TAB>
__device__ __forceinline__ void een_table_terms(const SYS &S, const TAB &T, const double *sp, int e,
                                                const double *tabs, int stride, double &gx, double &gy,
                                                double &gz, double &h, double &ks) {
  constexpr int NT = SYS::een_nterm, NA = SYS::natom, Ne = SYS::nelec;
  const double xi = sp[3 * e], yi = sp[3 * e + 1], zi = sp[3 * e + 2];
  const double ni = gram_norm(xi, yi, zi);
  const double *own = tabs + e * stride;
  for (int j = DERIV ? 0 : e + 1; j < Ne; ++j) {
    if (j == e) continue;
    const double xj = sp[3 * j], yj = sp[3 * j + 1], zj = sp[3 * j + 2];
    const double d2 = gram_d2_ee(S, xi, yi, zi, ni, xj, yj, zj, gram_norm(xj, yj, zj));
    const double irej = fast_rsqrt(d2), rej = d2 * irej;
    const double ux = (xi - xj) * irej, uy = (yi - yj) * irej, uz = (zi - zj) * irej;
    double G[NT], G1[NT], G2[NT];
    QMCB_UNROLL
    for (int m = 0; m < NT; ++m) {
      const double b2 = S.een_b2[m];
      const double dG = fast_rcp(fma(b2, rej, 1.0));
      const double adG = S.een_a2[m] * dG;
      G[m] = adG * rej;
      G1[m] = adG * dG;
      G2[m] = -2.0 * b2 * G1[m] * dG;
    }
    const double *par = tabs + j * stride;
    QMCB_UNROLL
    for (int A = 0; A < NA; ++A) {
      const double *qo = own + A * (1 + 3 * NT), *qp = par + A * (1 + 3 * NT);
      const double irE = qo[0];
      const double vx = (xi - T.atoms()[4 * A]) * irE, vy = (yi - T.atoms()[4 * A + 1]) * irE,
                   vz = (zi - T.atoms()[4 * A + 2]) * irE;
      const double cos2 = 2.0 * (ux * vx + uy * vy + uz * vz);
      double s0 = 0, s1 = 0, s3 = 0, sl = 0;
      QMCB_UNROLL
      for (int m = 0; m < NT; ++m) {
        const double FE = qo[1 + 3 * m];
        const double cFJ = S.een_c[m] * qp[1 + 3 * m];
        s0 = fma(cFJ * FE, G[m], s0);
        if (DERIV) {
          const double FE1 = qo[2 + 3 * m], FE2 = qo[3 + 3 * m];
          s1 = fma(cFJ * FE1, G[m], s1);
          s3 = fma(cFJ * FE, G1[m], s3);
          sl = fma(cFJ, fma(FE2, G[m], fma(FE, G2[m], FE1 * G1[m] * cos2)), sl);
        }
      }
      if (j > e) ks += s0;
      if (DERIV) {
        gx += s1 * vx + s3 * ux; gy += s1 * vy + s3 * uy; gz += s1 * vz + s3 * uz;
        h += sl + 2.0 * (s1 * irE + s3 * irej);
      }
    }
  }
}

// One-walker-per-thread variant: every electron pair is visited ONCE.  jv[k*jvs + e] receives
// gx, gy, gz, lap (k = 0..3, DERIV only); the walker totals of ln J, V_en, V_ee are returned.
// POT_EN = false: the caller adds the electron-nucleus potential itself (the specialised E_L kernel
// takes 1/r_eA from the basis-function loop, which forms the same distance anyway)
template <bool DERIV, bool POT, bool POT_EN = POT, class SYS, class TAB>
__device__ __forceinline__ void walker_terms(const SYS &S, const TAB &T, const double *sp, double *jv,
                                             int jvs, double &tks, double &tven, double &tvee) {
  const int Ne = S.nelec;
  if (DERIV)
    QMCB_UNROLL
    for (int e = 0; e < Ne; ++e) { jv[e] = 0.0; jv[jvs + e] = 0.0; jv[2 * jvs + e] = 0.0; jv[3 * jvs + e] = 0.0; }
  tks = 0.0; tven = 0.0; tvee = 0.0;
  const double w = S.jee_w;
  QMCB_UNROLL
  for (int i = 0; i < Ne; ++i) {
    const double xi = sp[3 * i], yi = sp[3 * i + 1], zi = sp[3 * i + 2];
    const double ni = gram_norm(xi, yi, zi);
    const bool up_i = i < S.nup;
    double gx = 0, gy = 0, gz = 0, h = 0, ks = 0, vee = 0, ven = 0;
    QMCB_UNROLL
    for (int j = i + 1; j < Ne; ++j) {
      const double xj = sp[3 * j], yj = sp[3 * j + 1], zj = sp[3 * j + 2];
      const double dx = xi - xj, dy = yi - yj, dz = zi - zj;
      const double s2 = dx * dx + dy * dy + dz * dz;
      if (POT) vee += fast_rsqrt(s2);
      if (S.use_jee) {
        // (the reference forms r_ij twice - Gram-form here, directly for the potential - and so do we: taking
        // 1/r_ij of the potential from the Gram value as well saves one reciprocal square root per pair, 2 % of
        // the LiH kernel, but carries the Gram form's cancellation error, up to 5e-12 Hartree per close pair,
        // into E_L: measured and not kept)
        const double d2 = gram_d2_ee(S, xi, yi, zi, ni, xj, yj, zj, gram_norm(xj, yj, zj));
        const double rinv = fast_rsqrt(d2);
        const double r = d2 * rinv;
        const double w0 = (up_i == (j < S.nup)) ? 0.25 : 0.5;
        const double den = fast_rcp(1.0 + w * r);
        ks += w0 * r * den;
        if (DERIV) {
          const double kp = w0 * den * den * rinv;
          const double px = kp * dx, py = kp * dy, pz = kp * dz;
          const double hp = 2.0 * kp * den * (s2 * rinv * rinv);
          gx += px; gy += py; gz += pz; h += hp;
          jv[j] -= px; jv[jvs + j] -= py; jv[2 * jvs + j] -= pz; jv[3 * jvs + j] += hp;
        }
      }
    }
    // nuclei
    QMCB_UNROLL
    for (int A = 0; A < S.natom; ++A) {
      const double xa = T.atoms()[4 * A], ya = T.atoms()[4 * A + 1], za = T.atoms()[4 * A + 2];
      const double dx = xi - xa, dy = yi - ya, dz = zi - za;
      const double s2 = dx * dx + dy * dy + dz * dz;
      if (POT_EN) ven -= T.atoms()[4 * A + 3] * fast_rsqrt(s2);
      if (S.use_jen) {
        const double wn = S.jen_w;
        const double d2n = gram_d2_en(xi, yi, zi, ni, xa, ya, za, gram_norm(xa, ya, za));
        const double r = d2n > 0.0 ? d2n * fast_rsqrt(d2n) : 0.0;   // electron on a nucleus: r = 0, kept finite by eps below
        const double den = fast_rcp(1.0 + wn * r);
        ks += r * den;
        if (DERIV) {
          const double invr = fast_rcp(r + QMCB_EPS);
          const double invr3 = fast_rcp(r * r * r + QMCB_EPS);
          const double kp = den * den * invr;
          gx += kp * dx; gy += kp * dy; gz += kp * dz;
          const double sdr2 = s2 * invr * invr, sd2r = 2.0 * s2 * invr3, den2 = den * den;
          h += den * sd2r - 2.0 * wn * den2 * sdr2 - wn * r * den2 * sd2r + 2.0 * wn * wn * r * den2 * den * sdr2;
        }
      }
    }
    if (DERIV) {
      gx += jv[i]; gy += jv[jvs + i]; gz += jv[2 * jvs + i]; h += jv[3 * jvs + i];
      jv[i] = gx; jv[jvs + i] = gy; jv[2 * jvs + i] = gz;
      jv[3 * jvs + i] = h + gx * gx + gy * gy + gz * gz;
    }
    tks += ks; tven += ven; tvee += vee;
  }
}

// ---------------------------------------------------------------------------------------
// Small dense LU-type kernels on one spin block (n x n), one thread per matrix.
// A(i,j) = mo[row0+i][cols[j]], B likewise.  Closed forms for n<=3; Gauss-Jordan with
// partial pivoting on an interleaved shared scratch otherwise (element stride `es`).
// Replaces torch.det / torch.inverse / btrace in slater_pooling.py:96-111,348-387,827-849.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void det_trace_small(int n, const double *A, const double *B, int ld,
                                                const int *cols, bool with_b, double &det, double &tr) {
  if (n == 1) {
    det = A[cols[0]];
    tr = with_b ? B[cols[0]] / det : 0.0;
  } else if (n == 2) {
    const double a00 = A[cols[0]], a01 = A[cols[1]], a10 = A[ld + cols[0]], a11 = A[ld + cols[1]];
    det = a00 * a11 - a01 * a10;
    if (with_b) {
      const double b00 = B[cols[0]], b01 = B[cols[1]], b10 = B[ld + cols[0]], b11 = B[ld + cols[1]];
      tr = (a11 * b00 - a01 * b10 - a10 * b01 + a00 * b11) / det;
    } else tr = 0.0;
  } else {  // n == 3
    double a[3][3], c[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) a[i][j] = A[i * ld + cols[j]];
    c[0][0] = a[1][1] * a[2][2] - a[1][2] * a[2][1];
    c[0][1] = a[1][2] * a[2][0] - a[1][0] * a[2][2];
    c[0][2] = a[1][0] * a[2][1] - a[1][1] * a[2][0];
    c[1][0] = a[0][2] * a[2][1] - a[0][1] * a[2][2];
    c[1][1] = a[0][0] * a[2][2] - a[0][2] * a[2][0];
    c[1][2] = a[0][1] * a[2][0] - a[0][0] * a[2][1];
    c[2][0] = a[0][1] * a[1][2] - a[0][2] * a[1][1];
    c[2][1] = a[0][2] * a[1][0] - a[0][0] * a[1][2];
    c[2][2] = a[0][0] * a[1][1] - a[0][1] * a[1][0];
    det = a[0][0] * c[0][0] + a[0][1] * c[0][1] + a[0][2] * c[0][2];
    tr = 0.0;
    if (with_b) {
      // inv(A)[j][i] = c[i][j]/det ; Tr(inv(A) B) = sum_ij inv[j][i] B[i][j] = sum_ij c[i][j] B[i][j]/det
      double t = 0.0;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) t += c[i][j] * B[i * ld + cols[j]];
      tr = t / det;
    }
  }
}

// inverse for n<=3 written to inv[i*n+j] (row-major), returns det
__device__ __forceinline__ double inverse_small(int n, const double *A, int ld, const int *cols, double *inv,
                                                int es) {
  if (n == 1) {
    const double d = A[cols[0]];
    inv[0] = 1.0 / d;
    return d;
  }
  if (n == 2) {
    const double a00 = A[cols[0]], a01 = A[cols[1]], a10 = A[ld + cols[0]], a11 = A[ld + cols[1]];
    const double det = a00 * a11 - a01 * a10, id = 1.0 / det;
    inv[0] = a11 * id; inv[es] = -a01 * id; inv[2 * es] = -a10 * id; inv[3 * es] = a00 * id;
    return det;
  }
  double a[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) a[i][j] = A[i * ld + cols[j]];
  const double c00 = a[1][1] * a[2][2] - a[1][2] * a[2][1];
  const double c01 = a[1][2] * a[2][0] - a[1][0] * a[2][2];
  const double c02 = a[1][0] * a[2][1] - a[1][1] * a[2][0];
  const double det = a[0][0] * c00 + a[0][1] * c01 + a[0][2] * c02, id = 1.0 / det;
  inv[0 * es] = c00 * id;
  inv[1 * es] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) * id;
  inv[2 * es] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * id;
  inv[3 * es] = c01 * id;
  inv[4 * es] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * id;
  inv[5 * es] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * id;
  inv[6 * es] = c02 * id;
  inv[7 * es] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) * id;
  inv[8 * es] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * id;
  return det;
}

// Gauss-Jordan with partial pivoting on [A | R] (n x (n+nr)), scratch element (i,j) at
// scr[(i*(n+nr)+j)*es].  On exit the right block holds inv(A) R.  Returns det(A).
__device__ inline double gauss_jordan(int n, int nr, double *scr, int es) {
  const int ldw = n + nr;
  double det = 1.0;
  for (int k = 0; k < n; ++k) {
    int piv = k;
    double best = fabs(scr[(k * ldw + k) * es]);
    for (int i = k + 1; i < n; ++i) {
      const double v = fabs(scr[(i * ldw + k) * es]);
      if (v > best) { best = v; piv = i; }
    }
    if (piv != k) {
      for (int j = 0; j < ldw; ++j) {
        const double t = scr[(k * ldw + j) * es];
        scr[(k * ldw + j) * es] = scr[(piv * ldw + j) * es];
        scr[(piv * ldw + j) * es] = t;
      }
      det = -det;
    }
    const double pv = scr[(k * ldw + k) * es];
    det *= pv;
    const double ip = fast_rcp(pv);
    for (int j = k + 1; j < ldw; ++j) scr[(k * ldw + j) * es] *= ip;
    for (int i = 0; i < n; ++i) {
      if (i == k) continue;
      const double f = scr[(i * ldw + k) * es];
      if (f != 0.0)
        for (int j = k + 1; j < ldw; ++j) scr[(i * ldw + j) * es] -= f * scr[(k * ldw + j) * es];
    }
  }
  return det;
}

// Warp-cooperative Gauss-Jordan with partial pivoting on [A | R] (n x (n+nr), row-major, leading
// dimension n+nr) in shared memory: lanes own columns, the pivot search is a warp arg-max
// (smallest index wins ties, like LAPACK).  On exit the right block holds inv(A) R.  Returns det(A)
// on every lane.  One warp per spin block: replaces the batched torch.det / torch.inverse of
// slater_pooling.py:96-111,348-387,827-849 for blocks larger than 3 x 3.
__device__ inline double warp_gauss_jordan(int n, int nr, double *m, int lane) {
  const int ldw = n + nr;
  double det = 1.0;
  for (int k = 0; k < n; ++k) {
    double best = -1.0;
    int piv = k;
    for (int i = k + lane; i < n; i += 32) {
      const double v = fabs(m[i * ldw + k]);
      if (v > best) { best = v; piv = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int op = __shfl_xor_sync(0xffffffffu, piv, o);
      if (ob > best || (ob == best && op < piv)) { best = ob; piv = op; }
    }
    if (piv != k) {
      for (int j = lane; j < ldw; j += 32) {
        const double t = m[k * ldw + j];
        m[k * ldw + j] = m[piv * ldw + j];
        m[piv * ldw + j] = t;
      }
      det = -det;
    }
    __syncwarp();
    const double pv = m[k * ldw + k];
    det *= pv;
    const double ip = fast_rcp(pv);
    __syncwarp();
    for (int j = k + 1 + lane; j < ldw; j += 32) m[k * ldw + j] *= ip;
    __syncwarp();
    for (int i = 0; i < n; ++i) {
      if (i == k) continue;
      const double f = m[i * ldw + k];
      for (int j = k + 1 + lane; j < ldw; j += 32) m[i * ldw + j] -= f * m[k * ldw + j];
    }
    __syncwarp();
  }
  return det;
}

// Register-resident LU with partial pivoting for N = 4..6 (fully unrolled): det(A) and, if
// WITH_B, Tr(A^-1 B) = sum_j (A^-1 b_j)_j by one forward/back substitution per column of B.
// Only A (N^2 doubles) lives in registers; B is read from shared memory through the row
// permutation.  A(i,j) = mo[row0+i][cols[j]], B likewise (ld = row stride).
// __noinline__: its register footprint must not leak into the callers' allocation.
template <int N, bool WITH_B>
__device__ __noinline__ void det_trace_reg(const double *A, const double *B, int ld, const int *cols,
                                           double &det_out, double &tr_out) {
  double a[N][N];
  int perm[N], cj[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    cj[j] = cols[j];
    perm[j] = j;
#pragma unroll
    for (int i = 0; i < N; ++i) a[i][j] = A[i * ld + cj[j]];
  }
  double det = 1.0;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    int piv = k;
    double best = fabs(a[k][k]);
#pragma unroll
    for (int i = k + 1; i < N; ++i) {
      const double v = fabs(a[i][k]);
      if (v > best) { best = v; piv = i; }
    }
#pragma unroll
    for (int i = k + 1; i < N; ++i) {
      const bool sw = piv == i;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const double t = a[k][j];
        a[k][j] = sw ? a[i][j] : t;
        a[i][j] = sw ? t : a[i][j];
      }
      const int tp = perm[k];
      perm[k] = sw ? perm[i] : tp;
      perm[i] = sw ? tp : perm[i];
    }
    if (piv != k) det = -det;
    det *= a[k][k];
    const double ip = fast_rcp(a[k][k]);
#pragma unroll
    for (int i = k + 1; i < N; ++i) {
      const double l = a[i][k] * ip;
      a[i][k] = l;                                   // L below the diagonal
#pragma unroll
      for (int j = k + 1; j < N; ++j) a[i][j] -= l * a[k][j];
    }
  }
  det_out = det;
  double tr = 0.0;
  if (WITH_B) {
    double inv_d[N];
#pragma unroll
    for (int i = 0; i < N; ++i) inv_d[i] = fast_rcp(a[i][i]);
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double x[N];
#pragma unroll
      for (int i = 0; i < N; ++i) {                   // forward: L y = P b_j
        double v = B[perm[i] * ld + cj[j]];
#pragma unroll
        for (int m = 0; m < i; ++m) v -= a[i][m] * x[m];
        x[i] = v;
      }
#pragma unroll
      for (int i = N - 1; i >= j; --i) {              // backward: U x = y, down to component j
        double v = x[i];
#pragma unroll
        for (int m = i + 1; m < N; ++m) v -= a[i][m] * x[m];
        x[i] = v * inv_d[i];
      }
      tr += x[j];
    }
  }
  tr_out = tr;
}

// Register-resident inverse for N = 4..6 (grad psi of CAS expansions: every unique spin block needs A^-1):
// the LU of det_trace_reg, then one forward/back substitution per unit vector.  inv(i,j) is written to
// m[(i * ldw + N + j) * es] - the [A | I] layout the shared-memory Gauss-Jordan leaves behind, so the
// epilogue reads either.  Returns det(A).
template <int N>
__device__ __noinline__ double inverse_reg(const double *A, int ld, const int *cols, double *m, int ldw, int es) {
  double a[N][N];
  int perm[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const int cj = cols[j];
    perm[j] = j;
#pragma unroll
    for (int i = 0; i < N; ++i) a[i][j] = A[i * ld + cj];
  }
  double det = 1.0;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    int piv = k;
    double best = fabs(a[k][k]);
#pragma unroll
    for (int i = k + 1; i < N; ++i) {
      const double v = fabs(a[i][k]);
      if (v > best) { best = v; piv = i; }
    }
#pragma unroll
    for (int i = k + 1; i < N; ++i) {
      const bool sw = piv == i;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const double t = a[k][j];
        a[k][j] = sw ? a[i][j] : t;
        a[i][j] = sw ? t : a[i][j];
      }
      const int tp = perm[k];
      perm[k] = sw ? perm[i] : tp;
      perm[i] = sw ? tp : perm[i];
    }
    if (piv != k) det = -det;
    det *= a[k][k];
    const double ip = fast_rcp(a[k][k]);
#pragma unroll
    for (int i = k + 1; i < N; ++i) {
      const double l = a[i][k] * ip;
      a[i][k] = l;
#pragma unroll
      for (int j = k + 1; j < N; ++j) a[i][j] -= l * a[k][j];
    }
  }
  double inv_d[N];
#pragma unroll
  for (int i = 0; i < N; ++i) inv_d[i] = fast_rcp(a[i][i]);
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double x[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {                     // forward: L y = P e_j
      double v = perm[i] == j ? 1.0 : 0.0;
#pragma unroll
      for (int q = 0; q < i; ++q) v -= a[i][q] * x[q];
      x[i] = v;
    }
#pragma unroll
    for (int i = N - 1; i >= 0; --i) {                // backward: U x = y
      double v = x[i];
#pragma unroll
      for (int q = i + 1; q < N; ++q) v -= a[i][q] * x[q];
      x[i] = v * inv_d[i];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) m[(i * ldw + N + j) * es] = x[i];
  }
  return det;
}

// Row-owner Gauss-Jordan for spin blocks of order n <= 16: TWO blocks per warp, one per half-warp.
// Lane hl of a half owns ROW hl of [A | R]: A in a[0..15], R in r[0..15].  The pivot COLUMN k is a
// compile-time loop index and the pivot ROW is a lane id, so no register is ever indexed dynamically
// (the column-owner variant below pays a select chain per row access: a third of the C4H6 E_L
// kernel, profiles/README.md).  Step k: arg-max |a[k]| over the rows not pivoted yet (four
// xor-shuffle rounds, ties to the lower row); the pivot lane broadcasts its row UNSCALED and every
// other row subtracts (a[k] / pivot) times it: two SHFL and one DFMA per element, no select.  Rows
// are neither swapped nor normalised: lane p remembers the column it pivoted (kc) and its pivot
// (pv); det = sign(kc) * prod pivots, and the lane with kc = k ends with pv * (row k of inv(A) R)
// in r[] - the caller divides.  All shuffles use the full mask (width 16): both halves run the same
// nmax = max order of the two blocks steps; a half whose block is smaller (or absent: n = 0) idles
// through the extra steps with zero multipliers.
template <bool WITH_R, int NP>
__device__ __forceinline__ double half_warp_gauss_jordan(int n, int nmax, double (&a)[NP], double (&r)[NP], int hl,
                                                         int &kc, double &ipiv) {
  // NP >= nmax: padded order (8, 12 or 16).  Padding columns hold zeros and are eliminated along with
  // the real ones: the inner loops carry no bounds test, so the NP independent shuffle + FMA chains
  // of a step overlap instead of running one basic block at a time.
  const unsigned full = 0xffffffffu;
  double det = 1.0;
  bool done = hl >= n;
  kc = 16 + hl;                      // rows beyond n: distinct, above every real column
  ipiv = 0.0;
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    if (k < nmax) {
      const bool on = k < n;
      // arg-max of |a[k]| over the rows not pivoted yet: ONE integer warp reduction per half-warp.  The
      // key is the high word of |a[k]| (for non-negative doubles the high word orders like the value:
      // sign 0 | exponent | top 20 mantissa bits) with 16 - row in its low 5 bits (ties -> lower row; never
      // 0 for a candidate), so a row within 2^-15 of the largest magnitude may be chosen instead of the
      // largest itself - an equally valid partial pivot (growth bound unchanged to 1 + 2^-15); NaN rows
      // compare as large and propagate.  Replaces four rounds of three shuffles + compare/select chains.
      unsigned key = 0u;
      if (!done && on) key = ((unsigned)__double2hiint(fabs(a[k])) & 0xffffffe0u) | (unsigned)(16 - hl);
      // (two full-mask reductions, one per half-warp: a half-mask reduction makes the compiler treat the
      // warp as diverged and wrap every later shuffle in WARPSYNC / ENDCOLLECTIVE)
      const bool upper = (threadIdx.x & 16) != 0;
      const unsigned k_lo = __reduce_max_sync(full, upper ? 0u : key);
      const unsigned k_hi = __reduce_max_sync(full, upper ? key : 0u);
      const unsigned kmax = upper ? k_hi : k_lo;
      const int l = kmax ? 16 - (int)(kmax & 31u) : 0;   // (0: this half has no active row left)
      const int p = l;
      const double pv = __shfl_sync(full, a[k], p, 16);
      double ipv;                      // 1 / pivot (either sign; pv = 0: inf, as the exact division)
      asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(ipv) : "d"(pv));
      { const double e = fma(-pv, ipv, 1.0); ipv = fma(ipv, fma(e, e, e), ipv); }
      { const double e = fma(-pv, ipv, 1.0); ipv = fma(ipv, e, ipv); }
      const bool me = on && hl == p;
      if (on) det *= pv;
      const double f = (me || !on) ? 0.0 : a[k] * ipv;      // multiplier of this row
#pragma unroll
      for (int j = k + 1; j < NP; ++j) a[j] = fma(-f, __shfl_sync(full, a[j], p, 16), a[j]);
      if (WITH_R) {
#pragma unroll
        for (int j = 0; j < NP; ++j) r[j] = fma(-f, __shfl_sync(full, r[j], p, 16), r[j]);
      }
      if (me) { done = true; kc = k; ipiv = ipv; }
    }
  }
  // sign of the permutation row -> pivoted column: parity of the number of inversions
  int inv = 0;
#pragma unroll
  for (int o = 0; o < 16; ++o) {
    const int oc = __shfl_sync(full, kc, o, 16);
    inv += (o < hl && oc > kc) ? 1 : 0;
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) inv += __shfl_xor_sync(full, inv, o, 16);
  return (inv & 1) ? -det : det;
}

// Register-resident warp Gauss-Jordan: lane j owns column j of [A | R] (n + nr <= 32, n <= NMAX),
// rows are registers.  Per elimination step: lane k finds the pivot in its registers, the pivot
// row index and the multipliers travel by warp shuffles - no shared-memory round trips.
// On exit lanes n..n+nr-1 hold the columns of inv(A) R.  Returns det(A) on every lane.
template <int NMAX>
__device__ __forceinline__ double warp_gauss_jordan_reg(int n, double (&col)[NMAX], int lane) {
  double det = 1.0;
#pragma unroll
  for (int k = 0; k < NMAX; ++k) {
    if (k < n) {
      double best = -1.0;
      int piv = k;
#pragma unroll
      for (int i = k; i < NMAX; ++i) {
        const double v = fabs(col[i]);
        if (i < n && v > best) { best = v; piv = i; }
      }
      piv = __shfl_sync(0xffffffffu, piv, k);
#pragma unroll
      for (int i = k + 1; i < NMAX; ++i) {
        const bool sw = piv == i;
        const double t = col[k];
        col[k] = sw ? col[i] : t;
        col[i] = sw ? t : col[i];
      }
      if (piv != k) det = -det;
      const double pv = __shfl_sync(0xffffffffu, col[k], k);
      det *= pv;
      const double pk = col[k] * (1.0 / pv);          // pivot row, scaled (meaningful for lanes > k)
#pragma unroll
      for (int i = 0; i < NMAX; ++i) {
        if (i != k && i < n) {
          const double f = __shfl_sync(0xffffffffu, col[i], k);   // element (i, k)
          if (lane > k) col[i] = fma(-f, pk, col[i]);
        }
      }
      if (lane > k) col[k] = pk;
    }
  }
  return det;
}
