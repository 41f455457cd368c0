// Building blocks shared by the structure-specialised kernels (spec_kernel.cuh: one walker per thread;
// spec_tile.cuh: warp-owned walker tiles): the kernel-parameter block, parameter reads, and the pieces
// the GENERATED shell walk calls with literal indices.  Compiled by NVRTC only (-DQMCB_SPEC) after a
// generated prelude that defines SPEC_* (spec.cu: prelude).
#pragma once

#ifndef SPEC_EEN_NTERM
#define SPEC_EEN_NTERM 0
#endif
#define SPEC_EEN_MAX 8           // QMCB_EEN_MAXTERM (include/qmcb.h)
struct SpecParams {
  double expc[8];
  double jee_w, jen_w, vnn;
  const double *etab_g;          // [QMCB_ETAB] 2^(j/QMCB_ETAB) (global; staged into shared memory)
  const double *dtab_g;          // the plan's double / int tables (tile kernels: MO weights, CI coefficients,
  const int *itab_g;             // occupation and configuration index tables; staged into shared memory)
  // Boys-Handy three-body Jastrow (run-time values; the term count is a compile-time constant)
  double een_a[SPEC_EEN_MAX], een_b[SPEC_EEN_MAX], een_a2[SPEC_EEN_MAX], een_b2[SPEC_EEN_MAX], een_c[SPEC_EEN_MAX];
  double v[SPEC_NV];
  static constexpr int nelec = SPEC_NE, nup = SPEC_NUP, ndown = SPEC_NDOWN, natom = SPEC_NATOM;
  static constexpr int use_jee = SPEC_USE_JEE, use_jen = SPEC_USE_JEN, gram_fma = SPEC_GRAM_FMA;
  static constexpr int een_nterm = SPEC_EEN_NTERM;
};

struct SpecVals {
  const SpecParams &P;
  int off;
  __device__ __forceinline__ double operator[](int i) const { return P.v[off + i]; }
};
struct SpecTab {
  const SpecParams &P;
  __device__ __forceinline__ SpecVals atoms() const { return SpecVals{P, SPEC_OFF_ATOM}; }
};

// ---- parameter reads.  NVVM hoists every load of a kernel parameter out of the walker and
// electron loops (they are loop invariant), which leaves ptxas with > 100 long-lived doubles: it
// spills them and feeds the FP64 pipe through LDL + R2UR instead of constant-bank operands
// (measured: 78 R2UR + 25 LDL per electron).  Reading the parameter with a volatile inline
// ld.param at the point of use keeps the load next to its consumer, where ptxas folds it into the
// c[0x0][offset] operand slot of DFMA/DMUL: no instruction, no register.
#define SPEC_V_BYTE0 432   // offsetof(SpecParams, v)
static_assert(sizeof(SpecParams) == SPEC_V_BYTE0 + 8 * SPEC_NV, "SpecParams layout");
// SPEC_KPFX: name prefix of the kernels of this translation unit ("spec_" one walker per thread,
// "spect_" warp tiles) - ld.param addresses a parameter through the kernel's own symbol
#ifndef SPEC_KPFX
#define SPEC_KPFX "spec_"
#endif
template <int MODE, int I>
__device__ __forceinline__ double spec_pv() {
  double v;
  if constexpr (MODE == MODE_PSI)
    asm volatile("ld.param.f64 %0, [" SPEC_KPFX "psi_param_0+%1];" : "=d"(v) : "n"(SPEC_V_BYTE0 + 8 * I));
  else if constexpr (MODE == MODE_ELOC)
    asm volatile("ld.param.f64 %0, [" SPEC_KPFX "eloc_param_0+%1];" : "=d"(v) : "n"(SPEC_V_BYTE0 + 8 * I));
  else if constexpr (MODE == MODE_GRAD)
    asm volatile("ld.param.f64 %0, [" SPEC_KPFX "grad_psi_param_0+%1];" : "=d"(v) : "n"(SPEC_V_BYTE0 + 8 * I));
  else if constexpr (MODE == MODE_BWD)
    asm volatile("ld.param.f64 %0, [" SPEC_KPFX "backward_param_0+%1];" : "=d"(v) : "n"(SPEC_V_BYTE0 + 8 * I));
  else if constexpr (MODE == MODE_BWD_ALL)
    asm volatile("ld.param.f64 %0, [" SPEC_KPFX "backward_all_param_0+%1];" : "=d"(v) : "n"(SPEC_V_BYTE0 + 8 * I));
  else
    asm volatile("ld.param.f64 %0, [" SPEC_KPFX "mh_param_0+%1];" : "=d"(v) : "n"(SPEC_V_BYTE0 + 8 * I));
  return v;
}
// AO channels contracted against the MO columns: psi / Metropolis 1 (ao), grad psi 4 (ao + gradient),
// E_L 2: ao and the folded kinetic channel K = lap ao + 2 grad ln J . grad ao + (lap J / J) ao
// (device.cuh: FoldJ) - B_kin = -1/2 K W needs two contractions instead of five.
template <int MODE>
__host__ __device__ constexpr int spec_nch() { return MODE == MODE_ELOC ? 2 : (MODE == MODE_GRAD ? 4 : 1); }
template <int MODE>
__host__ __device__ constexpr bool spec_deriv() { return MODE == MODE_ELOC || MODE == MODE_GRAD; }

// ---- building blocks the generated program calls (literal indices everywhere)
// exp(-a r^2) or exp(-a r) of one primitive, v[I] = -a.  Primitives of one atom with the same
// exponent (the s and p functions of a Pople SP shell) share ONE exponential: the generator emits
// the call once and hands the value to every primitive that uses it (spec.cu: walk).
// The lower clamp of the exponent (-708: 3e-308 instead of a denormal) is one integer min on the
// high word: more negative doubles have larger high words.  NaN arguments are the canonical
// positive NaN the FP64 pipe produces and pass through.
template <int MODE, int I>
__device__ __forceinline__ double spec_exp(const SpecParams &P, const double *et, double r2_or_r) {
  const double x = spec_pv<MODE, I>() * r2_or_r;
  const unsigned hi = min((unsigned)__double2hiint(x), 0xC0862000u);
  return exp_core(P, et, __hiloint2double((int)hi, __double2loint(x)));
}
// One primitive c exp(-a r^2) of a shell.  Every product of parameters is formed on the HOST
// (spec.cu: walk): v[I..I+4] = { -a, c, -2 a c, -, 4 a^2 c }, and the radial sums are kept as
//   S0 = sum c e,  S1 = sum (-2 a c) e,  T2 = sum (4 a^2 c) e        (e = exp(-a r^2))
// with grad R = S1 (x,y,z) and lap R = 3 S1 + T2 r^2 formed once per shell (spec_shell_end).  Each
// FP64 instruction then has exactly ONE parameter operand, which DFMA/DMUL read straight from the
// constant bank: 1 + 8 (exp) + 3 FP64-pipe instructions per primitive with derivatives.
template <int MODE, bool FIRST, int I>
__device__ __forceinline__ void spec_prim(double e, double &S0, double &S1, double &T2) {
  if (FIRST) S0 = spec_pv<MODE, I + 1>() * e; else S0 = fma(spec_pv<MODE, I + 1>(), e, S0);
  if (spec_deriv<MODE>()) {
    if (FIRST) S1 = spec_pv<MODE, I + 2>() * e; else S1 = fma(spec_pv<MODE, I + 2>(), e, S1);
  }
  if (MODE == MODE_ELOC) {
    if (FIRST) T2 = spec_pv<MODE, I + 4>() * e; else T2 = fma(spec_pv<MODE, I + 4>(), e, T2);
  }
}
// Other radial types (ADF-style bases; radial_functions.py:6-238,323-406), same conventions
// (grad R = S1 (x,y,z), lap R = S2); v[I..I+4] as written by spec.cu for the type:
//   sto_pure  c e^{-a r}:        { -a, c, -a c, a^2 c, - }   S0 = sum c e, S1 = sum(-a c) e, S2 = sum(a^2 c) e,
//                                 finished per shell: S2 += 2 S1 / r, S1 /= r
//   gto / sto c r^N e^{-a r^2 | -a r}: { -a, c, a, -, - } with the literal radial power N
template <int MODE, int I>
__device__ __forceinline__ void spec_prim_sto_pure(double e, double &S0, double &S1, double &S2) {
  S0 = fma(spec_pv<MODE, I + 1>(), e, S0);
  if (spec_deriv<MODE>()) S1 = fma(spec_pv<MODE, I + 2>(), e, S1);
  if (MODE == MODE_ELOC) S2 = fma(spec_pv<MODE, I + 3>(), e, S2);
}
template <int MODE, int I, bool GTO, int N>
__device__ __forceinline__ void spec_prim_power(double e, double r2, double r, double rinv, double &S0, double &S1,
                                                double &S2) {
  const double a = spec_pv<MODE, I + 2>();
  const double ce = spec_pv<MODE, I + 1>() * e;
  const double rn = ipow(r, N);
  S0 = fma(ce, rn, S0);
  if (spec_deriv<MODE>()) {
    const double nrnm2 = N == 0 ? 0.0 : N * rpow(r, rinv, N - 2);
    if (GTO) {
      S1 = fma(ce, nrnm2 - 2.0 * a * rn, S1);
      if (MODE == MODE_ELOC) S2 = fma(ce, nrnm2 * (N + 1) - 4.0 * a * N * rn + a * rn * (4.0 * a * r2 - 6.0), S2);
    } else {
      S1 = fma(ce, nrnm2 - a * rn * rinv, S1);
      if (MODE == MODE_ELOC) S2 = fma(ce, nrnm2 * (N + 1) - 2.0 * a * nrnm2 * r + a * rn * (a - 2.0 * rinv), S2);
    }
  }
}

// RT: 0 gto_pure, 1 gto, 2 sto_pure, 3 sto (QMCB_* radial types)
// per atom: gd = 2 grad ln J . (x,y,z)  (+ 3 for gto_pure, whose lap R = 3 S1 + T2 r^2 is never formed)
template <int MODE, int RT>
__device__ __forceinline__ double spec_gd(const FoldJ &fj, double x, double y, double z) {
  if (MODE != MODE_ELOC) return 0.0;
  return fma(fj.g2x, x, fma(fj.g2y, y, fma(fj.g2z, z, RT == 0 ? 3.0 : 0.0)));
}
// E_L: electron-nucleus potential -Z / r_eA of this (electron, atom), v[IZ] = Z (wf_base.py:72-95)
template <int MODE, int IZ, int RT>
__device__ __forceinline__ void spec_ven(double r2, double rinv, double &ven) {
  if (MODE == MODE_ELOC) ven = fma(-spec_pv<MODE, IZ>(), RT == 0 ? fast_rsqrt(r2) : rinv, ven);
}
// finishes the radial sums of a shell; E_L: returns the shell value of the folded kinetic channel
//   Wf = lap R + S1 (2 grad ln J . u) + (lap J / J) S0
template <int MODE, int RT>
__device__ __forceinline__ double spec_shell_end(double r2, double rinv, double gd, double lp, double S0, double &S1,
                                                 double &S2, double T2) {
  if (RT == 2 && spec_deriv<MODE>()) {
    if (MODE == MODE_ELOC) S2 = fma(2.0 * S1, rinv, S2);
    S1 *= rinv;
  }
  if (MODE != MODE_ELOC) return 0.0;
  if (RT == 0) return fma(S1, gd, fma(T2, r2, lp * S0));
  return fma(S1, gd, fma(lp, S0, S2));
}

#ifndef SPEC_MOW_SMEM
#define SPEC_MOW_SMEM 0     // 1: MO weights staged in shared memory (LDS) instead of the parameter block
#endif
#ifndef SPEC_MWLD
#define SPEC_MWLD SPEC_NMUP   // row stride of the shared-memory weights (tile kernels: the plan's padded, even stride)
#endif
#ifndef SPEC_EUNROLL
#define SPEC_EUNROLL 1      // electrons per trip of the electron loop
#endif
#ifndef SPEC_PREFETCH
#define SPEC_PREFETCH 0     // 1: every kernel cp.asyncs the next walker's coordinates while this one is computed
#endif
#ifndef SPEC_PREFETCH_ELOC
#define SPEC_PREFETCH_ELOC 1   // E_L kernel only (measured, LiH 1e6 walkers: 0.143 -> 0.137 ms; psi / Metropolis do not gain)
#endif
#ifndef SPEC_MINB_ELOC
#define SPEC_MINB_ELOC 3     // E_L: 168 registers keep the electron loop free of spills (measured 0.199 -> 0.194 ms)
#endif
template <int MODE, int AO, int NCH>
__device__ __forceinline__ void spec_emit(const double *mw, const double (&v)[NCH], double (&acc)[NCH][SPEC_NMUP]) {
  static_assert(NCH == spec_nch<MODE>(), "channel count");
  // parameter-gradient backward: the AO value of this electron is kept as well (mw = its AO row)
  // (accumulated: an AO may be a sum of several monomial components; the row is zeroed by the caller)
  if constexpr (MODE == MODE_BWD || MODE == MODE_BWD_ALL) const_cast<double *>(mw)[AO] += v[0];
  // (the MO weights of one AO are read once per column: SPEC_NMUP <= 8 one walker per thread, <= 16 warp tiles)
  double w[SPEC_NMUP];
  if (SPEC_MOW_SMEM && MODE != MODE_BWD && MODE != MODE_BWD_ALL) {
    // shared-memory weights: row stride SPEC_MWLD; when it is even the rows are 16-byte aligned: two per LDS.128
    if constexpr (SPEC_MWLD % 2 == 0) {
      const double2 *wr = reinterpret_cast<const double2 *>(mw + AO * SPEC_MWLD);
#pragma unroll
      for (int j = 0; j + 1 < SPEC_NMUP; j += 2) { const double2 t = wr[j >> 1]; w[j] = t.x; w[j + 1] = t.y; }
      if (SPEC_NMUP & 1) w[SPEC_NMUP - 1] = mw[AO * SPEC_MWLD + SPEC_NMUP - 1];
    } else {
#pragma unroll
      for (int j = 0; j < SPEC_NMUP; ++j) w[j] = mw[AO * SPEC_MWLD + j];
    }
  } else {
#define SPEC_W(J) if constexpr (SPEC_NMUP > J) w[J] = spec_pv<MODE, SPEC_OFF_MOW + AO * SPEC_NMUP + J>();
    SPEC_W(0) SPEC_W(1) SPEC_W(2) SPEC_W(3) SPEC_W(4) SPEC_W(5) SPEC_W(6) SPEC_W(7)
    SPEC_W(8) SPEC_W(9) SPEC_W(10) SPEC_W(11) SPEC_W(12) SPEC_W(13) SPEC_W(14) SPEC_W(15)
    static_assert(SPEC_NMUP <= 16, "at most 16 occupied MO columns per thread");
#undef SPEC_W
  }
#pragma unroll
  for (int j = 0; j < SPEC_NMUP; ++j)
#pragma unroll
    for (int c = 0; c < NCH; ++c) acc[c][j] = fma(v[c], w[j], acc[c][j]);
}

template <int MODE, int AO, int ISC, int NCH>
__device__ __forceinline__ void spec_s(const double *mw, double x, double y, double z, double S0, double S1, double Wf,
                                       const FoldJ &fj, double (&acc)[NCH][SPEC_NMUP]) {
  double v[NCH];
  v[0] = S0 * spec_pv<MODE, ISC>();
  if (MODE == MODE_ELOC) {
    v[NCH - 1] = Wf * spec_pv<MODE, ISC>();
  } else if (NCH > 1) {
    const double t = S1 * spec_pv<MODE, ISC>();
    v[1] = t * x; v[2] = t * y; v[3] = t * z;
  }
  spec_emit<MODE, AO>(mw, v, acc);
}

template <int MODE, int AO, int ISC, int NCH>
__device__ __forceinline__ void spec_p(const double *mw, double x, double y, double z, double S0, double S1, double Wf,
                                       const FoldJ &fj, double (&acc)[NCH][SPEC_NMUP]) {
  double v[NCH];
  const double R = S0 * spec_pv<MODE, ISC>();
  if (MODE == MODE_ELOC) {
    const double Wp = fma(2.0, S1, Wf) * spec_pv<MODE, ISC>();
    v[0] = R * x; v[NCH - 1] = fma(Wp, x, R * fj.g2x);
    spec_emit<MODE, AO>(mw, v, acc);
    v[0] = R * y; v[NCH - 1] = fma(Wp, y, R * fj.g2y);
    spec_emit<MODE, AO + 1>(mw, v, acc);
    v[0] = R * z; v[NCH - 1] = fma(Wp, z, R * fj.g2z);
    spec_emit<MODE, AO + 2>(mw, v, acc);
  } else if (NCH > 1) {
    const double t = S1 * spec_pv<MODE, ISC>();
    const double tx = t * x, ty = t * y, tz = t * z;
    v[0] = R * x; v[1] = fma(tx, x, R); v[2] = tx * y; v[3] = tx * z;
    spec_emit<MODE, AO>(mw, v, acc);
    v[0] = R * y; v[1] = ty * x; v[2] = fma(ty, y, R); v[3] = ty * z;
    spec_emit<MODE, AO + 1>(mw, v, acc);
    v[0] = R * z; v[1] = tz * x; v[2] = tz * y; v[3] = fma(tz, z, R);
    spec_emit<MODE, AO + 2>(mw, v, acc);
  } else {
    v[0] = R * x; spec_emit<MODE, AO>(mw, v, acc);
    v[0] = R * y; spec_emit<MODE, AO + 1>(mw, v, acc);
    v[0] = R * z; spec_emit<MODE, AO + 2>(mw, v, acc);
  }
}

template <int MODE, int AO, int ISC, int KK, int NCH>
__device__ __forceinline__ void spec_g(const double *mw, double x, double y, double z, double S0, double S1, double Wf,
                                       const FoldJ &fj, double (&acc)[NCH][SPEC_NMUP]) {
  double v[NCH];
  // literal powers: a few products
  if constexpr (MODE == MODE_ELOC) generic_component_fold(KK, spec_pv<MODE, ISC>(), x, y, z, S0, S1, Wf, fj, v);
  else generic_component<NCH>(KK, spec_pv<MODE, ISC>(), x, y, z, S0, S1, 0.0, v);
  spec_emit<MODE, AO>(mw, v, acc);
}

