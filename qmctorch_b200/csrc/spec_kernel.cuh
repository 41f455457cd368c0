// Structure-specialised walker kernel ("compiled walker program"), built at run time by NVRTC.
//
// The generic fused kernel (fused_impl.cuh) INTERPRETS the basis: it walks a packed shell program
// with loops, record loads and index arithmetic, and 60 % of its issued instructions are not
// FP64.  For small systems (one walker per thread, closed-form determinants) spec.cu generates
// the same program as straight-line code for ONE wave-function structure:
//   - the shell walk, the AO -> MO contraction, the determinants and the CI sum are emitted with
//     literal indices (no loops, no records, no index tables);
//   - every parameter (exponents, coefficients, MO weights, CI, Jastrow weights, nuclei) stays
//     RUN-TIME data in the kernel-parameter block P.v[...]: constant bank 0, a free operand of
//     DFMA, so an optimiser step only rewrites the parameter block - no recompilation;
//   - only the electron loop is kept rolled (instruction-cache footprint); per-electron rows go
//     through a private, odd-strided shared-memory slice exactly as in the generic THREAD tiles.
// The arithmetic is the generic kernel's (same device functions from device.cuh), so it is held
// to the same parity bar by the same tests.
//
// Compiled with -DQMCB_SPEC after a generated prelude that defines SPEC_* and the functions
// spec_aos<MODE>, spec_dets<WB>, spec_ci<MODE> (see spec.cu: walk).
#pragma once

// ---- the prelude has defined: SPEC_NE SPEC_NUP SPEC_NDOWN SPEC_NATOM SPEC_NMUP SPEC_NUU SPEC_NUD
//      SPEC_USE_JEE SPEC_USE_JEN SPEC_GRAM_FMA SPEC_NV SPEC_OFF_ATOM SPEC_OFF_MOW SPEC_OFF_CI
//      SPEC_THREADS SPEC_MINB

struct SpecParams {
  double expc[8];
  double jee_w, jen_w, vnn;
  const double *etab_g;          // [QMCB_ETAB] 2^(j/QMCB_ETAB) (global; staged into shared memory)
  double v[SPEC_NV];
  static constexpr int nelec = SPEC_NE, nup = SPEC_NUP, ndown = SPEC_NDOWN, natom = SPEC_NATOM;
  static constexpr int use_jee = SPEC_USE_JEE, use_jen = SPEC_USE_JEN, gram_fma = SPEC_GRAM_FMA;
  static constexpr int een_nterm = 0;
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
#define SPEC_V_BYTE0 96   // offsetof(SpecParams, v)
static_assert(sizeof(SpecParams) == SPEC_V_BYTE0 + 8 * SPEC_NV, "SpecParams layout");
template <int MODE, int I>
__device__ __forceinline__ double spec_pv() {
  double v;
  if constexpr (MODE == MODE_PSI)
    asm volatile("ld.param.f64 %0, [spec_psi_param_0+%1];" : "=d"(v) : "n"(SPEC_V_BYTE0 + 8 * I));
  else if constexpr (MODE == MODE_ELOC)
    asm volatile("ld.param.f64 %0, [spec_eloc_param_0+%1];" : "=d"(v) : "n"(SPEC_V_BYTE0 + 8 * I));
  else if constexpr (MODE == MODE_GRAD)
    asm volatile("ld.param.f64 %0, [spec_grad_psi_param_0+%1];" : "=d"(v) : "n"(SPEC_V_BYTE0 + 8 * I));
  else
    asm volatile("ld.param.f64 %0, [spec_mh_param_0+%1];" : "=d"(v) : "n"(SPEC_V_BYTE0 + 8 * I));
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
  // (the MO weights of one AO are read once per column: SPEC_NMUP <= 8)
  double w[SPEC_NMUP];
  if (SPEC_MOW_SMEM) {
#pragma unroll
    for (int j = 0; j < SPEC_NMUP; ++j) w[j] = mw[AO * SPEC_NMUP + j];
  } else {
#define SPEC_W(J) if constexpr (SPEC_NMUP > J) w[J] = spec_pv<MODE, SPEC_OFF_MOW + AO * SPEC_NMUP + J>();
    SPEC_W(0) SPEC_W(1) SPEC_W(2) SPEC_W(3) SPEC_W(4) SPEC_W(5) SPEC_W(6) SPEC_W(7)
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

// ---- generated: spec_aos<MODE>, spec_dets<WB>, spec_ci<MODE>, spec_grad<MODE>
SPEC_GENERATED_CODE

template <int MODE>
__device__ __forceinline__ void spec_body(const SpecParams &P, const FusedArgs &a) {
  constexpr int NCH = spec_nch<MODE>();
  constexpr int Ne = SPEC_NE, ne3 = 3 * SPEC_NE, NM = SPEC_NMUP, NUN = SPEC_NUU + SPEC_NUD;
  constexpr int NROW = MODE == MODE_ELOC ? 2 : (MODE == MODE_GRAD ? 4 : 1);   // mo | B_kin or mo | d mo/dx,dy,dz
  constexpr int SLICE = (3 * Ne + (spec_deriv<MODE>() ? 4 * Ne : 0) + NROW * Ne * NM) | 1;
  extern __shared__ __align__(16) double smem[];
  double *et = smem;
  for (int i = threadIdx.x; i < QMCB_ETAB; i += blockDim.x) et[i] = P.etab_g[i];
  constexpr int NMW = SPEC_MOW_SMEM ? ((SPEC_NV - SPEC_OFF_MOW + 1) & ~1) : 0;   // MO weights + CI, even
  double *mw = smem + QMCB_ETAB;
  for (int i = threadIdx.x; i < NMW; i += blockDim.x) mw[i] = i < SPEC_NV - SPEC_OFF_MOW ? P.v[SPEC_OFF_MOW + i] : 0.0;
  constexpr bool PF = SPEC_PREFETCH || (MODE == MODE_ELOC && SPEC_PREFETCH_ELOC);
  constexpr int SL = SLICE + (PF ? ne3 + (ne3 & 1) : 0);               // stays odd
  double *spos = smem + QMCB_ETAB + NMW + (size_t)threadIdx.x * SL;
  double *jv = spos + ne3;
  double *smo = jv + (spec_deriv<MODE>() ? 4 * Ne : 0);
  double *sB = smo + Ne * NM;
  double *snext = spos + SLICE;     // PF: landing zone of the next walker's coordinates
  const SpecTab T{P};
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t wfirst = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // cp.async (LDGSTS) copies one walker row global -> this thread's slice without registers
  auto prefetch = [&](int64_t wn) {
    if (PF && wn < a.W) {
      const unsigned dst = (unsigned)__cvta_generic_to_shared(snext);
      const double *src = a.pos + wn * ne3;
#pragma unroll
      for (int i = 0; i < ne3; ++i)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8 * i), "l"(src + i) : "memory");
    }
  };
  prefetch(wfirst);
  double st_s = 0.0, st_s2 = 0.0;     // fused energy statistics of this thread's walkers (E_L only)
  int st_nf = 0, st_nb = 0;
  for (int64_t w = wfirst; w < a.W; w += stride) {
    if (PF) asm volatile("cp.async.wait_all;" ::: "memory");
    // ---- coordinates (+ proposal)
    if (MODE == MODE_MH && !a.disp && a.proba_normal) {
      // one Philox call yields the four normals of a GLOBAL element quad (4q .. 4q+3): the draw of
      // an element does not depend on the tiling (same stream as the generic kernel)
      const int64_t g0 = w * ne3, g1 = g0 + ne3;
      int me = a.move_elec;
      if (me == -2) {
        if (a.elec_index) me = a.elec_index[w];
        else me = (int)(philox_u32(a.seed, a.offset, (uint64_t)w, 2u) % (unsigned)Ne);
      }
      for (int64_t q = g0 >> 2; 4 * q < g1; ++q) {
        double z[4];
        philox_normal4(a.seed, a.offset, (uint64_t)q, z);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const int64_t g = 4 * q + h;
          if (g < g0 || g >= g1) continue;
          const int i = (int)(g - g0), e = i / 3;
          double v = PF ? snext[i] : a.pos[g];
          if (me < 0 || me == e) v += a.scale * z[h];
          spos[i] = v;
        }
      }
    } else {
      int me = a.move_elec;
      if (MODE == MODE_MH && me == -2) {
        if (a.elec_index) me = a.elec_index[w];
        else me = (int)(philox_u32(a.seed, a.offset, (uint64_t)w, 2u) % (unsigned)Ne);
      }
#pragma unroll
      for (int i = 0; i < ne3; ++i) {
        double v = PF ? snext[i] : a.pos[w * ne3 + i];
        if (MODE == MODE_MH && (me < 0 || me == i / 3)) {
          double d;
          if (a.disp) d = a.disp[w * ne3 + i];
          else d = a.scale * (2.0 * philox_uniform(a.seed, a.offset, (uint64_t)(w * ne3 + i), 0u) - 1.0);
          v += d;
        }
        spos[i] = v;
      }
    }
    prefetch(w + stride);
    // ---- Jastrow gradient / Laplacian terms and potentials, every pair once
    double tks, tven, tvee;
    walker_terms<spec_deriv<MODE>(), (MODE == MODE_ELOC), false>(P, T, spos, jv, Ne, tks, tven, tvee);
    // ---- AO -> MO rows, one electron at a time (rolled: instruction-cache footprint)
#pragma unroll SPEC_EUNROLL
    for (int e = 0; e < Ne; ++e) {
      double acc[NCH][NM];
#pragma unroll
      for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int j = 0; j < NM; ++j) acc[c][j] = 0.0;
      FoldJ fj{0.0, 0.0, 0.0, 0.0};
      if (MODE == MODE_ELOC && (SPEC_USE_JEE || SPEC_USE_JEN))
        fj = FoldJ{2.0 * jv[e], 2.0 * jv[Ne + e], 2.0 * jv[2 * Ne + e], jv[3 * Ne + e]};
      spec_aos<MODE>(P, et, mw, spos[3 * e], spos[3 * e + 1], spos[3 * e + 2], fj, tven, acc);
      if (MODE == MODE_ELOC) {
        // B_kin = -1/2 (lap mo + 2 grad ln J . grad mo + (lap J / J) mo): the folded channel, projected
#pragma unroll
        for (int j = 0; j < NM; ++j) {
          smo[e * NM + j] = acc[0][j];
          sB[e * NM + j] = -0.5 * acc[NCH - 1][j];
        }
      } else if (MODE == MODE_GRAD) {
#pragma unroll
        for (int j = 0; j < NM; ++j) {
          smo[e * NM + j] = acc[0][j];
#pragma unroll
          for (int c = 1; c < NCH; ++c) sB[(c - 1) * Ne * NM + e * NM + j] = acc[c][j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < NM; ++j) smo[e * NM + j] = acc[0][j];
      }
    }
    if (MODE == MODE_GRAD) {
      // d psi / d r = J [ sum_u C_u sum_j inv_u[j][e] dmo[e][cols_u[j]] + grad_e(ln J) Sigma ]
      // (slater_jastrow.py:346-447): inverses, CI weights and the electron loop are generated
      const double J = (SPEC_USE_JEE || SPEC_USE_JEN) ? exp_clamped(P, et, tks) : 1.0;
      spec_grad<MODE>(mw, smo, sB, jv, J, a.pdf, a.out0 + w * ne3);
      continue;
    }
    // ---- determinants, traces, CI sum (generated, literal occupations)
    double det[NUN], tr[NUN];
    spec_dets<(MODE == MODE_ELOC)>(smo, sB, det, tr);
    double sig, ksig;
    spec_ci<MODE>(mw, det, tr, sig, ksig);
    const double J = (SPEC_USE_JEE || SPEC_USE_JEN) ? exp_clamped(P, et, tks) : 1.0;
    const double psi = J * sig;
    if (MODE == MODE_PSI) {
      a.out0[w] = psi;
    } else if (MODE == MODE_ELOC) {
      const double ekin = ksig / sig;
      const double el = ekin + tven + tvee + P.vnn;
      a.out0[w] = el;
      if (isfinite(el)) { st_s += el; st_s2 = fma(el, el, st_s2); ++st_nf; } else ++st_nb;
      if (a.out1) a.out1[w] = psi;
      if (a.out2) a.out2[w] = ekin;
    } else {
      double fxn = psi * psi;
      if (fxn == 0.0) fxn = a.eps;
      const double fx = a.out0[w];
      double df = fxn / fx;
      if (df > 1.0) df = 1.0;
      const double tau = a.tau ? a.tau[w] : philox_uniform(a.seed, a.offset, (uint64_t)w, 1u);
      const bool acc_ = (df - tau) >= 0.0;
      if (acc_) {
        a.out0[w] = fxn;   // fxn is never 0 here
#pragma unroll
        for (int i = 0; i < ne3; ++i) a.pos_rw[w * ne3 + i] = spos[i];
      }
      if (a.accept) a.accept[w] = acc_ ? 1 : 0;
      if (a.naccept) {
        // lanes leave the walker loop at different times
        const unsigned mask = __activemask();
        const int cnt = __reduce_add_sync(mask, acc_ ? 1 : 0);
        if ((int)(threadIdx.x & 31) == __ffs(mask) - 1 && cnt) atomicAdd(a.naccept, (unsigned long long)cnt);
      }
    }
  }
  // ---- fused statistics: fixed-order reduction of the CTA's walkers (warp butterflies, then the
  // warps in order) -> one partial per CTA; qmcb_local_energy_stats finishes with stats_stage2
  if (MODE == MODE_ELOC && a.stats_part) {
    double q[4] = {st_s, st_s2, (double)st_nf, (double)st_nb};
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) q[k] += __shfl_xor_sync(0xffffffffu, q[k], o);
    __syncthreads();                     // every thread is done with its slice: reuse it
    double *red = smem + QMCB_ETAB;
    if ((threadIdx.x & 31) == 0)
#pragma unroll
      for (int k = 0; k < 4; ++k) red[(threadIdx.x >> 5) * 4 + k] = q[k];
    __syncthreads();
    if (threadIdx.x < 4) {
      double t = 0.0;
      for (int wq = 0; wq < (int)(blockDim.x >> 5); ++wq) t += red[wq * 4 + threadIdx.x];
      a.stats_part[blockIdx.x * 4 + threadIdx.x] = t;
    }
    if (a.stats_ticket && a.stats_out) {
      // second stage without a second launch: the CTA that arrives last adds the per-CTA partials
      // in index order (one warp per quantity, lane-strided sums, fixed butterfly: the arithmetic
      // of stats_stage2, bitwise the same result whichever CTA is last).  atomicInc wraps the
      // counter back to zero for the next launch.
      __shared__ int last;
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) last = atomicInc(a.stats_ticket, gridDim.x - 1) == gridDim.x - 1;
      __syncthreads();
      if (last) {
        __threadfence();
        const int lane = threadIdx.x & 31;
        for (int q = threadIdx.x >> 5; q < 4; q += (int)(blockDim.x >> 5)) {
          double s = 0.0;
          for (int i = lane; i < (int)gridDim.x; i += 32) s += __ldcg(a.stats_part + i * 4 + q);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          if (lane == 0) a.stats_out[q] = s;
        }
      }
    }
  }
}

extern "C" __global__ void __launch_bounds__(SPEC_THREADS, SPEC_MINB)
    spec_psi(const __grid_constant__ SpecParams P, const FusedArgs a) { spec_body<MODE_PSI>(P, a); }
extern "C" __global__ void __launch_bounds__(SPEC_THREADS, SPEC_MINB_ELOC)
    spec_eloc(const __grid_constant__ SpecParams P, const FusedArgs a) { spec_body<MODE_ELOC>(P, a); }
extern "C" __global__ void __launch_bounds__(SPEC_THREADS, SPEC_MINB)
    spec_grad_psi(const __grid_constant__ SpecParams P, const FusedArgs a) { spec_body<MODE_GRAD>(P, a); }
extern "C" __global__ void __launch_bounds__(SPEC_THREADS, SPEC_MINB)
    spec_mh(const __grid_constant__ SpecParams P, const FusedArgs a) { spec_body<MODE_MH>(P, a); }
