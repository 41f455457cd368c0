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
// spec_aos<NCH>, spec_dets<WB>, spec_ci<WB> (see spec.cu: emit_source).
#pragma once

// ---- the prelude has defined: SPEC_NE SPEC_NUP SPEC_NDOWN SPEC_NATOM SPEC_NMUP SPEC_NUU SPEC_NUD
//      SPEC_USE_JEE SPEC_USE_JEN SPEC_GRAM_FMA SPEC_NV SPEC_OFF_ATOM SPEC_OFF_MOW SPEC_OFF_CI
//      SPEC_THREADS SPEC_MINB

struct SpecParams {
  double expc[8];
  double jee_w, jen_w, vnn;
  const double *etab_g;          // [64] 2^(j/64) (global; staged into shared memory)
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

// ---- building blocks the generated program calls (literal indices everywhere)
template <int NCH, bool FIRST>
__device__ __forceinline__ void spec_prim(const SpecParams &P, const double *et, double a, double c, double r2,
                                          double &S0, double &S1, double &S2) {
  const double ce = c * exp_neg(P, et, -a * r2);
  if (FIRST) S0 = ce; else S0 += ce;
  if (NCH > 1) {
    const double t = a * ce;
    const double u = t * fma(4.0 * a, r2, -6.0);
    if (FIRST) { S1 = -2.0 * t; S2 = u; } else { S1 = fma(-2.0, t, S1); S2 += u; }
  }
}

template <int NCH, int AO>
__device__ __forceinline__ void spec_emit(const SpecParams &P, const double (&v)[NCH], double (&acc)[NCH][SPEC_NMUP]) {
#pragma unroll
  for (int j = 0; j < SPEC_NMUP; ++j) {
    const double wj = P.v[SPEC_OFF_MOW + AO * SPEC_NMUP + j];
#pragma unroll
    for (int c = 0; c < NCH; ++c) acc[c][j] = fma(v[c], wj, acc[c][j]);
  }
}

template <int NCH, int AO>
__device__ __forceinline__ void spec_s(const SpecParams &P, double sc, double x, double y, double z, double S0,
                                       double S1, double S2, double (&acc)[NCH][SPEC_NMUP]) {
  double v[NCH];
  v[0] = S0 * sc;
  if (NCH > 1) {
    const double t = S1 * sc;
    v[1] = t * x; v[2] = t * y; v[3] = t * z;
    v[4] = S2 * sc;
  }
  spec_emit<NCH, AO>(P, v, acc);
}

template <int NCH, int AO>
__device__ __forceinline__ void spec_p(const SpecParams &P, double sc, double x, double y, double z, double S0,
                                       double S1, double S2, double (&acc)[NCH][SPEC_NMUP]) {
  double v[NCH];
  const double R = S0 * sc;
  if (NCH > 1) {
    const double t = S1 * sc, lf = fma(2.0, S1, S2) * sc;
    const double tx = t * x, ty = t * y, tz = t * z;
    v[0] = R * x; v[1] = fma(tx, x, R); v[2] = tx * y; v[3] = tx * z; v[4] = lf * x;
    spec_emit<NCH, AO>(P, v, acc);
    v[0] = R * y; v[1] = ty * x; v[2] = fma(ty, y, R); v[3] = ty * z; v[4] = lf * y;
    spec_emit<NCH, AO + 1>(P, v, acc);
    v[0] = R * z; v[1] = tz * x; v[2] = tz * y; v[3] = fma(tz, z, R); v[4] = lf * z;
    spec_emit<NCH, AO + 2>(P, v, acc);
  } else {
    v[0] = R * x; spec_emit<NCH, AO>(P, v, acc);
    v[0] = R * y; spec_emit<NCH, AO + 1>(P, v, acc);
    v[0] = R * z; spec_emit<NCH, AO + 2>(P, v, acc);
  }
}

template <int NCH, int AO, int KK>
__device__ __forceinline__ void spec_g(const SpecParams &P, double sc, double x, double y, double z, double S0,
                                       double S1, double S2, double (&acc)[NCH][SPEC_NMUP]) {
  double v[NCH];
  generic_component<NCH>(KK, sc, x, y, z, S0, S1, S2, v);   // literal powers: folds to a few products
  spec_emit<NCH, AO>(P, v, acc);
}

// ---- generated: spec_aos<NCH>, spec_dets<WB>, spec_ci<WB>
SPEC_GENERATED_CODE

template <int MODE>
__device__ __forceinline__ void spec_body(const SpecParams &P, const FusedArgs &a) {
  constexpr int NCH = MODE == MODE_ELOC ? 5 : 1;
  constexpr int Ne = SPEC_NE, ne3 = 3 * SPEC_NE, NM = SPEC_NMUP, NUN = SPEC_NUU + SPEC_NUD;
  constexpr int SLICE = (3 * Ne + (NCH > 1 ? 4 * Ne : 0) + (NCH > 1 ? 2 : 1) * Ne * NM) | 1;
  extern __shared__ __align__(16) double smem[];
  double *et = smem;
  for (int i = threadIdx.x; i < 64; i += blockDim.x) et[i] = P.etab_g[i];
  double *spos = smem + 64 + (size_t)threadIdx.x * SLICE;
  double *jv = spos + ne3;
  double *smo = jv + (NCH > 1 ? 4 * Ne : 0);
  double *sB = smo + Ne * NM;
  const SpecTab T{P};
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < a.W; w += stride) {
    // ---- coordinates (+ proposal)
    if (MODE == MODE_MH && !a.disp && a.proba_normal) {
      // one Philox call yields the two normals of a GLOBAL element pair (2p, 2p+1): the draw of an
      // element does not depend on the tiling (same stream as the generic kernel)
      const int64_t g0 = w * ne3, g1 = g0 + ne3;
      int me = a.move_elec;
      if (me == -2) {
        if (a.elec_index) me = a.elec_index[w];
        else me = (int)(philox_u32(a.seed, a.offset, (uint64_t)w, 2u) % (unsigned)Ne);
      }
      for (int64_t p = g0 >> 1; 2 * p < g1; ++p) {
        double z[2];
        philox_normal2(a.seed, a.offset, (uint64_t)p, z[0], z[1]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int64_t g = 2 * p + h;
          if (g < g0 || g >= g1) continue;
          const int i = (int)(g - g0), e = i / 3;
          double v = a.pos[g];
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
        double v = a.pos[w * ne3 + i];
        if (MODE == MODE_MH && (me < 0 || me == i / 3)) {
          double d;
          if (a.disp) d = a.disp[w * ne3 + i];
          else d = a.scale * (2.0 * philox_uniform(a.seed, a.offset, (uint64_t)(w * ne3 + i), 0u) - 1.0);
          v += d;
        }
        spos[i] = v;
      }
    }
    // ---- Jastrow gradient / Laplacian terms and potentials, every pair once
    double tks, tven, tvee;
    walker_terms<(NCH > 1), (MODE == MODE_ELOC)>(P, T, spos, jv, Ne, tks, tven, tvee);
    // ---- AO -> MO rows, one electron at a time (rolled: instruction-cache footprint)
#pragma unroll 1
    for (int e = 0; e < Ne; ++e) {
      double acc[NCH][NM];
#pragma unroll
      for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int j = 0; j < NM; ++j) acc[c][j] = 0.0;
      spec_aos<NCH>(P, et, spos[3 * e], spos[3 * e + 1], spos[3 * e + 2], acc);
      if (MODE == MODE_ELOC) {
        const double gx = jv[e], gy = jv[Ne + e], gz = jv[2 * Ne + e], lp = jv[3 * Ne + e];
#pragma unroll
        for (int j = 0; j < NM; ++j) {
          double b = acc[4][j];
          if (SPEC_USE_JEE || SPEC_USE_JEN)
            b += 2.0 * (gx * acc[1][j] + gy * acc[2][j] + gz * acc[3][j]) + lp * acc[0][j];
          smo[e * NM + j] = acc[0][j];
          sB[e * NM + j] = -0.5 * b;
        }
      } else {
#pragma unroll
        for (int j = 0; j < NM; ++j) smo[e * NM + j] = acc[0][j];
      }
    }
    // ---- determinants, traces, CI sum (generated, literal occupations)
    double det[NUN], tr[NUN];
    spec_dets<(MODE == MODE_ELOC)>(smo, sB, det, tr);
    double sig, ksig;
    spec_ci<(MODE == MODE_ELOC)>(P, det, tr, sig, ksig);
    const double J = (SPEC_USE_JEE || SPEC_USE_JEN) ? exp_clamped(P, et, tks) : 1.0;
    const double psi = J * sig;
    if (MODE == MODE_PSI) {
      a.out0[w] = psi;
    } else if (MODE == MODE_ELOC) {
      const double ekin = ksig / sig;
      a.out0[w] = ekin + tven + tvee + P.vnn;
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
}

extern "C" __global__ void __launch_bounds__(SPEC_THREADS, SPEC_MINB)
    spec_psi(const __grid_constant__ SpecParams P, const FusedArgs a) { spec_body<MODE_PSI>(P, a); }
extern "C" __global__ void __launch_bounds__(SPEC_THREADS, SPEC_MINB)
    spec_eloc(const __grid_constant__ SpecParams P, const FusedArgs a) { spec_body<MODE_ELOC>(P, a); }
extern "C" __global__ void __launch_bounds__(SPEC_THREADS, SPEC_MINB)
    spec_mh(const __grid_constant__ SpecParams P, const FusedArgs a) { spec_body<MODE_MH>(P, a); }
