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
// Compiled with -DQMCB_SPEC after a generated prelude that defines SPEC_*, after spec_common.cuh (the
// parameter block and the building blocks of the generated shell walk), with the generated functions
// spec_aos<MODE>, spec_dets<WB>, spec_ci<MODE>, spec_grad<MODE> (see spec.cu: walk) spliced in below.
#pragma once

// ---- the prelude has defined: SPEC_NE SPEC_NUP SPEC_NDOWN SPEC_NATOM SPEC_NMUP SPEC_NUU SPEC_NUD
//      SPEC_USE_JEE SPEC_USE_JEN SPEC_GRAM_FMA SPEC_NV SPEC_OFF_ATOM SPEC_OFF_MOW SPEC_OFF_CI
//      SPEC_THREADS SPEC_MINB

// ---- generated: spec_aos<MODE>, spec_dets<WB>, spec_ci<MODE>, spec_grad<MODE>
SPEC_GENERATED_CODE

template <int MODE>
__device__ __forceinline__ void spec_body(const SpecParams &P, const FusedArgs &a) {
  constexpr int NCH = spec_nch<MODE>();
  constexpr int Ne = SPEC_NE, ne3 = 3 * SPEC_NE, NM = SPEC_NMUP, NUN = SPEC_NUU + SPEC_NUD;
  constexpr int NROW = MODE == MODE_ELOC ? 2 : (MODE == MODE_GRAD ? 4 : 1);   // mo | B_kin or mo | d mo/dx,dy,dz
  constexpr int SLICE = (3 * Ne + (spec_deriv<MODE>() ? 4 * Ne : 0) + NROW * Ne * NM) | 1;
  extern __shared__ __align__(16) double smem[];
  // exp table, replicated QMCB_ETAB_REP times (device.cuh: exp_core); this thread reads replica lane % REP
  double *et0 = smem;
  for (int i = threadIdx.x; i < QMCB_ETAB * QMCB_ETAB_REP; i += blockDim.x) et0[i] = P.etab_g[i / QMCB_ETAB_REP];
  const double *et = et0 + (threadIdx.x & (QMCB_ETAB_REP - 1));
  constexpr int NMW = SPEC_MOW_SMEM ? ((SPEC_NV - SPEC_OFF_MOW + 1) & ~1) : 0;   // MO weights + CI, even
  double *mw = smem + QMCB_ETAB * QMCB_ETAB_REP;
  for (int i = threadIdx.x; i < NMW; i += blockDim.x) mw[i] = i < SPEC_NV - SPEC_OFF_MOW ? P.v[SPEC_OFF_MOW + i] : 0.0;
  constexpr bool PF = SPEC_PREFETCH || (MODE == MODE_ELOC && SPEC_PREFETCH_ELOC);
  constexpr int SL = SLICE + (PF ? ne3 + (ne3 & 1) : 0);               // stays odd
  double *spos = smem + QMCB_ETAB * QMCB_ETAB_REP + NMW + (size_t)threadIdx.x * SL;
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
    double *red = smem + QMCB_ETAB * QMCB_ETAB_REP;
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

// ---------------------------------------------------------------------------------------
// psi.backward(weight) for the parameters BASELINE config 3 trains (solver/solver.py:414-429; formulas:
// SURVEY appendix A.6): MO weights of the occupied columns, CI coefficients, Pade Jastrow weights.
// One walker per thread like the forward kernels: the walker's AO rows stay in its shared-memory slice,
// the generated spec_bwd forms the inverses, the CI weights C_u and G[e][m] = w J sum_u C_u inv_u[j(m)][e]
// with literal indices and adds AO[e][a] G[e][m] into REGISTER accumulators dW[a][m] that live for the
// whole kernel; one fixed-order block reduction at the end -> one partial per CTA, summed in index order
// by bwd_spec_reduce (backward.cu): bitwise reproducible.  The tile kernel backward_kernel (DMMA) remains
// the path for basis-parameter gradients, three-body weights and larger structures.
// ---------------------------------------------------------------------------------------
#ifndef SPEC_NCONF
#define SPEC_NCONF 1
#endif
// MODE_BWD_ALL adds the basis parameters (bas_exp, bas_coeffs): after the dW accumulation the
// AO row of an electron is overwritten by Gao[e][a] = sum_m G[e][m] W[a][m], and a second generated walk over
// the primitives (spec_aos_grad) adds  R_q Y_k Gao[e][a_k]  and  dR_q/dalpha Y_k Gao[e][a_k]  into one
// register accumulator per flat primitive; norms / coefficients are applied once at the end (spec_grad_scale).
#ifndef SPEC_NBAS
#define SPEC_NBAS 1
#endif
template <int MODE>
__device__ __forceinline__ void spec_bwd_body(const SpecParams &P, const FusedArgs &a) {
  constexpr bool ALL = MODE == MODE_BWD_ALL;
  constexpr int Ne = SPEC_NE, ne3 = 3 * SPEC_NE, NM = SPEC_NMUP, NAO = SPEC_NAO, NC = SPEC_NCONF;
  constexpr int NB = ALL ? SPEC_NBAS : 0;
  constexpr int NACC = NAO * NM + NC + 2 + 2 * NB;
  // slice: pos | mo rows | AO rows | landing zone of the next walker's coordinates (cp.async prefetch)
  constexpr int SL = (ne3 + Ne * NM + Ne * NAO + ne3) | 1;
  extern __shared__ __align__(16) double smem[];
  double *et0 = smem;
  for (int i = threadIdx.x; i < QMCB_ETAB * QMCB_ETAB_REP; i += blockDim.x) et0[i] = P.etab_g[i / QMCB_ETAB_REP];
  const double *et = et0 + (threadIdx.x & (QMCB_ETAB_REP - 1));
  double *spos = smem + QMCB_ETAB * QMCB_ETAB_REP + (size_t)threadIdx.x * SL;
  double *smo = spos + ne3, *sao = smo + Ne * NM, *snext = sao + Ne * NAO;
  auto prefetch = [&](int64_t wn) {
    if (wn < a.W) {
      const unsigned dst = (unsigned)__cvta_generic_to_shared(snext);
      const double *src = a.pos + wn * ne3;
#pragma unroll
      for (int i = 0; i < ne3; ++i)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8 * i), "l"(src + i) : "memory");
    }
  };
  const SpecTab T{P};
  __syncthreads();
  double dW[NAO][NM], dci[NC], djee = 0.0, djen = 0.0;
  double aC[ALL ? SPEC_NBAS : 1], aE[ALL ? SPEC_NBAS : 1];
#pragma unroll
  for (int i = 0; i < (ALL ? SPEC_NBAS : 1); ++i) { aC[i] = 0.0; aE[i] = 0.0; }
#pragma unroll
  for (int i = 0; i < NAO; ++i)
#pragma unroll
    for (int j = 0; j < NM; ++j) dW[i][j] = 0.0;
#pragma unroll
  for (int c = 0; c < NC; ++c) dci[c] = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  prefetch((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < a.W; w += stride) {
    asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
    for (int i = 0; i < ne3; ++i) spos[i] = snext[i];
    const double wgt = a.weight[w];
    prefetch(w + stride);
    // Jastrow exponent and its derivative w.r.t. the Pade weights (d/dw [w0 r / (1 + w r)] = -w0 r^2 / (1 + w r)^2)
    double ks = 0.0, dkee = 0.0, dken = 0.0;
#pragma unroll
    for (int i = 0; i < Ne; ++i) {
      const double xi = spos[3 * i], yi = spos[3 * i + 1], zi = spos[3 * i + 2];
      const double ni = gram_norm(xi, yi, zi);
      if (SPEC_USE_JEE) {
#pragma unroll
        for (int j = i + 1; j < Ne; ++j) {
          const double xj = spos[3 * j], yj = spos[3 * j + 1], zj = spos[3 * j + 2];
          const double d2 = gram_d2_ee(P, xi, yi, zi, ni, xj, yj, zj, gram_norm(xj, yj, zj));
          const double r = d2 * fast_rsqrt(d2);
          const double w0 = ((i < SPEC_NUP) == (j < SPEC_NUP)) ? 0.25 : 0.5;
          const double den = fast_rcp(1.0 + P.jee_w * r);
          const double t = w0 * r * den;
          ks += t;
          dkee -= t * r * den;
        }
      }
      if (SPEC_USE_JEN) {
#pragma unroll
        for (int A = 0; A < SPEC_NATOM; ++A) {
          const double xa = T.atoms()[4 * A], ya = T.atoms()[4 * A + 1], za = T.atoms()[4 * A + 2];
          const double d2 = gram_d2_en(xi, yi, zi, ni, xa, ya, za, gram_norm(xa, ya, za));
          const double r = d2 > 0.0 ? d2 * fast_rsqrt(d2) : 0.0;
          const double den = fast_rcp(1.0 + P.jen_w * r);
          ks += r * den;
          dken -= r * r * den * den;
        }
      }
    }
    // AO rows (kept) and MO rows of the occupied columns
    const FoldJ fj{0.0, 0.0, 0.0, 0.0};
    double ven = 0.0;
    for (int e = 0; e < Ne; ++e) {
      double acc[1][NM];
#pragma unroll
      for (int j = 0; j < NM; ++j) acc[0][j] = 0.0;
#pragma unroll
      for (int k = 0; k < NAO; ++k) sao[e * NAO + k] = 0.0;
      spec_aos<MODE>(P, et, sao + e * NAO, spos[3 * e], spos[3 * e + 1], spos[3 * e + 2], fj, ven, acc);
#pragma unroll
      for (int j = 0; j < NM; ++j) smo[e * NM + j] = acc[0][j];
    }
    const double J = (SPEC_USE_JEE || SPEC_USE_JEN) ? exp_clamped(P, et, ks) : 1.0;
    const double wJ = wgt * J;
    double sig;
    spec_bwd<MODE>(smo, sao, wJ, dW, dci, sig);
    djee = fma(wJ * sig, dkee, djee);
    djen = fma(wJ * sig, dken, djen);
    if constexpr (ALL) {
      for (int e = 0; e < Ne; ++e)
        spec_aos_grad<MODE>(P, et, sao + e * NAO, spos[3 * e], spos[3 * e + 1], spos[3 * e + 2], aC, aE);
    }
  }
  if constexpr (ALL) spec_grad_scale<MODE>(aC, aE);
  // ---- fixed-order reduction over the CTA: warp butterflies, then the warps in order
  double *red = smem + QMCB_ETAB * QMCB_ETAB_REP;       // [warps][NACC], reuses the slices
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  auto put = [&](int idx, double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp * NACC + idx] = v;
  };
#pragma unroll
  for (int i = 0; i < NAO; ++i)
#pragma unroll
    for (int j = 0; j < NM; ++j) put(i * NM + j, dW[i][j]);
#pragma unroll
  for (int c = 0; c < NC; ++c) put(NAO * NM + c, dci[c]);
  put(NAO * NM + NC, djee);
  put(NAO * NM + NC + 1, djen);
  if constexpr (ALL) {
#pragma unroll
    for (int i = 0; i < SPEC_NBAS; ++i) { put(NAO * NM + NC + 2 + i, aC[i]); put(NAO * NM + NC + 2 + SPEC_NBAS + i, aE[i]); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NACC; i += blockDim.x) {
    double t = 0.0;
    for (int wq = 0; wq < nwarp; ++wq) t += red[wq * NACC + i];
    a.bwd_part[(size_t)blockIdx.x * NACC + i] = t;
  }
}

extern "C" __global__ void __launch_bounds__(SPEC_THREADS, SPEC_MINB_ELOC)
    spec_backward(const __grid_constant__ SpecParams P, const FusedArgs a) { spec_bwd_body<MODE_BWD>(P, a); }
extern "C" __global__ void __launch_bounds__(SPEC_THREADS, 2)
    spec_backward_all(const __grid_constant__ SpecParams P, const FusedArgs a) { spec_bwd_body<MODE_BWD_ALL>(P, a); }
extern "C" __global__ void __launch_bounds__(SPEC_THREADS, SPEC_MINB)
    spec_psi(const __grid_constant__ SpecParams P, const FusedArgs a) { spec_body<MODE_PSI>(P, a); }
extern "C" __global__ void __launch_bounds__(SPEC_THREADS, SPEC_MINB_ELOC)
    spec_eloc(const __grid_constant__ SpecParams P, const FusedArgs a) { spec_body<MODE_ELOC>(P, a); }
extern "C" __global__ void __launch_bounds__(SPEC_THREADS, SPEC_MINB)
    spec_grad_psi(const __grid_constant__ SpecParams P, const FusedArgs a) { spec_body<MODE_GRAD>(P, a); }
extern "C" __global__ void __launch_bounds__(SPEC_THREADS, SPEC_MINB)
    spec_mh(const __grid_constant__ SpecParams P, const FusedArgs a) { spec_body<MODE_MH>(P, a); }
