// psi.backward(weight): sum over walkers of weight_w * d psi_w / d theta for
//   theta = MO weights, CI coefficients, basis exponents / contraction coefficients, Jastrow weights.
// Replaces the autograd backward of Solver.evaluate_grad_manual (solver/solver.py:414-429);
// formulas in SURVEY.md appendix A.6.
//
// One CTA owns TW walkers (rows = TW * nelec electron rows) per tile:
//   B0  coordinates -> smem
//   B1  thread (walker, electron): Jastrow exponent and its derivative w.r.t. the Pade weights
//   B2  thread (row): shell program, values only; stores the AO row, the harmonic factor of every
//       AO and, per grouped primitive, R_q and dR_q/d alpha
//   B2b MO rows from the AO rows;  B3 spin determinants + inverses;  B4 psi per walker
//   B5  G[row][m] = weight J sum_u C_u inv_u[j(m)][e];  U[row][a] = Y_a * sum_m G[row][m] W[a][m];
//       CI / Jastrow-weight sums by warp-striped, fixed-order reductions
//   B6  contractions over the rows on the FP64 tensor cores (mma.sync m8n8k4, DMMA):
//         dW[a][m]    += sum_row AO[row][a] G[row][m]
//         dC[q][a]    += sum_row R_q[row]   U[row][a]      (-> d/d bas_coeffs)
//         dE[q][a]    += sum_row dR_q[row]  U[row][a]      (-> d/d bas_exp)
//       every warp owns fixed 8x8 output tiles whose accumulators stay in registers for the
//       whole kernel; only tiles that contain a (primitive, AO) pair of the same shell are formed.
// The per-CTA partials are combined in index order by a second kernel: results are bitwise
// reproducible for a given W.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <set>

#include "device.cuh"
#include "spec.h"

#define BWD_MAXT 16   // output tiles per warp
#define BWD_WARP_LU_MIN 7   // spin blocks at least this large: one warp per block (see fused_impl.cuh)

struct BwdArgs {
  const double *pos, *weight;
  int64_t W;
  double *partial;
  const int *tiles;   // [ntile][2]
  int want_ao;
  int multi;          // some AO is a sum of several monomial components: AO rows accumulate (plan.cu)
  int tw, rows, lda, ldg, ldx, ppad, ntile_mo, ntile_ao, nslot, lu_conc;
};

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// fixed-order warp reduction (butterfly); every lane returns the total
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Jastrow exponent of one electron (pairs j>e and all nuclei) and d/dw of it
// een: d ln J / d(c, a, a', b, b') of every Boys-Handy term, pairs j>e: een[(5*m + k) * stride]
__device__ __forceinline__ void jastrow_value_dw(const DevSys &S, const Tab &T, const double *sp, int e,
                                                 double &ks, double &dkee, double &dken, double *een,
                                                 int stride) {
  const double xi = sp[3 * e], yi = sp[3 * e + 1], zi = sp[3 * e + 2];
  const double ni = __dadd_rn(__dadd_rn(__dmul_rn(xi, xi), __dmul_rn(yi, yi)), __dmul_rn(zi, zi));
  ks = 0.0; dkee = 0.0; dken = 0.0;
  if (S.use_jee) {
    const double w = S.jee_w;
    const bool up_i = e < S.nup;
    for (int j = e + 1; j < S.nelec; ++j) {
      const double xj = sp[3 * j], yj = sp[3 * j + 1], zj = sp[3 * j + 2];
      const double nj = __dadd_rn(__dadd_rn(__dmul_rn(xj, xj), __dmul_rn(yj, yj)), __dmul_rn(zj, zj));
      double dot;
      if (S.gram_fma) dot = __fma_rn(zi, zj, __fma_rn(yi, yj, __dmul_rn(xi, xj)));
      else dot = __dadd_rn(__dadd_rn(__dmul_rn(xi, xj), __dmul_rn(yi, yj)), __dmul_rn(zi, zj));
      const double r = sqrt(__dsub_rn(__dadd_rn(ni, nj), __dmul_rn(2.0, dot)));
      const double w0 = (up_i == (j < S.nup)) ? 0.25 : 0.5;
      const double den = 1.0 / (1.0 + w * r);
      ks += w0 * r * den;
      dkee -= w0 * r * r * den * den;      // d/dw [w0 r/(1+w r)]
    }
  }
  if (S.use_jen) {
    const double wn = S.jen_w;
    for (int A = 0; A < S.natom; ++A) {
      const double xa = T.atoms()[4 * A], ya = T.atoms()[4 * A + 1], za = T.atoms()[4 * A + 2];
      const double na = __dadd_rn(__dadd_rn(__dmul_rn(xa, xa), __dmul_rn(ya, ya)), __dmul_rn(za, za));
      const double dot = __fma_rn(zi, za, __fma_rn(yi, ya, __dmul_rn(xi, xa)));
      const double r = sqrt(__dsub_rn(__dadd_rn(ni, na), __dmul_rn(2.0, dot)));
      const double den = 1.0 / (1.0 + wn * r);
      ks += r * den;
      dken -= r * r * den * den;
    }
  }
  const int nt = S.een_nterm;
  if (nt > 0) {
    for (int k = 0; k < 5 * nt; ++k) een[k * stride] = 0.0;
    for (int j = e + 1; j < S.nelec; ++j) {
      const double xj = sp[3 * j], yj = sp[3 * j + 1], zj = sp[3 * j + 2];
      const double nj = gram_norm(xj, yj, zj);
      const double rej = sqrt(gram_d2_ee(S, xi, yi, zi, ni, xj, yj, zj, nj));
      for (int A = 0; A < S.natom; ++A) {
        const double xa = T.atoms()[4 * A], ya = T.atoms()[4 * A + 1], za = T.atoms()[4 * A + 2];
        const double na = gram_norm(xa, ya, za);
        const double rE = sqrt(gram_d2_en(xi, yi, zi, ni, xa, ya, za, na));
        const double rJ = sqrt(gram_d2_en(xj, yj, zj, nj, xa, ya, za, na));
        for (int m = 0; m < nt; ++m) {
          const double a = S.een_a[m], b = S.een_b[m], a2 = S.een_a2[m], b2 = S.een_b2[m], c = S.een_c[m];
          const double dE = 1.0 / (1.0 + b * rE), dJ = 1.0 / (1.0 + b * rJ), dG = 1.0 / (1.0 + b2 * rej);
          const double uE = rE * dE, uJ = rJ * dJ, uG = rej * dG;
          const double FE = a * uE, FJ = a * uJ, G = a2 * uG;
          const double P = FE * FJ * G;
          ks += c * P;
          double *q = een + 5 * m * stride;
          q[0] += P;                                        // d/dc
          q[stride] += c * (uE * FJ + FE * uJ) * G;         // d/da
          q[2 * stride] += c * FE * FJ * uG;                // d/da'
          q[3 * stride] -= c * P * (uE + uJ);               // d/db   (dF/db = -F r/(1+b r))
          q[4 * stride] -= c * P * uG;                      // d/db'
        }
      }
    }
  }
}

// shell program, values only; fills one row of AO, Y (into U) and X = [R_q | dR_q/dalpha]
// MOR > 0: the MO values of the (<= MOR) used columns are accumulated in registers on the fly and
// written to mo_row; MOR = 0: the caller forms them from the AO row afterwards.
template <int MOR>
__device__ __forceinline__ void backward_row(const DevSys &S, const Tab &T, double ex, double ey, double ez,
                                             double *ao, double *u, double *xr, int ppad, bool want_ao,
                                             double *mo_row, bool multi = false) {
  const double2 *rec = T.stream();
  const double *W = T.mow();
  const int nmup = S.nmup;
  double macc[MOR > 0 ? MOR : 1];
#pragma unroll
  for (int m = 0; m < (MOR > 0 ? MOR : 1); ++m) macc[m] = 0.0;
#define BWD_EMIT(idx, val)                                             \
  do {                                                                 \
    const double v_ = (val);                                           \
    if (multi) ao[idx] += v_; else ao[idx] = v_;                       \
    if (MOR > 0) {                                                     \
      _Pragma("unroll") for (int m = 0; m < MOR; ++m)                  \
        if (m < nmup) macc[m] = fma(v_, W[(idx)*nmup + m], macc[m]);   \
    }                                                                  \
  } while (0)
  const bool with_n = (S.radial_type == QMCB_GTO || S.radial_type == QMCB_STO);
  const bool gauss = (S.radial_type == QMCB_GTO || S.radial_type == QMCB_GTO_PURE);
  if (multi)
    for (int k = 0; k < S.nao; ++k) ao[k] = 0.0;
  int q = 0;
  for (int A = 0; A < S.natom; ++A) {
    const double x = ex - T.atoms()[4 * A], y = ey - T.atoms()[4 * A + 1], z = ez - T.atoms()[4 * A + 2];
    const double r2 = x * x + y * y + z * z;
    const double r = gauss && !with_n ? 0.0 : sqrt(r2);
    const int ns = T.ash()[A + 1] - T.ash()[A];
    for (int s = 0; s < ns; ++s) {
      const double hdr = rec->x;
      ++rec;
      const int nprim = __double2loint(hdr), ngrp = __double2hiint(hdr);
      double S0 = 0.0;
      for (int i = 0; i < nprim; ++i, ++q) {
        const double a = rec->x, c = rec->y;
        ++rec;
        double rn = 1.0;
        if (with_n) { rn = ipow(r, (int)rec->x); ++rec; }
        const double R = rn * exp_neg(S, T.etab(), gauss ? -a * r2 : -a * r);
        S0 += c * R;
        if (want_ao) {
          xr[q] = R;
          xr[ppad + q] = -(gauss ? r2 : r) * R;       // d R / d alpha (norm not differentiated)
        }
      }
      for (int g = 0; g < ngrp; ++g, ++rec) {
        const double2 gr = *rec;
        const int kk = __double2loint(gr.x), a0 = __double2hiint(gr.x);
        const double sc = gr.y;
        if (kk == 0) {
          BWD_EMIT(a0, S0 * sc);
          if (want_ao) u[a0] = 1.0;
        } else if (kk == (1 << 24)) {
          const double R = S0 * sc;
          BWD_EMIT(a0, R * x); BWD_EMIT(a0 + 1, R * y); BWD_EMIT(a0 + 2, R * z);
          if (want_ao) { u[a0] = x; u[a0 + 1] = y; u[a0 + 2] = z; }
        } else {
          const double Y = ipow(x, kk & 255) * ipow(y, (kk >> 8) & 255) * ipow(z, (kk >> 16) & 255);
          BWD_EMIT(a0, S0 * sc * Y);
          if (want_ao) u[a0] = Y;
        }
      }
    }
  }
  if (MOR > 0) {
#pragma unroll
    for (int m = 0; m < MOR; ++m)
      if (m < nmup) mo_row[m] = macc[m];
  }
#undef BWD_EMIT
}

__global__ void __launch_bounds__(512, 1) backward_kernel(const DevSys S, const BwdArgs a) {
  extern __shared__ __align__(16) double smem[];
  Tab T;
  double *ws = stage_tables(S, smem, T);
  const int Ne = S.nelec, ne3 = 3 * Ne, nmup = S.nmup, nun = S.nuu + S.nud;
  const int TW = a.tw, rows = a.rows, lda = a.lda, ldg = a.ldg, ldx = a.ldx;
  const int ntile = a.ntile_mo + a.ntile_ao;
  const int nmax = S.nup > S.ndown ? S.nup : S.ndown;
  const int inv_per = nmax <= 3 ? nmax * nmax : 2 * nmax * nmax;
  const int conc = a.lu_conc;
  // work area
  double *spos = ws;                          // [TW][3Ne]
  const int njv = 3 + 5 * S.een_nterm;
  double *jv = spos + TW * ne3;               // [njv][TW*Ne]  ks, dkee, dken, een derivatives
  double *sao = jv + njv * TW * Ne;           // [rows][lda]
  const bool wao = a.want_ao != 0;            // U and X exist only when basis gradients are wanted
  double *su = sao + rows * lda;              // [rows][lda]
  double *sg = su + (wao ? rows * lda : 0);   // [rows][ldg]
  double *sx = sg + rows * ldg;               // [rows][ldx]
  double *smo = sx + (wao ? rows * ldx : 0);  // [rows][nmup]
  double *sdet = smo + rows * nmup;           // [TW][nun]
  double *wj = sdet + TW * nun;               // [TW][4]  weight*J, Sigma, weight*psi
  const int nsc = S.nconf + 2 + 5 * S.een_nterm;   // scalar-type outputs: CI, Jastrow weights, e-e-n
  double *cacc = wj + TW * 4;                 // [nsc]
  double *scr = cacc + ((nsc + 1) & ~1);      // inverses [inv_per][conc]
  int *stiles = reinterpret_cast<int *>(scr + (size_t)inv_per * conc);
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nwarp = nthr >> 5;
  for (int i = tid; i < 2 * ntile; i += nthr) stiles[i] = a.tiles[i];
  for (int i = tid; i < rows * lda; i += nthr) { sao[i] = 0.0; if (wao) su[i] = 0.0; }
  for (int i = tid; i < rows * ldg; i += nthr) sg[i] = 0.0;
  for (int i = tid; wao && i < rows * ldx; i += nthr) sx[i] = 0.0;
  for (int i = tid; i < nsc; i += nthr) cacc[i] = 0.0;
  double c0[BWD_MAXT], c1[BWD_MAXT];
#pragma unroll
  for (int t = 0; t < BWD_MAXT; ++t) { c0[t] = 0.0; c1[t] = 0.0; }
  __syncthreads();

  const int64_t ntiles_w = (a.W + TW - 1) / TW;
  for (int64_t tile = blockIdx.x; tile < ntiles_w; tile += gridDim.x) {
    const int64_t w0 = tile * TW;
    const int tw = (int)((a.W - w0) < TW ? (a.W - w0) : TW);
    const int nrow = tw * Ne;
    for (int i = tid; i < tw * ne3; i += nthr) spos[i] = a.pos[w0 * ne3 + i];
    __syncthreads();
    // ---- B1 + B2: per electron row
    for (int it = tid; it < nrow; it += nthr) {
      const int wl = it / Ne, e = it - wl * Ne;
      const double *sp = spos + wl * ne3;
      double ks, dkee, dken;
      jastrow_value_dw(S, T, sp, e, ks, dkee, dken, jv + 3 * TW * Ne + it, TW * Ne);
      jv[it] = ks; jv[TW * Ne + it] = dkee; jv[2 * TW * Ne + it] = dken;
      if (nmup <= 2)
        backward_row<2>(S, T, sp[3 * e], sp[3 * e + 1], sp[3 * e + 2], sao + it * lda, su + it * lda,
                        sx + it * ldx, a.ppad, a.want_ao != 0, smo + it * nmup, a.multi != 0);
      else if (nmup <= 4)
        backward_row<4>(S, T, sp[3 * e], sp[3 * e + 1], sp[3 * e + 2], sao + it * lda, su + it * lda,
                        sx + it * ldx, a.ppad, a.want_ao != 0, smo + it * nmup, a.multi != 0);
      else if (nmup <= 8)
        backward_row<8>(S, T, sp[3 * e], sp[3 * e + 1], sp[3 * e + 2], sao + it * lda, su + it * lda,
                        sx + it * ldx, a.ppad, a.want_ao != 0, smo + it * nmup, a.multi != 0);
      else
        backward_row<0>(S, T, sp[3 * e], sp[3 * e + 1], sp[3 * e + 2], sao + it * lda, su + it * lda,
                        sx + it * ldx, a.ppad, a.want_ao != 0, nullptr, a.multi != 0);
    }
    // rows of a ragged last tile must not contribute
    for (int i = tid + nrow * ldg; i < rows * ldg; i += nthr) sg[i] = 0.0;
    for (int i = tid + nrow * lda; wao && i < rows * lda; i += nthr) su[i] = 0.0;
    __syncthreads();
    // ---- B2b: MO values of the used columns (only when they were not formed on the fly)
    for (int i = tid; nmup > 8 && i < nrow * nmup; i += nthr) {
      const int row = i / nmup, m = i - row * nmup;
      const double *ar = sao + row * lda;
      double acc = 0.0;
      for (int k = 0; k < S.nao; ++k) acc = fma(ar[k], T.mow()[k * nmup + m], acc);
      smo[i] = acc;
    }
    __syncthreads();
    // ---- B3: determinants and inverses per (walker, unique occupation): one thread (closed forms,
    // blocks <= 3x3, interleaved scratch) or one warp (warp_gauss_jordan, contiguous scratch)
    if (nmax >= BWD_WARP_LU_MIN) {
      for (int it = warp; it < tw * nun; it += nwarp) {
        const int wl = it / nun, u = it - wl * nun;
        const bool up = u < S.nuu;
        const int n = up ? S.nup : S.ndown;
        const int *cols = up ? T.ucu() + u * S.nup : T.ucd() + (u - S.nuu) * S.ndown;
        const double *A = smo + ((size_t)wl * Ne + (up ? 0 : S.nup)) * nmup;
        double *m = scr + (size_t)it * inv_per;
        const int ldw = 2 * n;
        double det = 1.0;
        if (n > 0 && n <= 16) {
          // register-resident: lane j owns column j of [A | I]; the inverse goes back to the scratch
          double col[16];
          const int jc = lane < n ? cols[lane] : 0;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            double v = 0.0;
            if (i < n) {
              if (lane < n) v = A[i * nmup + jc];
              else if (lane < ldw) v = i == lane - n ? 1.0 : 0.0;
            }
            col[i] = v;
          }
          det = warp_gauss_jordan_reg<16>(n, col, lane);
          if (lane >= n && lane < ldw) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (i < n) m[i * ldw + lane] = col[i];
          }
          __syncwarp();
        } else if (n > 0) {
          for (int idx = lane; idx < n * n; idx += 32) {
            const int i = idx / n, j = idx - i * n;
            m[i * ldw + j] = A[i * nmup + cols[j]];
            m[i * ldw + n + j] = i == j ? 1.0 : 0.0;
          }
          __syncwarp();
          det = warp_gauss_jordan(n, n, m, lane);
        }
        if (lane == 0) sdet[it] = det;
      }
    } else {
      for (int it = tid; it < tw * nun; it += nthr) {
        const int wl = it / nun, u = it - wl * nun;
        const bool up = u < S.nuu;
        const int n = up ? S.nup : S.ndown;
        const int *cols = up ? T.ucu() + u * S.nup : T.ucd() + (u - S.nuu) * S.ndown;
        const double *A = smo + ((size_t)wl * Ne + (up ? 0 : S.nup)) * nmup;
        double det = 1.0;
        if (n > 0 && n <= 3) det = inverse_small(n, A, nmup, cols, scr + it, conc);
        else if (n == 4) det = inverse_reg<4>(A, nmup, cols, scr + it, 2 * n, conc);   // register LU + unit-vector
        else if (n == 5) det = inverse_reg<5>(A, nmup, cols, scr + it, 2 * n, conc);   // solves (device.cuh): CAS
        else if (n == 6) det = inverse_reg<6>(A, nmup, cols, scr + it, 2 * n, conc);   // blocks, same scratch layout
        else if (n > 3) {
          double *m = scr + it;
          const int ldw = 2 * n;
          for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
              m[(i * ldw + j) * conc] = A[i * nmup + cols[j]];
              m[(i * ldw + n + j) * conc] = i == j ? 1.0 : 0.0;
            }
          det = gauss_jordan(n, n, m, conc);
        }
        sdet[it] = det;
      }
    }
    __syncthreads();
    // ---- B4: per walker
    for (int wl = tid; wl < tw; wl += nthr) {
      const double *dd = sdet + wl * nun;
      double sig = 0.0;
      for (int c = 0; c < S.nconf; ++c) sig += T.ci()[c] * dd[T.ciu()[c]] * dd[S.nuu + T.cid()[c]];
      double ks = 0.0;
      for (int e = 0; e < Ne; ++e) ks += jv[wl * Ne + e];
      const double J = (S.use_jee || S.use_jen || S.een_nterm > 0) ? exp_clamped(S, T.etab(), ks) : 1.0;
      const double wgt = a.weight[w0 + wl];
      wj[wl * 4] = wgt * J; wj[wl * 4 + 1] = sig; wj[wl * 4 + 2] = wgt * J * sig;
    }
    __syncthreads();
    // ---- B5a: G rows, U rows
    for (int it = tid; it < nrow; it += nthr) {
      const int wl = it / Ne, e = it - wl * Ne;
      const bool up = e < S.nup;
      const int n = up ? S.nup : S.ndown, el = up ? e : e - S.nup;
      const double *dd = sdet + wl * nun;
      double *g = sg + it * ldg;
      for (int m = 0; m < nmup; ++m) g[m] = 0.0;
      const double wJ = wj[wl * 4];
      const int nu = up ? S.nuu : S.nud;
      const bool contiguous = nmax >= BWD_WARP_LU_MIN;
      const int ild = (contiguous || n > 3) ? 2 * n : n, ioff = (contiguous || n > 3) ? n : 0, es = contiguous ? 1 : conc;
      for (int u = 0; u < nu; ++u) {
        double cu = 0.0;
        for (int c = 0; c < S.nconf; ++c) {
          if ((up ? T.ciu()[c] : T.cid()[c]) != u) continue;
          cu += T.ci()[c] * dd[up ? S.nuu + T.cid()[c] : T.ciu()[c]];
        }
        cu *= dd[up ? u : S.nuu + u] * wJ;
        if (cu == 0.0) continue;
        const int item = wl * nun + (up ? u : S.nuu + u);
        const double *inv = contiguous ? scr + (size_t)item * inv_per : scr + item;
        const int *cols = up ? T.ucu() + u * S.nup : T.ucd() + u * S.ndown;
        for (int j = 0; j < n; ++j) g[cols[j]] += cu * inv[(j * ild + ioff + el) * es];
      }
      if (a.want_ao) {
        double *ur = su + it * lda;
        for (int k = 0; k < S.nao; ++k) {
          double gao = 0.0;
          for (int m = 0; m < nmup; ++m) gao = fma(g[m], T.mow()[k * nmup + m], gao);
          ur[k] *= gao;
        }
      }
    }
    // ---- B5b: CI and Jastrow-weight sums (warp-striped, fixed order)
    for (int c = warp; c < nsc; c += nwarp) {
      double v = 0.0;
      if (c < S.nconf) {
        const int iu = T.ciu()[c], id = S.nuu + T.cid()[c];
        for (int wl = lane; wl < tw; wl += 32) v += wj[wl * 4] * sdet[wl * nun + iu] * sdet[wl * nun + id];
      } else {
        const double *dk = jv + (c - S.nconf + 1) * TW * Ne;
        for (int i = lane; i < nrow; i += 32) v += wj[(i / Ne) * 4 + 2] * dk[i];
      }
      v = warp_sum(v);
      if (lane == 0) cacc[c] += v;
    }
    __syncthreads();
    // ---- B6: contractions over the rows (DMMA); tile t belongs to warp t % nwarp
    {
      const int am = lane >> 2, kk = lane & 3;
      int slot = 0;
      for (int t = warp; t < ntile; t += nwarp, ++slot) {
        if (t >= a.ntile_mo && !a.want_ao) break;
        const bool mo = t < a.ntile_mo;
        const double *Am = mo ? sao : sx;
        const double *Bm = mo ? sg : su;
        const int la = mo ? lda : ldx, lb = mo ? ldg : lda;
        const int ao_ = stiles[2 * t] * 8 + am, bo_ = stiles[2 * t + 1] * 8 + am;
        double x0 = 0.0, x1 = 0.0;
        for (int r0 = 0; r0 < rows; r0 += 4) {
          const double av = Am[(r0 + kk) * la + ao_];
          const double bv = Bm[(r0 + kk) * lb + bo_];
          dmma884(x0, x1, av, bv);
        }
#pragma unroll
        for (int s2 = 0; s2 < BWD_MAXT; ++s2)
          if (s2 == slot) { c0[s2] += x0; c1[s2] += x1; }
      }
    }
    __syncthreads();
  }
  // ---- per-CTA partial: tiles [ntile][64], then CI, then the two Jastrow weights
  double *out = a.partial + (size_t)blockIdx.x * a.nslot;
  {
    int slot = 0;
    for (int t = warp; t < ntile; t += nwarp, ++slot) {
      double v0 = 0.0, v1 = 0.0;
#pragma unroll
      for (int s2 = 0; s2 < BWD_MAXT; ++s2)
        if (s2 == slot) { v0 = c0[s2]; v1 = c1[s2]; }
      const int row = lane >> 2, col = 2 * (lane & 3);
      out[t * 64 + row * 8 + col] = v0;
      out[t * 64 + row * 8 + col + 1] = v1;
    }
  }
  for (int i = tid; i < nsc; i += nthr) out[ntile * 64 + i] = cacc[i];
}

// Second stage: sum the CTA partials in index order and scatter to the caller's layouts.
__global__ void backward_reduce(const DevSys S, const double *partial, int ngrid, int nslot, const int *tiles,
                                int ntile_mo, int ntile_ao, int ppad, int nmo_full, double *g_mo, double *g_ci,
                                double *g_exp, double *g_coef, double *g_jee, double *g_jen, double *g_een) {
  const int ntile = ntile_mo + ntile_ao;
  const double *db = S.dblob;
  const int *ib = S.iblob;
  const int total = nslot;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    double v = 0.0;
    for (int g = 0; g < ngrid; ++g) v += partial[(size_t)g * nslot + i];
    if (i < ntile * 64) {
      const int t = i >> 6, r = (i >> 3) & 7, c = i & 7;
      const int rb = tiles[2 * t] * 8 + r, cb = tiles[2 * t + 1] * 8 + c;
      if (t < ntile_mo) {
        // row = AO index, col = position in the used-MO list
        if (g_mo && rb < S.nao && cb < S.nmu) g_mo[(size_t)rb * nmo_full + ib[S.o_used + cb]] = v;
      } else {
        // row = grouped primitive (second half: d/d alpha), col = AO index
        const bool dalpha = rb >= ppad;
        const int q = dalpha ? rb - ppad : rb;
        if (q < S.nprim && cb < S.nao) {
          // shell of q and component of AO cb inside that shell
          int s = 0;
          while (ib[S.o_spo + s + 1] <= q) ++s;
          const int k0 = ib[S.o_sco + s], k1 = ib[S.o_sco + s + 1];
          for (int k = k0; k < k1; ++k) {
            if (ib[S.o_cao + k] != cb) continue;
            const int ncomp = k1 - k0, qi = q - ib[S.o_spo + s];
            const int flat = ib[S.o_pflat + ib[S.o_pfo + s] + qi * ncomp + (k - k0)];
            if (dalpha) {
              if (g_exp) g_exp[flat] = v * db[S.o_coef + q] * db[S.o_cscale + k];
            } else if (g_coef) {
              g_coef[flat] = v * db[S.o_fnorm + flat];
            }
          }
        }
      }
    } else {
      const int c = i - ntile * 64;
      if (c < S.nconf) { if (g_ci) g_ci[c] = v; }
      else if (c == S.nconf) { if (g_jee) g_jee[0] = v; }
      else if (c == S.nconf + 1) { if (g_jen) g_jen[0] = v; }
      else if (g_een) {
        // slot (5*m + k), k = c, a, a', b, b'  ->  [num(2,nt) | denom(2,nt) | fc(nt)]
        const int idx = c - S.nconf - 2, m = idx / 5, k = idx - 5 * m, nt = S.een_nterm;
        const int dst = k == 0 ? 4 * nt + m : (k == 1 ? m : (k == 2 ? nt + m : (k == 3 ? 2 * nt + m : 3 * nt + m)));
        g_een[dst] = v;
      }
    }
  }
}

// Second stage of the structure-specialised backward (spec_kernel.cuh: spec_bwd_body): per-CTA partials
// [ngrid][nacc], nacc = nao * nmu + nconf + 2, summed in index order and scattered to the caller's layouts.
__global__ void bwd_spec_reduce(const DevSys S, const double *partial, int ngrid, int nacc, int nmo_full, double *g_mo,
                                double *g_ci, double *g_jee, double *g_jen, double *g_coef, double *g_exp) {
  // one warp per accumulator: lane-strided sums over the CTAs, then a fixed butterfly (deterministic)
  const int *ib = S.iblob;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
  for (int i = warp; i < nacc; i += nwarp) {
    double v = 0.0;
    for (int g = lane; g < ngrid; g += 32) v += partial[(size_t)g * nacc + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane != 0) continue;
    const int nw = S.nao * S.nmu;
    if (i < nw) {
      const int a = i / S.nmu, j = i - a * S.nmu;
      if (g_mo) g_mo[(size_t)a * nmo_full + ib[S.o_used + j]] = v;
    } else {
      const int c = i - nw;
      if (c < S.nconf) { if (g_ci) g_ci[c] = v; }
      else if (c == S.nconf) { if (g_jee) g_jee[0] = v; }
      else if (c == S.nconf + 1) { if (g_jen) g_jen[0] = v; }
      else {
        // spec_backward_all: [nbas] d/d bas_coeffs, then [nbas] d/d bas_exp, by flat primitive
        const int f = c - S.nconf - 2;
        if (f < S.nbas) { if (g_coef) g_coef[f] = v; }
        else if (g_exp) g_exp[f - S.nbas] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------
static int pad_ld(int n) {
  int p = ((n + 7) / 8) * 8;
  if (p % 16 == 0) p += 8;     // rows 8 doubles apart modulo 16: conflict-free fragment loads
  return p;
}

int qmcb_choose_backward(qmcb_plan *p) {
  const DevSys &S = p->sys;
  auto &b = p->bwd;
  b = qmcb_plan::BwdCfg{};
  b.lda = pad_ld(S.nao);
  b.ldg = pad_ld(S.nmup);
  b.ppad = ((S.nprim + 7) / 8) * 8;
  b.ldx = pad_ld(2 * b.ppad);
  // output tiles: MO gradient first
  std::vector<int> &tl = p->bwd_tiles;
  tl.clear();
  const int nab = (S.nao + 7) / 8, nmb = (S.nmu + 7) / 8;
  for (int i = 0; i < nab; ++i)
    for (int j = 0; j < nmb; ++j) { tl.push_back(i); tl.push_back(j); }
  b.ntile_mo = nab * nmb;
  // (primitive block, AO block) pairs that contain a same-shell pair, for both halves of X
  std::set<std::pair<int, int>> need;
  const int *spo = p->hi.data() + S.o_spo, *sco = p->hi.data() + S.o_sco, *cao = p->hi.data() + S.o_cao;
  for (int s = 0; s < S.nshell; ++s)
    for (int q = spo[s]; q < spo[s + 1]; ++q)
      for (int k = sco[s]; k < sco[s + 1]; ++k) {
        need.insert({q / 8, cao[k] / 8});
        need.insert({(b.ppad + q) / 8, cao[k] / 8});
      }
  for (auto &pr : need) { tl.push_back(pr.first); tl.push_back(pr.second); }
  b.ntile_ao = (int)need.size();
  const int ntile = b.ntile_mo + b.ntile_ao;
  b.nslot = ntile * 64 + S.nconf + 2 + 5 * S.een_nterm;
  const int nun = S.nuu + S.nud;
  const int nmax = S.nup > S.ndown ? S.nup : S.ndown;
  const int inv_per = nmax <= 3 ? nmax * nmax : 2 * nmax * nmax;
  const int budget = p->smem_optin - 1024;
  // two tilings: with the basis-parameter part (U, X resident) and without it (larger tiles)
  for (int wao = 1; wao >= 0; --wao) {
    auto &c = wao ? p->bwd : p->bwd0;
    if (!wao) c = p->bwd;
    c.tw = 0;
    for (int tw = 128; tw >= 1; tw = tw > 8 ? tw / 2 : tw - 1) {
      const int rows = ((tw * S.nelec + 3) / 4) * 4;
      int threads = ((rows + 31) / 32) * 32;
      const int ntile_act = b.ntile_mo + (wao ? b.ntile_ao : 0);
      const int need_warps = (ntile_act + BWD_MAXT - 1) / BWD_MAXT;
      if (threads < 32 * need_warps) threads = 32 * need_warps;
      { const char *em = getenv("QMCB_BWD_MINTHREADS"); const int mt = em ? atoi(em) : 256; if (threads < mt) threads = mt; }   // measured (ms, J+MO / all): H2O 1e5 walkers 128: 1.60 / 5.28, 256: 1.51 / 4.97, 384: 2.81 / 4.90; C4H6 2e4: 2.13 / 13.5, 256: 2.01 / 13.1
      if (threads > 512) continue;
      const int conc = tw * nun;
      size_t d = (size_t)table_doubles(S) + (size_t)tw * 3 * S.nelec +
                 (size_t)(3 + 5 * S.een_nterm) * tw * S.nelec +
                 (size_t)rows * ((wao ? 2 : 1) * b.lda + b.ldg + (wao ? b.ldx : 0) + S.nmup) + (size_t)tw * nun +
                 (size_t)tw * 4 + (size_t)((S.nconf + 2 + 5 * S.een_nterm + 1) & ~1) + (size_t)inv_per * conc +
                 (size_t)ntile + 2;
      const size_t sm = d * sizeof(double);
      // keep the tile small enough for two CTAs per SM when that is possible with >= 64 rows
      if ((int)sm <= budget && ((int)sm <= 100 * 1024 || rows <= 64)) {
        c.tw = tw; c.rows = rows; c.threads = threads; c.smem = (int)sm; c.lu_conc = conc;
        c.grid = 2 * p->sm_count;
        break;
      }
    }
  }
  // tw == 0: backward unavailable for this system; forward entry points still work
  return 0;
}

extern "C" int64_t qmcb_backward_workspace_bytes(const qmcb_plan *p, int64_t W) {
  if (!p) return 0;
  // bases with multi-monomial AOs (real spherical harmonics, l = 2): the basis-parameter gradients come from the
  // flat-primitive adjoint kernel (eloc_vjp.cu), which brings its own work areas
  const int64_t vjp = p->multi_component ? qmcb_local_energy_backward_workspace_bytes(p, W) : 0;
  // tile kernel: 2 CTAs per SM x nslot;  specialised kernel: up to 8 CTAs per SM x (nao nmu + nconf + 2)
  const int64_t spec = (int64_t)8 * p->sm_count * ((int64_t)p->sys.nao * p->sys.nmu + p->sys.nconf + 2 + 2 * (int64_t)p->sys.nbas);
  const int64_t tile = (p->bwd.tw == 0 && p->bwd0.tw == 0) ? 0 : (int64_t)2 * p->sm_count * p->bwd.nslot;
  const int64_t own = (spec > tile ? spec : tile) * (int64_t)sizeof(double);
  return own > vjp ? own : vjp;
}

extern "C" int qmcb_psi_backward(const qmcb_plan *p, const double *pos, const double *weight, int64_t W,
                                 double *g_mo, double *g_ci, double *g_bas_exp, double *g_bas_coeffs,
                                 double *g_jee_w, double *g_jen_w, double *g_een, void *workspace,
                                 void *stream) {
  if (!p || !p->d_dbl || !pos || !weight || !workspace || W <= 0) {
    qmcb_set_error("qmcb_psi_backward: bad arguments");
    return QMCB_EINVAL;
  }
  if ((g_bas_exp || g_bas_coeffs) && p->multi_component) {
    // An AO that is a sum of several monomials (real spherical harmonics of l = 2) does not fit the shell /
    // component staging of the kernels below; the adjoint kernel works on the flat primitive list and serves
    // any AO composition: basis-parameter gradients from there (psi weight only), the rest as usual.
    int rc = qmcb_local_energy_backward(p, pos, nullptr, weight, W, nullptr, nullptr, g_bas_exp, g_bas_coeffs, nullptr,
                                        nullptr, nullptr, workspace, stream);
    if (rc) return rc;
    if (!g_mo && !g_ci && !g_jee_w && !g_jen_w && !g_een) return 0;
    g_bas_exp = nullptr;
    g_bas_coeffs = nullptr;
  }
  const int want_ao = (g_bas_exp || g_bas_coeffs) ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  // Jastrow / MO / CI gradients of a one-walker-per-thread structure (BASELINE config 3): the
  // structure-specialised backward, register accumulators, no tile staging (QMCB_BWD_SPEC=0 disables)
  const char *env_spec = getenv("QMCB_BWD_SPEC");          // (read per call: tests switch between the two kernels)
  const bool spec_off = env_spec && atoi(env_spec) == 0;
  if (!g_een && !spec_off) {
    // (with basis-parameter gradients: spec_backward_all - a second generated walk over the primitives and one
    // register accumulator per flat primitive)
    cudaError_t e0;
    const size_t nmo_b = (size_t)p->sys.nao * p->sys.nmo * sizeof(double);
    if (g_mo && (e0 = cudaMemsetAsync(g_mo, 0, nmo_b, st)) != cudaSuccess) return qmcb_cuda_rc((int)e0, "backward.cu");
    FusedArgs fa{};
    fa.pos = pos; fa.W = W; fa.weight = weight; fa.bwd_part = (double *)workspace;
    int grid = 0;
    const int rc = qmcb_spec_launch(p, want_ao ? MODE_BWD_ALL : MODE_BWD, fa, stream, &grid);
    if (rc == 0) {
      const int nacc = p->sys.nao * p->sys.nmu + p->sys.nconf + 2 + (want_ao ? 2 * p->sys.nbas : 0);
      bwd_spec_reduce<<<(nacc + 3) / 4, 128, 0, st>>>(p->sys, (const double *)workspace, grid, nacc, p->sys.nmo, g_mo,
                                                      g_ci, g_jee_w, g_jen_w, g_bas_coeffs, g_bas_exp);
      return qmcb_cuda_rc((int)cudaGetLastError(), "backward.cu spec reduce");
    }
    if (rc != QMCB_SPEC_SKIP) return rc;
  }
  const auto &b = want_ao ? p->bwd : p->bwd0;
  if (b.tw == 0) {
    qmcb_set_error("qmcb_psi_backward: system does not fit the backward tiling");
    return QMCB_ESMEM;
  }
  BwdArgs a{};
  a.pos = pos; a.weight = weight; a.W = W; a.partial = (double *)workspace; a.tiles = p->d_bwd_tiles;
  a.want_ao = want_ao;
  a.multi = p->multi_component ? 1 : 0;
  a.tw = b.tw; a.rows = b.rows; a.lda = b.lda; a.ldg = b.ldg; a.ldx = b.ldx; a.ppad = b.ppad;
  a.ntile_mo = b.ntile_mo; a.ntile_ao = b.ntile_ao; a.nslot = b.nslot; a.lu_conc = b.lu_conc;
  cudaError_t e = cudaFuncSetAttribute(backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, b.smem);
  if (e != cudaSuccess) return qmcb_cuda_rc((int)e, "backward.cu");
  const int64_t ntile_w = (W + b.tw - 1) / b.tw;
  int grid = b.grid;
  if (grid > ntile_w) grid = (int)ntile_w;
  const size_t nmo_bytes = (size_t)p->sys.nao * p->sys.nmo * sizeof(double);
  if (g_mo && (e = cudaMemsetAsync(g_mo, 0, nmo_bytes, st)) != cudaSuccess) return qmcb_cuda_rc((int)e, "backward.cu");
  if (g_bas_exp && (e = cudaMemsetAsync(g_bas_exp, 0, p->sys.nbas * sizeof(double), st)) != cudaSuccess) return qmcb_cuda_rc((int)e, "backward.cu");
  if (g_bas_coeffs && (e = cudaMemsetAsync(g_bas_coeffs, 0, p->sys.nbas * sizeof(double), st)) != cudaSuccess)
    return qmcb_cuda_rc((int)e, "backward.cu");
  backward_kernel<<<grid, b.threads, b.smem, st>>>(p->sys, a);
  if ((e = cudaGetLastError()) != cudaSuccess) return qmcb_cuda_rc((int)e, "backward.cu");
  backward_reduce<<<(b.nslot + 127) / 128, 128, 0, st>>>(p->sys, a.partial, grid, b.nslot, p->d_bwd_tiles, b.ntile_mo,
                                                        b.ntile_ao, b.ppad, p->sys.nmo, g_mo, g_ci, g_bas_exp,
                                                        g_bas_coeffs, g_jee_w, g_jen_w, g_een);
  return qmcb_cuda_rc((int)cudaGetLastError(), "backward.cu launch");
}
