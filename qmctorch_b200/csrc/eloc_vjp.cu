// Adjoint of the local energy: sum over walkers of
//      wE_w * d E_L(R_w) / d theta   +   wP_w * d psi(R_w) / d theta
// for every wave-function parameter theta AND the atom coordinates.  This is what the reference obtains by
// back-propagating through WaveFunction.local_energy - Solver.evaluate_grad_auto (solver/solver.py:352-370:
// loss.backward() through E_L) and Solver.compute_forces (solver/solver.py:433-519: autograd.grad of E_L and of
// log psi^2 w.r.t. ao.atom_coords) - a second-order backward through the analytic AO derivatives, the
// AO -> MO products, torch.inverse / torch.det of every spin block, the trace trick and the Jastrow
// derivatives (wavefunction/slater_jastrow.py:312-344,449-482, pooling/slater_pooling.py:262-387).
//
// Here the adjoint is hand-derived for the dense middle of E_L and taken in FORWARD mode at the leaves:
//
//   leaves   AO channels (ao, d ao, lap ao)[e][a]        <- bas_exp, bas_coeffs, atom_coords
//            g_e = grad_e J / J, l_e = lap_e J / J, J    <- Pade weights (e-e, e-n)
//            V_en, V_nn                                   <- atom_coords
//   middle   K[e][a]  = lap ao + 2 g_e . d ao + l_e ao                   (folded kinetic channel)
//            MO = AO W,  B = -1/2 K W,   A_u = MO[rows_s, cols_u],  B_u likewise
//            D_u = det A_u,  t_u = Tr(A_u^-1 B_u),  S = sum_c c_c Du Dd,  T = sum_c c_c Du Dd (tu + td)
//            psi = J S,   E_L = T / S + V
//
//   reverse  S~ = wP J - wE T / S^2,  T~ = wE / S,  c~_c = S~ Du Dd + T~ Du Dd (tu + td),
//            D~_u = sum_c (S~ + T~ (tu + td)) c_c D_other,   t~_u = sum_c T~ c_c Du Dd,
//            A~_u = D~_u D_u A_u^-T - t~_u (A_u^-1 B_u A_u^-1)^T,   B~_u = t~_u A_u^-T,
//            scattered into M0[e][m] (adjoint of MO from the determinants) and Q[e][m] = -1/2 B~[e][m];
//            W~[a][m] = sum_e AO[e][a] M0[e][m] + K[e][a] Q[e][m];   QW = Q W^T,  MW = M0 W^T;
//            adjoints of the AO channels: (MW + l_e QW, 2 g_e QW, QW);  g~_e = 2 sum_a d ao QW,  l~_e = sum_a ao QW.
//
// The Jacobian of the leaves is block diagonal - a primitive's exponent only moves its own AO, an atom only its
// own shells, a Pade weight only its own term - so ONE dual-number evaluation per (electron, primitive) with
// the tangent directions (u_x, u_y, u_z, alpha), contracted at once with the channel adjoints, yields every
// basis-parameter and atom-coordinate derivative (third derivatives of the AOs included) without deriving them
// by hand, and one Dual<2> pass over the Pade terms yields the Jastrow-weight derivatives.
//
// Mapping: one WARP per walker (lanes stride over (electron, AO) / (electron, MO) / matrix entries), per-warp
// scratch in global memory (L1/L2 resident; sized for C4H6: 30 x 94 x 8 channels), phases separated by
// __syncwarp().  Parameter sums go to per-warp accumulators with a fixed owner lane per entry; walkers are
// assigned to warps by index and vjp_finish adds the per-warp partials in warp order: bitwise reproducible.
// General in the structure (any radial type, cartesian monomials, CAS expansions, n x n spin blocks with
// partial pivoting); the three-body Boys-Handy term enters through g_e, l_e, J (taken from the qmcb_jastrow
// kernel) and its OWN weights get no E_L derivative - the reference's graph drops the Laplacian's dependence
// on them too (jastrow_factor_electron_electron_nuclei.py:411-431: create_graph=False), so that gradient is
// not defined there either; the host API refuses it.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

#include "plan.h"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int kThreads = 128;        // 4 warps per CTA
constexpr int kMinBlocks = 3;        // CTAs per SM the kernel is compiled for: 168 registers (measured, first version: 2 -> 20.6, 3 -> 14.7 ms per 1e6 LiH walkers; final version: 3 -> 4.5, 4 (128 registers, spills) -> 5.5 ms)

// ---- dual numbers (forward-mode derivatives along N directions)
template <int N>
struct Dual {
  double v;
  double d[N];
};
template <int N> __device__ __forceinline__ Dual<N> mk(double v) {
  Dual<N> r; r.v = v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = 0.0;
  return r;
}
template <int N> __device__ __forceinline__ Dual<N> seed(double v, int dir) {
  Dual<N> r = mk<N>(v);
#pragma unroll
  for (int i = 0; i < N; ++i) if (i == dir) r.d[i] = 1.0;
  return r;
}
#define QMCB_DUAL_BIN(OP, VEXP, DEXP)                                                            \
  template <int N> __device__ __forceinline__ Dual<N> operator OP(const Dual<N> &a, const Dual<N> &b) { \
    Dual<N> r; r.v = VEXP;                                                                        \
    _Pragma("unroll") for (int i = 0; i < N; ++i) r.d[i] = DEXP;                                  \
    return r; }
QMCB_DUAL_BIN(+, a.v + b.v, a.d[i] + b.d[i])
QMCB_DUAL_BIN(-, a.v - b.v, a.d[i] - b.d[i])
QMCB_DUAL_BIN(*, a.v * b.v, a.d[i] * b.v + a.v * b.d[i])
#undef QMCB_DUAL_BIN
template <int N> __device__ __forceinline__ Dual<N> operator*(double s, const Dual<N> &a) {
  Dual<N> r; r.v = s * a.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = s * a.d[i];
  return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator*(const Dual<N> &a, double s) { return s * a; }
template <int N> __device__ __forceinline__ Dual<N> operator+(const Dual<N> &a, double s) { Dual<N> r = a; r.v += s; return r; }
template <int N> __device__ __forceinline__ Dual<N> operator+(double s, const Dual<N> &a) { return a + s; }
template <int N> __device__ __forceinline__ Dual<N> operator-(const Dual<N> &a, double s) { Dual<N> r = a; r.v -= s; return r; }
template <int N> __device__ __forceinline__ Dual<N> operator-(const Dual<N> &a) { return -1.0 * a; }
template <int N> __device__ __forceinline__ Dual<N> rcp(const Dual<N> &a) {
  Dual<N> r; r.v = 1.0 / a.v;
  const double m = -r.v * r.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = m * a.d[i];
  return r;
}
template <int N> __device__ __forceinline__ Dual<N> dsqrt(const Dual<N> &a) {
  Dual<N> r; r.v = sqrt(a.v);
  const double m = 0.5 / r.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = m * a.d[i];
  return r;
}
template <int N> __device__ __forceinline__ Dual<N> dexp(const Dual<N> &a) {
  Dual<N> r; r.v = exp(a.v);
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = r.v * a.d[i];
  return r;
}
__device__ __forceinline__ double rcp(double a) { return 1.0 / a; }
__device__ __forceinline__ double dsqrt(double a) { return sqrt(a); }
__device__ __forceinline__ double dexp(double a) { return exp(a); }

template <class T> __device__ __forceinline__ T ipow(const T &x, int k) {   // k >= 1
  T r = x;
  for (int i = 1; i < k; ++i) r = r * x;
  return r;
}
// x^k for any integer k given 1/x (k may be negative: r^(n-2) of the radial powers)
template <class T> __device__ __forceinline__ T rpow(const T &x, const T &xinv, int k, const T &one) {
  if (k == 0) return one;
  return k > 0 ? ipow(x, k) : ipow(xinv, -k);
}

// ---- one primitive R(r) x^kx y^ky z^kz without coefficient / norm: value, gradient, Laplacian
// (orbitals/radial_functions.py:6-406 and spherical_harmonics.py:102-199, the formulas of oracle/sj_oracle.py:
// ao_all).  T = double or a dual number; out[0] value, out[1..3] gradient, out[4] Laplacian.
template <class T>
__device__ __forceinline__ void primitive5(int rt, int kx, int ky, int kz, int n, const T &x, const T &y, const T &z,
                                           const T &al, const T &one, T (&out)[5]) {
  const T r2 = x * x + y * y + z * z;
  T R, R1, LR;     // R, (dR/dr)/r, lap R
  if (rt == QMCB_GTO_PURE) {
    R = dexp(-(al * r2));
    R1 = -2.0 * (al * R);
    LR = (al * R) * (4.0 * (al * r2) - 6.0);
  } else {
    const T r = dsqrt(r2);
    const T rinv = rcp(r);
    if (rt == QMCB_STO_PURE) {
      R = dexp(-(al * r));
      R1 = -(al * R) * rinv;
      LR = (al * R) * (al - 2.0 * rinv);
    } else {
      const bool gto = rt == QMCB_GTO;
      const T e = gto ? dexp(-(al * r2)) : dexp(-(al * r));
      const T rn = rpow(r, rinv, n, one);
      const T rnm2 = rpow(r, rinv, n - 2, one);
      const double nn = (double)n;
      R = rn * e;
      if (gto) {
        // R = r^n e^{-a r^2}:  R'/r = (n r^{n-2} - 2 a r^n) e,  lap R = (n(n+1) r^{n-2} - 2a(2n+3) r^n + 4 a^2 r^{n+2}) e
        R1 = (nn * rnm2 - 2.0 * (al * rn)) * e;
        LR = (nn * (nn + 1.0) * rnm2 - (2.0 * (2.0 * nn + 3.0)) * (al * rn) + 4.0 * ((al * al) * (rn * r2))) * e;
      } else {
        // R = r^n e^{-a r}:  R'/r = (n r^{n-2} - a r^{n-1}) e,  lap R = (n(n+1) r^{n-2} - 2a(n+1) r^{n-1} + a^2 r^n) e
        const T rnm1 = rn * rinv;
        R1 = (nn * rnm2 - al * rnm1) * e;
        LR = (nn * (nn + 1.0) * rnm2 - (2.0 * (nn + 1.0)) * (al * rnm1) + (al * al) * rn) * e;
      }
    }
  }
  // cartesian monomial Y, its gradient and Laplacian; s and p functions (most of every basis) without the
  // general power code
  const int Lk = kx + ky + kz;
  if (Lk == 0) {
    out[0] = R; out[1] = R1 * x; out[2] = R1 * y; out[3] = R1 * z; out[4] = LR;
    return;
  }
  if (Lk == 1) {
    const T &u = kx ? x : (ky ? y : z);
    const T R1u = R1 * u;
    out[0] = R * u;
    out[1] = R1u * x; out[2] = R1u * y; out[3] = R1u * z;
    if (kx) out[1] = out[1] + R; else if (ky) out[2] = out[2] + R; else out[3] = out[3] + R;
    out[4] = LR * u + 2.0 * R1u;
    return;
  }
  const T xk = kx ? ipow(x, kx) : one, yk = ky ? ipow(y, ky) : one, zk = kz ? ipow(z, kz) : one;
  const T Y = xk * yk * zk;
  T gx = one * 0.0, gy = gx, gz = gx, LY = gx;
  if (kx) { const T xm = kx > 1 ? ipow(x, kx - 1) : one; gx = (double)kx * (xm * (yk * zk)); }
  if (ky) { const T ym = ky > 1 ? ipow(y, ky - 1) : one; gy = (double)ky * (ym * (xk * zk)); }
  if (kz) { const T zm = kz > 1 ? ipow(z, kz - 1) : one; gz = (double)kz * (zm * (xk * yk)); }
  if (kx > 1) { const T xm = kx > 2 ? ipow(x, kx - 2) : one; LY = LY + (double)(kx * (kx - 1)) * (xm * (yk * zk)); }
  if (ky > 1) { const T ym = ky > 2 ? ipow(y, ky - 2) : one; LY = LY + (double)(ky * (ky - 1)) * (ym * (xk * zk)); }
  if (kz > 1) { const T zm = kz > 2 ? ipow(z, kz - 2) : one; LY = LY + (double)(kz * (kz - 1)) * (zm * (xk * yk)); }
  const double L = (double)(kx + ky + kz);
  const T R1Y = R1 * Y;
  out[0] = R * Y;
  out[1] = R1Y * x + R * gx;
  out[2] = R1Y * y + R * gy;
  out[3] = R1Y * z + R * gz;
  out[4] = LR * Y + (2.0 * L) * R1Y + R * LY;      // u . grad Y = L Y (Euler)
}

// ---- device view of the system for this kernel
struct VjpSys {
  int nelec, nup, ndown, natom, nbas, nao, nmo, nmu, nmup, nconf, nuu, nud, radial_type, use_jee, use_jen, has_j;
  double jee_w, jen_w;
  const double *atoms;                   // [natom][4] x y z Z
  const double *mow;                     // [nao][nmup] weights of the used MO columns
  const double *ci;                      // [nconf]
  const int *used, *ucu, *ucd, *ciu, *cid;
  const double *alpha, *cn, *norm;       // [nbas] exponent, norm * coeff, norm
  const int *patom, *pk, *pkr, *pao;     // [nbas] atom, kx | ky << 8 | kz << 16, radial power, AO
  const int *ao_start, *ao_prim;         // CSR: AO -> its flat primitives
  const int *ao_order;                   // AOs by decreasing contraction length (lockstep lanes get equal work)
};

// accumulator layout (doubles, per warp): W~ [nao][nmu] | ci [nconf] | exp [nbas] | coef [nbas] | primR [nbas][3] |
//                                         jee | jen | sum wE | V_en atoms [natom][3]
struct AccLayout {
  int o_w, o_ci, o_exp, o_coef, o_pr, o_jee, o_jen, o_sumE, o_ven, n;
  __host__ __device__ explicit AccLayout(const VjpSys &S) {
    o_w = 0; o_ci = o_w + S.nao * S.nmu; o_exp = o_ci + S.nconf; o_coef = o_exp + S.nbas; o_pr = o_coef + S.nbas;
    o_jee = o_pr + 3 * S.nbas; o_jen = o_jee + 1; o_sumE = o_jen + 1; o_ven = o_sumE + 1; n = o_ven + 3 * S.natom;
    n = (n + 1) & ~1;
  }
};
// scratch layout (doubles, per warp)
struct ScratchLayout {
  int o_ao, o_kc, o_mo, o_bk, o_m0, o_q, o_mw, o_qw, o_g, o_gb, o_det, o_inv, o_gj, o_col, n;
  __host__ __device__ explicit ScratchLayout(const VjpSys &S) {
    const int ea = S.nelec * S.nao, em = S.nelec * S.nmu, nun = S.nuu + S.nud;
    const int nmax = S.nup > S.ndown ? S.nup : S.ndown;
    o_ao = 0; o_kc = o_ao + 5 * ea; o_mo = o_kc + ea; o_bk = o_mo + em; o_m0 = o_bk + em; o_q = o_m0 + em;
    o_mw = o_q + em; o_qw = o_mw + ea; o_g = o_qw + ea;         // g: [4][Ne] (gx, gy, gz, l)
    o_gb = o_g + 4 * S.nelec;                                    // adjoints of g, l: [4][Ne]
    o_det = o_gb + 4 * S.nelec;                                  // D | t | D~ | t~ : [4][nun]
    o_inv = o_det + 4 * nun;                                     // per unique block: A^-1 [n][n] | A^-1 B A^-1 [n][n]
    o_gj = o_inv + 2 * (S.nuu * S.nup * S.nup + S.nud * S.ndown * S.ndown);
    o_col = o_gj + 3 * nmax * nmax;                              // Gauss-Jordan work [n][3n], pivot column [n]
    n = o_col + nmax;
    n = (n + 1) & ~1;
  }
};

// ---- lane groups.  G lanes (a power of two, 4 ... 32) own one walker, a warp works on 32 / G walkers at
// once; every loop below strides by G with the lane's position in its group and every reduction stays inside
// the group, so the groups of a warp run the same phases in lockstep and __syncwarp() separates them.
template <int G> __device__ __forceinline__ double group_sum(double v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
// sum over the groups of the warp of the value each group holds on the same group lane (fixed butterfly)
template <int G> __device__ __forceinline__ double across_groups(double v) {
#pragma unroll
  for (int o = G; o < 32; o <<= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
// Per-warp accumulators: for idx0 .. idx0 + n, entry i is owned by group lane i % G; the groups' values are
// added across the warp and group 0 adds the sum to acc (one owner lane per entry: no atomics, fixed order).
// `f(i)` returns this group's contribution to entry i.
template <int G, class F>
__device__ __forceinline__ void accumulate(double *acc, int idx0, int n, int sub, bool first_group, F f) {
  for (int i0 = 0; i0 < n; i0 += G) {
    const int i = i0 + sub;
    double v = i < n ? f(i) : 0.0;
    v = across_groups<G>(v);
    if (first_group && i < n) acc[idx0 + i] += v;
  }
}

// Group-cooperative Gauss-Jordan on M [n][3n] = [A | I | B] with partial pivoting (first largest |entry|: the
// pivot order does not depend on G) -> [I | A^-1 | A^-1 B]; returns det A.
template <int G>
__device__ double group_gauss_jordan3(double *M, double *col, int n, int sub) {
  const int ld = 3 * n;
  double det = 1.0;
  for (int k = 0; k < n; ++k) {
    double best = -1.0;
    int bi = k;
    for (int r = k + sub; r < n; r += G) {
      const double a = fabs(M[r * ld + k]);
      if (a > best) { best = a; bi = r; }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(FULL, best, o);
      const int oi = __shfl_xor_sync(FULL, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (bi != k) {
      for (int c = sub; c < ld; c += G) {
        const double t = M[k * ld + c];
        M[k * ld + c] = M[bi * ld + c];
        M[bi * ld + c] = t;
      }
      det = -det;
    }
    __syncwarp();
    const double p = M[k * ld + k];
    det *= p;
    const double pinv = 1.0 / p;
    for (int r = sub; r < n; r += G) col[r] = M[r * ld + k];
    __syncwarp();
    for (int c = sub; c < ld; c += G) M[k * ld + c] *= pinv;
    __syncwarp();
    for (int i = sub; i < n * ld; i += G) {
      const int r = i / ld, c = i - r * ld;
      if (r != k) M[i] = fma(-col[r], M[k * ld + c], M[i]);
    }
    __syncwarp();
  }
  return det;
}

// Small spin blocks (n <= 6: CAS expansions have many of them): ONE LANE per block, everything in registers -
// LU with partial pivoting (row swaps as selects: no dynamic register indexing), the inverse by one forward / back
// substitution per unit vector (written to Ai as it is formed), then X = A^-1 B row by row with its trace and
// Y = X A^-1.  The cooperative Gauss-Jordan spends ~10^4 cycles of dependent shared-memory round trips on a 5 x 5
// block; here the blocks of a walker are done side by side on different lanes.
template <int N>
__device__ __noinline__ void lane_block(const double *MO, const double *BK, int Nm, int r0, const int *cols, double *Ai,
                                        double *Yv, double &det_out, double &tr_out) {
  double a[N][N];
  int perm[N], cj[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    cj[j] = cols[j];
    perm[j] = j;
#pragma unroll
    for (int i = 0; i < N; ++i) a[i][j] = MO[(r0 + i) * Nm + cj[j]];
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
    const double ip = 1.0 / a[k][k];
#pragma unroll
    for (int i = k + 1; i < N; ++i) {
      const double l = a[i][k] * ip;
      a[i][k] = l;
#pragma unroll
      for (int j = k + 1; j < N; ++j) a[i][j] = fma(-l, a[k][j], a[i][j]);
    }
  }
  double inv_d[N];
#pragma unroll
  for (int i = 0; i < N; ++i) inv_d[i] = 1.0 / a[i][i];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double x[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {                     // forward: L y = P e_j
      double v = perm[i] == j ? 1.0 : 0.0;
#pragma unroll
      for (int q = 0; q < i; ++q) v = fma(-a[i][q], x[q], v);
      x[i] = v;
    }
#pragma unroll
    for (int i = N - 1; i >= 0; --i) {                // backward: U x = y
      double v = x[i];
#pragma unroll
      for (int q = i + 1; q < N; ++q) v = fma(-a[i][q], x[q], v);
      x[i] = v * inv_d[i];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) Ai[i * N + j] = x[i];
  }
  double tr = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double xr[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double v = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) v = fma(Ai[i * N + k], BK[(r0 + k) * Nm + cj[j]], v);
      xr[j] = v;
    }
    tr += xr[i];
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double v = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) v = fma(xr[k], Ai[k * N + j], v);
      Yv[i * N + j] = v;
    }
  }
  det_out = det;
  tr_out = tr;
}

// One flat primitive q against all electrons of the walker: sum_e (channel adjoints) . (channels of q and their
// tangents).  NDIR: tangent directions: 0 none, 1 (alpha), 4 (u_x, u_y, u_z, alpha).  Returns through sv
// (value part: the bas_coeffs derivative) and sd[NDIR].
template <int NDIR>
__device__ __forceinline__ void prim_leaf(const VjpSys &S, const ScratchLayout &SL, const double *x, const double *sc,
                                          int q, double &sv, double (&sd)[4]) {
  const int Ne = S.nelec, Na = S.nao;
  const double *g = sc + SL.o_g;
  const double *MW = sc + SL.o_mw, *QW = sc + SL.o_qw;
  const int A = S.patom[q], a = S.pao[q], pk = S.pk[q], n = S.pkr[q];
  const int kx = pk & 255, ky = (pk >> 8) & 255, kz = (pk >> 16) & 255;
  const double ax = S.atoms[4 * A], ay = S.atoms[4 * A + 1], az = S.atoms[4 * A + 2];
  sv = 0.0;
  sd[0] = sd[1] = sd[2] = sd[3] = 0.0;
  for (int e = 0; e < Ne; ++e) {
    const double ux = x[3 * e] - ax, uy = x[3 * e + 1] - ay, uz = x[3 * e + 2] - az;
    const double qw = QW[e * Na + a];
    const double c0 = fma(g[3 * Ne + e], qw, MW[e * Na + a]);
    const double c1 = 2.0 * g[e] * qw, c2 = 2.0 * g[Ne + e] * qw, c3 = 2.0 * g[2 * Ne + e] * qw;
    if constexpr (NDIR == 0) {
      double o[5];
      primitive5<double>(S.radial_type, kx, ky, kz, n, ux, uy, uz, S.alpha[q], 1.0, o);
      sv += c0 * o[0] + c1 * o[1] + c2 * o[2] + c3 * o[3] + qw * o[4];
    } else {
      typedef Dual<NDIR> T;
      T o[5];
      const T X = NDIR == 4 ? seed<NDIR>(ux, 0) : mk<NDIR>(ux), Y = NDIR == 4 ? seed<NDIR>(uy, 1) : mk<NDIR>(uy),
              Z = NDIR == 4 ? seed<NDIR>(uz, 2) : mk<NDIR>(uz);
      primitive5<T>(S.radial_type, kx, ky, kz, n, X, Y, Z, seed<NDIR>(S.alpha[q], NDIR - 1), mk<NDIR>(1.0), o);
      sv += c0 * o[0].v + c1 * o[1].v + c2 * o[2].v + c3 * o[3].v + qw * o[4].v;
#pragma unroll
      for (int j = 0; j < NDIR; ++j)
        sd[j] += c0 * o[0].d[j] + c1 * o[1].d[j] + c2 * o[2].d[j] + c3 * o[3].d[j] + qw * o[4].d[j];
    }
  }
}

template <int G, int NDIR>
__device__ __forceinline__ void prim_leaf_pass(const VjpSys &S, const ScratchLayout &SL, const AccLayout &AL,
                                               const double *x, const double *sc, double *acc, int sub, bool first_group,
                                               bool want_coef) {
  for (int q0 = 0; q0 < S.nbas; q0 += G) {
    const int q = q0 + sub;
    const bool on = q < S.nbas;
    double sv = 0.0, sd[4] = {0.0, 0.0, 0.0, 0.0};
    if (on) prim_leaf<NDIR>(S, SL, x, sc, q, sv, sd);
    const double cn = on ? S.cn[q] : 0.0;
    if (want_coef) {
      const double v = across_groups<G>(on ? S.norm[q] * sv : 0.0);
      if (first_group && on) acc[AL.o_coef + q] += v;
    }
    if constexpr (NDIR > 0) {
      const double v = across_groups<G>(cn * sd[NDIR - 1]);
      if (first_group && on) acc[AL.o_exp + q] += v;
    }
    if constexpr (NDIR == 4) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double v = across_groups<G>(cn * sd[k]);
        if (first_group && on) acc[AL.o_pr + 3 * q + k] -= v;      // d/dR_A = - d/du
      }
    }
  }
}

struct VjpArgs {
  const double *pos;       // [W][3 Ne] (the whole ensemble)
  const double *wE, *wP;   // [W] or nullptr
  int64_t W, w0, w1;       // this launch handles walkers [w0, w1)
  const double *J, *dJ, *d2J;   // Jastrow operator output of the chunk (index w - w0), or nullptr
  double *scratch, *acc;   // global: scratch per group (large systems), accumulators per warp
  int smem;                // 1: the work areas and accumulators of a warp live in shared memory
  int stop;                // profiling only (QMCB_VJP_STOP=k): leave the walker after phase k; 0 = full pass
  int want_mo, want_ci, want_exp, want_coef, want_jee, want_jen, want_atom;
};

template <int G>
__global__ void __launch_bounds__(kThreads, kMinBlocks) eloc_vjp_kernel(const VjpSys S_in, const VjpArgs a) {
  extern __shared__ __align__(16) double vjp_smem[];
  constexpr int GPW = 32 / G;                       // walkers (groups) per warp
  VjpSys S = S_in;
  if (!a.smem) {
    // large systems: the work areas stream through L1 and evict the small read-only tables every primitive
    // evaluation touches (7 loads per primitive went to L2); shared memory is unused in this mode, so the
    // tables live there: atoms | MO weights | alpha, norm*coeff, norm | int tables
    const int nb = S.nbas, nfi = 5 * nb + 2 * S.nao + 1;
    double *t_atoms = vjp_smem, *t_mow = t_atoms + 4 * S.natom, *t_fd = t_mow + S.nao * S.nmup;
    int *t_fi = reinterpret_cast<int *>(t_fd + 3 * nb);
    for (int i = threadIdx.x; i < 4 * S.natom; i += blockDim.x) t_atoms[i] = S_in.atoms[i];
    for (int i = threadIdx.x; i < S.nao * S.nmup; i += blockDim.x) t_mow[i] = S_in.mow[i];
    for (int i = threadIdx.x; i < 3 * nb; i += blockDim.x) t_fd[i] = S_in.alpha[i];
    for (int i = threadIdx.x; i < nfi; i += blockDim.x) t_fi[i] = S_in.patom[i];
    __syncthreads();
    S.atoms = t_atoms; S.mow = t_mow;
    S.alpha = t_fd; S.cn = t_fd + nb; S.norm = t_fd + 2 * nb;
    S.patom = t_fi; S.pk = t_fi + nb; S.pkr = t_fi + 2 * nb; S.pao = t_fi + 3 * nb;
    S.ao_start = t_fi + 4 * nb; S.ao_prim = t_fi + 4 * nb + S.nao + 1; S.ao_order = t_fi + 5 * nb + S.nao + 1;
  }
  const ScratchLayout SL(S);
  const AccLayout AL(S);
  const int lane = threadIdx.x & 31, sub = lane % G, grp = lane / G;
  const bool first_group = grp == 0;
  const int wic = threadIdx.x >> 5;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + wic;
  const int64_t nwarp = (int64_t)gridDim.x * (blockDim.x >> 5);
  // work area of the group and accumulators of the warp: shared memory when the system is small enough (the
  // phases of one walker are a chain of ~30 dependent steps: an L2 round trip per step otherwise dominates),
  // global scratch for large ones
  double *wbase = vjp_smem + (size_t)wic * (GPW * SL.n + AL.n);
  double *sc = a.smem ? wbase + grp * SL.n : a.scratch + (warp * GPW + grp) * SL.n;
  double *gacc = a.acc + warp * AL.n;
  double *acc = a.smem ? wbase + GPW * SL.n : gacc;
  if (a.smem) {
    for (int i = lane; i < AL.n; i += 32) acc[i] = 0.0;
    __syncwarp();
  }
  const int Ne = S.nelec, Na = S.nao, Nm = S.nmu, ld = S.nmup, nun = S.nuu + S.nud;
  double *AO = sc + SL.o_ao, *KC = sc + SL.o_kc, *MO = sc + SL.o_mo, *BK = sc + SL.o_bk, *M0 = sc + SL.o_m0,
         *Q = sc + SL.o_q, *MW = sc + SL.o_mw, *QW = sc + SL.o_qw, *g = sc + SL.o_g, *gb = sc + SL.o_gb,
         *Dt = sc + SL.o_det, *INV = sc + SL.o_inv, *GJ = sc + SL.o_gj, *col = sc + SL.o_col;
  // walkers of this launch go to (warp, group) slots by their index: a fixed assignment, so the sums are
  // bitwise reproducible.  The loop bound is warp uniform; a group past the end re-evaluates the last walker
  // with zero weights (every adjoint is proportional to them).
  for (int64_t wb = a.w0 + warp * GPW; wb < a.w1; wb += nwarp * GPW) {
    const bool live = wb + grp < a.w1;
    const int64_t w = live ? wb + grp : a.w1 - 1;
    const double *x = a.pos + w * 3 * Ne;
    const double wE = (live && a.wE) ? a.wE[w] : 0.0, wP = (live && a.wP) ? a.wP[w] : 0.0;
    // ---- Jastrow leaves of this walker: J, g = grad J / J, l = lap J / J
    double J = 1.0;
    if (S.has_j && a.J) {
      // (three-body term present: the product of all factors from the jastrow_kernel launch of this chunk)
      const int64_t wl = w - a.w0;
      J = a.J[wl];
      const double Jinv = 1.0 / J;
      for (int i = sub; i < 4 * Ne; i += G)
        g[i] = (i < 3 * Ne ? a.dJ[wl * 3 * Ne + i] : a.d2J[wl * Ne + (i - 3 * Ne)]) * Jinv;
    } else if (S.has_j) {
      // Pade factors only: ln J = K, grad J / J = grad K, lap J / J = lap K + |grad K|^2 evaluated here (the same
      // kernels the derivative pass below differentiates) - no operator launch, no J / dJ / d2J round trip
      double ksum = 0.0;
      for (int e = sub; e < Ne; e += G) {
        const double xe = x[3 * e], ye = x[3 * e + 1], ze = x[3 * e + 2];
        double kx = 0.0, ky = 0.0, kz = 0.0, kl = 0.0;
        if (S.use_jee) {
          for (int j = 0; j < Ne; ++j) {
            if (j == e) continue;
            const double dx = xe - x[3 * j], dy = ye - x[3 * j + 1], dz = ze - x[3 * j + 2];
            const double r = sqrt(dx * dx + dy * dy + dz * dz), rinv = 1.0 / r;
            const double w0 = ((e < S.nup) == (j < S.nup)) ? 0.25 : 0.5;
            const double den = 1.0 / (S.jee_w * r + 1.0);
            const double k1 = w0 * den * den, kr = k1 * rinv;
            ksum = fma(0.5 * w0 * r, den, ksum);
            kx = fma(kr, dx, kx); ky = fma(kr, dy, ky); kz = fma(kr, dz, kz);
            kl += -2.0 * S.jee_w * k1 * den + 2.0 * kr;
          }
        }
        if (S.use_jen) {
          for (int A = 0; A < S.natom; ++A) {
            const double dx = xe - S.atoms[4 * A], dy = ye - S.atoms[4 * A + 1], dz = ze - S.atoms[4 * A + 2];
            const double r = sqrt(dx * dx + dy * dy + dz * dz), rinv = 1.0 / r;
            const double den = 1.0 / (S.jen_w * r + 1.0);
            const double k1 = den * den, kr = k1 * rinv;
            ksum = fma(r, den, ksum);
            kx = fma(kr, dx, kx); ky = fma(kr, dy, ky); kz = fma(kr, dz, kz);
            kl += -2.0 * S.jen_w * k1 * den + 2.0 * kr;
          }
        }
        g[e] = kx; g[Ne + e] = ky; g[2 * Ne + e] = kz;
        g[3 * Ne + e] = kl + kx * kx + ky * ky + kz * kz;
      }
      J = exp(group_sum<G>(ksum));
    } else {
      for (int i = sub; i < 4 * Ne; i += G) g[i] = 0.0;
    }
    __syncwarp();
    if (a.stop == 7) continue;
    // ---- AO channels and the folded kinetic channel
    for (int it0 = sub; it0 < Ne * Na; it0 += G) {
      // items AO-major, AOs by decreasing contraction length: the lanes of a round (and the groups of the warp,
      // which run in lockstep) evaluate AOs of (nearly) the same length
      const int ko = it0 / Ne, e = it0 - ko * Ne, ao = S.ao_order[ko], it = e * Na + ao;
      double s[5] = {0, 0, 0, 0, 0};
      for (int k = S.ao_start[ao]; k < S.ao_start[ao + 1]; ++k) {
        const int q = S.ao_prim[k], A = S.patom[q], pk = S.pk[q];
        double o[5];
        primitive5<double>(S.radial_type, pk & 255, (pk >> 8) & 255, (pk >> 16) & 255, S.pkr[q],
                           x[3 * e] - S.atoms[4 * A], x[3 * e + 1] - S.atoms[4 * A + 1],
                           x[3 * e + 2] - S.atoms[4 * A + 2], S.alpha[q], 1.0, o);
        const double c = S.cn[q];
#pragma unroll
        for (int j = 0; j < 5; ++j) s[j] = fma(c, o[j], s[j]);
      }
#pragma unroll
      for (int j = 0; j < 5; ++j) AO[j * Ne * Na + it] = s[j];
      KC[it] = s[4] + 2.0 * (g[e] * s[1] + g[Ne + e] * s[2] + g[2 * Ne + e] * s[3]) + g[3 * Ne + e] * s[0];
    }
    __syncwarp();
    if (a.stop == 1) continue;
    // ---- MO = AO W, B = -1/2 K W (used columns)
    for (int it = sub; it < Ne * Nm; it += G) {
      const int e = it / Nm, m = it - e * Nm;
      double s0 = 0.0, s1 = 0.0;
      for (int ao = 0; ao < Na; ++ao) {
        const double wv = S.mow[ao * ld + m];
        s0 = fma(AO[e * Na + ao], wv, s0);
        s1 = fma(KC[e * Na + ao], wv, s1);
      }
      MO[it] = s0;
      BK[it] = -0.5 * s1;
    }
    __syncwarp();
    if (a.stop == 2) continue;
    // ---- spin blocks: inverse, determinant, trace, A^-1 B A^-1
    if (S.nup <= 6 && S.ndown <= 6) {
      // small blocks: one lane per unique block (lane_block), the blocks of the walker side by side
      for (int u = sub; u < nun; u += G) {
        const bool up = u < S.nuu;
        const int n = up ? S.nup : S.ndown, r0 = up ? 0 : S.nup;
        const int *cols = up ? S.ucu + u * S.nup : S.ucd + (u - S.nuu) * S.ndown;
        double *Ai = INV + (up ? 2 * S.nup * S.nup * u : 2 * S.nup * S.nup * S.nuu + 2 * S.ndown * S.ndown * (u - S.nuu));
        double *Yv = Ai + n * n;
        double det = 1.0, tr = 0.0;
        switch (n) {
          case 1: lane_block<1>(MO, BK, Nm, r0, cols, Ai, Yv, det, tr); break;
          case 2: lane_block<2>(MO, BK, Nm, r0, cols, Ai, Yv, det, tr); break;
          case 3: lane_block<3>(MO, BK, Nm, r0, cols, Ai, Yv, det, tr); break;
          case 4: lane_block<4>(MO, BK, Nm, r0, cols, Ai, Yv, det, tr); break;
          case 5: lane_block<5>(MO, BK, Nm, r0, cols, Ai, Yv, det, tr); break;
          case 6: lane_block<6>(MO, BK, Nm, r0, cols, Ai, Yv, det, tr); break;
          default: break;                               // n = 0: empty spin block, det = 1
        }
        Dt[u] = det;
        Dt[nun + u] = tr;
      }
      __syncwarp();
    } else {
      int off = 0;
      for (int u = 0; u < nun; ++u) {
        const bool up = u < S.nuu;
        const int n = up ? S.nup : S.ndown, r0 = up ? 0 : S.nup;
        const int *cols = up ? S.ucu + u * S.nup : S.ucd + (u - S.nuu) * S.ndown;
        const int ldg = 3 * n;
        for (int i = sub; i < n * n; i += G) {
          const int r = i / n, c = i - r * n;
          GJ[r * ldg + c] = MO[(r0 + r) * Nm + cols[c]];
          GJ[r * ldg + n + c] = r == c ? 1.0 : 0.0;
          GJ[r * ldg + 2 * n + c] = BK[(r0 + r) * Nm + cols[c]];
        }
        __syncwarp();
        const double det = group_gauss_jordan3<G>(GJ, col, n, sub);
        double tr = 0.0;
        for (int i = sub; i < n; i += G) tr += GJ[i * ldg + 2 * n + i];
        tr = group_sum<G>(tr);
        double *Ai = INV + off, *Yv = Ai + n * n;
        for (int i = sub; i < n * n; i += G) {
          const int r = i / n, c = i - r * n;
          Ai[i] = GJ[r * ldg + n + c];
          double s = 0.0;
          for (int k = 0; k < n; ++k) s = fma(GJ[r * ldg + 2 * n + k], GJ[k * ldg + n + c], s);
          Yv[i] = s;
        }
        if (sub == 0) { Dt[u] = det; Dt[nun + u] = tr; }
        off += 2 * n * n;
        __syncwarp();
      }
    }
    if (a.stop == 3) continue;
    // ---- CI sums
    double Ssum = 0.0, Tsum = 0.0;
    for (int c = sub; c < S.nconf; c += G) {
      const int iu = S.ciu[c], id = S.nuu + S.cid[c];
      const double dd = S.ci[c] * Dt[iu] * Dt[id];
      Ssum += dd;
      Tsum = fma(dd, Dt[nun + iu] + Dt[nun + id], Tsum);
    }
    Ssum = group_sum<G>(Ssum);
    Tsum = group_sum<G>(Tsum);
    const double Sinv = 1.0 / Ssum;
    const double Sb = wP * J - wE * Tsum * Sinv * Sinv, Tb = wE * Sinv;
    if (a.want_ci)
      accumulate<G>(acc, AL.o_ci, S.nconf, sub, first_group, [&](int c) {
        const int iu = S.ciu[c], id = S.nuu + S.cid[c];
        return Dt[iu] * Dt[id] * (Sb + Tb * (Dt[nun + iu] + Dt[nun + id]));
      });
    // adjoints of the determinants and traces of the unique blocks
    for (int u = sub; u < nun; u += G) {
      const bool up = u < S.nuu;
      double db = 0.0, tb = 0.0;
      for (int c = 0; c < S.nconf; ++c) {
        const int iu = S.ciu[c], id = S.nuu + S.cid[c];
        if ((up ? iu : id) != u) continue;
        const double other = Dt[up ? id : iu];
        db = fma((Sb + Tb * (Dt[nun + iu] + Dt[nun + id])) * S.ci[c], other, db);
        tb = fma(Tb * S.ci[c], Dt[iu] * Dt[id], tb);
      }
      Dt[2 * nun + u] = db;
      Dt[3 * nun + u] = tb;
    }
    for (int i = sub; i < Ne * Nm; i += G) { M0[i] = 0.0; Q[i] = 0.0; }
    __syncwarp();
    // ---- adjoint of MO (through the blocks) and Q = -1/2 adjoint of B
    {
      int off = 0;
      for (int u = 0; u < nun; ++u) {
        const bool up = u < S.nuu;
        const int n = up ? S.nup : S.ndown, r0 = up ? 0 : S.nup;
        const int *cols = up ? S.ucu + u * S.nup : S.ucd + (u - S.nuu) * S.ndown;
        const double *Ai = INV + off, *Yv = Ai + n * n;
        const double dD = Dt[2 * nun + u] * Dt[u], tb = Dt[3 * nun + u];
        for (int i = sub; i < n * n; i += G) {
          const int r = i / n, c = i - r * n;
          const int idx = (r0 + r) * Nm + cols[c];
          M0[idx] += dD * Ai[c * n + r] - tb * Yv[c * n + r];
          Q[idx] -= 0.5 * tb * Ai[c * n + r];
        }
        off += 2 * n * n;
        __syncwarp();
      }
    }
    if (a.stop == 4) continue;
    // ---- back through the projections
    for (int it = sub; it < Ne * Na; it += G) {
      const int e = it / Na, ao = it - e * Na;
      double s0 = 0.0, s1 = 0.0;
      for (int m = 0; m < Nm; ++m) {
        const double wv = S.mow[ao * ld + m];
        s0 = fma(M0[e * Nm + m], wv, s0);
        s1 = fma(Q[e * Nm + m], wv, s1);
      }
      MW[it] = s0;
      QW[it] = s1;
    }
    if (a.want_mo)
      accumulate<G>(acc, AL.o_w, Na * Nm, sub, first_group, [&](int it) {
        const int ao = it / Nm, m = it - ao * Nm;
        double s = 0.0;
        for (int e = 0; e < Ne; ++e)
          s = fma(AO[e * Na + ao], M0[e * Nm + m], fma(KC[e * Na + ao], Q[e * Nm + m], s));
        return s;
      });
    __syncwarp();
    if (a.stop == 5) continue;
    // ---- leaves: basis parameters and atom coordinates through the AO channels
    if (a.want_atom) prim_leaf_pass<G, 4>(S, SL, AL, x, sc, acc, sub, first_group, a.want_coef != 0);
    else if (a.want_exp) prim_leaf_pass<G, 1>(S, SL, AL, x, sc, acc, sub, first_group, a.want_coef != 0);
    else if (a.want_coef) prim_leaf_pass<G, 0>(S, SL, AL, x, sc, acc, sub, first_group, true);
    if (a.stop == 6) continue;
    // ---- leaves: Pade weights (e-e, e-n) through g, l and J
    if ((a.want_jee && S.use_jee) || (a.want_jen && S.use_jen)) {
      // adjoints of g_e, l_e from the kinetic channel: g~ = 2 sum_a d ao QW, l~ = sum_a ao QW
      for (int i = sub; i < 4 * Ne; i += G) {
        const int k = i / Ne, e = i - k * Ne;
        const double *ch = AO + (k < 3 ? (1 + k) : 0) * Ne * Na + e * Na;
        double s = 0.0;
        for (int ao = 0; ao < Na; ++ao) s = fma(ch[ao], QW[e * Na + ao], s);
        gb[i] = k < 3 ? 2.0 * s : s;
      }
      __syncwarp();
      typedef Dual<2> T;
      const double Kb = wP * J * Ssum;          // adjoint of ln J: J~ J with J~ = wP S
      double tj0 = 0.0, tj1 = 0.0;
      for (int e = sub; e < Ne; e += G) {
        const double xe = x[3 * e], ye = x[3 * e + 1], ze = x[3 * e + 2];
        const double lb = gb[3 * Ne + e];
        // l = sum_b lap K_b + |g|^2  ->  adjoint of grad K picks up 2 l~ g
        const double bx = fma(2.0 * lb, g[e], gb[e]), by = fma(2.0 * lb, g[Ne + e], gb[Ne + e]),
                     bz = fma(2.0 * lb, g[2 * Ne + e], gb[2 * Ne + e]);
        T ks = mk<2>(0.0), kx = ks, ky = ks, kz = ks, kl = ks, kn = ks;
        if (S.use_jee) {
          const T wj = seed<2>(S.jee_w, 0);
          for (int j = 0; j < Ne; ++j) {
            if (j == e) continue;
            const double dx = xe - x[3 * j], dy = ye - x[3 * j + 1], dz = ze - x[3 * j + 2];
            const double r = sqrt(dx * dx + dy * dy + dz * dz), rinv = 1.0 / r;
            const double w0 = ((e < S.nup) == (j < S.nup)) ? 0.25 : 0.5;     // pade_jastrow_kernel.py:34-66
            const T den = rcp(wj * r + 1.0);
            const T k1 = w0 * (den * den);                   // k'
            const T k2 = (-2.0 * w0) * (wj * (den * (den * den)));   // k''
            ks = ks + (w0 * r) * den;
            const T kr = k1 * rinv;
            kx = kx + kr * dx; ky = ky + kr * dy; kz = kz + kr * dz;
            kl = kl + k2 + 2.0 * kr;
          }
        }
        if (S.use_jen) {
          const T wn = seed<2>(S.jen_w, 1);
          for (int A = 0; A < S.natom; ++A) {
            const double dx = xe - S.atoms[4 * A], dy = ye - S.atoms[4 * A + 1], dz = ze - S.atoms[4 * A + 2];
            const double r = sqrt(dx * dx + dy * dy + dz * dz), rinv = 1.0 / r;
            const T den = rcp(wn * r + 1.0);
            const T k1 = den * den;
            const T k2 = -2.0 * (wn * (den * (den * den)));
            kn = kn + r * den;
            const T kr = k1 * rinv;
            kx = kx + kr * dx; ky = ky + kr * dy; kz = kz + kr * dz;
            kl = kl + k2 + 2.0 * kr;
          }
        }
        // ln J = 1/2 sum_e sum_{j != e} k_ee + sum_e sum_A k_en
        tj0 += Kb * 0.5 * ks.d[0] + bx * kx.d[0] + by * ky.d[0] + bz * kz.d[0] + lb * kl.d[0];
        tj1 += Kb * kn.d[1] + bx * kx.d[1] + by * ky.d[1] + bz * kz.d[1] + lb * kl.d[1];
      }
      tj0 = across_groups<G>(group_sum<G>(tj0));
      tj1 = across_groups<G>(group_sum<G>(tj1));
      if (lane == 0) { acc[AL.o_jee] += tj0; acc[AL.o_jen] += tj1; }
    }
    // ---- potentials: d V_en / d R_A = -Z_A (r_e - R_A) / r^3 ; V_nn is added once by vjp_finish (sum of wE)
    if (a.want_atom && a.wE) {
      accumulate<G>(acc, AL.o_ven, 3 * S.natom, sub, first_group, [&](int i) {
        const int A = i / 3, k = i - 3 * A;
        double s = 0.0;
        for (int e = 0; e < Ne; ++e) {
          const double dx = x[3 * e] - S.atoms[4 * A], dy = x[3 * e + 1] - S.atoms[4 * A + 1],
                       dz = x[3 * e + 2] - S.atoms[4 * A + 2];
          const double r2 = dx * dx + dy * dy + dz * dz, rinv = rsqrt(r2);
          s -= S.atoms[4 * A + 3] * (k == 0 ? dx : (k == 1 ? dy : dz)) * rinv * rinv * rinv;
        }
        return wE * s;
      });
      const double sE = across_groups<G>(sub == 0 ? wE : 0.0);
      if (lane == 0) acc[AL.o_sumE] += sE;
    }
    __syncwarp();
  }
  if (a.smem)
    for (int i = lane; i < AL.n; i += 32) gacc[i] += acc[i];
}

struct VjpOut {
  double *g_mo, *g_ci, *g_exp, *g_coef, *g_jee, *g_jen, *g_atom;
};

// Sums the per-warp accumulators in warp order (thread = entry: coalesced rows, fixed order) ...
__global__ void vjp_reduce(const double *acc, int nwarp, int n, double *tot) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int w = 0; w < nwarp; ++w) s += acc[(size_t)w * n + i];
  tot[i] = s;
}
// ... and maps the totals onto the outputs.
__global__ void vjp_finish(const VjpSys S, const double *tot, VjpOut o) {
  const AccLayout AL(S);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  if (o.g_mo)
    for (int i = tid; i < S.nao * S.nmo; i += nth) {
      const int ao = i / S.nmo, m = i - ao * S.nmo;
      int pos = -1;
      for (int j = 0; j < S.nmu; ++j) if (S.used[j] == m) pos = j;
      o.g_mo[i] = pos < 0 ? 0.0 : tot[AL.o_w + ao * S.nmu + pos];
    }
  if (o.g_ci) for (int i = tid; i < S.nconf; i += nth) o.g_ci[i] = tot[AL.o_ci + i];
  if (o.g_exp) for (int i = tid; i < S.nbas; i += nth) o.g_exp[i] = tot[AL.o_exp + i];
  if (o.g_coef) for (int i = tid; i < S.nbas; i += nth) o.g_coef[i] = tot[AL.o_coef + i];
  if (o.g_jee && tid == 0) o.g_jee[0] = tot[AL.o_jee];
  if (o.g_jen && tid == 0) o.g_jen[0] = tot[AL.o_jen];
  if (o.g_atom)
    for (int i = tid; i < 3 * S.natom; i += nth) {
      const int A = i / 3, k = i - 3 * A;
      double s = tot[AL.o_ven + i];
      for (int q = 0; q < S.nbas; ++q)
        if (S.patom[q] == A) s += tot[AL.o_pr + 3 * q + k];
      // nuclear repulsion (wf_base.py:97-116): d V_nn / d R_A = - sum_B Z_A Z_B (R_A - R_B) / |R_A - R_B|^3
      double v = 0.0;
      for (int B = 0; B < S.natom; ++B) {
        if (B == A) continue;
        const double dx = S.atoms[4 * A] - S.atoms[4 * B], dy = S.atoms[4 * A + 1] - S.atoms[4 * B + 1],
                     dz = S.atoms[4 * A + 2] - S.atoms[4 * B + 2];
        const double r2 = dx * dx + dy * dy + dz * dz, r = sqrt(r2);
        v -= S.atoms[4 * A + 3] * S.atoms[4 * B + 3] * (k == 0 ? dx : (k == 1 ? dy : dz)) / (r2 * r);
      }
      o.g_atom[i] = s + tot[AL.o_sumE] * v;
    }
}

constexpr int kChunk = 1 << 17;      // walkers per Jastrow-operator chunk

// launch shape: CTAs of kThreads; G lanes per walker by the size of the system (items per phase ~ nelec * nao);
// shared-memory work areas when at least one CTA fits, else global scratch with a whole warp per walker
struct VjpLaunch { int grid, threads, smem_bytes, use_smem, G; };
VjpLaunch launch_of(const qmcb_plan *p, const VjpSys &S) {
  const ScratchLayout SL(S);
  const AccLayout AL(S);
  const char *eg = getenv("QMCB_VJP_G");           // tuning: force the group size
  int G = 4;
  while (G < 32 && G * 6 < S.nelec * S.nao) G *= 2;
  if (eg && (atoi(eg) == 4 || atoi(eg) == 8 || atoi(eg) == 16 || atoi(eg) == 32)) G = atoi(eg);
  VjpLaunch L{};
  const size_t budget = (size_t)p->smem_optin - 1024, sm_total = (size_t)227 * 1024;
  const int reg_warps = 4 * kMinBlocks;              // resident warps the register budget of __launch_bounds__ allows
  // candidates: (G, warps per CTA); keep the one with the most resident warps per SM, then the smaller G
  int best_warps = 0;
  for (int g = G; g <= 32; g *= 2)
    for (int wpc = kThreads / 32; wpc >= 1; wpc /= 2) {
      const size_t per_cta = (size_t)wpc * ((size_t)(32 / g) * SL.n + AL.n) * 8;
      if (per_cta > budget) continue;
      int per_sm = (int)(sm_total / (per_cta + 1024));
      if (per_sm * wpc > reg_warps) per_sm = reg_warps / wpc;
      if (per_sm > 16) per_sm = 16;
      if (per_sm < 1) continue;
      // walkers in flight weigh more than warps: a smaller group wins ties
      if (per_sm * wpc > best_warps) {
        best_warps = per_sm * wpc;
        L.use_smem = 1; L.smem_bytes = (int)per_cta; L.grid = p->sm_count * per_sm; L.G = g; L.threads = 32 * wpc;
      }
    }
  if (best_warps >= 6) return L;
  // large systems: a whole warp per walker on global (L2-resident) scratch
  // (shared memory then holds the read-only tables, see the kernel prologue)
  const size_t tab = ((size_t)4 * S.natom + (size_t)S.nao * S.nmup + 3 * (size_t)S.nbas) * 8 +
                     ((size_t)5 * S.nbas + 2 * (size_t)S.nao + 1) * 4 + 16;
  L.use_smem = 0; L.smem_bytes = tab <= budget ? (int)tab : -1; L.G = 32; L.threads = kThreads; L.grid = p->sm_count * kMinBlocks;
  return L;
}

VjpSys make_sys(const qmcb_plan *p) {
  const DevSys &D = p->sys;
  VjpSys S{};
  S.nelec = D.nelec; S.nup = D.nup; S.ndown = D.ndown; S.natom = D.natom; S.nbas = D.nbas; S.nao = D.nao; S.nmo = D.nmo;
  S.nmu = D.nmu; S.nmup = D.nmup; S.nconf = D.nconf; S.nuu = D.nuu; S.nud = D.nud; S.radial_type = D.radial_type;
  S.use_jee = D.use_jee; S.use_jen = D.use_jen;
  S.has_j = D.use_jee || D.use_jen || D.een_nterm > 0;
  S.jee_w = D.jee_w; S.jen_w = D.jen_w;
  S.atoms = D.dblob + D.o_atoms; S.mow = D.dblob + D.o_mow; S.ci = D.dblob + D.o_ci;
  S.used = D.iblob + D.o_used; S.ucu = D.iblob + D.o_ucu; S.ucd = D.iblob + D.o_ucd;
  S.ciu = D.iblob + D.o_ciu; S.cid = D.iblob + D.o_cid;
  const double *fd = p->d_flat_dbl;
  const int *fi = p->d_flat_int;
  const int nb = D.nbas;
  S.alpha = fd; S.cn = fd + nb; S.norm = fd + 2 * nb;
  S.patom = fi; S.pk = fi + nb; S.pkr = fi + 2 * nb; S.pao = fi + 3 * nb;
  S.ao_start = fi + 4 * nb; S.ao_prim = fi + 4 * nb + D.nao + 1; S.ao_order = fi + 5 * nb + D.nao + 1;
  return S;
}

size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }

}  // namespace

extern "C" int64_t qmcb_local_energy_backward_workspace_bytes(const qmcb_plan *p, int64_t W) {
  if (!p || W < 0) return 0;
  VjpSys S = make_sys(p);
  const ScratchLayout SL(S);
  const AccLayout AL(S);
  const VjpLaunch LC = launch_of(p, S);
  const int64_t nwarp = (int64_t)LC.grid * (LC.threads / 32);
  const int64_t wc = W < kChunk ? W : kChunk;
  size_t n = align256((size_t)nwarp * (32 / LC.G) * SL.n * 8) + align256((size_t)nwarp * AL.n * 8) +
             align256((size_t)AL.n * 8);
  n += align256((size_t)wc * 8) + align256((size_t)wc * 3 * S.nelec * 8) + align256((size_t)wc * S.nelec * 8);
  return (int64_t)n + 256;
}

extern "C" int qmcb_local_energy_backward(const qmcb_plan *p, const double *pos, const double *w_eloc,
                                          const double *w_psi, int64_t W, double *g_mo, double *g_ci,
                                          double *g_bas_exp, double *g_bas_coeffs, double *g_jee_w, double *g_jen_w,
                                          double *g_atom_coords, void *workspace, void *stream) {
  if (!p || !p->d_dbl || !p->d_flat_dbl || !pos || W < 0 || !workspace || (!w_eloc && !w_psi)) {
    qmcb_set_error("qmcb_local_energy_backward: bad arguments");
    return QMCB_EINVAL;
  }
  if (p->sys.nup > 64 || p->sys.ndown > 64) {
    qmcb_set_error("qmcb_local_energy_backward: spin blocks larger than 64 x 64");
    return QMCB_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const VjpSys S = make_sys(p);
  const ScratchLayout SL(S);
  const AccLayout AL(S);
  const VjpLaunch LC = launch_of(p, S);
  const int grid = LC.grid;
  const int nwarp = grid * (LC.threads / 32);
  void (*kern)(const VjpSys, const VjpArgs) =
      LC.G == 4 ? eloc_vjp_kernel<4> : (LC.G == 8 ? eloc_vjp_kernel<8> : (LC.G == 16 ? eloc_vjp_kernel<16> : eloc_vjp_kernel<32>));
  if (LC.smem_bytes < 0) {
    qmcb_set_error("qmcb_local_energy_backward: basis tables exceed the shared memory of an SM");
    return QMCB_ESMEM;
  }
  {
    cudaError_t ea = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, LC.smem_bytes);
    if (ea != cudaSuccess) return qmcb_cuda_rc((int)ea, "qmcb_local_energy_backward smem");
  }
  const int64_t wc = W < kChunk ? W : kChunk;
  char *ws = (char *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  double *scratch = (double *)ws; ws += align256((size_t)nwarp * (32 / LC.G) * SL.n * 8);
  double *acc = (double *)ws; ws += align256((size_t)nwarp * AL.n * 8);
  double *tot = (double *)ws; ws += align256((size_t)AL.n * 8);
  double *J = (double *)ws; ws += align256((size_t)wc * 8);
  double *dJ = (double *)ws; ws += align256((size_t)wc * 3 * S.nelec * 8);
  double *d2J = (double *)ws;
  cudaError_t e = cudaMemsetAsync(acc, 0, (size_t)nwarp * AL.n * 8, st);
  if (e != cudaSuccess) return qmcb_cuda_rc((int)e, "qmcb_local_energy_backward memset");
  VjpArgs a{};
  a.pos = pos; a.wE = w_eloc; a.wP = w_psi; a.W = W;
  a.scratch = scratch; a.acc = acc; a.smem = LC.use_smem;
  { const char *es = getenv("QMCB_VJP_STOP"); a.stop = es ? atoi(es) : 0; }
  a.want_mo = g_mo != nullptr; a.want_ci = g_ci != nullptr; a.want_exp = g_bas_exp != nullptr;
  a.want_coef = g_bas_coeffs != nullptr; a.want_jee = g_jee_w != nullptr; a.want_jen = g_jen_w != nullptr;
  a.want_atom = g_atom_coords != nullptr;
  for (int64_t w0 = 0; w0 < W; w0 += wc) {
    const int64_t w1 = w0 + wc < W ? w0 + wc : W;
    if (S.has_j && p->sys.een_nterm > 0) {
      const int rc = qmcb_jastrow(p, pos + w0 * 3 * S.nelec, w1 - w0, 0, J, dJ, d2J, stream);
      if (rc) return rc;
      a.J = J; a.dJ = dJ; a.d2J = d2J;
    }
    a.w0 = w0; a.w1 = w1;
    kern<<<grid, LC.threads, LC.smem_bytes, st>>>(S, a);
    if ((e = cudaGetLastError()) != cudaSuccess) return qmcb_cuda_rc((int)e, "eloc_vjp_kernel launch");
  }
  VjpOut o{g_mo, g_ci, g_bas_exp, g_bas_coeffs, g_jee_w, g_jen_w, g_atom_coords};
  vjp_reduce<<<(AL.n + 127) / 128, 128, 0, st>>>(acc, nwarp, AL.n, tot);
  vjp_finish<<<8, 256, 0, st>>>(S, tot, o);
  return qmcb_cuda_rc((int)cudaGetLastError(), "vjp_finish launch");
}
