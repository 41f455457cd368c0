/* C-side smoke test of the drop-in boundary: compiled with a plain C compiler against include/qmcb.h
 * and linked with libqmcb.so - no torch, no C++.  Builds the H2 STO-3G wave function of BASELINE
 * config 1 by hand (public STO-3G table; MOs (phi1 +- phi2)/sqrt2, column-normalised like
 * scf/calculator/calculator_base.py:35-46), creates a host-only plan (device -1: table grouping only)
 * and - when a CUDA device is present - a device plan, evaluates psi and E_L on four walkers through
 * qmcb_psi / qmcb_local_energy and checks them against the closed form evaluated right here, and the atom-coordinate
 * derivative of psi from qmcb_local_energy_backward against a finite difference of that closed form.
 * Exit code 0 = pass (prints "ABI_OK host" or "ABI_OK device").
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "qmcb.h"

#define NW 4

static double dfact(int n) { double r = 1.0; for (; n > 1; n -= 2) r *= n; return r; }

/* closed form of the H2 STO-3G wave function of one walker for the given atom positions:
 * psi = J * sigma_g(r1) * sigma_g(r2), sigma_g = (phi_A + phi_B)/sqrt2, J = exp(0.5 r12 / (1 + r12)) */
static double psi_closed_form(const double *x, const double *atoms, const double *expo, const double *coef) {
  const double pi = 3.14159265358979323846, s = 1.0 / sqrt(2.0);
  double mo_e[2];
  for (int e = 0; e < 2; ++e) {
    double v = 0.0;
    for (int a = 0; a < 2; ++a) {
      const double dx = x[3 * e] - atoms[3 * a], dy = x[3 * e + 1] - atoms[3 * a + 1], dz = x[3 * e + 2] - atoms[3 * a + 2];
      const double r2 = dx * dx + dy * dy + dz * dz;
      for (int q = 0; q < 3; ++q) v += s * coef[q] * pow(2.0 * expo[q] / pi, 0.75) * exp(-expo[q] * r2);
    }
    mo_e[e] = v;
  }
  const double dx = x[0] - x[3], dy = x[1] - x[4], dz = x[2] - x[5];
  const double r12 = sqrt(dx * dx + dy * dy + dz * dz);
  return exp(0.5 * r12 / (1.0 + r12)) * mo_e[0] * mo_e[1];
}

int main(void) {
  const double expo[3] = {3.42525091, 0.62391373, 0.16885540};
  const double coef[3] = {0.15432897, 0.53532814, 0.44463454};
  const double pi = 3.14159265358979323846;
  double atom_coords[6] = {0, 0, -0.69, 0, 0, 0.69}, Z[2] = {1, 1};
  int32_t bas_atom[6], kx[6], ky[6], kz[6], kr[6], index_ctr[6];
  double bas_exp[6], bas_coeffs[6], bas_norm[6];
  for (int i = 0; i < 6; ++i) {
    bas_atom[i] = i / 3; index_ctr[i] = i / 3; kx[i] = ky[i] = kz[i] = kr[i] = 0;
    bas_exp[i] = expo[i % 3]; bas_coeffs[i] = coef[i % 3];
    bas_norm[i] = pow(2.0 * bas_exp[i] / pi, 0.75) / sqrt(dfact(-1));   /* norm_orbital.py:136-161, k = 0 */
  }
  const double s = 1.0 / sqrt(2.0);
  double mo[4] = {s, s, s, -s};                 /* [nao=2, nmo=2] */
  int32_t cfg_up[1] = {0}, cfg_down[1] = {0};
  double ci[1] = {1.0};
  qmcb_system sys;
  memset(&sys, 0, sizeof(sys));
  sys.nelec = 2; sys.nup = 1; sys.ndown = 1; sys.natom = 2; sys.nbas = 6; sys.nao = 2; sys.nmo = 2;
  sys.radial_type = QMCB_GTO_PURE; sys.contract = 1;
  sys.atom_coords = atom_coords; sys.atomic_number = Z; sys.bas_atom = bas_atom; sys.bas_exp = bas_exp;
  sys.bas_coeffs = bas_coeffs; sys.bas_norm = bas_norm; sys.bas_kx = kx; sys.bas_ky = ky; sys.bas_kz = kz;
  sys.bas_kr = kr; sys.index_ctr = index_ctr; sys.mo = mo; sys.nconf = 1; sys.cfg_up = cfg_up;
  sys.cfg_down = cfg_down; sys.ci = ci; sys.use_jee = 1; sys.jee_w = 1.0; sys.use_jen = 0; sys.jen_w = 0.0;
  sys.gram_fma = 0; sys.een_nterm = 0;

  if (qmcb_abi_version() != QMCB_ABI_VERSION) { fprintf(stderr, "ABI version mismatch\n"); return 2; }
  qmcb_plan *host = NULL;
  if (qmcb_plan_create(&sys, -1, &host) != 0) { fprintf(stderr, "host plan: %s\n", qmcb_last_error()); return 3; }
  /* two contracted s shells of three primitives each, one occupied MO column, one determinant per spin */
  if (qmcb_plan_info(host, 0) != 2 || qmcb_plan_info(host, 1) != 6 || qmcb_plan_info(host, 3) != 1 ||
      qmcb_plan_info(host, 4) != 1 || qmcb_plan_info(host, 5) != 1) {
    fprintf(stderr, "unexpected table grouping\n"); return 4;
  }
  /* a compute call on a host-only plan must fail loudly, not fall back */
  double dummy;
  if (qmcb_psi(host, &dummy, 1, &dummy, NULL) == 0) { fprintf(stderr, "host plan evaluated psi\n"); return 5; }
  qmcb_plan_destroy(host);
  /* bad input: inconsistent electron counts */
  sys.nup = 2;
  if (qmcb_plan_create(&sys, -1, &host) == 0) { fprintf(stderr, "inconsistent sizes accepted\n"); return 6; }
  sys.nup = 1;

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { printf("ABI_OK host (no CUDA device)\n"); return 0; }
  qmcb_plan *plan = NULL;
  if (qmcb_plan_create(&sys, 0, &plan) != 0) { fprintf(stderr, "device plan: %s\n", qmcb_last_error()); return 7; }
  double pos[NW * 6] = {0.1, 0.2, -0.5, -0.3, 0.1, 0.6,   0.4, -0.2, 0.9, 0.0, 0.3, -1.1,
                        -0.6, 0.5, 0.2, 0.7, -0.4, 0.1,   0.05, 0.0, -0.7, 0.0, -0.05, 0.72};
  double *d_pos, *d_psi, *d_el, psi[NW], el[NW];
  cudaMalloc((void **)&d_pos, sizeof(pos)); cudaMalloc((void **)&d_psi, sizeof(psi)); cudaMalloc((void **)&d_el, sizeof(el));
  cudaMemcpy(d_pos, pos, sizeof(pos), cudaMemcpyHostToDevice);
  if (qmcb_psi(plan, d_pos, NW, d_psi, NULL) != 0) { fprintf(stderr, "qmcb_psi: %s\n", qmcb_last_error()); return 8; }
  if (qmcb_local_energy(plan, d_pos, NW, d_el, NULL, NULL, NULL) != 0) { fprintf(stderr, "qmcb_local_energy: %s\n", qmcb_last_error()); return 9; }
  cudaMemcpy(psi, d_psi, sizeof(psi), cudaMemcpyDeviceToHost);
  cudaMemcpy(el, d_el, sizeof(el), cudaMemcpyDeviceToHost);
  for (int w = 0; w < NW; ++w) {
    const double ref = psi_closed_form(pos + w * 6, atom_coords, expo, coef);
    if (fabs(psi[w] - ref) > 1e-12 * fabs(ref)) { fprintf(stderr, "psi[%d] = %.17g, expected %.17g\n", w, psi[w], ref); return 10; }
    if (!(el[w] == el[w]) || fabs(el[w]) > 1e3) { fprintf(stderr, "E_L[%d] = %g\n", w, el[w]); return 11; }
  }
  /* adjoint entry point (grad="auto" / forces): d/dR_A of sum_w psi_w against a central finite difference of the
   * closed form (the e-e Jastrow factor does not depend on the atoms) */
  {
    double *d_w, *d_ga, *d_ws, ones[NW] = {1, 1, 1, 1}, ga[6];
    const int64_t nb = qmcb_local_energy_backward_workspace_bytes(plan, NW);
    cudaMalloc((void **)&d_w, sizeof(ones)); cudaMalloc((void **)&d_ga, sizeof(ga)); cudaMalloc((void **)&d_ws, (size_t)nb);
    cudaMemcpy(d_w, ones, sizeof(ones), cudaMemcpyHostToDevice);
    if (qmcb_local_energy_backward(plan, d_pos, NULL, d_w, NW, NULL, NULL, NULL, NULL, NULL, NULL, d_ga, d_ws, NULL) != 0) {
      fprintf(stderr, "qmcb_local_energy_backward: %s\n", qmcb_last_error()); return 12;
    }
    cudaMemcpy(ga, d_ga, sizeof(ga), cudaMemcpyDeviceToHost);
    for (int i = 0; i < 6; ++i) {
      const double h = 1e-5;
      double ap[6], am[6], fp = 0.0, fm = 0.0;
      memcpy(ap, atom_coords, sizeof(ap)); memcpy(am, atom_coords, sizeof(am));
      ap[i] += h; am[i] -= h;
      for (int w = 0; w < NW; ++w) { fp += psi_closed_form(pos + w * 6, ap, expo, coef); fm += psi_closed_form(pos + w * 6, am, expo, coef); }
      const double fd = (fp - fm) / (2.0 * h);
      if (fabs(ga[i] - fd) > 1e-7 * (fabs(fd) + 1e-3)) { fprintf(stderr, "d psi / d R[%d] = %.12g, finite difference %.12g\n", i, ga[i], fd); return 13; }
    }
    if (qmcb_local_energy_backward(plan, d_pos, NULL, NULL, NW, NULL, NULL, NULL, NULL, NULL, NULL, d_ga, d_ws, NULL) == 0) {
      fprintf(stderr, "adjoint without weights accepted\n"); return 14;
    }
    cudaFree(d_w); cudaFree(d_ga); cudaFree(d_ws);
  }
  cudaFree(d_pos); cudaFree(d_psi); cudaFree(d_el);
  qmcb_plan_destroy(plan);
  printf("ABI_OK device psi[0]=%.12g E_L[0]=%.12g\n", psi[0], el[0]);
  return 0;
}
