/*
 * qmcb.h - C ABI of the B200 (sm_100a) walker-parallel Slater-Jastrow hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  The
 * reference (QMCTorch v0.4.0) is pure Python and has no FFI of its own; each entry
 * point below replaces the Python method(s) cited next to it (paths relative to
 * qmctorch/), and INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Conventions
 *   - every array argument of a compute call is a DEVICE pointer owned by the caller
 *     (torch); FP64 unless stated; row-major; W = number of walkers.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *   - compute calls allocate nothing, never synchronise, and return 0 or a
 *     cudaError_t value (> 0) / a negative QMCB_E* code; qmcb_last_error() gives text.
 *   - the plan owns small device tables (basis, MO weights, configurations,
 *     Jastrow parameters); it is immutable between qmcb_plan_update() calls.
 */
#ifndef QMCB_H
#define QMCB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QMCB_ABI_VERSION 1

#define QMCB_EINVAL   (-1)  /* bad argument / unsupported configuration */
#define QMCB_ENOMEM   (-2)
#define QMCB_ESMEM    (-3)  /* problem does not fit the shared-memory tiling */

/* radial_type, atomic_orbitals.py:81-88 */
#define QMCB_GTO_PURE 0
#define QMCB_GTO      1
#define QMCB_STO_PURE 2
#define QMCB_STO      3

typedef struct qmcb_plan qmcb_plan;

/* Host-side description of one wave function: exactly the data the reference keeps in
 * AtomicOrbitals (atomic_orbitals.py:27-94), MolecularOrbitals (molecular_orbitals.py:40-74),
 * SlaterPooling/OrbitalConfigurations (slater_pooling.py:28-62) and the Pade Jastrow
 * kernels.  All pointers are HOST pointers, copied by qmcb_plan_create/update. */
typedef struct {
  int32_t nelec, nup, ndown;
  int32_t natom;
  int32_t nbas;            /* flat primitives (len(bas_exp)) */
  int32_t nao, nmo;
  int32_t radial_type;     /* QMCB_GTO_PURE ... */
  int32_t contract;        /* AtomicOrbitals.contract (atomic_orbitals.py:49) */
  const double *atom_coords;   /* [natom,3]  ao.atom_coords                     */
  const double *atomic_number; /* [natom]                                       */
  const int32_t *bas_atom;     /* [nbas] atom of each primitive (from nshells)  */
  const double *bas_exp;       /* [nbas] ao.bas_exp                             */
  const double *bas_coeffs;    /* [nbas] ao.bas_coeffs                          */
  const double *bas_norm;      /* [nbas] ao.norm_cst (frozen at construction)   */
  const int32_t *bas_kx, *bas_ky, *bas_kz, *bas_kr; /* [nbas] cartesian powers, radial power */
  const int32_t *index_ctr;    /* [nbas] AO index of each primitive             */
  const double *mo;            /* [nao,nmo] mo_scf * mo_modifier                */
  int32_t nconf;
  const int32_t *cfg_up;       /* [nconf,nup]   MO indices, wf.configs[0]       */
  const int32_t *cfg_down;     /* [nconf,ndown] MO indices, wf.configs[1]       */
  const double *ci;            /* [nconf] fc.weight                             */
  int32_t use_jee;  double jee_w;   /* e-e Pade Jastrow + its weight            */
  int32_t use_jen;  double jen_w;   /* e-n Pade Jastrow + its weight            */
  int32_t gram_fma;            /* 0: r_ij dot product unfused (ATen small-bmm path, Ne<=11),
                                  1: fma chain (MKL path); electron_electron_distance.py:177-190 */
  /* three-body Boys-Handy e-e-n Jastrow (elec_elec_nuclei/kernels/boys_handy_jastrow_kernel.py:8-93),
     exponents 1: K = sum_mu c_mu f_mu(r_iA) f_mu(r_jA) g_mu(r_ij), f = a r/(1+b r), g = a' r/(1+b' r) */
  int32_t een_nterm;           /* 0: term absent; <= QMCB_EEN_MAXTERM                   */
  const double *een_num;       /* [2,nterm] weight_num   (row 0: a, row 1: a')          */
  const double *een_denom;     /* [2,nterm] weight_denom (row 0: b, row 1: b')          */
  const double *een_fc;        /* [nterm]   fc.weight                                   */
} qmcb_system;

#define QMCB_EEN_MAXTERM 8

int         qmcb_abi_version(void);
const char *qmcb_last_error(void);

/* Build / refresh / free the device tables.  update() takes a full description again
 * (after opt.step(); the tables are a few KB). */
int  qmcb_plan_create(const qmcb_system *sys, int device, qmcb_plan **out);
int  qmcb_plan_update(qmcb_plan *plan, const qmcb_system *sys);
void qmcb_plan_destroy(qmcb_plan *plan);
/* introspection for tests: 0 nshell, 1 nprim(grouped), 2 ncomp, 3 nmo_used, 4 nuniq_up,
 * 5 nuniq_down, 6 walkers per CTA (local energy), 7 threads per CTA, 8 dynamic smem bytes,
 * 9 walkers per CTA (psi) [6-9: tiling of the generic kernels], 10-12 backward tiling,
 * 13: 1 when this plan runs structure-specialised (NVRTC) kernels - compiles / loads them now;
 *     0: generic kernels, qmcb_last_error() says why,
 * 14: kind of specialised kernel the structure is eligible for (0 none, 1 one walker per thread,
 *     2 warp tiles), 15: kind in use (compiles / loads now) */
int  qmcb_plan_info(const qmcb_plan *plan, int what);

/* --- fused hot path --------------------------------------------------------------- */

/* psi(R) = J * sum_n c_n Dup_n Ddown_n        replaces SlaterJastrow.forward
 * (wavefunction/slater_jastrow.py:243-286).   pos [W,3*nelec] -> psi [W] */
int qmcb_psi(const qmcb_plan *plan, const double *pos, int64_t W, double *psi, void *stream);

/* E_L = E_kin(Jacobi) + V_en + V_ee + V_nn     replaces WaveFunction.local_energy
 * (wavefunction/wf_base.py:184-215) with kinetic_energy_jacobi (slater_jastrow.py:312-344,
 * 449-482) and SlaterPooling.operator (pooling/slater_pooling.py:262-387).
 * pos [W,3*nelec] -> eloc [W]; psi, ekin optional (may be NULL) [W] */
int qmcb_local_energy(const qmcb_plan *plan, const double *pos, int64_t W,
                      double *eloc, double *psi, double *ekin, void *stream);

/* grad psi (pdf=0) or grad psi^2 (pdf=1)       replaces SlaterJastrow.gradients_jacobi
 * (slater_jastrow.py:346-447).  -> grad [W,3*nelec] */
int qmcb_grad_psi(const qmcb_plan *plan, const double *pos, int64_t W, int pdf,
                  double *grad, void *stream);

/* One Metropolis move for every walker, in place.   replaces the body of the loop in
 * Metropolis.__call__ (sampler/metropolis.py:134-160) with move (:227-254), _move (:256-277)
 * and _accept (:279-298) for logspace=False.
 *   pos [W,3*nelec], fx [W] (= psi^2 of pos, zeros already replaced by eps) are updated;
 *   disp [W,3*nelec] proposal displacement (injected draws) or NULL -> in-kernel Philox
 *       normal(0, sqrt(sigma)) (proba "normal") / uniform step*(2u-1) (proba "uniform");
 *   tau [W] uniform draws or NULL -> Philox;
 *   accept [W] uint8 out (may be NULL); naccept: device int64 counter, atomically incremented
 *       (may be NULL).
 *   move_elec: -1 all electrons, >=0 only that electron ("all-elec-iter"),
 *       -2 one random electron per walker (elec_index [W] int32 injected, or Philox if NULL). */
int qmcb_metropolis_step(const qmcb_plan *plan, double *pos, double *fx, int64_t W,
                         const double *disp, const double *tau, const int32_t *elec_index,
                         int move_elec, int proba_normal, double scale, double eps,
                         uint64_t seed, uint64_t offset,
                         uint8_t *accept, unsigned long long *naccept, void *stream);

/* psi.backward(weight): per-parameter sums over walkers of weight_w * d psi_w / d theta.
 * replaces the autograd backward in Solver.evaluate_grad_manual (solver/solver.py:414-429).
 * Any output pointer may be NULL.  Outputs are OVERWRITTEN (caller accumulates into .grad).
 *   g_mo [nao,nmo] (w.r.t. the effective weights mo_scf*mo_modifier), g_ci [nconf],
 *   g_bas_exp [nbas], g_bas_coeffs [nbas], g_jee_w [1], g_jen_w [1],
 *   g_een [5*nterm] = weight_num [2,nterm] | weight_denom [2,nterm] | fc [nterm];
 *   workspace: device scratch of qmcb_backward_workspace_bytes(plan, W) bytes. */
int64_t qmcb_backward_workspace_bytes(const qmcb_plan *plan, int64_t W);
int qmcb_psi_backward(const qmcb_plan *plan, const double *pos, const double *weight, int64_t W,
                      double *g_mo, double *g_ci, double *g_bas_exp, double *g_bas_coeffs,
                      double *g_jee_w, double *g_jen_w, double *g_een, void *workspace, void *stream);

/* Adjoint of the local energy (and of psi) w.r.t. the parameters AND the atom coordinates:
 *   out_theta = sum_w  w_eloc[w] * d E_L(R_w) / d theta  +  w_psi[w] * d psi(R_w) / d theta .
 * replaces the autograd backward THROUGH WaveFunction.local_energy in Solver.evaluate_grad_auto
 * (solver/solver.py:352-370, loss.backward() with grad="auto") and the two autograd.grad calls of
 * Solver.compute_forces (solver/solver.py:433-519: d E_L / d atom_coords, d log psi^2 / d atom_coords).
 *   w_eloc, w_psi [W]: either may be NULL (taken as zero), not both.
 *   g_mo [nao,nmo] (effective weights mo_scf*mo_modifier), g_ci [nconf], g_bas_exp [nbas],
 *   g_bas_coeffs [nbas], g_jee_w [1], g_jen_w [1], g_atom_coords [natom,3] - through the AOs, V_en and
 *   V_nn like the reference, whose Jastrow factors hold their own constant copy of the atom positions
 *   (jastrow_factor_electron_nuclei.py:40-41).  Any output may be NULL; outputs are OVERWRITTEN.
 *   The three-body (Boys-Handy) term contributes through grad J / J and lap J / J; its own weights
 *   have no output here (the reference's graph drops the Laplacian's dependence on them).
 *   workspace: qmcb_local_energy_backward_workspace_bytes(plan, W) bytes.  Deterministic. */
int64_t qmcb_local_energy_backward_workspace_bytes(const qmcb_plan *plan, int64_t W);
int qmcb_local_energy_backward(const qmcb_plan *plan, const double *pos, const double *w_eloc,
                               const double *w_psi, int64_t W, double *g_mo, double *g_ci,
                               double *g_bas_exp, double *g_bas_coeffs, double *g_jee_w, double *g_jen_w,
                               double *g_atom_coords, void *workspace, void *stream);

/* [sum E, sum E^2, count of finite, count of non-finite] -> out[4]; deterministic two-stage
 * reduction.  replaces torch.mean/var in SolverBase.single_point (solver/solver_base.py:371),
 * wf_base.py:217-229.  workspace: qmcb_stats_workspace_bytes(W). */
int64_t qmcb_stats_workspace_bytes(int64_t W);
int qmcb_energy_stats(const double *eloc, int64_t W, double *out4, void *workspace, void *stream);

/* qmcb_local_energy followed by qmcb_energy_stats in ONE call: eloc [W] (required), psi / ekin
 * optional, out4 as qmcb_energy_stats.  One VMC energy step of Solver.single_point
 * (solver/solver_base.py:355-371: local_energy per batch, then mean / var).  Structure-specialised
 * kernels reduce their walkers inside the E_L kernel (no second pass over eloc, no second launch:
 * the CTA that arrives last adds the per-CTA partials in index order; the plan keeps one arrival
 * counter per stream it has been called on (16; further streams take a separate second-stage
 * launch), so concurrent streams may share a plan - each with its OWN workspace and out4).
 * workspace: qmcb_stats_workspace_bytes(W). */
int qmcb_local_energy_stats(const qmcb_plan *plan, const double *pos, int64_t W, double *eloc,
                            double *psi, double *ekin, double *out4, void *workspace, void *stream);

/* --- operator-level entry points (sub-module parity; HBM-bound by construction) ---- */

/* AtomicOrbitals.forward(pos, derivative=[0,1,2]) (orbitals/atomic_orbitals.py:578-609):
 * ao [W,ne,nao], dao [W,ne,nao,3], d2ao [W,ne,nao]; dao/d2ao may be NULL (values only).
 * ne = nelec, or 1 when one_elec != 0 (pos is then [W,3]). */
int qmcb_ao(const qmcb_plan *plan, const double *pos, int64_t W, int one_elec,
            double *ao, double *dao, double *d2ao, void *stream);

/* MolecularOrbitals.forward (orbitals/molecular_orbitals.py:79-95): x [rows,nao] -> [rows,nmo] */
int qmcb_mo(const qmcb_plan *plan, const double *x, int64_t rows, double *out, void *stream);

/* Jastrow factor and derivatives, product of the configured terms
 * (jastrow_factor_electron_electron.py:124-260, jastrow_factor_electron_nuclei.py:60-161,
 * combine_jastrow.py:33-195): J [W], dJ [W,3,nelec], d2J [W,nelec] (dJ, d2J may be NULL).
 * which: 0 product of all, 1 e-e only, 2 e-n only, 3 e-e-n only. */
int qmcb_jastrow(const qmcb_plan *plan, const double *pos, int64_t W, int which,
                 double *J, double *dJ, double *d2J, void *stream);

/* SlaterPooling.forward / .operator (pooling/slater_pooling.py:64-80,262-387):
 * mo [W,nelec,nmo] -> dets [W,nconf];  bop [nop,W,nelec,nmo] -> trace [nop,W,nconf]
 * (bop/trace may be NULL). */
int qmcb_slater(const qmcb_plan *plan, const double *mo, const double *bop, int64_t nop,
                int64_t W, double *dets, double *trace, void *stream);

/* FP64 pipe probes for the roofline denominator: runs `iters` dependent-chain DFMA (kind 0),
 * DMMA m8n8k4 (kind 1) or both interleaved (kind 2) per thread on the whole GPU and returns flop count through
 * *flops; time it with events on `stream`. */
int qmcb_fp64_probe(int kind, int64_t iters, double *sink, double *flops, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* QMCB_H */
