// Host-side plan and the device table layout shared by all kernels.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/qmcb.h"

// Everything a kernel needs to know about the system, passed by value as a kernel
// parameter.  The tables themselves live in one double blob + one int blob in global
// memory; every CTA stages them in shared memory once.
struct DevSys {
  int nelec, nup, ndown, natom;
  int nshell, nprim, ncomp;      // grouped basis: shells -> primitives / components
  int nao, nmo;
  int nmu, nmup;                 // MO columns used by some configuration; padded count
  int nconf, nuu, nud;           // configurations; unique spin-up / spin-down occupations
  int radial_type, use_jee, use_jen, gram_fma;
  int nbas;                      // flat primitives (gradient outputs)
  double jee_w, jen_w, vnn;
  const double *dblob;
  const int *iblob;
  int ndbl, nint;
  // offsets into dblob (doubles)
  int o_atoms;   // [natom][4]  x y z Z
  int o_alpha;   // [nprim]
  int o_coef;    // [nprim]     coeff * norm of the first component
  int o_pn;      // [nprim]     radial power n (gto/sto), as double
  int o_cscale;  // [ncomp]     component scale relative to the shell's coef vector
  int o_mow;     // [nao][nmup] MO weights of the used columns (zero padded)
  int o_ci;      // [nconf]
  // offsets into iblob (ints)
  int o_ash;     // [natom+1]   first shell of each atom
  int o_spo;     // [nshell+1]  first primitive of each shell
  int o_sco;     // [nshell+1]  first component of each shell
  int o_ck;      // [ncomp]     kx | ky<<8 | kz<<16
  int o_cao;     // [ncomp]     AO index
  int o_used;    // [nmu]       MO index of used column
  int o_ucu;     // [nuu][nup]  unique up occupations, as positions in the used list
  int o_ucd;     // [nud][ndown]
  int o_ciu;     // [nconf]     unique-up index of each configuration
  int o_cid;     // [nconf]
  int o_pflat;   // per shell: [nprim_s][ncomp_s] flat primitive index; start at o_pfo[s]
  int o_pfo;     // [nshell+1]
  int o_fnorm;   // dblob: [nbas] norm of each flat primitive (gradient post-processing)
  int o_etab;    // dblob: [QMCB_ETAB] 2^(j/QMCB_ETAB), table of the exp() range reduction
  // packed "shell program" walked by the hot loops: 16-byte records (double2), see plan.cu
  int o_stream;  // offset into dblob (doubles, even)
  int nrec;      // number of 16-byte records
  // exp() polynomial + range-reduction constants.  Kernel parameters live in constant bank 0,
  // which DFMA can take as a direct operand: no UMOV pairs / LDC to materialise 64-bit immediates.
  double expc[16];
  // Boys-Handy three-body Jastrow parameters (constant bank): f = a r/(1+b r), g = a2 r/(1+b2 r)
  int een_nterm;
  double een_a[QMCB_EEN_MAXTERM], een_b[QMCB_EEN_MAXTERM], een_a2[QMCB_EEN_MAXTERM],
      een_b2[QMCB_EEN_MAXTERM], een_c[QMCB_EEN_MAXTERM];
};

struct LaunchCfg {
  int warp;      // 1: one warp owns a tile (no CTA barriers), 0: one CTA owns a tile
  int tw;        // walkers per tile
  int nblk;      // MO column blocks per electron
  int mb;        // MO columns per block (template)
  int threads;
  int smem;      // dynamic shared memory bytes
  int lu_conc;   // concurrent LU scratch slots (0: closed forms only)
};

struct qmcb_spec_state;   // spec.cu

struct qmcb_plan {
  int device = 0;
  bool multi_component = false;                  // some AO is a sum of several cartesian monomials (sph l = 2)
  uint64_t version = 0;                          // bumped by every table rebuild
  mutable qmcb_spec_state *spec = nullptr;       // structure-specialised kernels (lazy)
  DevSys sys{};
  std::vector<double> hd;
  std::vector<int> hi;
  double *d_dbl = nullptr;
  int *d_int = nullptr;
  size_t cap_dbl = 0, cap_int = 0;
  int sm_count = 148;
  int smem_optin = 227 * 1024;
  LaunchCfg cfg_psi{}, cfg_eloc{}, cfg_grad{};
  // backward (parameter gradients): contraction tiles over walker rows, see backward.cu
  struct BwdCfg {
    int tw = 0, rows = 0, threads = 0, smem = 0, lu_conc = 0;
    int lda = 0, ldg = 0, ldx = 0, ppad = 0;      // leading dimensions (doubles); ppad = nprim padded to 8
    int ntile_mo = 0, ntile_ao = 0;              // 8x8 output tiles
    int nslot = 0;                               // doubles per CTA partial
    int grid = 0;
  } bwd, bwd0;   // with / without the basis-parameter gradients
  std::vector<int> bwd_tiles;                    // [ntile][2] (row block, col block) ; MO first, then AO
  int *d_bwd_tiles = nullptr;
  size_t cap_bwd_tiles = 0;
  // full MO matrix for the operator-level entry point
  double *d_mo_full = nullptr;
  size_t cap_mo_full = 0;
  // arrival counters of the fused energy statistics (the CTA of an E_L kernel that arrives last adds the
  // partials); a counter is zero between launches (atomicInc wraps), so launches that share one must be
  // stream-ordered: every stream gets its own slot (qmcb_ticket_slot), and a stream that finds no free
  // slot takes the separate second-stage launch instead
  unsigned *d_ticket = nullptr;
  static constexpr int kTicketSlots = 16;
  mutable void *ticket_owner[kTicketSlots] = {};
  mutable bool ticket_used[kTicketSlots] = {};
  // flat primitive list in the reference's own order (qmcb_system: one entry per primitive per cartesian
  // monomial) for the adjoint of the local energy (eloc_vjp.cu): doubles alpha | norm*coeff | norm,
  // ints atom | kx,ky,kz packed | radial power | AO | CSR AO -> primitives (start [nao+1], list [nbas]) |
  // AOs by decreasing contraction length [nao]
  std::vector<double> flat_dbl;
  std::vector<int> flat_int;
  double *d_flat_dbl = nullptr;
  int *d_flat_int = nullptr;
  size_t cap_flat_dbl = 0, cap_flat_int = 0;
  // host copy of flat data needed by backward post-processing
  std::vector<int> index_ctr;
  std::vector<double> mo_full;
};

// Switches to the plan's device for the scope of a host entry point and restores the caller's
// current device on exit (a finalizer may call qmcb_plan_destroy at any time: it must not change
// the device the caller - e.g. torch - believes to be current).
struct DeviceGuard {
  int prev = -1;
  bool active = false;
  explicit DeviceGuard(int device) {
    if (device < 0) return;
    if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
    if (prev != device) { active = cudaSetDevice(device) == cudaSuccess; }
  }
  ~DeviceGuard() { if (active && prev >= 0) cudaSetDevice(prev); }
};

// device pointer of the arrival counter reserved for `stream` on this plan, or nullptr (all slots taken)
unsigned *qmcb_ticket_slot(const qmcb_plan *p, void *stream);
void qmcb_set_error(const std::string &msg);
// maps a cudaError_t to the ABI return code and records its text for qmcb_last_error()
int qmcb_cuda_rc(int cuda_error, const char *where);
int qmcb_build_tables(const qmcb_system *s, qmcb_plan *p);   // host grouping -> hd/hi/sys
int qmcb_choose_launch(qmcb_plan *p);
int qmcb_choose_backward(qmcb_plan *p);
// second stage of the energy statistics: n per-CTA partials [n][4] -> out4 (operators.cu)
int qmcb_stats_finish(const double *part, int n, double *out4, void *stream);
#define QMCB_STATS_MAX_PARTIALS 4096
