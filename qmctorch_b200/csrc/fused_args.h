// Argument block of the fused walker kernels (generic and structure-specialised).  Plain C types
// only: this header is also compiled by NVRTC.
#pragma once

enum { MODE_PSI = 0, MODE_ELOC = 1, MODE_GRAD = 2, MODE_MH = 3, MODE_BWD = 4, MODE_BWD_ALL = 5 };

// entries of the 2^(j/N) table of the exp() range reduction (device.cuh: exp_core); power of two
#ifndef QMCB_ETAB_LOG2
#define QMCB_ETAB_LOG2 7
#endif
#define QMCB_ETAB (1 << QMCB_ETAB_LOG2)

struct FusedArgs {
  const double *pos;   // [W,3Ne]   (MH: updated in place through pos_rw)
  double *pos_rw;
  int64_t W;
  double *out0;        // psi [W]           | eloc [W]      | grad [W,3Ne] | fx [W] (in/out)
  double *out1;        // -                 | psi or null   | -            | -
  double *out2;        // -                 | ekin or null  | -            | -
  int pdf;             // GRAD: return grad psi^2
  // Metropolis
  const double *disp;  // [W,3Ne] or null
  const double *tau;   // [W] or null
  const int *elec_index;
  int move_elec, proba_normal;
  double scale, eps;
  uint64_t seed, offset;
  uint8_t *accept;
  unsigned long long *naccept;
  // E_L: per-CTA partial statistics [gridDim.x][4] = sum, sum sq, n finite, n non-finite (or null)
  double *stats_part;
  // with both set, the CTA that arrives last adds the partials in index order -> stats_out[4]
  unsigned *stats_ticket;
  double *stats_out;
  // parameter-gradient backward of the one-walker-per-thread kernels (MODE_BWD): weight [W] of
  // psi.backward(weight), per-CTA partial sums [gridDim.x][SPEC_NAO * SPEC_NMUP + nconf + 2]
  const double *weight;
  double *bwd_part;
};

