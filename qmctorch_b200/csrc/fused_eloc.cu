// qmcb_local_energy: instantiates the fused kernel in MODE_ELOC.
#include "fused_impl.cuh"
#include "spec.h"

extern "C" int qmcb_local_energy(const qmcb_plan *p, const double *pos, int64_t W, double *eloc,
                                 double *psi, double *ekin, void *stream) {
  int rc = check(p, pos, W);
  if (rc || W == 0) return rc;
  FusedArgs a{};
  a.pos = pos; a.W = W; a.out0 = eloc; a.out1 = psi; a.out2 = ekin;
  rc = qmcb_spec_launch(p, MODE_ELOC, a, stream);   // structure-specialised kernel, when this plan has one
  if (rc != QMCB_SPEC_SKIP) return rc;
  return launch<MODE_ELOC>(p, p->cfg_eloc, a, (cudaStream_t)stream);
}
