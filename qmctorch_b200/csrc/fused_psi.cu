// qmcb_psi: instantiates the fused kernel in MODE_PSI.
#define QMCB_FUSED_MAIN
#include "fused_impl.cuh"
#include "spec.h"

extern "C" int qmcb_psi(const qmcb_plan *p, const double *pos, int64_t W, double *psi, void *stream) {
  int rc = check(p, pos, W);
  if (rc || W == 0) return rc;
  FusedArgs a{};
  a.pos = pos; a.W = W; a.out0 = psi;
  rc = qmcb_spec_launch(p, MODE_PSI, a, stream);   // structure-specialised kernel, when this plan has one
  if (rc != QMCB_SPEC_SKIP) return rc;
  return launch<MODE_PSI>(p, p->cfg_psi, a, (cudaStream_t)stream);
}
