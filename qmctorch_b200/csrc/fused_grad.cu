// qmcb_grad_psi: instantiates the fused kernel in MODE_GRAD.
#include "fused_impl.cuh"
#include "spec.h"

extern "C" int qmcb_grad_psi(const qmcb_plan *p, const double *pos, int64_t W, int pdf, double *grad,
                             void *stream) {
  int rc = check(p, pos, W);
  if (rc || W == 0) return rc;
  FusedArgs a{};
  a.pos = pos; a.W = W; a.out0 = grad; a.pdf = pdf;
  rc = qmcb_spec_launch(p, MODE_GRAD, a, stream);   // structure-specialised kernel, when this plan has one
  if (rc != QMCB_SPEC_SKIP) return rc;
  return launch<MODE_GRAD>(p, p->cfg_grad, a, (cudaStream_t)stream);
}
