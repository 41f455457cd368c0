// qmcb_grad_psi: instantiates the fused kernel in MODE_GRAD.
#include "fused_impl.cuh"

extern "C" int qmcb_grad_psi(const qmcb_plan *p, const double *pos, int64_t W, int pdf, double *grad,
                             void *stream) {
  int rc = check(p, pos, W);
  if (rc || W == 0) return rc;
  FusedArgs a{};
  a.pos = pos; a.W = W; a.out0 = grad; a.pdf = pdf;
  return launch<MODE_GRAD>(p, p->cfg_grad, a, (cudaStream_t)stream);
}
