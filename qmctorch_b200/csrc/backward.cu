// placeholder, replaced below
#include "plan.h"
extern "C" int64_t qmcb_backward_workspace_bytes(const qmcb_plan *, int64_t) { return 0; }
extern "C" int qmcb_psi_backward(const qmcb_plan *, const double *, const double *, int64_t, double *, double *,
                                 double *, double *, double *, double *, void *, void *) {
  qmcb_set_error("qmcb_psi_backward: not built yet");
  return QMCB_EINVAL;
}
