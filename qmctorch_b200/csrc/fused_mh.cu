// qmcb_metropolis_step: instantiates the fused kernel in MODE_MH.
#include "fused_impl.cuh"
#include "spec.h"

extern "C" int qmcb_metropolis_step(const qmcb_plan *p, double *pos, double *fx, int64_t W,
                                    const double *disp, const double *tau, const int32_t *elec_index,
                                    int move_elec, int proba_normal, double scale, double eps,
                                    uint64_t seed, uint64_t offset, uint8_t *accept,
                                    unsigned long long *naccept, void *stream) {
  int rc = check(p, pos, W);
  if (rc || W == 0) return rc;
  if (move_elec >= p->sys.nelec || move_elec < -2) { qmcb_set_error("qmcb_metropolis_step: move_elec"); return QMCB_EINVAL; }
  FusedArgs a{};
  a.pos = pos; a.pos_rw = pos; a.W = W; a.out0 = fx;
  a.disp = disp; a.tau = tau; a.elec_index = elec_index; a.move_elec = move_elec;
  a.proba_normal = proba_normal; a.scale = scale; a.eps = eps; a.seed = seed; a.offset = offset;
  a.accept = accept; a.naccept = naccept;
  rc = qmcb_spec_launch(p, MODE_MH, a, stream);   // structure-specialised kernel, when this plan has one
  if (rc != QMCB_SPEC_SKIP) return rc;
  return launch<MODE_MH>(p, p->cfg_psi, a, (cudaStream_t)stream);
}
