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

// E_L and its statistics in one pass.  With a structure-specialised kernel the first statistics
// stage is fused into the E_L kernel (per-CTA partials, fixed order); otherwise the generic E_L
// kernel is followed by the two statistics kernels of qmcb_energy_stats.
extern "C" int qmcb_local_energy_stats(const qmcb_plan *p, const double *pos, int64_t W, double *eloc,
                                       double *psi, double *ekin, double *out4, void *workspace, void *stream) {
  int rc = check(p, pos, W);
  if (rc) return rc;
  if (!eloc || !out4 || !workspace) { qmcb_set_error("qmcb_local_energy_stats: bad arguments"); return QMCB_EINVAL; }
  if (W == 0) return qmcb_energy_stats(eloc, 0, out4, workspace, stream);
  FusedArgs a{};
  a.pos = pos; a.W = W; a.out0 = eloc; a.out1 = psi; a.out2 = ekin;
  a.stats_part = (double *)workspace;
  // the last CTA of the specialised kernel finishes the reduction (grid <= SMs x CTAs per SM, far
  // below QMCB_STATS_MAX_PARTIALS); QMCB_STATS_2STAGE=1 keeps the separate second-stage launch
  static const bool two_stage = getenv("QMCB_STATS_2STAGE") && atoi(getenv("QMCB_STATS_2STAGE")) > 0;
  if (!two_stage && (int64_t)p->sm_count * 16 <= QMCB_STATS_MAX_PARTIALS) {
    // one arrival counter per (plan, stream): concurrent streams on one plan do not share a counter
    a.stats_ticket = qmcb_ticket_slot(p, stream);
    if (a.stats_ticket) a.stats_out = out4;
  }
  int grid = 0;
  rc = qmcb_spec_launch(p, MODE_ELOC, a, stream, &grid);
  if (rc == 0) {
    if (a.stats_ticket) return 0;
    if (grid > QMCB_STATS_MAX_PARTIALS) { qmcb_set_error("qmcb_local_energy_stats: grid exceeds the workspace"); return QMCB_EINVAL; }
    return qmcb_stats_finish((const double *)workspace, grid, out4, stream);
  }
  if (rc != QMCB_SPEC_SKIP) return rc;
  a.stats_part = nullptr; a.stats_ticket = nullptr; a.stats_out = nullptr;
  rc = launch<MODE_ELOC>(p, p->cfg_eloc, a, (cudaStream_t)stream);
  if (rc) return rc;
  return qmcb_energy_stats(eloc, W, out4, workspace, stream);
}
