// Structure-specialised (NVRTC) kernels: host interface, see spec.cu.
#pragma once
#include <string>

#include "fused_args.h"
#include "plan.h"

#define QMCB_SPEC_SKIP 0x7fff0001   // not handled: use the generic kernel

// Launches the specialised kernel for `mode` if this plan has one; QMCB_SPEC_SKIP otherwise.
// grid_out (optional) receives the number of CTAs launched (= partial statistics written).
int qmcb_spec_launch(const qmcb_plan *p, int mode, const FusedArgs &a, void *stream, int *grid_out = nullptr);
// 1 when the specialised kernels are compiled (and loaded on the plan's device); `why` = reason if not
int qmcb_spec_status(const qmcb_plan *p, std::string *why);
int qmcb_spec_eligible(const qmcb_plan *p);
int qmcb_spec_kind(const qmcb_plan *p);
void qmcb_spec_free(qmcb_plan *p);
