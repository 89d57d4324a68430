// K3 placeholder: filled in by the analysis milestone.
#include "gcb_internal.h"
extern "C" {
int gcb_analyse(const gcb_tgrid *, const float *, const gcb_analysis_opts *, gcb_analysis *)
{ return gcb::fail(GCB_ERR_STATE, "gcb_analyse: not built yet"); }
void gcb_analysis_free(gcb_analysis *) {}
int gcb_gridcalc(int, const float *const *, const float *, const gcb_tgrid *, float *, float *, int)
{ return gcb::fail(GCB_ERR_STATE, "gcb_gridcalc: not built yet"); }
}
