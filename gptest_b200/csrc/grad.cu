// grad.cu - hyper-parameter gradients of the regression likelihood (stage under construction).
#include "../../include/gpb200.h"
#include "gpb_context.cuh"

namespace gpb {
int gpr_nlml_grad_chunk(gpb_handle* h, const double*, int64_t, double, double*, double*, int32_t*) {
  h->err = "gradient stage not built yet";
  return -4;
}
}  // namespace gpb
