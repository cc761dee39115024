// laplace.cu - Laplace / Newton mode finding for GPc and GPpref (stage under construction).
#include "../../include/gpb200.h"
#include "gpb_context.cuh"

extern "C" {
int gpb_gpc_laplace(gpb_handle* h, const double*, const double*, int32_t, double, int32_t, int32_t, double*,
                    double*, int32_t*, double*, double*, int32_t*) {
  if (h) h->err = "gpc_laplace not built yet";
  return -4;
}
int gpb_gpc_predict(gpb_handle* h, const double*, int64_t, double*, double*, double*) {
  if (h) h->err = "gpc_predict not built yet";
  return -4;
}
int gpb_pref_laplace(gpb_handle* h, const int64_t*, const double*, int64_t, const double*, double, double, int32_t,
                     int32_t, int32_t, double*, double*, int32_t*, double*, double*, int32_t*) {
  if (h) h->err = "pref_laplace not built yet";
  return -4;
}
int gpb_pref_derivatives(gpb_handle* h, const int64_t*, const double*, int64_t, int64_t, const double*, double,
                         int32_t, double*, double*) {
  if (h) h->err = "pref_derivatives not built yet";
  return -4;
}
}
