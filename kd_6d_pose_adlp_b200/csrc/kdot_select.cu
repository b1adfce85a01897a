// placeholder -- replaced by the real selection kernel in the next milestone
#include "kdot_common.cuh"
extern "C" int kdot_select_cells(const float* const*, const float* const*, const int32_t*, const int32_t*, const float*,
                                 int, const float*, int, int, int, float, int, float, int, int32_t*, int32_t*, int32_t*,
                                 float*, float*, int32_t*, int32_t*, int32_t*, void*) {
  return KDOT_E_BADARG;
}
