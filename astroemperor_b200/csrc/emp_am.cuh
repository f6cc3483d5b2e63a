// emp_am.cuh — Hipparcos-Gaia astrometric block (emp_model.py:1232-1672).
// PLACEHOLDER until the device kernel lands: the handle refuses AM models instead of
// silently dropping the astrometric term.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include "../../include/emperor_b200.h"

static int fail(int code, const std::string& msg);

namespace emp {
struct AmDevice {
  int enabled = 0;
};
inline int am_upload(const EmpAmData*, AmDevice*) {
  return fail(EMP_EUNSUPPORTED, "astrometric block not implemented on the device path yet");
}
inline void am_free(AmDevice*) {}
inline int am_launch(AmDevice*, const EmpModelDesc*, const double*, int64_t, double*, cudaStream_t, int64_t*) {
  return fail(EMP_EUNSUPPORTED, "astrometric block not implemented on the device path yet");
}
}  // namespace emp
