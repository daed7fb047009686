// Entry points declared in include/hrmgpu.h whose kernels are not built yet: they fail loudly.
#include "internal.cuh"
using namespace hrmgpu;
struct hrmgpu_ctx : public Ctx {};
#define NOT_YET(c, what) ((c) ? fail((c), HRMGPU_E_STATE, what ": not implemented in this build") : HRMGPU_E_ARG)
extern "C" {
int hrmgpu_set_scene2d(hrmgpu_ctx* c, int, int, const double*, int) { return NOT_YET(c, "set_scene2d"); }
int hrmgpu_build_layers2d(hrmgpu_ctx* c, int, int, const double*, const double*, const double*, int) { return NOT_YET(c, "build_layers2d"); }
int hrmgpu_test_edges(hrmgpu_ctx* c, int, int, const double*, const double*, unsigned char*) { return NOT_YET(c, "test_edges"); }
int hrmgpu_build_bridge(hrmgpu_ctx* c, int, int, const double*, const double*, int) { return NOT_YET(c, "build_bridge"); }
int hrmgpu_build_bridge2d(hrmgpu_ctx* c, int, int, const double*, const double*, int) { return NOT_YET(c, "build_bridge2d"); }
int hrmgpu_test_bridge(hrmgpu_ctx* c, int, int, const double*, unsigned char*) { return NOT_YET(c, "test_bridge"); }
}
