// tcgen05 tensor-core convolution path — placeholder until the kernel lands.
#include "common.cuh"

struct rgbd_conv_tc_plan { int unused; };

extern "C" int rgbd_conv_tc_plan_create(const rgbd_conv_desc *, int32_t, rgbd_conv_tc_plan **) {
    rgbd_set_error("rgbd_conv_tc_plan_create: tensor-core path not built");
    return RGBD_E_UNSUPPORTED;
}
extern "C" int rgbd_conv_tc_run(const rgbd_conv_tc_plan *, void *) {
    rgbd_set_error("rgbd_conv_tc_run: tensor-core path not built");
    return RGBD_E_UNSUPPORTED;
}
extern "C" void rgbd_conv_tc_plan_destroy(rgbd_conv_tc_plan *) {}
