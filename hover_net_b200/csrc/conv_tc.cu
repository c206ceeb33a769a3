// tcgen05 convolution path -- placeholder until the kernel lands (plan always declines).
#include "cnn_kernels.h"
namespace hvn {
bool tc_plan(const ConvParams &, TcPlan &plan) { plan.ok = false; return false; }
void tc_launch(const ConvParams &, const TcPlan &, cudaStream_t) {}
}  // namespace hvn
