// rb_kernels_rt.cu -- unrolled kernels for any 7-joint serial chain; model constants read at run time from the
// kernel-parameter constant bank (RtModel<7>).  Serves 7-joint chains other than the compiled-in FR3.
#include "rb_kernels.cuh"

const RbOps* rb_ops_rt7() {
    static const RbOps ops = RbLaunch<RtModel<7>>::ops<RtModel<7, float>>("generic-7");
    return &ops;
}
