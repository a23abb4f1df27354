// rb_kernels_c32.cu -- kernels specialised at compile time on the synthetic 32-joint chain of BASELINE.json
// configs[4] (assets/chain32.urdf -> gen/model_chain32.h), long-chain layout (rb_kernels_long.cuh).
// rb_api.cu selects this family only when an uploaded chain equals the table bit for bit.
#include "rb_kernels_long.cuh"
#include "gen/model_chain32.h"

const RbOps* rb_ops_chain32() {
    static const RbOps ops = RbLaunchLong<CtModel<TabChain32>>::ops("chain32-specialised");
    return &ops;
}

const double* rb_chain32_table() {
    // filled once, thread-safely (function-local static initialisation): engines may be created from several threads
    struct Flat {
        double v[RB_MODEL_DOUBLES(32)];
        Flat() {
            for (int i = 0; i < 32; ++i)
                for (int k = 0; k < 24; ++k) v[i * 24 + k] = TabChain32::T[i][k];
            for (int k = 0; k < 3; ++k) v[32 * 24 + k] = TabChain32::G[k];
            for (int k = 0; k < 9; ++k) v[32 * 24 + 3 + k] = TabChain32::TIP[k];
        }
    };
    static const Flat flat;
    return flat.v;
}
