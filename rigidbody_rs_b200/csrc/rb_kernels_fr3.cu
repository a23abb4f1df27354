// rb_kernels_fr3.cu -- kernels specialised at compile time on the FR3 model (assets/fr3.urdf): the table in
// gen/model_fr3.h (written by rb_modelgen at build time) lets CtModel<> drop every multiplication by the exact
// zeros and +-1 of the arm's fixed joint rotations (all Rx(+-pi/2) or identity) and sparse joint offsets.
// rb_api.cu selects this family only when an uploaded chain equals the table bit for bit.
#include "rb_kernels.cuh"
#include "gen/model_fr3.h"

const RbOps* rb_ops_fr3() {
    static const RbOps ops = RbLaunch<CtModel<TabFr3>>::ops<CtModel<TabFr3, float>>("fr3-specialised");
    return &ops;
}

const double* rb_fr3_table() {
    // filled once, thread-safely (function-local static initialisation): engines may be created from several threads
    struct Flat {
        double v[RB_MODEL_DOUBLES(7)];
        Flat() {
            for (int i = 0; i < 7; ++i)
                for (int k = 0; k < 24; ++k) v[i * 24 + k] = TabFr3::T[i][k];
            for (int k = 0; k < 3; ++k) v[7 * 24 + k] = TabFr3::G[k];
            for (int k = 0; k < 9; ++k) v[7 * 24 + 3 + k] = TabFr3::TIP[k];
        }
    };
    static const Flat flat;
    return flat.v;
}
