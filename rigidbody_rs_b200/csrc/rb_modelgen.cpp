// rb_modelgen -- build-time tool: URDF -> compile-time model table header for CtModel<> (rb_dyn.cuh).
// usage: rb_modelgen <urdf> <TabName> <out.h>
#include <cstdio>
#include <fstream>
#include "rb_host_model.h"

int main(int argc, char** argv) {
    if (argc != 4) { fprintf(stderr, "usage: %s <urdf> <TabName> <out.h>\n", argv[0]); return 2; }
    RbHostModel m; std::string err;
    if (rb_model_from_urdf(argv[1], m, err) != RB_OK) { fprintf(stderr, "rb_modelgen: %s\n", err.c_str()); return 1; }
    std::ofstream f(argv[3]);
    f << rb_model_emit_header(m, argv[2]);
    return f.good() ? 0 : 1;
}
