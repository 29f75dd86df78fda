// One translation unit per generated node program: nvcc -DFL_VARIANT=<name> variant.cu
// (the straight-line bodies are ~8-10 k fp64 operations each; separate units compile in parallel).
#include "nlp_kernels.cuh"

#define FL_STR2(x) #x
#define FL_STR(x) FL_STR2(x)
#define FL_CAT2(a, b) a##b
#define FL_CAT(a, b) FL_CAT2(a, b)
#include FL_STR(generated/FL_CAT(gen_, FL_VARIANT).cuh)

extern "C" const fl_variant* FL_CAT(fl_variant_, FL_VARIANT)()
{
    static const fl_variant v = variant_ops<FL_CAT(gen_, FL_VARIANT)>::make();
    return &v;
}
