// Score-transport pathwise kernel, 384 threads per CTA x 32 columns (reads of up to 12 287 bases): see pathwise_tr_impl.cuh.
#define PWT_NT 384
#define PWT_FN(name) name##_wide
#include "pathwise_tr_impl.cuh"
