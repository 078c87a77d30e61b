// Score-transport pathwise kernel, 256 threads per CTA (reads of up to 8 191 bases): see pathwise_tr_impl.cuh.
#define PWT_NT 256
#define PWT_FN(name) name
#include "pathwise_tr_impl.cuh"
