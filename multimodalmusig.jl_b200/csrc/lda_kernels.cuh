// lda_kernels.cuh -- LDA variational-EM iteration (reference src/LDA.jl:69-224).
#pragma once
#include "det_math.cuh"

namespace mmsig {
struct LdaDev {
    int K, V;
    long long D, D_total;
};
}  // namespace mmsig
