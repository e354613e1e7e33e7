// hp_b200_shim.cpp -- the stub a maintainer of the reference adds in place of nndistance.cu + approxmatch.cu
// (utils/pytorch_structural_losses/setup.py:9-13): it defines the five launchers that the reference's own pybind glue
// declares at utils/pytorch_structural_losses/structural_loss.cpp:11-15 and forwards them to the C ABI of libhp_b200.so
// (include/hp_b200.h).  structural_loss.cpp itself -- tensor allocation, CHECK_INPUT, current stream -- stays untouched.
// build_shim.sh compiles exactly that pair (the reference's unmodified structural_loss.cpp + this file) and
// tests/test_native_binding.py runs the resulting StructuralLossesBackend module.
#include <cuda_runtime.h>

#include <stdexcept>
#include <string>

#include "hp_b200.h"

static void ok(int rc, const char *what) {
    // the reference's launchers throw std::runtime_error on a failed launch (approxmatch.cu:334-337); so does the shim
    if (rc != HP_OK) throw std::runtime_error(std::string(what) + ": " + hp_error_string(rc) + ": " + hp_last_error_message());
}

void approxmatch(int b, int n, int m, const float *xyz1, const float *xyz2, float *match, float *temp, cudaStream_t stream) {
    ok(hp_approxmatch(b, n, m, xyz1, xyz2, match, temp, (void *)stream), "approxmatch");
}
void matchcost(int b, int n, int m, const float *xyz1, const float *xyz2, float *match, float *out, cudaStream_t stream) {
    ok(hp_matchcost(b, n, m, xyz1, xyz2, match, out, (void *)stream), "matchcost");
}
void matchcostgrad(int b, int n, int m, const float *xyz1, const float *xyz2, const float *match, float *grad1, float *grad2,
                   cudaStream_t stream) {
    ok(hp_matchcostgrad(b, n, m, xyz1, xyz2, match, grad1, grad2, (void *)stream), "matchcostgrad");
}
void nndistance(int b, int n, const float *xyz, int m, const float *xyz2, float *result, int *result_i, float *result2,
                int *result2_i, cudaStream_t stream) {
    ok(hp_nndistance(b, n, xyz, m, xyz2, result, result_i, result2, result2_i, (void *)stream), "nndistance");
}
void nndistancegrad(int b, int n, const float *xyz1, int m, const float *xyz2, const float *grad_dist1, const int *idx1,
                    const float *grad_dist2, const int *idx2, float *grad_xyz1, float *grad_xyz2, cudaStream_t stream) {
    ok(hp_nndistancegrad(b, n, xyz1, m, xyz2, grad_dist1, idx1, grad_dist2, idx2, grad_xyz1, grad_xyz2, (void *)stream),
       "nndistancegrad");
}
