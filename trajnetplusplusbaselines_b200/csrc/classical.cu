// Classical crowd simulators -- placeholder until the persistent kernels land.
#include "common.cuh"

extern "C" {

int tb2_sf_simulate(const tb2_layout*, const tb2_sf_params*, const float*, float*, void*) {
    tb2::set_error("tb2_sf_simulate: not built yet");
    return TB2_ERR_UNSUPPORTED;
}

int tb2_orca_simulate(const tb2_layout*, const tb2_orca_params*, const float*, const float*,
                      const float*, const float*, float*, void*) {
    tb2::set_error("tb2_orca_simulate: not built yet");
    return TB2_ERR_UNSUPPORTED;
}

}
