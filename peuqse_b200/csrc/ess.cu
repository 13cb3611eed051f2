// Ensemble slice sampling -- placeholder until the per-walker kernel lands.
#include "common.cuh"

extern "C" int64_t peuqse_ess_work_bytes(int64_t n_walkers, int32_t k) { return 256 + n_walkers * (int64_t)(k + 8) * 8 * 4; }

extern "C" int peuqse_ess_run(const peuqse_problem*, const peuqse_ess_spec*, double*, double*, double*, double*, double*, void*,
                              void*) {
    return pq_set_error(PEUQSE_ERR_INVALID, "peuqse_ess_run: not implemented yet");
}
