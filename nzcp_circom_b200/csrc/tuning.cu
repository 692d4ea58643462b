// nzcp_tuning_set: the process-wide experiment / test knobs of the library (documented in include/nzcp_prover.h).  Kept in
// its own small translation unit so that adding a knob does not rebuild the heavy kernel files.
#include "api_util.cuh"

namespace nzcp {
extern std::atomic<int> g_tune_rounds;                                                   // msm.cu
extern std::atomic<int> g_tune_pair_k[kMsmMaxRounds];
extern std::atomic<int> g_tune_rounds_w, g_tune_rounds_h;
extern std::atomic<int> g_tune_pair_prefetch[2], g_tune_acc_prefetch, g_tune_pair_stage, g_tune_gather_hint, g_tune_sort_threads;
extern std::atomic<int> g_tune_ntt_tma;                                                  // ntt.cu
extern std::atomic<int> g_tune_c_h, g_tune_c_w, g_tune_rounds_b2;                        // prover.cu
extern std::atomic<int> g_tune_stage_mode, g_tune_stage_chunk_kb, g_tune_stage_threads;
}  // namespace nzcp

using namespace nzcp;

extern "C" {

int nzcp_tuning_set(const char* name, int value) {
  return api_guard([&] {
    if (!name) throw ApiError(NZCP_E_ARG, "null argument");
    const std::string k(name);
    if (k == "msm_rounds") g_tune_rounds.store(value);
    else if (k == "prover_rounds_w") g_tune_rounds_w.store(value);
    else if (k == "prover_rounds_h") g_tune_rounds_h.store(value);
    else if (k == "pair_k1") g_tune_pair_k[0].store(value);
    else if (k == "pair_k2") g_tune_pair_k[1].store(value);
    else if (k == "pair_k3") g_tune_pair_k[2].store(value);
    else if (k == "pair_prefetch_fwd") g_tune_pair_prefetch[0].store(value);
    else if (k == "pair_prefetch_bwd") g_tune_pair_prefetch[1].store(value);
    else if (k == "acc_prefetch") g_tune_acc_prefetch.store(value);
    else if (k == "pair_stage") g_tune_pair_stage.store(value);
    else if (k == "gather_hint") g_tune_gather_hint.store(value);
    else if (k == "sort_threads") g_tune_sort_threads.store(value);
    else if (k == "stage_mode") g_tune_stage_mode.store(value);
    else if (k == "stage_chunk_kb") g_tune_stage_chunk_kb.store(value);
    else if (k == "stage_threads") g_tune_stage_threads.store(value);
    else if (k == "ntt_tma") g_tune_ntt_tma.store(value);
    else if (k == "prover_rounds_b2") g_tune_rounds_b2.store(value);
    else if (k == "prover_c_h") g_tune_c_h.store(value);
    else if (k == "prover_c_w") g_tune_c_w.store(value);
    else throw ApiError(NZCP_E_ARG, "unknown tuning knob: " + k);
  });
}

}  // extern "C"
