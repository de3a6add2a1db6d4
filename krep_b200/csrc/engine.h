// engine.h — engine context and internal entry points shared by engine.cu / host_api.cu.
#pragma once
#include <mutex>
#include <vector>
#include "common.h"

namespace kb {

struct StageSlot
{
    uint8_t *buf = nullptr; // pinned
    cudaEvent_t ev = nullptr;
    bool in_flight = false;
};

struct Engine
{
    bool ready = false;
    int device = 0, sm_count = 0;
    cudaStream_t scan_stream = nullptr, copy_stream = nullptr;
    unsigned long long *d_counter = nullptr;
    uint64_t *h_counter = nullptr; // pinned
    uint64_t *d_keys[2] = {nullptr, nullptr};
    uint64_t key_cap = 0;
    void *d_sort_tmp = nullptr;
    size_t sort_tmp_bytes = 0;
    uint64_t *d_bounds = nullptr; // -c: line bounds, 2 per key
    uint64_t bounds_cap = 0;
    uint64_t *h_bounds = nullptr; // pinned
    uint64_t h_bounds_cap = 0;
    uint8_t *h_batch = nullptr; // pinned: texts of one krep_b200_search_batch call, packed
    uint64_t h_batch_cap = 0;
    // host-text entry points: device copy of the caller's buffer + pinned staging ring + key readback
    uint8_t *d_text = nullptr;
    uint64_t text_cap = 0;
    uint64_t *h_keys = nullptr;
    uint64_t h_keys_cap = 0;
    std::vector<StageSlot> stage;
    size_t stage_bytes = 0;
    std::vector<cudaEvent_t> ev_pool;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    std::vector<Plan *> plan_cache;
};

std::recursive_mutex &engine_mutex();
int engine_init(int device);
void engine_shutdown();
Plan *plan_build(const search_params_t *P, int algo, bool only_matching);
void plan_free(Plan *p);
int resolve_algo(const search_params_t *P, int algo); // host_api.cu: precondition fallbacks of the simd_* entries
int scan_shard(const Plan *plan, const krep_b200_shard_t *sh, int want_positions, cudaStream_t stream, ScanOut *out);
int ensure_keys(uint64_t cap);
int reset_counter(cudaStream_t stream);
int read_counter(cudaStream_t stream, uint64_t *count);
int sort_keys(uint64_t n, int end_bit, cudaStream_t stream, const uint64_t **sorted);
int key_end_bit(const Plan *plan, uint64_t max_offset);
void add_kernel_ms(float ms);
void reset_kernel_ms();

} // namespace kb
