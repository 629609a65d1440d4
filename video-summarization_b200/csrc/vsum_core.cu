// Error text, ABI version and launch accounting for libvsum_b200.
#include "vsum_common.cuh"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace vsum {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

char *error_buffer() { return g_err; }

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// SM partition between the evaluation stage and the scorer when they overlap on two streams.
static std::atomic<int> g_eval_sm_budget{0}, g_scorer_sm_reserve{0};
int eval_sm_budget() { return g_eval_sm_budget.load(std::memory_order_relaxed); }
int scorer_sm_reserve() { return g_scorer_sm_reserve.load(std::memory_order_relaxed); }

// Attention forward kernel: 2 = persistent two-query-tile kernel (vsum_attn2_tc05.cu, default), 1 = one 128-query tile
// per CTA (vsum_attn_tc05.cu).  VSUM_ATTN_KERNEL in the environment overrides the default at first use.
static std::atomic<int> g_attn_kernel{0};
int attention_kernel_version() {
    int v = g_attn_kernel.load(std::memory_order_relaxed);
    if (v == 0) {
        const char *e = getenv("VSUM_ATTN_KERNEL");
        v = (e && e[0] == '1') ? 1 : ((e && e[0] == '3') ? 3 : 2);
        g_attn_kernel.store(v, std::memory_order_relaxed);
    }
    return v;
}

// Feed-forward block: 2 = fused kernel (vsum_ffn_tc05.cu, default), 1 = fc1 and fc2 + LayerNorm as two GEMM launches.
// VSUM_FFN_KERNEL in the environment overrides the default at first use.
static std::atomic<int> g_ffn_kernel{0};
int ffn_kernel_version() {
    int v = g_ffn_kernel.load(std::memory_order_relaxed);
    if (v == 0) {
        const char *e = getenv("VSUM_FFN_KERNEL");
        v = (e && e[0] == '1') ? 1 : 2;
        g_ffn_kernel.store(v, std::memory_order_relaxed);
    }
    return v;
}

// Work counters of the persistent kernels' dynamic tile schedulers: every launch takes the next 64-byte slot (16 ints,
// zero on entry; the last CTA of the launch zeroes it again) of a per-device ring.  A slot comes round again after 4096
// launches, long after the launch that used it has retired.
int *sched_slot() {
    constexpr int SLOTS = 4096;
    static int *ring[64] = {};
    static std::atomic<unsigned> next[64];
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    dev &= 63;
    if (!ring[dev]) {
        std::lock_guard<std::mutex> lk(mu);
        if (!ring[dev]) {
            int *p = nullptr;
            if (cudaMalloc(&p, SLOTS * 64) != cudaSuccess || cudaMemset(p, 0, SLOTS * 64) != cudaSuccess) return nullptr;
            ring[dev] = p;
        }
    }
    return ring[dev] + 16 * (next[dev].fetch_add(1, std::memory_order_relaxed) % SLOTS);
}

// ---- profiling -------------------------------------------------------------------------------
struct ProfRecord { int cat; cudaEvent_t a, b; };
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfRecord> g_prof_recs;      // event pool, reused across sessions
static size_t g_prof_used = 0;

ProfScope::ProfScope(int category, cudaStream_t s) : idx(-1), stream(s) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (g_prof_used == g_prof_recs.size()) {
        ProfRecord r{};
        if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
        g_prof_recs.push_back(r);
    }
    idx = (int)g_prof_used++;
    g_prof_recs[idx].cat = category;
    cudaEventRecord(g_prof_recs[idx].a, stream);
}
ProfScope::~ProfScope() {
    if (idx >= 0) cudaEventRecord(g_prof_recs[idx].b, stream);
}

static const char *kProfNames[PROF_NUM] = {"embed_gemm", "qkv_gemm", "attention", "oproj_ln_gemm", "fc1_gemm",
                                           "fc2_ln_gemm", "ffn_fused", "shot_mean", "knapsack", "summary_mask", "overlap",
                                           "fscore_finalize", "other"};

}  // namespace vsum

extern "C" int vsum_profile_begin(void) {
    std::lock_guard<std::mutex> lk(vsum::g_prof_mu);
    vsum::g_prof_used = 0;
    vsum::g_prof_on = true;
    return VSUM_OK;
}
extern "C" int vsum_profile_end(float *ms_out, int32_t *count_out, int32_t ncat) {
    std::lock_guard<std::mutex> lk(vsum::g_prof_mu);
    vsum::g_prof_on = false;
    for (int i = 0; i < ncat; ++i) { if (ms_out) ms_out[i] = 0.f; if (count_out) count_out[i] = 0; }
    for (size_t i = 0; i < vsum::g_prof_used; ++i) {
        const vsum::ProfRecord &r = vsum::g_prof_recs[i];
        float ms = 0.f;
        VSUM_CUDA_OK(cudaEventSynchronize(r.b));
        VSUM_CUDA_OK(cudaEventElapsedTime(&ms, r.a, r.b));
        if (r.cat < ncat) { if (ms_out) ms_out[r.cat] += ms; if (count_out) count_out[r.cat] += 1; }
    }
    vsum::g_prof_used = 0;
    return VSUM_OK;
}
extern "C" int32_t vsum_profile_num_categories(void) { return vsum::PROF_NUM; }
extern "C" const char *vsum_profile_category_name(int32_t i) {
    return (i >= 0 && i < vsum::PROF_NUM) ? vsum::kProfNames[i] : "";
}

extern "C" int vsum_set_sm_partition(int32_t eval_sms, int32_t scorer_reserved_sms) {
    VSUM_REQUIRE(eval_sms >= 0 && scorer_reserved_sms >= 0 && scorer_reserved_sms < 128, VSUM_EINVAL,
                 "vsum_set_sm_partition: eval_sms=%d scorer_reserved_sms=%d", eval_sms, scorer_reserved_sms);
    vsum::g_eval_sm_budget.store(eval_sms);
    vsum::g_scorer_sm_reserve.store(scorer_reserved_sms);
    return VSUM_OK;
}

extern "C" int vsum_set_attention_kernel(int32_t version) {
    VSUM_REQUIRE(version >= 1 && version <= 3, VSUM_EINVAL, "vsum_set_attention_kernel: version %d (1, 2 or 3)", version);
    vsum::g_attn_kernel.store(version);
    return VSUM_OK;
}

extern "C" int vsum_set_ffn_kernel(int32_t version) {
    VSUM_REQUIRE(version == 1 || version == 2, VSUM_EINVAL, "vsum_set_ffn_kernel: version %d (1 or 2)", version);
    vsum::g_ffn_kernel.store(version);
    return VSUM_OK;
}

extern "C" int vsum_abi_version(void) { return 1; }
extern "C" const char *vsum_last_error(void) { return vsum::error_buffer(); }
extern "C" int64_t vsum_launch_count(void) { return vsum::g_launches.load(std::memory_order_relaxed); }
