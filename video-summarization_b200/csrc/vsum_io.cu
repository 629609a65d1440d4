// Data layer of the hot path's input side (host only): a packed, memory-mapped dataset file and a
// multi-threaded, padding-free collate.  Replaces what src/data/dataset.py:64-168 does with h5py, Python
// lists and pad_sequence: a batch is gathered straight into ONE caller-provided (pinned) buffer as packed
// rows plus cu_seqlens, which is the layout vsum_scorer_forward consumes -- no [bs, Nmax, 1024] padding
// with the 1000 sentinel, no mask round trip (src/train.py:115-118).
//
// File layout (little endian), written by vsum_b200/data/packed.py:
//   header  64 B : magic "VSPACK01", u32 version, u32 n_videos, u64 index_offset, u64 file_bytes, u32 feature_dim,
//                  u32 feature_dtype (0 = float32 as in the h5 files, 1 = bfloat16)
//   arrays       : per video features f32|bf16 [N,dim] (4096-aligned), gtscore f32[N], picks i32[N], change_points
//                  i32[S,2], user_summary f32|u8 [U,n_frames], user_scores f32[U,n_frames], video_rep f32[rep_dim]
//   index        : n_videos entries of 256 B (name, sizes, byte offsets; offset 0 = array absent)
#include "vsum_common.cuh"

#include <algorithm>
#include <atomic>
#include <numeric>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <vector>

namespace {

struct FileHeader {
    char magic[8]; uint32_t version, n_videos; uint64_t index_offset, file_bytes; uint32_t feature_dim, feature_dtype, pad[6];
};
static_assert(sizeof(FileHeader) == 64, "header is 64 bytes");

struct IndexEntry {
    char name[96];
    int32_t n_steps, n_frames, n_shots, n_users, rep_dim, has_user_scores, user_summary_dtype, pad0;
    uint64_t off[VSUM_PACK_NUM_ARRAYS];
    uint8_t pad1[256 - 96 - 8 * 4 - 8 * VSUM_PACK_NUM_ARRAYS];
};
static_assert(sizeof(IndexEntry) == 256, "index entries are 256 bytes");

}  // namespace

struct vsum_pack {
    int fd = -1;
    const uint8_t *base = nullptr;
    size_t bytes = 0;
    const FileHeader *hdr = nullptr;
    const IndexEntry *index = nullptr;
    bool pinned = false;          // base is a cudaHostAlloc'd copy of the file (VSUM_PACK_PINNED), not a mapping
};

static uint64_t array_bytes(const FileHeader *h, const IndexEntry &e, int kind) {
    switch (kind) {
        case VSUM_PACK_FEATURES: return (uint64_t)e.n_steps * h->feature_dim * (h->feature_dtype == VSUM_FEATURES_BF16 ? 2 : 4);
        case VSUM_PACK_GTSCORE: return (uint64_t)e.n_steps * 4;
        case VSUM_PACK_PICKS: return (uint64_t)e.n_steps * 4;
        case VSUM_PACK_CHANGE_POINTS: return (uint64_t)e.n_shots * 8;
        case VSUM_PACK_USER_SUMMARY: return (uint64_t)e.n_users * e.n_frames * (e.user_summary_dtype == 1 ? 1 : 4);
        case VSUM_PACK_USER_SCORES: return e.has_user_scores ? (uint64_t)e.n_users * e.n_frames * 4 : 0;
        case VSUM_PACK_VIDEO_REP: return (uint64_t)e.rep_dim * 4;
    }
    return 0;
}

// Parallel pread of the whole file into `dst` (page-locked): 8 workers over 64 MB slices.
static bool read_file_into(int fd, uint8_t *dst, size_t bytes) {
    const size_t slice = (size_t)64 << 20;
    const size_t n_slices = (bytes + slice - 1) / slice;
    std::atomic<size_t> next{0};
    std::atomic<bool> ok{true};
    auto work = [&]() {
        for (size_t i = next.fetch_add(1); i < n_slices && ok.load(); i = next.fetch_add(1)) {
            size_t at = i * slice;
            const size_t end = std::min(bytes, at + slice);
            while (at < end) {
                const ssize_t got = pread(fd, dst + at, end - at, (off_t)at);
                if (got <= 0) { ok.store(false); return; }
                at += (size_t)got;
            }
        }
    };
    const int nt = (int)std::max<size_t>(1, std::min<size_t>(8, n_slices));
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto &th : pool) th.join();
    return ok.load();
}

extern "C" int vsum_pack_open_ex(const char *path, int32_t residency, vsum_pack_t *out) {
    VSUM_REQUIRE(path && out, VSUM_EINVAL, "vsum_pack_open: null argument");
    VSUM_REQUIRE(residency == VSUM_PACK_MMAP || residency == VSUM_PACK_PINNED, VSUM_EINVAL, "vsum_pack_open: residency %d", residency);
    *out = nullptr;
    const int fd = ::open(path, O_RDONLY);
    VSUM_REQUIRE(fd >= 0, VSUM_EINVAL, "vsum_pack_open: cannot open %s", path);
    struct stat st;
    if (fstat(fd, &st) != 0 || (size_t)st.st_size < sizeof(FileHeader)) {
        ::close(fd);
        return vsum::set_error(VSUM_EINVAL, "vsum_pack_open: %s is not a pack file (too small)", path);
    }
    void *m = nullptr;
    if (residency == VSUM_PACK_PINNED) {
        // the dataset lives in page-locked memory: batches leave through the copy engine without a host-side gather
        const cudaError_t e = cudaHostAlloc(&m, (size_t)st.st_size, cudaHostAllocPortable);
        if (e != cudaSuccess) {
            ::close(fd);
            return vsum::set_error(VSUM_ENOMEM, "vsum_pack_open: cudaHostAlloc of %zu bytes for %s failed: %s", (size_t)st.st_size, path,
                                   cudaGetErrorString(e));
        }
        if (!read_file_into(fd, (uint8_t *)m, (size_t)st.st_size)) {
            cudaFreeHost(m); ::close(fd);
            return vsum::set_error(VSUM_EINVAL, "vsum_pack_open: reading %s failed", path);
        }
    } else {
        m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_SHARED, fd, 0);
        if (m == MAP_FAILED) {
            ::close(fd);
            return vsum::set_error(VSUM_ENOMEM, "vsum_pack_open: mmap of %s failed", path);
        }
    }
    auto *p = new vsum_pack;
    p->fd = fd; p->base = (const uint8_t *)m; p->bytes = (size_t)st.st_size; p->pinned = residency == VSUM_PACK_PINNED;
    p->hdr = (const FileHeader *)m;
    auto fail = [&](const char *why) {
        if (p->pinned) cudaFreeHost(m); else munmap(m, p->bytes);
        ::close(fd); delete p;
        return vsum::set_error(VSUM_EINVAL, "vsum_pack_open: %s: %s", path, why);
    };
    if (memcmp(p->hdr->magic, "VSPACK01", 8) != 0 || p->hdr->version != 1) return fail("bad magic or version");
    if (p->hdr->file_bytes != p->bytes) return fail("truncated file (size differs from the header)");
    if (p->hdr->feature_dtype != VSUM_FEATURES_F32 && p->hdr->feature_dtype != VSUM_FEATURES_BF16) return fail("unknown feature dtype");
    if (p->hdr->feature_dim == 0 || p->hdr->feature_dim > (1u << 20)) return fail("feature_dim out of range");
    // subtraction form: none of these can wrap for a crafted header
    if (p->hdr->index_offset > p->bytes || (uint64_t)p->hdr->n_videos > (p->bytes - p->hdr->index_offset) / sizeof(IndexEntry))
        return fail("index outside the file");
    p->index = (const IndexEntry *)(p->base + p->hdr->index_offset);
    for (uint32_t i = 0; i < p->hdr->n_videos; ++i) {
        const IndexEntry &e = p->index[i];
        if (e.n_steps < 0 || e.n_frames < 0 || e.n_shots < 0 || e.n_users < 0 || e.rep_dim < 0) return fail("negative size in the index");
        // n_users * n_frames * 4 and n_steps * feature_dim * 4 stay below 2^63 with these bounds (feature_dim <= 2^20)
        if ((uint64_t)e.n_users * (uint64_t)e.n_frames > (1ull << 40) || (uint64_t)e.n_steps > (1ull << 31)) return fail("array size out of range");
        for (int k = 0; k < VSUM_PACK_NUM_ARRAYS; ++k) {
            const uint64_t nb = array_bytes(p->hdr, e, k);
            if (e.off[k] && (e.off[k] > p->hdr->index_offset || nb > p->hdr->index_offset - e.off[k])) return fail("array outside the data region");
        }
    }
    if (p->pinned) { ::close(fd); p->fd = -1; }
    else madvise(m, p->bytes, MADV_WILLNEED);
    *out = p;
    return VSUM_OK;
}

extern "C" int vsum_pack_open(const char *path, vsum_pack_t *out) { return vsum_pack_open_ex(path, VSUM_PACK_MMAP, out); }

extern "C" void vsum_pack_close(vsum_pack_t p) {
    if (!p) return;
    if (p->pinned) cudaFreeHost((void *)p->base);
    else munmap((void *)p->base, p->bytes);
    if (p->fd >= 0) ::close(p->fd);
    delete p;
}

extern "C" int32_t vsum_pack_residency(vsum_pack_t p) { return p && p->pinned ? VSUM_PACK_PINNED : VSUM_PACK_MMAP; }
extern "C" int32_t vsum_pack_num_videos(vsum_pack_t p) { return p ? (int32_t)p->hdr->n_videos : 0; }
extern "C" int32_t vsum_pack_feature_dim(vsum_pack_t p) { return p ? (int32_t)p->hdr->feature_dim : 0; }
extern "C" int32_t vsum_pack_feature_dtype(vsum_pack_t p) { return p ? (int32_t)p->hdr->feature_dtype : 0; }

extern "C" int vsum_pack_video_info(vsum_pack_t p, int32_t i, vsum_pack_info *out) {
    VSUM_REQUIRE(p && out && i >= 0 && (uint32_t)i < p->hdr->n_videos, VSUM_EINVAL, "vsum_pack_video_info: bad video index %d", i);
    const IndexEntry &e = p->index[i];
    memcpy(out->name, e.name, sizeof(out->name));
    out->name[sizeof(out->name) - 1] = 0;
    out->n_steps = e.n_steps; out->n_frames = e.n_frames; out->n_shots = e.n_shots; out->n_users = e.n_users;
    out->rep_dim = e.rep_dim; out->has_user_scores = e.has_user_scores; out->user_summary_dtype = e.user_summary_dtype;
    return VSUM_OK;
}

extern "C" int vsum_pack_array(vsum_pack_t p, int32_t i, int32_t kind, const void **ptr, uint64_t *bytes) {
    VSUM_REQUIRE(p && ptr && bytes && i >= 0 && (uint32_t)i < p->hdr->n_videos && kind >= 0 && kind < VSUM_PACK_NUM_ARRAYS, VSUM_EINVAL,
                 "vsum_pack_array: bad argument (video %d, kind %d)", i, kind);
    const IndexEntry &e = p->index[i];
    *ptr = e.off[kind] ? p->base + e.off[kind] : nullptr;
    *bytes = e.off[kind] ? array_bytes(p->hdr, e, kind) : 0;
    return VSUM_OK;
}

// Padding-free collate: features of the listed videos back to back into features_out [sum N, dim] (and their
// gtscore into gtscore_out [sum N] when non-NULL), cu_seqlens_out[k] = first row of the k-th listed video.
// The byte range is split evenly over `threads` workers (each memcpy also faults the mapped pages in).
extern "C" int vsum_pack_collate(vsum_pack_t p, const int32_t *ids, int32_t n, int32_t threads, void *features_out,
                                 float *gtscore_out, int32_t *cu_seqlens_out) {
    VSUM_REQUIRE(p && ids && n >= 0 && features_out && cu_seqlens_out, VSUM_EINVAL, "vsum_pack_collate: null argument");
    const uint64_t row = (uint64_t)p->hdr->feature_dim * (p->hdr->feature_dtype == VSUM_FEATURES_BF16 ? 2 : 4);
    std::vector<uint64_t> dst(n + 1, 0);
    cu_seqlens_out[0] = 0;
    for (int k = 0; k < n; ++k) {
        VSUM_REQUIRE(ids[k] >= 0 && (uint32_t)ids[k] < p->hdr->n_videos, VSUM_EINVAL, "vsum_pack_collate: bad video index %d", ids[k]);
        const IndexEntry &e = p->index[ids[k]];
        VSUM_REQUIRE(e.off[VSUM_PACK_FEATURES] != 0, VSUM_EINVAL, "vsum_pack_collate: video %d has no features", ids[k]);
        VSUM_REQUIRE(!gtscore_out || e.off[VSUM_PACK_GTSCORE] != 0, VSUM_EINVAL, "vsum_pack_collate: video %d has no gtscore", ids[k]);
        dst[k + 1] = dst[k] + (uint64_t)e.n_steps * row;
        VSUM_REQUIRE(dst[k + 1] / row < (1ull << 31), VSUM_EUNSUPPORTED, "vsum_pack_collate: batch exceeds 2^31 frames");
        cu_seqlens_out[k + 1] = (int32_t)(dst[k + 1] / row);
    }
    const uint64_t total = dst[n];
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(threads, (int64_t)(total >> 20) + 1));
    auto work = [&](int t) {
        const uint64_t lo = total * t / nt, hi = total * (t + 1) / nt;        // my byte range of the packed output
        int k = (int)(std::upper_bound(dst.begin(), dst.end(), lo) - dst.begin()) - 1;
        for (uint64_t at = lo; at < hi && k < n; ++k) {
            const IndexEntry &e = p->index[ids[k]];
            const uint64_t end = std::min(hi, dst[k + 1]);
            if (end > at) memcpy((uint8_t *)features_out + at, p->base + e.off[VSUM_PACK_FEATURES] + (at - dst[k]), end - at);
            at = std::max(at, end);
        }
        if (gtscore_out)                                                        // small: strided over videos
            for (int k2 = t; k2 < n; k2 += nt) {
                const IndexEntry &e = p->index[ids[k2]];
                memcpy(gtscore_out + cu_seqlens_out[k2], p->base + e.off[VSUM_PACK_GTSCORE], (size_t)e.n_steps * 4);
            }
    };
    if (nt == 1) { work(0); return VSUM_OK; }
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto &th : pool) th.join();
    return VSUM_OK;
}

// ---- evaluation batches ---------------------------------------------------------------------------------------
static const int kEvalClassWidth[] = {256, 1024, 4096, 9728, 18944, 28672};     // = kKnapsackClassWidth (vsum_eval.cu)

extern "C" int vsum_pack_eval_collate(vsum_pack_t p, const int32_t *ids, int32_t n, void *blob_host, size_t blob_bytes,
                                      vsum_eval_batch_layout *lay) {
    VSUM_REQUIRE(p && lay && n >= 0 && (ids || n == 0), VSUM_EINVAL, "vsum_pack_eval_collate: null argument");
    memset(lay, 0, sizeof(*lay));
    // packed order: longest video first, stable (pipeline.pack_videos)
    std::vector<int32_t> vid(ids, ids + n);
    for (int k = 0; k < n; ++k) {
        VSUM_REQUIRE(vid[k] >= 0 && (uint32_t)vid[k] < p->hdr->n_videos, VSUM_EINVAL, "vsum_pack_eval_collate: bad video index %d", vid[k]);
        const IndexEntry &e = p->index[vid[k]];
        VSUM_REQUIRE(e.off[VSUM_PACK_FEATURES] && e.off[VSUM_PACK_PICKS] && e.off[VSUM_PACK_CHANGE_POINTS] && e.n_shots > 0, VSUM_EINVAL,
                     "vsum_pack_eval_collate: video %d lacks features, picks or change points", vid[k]);
    }
    std::stable_sort(vid.begin(), vid.end(), [&](int32_t a, int32_t b) { return p->index[a].n_steps > p->index[b].n_steps; });
    int us_dtype = -1;
    bool users = n > 0;
    std::vector<int64_t> cap(n), words(n);
    std::vector<int32_t> last_end(n);
    int64_t T = 0, shots = 0, frames = 0, bitw = 0, us_elems = 0, n_users = 0;
    int32_t max_steps = 0, max_cap = 0;
    for (int k = 0; k < n; ++k) {
        const IndexEntry &e = p->index[vid[k]];
        const int32_t *cps = (const int32_t *)(p->base + e.off[VSUM_PACK_CHANGE_POINTS]);
        last_end[k] = cps[2 * (e.n_shots - 1) + 1];
        VSUM_REQUIRE(last_end[k] >= 0, VSUM_EINVAL, "vsum_pack_eval_collate: video %d ends at frame %d", vid[k], last_end[k]);
        cap[k] = (int64_t)(int)((double)(last_end[k] + 1) * 0.15);                 // generate_summary.py:45-46
        int width = -1;
        for (int w : kEvalClassWidth) if (cap[k] + 1 <= w) { width = w; break; }
        VSUM_REQUIRE(width > 0, VSUM_EUNSUPPORTED, "vsum_pack_eval_collate: knapsack capacity %lld exceeds the largest kernel class", (long long)cap[k]);
        words[k] = (int64_t)e.n_shots * (width / 32);
        T += e.n_steps; shots += e.n_shots; frames += last_end[k] + 1; bitw += words[k];
        max_steps = std::max(max_steps, e.n_steps); max_cap = std::max<int32_t>(max_cap, (int32_t)cap[k]);
        if (!e.off[VSUM_PACK_USER_SUMMARY]) users = false;
        else {
            if (us_dtype < 0) us_dtype = e.user_summary_dtype;
            VSUM_REQUIRE(us_dtype == e.user_summary_dtype, VSUM_EUNSUPPORTED, "vsum_pack_eval_collate: mixed user-summary dtypes in one batch");
            us_elems += (int64_t)e.n_users * e.n_frames; n_users += e.n_users;
        }
    }
    VSUM_REQUIRE(T < ((int64_t)1 << 31) && shots < ((int64_t)1 << 31), VSUM_EUNSUPPORTED, "vsum_pack_eval_collate: batch exceeds 2^31 frames");
    lay->B = n; lay->T = T; lay->total_picks = T; lay->total_shots = (int32_t)shots; lay->summary_frames = frames; lay->bit_words = bitw;
    lay->max_steps = max_steps; lay->max_cap = max_cap; lay->user_summary_dtype = users && us_dtype == 1 ? VSUM_USER_SUMMARY_U8 : VSUM_USER_SUMMARY_F32;
    lay->total_users = users ? (int32_t)n_users : 0; lay->us_elems = users ? us_elems : 0;
    size_t off = 0;
    auto place = [&](int64_t &field, size_t bytes) { off = (off + 255) / 256 * 256; field = (int64_t)off; off += bytes; };
    place(lay->off_video_ids, (size_t)n * 4); place(lay->off_cu_steps, (size_t)(n + 1) * 4); place(lay->off_picks, (size_t)T * 4);
    place(lay->off_cu_picks, (size_t)(n + 1) * 4); place(lay->off_n_frames, (size_t)n * 4); place(lay->off_cps, (size_t)shots * 8);
    place(lay->off_cu_shots, (size_t)(n + 1) * 4); place(lay->off_bit_offsets, (size_t)(n + 1) * 8); place(lay->off_order, (size_t)n * 4);
    place(lay->off_sum_offsets, (size_t)(n + 1) * 8); place(lay->off_us_offsets, (size_t)(n + 1) * 8); place(lay->off_cu_users, (size_t)(n + 1) * 4);
    place(lay->off_us_cols, (size_t)n * 4);
    lay->blob_bytes = (int64_t)((off + 255) / 256 * 256);
    // knapsack order: largest capacity first (stable), one launch per run of equal class width
    std::vector<int32_t> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return cap[a] > cap[b]; });
    auto width_of = [&](int k) { for (int w : kEvalClassWidth) if (cap[k] + 1 <= w) return w; return -1; };
    for (int pos = 0; pos < n;) {
        const int w = width_of(order[pos]);
        int end = pos;
        while (end < n && width_of(order[end]) == w) ++end;
        VSUM_REQUIRE(lay->n_launches < 8, VSUM_EUNSUPPORTED, "vsum_pack_eval_collate: more knapsack classes than launch slots");
        lay->launch_first[lay->n_launches] = pos; lay->launch_count[lay->n_launches] = end - pos;
        lay->launch_max_cap[lay->n_launches] = (int32_t)cap[order[pos]];
        ++lay->n_launches;
        pos = end;
    }
    if (!blob_host) return VSUM_OK;
    VSUM_REQUIRE(blob_bytes >= (size_t)lay->blob_bytes, VSUM_ENOMEM, "vsum_pack_eval_collate: blob %zu < %lld bytes", blob_bytes, (long long)lay->blob_bytes);
    uint8_t *b = (uint8_t *)blob_host;
    int32_t *o_vid = (int32_t *)(b + lay->off_video_ids), *o_cu = (int32_t *)(b + lay->off_cu_steps), *o_picks = (int32_t *)(b + lay->off_picks);
    int32_t *o_cup = (int32_t *)(b + lay->off_cu_picks), *o_nf = (int32_t *)(b + lay->off_n_frames), *o_cps = (int32_t *)(b + lay->off_cps);
    int32_t *o_cus = (int32_t *)(b + lay->off_cu_shots), *o_ord = (int32_t *)(b + lay->off_order), *o_cuu = (int32_t *)(b + lay->off_cu_users);
    int32_t *o_cols = (int32_t *)(b + lay->off_us_cols);
    int64_t *o_bit = (int64_t *)(b + lay->off_bit_offsets), *o_sum = (int64_t *)(b + lay->off_sum_offsets), *o_uso = (int64_t *)(b + lay->off_us_offsets);
    o_cu[0] = o_cup[0] = o_cus[0] = o_cuu[0] = 0; o_bit[0] = o_sum[0] = o_uso[0] = 0;
    for (int k = 0; k < n; ++k) {
        const IndexEntry &e = p->index[vid[k]];
        o_vid[k] = vid[k]; o_ord[k] = order[k]; o_nf[k] = e.n_frames; o_cols[k] = users ? e.n_frames : 0;
        memcpy(o_picks + o_cu[k], p->base + e.off[VSUM_PACK_PICKS], (size_t)e.n_steps * 4);
        memcpy(o_cps + 2 * (size_t)o_cus[k], p->base + e.off[VSUM_PACK_CHANGE_POINTS], (size_t)e.n_shots * 8);
        o_cu[k + 1] = o_cup[k + 1] = o_cu[k] + e.n_steps;
        o_cus[k + 1] = o_cus[k] + e.n_shots;
        o_bit[k + 1] = o_bit[k] + words[k];
        o_sum[k + 1] = o_sum[k] + last_end[k] + 1;
        o_uso[k + 1] = o_uso[k] + (users ? (int64_t)e.n_users * e.n_frames : 0);
        o_cuu[k + 1] = o_cuu[k] + (users ? e.n_users : 0);
    }
    return VSUM_OK;
}

extern "C" int vsum_pack_h2d(vsum_pack_t p, const void *blob_host, const vsum_eval_batch_layout *lay, void *features_dev,
                             void *user_summary_dev, void *stream) {
    VSUM_REQUIRE(p && blob_host && lay, VSUM_EINVAL, "vsum_pack_h2d: null argument");
    VSUM_REQUIRE(p->pinned, VSUM_EINVAL, "vsum_pack_h2d: the pack must be opened with VSUM_PACK_PINNED (DMA out of page-locked memory)");
    VSUM_REQUIRE(!user_summary_dev || lay->total_users > 0, VSUM_EINVAL, "vsum_pack_h2d: the batch has no user summaries");
    const uint8_t *b = (const uint8_t *)blob_host;
    const int32_t *vid = (const int32_t *)(b + lay->off_video_ids), *cu = (const int32_t *)(b + lay->off_cu_steps);
    const int64_t *uso = (const int64_t *)(b + lay->off_us_offsets);
    const uint64_t row = (uint64_t)p->hdr->feature_dim * (p->hdr->feature_dtype == VSUM_FEATURES_BF16 ? 2 : 4);
    const uint64_t us_elt = lay->user_summary_dtype == VSUM_USER_SUMMARY_U8 ? 1 : 4;
    cudaStream_t s = (cudaStream_t)stream;
    for (int k = 0; k < lay->B; ++k) {
        VSUM_REQUIRE(vid[k] >= 0 && (uint32_t)vid[k] < p->hdr->n_videos, VSUM_EINVAL, "vsum_pack_h2d: corrupt blob (video %d)", vid[k]);
        const IndexEntry &e = p->index[vid[k]];
        // the blob must be the one vsum_pack_eval_collate wrote for this pack: row and element offsets advance by this video's sizes
        VSUM_REQUIRE(cu[k + 1] - cu[k] == e.n_steps && (!user_summary_dev || uso[k + 1] - uso[k] == (int64_t)e.n_users * e.n_frames), VSUM_EINVAL,
                     "vsum_pack_h2d: blob and pack disagree on video %d", vid[k]);
        if (features_dev)
            VSUM_CUDA_OK(cudaMemcpyAsync((uint8_t *)features_dev + (uint64_t)cu[k] * row, p->base + e.off[VSUM_PACK_FEATURES],
                                         (uint64_t)e.n_steps * row, cudaMemcpyHostToDevice, s));
        if (user_summary_dev)
            VSUM_CUDA_OK(cudaMemcpyAsync((uint8_t *)user_summary_dev + (uint64_t)uso[k] * us_elt, p->base + e.off[VSUM_PACK_USER_SUMMARY],
                                         (uint64_t)e.n_users * e.n_frames * us_elt, cudaMemcpyHostToDevice, s));
    }
    return VSUM_OK;
}
