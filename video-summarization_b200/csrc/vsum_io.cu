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

extern "C" int vsum_pack_open(const char *path, vsum_pack_t *out) {
    VSUM_REQUIRE(path && out, VSUM_EINVAL, "vsum_pack_open: null argument");
    *out = nullptr;
    const int fd = ::open(path, O_RDONLY);
    VSUM_REQUIRE(fd >= 0, VSUM_EINVAL, "vsum_pack_open: cannot open %s", path);
    struct stat st;
    if (fstat(fd, &st) != 0 || (size_t)st.st_size < sizeof(FileHeader)) {
        ::close(fd);
        return vsum::set_error(VSUM_EINVAL, "vsum_pack_open: %s is not a pack file (too small)", path);
    }
    void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_SHARED, fd, 0);
    if (m == MAP_FAILED) {
        ::close(fd);
        return vsum::set_error(VSUM_ENOMEM, "vsum_pack_open: mmap of %s failed", path);
    }
    auto *p = new vsum_pack;
    p->fd = fd; p->base = (const uint8_t *)m; p->bytes = (size_t)st.st_size;
    p->hdr = (const FileHeader *)m;
    auto fail = [&](const char *why) {
        munmap(m, p->bytes); ::close(fd); delete p;
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
    madvise(m, p->bytes, MADV_WILLNEED);
    *out = p;
    return VSUM_OK;
}

extern "C" void vsum_pack_close(vsum_pack_t p) {
    if (!p) return;
    munmap((void *)p->base, p->bytes);
    ::close(p->fd);
    delete p;
}

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
