// Rank correlations between the predicted frame scores and every user's frame scores on the GPU:
// Kendall tau-b and Spearman rho exactly as src/evaluation/compute_correlation.py:4-15 obtains them from
// scipy (stats.kendalltau / stats.spearmanr on rankdata(-pred), rankdata(-user)), for a batch of videos.
//
// The prediction is piecewise constant over the picks segments (compute_metrics.py:19-39), so its
// ranking is a sort of <= 8192 segment scores per video (shared by all users).  Per (video, user) pair,
// one CTA works on its own slice of a global scratch area (L2 resident):
//   1. gather the user's scores in prediction-class order                  (stable secondary key)
//   2. LSD radix sort by the user's score (4 passes x 8 bits)  -> order (y, class)
//   3. tie groups of y by scans -> average ranks, Spearman sums, ytie, joint ties  (exact int64)
//   4. discordant pairs = inversions of the class sequence: MSD one-bit stable partitions, the group
//      boundaries of every pass come from the per-video class histogram
//   5. tau = (tot - xtie - ytie + ntie - 2 dis) / sqrt(tot - xtie) / sqrt(tot - ytie)   (scipy's
//      expression order: bit-exact), rho from exact centred integer rank sums.
#include "vsum_kernels.cuh"

namespace vsum {
namespace {

constexpr int CT = 1024;
constexpr int MAXSEG = 8192;
constexpr unsigned FULL = 0xffffffffu;

struct VideoStats { int n, A, nb, head; long long xtie, sxx; };

__device__ __forceinline__ uint32_t order_key(float v) {       // ascending uint32 order == ascending float order
    if (v == 0.f) v = 0.f;                                        // -0.0 ties with +0.0 (numeric equality)
    const uint32_t u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u ^ 0x80000000u);
}

// ---- block-wide scans (blockDim.x == 1024; wbuf: 32 entries of shared memory) ---------------
template <class T, class Op>
__device__ __forceinline__ T block_scan_incl(T v, Op op, T *wbuf, T &total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const T y = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc = op(y, inc);
    }
    __syncthreads();
    if (lane == 31) wbuf[w] = inc;
    __syncthreads();
    const T wt = wbuf[lane];
    T winc = wt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const T y = __shfl_up_sync(FULL, winc, o);
        if (lane >= o) winc = op(y, winc);
    }
    total = __shfl_sync(FULL, winc, 31);
    const T prev = __shfl_sync(FULL, winc, (w + 31) & 31);      // inclusive value of warp w-1
    return w == 0 ? inc : op(prev, inc);
}
struct OpAdd { template <class T> __device__ T operator()(T a, T b) const { return a + b; } };
struct OpMax { __device__ int operator()(int a, int b) const { return max(a, b); } };
struct OpMin { __device__ int operator()(int a, int b) const { return min(a, b); } };
// segmented sum on (flag << 32 | value): a flag restarts the sum
struct OpSeg {
    __device__ unsigned long long operator()(unsigned long long a, unsigned long long b) const {
        if (b >> 32) return b;
        return (a & 0xffffffff00000000ull) | (unsigned long long)(uint32_t)((uint32_t)a + (uint32_t)b);
    }
};

// Exclusive block scan of segmented-sum items (0 is the identity of OpSeg); `total` = combination of all items.
__device__ __forceinline__ unsigned long long block_scan_seg_excl(unsigned long long v, unsigned long long *wbuf, unsigned long long &total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    OpSeg op;
    unsigned long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long y = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc = op(y, inc);
    }
    unsigned long long exc = __shfl_up_sync(FULL, inc, 1);
    if (lane == 0) exc = 0;
    __syncthreads();
    if (lane == 31) wbuf[w] = inc;
    __syncthreads();
    const unsigned long long wt = wbuf[lane];
    unsigned long long winc = wt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long y = __shfl_up_sync(FULL, winc, o);
        if (lane >= o) winc = op(y, winc);
    }
    total = __shfl_sync(FULL, winc, 31);
    const unsigned long long prev = __shfl_sync(FULL, winc, (w + 31) & 31);      // inclusive value of warp w-1
    return w == 0 ? exc : op(prev, exc);
}

__device__ __forceinline__ long long block_sum_ll(long long v, long long *wbuf) {
    long long t;
    block_scan_incl(v, OpAdd(), wbuf, t);
    return t;
}

// ---- kernel 1: per video, rank classes of the segment scores --------------------------------
// Segment k < N covers frames [picks[k], picks[k+1]) (the last one up to n_frames) with value scores[k];
// frames before picks[0] keep the 0 of np.zeros (compute_metrics.py:29): pseudo-segment N when picks[0] > 0.
__global__ void __launch_bounds__(CT)
corr_classes_kernel(const float *__restrict__ scores, const int32_t *__restrict__ cu_steps, const int32_t *__restrict__ picks,
                    const int32_t *__restrict__ n_frames, uint16_t *__restrict__ seg_class, int32_t *__restrict__ seg_dest,
                    int32_t *__restrict__ class_start, int32_t *__restrict__ class_dx2, VideoStats *__restrict__ vstats) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ long long wbuf64[32];
    __shared__ int wbuf[32];
    const int v = blockIdx.x, base = cu_steps[v], N = cu_steps[v + 1] - base, n = n_frames[v];
    const int head = (N > 0 && picks[base] > 0) ? 1 : 0;
    const int nseg = N + head;
    int P = 1;
    while (P < nseg) P <<= 1;
    unsigned long long *kv = reinterpret_cast<unsigned long long *>(smem_raw);         // [P] (key << 32 | segment)
    int *cls_of = reinterpret_cast<int *>(kv + P);                                     // [P] class of sorted position
    int *off_of = cls_of + P;                                                          // [P] first frame (class order)
    auto seg_len = [&](int k) -> int {
        if (k == N) return picks[base];                                                // head pseudo-segment
        const int lo = picks[base + k], hi = (k + 1 < N) ? picks[base + k + 1] : n;
        return max(hi - lo, 0);
    };
    for (int i = threadIdx.x; i < P; i += CT) {
        unsigned long long e = ~0ull;
        if (i < nseg) e = ((unsigned long long)order_key(i < N ? scores[base + i] : 0.f) << 32) | (unsigned)i;
        kv[i] = e;
    }
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < P; i += CT) {
                const int x = i ^ j;
                if (x > i) {
                    const unsigned long long a = kv[i], b = kv[x];
                    if (((i & k) == 0) ? (a > b) : (a < b)) { kv[i] = b; kv[x] = a; }
                }
            }
            __syncthreads();
        }
    // class ids (dense, ascending score) and frame offsets in class order: blocked scans over the sorted list
    const int per = (P + CT - 1) / CT, i0 = threadIdx.x * per;
    int fl = 0, ln = 0;
    for (int i = i0; i < min(i0 + per, nseg); ++i) {
        fl += (i == 0 || (kv[i] >> 32) != (kv[i - 1] >> 32)) ? 1 : 0;
        ln += seg_len((int)(uint32_t)kv[i]);
    }
    int tot_f, tot_l;
    const int ef = block_scan_incl(fl, OpAdd(), wbuf, tot_f) - fl;
    const int el = block_scan_incl(ln, OpAdd(), wbuf, tot_l) - ln;
    {
        int c = ef, o = el;
        for (int i = i0; i < min(i0 + per, nseg); ++i) {
            c += (i == 0 || (kv[i] >> 32) != (kv[i - 1] >> 32)) ? 1 : 0;
            cls_of[i] = c - 1;
            off_of[i] = o;
            o += seg_len((int)(uint32_t)kv[i]);
        }
    }
    __syncthreads();
    const int A = tot_f;
    uint16_t *sc = seg_class + base + v;
    int32_t *sd = seg_dest + base + v;
    int32_t *cs = class_start + base + 2 * v, *cd = class_dx2 + base + 2 * v;
    long long xtie = 0, sxx = 0;
    for (int i = threadIdx.x; i < nseg; i += CT) {
        const int k = (int)(uint32_t)kv[i];
        sc[k] = (uint16_t)cls_of[i];
        sd[k] = off_of[i];
        if (i == 0 || cls_of[i] != cls_of[i - 1]) {            // first segment of a class: find its frame count
            int e = i + 1;
            while (e < nseg && cls_of[e] == cls_of[i]) ++e;
            const int start = off_of[i], cnt = (e < nseg ? off_of[e] : tot_l) - start;
            cs[cls_of[i]] = start;
            const int dx2 = 2 * start + cnt - tot_l;             // 2 * average rank - (n + 1)
            cd[cls_of[i]] = dx2;
            xtie += (long long)cnt * (cnt - 1) / 2;
            sxx += (long long)cnt * dx2 * dx2;
        }
    }
    if (threadIdx.x == 0) cs[A] = tot_l;
    xtie = block_sum_ll(xtie, wbuf64);
    sxx = block_sum_ll(sxx, wbuf64);
    if (threadIdx.x == 0) {
        int nb = 0;
        while ((1 << nb) < A) ++nb;
        vstats[v] = VideoStats{tot_l, A, nb, head, xtie, sxx};
    }
}

// ---- kernel 2: one CTA per (video, user) -----------------------------------------------------
struct PairArgs {
    const float *user_scores; const int64_t *us_offsets; const int32_t *cu_users; const int32_t *us_cols;
    const int32_t *cu_steps; const int32_t *picks;
    const uint16_t *seg_class; const int32_t *seg_dest; const int32_t *class_start; const int32_t *class_dx2;
    const VideoStats *vstats;
    uint32_t *keyA, *keyB; uint16_t *clsA, *clsB; int32_t *sbuf;
    double *tau_out, *rho_out; int B;
};

__global__ void __launch_bounds__(CT, 1)
corr_pair_kernel(PairArgs a) {
    __shared__ int wbuf[32];
    __shared__ long long wbuf64[32];
    __shared__ unsigned long long wbufu[32];
    __shared__ int hist[256], bin_base[256], tile_cnt[256];
    __shared__ uint16_t wcnt[32][256], wpre[32][256];
    __shared__ int s_carry;
    __shared__ unsigned long long s_carry_seg;
    const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int v = find_segment(a.cu_users, a.B, p), u = p - a.cu_users[v];
    const VideoStats st = a.vstats[v];
    const int n = st.n;
    double *tau_out = a.tau_out + p, *rho_out = a.rho_out + p;
    if (a.us_cols[v] != n || n < 2) {                           // scipy: length mismatch raises / size < 2 -> nan
        if (tid == 0) { *tau_out = __longlong_as_double(0x7ff8000000000000ll); *rho_out = *tau_out; }
        return;
    }
    const int64_t off = a.us_offsets[v] + (int64_t)u * n;
    const float *y = a.user_scores + off;
    uint32_t *kA = a.keyA + off, *kB = a.keyB + off;
    uint16_t *cA = a.clsA + off, *cB = a.clsB + off;
    int32_t *sb = a.sbuf + off;
    const int base = a.cu_steps[v], N = a.cu_steps[v + 1] - base;
    const int32_t *pk = a.picks + base;
    const uint16_t *sc = a.seg_class + base + v;
    const int32_t *sd = a.seg_dest + base + v;
    const int32_t *cs = a.class_start + base + 2 * v, *cd = a.class_dx2 + base + 2 * v;

    // 1. gather in class order
    for (int f = tid; f < n; f += CT) {
        int k, begin;
        if (f < pk[0]) { k = N; begin = 0; }
        else { k = find_segment(pk, N, f); begin = pk[k]; }     // pk ascending; last segment runs to n
        const int j = sd[k] + (f - begin);
        kA[j] = order_key(y[f]);
        cA[j] = sc[k];
    }
    __syncthreads();

    // 2. LSD radix sort by key, stable, 8 bits per pass
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = pass * 8;
        const uint32_t *ks = (pass & 1) ? kB : kA; uint32_t *kd = (pass & 1) ? kA : kB;
        const uint16_t *csrc = (pass & 1) ? cB : cA; uint16_t *cdst = (pass & 1) ? cA : cB;
        if (tid < 256) hist[tid] = 0;
        for (int i = tid; i < 32 * 256; i += CT) (&wcnt[0][0])[i] = 0;
        __syncthreads();
        for (int j = tid; j < n; j += CT) atomicAdd(&hist[(ks[j] >> shift) & 255], 1);
        __syncthreads();
        {   // exclusive scan of the 256 bins (threads >= 256 contribute 0)
            int tot;
            const int hv = tid < 256 ? hist[tid] : 0;
            const int ex = block_scan_incl(hv, OpAdd(), wbuf, tot) - hv;
            if (tid < 256) bin_base[tid] = ex;
        }
        __syncthreads();
        for (int t0 = 0; t0 < n; t0 += CT) {
            const int j = t0 + tid;
            const bool act = j < n;
            uint32_t key = 0; uint16_t c = 0; int d = 0;
            if (act) { key = ks[j]; c = csrc[j]; d = (key >> shift) & 255; }
            const unsigned amask = __ballot_sync(FULL, act);
            int rank = 0;
            if (act) {
                const unsigned peers = __match_any_sync(amask, d);
                rank = __popc(peers & ((1u << lane) - 1));
                if (rank == 0) wcnt[warp][d] = (uint16_t)__popc(peers);
            }
            __syncthreads();
            if (tid < 256) {
                int run = 0;
#pragma unroll 8
                for (int w = 0; w < 32; ++w) { const int cnt = wcnt[w][tid]; wpre[w][tid] = (uint16_t)run; wcnt[w][tid] = 0; run += cnt; }
                tile_cnt[tid] = run;
            }
            __syncthreads();
            if (act) {
                const int pos = bin_base[d] + wpre[warp][d] + rank;
                kd[pos] = key; cdst[pos] = c;
            }
            __syncthreads();
            if (tid < 256) bin_base[tid] += tile_cnt[tid];
        }
        __syncthreads();
    }
    // sorted by (y, class): keys in kA, classes in cA (4 passes end in the A buffers)

    // 3a. forward sweep: start index of every y tie group (-> sb), ytie, joint ties
    long long ytie = 0, ntie = 0;
    if (tid == 0) s_carry = 0;
    __shared__ int s_carry2;
    if (tid == 0) s_carry2 = 0;
    __syncthreads();
    for (int t0 = 0; t0 < n; t0 += CT) {
        const int j = t0 + tid;
        int fs = -1, fj = -1;
        if (j < n) {
            const uint32_t kj = kA[j];
            const bool ny = j == 0 || kj != kA[j - 1];
            if (ny) fs = j;
            if (ny || cA[j] != cA[j - 1]) fj = j;
        }
        int tot;
        const int c1 = s_carry, c2 = s_carry2;
        int s = block_scan_incl(fs, OpMax(), wbuf, tot);
        const int t1 = tot;
        int sj = block_scan_incl(fj, OpMax(), wbuf, tot);
        s = max(s, c1); sj = max(sj, c2);
        if (j < n) { sb[j] = s; ytie += j - s; ntie += j - sj; }
        __syncthreads();
        if (tid == 0) { s_carry = max(c1, t1); s_carry2 = max(c2, tot); }
        __syncthreads();
    }
    // 3b. backward sweep: end of every tie group -> centred double ranks, Spearman sums
    long long sxy = 0, syy = 0;
    if (tid == 0) s_carry = n;
    __syncthreads();
    for (int t1 = n; t1 > 0; t1 -= CT) {
        const int j = t1 - 1 - tid;                              // thread order = descending index
        int fe = 0x7fffffff;
        if (j >= 0 && (j == n - 1 || kA[j] != kA[j + 1])) fe = j + 1;
        int tot;
        const int c1 = s_carry;
        const int e = min(block_scan_incl(fe, OpMin(), wbuf, tot), c1);
        if (j >= 0) {
            const long long dy2 = (long long)sb[j] + e - n;
            sxy += dy2 * cd[cA[j]];
            syy += dy2 * dy2;
        }
        __syncthreads();
        if (tid == 0) s_carry = min(c1, tot);
        __syncthreads();
    }

    // 4. inversions of the class sequence: MSD one-bit stable partitions, four consecutive elements per thread
    //    (tiles of 4096) and the class histogram in shared memory -- the group boundaries of every pass come from it
    long long dis = 0;
    const int A = st.A;
    extern __shared__ int s_cs[];                                  // class_start[0..A]
    for (int i = tid; i <= A; i += CT) s_cs[i] = cs[i];
    __syncthreads();
    auto cstart = [&](int c) -> int { return c >= A ? n : s_cs[c]; };
    for (int bit = st.nb - 1, it = 0; bit >= 0; --bit, ++it) {
        const uint16_t *src = (it & 1) ? cB : cA; uint16_t *dst = (it & 1) ? cA : cB;
        if (tid == 0) s_carry_seg = 0;
        __syncthreads();
        for (int t0 = 0; t0 < n; t0 += 4 * CT) {
            const int jb = t0 + tid * 4;
            int c[4], gs[4];
            unsigned long long pre[4], agg = 0;
            OpSeg op;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                unsigned long long item = 0;
                c[e] = 0; gs[e] = 0;
                if (jb + e < n) {
                    c[e] = src[jb + e];
                    gs[e] = cstart((c[e] >> (bit + 1)) << (bit + 1));
                    item = ((unsigned long long)(jb + e == gs[e] ? 1 : 0) << 32) | (unsigned)((c[e] >> bit) & 1);
                }
                pre[e] = agg;                                      // my items before e
                agg = op(agg, item);
            }
            unsigned long long tot;
            const unsigned long long carry = s_carry_seg;
            const unsigned long long before = op(carry, block_scan_seg_excl(agg, wbufu, tot));   // everything before my first item
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (jb + e >= n) continue;
                const int j = jb + e, one = (c[e] >> bit) & 1;
                const int ones_before = (j == gs[e]) ? 0 : (int)(uint32_t)op(before, pre[e]);   // within the group
                const int g0 = (c[e] >> (bit + 1)) << (bit + 1);
                const int zeros_total = cstart(g0 + (1 << bit)) - gs[e];
                int pos;
                if (one) pos = gs[e] + zeros_total + ones_before;
                else { pos = gs[e] + (j - gs[e] - ones_before); dis += ones_before; }
                dst[pos] = (uint16_t)c[e];
            }
            __syncthreads();
            if (tid == 0) s_carry_seg = op(carry, tot) & 0x00000000ffffffffull;   // flag consumed, keep the count
            __syncthreads();
        }
    }
    ytie = block_sum_ll(ytie, wbuf64);
    ntie = block_sum_ll(ntie, wbuf64);
    sxy = block_sum_ll(sxy, wbuf64);
    syy = block_sum_ll(syy, wbuf64);
    dis = block_sum_ll(dis, wbuf64);
    if (tid == 0) {
        const double nan = __longlong_as_double(0x7ff8000000000000ll);
        const long long tot = (long long)n * (n - 1) / 2;
        double tau = nan, rho = nan;
        if (st.xtie != tot && ytie != tot) {
            const long long cmd = tot - st.xtie - ytie + ntie - 2 * dis;
            tau = (double)cmd / sqrt((double)(tot - st.xtie)) / sqrt((double)(tot - ytie));
            tau = fmin(1.0, fmax(-1.0, tau));
        }
        if (st.sxx > 0 && syy > 0) {
            rho = (double)sxy / (sqrt((double)st.sxx) * sqrt((double)syy));
            rho = fmin(1.0, fmax(-1.0, rho));
        }
        *tau_out = tau; *rho_out = rho;
    }
}

// mean over the users of every video, in user order: sum(kendal) / len(kendal)   (compute_correlation.py:15)
__global__ void corr_finalize_kernel(const double *__restrict__ tau, const double *__restrict__ rho, const int32_t *__restrict__ cu_users,
                                     int B, double *__restrict__ kendall_out, double *__restrict__ spearman_out) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= B) return;
    const int u0 = cu_users[v], u1 = cu_users[v + 1];
    double a = 0.0, b = 0.0;
    for (int u = u0; u < u1; ++u) { a += tau[u]; b += rho[u]; }
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    kendall_out[v] = u1 > u0 ? a / (double)(u1 - u0) : nan;
    spearman_out[v] = u1 > u0 ? b / (double)(u1 - u0) : nan;
}

struct CorrWs {
    uint16_t *seg_class; int32_t *seg_dest, *class_start, *class_dx2; VideoStats *vstats;
    uint32_t *keyA, *keyB; uint16_t *clsA, *clsB; int32_t *sbuf; double *tau, *rho;
};
size_t carve_corr(int64_t total_elems, int64_t T, int B, int total_users, void *base, CorrWs &w) {
    Carver k{(uint8_t *)base};
    w.seg_class = k.get<uint16_t>(T + B); w.seg_dest = k.get<int32_t>(T + B);
    w.class_start = k.get<int32_t>(T + 2 * (size_t)B); w.class_dx2 = k.get<int32_t>(T + 2 * (size_t)B);
    w.vstats = k.get<VideoStats>(B);
    w.keyA = k.get<uint32_t>(total_elems); w.keyB = k.get<uint32_t>(total_elems);
    w.clsA = k.get<uint16_t>(total_elems); w.clsB = k.get<uint16_t>(total_elems);
    w.sbuf = k.get<int32_t>(total_elems);
    w.tau = k.get<double>(total_users); w.rho = k.get<double>(total_users);
    return align_up(k.off, 1024);
}

}  // namespace
}  // namespace vsum

using namespace vsum;

extern "C" size_t vsum_rank_correlation_workspace_bytes(int64_t total_user_elems, int64_t T, int32_t B, int32_t total_users) {
    if (total_user_elems < 0 || T < 0 || B <= 0 || total_users < 0) return 0;
    CorrWs w;
    return carve_corr(total_user_elems, T, B, total_users, nullptr, w);
}

extern "C" int vsum_rank_correlation(const float *scores, const int32_t *cu_steps, const int32_t *picks, const int32_t *n_frames,
                                     const float *user_scores, const int64_t *us_offsets, const int32_t *cu_users,
                                     const int32_t *us_cols, int32_t B, int64_t T, int32_t max_steps, int32_t total_users,
                                     int64_t total_user_elems, void *workspace, size_t workspace_bytes, double *kendall_out,
                                     double *spearman_out, double *per_user_tau, double *per_user_rho, void *stream) {
    VSUM_REQUIRE(B >= 0 && T >= 0 && total_users >= 0 && total_user_elems >= 0, VSUM_EINVAL, "vsum_rank_correlation: negative sizes");
    if (B == 0) return VSUM_OK;
    VSUM_REQUIRE(scores && cu_steps && picks && n_frames && user_scores && us_offsets && cu_users && us_cols && workspace &&
                 kendall_out && spearman_out, VSUM_EINVAL, "vsum_rank_correlation: null pointer");
    VSUM_REQUIRE(max_steps + 1 <= MAXSEG, VSUM_EUNSUPPORTED, "vsum_rank_correlation: %d steps per video exceed %d", max_steps, MAXSEG - 1);
    VSUM_REQUIRE(((uintptr_t)workspace & 1023) == 0 && workspace_bytes >= vsum_rank_correlation_workspace_bytes(total_user_elems, T, B, total_users),
                 VSUM_ENOMEM, "vsum_rank_correlation: workspace unaligned or too small");
    cudaStream_t s = (cudaStream_t)stream;
    CorrWs w;
    carve_corr(total_user_elems, T, B, total_users, workspace, w);
    int P = 1;
    while (P < max_steps + 1) P <<= 1;
    const size_t smem = (size_t)P * 16;
    VSUM_ONCE_PER_DEVICE(VSUM_CUDA_OK(cudaFuncSetAttribute(corr_classes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MAXSEG * 16));
                         VSUM_CUDA_OK(cudaFuncSetAttribute(corr_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (MAXSEG + 2) * (int)sizeof(int))));
    {
        ProfScope prof(PROF_OTHER, s);
        corr_classes_kernel<<<B, CT, smem, s>>>(scores, cu_steps, picks, n_frames, w.seg_class, w.seg_dest, w.class_start, w.class_dx2, w.vstats);
        VSUM_LAUNCH_OK("corr_classes_kernel");
    }
    double *tau = per_user_tau ? per_user_tau : w.tau, *rho = per_user_rho ? per_user_rho : w.rho;
    if (total_users > 0) {
        PairArgs a{user_scores, us_offsets, cu_users, us_cols, cu_steps, picks, w.seg_class, w.seg_dest, w.class_start, w.class_dx2,
                   w.vstats, w.keyA, w.keyB, w.clsA, w.clsB, w.sbuf, tau, rho, B};
        ProfScope prof(PROF_OTHER, s);
        corr_pair_kernel<<<total_users, CT, (size_t)(max_steps + 2) * sizeof(int), s>>>(a);
        VSUM_LAUNCH_OK("corr_pair_kernel");
    }
    corr_finalize_kernel<<<(unsigned)ceil_div(B, 128), 128, 0, s>>>(tau, rho, cu_users, B, kendall_out, spearman_out);
    VSUM_LAUNCH_OK("corr_finalize_kernel");
    return VSUM_OK;
}
