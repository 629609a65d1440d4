// Evaluation half of the hot path on sm_100a: shot pooling, 0/1 knapsack, summary mask and
// keyshot F-score.  All integer / fp32 / fp64 arithmetic is ordered exactly as the reference's
// numpy + Python code orders it, so results are bit-identical given identical scores
// (DESIGN.md "Bit-exactness").  These stages are latency / shared-memory bound, not GEMMs.
#include "vsum_common.cuh"

#include <climits>

namespace vsum {

// ---------------------------------------------------------------------------------------------
// K10  shot pooling  (reference: src/evaluation/generate_summary.py:25-46)
// ---------------------------------------------------------------------------------------------
// The reference materialises frame_scores[n_frames] (each sub-sampled score repeated up to the
// next pick) and calls ndarray.mean() on each shot's slice.  numpy's float32 pairwise summation
// visits the slice strictly left to right, so one cursor over the pick list reproduces the
// upsampled sequence without ever writing it to memory.
struct FrameStream {
    const float *scores;
    const int32_t *picks;
    int n_scores, n_picks, n_pos, n_frames;
    int f, seg, seg_end;
    float v;

    __device__ __forceinline__ int pos(int i) const { return i < n_picks ? __ldg(picks + i) : n_frames; }
    __device__ __forceinline__ void load_segment() {
        if (seg < 0) { v = 0.0f; seg_end = pos(0); }
        else if (seg >= n_pos - 1) { v = 0.0f; seg_end = INT_MAX; }       // past the last position
        else { v = (seg < n_scores) ? __ldg(scores + seg) : 0.0f; seg_end = pos(seg + 1); }  // line 32-33
    }
    __device__ void seek(int frame) {
        f = frame;
        int lo = -1, hi = n_pos;              // pos(lo) <= frame < pos(hi)
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (pos(mid) <= frame) lo = mid; else hi = mid;
        }
        seg = lo;
        load_segment();
    }
    __device__ __forceinline__ float next() {
        while (f >= seg_end) { ++seg; load_segment(); }
        ++f;
        return v;
    }
};

// numpy pairwise_sum leaf (n <= 128): 8 accumulators, fixed combine tree, sequential tail.
__device__ float leaf_sum(FrameStream &st, int n) {
    if (n < 8) {
        float r = -0.0f;
        for (int i = 0; i < n; ++i) r = __fadd_rn(r, st.next());
        return r;
    }
    float r0 = st.next(), r1 = st.next(), r2 = st.next(), r3 = st.next();
    float r4 = st.next(), r5 = st.next(), r6 = st.next(), r7 = st.next();
    const int full = n - (n % 8);
    for (int i = 8; i < full; i += 8) {
        r0 = __fadd_rn(r0, st.next()); r1 = __fadd_rn(r1, st.next());
        r2 = __fadd_rn(r2, st.next()); r3 = __fadd_rn(r3, st.next());
        r4 = __fadd_rn(r4, st.next()); r5 = __fadd_rn(r5, st.next());
        r6 = __fadd_rn(r6, st.next()); r7 = __fadd_rn(r7, st.next());
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r0, r1), __fadd_rn(r2, r3)),
                          __fadd_rn(__fadd_rn(r4, r5), __fadd_rn(r6, r7)));
    for (int i = full; i < n; ++i) res = __fadd_rn(res, st.next());
    return res;
}

// Full pairwise recursion (split at n/2 rounded down to a multiple of 8) with an explicit stack.
__device__ float pairwise_sum_stream(FrameStream &st, int n) {
    float left_val[32];
    int right_n[32];
    bool have_left[32];
    int sp = 0, cur = n;
    float val;
    for (;;) {
        while (cur > 128) {
            int n2 = cur >> 1;
            n2 -= n2 & 7;
            right_n[sp] = cur - n2; have_left[sp] = false; ++sp;
            cur = n2;
        }
        val = leaf_sum(st, cur);
        bool descend = false;
        while (sp > 0) {
            if (!have_left[sp - 1]) {
                left_val[sp - 1] = val; have_left[sp - 1] = true;
                cur = right_n[sp - 1];
                descend = true;
                break;
            }
            val = __fadd_rn(left_val[sp - 1], val);
            --sp;
        }
        if (!descend) return val;
    }
}

__global__ void __launch_bounds__(128)
shot_mean_kernel(const float *__restrict__ scores, const int32_t *__restrict__ cu_steps,
                 const int32_t *__restrict__ picks, const int32_t *__restrict__ cu_picks,
                 const int32_t *__restrict__ n_frames, const int32_t *__restrict__ cps,
                 const int32_t *__restrict__ cu_shots, int B, int S_total,
                 double *__restrict__ val_out, int32_t *__restrict__ wt_out,
                 int32_t *__restrict__ cap_out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S_total) return;
    const int v = find_segment(cu_shots, B, s);
    const int lo = __ldg(cps + 2 * s), hi = __ldg(cps + 2 * s + 1);
    wt_out[s] = hi - lo + 1;                                           // line 41
    if (s == __ldg(cu_shots + v + 1) - 1)                              // line 45-46
        cap_out[v] = (int32_t)((double)(hi + 1) * 0.15);

    FrameStream st;
    st.scores = scores + __ldg(cu_steps + v);
    st.n_scores = __ldg(cu_steps + v + 1) - __ldg(cu_steps + v);
    st.picks = picks + __ldg(cu_picks + v);
    st.n_picks = __ldg(cu_picks + v + 1) - __ldg(cu_picks + v);
    st.n_frames = __ldg(n_frames + v);
    st.n_pos = st.n_picks + ((st.n_picks == 0 || st.pos(st.n_picks - 1) != st.n_frames) ? 1 : 0);  // line 29-30
    const int stop = min(hi + 1, st.n_frames);                         // numpy slices clamp
    const int n = stop - lo;
    if (n <= 0) { val_out[s] = __longlong_as_double(0x7ff8000000000000LL); return; }
    st.seek(lo);
    const float total = __fadd_rn(0.0f, pairwise_sum_stream(st, n));   // add.reduce starts at +0
    val_out[s] = (double)__fdiv_rn(total, (float)n);                   // float32 mean -> .item()
}

// ---------------------------------------------------------------------------------------------
// K11  0/1 knapsack  (reference: src/evaluation/knapsack_implementation.py:11-28)
// ---------------------------------------------------------------------------------------------
// One video per CTA.  Thread t owns the capacities w = t + k*THREADS and keeps K[i][w] for them in
// REGISTERS across shots; shared memory holds a copy of the previous row only so that other threads
// can read K[i-1][w - wt].  Per shot and capacity that is one 8-byte shared load and one 8-byte shared
// store (two barriers per shot; storing only the capacities that changed was measured and is slower:
// the bookkeeping costs more instructions than the stores it saves).  Capacities below the shot's
// weight cannot change and are skipped.  take[i][w] = (K[i][w] != K[i-1][w]) -- the reference's own
// back-track test -- is packed with a warp ballot into a bit matrix in global memory (it stays in
// L2); words whose capacities are all below the weight are never written and never read, because
// the back-track (warp 0, 32 rows per probe) only probes rows with wt <= w.

// One shot's update of the capacities a thread owns (cur[k] = K[i][k THREADS + tid]); KLO >= 0: the chunk that contains
// w_min is known at compile time, KLO < 0: every chunk tests its capacities.  Each warp's ballot word (take bits of 32
// consecutive capacities) leaves through a predicated store of lane 0 (inline PTX: no divergence bookkeeping).
template <int THREADS, int EPT, int KLO>
__device__ __forceinline__ void knapsack_row_update(double (&cur)[EPT], const double *__restrict__ rd, uint32_t *__restrict__ bw,
                                                    double vi, int w_min, int tid, int lane) {
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        if (KLO >= 0 && k < KLO) continue;
        const double b = cur[k];                          // K[i-1][w]
        bool take;
        double a;
        if (KLO >= 0 && k > KLO) {
            a = vi + rd[k * THREADS];
            take = !(b >= a);
        } else {
            const bool can = k * THREADS + tid >= w_min;
            a = vi + (can ? rd[k * THREADS] : 0.0);
            take = can && !(b >= a);
        }
        // Python: m = max(a, b) keeps a unless b > a; take = (m != b).  With take = !(b >= a)
        // and m = take ? a : b this is identical for every input incl. NaN (values equal when a == b).
        cur[k] = take ? a : b;
        const unsigned word = __ballot_sync(0xffffffffu, take);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %2, 0;\n\t@p st.global.u32 [%0], %1;\n\t}"
                     ::"l"(bw + k * (THREADS / 32)), "r"(word), "r"(lane) : "memory");
    }
}

template <int THREADS, int EPT>
__global__ void __launch_bounds__(THREADS, 1)
knapsack_kernel(const double *__restrict__ val, const int32_t *__restrict__ wt,
                const int32_t *__restrict__ cu_shots, const int32_t *__restrict__ cap,
                const int64_t *__restrict__ bit_offsets, const int32_t *__restrict__ order,
                uint32_t *__restrict__ take_bits, uint8_t *__restrict__ selected_out, int n_videos) {
    extern __shared__ double row[];                       // THREADS * EPT capacities
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int WORDS = THREADS * EPT / 32;
  for (int slot = blockIdx.x; slot < n_videos; slot += gridDim.x) {      // persistent: a bounded number of CTAs walks the list
    const int v = order ? __ldg(order + slot) : slot;
    const int s0 = __ldg(cu_shots + v), S = __ldg(cu_shots + v + 1) - s0;
    const int W = __ldg(cap + v);
    __syncthreads();                                      // previous video's back-track is done with row / bits
    for (int i = tid; i < S; i += THREADS) selected_out[s0 + i] = 0;
    if (W < 0 || S <= 0) continue;
    {   // every video of a launch must belong to this kernel's class (its bit rows are padded to it)
        constexpr int kW[] = {256, 1024, 4096, 9728, 18944, 28672};
        int own = -1;
#pragma unroll
        for (int c = 5; c >= 0; --c) if (W + 1 <= kW[c]) own = kW[c];
        if (own != THREADS * EPT) __trap();
    }
    // The row is padded to this kernel's full width (shared memory and bit matrix alike): padding
    // cells are computed like real ones -- the recurrence only looks at lower capacities, so they
    // cannot influence K[i][w] for w <= W -- which removes every bounds test from the inner loop.
    uint32_t *bits = take_bits + __ldg(bit_offsets + v);
    double cur[EPT];                                      // K[i][w] for my capacities
#pragma unroll
    for (int k = 0; k < EPT; ++k) { cur[k] = 0.0; row[k * THREADS + tid] = 0.0; }   // K[0][*] = 0
    __syncthreads();

    int wi_next = __ldg(wt + s0);
    double vi_next = __ldg(val + s0);
    for (int i = 0; i < S; ++i) {
        const int wi = wi_next;
        const double vi = vi_next;
        if (i + 1 < S) {                                  // next shot's (weight, value): hide the L2 round trip
            wi_next = __ldg(wt + s0 + i + 1);
            vi_next = __ldg(val + s0 + i + 1);
        }
        if (wi < 0 || wi > W) continue;                   // line 16: no capacity can hold it (block-uniform)
        const int w_min = wi > 1 ? wi : 1;                // K[i][0] stays 0 (line 14)
        const double *rd = row + tid - wi;                // rd[k*THREADS] = K[i-1][w - wt]
        uint32_t *bw = bits + (int64_t)i * WORDS + warp;  // this warp's ballot word of chunk k: bw[k*THREADS/32]
        // Chunk k holds the capacities [k THREADS, k THREADS + THREADS): chunks below the one that contains w_min cannot
        // change (no instruction at all: their words are never written and never read), chunks above it need no capacity
        // test, only chunk k_lo = w_min / THREADS tests `w >= w_min` per thread.  The row update was ISSUE-bound at ~15
        // instructions per capacity and warp (ncu: 52 % of the issue slots, shared memory at 12 %), most of them the
        // per-capacity test; k_lo is 0 or 1 for every real shot length, so those two cases are compiled as straight-line
        // code with the test in ONE chunk (LDS.64, DADD, DSETP, 2 FSEL, VOTE and the ballot word's predicated store per
        // capacity), anything longer takes the generic form (profiles/r02_eval_kernels_ncu.txt).
        const int k_lo = w_min / THREADS;
        if (k_lo == 0) knapsack_row_update<THREADS, EPT, 0>(cur, rd, bw, vi, w_min, tid, lane);
        else if (k_lo == 1) knapsack_row_update<THREADS, EPT, 1>(cur, rd, bw, vi, w_min, tid, lane);
        else knapsack_row_update<THREADS, EPT, -1>(cur, rd, bw, vi, w_min, tid, lane);
        __syncthreads();                                  // everyone has read the old row
#pragma unroll
        for (int k = 0; k < EPT; ++k) row[k * THREADS + tid] = cur[k];
        __syncthreads();
    }

    if (tid < 32) {                                                     // back-track, line 23-28
        int w = W, i = S;
        while (i > 0) {
            const int r = i - 1 - lane;
            bool bit = false;
            if (r >= 0) {
                const int wr = __ldg(wt + s0 + r);
                // rows whose weight exceeds w never wrote (or could set) their bit
                if (wr >= 0 && wr <= w)
                    bit = (__ldcg(bits + (int64_t)r * WORDS + (w >> 5)) >> (w & 31)) & 1u;
            }
            const unsigned m = __ballot_sync(0xffffffffu, bit);
            if (m == 0) { i -= 32; continue; }
            const int rsel = i - 1 - (__ffs(m) - 1);                    // highest row that took
            if (lane == 0) selected_out[s0 + rsel] = 1;
            w -= __ldg(wt + s0 + rsel);
            i = rsel;
        }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K12  summary mask + overlap / F-score
// (reference: generate_summary.py:51-53, evaluation_metrics.py:12-33)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
summary_mask_kernel(const uint8_t *__restrict__ selected, const int32_t *__restrict__ cps,
                    const int32_t *__restrict__ cu_shots, const int64_t *__restrict__ sum_offsets,
                    int8_t *__restrict__ summary_out) {
    const int v = blockIdx.y;
    const int s0 = __ldg(cu_shots + v), S = __ldg(cu_shots + v + 1) - s0;
    if (S <= 0) return;
    const int len = __ldg(cps + 2 * (s0 + S - 1) + 1) + 1;              // line 51
    int8_t *out = summary_out + __ldg(sum_offsets + v);
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < len; f += gridDim.x * blockDim.x) {
        // last shot whose start <= f (shots are ascending and disjoint)
        int lo = 0, hi = S;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (__ldg(cps + 2 * (s0 + mid)) <= f) lo = mid; else hi = mid;
        }
        const bool inside = f >= __ldg(cps + 2 * (s0 + lo)) && f <= __ldg(cps + 2 * (s0 + lo) + 1);
        out[f] = (inside && selected[s0 + lo]) ? 1 : 0;
    }
}

// Frame indices of the keyshot summary in ascending order -- what generate_summary_image.py:74-76 builds with
// `[i for i, is_in in enumerate(summary) if is_in == 1]` -- straight from the selected shots (they are
// ascending and disjoint), one CTA per video.  frames_out holds video v at [out_offsets[v], out_offsets[v+1]).
__global__ void __launch_bounds__(256)
summary_frames_kernel(const uint8_t *__restrict__ selected, const int32_t *__restrict__ cps, const int32_t *__restrict__ cu_shots,
                      const int64_t *__restrict__ out_offsets, int32_t *__restrict__ frames_out, int32_t *__restrict__ counts_out) {
    __shared__ int s_off[256], s_warp[8], s_carry;
    const int v = blockIdx.x, s0 = __ldg(cu_shots + v), S = __ldg(cu_shots + v + 1) - s0;
    int32_t *out = frames_out + __ldg(out_offsets + v);
    const int64_t room = __ldg(out_offsets + v + 1) - __ldg(out_offsets + v);
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int c0 = 0; c0 < S; c0 += 256) {
        const int k = c0 + threadIdx.x;
        int len = 0;
        if (k < S && selected[s0 + k]) len = max(__ldg(cps + 2 * (s0 + k) + 1) - __ldg(cps + 2 * (s0 + k)) + 1, 0);
        int inc = len;                                   // inclusive scan over the chunk
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, inc, o); if ((threadIdx.x & 31) >= o) inc += y; }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = inc;
        __syncthreads();
        int wbase = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wbase += s_warp[w];
        const int carry = s_carry;
        s_off[threadIdx.x] = carry + wbase + inc - len;
        __syncthreads();
        for (int i = 0; i < min(256, S - c0); ++i) {     // all threads write one shot's frames at a time
            if (!selected[s0 + c0 + i]) continue;
            const int start = __ldg(cps + 2 * (s0 + c0 + i)), n = __ldg(cps + 2 * (s0 + c0 + i) + 1) - start + 1, o = s_off[i];
            for (int f = threadIdx.x; f < n; f += 256)
                if (o + f < room) out[o + f] = start + f;
        }
        __syncthreads();
        if (threadIdx.x == 255) s_carry = carry + wbase + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) counts_out[v] = s_carry;
}

__device__ __forceinline__ long long block_sum_i64(long long x, long long *smem) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) smem[warp] = x;
    __syncthreads();
    x = (lane < nw) ? smem[lane] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// This thread's share of sum(S) over one video's int8 mask: bytes up to the first aligned word, then four
// frames per 32-bit load (signed byte sum = dp4a with 0x01010101), then the tail.
__device__ __forceinline__ long long mask_sum_partial(const int8_t *__restrict__ sm, int slen) {
    int head = (int)((4 - ((uintptr_t)sm & 3)) & 3);
    if (head > slen) head = slen;
    const int nw = (slen - head) >> 2;
    int acc = 0;                                                     // |sum| <= 128 * slen / blockDim: fits
    if ((int)threadIdx.x < head) acc += sm[threadIdx.x];
    const int *w = reinterpret_cast<const int *>(sm + head);
    for (int q = threadIdx.x; q < nw; q += blockDim.x) acc = __dp4a(__ldg(w + q), 0x01010101, acc);
    const int c = head + 4 * nw + threadIdx.x;
    if (c < slen) acc += sm[c];
    return (long long)acc;
}

// One CTA per (video, user) row: o = sum(S & G), g = sum(G), s = sum(S) as int64
// (evaluation_metrics.py:20-25).  The user row is streamed with 16-byte loads.
__global__ void __launch_bounds__(256)
overlap_kernel(const int8_t *__restrict__ summary, const int64_t *__restrict__ sum_offsets,
               const float *__restrict__ user_summary, const int64_t *__restrict__ us_offsets,
               const int32_t *__restrict__ cu_users, const int32_t *__restrict__ us_cols, int B,
               long long *__restrict__ counts, int total_rows) {
    __shared__ long long red[32];
  for (int rowid = blockIdx.x; rowid < total_rows; rowid += gridDim.x) {
    const int v = find_segment(cu_users, B, rowid);
    const int u = rowid - __ldg(cu_users + v);
    const int cols = __ldg(us_cols + v);
    const int slen = (int)(__ldg(sum_offsets + v + 1) - __ldg(sum_offsets + v));
    const int8_t *sm = summary + __ldg(sum_offsets + v);
    const float *g = user_summary + __ldg(us_offsets + v) + (int64_t)u * cols;
    long long o_cnt = 0, g_cnt = 0, s_cnt = 0;

    // head (until g is 16-byte aligned), vector body, tail
    int head = (int)(((16 - ((uintptr_t)g & 15)) & 15) >> 2);
    if (head > cols) head = cols;
    const int nvec = (cols - head) >> 2;
    for (int c = threadIdx.x; c < head; c += blockDim.x) {
        const long long gi = (long long)__ldg(g + c);
        const long long si = c < slen ? (long long)sm[c] : 0;
        o_cnt += si & gi; g_cnt += gi;
    }
    const float4 *g4 = reinterpret_cast<const float4 *>(g + head);
    // Vector body, two 16-byte loads in flight per thread.  The four mask bytes of a group come from one or two
    // aligned 32-bit words (funnel shift); user values that are exactly 0.0 / 1.0 (every dataset's user summaries)
    // are counted from their bit pattern -- the float -> int64 conversion of the general path is an XU-pipe
    // sequence that made this streaming kernel instruction-bound (33 % of HBM peak before).
    const uintptr_t sm_addr = (uintptr_t)(sm + head);
    const uint32_t *smw = reinterpret_cast<const uint32_t *>(sm_addr & ~(uintptr_t)3);
    const int sm_shift = (int)(sm_addr & 3) * 8;
    // groups whose 4 frames all lie inside the summary (minus one when unaligned: the funnel shift reads the next word)
    const int full_vec = max(0, min(nvec, (slen - head) >> 2) - (sm_shift != 0 ? 1 : 0));
    auto masks_of = [&](int q) -> uint32_t {                        // mask bytes of frames head+4q .. head+4q+3
        const uint32_t w0 = __ldg(smw + q);
        if (sm_shift == 0) return w0;
        return __funnelshift_r(w0, __ldg(smw + q + 1), sm_shift);
    };
    auto add_group = [&](const float4 &x, int q) {
        const uint32_t b0 = __float_as_uint(x.x), b1 = __float_as_uint(x.y), b2 = __float_as_uint(x.z), b3 = __float_as_uint(x.w);
        auto is01 = [](uint32_t b) { return (b << 1) == 0u || b == 0x3f800000u; };
        const int c = head + 4 * q;
        if (q < full_vec && is01(b0) && is01(b1) && is01(b2) && is01(b3)) {
            const uint32_t m = masks_of(q);                          // summary bytes are 0 / 1
            const uint32_t gb = ((b0 >> 29) & 1u) | (((b1 >> 29) & 1u) << 8) | (((b2 >> 29) & 1u) << 16) | (((b3 >> 29) & 1u) << 24);
            g_cnt += __popc(gb);
            o_cnt += __popc(gb & m);
            return;
        }
        const long long g0 = (long long)x.x, g1 = (long long)x.y, g2 = (long long)x.z, g3 = (long long)x.w;
        g_cnt += g0 + g1 + g2 + g3;
        if (c < slen) o_cnt += (long long)sm[c] & g0;
        if (c + 1 < slen) o_cnt += (long long)sm[c + 1] & g1;
        if (c + 2 < slen) o_cnt += (long long)sm[c + 2] & g2;
        if (c + 3 < slen) o_cnt += (long long)sm[c + 3] & g3;
    };
    int q = threadIdx.x;
    const int bd = blockDim.x;
    for (; q + 3 * bd < nvec; q += 4 * bd) {              // four 16-byte loads in flight per thread, streamed once (two: 0.192 ms, four: 0.160 ms)
        const float4 x0 = __ldcs(g4 + q), x1 = __ldcs(g4 + q + bd), x2 = __ldcs(g4 + q + 2 * bd), x3 = __ldcs(g4 + q + 3 * bd);
        add_group(x0, q);
        add_group(x1, q + bd);
        add_group(x2, q + 2 * bd);
        add_group(x3, q + 3 * bd);
    }
    for (; q < nvec; q += bd) add_group(__ldcs(g4 + q), q);
    for (int c = head + 4 * nvec + threadIdx.x; c < cols; c += blockDim.x) {
        const long long gi = (long long)__ldg(g + c);
        const long long si = c < slen ? (long long)sm[c] : 0;
        o_cnt += si & gi; g_cnt += gi;
    }
    s_cnt = mask_sum_partial(sm, slen);

    o_cnt = block_sum_i64(o_cnt, red);
    g_cnt = block_sum_i64(g_cnt, red);
    s_cnt = block_sum_i64(s_cnt, red);
    if (threadIdx.x == 0) {
        counts[3 * (int64_t)rowid + 0] = o_cnt;
        counts[3 * (int64_t)rowid + 1] = g_cnt;
        counts[3 * (int64_t)rowid + 2] = s_cnt;
    }
  }
}

// The same counts for user summaries stored as uint8 (the packed dataset's lossless form of the 0/1 float32 rows:
// a quarter of the bytes over PCIe and out of HBM).  Sixteen frames per 16-byte load; the mask bytes of each
// four-frame word come from one or two aligned words of the int8 summary.
__global__ void __launch_bounds__(256)
overlap_u8_kernel(const int8_t *__restrict__ summary, const int64_t *__restrict__ sum_offsets,
                  const uint8_t *__restrict__ user_summary, const int64_t *__restrict__ us_offsets,
                  const int32_t *__restrict__ cu_users, const int32_t *__restrict__ us_cols, int B,
                  long long *__restrict__ counts, int total_rows) {
    __shared__ long long red[32];
  for (int rowid = blockIdx.x; rowid < total_rows; rowid += gridDim.x) {
    const int v = find_segment(cu_users, B, rowid);
    const int u = rowid - __ldg(cu_users + v);
    const int cols = __ldg(us_cols + v);
    const int slen = (int)(__ldg(sum_offsets + v + 1) - __ldg(sum_offsets + v));
    const int8_t *sm = summary + __ldg(sum_offsets + v);
    const uint8_t *g = user_summary + __ldg(us_offsets + v) + (int64_t)u * cols;
    unsigned o_acc = 0, g_acc = 0;                                   // per thread <= 255 * cols / 256 + slack

    int head = (int)((16 - ((uintptr_t)g & 15)) & 15);
    if (head > cols) head = cols;
    const int nvec = (cols - head) >> 4;
    auto scalar = [&](int c) {                                        // S & G on the integer values
        const int gi = (int)__ldg(g + c);
        g_acc += gi;
        if (c < slen) o_acc += (unsigned)((int)sm[c] & gi);
    };
    if ((int)threadIdx.x < head) scalar(threadIdx.x);
    const uintptr_t sm_addr = (uintptr_t)(sm + head);
    const uint32_t *smw = reinterpret_cast<const uint32_t *>(sm_addr & ~(uintptr_t)3);
    const int sm_shift = (int)(sm_addr & 3) * 8;
    // four-frame words that lie wholly inside the summary (minus one when unaligned: the funnel shift reads on)
    const int full_words = max(0, min(4 * nvec, (slen - head) >> 2) - (sm_shift != 0 ? 1 : 0));
    auto add_word = [&](uint32_t gw, int w) {
        if (w < full_words) {
            const uint32_t w0 = __ldg(smw + w);
            const uint32_t m = sm_shift == 0 ? w0 : __funnelshift_r(w0, __ldg(smw + w + 1), sm_shift);
            if (((gw | m) & 0xfefefefeu) == 0u) {                    // all eight bytes are 0 / 1
                g_acc += __popc(gw);
                o_acc += __popc(gw & m);
                return;
            }
        }
        const int c = head + 4 * w;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gi = (int)((gw >> (8 * j)) & 0xffu);
            g_acc += gi;
            if (c + j < slen) o_acc += (unsigned)((int)sm[c + j] & gi);
        }
    };
    const uint4 *g16 = reinterpret_cast<const uint4 *>(g + head);
    auto add_group = [&](const uint4 &x, int q) {
        add_word(x.x, 4 * q); add_word(x.y, 4 * q + 1); add_word(x.z, 4 * q + 2); add_word(x.w, 4 * q + 3);
    };
    int q = threadIdx.x;
    for (; q + (int)blockDim.x < nvec; q += 2 * blockDim.x) {
        const uint4 x0 = __ldcs(g16 + q), x1 = __ldcs(g16 + q + blockDim.x);          // streamed once (four in flight measured slower here: 0.100 -> 0.114 ms)
        add_group(x0, q);
        add_group(x1, q + blockDim.x);
    }
    if (q < nvec) add_group(__ldcs(g16 + q), q);
    { const int c = head + 16 * nvec + threadIdx.x; if (c < cols) scalar(c); }       // < 16 tail frames

    const long long o_cnt = block_sum_i64((long long)o_acc, red);
    const long long g_cnt = block_sum_i64((long long)g_acc, red);
    const long long s_cnt = block_sum_i64(mask_sum_partial(sm, slen), red);
    if (threadIdx.x == 0) {
        counts[3 * (int64_t)rowid + 0] = o_cnt;
        counts[3 * (int64_t)rowid + 1] = g_cnt;
        counts[3 * (int64_t)rowid + 2] = s_cnt;
    }
  }
}

// fp64 ratios in the reference's evaluation order (evaluation_metrics.py:23-33).
__global__ void __launch_bounds__(128)
fscore_finalize_kernel(const long long *__restrict__ counts, const int32_t *__restrict__ cu_users,
                       int B, int method, double *__restrict__ f_out,
                       double *__restrict__ per_user_out) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= B) return;
    const int u0 = __ldg(cu_users + v), U = __ldg(cu_users + v + 1) - u0;
    double acc = 0.0, best = 0.0;
    for (int u = 0; u < U; ++u) {
        const long long o = counts[3 * (int64_t)(u0 + u)], g = counts[3 * (int64_t)(u0 + u) + 1],
                        s = counts[3 * (int64_t)(u0 + u) + 2];
        const double p = __ddiv_rn((double)o, (double)s);                // 0/0 -> NaN like numpy
        const double r = __ddiv_rn((double)o, (double)g);
        double f;
        if (__dadd_rn(p, r) == 0.0) f = 0.0;
        else f = __ddiv_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, p), r), 100.0), __dadd_rn(p, r));
        if (per_user_out) per_user_out[u0 + u] = f;
        if (u == 0) { best = f; acc = __dadd_rn(0.0, f); }
        else { if (f > best) best = f; acc = __dadd_rn(acc, f); }
    }
    f_out[v] = U == 0 ? __longlong_as_double(0x7ff8000000000000LL)
                      : (method == VSUM_FSCORE_MAX ? best : __ddiv_rn(acc, (double)U));
}

}  // namespace vsum

// =============================================================================================
// C ABI
// =============================================================================================
using namespace vsum;

extern "C" int vsum_shot_mean(const float *scores, const int32_t *cu_steps, const int32_t *picks,
                              const int32_t *cu_picks, const int32_t *n_frames, const int32_t *cps,
                              const int32_t *cu_shots, int32_t B, int32_t S_total, double *val_out,
                              int32_t *wt_out, int32_t *cap_out, void *stream) {
    VSUM_REQUIRE(B >= 0 && S_total >= 0, VSUM_EINVAL, "vsum_shot_mean: negative sizes");
    if (B == 0 || S_total == 0) return VSUM_OK;
    VSUM_REQUIRE(scores && cu_steps && picks && cu_picks && n_frames && cps && cu_shots && val_out &&
                 wt_out && cap_out, VSUM_EINVAL, "vsum_shot_mean: null pointer");
    const int threads = 128;
    ProfScope prof(PROF_SHOT_MEAN, (cudaStream_t)stream);
    shot_mean_kernel<<<(unsigned)ceil_div(S_total, threads), threads, 0, (cudaStream_t)stream>>>(
        scores, cu_steps, picks, cu_picks, n_frames, cps, cu_shots, B, S_total, val_out, wt_out,
        cap_out);
    VSUM_LAUNCH_OK("shot_mean_kernel");
    return VSUM_OK;
}

// Kernel classes: a video of capacity W runs in the smallest class whose width holds W + 1 cells.
static const int kKnapsackClassWidth[] = {256, 1024, 4096, 9728, 18944, 28672};

extern "C" int32_t vsum_knapsack_class_width(int32_t capacity) {
    for (int w : kKnapsackClassWidth)
        if (capacity + 1 <= w) return w;
    return -1;
}

extern "C" int64_t vsum_knapsack_scratch_words(int32_t n_shots, int32_t capacity) {
    if (n_shots <= 0 || capacity < 0) return 0;
    const int64_t padded = vsum_knapsack_class_width(capacity);             // rows padded to the kernel class width
    return padded < 0 ? -1 : (int64_t)n_shots * (padded / 32);
}

template <int THREADS, int EPT>
static int launch_knapsack(const double *val, const int32_t *wt, const int32_t *cu_shots,
                           const int32_t *cap, const int64_t *bit_offsets, const int32_t *order,
                           int32_t B, int32_t max_cap, uint32_t *take_bits, uint8_t *selected_out,
                           cudaStream_t stream) {
    const size_t smem = (size_t)THREADS * EPT * sizeof(double);
    auto kern = knapsack_kernel<THREADS, EPT>;
    if (smem > 48 * 1024)
        VSUM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = B;
    if (const int budget = eval_sm_budget()) {             // CTAs of this kernel that fit one SM (smem / threads)
        const int per_sm = (int)max((size_t)1, min((size_t)(2048 / THREADS), (size_t)(227 * 1024) / (smem + 1024)));
        grid = min(B, budget * per_sm);
    }
    ProfScope prof(PROF_KNAPSACK, stream);
    kern<<<grid, THREADS, smem, stream>>>(val, wt, cu_shots, cap, bit_offsets, order, take_bits, selected_out, B);
    VSUM_LAUNCH_OK("knapsack_kernel");
    return VSUM_OK;
}

extern "C" int vsum_knapsack(const double *val, const int32_t *wt, const int32_t *cu_shots,
                             const int32_t *cap, const int64_t *bit_offsets, const int32_t *order,
                             int32_t B, int32_t max_cap, uint32_t *take_bits,
                             uint8_t *selected_out, void *stream) {
    VSUM_REQUIRE(B >= 0 && max_cap >= 0, VSUM_EINVAL, "vsum_knapsack: negative sizes");
    if (B == 0) return VSUM_OK;
    VSUM_REQUIRE(val && wt && cu_shots && cap && bit_offsets && take_bits && selected_out, VSUM_EINVAL,
                 "vsum_knapsack: null pointer");
    const int width = max_cap + 1;
    cudaStream_t s = (cudaStream_t)stream;
#define VSUM_KS(T, E) return launch_knapsack<T, E>(val, wt, cu_shots, cap, bit_offsets, order, B, max_cap, take_bits, selected_out, s)
    if (width <= 256) VSUM_KS(128, 2);
    if (width <= 1024) VSUM_KS(256, 4);
    if (width <= 4096) VSUM_KS(512, 8);
    if (width <= 9728) VSUM_KS(512, 19);
    if (width <= 18944) VSUM_KS(512, 37);
    if (width <= 28672) VSUM_KS(512, 56);
#undef VSUM_KS
    return set_error(VSUM_EUNSUPPORTED,
                     "vsum_knapsack: capacity %d needs %zu B of shared memory for the fp64 DP row; "
                     "the sm_100a kernel holds at most 28672 capacities (n_frames <= 191146)",
                     max_cap, (size_t)width * 8);
}

extern "C" int vsum_summary_fscore(const uint8_t *selected, const int32_t *cps,
                                   const int32_t *cu_shots, const void *user_summary, int32_t user_summary_dtype,
                                   const int64_t *us_offsets, const int32_t *cu_users,
                                   const int32_t *us_cols, int32_t B, int32_t total_users,
                                   int32_t method, int8_t *summary_out, const int64_t *sum_offsets,
                                   int64_t summary_total, int64_t *counts_ws, double *f_out,
                                   double *per_user_out, void *stream) {
    VSUM_REQUIRE(B >= 0 && total_users >= 0, VSUM_EINVAL, "vsum_summary_fscore: negative sizes");
    if (B == 0) return VSUM_OK;
    VSUM_REQUIRE(summary_out && sum_offsets && (!selected || (cps && cu_shots)), VSUM_EINVAL,
                 "vsum_summary_fscore: null pointer");
    VSUM_REQUIRE(method == VSUM_FSCORE_AVG || method == VSUM_FSCORE_MAX, VSUM_EINVAL,
                 "vsum_summary_fscore: method must be VSUM_FSCORE_AVG or VSUM_FSCORE_MAX");
    cudaStream_t s = (cudaStream_t)stream;
    if (selected) {   // selected == NULL: summary_out already holds the masks (evaluate_summary)
        const int64_t avg = summary_total / B + 1;
        dim3 grid((unsigned)max((int64_t)1, min((int64_t)64, ceil_div(avg, 256 * 4))), (unsigned)B);
        ProfScope prof(PROF_MASK, s);
        summary_mask_kernel<<<grid, 256, 0, s>>>(selected, cps, cu_shots, sum_offsets, summary_out);
        VSUM_LAUNCH_OK("summary_mask_kernel");
    }
    if (!f_out) return VSUM_OK;                                          // generate_summary only
    VSUM_REQUIRE(user_summary && us_offsets && cu_users && us_cols && counts_ws, VSUM_EINVAL,
                 "vsum_summary_fscore: null pointer");
    VSUM_REQUIRE(user_summary_dtype == VSUM_USER_SUMMARY_F32 || user_summary_dtype == VSUM_USER_SUMMARY_U8, VSUM_EINVAL,
                 "vsum_summary_fscore: user_summary_dtype must be VSUM_USER_SUMMARY_F32 or VSUM_USER_SUMMARY_U8");
    if (total_users > 0) {
        ProfScope prof(PROF_OVERLAP, s);
        int grid = total_users;
        if (const int budget = eval_sm_budget()) grid = min(total_users, budget * 8);
        if (user_summary_dtype == VSUM_USER_SUMMARY_U8)
            overlap_u8_kernel<<<grid, 256, 0, s>>>(summary_out, sum_offsets, static_cast<const uint8_t *>(user_summary),
                                                   us_offsets, cu_users, us_cols, B,
                                                   reinterpret_cast<long long *>(counts_ws), total_users);
        else
            overlap_kernel<<<grid, 256, 0, s>>>(summary_out, sum_offsets, static_cast<const float *>(user_summary),
                                                us_offsets, cu_users, us_cols, B,
                                                reinterpret_cast<long long *>(counts_ws), total_users);
        VSUM_LAUNCH_OK("overlap_kernel");
    }
    ProfScope prof(PROF_FSCORE, s);
    fscore_finalize_kernel<<<(unsigned)ceil_div(B, 128), 128, 0, s>>>(
        reinterpret_cast<const long long *>(counts_ws), cu_users, B, method, f_out, per_user_out);
    VSUM_LAUNCH_OK("fscore_finalize_kernel");
    return VSUM_OK;
}

extern "C" int vsum_summary_frames(const uint8_t *selected, const int32_t *cps, const int32_t *cu_shots,
                                   const int64_t *out_offsets, int32_t B, int32_t *frames_out, int32_t *counts_out,
                                   void *stream) {
    VSUM_REQUIRE(B >= 0, VSUM_EINVAL, "vsum_summary_frames: negative batch");
    if (B == 0) return VSUM_OK;
    VSUM_REQUIRE(selected && cps && cu_shots && out_offsets && frames_out && counts_out, VSUM_EINVAL, "vsum_summary_frames: null pointer");
    vsum::summary_frames_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(selected, cps, cu_shots, out_offsets, frames_out, counts_out);
    VSUM_LAUNCH_OK("summary_frames_kernel");
    return VSUM_OK;
}
