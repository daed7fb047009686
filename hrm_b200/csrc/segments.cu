// K3: per-sweep-line interval algebra -> collision-free segments, then enhanceFreeSegment.
//
// Replaces (reference):
//   FreeSpaceComputator::computeSweepLineFreeSegment   include/hrm/datastructure/FreeSpace-inl.h:101-129
//   Interval::unions / intersects / complements        src/datastructure/Interval.cpp:10-100
//   FreeSpaceComputator::computeFreeSegment            FreeSpace-inl.h:63-97
//   FreeSpaceComputator::enhanceFreeSegment            FreeSpace-inl.h:132-174
//
// One WARP per sweep line.  Pass 1 (k_line_segments / k_line_segments4): compact the non-NaN C-obstacle intervals of the
// line into shared memory, sort by start, prefix-max the ends with shuffles, and emit the gaps clipped to the arena
// interval ("original" segments) into a compact pool (a line allocates exactly its count with one atomicAdd).  Pass 2
// (k_enhance<false>) counts the degenerate segments enhanceFreeSegment appends from the two adjacent lines of the same
// x-plane; a scan turns the counts into CSR offsets; pass 3 (k_enhance<true>) writes originals + appended values in the
// reference's (j1, j2) loop order and sorts xL, xU, xM of every line but the last one of a plane independently, exactly
// like the reference.
//
// Nothing here synchronises with the host.  The pool and the payload have CAPACITIES (grown from the sizes of earlier
// builds); a pass that would overrun one writes nothing and only reports the size it needs (info[0] = pool entries,
// info[1] = segments), and finalize_segments() -- called by the first getter that needs the results -- grows the buffer
// and repeats the passes.  All passes run on the context's side stream `s2`, behind the event the fused kernel records, so
// that the free segments of build i are formed while the fused kernel of build i+1 runs (interval tables double-buffered).
#include "internal.cuh"

#include <algorithm>
#include <cfloat>

namespace hrmgpu {

constexpr int kSegWarps = 8;
constexpr int kSeg4Warps = 7;  // 4-line pass: 7 warps x 8 KB of lists (128 intervals per line) = 56 KB per CTA; 256-entry lists measured 3% slower per step

struct SegParams {
    const double* lo;   // [K][S][L]
    const double* up;
    double* orig;       // pool of original segments [orig_cap][2]
    long long* orig_off;  // [K][L] first pool entry of the line
    int* orig_cnt;      // [K][L]
    int* seg_cnt;       // [K][L]
    const long long* seg_off;  // [K*L+1]
    double* segL;
    double* segU;
    double* segM;
    unsigned long long* info;  // [0] pool entries needed (allocation cursor), [1] total segments (written by the scan)
    long long orig_cap, seg_cap;  // capacities of the pool / of segL, segU, segM (entries)
    int K, S, Sa, L, Lx, Ly;
    int warps;                     // warps per CTA in pass 1
    int sort_cap;                  // intervals per warp in shared memory
    int only_flagged;              // pass 1 fallback: process only the lines k_line_segments4 marked with orig_cnt = -1
};

// Ascending sort of n keys (with optional payload) by a warp; works for any n (virtual +inf padding).
template <bool PAYLOAD>
__device__ __forceinline__ void warp_bitonic(double* key, double* val, int n, int lane) {
    for (int k = 2; (k >> 1) < n; k <<= 1) {
        for (int t = lane; t < n; t += 32) {
            const int l = t ^ (k - 1);
            if (l > t && l < n) {
                const double a = key[t], b = key[l];
                if (a > b) {
                    key[t] = b;
                    key[l] = a;
                    if (PAYLOAD) {
                        const double va = val[t];
                        val[t] = val[l];
                        val[l] = va;
                    }
                }
            }
        }
        __syncwarp();
        for (int j = k >> 2; j > 0; j >>= 1) {
            for (int t = lane; t < n; t += 32) {
                const int l = t ^ j;
                if (l > t && l < n) {
                    const double a = key[t], b = key[l];
                    if (a > b) {
                        key[t] = b;
                        key[l] = a;
                        if (PAYLOAD) {
                            const double va = val[t];
                            val[t] = val[l];
                            val[l] = va;
                        }
                    }
                }
            }
            __syncwarp();
        }
    }
}

// Ascending sort of m <= 32 * SLOTS (start, end) pairs by start: every lane keeps its pairs in registers, counts for each
// of them the pairs that precede it (smaller start, or equal start and smaller position: a permutation) while the starts
// are broadcast from shared memory, then scatters its pairs to their ranks in place.  No dependent shared-memory passes
// and no barriers inside the counting loop, unlike the bitonic network (28 passes for 128 pairs); used for the short lists
// of k_line_segments4.  The order among equal starts does not affect the gaps (an interval's end is never below its
// start, so the second of two equal starts never opens a gap).
template <int SLOTS>
__device__ __forceinline__ void warp_rank_sort(double* ks, double* ke, int m, int lane) {
    double k[SLOTS], v[SLOTS];
    int r[SLOTS];
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        const int i = lane + 32 * s;
        k[s] = i < m ? ks[i] : DBL_MAX;
        v[s] = i < m ? ke[i] : 0.0;
        r[s] = 0;
    }
    for (int j = 0; j < m; ++j) {
        const double kj = ks[j];  // same address in all lanes: one broadcast
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) r[s] += (kj < k[s] || (kj == k[s] && j < lane + 32 * s)) ? 1 : 0;
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < SLOTS; ++s)
        if (lane + 32 * s < m) ks[r[s]] = k[s], ke[r[s]] = v[s];
    __syncwarp();
}
__device__ __forceinline__ void warp_sort_pairs(double* ks, double* ke, int m, int lane) {
    if (m <= 32) warp_rank_sort<1>(ks, ke, m, lane);
    else if (m <= 64) warp_rank_sort<2>(ks, ke, m, lane);
    else if (m <= 128) warp_rank_sort<4>(ks, ke, m, lane);
    else if (m <= 256) warp_rank_sort<8>(ks, ke, m, lane);
    else warp_bitonic<true>(ks, ke, m, lane);
}

// Original free segments of one line from its m compacted C-obstacle intervals (ks = starts, ke = ends, in shared
// memory) and its arena interval [as, ae]: sort by start, gaps of the union clipped to the arena.  The gaps are counted
// first, the line then takes exactly that many entries of the pool (one atomicAdd) and writes them; with at most 32
// candidates (the common case) the values stay in registers between the two steps.  A line that does not fit into the
// pool writes nothing: the cursor keeps counting, and the host repeats the pass with a pool of that size.
__device__ __forceinline__ void emit_line(const SegParams& p, long long gl, double* ks, double* ke, int m, double as, double ae,
                                          bool arena_ok, int lane) {
    int cnt = 0;
    const bool trivial = !arena_ok || m == 0;  // complements(): empty outer -> nothing; empty inner -> outer
    if (!trivial) warp_sort_pairs(ks, ke, m, lane);
    // gaps of the union: before interval i iff (running max of earlier ends) < start_i   (Interval.cpp:21-29)
    double gs = 0, ge = 0;
    bool keep = false;
    auto sweep = [&](double* out) -> int {  // out == nullptr: count only
        int c = 0;
        double carry = -DBL_MAX;  // running max of ends of intervals [0, base)
        for (int base = 0; base < m + 1; base += 32) {
            const int i = base + lane;
            double e = (i < m) ? ke[i] : -DBL_MAX;
            double inc = e;  // inclusive prefix max within this chunk
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                if (lane >= o) inc = fmax(inc, t);
            }
            double excl = __shfl_up_sync(0xFFFFFFFFu, inc, 1);
            excl = (lane == 0) ? carry : fmax(carry, excl);  // max of ends of intervals [0, i)
            keep = false;
            if (i <= m) {
                const bool is_gap = (i == 0) || (i == m) || (excl < ks[i]);
                if (is_gap) {
                    gs = (i == 0) ? -DBL_MAX : excl;       // Interval.cpp:80-84
                    ge = (i == m) ? DBL_MAX : ks[i];
                    gs = fmax(gs, as);                     // intersects({gap, outer}): max start, min end
                    ge = fmin(ge, ae);
                    keep = !(gs > ge);
                }
            }
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, keep);
            if (out && keep) {
                const int pos = c + __popc(bal & ((1u << lane) - 1u));
                out[2 * pos] = gs;
                out[2 * pos + 1] = ge;
            }
            c += __popc(bal);
            carry = fmax(carry, __shfl_sync(0xFFFFFFFFu, inc, 31));
        }
        return c;
    };
    if (!arena_ok)
        cnt = 0;
    else if (m == 0)
        cnt = 1;
    else
        cnt = sweep(nullptr);
    long long base = 0;
    if (lane == 0 && cnt > 0) base = static_cast<long long>(atomicAdd(p.info, static_cast<unsigned long long>(cnt)));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    const bool fits = base + cnt <= p.orig_cap;
    if (fits && cnt > 0) {
        double* out = p.orig + 2 * base;
        if (trivial) {
            if (lane == 0) out[0] = as, out[1] = ae;
        } else if (m + 1 <= 32) {  // single chunk: keep / gs / ge of the counting sweep are still in registers
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, keep);
            if (keep) {
                const int pos = __popc(bal & ((1u << lane) - 1u));
                out[2 * pos] = gs;
                out[2 * pos + 1] = ge;
            }
        } else {
            sweep(out);
        }
    }
    if (lane == 0) {
        p.orig_off[gl] = base;
        p.orig_cnt[gl] = fits ? cnt : 0;
    }
}

// Pass 1: original free segments of every line.  grid = ceil(K*L / warps) CTAs of `warps` warps.
__global__ void __launch_bounds__(kSegWarps * 32) k_line_segments(const SegParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= p.warps) return;
    const long long nl = static_cast<long long>(p.K) * p.L;
    // p.only_flagged: fallback pass behind k_line_segments4, a small grid walks all lines and redoes the overflowed ones
    for (long long gl = static_cast<long long>(blockIdx.x) * p.warps + warp; gl < nl; gl += static_cast<long long>(gridDim.x) * p.warps) {
        if (p.only_flagged && p.orig_cnt[gl] != -1) continue;
        const int k = static_cast<int>(gl / p.L);
        const int l = static_cast<int>(gl - static_cast<long long>(k) * p.L);
        double* ks = reinterpret_cast<double*>(smem_raw) + static_cast<size_t>(warp) * 2 * p.sort_cap;  // starts
        double* ke = ks + p.sort_cap;                                                                      // ends

        const double* lo = p.lo + static_cast<size_t>(k) * p.S * p.L + l;
        const double* up = p.up + static_cast<size_t>(k) * p.S * p.L + l;

        // arena: A = [max lo, min up] over non-NaN arena entries (Interval::intersects)
        double as = -DBL_MAX, ae = DBL_MAX;
        int na = 0;
        for (int j = lane; j < p.Sa; j += 32) {
            const double a = lo[static_cast<size_t>(j) * p.L], b = up[static_cast<size_t>(j) * p.L];
            if (!isnan(a) && !isnan(b)) {
                as = fmax(as, a);
                ae = fmin(ae, b);
                ++na;
            }
        }
    #pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            as = fmax(as, __shfl_xor_sync(0xFFFFFFFFu, as, o));
            ae = fmin(ae, __shfl_xor_sync(0xFFFFFFFFu, ae, o));
            na += __shfl_xor_sync(0xFFFFFFFFu, na, o);
        }
        const bool arena_ok = (na > 0) && !(as > ae);

        // obstacles: ordered compaction of the non-NaN intervals
        // (eight chunks of 32 surfaces are requested before the first one is consumed: the loads are strided by L doubles,
        // so the loop is bound by memory latency, not bandwidth)
        int m = 0;
        constexpr int kUnroll = 8;
        for (int base = p.Sa; base < p.S; base += 32 * kUnroll) {
            double a[kUnroll], b[kUnroll];
    #pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int j = base + 32 * u + lane;
                a[u] = b[u] = __longlong_as_double(0x7FF8000000000000LL);
                if (j < p.S) {
                    a[u] = __ldg(lo + static_cast<size_t>(j) * p.L);
                    b[u] = __ldg(up + static_cast<size_t>(j) * p.L);
                }
            }
    #pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const bool ok = !isnan(a[u]) && !isnan(b[u]);
                const unsigned bal = __ballot_sync(0xFFFFFFFFu, ok);
                if (ok) {
                    const int pos = m + __popc(bal & ((1u << lane) - 1u));
                    ks[pos] = a[u];
                    ke[pos] = b[u];
                }
                m += __popc(bal);
            }
        }
        __syncwarp();

        emit_line(p, gl, ks, ke, m, as, ae, arena_ok, lane);
        __syncwarp();
    }
}

// Pass 1, fast path (L % 4 == 0): one warp per group of 4 consecutive lines.  A lane reads the 32-byte sector that holds
// the four lines' entries of one surface (the tables are surface-major, so one line alone would use 8 of every 32 bytes
// fetched), and the warp compacts four interval lists at once, `cap4` entries each.  A line with more intervals than
// that is flagged (orig_cnt = -1) and redone by k_line_segments with full-size lists.
__global__ void __launch_bounds__(kSeg4Warps * 32, 2) k_line_segments4(const SegParams p, int cap4) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long grp = static_cast<long long>(blockIdx.x) * kSeg4Warps + warp;
    const long long gl0 = grp * 4;  // first global line id (k*L + l) of the group; L % 4 == 0 keeps a group inside one layer
    if (gl0 >= static_cast<long long>(p.K) * p.L) return;
    const int k = static_cast<int>(gl0 / p.L);
    const int l = static_cast<int>(gl0 - static_cast<long long>(k) * p.L);
    double* base_s = reinterpret_cast<double*>(smem_raw) + static_cast<size_t>(warp) * 8 * cap4;  // [4][2][cap4]
    const double* lo = p.lo + static_cast<size_t>(k) * p.S * p.L + l;
    const double* up = p.up + static_cast<size_t>(k) * p.S * p.L + l;
    const unsigned lt = (1u << lane) - 1u;

    double as[4], ae[4];
    int na[4], m[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) as[t] = -DBL_MAX, ae[t] = DBL_MAX, na[t] = 0, m[t] = 0;
    for (int j = lane; j < p.Sa; j += 32) {
        const double2 a01 = __ldg(reinterpret_cast<const double2*>(lo + static_cast<size_t>(j) * p.L));
        const double2 a23 = __ldg(reinterpret_cast<const double2*>(lo + static_cast<size_t>(j) * p.L) + 1);
        const double2 b01 = __ldg(reinterpret_cast<const double2*>(up + static_cast<size_t>(j) * p.L));
        const double2 b23 = __ldg(reinterpret_cast<const double2*>(up + static_cast<size_t>(j) * p.L) + 1);
        const double a[4] = {a01.x, a01.y, a23.x, a23.y}, b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
        for (int t = 0; t < 4; ++t)
            if (!isnan(a[t]) && !isnan(b[t])) as[t] = fmax(as[t], a[t]), ae[t] = fmin(ae[t], b[t]), ++na[t];
    }
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            as[t] = fmax(as[t], __shfl_xor_sync(0xFFFFFFFFu, as[t], o));
            ae[t] = fmin(ae[t], __shfl_xor_sync(0xFFFFFFFFu, ae[t], o));
            na[t] += __shfl_xor_sync(0xFFFFFFFFu, na[t], o);
        }

    constexpr int kUnroll = 4;  // 16 sector loads in flight per lane
    const double qnan = __longlong_as_double(0x7FF8000000000000LL);
    for (int base = p.Sa; base < p.S; base += 32 * kUnroll) {
        double2 a01[kUnroll], a23[kUnroll], b01[kUnroll], b23[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int j = base + 32 * u + lane;
            a01[u] = a23[u] = b01[u] = b23[u] = make_double2(qnan, qnan);
            if (j < p.S) {
                const double2* pa = reinterpret_cast<const double2*>(lo + static_cast<size_t>(j) * p.L);
                const double2* pb = reinterpret_cast<const double2*>(up + static_cast<size_t>(j) * p.L);
                a01[u] = __ldg(pa), a23[u] = __ldg(pa + 1), b01[u] = __ldg(pb), b23[u] = __ldg(pb + 1);
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const double a[4] = {a01[u].x, a01[u].y, a23[u].x, a23[u].y}, b[4] = {b01[u].x, b01[u].y, b23[u].x, b23[u].y};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const bool ok = !isnan(a[t]) && !isnan(b[t]);
                const unsigned bal = __ballot_sync(0xFFFFFFFFu, ok);
                if (ok) {
                    const int pos = m[t] + __popc(bal & lt);
                    if (pos < cap4) {
                        base_s[(2 * t) * cap4 + pos] = a[t];
                        base_s[(2 * t + 1) * cap4 + pos] = b[t];
                    }
                }
                m[t] += __popc(bal);
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        if (m[t] > cap4) {
            if (lane == 0) p.orig_cnt[gl0 + t] = -1;  // overflow: redone by the fallback pass
            continue;
        }
        emit_line(p, gl0 + t, base_s + (2 * t) * cap4, base_s + (2 * t + 1) * cap4, m[t], as[t], ae[t], (na[t] > 0) && !(as[t] > ae[t]), lane);
        __syncwarp();
    }
}

// Pass 1 from the per-line buckets the warp-specialised fused kernel fills (mink3d_ws.cu): a line's obstacle intervals are
// `m` compact (lo, up) pairs instead of a stride over all S surfaces of the dense tables -- ~65 x 16 bytes per line at the
// bench sizes instead of 1001 x 2 sectors.  One warp per line; a line whose bucket overflowed (m > bcap) is flagged and redone
// from the dense tables by k_line_segments.  The order of the pairs inside a bucket is arbitrary (atomic append), which the
// union does not depend on (see warp_rank_sort).  The arena interval still comes from the dense tables (Sa entries).
constexpr int kBucketWarps = 8;
constexpr int kBucketMax = 128;
__global__ void __launch_bounds__(kBucketWarps * 32) k_line_segments_b(const SegParams p, const int* __restrict__ bcnt, const double* __restrict__ bucket,
                                                                        int bcap) {
    __shared__ __align__(16) double lists[kBucketWarps][2][kBucketMax];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long gl = static_cast<long long>(blockIdx.x) * kBucketWarps + warp;
    if (gl >= static_cast<long long>(p.K) * p.L) return;
    const int m = bcnt[gl];
    if (m > bcap) {
        if (lane == 0) p.orig_cnt[gl] = -1;
        return;
    }
    const int k = static_cast<int>(gl / p.L);
    const int l = static_cast<int>(gl - static_cast<long long>(k) * p.L);
    const double* lo = p.lo + static_cast<size_t>(k) * p.S * p.L + l;
    const double* up = p.up + static_cast<size_t>(k) * p.S * p.L + l;
    double as = -DBL_MAX, ae = DBL_MAX;
    int na = 0;
    for (int j = lane; j < p.Sa; j += 32) {
        const double a = lo[static_cast<size_t>(j) * p.L], b = up[static_cast<size_t>(j) * p.L];
        if (!isnan(a) && !isnan(b)) as = fmax(as, a), ae = fmin(ae, b), ++na;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        as = fmax(as, __shfl_xor_sync(0xFFFFFFFFu, as, o));
        ae = fmin(ae, __shfl_xor_sync(0xFFFFFFFFu, ae, o));
        na += __shfl_xor_sync(0xFFFFFFFFu, na, o);
    }
    double* ks = lists[warp][0];
    double* ke = lists[warp][1];
    const double2* src = reinterpret_cast<const double2*>(bucket) + static_cast<size_t>(gl) * bcap;
    for (int i = lane; i < m; i += 32) {
        const double2 v = __ldg(src + i);
        ks[i] = v.x;
        ke[i] = v.y;
    }
    __syncwarp();
    emit_line(p, gl, ks, ke, m, as, ae, (na > 0) && !(as > ae), lane);
}

// The four append rules of enhanceFreeSegment for the ordered pair (line i, line i+1), segment j1 of
// line i = (s1,e1), segment j2 of line i+1 = (s2,e2).  Returns bit0: append to line i (value v_i),
// bit1: append to line i+1 (value v_n).                                      FreeSpace-inl.h:139-161
__device__ __forceinline__ int enhance_pair(double s1, double e1, double s2, double e2, double& v_i, double& v_n) {
    const double m1 = (s1 + e1) / 2.0, m2 = (s2 + e2) / 2.0;
    int r = 0;
    if (m1 < s2 && e1 >= s2) {
        v_i = s2;
        r |= 1;
    } else if (m1 > e2 && s1 <= e2) {
        v_i = e2;
        r |= 1;
    }
    if (m2 < s1 && e2 >= s1) {
        v_n = s1;
        r |= 2;
    } else if (m2 > e1 && s2 <= e1) {
        v_n = e1;
        r |= 2;
    }
    return r;
}

// Pass 2 (FILL=false): final count per line.  Pass 3 (FILL=true): write + sort.  One warp per line.
template <bool FILL>
__global__ void __launch_bounds__(kSegWarps * 32) k_enhance(const SegParams p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long gl = static_cast<long long>(blockIdx.x) * kSegWarps + warp;
    if (gl >= static_cast<long long>(p.K) * p.L) return;
    const int l = static_cast<int>(gl % p.L);
    const int iy = l % p.Ly;
    if (FILL && static_cast<long long>(p.info[1]) > p.seg_cap) return;  // payload too small: the host grows it and repeats this pass
    const int n0 = p.orig_cnt[gl];
    const double* s0 = p.orig + 2 * p.orig_off[gl];
    int cnt = n0;
    long long base_off = 0;
    double *oL = nullptr, *oU = nullptr, *oM = nullptr;
    if (FILL) {
        base_off = p.seg_off[gl];
        oL = p.segL + base_off;
        oU = p.segU + base_off;
        oM = p.segM + base_off;
        for (int j = lane; j < n0; j += 32) {
            const double s = s0[2 * j], e = s0[2 * j + 1];
            oL[j] = s;
            oU[j] = e;
            oM[j] = (s + e) / 2.0;  // FreeSpace-inl.h:84-88
        }
    }
    // (a) this line in the role "i+1" of the pair (iy-1, iy): reference processes that pair first
    if (iy > 0) {
        const int np = p.orig_cnt[gl - 1];
        const double* sp = p.orig + 2 * p.orig_off[gl - 1];
        const long long tot = static_cast<long long>(np) * n0;  // (j1 over previous line) x (j2 over this line)
        for (long long b0 = 0; b0 < tot; b0 += 32) {
            const long long q = b0 + lane;
            bool app = false;
            double v = 0;
            if (q < tot) {
                const int j1 = static_cast<int>(q / n0), j2 = static_cast<int>(q - static_cast<long long>(j1) * n0);
                double vi, vn;
                const int r = enhance_pair(sp[2 * j1], sp[2 * j1 + 1], s0[2 * j2], s0[2 * j2 + 1], vi, vn);
                app = (r & 2) != 0;
                v = vn;
            }
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, app);
            if (FILL && app) {
                const int pos = cnt + __popc(bal & ((1u << lane) - 1u));
                oL[pos] = v;
                oU[pos] = v;
                oM[pos] = v;
            }
            cnt += __popc(bal);
        }
    }
    // (b) this line in the role "i" of the pair (iy, iy+1)
    if (iy < p.Ly - 1) {
        const int nn = p.orig_cnt[gl + 1];
        const double* sn = p.orig + 2 * p.orig_off[gl + 1];
        const long long tot = static_cast<long long>(n0) * nn;
        for (long long b0 = 0; b0 < tot; b0 += 32) {
            const long long q = b0 + lane;
            bool app = false;
            double v = 0;
            if (q < tot) {
                const int j1 = static_cast<int>(q / nn), j2 = static_cast<int>(q - static_cast<long long>(j1) * nn);
                double vi, vn;
                const int r = enhance_pair(s0[2 * j1], s0[2 * j1 + 1], sn[2 * j2], sn[2 * j2 + 1], vi, vn);
                app = (r & 1) != 0;
                v = vi;
            }
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, app);
            if (FILL && app) {
                const int pos = cnt + __popc(bal & ((1u << lane) - 1u));
                oL[pos] = v;
                oU[pos] = v;
                oM[pos] = v;
            }
            cnt += __popc(bal);
        }
    }
    if (!FILL) {
        if (lane == 0) p.seg_cnt[gl] = cnt;
        return;
    }
    // independent ascending sorts of xL, xU, xM for every line but the last of its plane (FreeSpace-inl.h:165-170)
    if (iy < p.Ly - 1 && cnt > 1) {
        __syncwarp();
        warp_bitonic<false>(oL, nullptr, cnt, lane);
        warp_bitonic<false>(oU, nullptr, cnt, lane);
        warp_bitonic<false>(oM, nullptr, cnt, lane);
    }
}

// Exclusive scan of n counts into 64-bit CSR offsets, off[n] = info[1] = the total.  Chained tiles: a CTA takes the next tile of
// 8192 counts from a ticket counter (so the CTA holding the previous tile is already running), scans it locally while the previous
// tiles are still in flight, then thread 0 waits for the previous tile's inclusive total (state[1 + tile], stored +1 so that the
// zeroed state means "not there yet") and publishes its own.  A single CTA took 60-100 us for the 10^5 lines / line pairs of the
// dense maze scene, a visible share of plan().
constexpr int kScanItems = 8;
constexpr int kScanTile = 1024 * kScanItems;
__global__ void __launch_bounds__(1024) k_scan_counts(const int* cnt, long long* off, long long n, unsigned long long* info, unsigned long long* state) {
    __shared__ long long wsum[32];
    __shared__ long long prev_s;
    __shared__ unsigned long long tile_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) tile_s = atomicAdd(state, 1ULL);
    __syncthreads();
    const long long tile = static_cast<long long>(tile_s);
    const long long i0 = tile * kScanTile + static_cast<long long>(tid) * kScanItems;
    int v[kScanItems];
    long long sum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (i0 + k < n) ? cnt[i0 + k] : 0;
        sum += v[k];
    }
    long long inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        long long w = wsum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xFFFFFFFFu, w, o);
            if (lane >= o) w += t;
        }
        wsum[lane] = w;
        if (lane == 31) {  // w = the tile's total
            unsigned long long prev = 1ULL;
            if (tile > 0) {
                volatile unsigned long long* ps = state + tile;  // = state[1 + (tile - 1)]
                while ((prev = *ps) == 0ULL) {
                }
            }
            const long long before = static_cast<long long>(prev - 1ULL);
            atomicExch(state + 1 + tile, static_cast<unsigned long long>(before + w) + 1ULL);
            prev_s = before;
            if ((tile + 1) * kScanTile >= n) {  // the last tile
                off[n] = before + w;
                info[1] = static_cast<unsigned long long>(before + w);
            }
        }
    }
    __syncthreads();
    long long run = prev_s + (warp > 0 ? wsum[warp - 1] : 0) + inc - sum;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (i0 + k < n) off[i0 + k] = run;
        run += v[k];
    }
}

int launch_scan_counts(Ctx* c, const int* cnt, long long* off, long long n, unsigned long long* info, cudaStream_t st) {
    const long long tiles = n > 0 ? (n + kScanTile - 1) / kScanTile : 1;
    DevBuf<unsigned long long>& state = (st == c->s2) ? c->scan_state_side : c->scan_state_main;  // (the two streams may scan at the same time)
    int rc;
    if (state.n < static_cast<size_t>(tiles) + 2 && (rc = ensure_on(c, st, state, static_cast<size_t>(tiles) + 2 + 64))) return rc;  // (stream-ordered)
    HRMGPU_CUDA(c, cudaMemsetAsync(state.p, 0, (static_cast<size_t>(tiles) + 2) * sizeof(unsigned long long), st));
    k_scan_counts<<<static_cast<unsigned>(tiles), 1024, 0, st>>>(cnt, off, n, info, state.p);
    HRMGPU_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return HRMGPU_OK;
}

// Enqueues pass 1 .. 3 on the side stream (no host synchronisation).  `from_pass`: 1 = everything, 3 = the fill pass only
// (after the payload was grown).
static int enqueue_passes(Ctx* c, int from_pass) {
    SegParams p;
    p.lo = c->ivl_lo[c->ivl_cur].p;
    p.up = c->ivl_up[c->ivl_cur].p;
    p.K = c->K;
    p.S = c->S;
    p.Sa = c->Sa;
    p.L = c->L;
    p.Lx = (c->dim == 3) ? c->Lx : 1;
    p.Ly = c->Ly;
    const int So = c->S - c->Sa;
    const long long nl = static_cast<long long>(c->K) * c->L;
    p.orig = c->orig.p;
    p.orig_off = c->orig_off.p;
    p.orig_cnt = c->orig_cnt.p;
    p.seg_cnt = c->seg_cnt.p;
    p.seg_off = c->seg_off.p;
    p.segL = c->segdata.p;
    p.segU = c->segdata.p + c->seg_cap;
    p.segM = c->segdata.p + 2 * c->seg_cap;
    p.info = c->seg_info.p;
    p.orig_cap = c->orig_cap;
    p.seg_cap = c->seg_cap;
    cudaStream_t st = c->s2;
    const long long g2 = (nl + kSegWarps - 1) / kSegWarps;
    if (from_pass <= 1) {
        HRMGPU_CUDA(c, cudaMemsetAsync(c->seg_info.p, 0, 2 * sizeof(unsigned long long), st));
        // pass 1: shared memory holds `warps` x sort_cap (start,end) pairs
        p.sort_cap = So > 0 ? So : 1;
        const size_t per_warp = static_cast<size_t>(p.sort_cap) * 16;
        int warps = static_cast<int>((200 * 1024) / per_warp);
        if (warps < 1) return fail(c, HRMGPU_E_CAP, "more than 12800 C-obstacle surfaces per layer are not supported");
        if (warps > kSegWarps) warps = kSegWarps;
        p.warps = warps;
        const size_t smem = per_warp * warps;
        HRMGPU_CUDA(c, cudaFuncSetAttribute(k_line_segments, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        const long long g1 = (nl + warps - 1) / warps;
        p.only_flagged = 0;
        if (c->bucket_valid[c->ivl_cur] && c->bucket_cap <= kBucketMax && !c->force_seg_fallback) {
            // compact per-line buckets from the warp-specialised fused kernel; overflowed lines redone from the dense tables
            const long long gb = (nl + kBucketWarps - 1) / kBucketWarps;
            k_line_segments_b<<<static_cast<unsigned>(gb), kBucketWarps * 32, 0, st>>>(p, c->bcnt[c->ivl_cur].p, c->bucket[c->ivl_cur].p, c->bucket_cap);
            HRMGPU_CUDA(c, cudaGetLastError());
            p.only_flagged = 1;
            k_line_segments<<<static_cast<unsigned>(std::min<long long>(g1, c->num_sms)), warps * 32, smem, st>>>(p);
            HRMGPU_CUDA(c, cudaGetLastError());
            c->launches += 2;
        } else if (c->L % 4 == 0 && !c->force_seg_fallback) {
            // fast path: a warp per 4 lines with short lists, then the full-size kernel on a small grid for flagged lines
            const int cap4 = std::min(p.sort_cap, c->seg_fast_cap > 0 ? c->seg_fast_cap : 128);
            const size_t smem4 = static_cast<size_t>(kSeg4Warps) * 8 * cap4 * sizeof(double);
            HRMGPU_CUDA(c, cudaFuncSetAttribute(k_line_segments4, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem4)));
            const long long groups = nl / 4;
            k_line_segments4<<<static_cast<unsigned>((groups + kSeg4Warps - 1) / kSeg4Warps), kSeg4Warps * 32, smem4, st>>>(p, cap4);
            HRMGPU_CUDA(c, cudaGetLastError());
            c->launches += 1;
            if (cap4 < p.sort_cap) {
                p.only_flagged = 1;
                k_line_segments<<<static_cast<unsigned>(std::min<long long>(g1, c->num_sms)), warps * 32, smem, st>>>(p);
                HRMGPU_CUDA(c, cudaGetLastError());
                c->launches += 1;
            }
        } else {
            k_line_segments<<<static_cast<unsigned>(g1), warps * 32, smem, st>>>(p);
            HRMGPU_CUDA(c, cudaGetLastError());
            c->launches += 1;
        }
        k_enhance<false><<<static_cast<unsigned>(g2), kSegWarps * 32, 0, st>>>(p);
        HRMGPU_CUDA(c, cudaGetLastError());
        HRMGPU_CUDA(c, cudaGetLastError());
        int rc_scan;
        if ((rc_scan = launch_scan_counts(c, c->seg_cnt.p, c->seg_off.p, nl, c->seg_info.p, st))) return rc_scan;
        c->launches += 1;
    }
    k_enhance<true><<<static_cast<unsigned>(g2), kSegWarps * 32, 0, st>>>(p);
    HRMGPU_CUDA(c, cudaGetLastError());
    c->launches += 1;
    // the two sizes the host needs (16 bytes, pinned): read when a getter finalises the build
    HRMGPU_CUDA(c, cudaMemcpyAsync(c->h_info, c->seg_info.p, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    return HRMGPU_OK;
}

// Queues the free-segment passes of the build whose fused kernel was just enqueued on the main stream.
int launch_segments(Ctx* c) {
    const long long nl = static_cast<long long>(c->K) * c->L;
    c->seg_pending = false;
    c->h_off_valid = false;
    c->seg_total = 0;
    if (nl == 0) return HRMGPU_OK;
    int rc;
    cudaStream_t st = c->s2;
    // capacities: what the previous builds needed (x1.5), at least a few entries per line; a pass that overruns them reports
    // the size it needs and finalize_segments() repeats it
    long long want_orig = std::max<long long>(std::max(c->orig_cap, c->orig_hint), std::max<long long>(nl * 4, 4096));
    long long want_seg = std::max<long long>(std::max(c->seg_cap, c->seg_hint), std::max<long long>(nl * 6, 4096));
    if (c->seg_init_cap > 0) want_orig = std::max(c->orig_cap, c->seg_init_cap), want_seg = std::max(c->seg_cap, c->seg_init_cap);  // test hook
    // the side stream owns these buffers: (re)allocations are ordered on it
    if ((rc = ensure_on(c, st, c->orig, static_cast<size_t>(want_orig) * 2))) return rc;
    c->orig_cap = want_orig;
    if ((rc = ensure_on(c, st, c->segdata, static_cast<size_t>(want_seg) * 3))) return rc;
    c->seg_cap = want_seg;
    if ((rc = ensure_on(c, st, c->orig_off, static_cast<size_t>(nl)))) return rc;
    if ((rc = ensure_on(c, st, c->orig_cnt, static_cast<size_t>(nl)))) return rc;
    if ((rc = ensure_on(c, st, c->seg_cnt, static_cast<size_t>(nl)))) return rc;
    if ((rc = ensure_on(c, st, c->seg_off, static_cast<size_t>(nl) + 1))) return rc;
    if ((rc = ensure_on(c, st, c->seg_info, 2))) return rc;
    // behind the fused kernel of this build
    HRMGPU_CUDA(c, cudaEventRecord(c->ev_fused, c->stream));
    HRMGPU_CUDA(c, cudaStreamWaitEvent(st, c->ev_fused, 0));
    HRMGPU_CUDA(c, cudaEventRecord(c->ev2, st));
    if ((rc = enqueue_passes(c, 1))) return rc;
    HRMGPU_CUDA(c, cudaEventRecord(c->ev3, st));
    HRMGPU_CUDA(c, cudaEventRecord(c->ev_k3[c->ivl_cur], st));
    c->seg_pending = true;
    return HRMGPU_OK;
}

// Host side of the capacity protocol: waits for the passes of the last build, and if the pool or the payload was too small
// grows it and repeats them.  Every getter / consumer of the free segments calls this first.
int finalize_segments(Ctx* c) {
    if (!c->seg_pending) return HRMGPU_OK;
    for (int attempt = 0; attempt < 4; ++attempt) {
        HRMGPU_CUDA(c, cudaEventSynchronize(c->ev_k3[c->ivl_cur]));
        const long long need_orig = static_cast<long long>(c->h_info[0]), need_seg = static_cast<long long>(c->h_info[1]);
        int rc;
        if (need_orig > c->orig_cap) {
            c->orig_cap = need_orig + need_orig / 2;
            if ((rc = ensure_on(c, c->s2, c->orig, static_cast<size_t>(c->orig_cap) * 2))) return rc;
            if (need_seg > c->seg_cap) {  // (a lower bound: lines that did not fit counted as empty)
                c->seg_cap = need_seg + need_seg / 2;
                if ((rc = ensure_on(c, c->s2, c->segdata, static_cast<size_t>(c->seg_cap) * 3))) return rc;
            }
            if ((rc = enqueue_passes(c, 1))) return rc;
        } else if (need_seg > c->seg_cap) {
            c->seg_cap = need_seg + need_seg / 2;
            if ((rc = ensure_on(c, c->s2, c->segdata, static_cast<size_t>(c->seg_cap) * 3))) return rc;
            if ((rc = enqueue_passes(c, 3))) return rc;
        } else {
            c->seg_total = need_seg;
            c->seg_pending = false;
            // remember what this workload needs (with slack) so that the next build of a similar size fits at once
            c->orig_hint = std::max<long long>(c->orig_hint, need_orig + need_orig / 2);
            c->seg_hint = std::max<long long>(c->seg_hint, need_seg + need_seg / 2);
            return HRMGPU_OK;
        }
        c->seg_redos += 1;
        HRMGPU_CUDA(c, cudaEventRecord(c->ev_k3[c->ivl_cur], c->s2));
    }
    return fail(c, HRMGPU_E_CUDA, "free-segment passes did not converge");
}

// Loads this file's kernels up front (see preload_* in internal.cuh).
template <class K>
static void touch_kernel(K k) {
    cudaFuncAttributes a;
    if (cudaFuncGetAttributes(&a, k) != cudaSuccess) cudaGetLastError();
}
void preload_segments() {
    touch_kernel(k_line_segments), touch_kernel(k_line_segments4), touch_kernel(k_line_segments_b), touch_kernel(k_enhance<false>), touch_kernel(k_enhance<true>);
    touch_kernel(k_scan_counts);
}

}  // namespace hrmgpu
