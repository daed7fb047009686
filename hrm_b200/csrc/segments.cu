// K3: per-sweep-line interval algebra -> collision-free segments, then enhanceFreeSegment.
//
// Replaces (reference):
//   FreeSpaceComputator::computeSweepLineFreeSegment   include/hrm/datastructure/FreeSpace-inl.h:101-129
//   Interval::unions / intersects / complements        src/datastructure/Interval.cpp:10-100
//   FreeSpaceComputator::computeFreeSegment            FreeSpace-inl.h:63-97
//   FreeSpaceComputator::enhanceFreeSegment            FreeSpace-inl.h:132-174
//
// One WARP per sweep line.  Pass 1 (k_line_segments): compact the non-NaN C-obstacle intervals of the
// line into shared memory, sort by start with an arbitrary-n bitonic network, prefix-max the ends with
// shuffles, and emit the gaps clipped to the arena interval ("original" segments).  Pass 2
// (k_enhance_count) counts the degenerate segments enhanceFreeSegment appends from the two adjacent
// lines of the same x-plane; a scan turns the counts into CSR offsets; pass 3 (k_enhance_fill) writes
// originals + appended values in the reference's (j1, j2) loop order and sorts xL, xU, xM of every line
// but the last one of a plane independently, exactly like the reference.
#include "internal.cuh"

#include <cfloat>

namespace hrmgpu {

constexpr int kSegWarps = 8;

struct SegParams {
    const double* lo;   // [K][S][L]
    const double* up;
    double* orig;       // [K][L][cap][2]
    int* orig_cnt;      // [K][L]
    int* seg_cnt;       // [K][L]
    const long long* seg_off;  // [K*L+1]
    double* segL;
    double* segU;
    double* segM;
    int K, S, Sa, L, Lx, Ly, cap;  // cap = So + 1 original segments per line at most
    int warps;                     // warps per CTA in pass 1
    int sort_cap;                  // intervals per warp in shared memory
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

// Pass 1: original free segments of every line.  grid = ceil(K*L / warps) CTAs of `warps` warps.
__global__ void __launch_bounds__(kSegWarps * 32) k_line_segments(const SegParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long gl = static_cast<long long>(blockIdx.x) * p.warps + warp;  // global line id = k*L + l
    if (warp >= p.warps || gl >= static_cast<long long>(p.K) * p.L) return;
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
    // (four chunks of 32 surfaces are requested before the first one is consumed: the loads are strided by L doubles,
    // so the loop is bound by memory latency, not bandwidth)
    int m = 0;
    constexpr int kUnroll = 4;
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

    double* out = p.orig + (static_cast<size_t>(gl) * p.cap) * 2;
    int cnt = 0;
    if (!arena_ok) {
        cnt = 0;  // complements(): empty outer -> nothing
    } else if (m == 0) {
        if (lane == 0) out[0] = as, out[1] = ae;  // complements(): empty inner -> outer
        cnt = 1;
    } else {
        warp_bitonic<true>(ks, ke, m, lane);
        // gaps of the union: before interval i iff (running max of earlier ends) < start_i   (Interval.cpp:21-29)
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
            bool keep = false;
            double gs = 0, ge = 0;
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
            if (keep) {
                const int pos = cnt + __popc(bal & ((1u << lane) - 1u));
                out[2 * pos] = gs;
                out[2 * pos + 1] = ge;
            }
            cnt += __popc(bal);
            carry = fmax(carry, __shfl_sync(0xFFFFFFFFu, inc, 31));
        }
    }
    if (lane == 0) p.orig_cnt[gl] = cnt;
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
    const int n0 = p.orig_cnt[gl];
    const double* s0 = p.orig + static_cast<size_t>(gl) * p.cap * 2;
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
        const double* sp = p.orig + static_cast<size_t>(gl - 1) * p.cap * 2;
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
        const double* sn = p.orig + static_cast<size_t>(gl + 1) * p.cap * 2;
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

// Exclusive scan of K*L counts into 64-bit CSR offsets; single CTA (K*L is at most a few 10^5..10^6).
__global__ void __launch_bounds__(1024) k_scan_counts(const int* cnt, long long* off, long long n) {
    __shared__ long long wsum[32];
    __shared__ long long carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (long long base = 0; base < n; base += 1024) {
        const long long i = base + tid;
        long long v = (i < n) ? cnt[i] : 0;
        long long inc = v;
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
        }
        __syncthreads();
        const long long prefix = carry_s + (warp > 0 ? wsum[warp - 1] : 0) + inc - v;
        if (i < n) off[i] = prefix;
        __syncthreads();
        if (tid == 1023) carry_s = prefix + v;
        __syncthreads();
    }
    if (tid == 0) off[n] = carry_s;
}

int launch_segments(Ctx* c) {
    SegParams p;
    p.lo = c->ivl_lo.p;
    p.up = c->ivl_up.p;
    p.K = c->K;
    p.S = c->S;
    p.Sa = c->Sa;
    p.L = c->L;
    p.Lx = (c->dim == 3) ? c->Lx : 1;
    p.Ly = c->Ly;
    const int So = c->S - c->Sa;
    p.cap = So + 1;
    const long long nl = static_cast<long long>(c->K) * c->L;
    if (nl == 0) {
        c->seg_total = 0;
        c->h_seg_off.assign(1, 0);
        return HRMGPU_OK;
    }
    int rc;
    if ((rc = ensure(c, c->orig, static_cast<size_t>(nl) * p.cap * 2))) return rc;
    if ((rc = ensure(c, c->orig_cnt, static_cast<size_t>(nl)))) return rc;
    if ((rc = ensure(c, c->seg_cnt, static_cast<size_t>(nl)))) return rc;
    if ((rc = ensure(c, c->seg_off, static_cast<size_t>(nl) + 1))) return rc;
    p.orig = c->orig.p;
    p.orig_cnt = c->orig_cnt.p;
    p.seg_cnt = c->seg_cnt.p;
    p.seg_off = c->seg_off.p;
    p.segL = p.segU = p.segM = nullptr;

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
    k_line_segments<<<static_cast<unsigned>(g1), warps * 32, smem, c->stream>>>(p);
    HRMGPU_CUDA(c, cudaGetLastError());
    const long long g2 = (nl + kSegWarps - 1) / kSegWarps;
    k_enhance<false><<<static_cast<unsigned>(g2), kSegWarps * 32, 0, c->stream>>>(p);
    HRMGPU_CUDA(c, cudaGetLastError());
    k_scan_counts<<<1, 1024, 0, c->stream>>>(c->seg_cnt.p, c->seg_off.p, nl);
    HRMGPU_CUDA(c, cudaGetLastError());
    c->launches += 3;
    // the payload size is data dependent: read the offsets back (they are an API output anyway)
    c->h_seg_off.resize(static_cast<size_t>(nl) + 1);
    HRMGPU_CUDA(c, cudaMemcpyAsync(c->h_seg_off.data(), c->seg_off.p, sizeof(long long) * (nl + 1), cudaMemcpyDeviceToHost, c->stream));
    HRMGPU_CUDA(c, cudaStreamSynchronize(c->stream));
    c->seg_total = c->h_seg_off[static_cast<size_t>(nl)];
    const size_t tot = static_cast<size_t>(c->seg_total);
    if ((rc = ensure(c, c->segL, tot))) return rc;
    if ((rc = ensure(c, c->segU, tot))) return rc;
    if ((rc = ensure(c, c->segM, tot))) return rc;
    p.segL = c->segL.p;
    p.segU = c->segU.p;
    p.segM = c->segM.p;
    k_enhance<true><<<static_cast<unsigned>(g2), kSegWarps * 32, 0, c->stream>>>(p);
    HRMGPU_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return HRMGPU_OK;
}

}  // namespace hrmgpu
