// connectMultiSlice's vertex matching on the device (reference src/planners/HRM3D.cpp:158-224).
//
// For every layer pair (near = the closest earlier layer, cur) the reference walks the vertices m0 of `near` in order and,
// for each, the vertices m1 of `cur` from a moving lower bound n22: the first m1 within two line spacings in x and y
// (:195-200) whose interpolated motion is collision free (isMultiSliceTransitionFree, :367-392) becomes an edge, n22 = m1.
// A layer's vertices are its free-segment mid points PLUS the bridge vertices connectOneSlice inserted between them
// (bridgeVertex, HighwayRoadMap-inl.h:263-326), so the host hands over its vertex positions (a few 10^3 x 24 bytes) and the id
// range of every layer; what the host no longer does is enumerate, upload and replay the ~10^5 candidate pairs:
//   k_pair_candidates<false>  counts, per (pair, m0) -- one warp each --, the vertices of `cur` inside the window (ascending m1)
//   k_scan_counts             (segments.cu) turns the counts into offsets
//   k_pair_candidates<true>   writes pair index, both positions and m1 of every candidate
//   k_transitions3d           (queries.cu, unchanged) tests all candidates of all pairs in one launch
//   k_pair_replay             one warp per pair replays the reference's loop with its moving lower bound on the answers
// The window test is the reference's comparison on the same doubles, so the candidate sets, and with them the roadmap, are
// identical to the host enumeration they replace (hrm_host.hpp, round 1).
#include "internal.cuh"

namespace hrmgpu {

// One warp per (pair, m0): the lanes take 32 vertices of `cur` at a time, a ballot counts / places the ones inside the window, so the
// candidates of m0 stay ascending in m1.  (One thread per m0 walking all of `cur` left the grid at a few blocks: 0.6 ms per
// ProbHRM3D::plan() for 12 slice pairs.)
constexpr int kCandWarps = 4;
template <bool FILL>
__global__ void __launch_bounds__(32 * kCandWarps) k_pair_candidates(const PairParams p) {
    const int pr = blockIdx.y, lane = threadIdx.x & 31;
    const long long n0 = p.m0_first[pr + 1] - p.m0_first[pr];
    const long long v = static_cast<long long>(blockIdx.x) * kCandWarps + (threadIdx.x >> 5);
    if (v >= n0) return;
    const long long g = p.m0_first[pr] + v;
    const int m0 = p.range[4 * pr] + static_cast<int>(v);
    const int c0 = p.range[4 * pr + 2], c1 = p.range[4 * pr + 3];
    const double x0 = p.vtx[3 * static_cast<size_t>(m0)], y0 = p.vtx[3 * static_cast<size_t>(m0) + 1], z0 = p.vtx[3 * static_cast<size_t>(m0) + 2];
    const unsigned lt = (1u << lane) - 1u;
    int n = 0;
    long long c = FILL ? p.cand_first[g] : 0;
    for (int base = c0; base < c1; base += 32) {
        const int m1 = base + lane;
        double x1 = 0.0, y1 = 0.0;
        bool in = false;
        if (m1 < c1) {
            x1 = __ldg(p.vtx + 3 * static_cast<size_t>(m1)), y1 = __ldg(p.vtx + 3 * static_cast<size_t>(m1) + 1);
            in = !(fabs(x0 - x1) > p.thx || fabs(y0 - y1) > p.thy);  // HRM3D.cpp:195-200
        }
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, in);
        if (FILL) {
            if (in) {
                const long long e = c + __popc(bal & lt);
                p.c_pair[e] = pr;
                p.c_v0[3 * e] = x0, p.c_v0[3 * e + 1] = y0, p.c_v0[3 * e + 2] = z0;
                p.c_v1[3 * e] = x1, p.c_v1[3 * e + 1] = y1, p.c_v1[3 * e + 2] = __ldg(p.vtx + 3 * static_cast<size_t>(m1) + 2);
                p.c_m1[e] = m1;
            }
            c += __popc(bal);
        } else {
            n += __popc(bal);
        }
    }
    if (!FILL && lane == 0) p.cnt[g] = n;
}

// One warp per pair: m0 in order; the lanes look at 32 candidates of m0 at a time (they are ascending in m1), the first one at or
// above the moving lower bound whose transition is free is the match (HRM3D.cpp:186-222).  The loop is serial in m0 (n22 moves), so
// it runs at load latency: the candidate ranges of 32 vertices are fetched at once and, when they fit (kReplayCap), all their
// candidates are staged in shared memory with coalesced loads; else the first 32 candidates of vertex m0 + 1 are in flight while
// vertex m0 is decided (they do not depend on n22).
constexpr int kReplayCap = 2048;
__global__ void __launch_bounds__(32) k_pair_replay(const long long* m0_first, const long long* cand_first, const int* c_m1, const unsigned char* free_flag,
                                                    int* match) {
    __shared__ int s_m1[kReplayCap];
    __shared__ unsigned char s_fl[kReplayCap];
    const int pr = blockIdx.x, lane = threadIdx.x;
    const unsigned full = 0xFFFFFFFFu;
    int n22 = 0;  // (vertex ids are non-negative: the first candidate of the pair is at or above it, like n22 = the layer's first id)
    const long long g_begin = m0_first[pr], g_end = m0_first[pr + 1];
    for (long long g0 = g_begin; g0 < g_end; g0 += 32) {
        const int nv = static_cast<int>((g_end - g0 < 32) ? g_end - g0 : 32);
        const long long cf = (lane <= nv - 1) ? cand_first[g0 + lane] : 0;
        const long long cf_last = cand_first[g0 + nv];
        const long long r0 = __shfl_sync(full, cf, 0);
        const bool staged = cf_last - r0 <= kReplayCap;
        if (staged) {
            __syncwarp();
            for (long long j = r0 + lane; j < cf_last; j += 32) s_m1[j - r0] = c_m1[j], s_fl[j - r0] = free_flag[j];
            __syncwarp();
        }
        // first chunk of vertex 0 of this block
        long long b0 = r0, e0 = (nv > 1) ? __shfl_sync(full, cf, 1) : cf_last;
        int m1 = -1;
        unsigned char fl = 0;
        if (!staged && b0 + lane < e0) m1 = c_m1[b0 + lane], fl = free_flag[b0 + lane];
        for (int t = 0; t < nv; ++t) {
            long long b1 = 0, e1 = 0;
            int m1n = -1;
            unsigned char fln = 0;
            if (t + 1 < nv) {
                b1 = __shfl_sync(full, cf, t + 1);
                e1 = (t + 2 < nv) ? __shfl_sync(full, cf, t + 2) : cf_last;
                if (!staged && b1 + lane < e1) m1n = c_m1[b1 + lane], fln = free_flag[b1 + lane];  // issued before this vertex is decided
            }
            int found = -1;
            for (long long j0 = b0; j0 < e0 && found < 0; j0 += 32) {
                const long long j = j0 + lane;
                if (staged) {
                    m1 = -1, fl = 0;
                    if (j < e0) m1 = s_m1[j - r0], fl = s_fl[j - r0];
                } else if (j0 != b0) {  // (more than 32 candidates and no match so far)
                    m1 = -1, fl = 0;
                    if (j < e0) m1 = c_m1[j], fl = free_flag[j];
                }
                const bool ok = m1 >= 0 && m1 >= n22 && fl != 0;
                const unsigned bal = __ballot_sync(full, ok);
                if (bal) found = __shfl_sync(full, m1, __ffs(bal) - 1);
            }
            if (lane == 0) match[g0 + t] = found;
            if (found >= 0) n22 = found;
            b0 = b1, e0 = e1, m1 = m1n, fl = fln;
        }
    }
}

// ---- connectOneSlice's candidate segments on the device (HighwayRoadMap-inl.h:218-326, HRM3D.cpp:93-156) --------------------
// Vertex pairs of a layer in the reference's enumeration order: first, plane by plane, every segment of line (ix, iy) with every
// segment of line (ix, iy + 1); then every segment of line (ix, iy) with every segment of line (ix + 1, iy).  Item = (layer, part,
// line) with part 0 = in-plane, 1 = cross-plane; its pairs are consecutive, so an exclusive scan over the item counts gives every
// pair its place.  Each pair expands into the 5 segments of bridgeVertex {v1-v2, v1-vNew1, vNew1-v2, v1-vNew2, vNew2-v2}
// (vNew1 = (x1, y1, z2), vNew2 = (x2, y2, z1), :295-326), query e = 5 * pair + t.
template <bool FILL>
__global__ void __launch_bounds__(128) k_slice_pairs(const long long* off, const double* xM, const double* txs, const double* tys, int K, int Lx, int Ly,
                                                     const long long* item_first, int* cnt, double* qa, double* qb, int* qlayer) {
    const int L = Lx * Ly;
    const long long item = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (item >= 2LL * K * L) return;
    const int k = static_cast<int>(item / (2 * L));
    const int rem = static_cast<int>(item - static_cast<long long>(k) * 2 * L);
    const int part = rem / L, l = rem - part * L;
    const int ix = l / Ly, iy = l - ix * Ly;
    const bool has = part == 0 ? (iy + 1 < Ly) : (ix + 1 < Lx);
    const int l2 = part == 0 ? l + 1 : l + Ly;
    const long long* o = off + static_cast<long long>(k) * L;
    const long long a0 = o[l], a1 = o[l + 1];
    const long long b0 = has ? o[l2] : 0, b1 = has ? o[l2 + 1] : 0;
    const long long n1 = a1 - a0, n2 = b1 - b0;
    if (!FILL) {
        cnt[item] = static_cast<int>(n1 * n2);
        return;
    }
    if (n1 * n2 == 0) return;
    const double x1 = txs[1 + ix], y1 = tys[1 + iy];
    const double x2 = part == 0 ? x1 : txs[2 + ix], y2 = part == 0 ? tys[2 + iy] : y1;
    long long pr = item_first[item];
    for (long long j1 = a0; j1 < a1; ++j1) {
        const double z1 = xM[j1];
        for (long long j2 = b0; j2 < b1; ++j2, ++pr) {
            const double z2 = xM[j2];
            const double v1[3] = {x1, y1, z1}, v2[3] = {x2, y2, z2}, w1[3] = {x1, y1, z2}, w2[3] = {x2, y2, z1};
            const double* as[5] = {v1, v1, w1, v1, w2};
            const double* bs[5] = {v2, w1, v2, w2, v2};
#pragma unroll
            for (int t = 0; t < 5; ++t) {
                const size_t e = static_cast<size_t>(5 * pr + t);
#pragma unroll
                for (int d = 0; d < 3; ++d) qa[3 * e + d] = as[t][d], qb[3 * e + d] = bs[t][d];
                qlayer[e] = k;
            }
        }
    }
}
// connection code of every pair from the 5 answers of its segments: 0 direct, 1 through vNew1, 2 through vNew2, 3 none (:295-326)
__global__ void __launch_bounds__(256) k_pair_codes(const unsigned char* free5, long long n_pairs, unsigned char* code) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const unsigned char* f = free5 + 5 * i;
    code[i] = f[0] ? 0 : ((f[1] && f[2]) ? 1 : ((f[3] && f[4]) ? 2 : 3));
}

int launch_slice_pairs(Ctx* c, bool fill, const long long* item_first, int* cnt, double* qa, double* qb, int* qlayer) {
    const long long items = 2LL * c->K * c->L;
    const unsigned grid = static_cast<unsigned>((items + 127) / 128);
    const double* xM = c->segdata.p + 2 * c->seg_cap;
    if (fill)
        k_slice_pairs<true><<<grid, 128, 0, c->stream>>>(c->seg_off.p, xM, c->txs.p, c->tys.p, c->K, c->Lx, c->Ly, item_first, cnt, qa, qb, qlayer);
    else
        k_slice_pairs<false><<<grid, 128, 0, c->stream>>>(c->seg_off.p, xM, c->txs.p, c->tys.p, c->K, c->Lx, c->Ly, item_first, cnt, qa, qb, qlayer);
    HRMGPU_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return HRMGPU_OK;
}
int launch_pair_codes(Ctx* c, const unsigned char* free5, long long n_pairs, unsigned char* code) {
    if (n_pairs <= 0) return HRMGPU_OK;
    k_pair_codes<<<static_cast<unsigned>((n_pairs + 255) / 256), 256, 0, c->stream>>>(free5, n_pairs, code);
    HRMGPU_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return HRMGPU_OK;
}

int launch_pair_candidates(Ctx* c, const PairParams& p, long long max_n0, bool fill) {
    const dim3 grid(static_cast<unsigned>((max_n0 + kCandWarps - 1) / kCandWarps), static_cast<unsigned>(p.P));
    if (fill)
        k_pair_candidates<true><<<grid, 32 * kCandWarps, 0, c->stream>>>(p);
    else
        k_pair_candidates<false><<<grid, 32 * kCandWarps, 0, c->stream>>>(p);
    HRMGPU_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return HRMGPU_OK;
}

int launch_pair_replay(Ctx* c, int P, const long long* m0_first, const long long* cand_first, const int* c_m1, const unsigned char* free_flag, int* match) {
    k_pair_replay<<<static_cast<unsigned>(P), 32, 0, c->stream>>>(m0_first, cand_first, c_m1, free_flag, match);
    HRMGPU_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return HRMGPU_OK;
}

void preload_connect() {
    cudaFuncAttributes a;
    if (cudaFuncGetAttributes(&a, k_pair_candidates<false>) != cudaSuccess) cudaGetLastError();
    if (cudaFuncGetAttributes(&a, k_pair_candidates<true>) != cudaSuccess) cudaGetLastError();
    if (cudaFuncGetAttributes(&a, k_pair_replay) != cudaSuccess) cudaGetLastError();
    if (cudaFuncGetAttributes(&a, k_slice_pairs<false>) != cudaSuccess) cudaGetLastError();
    if (cudaFuncGetAttributes(&a, k_slice_pairs<true>) != cudaSuccess) cudaGetLastError();
    if (cudaFuncGetAttributes(&a, k_pair_codes) != cudaSuccess) cudaGetLastError();
}

}  // namespace hrmgpu
