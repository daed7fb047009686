// connectMultiSlice's vertex matching on the device (reference src/planners/HRM3D.cpp:158-224).
//
// For every layer pair (near = the closest earlier layer, cur) the reference walks the vertices m0 of `near` in order and,
// for each, the vertices m1 of `cur` from a moving lower bound n22: the first m1 within two line spacings in x and y
// (:195-200) whose interpolated motion is collision free (isMultiSliceTransitionFree, :367-392) becomes an edge, n22 = m1.
// A layer's vertices are its free-segment mid points PLUS the bridge vertices connectOneSlice inserted between them
// (bridgeVertex, HighwayRoadMap-inl.h:263-326), so the host hands over its vertex positions (a few 10^3 x 24 bytes) and the id
// range of every layer; what the host no longer does is enumerate, upload and replay the ~10^5 candidate pairs:
//   k_pair_candidates<false>  counts, per (pair, m0), the vertices of `cur` inside the window (ascending m1)
//   k_scan_counts             (segments.cu) turns the counts into offsets
//   k_pair_candidates<true>   writes pair index, both positions and m1 of every candidate
//   k_transitions3d           (queries.cu, unchanged) tests all candidates of all pairs in one launch
//   k_pair_replay             one warp per pair replays the reference's loop with its moving lower bound on the answers
// The window test is the reference's comparison on the same doubles, so the candidate sets, and with them the roadmap, are
// identical to the host enumeration they replace (hrm_host.hpp, round 1).
#include "internal.cuh"

namespace hrmgpu {

template <bool FILL>
__global__ void __launch_bounds__(128) k_pair_candidates(const PairParams p) {
    const int pr = blockIdx.y;
    const long long n0 = p.m0_first[pr + 1] - p.m0_first[pr];
    const long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (v >= n0) return;
    const long long g = p.m0_first[pr] + v;
    const int m0 = p.range[4 * pr] + static_cast<int>(v);
    const int c0 = p.range[4 * pr + 2], c1 = p.range[4 * pr + 3];
    const double x0 = p.vtx[3 * static_cast<size_t>(m0)], y0 = p.vtx[3 * static_cast<size_t>(m0) + 1];
    if (!FILL) {
        int n = 0;
        for (int m1 = c0; m1 < c1; ++m1) {
            const double x1 = __ldg(p.vtx + 3 * static_cast<size_t>(m1)), y1 = __ldg(p.vtx + 3 * static_cast<size_t>(m1) + 1);
            n += (fabs(x0 - x1) > p.thx || fabs(y0 - y1) > p.thy) ? 0 : 1;  // HRM3D.cpp:195-200
        }
        p.cnt[g] = n;
    } else {
        long long c = p.cand_first[g];
        const double z0 = p.vtx[3 * static_cast<size_t>(m0) + 2];
        for (int m1 = c0; m1 < c1; ++m1) {
            const double x1 = __ldg(p.vtx + 3 * static_cast<size_t>(m1)), y1 = __ldg(p.vtx + 3 * static_cast<size_t>(m1) + 1);
            if (fabs(x0 - x1) > p.thx || fabs(y0 - y1) > p.thy) continue;
            p.c_pair[c] = pr;
            p.c_v0[3 * c] = x0, p.c_v0[3 * c + 1] = y0, p.c_v0[3 * c + 2] = z0;
            p.c_v1[3 * c] = x1, p.c_v1[3 * c + 1] = y1, p.c_v1[3 * c + 2] = __ldg(p.vtx + 3 * static_cast<size_t>(m1) + 2);
            p.c_m1[c] = m1;
            ++c;
        }
    }
}

// One warp per pair: m0 in order; the lanes look at 32 candidates of m0 at a time (they are ascending in m1), the first one at or
// above the moving lower bound whose transition is free is the match (HRM3D.cpp:186-222).
__global__ void __launch_bounds__(32) k_pair_replay(const long long* m0_first, const long long* cand_first, const int* c_m1, const unsigned char* free_flag,
                                                    int* match) {
    const int pr = blockIdx.x, lane = threadIdx.x;
    int n22 = 0;  // (vertex ids are non-negative: the first candidate of the pair is at or above it, like n22 = the layer's first id)
    for (long long g = m0_first[pr]; g < m0_first[pr + 1]; ++g) {
        int found = -1;
        for (long long j0 = cand_first[g]; j0 < cand_first[g + 1] && found < 0; j0 += 32) {
            const long long j = j0 + lane;
            int m1 = -1;
            bool ok = false;
            if (j < cand_first[g + 1]) {
                m1 = c_m1[j];
                ok = m1 >= n22 && free_flag[j] != 0;
            }
            const unsigned b = __ballot_sync(0xFFFFFFFFu, ok);
            if (b) found = __shfl_sync(0xFFFFFFFFu, m1, __ffs(b) - 1);
        }
        if (lane == 0) match[g] = found;
        if (found >= 0) n22 = found;
    }
}

int launch_pair_candidates(Ctx* c, const PairParams& p, long long max_n0, bool fill) {
    const dim3 grid(static_cast<unsigned>((max_n0 + 127) / 128), static_cast<unsigned>(p.P));
    if (fill)
        k_pair_candidates<true><<<grid, 128, 0, c->stream>>>(p);
    else
        k_pair_candidates<false><<<grid, 128, 0, c->stream>>>(p);
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
}

}  // namespace hrmgpu
