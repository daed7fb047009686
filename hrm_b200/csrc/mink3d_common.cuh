// Shared by the two fused Minkowski + sweep kernels (mink3d.cu: one CTA per surface; mink3d_ws.cu: persistent,
// warp-specialised): launch parameters, the exact sweep-line classification of a vertex, small shared-memory helpers.
#pragma once
#include "devmath.cuh"
#include "internal.cuh"

#include <cfloat>
#include <cmath>
#include <type_traits>

namespace hrmgpu {

struct MinkParams {
    const Obj3* obj;       // [S0]
    const double* tab;     // [S0][8][n]
    const double* semi;    // [B][3]
    const double* R;       // [K][B][9]
    const double* off;     // [K][B][3]
    const double* txs;     // [Lx+2] sentinel padded
    const double* tys;     // [Ly+2]
    void* mesh;            // [K][So][3][M] Real
    double* lo;            // [K][So][L]
    double* up;
    unsigned long long* counters;
    long long* dbg;        // optional [grid][8] phase timestamps (profiling builds of tools/phase_times.py)
    int n, M, B, first_obj, n_objs, n_heavy;
    int Lx, Ly, L;
    int lpc, groups;                        // thread mapping
    int paired;                             // phase A handles rows in adjacent pairs (n even, lpc divides n/2)
    int seg_cap;                            // kept cells a warp can list per scan round (kCellCap / warps; smaller in tests)
    int semi_per_layer;                     // bridge builds: every layer (pair) has its own body semi-axes, semi [K][B][3]
    int reload;                             // re-sweep: the boundary points are read back from `mesh` instead of being computed
    // byte offsets of the shared-memory regions (host-computed, see smem_layout): read from the constant bank instead of
    // being re-derived from n, Lx, Ly wherever the compiler rematerialises a pointer
    int o_mats, o_cc, o_txs, o_tys, o_queue, o_vcode, o_m0, o_m1, o_qcount;
    double inv_dx, bias_x, inv_dy, bias_y;  // line-index guess g = x*inv_d + bias
    double bias_xq, bias_yq;                // bias + 1.5 * 2^36 + (1 - 2^-16): Q15.16 conversion and ceil folded into the FMA
    double zlow, zup;
};

// mink3d_ws.cu
bool ws_supported(const Ctx* c, int precision, bool do_sweep, bool reload, int lpc, int groups, int paired);
int launch_minksweep3d_ws(Ctx* c, const MinkParams& base, int precision, long long items, int* d_bcnt, double* d_bucket, int bcap);
void preload_mink3d_ws();

static_assert(kMaxGrid * kMaxGrid <= 65536, "cell ids are listed as 16-bit values");

// Exact classification of a coordinate against the sweep grid: smallest line index i in [0, Lc] with
// t_i >= x (Lc if none) and whether t_i == x, packed as 2*i + eq.  ts is sentinel padded:
// ts[0] = -inf, ts[1+i] = t_i, ts[Lc+1] = +inf.  Fast path: ceil of the affine guess when it is provably
// not within rounding distance of a grid line; otherwise an exact search on the t_i the host computed with
// the reference's formula (HRM3D.cpp:66-81).
static __device__ __noinline__ uint32_t line_code_slow(double x, const double* ts, int Lc, int i) {
    i = min(max(i, 0), Lc);
    while (i > 0 && ts[i] >= x) --i;
    while (ts[i + 1] < x) ++i;
    const uint32_t eq = (ts[i + 1] == x) ? 1u : 0u;
    return 2u * static_cast<uint32_t>(i) + eq;
}
// Fast part only: returns the code and sets ok = false when the exact search is required.
// g = (x - t_0) / d in line units is converted to Q11.20 fixed point by the magic-number addition (1.5 * 2^32: the low
// word of the sum is rint(g * 2^20)); ceil, the "within 2^-17 of a grid line" test and the range test are then integer
// operations, which keeps the FP64 pipe free for the boundary arithmetic.
__device__ __forceinline__ uint32_t line_code_fast(double x, int Lc, double inv_d, double bias, bool& ok) {
    const double g = fma(x, inv_d, bias);
    const double t = g + 6442450944.0;  // 1.5 * 2^32
    const int fx = __double2loint(t);
    const int i = (fx + 0xFFFFF) >> 20;                                   // ceil(g)
    const unsigned fr = static_cast<unsigned>(fx + 8) & 0xFFFFFu;         // < 16  <=>  |g - rint(g)| <= 8 * 2^-20
    const bool in_range = (__double2hiint(g) & 0x7FFFFFFF) < 0x40900000;  // |g| < 1024 (also rejects NaN / inf)
    ok = (fr >= 16u) && in_range;
    return 2u * static_cast<uint32_t>(__vimin_s32_relu(i, Lc));  // clamp to [0, Lc]
}
__device__ __forceinline__ uint32_t line_code(double x, const double* ts, int Lc, double inv_d, double bias) {
    bool ok;
    const uint32_t c = line_code_fast(x, Lc, inv_d, bias, ok);
    if (!ok) {
        const double g = fma(x, inv_d, bias);
        return line_code_slow(x, ts, Lc, (g > 0.0) ? ((g < 1000.0) ? static_cast<int>(c >> 1) : Lc) : 0);
    }
    return c;
}


// 32-bit shared-window addresses and explicit paired loads/stores: phase A keeps running addresses in registers and
// bumps them per column instead of letting the compiler re-derive the shared-memory carve-up inside the loop.
__device__ __forceinline__ uint32_t smem_u32(const void* ptr) { return static_cast<uint32_t>(__cvta_generic_to_shared(ptr)); }
__device__ __forceinline__ void lds_pair(uint32_t addr, double& a, double& b) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
__device__ __forceinline__ void lds_pair(uint32_t addr, float& a, float& b) {
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "r"(addr));
}
__device__ __forceinline__ void sts_codes(uint32_t addr, uint16_t c0, uint16_t c1) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(static_cast<uint32_t>(c0) | (static_cast<uint32_t>(c1) << 16)) : "memory");
}
__device__ __forceinline__ void sts_codes(uint32_t addr, uint32_t c0, uint32_t c1) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(c0), "r"(c1) : "memory");
}
// Fixed-point line index of the paired fast path.  t = fma(x, inv_d, bias_q) with bias_q = bias + 1.5 * 2^36 + (1 - 2^-16):
// the low word of t is w = rint(g * 2^16) + 0xFFFF (two's complement, g in line units), so its upper half is ceil(g);
// that holds while the high word of t is 0x42380000 with w's sign bit clear (0 <= g < 2^15) or 0x4237FFFF with it set
// (-2^15 <= g < 0): the signed 16-bit SIMD clamp that follows needs |g| < 2^15, not just the two high words (which alone
// would admit |g| < 2^16).  ok = false when g is within 4 * 2^-16 of a grid line (the computed w is off by at most one
// unit) or out of range: exact search instead.
__device__ __forceinline__ uint32_t line_q(double x, double inv_d, double bias_q, bool& ok) {
    const double t = fma(x, inv_d, bias_q);
    const uint32_t w = static_cast<uint32_t>(__double2loint(t));
    // (high word << 1) | sign bit of w is 0x84700000 for 0 <= g < 2^15 and 0x846FFFFF for -2^15 <= g < 0: one funnel shift
    // tests the range and the sign agreement at once
    const uint32_t u = __funnelshift_l(w, static_cast<uint32_t>(__double2hiint(t)), 1) - 0x846FFFFFu;
    ok = (((w + 5u) & 0xFFF8u) != 0u) && (u < 2u);
    return w;
}

template <int MODE>
struct RealOf {
    using type = double;
};
template <>
struct RealOf<HRMGPU_F32> {
    using type = float;
};

// Pair of Reals for 2-wide shared-memory loads of the per-column constants.
template <class Real>
struct alignas(2 * sizeof(Real)) Real2 {
    Real a, b;
};

// Fast modes split every boundary term into a row (omega) factor and a column (eta) factor:
//   x0 - off = ce1_c * V_r + C_c,   V_r = R1[:,0] a cw1_r + R1[:,1] b sw1_r,        C_c = R1[:,2] c se1_c + (p - off)
//   u        = ce2_c * U_r + UZ_c,  U_r = M1[:,0] cw2_r / a + M1[:,1] sw2_r / b,    UZ_c = M1[:,2] se2_c / c
//   w        = ce2_c * W_r + WZ_c,  W_r, WZ_c likewise with M2
// so a point costs 9 FMAs + the normalisation instead of 18.  A thread keeps V, U, W of its rows in registers; the column
// table holds {ce1, ce2, C, UZ, WZ} (11 values + pad), the 18 row-base coefficients live in `rowb`.
constexpr int kColStride = 12;


}  // namespace hrmgpu
