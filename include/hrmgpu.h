/* hrmgpu.h — C ABI of the B200-native free-space construction for the Highway RoadMap planner.
 *
 * This is the drop-in boundary behind the reference's FreeSpaceComputator seam
 * (ChirikjianLab/hrm include/hrm/datastructure/FreeSpace.h:45-128, FreeSpace3D.h:29-55) and the
 * planner-side edge / bridge tests (src/planners/HRM3D.cpp:301-425, HRM2D.cpp:199-282).  Plain C:
 * pointers and sizes only, no C++/torch types.  Every entry point names the reference interface
 * it replaces.  INTEGRATION.md shows the reference-side C++ binding.
 *
 * Conventions
 *  - All functions return HRMGPU_OK (0) or a negative hrmgpu_status; they never throw or abort.
 *    hrmgpu_last_error() returns a human-readable message for the last failure on that context.
 *  - A context is bound to one CUDA device and is NOT thread-safe (like the reference objects).
 *  - Caller owns host buffers; the context owns all device buffers.  Results of a build call stay
 *    resident in HBM until the next build call of the same kind or hrmgpu_destroy().
 *  - Streams.  A context has a MAIN stream (its own, or the one installed with hrmgpu_set_stream) and an internal SIDE
 *    stream.  hrmgpu_build_layers* / hrmgpu_resweep only QUEUE work and return: the fused Minkowski + sweep kernel on the
 *    main stream, the free-segment passes (and hrmgpu_allgather_layers, hrmgpu_fetch_segments_async) on the side stream
 *    behind it, so they overlap the fused kernel of the NEXT build (interval tables are double-buffered).  Every getter
 *    waits for what it returns; hrmgpu_join() makes the main stream wait for the side stream on the device (no host
 *    sync); hrmgpu_synchronize() waits for both on the host.
 *  - Surfaces are ordered j = object * B + body, arena objects first, then obstacles
 *    (FreeSpace-inl.h:42-60).  Sa = n_arena*B, So = n_obs*B, S = Sa+So.
 *  - 3D sweep lines are ordered l = ix * Ly + iy (HRM3D.cpp:63-91);  2D: l = iy.
 *  - Rotations are 3x3 row-major, exactly the matrix Eigen's Quaterniond::toRotationMatrix() gives
 *    for the (un-normalised) quaternion the reference holds (SuperQuadrics.cpp:78,109-110).
 *  - Boundary points come back in the reference's BoundaryPoints layout: 3 x M (2 x N) column-major,
 *    point index k = r + c*n with the omega index r fastest (SuperQuadrics.cpp:63-72).
 */
#ifndef HRMGPU_H
#define HRMGPU_H

#ifdef __cplusplus
extern "C" {
#endif

#define HRMGPU_ABI_VERSION 2

typedef struct hrmgpu_ctx hrmgpu_ctx;

typedef enum {
    HRMGPU_OK = 0,
    HRMGPU_E_ARG = -1,     /* invalid argument */
    HRMGPU_E_CUDA = -2,    /* CUDA runtime error (message in hrmgpu_last_error) */
    HRMGPU_E_STATE = -3,   /* call order violated (e.g. build before set_scene) */
    HRMGPU_E_CAP = -4,     /* a documented capacity limit was exceeded */
    HRMGPU_E_NOMEM = -5,   /* device or host allocation failed */
    HRMGPU_E_NODEVICE = -6 /* no usable CUDA device: there is NO CPU fallback */
} hrmgpu_status;

/* Arithmetic mode of the Minkowski + sweep kernels. */
typedef enum {
    HRMGPU_F64_STRICT = 0, /* IEEE op-for-op replica of the reference's FP64 arithmetic (no FMA contraction,
                              true divisions): bit-identical boundary points, identical hit decisions */
    HRMGPU_F64 = 1,        /* FP64 with FMA and hoisted column constants: <=1e-12 relative on points */
    HRMGPU_F32 = 2         /* FP32 throughput variant: <=1e-5 relative (to surface scale) on points */
} hrmgpu_precision;

/* ---- lifetime ------------------------------------------------------------------------------ */
int hrmgpu_abi_version(void);
int hrmgpu_device_count(void);
/* One context per GPU.  device = CUDA ordinal.  flags: reserved, pass 0. */
int hrmgpu_create(hrmgpu_ctx** out, int device, unsigned flags);
void hrmgpu_destroy(hrmgpu_ctx* ctx);
/* Back to the state right after hrmgpu_create -- no scene, no sweep grid, no layers, no bridge layers, counters zero, the
 * context's own stream -- but with its streams, events and device buffers kept: a planner object that is created per query
 * (the reference builds a FreeSpace3D per planner) can hand its context to the next one instead of paying for creation. */
int hrmgpu_reset(hrmgpu_ctx* ctx);
const char* hrmgpu_last_error(const hrmgpu_ctx* ctx);
/* Run all subsequent work of this context on an existing CUDA stream (cudaStream_t as void*);
 * NULL = the context's own stream.  Lets a host framework time / order the work with its own events. */
int hrmgpu_set_stream(hrmgpu_ctx* ctx, void* cuda_stream);
int hrmgpu_synchronize(hrmgpu_ctx* ctx);
/* Device-side join: work queued on the main stream after this call runs after everything queued so far on the side stream
 * (free-segment passes, exchange).  No host synchronisation.  Use it before recording an event that must cover a whole build. */
int hrmgpu_join(hrmgpu_ctx* ctx);

/* ---- scene --------------------------------------------------------------------------------- */
/* Replaces: the arena_/obstacle_ vectors a FreeSpace3D is constructed with (FreeSpace3D.cpp:6-10) and
 * the per-object parameter grids of SuperQuadrics::SuperQuadrics (SuperQuadrics.cpp:8-26).
 * sq: [n_arena + n_obs][17] = a,b,c, eps1,eps2, px,py,pz, R1[9] row-major.  n_grid = num_ (points per
 * parameter, M = n_grid^2).  Limits: 3 <= n_grid <= 200. */
int hrmgpu_set_scene3d(hrmgpu_ctx* ctx, int n_arena, int n_obs, const double* sq, int n_grid);
/* 2D twin (FreeSpace2D.cpp:6-10, SuperEllipse.cpp:10-18). se: [n_arena+n_obs][6] = a,b,eps,px,py,angle. */
int hrmgpu_set_scene2d(hrmgpu_ctx* ctx, int n_arena, int n_obs, const double* se, int num);
/* Replaces: FreeSpaceComputator::setup(numLine, lowBound, upBound) (FreeSpace-inl.h:20-39) plus the
 * sweep-line grids of HRM3D::sweepLineProcess (HRM3D.cpp:63-91) / HRM2D::sweepLineProcess (HRM2D.cpp:53-70).
 * lim = boundaryLimits {xlo,xhi,ylo,yhi,zlo,zhi} (2D: first four; Lx ignored, pass 1).
 * 3D requires Lx >= 2 and Ly >= 2 (with one line the reference's spacing is a division by zero). */
int hrmgpu_set_sweep(hrmgpu_ctx* ctx, const double* lim, int Lx, int Ly);

/* ---- C-layer construction (K1 + K2 + K3) ---------------------------------------------------- */
/* Replaces, for K layers in one call: computeCSpaceBoundary() (FreeSpace-inl.h:42-60 ->
 * MultiBodyTree3D::minkSum, MultiBodyTree3D.cpp:91-105 -> SuperQuadrics::getMinkSum3D,
 * SuperQuadrics.cpp:84-143), computeCSpaceBoundaryMesh() (FreeSpace3D.cpp:14-27),
 * computeIntersectionInterval() for every x-plane (FreeSpace3D.cpp:29-68 ->
 * intersectVerticalLineMesh3D, LineIntersection.cpp:56-171) and computeFreeSegment()
 * (FreeSpace-inl.h:63-174, Interval.cpp).
 *   body_semi [B][3]   semi-axes of the ellipsoidal robot bodies (base first)
 *   body_R    [K][B][9] per-layer world rotation of every body (row-major)
 *   body_off  [K][B][3] per-layer offset subtracted from the body's boundary (link centre; zeros for
 *                       the base) (MultiBodyTree3D.cpp:98-101)
 * Feeding the as-loaded pose for every k reproduces the reference snapshot (SURVEY.md finding 1);
 * feeding R(q_k)*R_link gives the per-orientation C-layers.  Host pointers; copied H2D inside. */
int hrmgpu_build_layers(hrmgpu_ctx* ctx, int K, int B, const double* body_semi, const double* body_R,
                        const double* body_off, int precision);
/* Same, with the three arrays already resident in device memory (device pointers). */
int hrmgpu_build_layers_dev(hrmgpu_ctx* ctx, int K, int B, const double* d_body_semi, const double* d_body_R,
                            const double* d_body_off, int precision);
/* 2D twin: computeCSpaceBoundary() -> MultiBodyTree2D::minkSum (MultiBodyTree2D.cpp:43-65) ->
 * SuperEllipse::getMinkSum2D (SuperEllipse.cpp:58-101); FreeSpace2D::computeIntersectionInterval
 * (FreeSpace2D.cpp:14-50); computeFreeSegment.  body_semi [B][2], body_angle [K][B], body_off [K][B][2]. */
int hrmgpu_build_layers2d(hrmgpu_ctx* ctx, int K, int B, const double* body_semi, const double* body_angle,
                          const double* body_off, int precision);
/* Refinement: HighwayRoadMap::refineExistRoadmap (include/hrm/planners/HighwayRoadMap-inl.h:178-215) doubles the sweep
 * lines and HRM3D/HRM2D::constructOneSlice (src/planners/HRM3D.cpp:24-54) reuses the cached C-boundaries
 * (sliceBoundAll_, sliceBoundMeshAll_).  Re-runs the sweep and the free-segment passes of ALL layers of the last build
 * on their meshes resident in device memory for a new grid; boundaries, edge and bridge tests are unaffected.
 * Interval tables are sized by the new line count (the reference overruns its tables here, SURVEY.md finding 6). */
int hrmgpu_resweep(hrmgpu_ctx* ctx, const double* boundaryLimits, int numLineX, int numLineY);

/* ---- results of the last build -------------------------------------------------------------- */
/* Sizes of the last build: out[0]=K, [1]=B, [2]=S, [3]=Sa, [4]=L, [5]=M (points per surface),
 * [6]=total segment count over all layers, [7]=dimension (2|3). */
int hrmgpu_get_dims(hrmgpu_ctx* ctx, long long* out8);
/* Replaces getCSpaceBoundary() (FreeSpace.h:70): boundary of one surface of layer k, dim x M column-major. */
int hrmgpu_get_boundary(hrmgpu_ctx* ctx, int k, int surface, double* xyz);
/* getCSpaceBoundary() for ALL surfaces of layer k in one call (what HRM3D::constructOneSlice copies into sliceBoundAll_,
 * HRM3D.cpp:36-42): xyz [S][M][dim] doubles, i.e. S consecutive dim x M column-major BoundaryPoints, arena surfaces first.
 * One device transpose + one device-to-host copy. */
int hrmgpu_get_boundary_layer(hrmgpu_ctx* ctx, int k, double* xyz);
/* Replaces getIntersectionInterval() (FreeSpace.h:74): lo/up [L][S] of layer k; NaN = no hit. */
int hrmgpu_get_intervals(hrmgpu_ctx* ctx, int k, double* lo, double* up);
/* Replaces getFreeSegment() (FreeSpace.h:80) for all lines of layer k: CSR offsets off[L+1] and
 * xL/xU/xM[cap].  Two-call convention: with xL==NULL only off[] (if non-NULL) is written.
 * Returns the layer's segment count (>= 0) or a negative status. */
int hrmgpu_get_segments(hrmgpu_ctx* ctx, int k, int* off, double* xL, double* xU, double* xM, int cap);
/* All layers at once: off [K*L+1] (global CSR, 64-bit), payload arrays sized out8[6]. */
int hrmgpu_get_segments_all(hrmgpu_ctx* ctx, long long* off, double* xL, double* xU, double* xM, long long cap);
/* Pipelined read-back of getFreeSegment() for all layers: queues, on the side stream behind the free-segment passes of the
 * last build, the copies of off [K*L+1] and of the first min(cap, internal capacity) entries of xL / xU / xM into the caller's
 * (pinned, for real overlap) host arrays, and returns at once (the number of entries per array it will copy, or a negative
 * status) -- the next hrmgpu_build_layers* may be issued immediately.
 * hrmgpu_wait_segments() blocks until those copies are done and returns the segment count (entries past it are unspecified),
 * or HRMGPU_E_CAP when the build held more segments than `cap` / the internal capacities: then read that build with
 * hrmgpu_get_segments_all(), which grows them (only possible before the next build call). */
long long hrmgpu_fetch_segments_async(hrmgpu_ctx* ctx, long long* off, double* xL, double* xU, double* xM, long long cap);
long long hrmgpu_wait_segments(hrmgpu_ctx* ctx);
/* Count of (line, surface) pairs that saw exactly ONE mesh hit in the last build.  The reference reads
 * hit[1] out of bounds there (FreeSpace3D.cpp:44-49,61-64); this library defines it as "no interval". */
int hrmgpu_get_single_hit_count(hrmgpu_ctx* ctx, long long* out);

/* ---- intra-layer edge tests (K4) ------------------------------------------------------------- */
/* Replaces HRM3D::isSameSliceTransitionFree (HRM3D.cpp:319-365 -> intersectLineMesh3D,
 * LineIntersection.cpp:7-54) against the C-obstacle meshes of layer k of the last 3D build, and
 * HRM2D::isSameSliceTransitionFree (HRM2D.cpp:215-232 -> isIntersectSegPolygon2D,
 * LineIntersection.cpp:173-213) for a 2D build.  v1,v2: [E][3] ([E][2] in 2D) segment end points.
 * free_out[e] = 1 if the transition is collision-free. */
int hrmgpu_test_edges(hrmgpu_ctx* ctx, int k, int E, const double* v1, const double* v2, unsigned char* free_out);
/* The same test for the candidate edges of SEVERAL layers in one launch: edge e is tested against the C-obstacles of layer
 * layer[e] of the last build.  The candidate edges of connectOneSlice (HighwayRoadMap-inl.h:218-261, HRM3D.cpp:116-156)
 * depend on the free segments only, not on the graph, so a planner collects them for every layer, asks once, and replays
 * the reference's loops on the answers. */
int hrmgpu_test_edges_multi(hrmgpu_ctx* ctx, int E, const int* layer, const double* v1, const double* v2, unsigned char* free_out);

/* ---- bridge layers (K5) ----------------------------------------------------------------------- */
/* Replaces HRM3D::bridgeSlice (HRM3D.cpp:301-317): for P layer pairs, the Minkowski sums of every
 * OBSTACLE with each body's tightly-fitted ellipsoid (centred at the origin).
 * tfe_semi [P][B][3], tfe_R [P][B][9].  Meshes stay resident until the next call. */
int hrmgpu_build_bridge(hrmgpu_ctx* ctx, int P, int B, const double* tfe_semi, const double* tfe_R, int precision);
/* computeTFE + bridgeSlice in one call (HRM3D.cpp:521-539, 301-317; TightFitEllipsoid.cpp:43-114): the tightly-fitted ellipsoid of
 * body b between the orientations quat_a[p][b] and quat_b[p][b] ((w, x, y, z), [P][B][4]; for a link: the layer's rotation times
 * the link's) with num_step interpolation steps is evaluated on the device for all (pair, body) in one launch and feeds the
 * bridge-layer build directly.  body_semi [B][3].  tfe_out (optional, [P][B][7] = semi-axes + quaternion) returns the ellipsoids. */
int hrmgpu_build_bridge_tfe(hrmgpu_ctx* ctx, int P, int B, const double* body_semi, const double* quat_a, const double* quat_b, int num_step,
                            int precision, double* tfe_out);
/* 2D twin (HRM2D.cpp:199-213): tfe_semi [P][B][2], tfe_angle [P][B]. */
int hrmgpu_build_bridge2d(hrmgpu_ctx* ctx, int P, int B, const double* tfe_semi, const double* tfe_angle, int precision);
/* Replaces isPtInCFree for all bodies of Q interpolated robot poses of pair p (HRM3D.cpp:367-425,
 * HRM2D.cpp:235-282): centres [Q][B][3] ([Q][B][2] in 2D) are the body centres at each pose;
 * free_out[q] = 1 iff every body centre is outside every bridge C-obstacle of its body. */
int hrmgpu_test_bridge(hrmgpu_ctx* ctx, int p, int Q, const double* centres, unsigned char* free_out);
/* The same test for poses of SEVERAL layer pairs in one launch (connectMultiSlice, HRM3D.cpp:158-224): pose q is tested
 * against the bridge layer of pair pair[q]. */
int hrmgpu_test_bridge_multi(hrmgpu_ctx* ctx, int Q, const int* pair, const double* centres, unsigned char* free_out);

/* Replaces isMultiSliceTransitionFree (HRM3D.cpp:367-392, HRM2D.cpp:235-265) for C candidate vertex pairs of any layer
 * pairs in ONE launch, with the interpolation done on the device.  The robot moves from v0[c] to v1[c] ([C][3], 2D [C][2]:
 * the positions of the two vertices) through T poses; pose s has the base centre  t = w0[s]*v0 + w1[s]*v1  (two rounded
 * products and one rounded sum, as interpolateSE3 / the 2D loop evaluate it with w1 = s*dt, w0 = 1 - s*dt) and body b > 0 sits
 * at link_off[pair[c]][s][b] + t, where link_off [P][T][B][dim] holds the link translations rotated by the pair's s-th
 * interpolated orientation (MultiBodyTree3D.cpp:34-53; every vertex of a layer carries the layer's orientation, so these are
 * the same for all candidates of a pair; row b = 0 is ignored).  free_out[c] = 1 iff every pose is free for every body. */
int hrmgpu_test_transitions(hrmgpu_ctx* ctx, int C, const int* pair, const double* v0, const double* v1, int T, const double* w0,
                            const double* w1, const double* link_off, unsigned char* free_out);
/* The same for an articulated robot (ProbHRM3D::connectMultiSlice, ProbHRM3D.cpp:105-162; poses set by
 * MultiBodyTree3D::robotTF(urdf, gBase, joints), MultiBodyTree3D.cpp:55-89): link i sits at gBase * FK(body i) * tf_i, whose position is
 * evaluated as  link_off + (chain_off + t)  with chain_off[pair][s][b] = R_s * FK_s(b).t (the chain's translation in the base frame of
 * pose s, added to the base position first) and link_off[pair][s][b] = (R_s * FK_s(b).R) * tf_b.t.  All vertices of a slice share the
 * slice's base orientation and joint angles, so both tables are per (pair, pose, body) like link_off above.  3D only. */
int hrmgpu_test_transitions_articulated(hrmgpu_ctx* ctx, int C, const int* pair, const double* v0, const double* v1, int T, const double* w0,
                                        const double* w1, const double* link_off, const double* chain_off, unsigned char* free_out);

/* connectOneSlice for ALL layers of the last build as ONE call (HighwayRoadMap-inl.h:218-326, HRM3D.cpp:93-156).  The vertices of a
 * layer are its free segments (mid points) in CSR order; the reference tries to connect, plane by plane, every vertex of line
 * (ix, iy) with every vertex of line (ix, iy + 1), then every vertex of line (ix, iy) with every vertex of line (ix + 1, iy), in
 * that order -- the pair order of this call.  Each pair is tried directly and through the two detour vertices of bridgeVertex
 * (:295-326): code_out[pair] = 0 the direct segment is free, 1 both segments through (x1, y1, z2) are, 2 both through
 * (x2, y2, z1) are, 3 none.  The 5 candidate segments of every pair are generated on the device from the free segments and
 * tested against the layer's C-obstacle meshes (same test as hrmgpu_test_edges); only the codes come back.
 * layer_first [K + 1] (optional) = index of every layer's first pair.  Returns the number of pairs, or a negative error code. */
long long hrmgpu_connect_slices3d(hrmgpu_ctx* ctx, unsigned char* code_out, long long cap, long long* layer_first);

/* connectMultiSlice's vertex matching as ONE call (HRM3D.cpp:158-224).  vertex_xyz [V][3] are the positions of the roadmap's
 * vertices (a layer's vertices are its free-segment mid points plus the bridge vertices connectOneSlice inserted, so the host
 * names them); pair p joins the vertices range4[p] = {near begin, near end, cur begin, cur end} (id ranges of the closest earlier
 * layer, :161-176, and of the current one) and is tested against bridge layer p of the last hrmgpu_build_bridge (P must equal
 * their number).  For every vertex m0 of the near range, in order, the vertices m1 of the cur range are scanned from the
 * reference's moving lower bound: the first one with |x0 - x1| <= thx, |y0 - y1| <= thy (:195-200) whose interpolated
 * transition is free (same T, w0, w1, link_off as hrmgpu_test_transitions) is its match, and the new lower bound.
 * match_out[first(p) + (m0 - near begin)] = m1 (vertex id) or -1, pairs concatenated.  The candidate enumeration, the transition
 * tests and the replay all run on the device; only the matches come back.  chain_off (NULL for rigid robots) selects the
 * articulated variant of hrmgpu_test_transitions_articulated.  Returns the number of entries written, or a negative error
 * code; n_candidates_out (optional) = transition tests made. */
long long hrmgpu_connect_pairs3d(hrmgpu_ctx* ctx, long long V, const double* vertex_xyz, int P, const int* range4, double thx, double thy, int T,
                                 const double* w0, const double* w1, const double* link_off, const double* chain_off, int* match_out,
                                 long long match_cap, long long* n_candidates_out);

/* ---- multi-GPU exchange ------------------------------------------------------------------------ */
/* C-layers shard by orientation: every rank (one process, one context, one GPU) builds its own block of layers with no
 * collective on the way.  The one exchange gives every rank the free segments (= roadmap vertices, HRM3D.cpp:93-113) of ALL
 * layers, which bridge-layer pairing and connectMultiSlice (HRM3D.cpp:158-224) need.  Meshes are never exchanged.
 *
 * hrmgpu_allgather_layers packs the last build into a fixed-size block
 *     int64 lines (= K*L of the sender), int64 total segments, int64 off[line_cap+1], double xL[seg_cap], xU[seg_cap], xM[seg_cap]
 * and issues ONE ncclAllGather of those blocks, all on the side stream (or on `cuda_stream`, a cudaStream_t, if non-NULL)
 * behind the free-segment passes: no size round trip, no host synchronisation, overlapped with the next build.  line_cap and
 * seg_cap must be the same on all ranks and line_cap >= K*L; a sender with more than seg_cap segments still sends its true
 * total, so every receiver sees that the capacity must be raised (hrmgpu_get_gathered returns HRMGPU_E_CAP).
 * nccl_comm: an ncclComm_t of the host framework, or NULL for the communicator hrmgpu_comm_init created.  NCCL is bound at run
 * time: the libnccl.so.2 already loaded in the process (e.g. PyTorch's) is used, else $HRMGPU_NCCL_LIB, else the system one. */
int hrmgpu_comm_unique_id(void* id128);                       /* ncclGetUniqueId: 128 bytes, to be broadcast by the host framework */
int hrmgpu_comm_init(hrmgpu_ctx* ctx, int n_ranks, int rank, const void* id128);   /* ncclCommInitRank on the context's device */
int hrmgpu_comm_destroy(hrmgpu_ctx* ctx);
int hrmgpu_allgather_layers(hrmgpu_ctx* ctx, void* nccl_comm, void* cuda_stream, long long line_cap, long long seg_cap);
/* Last exchange: out4 = {world size, own rank, line_cap, seg_cap}; *d_blocks (if non-NULL) = device address of the world-size
 * received blocks (layout above), valid until the next exchange. */
int hrmgpu_gathered_info(hrmgpu_ctx* ctx, long long* out4, void** d_blocks);
/* Rank `rank`'s block of the last exchange on the host: *lines_out = its K*L, off [lines+1] (capacity line_cap+1), xL/xU/xM
 * [cap].  Two-call convention as hrmgpu_get_segments.  Returns that rank's segment count or a negative status. */
long long hrmgpu_get_gathered(hrmgpu_ctx* ctx, int rank, long long* lines_out, long long* off, double* xL, double* xU, double* xM,
                              long long cap);
/* Packs the last build's free segments into one contiguous DEVICE buffer owned by the caller's framework (e.g. a torch
 * tensor): header [K*L+1] int64 CSR offsets followed by xL, xU, xM (doubles), exact sizes.  With d_buf == NULL returns the
 * required size in bytes.  Waits on the host for the free-segment passes (the size is data dependent).  The copies run on the
 * context's MAIN stream: with the context's own stream the call returns when the buffer is complete; with a stream installed
 * by hrmgpu_set_stream they are ordered on that stream and the caller must consume the buffer on it (or synchronise). */
long long hrmgpu_pack_segments_dev(hrmgpu_ctx* ctx, void* d_buf, long long cap_bytes);

/* ---- instrumentation ---------------------------------------------------------------------------- */
/* Number of kernels this context has launched since creation / since the last reset. */
long long hrmgpu_kernel_launches(hrmgpu_ctx* ctx, int reset);
/* Device time of the dominant kernel (fused Minkowski + sweep) of the last build call, measured with
 * CUDA events on the launching stream.  Synchronises.  Returns < 0 on error. */
double hrmgpu_last_minksweep_ms(hrmgpu_ctx* ctx);
/* Device time of the free-segment passes of the last build (CUDA events on the side stream).  Synchronises. */
double hrmgpu_last_segments_ms(hrmgpu_ctx* ctx);
/* How often free-segment passes had to be repeated because an internal capacity (grown from earlier builds) was too small. */
long long hrmgpu_segment_redos(hrmgpu_ctx* ctx, int reset);

#ifdef __cplusplus
}
#endif
#endif /* HRMGPU_H */
