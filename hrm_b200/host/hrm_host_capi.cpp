// Flat C entry points over the C++ host layer (hrm_host.hpp) so that ctypes-based parity tests — and any non-C++
// caller — can run the GPU-backed planners.  Argument layout mirrors the reference's scene files after
// parsePlanningConfig: superquadric rows [a,b,c,e1,e2,px,py,pz,qw,qx,qy,qz], superellipse rows [a,b,eps,px,py,angle],
// robot rows = base first, links relative to the base (resources/readme.md).
#include "hrm_host.hpp"

#include <cstring>

using namespace hrm;

static SuperQuadrics sqFromRow(const double* r, int n) {
    return SuperQuadrics({r[0], r[1], r[2]}, {r[3], r[4]}, {r[5], r[6], r[7]}, Quaternion{r[8], r[9], r[10], r[11]}, static_cast<Index>(n));
}
static SuperEllipse seFromRow(const double* r, int num) {
    return SuperEllipse({r[0], r[1]}, r[2], {r[3], r[4]}, r[5], static_cast<Index>(num));
}
static thread_local std::string g_err;

template <class Planner>
static void exportResult(Planner& hrm, long* info, double* vertex, long cap_v, int vdim, long* edge, long cap_e, long* path, long cap_p) {
    const PlanningResult& res = hrm.getPlanningResult();
    const Graph& g = res.graphStructure;
    info[0] = res.solved ? 1 : 0;
    info[1] = static_cast<long>(g.vertex.size());
    info[2] = static_cast<long>(g.edge.size());
    info[3] = static_cast<long>(res.solutionPath.PathId.size());
    info[4] = 0;
    info[5] = hrm.numEdgeTests();
    info[6] = static_cast<long>(hrm.freeSpace().kernelLaunches());
    info[7] = static_cast<long>(res.planningTime.totalTime * 1e6);
    for (size_t i = 0; i < g.vertex.size() && static_cast<long>(i) < cap_v; ++i)
        for (int d = 0; d < vdim; ++d) vertex[vdim * i + d] = g.vertex[i][d];
    for (size_t i = 0; i < g.edge.size() && static_cast<long>(i) < cap_e; ++i)
        edge[2 * i] = static_cast<long>(g.edge[i].first), edge[2 * i + 1] = static_cast<long>(g.edge[i].second);
    for (size_t i = 0; i < res.solutionPath.PathId.size() && static_cast<long>(i) < cap_p; ++i) path[i] = static_cast<long>(res.solutionPath.PathId[i]);
}

extern "C" {

const char* hrmhost_last_error(void) { return g_err.c_str(); }

// hrm::planners::HRM3D(robot, arena, obs, req).plan().  Returns 0, or -1 with hrmhost_last_error() set.
// info[8]: solved, n_vertex, n_edge, n_path, -, n_edge_tests, gpu_kernel_launches, total_time_us.
int hrmhost_plan3d(int n_arena, int n_obs, const double* sq, int n, int B, const double* robot, int numSlice, const double* quats,
                   const double* lim, int Lx, int Ly, int numPoint, const double* start, const double* goal, int per_orientation, int precision,
                   int device, long* info, double* vertex, long cap_v, long* edge, long cap_e, long* path, long cap_p) {
    try {
        std::vector<SuperQuadrics> arena, obs;
        for (int i = 0; i < n_arena; ++i) arena.push_back(sqFromRow(sq + 12 * i, n));
        for (int i = 0; i < n_obs; ++i) obs.push_back(sqFromRow(sq + 12 * (n_arena + i), n));
        SuperQuadrics base = sqFromRow(robot, n);
        std::vector<Quaternion> qs;
        for (int i = 0; i < numSlice; ++i) qs.push_back({quats[4 * i], quats[4 * i + 1], quats[4 * i + 2], quats[4 * i + 3]});
        base.setQuatSamples(qs);  // loadPreDefinedQuaternions (ParsePlanningSettings.cpp:85-106)
        MultiBodyTree3D rb(base);
        for (int b = 1; b < B; ++b) rb.addBody(sqFromRow(robot + 12 * b, n));
        PlanningRequest req;
        req.parameters.boundaryLimits.assign(lim, lim + 6);
        req.parameters.numSlice = static_cast<Index>(numSlice);
        req.parameters.numLineX = static_cast<Index>(Lx);
        req.parameters.numLineY = static_cast<Index>(Ly);
        req.parameters.numPoint = static_cast<Index>(numPoint);
        req.start.assign(start, start + 7);
        req.goal.assign(goal, goal + 7);
        planners::HRM3D hrm(rb, arena, obs, req, device, precision, per_orientation != 0);
        hrm.plan();
        exportResult(hrm, info, vertex, cap_v, 7, edge, cap_e, path, cap_p);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// hrm::planners::HRM2D(robot, arena, obs, req).plan()
int hrmhost_plan2d(int n_arena, int n_obs, const double* se, int num, int B, const double* robot, int numSlice, const double* lim, int Ly,
                   int numPoint, const double* start, const double* goal, int per_orientation, int precision, int device, long* info,
                   double* vertex, long cap_v, long* edge, long cap_e, long* path, long cap_p) {
    try {
        std::vector<SuperEllipse> arena, obs;
        for (int i = 0; i < n_arena; ++i) arena.push_back(seFromRow(se + 6 * i, num));
        for (int i = 0; i < n_obs; ++i) obs.push_back(seFromRow(se + 6 * (n_arena + i), num));
        MultiBodyTree2D rb(seFromRow(robot, num));
        for (int b = 1; b < B; ++b) rb.addBody(seFromRow(robot + 6 * b, num));
        PlanningRequest req;
        req.parameters.boundaryLimits.assign(lim, lim + 4);
        req.parameters.numSlice = static_cast<Index>(numSlice);
        req.parameters.numLineY = static_cast<Index>(Ly);
        req.parameters.numPoint = static_cast<Index>(numPoint);
        req.start.assign(start, start + 3);
        req.goal.assign(goal, goal + 3);
        planners::HRM2D hrm(rb, arena, obs, req, device, precision, per_orientation != 0);
        hrm.plan();
        exportResult(hrm, info, vertex, cap_v, 3, edge, cap_e, path, cap_p);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

}  // extern "C"
