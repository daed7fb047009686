// Flat C entry points over the C++ host layer (hrm_host.hpp) so that ctypes-based parity tests — and any non-C++
// caller — can run the GPU-backed planners.  Argument layout mirrors the reference's scene files after
// parsePlanningConfig: superquadric rows [a,b,c,e1,e2,px,py,pz,qw,qx,qy,qz], superellipse rows [a,b,eps,px,py,angle],
// robot rows = base first, links relative to the base (resources/readme.md).
#include "hrm_articulated.hpp"
#include "hrm_io.hpp"

#include <cstring>

using namespace hrm;

static SuperQuadrics sqFromRow(const double* r, int n) {
    return SuperQuadrics({r[0], r[1], r[2]}, {r[3], r[4]}, {r[5], r[6], r[7]}, Quaternion{r[8], r[9], r[10], r[11]}, static_cast<Index>(n));
}
static SuperEllipse seFromRow(const double* r, int num) {
    return SuperEllipse({r[0], r[1]}, r[2], {r[3], r[4]}, r[5], static_cast<Index>(num));
}
static thread_local std::string g_err;
static int g_max_refine = 0;  // refinement rounds hrmhost_plan3d / plan2d may run (HighwayRoadMap-inl.h:78-81); info[4] = rounds done

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
void hrmhost_set_max_refine(int rounds) { g_max_refine = rounds < 0 ? 0 : rounds; }
static int g_prob_refine_period = 60;  // ProbHRM3D.cpp:52-56
void hrmhost_set_prob_refine_period(int n) { g_prob_refine_period = n < 0 ? 0 : n; }

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
        hrm.setMaxRefine(static_cast<size_t>(g_max_refine));
        hrm.plan(g_max_refine > 0 ? 3600.0 : 0.0);  // the parity tests bound refinement by rounds; the time limit only has to be generous
        exportResult(hrm, info, vertex, cap_v, 7, edge, cap_e, path, cap_p);
        info[4] = static_cast<long>(hrm.numRefineDone());
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// The reference's per-layer call sequence (HRM3D.cpp:24-54,63-91) on a FreeSpaceLayer3D view of every built layer: per x-plane
// computeIntersectionInterval({tx, ty}) -> computeFreeSegment(ty) -> getFreeSegment().  Outputs, per layer, what the oracle's
// layer3d returns: lo / up [K][L][S], CSR off [K][L+1] and xL / xU / xM (capacity cap per layer, counts in cnt[K]).
int hrmhost_layer_views3d(int n_arena, int n_obs, const double* sq, int n, int B, const double* robot, int K, const double* quats,
                          const double* lim, int Lx, int Ly, int precision, int device, double* lo, double* up, int* off, double* xL, double* xU,
                          double* xM, long cap, long* cnt, double* bound0) {
    try {
        std::vector<SuperQuadrics> arena, obs;
        for (int i = 0; i < n_arena; ++i) arena.push_back(sqFromRow(sq + 12 * i, n));
        for (int i = 0; i < n_obs; ++i) obs.push_back(sqFromRow(sq + 12 * (n_arena + i), n));
        MultiBodyTree3D rb(sqFromRow(robot, n));
        for (int b = 1; b < B; ++b) rb.addBody(sqFromRow(robot + 12 * b, n));
        std::vector<MultiBodyTree3D> poses;
        for (int k = 0; k < K; ++k) {
            MultiBodyTree3D r = rb;
            SE3Transform g;
            g.R = toRotationMatrix(Quaternion{quats[4 * k], quats[4 * k + 1], quats[4 * k + 2], quats[4 * k + 3]});
            r.robotTF(g);
            poses.push_back(r);
        }
        FreeSpaceGPU3D fs(arena, obs, device);
        fs.setup(std::vector<Coordinate>(lim, lim + 6), static_cast<Index>(Lx), static_cast<Index>(Ly));
        fs.buildLayers(poses, precision);
        const int S = (n_arena + n_obs) * B, Sa = n_arena * B, L = Lx * Ly;
        for (int k = 0; k < K; ++k) {
            FreeSpaceLayer3D layer(fs, k);
            layer.computeCSpaceBoundary();
            const BoundaryInfo& b = layer.getCSpaceBoundary();
            layer.computeCSpaceBoundaryMesh(b);
            if (k == 0 && bound0) {
                size_t at = 0;
                for (const auto* grp : {&b.arena, &b.obstacle})
                    for (const auto& p : *grp) std::copy(p.begin(), p.end(), bound0 + at), at += p.size();
            }
            const std::vector<Coordinate> ty = layer.ty();
            long tot = 0;
            std::vector<Coordinate> txSoFar;
            for (int ix = 0; ix < Lx; ++ix) {
                txSoFar.push_back(layer.tx()[ix]);  // the reference grows tx plane by plane and sweeps at tx.back()
                layer.computeIntersectionInterval({txSoFar, ty});
                layer.computeFreeSegment(ty);
                const IntersectionInterval& iv = layer.getIntersectionInterval();
                const FreeSegment2D& seg = layer.getFreeSegment();
                for (int iy = 0; iy < Ly; ++iy) {
                    const size_t l = static_cast<size_t>(ix) * Ly + iy, base = (static_cast<size_t>(k) * L + l) * S;
                    for (int s = 0; s < Sa; ++s) lo[base + s] = iv.arenaLow[iy][s], up[base + s] = iv.arenaUpp[iy][s];
                    for (int s = Sa; s < S; ++s) lo[base + s] = iv.obstacleLow[iy][s - Sa], up[base + s] = iv.obstacleUpp[iy][s - Sa];
                    off[static_cast<size_t>(k) * (L + 1) + l] = static_cast<int>(tot);
                    for (size_t j = 0; j < seg.xM[iy].size(); ++j, ++tot)
                        if (tot < cap) {
                            const size_t o = static_cast<size_t>(k) * cap + tot;
                            xL[o] = seg.xL[iy][j], xU[o] = seg.xU[iy][j], xM[o] = seg.xM[iy][j];
                        }
                }
            }
            off[static_cast<size_t>(k) * (L + 1) + L] = static_cast<int>(tot);
            cnt[k] = tot;
        }
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
        hrm.setMaxRefine(static_cast<size_t>(g_max_refine));
        hrm.plan(g_max_refine > 0 ? 3600.0 : 0.0);  // the parity tests bound refinement by rounds; the time limit only has to be generous
        exportResult(hrm, info, vertex, cap_v, 3, edge, cap_e, path, cap_p);
        info[4] = static_cast<long>(hrm.numRefineDone());
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}


// ---- scene / config I/O (hrm_io.hpp) ---------------------------------------------------------------------------------
// parse2DCsvFile: rows flattened with a fixed stride; row_len[i] = number of cells of record i.  Returns the record
// count (which may exceed cap_rows: call again with a larger buffer) or -1.
long hrmhost_parse_csv(const char* filename, double* out, int stride, int* row_len, long cap_rows) {
    try {
        const auto data = parse2DCsvFile(filename);
        for (size_t i = 0; i < data.size() && static_cast<long>(i) < cap_rows; ++i) {
            row_len[i] = static_cast<int>(data[i].size());
            for (size_t j = 0; j < data[i].size() && static_cast<int>(j) < stride; ++j) out[static_cast<size_t>(stride) * i + j] = data[i][j];
        }
        return static_cast<long>(data.size());
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// parsePlanningConfig(objectType, envType, robotType, dim) with explicit resources / config directories
int hrmhost_parse_planning_config(const char* objectType, const char* envType, const char* robotType, const char* dim, const char* resources,
                                  const char* config) {
    try {
        Paths p = Paths::fromEnv();
        p.resources = resources;
        p.config = config;
        parsePlanningConfig(objectType, envType, robotType, dim, p);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// PlannerSetting::loadEnvironment + loadRobotMultiBody* + defineParameters on a config directory (prefix ends in '/').
// counts[8] = n_arena, n_obs, n_robot_bodies, n_end_points, n_quats, numLineX, numLineY, numSlice.  Object rows have
// stride 12 (3D: a,b,c,e1,e2,px,py,pz,qw,qx,qy,qz) or 6 (2D), end points stride 16, quats w,x,y,z, lim[6|4].
// Robot link rows are the links as stored in the tree (relative to the base, like the file).
int hrmhost_load_scene(const char* dim, const char* config_prefix, const char* quat_file, int num_param, int numLineX, int numLineY, long* counts,
                       double* arena, double* obstacle, double* robot, double* ends, double* quats, double* lim, long cap_rows) {
    try {
        PlannerParameter param;
        param.numLineX = static_cast<Index>(numLineX), param.numLineY = static_cast<Index>(numLineY);
        std::vector<std::vector<double>> endPts;
        auto put = [&](double* dst, long i, const std::vector<double>& row, int stride) {
            if (i < cap_rows)
                for (size_t j = 0; j < row.size() && static_cast<int>(j) < stride; ++j) dst[stride * i + static_cast<long>(j)] = row[j];
        };
        if (std::string(dim) == "3D") {
            PlannerSetting3D env(num_param);
            env.loadEnvironment(config_prefix);
            const MultiBodyTree3D rb = loadRobotMultiBody3D(config_prefix, quat_file, num_param);
            defineParameters(rb, env, param);
            auto row = [](const SuperQuadrics& s) {
                std::vector<double> r = s.getSemiAxis();
                r.insert(r.end(), s.getEpsilon().begin(), s.getEpsilon().end());
                r.insert(r.end(), s.getPosition().begin(), s.getPosition().end());
                r.insert(r.end(), s.getQuaternion().begin(), s.getQuaternion().end());
                return r;
            };
            for (size_t i = 0; i < env.getArena().size(); ++i) put(arena, static_cast<long>(i), row(env.getArena()[i]), 12);
            for (size_t i = 0; i < env.getObstacle().size(); ++i) put(obstacle, static_cast<long>(i), row(env.getObstacle()[i]), 12);
            put(robot, 0, row(rb.getBase()), 12);
            for (size_t i = 0; i < rb.getLinks().size(); ++i) put(robot, static_cast<long>(i) + 1, row(rb.getLinks()[i]), 12);
            const auto& qs = rb.getBase().getQuatSamples();
            for (size_t i = 0; i < qs.size(); ++i) put(quats, static_cast<long>(i), {qs[i][0], qs[i][1], qs[i][2], qs[i][3]}, 4);
            counts[0] = static_cast<long>(env.getArena().size()), counts[1] = static_cast<long>(env.getObstacle().size());
            counts[2] = static_cast<long>(rb.getNumLinks()) + 1, counts[4] = static_cast<long>(qs.size());
            endPts = env.getEndPoints();
        } else {
            PlannerSetting2D env(num_param);
            env.loadEnvironment(config_prefix);
            const MultiBodyTree2D rb = loadRobotMultiBody2D(config_prefix, num_param);
            defineParameters(rb, env, param);
            auto row = [](const SuperEllipse& s) {
                return std::vector<double>{s.getSemiAxis()[0], s.getSemiAxis()[1], s.getEpsilon(), s.getPosition()[0], s.getPosition()[1], s.getAngle()};
            };
            for (size_t i = 0; i < env.getArena().size(); ++i) put(arena, static_cast<long>(i), row(env.getArena()[i]), 6);
            for (size_t i = 0; i < env.getObstacle().size(); ++i) put(obstacle, static_cast<long>(i), row(env.getObstacle()[i]), 6);
            put(robot, 0, row(rb.getBase()), 6);
            for (size_t i = 0; i < rb.getLinks().size(); ++i) put(robot, static_cast<long>(i) + 1, row(rb.getLinks()[i]), 6);
            counts[0] = static_cast<long>(env.getArena().size()), counts[1] = static_cast<long>(env.getObstacle().size());
            counts[2] = static_cast<long>(rb.getNumLinks()) + 1, counts[4] = 0;
            endPts = env.getEndPoints();
        }
        for (size_t i = 0; i < endPts.size(); ++i) put(ends, static_cast<long>(i), endPts[i], 16);
        counts[3] = static_cast<long>(endPts.size());
        counts[5] = static_cast<long>(param.numLineX), counts[6] = static_cast<long>(param.numLineY), counts[7] = static_cast<long>(param.numSlice);
        for (size_t i = 0; i < param.boundaryLimits.size(); ++i) lim[i] = param.boundaryLimits[i];
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// storeGraphInfo + storePathInfo on a caller-supplied roadmap (the writers' formats can be checked without a GPU)
int hrmhost_store_result(const char* details_dir, const char* suffix, int vdim, long n_vertex, const double* vertex, long n_edge, const long* edge,
                         long n_path, const long* path) {
    try {
        Paths p = Paths::fromEnv();
        p.details = details_dir;
        Graph g;
        for (long i = 0; i < n_vertex; ++i) g.vertex.emplace_back(vertex + vdim * i, vertex + vdim * (i + 1));
        for (long i = 0; i < n_edge; ++i) g.edge.emplace_back(static_cast<Index>(edge[2 * i]), static_cast<Index>(edge[2 * i + 1]));
        SolutionPathInfo sp;
        for (long i = 0; i < n_path; ++i) {
            sp.PathId.push_back(static_cast<Index>(path[i]));
            sp.solvedPath.push_back(g.vertex.at(static_cast<size_t>(path[i])));
        }
        storeGraphInfo(g, suffix, p);
        storePathInfo(sp, suffix, p);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// ---- articulated robots (hrm_articulated.hpp) ---------------------------------------------------------------------------
// hrm::ParseURDF::getTransform(jointConfig, "body<k>") (body 0 = root link): out16 row-major 4x4; returns the joint count.
int hrmhost_fk(const char* urdf, const double* joints, int nj, int body, double* out16) {
    try {
        ParseURDF kdl(urdf);
        const SE3Transform f = kdl.getTransform(std::vector<double>(joints, joints + nj), body == 0 ? kdl.getRootName() : "body" + std::to_string(body));
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) out16[4 * i + j] = f.R[i][j];
            out16[4 * i + 3] = f.t[i];
        }
        out16[12] = out16[13] = out16[14] = 0.0, out16[15] = 1.0;
        return static_cast<int>(kdl.getNrOfJoints());
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// hrm::planners::ProbHRM3D(robot, urdf, arena, obs, req).plan().  samples [n_samples][4 + nj] replace the random draws
// (n_samples < 0: seeded RNG, `seed`); start/goal 7 + nj doubles; vertex stride 7 + nj; info[4] = C-slices built.
int hrmhost_plan_prob3d(int n_arena, int n_obs, const double* sq, int n, int B, const double* robot, const char* urdf, int nj, int n_samples,
                        const double* samples, unsigned long long seed, int batch, const double* lim, int Lx, int Ly, int numPoint,
                        const double* start, const double* goal, int per_orientation, int precision, int device, double time_limit, long* info,
                        double* vertex, long cap_v, long* edge, long cap_e, long* path, long cap_p) {
    try {
        std::vector<SuperQuadrics> arena, obs;
        for (int i = 0; i < n_arena; ++i) arena.push_back(sqFromRow(sq + 12 * i, n));
        for (int i = 0; i < n_obs; ++i) obs.push_back(sqFromRow(sq + 12 * (n_arena + i), n));
        MultiBodyTree3D rb(sqFromRow(robot, n));
        for (int b = 1; b < B; ++b) rb.addBody(sqFromRow(robot + 12 * b, n));
        PlanningRequest req;
        req.isRobotRigid = false;
        req.parameters.boundaryLimits.assign(lim, lim + 6);
        req.parameters.numLineX = static_cast<Index>(Lx);
        req.parameters.numLineY = static_cast<Index>(Ly);
        req.parameters.numPoint = static_cast<Index>(numPoint);
        req.start.assign(start, start + 7 + nj);
        req.goal.assign(goal, goal + 7 + nj);
        planners::ProbHRM3D hrm(rb, urdf, arena, obs, req, device, precision, per_orientation != 0);
        if (n_samples >= 0) {
            std::vector<std::vector<double>> draws;
            for (int i = 0; i < n_samples; ++i) draws.emplace_back(samples + static_cast<size_t>(i) * (4 + nj), samples + static_cast<size_t>(i + 1) * (4 + nj));
            hrm.setSampleList(draws);
        } else {
            hrm.setSeed(seed);
        }
        hrm.setBatch(static_cast<size_t>(batch));
        hrm.setRefinePeriod(static_cast<size_t>(g_prob_refine_period));
        hrm.plan(time_limit);
        exportResult(hrm, info, vertex, cap_v, 7 + nj, edge, cap_e, path, cap_p);
        info[4] = static_cast<long>(hrm.numSlices());
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

}  // extern "C"
