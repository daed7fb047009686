// ORACLE — TEST INFRASTRUCTURE ONLY. Nothing under hrm_b200/ may include, link or call this.
//
// CPU restatement of the planner layer of ChirikjianLab/hrm that sits on the hot path: HighwayRoadMap
// (template base), HRM3D and HRM2D — roadmap construction, intra-layer connection, bridge layers, graph
// search.  Boost.Graph is replaced by a plain A* over an adjacency list (same admissible heuristic, the
// result consumed downstream is only path membership / success).  Divergences from the reference are the
// documented ones of SURVEY.md App. C: C6 (interval tables resized on refinement), C7 (single-hit = no
// interval / not blocking), C12 (no time-seeded randomness), C18 (minQuat initialised to q_[0]).
#pragma once
#include "hrmo_freespace.hpp"

#include <chrono>
#include <cmath>
#include <functional>
#include <queue>

namespace hrmo {

// src/util/DistanceMetric.cpp:5-13
inline double vectorEuclidean(const std::vector<double>& v1, const std::vector<double>& v2) {
    double acc = 0.0;
    for (size_t i = 0; i < v1.size(); ++i) {
        const double d = v1[i] - v2[i];
        acc = acc + d * d;
    }
    return std::sqrt(acc);
}

// ---- edge / bridge predicates as free functions (shared by the planner classes and the C API) --------------------

// HRM3D.cpp:319-365 — segment v1->v2 against every C-obstacle mesh; single hit counts as not blocking (C7).
inline bool isSameSliceTransitionFree3D(const std::vector<MeshMatrix>& obstacle, const double v1[3], const double v2[3],
                                        long* singleHit = nullptr) {
    const Vec3 t1 = {v1[0], v1[1], v1[2]}, t2 = {v2[0], v2[1], v2[2]};
    Vec3 v12 = t2 - t1;
    normalize(v12);
    const Line3D line = {{t1.x, t1.y, t1.z, v12.x, v12.y, v12.z}};
    for (const auto& obs : obstacle) {
        const auto h = intersectLineMesh3D(line, obs);
        if (h.size() == 1 && singleHit) ++*singleHit;
        if (h.size() >= 2) {
            const double s0 = dot(h[0] - t1, h[0] - t2);
            const double s1 = dot(h[1] - t1, h[1] - t2);
            if ((s0 < 0) || (s1 < 0)) return false;
        }
    }
    return true;
}

// HRM2D.cpp:215-232
inline bool isSameSliceTransitionFree2D(const std::vector<Points>& obstacle, const double v1[2], const double v2[2]) {
    for (const auto& obs : obstacle)
        if (isIntersectSegPolygon2D(v1, v2, obs)) return false;
    return true;
}

// HRM3D.cpp:394-425 — body centre v against the bridge C-obstacles of that body.
inline bool isPtInCFree3D(const std::vector<MeshMatrix>& bridge, const double v[3], long* singleHit = nullptr) {
    const Line3D lineZ = {{v[0], v[1], v[2], 0, 0, 1}};
    for (const auto& bound : bridge) {
        const auto h = intersectVerticalLineMesh3D(lineZ, bound);
        if (h.size() == 1 && singleHit) ++*singleHit;
        if (h.size() >= 2) {
            if (v[2] > std::fmin(h[0].z, h[1].z) && v[2] < std::fmax(h[0].z, h[1].z)) return false;
        }
    }
    return true;
}

// HRM2D.cpp:267-282
inline bool isPtInCFree2D(const std::vector<Points>& bridge, const double v[2], long* singleHit = nullptr) {
    for (const auto& bound : bridge) {
        const auto h = intersectHorizontalLinePolygon2D(v[1], bound);
        if (h.size() == 1 && singleHit) ++*singleHit;
        if (h.size() >= 2) {
            if (v[0] > std::fmin(h[0], h[1]) && v[0] < std::fmax(h[0], h[1])) return false;
        }
    }
    return true;
}

// ---- planner data (include/hrm/planners/PlanningRequest.h, PlanningResult.h, HighwayRoadMap.h:30-42) ---------------
struct PlannerParameter {
    std::vector<double> boundaryLimits;
    size_t numSlice = 0, numLineX = 0, numLineY = 0, numPoint = 5, numSearchNeighbor = 10;
    double searchRadius = HALF_PI;
};
struct PlanningRequest {
    bool isRobotRigid = true;
    PlannerParameter parameters;
    std::vector<double> start, goal;
};
struct Graph {
    std::vector<std::vector<double>> vertex;
    std::vector<std::pair<size_t, size_t>> edge;
    std::vector<double> weight;
};
struct PlanningResult {
    bool solved = false;
    double buildTime = 0, searchTime = 0, totalTime = 0;
    Graph graphStructure;
    std::vector<int> PathId;
    std::vector<std::vector<double>> solvedPath;
    double cost = 0;
};
struct VertexIdx {
    size_t startId = 0, slice = 0;
    std::vector<size_t> plane;
    std::vector<std::vector<size_t>> line;
};

// A* on the undirected roadmap with the Euclidean heuristic of HighwayRoadMap-inl.h:135-147.
// Returns true and fills the predecessor map when the goal is reached.
inline bool astar(const Graph& g, size_t s, size_t t, std::vector<size_t>& pred) {
    const size_t n = g.vertex.size();
    std::vector<std::vector<std::pair<size_t, double>>> adj(n);
    for (size_t i = 0; i < g.edge.size(); ++i) {
        adj[g.edge[i].first].push_back({g.edge[i].second, g.weight[i]});
        adj[g.edge[i].second].push_back({g.edge[i].first, g.weight[i]});
    }
    std::vector<double> d(n, INFINITY);
    pred.assign(n, n);
    using QE = std::pair<double, size_t>;
    std::priority_queue<QE, std::vector<QE>, std::greater<QE>> pq;
    d[s] = 0;
    pred[s] = s;
    pq.push({vectorEuclidean(g.vertex[s], g.vertex[t]), s});
    std::vector<char> closed(n, 0);
    while (!pq.empty()) {
        const size_t u = pq.top().second;
        pq.pop();
        if (closed[u]) continue;
        closed[u] = 1;
        if (u == t) return true;
        for (const auto& e : adj[u]) {
            const double nd = d[u] + e.second;
            if (nd < d[e.first]) {
                d[e.first] = nd;
                pred[e.first] = u;
                pq.push({nd + vectorEuclidean(g.vertex[e.first], g.vertex[t]), e.first});
            }
        }
    }
    return false;
}

// ---- HighwayRoadMap base (include/hrm/planners/HighwayRoadMap-inl.h) -------------------------------------------------
template <class RobotType, class ObjectType>
struct HighwayRoadMap {
    RobotType robot_;
    std::vector<ObjectType> arena_, obs_;
    std::vector<double> start_, goal_;
    PlannerParameter param_;
    bool isRobotRigid_;
    PlanningResult res_;
    VertexIdx numVertex_;
    std::vector<VertexIdx> vertexIdx_;
    long singleHitCount = 0;
    long numEdgeTests = 0;
    bool isRefine_ = false;
    std::vector<std::vector<VertexIdx>> vertexIdxAll_;  // vertex index records of the previous refinement rounds
    size_t numRefineDone = 0;

    HighwayRoadMap(const RobotType& robot, const std::vector<ObjectType>& arena, const std::vector<ObjectType>& obs,
                   const PlanningRequest& req)
        : robot_(robot), arena_(arena), obs_(obs), start_(req.start), goal_(req.goal), param_(req.parameters),
          isRobotRigid_(req.isRobotRigid) {}
    virtual ~HighwayRoadMap() = default;

    virtual void sampleOrientations() = 0;
    virtual void constructOneSlice(size_t sliceIdx) = 0;
    virtual void connectMultiSlice() = 0;
    virtual void connectExistSlice(size_t sliceId) = 0;
    virtual bool isSameSliceTransitionFree(const std::vector<double>& v1, const std::vector<double>& v2) = 0;
    virtual std::vector<size_t> getNearestNeighborsOnGraph(const std::vector<double>& vertex, size_t k, double radius) = 0;

    // :92-107
    void buildRoadmap() {
        sampleOrientations();
        for (size_t i = 0; i < param_.numSlice; ++i) {
            constructOneSlice(i);
            numVertex_.slice = res_.graphStructure.vertex.size();
            vertexIdx_.push_back(numVertex_);
        }
        connectMultiSlice();
    }

    // :110-175 (C11: cost indexes weights by vertex id in the reference; success and path membership are what we keep)
    void search() {
        if (res_.graphStructure.vertex.empty()) return;
        const auto idx_s = getNearestNeighborsOnGraph(start_, param_.numSearchNeighbor, param_.searchRadius);
        const auto idx_g = getNearestNeighborsOnGraph(goal_, param_.numSearchNeighbor, param_.searchRadius);
        for (size_t idxS : idx_s)
            for (size_t idxG : idx_g) {
                std::vector<size_t> p;
                if (astar(res_.graphStructure, idxS, idxG, p)) {
                    res_.PathId.clear();
                    res_.cost = 0;
                    size_t cur = idxG;
                    res_.PathId.push_back(static_cast<int>(cur));
                    while (cur != idxS) {
                        const size_t pr = p[cur];
                        res_.cost += vectorEuclidean(res_.graphStructure.vertex[cur], res_.graphStructure.vertex[pr]);
                        cur = pr;
                        res_.PathId.push_back(static_cast<int>(cur));
                    }
                    std::reverse(res_.PathId.begin(), res_.PathId.end());
                    res_.solved = true;
                    return;
                }
            }
    }

    // :178-215.  "Done right" (SURVEY.md finding 6 / §8 f.4): the interval tables are re-sized for the doubled line count
    // (the reference overruns them and throws) and every slice is re-swept on ITS OWN cached boundary (the reference's
    // FreeSpace object would still hold the mesh of the last slice built).  Returns as soon as a path exists.
    void refineExistRoadmap() {
        isRefine_ = true;
        vertexIdxAll_.push_back(vertexIdx_);
        vertexIdx_.clear();
        param_.numLineX *= 2;
        param_.numLineY *= 2;
        for (size_t i = 0; i < param_.numSlice; ++i) {
            constructOneSlice(i);
            numVertex_.slice = res_.graphStructure.vertex.size();
            vertexIdx_.push_back(numVertex_);
            connectExistSlice(i);
            search();
            if (res_.solved) return;
        }
        isRefine_ = false;
    }

    // :65-89; the reference refines `while (!solved && time < limit)`: here at most maxRefine rounds (0 = none)
    void plan(size_t maxRefine = 0) {
        auto t0 = std::chrono::high_resolution_clock::now();
        buildRoadmap();
        res_.buildTime = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
        t0 = std::chrono::high_resolution_clock::now();
        search();
        res_.searchTime = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
        while (!res_.solved && numRefineDone < maxRefine) {
            refineExistRoadmap();
            ++numRefineDone;
        }
        res_.totalTime = res_.buildTime + res_.searchTime;
        if (res_.solved) {
            const size_t poseSize = res_.graphStructure.vertex.at(0).size();
            start_.resize(poseSize);
            goal_.resize(poseSize);
            res_.solvedPath.push_back(start_);
            for (int id : res_.PathId) res_.solvedPath.push_back(res_.graphStructure.vertex.at(static_cast<size_t>(id)));
            res_.solvedPath.push_back(goal_);
        }
    }

    // :295-326
    void bridgeVertex(size_t idx1, size_t idx2) {
        const auto v1 = res_.graphStructure.vertex.at(idx1);
        const auto v2 = res_.graphStructure.vertex.at(idx2);
        auto vNew1 = v1;
        vNew1.at(2) = v2.at(2);
        auto vNew2 = v2;
        vNew2.at(2) = v1.at(2);
        auto vNew = v1;
        if (isSameSliceTransitionFree(v1, vNew1) && isSameSliceTransitionFree(vNew1, v2)) {
            vNew = vNew1;
        } else if (isSameSliceTransitionFree(v1, vNew2) && isSameSliceTransitionFree(vNew2, v2)) {
            vNew = vNew2;
        } else {
            return;
        }
        const size_t idxNew = res_.graphStructure.vertex.size();
        res_.graphStructure.vertex.push_back(vNew);
        res_.graphStructure.edge.push_back({idx1, idxNew});
        res_.graphStructure.weight.push_back(vectorEuclidean(v1, vNew));
        res_.graphStructure.edge.push_back({idxNew, idx2});
        res_.graphStructure.weight.push_back(vectorEuclidean(vNew, v2));
    }

    // :218-261
    void connectOneSlice2D(const FreeSegment2D& freeSeg) {
        for (size_t i = 0; i < freeSeg.ty.size(); ++i) {
            const size_t n1 = numVertex_.plane.at(i);
            for (size_t j1 = 0; j1 < freeSeg.xM[i].size(); ++j1) {
                if (j1 != freeSeg.xM[i].size() - 1) {
                    if (std::fabs(freeSeg.xU[i][j1] - freeSeg.xL[i][j1 + 1]) < 1e-6) {
                        res_.graphStructure.edge.push_back({n1 + j1, n1 + j1 + 1});
                        res_.graphStructure.weight.push_back(
                            vectorEuclidean(res_.graphStructure.vertex[n1 + j1], res_.graphStructure.vertex[n1 + j1 + 1]));
                    }
                }
                if (i != freeSeg.ty.size() - 1) {
                    const size_t n2 = numVertex_.plane.at(i + 1);
                    for (size_t j2 = 0; j2 < freeSeg.xM[i + 1].size(); ++j2) {
                        if (isSameSliceTransitionFree(res_.graphStructure.vertex[n1 + j1], res_.graphStructure.vertex[n2 + j2])) {
                            res_.graphStructure.edge.push_back({n1 + j1, n2 + j2});
                            res_.graphStructure.weight.push_back(
                                vectorEuclidean(res_.graphStructure.vertex[n1 + j1], res_.graphStructure.vertex[n2 + j2]));
                        } else {
                            bridgeVertex(n1 + j1, n2 + j2);
                        }
                    }
                }
            }
        }
    }
};

// ---- HRM3D (src/planners/HRM3D.cpp) ------------------------------------------------------------------------------------
struct HRM3D : HighwayRoadMap<MultiBodyTree3D, SuperQuadrics> {
    FreeSpace3D freeSpace_;
    std::vector<Quat> q_;
    std::vector<Quat> quatSamples_;  // robot_.getBase().getQuatSamples()
    std::vector<std::vector<double>> v_;  // articulated robot shapes, one per C-slice (HRM3D.h: v_)
    BoundaryInfo sliceBound_;
    BoundaryMesh sliceBoundMesh_;
    std::vector<BoundaryInfo> sliceBoundAll_;
    std::vector<BoundaryMesh> sliceBoundMeshAll_;
    FreeSegment3D freeSegOneSlice_;
    std::vector<Ellipsoid3> tfe_;
    std::vector<std::vector<MeshMatrix>> bridgeSliceBound_;
    bool perOrientationFreeSpace_;  // false = reference snapshot (FreeSpace robot copy never re-oriented, finding 1)

    HRM3D(const MultiBodyTree3D& robot, const std::vector<SuperQuadrics>& arena, const std::vector<SuperQuadrics>& obs,
          const PlanningRequest& req, const std::vector<Quat>& quatSamples, bool perOrientationFreeSpace = false)
        : HighwayRoadMap(robot, arena, obs, req), freeSpace_(robot_, arena_, obs_), quatSamples_(quatSamples),
          perOrientationFreeSpace_(perOrientationFreeSpace) {
        freeSpace_.setup(static_cast<unsigned>(param_.numLineY), param_.boundaryLimits[4], param_.boundaryLimits[5]);  // :17-19
    }

    // :510-518
    virtual void setTransform(const std::vector<double>& v) {
        Tf3 g;
        g.R = quat_to_rot(Quat{v[3], v[4], v[5], v[6]});
        g.t[0] = v[0], g.t[1] = v[1], g.t[2] = v[2];
        robot_.robotTF(g);
    }

    // :58-61,427-441 (pre-defined samples only: C12)
    void sampleOrientations() override {
        param_.numSlice = quatSamples_.size();
        q_ = quatSamples_;
    }

    // :24-54
    void constructOneSlice(size_t sliceIdx) override {
        if (isRobotRigid_)
            setTransform({0.0, 0.0, 0.0, q_.at(sliceIdx).w, q_.at(sliceIdx).x, q_.at(sliceIdx).y, q_.at(sliceIdx).z});
        else
            setTransform(v_.at(sliceIdx));  // articulated: base pose + joint angles (ProbHRM3D)
        if (!isRefine_) {
            if (perOrientationFreeSpace_) freeSpace_.setRobot(robot_);
            freeSpace_.computeCSpaceBoundary();
            sliceBound_ = freeSpace_.cSpaceBoundary_;
            sliceBoundAll_.push_back(sliceBound_);
            freeSpace_.computeCSpaceBoundaryMesh(sliceBound_);
            sliceBoundMesh_ = freeSpace_.cSpaceBoundaryMesh_;
            sliceBoundMeshAll_.push_back(sliceBoundMesh_);
        } else {  // :44-47: cached boundary of this slice, re-swept with the current (doubled) line counts
            sliceBound_ = sliceBoundAll_.at(sliceIdx);
            sliceBoundMesh_ = sliceBoundMeshAll_.at(sliceIdx);
            freeSpace_.cSpaceBoundary_ = sliceBound_;
            freeSpace_.cSpaceBoundaryMesh_ = sliceBoundMesh_;
            freeSpace_.setup(static_cast<unsigned>(param_.numLineY), param_.boundaryLimits[4], param_.boundaryLimits[5]);
        }
        sweepLineProcess();
        connectOneSlice3D(freeSegOneSlice_);
    }

    // :63-91
    void sweepLineProcess() {
        freeSegOneSlice_ = sweepLineProcess3D(freeSpace_, param_.boundaryLimits.data(), param_.numLineX, param_.numLineY);
        singleHitCount = freeSpace_.singleHitCount;
    }

    // :93-113
    virtual void generateVertices(double tx, const FreeSegment2D& freeSeg) {
        numVertex_.plane.clear();
        for (size_t i = 0; i < freeSeg.ty.size(); ++i) {
            numVertex_.plane.push_back(res_.graphStructure.vertex.size());
            for (size_t j = 0; j < freeSeg.xM[i].size(); ++j)
                res_.graphStructure.vertex.push_back({tx, freeSeg.ty[i], freeSeg.xM[i][j], robot_.base.quat.w, robot_.base.quat.x,
                                                      robot_.base.quat.y, robot_.base.quat.z});
        }
        numVertex_.line.push_back(numVertex_.plane);
    }

    // :116-156
    void connectOneSlice3D(const FreeSegment3D& freeSeg) {
        numVertex_.line.clear();
        numVertex_.startId = res_.graphStructure.vertex.size();
        for (size_t i = 0; i < freeSeg.tx.size(); ++i) {
            generateVertices(freeSeg.tx.at(i), freeSeg.freeSegmentYZ.at(i));
            connectOneSlice2D(freeSeg.freeSegmentYZ.at(i));
        }
        numVertex_.slice = res_.graphStructure.vertex.size();
        for (size_t i = 0; i + 1 < freeSeg.tx.size(); ++i) {
            for (size_t j = 0; j < freeSeg.freeSegmentYZ.at(i).ty.size(); ++j) {
                const size_t n1 = numVertex_.line[i][j], n2 = numVertex_.line[i + 1][j];
                for (size_t k1 = 0; k1 < freeSeg.freeSegmentYZ.at(i).xM[j].size(); ++k1)
                    for (size_t k2 = 0; k2 < freeSeg.freeSegmentYZ.at(i + 1).xM[j].size(); ++k2) {
                        if (isSameSliceTransitionFree(res_.graphStructure.vertex[n1 + k1], res_.graphStructure.vertex[n2 + k2])) {
                            res_.graphStructure.edge.push_back({n1 + k1, n2 + k2});
                            res_.graphStructure.weight.push_back(
                                vectorEuclidean(res_.graphStructure.vertex[n1 + k1], res_.graphStructure.vertex[n2 + k2]));
                        } else {
                            bridgeVertex(n1 + k1, n2 + k2);
                        }
                    }
            }
        }
    }

    // :319-365
    bool isSameSliceTransitionFree(const std::vector<double>& v1, const std::vector<double>& v2) override {
        ++numEdgeTests;
        return isSameSliceTransitionFree3D(sliceBoundMesh_.obstacle, v1.data(), v2.data(), &singleHitCount);
    }

    // :521-539
    void computeTFE(const Quat& q1, const Quat& q2, std::vector<Ellipsoid3>& tfe) {
        tfe.clear();
        tfe.push_back(getTFE3D(robot_.base.semiAxis, q1, q2, param_.numPoint));
        for (size_t i = 0; i < robot_.getNumLinks(); ++i) {
            const Mat3& rotLink = robot_.tf.at(i).R;
            tfe.push_back(getTFE3D(robot_.link.at(i).semiAxis, rot_to_quat(quat_to_rot(q1) * rotLink),
                                   rot_to_quat(quat_to_rot(q2) * rotLink), param_.numPoint));
        }
    }

    // :301-317
    void bridgeSlice() {
        bridgeSliceBound_.resize(tfe_.size());
        for (size_t i = 0; i < tfe_.size(); ++i) {
            std::vector<MeshMatrix> bdMesh;
            for (const auto& obstacle : obs_) {
                const Points bd = obstacle.getMinkSum3D(tfe_.at(i).semi, tfe_.at(i).quat, +1);
                bdMesh.push_back(getMeshFromParamSurface(bd, obstacle.getNumParam()));
            }
            bridgeSliceBound_.at(i) = bdMesh;
        }
    }

    // :394-425
    bool isPtInCFree(size_t bdIdx, const double v[3]) { return isPtInCFree3D(bridgeSliceBound_.at(bdIdx), v, &singleHitCount); }

    // :367-392
    bool isMultiSliceTransitionFree(const std::vector<double>& v1, const std::vector<double>& v2) {
        const auto vInterp = interpolateCompoundSE3Rn(v1, v2, param_.numPoint);
        for (const auto& vStep : vInterp) {
            setTransform(vStep);
            if (!isPtInCFree(0, robot_.base.position)) return false;
            for (size_t j = 0; j < robot_.getNumLinks(); ++j)
                if (!isPtInCFree(j + 1, robot_.link[j].position)) return false;
        }
        return true;
    }

    // :158-224
    void connectMultiSlice() override {
        if (vertexIdx_.size() == 1) return;
        for (size_t i = 0; i < vertexIdx_.size(); ++i) {
            double minDist = INFINITY;
            int minIdx = 0;
            for (size_t j = 0; j != i && j < param_.numSlice; ++j) {  // C9: only earlier layers
                const double dist = angular_distance(q_.at(i), q_.at(j));
                if (dist < minDist) {
                    minDist = dist;
                    minIdx = static_cast<int>(j);
                }
            }
            size_t n22 = vertexIdx_.at(i).startId;
            const size_t n_2 = vertexIdx_.at(i).slice;
            const size_t start = vertexIdx_.at(minIdx).startId;
            const size_t n2 = vertexIdx_.at(minIdx).slice;
            computeTFE(q_.at(i), q_.at(minIdx), tfe_);
            bridgeSlice();
            for (size_t m0 = start; m0 < n2; ++m0) {
                const auto v1 = res_.graphStructure.vertex.at(m0);
                for (size_t m1 = n22; m1 < n_2; ++m1) {
                    const auto v2 = res_.graphStructure.vertex[m1];
                    if (std::fabs(v1.at(0) - v2.at(0)) >
                            2.0 * (param_.boundaryLimits[1] - param_.boundaryLimits[0]) / static_cast<double>(param_.numLineX) ||
                        std::fabs(v1.at(1) - v2.at(1)) >
                            2.0 * (param_.boundaryLimits[3] - param_.boundaryLimits[2]) / static_cast<double>(param_.numLineY))
                        continue;
                    if (isMultiSliceTransitionFree(v1, v2)) {
                        res_.graphStructure.edge.push_back({m0, m1});
                        res_.graphStructure.weight.push_back(vectorEuclidean(v1, v2));
                        n22 = m1;
                        break;
                    }
                }
            }
        }
    }

    // :226-270
    void connectExistSlice(size_t sliceId) override {
        const size_t startIdCur = vertexIdx_.at(sliceId).startId, endIdCur = vertexIdx_.at(sliceId).slice;
        size_t startIdExist = vertexIdxAll_.back().at(sliceId).startId;
        const size_t endIdExist = vertexIdxAll_.back().at(sliceId).slice;
        for (size_t m0 = startIdCur; m0 < endIdCur; ++m0) {
            const auto v1 = res_.graphStructure.vertex[m0];
            for (size_t m1 = startIdExist; m1 < endIdExist; ++m1) {
                const auto v2 = res_.graphStructure.vertex[m1];
                if (std::fabs(v1.at(0) - v2.at(0)) >
                    2.0 * std::fabs(param_.boundaryLimits[1] - param_.boundaryLimits[0]) / static_cast<double>(param_.numLineX))
                    continue;
                if (std::fabs(v1.at(1) - v2.at(1)) >
                    2.0 * std::fabs(param_.boundaryLimits[3] - param_.boundaryLimits[2]) / static_cast<double>(param_.numLineY))
                    continue;
                if (isSameSliceTransitionFree(v1, v2)) {
                    res_.graphStructure.edge.push_back({m0, m1});
                    res_.graphStructure.weight.push_back(vectorEuclidean(v1, v2));
                    startIdExist = m1;
                    break;
                }
            }
        }
    }

    // :443-508 with C18 (minQuat initialised to q_[0])
    std::vector<size_t> getNearestNeighborsOnGraph(const std::vector<double>& vertex, size_t k, double radius) override {
        std::vector<size_t> idx;
        const Quat queryQuat = {vertex[3], vertex[4], vertex[5], vertex[6]};
        Quat minQuat = q_[0];
        double minQuatDist = angular_distance(queryQuat, q_[0]);
        for (const auto& q : q_) {
            const double quatDist = angular_distance(queryQuat, q);
            if (quatDist < minQuatDist) {
                minQuatDist = quatDist;
                minQuat = q;
            }
        }
        std::vector<Quat> quatList;
        for (const auto& q : q_) {
            if (angular_distance(minQuat, q) < radius) quatList.push_back(q);
            if (quatList.size() >= k) break;
        }
        const auto& V = res_.graphStructure.vertex;
        for (const auto& quatCur : quatList) {
            size_t idxSlice = 0;
            double minEuclideanDist = INFINITY;
            for (size_t i = 0; i < V.size(); ++i) {
                const double euclideanDist = vectorEuclidean(vertex, V[i]);
                if (euclideanDist < minEuclideanDist && angular_distance(quatCur, Quat{V[i][3], V[i][4], V[i][5], V[i][6]}) < 1e-6) {
                    minEuclideanDist = euclideanDist;
                    idxSlice = i;
                }
            }
            if (std::abs(vertex[0] - V[idxSlice][0]) <
                    radius * (param_.boundaryLimits[1] - param_.boundaryLimits[0]) / static_cast<double>(param_.numLineX) &&
                std::abs(vertex[1] - V[idxSlice][1]) <
                    radius * (param_.boundaryLimits[3] - param_.boundaryLimits[2]) / static_cast<double>(param_.numLineY))
                idx.push_back(idxSlice);
        }
        return idx;
    }
};

// ---- HRM2D (src/planners/HRM2D.cpp) ------------------------------------------------------------------------------------
struct HRM2D : HighwayRoadMap<MultiBodyTree2D, SuperEllipse> {
    FreeSpace2D freeSpace_;
    std::vector<double> headings_;
    BoundaryInfo sliceBound_;
    std::vector<BoundaryInfo> sliceBoundAll_;
    FreeSegment2D freeSegOneSlice_;
    std::vector<Ellipse2> tfe_;
    std::vector<std::vector<Points>> bridgeSliceBound_;
    bool perOrientationFreeSpace_;

    HRM2D(const MultiBodyTree2D& robot, const std::vector<SuperEllipse>& arena, const std::vector<SuperEllipse>& obs,
          const PlanningRequest& req, bool perOrientationFreeSpace = false)
        : HighwayRoadMap(robot, arena, obs, req), freeSpace_(robot_, arena_, obs_), perOrientationFreeSpace_(perOrientationFreeSpace) {
        freeSpace_.setup(static_cast<unsigned>(param_.numLineY), param_.boundaryLimits[0], param_.boundaryLimits[1]);  // :13-15
    }

    // :344-350
    void setTransform(const std::vector<double>& v) {
        Tf2 g;
        g.R[0][0] = std::cos(v[2]), g.R[0][1] = -std::sin(v[2]), g.R[1][0] = std::sin(v[2]), g.R[1][1] = std::cos(v[2]);
        g.t[0] = v[0], g.t[1] = v[1];
        robot_.robotTF(g);
    }

    // :46-51
    void sampleOrientations() override {
        const double dr = 2 * PI / (static_cast<double>(param_.numSlice) - 1);
        for (size_t i = 0; i < param_.numSlice; ++i) headings_.push_back(-PI + dr * static_cast<double>(i));
    }

    // :20-42
    void constructOneSlice(size_t sliceIdx) override {
        setTransform({0.0, 0.0, headings_.at(sliceIdx)});
        if (!isRefine_) {
            if (perOrientationFreeSpace_) freeSpace_.setRobot(robot_);
            freeSpace_.computeCSpaceBoundary();
            sliceBound_ = freeSpace_.cSpaceBoundary_;
            sliceBoundAll_.push_back(sliceBound_);
        } else {  // HRM2D.cpp:29-31: cached boundary of this slice, re-swept with the doubled line count
            sliceBound_ = sliceBoundAll_.at(sliceIdx);
            freeSpace_.cSpaceBoundary_ = sliceBound_;
            freeSpace_.setup(static_cast<unsigned>(param_.numLineY), param_.boundaryLimits[0], param_.boundaryLimits[1]);
        }
        freeSegOneSlice_ = sweepLineProcess2D(freeSpace_, param_.boundaryLimits.data(), param_.numLineY);
        singleHitCount = freeSpace_.singleHitCount;
        generateVertices(0.0, freeSegOneSlice_);
        connectOneSlice2D(freeSegOneSlice_);
    }

    // :72-89
    void generateVertices(double, const FreeSegment2D& freeSeg) {
        numVertex_.plane.clear();
        numVertex_.line.clear();
        numVertex_.startId = res_.graphStructure.vertex.size();
        for (size_t i = 0; i < freeSeg.ty.size(); ++i) {
            numVertex_.plane.push_back(res_.graphStructure.vertex.size());
            for (size_t j = 0; j < freeSeg.xM[i].size(); ++j)
                res_.graphStructure.vertex.push_back({freeSeg.xM[i][j], freeSeg.ty[i], robot_.base.angle});
        }
        numVertex_.line.push_back(numVertex_.plane);
    }

    // :215-232
    bool isSameSliceTransitionFree(const std::vector<double>& v1, const std::vector<double>& v2) override {
        ++numEdgeTests;
        return isSameSliceTransitionFree2D(sliceBound_.obstacle, v1.data(), v2.data());
    }

    // :352-373
    void computeTFE(double thetaA, double thetaB, std::vector<Ellipse2>& tfe) {
        tfe.clear();
        tfe.push_back(getTFE2D(robot_.base.semiAxis, thetaA, thetaB, param_.numPoint));
        for (size_t i = 0; i < robot_.getNumLinks(); ++i) {
            const Tf2& g = robot_.tf.at(i);
            const double rl = std::atan2(g.R[1][0], g.R[0][0]);  // Rotation2Dd(Matrix2d).angle()
            const double cl = std::cos(rl), sl = std::sin(rl);
            auto ang = [&](double th) {  // Rotation2Dd(Rot(th).toRotationMatrix() * rotLink).angle(); Matrix * Rotation2D = matrix product
                const double c = std::cos(th), s = std::sin(th);
                const double m00 = c * cl + (-s) * sl, m10 = s * cl + c * sl;
                return std::atan2(m10, m00);
            };
            tfe.push_back(getTFE2D(robot_.link.at(i).semiAxis, ang(thetaA), ang(thetaB), param_.numPoint));
        }
    }

    // :199-213
    void bridgeSlice() {
        bridgeSliceBound_.resize(tfe_.size());
        for (size_t i = 0; i < tfe_.size(); ++i) {
            std::vector<Points> bd;
            for (const auto& ob : obs_) bd.push_back(ob.getMinkSum2D(tfe_.at(i).semi, tfe_.at(i).angle, +1));
            bridgeSliceBound_.at(i) = bd;
        }
    }

    // :235-265
    bool isMultiSliceTransitionFree(const std::vector<double>& v1, const std::vector<double>& v2) {
        const double dt = 1.0 / (static_cast<double>(param_.numPoint) - 1);
        for (size_t i = 0; i < param_.numPoint; ++i) {
            std::vector<double> vStep;
            for (size_t j = 0; j < v1.size(); ++j)
                vStep.push_back((1.0 - static_cast<double>(i) * dt) * v1[j] + static_cast<double>(i) * dt * v2[j]);
            setTransform(vStep);
            if (!isPtInCFree2D(bridgeSliceBound_.at(0), robot_.base.position, &singleHitCount)) return false;
            for (size_t j = 0; j < robot_.getNumLinks(); ++j)
                if (!isPtInCFree2D(bridgeSliceBound_.at(j + 1), robot_.link[j].position, &singleHitCount)) return false;
        }
        return true;
    }

    // :91-156
    void connectMultiSlice() override {
        if (vertexIdx_.size() == 1) return;
        const double distAdjacency = 2.0;
        for (size_t i = 0; i < vertexIdx_.size(); ++i) {
            const size_t startIdCur = vertexIdx_.at(i).startId, endIdCur = vertexIdx_.at(i).slice;
            size_t j;
            if (i == param_.numSlice - 1 && param_.numSlice != 2)
                j = 0;
            else
                j = i + 1;
            size_t startIdAdj = vertexIdx_.at(j).startId;
            const size_t endIdAdj = vertexIdx_.at(j).slice;
            computeTFE(headings_[i], headings_[j], tfe_);
            bridgeSlice();
            for (size_t m0 = startIdCur; m0 < endIdCur; ++m0) {
                const auto v1 = res_.graphStructure.vertex[m0];
                for (size_t m1 = startIdAdj; m1 < endIdAdj; ++m1) {
                    const auto v2 = res_.graphStructure.vertex[m1];
                    if (std::fabs(v1[1] - v2[1]) > distAdjacency * std::fabs(param_.boundaryLimits[3] - param_.boundaryLimits[2]) /
                                                       static_cast<double>(param_.numLineY))
                        continue;
                    if (isMultiSliceTransitionFree(v1, v2)) {
                        res_.graphStructure.edge.push_back({m0, m1});
                        res_.graphStructure.weight.push_back(vectorEuclidean(v1, v2));
                        startIdAdj = m1;
                        break;
                    }
                }
            }
        }
    }

    // HRM2D.cpp:158-197
    void connectExistSlice(size_t sliceId) override {
        const size_t startIdCur = vertexIdx_.at(sliceId).startId, endIdCur = vertexIdx_.at(sliceId).slice;
        size_t startIdExist = vertexIdxAll_.back().at(sliceId).startId;
        const size_t endIdExist = vertexIdxAll_.back().at(sliceId).slice;
        const double distAdjacency = 2.0;
        for (size_t m0 = startIdCur; m0 < endIdCur; ++m0) {
            const auto v1 = res_.graphStructure.vertex[m0];
            for (size_t m1 = startIdExist; m1 < endIdExist; ++m1) {
                const auto v2 = res_.graphStructure.vertex[m1];
                if (std::fabs(v1[1] - v2[1]) >
                    distAdjacency * std::fabs(param_.boundaryLimits[3] - param_.boundaryLimits[2]) / static_cast<double>(param_.numLineY))
                    continue;
                if (isSameSliceTransitionFree(v1, v2)) {
                    res_.graphStructure.edge.push_back({m0, m1});
                    res_.graphStructure.weight.push_back(vectorEuclidean(v1, v2));
                    startIdExist = m1;
                    break;
                }
            }
        }
    }

    // :284-342
    std::vector<size_t> getNearestNeighborsOnGraph(const std::vector<double>& vertex, size_t k, double radius) override {
        const auto& V = res_.graphStructure.vertex;
        std::vector<size_t> idx;
        double minAngle = V[0][2];
        double minAngleDist = std::fabs(vertex[2] - minAngle);
        for (const auto& vtx : V) {
            const double angleDist = std::fabs(vertex[2] - vtx[2]);
            if (angleDist < minAngleDist) {
                minAngleDist = angleDist;
                minAngle = vtx[2];
            }
        }
        std::vector<double> angList;
        for (double heading : headings_) {
            if (std::fabs(minAngle - heading) < radius) angList.push_back(heading);
            if (angList.size() >= k) break;
        }
        for (double angCur : angList) {
            size_t idxSlice = 0;
            double minEuclideanDist = vectorEuclidean(vertex, V[0]);
            for (size_t i = 0; i < V.size(); ++i) {
                const double euclideanDist = vectorEuclidean(vertex, V[i]);
                if ((euclideanDist < minEuclideanDist) && std::fabs(V[i][2] - angCur) < EPSILON) {
                    minEuclideanDist = euclideanDist;
                    idxSlice = i;
                }
            }
            if (std::abs(vertex[1] - V[idxSlice][1]) <
                radius * (param_.boundaryLimits[1] - param_.boundaryLimits[0]) / static_cast<double>(param_.numLineY))
                idx.push_back(idxSlice);
        }
        return idx;
    }
};

}  // namespace hrmo
