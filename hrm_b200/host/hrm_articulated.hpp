// Articulated robots on the GPU-backed free space: URDF revolute tree + forward kinematics and the
// probabilistic planner, with the reference's names --
//   hrm::ParseURDF                      include/hrm/util/ParseURDF.h, src/util/ParseURDF.cpp:10-33 (urdfdom + kdl_parser + KDL FK)
//   MultiBodyTree3D::robotTF(kdl, ...)  src/datastructure/MultiBodyTree3D.cpp:55-89 (template member in hrm_host.hpp)
//   hrm::planners::ProbHRM3D            include/hrm/planners/ProbHRM3D.h, src/planners/ProbHRM3D.cpp:7-246
// orocos-KDL / urdfdom are not vendored by the reference: a URDF joint is the frame Frame(Rot(axis, q), xyz) * Frame(RPY, 0)
// (kdl_parser's toKdl), joints are numbered depth first with children in joint-name order, poses multiply from the root.
//
// B200 design: the reference builds ONE C-slice per iteration.  The draws (base rotation + joint angles) do not depend on
// the roadmap, so this planner draws `batch` slices ahead, builds their layers (K1+K2+K3) in one `hrmgpu_build_layers`
// call, and then replays the reference's sequential loop -- vertices, intra-layer edges (K4 batches), nearest-slice
// bridge (host TFE over the interpolated articulated motion, K5) and graph search -- slice by slice, so the roadmap is the
// reference's for the same draws.
#pragma once
#include "hrm_host.hpp"

#include <fstream>
#include <random>
#include <sstream>

namespace hrm {

class ParseURDF {
  public:
    explicit ParseURDF(const std::string& urdfFile) {
        std::ifstream in(urdfFile);
        if (!in) throw std::invalid_argument("Failed to parse and construct KDL tree: cannot open " + urdfFile);
        std::stringstream buf;
        buf << in.rdbuf();
        std::string xml = buf.str();
        for (size_t a; (a = xml.find("<!--")) != std::string::npos;) {
            const size_t b = xml.find("-->", a);
            xml.erase(a, (b == std::string::npos ? xml.size() : b + 3) - a);
        }
        std::vector<JointDesc> joints;
        for (size_t p = 0; (p = xml.find("<joint ", p)) != std::string::npos;) {
            const size_t headEnd = xml.find('>', p), end = xml.find("</joint>", p);
            if (end == std::string::npos) break;
            const std::string head = xml.substr(p, headEnd - p + 1), body = xml.substr(headEnd + 1, end - headEnd - 1);
            JointDesc j;
            j.name = attr(head, "name"), j.type = attr(head, "type");
            j.parent = attr(element(body, "parent"), "link"), j.child = attr(element(body, "child"), "link");
            triple(attr(element(body, "origin"), "xyz"), j.xyz);
            triple(attr(element(body, "origin"), "rpy"), j.rpy);
            triple(attr(element(body, "axis"), "xyz"), j.axis);
            joints.push_back(j);
            p = end;
        }
        std::map<std::string, std::vector<const JointDesc*>> children;
        std::map<std::string, bool> isChild;
        for (const auto& j : joints) children[j.parent].push_back(&j), isChild[j.child] = true;
        for (auto& c : children)
            std::sort(c.second.begin(), c.second.end(), [](const JointDesc* a, const JointDesc* b) { return a->name < b->name; });
        for (const auto& j : joints)
            if (!isChild.count(j.parent)) root_ = j.parent;
        addChildren(root_, children);
    }

    size_t getNrOfJoints() const { return numJoints_; }
    const std::string& getRootName() const { return root_; }

    /** pose of link `bodyName` in the root frame for the joint angles `jointConfig` (rotation passed through a quaternion) */
    SE3Transform getTransform(const std::vector<double>& jointConfig, const std::string& bodyName) const {
        SE3Transform f = pose(jointConfig, bodyName);
        f.R = toRotationMatrix(toQuaternion(f.R));
        return f;
    }

  private:
    struct JointDesc {
        std::string name, type, parent, child;
        double xyz[3] = {0, 0, 0}, rpy[3] = {0, 0, 0}, axis[3] = {1, 0, 0};
    };
    struct Segment {
        std::string parent;
        Mat3 originR;
        Vec3 origin, axis;
        int q = -1;
    };
    static std::string attr(const std::string& tag, const std::string& key) {
        const std::string k = " " + key + "=\"";
        size_t p = tag.find(k);
        if (p == std::string::npos) return "";
        p += k.size();
        return tag.substr(p, tag.find('"', p) - p);
    }
    static std::string element(const std::string& body, const std::string& name) {
        const size_t a = body.find("<" + name);
        return a == std::string::npos ? std::string() : body.substr(a, body.find('>', a) - a + 1);
    }
    static void triple(const std::string& s, double out[3]) {
        if (s.empty()) return;
        std::istringstream in(s);
        in >> out[0] >> out[1] >> out[2];
    }
    static Mat3 rotAxis(const Vec3& v, double angle) {  // KDL::Rotation::Rot2
        const double ct = std::cos(angle), st = std::sin(angle), vt = 1.0 - ct;
        const double mvt0 = vt * v[0], mvt1 = vt * v[1], mvt2 = vt * v[2];
        const double mst0 = v[0] * st, mst1 = v[1] * st, mst2 = v[2] * st;
        const double mvt01 = mvt0 * v[1], mvt02 = mvt0 * v[2], mvt12 = mvt1 * v[2];
        Mat3 r;
        r[0] = {ct + mvt0 * v[0], -mst2 + mvt01, mst1 + mvt02};
        r[1] = {mst2 + mvt01, ct + mvt1 * v[1], -mst0 + mvt12};
        r[2] = {-mst1 + mvt02, mst0 + mvt12, ct + mvt2 * v[2]};
        return r;
    }
    static Mat3 rotRPY(double roll, double pitch, double yaw) {  // KDL::Rotation::RPY = Rz(yaw) Ry(pitch) Rx(roll)
        const double ca1 = std::cos(yaw), sa1 = std::sin(yaw), cb1 = std::cos(pitch), sb1 = std::sin(pitch), cc1 = std::cos(roll),
                     sc1 = std::sin(roll);
        Mat3 r;
        r[0] = {ca1 * cb1, ca1 * sb1 * sc1 - sa1 * cc1, ca1 * sb1 * cc1 + sa1 * sc1};
        r[1] = {sa1 * cb1, sa1 * sb1 * sc1 + ca1 * cc1, sa1 * sb1 * cc1 - ca1 * sc1};
        r[2] = {-sb1, cb1 * sc1, cb1 * cc1};
        return r;
    }
    void addChildren(const std::string& link, const std::map<std::string, std::vector<const JointDesc*>>& children) {
        const auto it = children.find(link);
        if (it == children.end()) return;
        for (const JointDesc* j : it->second) {
            Segment s;
            s.parent = j->parent;
            s.originR = rotRPY(j->rpy[0], j->rpy[1], j->rpy[2]);
            s.origin = {j->xyz[0], j->xyz[1], j->xyz[2]};
            const Vec3 a = matVec(s.originR, {j->axis[0], j->axis[1], j->axis[2]});
            const double nrm = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
            s.axis = {a[0] / nrm, a[1] / nrm, a[2] / nrm};
            if (j->type == "revolute" || j->type == "continuous") s.q = static_cast<int>(numJoints_++);
            segments_[j->child] = s;
            addChildren(j->child, children);
        }
    }
    SE3Transform pose(const std::vector<double>& q, const std::string& link) const {
        SE3Transform id;
        if (link == root_) return id;
        const auto it = segments_.find(link);
        if (it == segments_.end()) throw std::invalid_argument("Error in solving forward kinematics: no link " + link);
        const Segment& s = it->second;
        SE3Transform cur;  // Segment::pose(q) = Frame(Rot2(axis, q), origin) * Frame(originR, 0)
        const Mat3 eye = id.R;
        cur.R = matMul(s.q >= 0 ? rotAxis(s.axis, q.at(static_cast<size_t>(s.q))) : eye, s.originR);
        cur.t = s.origin;
        if (s.parent == root_) return cur;
        const SE3Transform par = pose(q, s.parent);
        SE3Transform out;
        out.R = matMul(par.R, cur.R);
        const Vec3 v = matVec(par.R, cur.t);
        out.t = {v[0] + par.t[0], v[1] + par.t[1], v[2] + par.t[2]};
        return out;
    }

    std::map<std::string, Segment> segments_;  // keyed by child link
    std::string root_;
    size_t numJoints_ = 0;
};

namespace planners {

class ProbHRM3D : public HRM3D {
  public:
    ProbHRM3D(const MultiBodyTree3D& robot, const std::string& urdfFile, const std::vector<SuperQuadrics>& arena,
              const std::vector<SuperQuadrics>& obs, const PlanningRequest& req, int device = 0, int precision = HRMGPU_F64_STRICT,
              bool perOrientation = false)
        : HRM3D(robot, arena, obs, req, device, precision, perOrientation), urdfFile_(urdfFile), kdl_(urdfFile) {
        q_.clear();
        param_.numSlice = 0;
    }

    /** The reference draws Quaterniond::UnitRandom() and ompl::RNG::uniformReal(-pi/2, pi/2) per slice (ProbHRM3D.cpp:67-96,
     * srand(time)): rows [qw,qx,qy,qz, joints...] given here replace those draws (slices 2, 3, ...); planning stops when the
     * list is exhausted.  Without a list the draws come from a seeded std::mt19937_64. */
    void setSampleList(const std::vector<std::vector<double>>& samples) { samples_ = samples, injected_ = true; }
    void setSeed(unsigned long long seed) { rng_.seed(seed); }
    void setBatch(size_t batch) { batch_ = batch > 0 ? batch : 1; }
    void setMaxSlices(size_t n) { maxSlices_ = n; }
    /** sweep lines of ALL slices doubled every `period` slices, at most numPoint times (ProbHRM3D.cpp:52-56: 60); 0 = never */
    void setRefinePeriod(size_t period) { refinePeriod_ = period; }
    Index numSlices() const { return param_.numSlice; }

    /** ProbHRM3D.cpp:18-65 */
    void plan(double timeLim = 0.0) override {
        auto t0 = std::chrono::high_resolution_clock::now();
        const auto& lim = param_.boundaryLimits;
        int Lx = 0, Ly = 0;
        std::vector<double> tx, ty;
        auto setGrid = [&]() {  // (again after every refinement: the line counts double)
            Lx = static_cast<int>(param_.numLineX), Ly = static_cast<int>(param_.numLineY);
            freeSpace_.setup(lim, param_.numLineX, param_.numLineY);
            const double dx = (lim[1] - lim[0]) / static_cast<double>(Lx - 1), dy = (lim[3] - lim[2]) / static_cast<double>(Ly - 1);
            tx.assign(Lx, 0.0), ty.assign(Ly, 0.0);
            for (int i = 0; i < Lx; ++i) tx[i] = lim[0] + static_cast<double>(i) * dx;
            for (int i = 0; i < Ly; ++i) ty[i] = lim[2] + static_cast<double>(i) * dy;
        };
        setGrid();

        param_.numSlice = 0;
        size_t batchStart = 0, batchEnd = 0;  // slices [batchStart, batchEnd) have their layers on the device
        std::vector<long long> off;
        std::vector<double> xL, xU, xM;
        std::vector<Quaternion> baseQ;
        do {
            const size_t s = param_.numSlice;
            if (s >= batchEnd) {  // draw the next batch of robot shapes and build their C-layers in one call
                size_t want = s < 2 ? 2 : batch_;
                if (refinePeriod_ > 0 && want > refinePeriod_ - s % refinePeriod_) want = refinePeriod_ - s % refinePeriod_;  // (a refinement rebuilds every layer)
                size_t have = 0;
                std::vector<MultiBodyTree3D> poses;
                baseQ.clear();
                while (have < want && drawShape(s + have)) {
                    setTransform(v_.at(s + have));
                    baseQ.push_back(robot_.getBase().getQuaternion());
                    poses.push_back(perOrientation_ ? robot_ : robot0_);  // the reference's FreeSpace copy keeps the as-loaded pose (finding 1)
                    ++have;
                }
                if (have == 0) break;  // injected draws exhausted
                {
                    StageScope ss(stage_, "buildLayers (K1-K3)");
                    freeSpace_.buildLayers(poses, precision_);
                    freeSpace_.getFreeSegmentsAll(off, xL, xU, xM);
                }
                batchStart = s, batchEnd = s + have;
                batchCode_.clear();
                if (deviceMatching_) {
                    // connectOneSlice's candidate segments of the WHOLE look-ahead batch in one call (they depend on the free
                    // segments only); the slices below replay the reference's loops on the cached codes, one after the other
                    StageScope ss(stage_, "K4 connectSlices (device, per batch)");
                    batchSp_.assign(have, SlicePairs());
                    size_t total = 0;
                    for (size_t k = 0; k < have; ++k) {
                        const long long* o = off.data() + k * static_cast<size_t>(Lx) * Ly;
                        auto segN = [&](int ix, int iy) { return static_cast<size_t>(o[ix * Ly + iy + 1] - o[ix * Ly + iy]); };
                        batchSp_[k].first = total;
                        for (int ix = 0; ix < Lx; ++ix)
                            for (int iy = 0; iy + 1 < Ly; ++iy) batchSp_[k].nInPlane += segN(ix, iy) * segN(ix, iy + 1);
                        batchSp_[k].P = batchSp_[k].nInPlane;
                        for (int ix = 0; ix + 1 < Lx; ++ix)
                            for (int iy = 0; iy < Ly; ++iy) batchSp_[k].P += segN(ix, iy) * segN(ix + 1, iy);
                        total += batchSp_[k].P;
                    }
                    std::vector<long long> layerFirst;
                    batchCode_ = freeSpace_.connectSlices3D(layerFirst, total, have);
                    if (batchCode_.size() != total) throw std::runtime_error("connectSlices3D: pair count mismatch");
                }
                staged_.clear();
                if (deviceMatching_ && batchBridge_) stageBatch(s, have, baseQ, tx, ty, off, xL, xU, xM);
            }
            // constructOneSlice: vertices + intra-layer edges of slice s
            const size_t k = s - batchStart;
            vertexTail_.assign(v_.at(s).begin() + 7, v_.at(s).end());
            q_.push_back({v_.at(s)[3], v_.at(s)[4], v_.at(s)[5], v_.at(s)[6]});
            if (!staged_.empty()) {
                // the slice was built with its batch (stageBatch): its vertices, intra-slice edges and bridge edges enter the roadmap
                // now, in the reference's order, so that search() sees exactly the roadmap the reference has after this slice
                StageScope ss(stage_, "append staged slice");
                StagedSlice& st = staged_.at(k);
                auto& g = res_.graphStructure;
                g.vertex.insert(g.vertex.end(), std::make_move_iterator(st.vertex.begin()), std::make_move_iterator(st.vertex.end()));
                g.edge.insert(g.edge.end(), st.edge.begin(), st.edge.end());
                g.weight.insert(g.weight.end(), st.weight.begin(), st.weight.end());
                vertexIdx_.push_back(st.vertexIdx);
                numEdgeTests_ += st.edgeTests;
                param_.numSlice++;
                for (const auto& e : st.bridge) addEdge(e.first, e.second);
            } else {
            {
                StageScope ss(stage_, "connectOneSlice");
                if (deviceMatching_) {
                    const SlicePairs& sp = batchSp_.at(k);
                    numEdgeTests_ += static_cast<long>(5 * sp.P);
                    const std::vector<int> codes(batchCode_.begin() + static_cast<long>(sp.first), batchCode_.begin() + static_cast<long>(sp.first + sp.P));
                    replaySlice3D(baseQ.at(k), tx, ty, off.data() + k * static_cast<size_t>(Lx) * Ly, xL, xU, xM, sp, codes);
                } else {
                    connectOneSlice3D(static_cast<int>(k), baseQ.at(k), tx, ty, off.data() + k * static_cast<size_t>(Lx) * Ly, xL, xU, xM);
                }
            }
            param_.numSlice++;
            if (param_.numSlice >= 2) {
                StageScope ss(stage_, "connectNearestSlice");
                connectNearestSlice();
            }
            }
            res_.planningTime.buildTime += seconds(t0);
            t0 = std::chrono::high_resolution_clock::now();
            search();
            stage_.add("search", seconds(t0));
            res_.planningTime.searchTime += seconds(t0);
            t0 = std::chrono::high_resolution_clock::now();
            res_.planningTime.totalTime = res_.planningTime.buildTime + res_.planningTime.searchTime;
            // ProbHRM3D.cpp:52-56: double the sweep lines of all slices every `refinePeriod_` slices (at most numPoint times)
            if (refinePeriod_ > 0 && param_.numSlice % refinePeriod_ == 0 && vertexIdxAll_.size() < param_.numPoint) {
                StageScope ss(stage_, "refineExistRoadmap");
                refineAllSlices(setGrid, tx, ty, timeLim, t0);
                batchStart = batchEnd = 0;  // the layers on the device are the refined ones: the next slice starts a new batch
            }
        } while (!res_.solved && param_.numSlice < maxSlices_ && (timeLim <= 0.0 || res_.planningTime.totalTime < timeLim));
        if (res_.solved) {
            const size_t pose = res_.graphStructure.vertex.at(0).size();
            start_.resize(pose), goal_.resize(pose);
            res_.solutionPath.solvedPath.push_back(start_);
            for (Index id : res_.solutionPath.PathId) res_.solutionPath.solvedPath.push_back(res_.graphStructure.vertex.at(id));
            res_.solutionPath.solvedPath.push_back(goal_);
            res_.solutionPath.interpolatedPath = getInterpolatedSolutionPath(param_.numPoint);
        }
        stage_.add("plan total", res_.planningTime.totalTime);
        stage_.print();
    }

  protected:
    static double seconds(std::chrono::high_resolution_clock::time_point t0) {
        return std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
    }

    /** refineExistRoadmap (HighwayRoadMap-inl.h:176-215) for the articulated planner: the C-layers of ALL slices so far are
     * rebuilt in ONE call at the doubled line counts (every slice on its own robot shape, SURVEY.md §8 f.4 -- the reference's
     * FreeSpace object only remembers the last slice), their candidate segments tested in one device call, and the
     * reference's per-slice loop (new vertices, connectOneSlice, connectExistSlice with the slice's previous vertices, search)
     * is replayed on the answers. */
    template <class SetGrid>
    void refineAllSlices(SetGrid& setGrid, std::vector<double>& tx, std::vector<double>& ty, double timeLim,
                         std::chrono::high_resolution_clock::time_point& t0) {
        vertexIdxAll_.push_back(vertexIdx_);
        vertexIdx_.clear();
        param_.numLineX *= 2;
        param_.numLineY *= 2;
        setGrid();
        const auto& lim = param_.boundaryLimits;
        const int Lx = static_cast<int>(param_.numLineX), Ly = static_cast<int>(param_.numLineY);
        const size_t K = param_.numSlice;
        std::vector<MultiBodyTree3D> poses;
        std::vector<Quaternion> baseQ;
        for (size_t k = 0; k < K; ++k) {
            setTransform(v_.at(k));
            baseQ.push_back(robot_.getBase().getQuaternion());
            poses.push_back(perOrientation_ ? robot_ : robot0_);
        }
        std::vector<long long> off;
        std::vector<double> xL, xU, xM;
        freeSpace_.buildLayers(poses, precision_);
        freeSpace_.getFreeSegmentsAll(off, xL, xU, xM);
        std::vector<SlicePairs> sp(K);
        size_t total = 0;
        for (size_t k = 0; k < K; ++k) {
            const long long* o = off.data() + k * static_cast<size_t>(Lx) * Ly;
            auto segN = [&](int ix, int iy) { return static_cast<size_t>(o[ix * Ly + iy + 1] - o[ix * Ly + iy]); };
            sp[k].first = total;
            for (int ix = 0; ix < Lx; ++ix)
                for (int iy = 0; iy + 1 < Ly; ++iy) sp[k].nInPlane += segN(ix, iy) * segN(ix, iy + 1);
            sp[k].P = sp[k].nInPlane;
            for (int ix = 0; ix + 1 < Lx; ++ix)
                for (int iy = 0; iy < Ly; ++iy) sp[k].P += segN(ix, iy) * segN(ix + 1, iy);
            total += sp[k].P;
        }
        std::vector<long long> layerFirst;
        const std::vector<unsigned char> code = freeSpace_.connectSlices3D(layerFirst, total, K);
        if (code.size() != total) throw std::runtime_error("connectSlices3D: pair count mismatch");
        const double thx = 2.0 * std::fabs(lim[1] - lim[0]) / static_cast<double>(param_.numLineX);
        const double thy = 2.0 * std::fabs(lim[3] - lim[2]) / static_cast<double>(param_.numLineY);
        for (size_t k = 0; k < K; ++k) {
            // generateVertices (ProbHRM3D.cpp:164-194) copies the joint angles of v_.back(), i.e. of the NEWEST slice, also while an
            // older slice is being refined: kept, so that the roadmap is the reference's (here v_ may already hold look-ahead draws)
            vertexTail_.assign(v_.at(K - 1).begin() + 7, v_.at(K - 1).end());
            numEdgeTests_ += static_cast<long>(5 * sp[k].P);
            const std::vector<int> codes(code.begin() + static_cast<long>(sp[k].first), code.begin() + static_cast<long>(sp[k].first + sp[k].P));
            replaySlice3D(baseQ[k], tx, ty, off.data() + k * static_cast<size_t>(Lx) * Ly, xL, xU, xM, sp[k], codes);
            const auto& V = res_.graphStructure.vertex;
            connectExistSliceReplay(
                k, [&](const std::vector<Coordinate>& a, const std::vector<Coordinate>& b) { return !(std::fabs(a[0] - b[0]) > thx) && !(std::fabs(a[1] - b[1]) > thy); },
                [&](const std::vector<std::pair<Index, Index>>& cand) {
                    std::vector<double> A, Bq;
                    for (const auto& c : cand)
                        for (int d = 0; d < 3; ++d) A.push_back(V[c.first][d]), Bq.push_back(V[c.second][d]);
                    return freeSpace_.testEdges(static_cast<int>(k), A, Bq, 3);
                });
            res_.planningTime.buildTime += seconds(t0);
            t0 = std::chrono::high_resolution_clock::now();
            search();
            res_.planningTime.searchTime += seconds(t0);
            t0 = std::chrono::high_resolution_clock::now();
            res_.planningTime.totalTime = res_.planningTime.buildTime + res_.planningTime.searchTime;
            if (res_.solved || (timeLim > 0.0 && res_.planningTime.totalTime > timeLim)) return;
        }
    }

    /** sampleOrientations (:67-102) for slice s; false when the injected list has no more rows */
    bool drawShape(size_t s) {
        if (s < v_.size()) return true;
        const size_t nj = kdl_.getNrOfJoints();
        std::vector<Coordinate> config{0.0, 0.0, 0.0};
        if (s == 0) {
            config.insert(config.end(), start_.begin() + 3, start_.begin() + 7 + static_cast<long>(nj));
        } else if (s == 1) {
            config.insert(config.end(), goal_.begin() + 3, goal_.begin() + 7 + static_cast<long>(nj));
        } else if (injected_) {
            if (s - 2 >= samples_.size()) return false;
            config.insert(config.end(), samples_[s - 2].begin(), samples_[s - 2].begin() + 4 + static_cast<long>(nj));
        } else {
            std::uniform_real_distribution<double> u(0.0, 1.0);  // Shoemake's uniform unit quaternion
            const double u1 = u(rng_), u2 = u(rng_), u3 = u(rng_), twoPi = 6.283185307179586;
            config.push_back(std::sqrt(u1) * std::cos(twoPi * u3));
            config.push_back(std::sqrt(1.0 - u1) * std::sin(twoPi * u2));
            config.push_back(std::sqrt(1.0 - u1) * std::cos(twoPi * u2));
            config.push_back(std::sqrt(u1) * std::sin(twoPi * u3));
            for (size_t i = 0; i < nj; ++i) config.push_back((2.0 * u(rng_) - 1.0) * maxJointAngle_);
        }
        v_.push_back(config);
        return true;
    }

    /** setTransform (:198-213) */
    void setTransform(const std::vector<Coordinate>& V) {
        SE3Transform g;
        g.R = toRotationMatrix({V[3], V[4], V[5], V[6]});
        g.t = {V[0], V[1], V[2]};
        robot_.robotTF(kdl_, g, std::vector<double>(V.begin() + 7, V.begin() + 7 + static_cast<long>(kdl_.getNrOfJoints())));
    }

    /** computeTFE (:216-246) */
    std::vector<Ellipsoid> computeTFE(const std::vector<Coordinate>& v1, const std::vector<Coordinate>& v2) {
        const auto vInterp = interpolateCompoundSE3Rn(v1, v2, param_.numPoint);
        setTransform(v1);
        auto bodies = robot_.getBodyShapes();
        std::vector<Ellipsoid> mvce;
        for (const auto& b : bodies) mvce.push_back({{b.getSemiAxis()[0], b.getSemiAxis()[1], b.getSemiAxis()[2]}, b.getQuaternion()});
        for (const auto& vStep : vInterp) {
            setTransform(vStep);
            bodies = robot_.getBodyShapes();
            for (size_t i = 0; i < bodies.size(); ++i)
                mvce[i] = getMVCE3D(mvce[i].semi, {bodies[i].getSemiAxis()[0], bodies[i].getSemiAxis()[1], bodies[i].getSemiAxis()[2]}, mvce[i].quat,
                                    bodies[i].getQuaternion());
        }
        return mvce;
    }

    /** One slice of a look-ahead batch, built ahead of its turn (stageBatch) */
    struct StagedSlice {
        std::vector<std::vector<Coordinate>> vertex;
        std::vector<std::pair<Index, Index>> edge;  // intra-slice edges (connectOneSlice)
        std::vector<double> weight;
        std::pair<Index, Index> vertexIdx;
        std::vector<std::pair<Index, Index>> bridge;  // edges to the nearest earlier slice (connectMultiSlice), in m0 order
        long edgeTests = 0;
    };

    /** Interpolation weights and the chain / link offsets of the interpolated poses between two slices (the part of
     * isMultiSliceTransitionFree that does not depend on the candidate, see connectNearestSlice) appended for one slice pair */
    void appendPairOffsets(const std::vector<Coordinate>& vNear, const std::vector<Coordinate>& vNew, std::vector<double>& linkOff,
                           std::vector<double>& chainOff) {
        const size_t steps = param_.numPoint, T = steps + 1;
        const size_t B = robot_.getNumLinks() + 1;
        const size_t base = linkOff.size();
        linkOff.resize(base + T * B * 3, 0.0), chainOff.resize(base + T * B * 3, 0.0);
        const auto vi = interpolateCompoundSE3Rn(vNear, vNew, steps);
        const auto& tf = robot_.getTF();
        const size_t nj = kdl_.getNrOfJoints();
        for (size_t s = 0; s < T; ++s) {
            const Mat3 R = toRotationMatrix({vi[s][3], vi[s][4], vi[s][5], vi[s][6]});
            const std::vector<double> joints(vi[s].begin() + 7, vi[s].begin() + 7 + static_cast<long>(nj));
            for (size_t l = 0; l + 1 < B; ++l) {
                const SE3Transform fk = kdl_.getTransform(joints, "body" + std::to_string(l + 1));
                const Vec3 c = matVec(R, fk.t);
                const Vec3 v = matVec(matMul(R, fk.R), tf[l].t);
                for (int d = 0; d < 3; ++d) chainOff[base + (s * B + (l + 1)) * 3 + d] = c[d], linkOff[base + (s * B + (l + 1)) * 3 + d] = v[d];
            }
        }
    }

    /** The whole look-ahead batch ahead of its turn (bridge + K5 of a batch in one launch each).  The robot
     * shapes of the batch are known, and neither the vertices of a slice nor its bridge to the nearest earlier slice depend on
     * the searches in between: (1) connectOneSlice of every slice is replayed on the batch's pair codes (vertex ids are
     * consecutive, bridging adds no vertices), (2) the tightly-fitted ellipsoids of ALL slice pairs go into ONE bridge build and
     * the candidate pairs of all of them into ONE hrmgpu_connect_pairs3d call, (3) everything is taken out of the roadmap again
     * and kept per slice; plan() appends slice after slice and searches in between, exactly like the reference. */
    void stageBatch(size_t s0, size_t have, const std::vector<Quaternion>& baseQ, const std::vector<double>& tx, const std::vector<double>& ty,
                    const std::vector<long long>& off, const std::vector<double>& xL, const std::vector<double>& xU, const std::vector<double>& xM) {
        auto& g = res_.graphStructure;
        const size_t V0 = g.vertex.size(), E0 = g.edge.size(), I0 = vertexIdx_.size();
        const long tests0 = numEdgeTests_;
        const size_t L = tx.size() * ty.size();
        std::vector<size_t> vEnd(have), eEnd(have);
        std::vector<long> tests(have);
        {
            StageScope ss(stage_, "connectOneSlice");
            for (size_t k = 0; k < have; ++k) {
                vertexTail_.assign(v_.at(s0 + k).begin() + 7, v_.at(s0 + k).end());
                const SlicePairs& sp = batchSp_.at(k);
                const std::vector<int> codes(batchCode_.begin() + static_cast<long>(sp.first), batchCode_.begin() + static_cast<long>(sp.first + sp.P));
                replaySlice3D(baseQ.at(k), tx, ty, off.data() + k * L, xL, xU, xM, sp, codes);
                vEnd[k] = g.vertex.size(), eEnd[k] = g.edge.size(), tests[k] = static_cast<long>(5 * sp.P);
            }
        }
        // bridges: slice s0 + k against its nearest earlier slice (ProbHRM3D.cpp:105-162)
        std::vector<std::vector<Ellipsoid>> tfe;
        std::vector<int> range4;
        std::vector<size_t> pairSlice;
        std::vector<double> linkOff, chainOff;
        {
            StageScope ss(stage_, "connectNearestSlice");
            auto tT = std::chrono::high_resolution_clock::now();
            for (size_t k = 0; k < have; ++k) {
                const size_t cur = s0 + k;
                if (cur == 0) continue;
                double minDist = INFINITY;
                size_t minIdx = 0;
                for (size_t i = 0; i < cur; ++i) {
                    const double dist = vectorEuclidean(v_.at(cur), v_.at(i));
                    if (dist < minDist) minDist = dist, minIdx = i;
                }
                const Index new0 = vertexIdx_.at(cur).first, new1 = vertexIdx_.at(cur).second;
                const Index start = vertexIdx_.at(minIdx).first, n1 = vertexIdx_.at(minIdx).second;
                if (n1 <= start || new1 <= new0) continue;
                tfe.push_back(computeTFE(v_.at(cur), v_.at(minIdx)));
                appendPairOffsets(g.vertex[start], g.vertex[new0], linkOff, chainOff);
                range4.insert(range4.end(), {static_cast<int>(start), static_cast<int>(n1), static_cast<int>(new0), static_cast<int>(new1)});
                pairSlice.push_back(k);
            }
            stage_.add("  TFE + FK of the interpolated poses (host)", seconds(tT));
        }
        std::vector<StagedSlice> out(have);
        if (!pairSlice.empty()) {
            StageScope ss(stage_, "connectNearestSlice");
            const auto& lim = param_.boundaryLimits;
            const double thx = 2.0 * (lim[1] - lim[0]) / static_cast<double>(param_.numLineX);
            const double thy = 2.0 * (lim[3] - lim[2]) / static_cast<double>(param_.numLineY);
            const size_t steps = param_.numPoint, T = steps + 1;
            std::vector<double> w0(T), w1(T);
            const double dt = 1.0 / static_cast<double>(steps);
            for (size_t t = 0; t < T; ++t) w1[t] = static_cast<double>(t) * dt, w0[t] = 1.0 - static_cast<double>(t) * dt;  // InterpolateSE3.cpp:5-29
            {
                StageScope s2(stage_, "  K5a buildBridge (batch)");
                freeSpace_.buildBridge(tfe, precision_);
            }
            std::vector<double> xyz(3 * g.vertex.size());
            for (size_t m = 0; m < g.vertex.size(); ++m)
                for (int d = 0; d < 3; ++d) xyz[3 * m + d] = g.vertex[m][d];
            std::vector<int> match;
            {
                StageScope s2(stage_, "  K5b connectPairs (device, batch)");
                match = freeSpace_.connectPairs3D(xyz, range4, thx, thy, w0, w1, linkOff, &numBridgeCandidates_, &chainOff);
            }
            size_t first = 0;
            for (size_t pi = 0; pi < pairSlice.size(); ++pi) {
                const Index start = static_cast<Index>(range4[4 * pi]), n1 = static_cast<Index>(range4[4 * pi + 1]);
                for (Index m0 = start; m0 < n1; ++m0)
                    if (match.at(first + (m0 - start)) >= 0) out[pairSlice[pi]].bridge.push_back({m0, static_cast<Index>(match[first + (m0 - start)])});
                first += n1 - start;
            }
        }
        // take the batch out of the roadmap again
        for (size_t k = 0; k < have; ++k) {
            const size_t vb = k ? vEnd[k - 1] : V0, eb = k ? eEnd[k - 1] : E0;
            out[k].vertex.assign(std::make_move_iterator(g.vertex.begin() + static_cast<long>(vb)), std::make_move_iterator(g.vertex.begin() + static_cast<long>(vEnd[k])));
            out[k].edge.assign(g.edge.begin() + static_cast<long>(eb), g.edge.begin() + static_cast<long>(eEnd[k]));
            out[k].weight.assign(g.weight.begin() + static_cast<long>(eb), g.weight.begin() + static_cast<long>(eEnd[k]));
            out[k].vertexIdx = vertexIdx_.at(I0 + k);
            out[k].edgeTests = tests[k];
        }
        g.vertex.resize(V0), g.edge.resize(E0), g.weight.resize(E0), vertexIdx_.resize(I0);
        numEdgeTests_ = tests0;
        staged_ = std::move(out);
    }

    /** connectMultiSlice (:105-162): the newest slice against its nearest earlier one; K5 batch + host replay */
    void connectNearestSlice() {
        if (param_.numSlice == 1) return;
        auto& V = res_.graphStructure.vertex;
        const auto& lim = param_.boundaryLimits;
        double minDist = INFINITY;
        size_t minIdx = 0;
        for (size_t i = 0; i + 1 < param_.numSlice; ++i) {
            const double dist = vectorEuclidean(v_.at(param_.numSlice - 1), v_.at(i));
            if (dist < minDist) minDist = dist, minIdx = i;
        }
        const Index new0 = vertexIdx_.at(param_.numSlice - 1).first, new1 = vertexIdx_.at(param_.numSlice - 1).second;
        const Index start = vertexIdx_.at(minIdx).first, n1 = vertexIdx_.at(minIdx).second;
        auto tT = std::chrono::high_resolution_clock::now();
        const auto tfe = computeTFE(v_.at(param_.numSlice - 1), v_.at(minIdx));
        stage_.add("  TFE (host)", seconds(tT));
        if (n1 <= start || new1 <= new0) return;
        {
            StageScope ss(stage_, "  K5a buildBridge");
            freeSpace_.buildBridge({tfe}, precision_);
        }
        tT = std::chrono::high_resolution_clock::now();
        const double thx = 2.0 * (lim[1] - lim[0]) / static_cast<double>(param_.numLineX);
        const double thy = 2.0 * (lim[3] - lim[2]) / static_cast<double>(param_.numLineY);
        const size_t steps = param_.numPoint, T = steps + 1;
        // isMultiSliceTransitionFree (HRM3D.cpp:367-392) for every candidate pair in ONE launch with the interpolation on the device.
        // Every vertex of a slice carries the slice's base orientation and joint angles, so the interpolated rotations, joint
        // angles and with them the chain / link offsets of robotTF(urdf, gBase, joints) are the same for all candidates of the
        // slice pair: T forward-kinematics evaluations per link instead of one per (candidate, pose).  Body centre of link i at
        // pose s = linkOff + (chainOff + t) exactly as robotTF adds them (a.t = R fk.t + t; position = (a.R tf.t) + a.t).
        const size_t B = robot_.getNumLinks() + 1;
        std::vector<double> w0(T), w1(T), linkOff(T * B * 3, 0.0), chainOff(T * B * 3, 0.0);
        {
            const double dt = 1.0 / static_cast<double>(steps);
            for (size_t s = 0; s < T; ++s) w1[s] = static_cast<double>(s) * dt, w0[s] = 1.0 - static_cast<double>(s) * dt;  // InterpolateSE3.cpp:5-29
            const auto vi = interpolateCompoundSE3Rn(V[start], V[new0], steps);  // rotation and joint part only depend on the slices
            const auto& tf = robot_.getTF();
            const size_t nj = kdl_.getNrOfJoints();
            for (size_t s = 0; s < T; ++s) {
                const Mat3 R = toRotationMatrix({vi[s][3], vi[s][4], vi[s][5], vi[s][6]});
                const std::vector<double> joints(vi[s].begin() + 7, vi[s].begin() + 7 + static_cast<long>(nj));
                for (size_t l = 0; l + 1 < B; ++l) {
                    const SE3Transform fk = kdl_.getTransform(joints, "body" + std::to_string(l + 1));
                    const Vec3 c = matVec(R, fk.t);
                    const Vec3 v = matVec(matMul(R, fk.R), tf[l].t);
                    for (int d = 0; d < 3; ++d) chainOff[(s * B + (l + 1)) * 3 + d] = c[d], linkOff[(s * B + (l + 1)) * 3 + d] = v[d];
                }
            }
        }
        if (deviceMatching_) {
            // candidate enumeration, transition tests (articulated variant) and the replay of the moving lower bound on the device
            stage_.add("  FK of the interpolated poses (host)", seconds(tT));
            std::vector<double> xyz(3 * V.size());
            for (size_t m = 0; m < V.size(); ++m)
                for (int d = 0; d < 3; ++d) xyz[3 * m + d] = V[m][d];
            const std::vector<int> range4 = {static_cast<int>(start), static_cast<int>(n1), static_cast<int>(new0), static_cast<int>(new1)};
            std::vector<int> match;
            {
                StageScope ss(stage_, "  K5b connectPairs (device)");
                match = freeSpace_.connectPairs3D(xyz, range4, thx, thy, w0, w1, linkOff, &numBridgeCandidates_, &chainOff);
            }
            for (Index m0 = start; m0 < n1; ++m0)
                if (match[m0 - start] >= 0) addEdge(m0, static_cast<Index>(match[m0 - start]));
            return;
        }
        std::vector<std::pair<Index, Index>> cand;
        std::vector<double> v0s, v1s;
        for (Index m0 = start; m0 < n1; ++m0)
            for (Index m1 = new0; m1 < new1; ++m1) {
                if (std::fabs(V[m0][0] - V[m1][0]) > thx || std::fabs(V[m0][1] - V[m1][1]) > thy) continue;
                cand.push_back({m0, m1});
                v0s.insert(v0s.end(), V[m0].begin(), V[m0].begin() + 3);
                v1s.insert(v1s.end(), V[m1].begin(), V[m1].begin() + 3);
            }
        stage_.add("  bridge candidates + FK (host)", seconds(tT));
        std::vector<unsigned char> ok;
        if (!cand.empty()) {
            StageScope ss(stage_, "  K5b testTransitions");
            ok = freeSpace_.testTransitions(std::vector<int>(cand.size(), 0), v0s, v1s, w0, w1, linkOff, &chainOff);
        }
        // replay with the reference's moving lower bound (candidates are listed by m0, ascending in m1)
        Index n12 = new0;
        size_t j = 0;
        for (Index m0 = start; m0 < n1; ++m0) {
            bool done = false;
            for (; j < cand.size() && cand[j].first == m0; ++j) {
                if (done || cand[j].second < n12 || !ok[j]) continue;
                addEdge(m0, cand[j].second);
                n12 = cand[j].second;
                done = true;  // (break in the reference: the first free partner at or above the bound)
            }
        }
    }

    std::vector<unsigned char> batchCode_;  // connectOneSlice pair codes of the look-ahead batch (device matching)
    std::vector<StagedSlice> staged_;       // the look-ahead batch built ahead of its turn (empty: slice by slice)
    bool batchBridge_ = std::getenv("HRM_PROB_PER_SLICE") == nullptr;  // (A/B switch: one bridge build + one matching call per slice)
    std::vector<SlicePairs> batchSp_;
    std::string urdfFile_;
    ParseURDF kdl_;
    MultiBodyTree3D robot0_ = robot_;  // the as-loaded robot
    std::vector<std::vector<Coordinate>> v_;  // robot shapes, one per slice
    std::vector<std::vector<double>> samples_;
    bool injected_ = false;
    std::mt19937_64 rng_{20220726ULL};
    size_t batch_ = 8;
    size_t refinePeriod_ = 60;
    size_t maxSlices_ = 4096;  // hard stop for the seeded-RNG mode without a time limit
    double maxJointAngle_ = PI / 2;
};

}  // namespace planners
}  // namespace hrm
