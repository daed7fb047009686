// TEST INFRASTRUCTURE ONLY (see oracle/README.md): CPU restatement of the reference's articulated path --
//   src/util/ParseURDF.cpp:10-33 + kdl_parser + KDL::TreeFkSolverPos_recursive   URDF -> revolute tree -> forward kinematics
//   src/datastructure/MultiBodyTree3D.cpp:55-89                                    robotTF(urdf, gBase, jointConfig)
//   src/planners/ProbHRM3D.cpp:7-246                                               ProbHRM3D
// orocos-KDL, urdfdom and tinyxml2 are third-party dependencies absent from /root/reference and from this image
// (package.xml: orocos_kdl, kdl_parser vendored under src/util/external without KDL itself).  Their published
// algorithms are restated: a URDF joint contributes Frame(Rot(axis, q), xyz) * Frame(RPY(rpy), 0) (kdl_parser's
// toKdl(joint): Joint(origin, M*axis, RotAxis) with f_tip = pose(0)^-1 * F_parent_jnt), KDL numbers the joints in the
// depth-first order in which kdl_parser adds the segments (children ordered by urdfdom's name-sorted joint map), and
// the recursive solver multiplies the segment poses from the root down.  Parity of this file with KDL is pinned by
// analytic identities only (tests/test_oracle_kat.py::test_fk_*): "parity unpinned" at the last bit.
#pragma once
#include "hrmo_planner.hpp"

#include <algorithm>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>

namespace hrmo {

struct UrdfJoint {
    std::string name, type, parent, child;
    double xyz[3] = {0, 0, 0}, rpy[3] = {0, 0, 0}, axis[3] = {1, 0, 0};
};

// ---- a deliberately small URDF reader: <joint name type> with <parent link>, <child link>, <origin xyz rpy>, <axis xyz>
inline std::string xmlAttr(const std::string& tag, const std::string& key) {
    const std::string k = key + "=\"";
    size_t p = tag.find(" " + k);
    if (p == std::string::npos) return "";
    p += k.size() + 1;
    const size_t e = tag.find('"', p);
    return tag.substr(p, e - p);
}
inline void parse3(const std::string& s, double out[3]) {
    if (s.empty()) return;
    std::istringstream in(s);
    in >> out[0] >> out[1] >> out[2];
}
inline std::vector<UrdfJoint> parseUrdfJoints(const std::string& file) {
    std::ifstream in(file);
    if (!in) throw std::invalid_argument("cannot open URDF " + file);
    std::stringstream buf;
    buf << in.rdbuf();
    std::string x = buf.str();
    for (size_t a; (a = x.find("<!--")) != std::string::npos;) {  // drop comments
        const size_t b = x.find("-->", a);
        x.erase(a, (b == std::string::npos ? x.size() : b + 3) - a);
    }
    std::vector<UrdfJoint> joints;
    for (size_t p = 0; (p = x.find("<joint ", p)) != std::string::npos;) {
        const size_t headEnd = x.find('>', p);
        const size_t end = x.find("</joint>", p);
        if (end == std::string::npos) break;
        const std::string head = x.substr(p, headEnd - p + 1), body = x.substr(headEnd + 1, end - headEnd - 1);
        UrdfJoint j;
        j.name = xmlAttr(head, "name");
        j.type = xmlAttr(head, "type");
        auto tagOf = [&](const std::string& name) {
            const size_t a = body.find("<" + name);
            if (a == std::string::npos) return std::string();
            return body.substr(a, body.find('>', a) - a + 1);
        };
        j.parent = xmlAttr(tagOf("parent"), "link");
        j.child = xmlAttr(tagOf("child"), "link");
        parse3(xmlAttr(tagOf("origin"), "xyz"), j.xyz);
        parse3(xmlAttr(tagOf("origin"), "rpy"), j.rpy);
        parse3(xmlAttr(tagOf("axis"), "xyz"), j.axis);
        joints.push_back(j);
        p = end;
    }
    return joints;
}

// KDL::Rotation::Rot2(unit axis, angle) -- Rodrigues' formula in KDL's arrangement
inline Mat3 rotAxisAngle(const double v[3], double angle) {
    const double ct = std::cos(angle), st = std::sin(angle), vt = 1.0 - ct;
    const double mvt0 = vt * v[0], mvt1 = vt * v[1], mvt2 = vt * v[2];
    const double mst0 = v[0] * st, mst1 = v[1] * st, mst2 = v[2] * st;
    const double mvt01 = mvt0 * v[1], mvt02 = mvt0 * v[2], mvt12 = mvt1 * v[2];
    Mat3 r;
    r.m[0][0] = ct + mvt0 * v[0], r.m[0][1] = -mst2 + mvt01, r.m[0][2] = mst1 + mvt02;
    r.m[1][0] = mst2 + mvt01, r.m[1][1] = ct + mvt1 * v[1], r.m[1][2] = -mst0 + mvt12;
    r.m[2][0] = -mst1 + mvt02, r.m[2][1] = mst0 + mvt12, r.m[2][2] = ct + mvt2 * v[2];
    return r;
}
// KDL::Rotation::RPY(roll, pitch, yaw) = Rz(yaw) * Ry(pitch) * Rx(roll)
inline Mat3 rotRPY(double roll, double pitch, double yaw) {
    const double ca1 = std::cos(yaw), sa1 = std::sin(yaw), cb1 = std::cos(pitch), sb1 = std::sin(pitch), cc1 = std::cos(roll),
                 sc1 = std::sin(roll);
    Mat3 r;
    r.m[0][0] = ca1 * cb1, r.m[0][1] = ca1 * sb1 * sc1 - sa1 * cc1, r.m[0][2] = ca1 * sb1 * cc1 + sa1 * sc1;
    r.m[1][0] = sa1 * cb1, r.m[1][1] = sa1 * sb1 * sc1 + ca1 * cc1, r.m[1][2] = sa1 * sb1 * cc1 - ca1 * sc1;
    r.m[2][0] = -sb1, r.m[2][1] = cb1 * sc1, r.m[2][2] = cb1 * cc1;
    return r;
}

struct KinematicTree {
    struct Segment {
        std::string name, parent;  // child link, parent link
        Mat3 originR;              // RPY of the joint origin
        double origin[3], axis[3];  // joint position and unit axis, both in the parent frame
        int q = -1;                // joint index (revolute / continuous), -1 for fixed
    };
    std::vector<Segment> seg;               // in KDL's segment order
    std::map<std::string, size_t> byName;   // child link -> segment
    std::string root;
    size_t nJoints = 0;

    explicit KinematicTree(const std::string& urdfFile) {
        const auto joints = parseUrdfJoints(urdfFile);
        std::map<std::string, std::vector<const UrdfJoint*>> children;  // parent link -> joints, name-sorted like urdfdom's map
        std::map<std::string, bool> isChild;
        for (const auto& j : joints) children[j.parent].push_back(&j), isChild[j.child] = true;
        for (auto& c : children)
            std::sort(c.second.begin(), c.second.end(), [](const UrdfJoint* a, const UrdfJoint* b) { return a->name < b->name; });
        for (const auto& j : joints)
            if (!isChild.count(j.parent)) root = j.parent;
        addChildren(root, children);
    }
    size_t getNrOfJoints() const { return nJoints; }

    // ParseURDF::getTransform: pose of link `bodyName` in the root frame; the rotation goes through a quaternion like the
    // reference (KDL GetQuaternion -> Eigen toRotationMatrix)
    Tf3 getTransform(const std::vector<double>& jointConfig, const std::string& bodyName) const {
        Tf3 f = poseOf(jointConfig, bodyName);
        f.R = quat_to_rot(rot_to_quat(f.R));
        return f;
    }

  private:
    void addChildren(const std::string& link, const std::map<std::string, std::vector<const UrdfJoint*>>& children) {
        const auto it = children.find(link);
        if (it == children.end()) return;
        for (const UrdfJoint* j : it->second) {
            Segment s;
            s.name = j->child, s.parent = j->parent;
            s.originR = rotRPY(j->rpy[0], j->rpy[1], j->rpy[2]);
            for (int d = 0; d < 3; ++d) s.origin[d] = j->xyz[d];
            const Vec3 a = s.originR * Vec3{j->axis[0], j->axis[1], j->axis[2]};
            const double nrm = std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z);
            s.axis[0] = a.x / nrm, s.axis[1] = a.y / nrm, s.axis[2] = a.z / nrm;
            if (j->type == "revolute" || j->type == "continuous") s.q = static_cast<int>(nJoints++);
            byName[s.name] = seg.size();
            seg.push_back(s);
            addChildren(j->child, children);
        }
    }
    Tf3 poseOf(const std::vector<double>& q, const std::string& link) const {
        Tf3 id;
        id.R = mat3_identity();
        id.t[0] = id.t[1] = id.t[2] = 0.0;
        if (link == root) return id;
        const auto it = byName.find(link);
        if (it == byName.end()) throw std::invalid_argument("no link named " + link);
        const Segment& s = seg[it->second];
        Tf3 cur;  // Segment::pose(q) = Frame(Rot2(axis, q), origin) * Frame(originR, 0)
        cur.R = (s.q >= 0 ? rotAxisAngle(s.axis, q.at(static_cast<size_t>(s.q))) : mat3_identity()) * s.originR;
        for (int d = 0; d < 3; ++d) cur.t[d] = s.origin[d];
        if (s.parent == root) return cur;
        return tf3_mul(poseOf(q, s.parent), cur);
    }
};

// MultiBodyTree3D::robotTF(kdl, gBase, jointConfig)                                MultiBodyTree3D.cpp:63-89
inline void robotTF(MultiBodyTree3D& robot, const KinematicTree& kdl, const Tf3& gBase, const std::vector<double>& jointConfig) {
    for (int i = 0; i < 3; ++i) robot.base.position[i] = gBase.t[i];
    robot.base.quat = rot_to_quat(gBase.R);
    for (size_t i = 0; i < robot.link.size(); ++i) {
        const Tf3 gl = tf3_mul(tf3_mul(gBase, kdl.getTransform(jointConfig, "body" + std::to_string(i + 1))), robot.tf.at(i));
        for (int d = 0; d < 3; ++d) robot.link[i].position[d] = gl.t[d];
        robot.link[i].quat = rot_to_quat(gl.R);
    }
}

// ---- ProbHRM3D (src/planners/ProbHRM3D.cpp) ---------------------------------------------------------------------------
// The reference draws the base rotation (Quaterniond::UnitRandom) and the joint angles (ompl::RNG) of every slice after
// the first two at run time (SURVEY.md finding 9); here the draws are an injected list, one row [qw,qx,qy,qz, joints...] per
// slice >= 2, and plan() stops when the path is found or the list is exhausted.  The sweep-line refinement every 60 slices
// (ProbHRM3D.cpp:52-56) is restated with the period as a parameter (60 in the reference; tests use small periods), on top of
// the base class's refineExistRoadmap ("done right", SURVEY.md §8 f.4: every slice re-swept on its own cached boundary).
struct ProbHRM3D : HRM3D {
    KinematicTree kdl_;
    std::vector<std::vector<double>> samples_;

    ProbHRM3D(const MultiBodyTree3D& robot, const std::string& urdfFile, const std::vector<SuperQuadrics>& arena,
              const std::vector<SuperQuadrics>& obs, const PlanningRequest& req, const std::vector<std::vector<double>>& samples,
              bool perOrientationFreeSpace = false)
        : HRM3D(robot, arena, obs, req, {}, perOrientationFreeSpace), kdl_(urdfFile), samples_(samples) {
        isRobotRigid_ = false;
    }

    // :18-65
    void planProb(size_t refinePeriod = 60) {
        param_.numSlice = 0;
        do {
            if (param_.numSlice >= 2 && param_.numSlice - 2 >= samples_.size()) break;  // injected draws exhausted
            sampleOrientations();
            constructOneSlice(param_.numSlice);
            if (!isRefine_) param_.numSlice++;
            numVertex_.slice = res_.graphStructure.vertex.size();
            vertexIdx_.push_back(numVertex_);
            if (param_.numSlice >= 2) connectMultiSlice();
            search();
            // :52-56 "Double the number of sweep lines for every 10 iterations" (the code says 60), at most numPoint times
            if (refinePeriod > 0 && param_.numSlice % refinePeriod == 0 && vertexIdxAll_.size() < param_.numPoint) refineExistRoadmap();
        } while (!res_.solved);
        if (res_.solved) {
            const size_t poseSize = res_.graphStructure.vertex.at(0).size();
            start_.resize(poseSize);
            goal_.resize(poseSize);
            res_.solvedPath.push_back(start_);
            for (int id : res_.PathId) res_.solvedPath.push_back(res_.graphStructure.vertex.at(static_cast<size_t>(id)));
            res_.solvedPath.push_back(goal_);
        }
    }

    // :67-102
    void sampleOrientations() override {
        const size_t nj = kdl_.getNrOfJoints();
        std::vector<double> config{0.0, 0.0, 0.0};
        if (param_.numSlice == 0) {
            q_.push_back(Quat{start_.at(3), start_.at(4), start_.at(5), start_.at(6)});
        } else if (param_.numSlice == 1) {
            q_.push_back(Quat{goal_.at(3), goal_.at(4), goal_.at(5), goal_.at(6)});
        } else {
            const auto& s = samples_.at(param_.numSlice - 2);
            q_.push_back(Quat{s.at(0), s.at(1), s.at(2), s.at(3)});
        }
        config.push_back(q_.back().w), config.push_back(q_.back().x), config.push_back(q_.back().y), config.push_back(q_.back().z);
        for (size_t i = 0; i < nj; ++i) {
            if (param_.numSlice == 0) config.push_back(start_.at(7 + i));
            else if (param_.numSlice == 1) config.push_back(goal_.at(7 + i));
            else config.push_back(samples_.at(param_.numSlice - 2).at(4 + i));
        }
        v_.push_back(config);
    }

    // :198-213
    void setTransform(const std::vector<double>& V) override {
        Tf3 g;
        g.R = quat_to_rot(Quat{V[3], V[4], V[5], V[6]});
        g.t[0] = V[0], g.t[1] = V[1], g.t[2] = V[2];
        std::vector<double> jointConfig(kdl_.getNrOfJoints());
        for (size_t i = 0; i < jointConfig.size(); ++i) jointConfig[i] = V[7 + i];
        robotTF(robot_, kdl_, g, jointConfig);
    }

    // :165-195
    void generateVertices(double tx, const FreeSegment2D& freeSeg) override {
        numVertex_.plane.clear();
        std::vector<double> vertex(v_.back().size());
        for (size_t i = 0; i < freeSeg.ty.size(); ++i) {
            numVertex_.plane.push_back(res_.graphStructure.vertex.size());
            for (size_t j = 0; j < freeSeg.xM[i].size(); ++j) {
                vertex[0] = tx, vertex[1] = freeSeg.ty[i], vertex[2] = freeSeg.xM[i][j];
                vertex[3] = robot_.base.quat.w, vertex[4] = robot_.base.quat.x, vertex[5] = robot_.base.quat.y, vertex[6] = robot_.base.quat.z;
                for (size_t m = 7; m < v_.back().size(); ++m) vertex[m] = v_.back()[m];
                res_.graphStructure.vertex.push_back(vertex);
            }
        }
        numVertex_.line.push_back(numVertex_.plane);
    }

    // :216-246
    void computeTFEArticulated(const std::vector<double>& v1, const std::vector<double>& v2, std::vector<Ellipsoid3>& tfe) {
        tfe.clear();
        const auto vInterp = interpolateCompoundSE3Rn(v1, v2, param_.numPoint);
        setTransform(v1);
        const size_t B = robot_.getNumLinks() + 1;
        auto body = [&](size_t i) -> const SuperQuadrics& { return i == 0 ? robot_.base : robot_.link[i - 1]; };
        std::vector<Ellipsoid3> mvce(B);
        for (size_t i = 0; i < B; ++i) {
            for (int d = 0; d < 3; ++d) mvce[i].semi[d] = body(i).semiAxis[d];
            mvce[i].quat = body(i).quat;
        }
        for (const auto& vStep : vInterp) {
            setTransform(vStep);
            for (size_t i = 0; i < B; ++i) mvce[i] = getMVCE3D(mvce[i].semi, body(i).semiAxis, mvce[i].quat, body(i).quat);
        }
        tfe = mvce;
    }

    // :105-162
    void connectMultiSlice() override {
        if (param_.numSlice == 1) return;
        double minDist = INFINITY;
        size_t minIdx = 0;
        for (size_t i = 0; i + 1 < param_.numSlice; ++i) {
            const double dist = vectorEuclidean(v_.back(), v_.at(i));
            if (dist < minDist) minDist = dist, minIdx = i;
        }
        size_t n12 = vertexIdx_.at(param_.numSlice - 1).startId;
        const size_t n2 = vertexIdx_.at(param_.numSlice - 1).slice;
        const size_t start = vertexIdx_.at(minIdx).startId;
        const size_t n1 = vertexIdx_.at(minIdx).slice;
        computeTFEArticulated(v_.back(), v_.at(minIdx), tfe_);
        bridgeSlice();
        for (size_t m0 = start; m0 < n1; ++m0) {
            const auto v1 = res_.graphStructure.vertex.at(m0);
            for (size_t m1 = n12; m1 < n2; ++m1) {
                const auto v2 = res_.graphStructure.vertex.at(m1);
                if (std::fabs(v1.at(0) - v2.at(0)) >
                        2.0 * (param_.boundaryLimits[1] - param_.boundaryLimits[0]) / static_cast<double>(param_.numLineX) ||
                    std::fabs(v1.at(1) - v2.at(1)) >
                        2.0 * (param_.boundaryLimits[3] - param_.boundaryLimits[2]) / static_cast<double>(param_.numLineY))
                    continue;
                if (isMultiSliceTransitionFree(v1, v2)) {
                    res_.graphStructure.edge.push_back({m0, m1});
                    res_.graphStructure.weight.push_back(vectorEuclidean(v1, v2));
                    n12 = m1;
                    break;
                }
            }
        }
    }
};

}  // namespace hrmo
