// Host-side C++ mirror of the reference's planner API for the accelerated path (same class and method names as
// ChirikjianLab/hrm where Eigen-free signatures allow it):
//
//   hrm::SuperQuadrics / SuperEllipse          include/hrm/geometry/SuperQuadrics.h:23-74, SuperEllipse.h:23-65  (parameters only)
//   hrm::MultiBodyTree3D / MultiBodyTree2D     include/hrm/datastructure/MultiBodyTree.h:12-63
//   hrm::FreeSpaceGPU3D / FreeSpaceGPU2D       drop-in for FreeSpace3D / FreeSpace2D (FreeSpace.h:45-128, FreeSpace3D.h:29-55),
//                                              every compute* call forwarded to the C ABI of include/hrmgpu.h (batched over layers)
//   hrm::getMVCE3D/2D, getTFE3D/2D, interpolate*   src/geometry/TightFitEllipsoid.cpp, src/util/InterpolateSE3.cpp  (host, tiny)
//   hrm::planners::HighwayRoadMap / HRM3D / HRM2D, PlanningRequest, PlanningResult   include/hrm/planners/*.h
//
// The GPU does K1-K5; the host enumerates candidates, replays the reference's loops with the returned booleans (so vertex
// ids and edge order are the reference's) and runs A*.  There is no CPU implementation of the kernels here: without a
// device every FreeSpaceGPU constructor throws.
#pragma once
#include "../../include/hrmgpu.h"
#include "hrm_math.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <functional>
#include <map>
#include <queue>
#include <stdexcept>
#include <string>
#include <utility>

namespace hrm {

// ---------------------------------------------------------------------------------------------------------------------
// geometric models (parameter holders; the boundary arithmetic lives in the CUDA kernels)
// ---------------------------------------------------------------------------------------------------------------------
class SuperQuadrics {
  public:
    SuperQuadrics() = default;
    SuperQuadrics(std::vector<double> semiAxis, std::vector<double> epsilon, std::vector<double> position, const Quaternion& quat,
                  Index num)
        : semiAxis_(std::move(semiAxis)), epsilon_(std::move(epsilon)), position_(std::move(position)), quat_(quat), num_(num) {}
    const std::vector<double>& getSemiAxis() const { return semiAxis_; }
    const std::vector<double>& getEpsilon() const { return epsilon_; }
    const std::vector<double>& getPosition() const { return position_; }
    const Quaternion& getQuaternion() const { return quat_; }
    Index getNumParam() const { return num_; }
    Index getNum() const { return num_ * num_; }
    void setPosition(const std::vector<double>& p) { position_ = p; }
    void setQuaternion(const Quaternion& q) { quat_ = q; }
    void setQuatSamples(const std::vector<Quaternion>& q) { qSample_ = q; }
    const std::vector<Quaternion>& getQuatSamples() const { return qSample_; }

  private:
    std::vector<double> semiAxis_{1, 1, 1}, epsilon_{1, 1}, position_{0, 0, 0};
    Quaternion quat_{1, 0, 0, 0};
    Index num_ = 0;
    std::vector<Quaternion> qSample_;
};

class SuperEllipse {
  public:
    SuperEllipse() = default;
    SuperEllipse(std::vector<double> semiAxis, double epsilon, std::vector<double> position, double angle, Index num)
        : semiAxis_(std::move(semiAxis)), position_(std::move(position)), epsilon_(epsilon), angle_(angle), num_(num) {}
    const std::vector<double>& getSemiAxis() const { return semiAxis_; }
    double getEpsilon() const { return epsilon_; }
    const std::vector<double>& getPosition() const { return position_; }
    double getAngle() const { return angle_; }
    Index getNum() const { return num_; }
    void setPosition(const std::vector<double>& p) { position_ = p; }
    void setAngle(double a) { angle_ = a; }

  private:
    std::vector<double> semiAxis_{1, 1}, position_{0, 0};
    double epsilon_ = 1, angle_ = 0;
    Index num_ = 0;
};

struct SE3Transform {
    Mat3 R{{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}};
    Vec3 t{0, 0, 0};
};

// reference src/datastructure/MultiBodyTree3D.cpp
class MultiBodyTree3D {
  public:
    explicit MultiBodyTree3D(const SuperQuadrics& base) : base_(base) {}
    void addBody(const SuperQuadrics& link) {  // :21-32
        link_.push_back(link);
        SE3Transform g;
        g.R = toRotationMatrix(link.getQuaternion());
        g.t = {link.getPosition()[0], link.getPosition()[1], link.getPosition()[2]};
        tf_.push_back(g);
    }
    void robotTF(const SE3Transform& g) {  // :34-53
        base_.setPosition({g.t[0], g.t[1], g.t[2]});
        base_.setQuaternion(toQuaternion(g.R));
        for (size_t i = 0; i < link_.size(); ++i) {
            const Mat3 R = matMul(g.R, tf_[i].R);
            const Vec3 v = matVec(g.R, tf_[i].t);
            link_[i].setPosition({v[0] + g.t[0], v[1] + g.t[1], v[2] + g.t[2]});
            link_[i].setQuaternion(toQuaternion(R));
        }
    }
    /** articulated body (:55-89): link i sits at gBase * FK("body<i+1>") * tf_i; Kdl = hrm::ParseURDF (hrm_articulated.hpp) */
    template <class Kdl>
    void robotTF(const Kdl& kdl, const SE3Transform& gBase, const std::vector<double>& jointConfig) {
        base_.setPosition({gBase.t[0], gBase.t[1], gBase.t[2]});
        base_.setQuaternion(toQuaternion(gBase.R));
        for (size_t i = 0; i < link_.size(); ++i) {
            const SE3Transform fk = kdl.getTransform(jointConfig, "body" + std::to_string(i + 1));
            SE3Transform a;  // gBase * fk
            a.R = matMul(gBase.R, fk.R);
            Vec3 v = matVec(gBase.R, fk.t);
            a.t = {v[0] + gBase.t[0], v[1] + gBase.t[1], v[2] + gBase.t[2]};
            const Mat3 R = matMul(a.R, tf_[i].R);
            v = matVec(a.R, tf_[i].t);
            link_[i].setPosition({v[0] + a.t[0], v[1] + a.t[1], v[2] + a.t[2]});
            link_[i].setQuaternion(toQuaternion(R));
        }
    }
    const SuperQuadrics& getBase() const { return base_; }
    const std::vector<SuperQuadrics>& getLinks() const { return link_; }
    const std::vector<SE3Transform>& getTF() const { return tf_; }
    Index getNumLinks() const { return link_.size(); }
    std::vector<SuperQuadrics> getBodyShapes() const {
        std::vector<SuperQuadrics> b{base_};
        b.insert(b.end(), link_.begin(), link_.end());
        return b;
    }
    // flat arrays of the current pose as MultiBodyTree3D::minkSum consumes it (:91-105): rotation of every body, link
    // centres subtracted from the link boundaries
    void bodyArrays(std::vector<double>& semi, std::vector<double>& R, std::vector<double>& off) const {
        const auto bodies = getBodyShapes();
        for (size_t b = 0; b < bodies.size(); ++b) {
            for (double v : bodies[b].getSemiAxis()) semi.push_back(v);
            const Mat3 Rb = toRotationMatrix(bodies[b].getQuaternion());
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) R.push_back(Rb[i][j]);
            for (int d = 0; d < 3; ++d) off.push_back(b == 0 ? 0.0 : bodies[b].getPosition()[d]);
        }
    }

  private:
    SuperQuadrics base_;
    std::vector<SuperQuadrics> link_;
    std::vector<SE3Transform> tf_;
};

struct SE2Transform {
    double R[2][2] = {{1, 0}, {0, 1}};
    double t[2] = {0, 0};
};

// reference src/datastructure/MultiBodyTree2D.cpp
class MultiBodyTree2D {
  public:
    explicit MultiBodyTree2D(const SuperEllipse& base) : base_(base) {}
    void addBody(const SuperEllipse& link) {  // :10-22
        link_.push_back(link);
        SE2Transform g;
        const double c = std::cos(link.getAngle()), s = std::sin(link.getAngle());
        g.R[0][0] = c, g.R[0][1] = -s, g.R[1][0] = s, g.R[1][1] = c;
        g.t[0] = link.getPosition()[0], g.t[1] = link.getPosition()[1];
        tf_.push_back(g);
    }
    void robotTF(double theta, double x, double y) {  // :24-41 with g = [Rot(theta) (x,y); 0 1]
        const double c = std::cos(theta), s = std::sin(theta);
        const double R[2][2] = {{c, -s}, {s, c}};
        base_.setPosition({x, y});
        base_.setAngle(std::atan2(R[1][0], R[0][0]));
        for (size_t i = 0; i < link_.size(); ++i) {
            const SE2Transform& l = tf_[i];
            const double m00 = R[0][0] * l.R[0][0] + R[0][1] * l.R[1][0];
            const double m10 = R[1][0] * l.R[0][0] + R[1][1] * l.R[1][0];
            link_[i].setPosition({(R[0][0] * l.t[0] + R[0][1] * l.t[1]) + x, (R[1][0] * l.t[0] + R[1][1] * l.t[1]) + y});
            link_[i].setAngle(std::atan2(m10, m00));
        }
    }
    const SuperEllipse& getBase() const { return base_; }
    const std::vector<SuperEllipse>& getLinks() const { return link_; }
    const std::vector<SE2Transform>& getTF() const { return tf_; }
    Index getNumLinks() const { return link_.size(); }
    // (semi, angle, offset) of every body as MultiBodyTree2D::minkSum derives them from the base angle (:43-65)
    void bodyArrays(std::vector<double>& semi, std::vector<double>& angle, std::vector<double>& off) const {
        for (double v : base_.getSemiAxis()) semi.push_back(v);
        angle.push_back(base_.getAngle());
        off.push_back(0.0), off.push_back(0.0);
        const double c = std::cos(base_.getAngle()), s = std::sin(base_.getAngle());
        for (size_t i = 0; i < link_.size(); ++i) {
            for (double v : link_[i].getSemiAxis()) semi.push_back(v);
            const SE2Transform& l = tf_[i];
            const double m00 = c * l.R[0][0] + (-s) * l.R[1][0];
            const double m10 = s * l.R[0][0] + c * l.R[1][0];
            angle.push_back(std::atan2(m10, m00));
            off.push_back(c * l.t[0] + (-s) * l.t[1]);
            off.push_back(s * l.t[0] + c * l.t[1]);
        }
    }

  private:
    SuperEllipse base_;
    std::vector<SuperEllipse> link_;
    std::vector<SE2Transform> tf_;
};

// ---------------------------------------------------------------------------------------------------------------------
// interpolation + tightly-fitted ellipsoids (host)
// ---------------------------------------------------------------------------------------------------------------------
inline std::vector<Quaternion> interpolateSlerp(const Quaternion& a, const Quaternion& b, Index numStep) {  // InterpolateSE3.cpp:89-103
    std::vector<Quaternion> out;
    const double dt = 1.0 / static_cast<double>(numStep);
    for (size_t i = 0; i <= numStep; ++i) out.push_back(slerp(a, static_cast<double>(i) * dt, b));
    return out;
}
inline std::vector<std::vector<Coordinate>> interpolateSE3(const std::vector<Coordinate>& v0, const std::vector<Coordinate>& v1,
                                                           Index numStep) {  // :5-29
    const auto q = interpolateSlerp({v0[3], v0[4], v0[5], v0[6]}, {v1[3], v1[4], v1[5], v1[6]}, numStep);
    const double dt = 1.0 / static_cast<double>(numStep);
    std::vector<std::vector<Coordinate>> out;
    for (size_t i = 0; i <= numStep; ++i) {
        std::vector<Coordinate> s;
        for (int j = 0; j < 3; ++j) s.push_back((1.0 - static_cast<double>(i) * dt) * v0[j] + static_cast<double>(i) * dt * v1[j]);
        s.insert(s.end(), q[i].begin(), q[i].end());
        out.push_back(s);
    }
    return out;
}

inline std::vector<std::vector<Coordinate>> interpolateRn(const std::vector<Coordinate>& v0, const std::vector<Coordinate>& v1, Index numStep) {  // :105-122
    const double dt = 1.0 / static_cast<double>(numStep);
    std::vector<std::vector<Coordinate>> out;
    for (size_t i = 0; i <= numStep; ++i) {
        std::vector<Coordinate> s;
        for (size_t j = 0; j < v0.size(); ++j) s.push_back((1.0 - static_cast<double>(i) * dt) * v0[j] + static_cast<double>(i) * dt * v1[j]);
        out.push_back(s);
    }
    return out;
}
inline std::vector<std::vector<Coordinate>> interpolateCompoundSE3Rn(const std::vector<Coordinate>& v0, const std::vector<Coordinate>& v1,
                                                                     Index numStep) {  // :31-65
    if (v0.size() == 7) return interpolateSE3(v0, v1, numStep);
    auto out = interpolateSE3(v0, v1, numStep);  // reads the first 7 entries
    const auto rn = interpolateRn({v0.begin() + 7, v0.end()}, {v1.begin() + 7, v1.end()}, numStep);
    for (size_t i = 0; i <= numStep; ++i) out[i].insert(out[i].end(), rn[i].begin(), rn[i].end());
    return out;
}

struct Ellipsoid {
    Vec3 semi;
    Quaternion quat;
};
inline Ellipsoid getMVCE3D(const Vec3& a, const Vec3& b, const Quaternion& quatA, const Quaternion& quatB) {  // TightFitEllipsoid.cpp:43-81
    const Mat3 Ra = toRotationMatrix(quatA), Rb = toRotationMatrix(quatB);
    const double r = std::fmin(b[0], std::fmin(b[1], b[2]));
    const Vec3 dg = {r / b[0], r / b[1], r / b[2]};
    const Vec3 dA = {std::pow(a[0], -2.0), std::pow(a[1], -2.0), std::pow(a[2], -2.0)};
    const Mat3 T = matMul(mulDiag(Rb, dg), transpose(Rb));
    const Mat3 Ti = inverse(T);
    const Mat3 aP = matMul(matMul(Ti, matMul(mulDiag(Ra, dA), transpose(Ra))), Ti);
    Vec3 ev;
    Mat3 U;
    symEig3(aP, ev, U);
    Vec3 dC;
    for (int i = 0; i < 3; ++i) dC[i] = std::pow(std::fmax(std::pow(ev[i], -0.5), r), -2.0);
    const Mat3 C = matMul(matMul(mulDiag(matMul(T, U), dC), transpose(U)), T);
    symEig3(C, ev, U);
    Ellipsoid e;
    e.quat = toQuaternion(U);
    for (int i = 0; i < 3; ++i) e.semi[i] = std::pow(ev[i], -0.5);
    return e;
}
inline Ellipsoid getTFE3D(const Vec3& a, const Quaternion& quatA, const Quaternion& quatB, Index numStep) {  // :98-114
    const auto qi = interpolateSlerp(quatA, quatB, numStep);
    Ellipsoid e = getMVCE3D(a, a, quatA, quatB);
    for (size_t i = 1; i < numStep; ++i) e = getMVCE3D(a, e.semi, qi.at(i), e.quat);
    return e;
}

struct Ellipse {
    double semi[2];
    double angle;
};
inline void symEig2(const double A[2][2], double ev[2], double U[2][2]) {
    const double a = A[0][0], b = 0.5 * (A[0][1] + A[1][0]), d = A[1][1];
    const double tr = a + d, df = a - d, rt = std::sqrt(df * df + 4.0 * b * b);
    ev[0] = 0.5 * (tr + rt), ev[1] = 0.5 * (tr - rt);
    double vx, vy;
    if (std::fabs(b) > 0.0) {
        vx = ev[0] - d, vy = b;
        if (std::fabs(vx) + std::fabs(vy) == 0.0) vx = b, vy = ev[0] - a;
    } else if (a >= d) {
        vx = 1.0, vy = 0.0;
    } else {
        vx = 0.0, vy = 1.0;
    }
    const double n = std::sqrt(vx * vx + vy * vy);
    vx /= n, vy /= n;
    U[0][0] = vx, U[1][0] = vy, U[0][1] = -vy, U[1][1] = vx;
}
inline Ellipse getMVCE2D(const double a[2], const double b[2], double thetaA, double thetaB) {  // TightFitEllipsoid.cpp:6-41
    const double Ra[2][2] = {{std::cos(thetaA), -std::sin(thetaA)}, {std::sin(thetaA), std::cos(thetaA)}};
    const double Rb[2][2] = {{std::cos(thetaB), -std::sin(thetaB)}, {std::sin(thetaB), std::cos(thetaB)}};
    const double r = std::fmin(b[0], b[1]);
    const double dg[2] = {r / b[0], r / b[1]};
    const double dA[2] = {std::pow(a[0], -2), std::pow(a[1], -2)};
    auto rdrt = [](const double R[2][2], const double d[2], double out[2][2]) {
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) out[i][j] = (R[i][0] * d[0]) * R[j][0] + (R[i][1] * d[1]) * R[j][1];
    };
    auto mul = [](const double X[2][2], const double Y[2][2], double out[2][2]) {
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) out[i][j] = X[i][0] * Y[0][j] + X[i][1] * Y[1][j];
    };
    double T[2][2], Am[2][2], tmp[2][2], aP[2][2];
    rdrt(Rb, dg, T);
    rdrt(Ra, dA, Am);
    const double idt = 1.0 / (T[0][0] * T[1][1] - T[0][1] * T[1][0]);
    const double Ti[2][2] = {{T[1][1] * idt, -T[0][1] * idt}, {-T[1][0] * idt, T[0][0] * idt}};
    mul(Ti, Am, tmp);
    mul(tmp, Ti, aP);
    double ev[2], U[2][2];
    symEig2(aP, ev, U);
    const double cP[2] = {std::max(std::pow(ev[0], -0.5), r), std::max(std::pow(ev[1], -0.5), r)};
    const double dC[2] = {std::pow(cP[0], -2), std::pow(cP[1], -2)};
    const double Ut[2][2] = {{U[0][0], U[1][0]}, {U[0][1], U[1][1]}};
    double TU[2][2], TUD[2][2], C[2][2];
    mul(T, U, TU);
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) TUD[i][j] = TU[i][j] * dC[j];
    mul(TUD, Ut, tmp);
    mul(tmp, T, C);
    symEig2(C, ev, U);
    Ellipse e;
    e.angle = std::acos(std::fmax(-1.0, std::fmin(1.0, U[0][0])));
    e.semi[0] = std::pow(ev[0], -0.5), e.semi[1] = std::pow(ev[1], -0.5);
    return e;
}
inline Ellipse getTFE2D(const double a[2], double thetaA, double thetaB, Index numStep) {  // :83-96
    Ellipse e = getMVCE2D(a, a, thetaA, thetaB);
    const double dt = 1.0 / (static_cast<double>(numStep) - 1);
    for (size_t i = 0; i < numStep; ++i) {
        const double ci = static_cast<double>(i);
        e = getMVCE2D(a, e.semi, (1 - ci * dt) * thetaA + ci * dt * thetaB, e.angle);
    }
    return e;
}

// ---------------------------------------------------------------------------------------------------------------------
// free-space computators backed by the GPU
// ---------------------------------------------------------------------------------------------------------------------
struct FreeSegment2D {  // include/hrm/datastructure/FreeSpace.h:29-41
    std::vector<Coordinate> ty;
    std::vector<std::vector<Coordinate>> xL, xU, xM;
};
struct FreeSegment3D {  // include/hrm/datastructure/FreeSpace3D.h:10-16
    std::vector<Coordinate> tx;
    std::vector<FreeSegment2D> freeSegmentYZ;
};
using BoundaryPoints = std::vector<double>;  // dim x M column-major, like Eigen::MatrixXd
struct BoundaryInfo {
    std::vector<BoundaryPoints> arena, obstacle;
};

class FreeSpaceGPU {
  public:
    /** The reference builds a FreeSpace3D per planner object; here the device context behind it is kept for the next planner
     * object on the same device (hrmgpu_reset: no state survives, streams / events / buffers do) -- creating one costs several
     * milliseconds, a whole plan() on the bundled maps about as much.  HRM_NO_CTX_CACHE=1 turns the cache off. */
    explicit FreeSpaceGPU(int device) : device_(device) {
        {
            std::lock_guard<std::mutex> lk(cache().m);
            auto& spare = cache().spare;
            for (size_t i = 0; i < spare.size(); ++i)
                if (spare[i].first == device) {
                    ctx_ = spare[i].second;
                    spare.erase(spare.begin() + static_cast<long>(i));
                    break;
                }
        }
        if (ctx_) return;
        const int rc = hrmgpu_create(&ctx_, device, 0);
        if (rc != HRMGPU_OK) throw std::runtime_error("hrmgpu_create failed (no sm_100 device? there is no CPU fallback), status " + std::to_string(rc));
    }
    virtual ~FreeSpaceGPU() {
        if (!ctx_) return;
        if (std::getenv("HRM_NO_CTX_CACHE") == nullptr && hrmgpu_reset(ctx_) == HRMGPU_OK) {
            std::lock_guard<std::mutex> lk(cache().m);
            if (cache().spare.size() < 2) {
                cache().spare.push_back({device_, ctx_});
                return;
            }
        }
        hrmgpu_destroy(ctx_);
    }
    FreeSpaceGPU(const FreeSpaceGPU&) = delete;
    FreeSpaceGPU& operator=(const FreeSpaceGPU&) = delete;

    /** setup(numLine, low, up) of the reference + the sweep grids of sweepLineProcess */
    void setup(const std::vector<Coordinate>& boundaryLimits, Index numLineX, Index numLineY) {
        check(hrmgpu_set_sweep(ctx_, boundaryLimits.data(), static_cast<int>(numLineX), static_cast<int>(numLineY)));
        lim_ = boundaryLimits, Lx_ = numLineX, Ly_ = numLineY;
    }
    /** the sweep-line coordinates of HRM3D::sweepLineProcess (HRM3D.cpp:66-81): t_i = lo + double(i) * d */
    void sweepGrid(std::vector<Coordinate>& tx, std::vector<Coordinate>& ty) const {
        tx.assign(Lx_, 0.0), ty.assign(Ly_, 0.0);
        if (lim_.size() >= 6) {
            const double dx = (lim_[1] - lim_[0]) / static_cast<double>(Lx_ - 1), dy = (lim_[3] - lim_[2]) / static_cast<double>(Ly_ - 1);
            for (Index i = 0; i < Lx_; ++i) tx[i] = lim_[0] + static_cast<double>(i) * dx;
            for (Index i = 0; i < Ly_; ++i) ty[i] = lim_[2] + static_cast<double>(i) * dy;
        }
    }
    /** refinement (HighwayRoadMap-inl.h:178-215): sweep + free segments of all built layers redone on their cached boundaries */
    void resweep(const std::vector<Coordinate>& boundaryLimits, Index numLineX, Index numLineY) {
        check(hrmgpu_resweep(ctx_, boundaryLimits.data(), static_cast<int>(numLineX), static_cast<int>(numLineY)));
        lim_ = boundaryLimits, Lx_ = numLineX, Ly_ = numLineY;
    }
    /** getCSpaceBoundary() of layer k */
    BoundaryInfo getCSpaceBoundary(int k) {
        long long d[8];
        check(hrmgpu_get_dims(ctx_, d));
        BoundaryInfo out;
        const size_t per = static_cast<size_t>(d[7] * d[5]);
        std::vector<double> all(per * static_cast<size_t>(d[2]));
        check(hrmgpu_get_boundary_layer(ctx_, k, all.data()));  // one device transpose + one copy for the whole layer
        for (int s = 0; s < d[2]; ++s)
            (s < d[3] ? out.arena : out.obstacle).emplace_back(all.begin() + static_cast<std::ptrdiff_t>(per * s), all.begin() + static_cast<std::ptrdiff_t>(per * (s + 1)));
        return out;
    }
    /** all free segments of the last build: global CSR over (layer, line) */
    void getFreeSegmentsAll(std::vector<long long>& off, std::vector<double>& xL, std::vector<double>& xU, std::vector<double>& xM) {
        long long d[8];
        check(hrmgpu_get_dims(ctx_, d));
        off.resize(static_cast<size_t>(d[0] * d[4]) + 1);
        xL.resize(static_cast<size_t>(d[6])), xU.resize(static_cast<size_t>(d[6])), xM.resize(static_cast<size_t>(d[6]));
        check(hrmgpu_get_segments_all(ctx_, off.data(), xL.data(), xU.data(), xM.data(), d[6]));
    }
    /** isSameSliceTransitionFree for a batch of segments against the C-obstacles of layer k */
    std::vector<unsigned char> testEdges(int k, const std::vector<double>& v1, const std::vector<double>& v2, int dim) {
        std::vector<unsigned char> out(v1.size() / dim);
        if (!out.empty()) check(hrmgpu_test_edges(ctx_, k, static_cast<int>(out.size()), v1.data(), v2.data(), out.data()));
        return out;
    }
    /** isPtInCFree of Q robot poses (body centres [Q][B][dim]) against the bridge layer of pair p */
    std::vector<unsigned char> testBridge(int p, const std::vector<double>& centres, int Q) {
        std::vector<unsigned char> out(static_cast<size_t>(Q));
        if (Q > 0) check(hrmgpu_test_bridge(ctx_, p, Q, centres.data(), out.data()));
        return out;
    }
    /** the same for the candidate edges of several layers in one launch: edge e against layer layer[e] */
    std::vector<unsigned char> testEdgesMulti(const std::vector<int>& layer, const std::vector<double>& v1, const std::vector<double>& v2) {
        std::vector<unsigned char> out(layer.size());
        if (!out.empty()) check(hrmgpu_test_edges_multi(ctx_, static_cast<int>(out.size()), layer.data(), v1.data(), v2.data(), out.data()));
        return out;
    }
    /** the same for poses of several layer pairs in one launch: pose q against the bridge layer of pair pair[q] */
    std::vector<unsigned char> testBridgeMulti(const std::vector<int>& pair, const std::vector<double>& centres) {
        std::vector<unsigned char> out(pair.size());
        if (!out.empty()) check(hrmgpu_test_bridge_multi(ctx_, static_cast<int>(out.size()), pair.data(), centres.data(), out.data()));
        return out;
    }
    /** isMultiSliceTransitionFree for candidate vertex pairs of any layer pairs in one launch, interpolation on the device:
     * v0, v1 [C][dim] vertex positions, w0/w1 [T] interpolation weights, linkOff [P][T][B][dim] rotated link translations */
    std::vector<unsigned char> testTransitions(const std::vector<int>& pair, const std::vector<double>& v0, const std::vector<double>& v1,
                                               const std::vector<double>& w0, const std::vector<double>& w1, const std::vector<double>& linkOff,
                                               const std::vector<double>* chainOff = nullptr) {
        std::vector<unsigned char> out(pair.size());
        if (out.empty()) return out;
        if (!chainOff)
            check(hrmgpu_test_transitions(ctx_, static_cast<int>(out.size()), pair.data(), v0.data(), v1.data(), static_cast<int>(w0.size()), w0.data(),
                                          w1.data(), linkOff.data(), out.data()));
        else  // articulated robot: body centre = linkOff + (chainOff + t)
            check(hrmgpu_test_transitions_articulated(ctx_, static_cast<int>(out.size()), pair.data(), v0.data(), v1.data(),
                                                      static_cast<int>(w0.size()), w0.data(), w1.data(), linkOff.data(), chainOff->data(), out.data()));
        return out;
    }
    /** connectMultiSlice's vertex matching in one call (hrmgpu_connect_pairs3d): vertexXYZ [V][3] roadmap vertex positions,
     * range4 [P][4] = {near begin, near end, cur begin, cur end} vertex id ranges of every pair; match[first(p) + (m0 - near begin)] =
     * id of the vertex m0 connects to, or -1.  Candidates, transition tests and the replay of the reference's moving lower bound run
     * on the device */
    std::vector<int> connectPairs3D(const std::vector<double>& vertexXYZ, const std::vector<int>& range4, double thx, double thy,
                                    const std::vector<double>& w0, const std::vector<double>& w1, const std::vector<double>& linkOff,
                                    long long* numCandidates = nullptr, const std::vector<double>* chainOff = nullptr) {
        size_t numNear = 0;
        for (size_t i = 0; i + 3 < range4.size(); i += 4) numNear += static_cast<size_t>(range4[i + 1] - range4[i]);
        std::vector<int> match(numNear > 0 ? numNear : 1, -1);
        const long long n = hrmgpu_connect_pairs3d(ctx_, static_cast<long long>(vertexXYZ.size() / 3), vertexXYZ.data(), static_cast<int>(range4.size() / 4),
                                                   range4.data(), thx, thy, static_cast<int>(w0.size()), w0.data(), w1.data(), linkOff.data(),
                                                   chainOff ? chainOff->data() : nullptr, match.data(), static_cast<long long>(match.size()), numCandidates);
        if (n < 0) check(static_cast<int>(n));
        match.resize(static_cast<size_t>(n));
        return match;
    }
    long long kernelLaunches() { return hrmgpu_kernel_launches(ctx_, 0); }
    /** {K, B, S, n_arena, L, M, total segments, dim} of the last build (hrmgpu_get_dims) */
    void dims(long long d[8]) { check(hrmgpu_get_dims(ctx_, d)); }
    hrmgpu_ctx* handle() { return ctx_; }
    void check(int rc) {
        if (rc < 0) throw std::runtime_error(std::string("hrmgpu: ") + hrmgpu_last_error(ctx_));
    }

  protected:
    struct CtxCache {
        std::mutex m;
        std::vector<std::pair<int, hrmgpu_ctx*>> spare;
        ~CtxCache() {
            for (auto& s : spare) hrmgpu_destroy(s.second);
        }
    };
    static CtxCache& cache() {
        static CtxCache c;
        return c;
    }
    int device_ = 0;
    hrmgpu_ctx* ctx_ = nullptr;
    std::vector<Coordinate> lim_;
    Index Lx_ = 0, Ly_ = 0;
};

class FreeSpaceGPU3D : public FreeSpaceGPU {
  public:
    FreeSpaceGPU3D(const std::vector<SuperQuadrics>& arena, const std::vector<SuperQuadrics>& obstacle, int device = 0)
        : FreeSpaceGPU(device) {
        std::vector<double> sq;
        auto push = [&](const SuperQuadrics& s) {
            for (double v : s.getSemiAxis()) sq.push_back(v);
            for (double v : s.getEpsilon()) sq.push_back(v);
            for (double v : s.getPosition()) sq.push_back(v);
            const Mat3 R = toRotationMatrix(s.getQuaternion());
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) sq.push_back(R[i][j]);
        };
        for (const auto& a : arena) push(a);
        for (const auto& o : obstacle) push(o);
        check(hrmgpu_set_scene3d(ctx_, static_cast<int>(arena.size()), static_cast<int>(obstacle.size()), sq.data(),
                                 static_cast<int>(arena.at(0).getNumParam())));
    }
    /** computeCSpaceBoundary + computeCSpaceBoundaryMesh + sweepLineProcess (intervals, free segments) for K robot poses */
    void buildLayers(const std::vector<MultiBodyTree3D>& poses, int precision) {
        std::vector<double> semi, R, off, semiK;
        for (size_t k = 0; k < poses.size(); ++k) {
            semiK.clear();
            poses[k].bodyArrays(semiK, R, off);
            if (k == 0) semi = semiK;
        }
        check(hrmgpu_build_layers(ctx_, static_cast<int>(poses.size()), static_cast<int>(semi.size() / 3), semi.data(), R.data(), off.data(),
                                  precision));
    }
    /** bridgeSlice for P layer pairs: tfe[p][b] */
    void buildBridge(const std::vector<std::vector<Ellipsoid>>& tfe, int precision) {
        std::vector<double> semi, R;
        for (const auto& pr : tfe)
            for (const auto& e : pr) {
                for (double v : e.semi) semi.push_back(v);
                const Mat3 Rb = toRotationMatrix(e.quat);
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) R.push_back(Rb[i][j]);
            }
        check(hrmgpu_build_bridge(ctx_, static_cast<int>(tfe.size()), static_cast<int>(tfe.at(0).size()), semi.data(), R.data(), precision));
    }
    /** connectOneSlice's pair codes of ALL layers in one call (hrmgpu_connect_slices3d): code[layerFirst[k] + i] of pair i of layer k
     * in the reference's enumeration order; 0 direct, 1 / 2 through a bridge vertex, 3 none */
    std::vector<unsigned char> connectSlices3D(std::vector<long long>& layerFirst, size_t numPairs, size_t numLayers) {
        std::vector<unsigned char> code(numPairs > 0 ? numPairs : 1);
        layerFirst.assign(numLayers + 1, 0);
        const long long n = hrmgpu_connect_slices3d(ctx_, code.data(), static_cast<long long>(numPairs), layerFirst.data());
        if (n < 0) check(static_cast<int>(n));
        code.resize(static_cast<size_t>(n));
        return code;
    }
    /** computeTFE + bridgeSlice for P layer pairs on the device (hrmgpu_build_bridge_tfe): bodySemi [B][3], quatA / quatB [P][B][4] */
    void buildBridgeTFE(const std::vector<double>& bodySemi, const std::vector<double>& quatA, const std::vector<double>& quatB, Index numStep,
                        int precision) {
        const int B = static_cast<int>(bodySemi.size() / 3);
        check(hrmgpu_build_bridge_tfe(ctx_, static_cast<int>(quatA.size() / (4 * static_cast<size_t>(B))), B, bodySemi.data(), quatA.data(), quatB.data(),
                                      static_cast<int>(numStep), precision, nullptr));
    }
};

struct IntersectionInterval {  // include/hrm/datastructure/FreeSpace.h:14-26: [line][surface] tables of one sweep plane
    std::vector<std::vector<Coordinate>> arenaLow, arenaUpp, obstacleLow, obstacleUpp;
};

/** One C-layer of a FreeSpaceGPU3D build behind the reference's per-layer FreeSpaceComputator / FreeSpace3D method set
 *  (FreeSpace.h:45-128, FreeSpace3D.h:29-55), in the call order of HRM3D::constructOneSlice / sweepLineProcess (HRM3D.cpp:24-54,63-91):
 *      computeCSpaceBoundary(); getCSpaceBoundary(); computeCSpaceBoundaryMesh(b); getCSpaceBoundaryMesh();
 *      for every x-plane: computeIntersectionInterval({tx, ty}); computeFreeSegment(ty); getFreeSegment();
 *  The arithmetic already ran for ALL layers in FreeSpaceGPU3D::buildLayers(); these methods hand out that layer's cached results
 *  (fetched from the device once, on first use), so the reference's planner loop can run on top of it unchanged. */
class FreeSpaceLayer3D {
  public:
    FreeSpaceLayer3D(FreeSpaceGPU3D& fs, int layer) : fs_(fs), k_(layer) { fs_.dims(d_); }
    void computeCSpaceBoundary() {}                                    // (done for all layers by buildLayers)
    const BoundaryInfo& getCSpaceBoundary() {
        if (bound_.arena.empty() && bound_.obstacle.empty()) bound_ = fs_.getCSpaceBoundary(k_);
        return bound_;
    }
    void computeCSpaceBoundaryMesh(const BoundaryInfo&) {}            // (the mesh connectivity is implicit on the device)
    /** tLine = {tx, ty} as the reference passes it: the x-plane is tx.back() (FreeSpace3D.cpp:29-33) */
    void computeIntersectionInterval(const std::vector<std::vector<Coordinate>>& tLine) {
        fetch();
        plane_ = planeOf(tLine.at(0).back());
        const int Ly = static_cast<int>(ty_.size()), S = static_cast<int>(d_[2]), Sa = static_cast<int>(d_[3]);
        auto fill = [&](std::vector<std::vector<Coordinate>>& t, const std::vector<double>& src, int s0, int s1) {
            t.assign(static_cast<size_t>(Ly), std::vector<Coordinate>(static_cast<size_t>(s1 - s0)));
            for (int iy = 0; iy < Ly; ++iy)
                for (int s = s0; s < s1; ++s) t[iy][s - s0] = src[(static_cast<size_t>(plane_) * Ly + iy) * S + s];
        };
        fill(intersect_.arenaLow, lo_, 0, Sa), fill(intersect_.arenaUpp, up_, 0, Sa);
        fill(intersect_.obstacleLow, lo_, Sa, S), fill(intersect_.obstacleUpp, up_, Sa, S);
    }
    const IntersectionInterval& getIntersectionInterval() const { return intersect_; }
    /** free segments of the plane selected by the last computeIntersectionInterval (FreeSpace-inl.h:63-97 incl. enhanceFreeSegment) */
    void computeFreeSegment(const std::vector<Coordinate>& ty) {
        fetch();
        const size_t Ly = ty_.size();
        segment_.ty = ty;
        segment_.xL.assign(Ly, {}), segment_.xU.assign(Ly, {}), segment_.xM.assign(Ly, {});
        for (size_t iy = 0; iy < Ly; ++iy) {
            const size_t l = static_cast<size_t>(plane_) * Ly + iy;
            segment_.xL[iy].assign(xL_.begin() + off_[l], xL_.begin() + off_[l + 1]);
            segment_.xU[iy].assign(xU_.begin() + off_[l], xU_.begin() + off_[l + 1]);
            segment_.xM[iy].assign(xM_.begin() + off_[l], xM_.begin() + off_[l + 1]);
        }
    }
    const FreeSegment2D& getFreeSegment() const { return segment_; }
    const std::vector<Coordinate>& tx() { fetch(); return tx_; }
    const std::vector<Coordinate>& ty() { fetch(); return ty_; }

  private:
    void fetch() {
        if (!off_.empty()) return;
        const int L = static_cast<int>(d_[4]), S = static_cast<int>(d_[2]);
        lo_.resize(static_cast<size_t>(L) * S), up_.resize(static_cast<size_t>(L) * S);
        fs_.check(hrmgpu_get_intervals(fs_.handle(), k_, lo_.data(), up_.data()));
        off_.resize(static_cast<size_t>(L) + 1);
        const int cnt = hrmgpu_get_segments(fs_.handle(), k_, off_.data(), nullptr, nullptr, nullptr, 0);
        fs_.check(cnt < 0 ? cnt : 0);
        xL_.resize(static_cast<size_t>(cnt)), xU_.resize(static_cast<size_t>(cnt)), xM_.resize(static_cast<size_t>(cnt));
        if (cnt > 0) fs_.check(hrmgpu_get_segments(fs_.handle(), k_, off_.data(), xL_.data(), xU_.data(), xM_.data(), cnt) < 0 ? -1 : 0);
        fs_.sweepGrid(tx_, ty_);
    }
    int planeOf(Coordinate x) const {
        for (size_t i = 0; i < tx_.size(); ++i)
            if (tx_[i] == x) return static_cast<int>(i);
        throw std::invalid_argument("computeIntersectionInterval: x is not a sweep plane of the grid given to setup()");
    }
    FreeSpaceGPU3D& fs_;
    int k_;
    long long d_[8];
    int plane_ = 0;
    BoundaryInfo bound_;
    IntersectionInterval intersect_;
    FreeSegment2D segment_;
    std::vector<double> lo_, up_, xL_, xU_, xM_, tx_, ty_;
    std::vector<int> off_;
};

class FreeSpaceGPU2D : public FreeSpaceGPU {
  public:
    FreeSpaceGPU2D(const std::vector<SuperEllipse>& arena, const std::vector<SuperEllipse>& obstacle, int device = 0) : FreeSpaceGPU(device) {
        std::vector<double> se;
        auto push = [&](const SuperEllipse& s) {
            se.push_back(s.getSemiAxis()[0]), se.push_back(s.getSemiAxis()[1]), se.push_back(s.getEpsilon());
            se.push_back(s.getPosition()[0]), se.push_back(s.getPosition()[1]), se.push_back(s.getAngle());
        };
        for (const auto& a : arena) push(a);
        for (const auto& o : obstacle) push(o);
        check(hrmgpu_set_scene2d(ctx_, static_cast<int>(arena.size()), static_cast<int>(obstacle.size()), se.data(),
                                 static_cast<int>(arena.at(0).getNum())));
    }
    void buildLayers(const std::vector<MultiBodyTree2D>& poses, int precision) {
        std::vector<double> semi, ang, off, semiK;
        for (size_t k = 0; k < poses.size(); ++k) {
            semiK.clear();
            poses[k].bodyArrays(semiK, ang, off);
            if (k == 0) semi = semiK;
        }
        check(hrmgpu_build_layers2d(ctx_, static_cast<int>(poses.size()), static_cast<int>(semi.size() / 2), semi.data(), ang.data(), off.data(),
                                    precision));
    }
    void buildBridge(const std::vector<std::vector<Ellipse>>& tfe, int precision) {
        std::vector<double> semi, ang;
        for (const auto& pr : tfe)
            for (const auto& e : pr) semi.push_back(e.semi[0]), semi.push_back(e.semi[1]), ang.push_back(e.angle);
        check(hrmgpu_build_bridge2d(ctx_, static_cast<int>(tfe.size()), static_cast<int>(tfe.at(0).size()), semi.data(), ang.data(), precision));
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// planners
// ---------------------------------------------------------------------------------------------------------------------
struct PlannerParameter {  // include/hrm/planners/PlanningRequest.h:10-32
    std::vector<Coordinate> boundaryLimits;
    Index numSlice = 0, numLineX = 0, numLineY = 0, numPoint = 5, numSearchNeighbor = 10;
    double searchRadius = HALF_PI;
};
struct PlanningRequest {  // :35-47
    bool isRobotRigid = true;
    PlannerParameter parameters;
    std::vector<Coordinate> start, goal;
};
using Edge = std::vector<std::pair<Index, Index>>;
struct Graph {  // include/hrm/planners/PlanningResult.h:14-23
    std::vector<std::vector<Coordinate>> vertex;
    Edge edge;
    std::vector<double> weight;
};
struct SolutionPathInfo {  // :26-38
    std::vector<Index> PathId;
    std::vector<std::vector<Coordinate>> solvedPath, interpolatedPath;
    double cost = 0;             // sum of the Euclidean lengths of the path's edges
    double costAsReference = 0;  // what the reference accumulates: edge weights indexed by VERTEX id
};
struct Time {  // :41-50
    double buildTime = 0, searchTime = 0, totalTime = 0;
};
struct PlanningResult {  // :53-65
    bool solved = false;
    Time planningTime;
    Graph graphStructure;
    SolutionPathInfo solutionPath;
};

namespace planners {

/** Connected components of the roadmap (union-find over the edge list), built once per search(), and -- only when some
 * (start, goal) candidate pair is connected at all -- its adjacency in CSR form (neighbours of a vertex in edge order, as pushing
 * both directions of every edge in turn produces them), shared by the A* runs of that search(). */
struct GraphIndex {
    mutable std::vector<size_t> off;  // (filled on demand: needAdjacency)
    mutable std::vector<std::pair<Index, double>> nbr;
    std::vector<Index> comp;
    explicit GraphIndex(const Graph& g) : g_(&g) {
        const Index n = g.vertex.size();
        comp.resize(n);
        for (Index i = 0; i < n; ++i) comp[i] = i;
        auto find = [&](Index x) {
            while (comp[x] != x) comp[x] = comp[comp[x]], x = comp[x];
            return x;
        };
        for (const auto& e : g.edge) {
            const Index a = find(e.first), b = find(e.second);
            if (a != b) comp[a < b ? b : a] = a < b ? a : b;
        }
        for (Index i = 0; i < n; ++i) comp[i] = find(i);
    }
    bool connected(Index a, Index b) const { return comp[a] == comp[b]; }
    void needAdjacency() const {
        if (!off.empty()) return;
        const Graph& g = *g_;
        const Index n = g.vertex.size();
        auto& o = off;
        auto& nb = nbr;
        o.assign(n + 1, 0);
        for (const auto& e : g.edge) ++o[e.first + 1], ++o[e.second + 1];
        for (Index i = 0; i < n; ++i) o[i + 1] += o[i];
        nb.resize(o[n]);
        std::vector<size_t> fill(o.begin(), o.end() - 1);
        for (size_t i = 0; i < g.edge.size(); ++i) {
            nb[fill[g.edge[i].first]++] = {g.edge[i].second, g.weight[i]};
            nb[fill[g.edge[i].second]++] = {g.edge[i].first, g.weight[i]};
        }
    }

  private:
    const Graph* g_;
};

/** A* with the Euclidean heuristic of HighwayRoadMap-inl.h:135-147 (Boost.Graph replaced by a binary heap).  The heuristic of a
 * vertex is evaluated once, when the vertex is first reached (the same value every later relaxation would compute). */
inline bool astar(const Graph& g, const GraphIndex& gi, Index s, Index t, std::vector<Index>& pred) {
    const Index n = g.vertex.size();
    if (!gi.connected(s, t)) return false;  // (the search below would exhaust s's component and fail)
    gi.needAdjacency();
    std::vector<double> d(n, INFINITY), h(n, -1.0);
    pred.assign(n, n);
    using QE = std::pair<double, Index>;
    std::priority_queue<QE, std::vector<QE>, std::greater<QE>> pq;
    d[s] = 0;
    pred[s] = s;
    pq.push({vectorEuclidean(g.vertex[s], g.vertex[t]), s});
    std::vector<char> closed(n, 0);
    while (!pq.empty()) {
        const Index u = pq.top().second;
        pq.pop();
        if (closed[u]) continue;
        closed[u] = 1;
        if (u == t) return true;
        for (size_t k = gi.off[u]; k < gi.off[u + 1]; ++k) {
            const auto& e = gi.nbr[k];
            const double nd = d[u] + e.second;
            if (nd < d[e.first]) {
                d[e.first] = nd;
                pred[e.first] = u;
                if (h[e.first] < 0.0) h[e.first] = vectorEuclidean(g.vertex[e.first], g.vertex[t]);
                pq.push({nd + h[e.first], e.first});
            }
        }
    }
    return false;
}
inline bool astar(const Graph& g, Index s, Index t, std::vector<Index>& pred) { return astar(g, GraphIndex(g), s, t, pred); }

/** Wall-clock per planner stage, printed at the end of plan() when the environment variable HRM_HOST_TIMING is set. */
struct StageTimes {
    std::vector<std::pair<std::string, double>> acc;
    void add(const char* name, double sec) {
        for (auto& a : acc)
            if (a.first == name) {
                a.second += sec;
                return;
            }
        acc.push_back({name, sec});
    }
    void print() const {
        if (!std::getenv("HRM_HOST_TIMING")) return;
        for (const auto& a : acc) std::fprintf(stderr, "[hrm timing] %-28s %9.3f ms\n", a.first.c_str(), a.second * 1e3);
    }
};
struct StageScope {
    StageTimes& t;
    const char* name;
    std::chrono::high_resolution_clock::time_point t0 = std::chrono::high_resolution_clock::now();
    StageScope(StageTimes& t_, const char* n) : t(t_), name(n) {}
    ~StageScope() { t.add(name, std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count()); }
};

class HighwayRoadMapBase {
  public:
    HighwayRoadMapBase(const PlanningRequest& req, int precision, bool perOrientation)
        : start_(req.start), goal_(req.goal), param_(req.parameters), precision_(precision), perOrientation_(perOrientation) {}
    virtual ~HighwayRoadMapBase() = default;

    /** HighwayRoadMap-inl.h:65-89: build, search, then -- exactly like the reference -- refine the roadmap while it is unsolved and
     *  the planning time is below timeLim (:78-81).  setMaxRefine(rounds) is an ADDITIONAL bound (default: none) that makes runs
     *  reproducible for the parity tests; it never replaces the time limit. */
    virtual void plan(double timeLim = 0.0) {
        auto t0 = std::chrono::high_resolution_clock::now();
        buildRoadmap();
        res_.planningTime.buildTime = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
        t0 = std::chrono::high_resolution_clock::now();
        search();
        stage_.add("search", std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count());
        res_.planningTime.searchTime = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
        res_.planningTime.totalTime = res_.planningTime.buildTime + res_.planningTime.searchTime;
        // HighwayRoadMap-inl.h:78-81
        while (!res_.solved && res_.planningTime.totalTime < timeLim && numRefineDone_ < maxRefine_) {
            t0 = std::chrono::high_resolution_clock::now();
            refineExistRoadmap();
            ++numRefineDone_;
            res_.planningTime.buildTime += std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
            res_.planningTime.totalTime = res_.planningTime.buildTime + res_.planningTime.searchTime;
        }
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
    /** HighwayRoadMap-inl.h:284-291: the solved path itself (HRM3D overrides) */
    virtual std::vector<std::vector<Coordinate>> getInterpolatedSolutionPath(Index /*num*/) { return res_.solutionPath.solvedPath; }
    const PlanningResult& getPlanningResult() const { return res_; }
    const PlannerParameter& getPlannerParameters() const { return param_; }
    long numEdgeTests() const { return numEdgeTests_; }
    long long numBridgeCandidates() const { return numBridgeCandidates_; }
    /** connectMultiSlice's vertex matching on the device (default) or enumerated on the host (the round-1 path; identical roadmaps) */
    void setDeviceMatching(bool on) { deviceMatching_ = on; }
    /** tightly-fitted ellipsoids of the bridge layers fitted on the device (default) or on the host (the round-1 path) */
    void setDeviceTfe(bool on) { deviceTfe_ = on; }
    void setMaxRefine(size_t rounds) { maxRefine_ = rounds; }
    size_t numRefineDone() const { return numRefineDone_; }

  protected:
    virtual void buildRoadmap() = 0;
    /** refineExistRoadmap (:178-215): doubled sweep lines on the cached C-boundaries, new vertices tied to the existing ones */
    virtual void refineExistRoadmap() {}
    virtual std::vector<Index> getNearestNeighborsOnGraph(const std::vector<Coordinate>& vertex, Index k, double radius) = 0;

    /** HighwayRoadMap-inl.h:110-175 */
    void search() {
        Graph& g = res_.graphStructure;
        if (g.vertex.empty()) return;
        auto tS = std::chrono::high_resolution_clock::now();
        const auto idxS = getNearestNeighborsOnGraph(start_, param_.numSearchNeighbor, param_.searchRadius);
        const auto idxG = getNearestNeighborsOnGraph(goal_, param_.numSearchNeighbor, param_.searchRadius);
        stage_.add("  search: nearest vertices", std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - tS).count());
        tS = std::chrono::high_resolution_clock::now();
        const GraphIndex gi(g);
        stage_.add("  search: adjacency + components", std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - tS).count());
        StageScope ssA(stage_, "  search: A*");
        for (Index s : idxS)
            for (Index t : idxG) {
                std::vector<Index> p;
                if (astar(g, gi, s, t, p)) {
                    auto& sp = res_.solutionPath;
                    sp.PathId.clear();
                    sp.cost = 0, sp.costAsReference = 0;
                    Index cur = t;
                    sp.PathId.push_back(cur);
                    while (cur != s) {
                        sp.cost += vectorEuclidean(g.vertex[cur], g.vertex[p[cur]]);
                        if (cur < g.weight.size()) sp.costAsReference += g.weight[cur];  // HighwayRoadMap-inl.h:155-157
                        cur = p[cur];
                        sp.PathId.push_back(cur);
                    }
                    std::reverse(sp.PathId.begin(), sp.PathId.end());
                    res_.solved = true;
                    return;
                }
            }
    }
    void addEdge(Index i, Index j) {
        res_.graphStructure.edge.push_back({i, j});
        res_.graphStructure.weight.push_back(vectorEuclidean(res_.graphStructure.vertex[i], res_.graphStructure.vertex[j]));
    }
    /** one (idx1, idx2) candidate of connectOneSlice: code 0 direct edge, 1 / 2 via bridgeVertex's vNew1 / vNew2 (:295-326), 3 none */
    void connectPair(Index idx1, Index idx2, int code, int zi) {
        auto& g = res_.graphStructure;
        if (code == 0) {
            addEdge(idx1, idx2);
        } else if (code == 1 || code == 2) {
            const auto v1 = g.vertex[idx1], v2 = g.vertex[idx2];
            auto vn = (code == 1) ? v1 : v2;
            vn[zi] = (code == 1) ? v2[zi] : v1[zi];
            const Index idn = g.vertex.size();
            g.vertex.push_back(vn);
            g.edge.push_back({idx1, idn});
            g.weight.push_back(vectorEuclidean(v1, vn));
            g.edge.push_back({idn, idx2});
            g.weight.push_back(vectorEuclidean(vn, v2));
        }
    }
    static int pairCode(bool f0, bool f1, bool f2, bool f3, bool f4) { return f0 ? 0 : ((f1 && f2) ? 1 : ((f3 && f4) ? 2 : 3)); }

    std::vector<Coordinate> start_, goal_;
    PlannerParameter param_;
    PlanningResult res_;
    std::vector<std::pair<Index, Index>> vertexIdx_;  // (startId, slice) per C-layer
    int precision_;
    bool perOrientation_;
    long numEdgeTests_ = 0;
    long long numBridgeCandidates_ = 0;
    bool deviceMatching_ = std::getenv("HRM_HOST_MATCHING") == nullptr;
    bool deviceTfe_ = std::getenv("HRM_HOST_TFE") == nullptr;
    StageTimes stage_;
    size_t maxRefine_ = static_cast<size_t>(-1), numRefineDone_ = 0;  // (no round bound unless setMaxRefine is called)
    std::vector<std::vector<std::pair<Index, Index>>> vertexIdxAll_;  // vertexIdx_ of the previous refinement rounds

    /** connectExistSlice (HRM3D.cpp:226-270, HRM2D.cpp:158-197): vertices of the refined layer against the same layer's
     * vertices of the previous round; `near(v1, v2)` is the sweep-line adjacency filter, edge tests in one K4 batch, replay with
     * the reference's moving lower bound. */
    template <class Near, class TestBatch>
    void connectExistSliceReplay(size_t sliceId, Near near, TestBatch testBatch) {
        const auto& V = res_.graphStructure.vertex;
        const Index cur0 = vertexIdx_.at(sliceId).first, cur1 = vertexIdx_.at(sliceId).second;
        const Index ex0 = vertexIdxAll_.back().at(sliceId).first, ex1 = vertexIdxAll_.back().at(sliceId).second;
        std::vector<std::pair<Index, Index>> cand;
        for (Index m0 = cur0; m0 < cur1; ++m0)
            for (Index m1 = ex0; m1 < ex1; ++m1)
                if (near(V[m0], V[m1])) cand.push_back({m0, m1});
        if (cand.empty()) return;
        const std::vector<unsigned char> ok = testBatch(cand);
        numEdgeTests_ += static_cast<long>(cand.size());
        std::map<std::pair<Index, Index>, bool> freePair;
        for (size_t c = 0; c < cand.size(); ++c) freePair[cand[c]] = ok[c] != 0;
        Index startIdExist = ex0;
        for (Index m0 = cur0; m0 < cur1; ++m0)
            for (Index m1 = startIdExist; m1 < ex1; ++m1) {
                const auto it = freePair.find({m0, m1});
                if (it == freePair.end() || !it->second) continue;
                addEdge(m0, m1);
                startIdExist = m1;
                break;
            }
    }
};

/** src/planners/HRM3D.cpp */
class HRM3D : public HighwayRoadMapBase {
  public:
    HRM3D(const MultiBodyTree3D& robot, const std::vector<SuperQuadrics>& arena, const std::vector<SuperQuadrics>& obs, const PlanningRequest& req,
          int device = 0, int precision = HRMGPU_F64_STRICT, bool perOrientation = false)
        : HighwayRoadMapBase(req, precision, perOrientation), robot_(robot), freeSpace_(arena, obs, device) {
        q_ = robot_.getBase().getQuatSamples();  // pre-defined samples (HRM3D.cpp:436-440); random sampling must be injected (App. C12)
        param_.numSlice = q_.size();
    }
    FreeSpaceGPU3D& freeSpace() { return freeSpace_; }

    /** HRM3D.cpp:272-299: about `num` interpolated poses per unit of average edge length along the solved path.  The step
     * length uses the geometric path cost (the reference divides its vertex-indexed cost, App. C11); a segment shorter than one
     * step contributes its two end poses (the reference interpolates with 1/0 there). */
    std::vector<std::vector<Coordinate>> getInterpolatedSolutionPath(Index num) override {
        const auto& sp = res_.solutionPath.solvedPath;
        std::vector<std::vector<Coordinate>> out;
        if (sp.size() < 2 || num == 0) return sp;
        double length = 0.0;
        for (size_t i = 0; i + 1 < sp.size(); ++i) length += vectorEuclidean(sp[i], sp[i + 1]);
        const double distanceStep = length / (static_cast<double>(num) * (static_cast<double>(sp.size()) - 1.0));
        for (size_t i = 0; i + 1 < sp.size(); ++i) {
            const int numStep = distanceStep > 0.0 ? static_cast<int>(vectorEuclidean(sp[i], sp[i + 1]) / distanceStep) : 0;
            if (numStep <= 0) {
                out.push_back(sp[i]), out.push_back(sp[i + 1]);
                continue;
            }
            const auto step = interpolateCompoundSE3Rn(sp[i], sp[i + 1], static_cast<Index>(numStep));
            out.insert(out.end(), step.begin(), step.end());
        }
        return out;
    }

  protected:
    void buildRoadmap() override {
        const auto& lim = param_.boundaryLimits;
        const int Lx = static_cast<int>(param_.numLineX), Ly = static_cast<int>(param_.numLineY);
        freeSpace_.setup(lim, param_.numLineX, param_.numLineY);
        // K1+K2+K3 for every orientation in ONE call
        std::vector<MultiBodyTree3D> poses;
        std::vector<Quaternion>& baseQ = baseQ_;
        baseQ.clear();
        MultiBodyTree3D rb = robot_;
        for (const auto& q : q_) {
            SE3Transform g;
            g.R = toRotationMatrix(q);
            rb.robotTF(g);  // HRM3D.cpp:27-31,510-518 (planner's robot)
            baseQ.push_back(rb.getBase().getQuaternion());
            poses.push_back(perOrientation_ ? rb : robot_);  // the reference's FreeSpace copy keeps the as-loaded pose (finding 1)
        }
        {
            StageScope ss(stage_, "buildLayers (K1-K3)");
            freeSpace_.buildLayers(poses, precision_);
        }
        std::vector<long long> off;
        std::vector<double> xL, xU, xM;
        {
            StageScope ss(stage_, "getFreeSegmentsAll");
            freeSpace_.getFreeSegmentsAll(off, xL, xU, xM);
        }
        const double dx = (lim[1] - lim[0]) / static_cast<double>(Lx - 1), dy = (lim[3] - lim[2]) / static_cast<double>(Ly - 1);
        std::vector<double> tx(Lx), ty(Ly);
        for (int i = 0; i < Lx; ++i) tx[i] = lim[0] + static_cast<double>(i) * dx;
        for (int i = 0; i < Ly; ++i) ty[i] = lim[2] + static_cast<double>(i) * dy;
        {
            // connectOneSlice for every layer: the candidate segments depend on the free segments only, so they are collected
            // for ALL layers, tested in ONE K4 launch, and the reference's loops are replayed layer by layer on the answers
            StageScope ss(stage_, "connectOneSlice (all)");
            if (deviceMatching_) {
                // the 5 candidate segments of every vertex pair are generated and tested on the device from the free segments
                // (hrmgpu_connect_slices3d); the host counts the pairs of every layer and replays the reference's loops on the codes
                std::vector<SlicePairs> sp(q_.size());
                size_t total = 0;
                for (size_t k = 0; k < q_.size(); ++k) {
                    const long long* o = off.data() + k * Lx * Ly;
                    auto segN = [&](int ix, int iy) { return static_cast<size_t>(o[ix * Ly + iy + 1] - o[ix * Ly + iy]); };
                    sp[k].first = total;
                    for (int ix = 0; ix < Lx; ++ix)
                        for (int iy = 0; iy + 1 < Ly; ++iy) sp[k].nInPlane += segN(ix, iy) * segN(ix, iy + 1);
                    sp[k].P = sp[k].nInPlane;
                    for (int ix = 0; ix + 1 < Lx; ++ix)
                        for (int iy = 0; iy < Ly; ++iy) sp[k].P += segN(ix, iy) * segN(ix + 1, iy);
                    total += sp[k].P;
                }
                std::vector<unsigned char> code;
                std::vector<long long> layerFirst;
                {
                    StageScope st(stage_, "  K4 connectSlices (device)");
                    code = freeSpace_.connectSlices3D(layerFirst, total, q_.size());
                }
                numEdgeTests_ += static_cast<long>(5 * code.size());
                for (size_t k = 0; k < q_.size(); ++k) {
                    if (static_cast<size_t>(layerFirst[k]) != sp[k].first) throw std::runtime_error("connectSlices3D: pair count mismatch");
                    const std::vector<int> codes(code.begin() + sp[k].first, code.begin() + sp[k].first + sp[k].P);
                    replaySlice3D(baseQ[k], tx, ty, off.data() + k * Lx * Ly, xL, xU, xM, sp[k], codes);
                }
            } else {
            std::vector<double> A, Bq;
            std::vector<int> layerOf;
            std::vector<SlicePairs> sp(q_.size());
            for (size_t k = 0; k < q_.size(); ++k) {
                sp[k] = gatherSlicePairs3D(tx, ty, off.data() + k * Lx * Ly, xM, A, Bq);
                layerOf.insert(layerOf.end(), 5 * sp[k].P, static_cast<int>(k));
            }
            std::vector<unsigned char> f;
            {
                StageScope st(stage_, "  K4 testEdges");
                f = freeSpace_.testEdgesMulti(layerOf, A, Bq);
            }
            numEdgeTests_ += static_cast<long>(f.size());
            for (size_t k = 0; k < q_.size(); ++k)
                replaySlice3D(baseQ[k], tx, ty, off.data() + k * Lx * Ly, xL, xU, xM, sp[k], pairCodes(f.data() + sp[k].first, sp[k].P));
            }
        }
        StageScope ss(stage_, "connectMultiSlice");
        connectMultiSlice();
    }

    void refineExistRoadmap() override {
        vertexIdxAll_.push_back(vertexIdx_);
        vertexIdx_.clear();
        param_.numLineX *= 2;
        param_.numLineY *= 2;
        const auto& lim = param_.boundaryLimits;
        const int Lx = static_cast<int>(param_.numLineX), Ly = static_cast<int>(param_.numLineY);
        freeSpace_.resweep(lim, param_.numLineX, param_.numLineY);  // K2 + K3 of every layer in one call, K1 untouched
        std::vector<long long> off;
        std::vector<double> xL, xU, xM;
        freeSpace_.getFreeSegmentsAll(off, xL, xU, xM);
        const double dx = (lim[1] - lim[0]) / static_cast<double>(Lx - 1), dy = (lim[3] - lim[2]) / static_cast<double>(Ly - 1);
        std::vector<double> tx(Lx), ty(Ly);
        for (int i = 0; i < Lx; ++i) tx[i] = lim[0] + static_cast<double>(i) * dx;
        for (int i = 0; i < Ly; ++i) ty[i] = lim[2] + static_cast<double>(i) * dy;
        const double thx = 2.0 * std::fabs(lim[1] - lim[0]) / static_cast<double>(param_.numLineX);
        const double thy = 2.0 * std::fabs(lim[3] - lim[2]) / static_cast<double>(param_.numLineY);
        for (size_t k = 0; k < q_.size(); ++k) {
            connectOneSlice3D(static_cast<int>(k), baseQ_[k], tx, ty, off.data() + k * static_cast<size_t>(Lx) * Ly, xL, xU, xM);
            const auto& V = res_.graphStructure.vertex;
            connectExistSliceReplay(
                k, [&](const std::vector<Coordinate>& a, const std::vector<Coordinate>& b) { return !(std::fabs(a[0] - b[0]) > thx) && !(std::fabs(a[1] - b[1]) > thy); },
                [&](const std::vector<std::pair<Index, Index>>& cand) {
                    std::vector<double> A, Bq;
                    for (const auto& c : cand)
                        for (int d = 0; d < 3; ++d) A.push_back(V[c.first][d]), Bq.push_back(V[c.second][d]);
                    return freeSpace_.testEdges(static_cast<int>(k), A, Bq, 3);
                });
            search();
            if (res_.solved) return;
        }
    }

    /** Candidate segments of one layer in connectOneSlice's enumeration order: P pairs, each expanded into the 5 segments
     * {v1-v2, v1-vNew1, vNew1-v2, v1-vNew2, vNew2-v2} of bridgeVertex (HighwayRoadMap-inl.h:295-326), stored variant-major:
     * segment (t, i) is entry first + t * P + i of the global arrays. */
    struct SlicePairs {
        size_t first = 0, nInPlane = 0, P = 0;
    };
    SlicePairs gatherSlicePairs3D(const std::vector<double>& tx, const std::vector<double>& ty, const long long* o, const std::vector<double>& xM,
                                  std::vector<double>& A, std::vector<double>& Bq) const {
        const int Lx = static_cast<int>(tx.size()), Ly = static_cast<int>(ty.size());
        auto seg0 = [&](int ix, int iy) { return o[ix * Ly + iy]; };
        auto segN = [&](int ix, int iy) { return static_cast<int>(o[ix * Ly + iy + 1] - o[ix * Ly + iy]); };
        SlicePairs sp;
        sp.first = A.size() / 3;
        for (int ix = 0; ix < Lx; ++ix)
            for (int iy = 0; iy + 1 < Ly; ++iy) sp.nInPlane += static_cast<size_t>(segN(ix, iy)) * segN(ix, iy + 1);
        sp.P = sp.nInPlane;
        for (int ix = 0; ix + 1 < Lx; ++ix)
            for (int iy = 0; iy < Ly; ++iy) sp.P += static_cast<size_t>(segN(ix, iy)) * segN(ix + 1, iy);
        const size_t P = sp.P;
        A.resize(3 * (sp.first + 5 * P));
        Bq.resize(3 * (sp.first + 5 * P));
        double* a = A.data() + 3 * sp.first;
        double* b = Bq.data() + 3 * sp.first;
        size_t i = 0;
        auto put = [&](double x1, double y1, double z1, double x2, double y2, double z2) {
            const double v1[3] = {x1, y1, z1}, v2[3] = {x2, y2, z2};
            const double n1[3] = {x1, y1, z2}, n2[3] = {x2, y2, z1};  // vNew1, vNew2 (HighwayRoadMap-inl.h:301-304)
            const double* as[5] = {v1, v1, n1, v1, n2};
            const double* bs[5] = {v2, n1, v2, n2, v2};
            for (int t = 0; t < 5; ++t)
                for (int d = 0; d < 3; ++d) a[(t * P + i) * 3 + d] = as[t][d], b[(t * P + i) * 3 + d] = bs[t][d];
            ++i;
        };
        for (int ix = 0; ix < Lx; ++ix)
            for (int iy = 0; iy + 1 < Ly; ++iy)
                for (int j1 = 0; j1 < segN(ix, iy); ++j1)
                    for (int j2 = 0; j2 < segN(ix, iy + 1); ++j2)
                        put(tx[ix], ty[iy], xM[seg0(ix, iy) + j1], tx[ix], ty[iy + 1], xM[seg0(ix, iy + 1) + j2]);
        for (int ix = 0; ix + 1 < Lx; ++ix)
            for (int iy = 0; iy < Ly; ++iy)
                for (int k1 = 0; k1 < segN(ix, iy); ++k1)
                    for (int k2 = 0; k2 < segN(ix + 1, iy); ++k2)
                        put(tx[ix], ty[iy], xM[seg0(ix, iy) + k1], tx[ix + 1], ty[iy], xM[seg0(ix + 1, iy) + k2]);
        return sp;
    }
    /** connection code of every pair from the 5 answers of its segments (variant-major, see SlicePairs) */
    static std::vector<int> pairCodes(const unsigned char* f, size_t P) {
        std::vector<int> codes(P);
        for (size_t i = 0; i < P; ++i) codes[i] = pairCode(f[i], f[P + i], f[2 * P + i], f[3 * P + i], f[4 * P + i]);
        return codes;
    }

    /** generateVertices + connectOneSlice2D per plane + cross-plane connections (HRM3D.cpp:93-156) of ONE layer with its own
     * K4 call (refinement and ProbHRM3D build layer by layer; HRM3D::buildRoadmap batches all layers instead) */
    void connectOneSlice3D(int k, const Quaternion& q, const std::vector<double>& tx, const std::vector<double>& ty, const long long* o,
                           const std::vector<double>& xL, const std::vector<double>& xU, const std::vector<double>& xM) {
        std::vector<double> A, Bq;
        const SlicePairs sp = gatherSlicePairs3D(tx, ty, o, xM, A, Bq);
        std::vector<unsigned char> f;
        if (sp.P) {
            StageScope ss(stage_, "  K4 testEdges");
            f = freeSpace_.testEdges(k, A, Bq, 3);
        }
        numEdgeTests_ += static_cast<long>(5 * sp.P);
        replaySlice3D(q, tx, ty, o, xL, xU, xM, sp, pairCodes(f.data(), sp.P));
    }

    /** the reference's loops of one layer replayed on the pair codes, so that vertex ids and edge order are the reference's */
    void replaySlice3D(const Quaternion& q, const std::vector<double>& tx, const std::vector<double>& ty, const long long* o,
                       const std::vector<double>& xL, const std::vector<double>& xU, const std::vector<double>& xM, const SlicePairs& sp,
                       const std::vector<int>& codes) {
        const int Lx = static_cast<int>(tx.size()), Ly = static_cast<int>(ty.size());
        auto seg0 = [&](int ix, int iy) { return o[ix * Ly + iy]; };
        auto segN = [&](int ix, int iy) { return static_cast<int>(o[ix * Ly + iy + 1] - o[ix * Ly + iy]); };
        const size_t nInPlane = sp.nInPlane;
        auto& V = res_.graphStructure.vertex;
        const Index startId = V.size();
        std::vector<std::vector<Index>> lineIds;
        size_t pc = 0;
        for (int ix = 0; ix < Lx; ++ix) {
            std::vector<Index> plane;
            for (int iy = 0; iy < Ly; ++iy) {  // generateVertices
                plane.push_back(V.size());
                for (int j = 0; j < segN(ix, iy); ++j) {
                    V.push_back({tx[ix], ty[iy], xM[seg0(ix, iy) + j], q[0], q[1], q[2], q[3]});
                    V.back().insert(V.back().end(), vertexTail_.begin(), vertexTail_.end());  // joint angles (ProbHRM3D.cpp:185-188)
                }
            }
            lineIds.push_back(plane);
            // connectOneSlice2D (HighwayRoadMap-inl.h:218-261); its pair order is the enumeration order above restricted to plane ix
            std::vector<size_t> base(Ly, 0);  // index of the first pair of (ix, iy) in `pairs`
            {
                size_t acc = pc;
                for (int iy = 0; iy + 1 < Ly; ++iy) base[iy] = acc, acc += static_cast<size_t>(segN(ix, iy)) * segN(ix, iy + 1);
                pc = acc;
            }
            for (int iy = 0; iy < Ly; ++iy) {
                const Index n1 = plane[iy];
                const int cnt = segN(ix, iy);
                for (int j1 = 0; j1 < cnt; ++j1) {
                    if (j1 != cnt - 1 && std::fabs(xU[seg0(ix, iy) + j1] - xL[seg0(ix, iy) + j1 + 1]) < 1e-6) addEdge(n1 + j1, n1 + j1 + 1);
                    if (iy != Ly - 1) {
                        const Index n2 = plane[iy + 1];
                        const int c2 = segN(ix, iy + 1);
                        for (int j2 = 0; j2 < c2; ++j2) connectPair(n1 + j1, n2 + j2, codes[base[iy] + static_cast<size_t>(j1) * c2 + j2], 2);
                    }
                }
            }
        }
        size_t pi = nInPlane;
        for (int ix = 0; ix + 1 < Lx; ++ix)  // HRM3D.cpp:131-155
            for (int iy = 0; iy < Ly; ++iy) {
                const Index n1 = lineIds[ix][iy], n2 = lineIds[ix + 1][iy];
                for (int k1 = 0; k1 < segN(ix, iy); ++k1)
                    for (int k2 = 0; k2 < segN(ix + 1, iy); ++k2) connectPair(n1 + k1, n2 + k2, codes[pi++], 2);
            }
        vertexIdx_.push_back({startId, V.size()});  // numVertex_.slice is re-read after constructOneSlice (HighwayRoadMap-inl.h:100-102)
    }

    /** HRM3D.cpp:158-224 with computeTFE (:521-539), bridgeSlice (:301-317), isMultiSliceTransitionFree (:367-392) */
    void connectMultiSlice() {
        if (vertexIdx_.size() == 1) return;
        auto& V = res_.graphStructure.vertex;
        const auto& lim = param_.boundaryLimits;
        const size_t B = 1 + robot_.getNumLinks();
        const auto bodies = robot_.getBodyShapes();
        std::vector<int> nearest;
        std::vector<std::vector<Ellipsoid>> tfe;
        std::vector<double> tfeQa, tfeQb, bodySemi;  // device TFE: the two orientations of every (pair, body)
        for (size_t b = 0; b < B; ++b)
            for (int d = 0; d < 3; ++d) bodySemi.push_back(bodies[b].getSemiAxis()[d]);
        auto tTfe = std::chrono::high_resolution_clock::now();
        for (size_t i = 0; i < vertexIdx_.size(); ++i) {
            double minDist = INFINITY;
            int minIdx = 0;
            for (size_t j = 0; j != i && j < param_.numSlice; ++j) {  // only earlier layers (App. C9)
                const double d = angularDistance(q_[i], q_[j]);
                if (d < minDist) minDist = d, minIdx = static_cast<int>(j);
            }
            nearest.push_back(minIdx);
            if (deviceTfe_) {  // HRM3D.cpp:521-539: the orientations only; the ellipsoids are fitted on the device
                for (size_t b = 0; b < B; ++b) {
                    Quaternion qa = q_[i], qb = q_[minIdx];
                    if (b > 0) {
                        const Mat3& Rl = robot_.getTF()[b - 1].R;
                        qa = toQuaternion(matMul(toRotationMatrix(q_[i]), Rl));
                        qb = toQuaternion(matMul(toRotationMatrix(q_[minIdx]), Rl));
                    }
                    tfeQa.insert(tfeQa.end(), qa.begin(), qa.end());
                    tfeQb.insert(tfeQb.end(), qb.begin(), qb.end());
                }
                continue;
            }
            std::vector<Ellipsoid> pr;
            for (size_t b = 0; b < B; ++b) {
                const Vec3 semi = {bodies[b].getSemiAxis()[0], bodies[b].getSemiAxis()[1], bodies[b].getSemiAxis()[2]};
                if (b == 0) {
                    pr.push_back(getTFE3D(semi, q_[i], q_[minIdx], param_.numPoint));
                } else {
                    const Mat3& Rl = robot_.getTF()[b - 1].R;
                    pr.push_back(getTFE3D(semi, toQuaternion(matMul(toRotationMatrix(q_[i]), Rl)),
                                          toQuaternion(matMul(toRotationMatrix(q_[minIdx]), Rl)), param_.numPoint));
                }
            }
            tfe.push_back(pr);
        }
        stage_.add(deviceTfe_ ? "  TFE orientations (host)" : "  TFE (host)", std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - tTfe).count());
        {
            StageScope ss(stage_, deviceTfe_ ? "  K5a TFE + buildBridge (device)" : "  K5a buildBridge");
            if (deviceTfe_)
                freeSpace_.buildBridgeTFE(bodySemi, tfeQa, tfeQb, param_.numPoint, precision_);
            else
                freeSpace_.buildBridge(tfe, precision_);
        }
        const double thx = 2.0 * (lim[1] - lim[0]) / static_cast<double>(param_.numLineX);
        const double thy = 2.0 * (lim[3] - lim[2]) / static_cast<double>(param_.numLineY);
        const size_t steps = param_.numPoint;
        // Candidates of every layer pair (vertices within two line spacings, HRM3D.cpp:195-206), tested in ONE K5 launch.
        // isMultiSliceTransitionFree (:367-392) interpolates v1 -> v2 and moves the robot to every pose (setTransform): all
        // vertices of a layer carry the layer's quaternion, so the slerp rotations R_s and the rotated link offsets R_s * t_l
        // are the same for every candidate of a pair; they are evaluated once per pair with the operations robotTF performs,
        // and the kernel forms the body centres (interpolated position + offset) itself.
        // The reference scans all |layer i| x |layer j| vertex pairs for the ones within two line spacings; here every layer's
        // vertices are bucketed by sweep line (their x, y are grid values), a vertex looks only into the buckets around its own
        // line, and the exact distance test and the ascending vertex order of the reference's inner loop are kept.
        struct PairCand {
            Index start = 0, n2 = 0, cur0 = 0, cur1 = 0;
            std::vector<size_t> first;  // candidates of vertex m0 = start + r: [first[r], first[r + 1]) of `cand`, ascending in m1
        };
        const size_t P = vertexIdx_.size(), T = steps + 1;
        const int Lx = static_cast<int>(param_.numLineX), Ly = static_cast<int>(param_.numLineY);
        const double dx = (lim[1] - lim[0]) / static_cast<double>(Lx - 1), dy = (lim[3] - lim[2]) / static_cast<double>(Ly - 1);
        auto cellOf = [&](const std::vector<Coordinate>& v, int& ix, int& iy) {
            ix = std::min(std::max(static_cast<int>(std::lround((v[0] - lim[0]) / dx)), 0), Lx - 1);
            iy = std::min(std::max(static_cast<int>(std::lround((v[1] - lim[2]) / dy)), 0), Ly - 1);
        };
        const int wx = static_cast<int>(std::floor(thx / dx)) + 1, wy = static_cast<int>(std::floor(thy / dy)) + 1;  // window half-widths (lines)
        std::vector<PairCand> pc(P);
        std::vector<int> pairOf;
        std::vector<double> v0s, v1s, w0(T), w1(T), linkOff(P * T * B * 3, 0.0);
        std::vector<std::pair<Index, Index>> cand;
        const auto& tf = robot_.getTF();
        const double dt = 1.0 / static_cast<double>(steps);
        for (size_t s = 0; s < T; ++s) w1[s] = static_cast<double>(s) * dt, w0[s] = 1.0 - static_cast<double>(s) * dt;  // InterpolateSE3.cpp:5-29
        if (deviceMatching_ && vertexTail_.empty()) {
            // The candidate enumeration, the transition tests and the replay of the moving lower bound run on the device
            // (hrmgpu_connect_pairs3d) from the vertex positions and the id range of every layer; the host only evaluates the
            // per-pair link offsets and appends the matches as edges, in the reference's order.
            std::vector<int> range4(4 * P);
            std::vector<double> xyz(3 * V.size());
            {
                StageScope ss(stage_, "  link offsets (host)");
                for (size_t m = 0; m < V.size(); ++m)
                    for (int d = 0; d < 3; ++d) xyz[3 * m + d] = V[m][d];
                for (size_t i = 0; i < P; ++i) {
                    const Index start = vertexIdx_[nearest[i]].first, n2 = vertexIdx_[nearest[i]].second;
                    range4[4 * i] = static_cast<int>(start), range4[4 * i + 1] = static_cast<int>(n2);
                    range4[4 * i + 2] = static_cast<int>(vertexIdx_[i].first), range4[4 * i + 3] = static_cast<int>(vertexIdx_[i].second);
                    if (n2 <= start || vertexIdx_[i].second <= vertexIdx_[i].first) {
                        range4[4 * i + 1] = range4[4 * i];  // (an empty side: no candidates)
                        continue;
                    }
                    const auto& va = V[start];
                    const auto& vb = V[vertexIdx_[i].first];
                    const auto qs = interpolateSlerp({va[3], va[4], va[5], va[6]}, {vb[3], vb[4], vb[5], vb[6]}, steps);
                    for (size_t s = 0; s < T; ++s) {
                        const Mat3 R = toRotationMatrix(qs[s]);
                        for (size_t l = 0; l < tf.size(); ++l) {
                            const Vec3 v = matVec(R, tf[l].t);
                            for (int d = 0; d < 3; ++d) linkOff[((i * T + s) * B + (l + 1)) * 3 + d] = v[d];
                        }
                    }
                }
            }
            std::vector<int> match;
            {
                StageScope ss(stage_, "  K5b connectPairs (device)");
                match = freeSpace_.connectPairs3D(xyz, range4, thx, thy, w0, w1, linkOff, &numBridgeCandidates_);
            }
            size_t g = 0;
            for (size_t i = 0; i < P; ++i)
                for (int m0 = range4[4 * i]; m0 < range4[4 * i + 1]; ++m0, ++g)
                    if (match[g] >= 0) addEdge(static_cast<Index>(m0), static_cast<Index>(match[g]));
            return;
        }
        {
            StageScope ss(stage_, "  bridge candidates (host)");
            std::vector<std::vector<Index>> cell(static_cast<size_t>(Lx) * Ly);
            std::vector<Index> near;
            for (size_t i = 0; i < P; ++i) {
                PairCand& c = pc[i];
                c.cur0 = vertexIdx_[i].first, c.cur1 = vertexIdx_[i].second;
                c.start = vertexIdx_[nearest[i]].first, c.n2 = vertexIdx_[nearest[i]].second;
                if (c.n2 <= c.start || c.cur1 <= c.cur0) continue;
                const auto& va = V[c.start];
                const auto& vb = V[c.cur0];
                const auto qs = interpolateSlerp({va[3], va[4], va[5], va[6]}, {vb[3], vb[4], vb[5], vb[6]}, steps);
                for (size_t s = 0; s < T; ++s) {
                    const Mat3 R = toRotationMatrix(qs[s]);
                    for (size_t l = 0; l < tf.size(); ++l) {
                        const Vec3 v = matVec(R, tf[l].t);
                        for (int d = 0; d < 3; ++d) linkOff[((i * T + s) * B + (l + 1)) * 3 + d] = v[d];
                    }
                }
                for (auto& cl : cell) cl.clear();
                for (Index m1 = c.cur0; m1 < c.cur1; ++m1) {  // ascending ids: every bucket stays sorted
                    int ix, iy;
                    cellOf(V[m1], ix, iy);
                    cell[static_cast<size_t>(ix) * Ly + iy].push_back(m1);
                }
                c.first.assign(c.n2 - c.start + 1, cand.size());
                for (Index m0 = c.start; m0 < c.n2; ++m0) {
                    const auto& v0 = V[m0];
                    int ix0, iy0;
                    cellOf(v0, ix0, iy0);
                    near.clear();
                    for (int ix = std::max(ix0 - wx, 0); ix <= std::min(ix0 + wx, Lx - 1); ++ix)
                        for (int iy = std::max(iy0 - wy, 0); iy <= std::min(iy0 + wy, Ly - 1); ++iy)
                            for (Index m1 : cell[static_cast<size_t>(ix) * Ly + iy]) {
                                const auto& v1 = V[m1];
                                if (std::fabs(v0[0] - v1[0]) > thx || std::fabs(v0[1] - v1[1]) > thy) continue;  // HRM3D.cpp:195-200
                                near.push_back(m1);
                            }
                    std::sort(near.begin(), near.end());
                    c.first[m0 - c.start] = cand.size();
                    for (Index m1 : near) {
                        const auto& v1 = V[m1];
                        cand.push_back({m0, m1});
                        v0s.insert(v0s.end(), v0.begin(), v0.begin() + 3);
                        v1s.insert(v1s.end(), v1.begin(), v1.begin() + 3);
                        pairOf.push_back(static_cast<int>(i));
                    }
                }
                c.first[c.n2 - c.start] = cand.size();
            }
        }
        std::vector<unsigned char> ok;
        {
            StageScope ss(stage_, "  K5b testTransitions");
            ok = freeSpace_.testTransitions(pairOf, v0s, v1s, w0, w1, linkOff);
        }
        for (size_t i = 0; i < P; ++i) {
            const PairCand& c = pc[i];
            if (c.first.empty()) continue;
            Index n22 = c.cur0;  // replay with the reference's moving lower bound
            for (Index m0 = c.start; m0 < c.n2; ++m0)
                for (size_t j = c.first[m0 - c.start]; j < c.first[m0 - c.start + 1]; ++j) {
                    const Index m1 = cand[j].second;
                    if (m1 < n22 || !ok[j]) continue;
                    addEdge(m0, m1);
                    n22 = m1;
                    break;
                }
        }
    }

    /** HRM3D.cpp:443-508 with minQuat initialised to q_[0] (App. C18) */
    std::vector<Index> getNearestNeighborsOnGraph(const std::vector<Coordinate>& vertex, Index k, double radius) override {
        const auto& V = res_.graphStructure.vertex;
        const auto& lim = param_.boundaryLimits;
        const Quaternion qq = {vertex[3], vertex[4], vertex[5], vertex[6]};
        Quaternion minQuat = q_[0];
        double minD = angularDistance(qq, q_[0]);
        for (const auto& q : q_) {
            const double d = angularDistance(qq, q);
            if (d < minD) minD = d, minQuat = q;
        }
        std::vector<Quaternion> list;
        for (const auto& q : q_) {
            if (angularDistance(minQuat, q) < radius) list.push_back(q);
            if (list.size() >= k) break;
        }
        // The reference scans every vertex for the ones whose orientation is (numerically) qc.  Vertices carry their layer's
        // quaternion verbatim and are stored layer after layer, so the orientation test is made once per run of identical
        // quaternions instead of once per vertex (same comparisons on the same values, same vertex order: same result).
        std::vector<std::pair<size_t, size_t>> runs;
        for (size_t i = 0; i < V.size();) {
            size_t j = i + 1;
            while (j < V.size() && V[j][3] == V[i][3] && V[j][4] == V[i][4] && V[j][5] == V[i][5] && V[j][6] == V[i][6]) ++j;
            runs.push_back({i, j});
            i = j;
        }
        std::vector<Index> idx;
        for (const auto& qc : list) {
            Index idxSlice = 0;
            double best = INFINITY;
            for (const auto& run : runs) {
                if (!(angularDistance(qc, {V[run.first][3], V[run.first][4], V[run.first][5], V[run.first][6]}) < 1e-6)) continue;
                for (size_t i = run.first; i < run.second; ++i) {
                    const double d = vectorEuclidean(vertex, V[i]);
                    if (d < best) best = d, idxSlice = i;
                }
            }
            if (std::abs(vertex[0] - V[idxSlice][0]) < radius * (lim[1] - lim[0]) / static_cast<double>(param_.numLineX) &&
                std::abs(vertex[1] - V[idxSlice][1]) < radius * (lim[3] - lim[2]) / static_cast<double>(param_.numLineY))
                idx.push_back(idxSlice);
        }
        return idx;
    }

    MultiBodyTree3D robot_;
    FreeSpaceGPU3D freeSpace_;
    std::vector<Quaternion> q_;
    std::vector<Coordinate> vertexTail_;  // appended to every vertex of the slice under construction (empty for rigid bodies)
    std::vector<Quaternion> baseQ_;       // base quaternion of every layer as the vertices carry it
};

/** src/planners/HRM2D.cpp */
class HRM2D : public HighwayRoadMapBase {
  public:
    HRM2D(const MultiBodyTree2D& robot, const std::vector<SuperEllipse>& arena, const std::vector<SuperEllipse>& obs, const PlanningRequest& req,
          int device = 0, int precision = HRMGPU_F64_STRICT, bool perOrientation = false)
        : HighwayRoadMapBase(req, precision, perOrientation), robot_(robot), freeSpace_(arena, obs, device) {}
    FreeSpaceGPU2D& freeSpace() { return freeSpace_; }

  protected:
    void buildRoadmap() override {
        const auto& lim = param_.boundaryLimits;
        const int Ly = static_cast<int>(param_.numLineY);
        const double dr = 2 * PI / (static_cast<double>(param_.numSlice) - 1);  // HRM2D.cpp:46-51
        for (size_t i = 0; i < param_.numSlice; ++i) headings_.push_back(-PI + dr * static_cast<double>(i));
        freeSpace_.setup(lim, 1, param_.numLineY);
        std::vector<MultiBodyTree2D> poses;
        std::vector<double>& baseAngle = baseAngle_;
        baseAngle.clear();
        MultiBodyTree2D rb = robot_;
        for (double th : headings_) {
            rb.robotTF(th, 0.0, 0.0);
            baseAngle.push_back(rb.getBase().getAngle());
            poses.push_back(perOrientation_ ? rb : robot_);
        }
        freeSpace_.buildLayers(poses, precision_);
        std::vector<long long> off;
        std::vector<double> xL, xU, xM;
        freeSpace_.getFreeSegmentsAll(off, xL, xU, xM);
        const double dy = (lim[3] - lim[2]) / (static_cast<double>(Ly) - 1);  // :56-61
        std::vector<double> ty(Ly);
        for (int i = 0; i < Ly; ++i) ty[i] = lim[2] + static_cast<double>(i) * dy;
        for (size_t k = 0; k < headings_.size(); ++k) connectOneSliceLayer(k, Ly, ty, off, xL, xU, xM);
        connectMultiSlice();
    }

    /** generateVertices (:72-89) + connectOneSlice2D of layer k: edge tests of the layer in one K4 batch, host replay */
    void connectOneSliceLayer(size_t k, int Ly, const std::vector<double>& ty, const std::vector<long long>& off, const std::vector<double>& xL,
                              const std::vector<double>& xU, const std::vector<double>& xM) {
        auto& V = res_.graphStructure.vertex;
        const std::vector<double>& baseAngle = baseAngle_;
        const long long* o = off.data() + k * Ly;
        auto cnt = [&](int iy) { return static_cast<int>(o[iy + 1] - o[iy]); };
        std::vector<double> A, Bq;
        for (int iy = 0; iy + 1 < Ly; ++iy)
            for (int j1 = 0; j1 < cnt(iy); ++j1)
                for (int j2 = 0; j2 < cnt(iy + 1); ++j2) {
                    A.push_back(xM[o[iy] + j1]), A.push_back(ty[iy]);
                    Bq.push_back(xM[o[iy + 1] + j2]), Bq.push_back(ty[iy + 1]);
                }
        // bridgeVertex swaps index 2 = theta in 2D: vNew1 == v1, vNew2 == v2 as points, so it can never add a vertex (App. C8)
        const auto f = freeSpace_.testEdges(static_cast<int>(k), A, Bq, 2);
        numEdgeTests_ += static_cast<long>(f.size());
        const Index startId = V.size();
        std::vector<Index> plane;
        for (int iy = 0; iy < Ly; ++iy) {  // generateVertices (:72-89)
            plane.push_back(V.size());
            for (int j = 0; j < cnt(iy); ++j) V.push_back({xM[o[iy] + j], ty[iy], baseAngle[k]});
        }
        size_t pi = 0;
        for (int iy = 0; iy < Ly; ++iy) {  // connectOneSlice2D
            const Index n1 = plane[iy];
            for (int j1 = 0; j1 < cnt(iy); ++j1) {
                if (j1 != cnt(iy) - 1 && std::fabs(xU[o[iy] + j1] - xL[o[iy] + j1 + 1]) < 1e-6) addEdge(n1 + j1, n1 + j1 + 1);
                if (iy != Ly - 1)
                    for (int j2 = 0; j2 < cnt(iy + 1); ++j2)
                        if (f[pi++]) addEdge(n1 + j1, plane[iy + 1] + j2);
            }
        }
        vertexIdx_.push_back({startId, V.size()});
    }

    void refineExistRoadmap() override {
        vertexIdxAll_.push_back(vertexIdx_);
        vertexIdx_.clear();
        param_.numLineX *= 2;
        param_.numLineY *= 2;
        const auto& lim = param_.boundaryLimits;
        const int Ly = static_cast<int>(param_.numLineY);
        freeSpace_.resweep(lim, 1, param_.numLineY);
        std::vector<long long> off;
        std::vector<double> xL, xU, xM;
        freeSpace_.getFreeSegmentsAll(off, xL, xU, xM);
        const double dy = (lim[3] - lim[2]) / (static_cast<double>(Ly) - 1);
        std::vector<double> ty(Ly);
        for (int i = 0; i < Ly; ++i) ty[i] = lim[2] + static_cast<double>(i) * dy;
        const double th = 2.0 * std::fabs(lim[3] - lim[2]) / static_cast<double>(param_.numLineY);
        for (size_t k = 0; k < headings_.size(); ++k) {
            connectOneSliceLayer(k, Ly, ty, off, xL, xU, xM);
            const auto& V = res_.graphStructure.vertex;
            connectExistSliceReplay(
                k, [&](const std::vector<Coordinate>& a, const std::vector<Coordinate>& b) { return !(std::fabs(a[1] - b[1]) > th); },
                [&](const std::vector<std::pair<Index, Index>>& cand) {
                    std::vector<double> A, Bq;
                    for (const auto& c : cand)
                        for (int d = 0; d < 2; ++d) A.push_back(V[c.first][d]), Bq.push_back(V[c.second][d]);
                    return freeSpace_.testEdges(static_cast<int>(k), A, Bq, 2);
                });
            search();
            if (res_.solved) return;
        }
    }

    void connectMultiSlice() {  // HRM2D.cpp:91-156
        if (vertexIdx_.size() == 1) return;
        auto& V = res_.graphStructure.vertex;
        const auto& lim = param_.boundaryLimits;
        const size_t B = 1 + robot_.getNumLinks(), ns = param_.numSlice;
        std::vector<size_t> adj;
        std::vector<std::vector<Ellipse>> tfe;
        for (size_t i = 0; i < vertexIdx_.size(); ++i) {
            const size_t j = (i == ns - 1 && ns != 2) ? 0 : i + 1;
            adj.push_back(j);
            const double tha = headings_[i], thb = headings_[j];
            std::vector<Ellipse> pr;
            for (size_t b = 0; b < B; ++b) {  // computeTFE (:352-373)
                if (b == 0) {
                    const double a[2] = {robot_.getBase().getSemiAxis()[0], robot_.getBase().getSemiAxis()[1]};
                    pr.push_back(getTFE2D(a, tha, thb, param_.numPoint));
                } else {
                    const SE2Transform& g = robot_.getTF()[b - 1];
                    const double rl = std::atan2(g.R[1][0], g.R[0][0]), cl = std::cos(rl), sl = std::sin(rl);
                    auto ang = [&](double th) {
                        const double c = std::cos(th), s = std::sin(th);
                        const double m00 = c * cl + (-s) * sl, m10 = s * cl + c * sl;
                        return std::atan2(m10, m00);
                    };
                    const double a[2] = {robot_.getLinks()[b - 1].getSemiAxis()[0], robot_.getLinks()[b - 1].getSemiAxis()[1]};
                    pr.push_back(getTFE2D(a, ang(tha), ang(thb), param_.numPoint));
                }
            }
            tfe.push_back(pr);
        }
        freeSpace_.buildBridge(tfe, precision_);
        const double th = 2.0 * std::fabs(lim[3] - lim[2]) / static_cast<double>(param_.numLineY);
        const size_t steps = param_.numPoint;
        const double dt = 1.0 / (static_cast<double>(steps) - 1);
        // All layer pairs in one K5 launch (see HRM3D::connectMultiSlice): the interpolated heading of step s is the same for
        // every candidate of a pair (the vertices of a layer share their heading), so the rotated link translations
        // Rot(theta_s) * t_l of robotTF (:24-41) are evaluated once per pair; the kernel adds the interpolated position.
        struct PairCand {
            Index sCur = 0, eCur = 0, sAdj = 0, eAdj = 0;
            size_t first = 0, count = 0;
            std::vector<unsigned char> state;  // [(m0 - sCur) * (eAdj - sAdj) + (m1 - sAdj)]: 0 no candidate, 1 blocked, 2 free
        };
        const size_t P = vertexIdx_.size(), T = steps;
        std::vector<PairCand> pc(P);
        std::vector<int> pairOf;
        std::vector<double> v0s, v1s, w0(T), w1(T), linkOff(P * T * B * 2, 0.0);
        std::vector<std::pair<Index, Index>> cand;
        const auto& tf = robot_.getTF();
        for (size_t s = 0; s < T; ++s) w1[s] = static_cast<double>(s) * dt, w0[s] = 1.0 - static_cast<double>(s) * dt;  // :237-247
        for (size_t i = 0; i < P; ++i) {
            PairCand& c = pc[i];
            c.sCur = vertexIdx_[i].first, c.eCur = vertexIdx_[i].second;
            c.sAdj = vertexIdx_[adj[i]].first, c.eAdj = vertexIdx_[adj[i]].second;
            c.first = cand.size();
            if (c.eCur <= c.sCur || c.eAdj <= c.sAdj) continue;
            c.state.assign((c.eCur - c.sCur) * (c.eAdj - c.sAdj), 0);
            for (size_t s = 0; s < T; ++s) {
                const double theta = w0[s] * V[c.sCur][2] + w1[s] * V[c.sAdj][2];
                const double cs = std::cos(theta), sn = std::sin(theta);
                const double R[2][2] = {{cs, -sn}, {sn, cs}};
                for (size_t l = 0; l < tf.size(); ++l) {
                    linkOff[((i * T + s) * B + (l + 1)) * 2] = R[0][0] * tf[l].t[0] + R[0][1] * tf[l].t[1];
                    linkOff[((i * T + s) * B + (l + 1)) * 2 + 1] = R[1][0] * tf[l].t[0] + R[1][1] * tf[l].t[1];
                }
            }
            for (Index m0 = c.sCur; m0 < c.eCur; ++m0)
                for (Index m1 = c.sAdj; m1 < c.eAdj; ++m1) {
                    if (std::fabs(V[m0][1] - V[m1][1]) > th) continue;
                    c.state[(m0 - c.sCur) * (c.eAdj - c.sAdj) + (m1 - c.sAdj)] = 1;
                    cand.push_back({m0, m1});
                    v0s.insert(v0s.end(), V[m0].begin(), V[m0].begin() + 2);
                    v1s.insert(v1s.end(), V[m1].begin(), V[m1].begin() + 2);
                    pairOf.push_back(static_cast<int>(i));
                }
            c.count = cand.size() - c.first;
        }
        const auto ok = freeSpace_.testTransitions(pairOf, v0s, v1s, w0, w1, linkOff);
        for (size_t i = 0; i < P; ++i) {
            PairCand& c = pc[i];
            if (c.count == 0) continue;
            const Index w = c.eAdj - c.sAdj;
            for (size_t j = 0; j < c.count; ++j)
                if (ok[c.first + j]) c.state[(cand[c.first + j].first - c.sCur) * w + (cand[c.first + j].second - c.sAdj)] = 2;
            Index sAdj = c.sAdj;
            for (Index m0 = c.sCur; m0 < c.eCur; ++m0)
                for (Index m1 = sAdj; m1 < c.eAdj; ++m1)
                    if (c.state[(m0 - c.sCur) * w + (m1 - c.sAdj)] == 2) {
                        addEdge(m0, m1);
                        sAdj = m1;
                        break;
                    }
        }
    }

    std::vector<Index> getNearestNeighborsOnGraph(const std::vector<Coordinate>& vertex, Index k, double radius) override {  // :284-342
        const auto& V = res_.graphStructure.vertex;
        const auto& lim = param_.boundaryLimits;
        double minAngle = V[0][2], minD = std::fabs(vertex[2] - minAngle);
        for (const auto& v : V) {
            const double d = std::fabs(vertex[2] - v[2]);
            if (d < minD) minD = d, minAngle = v[2];
        }
        std::vector<double> angList;
        for (double h : headings_) {
            if (std::fabs(minAngle - h) < radius) angList.push_back(h);
            if (angList.size() >= k) break;
        }
        std::vector<Index> idx;
        for (double ac : angList) {
            Index idxSlice = 0;
            double best = vectorEuclidean(vertex, V[0]);
            for (size_t i = 0; i < V.size(); ++i) {
                const double d = vectorEuclidean(vertex, V[i]);
                if (d < best && std::fabs(V[i][2] - ac) < EPSILON) best = d, idxSlice = i;
            }
            if (std::abs(vertex[1] - V[idxSlice][1]) < radius * (lim[1] - lim[0]) / static_cast<double>(param_.numLineY)) idx.push_back(idxSlice);
        }
        return idx;
    }

    MultiBodyTree2D robot_;
    FreeSpaceGPU2D freeSpace_;
    std::vector<double> headings_;
    std::vector<double> baseAngle_;  // base angle of every layer as the vertices carry it
};

}  // namespace planners
}  // namespace hrm
