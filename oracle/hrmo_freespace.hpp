// ORACLE — TEST INFRASTRUCTURE ONLY. Nothing under hrm_b200/ may include, link or call this.
//
// CPU restatement of the free-space layer of ChirikjianLab/hrm (the drop-in seam):
// Interval algebra, MultiBodyTree{2D,3D}::minkSum, FreeSpaceComputator / FreeSpace2D / FreeSpace3D.
// Each function cites the reference file:line it follows.
#pragma once
#include "hrmo_geometry.hpp"

#include <cfloat>

namespace hrmo {

// ---------------------------------------------------------------------------------------------
// Interval (src/datastructure/Interval.cpp)
// ---------------------------------------------------------------------------------------------
struct Interval {
    double s = 0, e = 0;
};

// Interval.cpp:10-32
inline std::vector<Interval> unions(const std::vector<Interval>& ins) {
    if (ins.empty()) return {};
    auto c = ins;
    std::vector<Interval> res;
    std::sort(c.begin(), c.end(), [](Interval a, Interval b) { return a.s < b.s; });
    res.push_back(c.at(0));
    for (const auto el : c) {
        if (res.back().e < el.s)
            res.push_back(el);
        else
            res.back().e = std::max(res.back().e, el.e);
    }
    return res;
}

// Interval.cpp:34-58
inline std::vector<Interval> intersects(const std::vector<Interval>& ins) {
    if (ins.empty()) return {};
    std::vector<Interval> res;
    auto c = ins;
    Interval buff;
    std::sort(c.begin(), c.end(), [](Interval a, Interval b) { return a.s < b.s; });
    buff.s = c.back().s;
    std::sort(c.begin(), c.end(), [](Interval a, Interval b) { return a.e < b.e; });
    buff.e = c.front().e;
    if (buff.s > buff.e) return {};
    res.push_back(buff);
    return res;
}

// Interval.cpp:60-100
inline std::vector<Interval> complements(const std::vector<Interval>& outer, const std::vector<Interval>& inner) {
    if (outer.empty()) return {};
    if (inner.empty()) return outer;
    auto ic = inner;
    std::vector<Interval> res, comp, intBuff, intsect;
    std::sort(ic.begin(), ic.end(), [](Interval a, Interval b) { return a.s < b.s; });
    comp.push_back({-std::numeric_limits<double>::max(), ic.at(0).s});
    for (size_t i = 0; i < ic.size() - 1; ++i) comp.push_back({ic.at(i).e, ic.at(i + 1).s});
    comp.push_back({ic.back().e, std::numeric_limits<double>::max()});
    for (auto curr : comp) {
        intBuff.push_back(curr);
        intBuff.push_back(outer.at(0));
        intsect = intersects(intBuff);
        if (!intsect.empty()) res.push_back(intsect.at(0));
        intBuff.clear();
    }
    return res;
}

// ---------------------------------------------------------------------------------------------
// MultiBodyTree3D / 2D (src/datastructure/MultiBodyTree3D.cpp, MultiBodyTree2D.cpp)
// ---------------------------------------------------------------------------------------------
struct Tf3 {
    Mat3 R;
    double t[3];
};
inline Tf3 tf3_mul(const Tf3& a, const Tf3& b) {
    Tf3 r;
    r.R = a.R * b.R;
    const Vec3 v = a.R * Vec3{b.t[0], b.t[1], b.t[2]};
    r.t[0] = v.x + a.t[0];
    r.t[1] = v.y + a.t[1];
    r.t[2] = v.z + a.t[2];
    return r;
}

struct MultiBodyTree3D {
    SuperQuadrics base;
    std::vector<SuperQuadrics> link;
    std::vector<Tf3> tf;  // link transforms relative to base

    // MultiBodyTree3D.cpp:21-32
    void addBody(const SuperQuadrics& l) {
        link.push_back(l);
        Tf3 g;
        g.R = quat_to_rot(l.quat);
        for (int i = 0; i < 3; ++i) g.t[i] = l.position[i];
        tf.push_back(g);
    }
    size_t getNumLinks() const { return link.size(); }

    // MultiBodyTree3D.cpp:34-53
    void robotTF(const Tf3& g) {
        for (int i = 0; i < 3; ++i) base.position[i] = g.t[i];
        base.quat = rot_to_quat(g.R);
        for (size_t i = 0; i < link.size(); ++i) {
            const Tf3 gl = tf3_mul(g, tf.at(i));
            for (int d = 0; d < 3; ++d) link[i].position[d] = gl.t[d];
            link[i].quat = rot_to_quat(gl.R);
        }
    }

    // MultiBodyTree3D.cpp:91-105
    std::vector<Points> minkSum(const SuperQuadrics& s1, int k) const {
        std::vector<Points> mink;
        mink.push_back(s1.getMinkSum3D(base.semiAxis, base.quat, k));
        for (size_t i = 0; i < link.size(); ++i) {
            Points p = s1.getMinkSum3D(link[i].semiAxis, link[i].quat, k);
            for (int kk = 0; kk < p.num; ++kk)
                for (int d = 0; d < 3; ++d) p.at(d, kk) = p.at(d, kk) - link[i].position[d];
            mink.push_back(std::move(p));
        }
        return mink;
    }
};

struct Tf2 {
    double R[2][2];
    double t[2];
};

struct MultiBodyTree2D {
    SuperEllipse base;
    std::vector<SuperEllipse> link;
    std::vector<Tf2> tf;

    // MultiBodyTree2D.cpp:10-22
    void addBody(const SuperEllipse& l) {
        link.push_back(l);
        Tf2 g;
        const double c = std::cos(l.angle), s = std::sin(l.angle);
        g.R[0][0] = c, g.R[0][1] = -s, g.R[1][0] = s, g.R[1][1] = c;
        g.t[0] = l.position[0], g.t[1] = l.position[1];
        tf.push_back(g);
    }
    size_t getNumLinks() const { return link.size(); }

    // MultiBodyTree2D.cpp:24-41.  Rotation2Dd(Matrix2d) -> angle = atan2(m10, m00).
    void robotTF(const Tf2& g) {
        base.position[0] = g.t[0];
        base.position[1] = g.t[1];
        base.angle = std::atan2(g.R[1][0], g.R[0][0]);
        for (size_t i = 0; i < link.size(); ++i) {
            double R[2][2];
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 2; ++b) R[a][b] = g.R[a][0] * tf[i].R[0][b] + g.R[a][1] * tf[i].R[1][b];
            link[i].position[0] = (g.R[0][0] * tf[i].t[0] + g.R[0][1] * tf[i].t[1]) + g.t[0];
            link[i].position[1] = (g.R[1][0] * tf[i].t[0] + g.R[1][1] * tf[i].t[1]) + g.t[1];
            link[i].angle = std::atan2(R[1][0], R[0][0]);
        }
    }

    // Per-body (angle, offset) exactly as MultiBodyTree2D.cpp:43-65 derives them from the base angle.
    void bodyPose(size_t i, double& ang, double off[2]) const {
        if (i == 0) {
            ang = base.angle;
            off[0] = off[1] = 0.0;
            return;
        }
        const Tf2& g = tf.at(i - 1);
        const double c = std::cos(base.angle), s = std::sin(base.angle);
        // rotLink = rotBase * tf.topLeftCorner(2,2) (Rotation2D * Matrix2d -> matrix product; angle via atan2)
        const double m00 = c * g.R[0][0] + (-s) * g.R[1][0];
        const double m10 = s * g.R[0][0] + c * g.R[1][0];
        ang = std::atan2(m10, m00);
        off[0] = c * g.t[0] + (-s) * g.t[1];
        off[1] = s * g.t[0] + c * g.t[1];
    }

    // MultiBodyTree2D.cpp:43-65
    std::vector<Points> minkSum(const SuperEllipse& s1, int k) const {
        std::vector<Points> mink;
        mink.push_back(s1.getMinkSum2D(base.semiAxis, base.angle, k));
        for (size_t i = 0; i < link.size(); ++i) {
            double ang, off[2];
            bodyPose(i + 1, ang, off);
            Points p = s1.getMinkSum2D(link[i].semiAxis, ang, k);
            for (int kk = 0; kk < p.num; ++kk) {
                p.at(0, kk) = p.at(0, kk) - off[0];
                p.at(1, kk) = p.at(1, kk) - off[1];
            }
            mink.push_back(std::move(p));
        }
        return mink;
    }
};

// ---------------------------------------------------------------------------------------------
// FreeSpaceComputator (include/hrm/datastructure/FreeSpace.h, FreeSpace-inl.h)
// ---------------------------------------------------------------------------------------------
struct BoundaryInfo {
    std::vector<Points> arena, obstacle;
};
struct IntersectionInterval {
    std::vector<std::vector<double>> arenaLow, arenaUpp, obstacleLow, obstacleUpp;
};
struct FreeSegment2D {
    std::vector<double> ty;
    std::vector<std::vector<double>> xL, xU, xM;
};
struct FreeSegment3D {
    std::vector<double> tx;
    std::vector<FreeSegment2D> freeSegmentYZ;
};
struct BoundaryMesh {
    std::vector<MeshMatrix> arena, obstacle;
};

template <class RobotType, class ObjectType>
struct FreeSpaceComputator {
    RobotType robot_;                        // BY VALUE (FreeSpace.h:106; SURVEY.md finding 1)
    const std::vector<ObjectType>* arena_;   // by reference (FreeSpace.h:109-112)
    const std::vector<ObjectType>* obstacle_;
    BoundaryInfo cSpaceBoundary_;
    IntersectionInterval intersect_;
    FreeSegment2D segment_;
    double lowBound_ = 0, upBound_ = 0;
    long singleHitCount = 0;  // SURVEY.md finding 7: defined here as "no interval", counted

    FreeSpaceComputator(const RobotType& robot, const std::vector<ObjectType>& arena,
                        const std::vector<ObjectType>& obstacle)
        : robot_(robot), arena_(&arena), obstacle_(&obstacle) {}
    virtual ~FreeSpaceComputator() = default;

    // FreeSpace-inl.h:20-39 (divergence C6: callable again when the line count changes)
    void setup(unsigned numLine, double lowBound, double upBound) {
        lowBound_ = lowBound;
        upBound_ = upBound;
        const size_t numArena = (1 + robot_.getNumLinks()) * arena_->size();
        const size_t numObstacle = (1 + robot_.getNumLinks()) * obstacle_->size();
        intersect_.arenaLow.assign(numLine, std::vector<double>(numArena, lowBound));
        intersect_.arenaUpp.assign(numLine, std::vector<double>(numArena, upBound));
        intersect_.obstacleLow.assign(numLine, std::vector<double>(numObstacle, NAN));
        intersect_.obstacleUpp.assign(numLine, std::vector<double>(numObstacle, NAN));
    }
    void setRobot(const RobotType& r) { robot_ = r; }

    // FreeSpace-inl.h:42-60
    void computeCSpaceBoundary() {
        cSpaceBoundary_.arena.clear();
        cSpaceBoundary_.obstacle.clear();
        for (const auto& part : *arena_)
            for (auto& b : robot_.minkSum(part, -1)) cSpaceBoundary_.arena.push_back(b);
        for (const auto& part : *obstacle_)
            for (auto& b : robot_.minkSum(part, +1)) cSpaceBoundary_.obstacle.push_back(b);
    }

    // FreeSpace-inl.h:101-129
    std::vector<Interval> computeSweepLineFreeSegment(size_t lineIdx) {
        std::vector<Interval> obsSegment, arenaSegment;
        for (size_t j = 0; j < intersect_.arenaLow.at(lineIdx).size(); ++j)
            if (!std::isnan(intersect_.arenaLow[lineIdx][j]) && !std::isnan(intersect_.arenaUpp[lineIdx][j]))
                arenaSegment.push_back({intersect_.arenaLow[lineIdx][j], intersect_.arenaUpp[lineIdx][j]});
        for (size_t j = 0; j < intersect_.obstacleLow.at(lineIdx).size(); ++j)
            if (!std::isnan(intersect_.obstacleLow[lineIdx][j]) && !std::isnan(intersect_.obstacleUpp[lineIdx][j]))
                obsSegment.push_back({intersect_.obstacleLow[lineIdx][j], intersect_.obstacleUpp[lineIdx][j]});
        return complements(intersects(arenaSegment), unions(obsSegment));
    }

    // FreeSpace-inl.h:63-97
    void computeFreeSegment(const std::vector<double>& ty) {
        segment_.xL.clear();
        segment_.xU.clear();
        segment_.xM.clear();
        segment_.ty = ty;
        for (size_t i = 0; i < ty.size(); ++i) {
            const std::vector<Interval> iv = computeSweepLineFreeSegment(i);
            std::vector<double> xL, xU, xM;
            for (size_t j = 0; j < iv.size(); ++j) {
                xL.push_back(iv[j].s);
                xU.push_back(iv[j].e);
                xM.push_back((iv[j].s + iv[j].e) / 2.0);
            }
            segment_.xL.push_back(xL);
            segment_.xU.push_back(xU);
            segment_.xM.push_back(xM);
        }
        enhanceFreeSegment();
    }

    // FreeSpace-inl.h:132-174
    void enhanceFreeSegment() {
        FreeSegment2D en = segment_;
        if (segment_.ty.empty()) return;
        for (size_t i = 0; i < segment_.ty.size() - 1; ++i) {
            for (size_t j1 = 0; j1 < segment_.xM[i].size(); ++j1) {
                for (size_t j2 = 0; j2 < segment_.xM[i + 1].size(); ++j2) {
                    if (en.xM[i][j1] < en.xL[i + 1][j2] && en.xU[i][j1] >= en.xL[i + 1][j2]) {
                        en.xU[i].push_back(en.xL[i + 1][j2]);
                        en.xL[i].push_back(en.xL[i + 1][j2]);
                        en.xM[i].push_back(en.xL[i + 1][j2]);
                    } else if (en.xM[i][j1] > en.xU[i + 1][j2] && en.xL[i][j1] <= en.xU[i + 1][j2]) {
                        en.xU[i].push_back(en.xU[i + 1][j2]);
                        en.xL[i].push_back(en.xU[i + 1][j2]);
                        en.xM[i].push_back(en.xU[i + 1][j2]);
                    }
                    if (en.xM[i + 1][j2] < en.xL[i][j1] && en.xU[i + 1][j2] >= en.xL[i][j1]) {
                        en.xU[i + 1].push_back(en.xL[i][j1]);
                        en.xL[i + 1].push_back(en.xL[i][j1]);
                        en.xM[i + 1].push_back(en.xL[i][j1]);
                    } else if (en.xM[i + 1][j2] > en.xU[i][j1] && en.xL[i + 1][j2] <= en.xU[i][j1]) {
                        en.xU[i + 1].push_back(en.xU[i][j1]);
                        en.xL[i + 1].push_back(en.xU[i][j1]);
                        en.xM[i + 1].push_back(en.xU[i][j1]);
                    }
                }
            }
            std::sort(en.xL[i].begin(), en.xL[i].end(), [](double a, double b) { return a < b; });
            std::sort(en.xU[i].begin(), en.xU[i].end(), [](double a, double b) { return a < b; });
            std::sort(en.xM[i].begin(), en.xM[i].end(), [](double a, double b) { return a < b; });
        }
        segment_ = en;
    }
};

// src/datastructure/FreeSpace3D.cpp
struct FreeSpace3D : FreeSpaceComputator<MultiBodyTree3D, SuperQuadrics> {
    BoundaryMesh cSpaceBoundaryMesh_;
    using FreeSpaceComputator::FreeSpaceComputator;

    // FreeSpace3D.cpp:14-27
    void computeCSpaceBoundaryMesh(const BoundaryInfo& bound) {
        cSpaceBoundaryMesh_.arena.resize(bound.arena.size());
        cSpaceBoundaryMesh_.obstacle.resize(bound.obstacle.size());
        for (size_t i = 0; i < bound.arena.size(); ++i)
            cSpaceBoundaryMesh_.arena[i] = getMeshFromParamSurface(bound.arena[i], arena_->at(0).getNumParam());
        for (size_t i = 0; i < bound.obstacle.size(); ++i)
            cSpaceBoundaryMesh_.obstacle[i] = getMeshFromParamSurface(bound.obstacle[i], obstacle_->at(0).getNumParam());
    }

    // FreeSpace3D.cpp:29-68.  tx = tLine[0].back(), ty = tLine[1].
    void computeIntersectionInterval(double tx, const std::vector<double>& ty) {
        for (size_t i = 0; i < ty.size(); ++i) {
            const Line3D lineZ = {{tx, ty[i], 0, 0, 0, 1}};
            for (size_t j = 0; j < cSpaceBoundary_.arena.size(); ++j) {
                const auto h = intersectVerticalLineMesh3D(lineZ, cSpaceBoundaryMesh_.arena[j]);
                if (h.size() == 1) ++singleHitCount;
                if (h.size() < 2) {  // reference: empty(); the 1-hit case reads h[1] out of bounds (finding 7)
                    intersect_.arenaLow.at(i).at(j) = lowBound_;
                    intersect_.arenaUpp.at(i).at(j) = upBound_;
                } else {
                    intersect_.arenaLow[i][j] = std::fmin(lowBound_, std::fmin(h[0].z, h[1].z));
                    intersect_.arenaUpp[i][j] = std::fmax(upBound_, std::fmax(h[0].z, h[1].z));
                }
            }
            for (size_t j = 0; j < cSpaceBoundary_.obstacle.size(); ++j) {
                const auto h = intersectVerticalLineMesh3D(lineZ, cSpaceBoundaryMesh_.obstacle[j]);
                if (h.size() == 1) ++singleHitCount;
                if (h.size() < 2) {
                    intersect_.obstacleLow.at(i).at(j) = NAN;
                    intersect_.obstacleUpp.at(i).at(j) = NAN;
                } else {
                    intersect_.obstacleLow[i][j] = std::fmin(h[0].z, h[1].z);
                    intersect_.obstacleUpp[i][j] = std::fmax(h[0].z, h[1].z);
                }
            }
        }
    }
};

// src/datastructure/FreeSpace2D.cpp
struct FreeSpace2D : FreeSpaceComputator<MultiBodyTree2D, SuperEllipse> {
    using FreeSpaceComputator::FreeSpaceComputator;

    // FreeSpace2D.cpp:14-50
    void computeIntersectionInterval(const std::vector<double>& ty) {
        for (size_t i = 0; i < ty.size(); ++i) {
            for (size_t j = 0; j < cSpaceBoundary_.arena.size(); ++j) {
                const auto h = intersectHorizontalLinePolygon2D(ty[i], cSpaceBoundary_.arena[j]);
                if (h.size() == 1) ++singleHitCount;
                if (h.size() < 2) {
                    intersect_.arenaLow.at(i).at(j) = lowBound_;
                    intersect_.arenaUpp.at(i).at(j) = upBound_;
                } else {
                    intersect_.arenaLow[i][j] = std::fmin(lowBound_, std::fmin(h[0], h[1]));
                    intersect_.arenaUpp[i][j] = std::fmax(upBound_, std::fmax(h[0], h[1]));
                }
            }
            for (size_t j = 0; j < cSpaceBoundary_.obstacle.size(); ++j) {
                const auto h = intersectHorizontalLinePolygon2D(ty[i], cSpaceBoundary_.obstacle[j]);
                if (h.size() == 1) ++singleHitCount;
                if (h.size() < 2) {
                    intersect_.obstacleLow.at(i).at(j) = NAN;
                    intersect_.obstacleUpp.at(i).at(j) = NAN;
                } else {
                    intersect_.obstacleLow[i][j] = std::fmin(h[0], h[1]);
                    intersect_.obstacleUpp[i][j] = std::fmax(h[0], h[1]);
                }
            }
        }
    }
};

// src/planners/HRM3D.cpp:63-91 — sweep-line grids + per-plane free segments for one C-layer.
inline FreeSegment3D sweepLineProcess3D(FreeSpace3D& fs, const double lim[6], size_t numLineX, size_t numLineY) {
    std::vector<double> ty(numLineY);
    const double dx = (lim[1] - lim[0]) / static_cast<double>(numLineX - 1);
    const double dy = (lim[3] - lim[2]) / static_cast<double>(numLineY - 1);
    for (size_t i = 0; i < numLineY; ++i) ty[i] = lim[2] + static_cast<double>(i) * dy;
    FreeSegment3D out;
    for (size_t i = 0; i < numLineX; ++i) {
        out.tx.push_back(lim[0] + static_cast<double>(i) * dx);
        fs.computeIntersectionInterval(out.tx.back(), ty);
        fs.computeFreeSegment(ty);
        out.freeSegmentYZ.push_back(fs.segment_);
    }
    return out;
}

// src/planners/HRM2D.cpp:53-70
inline FreeSegment2D sweepLineProcess2D(FreeSpace2D& fs, const double lim[4], size_t numLineY) {
    std::vector<double> ty(numLineY);
    const double dy = (lim[3] - lim[2]) / (static_cast<double>(numLineY) - 1);
    for (size_t i = 0; i < numLineY; ++i) ty[i] = lim[2] + static_cast<double>(i) * dy;
    fs.computeIntersectionInterval(ty);
    fs.computeFreeSegment(ty);
    return fs.segment_;
}

}  // namespace hrmo
