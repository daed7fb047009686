// Scene / configuration input and result output of the host layer: the reference's CSV formats on both sides of
// the GPU path.  Same function names, argument meaning and file layout as
//   src/util/Parse2dCsvFile.cpp:3-43                       parse2DCsvFile
//   src/test/util/ParsePlanningSettings.cpp:6-263          loadVectorGeometry, loadRobotMultiBody2D/3D,
//                                                          loadPreDefinedQuaternions, defineParameters,
//                                                          parsePlanningConfig, parseGeometricModel, parseStartGoalConfig
//   include/hrm/test/util/ParsePlanningSettings.h:76-171   computeObstacleMinSize, PlannerSetting<T>
//   src/test/util/DisplayPlanningData.cpp:8-91             display*/store* writers
//   include/hrm/test/util/DisplayPlanningData-inl.h:12-66  storeRoutines
// The reference bakes RESOURCES_PATH / CONFIG_PATH / SOLUTION_DETAILS_PATH in at configure time (config.h.in); here
// they are run-time arguments (hrm::Paths), defaulting to the environment variables HRM_RESOURCES_PATH,
// HRM_CONFIG_PATH, HRM_RESULT_PATH.
#pragma once
#include "hrm_host.hpp"

#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

namespace hrm {

struct Paths {
    std::string resources, config, details, benchmark;
    static Paths fromEnv() {
        auto env = [](const char* k, const char* dflt) {
            const char* v = std::getenv(k);
            return std::string(v ? v : dflt);
        };
        Paths p;
        p.resources = env("HRM_RESOURCES_PATH", "resources");
        p.config = env("HRM_CONFIG_PATH", "config");
        const std::string result = env("HRM_RESULT_PATH", "result");
        p.details = result + "/details";
        p.benchmark = result + "/benchmark";
        return p;
    }
};

// Cells go through std::stof — single precision — exactly like the reference; a cell that does not start with a
// number is reported and skipped, a line starting with '#' is a comment, an empty line yields an empty record.
inline std::vector<std::vector<double>> parse2DCsvFile(const std::string& filename) {
    std::vector<std::vector<double>> data;
    std::ifstream in(filename);
    int lineNo = 0;
    std::string row;
    while (in) {
        ++lineNo;
        if (!std::getline(in, row)) break;
        if (!row.empty() && row[0] == '#') continue;
        std::istringstream cells(row);
        std::vector<double> record;
        std::string cell;
        while (std::getline(cells, cell, ',')) {
            try {
                record.emplace_back(std::stof(cell));
            } catch (const std::invalid_argument&) {
                std::cout << "NaN found in file " << filename << " line " << lineNo << std::endl;
            }
        }
        data.push_back(record);
    }
    if (!in.eof()) {
        std::cerr << "Could not read file " << filename << "\n";
        throw std::invalid_argument("File not found.");
    }
    return data;
}

// Quaterniond(AngleAxisd(angle, axis.normalized()))
inline Quaternion angleAxisToQuaternion(double angle, double x, double y, double z) {
    const double n2 = x * x + y * y + z * z;
    if (n2 > 0.0) {
        const double n = std::sqrt(n2);
        x = x / n, y = y / n, z = z / n;
    }
    const double ha = 0.5 * angle, s = std::sin(ha);
    return {std::cos(ha), s * x, s * y, s * z};
}

// resources/*.csv (angle-axis) -> config/*.csv (quaternion); values are written with the default ostream format
// (6 significant digits), which is part of what the planner later sees.
inline void parseGeometricModel(const std::string& inputFilename, const std::string& outputFilename) {
    const auto objects = parse2DCsvFile(inputFilename);
    std::ofstream out(outputFilename);
    if (!out) throw std::invalid_argument("cannot write " + outputFilename);
    for (const auto& object : objects) {
        if (object.size() == 12) {
            const Quaternion q = angleAxisToQuaternion(object[11], object[8], object[9], object[10]);
            for (int i = 0; i < 8; ++i) out << object[i] << ',';
            out << q[0] << ',' << q[1] << ',' << q[2] << ',' << q[3] << "\n";
        } else {
            if (object.empty()) throw std::out_of_range("parseGeometricModel: empty record in " + inputFilename);
            for (size_t i = 0; i + 1 < object.size(); ++i) out << object[i] << ',';
            out << object.back() << "\n";
        }
    }
}

inline void parseStartGoalConfig(const std::string& robotType, const std::string& dim, const std::string& inputFilename,
                                 const std::string& outputFilename) {
    const auto configs = parse2DCsvFile(inputFilename);
    std::ofstream out(outputFilename);
    if (!out) throw std::invalid_argument("cannot write " + outputFilename);
    for (auto config : configs) {
        if (dim == "2D") {
            out << config.at(0) << ',' << config.at(1) << ',' << config.at(2) << "\n";
        } else if (dim == "3D") {
            size_t numDOF = 6;
            const Quaternion q = angleAxisToQuaternion(config.at(6), config.at(3), config.at(4), config.at(5));
            config.at(3) = q[0], config.at(4) = q[1], config.at(5) = q[2], config.at(6) = q[3];
            if (robotType == "snake") numDOF = 9;
            else if (robotType == "tree") numDOF = 12;
            for (size_t i = 0; i < numDOF; ++i) out << config.at(i) << ',';
            out << config.at(numDOF) << "\n";
        } else {
            std::cerr << "Only '2D' and '3D' are supported." << std::endl;
        }
    }
}

inline void parsePlanningConfig(const std::string& objectType, const std::string& envType, const std::string& robotType,
                                const std::string& dim, const Paths& paths = Paths::fromEnv()) {
    const std::string env = paths.resources + "/" + dim + "/env_" + objectType + "_" + envType + "_" + dim;
    parseGeometricModel(env + "_arena.csv", paths.config + "/arena_config_" + dim + ".csv");
    parseGeometricModel(env + "_obstacle.csv", paths.config + "/obstacle_config_" + dim + ".csv");
    parseGeometricModel(paths.resources + "/" + dim + "/robot_" + robotType + "_" + dim + ".csv", paths.config + "/robot_config_" + dim + ".csv");
    parseStartGoalConfig(robotType, dim, paths.resources + "/" + dim + "/setting_" + objectType + "_" + envType + "_" + dim + ".csv",
                         paths.config + "/end_points_" + dim + ".csv");
}

inline void loadVectorGeometry(const std::vector<std::vector<double>>& objectConfig, int numCurvePoint, std::vector<SuperEllipse>& object) {
    object.clear();
    for (const auto& c : objectConfig) object.emplace_back(std::vector<double>{c.at(0), c.at(1)}, c.at(2), std::vector<double>{c.at(3), c.at(4)}, c.at(5), Index(numCurvePoint));
}
inline void loadVectorGeometry(const std::vector<std::vector<double>>& objectConfig, int numSurfPointParam, std::vector<SuperQuadrics>& object) {
    object.clear();
    for (const auto& c : objectConfig)
        object.emplace_back(std::vector<double>{c.at(0), c.at(1), c.at(2)}, std::vector<double>{c.at(3), c.at(4)},
                            std::vector<double>{c.at(5), c.at(6), c.at(7)}, Quaternion{c.at(8), c.at(9), c.at(10), c.at(11)}, Index(numSurfPointParam));
}
template <class ObjectType>
inline void loadVectorGeometry(const std::string& configFilename, int numPointParam, std::vector<ObjectType>& object) {
    loadVectorGeometry(parse2DCsvFile(configFilename), numPointParam, object);
}

// rows are x,y,z,w in the SO3_sequence files; "0" means "sample at run time"
inline void loadPreDefinedQuaternions(const std::string& quaternionFilename, SuperQuadrics& robotBase) {
    if (quaternionFilename == "0") {
        std::cout << "Will generate uniform random rotations from SO(3)" << std::endl;
        return;
    }
    std::vector<Quaternion> qSample;
    for (const auto& s : parse2DCsvFile(quaternionFilename)) qSample.push_back({s.at(3), s.at(0), s.at(1), s.at(2)});
    robotBase.setQuatSamples(qSample);
}

inline MultiBodyTree2D loadRobotMultiBody2D(const std::string& pathPrefix, int numCurvePoint) {
    std::vector<SuperEllipse> parts;
    loadVectorGeometry(pathPrefix + "robot_config_2D.csv", numCurvePoint, parts);
    MultiBodyTree2D robot(parts.at(0));
    for (size_t i = 1; i < parts.size(); ++i) robot.addBody(parts[i]);
    return robot;
}
inline MultiBodyTree3D loadRobotMultiBody3D(const std::string& pathPrefix, const std::string& quaternionFilename, int numSurfPointParam) {
    std::vector<SuperQuadrics> parts;
    loadVectorGeometry(pathPrefix + "robot_config_3D.csv", numSurfPointParam, parts);
    loadPreDefinedQuaternions(quaternionFilename, parts.at(0));
    MultiBodyTree3D robot(parts.at(0));
    for (size_t i = 1; i < parts.size(); ++i) robot.addBody(parts[i]);
    return robot;
}

template <typename ObjectType>
double computeObstacleMinSize(const std::vector<ObjectType>& obstacles) {
    double minSize = obstacles.at(0).getSemiAxis().at(0);
    for (const auto& obs : obstacles)
        for (double s : obs.getSemiAxis())
            if (s < minSize) minSize = s;
    return minSize;
}

template <class ObjectType>
class PlannerSetting {
  public:
    explicit PlannerSetting(int numPointParam) : dim_(std::is_same<ObjectType, SuperEllipse>::value ? "2D" : "3D"), numPointParam_(numPointParam) {}
    void setArena(std::vector<ObjectType> arena) { arena_ = std::move(arena); }
    void setObstacle(std::vector<ObjectType> obstacle) { obstacle_ = std::move(obstacle); }
    void setEndPoints(std::vector<std::vector<double>> endPoints) { endPoints_ = std::move(endPoints); }
    const std::vector<ObjectType>& getArena() const { return arena_; }
    const std::vector<ObjectType>& getObstacle() const { return obstacle_; }
    const std::vector<std::vector<double>>& getEndPoints() const { return endPoints_; }
    int getNumParam() const { return numPointParam_; }
    void loadEnvironment(const std::string& pathPrefix) {
        loadVectorGeometry(pathPrefix + "arena_config_" + dim_ + ".csv", numPointParam_, arena_);
        loadVectorGeometry(pathPrefix + "obstacle_config_" + dim_ + ".csv", numPointParam_, obstacle_);
        endPoints_ = parse2DCsvFile(pathPrefix + "end_points_" + dim_ + ".csv");
    }

  protected:
    std::string dim_;
    std::vector<ObjectType> arena_, obstacle_;
    std::vector<std::vector<double>> endPoints_;
    const int numPointParam_;
};
using PlannerSetting2D = PlannerSetting<SuperEllipse>;
using PlannerSetting3D = PlannerSetting<SuperQuadrics>;

// 2D: margin 1.5 x the base's first semi-axis on both axes; numLineY = int(bound_y / min obstacle semi-axis)
inline void defineParameters(const MultiBodyTree2D& robot, const PlannerSetting2D& env2D, PlannerParameter& param) {
    const double f = 1.5;
    const auto& arena = env2D.getArena().at(0);
    const std::vector<Coordinate> bound = {arena.getSemiAxis().at(0) - f * robot.getBase().getSemiAxis().at(0),
                                           arena.getSemiAxis().at(1) - f * robot.getBase().getSemiAxis().at(0)};
    param.boundaryLimits = {arena.getPosition().at(0) - bound[0], arena.getPosition().at(0) + bound[0], arena.getPosition().at(1) - bound[1],
                            arena.getPosition().at(1) + bound[1]};
    if (param.numLineY == 0) param.numLineY = static_cast<Index>(static_cast<int>(bound[1] / computeObstacleMinSize(env2D.getObstacle())));
}
// 3D: margin 1.2 x; numSlice from the quaternion samples; line counts = floor(bound / min obstacle semi-axis)
inline void defineParameters(const MultiBodyTree3D& robot, const PlannerSetting3D& env3D, PlannerParameter& param) {
    const double f = 1.2;
    const auto& arena = env3D.getArena().at(0);
    const double r = robot.getBase().getSemiAxis().at(0);
    const std::vector<double> bound = {arena.getSemiAxis().at(0) - f * r, arena.getSemiAxis().at(1) - f * r, arena.getSemiAxis().at(2) - f * r};
    param.boundaryLimits = {arena.getPosition().at(0) - bound[0], arena.getPosition().at(0) + bound[0], arena.getPosition().at(1) - bound[1],
                            arena.getPosition().at(1) + bound[1], arena.getPosition().at(2) - bound[2], arena.getPosition().at(2) + bound[2]};
    param.numSlice = robot.getBase().getQuatSamples().size();
    if (param.numLineX == 0 || param.numLineY == 0) {
        const double minSize = computeObstacleMinSize(env3D.getObstacle());
        param.numLineX = static_cast<Index>(std::floor(bound[0] / minSize));
        param.numLineY = static_cast<Index>(std::floor(bound[1] / minSize));
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// result output
// ---------------------------------------------------------------------------------------------------------------------
inline void displayPlanningTimeInfo(const Time& time) {
    if (time.buildTime > 0) std::cout << "Roadmap build time: " << time.buildTime << "s" << std::endl;
    if (time.searchTime > 0) std::cout << "Path search time: " << time.searchTime << "s" << std::endl;
    std::cout << "Total Planning Time: " << time.totalTime << 's' << std::endl;
}
inline void displayGraphInfo(const Graph& graph) {
    std::cout << "Number of valid configurations: " << graph.vertex.size() << std::endl;
    std::cout << "Number of valid edges: " << graph.edge.size() << std::endl;
}
inline void displayPathInfo(const SolutionPathInfo& path) {
    std::cout << "Path length: " << path.solvedPath.size() << std::endl;
    std::cout << "Path cost: " << path.cost << std::endl;
}

inline void writeRows(const std::string& filename, const std::vector<std::vector<Coordinate>>& rows) {
    std::ofstream out(filename);
    if (!out) throw std::invalid_argument("cannot write " + filename);
    for (const auto& row : rows) {
        for (double v : row) out << v << ',';  // trailing comma, like the reference
        out << "\n";
    }
}
inline void storeGraphInfo(const Graph& graph, const std::string& suffix, const Paths& paths = Paths::fromEnv()) {
    writeRows(paths.details + "/vertex_" + suffix + ".csv", graph.vertex);
    std::ofstream out(paths.details + "/edge_" + suffix + ".csv");
    if (!out) throw std::invalid_argument("cannot write edge file under " + paths.details);
    for (const auto& e : graph.edge) out << e.first << ',' << e.second << "\n";
}
inline void storePathInfo(const SolutionPathInfo& path, const std::string& suffix, const Paths& paths = Paths::fromEnv()) {
    {
        std::ofstream out(paths.details + "/path_id_" + suffix + ".csv");
        if (!out) throw std::invalid_argument("cannot write path file under " + paths.details);
        for (Index id : path.PathId) out << id << ',';
    }
    writeRows(paths.details + "/solution_path_" + suffix + ".csv", path.solvedPath);
    writeRows(paths.details + "/interpolated_path_" + suffix + ".csv", path.interpolatedPath);
}

// A boundary is written the way Eigen streams a 3 x M (2 x M) matrix: one text row per coordinate, entries separated
// by a single space and right-aligned to the widest entry of that matrix.
inline void writeBoundary(std::ostream& out, const std::vector<double>& pts, int dim) {
    const size_t M = pts.size() / static_cast<size_t>(dim);
    std::vector<std::string> cell(pts.size());
    size_t width = 0;
    for (size_t i = 0; i < pts.size(); ++i) {
        std::ostringstream s;
        s << pts[i];
        cell[i] = s.str();
        width = std::max(width, cell[i].size());
    }
    for (int d = 0; d < dim; ++d) {
        for (size_t k = 0; k < M; ++k) {
            if (k) out << ' ';
            const std::string& c = cell[k * static_cast<size_t>(dim) + static_cast<size_t>(d)];
            out << std::string(width - c.size(), ' ') << c;
        }
        if (d + 1 < dim) out << "\n";
    }
}

// mink_bound_<suffix>.csv (layer 0: obstacles first, then arenas) and segment_<suffix>.csv (tx,ty,xL,xM,xU per free
// segment of layer 0).  origin_bound_* (the un-inflated shapes) is not produced: it is outside the GPU path.
template <class Planner>
void storeRoutines(Planner& planner, const std::string& suffix, const Paths& paths = Paths::fromEnv()) {
    auto& fs = planner.freeSpace();
    long long d[8];
    fs.dims(d);
    const int dim = d[7] == 3 ? 3 : 2;
    const BoundaryInfo b = fs.getCSpaceBoundary(0);
    {
        std::ofstream out(paths.details + "/mink_bound_" + suffix + ".csv");
        if (!out) throw std::invalid_argument("cannot write boundary file under " + paths.details);
        for (const auto& o : b.obstacle) writeBoundary(out, o, dim), out << "\n";
        for (const auto& a : b.arena) writeBoundary(out, a, dim), out << "\n";
    }
    std::vector<long long> off;
    std::vector<double> xL, xU, xM;
    fs.getFreeSegmentsAll(off, xL, xU, xM);
    const PlannerParameter& prm = planner.getPlannerParameters();
    const auto& lim = prm.boundaryLimits;
    const size_t Ly = prm.numLineY, Lx = dim == 3 ? prm.numLineX : 1;
    auto grid = [](double lo, double hi, size_t n) {  // same expression as the planners' sweep grids
        std::vector<Coordinate> t(n);
        const double step = (hi - lo) / static_cast<double>(n - 1);
        for (size_t i = 0; i < n; ++i) t[i] = lo + static_cast<double>(i) * step;
        return t;
    };
    const std::vector<Coordinate> tx = dim == 3 ? grid(lim[0], lim[1], Lx) : std::vector<Coordinate>{0.0};
    const std::vector<Coordinate> ty = grid(lim[2], lim[3], Ly);
    std::ofstream out(paths.details + "/segment_" + suffix + ".csv");
    if (!out) throw std::invalid_argument("cannot write segment file under " + paths.details);
    const char sep = dim == 3 ? ',' : ' ';  // TestHRM2D.cpp:78-87 writes the planar dump space-separated
    for (size_t i = 0; i < Lx; ++i)
        for (size_t j = 0; j < Ly; ++j)
            for (long long k = off[i * Ly + j]; k < off[i * Ly + j + 1]; ++k) {
                const size_t s = static_cast<size_t>(k);
                if (dim == 3) out << tx[i] << sep;
                out << ty[j] << sep << xL[s] << sep << xM[s] << sep << xU[s] << "\n";
            }
}

}  // namespace hrm
