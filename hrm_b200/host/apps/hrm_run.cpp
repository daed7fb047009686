// hrm_run — one driver for the three GPU-backed planners, written from the on-disk schemas alone (the
// resources/*.csv inputs and the result/benchmark/time_*.csv statistics the reference's plotting scripts read).
//
//     hrm_run <hrm3d|hrm2d|prob3d> map=<name> robot=<name> [trials=1] [time=60] [slices=60] [so3=icosahedron]
//             [lx=0] [ly=0] [precision=f64_strict|f64|f32] [per_orientation=0|1] [seed=N] [batch=8] [details=0|1]
//
// Every planner kind is one row of kKinds: its dimension, object family, statistics file and the columns of that file as
// (name, getter) pairs; main() is a single generic loop over trials that fills a Trial record and lets the column table
// write it.  Paths come from HRM_RESOURCES_PATH / HRM_CONFIG_PATH / HRM_RESULT_PATH (hrm_io.hpp).
#include "../hrm_articulated.hpp"
#include "../hrm_io.hpp"

#include <functional>
#include <map>
#include <memory>

namespace {

struct Trial {
    bool solved = false;
    hrm::Time time;
    hrm::PlannerParameter param;
    size_t vertices = 0, edges = 0, pathNodes = 0;
    double cost = 0.0;
};
struct Column {
    const char* name;
    std::function<double(const Trial&)> get;
    bool integral;
};
const Column cSuccess{"SUCCESS", [](const Trial& t) { return t.solved ? 1.0 : 0.0; }, true};
const Column cBuild{"BUILD_TIME", [](const Trial& t) { return t.time.buildTime; }, false};
const Column cSearch{"SEARCH_TIME", [](const Trial& t) { return t.time.searchTime; }, false};
const Column cPlan{"PLAN_TIME", [](const Trial& t) { return t.time.totalTime; }, false};
const Column cLayers{"N_LAYERS", [](const Trial& t) { return double(t.param.numSlice); }, true};
const Column cNx{"N_X", [](const Trial& t) { return double(t.param.numLineX); }, true};
const Column cNy{"N_Y", [](const Trial& t) { return double(t.param.numLineY); }, true};
const Column cNode{"GRAPH_NODE", [](const Trial& t) { return double(t.vertices); }, true};
const Column cEdge{"GRAPH_EDGE", [](const Trial& t) { return double(t.edges); }, true};
const Column cPath{"PATH_NODE", [](const Trial& t) { return double(t.pathNodes); }, true};

struct Kind {
    const char* dim;
    const char* family;     // object family in the resource file names
    const char* statsFile;  // under result/benchmark
    const char* suffix;     // of the detail dumps
    int numParam;           // boundary samples per parameter
    std::vector<Column> columns;
};
const std::map<std::string, Kind> kKinds = {
    {"hrm3d", {"3D", "superquadrics", "time_high_3D.csv", "hrm_3D", 10, {cSuccess, cBuild, cSearch, cPlan, cLayers, cNx, cNy, cNode, cEdge, cPath}}},
    {"hrm2d", {"2D", "superellipse", "time_hrm_2D.csv", "hrm_2D", 50, {cBuild, cSearch, cPlan, cNode, cEdge, cPath}}},
    {"prob3d", {"3D", "superquadrics", "time_prob_high_3D.csv", "prob_hrm_3D", 10, {cSuccess, cPlan, cLayers, cNx, cNy, cNode, cEdge, cPath}}},
};

struct Options {
    std::map<std::string, std::string> kv;
    std::string str(const std::string& k, const std::string& dflt) const {
        auto it = kv.find(k);
        return it == kv.end() ? dflt : it->second;
    }
    long num(const std::string& k, long dflt) const {
        auto it = kv.find(k);
        return it == kv.end() ? dflt : std::strtol(it->second.c_str(), nullptr, 10);
    }
};

template <class Planner>
Trial record(Planner& planner, double timeLimit, const Kind& kind, bool details, const hrm::Paths& paths) {
    planner.plan(timeLimit);
    const auto res = planner.getPlanningResult();
    Trial t;
    t.solved = res.solved;
    t.time = res.planningTime;
    t.param = planner.getPlannerParameters();
    t.vertices = res.graphStructure.vertex.size();
    t.edges = res.graphStructure.edge.size();
    t.pathNodes = res.solutionPath.PathId.size();
    t.cost = res.solutionPath.cost;
    if (details) {
        hrm::storeGraphInfo(res.graphStructure, kind.suffix, paths);
        hrm::storePathInfo(res.solutionPath, kind.suffix, paths);
        hrm::storeRoutines(planner, kind.suffix, paths);
    }
    return t;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 2 || !kKinds.count(argv[1])) {
        std::cerr << "usage: hrm_run <hrm3d|hrm2d|prob3d> map=<name> robot=<name> [trials=N] [time=S] [slices=N] [so3=<method>] [lx=N] [ly=N]"
                     " [precision=f64_strict|f64|f32] [per_orientation=0|1] [seed=N] [batch=N] [details=0|1]"
                  << std::endl;
        return 1;
    }
    const std::string which = argv[1];
    const Kind& kind = kKinds.at(which);
    Options opt;
    for (int i = 2; i < argc; ++i) {
        const std::string a = argv[i];
        const size_t eq = a.find('=');
        if (eq == std::string::npos) {
            std::cerr << "hrm_run: expected key=value, got '" << a << "'" << std::endl;
            return 1;
        }
        opt.kv[a.substr(0, eq)] = a.substr(eq + 1);
    }
    if (!opt.kv.count("map") || !opt.kv.count("robot")) {
        std::cerr << "hrm_run: map= and robot= are required" << std::endl;
        return 1;
    }
    const std::string map = opt.str("map", ""), robotName = opt.str("robot", "");
    const std::string pr = opt.str("precision", "f64_strict");
    const int precision = pr == "f32" ? HRMGPU_F32 : (pr == "f64" ? HRMGPU_F64 : HRMGPU_F64_STRICT);
    const bool perOrientation = opt.num("per_orientation", 0) != 0, details = opt.num("details", 0) != 0;
    const long trials = opt.num("trials", 1);
    const double timeLimit = double(opt.num("time", which == "hrm2d" ? 10 : 60));
    const hrm::Paths paths = hrm::Paths::fromEnv();

    try {
        hrm::parsePlanningConfig(kind.family, map, robotName, kind.dim, paths);
        const std::string cfg = paths.config + "/";
        hrm::PlannerParameter param;
        param.numLineX = size_t(opt.num("lx", 0));
        param.numLineY = size_t(opt.num("ly", 0));
        hrm::PlanningRequest req;
        std::vector<Trial> done;
        std::function<Trial(long)> runTrial;

        // the planner kinds differ only in how the robot / environment are loaded and which planner object is built
        hrm::PlannerSetting3D env3(kind.numParam);
        hrm::PlannerSetting2D env2(kind.numParam);
        std::unique_ptr<hrm::MultiBodyTree3D> robot3;
        std::unique_ptr<hrm::MultiBodyTree2D> robot2;
        if (std::string(kind.dim) == "3D") {
            env3.loadEnvironment(cfg);
            std::string quatFile = "0";
            if (which == "hrm3d") {
                const long slices = opt.num("slices", 60);
                quatFile = paths.resources + "/SO3_sequence/q_" + opt.str("so3", "icosahedron") + "_" + std::to_string(slices) + ".csv";
            }
            robot3 = std::make_unique<hrm::MultiBodyTree3D>(hrm::loadRobotMultiBody3D(cfg, quatFile, kind.numParam));
            if (which == "hrm3d" && robot3->getBase().getQuatSamples().empty())
                throw std::invalid_argument("no orientation samples in " + quatFile + " (run-time SO(3) sampling is not part of the GPU path)");
            hrm::defineParameters(*robot3, env3, param);
            if (which == "prob3d") param.numSlice = 0, req.isRobotRigid = false;
            req.start = env3.getEndPoints().at(0), req.goal = env3.getEndPoints().at(1);
        } else {
            env2.loadEnvironment(cfg);
            robot2 = std::make_unique<hrm::MultiBodyTree2D>(hrm::loadRobotMultiBody2D(cfg, kind.numParam));
            param.numSlice = size_t(opt.num("slices", 20));
            param.numPoint = 5;
            hrm::defineParameters(*robot2, env2, param);
            req.start = env2.getEndPoints().at(0), req.goal = env2.getEndPoints().at(1);
        }
        req.parameters = param;
        std::cout << which << ": map " << map << ", robot " << robotName << ", " << param.numSlice << " C-slices, sweep lines " << param.numLineX << " x "
                  << param.numLineY << std::endl;

        if (which == "hrm3d") {
            runTrial = [&](long) {
                hrm::planners::HRM3D p(*robot3, env3.getArena(), env3.getObstacle(), req, 0, precision, perOrientation);
                return record(p, timeLimit, kind, details, paths);
            };
        } else if (which == "prob3d") {
            const std::string urdf = paths.resources + "/3D/urdf/" + robotName + ".urdf";
            const unsigned long long seed = static_cast<unsigned long long>(opt.num("seed", 20220726));
            const size_t batch = size_t(opt.num("batch", 8));
            runTrial = [&, urdf, seed, batch](long i) {
                hrm::planners::ProbHRM3D p(*robot3, urdf, env3.getArena(), env3.getObstacle(), req, 0, precision, perOrientation);
                p.setSeed(seed + static_cast<unsigned long long>(i));
                p.setBatch(batch);
                return record(p, timeLimit, kind, details, paths);
            };
        } else {
            runTrial = [&](long) {
                hrm::planners::HRM2D p(*robot2, env2.getArena(), env2.getObstacle(), req, 0, precision, perOrientation);
                return record(p, timeLimit, kind, details, paths);
            };
        }

        for (long i = 0; i < trials; ++i) {
            const Trial t = runTrial(i);
            std::cout << "trial " << i + 1 << '/' << trials << ": solved=" << t.solved << " build=" << t.time.buildTime << "s search=" << t.time.searchTime
                      << "s total=" << t.time.totalTime << "s slices=" << t.param.numSlice << " lines=" << t.param.numLineX << 'x' << t.param.numLineY
                      << " vertices=" << t.vertices << " edges=" << t.edges << " path_nodes=" << t.pathNodes << " cost=" << t.cost << std::endl;
            done.push_back(t);
        }

        std::ofstream stats(paths.benchmark + "/" + kind.statsFile);
        if (!stats) throw std::invalid_argument("cannot write under " + paths.benchmark);
        for (size_t c = 0; c < kind.columns.size(); ++c) stats << (c ? "," : "") << kind.columns[c].name;
        stats << "\n";
        for (const Trial& t : done) {
            for (size_t c = 0; c < kind.columns.size(); ++c) {
                const double v = kind.columns[c].get(t);
                stats << (c ? "," : "");
                if (kind.columns[c].integral)
                    stats << static_cast<long long>(v);
                else
                    stats << v;
            }
            stats << "\n";
        }
    } catch (const std::exception& e) {
        std::cerr << "hrm_run: " << e.what() << std::endl;
        return 2;
    }
    return 0;
}
