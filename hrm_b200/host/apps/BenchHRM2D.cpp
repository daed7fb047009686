// Command-line twin of the reference's test/benchmark/BenchHRM2D.cpp:14-124 on the GPU-backed planner: same positional
// arguments, console lines and time_hrm_2D.csv columns.  Paths: HRM_RESOURCES_PATH / HRM_CONFIG_PATH / HRM_RESULT_PATH;
// HRM_PRECISION=f64_strict|f64|f32, HRM_PER_ORIENTATION=1.
#include "../hrm_io.hpp"

#include <cstring>

int main(int argc, char** argv) {
    if (argc >= 5) {
        std::cout << "Benchmark: Highway RoadMap for 2D rigid-body planning" << std::endl;
        std::cout << "----------" << std::endl;
    } else {
        std::cerr << "Usage: Please add 1) Map type 2) Robot type 3) Num of trials 4) Num of slices 5) [optional] Num of sweep lines" << std::endl;
        return 1;
    }
    const std::string mapType = argv[1], robotType = argv[2];
    const int numTrial = atoi(argv[3]);
    const int numSlice = atoi(argv[4]);
    const int numLineY = argc > 5 ? atoi(argv[5]) : 0;
    const int NUM_CURVE_PARAM = 50;
    const double MAX_PLAN_TIME = 10.0;

    const hrm::Paths paths = hrm::Paths::fromEnv();
    int precision = HRMGPU_F64_STRICT;
    if (const char* p = std::getenv("HRM_PRECISION")) precision = !strcmp(p, "f32") ? HRMGPU_F32 : (!strcmp(p, "f64") ? HRMGPU_F64 : HRMGPU_F64_STRICT);
    const bool perOrientation = std::getenv("HRM_PER_ORIENTATION") && atoi(std::getenv("HRM_PER_ORIENTATION")) != 0;

    try {
        hrm::parsePlanningConfig("superellipse", mapType, robotType, "2D", paths);
        hrm::MultiBodyTree2D robot = hrm::loadRobotMultiBody2D(paths.config + "/", NUM_CURVE_PARAM);
        hrm::PlannerSetting2D env2D(NUM_CURVE_PARAM);
        env2D.loadEnvironment(paths.config + "/");

        hrm::PlannerParameter param;
        param.numSlice = static_cast<size_t>(numSlice);
        param.numLineY = static_cast<size_t>(numLineY);
        param.numPoint = 5;
        hrm::defineParameters(robot, env2D, param);
        std::cout << "Initial number of C-slices: " << param.numSlice << std::endl;
        std::cout << "Initial number of sweep lines: " << param.numLineY << std::endl;
        std::cout << "----------" << std::endl;

        hrm::PlanningRequest req;
        req.parameters = param;
        req.start = env2D.getEndPoints().at(0);
        req.goal = env2D.getEndPoints().at(1);

        std::cout << "Start benchmark..." << std::endl;
        std::cout << " Map type: [" << mapType << "]; Robot type: [" << robotType << "]" << std::endl;
        std::vector<std::vector<double>> timeStatistics;
        for (int i = 0; i < numTrial; i++) {
            std::cout << "Number of trials: " << i + 1 << std::endl;
            hrm::planners::HRM2D hrm(robot, env2D.getArena(), env2D.getObstacle(), req, 0, precision, perOrientation);
            hrm.plan(MAX_PLAN_TIME);
            const auto res = hrm.getPlanningResult();
            timeStatistics.push_back({res.planningTime.buildTime, res.planningTime.searchTime, res.planningTime.totalTime,
                                      static_cast<double>(res.graphStructure.vertex.size()), static_cast<double>(res.graphStructure.edge.size()),
                                      static_cast<double>(res.solutionPath.PathId.size())});
            std::cout << "Roadmap build time: " << res.planningTime.buildTime << "s" << std::endl;
            std::cout << "Path search time: " << res.planningTime.searchTime << "s" << std::endl;
            std::cout << "Total Planning Time: " << res.planningTime.totalTime << 's' << std::endl;
            std::cout << "Number of valid configurations: " << res.graphStructure.vertex.size() << std::endl;
            std::cout << "Number of valid edges: " << res.graphStructure.edge.size() << std::endl;
            std::cout << "Number of configurations in Path: " << res.solutionPath.PathId.size() << std::endl;
            std::cout << "Cost: " << res.solutionPath.cost << std::endl;
        }
        std::ofstream fileTimeStatistics(paths.benchmark + "/time_hrm_2D.csv");
        if (!fileTimeStatistics) throw std::invalid_argument("cannot write under " + paths.benchmark);
        fileTimeStatistics << "BUILD_TIME" << ',' << "SEARCH_TIME" << ',' << "PLAN_TIME" << ',' << "GRAPH_NODE" << ',' << "GRAPH_EDGE" << ','
                           << "PATH_NODE"
                           << "\n";
        for (const auto& t : timeStatistics) fileTimeStatistics << t[0] << ',' << t[1] << ',' << t[2] << ',' << t[3] << ',' << t[4] << ',' << t[5] << "\n";
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << std::endl;
        return 2;
    }
    return 0;
}
