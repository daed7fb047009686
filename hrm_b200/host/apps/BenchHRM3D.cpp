// Command-line twin of the reference's test/benchmark/BenchHRM3D.cpp:11-121 on the GPU-backed planner: same positional
// arguments, same console lines, same time_high_3D.csv columns.  Paths come from HRM_RESOURCES_PATH / HRM_CONFIG_PATH /
// HRM_RESULT_PATH (hrm_io.hpp) instead of configure-time macros; HRM_PRECISION=f64_strict|f64|f32 and
// HRM_PER_ORIENTATION=1 select the arithmetic mode and the per-orientation robot pose.
#include "../hrm_io.hpp"

#include <cstring>

int main(int argc, char** argv) {
    if (argc >= 7) {
        std::cout << "Benchmark: Highway RoadMap for 3D rigid-body planning" << std::endl;
        std::cout << "----------" << std::endl;
    } else {
        std::cerr << "Usage: Please add 1) Map type 2) Robot type 3) Num of trials "
                     "4) Max planning time 5) Num of slices 6) Method for "
                     "pre-defined SO(3) samples 7) [optional] Num of sweep lines "
                     "(x-direction) 8) [optional] Num of sweep lines (y-direction)"
                  << std::endl;
        return 1;
    }
    const std::string mapType = argv[1], robotType = argv[2];
    const auto numTrial = size_t(atoi(argv[3]));
    const auto MAX_PLAN_TIME = double(atoi(argv[4]));
    const int numSlice = atoi(argv[5]);
    const std::string methodSO3 = argv[6];
    const int numLineX = argc > 8 ? atoi(argv[7]) : 0;
    const int numLineY = argc > 8 ? atoi(argv[8]) : 0;
    const int NUM_SURF_PARAM = 10;

    const hrm::Paths paths = hrm::Paths::fromEnv();
    int precision = HRMGPU_F64_STRICT;
    if (const char* p = std::getenv("HRM_PRECISION")) precision = !strcmp(p, "f32") ? HRMGPU_F32 : (!strcmp(p, "f64") ? HRMGPU_F64 : HRMGPU_F64_STRICT);
    const bool perOrientation = std::getenv("HRM_PER_ORIENTATION") && atoi(std::getenv("HRM_PER_ORIENTATION")) != 0;

    try {
        hrm::parsePlanningConfig("superquadrics", mapType, robotType, "3D", paths);
        hrm::PlannerSetting3D env3D(NUM_SURF_PARAM);
        env3D.loadEnvironment(paths.config + "/");

        std::string quaternionFilename = "0";
        if (strcmp(argv[6], "0") != 0) quaternionFilename = paths.resources + "/SO3_sequence/q_" + methodSO3 + "_" + std::to_string(numSlice) + ".csv";
        auto robot = hrm::loadRobotMultiBody3D(paths.config + "/", quaternionFilename, NUM_SURF_PARAM);
        if (robot.getBase().getQuatSamples().empty()) {
            std::cerr << "Run-time SO(3) sampling is not part of the GPU path: pass a pre-defined sequence (e.g. icosahedron 60)" << std::endl;
            return 1;
        }

        hrm::PlannerParameter param;
        param.numSlice = size_t(numSlice);
        param.numLineX = size_t(numLineX);
        param.numLineY = size_t(numLineY);
        hrm::defineParameters(robot, env3D, param);
        std::cout << "Initial number of C-slices: " << param.numSlice << std::endl;
        std::cout << "Initial number of sweep lines: {" << param.numLineX << ", " << param.numLineY << '}' << std::endl;
        std::cout << "----------" << std::endl;

        hrm::PlanningRequest req;
        req.parameters = param;
        req.start = env3D.getEndPoints().at(0);
        req.goal = env3D.getEndPoints().at(1);

        std::ofstream fileTimeStatistics(paths.benchmark + "/time_high_3D.csv");
        if (!fileTimeStatistics) throw std::invalid_argument("cannot write under " + paths.benchmark);
        fileTimeStatistics << "SUCCESS" << ',' << "BUILD_TIME" << ',' << "SEARCH_TIME" << ',' << "PLAN_TIME" << ',' << "N_LAYERS" << ',' << "N_X" << ','
                           << "N_Y" << ',' << "GRAPH_NODE" << ',' << "GRAPH_EDGE" << ',' << "PATH_NODE"
                           << "\n";
        std::cout << "Start benchmark..." << std::endl;
        std::cout << " Map type: [" << mapType << "]; Robot type: [" << robotType << "]" << std::endl;
        for (size_t i = 0; i < numTrial; i++) {
            std::cout << "Number of trials: " << i + 1 << std::endl;
            hrm::planners::HRM3D hrm(robot, env3D.getArena(), env3D.getObstacle(), req, 0, precision, perOrientation);
            hrm.plan(MAX_PLAN_TIME);
            const auto res = hrm.getPlanningResult();
            hrm::displayPlanningTimeInfo(res.planningTime);
            hrm::displayGraphInfo(res.graphStructure);
            hrm::displayPathInfo(res.solutionPath);
            std::cout << "Final number of C-slices: " << hrm.getPlannerParameters().numSlice << std::endl;
            std::cout << "Final number of sweep lines: {" << hrm.getPlannerParameters().numLineX << ", " << hrm.getPlannerParameters().numLineY << '}'
                      << std::endl;
            std::cout << "==========" << std::endl;
            fileTimeStatistics << static_cast<int>(res.solved) << ',' << res.planningTime.buildTime << ',' << res.planningTime.searchTime << ','
                               << res.planningTime.totalTime << ',' << hrm.getPlannerParameters().numSlice << ','
                               << hrm.getPlannerParameters().numLineX << ',' << hrm.getPlannerParameters().numLineY << ','
                               << res.graphStructure.vertex.size() << ',' << res.graphStructure.edge.size() << ',' << res.solutionPath.PathId.size()
                               << "\n";
            if (i + 1 == numTrial && std::getenv("HRM_STORE_DETAILS")) {  // what demo/DemoHRMRigidBodyPlanning3D.cpp:55-61 stores
                hrm::storeGraphInfo(res.graphStructure, "hrm_3D", paths);
                hrm::storePathInfo(res.solutionPath, "hrm_3D", paths);
                hrm::storeRoutines(hrm, "hrm_3D", paths);
            }
        }
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << std::endl;
        return 2;
    }
    return 0;
}
