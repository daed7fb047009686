// Command-line twin of the reference's test/benchmark/BenchProbHRM3D.cpp:10-112 on the GPU-backed planner: same positional
// arguments, console lines and time_prob_high_3D.csv columns.  Paths: HRM_RESOURCES_PATH / HRM_CONFIG_PATH / HRM_RESULT_PATH;
// HRM_PRECISION=f64_strict|f64|f32, HRM_PER_ORIENTATION=1, HRM_SEED (draws come from a seeded generator instead of
// srand(time)), HRM_BATCH (C-layers built ahead per GPU call).
#include "../hrm_articulated.hpp"
#include "../hrm_io.hpp"

#include <cstring>

int main(int argc, char** argv) {
    if (argc >= 5) {
        std::cout << "Benchmark: Probabilistic Highway RoadMap for 3D articulated-body planning" << std::endl;
        std::cout << "----------" << std::endl;
    } else {
        std::cerr << "Usage: Please add 1) Map type 2) Robot type 3) Num of trials 4) Max planning time (in seconds) 5) [optional] "
                     "Num of sweep lines (x-direction) 6) [optional] Num of sweep lines (y-direction)"
                  << std::endl;
        return 1;
    }
    const std::string mapType = argv[1], robotType = argv[2];
    const auto numTrial = size_t(atoi(argv[3]));
    const auto MAX_PLAN_TIME = double(atoi(argv[4]));
    const int numLineX = argc > 6 ? atoi(argv[5]) : 0;
    const int numLineY = argc > 6 ? atoi(argv[6]) : 0;
    const int NUM_SURF_PARAM = 10;

    const hrm::Paths paths = hrm::Paths::fromEnv();
    int precision = HRMGPU_F64_STRICT;
    if (const char* p = std::getenv("HRM_PRECISION")) precision = !strcmp(p, "f32") ? HRMGPU_F32 : (!strcmp(p, "f64") ? HRMGPU_F64 : HRMGPU_F64_STRICT);
    const bool perOrientation = std::getenv("HRM_PER_ORIENTATION") && atoi(std::getenv("HRM_PER_ORIENTATION")) != 0;
    const unsigned long long seed = std::getenv("HRM_SEED") ? strtoull(std::getenv("HRM_SEED"), nullptr, 10) : 20220726ULL;
    const size_t batch = std::getenv("HRM_BATCH") ? size_t(atoi(std::getenv("HRM_BATCH"))) : 8;

    try {
        hrm::parsePlanningConfig("superquadrics", mapType, robotType, "3D", paths);
        hrm::PlannerSetting3D env3D(NUM_SURF_PARAM);
        env3D.loadEnvironment(paths.config + "/");
        hrm::MultiBodyTree3D robot = hrm::loadRobotMultiBody3D(paths.config + "/", "0", NUM_SURF_PARAM);
        const std::string urdfFile = paths.resources + "/3D/urdf/" + robotType + ".urdf";

        hrm::PlannerParameter param;
        param.numSlice = 0;
        param.numLineX = size_t(numLineX);
        param.numLineY = size_t(numLineY);
        hrm::defineParameters(robot, env3D, param);
        std::cout << "Initial number of sweep lines: {" << param.numLineX << ", " << param.numLineY << '}' << std::endl;
        std::cout << "----------" << std::endl;

        hrm::PlanningRequest req;
        req.isRobotRigid = false;
        req.parameters = param;
        req.start = env3D.getEndPoints().at(0);
        req.goal = env3D.getEndPoints().at(1);

        std::ofstream fileTimeStatistics(paths.benchmark + "/time_prob_high_3D.csv");
        if (!fileTimeStatistics) throw std::invalid_argument("cannot write under " + paths.benchmark);
        fileTimeStatistics << "SUCCESS" << ',' << "PLAN_TIME" << ',' << "N_LAYERS" << ',' << "N_X" << ',' << "N_Y" << ',' << "GRAPH_NODE" << ','
                           << "GRAPH_EDGE" << ',' << "PATH_NODE"
                           << "\n";
        std::cout << "Start benchmark..." << std::endl;
        std::cout << " Map type: [" << mapType << "]; Robot type: [" << robotType << "]" << std::endl;
        for (size_t i = 0; i < numTrial; i++) {
            std::cout << "Number of trials: " << i + 1 << std::endl;
            hrm::planners::ProbHRM3D probHRM(robot, urdfFile, env3D.getArena(), env3D.getObstacle(), req, 0, precision, perOrientation);
            probHRM.setSeed(seed + i);
            probHRM.setBatch(batch);
            probHRM.plan(MAX_PLAN_TIME);
            const auto res = probHRM.getPlanningResult();
            const auto prm = probHRM.getPlannerParameters();
            hrm::displayPlanningTimeInfo(res.planningTime);
            hrm::displayGraphInfo(res.graphStructure);
            hrm::displayPathInfo(res.solutionPath);
            std::cout << "Final number of C-slices: " << prm.numSlice << std::endl;
            std::cout << "Final number of sweep lines: {" << prm.numLineX << ", " << prm.numLineY << '}' << std::endl;
            std::cout << "==========" << std::endl;
            fileTimeStatistics << static_cast<int>(res.solved) << ',' << res.planningTime.totalTime << ',' << prm.numSlice << ',' << prm.numLineX
                               << ',' << prm.numLineY << ',' << res.graphStructure.vertex.size() << ',' << res.graphStructure.edge.size() << ','
                               << res.solutionPath.PathId.size() << "\n";
        }
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << std::endl;
        return 2;
    }
    return 0;
}
