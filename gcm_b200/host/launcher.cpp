// gcmb_exe — command-line entry with the options of the reference's launcher (src/launcher/main.cpp:22-71,
// src/launcher/getopt_wrapper.hpp): --task (-t) names a task, --out (-o) the output directory.  A task is a
// plain-text task file (grammar in task_file.cpp) or the id of one shipped in gcm_b200/tasks/ (cubic2d, cubic3d,
// acoustic, ndi_empty, ndi, titan: the reference launcher's cubic demo tasks; cubeAcs, cubeEls: its simplex cube tasks;
// simplex_plate).
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <getopt.h>
#include <libgen.h>
#include <unistd.h>

#include "gcmb_host.hpp"

namespace {

std::string readFile(const std::string& path) {
	std::ifstream in(path);
	if (!in) { return std::string(); }
	std::stringstream ss;
	ss << in.rdbuf();
	return ss.str();
}

std::string executableDirectory() {
	char buf[4096];
	const ssize_t n = readlink("/proc/self/exe", buf, sizeof buf - 1);
	if (n <= 0) { return "."; }
	buf[n] = 0;
	return dirname(buf);
}

}  // namespace

int main(int argc, char** argv) {
	std::string taskId, outputDirectory;
	int device = 0;
	bool verbose = true;
	static struct option options[] = {
		{"task", required_argument, 0, 't'}, {"out", required_argument, 0, 'o'},
		{"device", required_argument, 0, 'd'}, {"quiet", no_argument, 0, 'q'}, {0, 0, 0, 0}};
	while (true) {
		int index = 0;
		const int c = getopt_long_only(argc, argv, "t:o:d:q", options, &index);
		if (c == -1) { break; }
		if (c == 't') { taskId = optarg; }
		else if (c == 'o') { outputDirectory = optarg; }
		else if (c == 'd') { device = std::atoi(optarg); }
		else if (c == 'q') { verbose = false; }
	}
	std::string text = readFile(taskId);
	if (text.empty()) { text = readFile(executableDirectory() + "/tasks/" + taskId + ".task"); }
	if (text.empty()) {
		std::fprintf(stderr, "FATAL gcm.main - Invalid task file\n");   // main.cpp:48-50
		return -1;
	}
	try {
		gcmb::Task task = gcmb::parseTaskText(text);
		task.globalSettings.outputDirectory = outputDirectory;
		task.globalSettings.verboseTimeSteps = verbose;
		task.device.device = device;
		const auto t1 = std::chrono::high_resolution_clock::now();
		auto engine = gcmb::createEngine(task);
		engine->run();
		const auto t2 = std::chrono::high_resolution_clock::now();
		std::printf("INFO gcm.main - Time of calculation, microseconds = %lld\n",
				(long long) std::chrono::duration_cast<std::chrono::microseconds>(t2 - t1).count());
		std::printf("INFO gcm.main - steps = %d, time = %.17g\n", engine->stepsDone(), gcmb::Clock::Time());
		if (auto cubic = std::dynamic_pointer_cast<gcmb::cubic::EngineBase>(engine)) {
			for (const auto& body : task.bodies) {
				double sum = 0;
				if (gcmb_cubic_checksum(cubic->getMesh(body.first)->handle(), &sum) == GCMB_OK) {
					std::printf("INFO gcm.main - body %zu checksum = %.17g\n", body.first, sum);
				}
			}
		} else if (auto simplex = std::dynamic_pointer_cast<gcmb::simplex::Engine>(engine)) {
			for (const auto& body : task.bodies) {
				const auto& pde = simplex->getMesh(body.first)->pdeAll();
				const int M = simplex->getMesh(body.first)->M;
				double sum = 0;
				for (size_t i = 0; i < pde.size(); i++) { sum += (double) (i % (size_t) M + 1) * pde[i]; }
				std::printf("INFO gcm.main - body %zu vertices = %zu checksum = %.17g\n", body.first, pde.size() / (size_t) M, sum);
			}
			std::printf("INFO gcm.main - node computations the reference would have thrown on = %d\n", simplex->errorCount());
		}
	} catch (const std::exception& e) {
		std::fprintf(stderr, "FATAL gcm.main - %s\n", e.what());
		return 1;
	}
	return 0;
}
