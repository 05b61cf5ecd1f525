// host_demo.cpp -- drives include/fluid_simulation.hpp (the compiled-language host mirror of the reference
// component) through the C ABI: N Update() frames of the 128-circle scene configuration with a velocity-emitting
// custom source, then prints checksums of the fields.  The CPU tier links it against the host-emulated core and
// compares the numbers with the Python mirror; on a B200 the same program links against libfluidsolver.so.
#include <cstdio>
#include <cstdlib>

#include "fluid_simulation.hpp"

int main(int argc, char **argv) {
    const int size = argc > 1 ? atoi(argv[1]) : 48;
    const int frames = argc > 2 ? atoi(argv[2]) : 3;
    const int depth = argc > 3 ? atoi(argv[3]) : 1;
    try {
        fluidsim::FluidSimulation sim;
        sim.size = size;
        sim.depth = depth;
        sim.useCudaGraph = false;
        sim.enableCustomSource = true;
        sim.sourceEmitsVelocity = true;
        sim.sourceDirection = 90.0f;
        sim.sourceRadius = 2.0f;
        sim.sourcePositionY = 0.2f;
        sim.ResetSimulation();
        if (argc > 4) sim.obstacleShape = (fluidsim::ObstacleShape)atoi(argv[4]);
        sim.SetupObstacles();
        for (int f = 0; f < frames; f++) {
            sim.AddForceToArea(0.3f * size + f, 0.6f * size, 2.5f, -1.25f, 3.0f);
            sim.Update();
        }
        const fs_field fields[] = {FS_DENSITY, FS_VX, FS_VY, FS_PRESSURE};
        const char *names[] = {"density", "vx", "vy", "pressure"};
        for (int k = 0; k < 4; k++) {
            const std::vector<float> a = sim.Field(fields[k]);
            double sum = 0, sq = 0;
            for (float v : a) { sum += v; sq += (double)v * v; }
            printf("%s %.9e %.9e\n", names[k], sum, sq);
        }
        float mean, mx;
        sim.Metrics(&mean, &mx);
        printf("metrics %.9e %.9e\n", mean, mx);
        sim.colorMode = 2; // DensityBased
        const std::vector<float> rgba = sim.UpdateVisualization();
        double csum = 0, csq = 0;
        for (float v : rgba) { csum += v; csq += (double)v * v; }
        printf("rgba %.9e %.9e\n", csum, csq);
        long painted = 0;
        for (uint8_t p : sim.DrawStreamlines()) painted += p;
        printf("streamline_pixels %ld\n", painted);
        long obst = 0;
        for (uint8_t o : sim.Obstacles()) obst += o;
        printf("obstacle_cells %ld\n", obst);
    } catch (const std::exception &e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
