// fluid_simulation.hpp -- compiled-language host side above the C ABI (include/fluidsolver.h).
//
// The reference's host language is C# (Unity); no C# toolchain exists in the build image, so this header-only
// C++17 class is the compilable twin of Assets/Plugin/FluidSimulationNative.cs: the same field names, the same
// public methods (SetPaused, GetSourcePosition, SetSourcePosition) plus Step/AddDensity/AddVelocity, and the
// managed-side logic of the reference (Assets/Scripts/FluidSim.cs): parameter scaling :216-222/:554-556, the
// custom-source disc :485-533, the obstacle flood fill :302-388, the Update() order :390-450.
// Everything numerical happens behind fs_* (libfluidsolver.so on a B200; tests/host_emul on the CPU tier).
#pragma once
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "fluidsolver.h"

namespace fluidsim {

enum class ObstacleShape { Circle, Rectangle, Airfoil };

class FluidSimulation {
public:
    // inspector fields of the reference (FluidSim.cs:19-31, :34-55, :96-110)
    bool paused = false;
    int size = 128;
    int depth = 1; // 3D extension; 1 = the reference's 2D solver
    float physicalSize = 1.0f, resolutionMultiplier = 1.0f;
    float diffusion = 0.0001f, viscosity = 0.0001f, timeStep = 0.1f;
    bool autoAdjustParameters = true;
    bool enableCustomSource = false, sourceEmitsVelocity = false, sourcePulsing = false;
    float sourceStrength = 100.0f, sourceDirection = 0.0f, sourceVelocity = 10.0f, sourceRadius = 1.0f, sourcePulseRate = 1.0f;
    float sourcePositionX = 0.5f, sourcePositionY = 0.5f, sourcePositionZ = 0.5f;
    bool enableObstacle = true;
    ObstacleShape obstacleShape = ObstacleShape::Circle;
    float obstaclePositionX = 0.5f, obstaclePositionY = 0.5f, obstacleRadius = 0.1f, obstacleWidth = 0.2f, obstacleHeight = 0.2f;
    int itersDiffuse = 20, itersPressure = 20, solverKind = FS_JACOBI, deviceId = 0;
    bool useCudaGraph = true;

    FluidSimulation() = default;
    FluidSimulation(const FluidSimulation &) = delete;
    FluidSimulation &operator=(const FluidSimulation &) = delete;
    ~FluidSimulation() { fs_destroy(solver_); }

    int CurrentSize() const { return currentSize_; }
    const std::vector<uint8_t> &Obstacles() const { return obstacles_; }

    void SetPaused(bool Paused) { paused = Paused; }                                                  // :149
    std::pair<float, float> GetSourcePosition() const { return {sourcePositionX * currentSize_, sourcePositionY * currentSize_}; } // :979
    void SetSourcePosition(float x, float y) {                                                        // :984
        sourcePositionX = clamp01(x / currentSize_);
        sourcePositionY = clamp01(y / currentSize_);
    }

    // ResetSimulation + SetupObstacles, :213-235, :299
    void ResetSimulation() {
        fs_destroy(solver_);
        solver_ = nullptr;
        currentSize_ = (int)std::nearbyint(size * resolutionMultiplier); // Mathf.RoundToInt: half to even
        currentDepth_ = depth <= 1 ? 1 : (int)std::nearbyint(depth * resolutionMultiplier);
        cellSize_ = physicalSize / currentSize_;
        dtScale_ = autoAdjustParameters ? 128.0f / currentSize_ : 1.0f;
        fs_params p{};
        p.abi_version = FS_ABI_VERSION;
        p.nx = p.ny = currentSize_;
        p.nz = currentDepth_;
        p.iters_diffuse = itersDiffuse;
        p.iters_pressure = itersPressure;
        p.solver_kind = solverKind;
        p.enable_obstacle = enableObstacle ? 1 : 0;
        p.cell_size = cellSize_;
        p.raw_viscosity = viscosity;
        p.device_id = deviceId;
        p.slab_rank = 0;
        p.slab_count = 1;
        p.use_cuda_graph = useCudaGraph ? 1 : 0;
        if (fs_create(&p, &solver_) != FS_OK) throw std::runtime_error(std::string("fs_create: ") + fs_last_error(nullptr));
        SetupObstacles();
    }

    // SetupObstacles / RecursiveFloodFill / IsInsideShape, :302-388 (iterative flood fill; 3D: the 2D mask extruded)
    void SetupObstacles() {
        const int n = currentSize_;
        std::vector<uint8_t> plane((size_t)n * n, 0);
        if (enableObstacle) {
            const float extent = (obstacleShape == ObstacleShape::Circle ? obstacleRadius : obstacleWidth) * n;
            std::vector<std::pair<int, int>> todo;
            todo.emplace_back((int)std::nearbyint(obstaclePositionX * n), (int)std::nearbyint(obstaclePositionY * n));
            while (!todo.empty()) {
                const auto [x, y] = todo.back();
                todo.pop_back();
                if (x < 0 || x >= n || y < 0 || y >= n || plane[x + (size_t)y * n] || !inside(x, y, extent)) continue;
                plane[x + (size_t)y * n] = 1;
                todo.emplace_back(x + 1, y); todo.emplace_back(x - 1, y); todo.emplace_back(x, y + 1); todo.emplace_back(x, y - 1);
            }
        }
        obstacles_.assign((size_t)n * n * currentDepth_, 0);
        for (int k = 0; k < currentDepth_; k++) std::copy(plane.begin(), plane.end(), obstacles_.begin() + (size_t)k * n * n);
        check(fs_set_obstacles(solver_, obstacles_.data(), (int64_t)obstacles_.size()));
    }

    void AddDensity(float x, float y, float amount, float z = 0.0f) { check(fs_add_density(solver_, x, y, z, amount)); }       // :723
    void AddVelocity(float x, float y, float ax, float ay, float z = 0.0f, float az = 0.0f) { check(fs_add_velocity(solver_, x, y, z, ax, ay, az)); } // :731

    // Simulate, :551-570: the scaling stays on the host, the step is one native call
    void Step() {
        const float dt = autoAdjustParameters ? timeStep * dtScale_ : timeStep;
        const float diff = autoAdjustParameters ? diffusion / resolutionMultiplier : diffusion;
        const float visc = autoAdjustParameters ? viscosity / resolutionMultiplier : viscosity;
        check(fs_step(solver_, dt, visc, diff));
    }

    // Update, :390-450 (without input and rendering): sources first, then the step
    void Update(float deltaTime = 1.0f / 60.0f) {
        if (paused) return;
        elapsedTime_ += deltaTime;
        if (enableCustomSource) UpdateCustomSource();
        Step();
    }

    std::vector<float> Field(fs_field f) {
        std::vector<float> out((size_t)currentSize_ * currentSize_ * currentDepth_);
        check(fs_get_field(solver_, f, out.data(), (int64_t)out.size()));
        return out;
    }
    void Metrics(float *meanDensity, float *maxSpeed) { check(fs_get_metrics(solver_, meanDensity, maxSpeed, nullptr)); }
    fs_solver *Handle() { return solver_; }

private:
    fs_solver *solver_ = nullptr;
    int currentSize_ = 0, currentDepth_ = 1;
    float cellSize_ = 0, dtScale_ = 1, elapsedTime_ = 0;
    std::vector<uint8_t> obstacles_;

    static float clamp01(float v) { return v < 0 ? 0 : (v > 1 ? 1 : v); }
    void check(int rc) const {
        if (rc != FS_OK) throw std::runtime_error(std::string("fluidsolver: ") + fs_last_error(solver_));
    }
    bool inside(int x, int y, float extent) const { // IsInsideShape :353-388
        const float cx = obstaclePositionX * currentSize_, cy = obstaclePositionY * currentSize_;
        switch (obstacleShape) {
        case ObstacleShape::Circle: return (x - cx) * (x - cx) + (y - cy) * (y - cy) < extent * extent;
        case ObstacleShape::Rectangle: {
            const float hw = obstacleWidth * currentSize_ * 0.5f, hh = obstacleHeight * currentSize_ * 0.5f;
            return x > cx - hw && x < cx + hw && y > cy - hh && y < cy + hh;
        }
        default: {
            const float chord = 2 * obstacleWidth * currentSize_, t = 0.15f;
            const float u = (x - cx + chord / 2) / chord, v = (y - cy) / chord;
            if (u < 0 || u > 1 || std::fabs(v) > t) return false;
            const float half = 5 * t * (0.2969f * std::sqrt(u) - 0.1260f * u - 0.3516f * u * u + 0.2843f * u * u * u - 0.1015f * u * u * u * u);
            return std::fabs(v) <= half;
        }
        }
    }
    // UpdateCustomSource, :485-533: the disc of AddDensity/AddVelocity calls as ONE batched native call
    void UpdateCustomSource() {
        const float srcX = sourcePositionX * currentSize_, srcY = sourcePositionY * currentSize_;
        const float pulse = sourcePulsing ? std::fabs(std::sin(elapsedTime_ * sourcePulseRate * 3.14159274f)) : 1.0f;
        const float strength = sourceStrength * pulse * resolutionMultiplier;
        const float r = sourceRadius * resolutionMultiplier;
        std::vector<float> xs, ys, zs, ds, ax, ay;
        for (int i = std::max(0, (int)std::floor(srcX - r)); i <= std::min(currentSize_ - 1, (int)std::ceil(srcX + r)); i++)
            for (int j = std::max(0, (int)std::floor(srcY - r)); j <= std::min(currentSize_ - 1, (int)std::ceil(srcY + r)); j++) {
                const float dist = std::sqrt((i - srcX) * (i - srcX) + (j - srcY) * (j - srcY));
                if (dist > r) continue;
                const float falloff = 1.0f - dist / r;
                xs.push_back((float)i); ys.push_back((float)j); zs.push_back(sourcePositionZ * currentDepth_);
                ds.push_back(strength * falloff);
                const float ang = sourceDirection * 0.0174532924f; // Mathf.Deg2Rad
                ax.push_back(std::cos(ang) * sourceVelocity * resolutionMultiplier * falloff);
                ay.push_back(std::sin(ang) * sourceVelocity * resolutionMultiplier * falloff);
            }
        if (xs.empty()) return;
        check(fs_add_source_cells(solver_, (int64_t)xs.size(), xs.data(), ys.data(), zs.data(), ds.data(),
                                  sourceEmitsVelocity ? ax.data() : nullptr, sourceEmitsVelocity ? ay.data() : nullptr, nullptr));
    }
};

} // namespace fluidsim
