// fluid_simulation.hpp -- compiled-language host side above the C ABI (include/fluidsolver.h).
//
// The reference's host language is C# (Unity); no C# toolchain exists in the build image, so this header-only
// C++17 class is the compilable twin of Assets/Plugin/FluidSimulationNative.cs: the same field names, the same
// public methods (SetPaused, GetSourcePosition, SetSourcePosition) plus Step/AddDensity/AddVelocity, and the
// managed-side logic of the reference (Assets/Scripts/FluidSim.cs): parameter scaling :216-222/:554-556, the
// custom-source disc :485-533, AddForceToArea :452-483, the obstacle shape parameters :302-388, the Update() order
// :390-450, and the per-frame consumers UpdateVisualization :755-866 / DrawStreamlines :886-959 (device jobs + host Bresenham).
// Everything numerical happens behind fs_* (libfluidsolver.so on a B200; tests/host_emul on the CPU tier).
#pragma once
#include <algorithm>
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
    float obstaclePositionX = 0.5f, obstaclePositionY = 0.5f, obstaclePositionZ = 0.5f;
    float obstacleRadius = 0.1f, obstacleWidth = 0.2f, obstacleHeight = 0.2f;
    // visualisation (:57-95): the fields UpdateVisualization / DrawStreamlines pass to their jobs
    int colorMode = 0; // ColorMode: 0 SingleColor, 1 Gradient, 2 DensityBased, 3 PressureBased, 4 Streamlines
    float colourIntensity = 1.0f, mediumDensityThreshold = 50.0f, highDensityThreshold = 200.0f;
    float lowPressureThreshold = -50.0f, highPressureThreshold = 50.0f;
    bool visualizeSourcePosition = true;
    int streamlineDensity = 4;
    float streamlineScale = 1.0f, streamlineThickness = 1.0f, viewSliceZ = 0.5f;
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

    // SetupObstacles, :302-327: RecursiveFloodFill over IsInsideShape runs on the device (fs_build_obstacles); only the
    // shape parameters, computed as the reference computes them (:308-324, :355-370), cross the boundary.
    void SetupObstacles() {
        const int n = currentSize_;
        obstacles_.assign((size_t)n * n * currentDepth_, 0);
        if (!enableObstacle) { check(fs_set_obstacles(solver_, obstacles_.data(), (int64_t)obstacles_.size())); return; }
        fs_obstacle_shape sh{};
        sh.kind = (int)obstacleShape;
        sh.center_x = obstaclePositionX * n; sh.center_y = obstaclePositionY * n; sh.center_z = obstaclePositionZ * currentDepth_;
        sh.radius = obstacleRadius * n; sh.width = obstacleWidth * n; sh.height = obstacleHeight * n;
        sh.depth = obstacleWidth * currentDepth_;
        sh.seed_x = (int)std::nearbyint(obstaclePositionX * n); sh.seed_y = (int)std::nearbyint(obstaclePositionY * n);
        sh.seed_z = currentDepth_ > 1 ? (int)std::nearbyint(obstaclePositionZ * currentDepth_) : 0;
        check(fs_build_obstacles(solver_, &sh, &obstacleCells_));
        check(fs_get_obstacles(solver_, obstacles_.data(), (int64_t)obstacles_.size()));
    }
    int64_t ObstacleCells() const { return obstacleCells_; }

    // AddForceToArea, :452-483: velocity with linear fall-off, density inside 0.3 r -- ONE batched native call
    void AddForceToArea(float centerX, float centerY, float forceX, float forceY, float radius) {
        const int n = currentSize_;
        auto clampi = [n](int v) { return v < 0 ? 0 : (v > n - 1 ? n - 1 : v); };
        std::vector<float> xs, ys, zs, ds, ax, ay;
        for (int x = clampi((int)(centerX - radius)); x <= clampi((int)(centerX + radius)); x++)
            for (int y = clampi((int)(centerY - radius)); y <= clampi((int)(centerY + radius)); y++) {
                const float dx = (float)x - centerX, dy = (float)y - centerY;
                const float distance = (float)std::sqrt((double)(dx * dx + dy * dy)); // Vector2.Distance
                if (distance > radius) continue;
                const float falloff = 1 - (distance / radius);
                xs.push_back((float)x); ys.push_back((float)y); zs.push_back(sourcePositionZ * currentDepth_);
                ax.push_back(forceX * falloff); ay.push_back(forceY * falloff);
                ds.push_back(distance < radius * 0.3f ? sourceStrength * falloff : 0.0f);
            }
        if (xs.empty()) return;
        check(fs_add_source_cells(solver_, (int64_t)xs.size(), xs.data(), ys.data(), zs.data(), ds.data(), ax.data(), ay.data(), nullptr));
    }

    // UpdateVisualization, :755-853: UpdateVisualizationJob on the device; returns currentSize^2 RGBA floats (Color[] layout)
    fs_vis_params VisParams() const {
        fs_vis_params v{};
        const float white[4] = {1, 1, 1, 1}, blue[4] = {0, 0, 1, 1}, green[4] = {0, 1, 0, 1}, red[4] = {1, 0, 0, 1},
                    gray[4] = {0.5f, 0.5f, 0.5f, 1}, yellow[4] = {1, 0.92156863f, 0.01568628f, 1};
        auto set = [](float *d, const float *c) { for (int i = 0; i < 4; i++) d[i] = c[i]; };
        v.color_mode = colorMode; v.visualize_source_position = visualizeSourcePosition; v.enable_custom_source = enableCustomSource;
        v.source_x = sourcePositionX * currentSize_; v.source_y = sourcePositionY * currentSize_; v.visual_marker_radius = 3.0f; // :805-807
        v.colour_intensity = colourIntensity;
        v.medium_density_threshold = mediumDensityThreshold; v.high_density_threshold = highDensityThreshold;
        v.low_pressure_threshold = lowPressureThreshold; v.high_pressure_threshold = highPressureThreshold;
        set(v.fluid_color, white); set(v.obstacle_color, gray); set(v.source_position_color, yellow);
        set(v.low_density_color, blue); set(v.medium_density_color, green); set(v.high_density_color, red);
        set(v.low_pressure_color, blue); set(v.neutral_pressure_color, white); set(v.high_pressure_color, red);
        v.gradient_key_count = 2; set(v.gradient_colors[0], blue); set(v.gradient_colors[1], red);           // Start(): :188-203
        v.gradient_times[0] = 0.0f; v.gradient_times[1] = 1.0f;
        v.z_slice = ViewSlice();
        return v;
    }
    std::vector<float> UpdateVisualization() {
        const fs_vis_params v = VisParams();
        std::vector<float> rgba((size_t)currentSize_ * currentSize_ * 4);
        check(fs_render_rgba(solver_, &v, rgba.data(), (int64_t)rgba.size()));
        return rgba;
    }

    // DrawStreamlines, :886-959: the glyph jobs run on the device (fs_streamlines); Bresenham drawing (:1765-1849) here.
    // Returns a currentSize^2 coverage mask (1 where streamlineColor is painted).
    std::vector<uint8_t> DrawStreamlines() {
        const int n = currentSize_;
        const int skip = std::max(1, n / (streamlineDensity * 10));
        const int64_t count = (int64_t)(n / skip) * (n / skip);
        std::vector<float> seg((size_t)count * 4);
        check(fs_streamlines(solver_, skip, streamlineScale, ViewSlice(), seg.data(), count));
        std::vector<uint8_t> tex((size_t)n * n, 0);
        const int halfThick = (int)std::floor(streamlineThickness / 2);
        for (int64_t i = 0; i < count; i++) {
            if (seg[4 * i] < 0) continue;
            int x0 = (int)seg[4 * i], y0 = (int)seg[4 * i + 1], x1 = (int)std::nearbyint(seg[4 * i + 2]), y1 = (int)std::nearbyint(seg[4 * i + 3]);
            const bool steep = std::abs(y1 - y0) > std::abs(x1 - x0);
            if (steep) { std::swap(x0, y0); std::swap(x1, y1); }
            if (x0 > x1) { std::swap(x0, x1); std::swap(y0, y1); }
            const int dx = x1 - x0, dy = std::abs(y1 - y0), ystep = y0 < y1 ? 1 : -1;
            int error = dx / 2, y = y0;
            for (int x = x0; x <= x1; x++) {
                for (int tx = -halfThick; tx <= halfThick; tx++)
                    for (int ty = -halfThick; ty <= halfThick; ty++) {
                        const int px = steep ? y + tx : x + tx, py = steep ? x + ty : y + ty;
                        if (px >= 0 && px < n && py >= 0 && py < n) tex[px + (size_t)py * n] = 1;
                    }
                error -= dy;
                if (error < 0) { y += ystep; error += dx; }
            }
        }
        return tex;
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
    int64_t obstacleCells_ = 0;
    int ViewSlice() const {
        if (currentDepth_ <= 1) return 0;
        const int z = (int)std::nearbyint(viewSliceZ * currentDepth_);
        return z < 0 ? 0 : (z > currentDepth_ - 1 ? currentDepth_ - 1 : z);
    }

    static float clamp01(float v) { return v < 0 ? 0 : (v > 1 ? 1 : v); }
    void check(int rc) const {
        if (rc != FS_OK) throw std::runtime_error(std::string("fluidsolver: ") + fs_last_error(solver_));
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
