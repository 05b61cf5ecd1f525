/* fluid_oracle.h -- TEST INFRASTRUCTURE (CPU oracle), see fluid_oracle.c.  Not shipped, never
 * linked into libfluidsolver.so. */
#ifndef FLUID_ORACLE_H
#define FLUID_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fo_state {
    int nx, ny, nz;          /* nz == 1: the reference's 2D solver */
    int iters_diffuse;       /* 20 in the reference (FluidSim.cs:1310, :1378) */
    int iters_pressure;      /* 20 in the reference (FluidSim.cs:1594) */
    int red_black;           /* 0 = Jacobi (reference), 1 = red-black pressure solve */
    int enable_obstacle;     /* FluidSim.cs:97, :567 */
    float cell_size;         /* FluidSim.cs:219 */
    float raw_viscosity;     /* FluidSim.cs:664 uses the unscaled viscosity */
    float *density, *vx, *vy, *vz, *vx0, *vy0, *vz0, *pressure; /* caller-owned, nx*ny*nz each */
    const uint8_t *obstacles;                                   /* 1 byte per cell */
} fo_state;

/* parameters of UpdateVisualizationJob (FluidSim.cs:1859-1886, :799-829); same field order as fs_vis_params */
typedef struct fo_vis_params {
    int32_t color_mode, visualize_source_position, enable_custom_source, gradient_key_count, z_slice;
    float source_x, source_y, visual_marker_radius, colour_intensity;
    float medium_density_threshold, high_density_threshold, low_pressure_threshold, high_pressure_threshold;
    float fluid_color[4], obstacle_color[4], source_position_color[4];
    float low_density_color[4], medium_density_color[4], high_density_color[4];
    float low_pressure_color[4], neutral_pressure_color[4], high_pressure_color[4];
    float gradient_colors[8][4];
    float gradient_times[8];
} fo_vis_params;
void fo_visualize(int nx, int ny, const float *density, const float *pressure, const uint8_t *obstacles,
                  const fo_vis_params *vp, float *out);

void fo_streamlines(int nx, int ny, int skip, float streamlineScale, const float *velocX, const float *velocY,
                    const uint8_t *obstacles, float *out);

void fo_set_bnd(int nx, int ny, int nz, int b, float *x, const uint8_t *obs);
void fo_diffuse_coeffs(int n, float diff, float dt, float *a, float *c);
void fo_diffuse_smooth(int nx, int ny, int nz, int b, float *x, const float *x0, float a, float c,
                       const uint8_t *obs, int iters);
void fo_lin_solve(int nx, int ny, int nz, int b, float *x, const float *x0, float a, float c,
                  const uint8_t *obs, int iters);
void fo_lin_solve_rb(int nx, int ny, int nz, int b, float *x, const float *x0, float a, float c,
                     const uint8_t *obs, int iters);
void fo_diffuse(int nx, int ny, int nz, int b, float *x, const float *x0, float diff, float dt,
                const uint8_t *obs, int iters);
void fo_divergence(int nx, int ny, int nz, float *div, const float *vx, const float *vy, const float *vz,
                   const uint8_t *obs);
void fo_subtract_gradient(int nx, int ny, int nz, float *vx, float *vy, float *vz, const float *p,
                          const uint8_t *obs);
void fo_project(int nx, int ny, int nz, float *vx, float *vy, float *vz, float *p, const uint8_t *obs,
                int iters, int red_black);
void fo_advect(int nx, int ny, int nz, int b, float *d, const float *d0, const float *vx, const float *vy,
               const float *vz, float dt, const uint8_t *obs);
void fo_enforce_obstacles(int nx, int ny, int nz, float *vx, float *vy, float *vz, const uint8_t *obs,
                          float cell, float rawvisc);
long long fo_cell_index(int nx, int ny, int nz, float x, float y, float z);
void fo_step(fo_state *s, float dt, float visc, float diff);
/* Same step, same bits, with the reference's EXECUTION structure (flat static-64 job loops, single-threaded
 * boundary scans, per-call allocate-and-copy): ref_faithful3d.c, the "port-faithful" CPU baseline. */
void rf_step(fo_state *s, float dt, float visc, float diff);
void fo_metrics(const fo_state *s, float *mean_density, float *max_speed);

/* literal 2D restatement (ref2d.c) */
void r2_boundary(int size, int b, float *x, const uint8_t *obstacles);
void r2_diffuse_with_jobs(int size, int b, float *x, const float *x0, float diff, float dt,
                          const uint8_t *obstacles, int iters);
void r2_linear_solve_with_jobs(int size, int b, float *x, const float *x0, float a, float c,
                               const uint8_t *obstacles, int iters);
void r2_diffuse(int size, int b, float *x, const float *x0, float diff, float dt, const uint8_t *obstacles,
                int iters);
void r2_project_with_jobs(int size, float *velocX, float *velocY, float *p, const uint8_t *obstacles,
                          int iters);
void r2_advect_with_jobs(int size, int b, float *d, const float *d0, const float *velocX, const float *velocY,
                         float dt, const uint8_t *obstacles);
void r2_enforce_obstacles(int size, float *velocityX, float *velocityY, const uint8_t *obstacles,
                          float cellSize, float viscosity);
void r2_simulate(int size, float *density, float *velocityX, float *velocityY, float *velocityX0,
                 float *velocityY0, float *pressure, const uint8_t *obstacles, float dt, float visc, float diff,
                 int enableObstacle, float cellSize, float rawViscosity, int iters);
void r2_add_density(int size, float *density, float x, float y, float amount);
void r2_add_velocity(int size, float *velocityX, float *velocityY, float x, float y, float ax, float ay);

#ifdef __cplusplus
}
#endif
#endif
