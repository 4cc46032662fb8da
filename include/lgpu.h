/* include/lgpu.h — C ABI of the B200 (sm_100a) particle-step library, liblgpu.so.
 *
 * This is the thin device boundary that the host-side Lustrine API (Simulation /
 * simulate_fun / LustrineWrapper, see lustrine_b200/host/) calls into.  Everything
 * is extern "C", plain pointers and sizes, int status codes (0 = ok).  Host buffers
 * are caller-owned; vec3 data is the reference's AoS layout (glm::vec3 = 3 packed
 * floats), exactly what Lustrine::Simulation::positions / velocities hold
 * (reference src/Simulation.hpp:188-193,173).
 *
 * What each entry point replaces in the reference (paths relative to the reference root):
 *   lgpu_create            grid/kernel constants of init_simulation, src/Lustrine.cpp:251-284
 *   lgpu_upload_sand       the sand fill of init_simulation, src/Lustrine.cpp:144-172 (+ velocities :241)
 *   lgpu_upload_solids     the solid fill, src/Lustrine.cpp:174-235, and the static grid cache,
 *                          src/neighbors/Neighbors.cpp:266-272
 *   lgpu_append_sand       the particle-source spawn, src/Lustrine.cpp:771-782
 *   lgpu_step_fluid        Lustrine::simulate_fluid, src/Simulate.cpp:27-115
 *                          (find_neighbors_uniform_grid, src/neighbors/Neighbors.cpp:366-450;
 *                           cubic_kernel / cubic_kernel_grad, src/Kernels.cpp:6-41; s_coor, src/Simulate.cpp:7-9)
 *   lgpu_step_sand         Lustrine::simulate_sand / simulate_sand_credits, src/Simulate.cpp:156-325,327-510
 *                          (find_neighbors_uniform_grid_v1, src/neighbors/Neighbors.cpp:262-364;
 *                           Sorting::counting_sort, src/neighbors/Sorting.cpp:11-33)
 *   lgpu_download_sand     the caller's direct reads of simulation.positions and
 *                          Wrapper::simulation_bind_positions_copy, src/LustrineWrapper.cpp:377-379
 *   lgpu_cell_count        query_cell_num_particles, src/Lustrine.cpp:864-914
 *   lgpu_remove_in_cells   the sink pass of Lustrine::simulate, src/Lustrine.cpp:806-836
 *   lgpu_aabb_first_k      the player-AABB scan of set_particles_box_colliders_positions,
 *                          src/BulletPhysics.cpp:602-652
 *   lgpu_eval_kernel       W / gradW / poly6 / spiky / s_coor tables, src/Kernels.cpp, src/Simulate.cpp:7-9
 *   lgpu_counting_sort     Sorting::counting_sort on caller keys (known-answer test of
 *                          experiments/unit_tests/main.cpp:46-89)
 *
 * There is NO CPU fallback behind this interface: every call needs a CUDA device and
 * fails with LGPU_ERR_CUDA otherwise.
 */
#ifndef LGPU_H
#define LGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define LGPU_API __attribute__((visibility("default")))
#else
#define LGPU_API
#endif

#define LGPU_OK 0
#define LGPU_ERR_CUDA 1      /* a CUDA runtime call failed; see lgpu_last_error() */
#define LGPU_ERR_ARG 2       /* bad argument */
#define LGPU_ERR_CAPACITY 3  /* more particles than the context was created for */

typedef struct lgpu_ctx lgpu_ctx;

typedef struct lgpu_config {
    int domain[3];             /* SimulationParameters::X,Y,Z (ints, src/Simulation.hpp:83-92) */
    float particle_radius;     /* SimulationParameters::particleRadius */
    float particle_diameter;   /* SimulationParameters::particleDiameter */
    float kernel_radius_scale; /* 3.1f in init_simulation (src/Lustrine.cpp:253), argument of
                                  init_simulation_extra_parameters (:545) */
    int capacity_sand;         /* max sand particles ever resident (owned + ghosts) */
    int capacity_solid;        /* max solid (static boundary) particles */
    int max_neighbors;         /* width of the per-particle neighbour table; 0 = default (32).
                                  Longer lists fall back to a stencil re-walk (always correct). */
    int device;                /* CUDA device ordinal, -1 = current device */
    /* spatial slab owned by this context, in grid-cell x coordinates [x_lo, x_hi);
       x_lo = x_hi = 0 means the whole grid (single GPU).  See lgpu_halo_* below. */
    int slab_x_lo, slab_x_hi;
    void* stream;              /* cudaStream_t to launch on, NULL = a private non-blocking stream */
    int halo_capacity;         /* slab mode: max particles per neighbour and substep in flight (migrants, ghost
                                  copies); 0 = max(65536, capacity_sand / 4).  Must be equal on all slabs. */
    int slab_ghost_columns;    /* slab mode: width of the ghost layer in cell columns, 1 (0 = default) or 2.  With 2 the
                                  lambdas of the inner ghost column are computed locally (their neighbours are all
                                  present), so a fluid substep needs K - 1 ghost refreshes instead of 2K - 1.  Must be
                                  equal on all slabs; every slab must own at least that many columns. */
    int slab_guard_columns;    /* slab mode, a plan CROPPED to the occupied columns (the first slab starts at x_lo > 0 or the
                                  last one ends at x_hi < grid x: the empty part of the domain costs no cells): the particles
                                  in the outer `slab_guard_columns` owned columns of such an open side are counted every
                                  substep (lgpu_slab_edge) so that the host re-plans before any particle leaves the planned
                                  columns (which fails the step with LGPU_ERR_CAPACITY).  0 = not counted. */
} lgpu_config;

/* Per-step scalars.  The reference re-reads them from the public Simulation struct on
 * every call (demos mutate them live, experiments/fluid/fluid.cpp:177-189), so they are
 * passed by value on every step.  Defaults: lgpu_default_step_params (src/Simulation.hpp:147-171,229-232;
 * src/Simulate.cpp:159-163). */
typedef struct lgpu_step_params {
    float dt;
    float gravity[3];
    float rest_density, mass, relaxation_epsilon;
    float s_corr_dq, s_corr_k, s_corr_n;
    int iterations;            /* solver iterations per substep. fluid: 1 in the reference; sand: 4 */
    int literal_lambda_index;  /* 1 = reference behaviour lambdas[loop counter] (src/Simulate.cpp:97);
                                  0 = lambdas[neighbour] */
    int exact_math;            /* 1 = every fp32 operation separately rounded in the reference's order
                                  (parity mode); 0 = FMA contraction + fast reciprocal/rsqrt (throughput mode,
                                  same results to ~1e-6 relative) */
    int sph_kernel;            /* 0 = cubic spline (what the reference wires, src/Lustrine.cpp:247-248);
                                  1 = poly6 density + spiky gradient (src/Kernels.cpp:43-67) */
    /* sand only */
    float player_position[3];  /* Bullet::Simulation::player_position, published by the host rigid-body step */
    int attract_flag, blow_flag, prev_attract_flag;
    float attract_radius, blow_radius, attract_coeff, blow_coeff;
    float collision_coeff, friction_coeff, mu_s, mu_k;
    int credits;               /* 1 = simulate_sand_credits semantics (bit 1 of the flag word = no gravity) */
} lgpu_step_params;

LGPU_API void lgpu_default_step_params(lgpu_step_params* p);

LGPU_API int lgpu_create(const lgpu_config* cfg, lgpu_ctx** out);
LGPU_API void lgpu_destroy(lgpu_ctx* ctx);
LGPU_API const char* lgpu_last_error(void);

/* Grid constants as the reference computes them (src/Lustrine.cpp:253-266). */
typedef struct lgpu_grid_info {
    int grid[3];
    int num_cells;
    float cell_size, kernel_radius, cubic_k, cubic_l;
} lgpu_grid_info;
LGPU_API int lgpu_get_grid(const lgpu_ctx* ctx, lgpu_grid_info* out);

/* State transfer.  pos/vel are n*3 floats (AoS), flags n ints (Simulation::attracted); vel and
 * flags may be NULL (zeros).  After upload, particle i sits in reference slot i. */
LGPU_API int lgpu_upload_sand(lgpu_ctx* ctx, int n, const float* pos, const float* vel, const int* flags);
LGPU_API int lgpu_upload_solids(lgpu_ctx* ctx, int n, const float* pos);
LGPU_API int lgpu_append_sand(lgpu_ctx* ctx, int n, const float* pos, const float* vel, const int* flags);
/* Downloads the sand state in the reference's storage order: after a sand step that is the
 * stable cell-sorted order the reference itself permutes its arrays into
 * (src/neighbors/Neighbors.cpp:288-304); after a fluid step it is the unchanged upload order.
 * Any pointer may be NULL. */
LGPU_API int lgpu_download_sand(lgpu_ctx* ctx, float* pos, float* vel, int* flags);
/* The same, with the positions delivered to TWO host arrays by the copy engine (the reference leaves
 * positions_star == positions after a step, src/Simulate.cpp:111,318: the drop-in fills both without a host
 * memcpy).  pos2 may be NULL. */
LGPU_API int lgpu_download_sand2(lgpu_ctx* ctx, float* pos, float* pos2, float* vel, int* flags);
/* Page-locks caller-owned host memory (the host arrays of Lustrine::Simulation, src/Lustrine.cpp:129-141) so that
 * uploads and downloads run at the PCIe rate instead of through the driver's bounce buffers.  Optional: every
 * transfer also works with pageable memory.  lgpu_host_is_pinned: 1 if ptr lies in page-locked memory, else 0. */
LGPU_API int lgpu_host_register(void* ptr, size_t bytes);
LGPU_API int lgpu_host_unregister(void* ptr);
LGPU_API int lgpu_host_is_pinned(const void* ptr);
LGPU_API int lgpu_num_sand(const lgpu_ctx* ctx);
LGPU_API int lgpu_num_solids(const lgpu_ctx* ctx);

/* One substep: predict, grid build, `iterations` solver iterations, velocity commit.
 * Asynchronous on the context's stream; lgpu_sync (or any download) waits for it. */
LGPU_API int lgpu_step_fluid(lgpu_ctx* ctx, const lgpu_step_params* p);
LGPU_API int lgpu_step_sand(lgpu_ctx* ctx, const lgpu_step_params* p);
LGPU_API int lgpu_sync(lgpu_ctx* ctx);

/* Device time of the last `lgpu_step_*` call (CUDA events on the context's stream).  phase 0 =
 * the whole step (always available).  With phase timing on, also the sum over the launches of
 * one kind: 1 predict+key+histogram, 2 cell prefix sum, 3 scatter+stable reorder, 4 neighbour
 * table, 6 density/lambda kernels, 7 delta-p (fluid) / contact (sand) kernels, 5 = 6 + 7,
 * 8 halo refresh messages of slab mode. */
LGPU_API int lgpu_last_step_ms(lgpu_ctx* ctx, int phase, float* ms);
/* Kernels launched by this context since creation (for bench.py's gpu_launches). */
LGPU_API long lgpu_launch_count(const lgpu_ctx* ctx);
/* Enables per-phase event timing (adds event records to every step). */
LGPU_API int lgpu_set_phase_timing(lgpu_ctx* ctx, int on);
/* 1 = run each step as a CUDA graph: the launches of the substep are captured from the context's stream
 * (cudaStreamBeginCapture) and replayed with one cudaGraphLaunch while the particle count, the mode and the
 * step parameters are unchanged; otherwise the instantiated graph is updated in place or rebuilt.  Not
 * available in slab mode (LGPU_ERR_ARG).  lgpu_graph_stats: how often a step was captured / replayed. */
LGPU_API int lgpu_set_use_graph(lgpu_ctx* ctx, int on);
LGPU_API int lgpu_graph_stats(const lgpu_ctx* ctx, long* captures, long* replays);
/* Tuning / test hook: capacity (in particles) of the shared-memory stage a thread block may use for
 * its neighbourhood; blocks that need more read their neighbours through L1/L2 instead ("virtual
 * slots").  Clamped to the compiled maximum; results do not depend on it. */
LGPU_API int lgpu_set_stage_slots(lgpu_ctx* ctx, int slots);
/* Test hook: 1 = run the fast-arithmetic fluid step with the generic solver kernels (the ones that also
 * serve exact_math / poly6 / literal_lambda_index) instead of the specialised k_fluid_*_fast pair. */
LGPU_API int lgpu_set_generic_kernels(lgpu_ctx* ctx, int on);

/* ---- spatial slabs: one context per GPU owns the cell columns [slab_x_lo, slab_x_hi) ----
 * No reference counterpart (the reference is single-threaded); the per-particle arithmetic is the
 * single-GPU one.  Wiring: every context exports its peer-visible arena (one CUDA IPC handle, or the
 * plain pointer for contexts of the same process) and connects to the arenas of its left (side 0)
 * and right (side 1) neighbours.  Then lgpu_step_fluid / lgpu_step_sand run the slab protocol:
 * migration + ghost copies after predict, ghost refresh after every solver pass, all written by the
 * sender's kernels into the receiver's memory over NVLink.  literal_lambda_index must be 0.
 * Particles carry caller-given global ids (lgpu_slab_upload / lgpu_slab_download).
 * lgpu_slab_step_begin/_end split a substep so that ONE host thread can drive several contexts
 * (tests on a single GPU): call _begin on all of them, then _end on all of them. */
/* Host-side planning for a C / C++ host (lustrine_b200/slabs.py holds the same logic for the Python host of the tests and
 * of bench.py; tests/test_slabs.py checks the two against each other).  No device is touched.
 *   column_hist[grid_x]  particles per global cell column (x / cell_size, truncated) — what the ranks obtain with one
 *                        all-reduce of an int array
 *   lgpu_plan_slabs      boundaries bounds[0..world]: rank k owns the columns [bounds[k], bounds[k + 1]), equal particle
 *                        counts as far as whole columns allow, every rank at least min_columns wide (>= the ghost layer).
 *                        margin < 0: the plan covers [0, grid_x); margin >= 0: it is cropped to the occupied columns plus
 *                        `margin` free columns on either side (create the contexts with slab_guard_columns =
 *                        lgpu_slab_guard_columns(margin) and re-plan when lgpu_slab_edge reports particles near the end).
 *   lgpu_slab_capacity   capacity_sand for every context of the plan (equal on all ranks): owned + twice the ghost
 *                        columns of the fullest slab, times `factor` (head room until the next re-plan), + 4096.
 * Returns LGPU_ERR_ARG if grid_x cannot give every rank min_columns columns. */
LGPU_API int lgpu_plan_slabs(const long long* column_hist, int grid_x, int world, int min_columns, int margin, int* bounds);
LGPU_API int lgpu_slab_guard_columns(int margin);
LGPU_API long long lgpu_slab_capacity(const long long* column_hist, int grid_x, const int* bounds, int world, int ghost_columns, double factor);
LGPU_API int lgpu_slab_export(lgpu_ctx* ctx, unsigned char handle[64], void** local_ptr, size_t* bytes);
LGPU_API int lgpu_slab_connect(lgpu_ctx* ctx, int side, const unsigned char handle[64], void* same_process_ptr);
/* out: x_lo, x_hi, local grid X, local cells, owned particles, ghost particles, halo capacity, x offset */
LGPU_API int lgpu_slab_info(const lgpu_ctx* ctx, int out[8]);
/* out[0] = particles of the last substep inside the guard columns of an open side of a cropped plan (see
 * lgpu_config::slab_guard_columns; > 0: time to re-plan), out[1] = 1 if this slab has an open side */
LGPU_API int lgpu_slab_edge(const lgpu_ctx* ctx, int out[2]);
LGPU_API int lgpu_slab_upload(lgpu_ctx* ctx, int n, const float* pos, const float* vel, const int* flags, const int* ids);
LGPU_API int lgpu_slab_download(lgpu_ctx* ctx, float* pos, float* vel, int* flags, int* ids, int* n_out);
LGPU_API int lgpu_slab_step_begin(lgpu_ctx* ctx, const lgpu_step_params* p, int mode /* 1 fluid, 2 sand */);
LGPU_API int lgpu_slab_step_end(lgpu_ctx* ctx);

/* Grid queries on the device grid of the last step ("next" rows). */
LGPU_API int lgpu_cell_count(lgpu_ctx* ctx, const int lo[3], const int hi[3], int include_solid, int* count);
LGPU_API int lgpu_remove_in_cells(lgpu_ctx* ctx, const int* cell_ids, int n_cells, int* removed);
LGPU_API int lgpu_aabb_first_k(lgpu_ctx* ctx, const float center[3], const float half[3], int k, float* out_pos, int* out_n);

/* Stage dumps for the parity tests (all in sorted-slot order of the last grid build unless noted).
 *   LGPU_DUMP_KEYS       int[n]   cell id per sorted slot
 *   LGPU_DUMP_PERM       int[n]   reference slot (before this step's reorder) of each sorted slot
 *   LGPU_DUMP_ORIG       int[n]   reference slot (current) of each sorted slot
 *   LGPU_DUMP_NBR_COUNT  int[n]   neighbour-list length (self entries excluded)
 *   LGPU_DUMP_NBR        int[sum] neighbour lists, concatenated, list order; sand = sorted slot,
 *                                  solid = n + (solid upload index)
 *   LGPU_DUMP_DENSITY    float[n], LGPU_DUMP_LAMBDA float[n]
 *   LGPU_DUMP_PSTAR      float[3n] current predicted positions
 *   LGPU_DUMP_CELL_START int[num_cells+1]
 *   LGPU_DUMP_COUNTERS   long[4]: key violations, neighbour-table overflows (fallback walks), 0, 0
 */
enum {
    LGPU_DUMP_KEYS = 0, LGPU_DUMP_PERM = 1, LGPU_DUMP_ORIG = 2, LGPU_DUMP_NBR_COUNT = 3, LGPU_DUMP_NBR = 4,
    LGPU_DUMP_DENSITY = 5, LGPU_DUMP_LAMBDA = 6, LGPU_DUMP_PSTAR = 7, LGPU_DUMP_CELL_START = 8,
    LGPU_DUMP_COUNTERS = 9
};
LGPU_API int lgpu_dump(lgpu_ctx* ctx, int what, void* out, size_t out_bytes);

/* Function tables evaluated on the device with the context's constants.
 *   which: 0 cubic W(r) (n floats in, n out), 1 cubic gradW (3n in, 3n out), 2 poly6(r),
 *          3 spiky (3n in, 3n out), 4 s_coor(r);  exact_math as in lgpu_step_params. */
LGPU_API int lgpu_eval_kernel(lgpu_ctx* ctx, const lgpu_step_params* p, int which, const float* in, int n, float* out);

/* Stable counting sort of caller keys in [0, num_cells): sorted[p] = index of the p-th element. */
LGPU_API int lgpu_counting_sort(const int* keys, int n, int num_cells, int* sorted, int device);

#ifdef __cplusplus
}
#endif
#endif /* LGPU_H */
