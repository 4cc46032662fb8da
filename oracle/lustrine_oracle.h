/* oracle/lustrine_oracle.h — TEST INFRASTRUCTURE ONLY (see lustrine_oracle.c). */
#ifndef LUSTRINE_ORACLE_H
#define LUSTRINE_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lo_sim {
    /* geometry — src/Lustrine.cpp:105-107,251-266 */
    float domainX, domainY, domainZ;
    float particleRadius, particleDiameter;
    float kernelRadius, kernelFactor, cell_size;
    float cubic_kernel_k, cubic_kernel_l;
    int gridX, gridY, gridZ, num_grid_cells;
    /* scalars — src/Simulation.hpp:147-171,229-232 */
    float rest_density, mass, relaxation_epsilon;
    float s_corr_dq, s_corr_k, s_corr_n;
    float gravity[3];
    float time_step;
    float attract_radius, blow_radius, attract_coeff, blow_coeff;
    float player_position[3];
    int attract_flag, blow_flag, prev_attract_flag;
    /* particles: sand occupies [0, n_sand), solids [n_sand, n_sand + n_solid).
       (The reference keeps solids at [total_allocated - n_solid, total_allocated);
       only the relative order matters for any result.) */
    int n_sand, n_solid, capacity;
    float *positions, *positions_star, *positions_tmp, *velocities; /* xyz triples */
    int* attracted;
    float *lambdas, *densities;
    /* neighbour lists in the reference's list order, CSR */
    long* nbr_offsets; /* n_sand + 1 */
    int* nbr;
    long nbr_capacity;
    /* stage outputs */
    int* keys;          /* cell id per sand particle at the last grid build */
    int* sorted_index;  /* v1: permutation applied at the last grid build (sorted slot -> previous index) */
    /* scratch */
    int *cell_counts, *cell_start, *cell_items;
    int *solid_cell_start, *solid_cell_items;
    int solid_grid_built;
    float* scratch3;
    int* scratchi;
    long violations; /* keys outside [0, num_grid_cells) seen (the reference has no clamp, F10) */
    int sph_kernel;  /* which functions Simulation::W / gradW point at: 0 = cubic_kernel / cubic_kernel_grad (what
                        init_simulation wires, src/Lustrine.cpp:247-248), 1 = poly6_kernel(float) / spiky_kernel
                        (src/Kernels.cpp:43-67; never wired by the reference itself) */
} lo_sim;

lo_sim* lo_create(int X, int Y, int Z, float radius, float diameter, float kernel_radius_scale,
                  int sand_capacity, int n_solid);
void lo_destroy(lo_sim* s);
void lo_set_sand(lo_sim* s, int n, const float* pos, const float* vel, const int* attracted);
void lo_set_solid(lo_sim* s, const float* pos);

int lo_cell_id(const lo_sim* s, float x, float y, float z);
void lo_counting_sort(int* counts, const int* keys, long n, long num_cells, int* sorted);

float lo_cubic_kernel(const lo_sim* s, float r);
void lo_cubic_kernel_grad(const lo_sim* s, const float r[3], float out[3]);
float lo_poly6_kernel(const lo_sim* s, float r);
void lo_spiky_kernel(const lo_sim* s, const float r[3], float out[3]);
float lo_s_coor(const lo_sim* s, float rl);

void lo_find_neighbors_v0(lo_sim* s); /* fluid: no reorder, lists include self once */
void lo_find_neighbors_v1(lo_sim* s); /* sand: stable reorder, half stencil, lists include self twice */

/* staged solver entry points */
void lo_fluid_predict(lo_sim* s, float dt);
void lo_fluid_lambda(lo_sim* s);
void lo_fluid_deltap(lo_sim* s, int jacobi, int literal_lambda_index);
void lo_fluid_commit(lo_sim* s);
/* whole step: jacobi=0,K=1 is the literal reference (Gauss-Seidel delta-p) */
void lo_step_fluid(lo_sim* s, float dt, int iterations, int jacobi, int literal_lambda_index);

void lo_sand_predict(lo_sim* s, float dt, int credits);
void lo_sand_iteration(lo_sim* s, int credits);
void lo_sand_commit(lo_sim* s);
void lo_step_sand(lo_sim* s, float dt, int iterations, int credits);

#ifdef __cplusplus
}
#endif
#endif
