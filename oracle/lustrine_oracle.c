/* oracle/lustrine_oracle.c — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement ("port") of the particle step of GrapixLeGrand/Lustrine, written
 * from the reference's behaviour, each function citing the reference file:line it
 * follows (paths relative to the reference root).  It is the checker for the CUDA path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product never calls into this file.
 *
 * Pinning: the reference stores no golden vectors for this path, so this
 * port is pinned against the reference ITSELF: tests/test_oracle_vs_reference.py runs
 * oracle/_ref/libref_lit.so (the unmodified reference sources compiled with
 * -O2 -ffp-contract=off) on the same seeded inputs and requires bit-identical keys,
 * permutations, neighbour lists (order included), lambdas, positions and velocities;
 * tests/golden/ holds vectors generated from that library by tests/golden/make_golden.py.
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off (no fast-math, no FMA contraction) so that
 * every fp32 operation is a separately rounded IEEE operation, like the parity build of
 * the reference.  All arithmetic is fp32 unless a comment says double.
 */
#include "lustrine_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- glm 0.9.9.8 vec3 semantics (thirdparty/glm-0.9.9.8/glm/detail/func_geometric.inl:48-55,
 * :8-14, :82-90; func_exponential.inl:136-139) ---- */
typedef struct { float x, y, z; } v3;
static inline v3 v3_make(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 v3_load(const float* p) { v3 r = {p[0], p[1], p[2]}; return r; }
static inline void v3_store(float* p, v3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_scale(v3 a, float s) { return v3_make(a.x * s, a.y * s, a.z * s); }
static inline v3 v3_div(v3 a, float s) { return v3_make(a.x / s, a.y / s, a.z / s); }
static inline v3 v3_neg(v3 a) { return v3_make(-a.x, -a.y, -a.z); }
static inline float v3_dot(v3 a, v3 b) { float tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z; return (tx + ty) + tz; }
static inline float v3_length(v3 a) { return sqrtf(v3_dot(a, a)); }
static inline v3 v3_normalize(v3 a) { return v3_scale(a, 1.0f / sqrtf(v3_dot(a, a))); }
static inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

static const double LO_PI = 3.14159265358979323846; /* src/Lustrine.cpp:16 */

lo_sim* lo_create(int X, int Y, int Z, float radius, float diameter, float kernel_radius_scale,
                  int sand_capacity, int n_solid) {
    lo_sim* s = (lo_sim*)calloc(1, sizeof(lo_sim));
    /* src/Lustrine.cpp:105-107 (ints assigned to floats) */
    s->domainX = (float)X; s->domainY = (float)Y; s->domainZ = (float)Z;
    s->particleRadius = radius; s->particleDiameter = diameter;
    /* src/Lustrine.cpp:253-254 (3.1f) / :545 (kernel_radius_scale) */
    s->kernelRadius = kernel_radius_scale * radius;
    s->cell_size = 1.0f * s->kernelRadius;
    s->kernelFactor = 0.5f;
    /* src/Lustrine.cpp:257-259: std::pow(float,int) promotes to double; k,l computed in double */
    float h3 = (float)pow((double)s->kernelRadius, 3.0);
    s->cubic_kernel_k = (float)(8.0f / (LO_PI * h3));
    s->cubic_kernel_l = (float)(48.0f / (LO_PI * h3));
    /* src/Lustrine.cpp:261-266 */
    s->gridX = (int)(s->domainX / s->cell_size) + 1;
    s->gridY = (int)(s->domainY / s->cell_size) + 1;
    s->gridZ = (int)(s->domainZ / s->cell_size) + 1;
    s->num_grid_cells = s->gridX * s->gridY * s->gridZ;
    /* src/Simulation.hpp:147-171,229-232 */
    s->rest_density = 24.0f; s->mass = 5.0f; s->relaxation_epsilon = 10.0f;
    s->s_corr_dq = 0.5f; s->s_corr_k = 1.0f; s->s_corr_n = 4.0f;
    s->gravity[0] = 0.0f; s->gravity[1] = -10.0f; s->gravity[2] = 0.0f;
    s->time_step = 0.01f;
    s->attract_radius = 1.5f; s->blow_radius = 2.0f; s->attract_coeff = 1000.0f; s->blow_coeff = 500.0f;
    s->n_sand = 0; s->n_solid = n_solid; s->capacity = sand_capacity;
    size_t tot = (size_t)sand_capacity + (size_t)n_solid;
    s->positions = (float*)calloc(3 * tot, sizeof(float));
    s->positions_star = (float*)calloc(3 * tot, sizeof(float));
    s->positions_tmp = (float*)calloc(3 * tot, sizeof(float));
    s->velocities = (float*)calloc(3 * tot, sizeof(float));
    s->attracted = (int*)calloc(tot, sizeof(int));
    s->lambdas = (float*)calloc(tot, sizeof(float));
    s->densities = (float*)calloc(tot, sizeof(float));
    s->nbr_offsets = (long*)calloc(tot + 1, sizeof(long));
    s->keys = (int*)calloc(tot, sizeof(int));
    s->sorted_index = (int*)calloc(tot, sizeof(int));
    s->cell_counts = (int*)calloc((size_t)s->num_grid_cells + 1, sizeof(int));
    s->cell_start = (int*)calloc((size_t)s->num_grid_cells + 1, sizeof(int));
    s->cell_items = (int*)calloc(tot, sizeof(int));
    s->solid_cell_start = (int*)calloc((size_t)s->num_grid_cells + 1, sizeof(int));
    s->solid_cell_items = (int*)calloc(n_solid > 0 ? n_solid : 1, sizeof(int));
    s->scratch3 = (float*)calloc(3 * tot, sizeof(float));
    s->scratchi = (int*)calloc(tot + 1, sizeof(int));
    return s;
}

void lo_destroy(lo_sim* s) {
    free(s->positions); free(s->positions_star); free(s->positions_tmp); free(s->velocities);
    free(s->attracted); free(s->lambdas); free(s->densities); free(s->nbr_offsets); free(s->nbr);
    free(s->keys); free(s->sorted_index); free(s->cell_counts); free(s->cell_start); free(s->cell_items);
    free(s->solid_cell_start); free(s->solid_cell_items); free(s->scratch3); free(s->scratchi);
    free(s);
}

void lo_set_sand(lo_sim* s, int n, const float* pos, const float* vel, const int* attracted) {
    s->n_sand = n;
    /* solids live right after the sand block: move them if the sand count changed */
    if (pos) { memcpy(s->positions, pos, sizeof(float) * 3 * n); memcpy(s->positions_star, pos, sizeof(float) * 3 * n); }
    if (vel) memcpy(s->velocities, vel, sizeof(float) * 3 * n);
    else memset(s->velocities, 0, sizeof(float) * 3 * n);
    if (attracted) memcpy(s->attracted, attracted, sizeof(int) * n);
    else memset(s->attracted, 0, sizeof(int) * n);
}

/* Solids are stored in a separate block addressed as index capacity + k, reported to
 * callers as n_sand + k (see lo_nbr_public_index). */
void lo_set_solid(lo_sim* s, const float* pos) {
    float* base = s->positions + 3 * (size_t)s->capacity;
    memcpy(base, pos, sizeof(float) * 3 * s->n_solid);
    memcpy(s->positions_star + 3 * (size_t)s->capacity, pos, sizeof(float) * 3 * s->n_solid);
    memcpy(s->positions_tmp + 3 * (size_t)s->capacity, pos, sizeof(float) * 3 * s->n_solid);
    s->solid_grid_built = 0;
}

/* ---- src/neighbors/Utils.hpp:24-33: get_cell_id, no clamp, true division, C truncation ---- */
int lo_cell_id(const lo_sim* s, float x, float y, float z) {
    float px = x / s->cell_size, py = y / s->cell_size, pz = z / s->cell_size;
    return ((int)py) * s->gridX * s->gridZ + ((int)px) * s->gridZ + ((int)pz);
}

static int checked_key(lo_sim* s, const float* p) {
    int id = lo_cell_id(s, p[0], p[1], p[2]);
    if (id < 0 || id >= s->num_grid_cells) { /* the reference would index out of bounds here (F10) */
        s->violations++;
        id = id < 0 ? 0 : s->num_grid_cells - 1;
    }
    return id;
}

/* ---- src/neighbors/Sorting.cpp:11-33: Sorting::counting_sort ---- */
void lo_counting_sort(int* counts, const int* keys, long n, long num_cells, int* sorted) {
    memset(counts, 0, (size_t)(num_cells + 1) * sizeof(int));
    for (long i = 0; i < n; i++) counts[keys[i]]++;
    for (long c = 1; c <= num_cells; c++) counts[c] += counts[c - 1];
    for (long i = n - 1; i >= 0; i--) {
        sorted[counts[keys[i]] - 1] = (int)i;
        counts[keys[i]]--;
    }
}

/* ---- src/Kernels.cpp:6-19: cubic_kernel ---- */
float lo_cubic_kernel(const lo_sim* s, float r) {
    float q = (r * s->kernelFactor) / s->kernelRadius;
    float result = 0.0f;
    if (q <= 1.0) {
        if (q <= 0.5) {
            float q2 = q * q;
            float q3 = q2 * q;
            result = s->cubic_kernel_k * (6.0f * q3 - 6.0f * q2 + 1.0f);
        } else {
            result = s->cubic_kernel_k * (2.0f * powf(1.0f - q, 3.0f));
        }
    }
    return result;
}

/* ---- src/Kernels.cpp:26-41: cubic_kernel_grad ---- */
static inline v3 cubic_grad(const lo_sim* s, v3 r) {
    v3 result = v3_make(0.0f, 0.0f, 0.0f);
    float rl = v3_length(r) * s->kernelFactor;
    float q = rl / s->kernelRadius;
    if (rl > 1.0e-5 && q <= 1.0) {
        v3 grad_q = v3_scale(r, 1.0f / (rl * s->kernelRadius));
        if (q <= 0.5) {
            result = v3_scale(grad_q, s->cubic_kernel_l * q * (3.0f * q - 2.0f));
        } else {
            float f = 1.0f - q;
            result = v3_scale(grad_q, s->cubic_kernel_l * (-f * f));
        }
    }
    return result;
}
void lo_cubic_kernel_grad(const lo_sim* s, const float r[3], float out[3]) { v3_store(out, cubic_grad(s, v3_load(r))); }

/* ---- src/Kernels.cpp:43-51: poly6_kernel(float); std::pow(float,int) promotes to double ---- */
float lo_poly6_kernel(const lo_sim* s, float r) {
    float result = 0.0f;
    float hf = s->kernelRadius * s->kernelFactor;
    if (r <= s->kernelRadius) {
        double a = 315.0f / (64.0f * 3.14f * pow((double)hf, 9.0));
        double b = pow(pow((double)hf, 2.0) - pow((double)(s->kernelFactor * r), 2.0), 3.0);
        result = (float)(a * b);
    }
    return result;
}

/* ---- src/Kernels.cpp:57-67: spiky_kernel ---- */
void lo_spiky_kernel(const lo_sim* s, const float rr[3], float out[3]) {
    v3 r = v3_load(rr);
    v3 result = v3_make(0.0f, 0.0f, 0.0f);
    float rl = v3_length(r);
    if (rl > 0.0 && rl <= s->kernelRadius) {
        float hf = s->kernelRadius * s->kernelFactor;
        float temp = (float)((15.0f / (3.14f * pow((double)hf, 6.0))) * pow((double)(hf - (rl * s->kernelFactor)), 2.0));
        result = v3_scale(v3_div(r, rl * s->kernelFactor), temp);
    }
    v3_store(out, result);
}

/* ---- src/Simulate.cpp:7-9: s_coor; std::pow(float,float) = powf ---- */
float lo_s_coor(const lo_sim* s, float rl) {
    return -s->s_corr_k * powf(lo_cubic_kernel(s, rl) / lo_cubic_kernel(s, s->s_corr_dq), s->s_corr_n);
}

/* ---- grid helpers ---- */
static void build_solid_grid(lo_sim* s) {
    /* src/neighbors/Neighbors.cpp:266-272: solids binned once, ascending index per cell */
    if (s->solid_grid_built) return;
    int C = s->num_grid_cells;
    int* cnt = s->solid_cell_start;
    memset(cnt, 0, ((size_t)C + 1) * sizeof(int));
    int* keys = (int*)malloc(sizeof(int) * (s->n_solid > 0 ? s->n_solid : 1));
    for (int k = 0; k < s->n_solid; k++) {
        keys[k] = checked_key(s, s->positions_star + 3 * ((size_t)s->capacity + k));
        cnt[keys[k] + 1]++;
    }
    for (int c = 0; c < C; c++) cnt[c + 1] += cnt[c];
    int* cur = (int*)malloc(sizeof(int) * ((size_t)C + 1));
    memcpy(cur, cnt, sizeof(int) * ((size_t)C + 1));
    for (int k = 0; k < s->n_solid; k++) s->solid_cell_items[cur[keys[k]]++] = k;
    free(cur); free(keys);
    s->solid_grid_built = 1;
}

static void build_sand_grid(lo_sim* s) {
    /* cell -> ascending sand indices (push_back order of Neighbors.cpp:296-304 / :374-378) */
    int C = s->num_grid_cells, n = s->n_sand;
    int* st = s->cell_start;
    memset(st, 0, ((size_t)C + 1) * sizeof(int));
    for (int i = 0; i < n; i++) st[s->keys[i] + 1]++;
    for (int c = 0; c < C; c++) st[c + 1] += st[c];
    int* cur = s->cell_counts;
    memcpy(cur, st, sizeof(int) * ((size_t)C + 1));
    for (int i = 0; i < n; i++) s->cell_items[cur[s->keys[i]]++] = i;
}

static void reserve_nbr(lo_sim* s, long total) {
    if (total > s->nbr_capacity) {
        free(s->nbr);
        s->nbr_capacity = total + total / 4 + 1024;
        s->nbr = (int*)malloc(sizeof(int) * (size_t)s->nbr_capacity);
    }
}

static inline const float* pstar(const lo_sim* s, int j) {
    /* public index j: sand [0,n_sand), solid n_sand + k stored at capacity + k */
    return s->positions_star + 3 * (size_t)(j < s->n_sand ? j : (j - s->n_sand) + s->capacity);
}

/* ---- src/neighbors/Neighbors.cpp:366-450: find_neighbors_uniform_grid (v0, fluid) ----
 * Cell lists hold sand (ascending) then solids (ascending) (:374-384).  For every sand
 * particle the 27-stencil is walked y-outer, x, z-inner (:401-403), skipping cells outside
 * the grid (:405-412), and every item of the neighbour cell that satisfies
 * dot(t,t) <= h*h is appended (:431-437) — self included.  The reference iterates cell by
 * cell; for a fixed particle the append order is identical to this per-particle walk. */
void lo_find_neighbors_v0(lo_sim* s) {
    const int n = s->n_sand, gX = s->gridX, gY = s->gridY, gZ = s->gridZ;
    s->solid_grid_built = 0; /* v0 re-bins solids every call (:380-384) */
    build_solid_grid(s);
    for (int i = 0; i < n; i++) s->keys[i] = checked_key(s, s->positions_star + 3 * (size_t)i);
    build_sand_grid(s);
    const float h2 = s->kernelRadius * s->kernelRadius;
    for (int pass = 0; pass < 2; pass++) {
        long total = 0;
        for (int i = 0; i < n; i++) {
            int id = s->keys[i];
            int yy = id / (gX * gZ);
            int rem = id - yy * gX * gZ;
            int xx = rem / gZ, zz = rem % gZ;
            v3 self = v3_load(s->positions_star + 3 * (size_t)i);
            if (pass == 1) s->nbr_offsets[i] = total;
            for (int y = -1; y <= 1; y++) for (int x = -1; x <= 1; x++) for (int z = -1; z <= 1; z++) {
                if (xx + x < 0 || xx + x >= gX || yy + y < 0 || yy + y >= gY || zz + z < 0 || zz + z >= gZ) continue;
                int c = (yy + y) * gX * gZ + (xx + x) * gZ + (zz + z);
                for (int t = s->cell_start[c]; t < s->cell_start[c + 1]; t++) {
                    int j = s->cell_items[t];
                    v3 d = v3_sub(self, v3_load(s->positions_star + 3 * (size_t)j));
                    if (v3_dot(d, d) <= h2) { if (pass == 1) s->nbr[total] = j; total++; }
                }
                for (int t = s->solid_cell_start[c]; t < s->solid_cell_start[c + 1]; t++) {
                    int k = s->solid_cell_items[t];
                    v3 d = v3_sub(self, v3_load(s->positions_star + 3 * ((size_t)s->capacity + k)));
                    if (v3_dot(d, d) <= h2) { if (pass == 1) s->nbr[total] = n + k; total++; }
                }
            }
        }
        if (pass == 0) reserve_nbr(s, total);
        else s->nbr_offsets[n] = total;
    }
}

/* ---- src/neighbors/Neighbors.cpp:262-364: find_neighbors_uniform_grid_v1 (sand) ---- */
void lo_find_neighbors_v1(lo_sim* s) {
    const int n = s->n_sand, gX = s->gridX, gY = s->gridY, gZ = s->gridZ;
    build_solid_grid(s); /* :266-272, cached */
    for (int i = 0; i < n; i++) s->keys[i] = checked_key(s, s->positions_star + 3 * (size_t)i); /* :275-278 */
    lo_counting_sort(s->cell_counts, s->keys, n, s->num_grid_cells, s->sorted_index);            /* :280-286 */
    /* :288-300 gather positions, positions_star, velocities, attracted through the permutation */
    float* arrays[3] = {s->positions, s->positions_star, s->velocities};
    for (int a = 0; a < 3; a++) {
        memcpy(s->scratch3, arrays[a], sizeof(float) * 3 * (size_t)n);
        for (int i = 0; i < n; i++) memcpy(arrays[a] + 3 * (size_t)i, s->scratch3 + 3 * (size_t)s->sorted_index[i], 3 * sizeof(float));
    }
    memcpy(s->scratchi, s->attracted, sizeof(int) * (size_t)n);
    for (int i = 0; i < n; i++) s->attracted[i] = s->scratchi[s->sorted_index[i]];
    /* :301-303 cell id recomputed from the moved positions_star; cells = solids then sand */
    for (int i = 0; i < n; i++) s->keys[i] = checked_key(s, s->positions_star + 3 * (size_t)i);
    build_sand_grid(s);
    const float h2 = s->kernelRadius * s->kernelRadius;
    long* cursor = (long*)malloc(sizeof(long) * ((size_t)n + 1));
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 0) memset(s->scratchi, 0, sizeof(int) * ((size_t)n + 1));
        for (int i = 0; i < n; i++) { /* :306-361 */
            int id = s->keys[i];
            int yy = id / (gX * gZ);
            int rem = id - yy * gX * gZ;
            int zz = rem % gZ, xx = rem / gZ;
            int ylo = -1, yhi = 1, xlo = -1, xhi = 1, zlo = -1, zhi = 1;
            if (yy + ylo < 0) ylo = 0; if (yy + yhi >= gY) yhi = 0;
            if (xx + xlo < 0) xlo = 0; if (xx + xhi >= gX) xhi = 0;
            if (zz + zlo < 0) zlo = 0; if (zz + zhi >= gZ) zhi = 0;
            v3 self = v3_load(s->positions_star + 3 * (size_t)i);
            for (int y = ylo; y <= yhi; y++) for (int x = xlo; x <= xhi; x++) for (int z = zlo; z <= zhi; z++) {
                int c = (yy + y) * gX * gZ + (xx + x) * gZ + (zz + z);
                /* solids first (copied from the static cache, :293), all have index >= ptr_solid_start > i */
                for (int t = s->solid_cell_start[c]; t < s->solid_cell_start[c + 1]; t++) {
                    int k = s->solid_cell_items[t];
                    v3 d = v3_sub(self, v3_load(s->positions_star + 3 * ((size_t)s->capacity + k)));
                    if (v3_dot(d, d) <= h2) {
                        if (pass == 0) s->scratchi[i]++; else s->nbr[cursor[i]++] = n + k;
                    }
                }
                for (int t = s->cell_start[c]; t < s->cell_start[c + 1]; t++) {
                    int j = s->cell_items[t];
                    if (j < i) continue; /* :343-345 */
                    v3 d = v3_sub(self, v3_load(s->positions_star + 3 * (size_t)j));
                    if (v3_dot(d, d) <= h2) { /* :349-353: push j to i, and i to j (also when j == i) */
                        if (pass == 0) { s->scratchi[i]++; s->scratchi[j]++; }
                        else { s->nbr[cursor[i]++] = j; s->nbr[cursor[j]++] = i; }
                    }
                }
            }
        }
        if (pass == 0) {
            long total = 0;
            for (int i = 0; i < n; i++) { s->nbr_offsets[i] = total; cursor[i] = total; total += s->scratchi[i]; }
            s->nbr_offsets[n] = total;
            reserve_nbr(s, total);
        }
    }
    free(cursor);
}

/* ================= fluid: src/Simulate.cpp:27-115 ================= */

/* simulation->W / simulation->gradW (src/Simulation.hpp:139-140): function pointers in the reference */
static inline float sim_W(const lo_sim* s, float r) { return s->sph_kernel ? lo_poly6_kernel(s, r) : lo_cubic_kernel(s, r); }
static inline v3 sim_gradW(const lo_sim* s, v3 d) {
    if (!s->sph_kernel) return cubic_grad(s, d);
    float in[3], out[3];
    v3_store(in, d);
    lo_spiky_kernel(s, in, out);
    return v3_load(out);
}
/* s_coor through simulation->W (src/Simulate.cpp:7-9) */
static inline float sim_s_coor(const lo_sim* s, float rl) {
    return -s->s_corr_k * powf(sim_W(s, rl) / sim_W(s, s->s_corr_dq), s->s_corr_n);
}

void lo_fluid_predict(lo_sim* s, float dt) {
    dt = clampf(dt, 0.001f, 0.01f); /* :31 */
    s->time_step = dt;
    v3 g = v3_load(s->gravity);
    for (int i = 0; i < s->n_sand; i++) { /* :48-51 */
        v3 v = v3_add(v3_load(s->velocities + 3 * (size_t)i), v3_scale(v3_scale(g, s->mass), dt));
        v3_store(s->velocities + 3 * (size_t)i, v);
        v3_store(s->positions_star + 3 * (size_t)i, v3_add(v3_load(s->positions + 3 * (size_t)i), v3_scale(v, dt)));
    }
    lo_find_neighbors_v0(s); /* :54 */
}

void lo_fluid_lambda(lo_sim* s) { /* :58-88 */
    for (int i = 0; i < s->n_sand; i++) {
        v3 pi = v3_load(s->positions_star + 3 * (size_t)i);
        float density = 0.0f;
        for (long t = s->nbr_offsets[i]; t < s->nbr_offsets[i + 1]; t++) {
            v3 ij = v3_sub(pi, v3_load(pstar(s, s->nbr[t])));
            density += s->mass * sim_W(s, v3_length(ij));
        }
        density += s->mass * sim_W(s, 0.0f);
        s->densities[i] = density;
        float constraint_i = (float)((double)(density / s->rest_density) - 1.0); /* :69 */
        float sum = 0.0f;
        v3 gi = v3_make(0.0f, 0.0f, 0.0f);
        for (long t = s->nbr_offsets[i]; t < s->nbr_offsets[i + 1]; t++) {
            v3 d = v3_sub(pi, v3_load(pstar(s, s->nbr[t])));
            v3 g = v3_scale(sim_gradW(s, d), -(s->mass / s->rest_density));
            sum += v3_dot(g, g);
            gi = v3_sub(gi, g);
        }
        sum += v3_dot(gi, gi);
        s->lambdas[i] = 0.0f;
        if (sum > 0.0) s->lambdas[i] = -constraint_i / (sum + s->relaxation_epsilon);
    }
}

/* src/Simulate.cpp:13-24 resolve_collision: returns epsilon_collision (0.01), not min (F9) */
static inline float resolve_collision(float value, float lo, float hi) {
    const float eps = 0.01f;
    if (value <= lo) return eps;
    if (value > hi) return hi - eps;
    return value;
}

void lo_fluid_deltap(lo_sim* s, int jacobi, int literal_lambda_index) { /* :90-108 */
    const int X = (int)s->domainX, Y = (int)s->domainY, Z = (int)s->domainZ; /* :35-37 int truncation */
    const float r = s->particleRadius;
    float* out = jacobi ? s->scratch3 : s->positions_star;
    for (int i = 0; i < s->n_sand; i++) {
        v3 pi = v3_load(s->positions_star + 3 * (size_t)i);
        v3 f = v3_make(0.0f, 0.0f, 0.0f);
        long b = s->nbr_offsets[i], e = s->nbr_offsets[i + 1];
        for (long t = b; t < e; t++) {
            v3 ij = v3_sub(pi, v3_load(pstar(s, s->nbr[t])));
            /* F4: the reference reads lambdas[j] with j the loop counter */
            float lj = literal_lambda_index ? s->lambdas[t - b] : (s->nbr[t] < s->n_sand ? s->lambdas[s->nbr[t]] : 0.0f);
            float w = (s->lambdas[i] + lj) + sim_s_coor(s, v3_length(ij));
            f = v3_add(f, v3_scale(sim_gradW(s, ij), w));
        }
        f = v3_div(f, s->rest_density);
        v3 p = v3_add(pi, f);
        p.x = resolve_collision(p.x, r, X - r);
        p.y = resolve_collision(p.y, r, Y - r);
        p.z = resolve_collision(p.z, r, Z - r);
        v3_store(out + 3 * (size_t)i, p);
    }
    if (jacobi) memcpy(s->positions_star, s->scratch3, sizeof(float) * 3 * (size_t)s->n_sand);
}

void lo_fluid_commit(lo_sim* s) { /* :110-111 */
    for (int i = 0; i < s->n_sand; i++) {
        v3 ps = v3_load(s->positions_star + 3 * (size_t)i);
        v3_store(s->velocities + 3 * (size_t)i, v3_div(v3_sub(ps, v3_load(s->positions + 3 * (size_t)i)), s->time_step));
        v3_store(s->positions + 3 * (size_t)i, ps);
    }
}

void lo_step_fluid(lo_sim* s, float dt, int iterations, int jacobi, int literal_lambda_index) {
    lo_fluid_predict(s, dt);
    if (!jacobi) {
        /* literal reference: velocity/position commit is interleaved with the in-place
         * delta-p loop (:103-111); interleaving does not change any value because
         * positions[] is never read by later particles. */
        lo_fluid_lambda(s);
        lo_fluid_deltap(s, 0, literal_lambda_index);
        lo_fluid_commit(s);
        return;
    }
    for (int it = 0; it < iterations; it++) {
        lo_fluid_lambda(s);
        lo_fluid_deltap(s, 1, literal_lambda_index);
    }
    lo_fluid_commit(s);
}

/* ================= sand: src/Simulate.cpp:156-325 (credits variant :327-510) ================= */

static inline float blow_kernel(float x, float kernel_radius) { /* :143-154 */
    if (x > kernel_radius) return 0.0f;
    return 1.0f - x / kernel_radius;
}

void lo_sand_predict(lo_sim* s, float dt, int credits) {
    s->time_step = dt; /* :168 raw dt */
    const float r = s->particleRadius;
    v3 g = v3_load(s->gravity);
    v3 player = v3_load(s->player_position);
    v3 lo = v3_make(r, r, r);
    v3 hi = v3_sub(v3_make(s->domainX, s->domainY, s->domainZ), v3_make(r, r, r));
    for (int i = 0; i < s->n_sand; i++) { /* :188-213 / :359-390 */
        float w = 1.0f / s->mass;
        v3 v = v3_load(s->velocities + 3 * (size_t)i);
        v3 p = v3_load(s->positions + 3 * (size_t)i);
        int a = s->attracted[i];
        if (credits && (a & 2)) v = v3_make(0.0f, -1.0f, 0.0f); /* :362-364 */
        v = v3_add(v, v3_scale(g, dt));
        if (s->prev_attract_flag && !s->attract_flag) a = credits ? (a & ~1) : 0; /* :191-193 / :366-368 */
        if (s->attract_flag) {
            if (v3_length(v3_sub(player, p)) < s->attract_radius) a = credits ? ((a | 1) & ~2) : 1;
            if (credits ? (a & 1) : a) {
                v3 to = v3_sub(v3_add(player, v3_make(0.0f, 1.5f, 0.0f)), p);
                v3 t = v3_scale(v3_normalize(to), 1.0f /* attract_kernel :138 */);
                t = v3_scale(t, s->attract_coeff); t = v3_scale(t, r); t = v3_scale(t, dt); t = v3_scale(t, w);
                v = v3_add(v, t);
            }
        }
        if (s->blow_flag) {
            v3 to = v3_sub(player, p);
            float len = v3_length(to);
            if (len < s->blow_radius) {
                v3 t = v3_scale(v3_neg(v3_normalize(to)), blow_kernel(len, s->blow_radius));
                t = v3_scale(t, s->blow_coeff); t = v3_scale(t, r); t = v3_scale(t, w);
                v = v3_add(v, t);
                if (credits) a &= ~2; /* :383 */
            }
        }
        s->attracted[i] = a;
        v3_store(s->velocities + 3 * (size_t)i, v);
        v3 ps = v3_add(p, v3_scale(v, dt));
        ps.x = clampf(ps.x, lo.x, hi.x); ps.y = clampf(ps.y, lo.y, hi.y); ps.z = clampf(ps.z, lo.z, hi.z); /* :212 */
        v3_store(s->positions_star + 3 * (size_t)i, ps);
    }
    lo_find_neighbors_v1(s); /* :215 */
    memcpy(s->positions_tmp, s->positions_star, sizeof(float) * 3 * (size_t)s->n_sand); /* :217-223 */
    s->prev_attract_flag = s->attract_flag; /* :323 (per call; moved here, nothing reads it in between) */
}

void lo_sand_iteration(lo_sim* s, int credits) { /* one pass of :228-310 */
    const float collision_coeff = 0.8f, friction_coeff = 0.7f;
    const float mu_s = credits ? 0.8f : 0.95f, mu_k = credits ? 0.7f : 0.9f; /* :162-163 / :333-334 */
    const float D = s->particleDiameter, r = s->particleRadius;
    v3 lo = v3_make(r, r, r);
    v3 hi = v3_sub(v3_make(s->domainX, s->domainY, s->domainZ), v3_make(r, r, r));
    const int n = s->n_sand;
    for (int i = 0; i < n; i++) {
        v3 deltap = v3_make(0.0f, 0.0f, 0.0f);
        v3 pi = v3_load(s->positions_tmp + 3 * (size_t)i);
        v3 xi_old = v3_load(s->positions + 3 * (size_t)i);
        for (long t = s->nbr_offsets[i]; t < s->nbr_offsets[i + 1]; t++) {
            int j = s->nbr[t];
            if (i == j) continue;
            size_t js = j < n ? (size_t)j : (size_t)(j - n) + s->capacity;
            v3 pj = v3_load(s->positions_tmp + 3 * js);
            v3 ij = v3_sub(pi, pj);
            if (v3_length(ij) == 0.0f) ij = v3_make(0.0f, 0.00001f, 0.0f);
            float len = v3_length(ij);
            if (len > D) continue;
            v3 tmp, xjdelta, nrm;
            if (j < n) { /* sand-sand :242-265 */
                float mass = s->mass, nmass = s->mass;
                float sc = collision_coeff * nmass / (mass + nmass) * (len - D);
                tmp = v3_div(v3_scale(ij, sc), len);
                xjdelta = v3_sub(v3_add(pj, tmp), v3_load(s->positions + 3 * js));
                nrm = v3_sub(v3_sub(pi, tmp), v3_add(pj, tmp));
            } else { /* sand-solid :266-286 */
                float sc = collision_coeff * (len - D);
                tmp = v3_div(v3_scale(ij, sc), len);
                xjdelta = v3_make(0.0f, 0.0f, 0.0f);
                nrm = v3_sub(v3_sub(pi, tmp), pj);
            }
            deltap = v3_sub(deltap, tmp);
            float d = v3_length(tmp);
            v3 xidelta = v3_sub(v3_sub(pi, tmp), xi_old);
            nrm = v3_normalize(nrm);
            v3 rel = v3_sub(xidelta, xjdelta);
            v3 xtan = v3_sub(rel, v3_scale(nrm, v3_dot(rel, nrm)));
            float lt = v3_length(xtan) + 1e-9f; /* avoid0 :134 */
            if ((d * mu_s) > lt) {
                deltap = v3_sub(deltap, v3_scale(xtan, friction_coeff));
            } else {
                float ratio = (mu_k * d / lt) < 1 ? (mu_k * d / lt) : 1;
                deltap = v3_sub(deltap, v3_scale(v3_scale(xtan, friction_coeff), ratio));
            }
            if (credits && (s->attracted[i] & 2)) s->attracted[i] &= ~2; /* :463-470 */
        }
        v3 ps = v3_add(pi, deltap); /* :288 */
        ps.x = clampf(ps.x, lo.x, hi.x); ps.y = clampf(ps.y, lo.y, hi.y); ps.z = clampf(ps.z, lo.z, hi.z); /* :307 */
        v3_store(s->positions_star + 3 * (size_t)i, ps);
    }
    memcpy(s->positions_tmp, s->positions_star, sizeof(float) * 3 * (size_t)n); /* :310 */
}

void lo_sand_commit(lo_sim* s) { /* :316-319 */
    lo_fluid_commit(s);
}

void lo_step_sand(lo_sim* s, float dt, int iterations, int credits) {
    lo_sand_predict(s, dt, credits);
    for (int it = 0; it < iterations; it++) lo_sand_iteration(s, credits);
    lo_sand_commit(s);
}
