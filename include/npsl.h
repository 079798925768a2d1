/*
 * npsl.h -- C ABI of the B200-native force / energy / integrator path of np-shape-lab.
 *
 * This is the drop-in boundary. The reference has no plugin or FFI layer; its boundary is the
 * in-process C++ call surface of md_interface (src/newmd.cpp). Each entry point below replaces
 * the body of one of those calls; the reference file:line it stands for is given beside it.
 * INTEGRATION.md shows the adapter a maintainer of the reference would add on their side.
 *
 * Conventions
 *   - plain pointers and sizes only; all host arrays are caller-owned, double precision,
 *     vectors are [n][3] row-major (x,y,z) exactly like a packed array of VECTOR3D-as-double;
 *   - every call returns 0 on success and a negative npsl_status on failure; the message is
 *     available from npsl_last_error(); the library never calls exit()/abort();
 *   - a context is single-owner and not thread-safe (the reference's callers are
 *     single-threaded, src/newmd.cpp); it is bound to the CUDA device that is current when it
 *     is created; all device memory is allocated once at creation;
 *   - there is no CPU fallback: without a CUDA device npsl_create fails with NPSL_ERR_CUDA.
 */
#ifndef NPSL_H
#define NPSL_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NPSL_ABI_VERSION 2
#define NPSL_MAX_CHAIN 16   /* links per Nose-Hoover chain, incl. the Q=0 dummy tail */
#define NPSL_MAX_DEGREE 8   /* maximum vertex valence supported by the gather kernels */
#define NPSL_NCCL_ID_BYTES 128

typedef enum {
    NPSL_OK = 0,
    NPSL_ERR_ARG = -1,      /* bad argument / unsupported mesh */
    NPSL_ERR_CUDA = -2,     /* CUDA runtime error (incl. no device, extension not usable) */
    NPSL_ERR_NCCL = -3,     /* NCCL error or libnccl not loadable (sharded contexts only) */
    NPSL_ERR_STATE = -4,    /* call sequence error (e.g. step before force_init) */
    NPSL_ERR_NUMERIC = -5   /* non-finite energy detected at a read-back */
} npsl_status;

typedef struct npsl_ctx npsl_ctx;

/* Static topology, as INTERFACE::discretize leaves it (src/interface.cpp:70-156).
 * Flattens VERTEX/EDGE/FACE pointer links (src/vertex.h:61-62, src/edge.h:38-39, src/face.h:35-36). */
typedef struct {
    int32_t n_vertices, n_edges, n_faces;
    const int32_t *edge_v; /* [n_edges][2]  EDGE::itsV indices                              */
    const int32_t *face_v; /* [n_faces][3]  FACE::itsV indices in the stored winding
                              (already reversed where the file's sign column is -1,
                              src/interface.cpp:121-129)                                    */
    const double *l0;      /* [n_edges]     EDGE::l0, initial edge lengths                  */
} npsl_mesh_desc;

/* One thermostat link, THERMOSTAT{Q,T,dof,xi,eta} (src/thermostat.h:28-35). */
typedef struct {
    double Q, T, dof, xi, eta;
} npsl_bath_link;

/* Reduced-unit parameters, already derived on the host exactly as main() does
 * (src/main.cpp:196-273). */
typedef struct {
    double scalefactor;         /* epsilon_water*lB_water/R            src/main.cpp:210       */
    double em;                  /* INTERFACE::em                       src/interface.cpp:22   */
    double inv_kappa_out;       /* Debye length                        src/main.cpp:235-239   */
    double elj;                 /* LJ strength                         src/main.cpp:273       */
    double lj_length_mesh_mesh; /* 0.1*avg_edge_length                 src/main.cpp:269-270   */
    double bkappa;              /* bending modulus (-b)                                        */
    double sconstant;           /* stretching modulus (-s)                                     */
    double sigma_a;             /* reduced surface tension             src/main.cpp:202-203   */
    double sigma_v;             /* reduced volume tension              src/main.cpp:206-207   */
    double ref_volume;          /* INTERFACE::ref_volume               src/main.cpp:265       */
    double avg_edge_length;     /* INTERFACE::avg_edge_length          src/interface.cpp:148-155 */
    double dt;                  /* CONTROL::timestep                                           */
    int32_t buckling;           /* 0: -B n, 1: -B y (stretching form, src/vertex.cpp:61-88)   */
    int32_t chain;              /* links per chain incl. dummy (= -C)  src/main.cpp:276-290   */
    npsl_bath_link mesh_bath[NPSL_MAX_CHAIN];
    npsl_bath_link ions_bath[NPSL_MAX_CHAIN]; /* evolves even with zero ions, src/newmd.cpp:101 */
} npsl_params;

/* Energies as INTERFACE::compute_energy reports them (src/interface.cpp:964-1137) plus the
 * bath terms md_interface adds (src/newmd.cpp:178-192). */
typedef struct {
    double kinetic;        /* netMeshKE                                             */
    double bending;        /* bE   src/interface.cpp:990-995                        */
    double stretching;     /* sE   src/interface.cpp:1000-1015                      */
    double tension;        /* tE   src/interface.cpp:1023                           */
    double lj;             /* ljE  src/interface.cpp:1046-1099                      */
    double electrostatic;  /* esE                                                   */
    double volume;         /* volE src/interface.cpp:1030                           */
    double potential;      /* INTERFACE::penergy                                    */
    double energy;         /* INTERFACE::energy = KE + penergy                      */
    double mesh_bath_ke, mesh_bath_pe, ions_bath_ke, ions_bath_pe; /* src/newenergies.h:28-46 */
    double total_area, total_volume;                               /* INTERFACE::total_*      */
    double ions_kinetic;   /* INTERFACE::netIonKE (0 without counterions); energy = kinetic + ions_kinetic + potential */
} npsl_energies;

/* ---- life cycle ------------------------------------------------------------------------ */

/* Single-GPU context on the current CUDA device. Replaces the lazy set-up the reference does
 * by walking INTERFACE on every call (src/newforces.cpp:140-143). */
int npsl_create(const npsl_mesh_desc *mesh, const npsl_params *params, npsl_ctx **out);

/* One rank of a sharded job (one process per GPU). The units of the pair interaction are split over the ranks; under
 * the "rows" plan the owned vertex range follows the reference's MPI decomposition (src/main.cpp:448-455) with the
 * per-rank range rounded up to a warp of vertices, under the "replicated" plan every rank integrates every vertex
 * (npsl_shard_plan). nccl_unique_id: NPSL_NCCL_ID_BYTES bytes obtained from npsl_nccl_unique_id() on rank 0 and
 * distributed by the caller.
 *
 * COLLECTIVE CALLS. Like the reference's force_calculation / compute_energy, whose MPI collectives make every rank call
 * them in lock-step (src/newforces.cpp:268-271, src/interface.cpp:1092-1094), every entry point that evaluates forces
 * -- npsl_force_init, npsl_force, npsl_force_calculation, npsl_step, npsl_step_timing*, npsl_step_profile,
 * npsl_energy, npsl_local_energies, npsl_md_run, npsl_md_step_host, npsl_md_step_packed, npsl_constraint_init -- and
 * npsl_download_state / npsl_download_force_components (rows plan: they gather) must be issued by ALL ranks of the
 * context, the same number of times and in the same order (do NOT guard them with rank == 0 the way the reference
 * guards its file output). The kernels of a rank wait on flags written by its peers; a wait whose peer never arrives
 * gives up after 30 s and the next npsl_sync / npsl_download_state / npsl_energy returns NPSL_ERR_STATE. */
int npsl_create_sharded(const npsl_mesh_desc *mesh, const npsl_params *params, int rank, int world,
                        const void *nccl_unique_id, npsl_ctx **out);
int npsl_nccl_unique_id(void *out_id);

void npsl_destroy(npsl_ctx *ctx);

/* Message of the most recent failure (ctx may be NULL for creation failures). */
const char *npsl_last_error(const npsl_ctx *ctx);
int npsl_abi_version(void);

/* ---- state ----------------------------------------------------------------------------- */

/* VERTEX::posvec / velvec / q -> device. Any pointer may be NULL (left unchanged). */
int npsl_upload_state(npsl_ctx *ctx, const double *pos, const double *vel, const double *q);

/* device -> VERTEX::posvec / velvec / forvec. Any pointer may be NULL. */
int npsl_download_state(npsl_ctx *ctx, double *pos, double *vel, double *force);

/* Force components of the most recent npsl_force_init / npsl_force call:
 * pair (forVecMeshGather), VERTEX::bforce, sforce, TForce, VolTForce (src/newforces.cpp:279-280).
 * Any pointer may be NULL. */
int npsl_download_force_components(npsl_ctx *ctx, double *pair, double *bending, double *stretching,
                                   double *tension, double *volume_tension);

/* FACE::itsarea / subtends_volume / itsdirection and EDGE::itsS of the most recent force call. */
int npsl_download_mesh_fields(npsl_ctx *ctx, double *face_area, double *face_volume, double *face_dir,
                              double *edge_S);

int npsl_get_thermostat(npsl_ctx *ctx, npsl_bath_link *mesh_bath, npsl_bath_link *ions_bath);
/* Also the annealing hook: T and Q of every link are taken from the arrays (src/newmd.cpp:295-298). */
int npsl_set_thermostat(npsl_ctx *ctx, const npsl_bath_link *mesh_bath, const npsl_bath_link *ions_bath);

/* ---- the path ---------------------------------------------------------------------------- */

/* INTERFACE::dressup (src/interface.cpp:33-67): face direction/area/volume, total area, total
 * volume at the current positions. */
int npsl_dressup(npsl_ctx *ctx, double *total_area, double *total_volume);

/* force_calculation_init (src/newforces.cpp:4-135) and force_calculation (:137-289): all forces at
 * the current positions; velocities untouched. Both also refresh the kinetic energy the
 * thermostat consumes (compute_kinetic_energy, src/newmd.cpp:52). */
int npsl_force_init(npsl_ctx *ctx);
int npsl_force(npsl_ctx *ctx);

/* n iterations of the loop body of md_interface (src/newmd.cpp:96-170): reverse half-chain,
 * half-kick, drift, dressup, force_calculation, half-kick, KE, forward half-chain. Runs entirely on
 * the device (one CUDA-graph launch per step); asynchronous until a download or npsl_sync. */
int npsl_step(npsl_ctx *ctx, int n_steps);

/* INTERFACE::compute_energy (src/interface.cpp:964-1137) + bath energies at the current state. */
int npsl_energy(npsl_ctx *ctx, npsl_energies *out);

int npsl_sync(npsl_ctx *ctx);

/* ---- host-buffer entry points (what a caller holding the reference's AoS mesh uses) -------- */

/* force_calculation with host buffers (src/newforces.cpp:137-289 as md_interface calls it,
 * src/newmd.cpp:143): positions [n][3] host -> device, every force term, total force [n][3]
 * (VERTEX::forvec) device -> host. pos == NULL keeps the device positions, force == NULL skips the
 * read-back. Host<->device copies go through pinned staging owned by the context. */
int npsl_force_calculation(npsl_ctx *ctx, const double *pos, double *force);

/* md_interface with host buffers (src/newmd.cpp:9-323 without its file writers): upload
 * pos/vel/q (NULL = keep), force_calculation_init, n_steps loop iterations, compute_energy at the
 * final state, download positions / velocities (NULL = skip). */
int npsl_md_run(npsl_ctx *ctx, const double *pos, const double *vel, const double *q, int n_steps,
                npsl_energies *energies_out, double *pos_out, double *vel_out);

/* The loop body of md_interface with the reference's host-resident state (src/newmd.cpp:96-170;
 * the reference keeps VERTEX::posvec / velvec / forvec in host memory between steps): positions,
 * velocities and forces [n][3] host -> device, n_steps iterations, the three arrays device -> host
 * in place, and the kinetic energy of the final velocities (compute_kinetic_energy,
 * src/newmd.cpp:156). force must hold the forces at pos (as left by npsl_force_calculation or by a
 * previous call). The thermostat chains stay on the device (npsl_get/set_thermostat). */
int npsl_md_step_host(npsl_ctx *ctx, double *pos, double *vel, double *force, int n_steps, double *kinetic_out);

/* The same loop body over ONE page-locked block owned by the context, laid out [posvec n*3 | velvec n*3 | forvec n*3 |
 * kinetic energy]: npsl_host_state hands out the four pointers into it (stable for the life of the context); the
 * caller -- the adapter that flattens VERTEX::posvec / velvec / forvec (src/vertex.h:60-70) anyway -- keeps its state
 * there. npsl_md_step_packed copies the block in with one transfer, advances n_steps and copies it back (for one step
 * on a single-GPU or replicated-plan context the kernels read and write the packed image directly, the positions
 * travel back beside the force evaluation and velocities + forces + kinetic energy in one transfer at the end);
 * returns after the block holds the new state. Collective on sharded contexts like npsl_step. */
int npsl_host_state(npsl_ctx *ctx, double **pos, double **vel, double **force, double **kinetic);
int npsl_md_step_packed(npsl_ctx *ctx, int n_steps);

/* Rigid volume constraint of md_interface with geomConstraint 'V', linear form (src/functions.h:33-171): SHAKE
 * after the drift (src/newmd.cpp:126), RATTLE after the second half-kick (:152), both inside npsl_step.
 * npsl_constraint_init is the SHAKE + RATTLE pass before the first step (src/newmd.cpp:33-38); call it after
 * npsl_force_init. Available on single-GPU and on sharded contexts (under the rows plan VERTEX::neg_GradVi and the
 * velocities are all-gathered for the face sums, which every rank then evaluates in the same order). The multipliers of
 * the last pass are available for diagnostics. */
int npsl_set_volume_constraint(npsl_ctx *ctx, int on);
int npsl_constraint_init(npsl_ctx *ctx);
int npsl_constraint_multipliers(npsl_ctx *ctx, double *shake_lambda, double *rattle_mu);

/* ---- the pair kernel's radial-factor table ----------------------------------------------------------------
 * k_pair_sym evaluates G(s) = e^{-r/lambda} (1/r^3 + 1/(lambda r^2)), s = r^2 -- the scalar the reference multiplies r_vec
 * with (src/newforces.cpp:173-174) -- from a piecewise degree-5 polynomial in s built at context creation for the
 * context's Debye length (csrc/pair_sym_kernel.cuh). npsl_pair_poly_eval rebuilds that table on the host for
 * inv_kappa_out and returns, for each s[i] in [2^-21, 2^5), the value the device's Horner evaluation gives and the
 * extended-precision value of G: the accuracy of the table can be checked without a GPU. Either output may be NULL.
 * npsl_pair_math: how the context's symmetric pair kernel evaluates G -- 0 rsqrt / exp to 3e-16, 1 the table, 2 the
 * reduced-accuracy exact scheme (csrc/pair_sym_kernel.cuh; environment NPSL_SYM_MATH overrides the default). */
int npsl_pair_poly_eval(double inv_kappa_out, const double *s, double *table_value, double *exact_value, int n);
/* The default evaluation (npsl_pair_math == 2) restated on the host, operation for operation: s[i] = r^2 in units of
 * lambda^2 (the kernel pre-scales the coordinates); value = e^{-r} (1/r^3 + 1/r^2) as the device computes it from a
 * reciprocal-square-root seed cut to seed_bits mantissa bits (the device's MUFU.RSQ64H seed is better than 20 bits),
 * exact_value = the extended-precision value. Pins the scheme's error budget (< 2e-12 against the path's 1e-10), the
 * index-biased exponent table and the polynomial constants without a GPU. Either output may be NULL. */
int npsl_pair_fast_eval(const double *s, double *value, double *exact_value, int n, int seed_bits);
int npsl_pair_math(const npsl_ctx *ctx);
/* Shape of the pair kernel in use (rows per thread, threads per CTA) and the FP64-pipe instructions it executes per
 * pair evaluation, as counted in the SASS of its inner loop (tools/sass_loops.py): per UNORDERED pair for the symmetric
 * kernel, per ordered pair for k_pair. bench.py derives the executed-instruction roofline from it. */
int npsl_pair_kernel_info(const npsl_ctx *ctx, int32_t *rows_per_thread, int32_t *threads, double *fp64_instr_per_pair);

/* ---- introspection (measurement only) --------------------------------------------------- */

/* Kernel launches issued by this context since creation (for bench.py's gpu_launches). */
int64_t npsl_launch_count(const npsl_ctx *ctx);
/* ---- explicit counterions ----------------------------------------------------------------------------------
 * vector<PARTICLE> counterions as main() hands it to force_calculation / compute_energy / md_interface
 * (src/particle.h; masses are 1 like the reference's): pos[n][3], vel[n][3] (NULL = at rest), q[n]; diameter = the
 * ion-ion LJ length (counterions[0].diameter, src/newforces.cpp:214), lj_length_mesh_ions = INTERFACE::lj_length_mesh_ions
 * (src/main.cpp:271-272). From then on every force evaluation adds the mesh-ion and ion-ion terms
 * (src/newforces.cpp:44-65,179-250), npsl_energy their energies (src/interface.cpp:1059-1089) and the ions' kinetic
 * energy, and npsl_step integrates the ions with their own Nose-Hoover chain (the caller sets ions_bath[0].dof = 3 n
 * like src/main.cpp:283). n = 0 removes the ions. Call before npsl_force_init. On sharded contexts every rank is given
 * the same ions and evaluates the (small) mesh-ion / ion-ion terms and the ion integrator redundantly, in identical
 * order, while the reference splits the ion loops by rank too (src/newforces.cpp:203-250, src/main.cpp:456-465). */
int npsl_set_ions(npsl_ctx *ctx, int n_ions, const double *pos, const double *vel, const double *q, double diameter,
                  double lj_length_mesh_ions);
/* PARTICLE::posvec / velvec / forvec of the counterions <- device; any pointer may be NULL. */
int npsl_download_ions(npsl_ctx *ctx, double *pos, double *vel, double *force);

/* ---- data feeds of the reference's writers ----------------------------------------------------------------
 * Vertex areas (1/3 of the incident face areas) and unit normals (area-weighted face normals), i.e.
 * VERTEX::compute_area_normal for every vertex (src/vertex.cpp:7-18; src/interface.cpp:39-41, src/newmd.cpp:232-235)
 * at the positions on the device: area[n], normal[n][3]; either may be NULL. interface_movie prints the area and
 * charge / area per vertex (src/functions.cpp:50-84). */
int npsl_vertex_geometry(npsl_ctx *ctx, double *area, double *normal);
/* Per-face energy maps of INTERFACE::compute_local_energies and compute_local_energies_by_component
 * (src/interface.cpp:1140-1337) at the positions on the device, one value per face (the g channel of the
 * reference's local_*_E.off files): electrostatic = sum over the face's vertices of their pair energy with every other
 * vertex; bending / stretching = sum over the face's three edges (stretching in its l0 form, as the reference does
 * here whatever -B says); elastic = stretching + bending. Any output may be NULL. Needs a force evaluation at the
 * current positions (npsl_force_init / npsl_force / npsl_energy); runs the energy-enabled pass itself when the last
 * evaluation did not keep the per-vertex energies. */
int npsl_local_energies(npsl_ctx *ctx, double *electrostatic, double *elastic, double *bending, double *stretching);

/* Average device time of the pair kernel over the launches since the last reset, measured with
 * CUDA events on the context's stream when timing is enabled. */
int npsl_pair_timing(npsl_ctx *ctx, int enable);
int npsl_pair_timing_read(npsl_ctx *ctx, double *mean_ms, int64_t *launches);
/* Device time in ms of n_steps steps (one CUDA-graph launch each), measured with CUDA events on the
 * context's own stream. flush_l2 != 0: a buffer larger than L2 is overwritten before every step,
 * outside that step's event bracket; the result is the sum of the per-step brackets. */
int npsl_step_timing(npsl_ctx *ctx, int n_steps, int flush_l2, double *total_ms);
/* Same measurement, one value per step: lead_in untimed steps (they absorb the skew between the ranks of a sharded job
 * after a host-side barrier), then n_steps timed ones, all submitted to the stream in one go; per_step_ms[n_steps].
 * Nothing is allocated inside the call once a call with the same flush_l2 and at least n_steps has been made. */
int npsl_step_timing_list(npsl_ctx *ctx, int n_steps, int flush_l2, int lead_in, double *per_step_ms);
/* Per-segment device time of a step (CUDA events between the launches of n_steps directly launched steps),
 * formatted as text into out. Measurement only. */
int npsl_step_profile(npsl_ctx *ctx, int n_steps, char *out, int out_len);
/* Override the pair kernel's launch plan (rows per thread in {1,2,4}, column splits >= 1); 0 keeps
 * the automatic choice for that field. Results do not depend on rows_per_thread; they depend on
 * col_splits only through the (fixed, ordered) grouping of the column sum. */
int npsl_set_pair_plan(npsl_ctx *ctx, int rows_per_thread, int col_splits);
/* Pair kernel selection: 0 automatic (symmetric from 4096 vertices up), 1 ordered -- k_pair walks all
 * N(N-1) ordered pairs like the reference loop --, 2 symmetric -- k_pair_sym evaluates each unordered pair
 * once and applies it to both vertices. npsl_pair_mode returns the kernel in use (1 or 2). */
int npsl_set_pair_mode(npsl_ctx *ctx, int mode);
int npsl_pair_mode(const npsl_ctx *ctx);
/* How a sharded context exchanges positions, pair-force sums and kinetic-energy partials each step:
 * 0 single GPU, 1 NCCL all-gathers, 2 direct stores into peer memory over NVLink (CUDA IPC; default when the
 * buffers can be mapped; NPSL_EXCHANGE=nccl in the environment forces 1). */
int npsl_exchange_mode(const npsl_ctx *ctx);
/* The work decomposition of the symmetric pair kernel, computed on the host without touching a device (for tests and
 * tools): the units (row block, column chunk) that rank `rank` of `world` computes, in launch order. Per unit four
 * int32 in `units` (at most max_units are written): row block, chunk, first column, one past the last column; a unit
 * covers the pairs (i, j) with i in the row block, j in [first, last) and j > i. geometry (6 int32, may be NULL): rows per
 * row block, column granularity, number of chunks, number of full-width chunks, tiles per full-width and per
 * half-width chunk. Returns the number of units of that rank (>= 0) or a negative status. The decomposition depends on
 * the vertex count only; ranks own the virtual ranks [8 rank / world, 8 (rank + 1) / world). */
int npsl_plan_pair_units(int n_vertices, int rank, int world, int32_t *units, int max_units, int32_t *geometry);
/* Sharding plan of a multi-GPU context: 0 single GPU; 1 rows -- vertices are owned in slices like the reference's
 * MPI ranks (src/main.cpp:448-455), three exchanges per step; 2 replicated integrator -- only the pair term is
 * split and every rank integrates all vertices, as the reference does after its all_gather of forces
 * (src/newforces.cpp:268-271), one exchange per step. Automatic: 2 for charged meshes up to 100 000 vertices when
 * peer memory is available; NPSL_SHARD=rows|replicated in the environment overrides. Trajectories are bit-identical
 * under both plans and on any GPU count. */
int npsl_shard_plan(const npsl_ctx *ctx);
/* Vertex range this context integrates, half-open (all vertices under the replicated plan). */
int npsl_owned_rows(const npsl_ctx *ctx, int32_t *lo, int32_t *hi);
/* Launch geometry of the pair kernel: rows per thread, row tiles x column splits, columns per CTA. */
int npsl_pair_geometry(const npsl_ctx *ctx, int32_t *rows_per_thread, int32_t *row_tiles, int32_t *col_splits,
                       int32_t *cols_per_cta);
/* Pure-DFMA microbenchmark: measured FP64 FMA instructions/s of the current device. */
int npsl_fp64_peak(double *dfma_per_s, double *sm_clock_mhz);

#ifdef __cplusplus
}
#endif
#endif /* NPSL_H */
