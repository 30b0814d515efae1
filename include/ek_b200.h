/*
 * ek_b200.h -- C ABI of the B200-native coupled electrokinetic time step.
 *
 * Drop-in boundary for the simulation loop of gyf135/EK-PNP-3D
 * (main.cu:189-224): D3Q27 TRT lattice-Boltzmann fluid + cation + anion +
 * temperature, coupled to the spectral Poisson solve.  Plain pointers and
 * sizes only; every entry point returns an ek_status instead of calling
 * exit() as the reference does (LBM.cu:35-53, LBM.h:187-208).
 *
 * "Replaces" cites the reference interface each entry point stands in for
 * (paths relative to the reference repository).
 *
 * Threading: one handle = one simulation on one CUDA device; a handle is not
 * thread-safe (the reference is single-threaded, default stream only).
 */
#ifndef EK_B200_H
#define EK_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define EK_B200_ABI_VERSION 1

typedef enum ek_status {
    EK_OK = 0,
    EK_ERR_INVALID = 1,   /* bad argument / unsupported grid */
    EK_ERR_CUDA = 2,      /* CUDA runtime error (see ek_last_error) */
    EK_ERR_CUFFT = 3,     /* cuFFT error */
    EK_ERR_STATE = 4,     /* call out of order (e.g. step before init) */
    EK_ERR_NOMEM = 5
} ek_status;

/* Input parameters.  Replaces the compile-time constants of LBM.h:29-125
 * (SURVEY.md App. B); same names, same meaning, now runtime values.
 * NX is the fastest-varying index of every array (LBM.cu:17-30). */
typedef struct ek_params {
    int NX, NY, NZ;                  /* LBM.h:32-35; z includes both wall planes */
    double Lx, Ly, Lz;               /* LBM.h:40-42; must be NX*dx, NY*dy, (NZ-1)*dz */
    double dx, dy, dz;               /* LBM.h:43-45 */
    double uw, exf;                  /* LBM.h:47-50 top-wall speed, body force along x */
    double CFL, dt, cs_square, rho0; /* LBM.h:51-54 */
    double chargeinf;                /* LBM.h:56 bulk concentration */
    double voltage, voltage2;        /* LBM.h:60,62 wall (zeta) potentials, z=0 / z=NZ-1 */
    double Ext, eps;                 /* LBM.h:64-65 external field along x, permittivity */
    double diffu, nu, K;             /* LBM.h:66-70 cation diffusivity, viscosity, cation mobility */
    double diffun, Kn;               /* LBM.h:73-76 anion diffusivity and mobility */
    double kB, electron, roomT, convertCtoCharge, PB_omega; /* LBM.h:87-91 */
    double D, Ra, TH;                /* LBM.h:95-98 thermal diffusivity, buoyancy coefficient, bottom temperature */
    double w0, ws, wa, wd;           /* LBM.h:109-112 D3Q27 weights */
    double V, VC, VCn, VT;           /* LBM.h:115-118 TRT magic products */
    int pb_iters;                    /* LBM.cu:89 Poisson-Boltzmann start-up iterations (501) */
} ek_params;

/* Macroscopic arrays, N = NX*NY*NZ doubles each, index NX*(NY*z+y)+x
 * (LBM.cu:22-25): the reference's dump contract (LBM.cu:2511-2521). */
typedef enum ek_field {
    EK_RHO = 0, EK_UX, EK_UY, EK_UZ, EK_CHARGE, EK_CHARGEN,
    EK_PHI, EK_T, EK_EX, EK_EY, EK_EZ, EK_NFIELDS
} ek_field;

/* population sets, in the order of the reference's arrays f, h, hn, temp */
typedef enum ek_set { EK_FLUID = 0, EK_CATION, EK_ANION, EK_TEMPERATURE, EK_NSETS } ek_set;

/* streaming schemes (ek_set_option "stream_mode") */
#define EK_STREAM_AA 0    /* in-place A-A pattern, one lattice (default)      */
#define EK_STREAM_PUSH 1  /* two lattices, collide-and-push                    */

typedef struct ek_handle ek_handle;

/* LBM.h as shipped (50x8x51 microchannel). */
void ek_default_params(ek_params *p);

/* Replaces main.cu:58-152 (device selection, 27 cudaMallocs, cuFFT plan,
 * wavenumber tables).  device < 0 keeps the current device. */
ek_status ek_create(const ek_params *p, int device, ek_handle **out);
/* Replaces main.cu:259-291. */
ek_status ek_destroy(ek_handle *h);

/* Replaces initialization() (LBM.h:159, LBM.cu:68-146): uniform state, then
 * pb_iters under-relaxed Poisson-Boltzmann iterations, entirely on device. */
ek_status ek_init_fields(ek_handle *h);
/* Replaces read_data()'s upload (LBM.cu:2629-2671) / a caller-made state:
 * fields[id] may be NULL to keep the current array.  src_on_device selects
 * cudaMemcpy direction. */
ek_status ek_set_fields(ek_handle *h, const double *const fields[EK_NFIELDS], int src_on_device);
/* Replaces init_equilibrium() (LBM.h:162, LBM.cu:150-463): populations of
 * the four sets from rho,u,c+,c-,T and E. */
ek_status ek_init_equilibrium(ek_handle *h);
/* ek_init_fields + ek_init_equilibrium: the reference's start of run with
 * flag == 0 (main.cu:165-174). */
ek_status ek_init(ek_handle *h);

/* nsteps iterations of main.cu:189-200: stream_collide_save() (LBM.h:165,
 * LBM.cu:465-481) followed by fast_Poisson() (LBM.h:176, poisson.cu:75-103).
 * Asynchronous on the handle's stream.  After it returns (and ek_sync) the
 * eleven macroscopic arrays hold what the reference's arrays hold after the
 * same number of loop iterations. */
ek_status ek_step(ek_handle *h, int nsteps);
/* The two halves separately (same contract as the reference functions):
 * one LBM pass that leaves c+ - c- for the solver, then the solve. */
ek_status ek_stream_collide_save(ek_handle *h, int write_fields);
ek_status ek_fast_poisson(ek_handle *h, int write_efield);
/* the LBM pass split into launches over z-chunk ranges [zblock0, zblock1) of
 * "zchunk" planes each (0,0 = all); last != 0 on the final launch of a pass.
 * Lets the host pipeline the distributed Poisson stage against the LBM pass. */
ek_status ek_stream_collide_save_range(ek_handle *h, int write_fields, int zblock0, int zblock1, int last);
/* ... and over a subset of the 32-column x-tiles of every row: xtiles = 0 all, 1 the two boundary tiles (first and
 * last), 2 the interior ones -- boundary tiles first, so that the halo exchange travels under the interior launch */
ek_status ek_stream_collide_save_part(ek_handle *h, int write_fields, int zblock0, int zblock1, int xtiles, int last);

/* ek_step bracketed by CUDA events on the handle's stream; blocks until done
 * and returns the device time of the nsteps steps in milliseconds. */
ek_status ek_step_timed(ek_handle *h, int nsteps, float *ms);

ek_status ek_sync(ek_handle *h);

/* One whole job on HOST arrays (pinned memory recommended), N doubles each in the reference's layout: upload
 * the eleven arrays (read_data()'s upload, LBM.cu:2629-2671), init_equilibrium() (main.cu:174), nsteps
 * iterations of main.cu:189-200, download the eleven arrays (the copies of save_data_tecplot,
 * LBM.cu:2511-2521) -- with the upload pipelined in plane groups against init_equilibrium and the first
 * LBM pass, and the download of rho, u, c+, c-, T against the last LBM pass.  Blocking; same results as
 * ek_set_fields + ek_init_equilibrium + ek_step + ek_get_field, bit for bit. */
ek_status ek_run_from_host(ek_handle *h, const double *const in[EK_NFIELDS], int nsteps, double *const out[EK_NFIELDS]);

/* Replaces the cudaMemcpy D2H calls of save_data_tecplot/current/record_umax
 * (LBM.cu:2511-2521, main.cu:212-214, LBM.cu:2720-2722). */
ek_status ek_get_field(ek_handle *h, int id, double *dst, int dst_on_device);
/* Device pointer of a macroscopic array (owned by the handle). */
ek_status ek_field_ptr(ek_handle *h, int id, double **dev_ptr);
/* Make the handle use a caller-owned device array (reference layout, NX even)
 * as its macroscopic array `id` -- what the reference's main() allocates at
 * main.cu:94-106.  The handle never frees it.  ek_mark_fields_ready() declares
 * that the adopted arrays now hold an initial state written by the caller
 * (e.g. by the reference's own initialization(), LBM.cu:68). */
ek_status ek_adopt_field(ek_handle *h, int id, double *dev_ptr);
ek_status ek_mark_fields_ready(ek_handle *h);
/* Recompute c+ - c- from the charge/chargen arrays (after a caller wrote them,
 * e.g. the reference's gpu_PBE, LBM.cu:139-146) before ek_fast_poisson. */
ek_status ek_refresh_charge_difference(ek_handle *h);
/* Pre-collision populations of one set in the reference's layout:
 * 27*N doubles, [d][z][y][x], d = 0 the rest population (f0|f1 of
 * LBM.cu:17-30 back to back).  For tests. */
ek_status ek_get_populations(ek_handle *h, int set, double *dst, int dst_on_device);

/* The (kx,ky,kz) = (0,0,0) mode of the extended right-hand side.  It is zero
 * by oddness in exact arithmetic; the reference divides whatever rounding
 * residue its cuFFT Z2Z leaves there by mu := 1 (poisson.cu:176-177) while
 * every other mode is divided by mu ~ 1e15, i.e. it adds an
 * implementation-dependent constant to the interior potential (DESIGN.md,
 * "DC artefact").  EK_DC_ZERO (default) enforces the exact value 0;
 * EK_DC_LITERAL keeps this library's own residue with mu = 1;
 * EK_DC_PRESCRIBED uses ghat0 as the forward coefficient (test hook: replay
 * the residue recorded from a reference run). */
#define EK_DC_ZERO 0
#define EK_DC_LITERAL 1
#define EK_DC_PRESCRIBED 2
ek_status ek_set_poisson_dc(ek_handle *h, int mode, double ghat0);

/* options: "stream_mode" (EK_STREAM_*; before ek_init*), "zchunk",
 * "profile" (1: time every LBM/Poisson launch with CUDA events),
 * "graph" (ek_step replays a CUDA graph of two coupled steps: 1 on, 0 off,
 * -1 automatic = grids below 4 M cells, which are launch-latency bound),
 * "kernel" (0 default: z-walking CTAs with the lean deep-interior node path, the odd A-A step with
 * the lattice row stride as a compile-time immediate where an instantiation exists; 4: the generic
 * lean kernel for every row length; 3: general node path everywhere).
 * Cross-check build only (libek_b200_xcheck.so, ek_is_xcheck_build()):
 * "poisson_path" 1 = the reference's odd-extension 3-D FFT (poisson.cu:75-103
 * literally), "kernel" 1/2 = eight-/five-warp LBM kernels, 5/6 = x-marching rows for the odd
 * A-A step (aligned stores / aligned loads and stores), EK_DC_LITERAL. */
ek_status ek_set_option(ek_handle *h, const char *key, long long value);
/* counters: "steps", "zchunk", "lbm_launches", "poisson_launches", "kernel_launches";
 * times (ms, profile on): "lbm_ms", "poisson_ms" */
ek_status ek_get_counter(ek_handle *h, const char *key, double *value);
ek_status ek_reset_counters(ek_handle *h);

/* cudaStream_t of the handle, as void*. */
void *ek_stream(ek_handle *h);
const char *ek_last_error(ek_handle *h);
int ek_abi_version(void);
int ek_is_xcheck_build(void);   /* 1 in libek_b200_xcheck.so (test-only kernel variants + Poisson path 1), 0 in the product */
int ek_device_count(void);   /* CUDA devices visible to the process (0: none -- there is no CPU path) */

/* Diagnostics on device.  Replace current() (LBM.cu:2674-2710) and the
 * reduction of record_umax() (LBM.cu:2712-2753). */
ek_status ek_wall_current(ek_handle *h, double *current);
ek_status ek_max_uz(ek_handle *h, double *umax);

/* Field dumps in the reference's formats (LBM.cu:2492-2627), including the
 * dump-time wall extrapolation of LBM.cu:2527-2542.  first != 0 writes the
 * VARIABLES header. */
ek_status ek_save_data_tecplot(ek_handle *h, const char *path, double time, int append, int first);
ek_status ek_save_data_end(ek_handle *h, const char *path, double time);
/* Restart.  ek_read_data = read_data() (LBM.cu:2629-2671): restores the macroscopic
 * arrays from the text file of save_data_end (6 decimals, as the reference); the
 * caller then calls ek_init_equilibrium() like main.cu:161-176.  The checkpoint pair
 * is the lossless alternative: fields and pre-collision populations in the
 * reference's natural order, independent of the in-place layout and A-A parity. */
ek_status ek_read_data(ek_handle *h, const char *path, double *time);
ek_status ek_checkpoint_save(ek_handle *h, const char *path, double time);
ek_status ek_checkpoint_load(ek_handle *h, const char *path, double *time);
ek_status ek_set_populations(ek_handle *h, int set, const double *host_src);
ek_status ek_populations_restored(ek_handle *h);

/* ------------------------------------------------------------------------
 * Multi-GPU: x-slab decomposition (new; the reference is single-GPU,
 * main.cu:58).  One handle per slab/process/GPU; the host moves the buffers
 * between ranks (ek-pnp-3d_b200/slab.py over torch.distributed/NCCL).
 * `global` describes the whole domain; rank r owns columns
 * [r*NX/nranks, (r+1)*NX/nranks).  Slabs use the A-A scheme.
 * ------------------------------------------------------------------------ */
ek_status ek_create_slab(const ek_params *global, int device, int rank, int nranks, ek_handle **out);
/* run on a caller-provided cudaStream_t (e.g. torch's current stream) */
ek_status ek_set_stream(ek_handle *h, void *stream);
/* switch streams without draining the old one (the caller orders them with events) */
ek_status ek_switch_stream(ek_handle *h, void *stream);
ek_status ek_ensure_allocated(ek_handle *h);
int ek_row_pitch(ek_handle *h);          /* doubles per row of a field array (>= NX, ghosts included) */
int ek_lbm_parity(ek_handle *h);         /* 1 after an even (local) A-A step, 0 after an odd one */
/* population halos: 4 sets x 9 populations x NY x NZ doubles per face.
 * phase 0: after an even step (boundary columns -> neighbours' ghost columns),
 * phase 1: after an odd step (ghost columns -> neighbours' boundary columns). */
long long ek_halo_doubles(ek_handle *h);
ek_status ek_halo_pack(ek_handle *h, int phase, double *to_left, double *to_right);
ek_status ek_halo_unpack(ek_handle *h, int phase, const double *from_left, const double *from_right);
/* one ghost column of phi per face (NY x NZ doubles) for the fused E = -grad(phi) */
ek_status ek_phi_halo_pack(ek_handle *h, double *to_left, double *to_right);
ek_status ek_phi_halo_unpack(ek_handle *h, const double *from_left, const double *from_right);
/* the same for the planes [z0, z1) only; the buffers keep the full [z][y] layout, so the
 * planes of a range are the contiguous slice [z0*NY, z1*NY) */
ek_status ek_phi_halo_pack_range(ek_handle *h, int z0, int z1, double *to_left, double *to_right);
ek_status ek_phi_halo_unpack_range(ek_handle *h, int z0, int z1, const double *from_left, const double *from_right);
/* distributed Poisson stage: c+ - c- array, the z-solve on a block of ky rows
 * of the full-x spectrum [NZ-2][kyl][NXglobal] complex, and the epilogue
 * (wall planes of phi, flags) once the host has written phi's interior */
ek_status ek_dq_ptr(ek_handle *h, double **dev_ptr);
ek_status ek_zsolve_columns(ek_handle *h, double *spec, int ky0, int kyl);
ek_status ek_poisson_finish(ek_handle *h, int set_walls);
ek_status ek_compute_efield(ek_handle *h);
/* The distributed fast_Poisson() (replaces poisson.cu:75-103 on a domain split
 * along x; ek_slab_poisson.cu): y-transform of my columns, slab transpose,
 * x-transform + z-solve + inverse x-transform of my ky rows, transpose back,
 * inverse y-transform into phi.  The planes are processed in `nchunks` groups
 * of the LBM kernel's z-blocks so that the host can pipeline the transposes
 * (and the LBM launches that produce the planes) against the transforms.
 * Per chunk k the host runs
 *     forward(k); all-to-all(send_k -> recv_k); gather_x(k)
 * then solve() once, then per chunk
 *     scatter_x(k); all-to-all(send_k -> recv_k); backward(k)
 * and ek_poisson_finish().  send/recv hold `count` complex doubles in nranks
 * equal parts (part i travels to / comes from rank i). */
ek_status ek_slab_poisson_setup(ek_handle *h, int nchunks);   /* nchunks <= 0: automatic (7 chunks 1:2:3:4:3:2:1 of the z-blocks, or 4 equal ones) */
int ek_slab_poisson_chunks(ek_handle *h);
/* The chunk plan ek_slab_poisson_setup uses, as pure host arithmetic: bounds[0..K] in LBM z-blocks (bounds needs 17
 * entries), returns K.  sizes_csv: explicit chunk sizes "2,3,4,4,2,1" (what the EK_POISSON_CHUNK_BLOCKS environment
 * variable holds) or NULL. */
int ek_slab_poisson_plan_chunks(int nblocks, int nchunks, const char *sizes_csv, int *bounds);
ek_status ek_slab_poisson_chunk(ek_handle *h, int k, int *block0, int *block1, void **send, void **recv,
                                long long *count);
ek_status ek_slab_poisson_forward(ek_handle *h, int k);
ek_status ek_slab_poisson_gather_x(ek_handle *h, int k);
ek_status ek_slab_poisson_solve(ek_handle *h);
ek_status ek_slab_poisson_scatter_x(ek_handle *h, int k);
ek_status ek_slab_poisson_backward(ek_handle *h, int k);
/* Way back with the two ghost columns of phi travelling inside transpose 2 (no separate phi halo exchange):
 * after enable_ghosts the host runs scatter_xg(k); all-to-all(chunk_back buffers); backward_g(k) per chunk */
ek_status ek_slab_poisson_enable_ghosts(ek_handle *h);
ek_status ek_slab_poisson_chunk_back(ek_handle *h, int k, void **send, void **recv, long long *count);
ek_status ek_slab_poisson_scatter_xg(ek_handle *h, int k);
ek_status ek_slab_poisson_backward_g(ek_handle *h, int k);
/* Direct peer-memory transport of the two transposes (one node, NVLink): every
 * rank maps the pencil and receive buffers of every other rank -- CUDA IPC
 * handles exported here and exchanged by the host (ipc_bytes() bytes per rank),
 * or plain pointers for slabs of the same process -- and the re-blocking kernel
 * writes its rows straight into them:
 *     push_x(k)    = all-to-all 1 + gather_x(k)
 *     push_back(k) = scatter_x(k) + all-to-all 2
 * The host provides a cross-rank barrier after the push_x of all chunks and
 * another after the push_back of all chunks. */
int ek_slab_poisson_ipc_bytes(void);
ek_status ek_slab_poisson_ipc_export(ek_handle *h, void *handles);
ek_status ek_slab_poisson_ipc_import(ek_handle *h, int rank, const void *handles);
ek_status ek_slab_poisson_set_peer(ek_handle *h, int rank, void *X, void *R);
ek_status ek_slab_poisson_my_buffers(ek_handle *h, void **X, void **R);
ek_status ek_slab_poisson_push_x(ek_handle *h, int k);
ek_status ek_slab_poisson_push_back(ek_handle *h, int k);
/* pushes by strided 3-D copies on the copy engines (NVLink DMA, no SM time) instead of the kernel */
ek_status ek_slab_poisson_set_dma(ek_handle *h, int on);
/* the pieces of initialization() (LBM.cu:68-146) for a host-driven PB loop */
ek_status ek_init_uniform(ek_handle *h);
ek_status ek_pbe(ek_handle *h);
ek_status ek_pbe_relax(ek_handle *h);

/* ------------------------------------------------------------------------
 * Multi-GPU from ONE host process (ek_multi.cu): the x-slab path driven natively,
 * for a C++ caller such as the reference's main().  One slab per entry of
 * `devices` (entries may repeat); halos and the Poisson transposes go through
 * CUDA peer access inside the process, cross-device ordering through events.
 * Same call sequence as the single-GPU handle; ek_multi_slab() gives access to
 * the per-slab handles (options, counters).  Global arrays are in the
 * reference's layout NX*(NY*z+y)+x over the WHOLE domain, in host memory.
 * ------------------------------------------------------------------------ */
typedef struct ek_multi ek_multi;
ek_status ek_multi_create(const ek_params *global, int nslabs, const int *devices, int poisson_chunks, ek_multi **out);
ek_status ek_multi_destroy(ek_multi *m);
ek_status ek_multi_init_fields(ek_multi *m);        /* initialization(), LBM.cu:68-146 */
ek_status ek_multi_init_equilibrium(ek_multi *m);   /* init_equilibrium(), LBM.cu:150-463 */
ek_status ek_multi_init(ek_multi *m);
ek_status ek_multi_step(ek_multi *m, int nsteps);   /* main.cu:189-200 */
ek_status ek_multi_step_timed(ek_multi *m, int nsteps, float *ms);
ek_status ek_multi_sync(ek_multi *m);
ek_status ek_multi_get_field(ek_multi *m, int id, double *host_global);
ek_status ek_multi_set_fields(ek_multi *m, const double *const host_global[EK_NFIELDS]);
ek_status ek_multi_wall_current(ek_multi *m, double *current);
ek_status ek_multi_max_uz(ek_multi *m, double *umax);
ek_status ek_multi_save_data_tecplot(ek_multi *m, const char *path, double time, int append, int first);
ek_status ek_multi_save_data_end(ek_multi *m, const char *path, double time);
/* restart of the whole domain in the formats of ek_read_data / ek_checkpoint_* (files are interchangeable
 * between single- and multi-GPU runs; one population set at a time is staged in host memory) */
ek_status ek_multi_read_data(ek_multi *m, const char *path, double *time);
ek_status ek_multi_checkpoint_save(ek_multi *m, const char *path, double time);
ek_status ek_multi_checkpoint_load(ek_multi *m, const char *path, double *time);
ek_status ek_multi_set_pipeline(ek_multi *m, int on);   /* 1: overlap streams (default), 0: in sequence */
int ek_multi_slabs(ek_multi *m);
ek_handle *ek_multi_slab(ek_multi *m, int s);
const char *ek_multi_last_error(ek_multi *m);

/* ------------------------------------------------------------------------
 * Multi-GPU with ONE PROCESS PER GPU (ek_rank.cu): the layout of torchrun / mpirun launchers
 * and of bench.py --gpus N.  Every process owns one x-slab; halos travel as ncclSend/ncclRecv
 * between ring neighbours, the Poisson transposes as grouped send/recv with all ranks, on side
 * streams of the rank so that they overlap the LBM launches -- all of it driven from C++, no host
 * language between the launches of a step.  NCCL is loaded at run time (libnccl.so.2).
 * Rank 0 creates the id block and the launcher broadcasts it:
 *     char id[ek_rank_nccl_id_bytes()];  if (rank == 0) ek_rank_nccl_unique_id(id);  <broadcast id>
 *     ek_rank_create(&global, device, rank, nranks, id, 0, &r);          (collective)
 *     ek_rank_init(r);  ek_rank_step(r, n);  ek_get_field(ek_rank_slab(r), ...)   -- this rank's columns
 * Every ek_rank_init*, ek_rank_step* call is collective (same arguments on every rank).
 * ------------------------------------------------------------------------ */
typedef struct ek_rank ek_rank;
int ek_rank_nccl_id_bytes(void);
ek_status ek_rank_nccl_unique_id(void *id);
int ek_rank_nccl_version(void);                          /* e.g. 22809; 0: libnccl could not be loaded */
ek_status ek_rank_create(const ek_params *global, int device, int rank, int nranks, const void *nccl_id,
                         int poisson_chunks, ek_rank **out);
ek_status ek_rank_destroy(ek_rank *r);
ek_handle *ek_rank_slab(ek_rank *r);                     /* this rank's slab: fields, options, counters */
int ek_rank_chunks(ek_rank *r);
ek_status ek_rank_set_pipeline(ek_rank *r, int overlap, int overlap_back);
/* 1: launch the boundary x-tiles of the LBM pass first and send the population halos under the interior launches
 * (default 0: the halos travel next to the Poisson stage; measured faster at >= 8 M cells per GPU) */
ek_status ek_rank_set_boundary_first(ek_rank *r, int on);
ek_status ek_rank_init_fields(ek_rank *r);               /* initialization(), LBM.cu:68-146 */
ek_status ek_rank_init_equilibrium(ek_rank *r);          /* init_equilibrium(), LBM.cu:150-463 */
ek_status ek_rank_init(ek_rank *r);
ek_status ek_rank_step(ek_rank *r, int nsteps);          /* main.cu:189-200 */
ek_status ek_rank_step_timed(ek_rank *r, int nsteps, float *ms);   /* this rank's device time; max over ranks = the job's */
ek_status ek_rank_sync(ek_rank *r);
ek_status ek_rank_get_counter(ek_rank *r, const char *key, double *value);   /* + "nccl_groups" */
/* phase split of nsteps steps as a JSON object of ms per step; sequential != 0 disables every overlap */
ek_status ek_rank_profile(ek_rank *r, int nsteps, int sequential, char *json, int cap);
const char *ek_rank_last_error(ek_rank *r);

#ifdef __cplusplus
}
#endif
#endif /* EK_B200_H */
