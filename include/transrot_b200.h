/*
 * transrot_b200.h — C ABI of the B200 annealing backend for TransRot.
 *
 * The reference (steventopper/TransRot v1.9.2, Java) has no plugin/FFI interface: the
 * seam is carved at the two calls the Java host makes into the hot path,
 *     Main.sawtoothAnneal(Space, double)   reference src/main/Main.java:81,177-263
 *     Space.calcEnergy()                   reference src/main/Space.java:648-658
 *                                          (called at Main.java:71,256 and Space.java:554,745)
 * The host keeps getArgs / readDB / parseConfig / readParams / setupPairVals / readInput and
 * every write* method; it flattens `space` into the arrays below, calls this library, and
 * rebuilds its output files from the returned buffers (INTEGRATION.md shows the Panama
 * binding).  Plain C, no callbacks, caller owns every buffer; the library copies in during
 * tr_set_* / tr_init_* and out during tr_get_*.
 *
 * Conventions
 *   - every function returns 0 (TR_OK) or a negative TR_E*; the text is in tr_last_error().
 *     The host turns non-zero into the RuntimeException("Error: ...") the reference would
 *     have thrown (Main.java:85-87).  Nothing aborts, nothing throws across the ABI.
 *   - a tr_ctx is bound to ONE CUDA device and used by one thread at a time; distinct
 *     contexts are independent (one per GPU for multi-GPU hosts).
 *   - coordinates are double, AoS [chain][site][xyz], sites in the order of `space`
 *     (Space.java:27): radius-sorted insertion order of Space.add (Space.java:46-61) or
 *     Input.xyz order.  Ghost ('*') sites are ordinary sites here; the host filters them on
 *     output (Space.java:574,748).
 *   - there is NO CPU fallback: without a CUDA device tr_create fails with TR_ENODEVICE.
 */
#ifndef TRANSROT_B200_H
#define TRANSROT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TR_OK          0
#define TR_EINVAL     -1   /* bad argument / call order                         */
#define TR_ENODEVICE  -2   /* no usable CUDA device (no CPU fallback exists)    */
#define TR_ECUDA      -3   /* a CUDA runtime call or kernel failed              */
#define TR_ELIMIT     -4   /* system exceeds a compiled-in limit (see below)    */
#define TR_ENOMEM     -5
#define TR_ERNG       -6   /* a chain was abandoned: its random-stream ring could not be refilled (tr_get_summary) */
#define TR_ENCCL      -7   /* NCCL could not be loaded or a collective failed (tr_*_multi)                        */

#define TR_KERNEL_AUTO 0   /* tr_set_kernel: let the library choose                                          */
#define TR_KERNEL_TEAM 1   /* team kernel (anneal.cuh): every lane of a chain's team runs the control flow   */
#define TR_KERNEL_CREW 2   /* warp-specialised kernel (crew.cuh): one control warp + position-major energy teams */
#define TR_KERNEL_CREW2 3  /* crew2.cuh: three control warps + particle-major teams + pair-energy cache          */
#define TR_KERNEL_TEAM_PC 4 /* team kernel on particle-major positions with the pair-energy cache in global memory */

#define TR_MAX_MOL_SITES  16    /* sites in one particle                        */
#define TR_MAX_SITES      1536  /* sites in `space`                             */
#define TR_MAX_TYPES      64    /* dbase atom slots                             */

typedef struct tr_ctx tr_ctx;

/* `space`, flattened.  Replaces the object graph walked by
 * Space.move/rotate/calcEnergy (Space.java:648-734). */
typedef struct tr_system {
    int32_t n_mol;                /* space.size()                                              */
    int32_t n_sites;              /* sum of atoms.size()                                       */
    int32_t n_types;              /* dbase atom slots = UUIDs (Atom.java:35,54)                */
    int32_t n_moveable;           /* moveableMolecules.size() (Space.java:28,219)              */
    const int32_t *mol_site_off;  /* [n_mol+1] first site of each particle                     */
    const double  *mol_radius;    /* [n_mol]   Molecule.radius (propagate only; may be NULL)   */
    const int32_t *site_type;     /* [n_sites] dbase atom slot of each site                    */
    const double  *site_q;        /* [n_sites] Atom.q; kTimesQ = 332.0637133*q is formed here  */
    const double  *site_def;      /* [n_sites*3] Atom.defX/Y/Z (propagate only; may be NULL)   */
    const double  *site_mass;     /* [n_sites] Atom.mass (centre of mass)                      */
    const int32_t *site_ghost;    /* [n_sites] 1 = excluded from the centre of mass            */
    const int32_t *moveable_idx;  /* [n_moveable] indices into space, moveableMolecules order  */
    const double  *pair_abcd;     /* [n_types*n_types*4] pairwiseDbase (Space.java:30)         */
} tr_system;

/* The Config fields Main.sawtoothAnneal reads (Config.java:12-33).  One per parameter set;
 * chains pick one by index, so seeds x parameter sweeps share a launch. */
typedef struct tr_schedule {
    double  max_temp;                  /* Config.maxTemp                        */
    double  temp_decrease_per_tooth;   /* Config.tempDecreasePerTooth           */
    double  max_trans;                 /* Config.maxTrans                       */
    double  magwalk_factor_trans;      /* Config.magwalkFactorTrans             */
    double  magwalk_prob_trans;        /* Config.magwalkProbTrans               */
    double  max_rot;                   /* Config.maxRot                         */
    double  magwalk_prob_rot;          /* Config.magwalkProbRot                 */
    int32_t moves_per_point;           /* Config.movePerPt                      */
    int32_t points_per_tooth;          /* Config.ptsPerTooth                    */
    int32_t points_added_per_tooth;    /* Config.ptsAddedPerTooth               */
    int32_t num_teeth;                 /* Config.numTeeth                       */
    int32_t eq_configs;                /* Config.eqConfigs                      */
    int32_t finale;                    /* Config.finale                         */
    int32_t static_temp;               /* Config.staticTemp                     */
    int32_t write_conf_heat_capacities;/* Config.writeConfHeatCapacitiesEnabled: the `continue`
                                          at Main.java:228 skips the movie frame when round(t,2)==0 */
} tr_schedule;

/* One record per iteration of the y-loop (Main.java:191): what writeAcceptance,
 * writeConfigurationalHeatCapacity and writeMovie need (Main.java:225-240). */
typedef struct tr_point_rec {
    int32_t tooth;
    int32_t point;
    int32_t accepted;        /* cumulative within the tooth (Main.java:189,223) */
    int32_t total;
    int32_t movie_written;   /* 0 when Main.java:228 skipped writeMovie         */
    int32_t finale;          /* 1 inside the 0 K finale repeat                   */
    double  temp;            /* t given to writeAcceptance                       */
    double  total_energy;    /* Σ startingEnergy  (Main.java:217)                */
    double  total_squared_energy;
    double  movie_energy;    /* calcEnergy() inside writeMovie (Space.java:745)  */
} tr_point_rec;

/* One record per iteration of the z-loop (Main.java:195); first `trace_moves` moves of each
 * chain when tracing is enabled.  With Static Temperature this is the energies.txt stream. */
typedef struct tr_trace_rec {
    int32_t mol;
    int8_t  kind;        /* 0 = Space.move, 1 = Space.rotate */
    int8_t  magwalk;
    int8_t  bounds;      /* rejected by the 0.75*spaceLen box (Space.java:663,706) */
    int8_t  accepted;
    int32_t axis;        /* Molecule.java:68, or -1 */
    int32_t drew_rand;
    double  e_start;
    double  e_end;
    double  dE;
    double  rand;        /* NaN when no Metropolis draw was made */
    double  temp;
    double  running;     /* startingEnergy after the move (Main.java:215) */
} tr_trace_rec;

typedef struct tr_chain_summary {
    double  start_energy;     /* calcEnergy() of the initial structure (Main.java:71)      */
    double  running_energy;   /* startingEnergy accumulator (Main.java:215)                */
    double  min_energy;       /* Space.minEnergy (Space.java:614-617)                      */
    double  space_len;        /* Config.spaceLen after the propagate retries (Main.java:59) */
    int64_t moves;            /* z-loop iterations executed                                */
    int64_t evaluated;        /* of which reached the energy loop (not bounds-rejected)    */
    int64_t accepted;         /* accepted moves over the whole run                         */
    int64_t pair_evals;       /* site-pair evaluations executed inside move/rotate          */
    int32_t min_tooth;        /* Space.minEnergyToothNum                                   */
    int32_t snapshots;        /* write(n) calls made so far (incl. write(0))               */
    int32_t points;           /* point records emitted so far                              */
    int32_t done;
    int32_t prop_retries;     /* propagate() failures before placement (Main.java:58-60)   */
    int32_t pad_;
} tr_chain_summary;

typedef struct tr_minimum {   /* what writeMinEnergy needs (Space.java:836-844)             */
    double  energy;
    int64_t chain;            /* global chain id = first_chain + local index               */
    int32_t tooth;
    int32_t pad_;
} tr_minimum;

/* ---- lifetime ------------------------------------------------------------------------- */
int  tr_create(int device, tr_ctx **out);
void tr_destroy(tr_ctx *ctx);
/* Message for the last failing call on ctx (ctx == NULL: last failing tr_create). */
const char *tr_last_error(const tr_ctx *ctx);
/* Launch on this cudaStream_t instead of the context's own stream (0 = back to own). */
int  tr_set_stream(tr_ctx *ctx, void *cuda_stream);
int  tr_synchronize(tr_ctx *ctx);

/* ---- problem definition ---------------------------------------------------------------- */
int  tr_set_system(tr_ctx *ctx, const tr_system *sys);
int  tr_set_schedules(tr_ctx *ctx, const tr_schedule *sched, int32_t n_sched);
/* Options, set before tr_init_*:  trace_moves = decision-trace records kept per chain;
 * record_points / record_snapshots = keep per-point records / write(n) structures;
 * threads_per_chain = 0 (auto), 16 (half-warp per chain, <= 64 sites), 32 (warp per chain,
 * <= 192 sites) or 128/256 (block per chain). */
int  tr_set_options(tr_ctx *ctx, int64_t trace_moves, int32_t record_points,
                    int32_t record_snapshots, int32_t threads_per_chain);
/* Keep the structure writeMovie appends after every point (Space.java:740-790), set before tr_init_*.  Off by
 * default: one [S][3] frame per point and chain (needs record_points). */
int  tr_set_record_movie(tr_ctx *ctx, int32_t on);

/* Kernel family and launch geometry, set before tr_init_* (0 / TR_KERNEL_AUTO = automatic).  For experiments and
 * tests; the automatic choice is what bench.py times. */
int  tr_set_kernel(tr_ctx *ctx, int32_t kernel, int32_t chains_per_block, int32_t blocks_per_sm);
/* The kernel family and variant the live chain batch runs with, e.g. "crew_kernel<L=16,SPL=4>" ("" before tr_init_*). */
const char *tr_kernel_name(const tr_ctx *ctx);
/* sizeof / offsetof of the structs above as this library was compiled: {sizeof(tr_system), offsetof(mol_site_off),
 * offsetof(pair_abcd), sizeof(tr_schedule), offsetof(moves_per_point), offsetof(write_conf_heat_capacities),
 * sizeof(tr_point_rec), offsetof(temp), offsetof(movie_energy), sizeof(tr_trace_rec), offsetof(e_start),
 * offsetof(running), sizeof(tr_chain_summary), offsetof(moves), offsetof(min_tooth), sizeof(tr_minimum),
 * offsetof(chain), offsetof(tooth)}.  Returns the number of entries (18); fills out[] when cap is large enough.
 * A binding checks its own layouts against these at load time. */
int  tr_abi_sizes(int32_t *out, int32_t cap);

/* Seeded start = Main.java:34-37 (seed fan-out) + Main.java:58-60 / Space.java:72-117
 * (propagate with box growth and the random initial rotation), all on the device.
 * sched_idx may be NULL (all chains use schedule 0).  first_chain only labels results. */
int  tr_init_chains_seeded(tr_ctx *ctx, int64_t n_chains, int64_t first_chain, const int64_t *seeds,
                           const int32_t *sched_idx, double space_len, int32_t max_prop_failures);
/* Explicit start = "Use Input.xyz" (Space.java:178-233): sites holds n_structs structures
 * (n_structs == 1: shared by all chains, or == n_chains); centres of mass are formed as
 * Molecule.setCenterOfMass does (Molecule.java:137-156).  RNG: seeds (fan-out only) or, if
 * mt_state != NULL, explicit states [n_chains][3][625] = streams Main.r, Space.r, Molecule.r,
 * 624 words + mti each. */
int  tr_init_chains_explicit(tr_ctx *ctx, int64_t n_chains, int64_t first_chain, const double *sites,
                             int64_t n_structs, const int64_t *seeds, const uint32_t *mt_state,
                             const int32_t *sched_idx, double space_len);

/* ---- the hot path ------------------------------------------------------------------------ */
/* Main.sawtoothAnneal for every chain.  max_moves > 0 advances each chain by at most that
 * many z-loop iterations and returns (resumable); max_moves <= 0 runs to completion.
 * Asynchronous on the context's stream; tr_get_* and tr_synchronize wait for it. */
int  tr_run(tr_ctx *ctx, int64_t max_moves);
/* Device time of the most recent tr_run kernel (CUDA events on the launch stream), and the
 * number of kernels this context has launched since creation. */
int  tr_last_run_ms(tr_ctx *ctx, float *ms);
int  tr_launch_count(tr_ctx *ctx, int64_t *n);

/* ---- results ------------------------------------------------------------------------------ */
int  tr_get_summary(tr_ctx *ctx, tr_chain_summary *out /* [n_chains] */);
int  tr_get_sites(tr_ctx *ctx, double *sites /* [n_chains][S][3] */, double *com /* [n_chains][N][3] or NULL */);
int  tr_get_points(tr_ctx *ctx, tr_point_rec *out, int64_t cap_per_chain);
int  tr_get_snapshots(tr_ctx *ctx, double *energy /* [n_chains][cap] */, int32_t *tooth /* same */,
                      double *sites /* [n_chains][cap][S][3] or NULL */, int64_t cap_per_chain);
/* The movie frames: frame p of a chain belongs to its point record p (movie_written says whether the reference
 * appended it, Main.java:228). */
int  tr_get_movie(tr_ctx *ctx, double *sites /* [n_chains][cap][S][3] */, int64_t cap_per_chain);
int  tr_get_trace(tr_ctx *ctx, tr_trace_rec *out /* [n_chains][trace_moves] */);
int  tr_get_rng_state(tr_ctx *ctx, uint32_t *out /* [n_chains][3][625] */);
/* Per-chain minimum over the write(n) snapshots (Space.java:614-617), packed on the device.
 * host_out and/or device_out (a device pointer, e.g. the tensor handed to an NCCL gather). */
int  tr_pack_minima(tr_ctx *ctx, tr_minimum *host_out, void *device_out);
/* The structure written by write(min_tooth) for one chain. */
int  tr_get_min_structure(tr_ctx *ctx, int64_t chain, double *sites /* [S][3] */, double *energy, int32_t *tooth);
/* The same structure for chains first .. first+count-1 (local indices), packed on the device: [count][S][3] (needs
 * record_snapshots).  host_out and/or device_out (a device pointer, e.g. the send buffer of an NCCL all-gather of
 * final structures). */
int  tr_pack_min_structures(tr_ctx *ctx, int64_t first, int64_t count, double *host_out, void *device_out);
int  tr_max_points(tr_ctx *ctx, int64_t *points_per_chain, int64_t *snapshots_per_chain);

/* ---- multi-GPU: one host thread, all GPUs of a box (SURVEY.md §8b,e) -------------------------------
 * The reference is one JVM with one thread (Main.java:17), so a Java host drives every GPU from one process: a
 * tr_multi owns one tr_ctx per device plus an NCCL communicator over them (ncclCommInitAll; NCCL is bound at run
 * time, and not at all for n == 1).  Chains are sharded in contiguous GLOBAL ranges (tr_multi_chain_range); seeds and
 * parameter sets are indexed by the global chain id, so per-chain results do not depend on the number of devices.
 * There is no traffic while the chains anneal.  tr_multi_ctx(m, i) gives the per-device context for every tr_get_* /
 * probe call of this header (local chain j of device i is global chain first_chain + lo_i + j). */
typedef struct tr_multi tr_multi;
int  tr_create_multi(const int32_t *devices, int32_t n, tr_multi **out);
void tr_destroy_multi(tr_multi *m);
const char *tr_multi_last_error(const tr_multi *m);     /* m == NULL: last failing tr_create_multi */
int32_t tr_multi_size(const tr_multi *m);
tr_ctx *tr_multi_ctx(tr_multi *m, int32_t i);
void tr_multi_chain_range(int32_t rank, int32_t world, int64_t n_total, int64_t *lo, int64_t *hi);
int  tr_multi_set_system(tr_multi *m, const tr_system *sys);
int  tr_multi_set_schedules(tr_multi *m, const tr_schedule *sched, int32_t n_sched);
int  tr_multi_set_options(tr_multi *m, int64_t trace_moves, int32_t record_points, int32_t record_snapshots, int32_t threads_per_chain);
/* seeds / sched_idx hold ALL n_chains entries; device i takes its range. */
int  tr_multi_init_chains_seeded(tr_multi *m, int64_t n_chains, int64_t first_chain, const int64_t *seeds,
                                 const int32_t *sched_idx, double space_len, int32_t max_prop_failures);
int  tr_multi_run(tr_multi *m, int64_t max_moves);       /* every device, asynchronously */
int  tr_multi_synchronize(tr_multi *m);
int  tr_multi_last_run_ms(tr_multi *m, float *ms_max);   /* max over devices of tr_last_run_ms */
/* The one exchange of the path (what Space.writeMinEnergy needs over all chains, Space.java:836-844): every device
 * packs its chains' minima, an NCCL all-gather leaves all n_chains records on every device, each device finds the
 * global winner (lowest energy, ties to the lowest chain id), and the owner broadcasts the winning structure.
 * all: [n_chains] records ordered by global chain id, or NULL; best: the winner, or NULL; best_sites: [S][3], or NULL. */
int  tr_gather_minima(tr_multi *m, tr_minimum *all, tr_minimum *best, double *best_sites);
/* NCCL all-gather of EVERY chain's minimum structure: sites[n_chains][S][3] (BASELINE config #4). */
int  tr_gather_min_structures(tr_multi *m, double *sites);

/* ---- parity probes: Space.calcEnergy and the eStart/eEnd loop on supplied geometries ----- */
int  tr_total_energy(tr_ctx *ctx, const double *sites /* [n][S][3] */, int64_t n, double *energy /* [n] */);
/* For structure i: particle mol[i] is given trial site coordinates trial[i][TR_MAX_MOL_SITES][3]
 * (first atoms.size() rows used); returns eStart and eEnd of Space.java:711-718 and, if d_e is
 * not NULL, eEnd - eStart exactly as the anneal kernel forms it for the Metropolis test. */
int  tr_delta_energy(tr_ctx *ctx, const double *sites, const int32_t *mol, const double *trial,
                     int64_t n, double *e_start, double *e_end, double *d_e);
/* First k 32-bit outputs of the three streams for each seed after the Main.java:34-37 fan-out:
 * out[n][3][k], drawn through the same shared-memory ring generator the anneal kernel uses
 * (in move-sized bursts, with one suspend/resume in the middle). */
int  tr_rng_words(tr_ctx *ctx, const int64_t *seeds, int64_t n, int32_t k, uint32_t *out);

/* ---- measurement --------------------------------------------------------------------------- */
/* DFMA micro-benchmark: sustained FP64 FMA rate of this device in TFLOP/s (2 flops per FMA),
 * the denominator of the pair-evaluation roofline. */
int  tr_measure_fp64_peak(tr_ctx *ctx, double *tflops, double *sm_clock_mhz_est);

#ifdef __cplusplus
}
#endif
#endif
