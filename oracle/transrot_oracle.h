/*
 * transrot_oracle.h — CPU restatement of TransRot's annealing hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under transrot_b200/ may include, link or
 * load this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs use it, as the checker / the CPU baseline.
 *
 * PARITY UNPINNED: the reference (steventopper/TransRot v1.9.2) ships no tests,
 * no golden outputs and cannot run in this image (no JVM).  This restatement is
 * pinned only by (i) the canonical mt19937ar known-answer vector and
 * numpy.random.RandomState for the RNG, (ii) the six combination-rule goldens in
 * the reference's src/interaction_params.txt:6,7,12, and (iii) an independent
 * Python restatement (oracle/pyref.py).  Transcendentals use glibc libm where the
 * reference uses java.lang.Math (both <= 1 ulp; not bit-identical by contract).
 *
 * All file:line citations are into /root/reference/src/main/.
 */
#ifndef TRANSROT_ORACLE_H
#define TRANSROT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OR_MT_N 624

/* commons-math3 3.6.1 MersenneTwister state. */
typedef struct or_mt {
    uint32_t mt[OR_MT_N];
    int32_t mti;
} or_mt;

void     or_mt_set_seed_int(or_mt *g, uint32_t seed);
void     or_mt_set_seed_array(or_mt *g, const uint32_t *key, int len);
void     or_mt_set_seed_long(or_mt *g, int64_t seed);
uint32_t or_mt_next(or_mt *g, int bits);
double   or_mt_next_double(or_mt *g);
int32_t  or_mt_next_int(or_mt *g, int32_t n);
int64_t  or_mt_next_long(or_mt *g);

/* Config.java:12-33 — the fields the hot path reads. */
typedef struct or_config {
    double  maxTemp;
    double  tempDecreasePerTooth;
    double  maxTrans;
    double  magwalkFactorTrans;
    double  magwalkProbTrans;
    double  maxRot;
    double  magwalkProbRot;
    double  spaceLen;
    int32_t movePerPt;
    int32_t ptsPerTooth;
    int32_t ptsAddedPerTooth;
    int32_t numTeeth;
    int32_t maxPropFailures;
    int32_t eqConfigs;
    int32_t finale;
    int32_t staticTemp;
    int32_t writeConfHeatCapacitiesEnabled;
    int32_t pad_;
} or_config;

/* One record per iteration of the z-loop at Main.java:195. */
typedef struct or_trace_rec {
    int32_t mol;        /* index into `space` of the picked particle            */
    int8_t  kind;       /* 0 = Space.move, 1 = Space.rotate                      */
    int8_t  magwalk;    /* 1 = magnified step chosen (Main.java:204,212)         */
    int8_t  bounds;     /* 1 = rejected by the 0.75*spaceLen box test            */
    int8_t  accepted;   /* values.getSecond()                                    */
    int32_t axis;       /* rotation axis c (Molecule.java:68) or -1              */
    int32_t drew_rand;  /* 1 = Metropolis nextDouble was consumed                */
    double  e_start;    /* eStart (0 when bounds-rejected)                       */
    double  e_end;      /* eEnd                                                  */
    double  dE;         /* eEnd - eStart as computed (even when then rejected)   */
    double  rand;       /* the Metropolis draw, NaN if not drawn                 */
    double  temp;       /* t in effect                                           */
    double  running;    /* startingEnergy after `+= values.getFirst()`           */
} or_trace_rec;

/* One record per iteration of the y-loop at Main.java:191. */
typedef struct or_point_rec {
    int32_t tooth;          /* x (the finale repeats numTeeth-1)                 */
    int32_t point;          /* y                                                 */
    int32_t accepted;       /* cumulative within the tooth (Main.java:189,223)   */
    int32_t total;          /* cumulative within the tooth                       */
    int32_t movie_written;  /* 0 when the `continue` at Main.java:228 skipped it */
    int32_t finale;         /* 1 inside the 0 K finale repeat                    */
    double  temp;           /* t passed to writeAcceptance (Main.java:225)       */
    double  totalEnergy;    /* Σ startingEnergy (Main.java:217)                  */
    double  totalSquaredEnergy;
    double  movieEnergy;    /* calcEnergy() inside writeMovie (Space.java:745)   */
} or_point_rec;

typedef struct or_space or_space;

or_space *or_new(void);
void      or_free(or_space *s);

/* readDB equivalent for one particle (Space.java:162-167): params = natoms x 9
 * (x y z a b c d q mass), ghost[i] = symbol contains '*'.  Returns species id. */
int  or_dbase_add(or_space *s, double radius, int natoms, const double *params, const int32_t *ghost);
/* Space.add(String,int) (Space.java:46-69): radius-sorted insertion into `list`. */
int  or_add(or_space *s, int species, int count);
/* Space.setupPairVals (Space.java:346-366). */
void or_setup_pair_vals(or_space *s);
/* Explicit table entry, as readParams would store it (Space.java:319-320); symmetric put. */
void or_set_pair(or_space *s, int type1, int type2, double A, double B, double C, double D);
int  or_num_types(const or_space *s);
void or_get_pair_table(const or_space *s, double *out /* ntypes*ntypes*4 */);

void or_set_config(or_space *s, const or_config *c);
void or_get_config(const or_space *s, or_config *c);
/* Main.java:34-37 seed fan-out into Main.r, Space.r, Molecule.r. */
void or_set_seed(or_space *s, int64_t seed);
void or_get_stream_seeds(const or_space *s, int64_t out[3]);   /* M, S, R */
void or_get_rng(const or_space *s, or_mt out[3]);               /* M, S, R */
void or_set_rng(or_space *s, const or_mt in[3]);

/* Space.propagate (Space.java:72-117): returns 1 on success. */
int  or_propagate(or_space *s);
/* Main.java:58-60: while(!propagate()) spaceLen *= 1.1.  Returns #retries. */
int  or_propagate_until_placed(or_space *s);
/* readInput equivalent (Space.java:200-220): groups of (species,count,frozen), coords = sites x 3. */
int  or_read_input(or_space *s, int ngroups, const int32_t *species, const int32_t *counts,
                   const int32_t *frozen, const double *coords);

int  or_num_molecules(const or_space *s);
int  or_num_moveable(const or_space *s);
int  or_num_sites(const or_space *s);
void or_get_layout(const or_space *s, int32_t *mol_species, int32_t *mol_site_off /* n+1 */,
                   int32_t *site_type, int32_t *moveable_idx);
void or_get_coords(const or_space *s, double *sites /* S*3 */, double *com /* N*3 */);
void or_set_coords(or_space *s, const double *sites, const double *com);

/* Space.calcEnergy (Space.java:648-658). */
double or_calc_energy(const or_space *s);
/* Test aid: calcEnergy plus the sum of |pair term|, the conditioning scale of that sum. */
double or_calc_energy_abs(const or_space *s, double *abs_sum);
/* Molecule.calcEnergy(m -> n) (Molecule.java:99-107). */
double or_mol_energy(const or_space *s, int m, int n);
/* Space.move / Space.rotate (Space.java:660-734); rec may be NULL. */
int  or_move(or_space *s, int m, double maxD, double temp, double *dE, or_trace_rec *rec);
int  or_rotate(or_space *s, int m, double maxRot, double temp, double *dE, or_trace_rec *rec);

/* Main.sawtoothAnneal (Main.java:177-263).  Buffers may be NULL / cap 0.
 *   trace      : first trace_cap moves
 *   points     : one per point, first points_cap
 *   snapshots  : write(n) calls: snap_tooth[k] = n, snap_energy[k], snap_coords[k*S*3]
 * Returns the final running energy (startingEnergy).  The initial write(0)
 * (Main.java:64) is NOT part of this call; use or_write_snapshot. */
typedef struct or_anneal_out {
    or_trace_rec *trace;   int64_t trace_cap;   int64_t trace_len;
    or_point_rec *points;  int64_t points_cap;  int64_t points_len;
    int32_t *snap_tooth;   double *snap_energy; double *snap_coords;
    int64_t snap_cap;      int64_t snap_len;
    double *trace_abs;     /* [trace_cap] or NULL: sum of |pair term| over the eStart and eEnd loops of each move,
                              the conditioning scale for comparing eStart/eEnd/dE (test aid, not in the reference) */
    int64_t moves;         /* iterations of the z-loop                     */
    int64_t evaluated;     /* moves that reached the energy loop           */
    int64_t pair_evals;    /* Atom.calcEnergy + calcTempEnergy calls made  */
} or_anneal_out;

double or_sawtooth_anneal(or_space *s, double startingEnergy, or_anneal_out *out);
/* Space.write(n) bookkeeping (Space.java:554,614-617): energy + min tracking. */
double or_write_snapshot(or_space *s, int n);
void   or_get_min(const or_space *s, double *minEnergy, int32_t *minTooth);

#ifdef __cplusplus
}
#endif
#endif
