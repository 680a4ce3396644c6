// Device-side PODs shared by the kernels and the C-ABI glue.
#pragma once
#include <stdint.h>
#include "../../include/transrot_b200.h"
#include "mt19937.cuh"

namespace tr {

constexpr double K_VALUE = 332.0637133;            // Atom.java:10
constexpr double BOLTZMANN_CONSTANT = 0.0019872;   // Space.java:32
constexpr double JAVA_TWO_PI = 2 * 3.141592653589793;   // `2 * Math.PI` (Main.java:204, Space.java:112)

// Chain-independent tables, all in global memory (copied to shared memory per block).
struct DevSystem {
    int n_mol, n_sites, n_types, n_moveable;
    int has_exp;                 // any pair with A != 0: keep the A*exp(-B*r) term
    int pad_;
    int n_pos;                   // positions = 32 * W * SPL of the kernel variant in use
    int n_lj_pos;                // positions [0, n_lj_pos) hold the sites with any non-zero C or D
    FastMod pick_mod;            // nextInt(n_moveable) without a divide
    const int *mol_off;          // [n_mol+1]
    const int *site_info;        // [n_sites]  mol | type << 16 | index-in-particle << 24
    int row_stride;              // bytes per row of the shared-memory pair tables = (n_types + 1) * 16
    int zero_col;                // byte offset of the all-zero column = n_types * 16
    const int *pos_site;         // [n_pos]    site held at each (thread, slot) position, or -1
    const int *site_pos;         // [n_sites]  inverse of pos_site
    const int *pos_info;         // [n_pos]    site_info of the site at each position; empty: mol = 0xffff, type = n_types
    const double *site_q;        // [n_sites]
    const double2 *pair_cd;      // [n_types*n_types] {C, D}
    const double2 *pair_ab;      // [n_types*n_types] {A, B}
    const int *moveable;         // [n_moveable]
    const double *mol_radius;    // [n_mol]
    const double *site_def;      // [n_sites*3]
    const double *site_mass;     // [n_sites]
    const int *site_ghost;       // [n_sites]
};

// Everything Main.sawtoothAnneal keeps in locals (Main.java:178-194), so a chain can be
// suspended after any move and resumed by the next launch.
struct ChainCtrl {
    double t, saveT, delT;             // Main.java:178-179,188
    double maxD, maxRot;               // :185-186 (divided by 100 in the finale, :249-250)
    double probRot, probTrans;         // :183-184
    double running;                    // startingEnergy (:215)
    double totalE, totalSqE;           // :193-194
    double space_len;                  // Config.spaceLen after propagate retries
    double start_energy;
    double min_energy;                 // Space.minEnergy
    long long moves, evaluated, accepted_all, pair_evals;
    int x, y, z;                       // loop indices of :187,191,195
    int ptsPerTooth;                   // :182,244
    int accepted, total;               // :189-190
    int isDoubleCycle;                 // :181
    int started;                       // initial calcEnergy + write(0) done
    int done;
    int n_points, n_snaps;             // records emitted
    int min_tooth;                     // Space.minEnergyToothNum
    int sched;                         // index into the schedule table
    int prop_retries;
    MtCursor rng[3];                   // streams M (Main.r), S (Space.r), R (Molecule.r)
    int pad_;
};

struct KParams {
    DevSystem sys;
    const tr_schedule *sched;
    ChainCtrl *ctrl;            // [n_chains]
    double *sites;              // [n_chains][S][3]
    double *com;                // [n_chains][N][3]
    uint32_t *mt;               // [n_chains][3][2][624] ping-pong blocks
    tr_point_rec *points;       // [n_chains][max_points] or null
    double *snap_energy;        // [n_chains][max_snaps]
    int *snap_tooth;            // [n_chains][max_snaps]
    double *snap_sites;         // [n_chains][max_snaps][S][3] or null
    double *movie_sites;        // [n_chains][max_points][S][3] or null
    tr_trace_rec *trace;        // [n_chains][trace_moves] or null
    double *pair_cache;         // [n_chains][tri_count] particle-pair energies (crew2 kernel) or null
    long long n_chains;
    long long max_moves;        // <= 0: run to completion
    long long trace_moves;
    int max_points, max_snaps;
};

enum { STREAM_M = 0, STREAM_S = 1, STREAM_R = 2 };

}  // namespace tr
