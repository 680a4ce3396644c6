/*
 * transrot_oracle.c — plain-C restatement of the TransRot annealing hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see transrot_oracle.h).  PARITY UNPINNED: the
 * reference has no tests and no JVM exists in this image; see the header.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared (oracle/Makefile).
 * -ffp-contract=off matters: Java never fuses a*b+c.
 *
 * Citations are into /root/reference/src/main/.
 */
#include "transrot_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ RNG --- */
/* commons-math3 3.6.1 random/MersenneTwister.class + BitsStreamGenerator.class,
 * restated from the disassembly recorded in SURVEY.md Appendix B. */

#define MT_M 397

void or_mt_set_seed_int(or_mt *g, uint32_t seed)
{
    /* setSeed(int): mt[i] = 1812433253 * (mt[i-1] ^ (mt[i-1] >>> 30)) + i */
    g->mt[0] = seed;
    for (int i = 1; i < OR_MT_N; i++) {
        uint32_t p = g->mt[i - 1];
        g->mt[i] = 1812433253u * (p ^ (p >> 30)) + (uint32_t)i;
    }
    g->mti = OR_MT_N;
}

void or_mt_set_seed_array(or_mt *g, const uint32_t *key, int len)
{
    /* setSeed(int[]) == mt19937ar init_by_array, all arithmetic mod 2^32 */
    or_mt_set_seed_int(g, 19650218u);
    int i = 1, j = 0;
    int k = OR_MT_N > len ? OR_MT_N : len;
    for (; k != 0; k--) {
        uint32_t p = g->mt[i - 1];
        g->mt[i] = (g->mt[i] ^ ((p ^ (p >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
        i++; j++;
        if (i >= OR_MT_N) { g->mt[0] = g->mt[OR_MT_N - 1]; i = 1; }
        if (j >= len) j = 0;
    }
    for (k = OR_MT_N - 1; k != 0; k--) {
        uint32_t p = g->mt[i - 1];
        g->mt[i] = (g->mt[i] ^ ((p ^ (p >> 30)) * 1566083941u)) - (uint32_t)i;
        i++;
        if (i >= OR_MT_N) { g->mt[0] = g->mt[OR_MT_N - 1]; i = 1; }
    }
    g->mt[0] = 0x80000000u;
}

void or_mt_set_seed_long(or_mt *g, int64_t seed)
{
    /* setSeed(long) = setSeed(new int[]{(int)(s>>>32), (int)(s & 0xffffffff)}): high word first */
    uint32_t key[2] = { (uint32_t)((uint64_t)seed >> 32), (uint32_t)((uint64_t)seed & 0xffffffffu) };
    or_mt_set_seed_array(g, key, 2);
}

uint32_t or_mt_next(or_mt *g, int bits)
{
    static const uint32_t MAG01[2] = { 0u, 0x9908b0dfu };
    uint32_t y;
    if (g->mti >= OR_MT_N) {
        int k;
        for (k = 0; k < OR_MT_N - MT_M; k++) {
            y = (g->mt[k] & 0x80000000u) | (g->mt[k + 1] & 0x7fffffffu);
            g->mt[k] = g->mt[k + MT_M] ^ (y >> 1) ^ MAG01[y & 1u];
        }
        for (; k < OR_MT_N - 1; k++) {
            y = (g->mt[k] & 0x80000000u) | (g->mt[k + 1] & 0x7fffffffu);
            g->mt[k] = g->mt[k + (MT_M - OR_MT_N)] ^ (y >> 1) ^ MAG01[y & 1u];
        }
        y = (g->mt[OR_MT_N - 1] & 0x80000000u) | (g->mt[0] & 0x7fffffffu);
        g->mt[OR_MT_N - 1] = g->mt[MT_M - 1] ^ (y >> 1) ^ MAG01[y & 1u];
        g->mti = 0;
    }
    y = g->mt[g->mti++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return bits >= 32 ? y : (y >> (32 - bits));
}

double or_mt_next_double(or_mt *g)
{
    /* BitsStreamGenerator.nextDouble: (((long)next(26) << 26) | next(26)) * 0x1.0p-52 */
    uint64_t high = (uint64_t)or_mt_next(g, 26) << 26;
    uint64_t low = or_mt_next(g, 26);
    return (double)(int64_t)(high | low) * 0x1.0p-52;
}

int32_t or_mt_next_int(or_mt *g, int32_t n)
{
    /* BitsStreamGenerator.nextInt(int) */
    if ((n & -n) == n)
        return (int32_t)(((int64_t)n * (int64_t)or_mt_next(g, 31)) >> 31);
    int32_t bits, val;
    do {
        bits = (int32_t)or_mt_next(g, 31);
        val = bits % n;
        /* `bits - val + (n-1) < 0` in 32-bit two's-complement (Java int overflow) */
    } while ((int32_t)((uint32_t)bits - (uint32_t)val + (uint32_t)(n - 1)) < 0);
    return val;
}

int64_t or_mt_next_long(or_mt *g)
{
    uint64_t high = (uint64_t)or_mt_next(g, 32) << 32;
    uint64_t low = or_mt_next(g, 32);
    return (int64_t)(high | low);
}

/* ---------------------------------------------------------- data model --- */

#define K_VALUE 332.0637133            /* Atom.java:10 */
#define BOLTZMANN_CONSTANT 0.0019872   /* Space.java:32 */
#define JAVA_MATH_E 2.718281828459045
#define JAVA_MATH_PI 3.141592653589793

/* HotSpot evaluates Math.pow(x, 2.0) as x*x (LibraryCallKit::inline_math_pow and the
 * x86_64 pow stub both special-case y == 2); other exponents go to the <=1ulp pow. */
static inline double java_pow(double x, double y)
{
    if (y == 2.0) return x * x;
    return pow(x, y);
}

typedef struct atom {     /* Atom.java:8-32 */
    double x, y, z;
    double defX, defY, defZ;
    double tempx, tempy, tempz;
    double a, b, c, d, q, kTimesQ, mass;
    int type;   /* stands for the UUID: one id per dbase atom slot (Atom.java:35,54,73) */
    int ghost;  /* symbol.contains("*") */
} atom;

typedef struct molecule {  /* Molecule.java:15-24 */
    int species;
    double x, y, z, tempx, tempy, tempz, radius;
    int natoms;
    atom *atoms;
} molecule;

struct or_space {
    molecule **dbase; int ndbase;
    molecule **list;  int nlist;
    molecule **space; int nspace;
    molecule **moveable; int nmoveable;
    int ntypes;
    double *pair;      /* pairwiseDbase: [type1][type2][A,B,C,D] */
    or_mt rM, rS, rR;  /* Main.r, Space.r, Molecule.r */
    int64_t stream_seed[3];
    or_config cfg;
    double minEnergy; int minEnergyToothNum;   /* Space.java:23-24 */
    int64_t pair_evals;
    double abs_acc;    /* test aid: running sum of |pair term| */
};

static void atom_reset_temps(atom *a) { a->tempx = a->x; a->tempy = a->y; a->tempz = a->z; } /* Atom.java:121 */
static void atom_set_temps(atom *a)   { a->x = a->tempx; a->y = a->tempy; a->z = a->tempz; } /* Atom.java:127 */

static void mol_reset_temps(molecule *m)   /* Molecule.java:119-126 */
{
    m->tempx = m->x; m->tempy = m->y; m->tempz = m->z;
    for (int i = 0; i < m->natoms; i++) atom_reset_temps(&m->atoms[i]);
}

static void mol_set_temps(molecule *m)     /* Molecule.java:128-135 */
{
    m->x = m->tempx; m->y = m->tempy; m->z = m->tempz;
    for (int i = 0; i < m->natoms; i++) atom_set_temps(&m->atoms[i]);
}

static void mol_set_center_of_mass(molecule *m)   /* Molecule.java:137-156 */
{
    double totMass = 0, massPosX = 0, massPosY = 0, massPosZ = 0;
    for (int i = 0; i < m->natoms; i++) {
        atom *a = &m->atoms[i];
        if (a->ghost) continue;
        totMass += a->mass;
        massPosX += a->x * a->mass;
        massPosY += a->y * a->mass;
        massPosZ += a->z * a->mass;
    }
    m->x = massPosX / totMass;
    m->y = massPosY / totMass;
    m->z = massPosZ / totMass;
    mol_reset_temps(m);
}

static void mol_put(molecule *m, double x, double y, double z)   /* Molecule.java:51-62 */
{
    m->x = x; m->y = y; m->z = z;
    for (int i = 0; i < m->natoms; i++) {
        atom *a = &m->atoms[i];
        a->x = a->defX + x;
        a->y = a->defY + y;
        a->z = a->defZ + z;
    }
    mol_reset_temps(m);
    mol_set_center_of_mass(m);
}

static molecule *mol_copy(const molecule *src)   /* Molecule(Molecule) Molecule.java:33-44 */
{
    molecule *m = (molecule *)malloc(sizeof(molecule));
    *m = *src;
    m->atoms = (atom *)malloc(sizeof(atom) * (size_t)src->natoms);
    memcpy(m->atoms, src->atoms, sizeof(atom) * (size_t)src->natoms);
    for (int i = 0; i < m->natoms; i++) atom_reset_temps(&m->atoms[i]);  /* Atom(Atom) Atom.java:67-88 */
    mol_reset_temps(m);
    return m;
}

static void mol_free(molecule *m) { if (m) { free(m->atoms); free(m); } }

or_space *or_new(void)
{
    or_space *s = (or_space *)calloc(1, sizeof(or_space));
    s->minEnergy = 1.7976931348623157e308;   /* Double.MAX_VALUE, Space.java:23 */
    return s;
}

void or_free(or_space *s)
{
    if (!s) return;
    for (int i = 0; i < s->ndbase; i++) mol_free(s->dbase[i]);
    for (int i = 0; i < s->nlist; i++) mol_free(s->list[i]);
    for (int i = 0; i < s->nspace; i++) mol_free(s->space[i]);
    free(s->dbase); free(s->list); free(s->space); free(s->moveable); free(s->pair);
    free(s);
}

int or_dbase_add(or_space *s, double radius, int natoms, const double *p, const int32_t *ghost)
{
    /* new Atom(name, params) (Atom.java:53-70) + new Molecule(name, radius, atoms) (Molecule.java:25-31) */
    molecule *m = (molecule *)calloc(1, sizeof(molecule));
    m->species = s->ndbase;
    m->radius = radius;
    m->natoms = natoms;
    m->atoms = (atom *)calloc((size_t)natoms, sizeof(atom));
    for (int i = 0; i < natoms; i++) {
        atom *a = &m->atoms[i];
        const double *q = p + 9 * i;
        a->x = q[0]; a->y = q[1]; a->z = q[2];
        a->defX = a->x; a->defY = a->y; a->defZ = a->z;
        atom_reset_temps(a);
        a->a = q[3]; a->b = q[4]; a->c = q[5]; a->d = q[6]; a->q = q[7];
        a->mass = q[8];
        a->kTimesQ = K_VALUE * a->q;
        a->type = s->ntypes++;
        a->ghost = ghost ? ghost[i] : 0;
    }
    mol_reset_temps(m);
    mol_set_center_of_mass(m);
    s->dbase = (molecule **)realloc(s->dbase, sizeof(molecule *) * (size_t)(s->ndbase + 1));
    s->dbase[s->ndbase] = m;
    return s->ndbase++;
}

static void list_insert(or_space *s, int pos, molecule *m)
{
    s->list = (molecule **)realloc(s->list, sizeof(molecule *) * (size_t)(s->nlist + 1));
    memmove(&s->list[pos + 1], &s->list[pos], sizeof(molecule *) * (size_t)(s->nlist - pos));
    s->list[pos] = m;
    s->nlist++;
}

int or_add(or_space *s, int species, int count)   /* Space.java:46-69 */
{
    if (species < 0 || species >= s->ndbase) return 0;
    for (int c = 0; c < count; c++) {
        molecule *m = mol_copy(s->dbase[species]);
        int placed = 0;
        for (int n = 0; n < s->nlist; n++) {
            if (m->radius > s->list[n]->radius) { list_insert(s, n, m); placed = 1; break; }
        }
        if (!placed) list_insert(s, s->nlist, m);
    }
    return 1;
}

static void pair_alloc(or_space *s)
{
    if (!s->pair) s->pair = (double *)calloc((size_t)s->ntypes * (size_t)s->ntypes * 4, sizeof(double));
}

void or_setup_pair_vals(or_space *s)   /* Space.java:346-366 */
{
    pair_alloc(s);
    for (int i = 0; i < s->ndbase; i++)
        for (int ai = 0; ai < s->dbase[i]->natoms; ai++)
            for (int j = 0; j < s->ndbase; j++)
                for (int aj = 0; aj < s->dbase[j]->natoms; aj++) {
                    const atom *a1 = &s->dbase[i]->atoms[ai], *a2 = &s->dbase[j]->atoms[aj];
                    double A = sqrt(a1->a * a2->a);
                    double B = (a1->b + a2->b) / 2;
                    double C = sqrt(a1->c * a2->c);
                    double D = sqrt(a1->d * a2->d);
                    double *v1 = s->pair + ((size_t)a1->type * (size_t)s->ntypes + (size_t)a2->type) * 4;
                    double *v2 = s->pair + ((size_t)a2->type * (size_t)s->ntypes + (size_t)a1->type) * 4;
                    v1[0] = A; v1[1] = B; v1[2] = C; v1[3] = D;
                    v2[0] = A; v2[1] = B; v2[2] = C; v2[3] = D;
                }
}

void or_set_pair(or_space *s, int t1, int t2, double A, double B, double C, double D)
{
    pair_alloc(s);
    double *v1 = s->pair + ((size_t)t1 * (size_t)s->ntypes + (size_t)t2) * 4;
    double *v2 = s->pair + ((size_t)t2 * (size_t)s->ntypes + (size_t)t1) * 4;
    v1[0] = A; v1[1] = B; v1[2] = C; v1[3] = D;
    v2[0] = A; v2[1] = B; v2[2] = C; v2[3] = D;
}

int or_num_types(const or_space *s) { return s->ntypes; }

void or_get_pair_table(const or_space *s, double *out)
{
    memcpy(out, s->pair, sizeof(double) * (size_t)s->ntypes * (size_t)s->ntypes * 4);
}

void or_set_config(or_space *s, const or_config *c) { s->cfg = *c; }
void or_get_config(const or_space *s, or_config *c) { *c = s->cfg; }

void or_set_seed(or_space *s, int64_t seed)   /* Main.java:24,34-37 */
{
    or_mt g;
    or_mt_set_seed_long(&g, seed);
    s->stream_seed[0] = or_mt_next_long(&g); or_mt_set_seed_long(&s->rM, s->stream_seed[0]);
    s->stream_seed[1] = or_mt_next_long(&g); or_mt_set_seed_long(&s->rS, s->stream_seed[1]);
    s->stream_seed[2] = or_mt_next_long(&g); or_mt_set_seed_long(&s->rR, s->stream_seed[2]);
}

void or_get_stream_seeds(const or_space *s, int64_t out[3]) { memcpy(out, s->stream_seed, sizeof(int64_t) * 3); }
void or_get_rng(const or_space *s, or_mt out[3]) { out[0] = s->rM; out[1] = s->rS; out[2] = s->rR; }
void or_set_rng(or_space *s, const or_mt in[3])  { s->rM = in[0]; s->rS = in[1]; s->rR = in[2]; }

/* ------------------------------------------------------------- energy --- */

static inline double atom_energy_at(or_space *s, const atom *self, double sx, double sy, double sz, const atom *other)
{
    /* Atom.calcEnergy / calcTempEnergy (Atom.java:91-119): `self` is `this` */
    const double *v = s->pair + ((size_t)self->type * (size_t)s->ntypes + (size_t)other->type) * 4;
    double A = v[0], B = v[1], C = v[2], D = v[3];
    double r = sqrt(java_pow(other->x - sx, 2) + java_pow(other->y - sy, 2) + java_pow(other->z - sz, 2));
    s->pair_evals++;
    double e = (self->kTimesQ * other->q / r) + A * exp(-1 * B * r) - (C / java_pow(r, 6)) + (D / java_pow(r, 12));
    s->abs_acc += fabs(e);
    return e;
}

static double mol_calc_energy(or_space *s, const molecule *m, const molecule *n)   /* Molecule.java:99-107 */
{
    double energy = 0;
    for (int i = 0; i < m->natoms; i++)
        for (int j = 0; j < n->natoms; j++)
            energy += atom_energy_at(s, &m->atoms[i], m->atoms[i].x, m->atoms[i].y, m->atoms[i].z, &n->atoms[j]);
    return energy;
}

static double mol_calc_temp_energy(or_space *s, const molecule *m, const molecule *n)   /* Molecule.java:109-117 */
{
    double energy = 0;
    for (int i = 0; i < m->natoms; i++)
        for (int j = 0; j < n->natoms; j++)
            energy += atom_energy_at(s, &m->atoms[i], m->atoms[i].tempx, m->atoms[i].tempy, m->atoms[i].tempz, &n->atoms[j]);
    return energy;
}

double or_calc_energy(const or_space *cs)   /* Space.java:648-658 */
{
    or_space *s = (or_space *)cs;
    double totEnergy = 0;
    for (int x = 0; x < s->nspace; x++)
        for (int y = 0; y < s->nspace; y++)
            if (y > x) totEnergy += mol_calc_energy(s, s->space[x], s->space[y]);
    return totEnergy;
}

double or_calc_energy_abs(const or_space *cs, double *abs_sum)   /* test aid: calcEnergy + sum of |pair term| */
{
    or_space *s = (or_space *)cs;
    s->abs_acc = 0.0;
    double e = or_calc_energy(cs);
    if (abs_sum) *abs_sum = s->abs_acc;
    return e;
}

double or_mol_energy(const or_space *cs, int m, int n)
{
    or_space *s = (or_space *)cs;
    return mol_calc_energy(s, s->space[m], s->space[n]);
}

/* -------------------------------------------------------------- moves --- */

static int mol_rotate_temp(or_space *s, molecule *m, double maxRot, double *rotOut)   /* Molecule.java:64-97 */
{
    if (m->natoms <= 1) return -1;
    double rotAmount = (or_mt_next_double(&s->rR) - 0.5) * 2 * maxRot;
    int c = or_mt_next_int(&s->rR, 3);
    double x = m->x, y = m->y, z = m->z;
    for (int i = 0; i < m->natoms; i++) {
        atom *a = &m->atoms[i];
        a->x -= x; a->y -= y; a->z -= z;
        switch (c) {
        case 0:
            a->tempx = a->x;
            a->tempy = cos(rotAmount) * a->y + sin(rotAmount) * a->z;
            a->tempz = -1 * sin(rotAmount) * a->y + cos(rotAmount) * a->z;
            break;
        case 1:
            a->tempx = cos(rotAmount) * a->x - sin(rotAmount) * a->z;
            a->tempy = a->y;
            a->tempz = sin(rotAmount) * a->x + cos(rotAmount) * a->z;
            break;
        case 2:
            a->tempx = cos(rotAmount) * a->x + sin(rotAmount) * a->y;
            a->tempy = -1 * sin(rotAmount) * a->x + cos(rotAmount) * a->y;
            a->tempz = a->z;
            break;
        }
        a->x += x; a->y += y; a->z += z;
        a->tempx += x; a->tempy += y; a->tempz += z;
    }
    if (rotOut) *rotOut = rotAmount;
    return c;
}

static int out_of_bounds(const or_space *s, const molecule *m)   /* Space.java:663,706 */
{
    double L = s->cfg.spaceLen;
    for (int i = 0; i < m->natoms; i++) {
        const atom *a = &m->atoms[i];
        if (a->tempx > L * 0.75 || a->tempx < L * -0.75 || a->tempy > L * 0.75 || a->tempy < L * -0.75 ||
            a->tempz > L * 0.75 || a->tempz < L * -0.75)
            return 1;
    }
    return 0;
}

static void rec_init(or_trace_rec *rec, int mol, int kind, double temp)
{
    if (!rec) return;
    rec->mol = mol; rec->kind = (int8_t)kind; rec->bounds = 0; rec->accepted = 0;
    rec->axis = -1; rec->drew_rand = 0;
    rec->e_start = 0; rec->e_end = 0; rec->dE = 0; rec->rand = NAN; rec->temp = temp;
}

/* Shared tail of Space.rotate (:668-689) and Space.move (:711-733). */
static int evaluate_and_decide(or_space *s, int mi, double temp, double *dE, or_trace_rec *rec)
{
    molecule *m = s->space[mi];
    double eStart = 0, eEnd = 0;
    for (int n = 0; n < s->nspace; n++) {
        if (s->space[n] != m) {
            eStart += mol_calc_energy(s, m, s->space[n]);
            eEnd += mol_calc_temp_energy(s, m, s->space[n]);
        }
    }
    if (rec) { rec->e_start = eStart; rec->e_end = eEnd; rec->dE = eEnd - eStart; }
    if (eEnd - eStart > 0) {
        double rnd = or_mt_next_double(&s->rS);
        if (rec) { rec->drew_rand = 1; rec->rand = rnd; }
        if (temp == 0) {
            mol_reset_temps(m);
            *dE = 0.0;
            return 0;
        }
        double test = java_pow(JAVA_MATH_E, -1 * (eEnd - eStart) / (BOLTZMANN_CONSTANT * temp));
        if (rnd > test) {
            mol_reset_temps(m);
            *dE = 0.0;
            return 0;
        }
    }
    mol_set_temps(m);
    *dE = eEnd - eStart;
    if (rec) rec->accepted = 1;
    return 1;
}

int or_rotate(or_space *s, int mi, double maxRot, double temp, double *dE, or_trace_rec *rec)   /* Space.java:660-690 */
{
    molecule *m = s->space[mi];
    rec_init(rec, mi, 1, temp);
    int c = mol_rotate_temp(s, m, maxRot, NULL);
    if (rec) rec->axis = c;
    if (out_of_bounds(s, m)) {
        mol_reset_temps(m);
        if (rec) rec->bounds = 1;
        *dE = 0.0;
        return 0;
    }
    return evaluate_and_decide(s, mi, temp, dE, rec);
}

int or_move(or_space *s, int mi, double maxD, double temp, double *dE, or_trace_rec *rec)   /* Space.java:692-734 */
{
    molecule *m = s->space[mi];
    rec_init(rec, mi, 0, temp);
    double x = (or_mt_next_double(&s->rS) * 2 * maxD) - maxD;
    double y = (or_mt_next_double(&s->rS) * 2 * maxD) - maxD;
    double z = (or_mt_next_double(&s->rS) * 2 * maxD) - maxD;
    m->tempx += x; m->tempy += y; m->tempz += z;
    for (int i = 0; i < m->natoms; i++) {
        atom *a = &m->atoms[i];
        a->tempx += x; a->tempy += y; a->tempz += z;
    }
    if (out_of_bounds(s, m)) {
        mol_reset_temps(m);
        if (rec) rec->bounds = 1;
        *dE = 0.0;
        return 0;
    }
    return evaluate_and_decide(s, mi, temp, dE, rec);
}

/* ------------------------------------------------------- initial state --- */

static void space_push(or_space *s, molecule *m, int moveable)
{
    s->space = (molecule **)realloc(s->space, sizeof(molecule *) * (size_t)(s->nspace + 1));
    s->space[s->nspace++] = m;
    if (moveable) {
        s->moveable = (molecule **)realloc(s->moveable, sizeof(molecule *) * (size_t)(s->nmoveable + 1));
        s->moveable[s->nmoveable++] = m;
    }
}

int or_propagate(or_space *s)   /* Space.java:72-117 */
{
    for (int li = 0; li < s->nlist; li++) {
        molecule *m = s->list[li];
        int c = 0;
        for (;;) {
            double L = s->cfg.spaceLen;
            double x = (or_mt_next_double(&s->rS) * L) - (L / 2);
            double y = (or_mt_next_double(&s->rS) * L) - (L / 2);
            double z = (or_mt_next_double(&s->rS) * L) - (L / 2);
            if (s->nspace == 0) {
                mol_put(m, x, y, z);
                space_push(s, m, 1);
                break;
            }
            int ok = 1;
            for (int n = 0; n < s->nspace; n++) {
                const molecule *o = s->space[n];
                double min = m->radius + o->radius;
                double d = sqrt(java_pow(o->x - x, 2) + java_pow(o->y - y, 2) + java_pow(o->z - z, 2));
                if (d < min) { c++; ok = 0; break; }
            }
            if (ok) {
                mol_put(m, x, y, z);
                space_push(s, m, 1);
                break;
            }
            if (c >= s->cfg.maxPropFailures) {
                /* space.clear(); moveableMolecules.clear(): molecules stay owned by `list` */
                s->nspace = 0;
                s->nmoveable = 0;
                return 0;
            }
        }
    }
    for (int n = 0; n < s->nspace; n++) {
        mol_rotate_temp(s, s->space[n], 2 * JAVA_MATH_PI, NULL);
        mol_set_temps(s->space[n]);
    }
    s->nlist = 0;   /* list.clear(): ownership moved to `space` */
    return 1;
}

int or_propagate_until_placed(or_space *s)   /* Main.java:58-60 */
{
    int retries = 0;
    while (!or_propagate(s)) {
        s->cfg.spaceLen *= 1.1;
        retries++;
    }
    return retries;
}

int or_read_input(or_space *s, int ngroups, const int32_t *species, const int32_t *counts,
                  const int32_t *frozen, const double *coords)   /* Space.java:189-221 */
{
    const double *p = coords;
    for (int g = 0; g < ngroups; g++) {
        if (species[g] < 0 || species[g] >= s->ndbase) return 0;
        for (int i = 0; i < counts[g]; i++) {
            molecule *m = mol_copy(s->dbase[species[g]]);
            for (int j = 0; j < m->natoms; j++) {
                m->atoms[j].x = p[0]; m->atoms[j].y = p[1]; m->atoms[j].z = p[2];
                p += 3;
            }
            mol_set_center_of_mass(m);
            space_push(s, m, !frozen[g]);
        }
    }
    return 1;
}

int or_num_molecules(const or_space *s) { return s->nspace; }
int or_num_moveable(const or_space *s) { return s->nmoveable; }

int or_num_sites(const or_space *s)
{
    int n = 0;
    for (int i = 0; i < s->nspace; i++) n += s->space[i]->natoms;
    return n;
}

void or_get_layout(const or_space *s, int32_t *mol_species, int32_t *mol_site_off, int32_t *site_type, int32_t *moveable_idx)
{
    int off = 0;
    for (int i = 0; i < s->nspace; i++) {
        if (mol_species) mol_species[i] = s->space[i]->species;
        if (mol_site_off) mol_site_off[i] = off;
        for (int j = 0; j < s->space[i]->natoms; j++)
            if (site_type) site_type[off + j] = s->space[i]->atoms[j].type;
        off += s->space[i]->natoms;
    }
    if (mol_site_off) mol_site_off[s->nspace] = off;
    if (moveable_idx)
        for (int k = 0; k < s->nmoveable; k++)
            for (int i = 0; i < s->nspace; i++)
                if (s->space[i] == s->moveable[k]) { moveable_idx[k] = i; break; }
}

void or_get_coords(const or_space *s, double *sites, double *com)
{
    int off = 0;
    for (int i = 0; i < s->nspace; i++) {
        const molecule *m = s->space[i];
        if (com) { com[3 * i] = m->x; com[3 * i + 1] = m->y; com[3 * i + 2] = m->z; }
        for (int j = 0; j < m->natoms; j++, off++)
            if (sites) { sites[3 * off] = m->atoms[j].x; sites[3 * off + 1] = m->atoms[j].y; sites[3 * off + 2] = m->atoms[j].z; }
    }
}

void or_set_coords(or_space *s, const double *sites, const double *com)
{
    int off = 0;
    for (int i = 0; i < s->nspace; i++) {
        molecule *m = s->space[i];
        if (com) { m->x = com[3 * i]; m->y = com[3 * i + 1]; m->z = com[3 * i + 2]; }
        for (int j = 0; j < m->natoms; j++, off++)
            if (sites) { m->atoms[j].x = sites[3 * off]; m->atoms[j].y = sites[3 * off + 1]; m->atoms[j].z = sites[3 * off + 2]; }
        mol_reset_temps(m);
    }
}

/* ----------------------------------------------------------- schedule --- */

double or_write_snapshot(or_space *s, int n)   /* Space.java:554,614-617 */
{
    double toothEnergy = or_calc_energy(s);
    if (toothEnergy < s->minEnergy) {
        s->minEnergy = toothEnergy;
        s->minEnergyToothNum = n;
    }
    return toothEnergy;
}

void or_get_min(const or_space *s, double *minEnergy, int32_t *minTooth)
{
    *minEnergy = s->minEnergy;
    *minTooth = s->minEnergyToothNum;
}

static int mol_index(const or_space *s, const molecule *m)
{
    for (int i = 0; i < s->nspace; i++) if (s->space[i] == m) return i;
    return -1;
}

double or_sawtooth_anneal(or_space *s, double startingEnergy, or_anneal_out *out)   /* Main.java:177-263 */
{
    const or_config *C = &s->cfg;
    int numTeeth = C->numTeeth;
    int ptsPerTooth = C->ptsPerTooth;
    if (C->staticTemp) { numTeeth = 1; ptsPerTooth = 1; }   /* Main.java:74-76 */
    double t = C->maxTemp;
    double saveT = t;
    int isDoubleCycle = 0;
    double magwalkProbRot = C->magwalkProbRot;
    double magwalkProbTrans = C->magwalkProbTrans;
    double maxD = C->maxTrans;
    double maxRot = C->maxRot;
    int nsites = or_num_sites(s);
    or_anneal_out dummy;
    if (!out) { memset(&dummy, 0, sizeof dummy); out = &dummy; }
    out->trace_len = out->points_len = out->snap_len = 0;
    out->moves = out->evaluated = out->pair_evals = 0;

    for (int x = 0; x < numTeeth; x++) {
        double delT = t / (ptsPerTooth - 1);
        int accepted = 0;
        int total = 0;
        for (int y = 0; y < ptsPerTooth; y++) {
            double totalEnergy = 0;
            double totalSquaredEnergy = 0;
            for (int z = 0; z < C->movePerPt; z++) {
                total++;
                int mi = mol_index(s, s->moveable[or_mt_next_int(&s->rS, s->nmoveable)]);   /* Space.java:735-738 */
                molecule *m = s->space[mi];
                or_trace_rec rec;
                double dE;
                int acc;
                int magwalk;
                int64_t pe = s->pair_evals;
                s->abs_acc = 0.0;
                if (or_mt_next_double(&s->rM) >= 0.5 && m->natoms > 1) {
                    if (or_mt_next_double(&s->rM) >= magwalkProbRot) {
                        magwalk = 0;
                        acc = or_rotate(s, mi, maxRot, t, &dE, &rec);
                    } else {
                        magwalk = 1;
                        acc = or_rotate(s, mi, 2 * JAVA_MATH_PI, t, &dE, &rec);
                    }
                } else {
                    if (or_mt_next_double(&s->rM) >= magwalkProbTrans) {
                        magwalk = 0;
                        acc = or_move(s, mi, maxD, t, &dE, &rec);
                    } else {
                        magwalk = 1;
                        acc = or_move(s, mi, maxD * C->magwalkFactorTrans, t, &dE, &rec);
                    }
                }
                startingEnergy += dE;
                if (!C->staticTemp || z > C->eqConfigs) {
                    totalEnergy += startingEnergy;
                    totalSquaredEnergy += java_pow(startingEnergy, 2);
                }
                accepted += acc;
                out->moves++;
                out->pair_evals += s->pair_evals - pe;
                if (!rec.bounds) out->evaluated++;
                if (out->trace && out->trace_len < out->trace_cap) {
                    rec.magwalk = (int8_t)magwalk;
                    rec.running = startingEnergy;
                    if (out->trace_abs) out->trace_abs[out->trace_len] = s->abs_acc;
                    out->trace[out->trace_len++] = rec;
                }
            }
            or_point_rec *pr = (out->points && out->points_len < out->points_cap) ? &out->points[out->points_len] : NULL;
            if (pr) {
                pr->tooth = x; pr->point = y; pr->accepted = accepted; pr->total = total;
                pr->finale = isDoubleCycle; pr->temp = t;
                pr->totalEnergy = totalEnergy; pr->totalSquaredEnergy = totalSquaredEnergy;
                pr->movie_written = 0; pr->movieEnergy = NAN;
                out->points_len++;
            }
            if (C->writeConfHeatCapacitiesEnabled) {
                /* BigDecimal(t).setScale(2, HALF_UP) == 0  <=>  t < 0.005 for t >= 0 (exact binary value rounded) */
                if (t < 0.005 && t > -0.005) continue;   /* Main.java:227-228 */
            }
            if (!C->staticTemp) t -= delT;
            if (t < 0) t = 0;
            double me = or_calc_energy(s);   /* writeMovie -> calcEnergy (Space.java:745) */
            if (pr) { pr->movie_written = 1; pr->movieEnergy = me; }
        }
        saveT *= C->tempDecreasePerTooth;
        t = saveT;
        ptsPerTooth += C->ptsAddedPerTooth;
        if (x == numTeeth - 1 && !isDoubleCycle && C->finale) {
            t = 0;
            isDoubleCycle = 1;
            x--;
            maxD /= 100;
            maxRot /= 100;
            magwalkProbRot = 0;
            magwalkProbTrans = 0;
        } else {
            double e = or_write_snapshot(s, x + 1);   /* write(x+1), Main.java:255; the log line at :256
                                                          recomputes the same calcEnergy() and is not repeated */
            if (out->snap_len < out->snap_cap) {
                int64_t k = out->snap_len++;
                if (out->snap_tooth) out->snap_tooth[k] = x + 1;
                if (out->snap_energy) out->snap_energy[k] = e;
                if (out->snap_coords) or_get_coords(s, out->snap_coords + (size_t)k * (size_t)nsites * 3, NULL);
            }
        }
    }
    return startingEnergy;
}
