package main;

import org.apache.commons.math3.random.MersenneTwister;
import org.apache.commons.math3.util.Pair;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemoryLayout;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.StructLayout;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;
import java.lang.reflect.Field;
import java.math.BigDecimal;
import java.math.RoundingMode;
import java.util.ArrayList;
import java.util.HashMap;
import java.util.List;
import java.util.Map;
import java.util.UUID;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_FLOAT;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

/**
 * Panama FFM binding of libtransrot_b200.so (include/transrot_b200.h) and the host half of the seam: replaces the body of
 * Main.sawtoothAnneal (Main.java:177-263) by one device run and rebuilds every output file the loop would have written from
 * the returned buffers, through the reference's own writers (Space.write / writeMovie / writeAcceptance / ... , Space.java:550-859).
 *
 * JDK 22+ (java.lang.foreign is final there).  New file in the reference's package `main`; needs the four package-private
 * accessors java/transrot_b200.patch adds to Space.  NOT COMPILED in this repository's images (no JDK there): the Python twin
 * that exercises the same call sequence against a GPU is transrot_b200/engine.py + transrot_b200/replay.py.
 *
 * Usage (what the patch does to Main.main): {@code TransRotB200.sawtoothAnneal(s, startingEnergy)} instead of
 * {@code sawtoothAnneal(s, startingEnergy)}.
 */
public final class TransRotB200 implements AutoCloseable {

    // ------------------------------------------------------------------------------------------ library and downcall handles

    private static final Linker LINKER = Linker.nativeLinker();
    private static final Arena LIB_ARENA = Arena.global();
    private static final SymbolLookup LIB = SymbolLookup.libraryLookup(System.getProperty("transrot.b200.lib", "libtransrot_b200.so"), LIB_ARENA);

    private static MethodHandle handle(String name, FunctionDescriptor fd) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(() -> new RuntimeException("Error: " + name + " not found in libtransrot_b200.so")), fd);
    }

    private static final MethodHandle TR_CREATE = handle("tr_create", FunctionDescriptor.of(JAVA_INT, JAVA_INT, ADDRESS));
    private static final MethodHandle TR_DESTROY = handle("tr_destroy", FunctionDescriptor.ofVoid(ADDRESS));
    private static final MethodHandle TR_LAST_ERROR = handle("tr_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
    private static final MethodHandle TR_SYNCHRONIZE = handle("tr_synchronize", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    private static final MethodHandle TR_SET_SYSTEM = handle("tr_set_system", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle TR_SET_SCHEDULES = handle("tr_set_schedules", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT));
    private static final MethodHandle TR_SET_OPTIONS = handle("tr_set_options", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, JAVA_INT, JAVA_INT, JAVA_INT));
    private static final MethodHandle TR_SET_RECORD_MOVIE = handle("tr_set_record_movie", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
    private static final MethodHandle TR_INIT_EXPLICIT = handle("tr_init_chains_explicit",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, JAVA_LONG, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS, JAVA_DOUBLE));
    private static final MethodHandle TR_INIT_SEEDED = handle("tr_init_chains_seeded",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, JAVA_LONG, ADDRESS, ADDRESS, JAVA_DOUBLE, JAVA_INT));
    private static final MethodHandle TR_RUN = handle("tr_run", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG));
    private static final MethodHandle TR_LAST_RUN_MS = handle("tr_last_run_ms", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle TR_GET_SUMMARY = handle("tr_get_summary", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle TR_GET_SITES = handle("tr_get_sites", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle TR_MAX_POINTS = handle("tr_max_points", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle TR_GET_POINTS = handle("tr_get_points", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG));
    private static final MethodHandle TR_GET_SNAPSHOTS = handle("tr_get_snapshots", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG));
    private static final MethodHandle TR_GET_MOVIE = handle("tr_get_movie", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG));
    private static final MethodHandle TR_GET_TRACE = handle("tr_get_trace", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle TR_GET_RNG_STATE = handle("tr_get_rng_state", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle TR_PACK_MINIMA = handle("tr_pack_minima", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle TR_GET_MIN_STRUCTURE = handle("tr_get_min_structure", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle TR_TOTAL_ENERGY = handle("tr_total_energy", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS));
    private static final MethodHandle TR_ABI_SIZES = handle("tr_abi_sizes", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
    private static final MethodHandle TR_KERNEL_NAME = handle("tr_kernel_name", FunctionDescriptor.of(ADDRESS, ADDRESS));
    // multi-GPU (one JVM thread drives every GPU of the box; the final gather runs over NCCL inside the library)
    private static final MethodHandle TR_CREATE_MULTI = handle("tr_create_multi", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
    private static final MethodHandle TR_DESTROY_MULTI = handle("tr_destroy_multi", FunctionDescriptor.ofVoid(ADDRESS));
    private static final MethodHandle TR_MULTI_LAST_ERROR = handle("tr_multi_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
    private static final MethodHandle TR_MULTI_CTX = handle("tr_multi_ctx", FunctionDescriptor.of(ADDRESS, ADDRESS, JAVA_INT));
    private static final MethodHandle TR_MULTI_SET_SYSTEM = handle("tr_multi_set_system", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle TR_MULTI_SET_SCHEDULES = handle("tr_multi_set_schedules", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT));
    private static final MethodHandle TR_MULTI_SET_OPTIONS = handle("tr_multi_set_options", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, JAVA_INT, JAVA_INT, JAVA_INT));
    private static final MethodHandle TR_MULTI_INIT_SEEDED = handle("tr_multi_init_chains_seeded",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, JAVA_LONG, ADDRESS, ADDRESS, JAVA_DOUBLE, JAVA_INT));
    private static final MethodHandle TR_MULTI_RUN = handle("tr_multi_run", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG));
    private static final MethodHandle TR_MULTI_SYNCHRONIZE = handle("tr_multi_synchronize", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    private static final MethodHandle TR_GATHER_MINIMA = handle("tr_gather_minima", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle TR_GATHER_MIN_STRUCTURES = handle("tr_gather_min_structures", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    // the rest of the header, bound for completeness (tools, experiments and parity probes; sawtoothAnneal does not need them)
    static final MethodHandle TR_SET_STREAM = handle("tr_set_stream", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle TR_SET_KERNEL = handle("tr_set_kernel", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT));
    static final MethodHandle TR_LAUNCH_COUNT = handle("tr_launch_count", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle TR_PACK_MIN_STRUCTURES = handle("tr_pack_min_structures", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, JAVA_LONG, ADDRESS, ADDRESS));
    static final MethodHandle TR_MULTI_SIZE = handle("tr_multi_size", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    static final MethodHandle TR_MULTI_CHAIN_RANGE = handle("tr_multi_chain_range", FunctionDescriptor.ofVoid(JAVA_INT, JAVA_INT, JAVA_LONG, ADDRESS, ADDRESS));
    static final MethodHandle TR_MULTI_LAST_RUN_MS = handle("tr_multi_last_run_ms", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle TR_DELTA_ENERGY = handle("tr_delta_energy", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS));
    static final MethodHandle TR_RNG_WORDS = handle("tr_rng_words", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, JAVA_INT, ADDRESS));
    static final MethodHandle TR_MEASURE_FP64_PEAK = handle("tr_measure_fp64_peak", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));

    // ------------------------------------------------------------------------------------------ struct layouts (include/transrot_b200.h)

    static final StructLayout TR_SYSTEM = MemoryLayout.structLayout(
            JAVA_INT.withName("n_mol"), JAVA_INT.withName("n_sites"), JAVA_INT.withName("n_types"), JAVA_INT.withName("n_moveable"),
            ADDRESS.withName("mol_site_off"), ADDRESS.withName("mol_radius"), ADDRESS.withName("site_type"), ADDRESS.withName("site_q"),
            ADDRESS.withName("site_def"), ADDRESS.withName("site_mass"), ADDRESS.withName("site_ghost"), ADDRESS.withName("moveable_idx"),
            ADDRESS.withName("pair_abcd"));
    static final StructLayout TR_SCHEDULE = MemoryLayout.structLayout(
            JAVA_DOUBLE.withName("max_temp"), JAVA_DOUBLE.withName("temp_decrease_per_tooth"), JAVA_DOUBLE.withName("max_trans"),
            JAVA_DOUBLE.withName("magwalk_factor_trans"), JAVA_DOUBLE.withName("magwalk_prob_trans"), JAVA_DOUBLE.withName("max_rot"),
            JAVA_DOUBLE.withName("magwalk_prob_rot"),
            JAVA_INT.withName("moves_per_point"), JAVA_INT.withName("points_per_tooth"), JAVA_INT.withName("points_added_per_tooth"),
            JAVA_INT.withName("num_teeth"), JAVA_INT.withName("eq_configs"), JAVA_INT.withName("finale"), JAVA_INT.withName("static_temp"),
            JAVA_INT.withName("write_conf_heat_capacities"));
    static final StructLayout TR_POINT_REC = MemoryLayout.structLayout(
            JAVA_INT.withName("tooth"), JAVA_INT.withName("point"), JAVA_INT.withName("accepted"), JAVA_INT.withName("total"),
            JAVA_INT.withName("movie_written"), JAVA_INT.withName("finale"),
            JAVA_DOUBLE.withName("temp"), JAVA_DOUBLE.withName("total_energy"), JAVA_DOUBLE.withName("total_squared_energy"),
            JAVA_DOUBLE.withName("movie_energy"));
    static final StructLayout TR_TRACE_REC = MemoryLayout.structLayout(
            JAVA_INT.withName("mol"), JAVA_BYTE.withName("kind"), JAVA_BYTE.withName("magwalk"), JAVA_BYTE.withName("bounds"), JAVA_BYTE.withName("accepted"),
            JAVA_INT.withName("axis"), JAVA_INT.withName("drew_rand"),
            JAVA_DOUBLE.withName("e_start"), JAVA_DOUBLE.withName("e_end"), JAVA_DOUBLE.withName("dE"), JAVA_DOUBLE.withName("rand"),
            JAVA_DOUBLE.withName("temp"), JAVA_DOUBLE.withName("running"));
    static final StructLayout TR_CHAIN_SUMMARY = MemoryLayout.structLayout(
            JAVA_DOUBLE.withName("start_energy"), JAVA_DOUBLE.withName("running_energy"), JAVA_DOUBLE.withName("min_energy"), JAVA_DOUBLE.withName("space_len"),
            JAVA_LONG.withName("moves"), JAVA_LONG.withName("evaluated"), JAVA_LONG.withName("accepted"), JAVA_LONG.withName("pair_evals"),
            JAVA_INT.withName("min_tooth"), JAVA_INT.withName("snapshots"), JAVA_INT.withName("points"), JAVA_INT.withName("done"),
            JAVA_INT.withName("prop_retries"), JAVA_INT.withName("pad_"));
    static final StructLayout TR_MINIMUM = MemoryLayout.structLayout(
            JAVA_DOUBLE.withName("energy"), JAVA_LONG.withName("chain"), JAVA_INT.withName("tooth"), JAVA_INT.withName("pad_"));

    private static long off(StructLayout l, String field) {
        return l.byteOffset(MemoryLayout.PathElement.groupElement(field));
    }

    /** The layouts above against the library's own sizeof / offsetof (tr_abi_sizes): a mismatch fails at load time, not at run time. */
    static {
        try (Arena a = Arena.ofConfined()) {
            int n = (int) TR_ABI_SIZES.invokeExact(MemorySegment.NULL, 0);
            MemorySegment buf = a.allocate(JAVA_INT, n);
            int got = (int) TR_ABI_SIZES.invokeExact(buf, n);
            long[] want = {
                    TR_SYSTEM.byteSize(), off(TR_SYSTEM, "mol_site_off"), off(TR_SYSTEM, "pair_abcd"),
                    TR_SCHEDULE.byteSize(), off(TR_SCHEDULE, "moves_per_point"), off(TR_SCHEDULE, "write_conf_heat_capacities"),
                    TR_POINT_REC.byteSize(), off(TR_POINT_REC, "temp"), off(TR_POINT_REC, "movie_energy"),
                    TR_TRACE_REC.byteSize(), off(TR_TRACE_REC, "e_start"), off(TR_TRACE_REC, "running"),
                    TR_CHAIN_SUMMARY.byteSize(), off(TR_CHAIN_SUMMARY, "moves"), off(TR_CHAIN_SUMMARY, "min_tooth"),
                    TR_MINIMUM.byteSize(), off(TR_MINIMUM, "chain"), off(TR_MINIMUM, "tooth")};
            if (got != want.length) throw new RuntimeException("Error: tr_abi_sizes reports " + got + " entries, this binding knows " + want.length);
            for (int i = 0; i < want.length; i++)
                if (buf.getAtIndex(JAVA_INT, i) != want[i])
                    throw new RuntimeException("Error: struct layout entry " + i + " differs: library " + buf.getAtIndex(JAVA_INT, i) + ", binding " + want[i]);
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException("Error: cannot query tr_abi_sizes: " + t);
        }
    }

    // ------------------------------------------------------------------------------------------ one context = one GPU

    private final MemorySegment ctx;
    private final Arena arena = Arena.ofConfined();     // everything passed to tr_set_* lives until close()
    private int nSites, nMol;

    public TransRotB200(int device) {
        try {
            MemorySegment out = arena.allocate(ADDRESS);
            int rc = (int) TR_CREATE.invokeExact(device, out);
            if (rc != 0) throw new RuntimeException(string((MemorySegment) TR_LAST_ERROR.invokeExact(MemorySegment.NULL)));
            ctx = out.get(ADDRESS, 0);
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException("Error: tr_create: " + t);
        }
    }

    @Override
    public void close() {
        try {
            TR_DESTROY.invokeExact(ctx);
        } catch (Throwable ignored) {
        }
        arena.close();
    }

    private static String string(MemorySegment cstr) {
        return cstr.equals(MemorySegment.NULL) ? "Error: unknown" : cstr.reinterpret(4096).getString(0);
    }

    /** Non-zero status -> the RuntimeException("Error: ...") the reference would have thrown (Main.java:85-87). */
    private void check(int rc) {
        if (rc == 0) return;
        try {
            throw new RuntimeException(string((MemorySegment) TR_LAST_ERROR.invokeExact(ctx)));
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException("Error: library call failed with status " + rc);
        }
    }

    // ------------------------------------------------------------------------------------------ Space -> tr_system

    /** Flatten `space` (Space.java:27): what Space.move / rotate / calcEnergy walk as an object graph (Space.java:648-734). */
    public void setSystem(Space s) throws Throwable {
        int[] counts = new int[2];
        MemorySegment sys = buildSystem(s, arena, counts);
        nMol = counts[0];
        nSites = counts[1];
        check((int) TR_SET_SYSTEM.invokeExact(ctx, sys));
    }

    /** The tr_system block (and the arrays it points to) in `arena`; counts = {n_mol, n_sites}. */
    static MemorySegment buildSystem(Space s, Arena arena, int[] counts) {
        List<Molecule> space = s.b200Space();
        List<Molecule> moveable = s.b200Moveable();
        List<Molecule> dbase = s.b200Dbase();
        HashMap<Pair<UUID, UUID>, double[]> pairs = s.b200Pairs();
        // dbase atom slot = site type: one id per UUID minted in readDB (Atom.java:35,54), copied into instances (Atom.java:73)
        Map<UUID, Integer> typeOf = new HashMap<>();
        List<UUID> types = new ArrayList<>();
        for (Molecule m : dbase)
            for (Atom a : m.atoms)
                if (!typeOf.containsKey(a.uuid)) { typeOf.put(a.uuid, types.size()); types.add(a.uuid); }
        int nMol = space.size();
        int nSites = 0;
        for (Molecule m : space) nSites += m.atoms.size();
        counts[0] = nMol;
        counts[1] = nSites;
        int nTypes = types.size();
        MemorySegment molOff = arena.allocate(JAVA_INT, nMol + 1), radius = arena.allocate(JAVA_DOUBLE, nMol);
        MemorySegment type = arena.allocate(JAVA_INT, nSites), q = arena.allocate(JAVA_DOUBLE, nSites), def = arena.allocate(JAVA_DOUBLE, 3L * nSites);
        MemorySegment mass = arena.allocate(JAVA_DOUBLE, nSites), ghost = arena.allocate(JAVA_INT, nSites);
        MemorySegment mov = arena.allocate(JAVA_INT, moveable.size()), abcd = arena.allocate(JAVA_DOUBLE, 4L * nTypes * nTypes);
        int site = 0;
        for (int i = 0; i < nMol; i++) {
            Molecule m = space.get(i);
            molOff.setAtIndex(JAVA_INT, i, site);
            radius.setAtIndex(JAVA_DOUBLE, i, m.radius);
            for (Atom a : m.atoms) {
                Integer t = typeOf.get(a.uuid);
                if (t == null) throw new RuntimeException("Error: atom " + a.symbol + " of " + m.name + " is not a dbase atom");
                type.setAtIndex(JAVA_INT, site, t);
                q.setAtIndex(JAVA_DOUBLE, site, a.q);
                def.setAtIndex(JAVA_DOUBLE, 3L * site, a.defX); def.setAtIndex(JAVA_DOUBLE, 3L * site + 1, a.defY); def.setAtIndex(JAVA_DOUBLE, 3L * site + 2, a.defZ);
                mass.setAtIndex(JAVA_DOUBLE, site, a.mass);
                ghost.setAtIndex(JAVA_INT, site, a.symbol.contains("*") ? 1 : 0);          // Molecule.java:143
                site++;
            }
        }
        molOff.setAtIndex(JAVA_INT, nMol, site);
        for (int k = 0; k < moveable.size(); k++) {
            int idx = -1;
            for (int i = 0; i < nMol; i++) if (space.get(i) == moveable.get(k)) idx = i;         // identity, like Space.move's `n != m`
            mov.setAtIndex(JAVA_INT, k, idx);
        }
        for (int i = 0; i < nTypes; i++)
            for (int j = 0; j < nTypes; j++) {
                double[] v = pairs.get(new Pair<>(types.get(i), types.get(j)));                  // Atom.java:92-93
                for (int c = 0; c < 4; c++) abcd.setAtIndex(JAVA_DOUBLE, 4L * (i * nTypes + j) + c, v == null ? Double.NaN : v[c]);
            }
        MemorySegment sys = arena.allocate(TR_SYSTEM);
        sys.set(JAVA_INT, off(TR_SYSTEM, "n_mol"), nMol);
        sys.set(JAVA_INT, off(TR_SYSTEM, "n_sites"), nSites);
        sys.set(JAVA_INT, off(TR_SYSTEM, "n_types"), nTypes);
        sys.set(JAVA_INT, off(TR_SYSTEM, "n_moveable"), moveable.size());
        sys.set(ADDRESS, off(TR_SYSTEM, "mol_site_off"), molOff);
        sys.set(ADDRESS, off(TR_SYSTEM, "mol_radius"), radius);
        sys.set(ADDRESS, off(TR_SYSTEM, "site_type"), type);
        sys.set(ADDRESS, off(TR_SYSTEM, "site_q"), q);
        sys.set(ADDRESS, off(TR_SYSTEM, "site_def"), def);
        sys.set(ADDRESS, off(TR_SYSTEM, "site_mass"), mass);
        sys.set(ADDRESS, off(TR_SYSTEM, "site_ghost"), ghost);
        sys.set(ADDRESS, off(TR_SYSTEM, "moveable_idx"), mov);
        sys.set(ADDRESS, off(TR_SYSTEM, "pair_abcd"), abcd);
        return sys;
    }

    /** The Config statics Main.sawtoothAnneal reads (Config.java:12-33) as one tr_schedule. */
    public void setScheduleFromConfig() throws Throwable {
        check((int) TR_SET_SCHEDULES.invokeExact(ctx, buildSchedule(arena), 1));
    }

    static MemorySegment buildSchedule(Arena arena) {
        MemorySegment sc = arena.allocate(TR_SCHEDULE);
        sc.set(JAVA_DOUBLE, off(TR_SCHEDULE, "max_temp"), Config.maxTemp);
        sc.set(JAVA_DOUBLE, off(TR_SCHEDULE, "temp_decrease_per_tooth"), Config.tempDecreasePerTooth);
        sc.set(JAVA_DOUBLE, off(TR_SCHEDULE, "max_trans"), Config.maxTrans);
        sc.set(JAVA_DOUBLE, off(TR_SCHEDULE, "magwalk_factor_trans"), Config.magwalkFactorTrans);
        sc.set(JAVA_DOUBLE, off(TR_SCHEDULE, "magwalk_prob_trans"), Config.magwalkProbTrans);
        sc.set(JAVA_DOUBLE, off(TR_SCHEDULE, "max_rot"), Config.maxRot);
        sc.set(JAVA_DOUBLE, off(TR_SCHEDULE, "magwalk_prob_rot"), Config.magwalkProbRot);
        sc.set(JAVA_INT, off(TR_SCHEDULE, "moves_per_point"), Config.movePerPt);
        sc.set(JAVA_INT, off(TR_SCHEDULE, "points_per_tooth"), Config.ptsPerTooth);
        sc.set(JAVA_INT, off(TR_SCHEDULE, "points_added_per_tooth"), Config.ptsAddedPerTooth);
        sc.set(JAVA_INT, off(TR_SCHEDULE, "num_teeth"), Config.numTeeth);
        sc.set(JAVA_INT, off(TR_SCHEDULE, "eq_configs"), Config.eqConfigs);
        sc.set(JAVA_INT, off(TR_SCHEDULE, "finale"), Config.finale ? 1 : 0);
        sc.set(JAVA_INT, off(TR_SCHEDULE, "static_temp"), Config.staticTemp ? 1 : 0);
        sc.set(JAVA_INT, off(TR_SCHEDULE, "write_conf_heat_capacities"), Config.writeConfHeatCapacitiesEnabled ? 1 : 0);
        return sc;
    }

    /** Space.calcEnergy() (Space.java:648-658) of the structure the Space holds now, on the device (call setSystem first). */
    public double calcEnergy(Space s) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment sites = a.allocate(JAVA_DOUBLE, 3L * nSites), e = a.allocate(JAVA_DOUBLE);
            long k = 0;
            for (Molecule m : s.b200Space())
                for (Atom at : m.atoms) {
                    sites.setAtIndex(JAVA_DOUBLE, k, at.x); sites.setAtIndex(JAVA_DOUBLE, k + 1, at.y); sites.setAtIndex(JAVA_DOUBLE, k + 2, at.z);
                    k += 3;
                }
            check((int) TR_TOTAL_ENERGY.invokeExact(ctx, sites, 1L, e));
            return e.get(JAVA_DOUBLE, 0);
        }
    }

    /** Which kernel the library chose for the live chain batch, and the device time of the last tr_run in milliseconds. */
    public String kernelName() throws Throwable {
        return string((MemorySegment) TR_KERNEL_NAME.invokeExact(ctx));
    }

    public float lastRunMs() throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment ms = a.allocate(JAVA_FLOAT);
            check((int) TR_LAST_RUN_MS.invokeExact(ctx, ms));
            return ms.get(JAVA_FLOAT, 0);
        }
    }

    /** The minimum over the write(n) snapshots of chain `chain` (Space.java:614-617) with its structure: {energy, tooth}. */
    public double[] minimum(long chain, double[] sitesOut) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment x = a.allocate(JAVA_DOUBLE, 3L * nSites), e = a.allocate(JAVA_DOUBLE), t = a.allocate(JAVA_INT);
            check((int) TR_GET_MIN_STRUCTURE.invokeExact(ctx, chain, x, e, t));
            for (int i = 0; i < 3 * nSites && i < sitesOut.length; i++) sitesOut[i] = x.getAtIndex(JAVA_DOUBLE, i);
            return new double[]{e.get(JAVA_DOUBLE, 0), t.get(JAVA_INT, 0)};
        }
    }

    // ------------------------------------------------------------------------------------------ RNG state of the three static generators

    /** MersenneTwister keeps `int[] mt` (624 words) and `int mti` private (commons-math3 3.6.1): read them by reflection. */
    private static void exportState(MersenneTwister g, MemorySegment dst, long wordOffset) throws ReflectiveOperationException {
        Field fmt = MersenneTwister.class.getDeclaredField("mt");
        Field fmti = MersenneTwister.class.getDeclaredField("mti");
        fmt.setAccessible(true);
        fmti.setAccessible(true);
        int[] mt = (int[]) fmt.get(g);
        for (int i = 0; i < 624; i++) dst.setAtIndex(JAVA_INT, wordOffset + i, mt[i]);
        dst.setAtIndex(JAVA_INT, wordOffset + 624, fmti.getInt(g));
    }

    private static void importState(MersenneTwister g, MemorySegment src, long wordOffset) throws ReflectiveOperationException {
        Field fmt = MersenneTwister.class.getDeclaredField("mt");
        Field fmti = MersenneTwister.class.getDeclaredField("mti");
        fmt.setAccessible(true);
        fmti.setAccessible(true);
        int[] mt = (int[]) fmt.get(g);
        for (int i = 0; i < 624; i++) mt[i] = src.getAtIndex(JAVA_INT, wordOffset + i);
        fmti.setInt(g, src.getAtIndex(JAVA_INT, wordOffset + 624));
    }

    private static MersenneTwister staticGenerator(Class<?> owner) throws ReflectiveOperationException {
        Field f = owner.getDeclaredField("r");          // Main.r (Main.java:15), Space.r (Space.java:31), Molecule.r (Molecule.java:13)
        f.setAccessible(true);
        return (MersenneTwister) f.get(null);
    }

    // ------------------------------------------------------------------------------------------ the replaced call

    private static void putCoordinates(Space s, MemorySegment sites, long base) {
        long k = base;
        for (Molecule m : s.b200Space())
            for (Atom a : m.atoms) {
                a.x = sites.getAtIndex(JAVA_DOUBLE, k); a.y = sites.getAtIndex(JAVA_DOUBLE, k + 1); a.z = sites.getAtIndex(JAVA_DOUBLE, k + 2);
                a.tempx = a.x; a.tempy = a.y; a.tempz = a.z;
                k += 3;
            }
    }

    /**
     * Drop-in for Main.sawtoothAnneal(Space, double) (Main.java:81,177-263): one chain, started from the Space as Main.main left
     * it (after propagate / readInput, write(0) and calcEnergy()) with the three generators in the state they are in.
     */
    public static void sawtoothAnneal(Space s, double startingEnergy) {
        try (TransRotB200 b = new TransRotB200(Integer.getInteger("transrot.b200.device", 0)); Arena a = Arena.ofConfined()) {
            b.setSystem(s);
            b.setScheduleFromConfig();
            boolean wantTrace = Config.staticTemp && Config.writeEnergiesEnabled;
            long traceMoves = wantTrace ? Config.movePerPt : 0;
            b.check((int) TR_SET_OPTIONS.invokeExact(b.ctx, traceMoves, 1, 1, 0));
            b.check((int) TR_SET_RECORD_MOVIE.invokeExact(b.ctx, 1));
            // start structure and generator states
            int S = b.nSites;
            MemorySegment sites = a.allocate(JAVA_DOUBLE, 3L * S);
            long k = 0;
            for (Molecule m : s.b200Space())
                for (Atom at : m.atoms) {
                    sites.setAtIndex(JAVA_DOUBLE, k, at.x); sites.setAtIndex(JAVA_DOUBLE, k + 1, at.y); sites.setAtIndex(JAVA_DOUBLE, k + 2, at.z);
                    k += 3;
                }
            MemorySegment rng = a.allocate(JAVA_INT, 3L * 625);
            exportState(staticGenerator(Main.class), rng, 0);
            exportState(staticGenerator(Space.class), rng, 625);
            exportState(staticGenerator(Molecule.class), rng, 1250);
            b.check((int) TR_INIT_EXPLICIT.invokeExact(b.ctx, 1L, 0L, sites, 1L, MemorySegment.NULL, rng, MemorySegment.NULL, Config.spaceLen));
            b.check((int) TR_RUN.invokeExact(b.ctx, 0L));
            b.check((int) TR_SYNCHRONIZE.invokeExact(b.ctx));
            b.replayOutputs(s, a, wantTrace);
            // leave the Space and the generators as the reference loop would have
            b.check((int) TR_GET_SITES.invokeExact(b.ctx, sites, MemorySegment.NULL));
            putCoordinates(s, sites, 0);
            b.check((int) TR_GET_RNG_STATE.invokeExact(b.ctx, rng));
            importState(staticGenerator(Main.class), rng, 0);
            importState(staticGenerator(Space.class), rng, 625);
            importState(staticGenerator(Molecule.class), rng, 1250);
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException("Error: " + t);
        }
    }

    /**
     * The file writes of the y-loop (Main.java:191-258) from the returned records, through the reference's own writers: per point
     * writeAcceptance (:225), the heat-capacity line (:226-233), writeMovie (:240); per tooth write(x+1) and the log line
     * (:254-257); energies.txt per move in Static Temperature mode (:220-222); writeMinEnergy (:260-262).  The writers call
     * calcEnergy() themselves (Space.java:554,745): the structure is put back into the Space first, so every file is
     * self-consistent and Space.minEnergy / minEnergyToothNum (Space.java:614-617) are tracked by Space.write as always.
     */
    private void replayOutputs(Space s, Arena a, boolean wantTrace) throws Throwable {
        MemorySegment caps = a.allocate(JAVA_LONG, 2);
        check((int) TR_MAX_POINTS.invokeExact(ctx, caps.asSlice(0, 8), caps.asSlice(8, 8)));
        long maxPoints = caps.getAtIndex(JAVA_LONG, 0), maxSnaps = caps.getAtIndex(JAVA_LONG, 1);
        long S3 = 3L * nSites;
        MemorySegment summary = a.allocate(TR_CHAIN_SUMMARY);
        check((int) TR_GET_SUMMARY.invokeExact(ctx, summary));
        int nPoints = summary.get(JAVA_INT, off(TR_CHAIN_SUMMARY, "points"));
        int nSnaps = summary.get(JAVA_INT, off(TR_CHAIN_SUMMARY, "snapshots"));
        MemorySegment points = a.allocate(TR_POINT_REC, maxPoints);
        check((int) TR_GET_POINTS.invokeExact(ctx, points, maxPoints));
        MemorySegment snapE = a.allocate(JAVA_DOUBLE, maxSnaps), snapT = a.allocate(JAVA_INT, maxSnaps), snapX = a.allocate(JAVA_DOUBLE, maxSnaps * S3);
        check((int) TR_GET_SNAPSHOTS.invokeExact(ctx, snapE, snapT, snapX, maxSnaps));
        MemorySegment movie = a.allocate(JAVA_DOUBLE, maxPoints * S3);
        check((int) TR_GET_MOVIE.invokeExact(ctx, movie, maxPoints));
        if (wantTrace) {
            long n = Config.movePerPt;
            MemorySegment trace = a.allocate(TR_TRACE_REC, n);
            check((int) TR_GET_TRACE.invokeExact(ctx, trace));
            long moves = summary.get(JAVA_LONG, off(TR_CHAIN_SUMMARY, "moves"));
            for (long i = 0; i < Math.min(n, moves); i++)
                s.writeEnergy(trace.get(JAVA_DOUBLE, i * TR_TRACE_REC.byteSize() + off(TR_TRACE_REC, "running")));      // Main.java:220-222
        }
        int snap = 1;                                   // snapshot 0 is write(0): Main.main wrote it already (Main.java:64)
        for (int p = 0; p < nPoints; p++) {
            long base = p * TR_POINT_REC.byteSize();
            int tooth = points.get(JAVA_INT, base + off(TR_POINT_REC, "tooth"));
            double t = points.get(JAVA_DOUBLE, base + off(TR_POINT_REC, "temp"));
            s.writeAcceptance(t, points.get(JAVA_INT, base + off(TR_POINT_REC, "accepted")), points.get(JAVA_INT, base + off(TR_POINT_REC, "total")));
            if (Config.writeConfHeatCapacitiesEnabled) {
                double roundedTemperature = new BigDecimal(t).setScale(2, RoundingMode.HALF_UP).doubleValue();
                if (roundedTemperature != 0) {
                    double avgEnergy = points.get(JAVA_DOUBLE, base + off(TR_POINT_REC, "total_energy")) / Config.movePerPt;
                    double avgSquaredEnergy = points.get(JAVA_DOUBLE, base + off(TR_POINT_REC, "total_squared_energy")) / Config.movePerPt;
                    double cv = (avgSquaredEnergy - Math.pow(avgEnergy, 2)) / (Space.BOLTZMANN_CONSTANT * Math.pow(t, 2));
                    s.writeConfigurationalHeatCapacity(roundedTemperature, avgEnergy, cv);
                }
            }
            if (points.get(JAVA_INT, base + off(TR_POINT_REC, "movie_written")) != 0) {
                putCoordinates(s, movie, p * S3);
                s.writeMovie(tooth, tooth + 1);
            }
            // end of a tooth that writes an Output file: the next write(n) snapshot belongs here (the pre-finale tooth writes none)
            boolean lastOfTooth = p + 1 == nPoints || points.get(JAVA_INT, (p + 1) * TR_POINT_REC.byteSize() + off(TR_POINT_REC, "tooth")) != tooth
                    || points.get(JAVA_INT, (p + 1) * TR_POINT_REC.byteSize() + off(TR_POINT_REC, "finale")) != points.get(JAVA_INT, base + off(TR_POINT_REC, "finale"));
            if (lastOfTooth && snap < nSnaps && snapT.getAtIndex(JAVA_INT, snap) == tooth + 1
                    && !(Config.finale && points.get(JAVA_INT, base + off(TR_POINT_REC, "finale")) == 0 && tooth == Config.numTeeth - 1)) {
                putCoordinates(s, snapX, snap * S3);
                s.write(tooth + 1);
                s.log("Energy at end of tooth " + (tooth + 1) + ": " + s.calcEnergy());
                snap++;
            }
        }
        if (Config.numTeeth > 1) s.writeMinEnergy();
    }

    // ------------------------------------------------------------------------------------------ batched, multi-GPU use (thousands of seeds)

    /**
     * Every GPU of the box from this one thread: seeds.length seeded chains of the system in `s` (Space.propagate runs on the
     * devices), sharded in contiguous global ranges and annealed concurrently; the only exchange is the NCCL gather of per-chain
     * minima inside the library (tr_gather_minima).  bestSites = double[3 * S] receives the winning structure; the return value
     * is {energy, global chain index, tooth}.
     */
    public static double[] annealBatchAllGpus(Space s, int[] devices, long[] seeds, double[] bestSites) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment dev = a.allocate(JAVA_INT, devices.length);
            for (int i = 0; i < devices.length; i++) dev.setAtIndex(JAVA_INT, i, devices[i]);
            MemorySegment out = a.allocate(ADDRESS);
            int rc = (int) TR_CREATE_MULTI.invokeExact(dev, devices.length, out);
            if (rc != 0) throw new RuntimeException(string((MemorySegment) TR_MULTI_LAST_ERROR.invokeExact(MemorySegment.NULL)));
            MemorySegment m = out.get(ADDRESS, 0);
            try {
                int[] counts = new int[2];
                multiCheck(m, (int) TR_MULTI_SET_SYSTEM.invokeExact(m, buildSystem(s, a, counts)));
                multiCheck(m, (int) TR_MULTI_SET_SCHEDULES.invokeExact(m, buildSchedule(a), 1));
                multiCheck(m, (int) TR_MULTI_SET_OPTIONS.invokeExact(m, 0L, 1, 1, 0));
                MemorySegment sd = a.allocate(JAVA_LONG, seeds.length);
                for (int i = 0; i < seeds.length; i++) sd.setAtIndex(JAVA_LONG, i, seeds[i]);
                multiCheck(m, (int) TR_MULTI_INIT_SEEDED.invokeExact(m, (long) seeds.length, 0L, sd, MemorySegment.NULL, Config.spaceLen, Config.maxPropFailures));
                multiCheck(m, (int) TR_MULTI_RUN.invokeExact(m, 0L));
                multiCheck(m, (int) TR_MULTI_SYNCHRONIZE.invokeExact(m));
                MemorySegment best = a.allocate(TR_MINIMUM), sites = a.allocate(JAVA_DOUBLE, 3L * counts[1]);
                multiCheck(m, (int) TR_GATHER_MINIMA.invokeExact(m, MemorySegment.NULL, best, sites));
                for (int i = 0; i < 3 * counts[1] && i < bestSites.length; i++) bestSites[i] = sites.getAtIndex(JAVA_DOUBLE, i);
                return new double[]{best.get(JAVA_DOUBLE, off(TR_MINIMUM, "energy")), (double) best.get(JAVA_LONG, off(TR_MINIMUM, "chain")),
                        (double) best.get(JAVA_INT, off(TR_MINIMUM, "tooth"))};
            } finally {
                TR_DESTROY_MULTI.invokeExact(m);
            }
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException("Error: " + t);
        }
    }

    private static void multiCheck(MemorySegment m, int rc) throws Throwable {
        if (rc != 0) throw new RuntimeException(string((MemorySegment) TR_MULTI_LAST_ERROR.invokeExact(m)));
    }
}
