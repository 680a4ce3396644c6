"""CPU tests of the host-side mirror (parsers, flattening, sharding) and of the C-ABI library surface."""
import ctypes
import os
import re

import numpy as np
import pytest

from transrot_b200 import (Config, EXPORTED_SYMBOLS, LIB_PATH, SystemBuilder, TransRotError, chain_range, global_minimum,
                           owner_of, read_dbase, read_input, read_params, setup_pair_vals, system_from_config)
from transrot_b200.engine import MINIMUM_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_read_dbase_shipped(dbase):
    assert [p.name for p in dbase] == ["NH4+", "Cl-", "H2O"]
    assert [len(p.atoms) for p in dbase] == [5, 1, 3] and [p.radius for p in dbase] == [2.0, 1.0, 2.0]
    assert [a.type_id for p in dbase for a in p.atoms] == list(range(9))
    o = dbase[2].atoms[0]
    assert (o.symbol, o.z, o.c, o.d, o.q, o.mass) == ("O", 0.96, 595.0550, 582002.6617, -0.8340, 15.999)


def test_read_dbase_errors(tmp_path):
    def check(text, msg):
        f = tmp_path / "db.txt"
        f.write_text(text)
        with pytest.raises(TransRotError, match=re.escape(msg)):
            read_dbase(str(f))

    check("x\n", "Error in line 1 of database file: Expected atom count.")
    check("1\nCl-  1.0\n", "Unexpected end of file.")
    check("1\nCl- 1.0\nCl  1.0  0.0  0.0  0.0  0.0  1.0  2.0  -1.0  35.45\n", "Unexpected format for particle definition.")
    check("1\nG  1.0\nX*  1.0  0.0  0.0  0.0  0.0  1.0  2.0  -1.0  35.45\n", "A particle may not be comprised of only ghost atoms.")
    check("1\nA  1.0\nX  1.0  0.0  0.0  0.0  0.0  1.0  2.0  -1.0  35.45\n1\nA  1.0\nX  1.0  0.0  0.0  0.0  0.0  1.0  2.0  -1.0  35.45\n",
          "Particle identifiers must be unique.")
    with pytest.raises(TransRotError, match="not found"):
        read_dbase(str(tmp_path / "missing.txt"))


def test_parse_config_shipped(sample_cfg):
    c = sample_cfg
    assert (c.maxTemp, c.movePerPt, c.ptsPerTooth, c.ptsAddedPerTooth, c.numTeeth) == (10000.0, 20000, 10, 0, 4)
    assert (c.tempDecreasePerTooth, c.maxTrans, c.magwalkFactorTrans, c.magwalkProbTrans) == (0.8, 0.25, 10.0, 0.1)
    assert (c.maxRot, c.magwalkProbRot, c.spaceLen, c.maxPropFailures) == (0.523599, 0.1, 5.0, 20)
    assert (c.useInput, c.finale, c.staticTemp, c.writeConfHeatCapacitiesEnabled, c.eqConfigs) == (False, False, False, True, 1000)
    assert c.particles == [("NH4+", 4), ("Cl-", 4)] and c.num_points() == 40 and c.total_moves() == 800000


def test_parse_config_errors(tmp_path, golden_dir):
    good = open(os.path.join(golden_dir, "sample8.config.txt")).read()

    def check(text, msg):
        f = tmp_path / "c.txt"
        f.write_text(text)
        with pytest.raises(TransRotError, match=re.escape(msg)):
            Config.parseConfig(str(f))

    check(good.replace("Number of Teeth:                                      4", "Number of Teeth:                                      4.5"),
          "Unexpected type for 'Number of Teeth' configuration")
    check(good.replace("0K Finale (true/false):                               false\n", ""), "Missing setting '0K Finale'.")
    check(good + "\nMax Temperature (K):  5.0\n", "Particle counts must be defined after all configuration settings")
    check(good.replace("Cl-  4", "NH4+  4"), "Particle counts can only be specified once per particle.")
    check(good.replace("Max Rotation (Radians)", "Max Spin (Radians)"), "Configuration setting 'Max Spin' not expected.")


def test_space_add_sorts_by_radius(dbase, sample_cfg):
    """Space.add (Space.java:46-61): strictly larger radius goes first, ties keep insertion order."""
    s = system_from_config(dbase, sample_cfg.copy(particles=[("Cl-", 2), ("H2O", 2), ("NH4+", 1)]))
    assert [dbase[i].name for i in s.mol_species] == ["H2O", "H2O", "NH4+", "Cl-", "Cl-"]
    assert list(s.mol_site_off) == [0, 3, 6, 11, 12, 13] and s.n_sites == 13 and s.n_types == 9
    assert list(s.site_type) == [6, 7, 8, 6, 7, 8, 0, 1, 2, 3, 4, 5, 5]
    assert s.pair_evals_per_move() == pytest.approx(np.mean([2 * n * (13 - n) for n in (3, 3, 5, 1, 1)]))
    assert system_from_config(dbase, sample_cfg.copy(particles=[("H2O", 20)])).pair_evals_per_move() == 342.0


def test_read_input_and_frozen(dbase, golden_dir, tmp_path):
    s, xyz = read_input(os.path.join(golden_dir, "sample8.input.xyz"), dbase)
    assert xyz.shape == (24, 3) and s.n_mol == 8 and list(s.moveable_idx) == list(range(8))
    text = open(os.path.join(golden_dir, "sample8.input.xyz")).read().replace("4 NH4+ | 4 Cl-", "Energy: -460.82 Kcal/mole  3 NH4+ | 1F NH4+ | 4 Cl-")
    f = tmp_path / "in.xyz"
    f.write_text(text)
    s2, xyz2 = read_input(str(f), dbase)
    assert list(s2.moveable_idx) == [0, 1, 2, 4, 5, 6, 7] and np.array_equal(xyz, xyz2)
    f.write_text(text.replace("          24", "          23"))
    with pytest.raises(TransRotError, match="Atom count at top of input file"):
        read_input(str(f), dbase)


def test_read_params_matches_combination_rules(dbase, golden_dir):
    """The shipped explicit table holds exactly what setupPairVals computes (reference src/interaction_params.txt)."""
    tab = read_params(os.path.join(golden_dir, "shipped.interaction_params.txt"), dbase)
    assert np.array_equal(tab, setup_pair_vals(dbase))


def test_chain_sharding_partition():
    for world, n in [(1, 7), (2, 4096), (8, 32768), (3, 10), (4, 3)]:
        ranges = [chain_range(r, world, n) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
        assert max(hi - lo for lo, hi in ranges) - min(hi - lo for lo, hi in ranges) <= 1
        for c in (0, n // 2, n - 1):
            lo, hi = ranges[owner_of(c, world, n)]
            assert lo <= c < hi
    m = np.zeros(4, dtype=MINIMUM_DTYPE)
    m["energy"] = [-3.0, -7.5, -7.5, -1.0]
    m["chain"] = [10, 11, 12, 13]
    m["tooth"] = [1, 4, 2, 0]
    assert global_minimum(m) == (-7.5, 11, 4)


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads on a GPU-less host and exports everything include/transrot_b200.h declares."""
    assert os.path.exists(LIB_PATH), "libtransrot_b200.so is not built (python -c 'import __graft_entry__ as g; g.build()')"
    header = open(os.path.join(ROOT, "include", "transrot_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|void|int32_t|const char \*|tr_ctx \*)\s*\*?\s*(tr_[a-z0-9_]+)\s*\(", header, flags=re.M))
    assert declared == set(EXPORTED_SYMBOLS), declared ^ set(EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name


def test_binding_layouts_match_the_library():
    """tr_abi_sizes: the struct layouts of the ctypes / numpy binding are checked against the library's own sizeof /
    offsetof at load time (engine._check_abi) - no GPU needed."""
    from transrot_b200 import load_library

    L = load_library()
    n = L.tr_abi_sizes(None, 0)
    assert n == 18
    buf = (ctypes.c_int32 * n)()
    assert L.tr_abi_sizes(buf, n) == n
    assert list(buf)[0] == 88 and list(buf)[3] == 88 and list(buf)[6] == 56 and list(buf)[9] == 64 and list(buf)[15] == 24


def test_multi_chain_range_matches_the_python_sharding():
    """tr_multi_chain_range (C) and sharding.chain_range (Python, used under torchrun) are the same partition."""
    from transrot_b200 import load_library

    L = load_library()
    for n, world in [(4096, 8), (10, 3), (7, 7), (32768, 8), (5, 2)]:
        for r in range(world):
            lo, hi = ctypes.c_int64(), ctypes.c_int64()
            L.tr_multi_chain_range(r, world, n, ctypes.byref(lo), ctypes.byref(hi))
            assert (lo.value, hi.value) == chain_range(r, world, n)


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product path refuses to run (tr_create fails with a message); it never computes on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from transrot_b200 import Engine

    with pytest.raises(TransRotError, match="no CPU fallback|CUDA"):
        Engine(0)


def test_product_package_does_not_import_the_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "transrot_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "oracle" not in text.replace("the oracle", "").replace("CPU oracle", "").lower() or f == "host.py", f


def test_bench_workload_sizes_are_whole_sweeps():
    """bench.py's chain counts: config #3 is 64 parameter sets x SWEEP_SEEDS seeds per GPU, every set equally often on every GPU
    (the set index depends on the GLOBAL chain id, so the sweep is the same at 1, 2, 4 and 8 GPUs)."""
    import importlib.util
    import os

    import numpy as np

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.CHAINS_PER_GPU["nh4cl32"] == 64 * bench.SWEEP_SEEDS
    assert bench.CHAINS_PER_GPU["h2o20"] == 4096                      # BASELINE.json configs[1]
    for rank in range(8):
        first = rank * bench.CHAINS_PER_GPU["nh4cl32"]
        idx = bench.sched_index(None, np.arange(first, first + bench.CHAINS_PER_GPU["nh4cl32"]), 64)
        assert np.array_equal(np.bincount(idx, minlength=64), np.full(64, bench.SWEEP_SEEDS))


def test_java_struct_layouts_match_the_header():
    """java/main/TransRotB200.java cannot be compiled here (no JDK), so its Panama StructLayouts are checked as text: same field
    names, order and types as the structs of include/transrot_b200.h, every member naturally aligned without implicit padding
    (MemoryLayout.structLayout adds none and rejects misaligned members), same total size as the C compiler's layout."""
    import os
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "transrot_b200.h")).read()
    java = open(os.path.join(root, "java", "main", "TransRotB200.java")).read()
    c_size = {"int32_t": 4, "int8_t": 1, "double": 8, "int64_t": 8, "uint32_t": 4, "ptr": 8}
    j_size = {"JAVA_INT": 4, "JAVA_BYTE": 1, "JAVA_DOUBLE": 8, "JAVA_LONG": 8, "ADDRESS": 8}
    j_of_c = {"int32_t": "JAVA_INT", "int8_t": "JAVA_BYTE", "double": "JAVA_DOUBLE", "int64_t": "JAVA_LONG", "ptr": "ADDRESS"}
    for cname, jname in (("tr_system", "TR_SYSTEM"), ("tr_schedule", "TR_SCHEDULE"), ("tr_point_rec", "TR_POINT_REC"),
                         ("tr_trace_rec", "TR_TRACE_REC"), ("tr_chain_summary", "TR_CHAIN_SUMMARY"), ("tr_minimum", "TR_MINIMUM")):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), header, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        c_fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            m = re.match(r"(?:const\s+)?(\w+)\s*(\*?)\s*(\w+)$", decl)
            assert m, (cname, decl)
            c_fields.append(("ptr" if m.group(2) else m.group(1), m.group(3)))
        jbody = re.search(r"StructLayout %s = MemoryLayout\.structLayout\((.*?)\);" % jname, java, re.S).group(1)
        j_fields = re.findall(r"(JAVA_INT|JAVA_BYTE|JAVA_DOUBLE|JAVA_LONG|ADDRESS)\.withName\(\"(\w+)\"\)", jbody)
        assert "paddingLayout" not in jbody
        assert [n for _, n in j_fields] == [n for _, n in c_fields], cname
        assert [t for t, _ in j_fields] == [j_of_c[t] for t, _ in c_fields], cname
        off, biggest = 0, 1
        for t, n in j_fields:
            assert off % j_size[t] == 0, f"{jname}.{n} would need implicit padding at offset {off}"
            off += j_size[t]
            biggest = max(biggest, j_size[t])
        assert off % biggest == 0, f"{jname}: C pads the tail, the Java layout does not"
        assert off == sum(c_size[t] for t, _ in c_fields)


def test_java_binding_declares_every_function_of_the_header():
    """Every prototype of include/transrot_b200.h has a downcall handle in java/main/TransRotB200.java with the same arity and the
    Panama value layouts its C types map to (checked as text: the image has no JDK)."""
    import os
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "transrot_b200.h")).read(), flags=re.S)
    java = open(os.path.join(root, "java", "main", "TransRotB200.java")).read()
    scalar = {"int": "JAVA_INT", "int32_t": "JAVA_INT", "int64_t": "JAVA_LONG", "long long": "JAVA_LONG", "double": "JAVA_DOUBLE", "float": "JAVA_FLOAT"}

    def layout(decl, is_return=False):
        decl = decl.replace("const", " ").strip()
        if "*" in decl:
            return "ADDRESS"
        toks = decl.split()
        ty = " ".join(toks if is_return or len(toks) == 1 else toks[:-1])
        return "VOID" if ty == "void" else scalar[ty]

    protos = re.findall(r"^\s*((?:const\s+)?\w+(?:\s+\w+)*\s*\*?)\s*(tr_\w+)\s*\(([^;{]*?)\)\s*;", header, re.M | re.S)
    assert len(protos) >= 47
    handles = {}
    for m in re.finditer(r'handle\("(tr_\w+)",\s*FunctionDescriptor\.(of|ofVoid)\(([^;]*?)\)\);', java, re.S):
        handles[m.group(1)] = (m.group(2), [a.strip() for a in m.group(3).replace("\n", " ").split(",") if a.strip()])
    for ret, name, args in protos:
        params = [] if args.strip() in ("", "void") else [layout(a) for a in args.replace("\n", " ").split(",")]
        r = layout(ret, is_return=True)
        assert name in handles, f"{name} has no downcall handle"
        kind, jargs = handles[name]
        assert (kind == "ofVoid") == (r == "VOID"), name
        assert jargs == ([] if r == "VOID" else [r]) + params, (name, jargs, params)
    assert set(handles) == {p[1] for p in protos}
