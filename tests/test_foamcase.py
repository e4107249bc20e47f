"""OpenFOAM case I/O of the standalone engine (yade-openfoam-coupling_b200/foamcase.py), on the CPU:
the stock OpenFOAM-6 cavity tutorial's dictionaries and field files as they are on disk (tests/cases_foam/cavity, restated
from tutorials/incompressible/icoFoam/cavity/cavity) + a polyMesh in blockMesh's conventions -> the case description ->
the oracle reproduces the tutorial's public solver log from it; time directories round-trip."""
import importlib.util
import os
import shutil

import numpy as np
import pytest

from oracle import meshgen, port
from tests.test_fv_oracle import CAVITY_LOG, sig6

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BIN_PATH = os.path.join(ROOT, "yade-openfoam-coupling_b200", "foamYadeB200")


def _foamcase():
    spec = importlib.util.spec_from_file_location("fy_foamcase", os.path.join(ROOT, "yade-openfoam-coupling_b200", "foamcase.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


fc = _foamcase()
CAVITY_PATCHES = [("movingWall", ["ymax"]), ("fixedWalls", ["xmin", "xmax", "ymin"]), ("frontAndBack", ["zmin", "zmax"])]


def cavity_case(tmp_path, n=(20, 20, 1)):
    case = str(tmp_path / "cavity")
    shutil.copytree(os.path.join(HERE, "cases_foam", "cavity"), case)
    fc.write_box_poly_mesh(case, n, (0.1, 0.1, 0.01), patches=CAVITY_PATCHES,
                           patch_types={"movingWall": "wall", "fixedWalls": "wall", "frontAndBack": "empty"})
    return case


def test_tutorial_directory_before_blockMesh_has_been_run(tmp_path):
    """the stock cavity directory as OpenFOAM ships it -- system/blockMeshDict, no constant/polyMesh -- reads to the same
    case as after `blockMesh` (here: after write_box_poly_mesh), up to the numbering of the boundary faces inside a patch"""
    with_pm = fc.load_case(cavity_case(tmp_path))
    bare = str(tmp_path / "bare")
    shutil.copytree(os.path.join(HERE, "cases_foam", "cavity"), bare)
    assert not os.path.exists(os.path.join(bare, "constant", "polyMesh"))
    c = fc.load_case(bare)
    assert c["box"]["n"] == with_pm["box"]["n"] == (20, 20, 1)
    np.testing.assert_allclose(c["box"]["L"], with_pm["box"]["L"], rtol=1e-15)
    np.testing.assert_allclose(c["box"]["origin"], with_pm["box"]["origin"], rtol=0, atol=1e-18)
    key = lambda q: (q["name"], q["type"], sorted(q["sides"]), q["start"], q["nFaces"], q["bcU"], q["valueU"], q["bcP"], q["valueP"])   # noqa: E731
    assert [key(q) for q in c["patches"]] == [key(q) for q in with_pm["patches"]]
    assert c["piso"] == with_pm["piso"] and c["nu"] == with_pm["nu"]
    m = fc.build_mesh(c, meshgen.hex_box_ldu, meshgen.set_bc, meshgen)
    ref = meshgen.hex_box_ldu(20, 20, 1, 0.1, 0.1, 0.01, patches=CAVITY_PATCHES)
    assert np.array_equal(m["owner"], ref["owner"]) and np.array_equal(m["V"], ref["V"])
    # time directories can be written without a polyMesh on disk
    tname = fc.write_time(c, 0.005, np.zeros((400, 3)), np.arange(400.0))
    assert np.array_equal(fc.load_case(bare, time=tname)["p"], np.arange(400.0))
    # refusals
    bmd = os.path.join(bare, "system", "blockMeshDict")
    txt = open(bmd).read()
    open(bmd, "w").write(txt.replace("simpleGrading (1 1 1)", "simpleGrading (2 1 1)"))
    with pytest.raises(fc.FoamCaseError, match="graded"):
        fc.load_case(bare)
    open(bmd, "w").write(txt.replace("            (4 5 6 7)\n", ""))
    with pytest.raises(fc.FoamCaseError, match="without a patch"):
        fc.load_case(bare)
    open(bmd, "w").write(txt)
    if os.path.exists(BIN_PATH):
        import json
        import subprocess
        r = subprocess.run([BIN_PATH, "-case", bare, "-dump"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        cc = json.loads(r.stdout)
        assert tuple(cc["n"]) == (20, 20, 1) and [q["sides"] for q in cc["patches"]] == [q["sides"] for q in c["patches"]]
        assert [(q["start"], q["nFaces"]) for q in cc["patches"]] == [(q["start"], q["nFaces"]) for q in c["patches"]]
        np.testing.assert_allclose(cc["L"], c["box"]["L"], rtol=1e-15)


def test_dictionary_syntax():
    d = fc.parse_dict('''
        FoamFile { version 2.0; format ascii; class dictionary; }   // header
        /* block
           comment */
        nu              [0 2 -1 0 0 0 0] 0.01;
        nuOld           nuOld [0 2 -1 0 0 0 0] 1e-06;
        g               (0 0 -9.81);
        solvers { p { solver PCG; tolerance 1e-06; relTol 0.05; } pFinal { $p; relTol 0; } "(U|k)" { solver smoothSolver; } }
        on yes; name "quoted string"; list 3 (1 2 3);
    ''')
    assert fc.scalar_of(d["nu"], "nu") == 0.01 and fc.scalar_of(d["nuOld"], "nuOld") == 1e-6
    assert fc.vector_of(d["g"], "g") == (0.0, 0.0, -9.81)
    assert d["solvers"]["pFinal"] == {"solver": "PCG", "tolerance": 1e-06, "relTol": 0}
    assert d["solvers"]["p"]["relTol"] == 0.05 and "(U|k)" in d["solvers"]
    assert d["on"] == "yes" and d["name"] == "quoted string" and list(d["list"][1]) == [1, 2, 3]
    with pytest.raises(fc.FoamCaseError):
        fc.parse_dict("a { b 1; ")
    with pytest.raises(fc.FoamCaseError):
        fc.parse_dict("a { $missing; }")


def test_cavity_case_is_read_as_the_tutorial_defines_it(tmp_path):
    case = fc.load_case(cavity_case(tmp_path))
    assert case["box"]["n"] == (20, 20, 1)
    np.testing.assert_allclose(case["box"]["L"], (0.1, 0.1, 0.01), rtol=1e-15)
    by = {p["name"]: p for p in case["patches"]}
    assert by["movingWall"]["sides"] == ["ymax"] and by["movingWall"]["bcU"] == "fixedValue" and by["movingWall"]["valueU"] == (1.0, 0.0, 0.0)
    assert sorted(by["fixedWalls"]["sides"]) == ["xmax", "xmin", "ymin"] and by["fixedWalls"]["bcU"] == "fixedValue"        # noSlip
    assert by["fixedWalls"]["valueU"] == (0.0, 0.0, 0.0) and by["fixedWalls"]["bcP"] == "zeroGradient"
    assert by["frontAndBack"]["bcU"] == by["frontAndBack"]["bcP"] == "empty"
    assert case["nu"] == 0.01 and case["control"]["deltaT"] == 0.005 and case["control"]["endTime"] == 0.5
    assert case["control"]["writeInterval"] == 20
    assert case["piso"] == dict(nCorrectors=2, nNonOrthogonalCorrectors=0, momentumPredictor=1, pRefCell=0, pRefValue=0.0, pTol=1e-6,
                                pRelTol=0.05, pFinalTol=1e-6, pFinalRelTol=0.0, UTol=1e-5, URelTol=0.0, maxIter=1000,
                                preconditioner="DIC")
    assert not np.any(case["U"]) and not np.any(case["p"]) and case["U"].shape == (400, 3)
    # the mesh the description builds IS the synthetic generator's (same arrays, bit for bit)
    m = fc.build_mesh(case, meshgen.hex_box_ldu, meshgen.set_bc, meshgen)
    ref = meshgen.hex_box_ldu(20, 20, 1, 0.1, 0.1, 0.01, patches=CAVITY_PATCHES)
    for k in ("owner", "neighbour", "Sf", "magSf", "weights", "deltaCoeffs", "V"):
        assert np.array_equal(m[k], ref[k]), k
    np.testing.assert_allclose(m["C"], ref["C"], rtol=0, atol=1e-17)


@pytest.mark.skipif(not port.available(), reason="oracle/_build/liboracle.so not built")
def test_cavity_case_from_disk_reproduces_the_tutorial_log(tmp_path):
    """the public log.icoFoam digits (tests/test_fv_oracle.py) from the case DIRECTORY: reader -> mesh + controls -> oracle"""
    case = fc.load_case(cavity_case(tmp_path))
    m = fc.build_mesh(case, meshgen.hex_box_ldu, meshgen.set_bc, meshgen)
    ctl = dict(case["piso"])
    O = port.IcoOracle(m, nu=case["nu"], **ctl)
    O.field("U")[:] = case["U"]
    O.field("p")[:] = case["p"]
    O.create_phi()
    dt = case["control"]["deltaT"]
    for ref in CAVITY_LOG:
        O.pre(dt)
        O.solve(dt)
        st = O.stats()
        assert (sig6(st["meanCoNum"]), sig6(st["CoNum"])) == ref["Co"]
        for j, k in ((0, "Ux"), (1, "Uy")):
            assert (sig6(st["U"][j]["initial"]), sig6(st["U"][j]["final"]), st["U"][j]["iters"]) == ref[k], k
        for j, k in ((0, "p1"), (1, "p2")):
            assert (sig6(st["p"][j]["initial"]), sig6(st["p"][j]["final"]), st["p"][j]["iters"]) == ref[k], k
    # runTime.write(): the time directory reads back as written; at writePrecision 17 bit for bit
    U, p = O.field("U").copy(), O.field("p").copy()
    O.close()
    case["control"]["writePrecision"] = 17
    tname = fc.write_time(case, 3 * dt, U, p)
    assert tname == "0.015"
    back = fc.load_case(case["case_dir"], time=tname)
    assert np.array_equal(back["U"], U) and np.array_equal(back["p"], p)
    assert [(q["name"], q["bcU"], q["valueU"], q["bcP"]) for q in back["patches"]] == \
           [(q["name"], q["bcU"], q["valueU"], q["bcP"]) for q in case["patches"]]
    case["control"]["writePrecision"] = 6
    fc.write_time(case, 4 * dt, U, p)
    back6 = fc.load_case(case["case_dir"], time="0.02")
    assert np.abs(back6["U"] - U).max() <= 5e-6 * np.abs(U).max()


def test_3d_channel_case_with_nonuniform_fields_and_pimple_dictionaries(tmp_path):
    """a 3-D box with inlet / outlet / walls, nonuniform internal fields, the pimpleFoamYade property names, a PIMPLE
    dictionary with outer correctors and relaxationFactors, constant/g"""
    case_dir = str(tmp_path / "chan")
    n, L, org = (6, 5, 4), (2.0, 1.0, 0.5), (-1.0, 0.25, 3.0)
    patches = [("inlet", ["xmin"]), ("outlet", ["xmax"]), ("walls", ["ymin", "ymax", "zmin", "zmax"])]
    fc.write_box_poly_mesh(case_dir, n, L, org, patches, {"inlet": "patch", "outlet": "patch", "walls": "wall"})
    N = 120
    rng = np.random.default_rng(3)
    U, p = rng.standard_normal((N, 3)), rng.standard_normal(N)
    pl = [dict(name=a, sides=b) for a, b in patches]
    fc.write_field(case_dir, "0", "Uc", U, (0, 1, -1, 0, 0, 0, 0), pl,
                   {"inlet": ("fixedValue", np.array([0.3, 0, 0])), "outlet": ("zeroGradient", None), "walls": ("noSlip", None)}, prec=17)
    fc.write_field(case_dir, "0", "p", p, (0, 2, -2, 0, 0, 0, 0), pl,
                   {"inlet": ("zeroGradient", None), "outlet": ("fixedValue", 0.25), "walls": ("fixedFluxPressure", 0.0)}, prec=17)
    os.makedirs(os.path.join(case_dir, "system"))
    open(os.path.join(case_dir, "constant", "transportProperties"), "w").write(
        "FoamFile{version 2.0;format ascii;class dictionary;object transportProperties;}\n"
        "partDensity partDensity [1 -3 0 0 0 0 0] 2500;\nrhocValue rhocValue [1 -3 0 0 0 0 0] 1000;\nnuValue nuValue [0 2 -1 0 0 0 0] 1e-06;\n")
    open(os.path.join(case_dir, "constant", "g"), "w").write(
        "FoamFile{version 2.0;format ascii;class uniformDimensionedVectorField;object g;}\ndimensions [0 1 -2 0 0 0 0];\nvalue (0 0 -9.81);\n")
    open(os.path.join(case_dir, "system", "controlDict"), "w").write(
        "FoamFile{version 2.0;format ascii;class dictionary;object controlDict;}\napplication pimpleFoamYade;\nstartTime 0;\nendTime 1;\n"
        "deltaT 1e-3;\nwriteControl timeStep;\nwriteInterval 100;\n")
    open(os.path.join(case_dir, "system", "fvSolution"), "w").write(
        "FoamFile{version 2.0;format ascii;class dictionary;object fvSolution;}\n"
        "solvers { p { solver PCG; preconditioner diagonal; tolerance 1e-7; relTol 0.01; } pFinal { $p; relTol 0; }\n"
        '          "(Uc|k)" { solver smoothSolver; smoother symGaussSeidel; tolerance 1e-6; relTol 0.1; } }\n'
        "PIMPLE { nOuterCorrectors 3; nCorrectors 2; momentumPredictor no; nNonOrthogonalCorrectors 1; pRefCell 7; pRefValue 1.5; }\n"
        'relaxationFactors { equations { "Uc.*" 0.7; UcFinal 1; } fields { p 0.3; pFinal 1; } }\n')
    case = fc.load_case(case_dir, solver="pimpleFoamYade")
    assert case["Uname"] == "Uc" and np.array_equal(case["U"], U) and np.array_equal(case["p"], p)
    assert case["box"]["n"] == n and np.allclose(case["box"]["origin"], org) and np.allclose(case["box"]["L"], L)
    by = {q["name"]: q for q in case["patches"]}
    assert by["inlet"]["bcU"] == "fixedValue" and by["inlet"]["valueU"] == (0.3, 0.0, 0.0) and by["outlet"]["bcU"] == "zeroGradient"
    assert by["outlet"]["bcP"] == "fixedValue" and by["outlet"]["valueP"] == 0.25 and by["walls"]["bcP"] == "fixedFluxPressure"
    assert case["nu"] == 1e-6 and case["props"] == dict(nuValue=1e-6, rhocValue=1000.0, partDensity=2500.0) and case["g"] == (0.0, 0.0, -9.81)
    assert case["piso"]["preconditioner"] == "diagonal" and case["piso"]["momentumPredictor"] == 0 and case["piso"]["pRefCell"] == 7
    assert case["piso"]["nNonOrthogonalCorrectors"] == 1 and case["piso"]["URelTol"] == 0.1 and case["piso"]["pFinalRelTol"] == 0.0
    assert case["pimple"] == dict(nOuterCorrectors=3, relaxU=0.7, relaxUFinal=1.0, relaxP=0.3, relaxPFinal=1.0)
    m = fc.build_mesh(case, meshgen.hex_box_ldu, meshgen.set_bc, meshgen)
    ref = meshgen.hex_box_ldu(*n, *L, origin=org, patches=patches)
    assert np.array_equal(m["owner"], ref["owner"]) and np.allclose(m["C"], ref["C"], rtol=0, atol=1e-15)


def test_unsupported_cases_are_refused_with_the_reason(tmp_path):
    case_dir = cavity_case(tmp_path)
    # a patch covering part of a side
    pm = os.path.join(case_dir, "constant", "polyMesh", "boundary")
    txt = open(pm).read()
    open(pm, "w").write(txt.replace("nFaces          20;", "nFaces          10;", 1))
    with pytest.raises(fc.FoamCaseError, match="whole sides|no patch|lie on no box side"):
        fc.load_case(case_dir)
    open(pm, "w").write(txt)
    # binary files
    pts = os.path.join(case_dir, "constant", "polyMesh", "points")
    t2 = open(pts).read()
    open(pts, "w").write(t2.replace("format      ascii;", "format      binary;", 1))
    with pytest.raises(fc.FoamCaseError, match="ascii"):
        fc.load_case(case_dir)
    open(pts, "w").write(t2)
    # a boundary condition the device path does not implement
    u = os.path.join(case_dir, "0", "U")
    t3 = open(u).read()
    open(u, "w").write(t3.replace("type            noSlip;", "type            slip;"))
    with pytest.raises(fc.FoamCaseError, match="slip"):
        fc.load_case(case_dir)
    open(u, "w").write(t3)
    # a solver the engine does not have
    fs = os.path.join(case_dir, "system", "fvSolution")
    t4 = open(fs).read()
    open(fs, "w").write(t4.replace("solver          PCG;", "solver          GAMG;"))
    with pytest.raises(fc.FoamCaseError, match="GAMG"):
        fc.load_case(case_dir)
    open(fs, "w").write(t4)
    # a discretisation scheme the kernels do not implement
    sch = os.path.join(case_dir, "system", "fvSchemes")
    t5 = open(sch).read()
    open(sch, "w").write(t5.replace("div(phi,U)      Gauss linear;", "div(phi,U)      Gauss upwind;"))
    with pytest.raises(fc.FoamCaseError, match="upwind"):
        fc.load_case(case_dir)
    if os.path.exists(BIN_PATH):
        import subprocess
        r = subprocess.run([BIN_PATH, "-case", case_dir, "-dump"], capture_output=True, text=True)
        assert r.returncode == 3 and "upwind" in r.stderr
    open(sch, "w").write(t5.replace("default         Euler;", "default         backward;"))
    with pytest.raises(fc.FoamCaseError, match="backward"):
        fc.load_case(case_dir)
    open(sch, "w").write(t5.replace("default         orthogonal;", "default         corrected;"))
    fc.load_case(case_dir)                                # (the same scheme on an orthogonal box)
    open(sch, "w").write(t5)
    fc.load_case(case_dir)


# ---- the C++ twin (host/foamCase.H behind the standalone solver binary foamYadeB200) ------------------------------
BIN = os.path.join(ROOT, "yade-openfoam-coupling_b200", "foamYadeB200")


def _fnv(a):
    h = 1469598103934665603
    for b in np.ascontiguousarray(a).tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


def _mesh_py():
    spec = importlib.util.spec_from_file_location("fy_mesh", os.path.join(ROOT, "yade-openfoam-coupling_b200", "mesh.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _dump(case_dir, *extra):
    import json
    import subprocess
    r = subprocess.run([BIN, "-case", case_dir, "-dump"] + list(extra), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return json.loads(r.stdout)


@pytest.mark.skipif(not os.path.exists(BIN), reason="foamYadeB200 not built")
def test_cpp_case_reader_agrees_with_the_python_one(tmp_path):
    """host/foamCase.H (C++, what the standalone solver binary uses) and foamcase.py read the same case to the same
    description, and build the same fy_mesh_desc arrays bit for bit (FNV hashes of the raw bytes) -- for the stock cavity
    tutorial and for a 3-D inlet / outlet / fixedFluxPressure case with nonuniform fields and the pimple dictionaries."""
    mesh_py = _mesh_py()
    code = {"fixedValue": 0, "zeroGradient": 1, "empty": 2, "fixedFluxPressure": 3}
    cases = [(cavity_case(tmp_path), "icoFoamYade")]
    # the 3-D case of the test above, written again
    chan = str(tmp_path / "chan3")
    n, L, org = (6, 5, 4), (2.0, 1.0, 0.5), (-1.0, 0.25, 3.0)
    patches = [("inlet", ["xmin"]), ("outlet", ["xmax"]), ("walls", ["ymin", "ymax", "zmin", "zmax"])]
    fc.write_box_poly_mesh(chan, n, L, org, patches, {"inlet": "patch", "outlet": "patch", "walls": "wall"})
    rng = np.random.default_rng(3)
    U, p = rng.standard_normal((120, 3)), rng.standard_normal(120)
    pl = [dict(name=a, sides=b) for a, b in patches]
    fc.write_field(chan, "0", "Uc", U, (0, 1, -1, 0, 0, 0, 0), pl,
                   {"inlet": ("fixedValue", np.array([0.3, 0, 0])), "outlet": ("zeroGradient", None), "walls": ("noSlip", None)}, prec=17)
    fc.write_field(chan, "0", "p", p, (0, 2, -2, 0, 0, 0, 0), pl,
                   {"inlet": ("zeroGradient", None), "outlet": ("fixedValue", 0.25), "walls": ("fixedFluxPressure", 0.0)}, prec=17)
    os.makedirs(os.path.join(chan, "system"))
    open(os.path.join(chan, "constant", "transportProperties"), "w").write(
        "FoamFile{version 2.0;format ascii;class dictionary;object transportProperties;}\n"
        "partDensity partDensity [1 -3 0 0 0 0 0] 2400;\nrhocValue rhocValue [1 -3 0 0 0 0 0] 998;\nnuValue nuValue [0 2 -1 0 0 0 0] 1e-06;\n")
    open(os.path.join(chan, "constant", "g"), "w").write(
        "FoamFile{version 2.0;format ascii;class uniformDimensionedVectorField;object g;}\ndimensions [0 1 -2 0 0 0 0];\nvalue (0 0 -9.81);\n")
    open(os.path.join(chan, "system", "controlDict"), "w").write(
        "FoamFile{version 2.0;format ascii;class dictionary;object controlDict;}\napplication pimpleFoamYade;\nstartTime 0;\nendTime 1;\n"
        "deltaT 1e-3;\nwriteControl timeStep;\nwriteInterval 100;\nwritePrecision 12;\n")
    open(os.path.join(chan, "system", "fvSolution"), "w").write(
        "FoamFile{version 2.0;format ascii;class dictionary;object fvSolution;}\n"
        "solvers { p { solver PCG; preconditioner diagonal; tolerance 1e-7; relTol 0.01; } pFinal { $p; relTol 0; }\n"
        '          "(Uc|k)" { solver smoothSolver; smoother symGaussSeidel; tolerance 1e-6; relTol 0.1; } }\n'
        "PIMPLE { nOuterCorrectors 3; nCorrectors 2; momentumPredictor no; nNonOrthogonalCorrectors 1; pRefCell 7; pRefValue 1.5; }\n"
        'relaxationFactors { equations { "Uc.*" 0.7; UcFinal 1; } fields { p 0.3; pFinal 1; } }\n')
    cases.append((chan, "pimpleFoamYade"))
    for case_dir, solver in cases:
        py = fc.load_case(case_dir, solver=solver)
        cc = _dump(case_dir, "-solver", solver)
        assert tuple(cc["n"]) == py["box"]["n"] and cc["Uname"] == py["Uname"]
        np.testing.assert_allclose(cc["origin"], py["box"]["origin"], rtol=0, atol=0)
        np.testing.assert_allclose(cc["L"], py["box"]["L"], rtol=0, atol=0)
        assert cc["nu"] == py["nu"] and tuple(cc["g"]) == tuple(py["g"])
        assert cc["deltaT"] == py["control"]["deltaT"] and cc["endTime"] == py["control"]["endTime"]
        assert cc["writeInterval"] == py["control"]["writeInterval"] and cc["writePrecision"] == py["control"]["writePrecision"]
        pre = {"DIC": 0, "diagonal": 1, "none": 2}
        for k, v in py["piso"].items():
            assert cc["piso"][k] == (pre[v] if k == "preconditioner" else v), k
        assert cc["pimple"] == py["pimple"]
        assert len(cc["patches"]) == len(py["patches"])
        for a, b in zip(cc["patches"], py["patches"]):
            assert (a["name"], a["type"], a["sides"], a["start"], a["nFaces"]) == (b["name"], b["type"], b["sides"], b["start"], b["nFaces"])
            assert (a["bcU"], a["bcP"]) == (code[b["bcU"]], code[b["bcP"]]) and tuple(a["valueU"]) == b["valueU"] and a["valueP"] == b["valueP"]
        if solver == "pimpleFoamYade":
            assert cc["rhoP"] == 2400.0 and cc["rhoF"] == 998.0
        m = fc.build_mesh(py, mesh_py.box_mesh, mesh_py.set_bc, mesh_py)
        hs = cc["hash"]
        assert hs["U"] == _fnv(py["U"]) and hs["p"] == _fnv(py["p"])
        for k in ("C", "V", "owner", "neighbour", "Sf", "magSf", "deltaCoeffs"):
            assert hs[k] == _fnv(m[k]), k
        for q, pt in enumerate(m["patches"]):
            assert hs["faceCells%d" % q] == _fnv(pt["faceCells"]) and hs["bSf%d" % q] == _fnv(pt["Sf"])
            assert hs["bDeltaCoeffs%d" % q] == _fnv(pt["deltaCoeffs"])
    # the writer: the fields as read, written as a time directory by the C++ side, read back by the Python side
    import subprocess
    r = subprocess.run([BIN, "-case", chan, "-solver", "pimpleFoamYade", "-writeNow", "0.25"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    back = fc.load_case(chan, time="0.25", solver="pimpleFoamYade")
    assert np.abs(back["U"] - U).max() <= 1e-11 * np.abs(U).max() and np.abs(back["p"] - p).max() <= 1e-11 * np.abs(p).max()
    assert [(q["name"], q["bcU"], q["valueU"], q["bcP"], q["valueP"]) for q in back["patches"]] == \
           [(q["name"], q["bcU"], q["valueU"], q["bcP"], q["valueP"]) for q in fc.load_case(chan, solver="pimpleFoamYade")["patches"]]
    # the same case with only a blockMeshDict on disk: same description; the fixedFluxPressure patch values the C++ writer
    # derives from the box equal the ones the Python writer derives
    import shutil as _sh
    _sh.rmtree(os.path.join(chan, "constant", "polyMesh"))
    _sh.rmtree(os.path.join(chan, "0.25"))
    open(os.path.join(chan, "system", "blockMeshDict"), "w").write(
        "FoamFile{version 2.0;format ascii;class dictionary;object blockMeshDict;}\nconvertToMeters 0.5;\n"
        "vertices ((-2 0.5 6) (2 0.5 6) (2 2.5 6) (-2 2.5 6) (-2 0.5 7) (2 0.5 7) (2 2.5 7) (-2 2.5 7));\n"
        "blocks (hex (0 1 2 3 4 5 6 7) (6 5 4) simpleGrading (1 1 1));\nedges ();\n"
        "boundary (inlet {type patch; faces ((0 4 7 3));} outlet {type patch; faces ((1 2 6 5));}\n"
        "          walls {type wall; faces ((0 1 5 4) (3 7 6 2) (0 3 2 1) (4 5 6 7));});\n")
    py2 = fc.load_case(chan, solver="pimpleFoamYade")
    cc2 = _dump(chan, "-solver", "pimpleFoamYade")
    assert py2["box"]["n"] == (6, 5, 4) and tuple(cc2["n"]) == (6, 5, 4)
    np.testing.assert_allclose(py2["box"]["origin"], org, rtol=0, atol=1e-15)
    np.testing.assert_allclose(cc2["origin"], org, rtol=0, atol=1e-15)
    assert [q["sides"] for q in cc2["patches"]] == [q["sides"] for q in py2["patches"]] == [["xmin"], ["xmax"], ["ymin", "ymax", "zmin", "zmax"]]
    assert [(q["start"], q["nFaces"]) for q in cc2["patches"]] == [(q["start"], q["nFaces"]) for q in py2["patches"]]
    r = subprocess.run([BIN, "-case", chan, "-solver", "pimpleFoamYade", "-writeNow", "0.25"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    d_cpp = fc.read_dict(os.path.join(chan, "0.25", "p"))
    os.rename(os.path.join(chan, "0.25"), os.path.join(chan, "0.25_cpp"))
    py2["control"]["writePrecision"] = 12
    fc.write_time(py2, 0.25, py2["U"], py2["p"])
    d_py = fc.read_dict(os.path.join(chan, "0.25", "p"))
    w_cpp = [float(x) for x in d_cpp["boundaryField"]["walls"]["value"][-1]]
    w_py = [float(x) for x in d_py["boundaryField"]["walls"]["value"][-1]]
    assert len(w_cpp) == 2 * 6 * 4 + 2 * 6 * 5 and w_cpp == w_py
    # refusals carry the reason
    u = os.path.join(cases[0][0], "0", "U")
    t3 = open(u).read()
    open(u, "w").write(t3.replace("type            noSlip;", "type            slip;"))
    r = subprocess.run([BIN, "-case", cases[0][0], "-dump"], capture_output=True, text=True)
    assert r.returncode == 3 and "slip" in r.stderr
