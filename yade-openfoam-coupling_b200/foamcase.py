"""OpenFOAM case directories for the standalone engine.

The reference's solvers get their mesh, fields and settings from OpenFOAM's runtime (icoFoamYade/createFields.H:3-45,
166-168, pimpleFoamYade/createFields.H:3-107) and write time directories through runTime.write() (icoFoamYade.C:142).
With OpenFOAM present the C++ host class (host/FoamYadeB200.H) is handed the fvMesh OpenFOAM has read; WITHOUT it this
module reads the same ASCII files, so that a case prepared for the reference runs on the engine as it lies on disk:

    constant/polyMesh/{points,faces,owner,neighbour,boundary}   -> the hex box behind it (cells x fastest, as blockMesh
                                                                    writes them) and which box sides each patch covers
    0/U, 0/p                                                     -> internal fields + patch types / values
    constant/transportProperties (nu | nuValue, partDensity, rhocValue), constant/g
    system/controlDict (deltaT, startTime, endTime, writeInterval, writePrecision)
    system/fvSolution (solvers p / pFinal / U, PISO | PIMPLE, relaxationFactors)
and writes <time>/U, <time>/p, <time>/phi back in OpenFOAM's format.  Scope: what the device FV path supports -- uniform
hex boxes whose patches are unions of whole box sides; patch types fixedValue / noSlip / zeroGradient / empty /
fixedFluxPressure; ASCII files.  Anything else raises FoamCaseError naming the file and the entry.

Nothing here touches the GPU or the oracle: `load_case` returns a plain description; `build_mesh(case, box_mesh, set_bc,
consts)` turns it into a mesh dict with the caller's generator (the package's `box_mesh` for the engine,
`oracle.meshgen.hex_box_ldu` in the tests)."""
import os
import re

import numpy as np

SIDES = ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")


class FoamCaseError(ValueError):
    pass


# ---------------------------------------------------------------------------------------------------------------
# dictionary syntax
# ---------------------------------------------------------------------------------------------------------------
_PUNCT = "{}()[];"


def _tokenize(t):
    """words, numbers, strings and the punctuation { } ( ) [ ] ;  -- a WORD (it starts with a letter, `_`, `$` or `.`) keeps
    balanced parentheses that follow it without white space, as OpenFOAM's keywords do: div(phi,U), grad(p)"""
    toks, i, n = [], 0, len(t)
    while i < n:
        c = t[i]
        if c.isspace():
            i += 1
        elif c == '"':
            j = i + 1
            while j < n and t[j] != '"':
                j += 2 if t[j] == "\\" else 1
            toks.append(t[i:j + 1])
            i = j + 1
        elif c in _PUNCT:
            toks.append(c)
            i += 1
        else:
            j, depth, wordy = i, 0, (c.isalpha() or c in "_$.")
            while j < n:
                ch = t[j]
                if ch == "(" and wordy and j > i:
                    depth += 1
                elif ch == ")" and depth > 0:
                    depth -= 1
                elif depth == 0 and (ch.isspace() or ch in _PUNCT or ch == '"'):
                    break
                j += 1
            toks.append(t[i:j])
            i = j
    return toks


def _strip_comments(text):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return re.sub(r"//[^\n]*", " ", text)


class FoamList(list):
    """a parenthesised list; `.prefix` keeps what stood in front of it (e.g. 'List<vector>' and the count)"""
    prefix = ()


def _parse_value_tokens(toks, pos, end_chars):
    """values up to one of end_chars at nesting level 0 -> (list of values, position of the terminator)"""
    vals = []
    while pos < len(toks) and toks[pos] not in end_chars:
        t = toks[pos]
        if t == "(":
            lst, pos = _parse_value_tokens(toks, pos + 1, (")",))
            out = FoamList(lst)
            vals.append(out)
            pos += 1
        elif t == "[":
            lst, pos = _parse_value_tokens(toks, pos + 1, ("]",))
            vals.append(("dimensions", tuple(lst)))
            pos += 1
        elif t == "{":
            d, pos = _parse_dict_tokens(toks, pos + 1)
            vals.append(d)
            pos += 1
        else:
            vals.append(_atom(t))
            pos += 1
    if pos >= len(toks):
        raise FoamCaseError("unterminated entry (expected one of %r)" % (end_chars,))
    return vals, pos


def _atom(t):
    if t.startswith('"'):
        return t[1:-1]
    try:
        return int(t)
    except ValueError:
        pass
    try:
        return float(t)
    except ValueError:
        return t


def _lookup(name, scopes):
    for sc in scopes:
        if name in sc:
            return sc[name]
    return None


def _parse_dict_tokens(toks, pos, parents=()):
    """entries up to the closing brace (or the end of the token stream at top level)"""
    d = {}
    scopes = (d,) + tuple(parents)
    while pos < len(toks) and toks[pos] != "}":
        key = toks[pos]
        pos += 1
        if key.startswith("$") and pos < len(toks) and toks[pos] == ";":        # `$p;` : merge a dictionary of an enclosing scope
            src = _lookup(key[1:], scopes)
            if not isinstance(src, dict):
                raise FoamCaseError("macro %s does not name a dictionary of the enclosing scopes" % key)
            d.update({k: v for k, v in src.items()})
            pos += 1
            continue
        if key.startswith('"'):
            key = key[1:-1]
        if pos < len(toks) and toks[pos] == "{":
            sub, pos = _parse_dict_tokens(toks, pos + 1, scopes)
            if pos >= len(toks):
                raise FoamCaseError("dictionary %s is not closed" % key)
            d[key] = sub
            pos += 1
            continue
        vals, pos = _parse_value_tokens(toks, pos, (";",))
        pos += 1
        vals = [(_lookup(v[1:], scopes) if _lookup(v[1:], scopes) is not None else v) if isinstance(v, str) and v.startswith("$") else v
                for v in vals]
        d[key] = vals[0] if len(vals) == 1 else vals
    return d, pos


def parse_dict(text):
    """An OpenFOAM dictionary file (FoamFile header included) as nested dicts.  `key v1 v2 ...;` keeps a list of values,
    a dimensioned entry `nu [0 2 -1 0 0 0 0] 0.01;` becomes [('dimensions', (...)), 0.01]."""
    toks = _tokenize(_strip_comments(text))
    d, pos = _parse_dict_tokens(toks, 0)
    if pos != len(toks):
        raise FoamCaseError("unbalanced braces")
    return d


def read_dict(path):
    try:
        with open(path) as f:
            return parse_dict(f.read())
    except FoamCaseError as e:
        raise FoamCaseError("%s: %s" % (path, e))


def scalar_of(entry, what):
    """value of `key 0.01;`, `key [dims] 0.01;` or `key key [dims] 0.01;` (the old dimensionedScalar form)"""
    vals = entry if isinstance(entry, list) and not isinstance(entry, FoamList) else [entry]
    nums = [v for v in vals if isinstance(v, (int, float)) and not isinstance(v, bool)]
    if len(nums) != 1:
        raise FoamCaseError("%s: expected one number, got %r" % (what, entry))
    return float(nums[0])


def vector_of(entry, what):
    vals = entry if isinstance(entry, list) and not isinstance(entry, FoamList) else [entry]
    for v in vals:
        if isinstance(v, FoamList) and len(v) == 3:
            return tuple(float(x) for x in v)
    raise FoamCaseError("%s: expected a vector (x y z), got %r" % (what, entry))


# ---------------------------------------------------------------------------------------------------------------
# polyMesh
# ---------------------------------------------------------------------------------------------------------------
def _split_header(text, path):
    text = _strip_comments(text)
    m = re.search(r"FoamFile\s*\{(.*?)\}", text, flags=re.S)
    head = parse_dict(m.group(1)) if m else {}
    if head.get("format", "ascii") != "ascii":
        raise FoamCaseError("%s: format %s (only ascii files are read)" % (path, head.get("format")))
    return head, (text[m.end():] if m else text)


def _numbers(body, path, dtype):
    """the numbers of `N ( ... )` with every parenthesis dropped -> (N, flat array)"""
    m = re.search(r"(\d+)\s*\(", body)
    if not m:
        raise FoamCaseError("%s: no `N (` list found" % path)
    n = int(m.group(1))
    end = body.rfind(")")
    flat = body[m.end():end].replace("(", " ").replace(")", " ")
    arr = np.array(flat.split(), dtype=dtype)
    return n, arr


def read_poly_mesh(case_dir):
    pm = os.path.join(case_dir, "constant", "polyMesh")
    out = {}
    for name, dtype in (("points", np.float64), ("faces", np.int64), ("owner", np.int64), ("neighbour", np.int64)):
        path = os.path.join(pm, name)
        if not os.path.exists(path):
            raise FoamCaseError("%s: missing" % path)
        with open(path) as f:
            head, body = _split_header(f.read(), path)
        n, arr = _numbers(body, path, dtype)
        if name == "points":
            if arr.size != 3 * n:
                raise FoamCaseError("%s: %d numbers for %d points" % (path, arr.size, n))
            out[name] = arr.reshape(n, 3)
        elif name == "faces":
            if arr.size != 5 * n or np.any(arr.reshape(n, 5)[:, 0] != 4):
                raise FoamCaseError("%s: only quadrilateral faces `4(a b c d)` (hex meshes) are supported" % path)
            out[name] = arr.reshape(n, 5)[:, 1:]
        else:
            if arr.size != n:
                raise FoamCaseError("%s: %d labels, header says %d" % (path, arr.size, n))
            out[name] = arr
    path = os.path.join(pm, "boundary")
    with open(path) as f:
        head, body = _split_header(f.read(), path)
    m = re.search(r"(\d+)\s*\(", body)
    if not m:
        raise FoamCaseError("%s: no patch list" % path)
    d = parse_dict(body[m.end():body.rfind(")")])
    out["boundary"] = [(k, v) for k, v in d.items()]
    if len(out["boundary"]) != int(m.group(1)):
        raise FoamCaseError("%s: %d patches listed, %d found" % (path, int(m.group(1)), len(out["boundary"])))
    return out


def detect_hex_box(pm, rtol=1e-9):
    """The uniform hex box behind a polyMesh: (n, origin, L, patches) with patches = [(name, type, [box sides])].
    Cells must be numbered x fastest (blockMesh's order for one block) and every patch must cover whole sides."""
    pts, faces, own, nei = pm["points"], pm["faces"], pm["owner"], pm["neighbour"]
    nF, nFi = faces.shape[0], nei.shape[0]
    if own.shape[0] != nF:
        raise FoamCaseError("polyMesh: owner has %d entries for %d faces" % (own.shape[0], nF))
    N = int(own.max()) + 1
    lo, hi = pts.min(0), pts.max(0)
    L = hi - lo
    fc = pts[faces].mean(1)                                   # face centres (planar quads)
    # cell centres = mean of the cell's face centres (exact for a box cell)
    C = np.zeros((N, 3))
    cnt = np.zeros(N)
    np.add.at(C, own, fc)
    np.add.at(cnt, own, 1.0)
    np.add.at(C, nei, fc[:nFi])
    np.add.at(cnt, nei, 1.0)
    if np.any(cnt != 6):
        raise FoamCaseError("polyMesh: not every cell has six faces (hex meshes only)")
    C /= 6.0
    n = []
    for d in range(3):
        # distinct centre coordinates along d (tolerant unique)
        u = np.unique(np.round((C[:, d] - lo[d]) / (L[d] if L[d] > 0 else 1.0) / rtol).astype(np.int64))
        n.append(len(u))
    nx, ny, nz = n
    if nx * ny * nz != N:
        raise FoamCaseError("polyMesh: %d cells are not an %d x %d x %d box" % (N, nx, ny, nz))
    h = L / np.array(n)
    c = np.arange(N)
    i, j, k = c % nx, (c // nx) % ny, c // (nx * ny)
    want = lo + (np.stack([i, j, k], 1) + 0.5) * h
    if np.abs(C - want).max() > 1e-7 * h.min():
        raise FoamCaseError("polyMesh: cells are not a uniform box numbered x fastest (blockMesh single-block order)")
    # internal faces must be in upper-triangular order with the box's neighbours
    if np.any(own[:nFi] >= nei) or np.any(np.diff(own[:nFi]) < 0):
        raise FoamCaseError("polyMesh: internal faces are not in upper-triangular (owner-sorted) order")
    patches, covered = [], {}
    for name, pd in pm["boundary"]:
        start, nf = int(pd["startFace"]), int(pd["nFaces"])
        sides = []
        if nf:
            f = fc[start:start + nf]
            on = {"xmin": np.abs(f[:, 0] - lo[0]) < 1e-7 * h[0], "xmax": np.abs(f[:, 0] - hi[0]) < 1e-7 * h[0],
                  "ymin": np.abs(f[:, 1] - lo[1]) < 1e-7 * h[1], "ymax": np.abs(f[:, 1] - hi[1]) < 1e-7 * h[1],
                  "zmin": np.abs(f[:, 2] - lo[2]) < 1e-7 * h[2], "zmax": np.abs(f[:, 2] - hi[2]) < 1e-7 * h[2]}
            size = {"xmin": ny * nz, "xmax": ny * nz, "ymin": nx * nz, "ymax": nx * nz, "zmin": nx * ny, "zmax": nx * ny}
            total = 0
            for s in SIDES:
                m = int(on[s].sum())
                if m == 0:
                    continue
                if m != size[s]:
                    raise FoamCaseError("polyMesh: patch %s covers %d of the %d faces of side %s (whole sides only)" % (name, m, size[s], s))
                if s in covered:
                    raise FoamCaseError("polyMesh: side %s belongs to patches %s and %s" % (s, covered[s], name))
                covered[s] = name
                sides.append(s)
                total += m
            if total != nf:
                raise FoamCaseError("polyMesh: patch %s has faces that lie on no box side" % name)
        patches.append((name, str(pd.get("type", "patch")), sides, start, nf))
    if len(covered) != 6:
        raise FoamCaseError("polyMesh: box sides without a patch: %s" % sorted(set(SIDES) - set(covered)))
    return dict(n=(nx, ny, nz), origin=tuple(lo), L=tuple(L), patches=patches)


def read_block_mesh_dict(case_dir):
    """The box straight from system/blockMeshDict (what a tutorial directory holds BEFORE `blockMesh` has been run): one `hex`
    block with axis-aligned edges (local x1, x2, x3 = global x, y, z, so that blockMesh's cell order is x fastest),
    uniform grading, every block face listed in a boundary patch.  Returns what detect_hex_box returns; boundary faces are
    numbered patch by patch after the internal faces, side by side, by increasing owner cell (this module's own order --
    blockMesh's order inside a patch differs, which no patch type supported here can see)."""
    path = os.path.join(case_dir, "system", "blockMeshDict")
    if not os.path.exists(path):
        path = os.path.join(case_dir, "constant", "polyMesh", "blockMeshDict")
    if not os.path.exists(path):
        raise FoamCaseError("%s: neither constant/polyMesh nor a blockMeshDict" % case_dir)
    d = read_dict(path)
    scale = float(d.get("convertToMeters", d.get("scale", 1.0)))
    try:
        verts = np.array([[float(x) for x in v] for v in d["vertices"]], dtype=np.float64) * scale
        blocks = d["blocks"]
    except (KeyError, TypeError, ValueError):
        raise FoamCaseError("%s: vertices / blocks missing or malformed" % path)
    if sum(1 for b in blocks if b == "hex") != 1 or len(blocks) < 3:
        raise FoamCaseError("%s: exactly one hex block is supported" % path)
    k = list(blocks).index("hex")
    hexv, ncell = [int(x) for x in blocks[k + 1]], [int(x) for x in blocks[k + 2]]
    grading = [g for g in blocks[k + 3:] if isinstance(g, FoamList)]
    if len(hexv) != 8 or len(ncell) != 3:
        raise FoamCaseError("%s: hex block needs 8 vertices and 3 cell counts" % path)
    if grading and any(float(x) != 1.0 for x in grading[0]):
        raise FoamCaseError("%s: graded blocks are not supported (uniform cells only)" % path)
    v = verts[hexv]
    lo, hi = v.min(0), v.max(0)
    want = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], dtype=np.float64)
    if np.abs(v - (lo + want * (hi - lo))).max() > 1e-12 * np.abs(hi - lo).max():
        raise FoamCaseError("%s: the block is not an axis-aligned box with local axes x1 x2 x3 = x y z" % path)
    nx, ny, nz = ncell
    size = {"xmin": ny * nz, "xmax": ny * nz, "ymin": nx * nz, "ymax": nx * nz, "zmin": nx * ny, "zmax": nx * ny}
    bnd = d.get("boundary")
    if not isinstance(bnd, FoamList):
        raise FoamCaseError("%s: no boundary list" % path)
    patches, covered = [], {}
    start = 3 * nx * ny * nz - nx * ny - ny * nz - nx * nz
    it = iter(bnd)
    for name in it:
        pd = next(it, None)
        if not isinstance(name, str) or not isinstance(pd, dict):
            raise FoamCaseError("%s: boundary must alternate patch names and dictionaries" % path)
        sides = []
        for f in pd.get("faces", []):
            fv = verts[[int(x) for x in f]]
            side = None
            for ax, nm in enumerate("xyz"):
                if np.all(np.abs(fv[:, ax] - lo[ax]) <= 1e-12 * (hi[ax] - lo[ax])):
                    side = nm + "min"
                elif np.all(np.abs(fv[:, ax] - hi[ax]) <= 1e-12 * (hi[ax] - lo[ax])):
                    side = nm + "max"
            if side is None:
                raise FoamCaseError("%s: patch %s: face %s is not a side of the block" % (path, name, list(f)))
            if side in covered:
                raise FoamCaseError("%s: side %s belongs to patches %s and %s" % (path, side, covered[side], name))
            covered[side] = name
            sides.append(side)
        nf = sum(size[s_] for s_ in sides)
        patches.append((name, str(pd.get("type", "patch")), sides, start, nf))
        start += nf
    if len(covered) != 6:
        raise FoamCaseError("%s: block faces without a patch: %s (defaultFaces is not supported)" % (path, sorted(set(SIDES) - set(covered))))
    return dict(n=(nx, ny, nz), origin=tuple(lo), L=tuple(hi - lo), patches=patches, from_block_mesh_dict=True)


def read_box(case_dir):
    """the hex box of a case: from constant/polyMesh when it exists, else from the blockMeshDict"""
    if os.path.exists(os.path.join(case_dir, "constant", "polyMesh", "points")):
        return detect_hex_box(read_poly_mesh(case_dir))
    return read_block_mesh_dict(case_dir)


# ---------------------------------------------------------------------------------------------------------------
# fields
# ---------------------------------------------------------------------------------------------------------------
def _field_values(entry, ncomp, count, what):
    """`uniform v` / `nonuniform List<T> N ( ... )` -> array [count][ncomp] (or [count])"""
    vals = entry if isinstance(entry, list) and not isinstance(entry, FoamList) else [entry]
    kind = vals[0]
    if kind == "uniform":
        v = vals[1]
        a = np.array(list(v) if isinstance(v, FoamList) else [v], dtype=np.float64)
        if a.size != ncomp:
            raise FoamCaseError("%s: uniform value with %d components, expected %d" % (what, a.size, ncomp))
        return np.tile(a, (count, 1)) if ncomp > 1 else np.full(count, a[0])
    if kind == "nonuniform":
        lst = [v for v in vals if isinstance(v, FoamList)]
        if not lst:
            raise FoamCaseError("%s: nonuniform without a list" % what)
        a = np.array([list(x) if isinstance(x, FoamList) else x for x in lst[0]], dtype=np.float64)
        a = a.reshape(-1, ncomp) if ncomp > 1 else a.reshape(-1)
        if a.shape[0] != count:
            raise FoamCaseError("%s: %d values for %d entries" % (what, a.shape[0], count))
        return a
    raise FoamCaseError("%s: expected uniform / nonuniform, got %r" % (what, kind))


def read_field(path, ncomp, n_cells, patches):
    """-> (internal [N][ncomp], {patch: (type, value or None, raw dict)})"""
    with open(path) as f:
        text = _strip_comments(f.read())
    # a large nonuniform internalField is pulled out before the generic parser sees it (speed)
    internal = None
    m = re.search(r"internalField\s+nonuniform\s+List<\w+>\s*(\d+)\s*\(", text)
    if m:
        cnt = int(m.group(1))
        depth, p = 1, m.end()
        while depth:
            ch = text[p]
            depth += ch == "("
            depth -= ch == ")"
            p += 1
        flat = text[m.end():p - 1].replace("(", " ").replace(")", " ")
        a = np.array(flat.split(), dtype=np.float64)
        if a.size != cnt * ncomp or cnt != n_cells:
            raise FoamCaseError("%s: internalField has %d numbers for %d cells x %d" % (path, a.size, n_cells, ncomp))
        internal = a.reshape(cnt, ncomp) if ncomp > 1 else a
        text = text[:m.start()] + text[text.index(";", p - 1) + 1:]
    try:
        d = parse_dict(text)
    except FoamCaseError as e:
        raise FoamCaseError("%s: %s" % (path, e))
    if internal is None:
        if "internalField" not in d:
            raise FoamCaseError("%s: no internalField" % path)
        internal = _field_values(d["internalField"], ncomp, n_cells, path + ": internalField")
    bf = d.get("boundaryField")
    if not isinstance(bf, dict):
        raise FoamCaseError("%s: no boundaryField" % path)
    out = {}
    for name, ptype, sides, start, nf in patches:
        pd = match_key(bf, name)                              # (regular-expression keys: "(left|right)", ".*")
        if pd is None:
            raise FoamCaseError("%s: boundaryField has no entry for patch %s" % (path, name))
        val = None
        if "value" in pd:
            v = _field_values(pd["value"], ncomp, max(nf, 1), "%s: %s.value" % (path, name))
            if nf and np.abs(v - v[0]).max() > 0:
                if pd.get("type") == "fixedValue":
                    raise FoamCaseError("%s: patch %s: a non-uniform fixedValue is not supported" % (path, name))
            val = v[0]
        out[name] = (str(pd.get("type")), val, pd)
    return internal, out


# ---------------------------------------------------------------------------------------------------------------
# the case
# ---------------------------------------------------------------------------------------------------------------
def match_key(d, name):
    """OpenFOAM's dictionary lookup: an exact keyword first, then the regular-expression keywords, last defined first"""
    if name in d:
        return d[name]
    for k in reversed(list(d.keys())):
        try:
            if re.fullmatch(k, name):
                return d[k]
        except re.error:
            continue
    return None


def _solver_entry(sol, name, path):
    e = match_key(sol, name)
    if e is None:
        raise FoamCaseError("%s: solvers has no entry for %s" % (path, name))
    return e


def _words(entry):
    vals = entry if isinstance(entry, list) and not isinstance(entry, FoamList) else [entry]
    return [str(v) for v in vals]


def check_schemes(case_dir):
    """system/fvSchemes must ask for what the kernels implement (the stock cavity set): Euler ddt; Gauss linear gradients,
    divergences and laplacians; linear interpolation; the surface-normal gradient and the laplacian's may be orthogonal,
    uncorrected or corrected (the same thing on the orthogonal boxes this path supports).  A missing file is accepted
    (the standalone engine has no other schemes to choose from); anything else is refused with the entry named."""
    path = os.path.join(case_dir, "system", "fvSchemes")
    if not os.path.exists(path):
        return
    d = read_dict(path)
    ng = ("orthogonal", "uncorrected", "corrected")

    def ok(section, key, words):
        if section == "ddtSchemes":
            return words == ["Euler"]
        if section == "gradSchemes":
            return words == ["Gauss", "linear"]
        if section == "divSchemes":
            return words == ["none"] or words == ["Gauss", "linear"]
        if section == "laplacianSchemes":
            return words[:2] == ["Gauss", "linear"] and len(words) == 3 and words[2] in ng
        if section == "interpolationSchemes":
            return words == ["linear"]
        if section == "snGradSchemes":
            return len(words) == 1 and words[0] in ng
        return True
    for section, entries in d.items():
        if section == "FoamFile" or not isinstance(entries, dict):
            continue
        for key, val in entries.items():
            w = _words(val)
            if not ok(section, key, w):
                raise FoamCaseError("%s: %s { %s %s; } is not implemented (Euler; Gauss linear; linear; orthogonal)" % (path, section, key, " ".join(w)))


def load_case(case_dir, time="0", solver="icoFoamYade"):
    """Everything the engine needs from a case directory.  solver: 'icoFoamYade' (fields U, p; transportProperties nu; PISO
    dictionary) or 'pimpleFoamYade' (fields Uc | U, p; nuValue / rhocValue / partDensity; PIMPLE dictionary; constant/g)."""
    check_schemes(case_dir)
    box = read_box(case_dir)
    nx, ny, nz = box["n"]
    N = nx * ny * nz
    pimple = solver == "pimpleFoamYade"
    tdir = os.path.join(case_dir, time)
    uname = "Uc" if pimple and os.path.exists(os.path.join(tdir, "Uc")) else "U"
    U, bU = read_field(os.path.join(tdir, uname), 3, N, box["patches"])
    p, bP = read_field(os.path.join(tdir, "p"), 1, N, box["patches"])
    patch_bc = []
    for name, ptype, sides, start, nf in box["patches"]:
        tU, vU, _ = bU[name]
        tP, vP, _ = bP[name]
        if ptype == "empty" or tU == "empty" or tP == "empty":
            if not (tU == "empty" and tP == "empty"):
                raise FoamCaseError("patch %s: empty must be set on the patch and on both fields" % name)
            kU, kP, vU, vP = "empty", "empty", (0.0, 0.0, 0.0), 0.0
        else:
            if tU in ("fixedValue", "movingWallVelocity"):
                if vU is None:
                    raise FoamCaseError("%s: patch %s: fixedValue without value" % (uname, name))
                kU = "fixedValue"
            elif tU == "noSlip":
                kU, vU = "fixedValue", np.zeros(3)
            elif tU in ("zeroGradient", "inletOutlet", "pressureInletOutletVelocity"):
                if tU != "zeroGradient":
                    raise FoamCaseError("%s: patch %s: type %s is not supported (zeroGradient is)" % (uname, name, tU))
                kU, vU = "zeroGradient", np.zeros(3)
            else:
                raise FoamCaseError("%s: patch %s: type %s is not supported" % (uname, name, tU))
            if tP == "zeroGradient":
                kP, vP = "zeroGradient", 0.0
            elif tP == "fixedValue":
                if vP is None:
                    raise FoamCaseError("p: patch %s: fixedValue without value" % name)
                kP = "fixedValue"
            elif tP == "fixedFluxPressure":
                kP, vP = "fixedFluxPressure", 0.0
            else:
                raise FoamCaseError("p: patch %s: type %s is not supported" % (name, tP))
        patch_bc.append(dict(name=name, type=ptype, sides=sides, start=start, nFaces=nf, bcU=kU,
                             valueU=tuple(float(x) for x in np.atleast_1d(vU)), bcP=kP, valueP=float(vP)))
    tp_path = os.path.join(case_dir, "constant", "transportProperties")
    tp = read_dict(tp_path)
    props = {}
    if pimple:
        for k in ("nuValue", "rhocValue", "partDensity"):
            if k in tp:
                props[k] = scalar_of(tp[k], "%s: %s" % (tp_path, k))
        nu = props.get("nuValue", scalar_of(tp["nu"], tp_path + ": nu") if "nu" in tp else None)
    else:
        nu = scalar_of(tp["nu"], tp_path + ": nu") if "nu" in tp else None
        for k in ("partDensity", "fluidDensity"):
            if k in tp:
                props[k] = scalar_of(tp[k], "%s: %s" % (tp_path, k))
    if nu is None:
        raise FoamCaseError("%s: no viscosity entry (nu / nuValue)" % tp_path)
    g = (0.0, 0.0, 0.0)
    gp = os.path.join(case_dir, "constant", "g")
    if os.path.exists(gp):
        g = vector_of(read_dict(gp)["value"], gp + ": value")
    cd_path = os.path.join(case_dir, "system", "controlDict")
    cd = read_dict(cd_path)
    control = dict(deltaT=scalar_of(cd["deltaT"], cd_path + ": deltaT"), startTime=float(cd.get("startTime", 0)),
                   endTime=scalar_of(cd["endTime"], cd_path + ": endTime"), writeControl=str(cd.get("writeControl", "timeStep")),
                   writeInterval=float(cd.get("writeInterval", 1)), writePrecision=int(cd.get("writePrecision", 6)),
                   application=str(cd.get("application", solver)))
    fs_path = os.path.join(case_dir, "system", "fvSolution")
    fs = read_dict(fs_path)
    sol = fs.get("solvers", {})
    sp, su = _solver_entry(sol, "p", fs_path), _solver_entry(sol, uname, fs_path)
    try:
        spf = _solver_entry(sol, "pFinal", fs_path)
    except FoamCaseError:
        spf = sp
    if sp.get("solver") != "PCG":
        raise FoamCaseError("%s: p solver %s is not supported (PCG is)" % (fs_path, sp.get("solver")))
    if su.get("solver") != "smoothSolver" or su.get("smoother") != "symGaussSeidel":
        raise FoamCaseError("%s: %s solver must be smoothSolver / symGaussSeidel" % (fs_path, uname))
    algo = fs.get("PIMPLE" if pimple else "PISO", {})
    piso = dict(nCorrectors=int(algo.get("nCorrectors", 1 if pimple else 2)),
                nNonOrthogonalCorrectors=int(algo.get("nNonOrthogonalCorrectors", 0)),
                momentumPredictor=1 if str(algo.get("momentumPredictor", "yes")) in ("yes", "on", "true", "1") else 0,
                pRefCell=int(algo.get("pRefCell", 0)), pRefValue=float(algo.get("pRefValue", 0.0)),
                pTol=float(sp.get("tolerance", 1e-6)), pRelTol=float(sp.get("relTol", 0.0)),
                pFinalTol=float(spf.get("tolerance", 1e-6)), pFinalRelTol=float(spf.get("relTol", 0.0)),
                UTol=float(su.get("tolerance", 1e-6)), URelTol=float(su.get("relTol", 0.0)),
                maxIter=int(sp.get("maxIter", 1000)), preconditioner=str(sp.get("preconditioner", "DIC")))
    if piso["preconditioner"] not in ("DIC", "diagonal", "none"):
        raise FoamCaseError("%s: p preconditioner %s is not supported" % (fs_path, piso["preconditioner"]))
    rf = fs.get("relaxationFactors", {})
    eq, fl = rf.get("equations", {}), rf.get("fields", {})

    def factor(d, key):
        v = match_key(d, key)
        return 0.0 if v is None else float(v)
    pimple_ctl = dict(nOuterCorrectors=int(algo.get("nOuterCorrectors", 1)), relaxU=factor(eq, uname), relaxUFinal=factor(eq, uname + "Final"),
                      relaxP=factor(fl, "p"), relaxPFinal=factor(fl, "pFinal"))
    return dict(case_dir=case_dir, solver=solver, box=box, patches=patch_bc, U=U, p=p, Uname=uname, nu=nu, props=props, g=g,
                control=control, piso=piso, pimple=pimple_ctl)


def build_mesh(case, box_mesh, set_bc, consts):
    """mesh dict from a generator pair: box_mesh(nx, ny, nz, lx, ly, lz, origin=..., patches=...) and set_bc(mesh, name, ...);
    consts = an object with BC_FIXED_VALUE / BC_ZERO_GRADIENT / BC_EMPTY / BC_FIXED_FLUX_PRESSURE"""
    b = case["box"]
    groups = [(p["name"], p["sides"]) for p in case["patches"] if p["sides"]]
    m = box_mesh(*b["n"], *b["L"], origin=b["origin"], patches=groups)
    code = {"fixedValue": consts.BC_FIXED_VALUE, "zeroGradient": consts.BC_ZERO_GRADIENT, "empty": consts.BC_EMPTY,
            "fixedFluxPressure": consts.BC_FIXED_FLUX_PRESSURE}
    for p in case["patches"]:
        if p["sides"]:
            set_bc(m, p["name"], bcU=code[p["bcU"]], valueU=p["valueU"], bcP=code[p["bcP"]], valueP=p["valueP"])
    return m


# ---------------------------------------------------------------------------------------------------------------
# writing
# ---------------------------------------------------------------------------------------------------------------
_HEADER = """/*--------------------------------*- C++ -*----------------------------------*\\
  =========                 |
  \\\\      /  F ield         | OpenFOAM: The Open Source CFD Toolbox
   \\\\    /   O peration     | Website:  https://openfoam.org
    \\\\  /    A nd           | Version:  6
     \\\\/     M anipulation  |
\\*---------------------------------------------------------------------------*/
FoamFile
{
    version     2.0;
    format      ascii;
    class       %s;
    location    "%s";
    object      %s;
}
// * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * //

"""


def _fmt(x, prec):
    return ("%." + str(prec) + "g") % x


def _list(a, prec):
    a = np.asarray(a)
    if a.ndim == 1:
        return "\n".join(_fmt(v, prec) for v in a)
    return "\n".join("(" + " ".join(_fmt(v, prec) for v in row) + ")" for row in a)


def time_name(t, prec=6):
    """runTime.timeName(): the time value at timePrecision significant digits"""
    s = ("%." + str(prec) + "g") % t
    return s


def write_field(case_dir, tname, obj, values, dims, patches, patch_entries, prec=6, cls=None):
    """One field file <case>/<time>/<obj>.  patch_entries: {patch: (type, value array or None)}"""
    values = np.asarray(values)
    vec = values.ndim == 2
    cls = cls or ("volVectorField" if vec else "volScalarField")
    os.makedirs(os.path.join(case_dir, tname), exist_ok=True)
    with open(os.path.join(case_dir, tname, obj), "w") as f:
        f.write(_HEADER % (cls, tname, obj))
        f.write("dimensions      [%s];\n\n" % " ".join(str(d) for d in dims))
        f.write("internalField   nonuniform List<%s> \n%d\n(\n%s\n)\n;\n\n" % ("vector" if vec else "scalar", values.shape[0], _list(values, prec)))
        f.write("boundaryField\n{\n")
        for p in patches:
            ptype, val = patch_entries[p["name"]]
            f.write("    %s\n    {\n        type            %s;\n" % (p["name"], ptype))
            if val is not None:
                val = np.asarray(val)
                if val.ndim == (2 if vec else 1):
                    f.write("        value           nonuniform List<%s> \n%d\n(\n%s\n)\n;\n" % ("vector" if vec else "scalar", val.shape[0], _list(val, prec)))
                elif vec:
                    f.write("        value           uniform (%s);\n" % " ".join(_fmt(v, prec) for v in val))
                else:
                    f.write("        value           uniform %s;\n" % _fmt(float(val), prec))
            f.write("    }\n")
        f.write("}\n\n\n// ************************************************************************* //\n")


def boundary_owner_cells(case):
    """owner cell of every boundary face in the polyMesh's own face order, per patch (for zeroGradient patch values)"""
    if case["box"].get("from_block_mesh_dict"):
        # no polyMesh on disk: this module's own boundary order (side by side, by increasing owner cell; build_mesh's)
        nx, ny, nz = case["box"]["n"]
        c = np.arange(nx * ny * nz)
        i, j, k = c % nx, (c // nx) % ny, c // (nx * ny)
        on = dict(xmin=i == 0, xmax=i == nx - 1, ymin=j == 0, ymax=j == ny - 1, zmin=k == 0, zmax=k == nz - 1)
        return {p["name"]: np.concatenate([c[on[s_]] for s_ in p["sides"]] or [c[:0]]) for p in case["patches"]}
    pm = read_poly_mesh(case["case_dir"])
    return {p["name"]: pm["owner"][p["start"]:p["start"] + p["nFaces"]] for p in case["patches"]}


def write_time(case, t, U, p, phi=None, owners=None):
    """runTime.write() (icoFoamYade.C:142): <time>/U (or Uc), <time>/p [, <time>/phi is left to the caller's face order].
    fixedValue / noSlip / empty patches are written as they were read; zeroGradient (and fixedFluxPressure) patches get
    the owner cells' values, as OpenFOAM evaluates them."""
    prec = case["control"]["writePrecision"]
    tname = time_name(t)
    owners = owners if owners is not None else boundary_owner_cells(case)
    eU, eP = {}, {}
    for pt in case["patches"]:
        oc = owners[pt["name"]]
        if pt["bcU"] == "empty":
            eU[pt["name"]] = ("empty", None)
            eP[pt["name"]] = ("empty", None)
            continue
        eU[pt["name"]] = ("fixedValue", np.array(pt["valueU"])) if pt["bcU"] == "fixedValue" else ("zeroGradient", None)
        if pt["bcP"] == "fixedValue":
            eP[pt["name"]] = ("fixedValue", pt["valueP"])
        elif pt["bcP"] == "fixedFluxPressure":
            eP[pt["name"]] = ("fixedFluxPressure", np.asarray(p)[oc])
        else:
            eP[pt["name"]] = ("zeroGradient", None)
    write_field(case["case_dir"], tname, case["Uname"], U, (0, 1, -1, 0, 0, 0, 0), case["patches"], eU, prec)
    write_field(case["case_dir"], tname, "p", p, (0, 2, -2, 0, 0, 0, 0), case["patches"], eP, prec)
    return tname


# ---------------------------------------------------------------------------------------------------------------
# a box polyMesh on disk (what `blockMesh` writes for a single block): lets cases be prepared without OpenFOAM
# ---------------------------------------------------------------------------------------------------------------
def write_box_poly_mesh(case_dir, n, L, origin=(0.0, 0.0, 0.0), patches=None, patch_types=None, prec=17):
    """constant/polyMesh/{points,faces,owner,neighbour,boundary} of an nx x ny x nz box in blockMesh's conventions: points
    and cells x fastest, internal faces in upper-triangular order, boundary faces patch by patch, face normals outward
    (owner -> neighbour for internal faces).  patches: [(name, [sides])]; patch_types: {name: 'wall' | 'patch' | 'empty'}."""
    nx, ny, nz = n
    h = np.array(L, dtype=np.float64) / np.array(n)
    px, py, pz = nx + 1, ny + 1, nz + 1
    kk, jj, ii = np.meshgrid(np.arange(pz), np.arange(py), np.arange(px), indexing="ij")
    pts = np.stack([origin[0] + ii.reshape(-1) * h[0], origin[1] + jj.reshape(-1) * h[1], origin[2] + kk.reshape(-1) * h[2]], 1)

    def pid(i, j, k):
        return i + px * (j + py * k)
    N = nx * ny * nz
    c = np.arange(N)
    i, j, k = c % nx, (c // nx) % ny, c // (nx * ny)

    def face(side, i, j, k):
        # quad of the cell's face on `side`, vertices ordered so that the normal points out of the cell
        if side == "xmax":
            return [pid(i + 1, j, k), pid(i + 1, j + 1, k), pid(i + 1, j + 1, k + 1), pid(i + 1, j, k + 1)]
        if side == "xmin":
            return [pid(i, j, k), pid(i, j, k + 1), pid(i, j + 1, k + 1), pid(i, j + 1, k)]
        if side == "ymax":
            return [pid(i, j + 1, k), pid(i, j + 1, k + 1), pid(i + 1, j + 1, k + 1), pid(i + 1, j + 1, k)]
        if side == "ymin":
            return [pid(i, j, k), pid(i + 1, j, k), pid(i + 1, j, k + 1), pid(i, j, k + 1)]
        if side == "zmax":
            return [pid(i, j, k + 1), pid(i + 1, j, k + 1), pid(i + 1, j + 1, k + 1), pid(i, j + 1, k + 1)]
        return [pid(i, j, k), pid(i, j + 1, k), pid(i + 1, j + 1, k), pid(i + 1, j, k)]
    has = np.stack([i < nx - 1, j < ny - 1, k < nz - 1], 1)
    quads = np.stack([np.stack(face(s, i, j, k), 1) for s in ("xmax", "ymax", "zmax")], 1)       # [N][3][4]
    nb = np.stack([c + 1, c + nx, c + nx * ny], 1)
    sel = has.reshape(-1)
    faces = [quads.reshape(-1, 4)[sel]]
    owner = [np.repeat(c[:, None], 3, 1).reshape(-1)[sel]]
    neigh = nb.reshape(-1)[sel]
    on = dict(xmin=i == 0, xmax=i == nx - 1, ymin=j == 0, ymax=j == ny - 1, zmin=k == 0, zmax=k == nz - 1)
    groups = patches if patches is not None else [(s, [s]) for s in SIDES]
    blist, start = [], faces[0].shape[0]
    for name, sides in groups:
        nf = 0
        for s in sides:
            cc = c[on[s]]
            faces.append(np.stack(face(s, i[on[s]], j[on[s]], k[on[s]]), 1))
            owner.append(cc)
            nf += cc.shape[0]
        blist.append((name, (patch_types or {}).get(name, "wall"), nf, start))
        start += nf
    faces, owner = np.concatenate(faces), np.concatenate(owner)
    pm = os.path.join(case_dir, "constant", "polyMesh")
    os.makedirs(pm, exist_ok=True)
    note = "nPoints:%d  nCells:%d  nFaces:%d  nInternalFaces:%d" % (pts.shape[0], N, faces.shape[0], neigh.shape[0])

    def put(name, cls, body, with_note=False):
        with open(os.path.join(pm, name), "w") as f:
            hd = _HEADER % (cls, "constant/polyMesh", name)
            if with_note:
                hd = hd.replace("    location", '    note        "%s";\n    location' % note)
            f.write(hd + body + "\n\n// ************************************************************************* //\n")
    put("points", "vectorField", "%d\n(\n%s\n)\n" % (pts.shape[0], _list(pts, prec)))
    put("faces", "faceList", "%d\n(\n%s\n)\n" % (faces.shape[0], "\n".join("4(%d %d %d %d)" % tuple(q) for q in faces)))
    put("owner", "labelList", "%d\n(\n%s\n)\n" % (owner.shape[0], "\n".join(str(int(v)) for v in owner)), True)
    put("neighbour", "labelList", "%d\n(\n%s\n)\n" % (neigh.shape[0], "\n".join(str(int(v)) for v in neigh)), True)
    body = "%d\n(\n" % len(blist)
    for name, ptype, nf, st in blist:
        body += "    %s\n    {\n        type            %s;\n" % (name, ptype)
        if ptype == "wall":
            body += "        inGroups        1(wall);\n"
        body += "        nFaces          %d;\n        startFace       %d;\n    }\n" % (nf, st)
    put("boundary", "polyBoundaryMesh", body + ")\n")
