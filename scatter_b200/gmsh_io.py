"""gmsh 2.2 ASCII reader / writer (the only mesh format the reference accepts, `scatter/mesher.py:55-228`).

Layout (all shipped meshes are `2.2 0 8` with exactly two tags per element):

    $MeshFormat / 2.2 0 8 / $EndMeshFormat
    $PhysicalNames / n / <dim> <tag> "<name>" ... / $EndPhysicalNames
    $Nodes / Nn / <id> <x> <y> <z> ... / $EndNodes
    $Elements / Ne / <id> <type> 2 <phys-tag> <geom-tag> <node ids...> ... / $EndElements
"""
from __future__ import annotations

import numpy as np

GMSH_TO_TYPE = {2: "tri3", 9: "tri6", 3: "quad4", 5: "hexa8", 17: "hexa20", 4: "tetra4", 11: "tetra10"}
TYPE_TO_GMSH = {v: k for k, v in GMSH_TO_TYPE.items()}
TYPE_TO_GMSH["quad8"] = 16      # not accepted as a domain element by the reference (mesher.py:181); writer only


# nodes per gmsh element type (all types gmsh 2.2 can write for points, lines, triangles, quads, tets, hexes, prisms, pyramids)
_GMSH_NNODES = {1: 2, 2: 3, 3: 4, 4: 4, 5: 8, 6: 6, 7: 5, 8: 3, 9: 6, 10: 9, 11: 10, 12: 27, 13: 18, 14: 14, 15: 1, 16: 8, 17: 20,
                18: 15, 19: 13}


def _section_text(text: str, start: str, end: str):
    i0 = text.find(start)
    i1 = text.find(end)
    if i0 < 0 or i1 < 0:
        return None
    return text[text.index("\n", i0) + 1:i1]


def read_msh(path: str) -> dict:
    """Whole-array reader: every section is tokenised in one call and the (ragged) element records are decoded block by
    block -- a run of records of the same type and tag count is one reshape -- so a 16 M-element file costs seconds, not
    the minute a per-line loop takes."""
    with open(path, "r") as f:
        text = f.read()
    # physical names: [dim(float), tag(int), name(str)] like utils.search_idx + mesher.py:146-147
    names = []
    sec = _section_text(text, "$PhysicalNames", "$EndPhysicalNames")
    if sec is not None:
        for l in sec.splitlines()[1:]:
            t = l.split()
            if t:
                names.append([float(t[0]), int(float(t[1])), " ".join(t[2:]).replace('"', "")])
    sec = _section_text(text, "$Nodes", "$EndNodes")
    if sec is None:
        raise SystemExit("ERROR: the mesh file has no $Nodes section")
    flat = np.fromstring(sec, sep=" ")
    nodes = flat[1:].reshape(-1, 4)
    sec = _section_text(text, "$Elements", "$EndElements")
    if sec is None:
        raise SystemExit("ERROR: the mesh file has no $Elements section")
    flat = np.fromstring(sec, sep=" ", dtype=np.int64)[1:]        # records: id type ntags tags... node ids...
    merged = []
    p, n = 0, len(flat)
    while p < n:
        et, ntags = int(flat[p + 1]), int(flat[p + 2])
        if ntags != 2:
            raise SystemExit("ERROR: gmsh elements must carry exactly 2 tags")
        if et not in _GMSH_NNODES:
            raise SystemExit(f"ERROR: gmsh element type {et} not supported")
        L = 3 + ntags + _GMSH_NNODES[et]
        rows = flat[p:p + ((n - p) // L) * L].reshape(-1, L)
        same = (rows[:, 1] == et) & (rows[:, 2] == ntags)
        k = len(rows) if same.all() else int(np.argmin(same))     # records up to the first one of another kind
        blk = rows[:k]
        # consecutive blocks of equal type are one block (ids are irrelevant for the reference, order is)
        if merged and merged[-1][0] == et:
            merged[-1] = (et, np.concatenate([merged[-1][1], blk[:, 3]]), np.vstack([merged[-1][2], blk[:, 3 + ntags:]]))
        else:
            merged.append((et, blk[:, 3].copy(), blk[:, 3 + ntags:].copy()))
        p += k * L
    return {"physical_names": names, "nodes": nodes, "elements": merged}


def write_msh(path: str, nodes: np.ndarray, elem: np.ndarray, tags: np.ndarray, physical_names: list, element_type: str):
    """nodes (Nn,4) [id,x,y,z]; elem (Ne,nne) node ids; tags (Ne,) physical tag; physical_names [[dim, tag, name],...]."""
    code = TYPE_TO_GMSH[element_type]
    with open(path, "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n")
        f.write(f"$PhysicalNames\n{len(physical_names)}\n")
        for d, t, n in physical_names:
            f.write(f'{int(d)} {int(t)} "{n}"\n')
        f.write("$EndPhysicalNames\n")
        f.write(f"$Nodes\n{len(nodes)}\n")
        for r in nodes:
            f.write(f"{int(r[0])} {float(r[1])!r} {float(r[2])!r} {float(r[3])!r}\n")
        f.write("$EndNodes\n")
        f.write(f"$Elements\n{len(elem)}\n")
        for i, (e, t) in enumerate(zip(elem, tags)):
            f.write(f"{i + 1} {code} 2 {int(t)} {int(t)} " + " ".join(str(int(n)) for n in e) + "\n")
        f.write("$EndElements\n")
