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


def _section(lines, start, end):
    i0 = next(i for i, l in enumerate(lines) if l.startswith(start))
    i1 = next(i for i, l in enumerate(lines) if l.startswith(end))
    return i0, i1


def read_msh(path: str) -> dict:
    with open(path, "r") as f:
        lines = f.readlines()
    # physical names: [dim(float), tag(int), name(str)] like utils.search_idx + mesher.py:146-147
    names = []
    try:
        i0, i1 = _section(lines, "$PhysicalNames", "$EndPhysicalNames")
        for l in lines[i0 + 2:i1]:
            t = l.split()
            names.append([float(t[0]), int(float(t[1])), " ".join(t[2:]).replace('"', "")])
    except StopIteration:
        pass
    i0, i1 = _section(lines, "$Nodes", "$EndNodes")
    nodes = np.array(" ".join(lines[i0 + 2:i1]).split(), dtype=float).reshape(-1, 4)
    i0, i1 = _section(lines, "$Elements", "$EndElements")
    rows = [np.array(l.split(), dtype=np.int64) for l in lines[i0 + 2:i1] if l.strip()]
    # consecutive rows of the same gmsh type form one block (file order is kept)
    blocks = []
    k = 0
    while k < len(rows):
        j = k
        while j < len(rows) and rows[j][1] == rows[k][1] and len(rows[j]) == len(rows[k]):
            j += 1
        blk = np.vstack(rows[k:j])
        ntags = int(blk[0, 2])
        if ntags != 2:
            raise SystemExit("ERROR: gmsh elements must carry exactly 2 tags")
        blocks.append((int(blk[0, 1]), blk[:, 3].copy(), blk[:, 3 + ntags:].copy()))
        k = j
    # merge blocks of equal type (ids are irrelevant for the reference, order is)
    merged = []
    for b in blocks:
        if merged and merged[-1][0] == b[0]:
            merged[-1] = (b[0], np.concatenate([merged[-1][1], b[1]]), np.vstack([merged[-1][2], b[2]]))
        else:
            merged.append(b)
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
