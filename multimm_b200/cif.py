"""Structure I/O in the reference's on-disk formats (the .cif outputs are part of the contract).

* init / minimised mmCIF: 13 whitespace-separated atom_site columns, coordinates in Angstrom
  (initial_structure_tools.py:292-358; the minimised file is written by OpenMM's PDBxFile in
  the reference, model.py:890-894 — here by ``write_mmcif`` with the same column layout);
* per-chromosome mmCIF (write_mmcif_chrom, initial_structure_tools.py:417-458);
* reader that keeps ``ATOM`` rows and columns 10-12 (utils.py:168-205), plus a reader that also
  takes ``HETATM`` rows the way PDBxFile does when it loads the start structure (model.py:753);
* PSF (initial_structure_tools.py:461-484).

Vectorised with numpy: the reference's per-bead Python loops cost minutes at 2e5 beads.
"""
from __future__ import annotations

import numpy as np

ATOM_SITE_COLUMNS = ("group_PDB", "id", "type_symbol", "label_atom_id", "label_alt_id", "label_comp_id",
                     "label_asym_id", "label_entity_id", "label_seq_id", "pdbx_PDB_ins_code", "Cartn_x",
                     "Cartn_y", "Cartn_z")
STRUCT_CONN_COLUMNS = ("id", "conn_type_id", "ptnr1_label_comp_id", "ptnr1_label_asym_id", "ptnr1_label_seq_id",
                       "ptnr1_label_atom_id", "ptnr2_label_comp_id", "ptnr2_label_asym_id", "ptnr2_label_seq_id",
                       "ptnr2_label_atom_id")


def atom_header() -> str:
    lines = ["data_MultiMM", "# ", "_entry.id MultiMM", "# ",
             "_audit_conform.dict_name       mmcif_pdbx.dic ", "_audit_conform.dict_version    5.296 ",
             "_audit_conform.dict_location   http://mmcif.pdb.org/dictionaries/ascii/mmcif_pdbx.dic ",
             "# ----------- ATOMS ----------------", "loop_"]
    lines += [f"_atom_site.{c} " for c in ATOM_SITE_COLUMNS[:-1]] + [f"_atom_site.{ATOM_SITE_COLUMNS[-1]}"]
    return "\n".join(lines) + "\n"


def conn_header() -> str:
    return "\n".join(["#", "loop_"] + [f"_struct_conn.{c}" for c in STRUCT_CONN_COLUMNS]) + "\n"


def _chain_index(n: int, chrom_ends: np.ndarray) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """chain_idx, is_end (i in chrom_ends), is_before_end (i in chrom_ends - 1) for i in [0, n)."""
    i = np.arange(n)
    ce = np.asarray(chrom_ends)
    at_end = np.isin(i, ce)
    before = np.isin(i, ce - 1)
    chain = np.searchsorted(ce, i) + at_end  # initial_structure_tools.py:302-304
    return chain, at_end, before


def _fmt(coords: np.ndarray, decimals: int) -> list[np.ndarray]:
    return [np.char.mod(f"%.{decimals}f", coords[:, d]) for d in range(3)]


def _join(cols) -> str:
    out = cols[0]
    for c in cols[1:]:
        out = np.char.add(np.char.add(out, " "), c)
    return "\n".join(out.tolist()) + "\n"


def write_mmcif(coords_angstrom, chrom_ends, path, hetatm_ends=True, connections=True, decimals=3):
    """Whole-model mmCIF.  Beads at chrom_ends[k] and chrom_ends[k]-1 are ALB/CB (HETATM when
    hetatm_ends, as in build_init_mmcif; plain ATOM as in write_mmcif), the rest ALA/CA; the chain
    letter is chr(65 + chain_idx).  One %-format per line over pre-built Python lists: ~1.5 us per
    bead (the reference concatenates strings bead by bead in O(N^2))."""
    xyz = np.asarray(coords_angstrom, dtype=np.float64)
    n = len(xyz)
    chain, at_end, before = _chain_index(n, chrom_ends)
    special = at_end | before
    letters = [chr(65 + int(c)) for c in range(int(chain.max()) + 1)] if n else []
    chain_l = chain.tolist()
    letter_l = [letters[c] for c in chain_l]
    ids = range(1, n + 1)
    if special.any():
        sp = special.tolist()
        group = ["HETATM" if (s and hetatm_ends) else "ATOM" for s in sp]
        atom = ["CB" if s else "CA" for s in sp]
        res = ["ALB" if s else "ALA" for s in sp]
    else:
        group, atom, res = ["ATOM"] * n, ["CA"] * n, ["ALA"] * n
    fmt = f"%s %d D %s . %s %s %d %d ? %.{decimals}f %.{decimals}f %.{decimals}f\n"
    x, y, z = xyz[:, 0].tolist(), xyz[:, 1].tolist(), xyz[:, 2].tolist()
    parts = [atom_header()]
    parts.extend([fmt % t for t in zip(group, ids, atom, res, letter_l, chain_l, ids, x, y, z)])
    if connections and n > 1:
        keep = ~before[:-1]  # no connection out of the bead just before a chromosome end
        idx = np.nonzero(keep)[0].tolist()
        at_end_l, before_l = at_end.tolist(), before.tolist()
        parts.append("\n" + conn_header())
        parts.extend(["D%d covale %s %s %d %s %s %s %d %s\n" % (
            i + 1, "ALB" if at_end_l[i] else "ALA", letter_l[i], i + 1, "CB" if at_end_l[i] else "CA",
            "ALB" if before_l[i + 1] else "ALA", letter_l[i], i + 2, "CB" if before_l[i + 1] else "CA") for i in idx])
    with open(path, "w") as f:
        f.write("".join(parts))


def write_mmcif_chrom(coords_angstrom, path, decimals=3):
    """One chromosome: chain A, entity 1, first and last bead ALB (initial_structure_tools.py:417-458)."""
    xyz = np.asarray(coords_angstrom, dtype=np.float64)
    n = len(xyz)
    res = ["ALA"] * n
    if n:
        res[0] = res[-1] = "ALB"
    fmt = f"ATOM %d D CA . %s A 1 %d ? %.{decimals}f %.{decimals}f %.{decimals}f\n"
    ids = range(1, n + 1)
    parts = [atom_header()]
    parts.extend([fmt % t for t in zip(ids, res, ids, xyz[:, 0].tolist(), xyz[:, 1].tolist(), xyz[:, 2].tolist())])
    if n > 1:
        parts.append(conn_header())
        parts.extend(["D%d covale %s A %d CA %s A %d CA\n" % (i + 1, res[i], i + 1, res[i + 1], i + 2)
                      for i in range(n - 1)])
    with open(path, "w") as f:
        f.write("".join(parts))


def read_cif_coordinates(path, include_hetatm=False) -> np.ndarray:
    """(N,3) coordinates as written (Angstrom).  include_hetatm=False is get_coordinates_cif
    (utils.py:168-205: ATOM rows only, columns 10-12); True is what PDBxFile sees."""
    rows = []
    with open(path) as f:
        for line in f:
            if line.startswith("ATOM") or (include_hetatm and line.startswith("HETATM")):
                c = line.split()
                try:
                    rows.append((float(c[10]), float(c[11]), float(c[12])))
                except (IndexError, ValueError):
                    continue
    return np.array(rows, dtype=np.float64).reshape(-1, 3)


def write_psf(n: int, path, title="No title provided"):
    assert len(title) < 40, "provided title in psf file is too long."
    atoms = ["%8d BEAD %-5d ALA  CA   A      0.000000        1.00 0           0\n" % (a, a) for a in range(1, n + 1)]
    bonds = ["%8d%8d\n" % (a, a + 1) for a in range(1, n)]
    with open(path, "w") as f:
        f.writelines(["PSF CMAP\n", "\n", "      1 !NTITLE\n", f"REMARKS {title}\n", "\n", f"{n:>8} !NATOM\n"])
        f.writelines(atoms)
        f.writelines(["\n", f"{n - 1:>8} !NBOND: bonds\n"])
        f.writelines(bonds)


class DCDWriter:
    """Minimal CHARMM/NAMD DCD trajectory writer (what OpenMM's DCDReporter produces for
    model.py:920-925): little-endian, no unit cell, coordinates in Angstrom as float32."""

    def __init__(self, path, n_atoms: int, dt_ps: float, interval: int, first_step: int = 0):
        self.fh = open(path, "wb")
        self.n_atoms, self.frames = int(n_atoms), 0
        icntrl = np.zeros(20, dtype="<i4")
        icntrl[1], icntrl[2] = first_step, interval
        icntrl[9] = np.array([dt_ps / 0.04888821], dtype="<f4").view("<i4")[0]  # AKMA time units
        icntrl[19] = 24
        title = b"Created by multimm_b200".ljust(80)
        self.fh.write(np.array([84], "<i4").tobytes() + b"CORD" + icntrl.tobytes() + np.array([84], "<i4").tobytes())
        self.fh.write(np.array([84, 1], "<i4").tobytes() + title + np.array([84], "<i4").tobytes())
        self.fh.write(np.array([4, self.n_atoms, 4], "<i4").tobytes())

    def write(self, coords_angstrom):
        c = np.asarray(coords_angstrom, dtype="<f4")
        if c.shape != (self.n_atoms, 3):
            raise ValueError("DCD frame has the wrong shape")
        nb = np.array([4 * self.n_atoms], "<i4").tobytes()
        for d in range(3):
            self.fh.write(nb + np.ascontiguousarray(c[:, d]).tobytes() + nb)
        self.frames += 1
        pos = self.fh.tell()
        self.fh.seek(8)  # NSET, first control word after 'CORD'
        self.fh.write(np.array([self.frames], "<i4").tobytes())
        self.fh.seek(pos)

    def close(self):
        self.fh.close()


def read_dcd(path):
    """Frames of a DCD written by DCDWriter -> (n_frames, n_atoms, 3) float32 (tests)."""
    raw = np.fromfile(path, dtype=np.uint8)
    n_frames = int(raw[8:12].view("<i4")[0])
    off = 4 + 84 + 4
    ntitle_block = int(raw[off:off + 4].view("<i4")[0])
    off += 4 + ntitle_block + 4
    n_atoms = int(raw[off + 4:off + 8].view("<i4")[0])
    off += 12
    out = np.empty((n_frames, n_atoms, 3), dtype=np.float32)
    for f in range(n_frames):
        for d in range(3):
            out[f, :, d] = raw[off + 4:off + 4 + 4 * n_atoms].view("<f4")
            off += 8 + 4 * n_atoms
    return out
