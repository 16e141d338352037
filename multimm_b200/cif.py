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
    letter is chr(65 + chain_idx)."""
    xyz = np.asarray(coords_angstrom, dtype=np.float64)
    n = len(xyz)
    chain, at_end, before = _chain_index(n, chrom_ends)
    special = at_end | before
    ids = np.char.mod("%d", np.arange(1, n + 1))
    group = np.where(special & hetatm_ends, "HETATM", "ATOM")
    atom = np.where(special, "CB", "CA")
    res = np.where(special, "ALB", "ALA")
    letter = np.array([chr(65 + int(c)) for c in range(int(chain.max()) + 1)])[chain]
    fx, fy, fz = _fmt(xyz, decimals)
    n_like = np.full(n, "D")
    text = atom_header() + _join([group, ids, n_like, atom, np.full(n, "."), res, letter,
                                  np.char.mod("%d", chain), ids, np.full(n, "?"), fx, fy, fz])
    if connections and n > 1:
        i = np.arange(n - 1)
        keep = ~before[:-1]  # no connection out of the bead just before a chromosome end
        i = i[keep]
        res1 = np.where(at_end[i], "ALB", "ALA")
        at1 = np.where(at_end[i], "CB", "CA")
        res2 = np.where(before[i + 1], "ALB", "ALA")
        at2 = np.where(before[i + 1], "CB", "CA")
        cl = letter[i]
        text += "\n" + conn_header() + _join([
            np.char.add("D", np.char.mod("%d", i + 1)), np.full(len(i), "covale"), res1, cl,
            np.char.mod("%d", i + 1), at1, res2, cl, np.char.mod("%d", i + 2), at2])
    with open(path, "w") as f:
        f.write(text)


def write_mmcif_chrom(coords_angstrom, path, decimals=3):
    """One chromosome: chain A, entity 1, first and last bead ALB (initial_structure_tools.py:417-458)."""
    xyz = np.asarray(coords_angstrom, dtype=np.float64)
    n = len(xyz)
    ids = np.char.mod("%d", np.arange(1, n + 1))
    edge = (np.arange(n) == 0) | (np.arange(n) == n - 1)
    res = np.where(edge, "ALB", "ALA")
    fx, fy, fz = _fmt(xyz, decimals)
    text = atom_header() + _join([np.full(n, "ATOM"), ids, np.full(n, "D"), np.full(n, "CA"), np.full(n, "."), res,
                                  np.full(n, "A"), np.full(n, "1"), ids, np.full(n, "?"), fx, fy, fz])
    if n > 1:
        i = np.arange(n - 1)
        text += conn_header() + _join([
            np.char.add("D", np.char.mod("%d", i + 1)), np.full(n - 1, "covale"), res[:-1], np.full(n - 1, "A"),
            np.char.mod("%d", i + 1), np.full(n - 1, "CA"), res[1:], np.full(n - 1, "A"),
            np.char.mod("%d", i + 2), np.full(n - 1, "CA")])
    with open(path, "w") as f:
        f.write(text)


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
    k = np.arange(1, n + 1)
    atoms = [f"{a:>8} BEAD {a:<5} ALA  CA   A      0.000000        1.00 0           0\n" for a in k]
    bonds = [f"{a:>8}{a + 1:>8}\n" for a in k[:-1]]
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
