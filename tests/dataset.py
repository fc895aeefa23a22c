"""TEST INFRASTRUCTURE: the synthetic stand-in for BASELINE.json configs[0] (SURVEY.md 8d "Config 1").

The reference's own test downloads 12 viral references and 6 MAGs from NCBI
(/root/reference/test/references.tsv:3-14, test/mags.txt:3-8); there is no network here, so the same
SHAPE is generated from fixed seeds: 4 species x {4, 2, 2, 4} genomes under the 7-level labels copied
from references.tsv (8 poxvirus-like genomes of ~197 kbp, 4 RSV-like of ~15 kbp, each genome a 1-3 %
mutation of its species ancestor, multi-contig, with an N run) and 6 MAGs (5 mutated 2-10 % from
references, 1 unrelated).  Everything is deterministic, so the committed golden files under
tests/golden/config1/ stay valid.
"""
import os
from typing import Dict, List, Tuple

import numpy as np

from tests import util

POX = "k__Viruses|p__Nucleocytoviricota|c__Pokkesviricetes|o__Chitovirales|f__Poxviridae|g__Orthopoxvirus|s__"
RSV = "k__Viruses|p__Negarnaviricota|c__Monjiviricetes|o__Mononegavirales|f__Pneumoviridae|g__Orthopneumovirus|s__"
SPECIES = [  # (label, ancestor seed, length, genomes)
    (POX + "Monkeypox_virus", 0xA11CE000, 197_000, 4),
    (POX + "Skunkpox_virus", 0xA11CE001, 199_500, 2),
    (POX + "Raccoonpox_virus", 0xA11CE002, 196_200, 2),
    (RSV + "Human_orthopneumovirus", 0xA11CE003, 15_200, 4),
]
K, FILTER_BITS, MIN_OCC = 31, 1 << 22, 1


def _contigs(codes: np.ndarray, name: str, n_contigs: int, n_run: bool) -> bytes:
    """FASTA with `n_contigs` records; the first record gets a run of 25 N in the middle."""
    cuts = np.linspace(0, len(codes), n_contigs + 1).astype(int)
    out = []
    for c in range(n_contigs):
        body = util.codes_to_fasta(codes[cuts[c]:cuts[c + 1]], "%s_contig%d" % (name, c + 1), 70)
        if c == 0 and n_run:
            lines = body.split(b"\n")
            lines.insert(len(lines) // 2, b"N" * 25)
            body = b"\n".join(lines)
        out.append(body)
    return b"".join(out)


def references() -> List[Tuple[str, str, bytes]]:
    """[(genome name, taxonomy, FASTA bytes)]"""
    out = []
    g = 0
    for label, seed, length, count in SPECIES:
        ancestor = util.synthetic_codes(seed, length)
        for i in range(count):
            codes = util.mutate(ancestor, 0.01 * (1 + i % 3), 1000 + g)
            name = "GCA_%09d.1_ref%d" % (24732825 + 7919 * g, g)
            out.append((name, label, _contigs(codes, name, 1 + g % 3, g % 2 == 0)))
            g += 1
    return out


def mags() -> List[Tuple[str, bytes]]:
    refs = references()
    out = []
    picks = [(0, 0.02), (5, 0.04), (7, 0.06), (9, 0.10), (11, 0.03)]
    for m, (src, rate) in enumerate(picks):
        # recover codes of the source by re-generating it (cheap) -- mutate its ancestor the same way, then again
        label, seed, length, count = next(s for s in SPECIES if s[0] == refs[src][1])
        first = sum(s[3] for s in SPECIES[:[s[0] for s in SPECIES].index(label)])
        i = src - first
        codes = util.mutate(util.synthetic_codes(seed, length), 0.01 * (1 + i % 3), 1000 + src)
        codes = util.mutate(codes, rate, 5000 + m)
        name = "GCA_%09d.1_mag%d" % (11395005 + 104729 * m, m)
        out.append((name, _contigs(codes, name, 2 + m % 2, m % 2 == 1)))
    name = "GCA_002817435.1_mag5"
    out.append((name, _contigs(util.synthetic_codes(0xBADC0DE, 30_000), name, 2, False)))
    return out


def write(folder: str) -> Tuple[str, str, Dict[str, str]]:
    """Write FASTA files + references.tsv + mags.txt under `folder` -> (references.tsv, mags.txt, {name: path})."""
    os.makedirs(os.path.join(folder, "genomes"), exist_ok=True)
    paths = {}
    with open(os.path.join(folder, "references.tsv"), "w") as tsv:
        tsv.write("# synthetic stand-in for test/references.tsv\n# path\ttaxonomy\n")
        for name, label, fasta in references():
            p = os.path.join(folder, "genomes", name + ".fna")
            with open(p, "wb") as f:
                f.write(fasta)
            paths[name] = p
            tsv.write("%s\t%s\n" % (p, label))
    with open(os.path.join(folder, "mags.txt"), "w") as txt:
        txt.write("")
        for name, fasta in mags():
            p = os.path.join(folder, "genomes", name + ".fna")
            with open(p, "wb") as f:
                f.write(fasta)
            paths[name] = p
            txt.write(p + "\n")
    return os.path.join(folder, "references.tsv"), os.path.join(folder, "mags.txt"), paths
