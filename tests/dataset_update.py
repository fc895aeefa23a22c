"""TEST INFRASTRUCTURE: a synthetic data set for the `metasbt update` path (is_known, characterize,
_estimate_boundaries).  Unlike the config-1 stand-in (whose species are unrelated, as the reference's viral test set is)
it needs RELATED taxa, so that clusters above the species level have finite ANI boundaries:

    root ancestor R (80 kbp)
      genus ancestor   = R mutated 5 %           (2 genera under one family)
      species ancestor = genus ancestor mutated 5 %   (3 species per genus)
      genome           = species ancestor mutated 1 % (4 / 3 / 3 genomes per species)

which gives ANI distances of about 0.02 inside a species, 0.13 between species of a genus and 0.25 between genera.
The six MAGs exercise every branch of the update: one inside a known species, three forming a new species of a known
genus, one forming a new genus of the known family, one unrelated (stays unassigned).  Deterministic (fixed seeds): the
golden snapshot tests/golden/update_reference.json, produced by the unmodified reference Python, stays valid.
"""
import os
from typing import Dict, List, Tuple

from tests import util

K, FILTER_BITS, MIN_OCC = 31, 1 << 22, 1
LENGTH = 80_000
LINEAGE = "k__Synthia|p__Synthota|c__Synthia_c|o__Synthales|f__Synthaceae"
GENERA = [("g__Alphasynth", 11), ("g__Betasynth", 12)]
COUNTS = [4, 3, 3]  # genomes per species, three species per genus


def _root():
    return util.synthetic_codes(0xD0C0FFEE, LENGTH)


def _genus(gi: int, rate: float = 0.05):
    return util.mutate(_root(), rate, 100 + gi)


def _species(gi: int, si: int):
    return util.mutate(_genus(gi), 0.05, 200 + 10 * gi + si)


def references() -> List[Tuple[str, str, bytes]]:
    out = []
    n = 0
    for gi, (genus, _) in enumerate(GENERA):
        for si, count in enumerate(COUNTS):
            label = "%s|%s|s__%s_sp%d" % (LINEAGE, genus, genus[3:], si + 1)
            anc = _species(gi, si)
            for c in range(count):
                name = "SYN_%03d_ref" % n
                codes = util.mutate(anc, 0.01, 1000 + n)
                out.append((name, label, util.codes_to_fasta(codes, name + "_contig1", 70)))
                n += 1
    return out


def mags() -> List[Tuple[str, bytes]]:
    out = []

    def put(name, codes, contigs=1):
        cut = len(codes) // contigs
        body = b"".join(util.codes_to_fasta(codes[i * cut:(i + 1) * cut if i + 1 < contigs else len(codes)],
                                            "%s_contig%d" % (name, i + 1), 60) for i in range(contigs))
        out.append((name, body))

    # a: 1 % away from the first genome of the first species -> inside a known species
    first = util.mutate(_species(0, 0), 0.01, 1000)
    put("MAG_a_known", util.mutate(first, 0.01, 7001))
    # b, c, d: a new species of genus Alphasynth (3 % from the genus ancestor), three MAGs 1 % from it
    new_species = util.mutate(_genus(0), 0.03, 7100)
    put("MAG_b_newsp", util.mutate(new_species, 0.01, 7101), contigs=2)
    put("MAG_c_newsp", util.mutate(new_species, 0.01, 7102))
    put("MAG_d_newsp", util.mutate(new_species, 0.01, 7103), contigs=3)
    # e: a new genus of the family (3 % from the root, then 3 %)
    new_genus = util.mutate(util.mutate(_root(), 0.03, 7200), 0.03, 7201)
    put("MAG_e_newgen", util.mutate(new_genus, 0.01, 7202))
    # f: unrelated
    put("MAG_f_alien", util.synthetic_codes(0xBADC0DE, 30_000), contigs=2)
    return out


def write(folder: str) -> Tuple[str, str, Dict[str, str]]:
    os.makedirs(os.path.join(folder, "genomes"), exist_ok=True)
    paths = {}
    with open(os.path.join(folder, "references.tsv"), "w") as tsv:
        tsv.write("# path\ttaxonomy\n")
        for name, label, fasta in references():
            p = os.path.join(folder, "genomes", name + ".fna")
            with open(p, "wb") as f:
                f.write(fasta)
            paths[name] = p
            tsv.write("%s\t%s\n" % (p, label))
    with open(os.path.join(folder, "mags.txt"), "w") as txt:
        for name, fasta in mags():
            p = os.path.join(folder, "genomes", name + ".fna")
            with open(p, "wb") as f:
                f.write(fasta)
            paths[name] = p
            txt.write(p + "\n")
    return os.path.join(folder, "references.tsv"), os.path.join(folder, "mags.txt"), paths
