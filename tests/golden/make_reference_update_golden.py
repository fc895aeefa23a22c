"""Regenerates tests/golden/update_reference.json by running the UNMODIFIED reference Python
(/root/reference/metasbt/core.py: Database.is_known / add / characterize / _estimate_boundaries / update, on top of
the index + profile path) on the related-taxa data set of tests/dataset_update.py, in THIS container.

The calls follow the reference's own `update` command (metasbt.py:1730-1920) step by step: profile every MAG with the
CLI defaults (uncertainty 20, pruning 0), Database._is_known per MAG, Database.add with the answer, characterize(),
update().  Substitutions are those of make_reference_golden.py (a stand-in `howdesbt` served by the CPU oracle; `Bio`
stubbed) plus `fastcluster.linkage` -> `scipy.cluster.hierarchy.linkage` (fastcluster is not installed; it implements
the same average-linkage algorithm and the same output format).  MAGs are processed in sorted order (the reference
iterates a set).

    PYTHONHASHSEED=0 python tests/golden/make_reference_update_golden.py
"""
import json
import os
import stat
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
REFERENCE = "/root/reference"


def main():
    import scipy.cluster.hierarchy as hier

    import make_reference_golden as base
    from oracle import oracle as O
    from tests import dataset_update as D
    from tests import snapshot
    O.build()
    base.install_stubs()
    sys.modules["fastcluster"].linkage = hier.linkage
    sys.path.insert(0, REFERENCE)
    import metasbt.core as ref  # the unmodified reference

    with tempfile.TemporaryDirectory() as tmp:
        bindir = os.path.join(tmp, "bin")
        os.makedirs(bindir)
        exe = os.path.join(bindir, "howdesbt")
        with open(exe, "w") as f:
            f.write("#!/bin/sh\nexec %s %s \"$@\"\n" % (sys.executable, os.path.join(HERE, "refshim", "fake_howdesbt.py")))
        os.chmod(exe, os.stat(exe).st_mode | stat.S_IEXEC)
        os.environ["PATH"] = bindir + os.pathsep + os.environ["PATH"]
        os.environ["MSBT_REPO_ROOT"] = ROOT

        workdir = os.path.join(tmp, "w")
        os.makedirs(workdir)
        refs_tsv, mags_txt, _ = D.write(os.path.join(workdir, "inputs"))
        references = {}
        for line in open(refs_tsv):
            line = line.strip()
            if line and not line.startswith("#"):
                p, t = line.split("\t")[:2]
                references[p] = t
        db = ref.Database("db", os.path.join(workdir, "db"), os.path.join(workdir, "tmp"), flat=True, nproc=2)
        genomes = sorted(references)
        db.set_configs(genomes, min_kmer_occurrence=D.MIN_OCC, kmer_size=D.K, kmer_max=32, filter_size=D.FILTER_BITS,
                       filter_expand_by=50.0)
        for g in genomes:
            db.add(g, reference=True, taxonomy=references[g])
        db.update()

        # metasbt.py:1836-1920
        db = ref.Database("db", os.path.join(workdir, "db"), os.path.join(workdir, "tmp"), flat=True, nproc=2)
        mags = sorted(l.strip() for l in open(mags_txt) if l.strip())
        for g in mags:
            name = os.path.splitext(os.path.basename(g))[0]
            sk = ref.Entry(db, name, name, "genome").sketch(g)
            ref.Database._profile(db, g, sk, 20.0, 0.0)
        species_assignments = {}
        for g in mags:
            path, taxonomy = ref.Database._is_known(db, g)
            species_assignments[path] = taxonomy
        for g in species_assignments:
            db.add(g, reference=False, taxonomy=species_assignments[g])
        characterized, unassigned = db.characterize()
        db.update()
        snap = snapshot.snapshot(workdir)
        snap["update"] = {
            "species_assignments": {os.path.basename(k): v for k, v in species_assignments.items()},
            "characterized": {k: sorted(v) for k, v in characterized.items()},
            "unassigned": sorted(unassigned)}
    with open(os.path.join(HERE, "update_reference.json"), "w") as f:
        json.dump(snap, f, indent=0, sort_keys=True)
    print("wrote update_reference.json:", len(snap["filters"]), "filters,", len(snap["profiles"]), "profiles")
    print(json.dumps(snap["update"], indent=1))


if __name__ == "__main__":
    main()
