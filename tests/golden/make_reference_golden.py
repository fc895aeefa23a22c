"""Regenerates tests/golden/config1_reference.json by running the UNMODIFIED reference Python
(/root/reference/metasbt/core.py: Database.set_configs / add / update / _profile, Entry.sketch / index / get_density /
get_boundaries, Database.dist) on the config-1 stand-in data set, in THIS container.

What is real and what is substituted:
* real     : every line of the reference's own host logic (taxonomy tree, MSBT ids, file layout, report, ANI formula,
             profile descent, boundaries, centroid choice) -- imported from /root/reference, not restated;
* shimmed  : the `howdesbt` executable (absent) -> tests/golden/refshim/fake_howdesbt.py, a thin argv/stdout wrapper over
             the CPU oracle; `fastcluster` and `Bio` (absent, only imported) -> minimal stubs below.
The snapshot is compared with this repo's own Database/Entry/CLI in tests/test_core.py
(test_config1_matches_reference_python): Tier-2 parity of SURVEY.md 8c -- the profile tables and clusters.tsv are what
reference MetaSBT prints given the same filter bits and counts.

    python tests/golden/make_reference_golden.py          (needs /root/reference; run with PYTHONHASHSEED=0)
"""
import json
import os
import stat
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REFERENCE = "/root/reference"


def install_stubs():
    fc = types.ModuleType("fastcluster")
    fc.linkage = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("fastcluster stub"))
    sys.modules["fastcluster"] = fc

    class Seq(str):
        pass

    class SeqRecord:
        def __init__(self, seq, id="", name="", description=""):
            self.seq, self.id, self.name, self.description = seq, id, name, description

    def parse(path, format="fasta"):
        recs, cur = [], None
        with open(path) as f:
            for line in f:
                line = line.rstrip("\n").rstrip("\r")
                if line.startswith(">"):
                    head = line[1:]
                    cur = SeqRecord("", id=head.split()[0] if head.split() else "", name=head.split()[0] if head.split() else "",
                                    description=head)
                    cur._parts = []
                    recs.append(cur)
                elif cur is not None:
                    cur._parts.append(line.strip())
        for r in recs:
            r.seq = Seq("".join(r._parts))
        return iter(recs)

    def write(record, handle, format="fasta"):
        handle.write(">%s\n" % (record.description or record.id))
        s = str(record.seq)
        for i in range(0, len(s), 60):
            handle.write(s[i:i + 60] + "\n")
        return 1

    bio = types.ModuleType("Bio")
    seqio = types.ModuleType("Bio.SeqIO")
    seqio.parse, seqio.write = parse, write
    seqm = types.ModuleType("Bio.Seq")
    seqm.Seq = Seq
    recm = types.ModuleType("Bio.SeqRecord")
    recm.SeqRecord = SeqRecord
    bio.SeqIO, bio.Seq, bio.SeqRecord = seqio, seqm, recm
    sys.modules.update({"Bio": bio, "Bio.SeqIO": seqio, "Bio.Seq": seqm, "Bio.SeqRecord": recm})


def main():
    from oracle import oracle as O
    from tests import dataset, snapshot
    O.build()
    install_stubs()
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
        refs_tsv, mags_txt, _ = dataset.write(os.path.join(workdir, "inputs"))
        references = {}
        for line in open(refs_tsv):
            line = line.strip()
            if line and not line.startswith("#"):
                p, t = line.split("\t")[:2]
                references[p] = t
        # metasbt.py:485-546 (`index`), with the explicit k / filter size path and sorted genomes
        db = ref.Database("db", os.path.join(workdir, "db"), os.path.join(workdir, "tmp"), flat=True, nproc=2)
        genomes = sorted(references)
        db.set_configs(genomes, min_kmer_occurrence=dataset.MIN_OCC, kmer_size=dataset.K, kmer_max=32,
                       filter_size=dataset.FILTER_BITS, filter_expand_by=50.0)
        for g in genomes:
            db.add(g, reference=True, taxonomy=references[g])
        db.update()
        # metasbt.py:1113-1143 (`profile`) and :1231-1242 (`sketch`), CLI defaults uncertainty=20, pruning=0
        db = ref.Database("db", os.path.join(workdir, "db"), os.path.join(workdir, "tmp"), flat=True, nproc=2)
        mags = [l.strip() for l in open(mags_txt) if l.strip()]
        for g in mags:
            name = os.path.splitext(os.path.basename(g))[0]
            sk = ref.Entry(db, name, name, "genome").sketch(g)
            ref.Database._profile(db, g, sk, 20.0, 0.0)
        snap = snapshot.snapshot(workdir)
    with open(os.path.join(HERE, "config1_reference.json"), "w") as f:
        json.dump(snap, f, indent=0, sort_keys=True)
    print("wrote config1_reference.json:", len(snap["filters"]), "filters,", len(snap["profiles"]), "profiles")


if __name__ == "__main__":
    main()
