"""`metasbt index | profile | sketch | update`: the sub-commands of the reference CLI that reach the hot
path (/root/reference/metasbt/metasbt.py: index :319-553, profile :1020-1143, sketch :1145-1244, update :1730-1920),
flag for flag and default for default, over the in-process GPU operators.

What is deliberately different from the reference CLI:
* no `shutil.which` gate on busco/checkm/checkv/kitsune/ntcard/kraken2-build/howdesbt and no GitHub
  release lookup (metasbt.py:100-121): none of those programs is used by this path;
* the per-genome loops (metasbt.py:542-543, 1231-1242) and the `mp.Pool` fan-out (metasbt.py:1124-1143)
  are batch calls: forking after CUDA initialisation is not possible and not needed;
* `--completeness`, `--contamination`, `--pack` are parsed for compatibility and rejected when they would do
  something (they belong to subsystems outside SURVEY.md 8).
"""

from __future__ import annotations

import argparse
import errno
import os
import sys
import time
from typing import List, Optional, Sequence

from .core import Database, Entry, __version__

COMMANDS = ("index", "profile", "sketch", "update")


def _index_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(prog="index", description="Index a set of reference genomes.",
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    general = parser.add_argument_group("General arguments")
    general.add_argument("--workdir", required=True, type=os.path.abspath, help="Path to the working directory.")
    general.add_argument("--database", required=True, type=str, help="The database name.")
    general.add_argument("--references", required=True, type=os.path.abspath,
                         help="Path to the tab-separated-values file with the list of reference genome paths and their "
                              "fully defined taxonomic labels.")
    general.add_argument("--dereplicate", type=float, default=0.0, help="Dereplicate genomes based on their ANI distance.")
    general.add_argument("--nproc", type=int, default=os.cpu_count(), help="Process the input genomes in parallel.")
    general.add_argument("--pack", action="store_true", default=False, help="Pack the database into a compressed tarball.")
    filt = parser.add_argument_group("Estimate a proper bloom filter size")
    filt.add_argument("--filter-size", type=int, required=False, dest="filter_size", help="The bloom filter size.")
    filt.add_argument("--increase-filter-size", type=float, default=50.0, dest="increase_filter_size",
                      help="Increase the estimated filter size by the specified percentage.")
    filt.add_argument("--min-kmer-occurrences", type=int, default=2, dest="min_kmer_occurrences",
                      help="Minimum number of occurrences of kmers to be considered for building the bloom filter files.")
    kmer = parser.add_argument_group("Estimate the optimal kmer size")
    kmer.add_argument("--kmer-size", type=int, required=False, dest="kmer_size", help="The kmer size.")
    kmer.add_argument("--limit-kmer-size", type=int, default=32, dest="limit_kmer_size",
                      help="Limit the estimation of the optimal kmer size with kitsune to this size at most.")
    quality = parser.add_argument_group("Assess the quality of input genomes")
    quality.add_argument("--completeness", type=float, required=False, default=0.0,
                         help="Percentage threshold on genomes completeness.")
    quality.add_argument("--contamination", type=float, required=False, default=100.0,
                         help="Percentage threshold on genomes contamination.")
    return parser


def _genomes_parser(prog: str, argv: Sequence[str], profile: bool) -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(prog=prog, formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("--workdir", required=True, type=os.path.abspath, help="Path to the working directory.")
    parser.add_argument("--database", default="MetaSBT", type=str, help="The database name.")
    parser.add_argument("--genome", required="--genomes" not in argv, type=os.path.abspath, help="Path to the input genome.")
    parser.add_argument("--genomes", required="--genome" not in argv, type=os.path.abspath,
                        help="Path to the file with a list of paths to the input genomes.")
    if profile:
        parser.add_argument("--uncertainty", required=False, type=float, default=20.0,
                            help="Uncertainty percentage for considering multiple best hits.")
        parser.add_argument("--pruning-threshold", dest="pruning_threshold", required=False, type=float, default=0.0,
                            help="Threshold for pruning the Sequence Bloom Tree.")
    parser.add_argument("--nproc", required=False, type=int, default=os.cpu_count(), help="Process the input genomes in parallel.")
    return parser


class MetaSBT(object):
    """Mirror of the reference's `MetaSBT` command object (metasbt.py:55-152) for the hot-path commands."""

    def __init__(self, argv: Optional[Sequence[str]] = None, backend=None) -> None:
        self.database: Optional[Database] = None
        self._backend = backend  # tests inject a checker backend; None -> CudaBackend (no CPU fallback)
        self.start_at = time.time()
        argv = list(sys.argv[1:] if argv is None else argv)
        parser = argparse.ArgumentParser(prog="metasbt", description=f"MetaSBT (B200 hot path) v{__version__}")
        parser.add_argument("command", choices=COMMANDS, help="The sub-command to run.")
        parser.add_argument("-v", "--version", action="version", version=f"metasbt version {__version__}")
        args = parser.parse_args(argv[:1])
        self.result = getattr(self, args.command)(argv[1:])

    # ------------------------------------------------------------------ index (metasbt.py:319-553)
    def index(self, argv: List[str]) -> None:
        args = _index_parser().parse_args(argv)
        os.makedirs(args.workdir, exist_ok=True)
        db_dir = os.path.join(args.workdir, args.database)
        if os.path.isdir(db_dir):
            raise Exception(f"A database named '{args.database}' already exist under {args.workdir}")
        if args.completeness > 0.0 or args.contamination < 100.0:
            raise NotImplementedError("quality control (checkm/checkv/busco) is outside the B200 hot path")
        if args.pack:
            raise NotImplementedError("--pack (tar + sha256sum) is outside the B200 hot path")
        tmp_dir = os.path.join(args.workdir, "tmp")
        self.database = Database(args.database, db_dir, tmp_dir, flat=True, nproc=args.nproc, backend=self._backend)
        references = dict()
        with open(args.references) as input_table:
            for line in input_table:
                line = line.strip()
                if line and not line.startswith("#"):
                    line_split = line.split("\t")
                    references[line_split[0]] = line_split[1]
        # the reference iterates a set of paths (metasbt.py:509): order decided by PYTHONHASHSEED.
        # Sorted order here makes MSBT id assignment reproducible.
        genomes = sorted(references.keys())
        self.database.set_configs(genomes, min_kmer_occurrence=args.min_kmer_occurrences, kmer_size=args.kmer_size,
                                  kmer_max=args.limit_kmer_size, filter_size=args.filter_size,
                                  filter_expand_by=args.increase_filter_size)
        if args.dereplicate > 0.0:
            # metasbt.py:533-535: input genomes against themselves (one all-pairs launch)
            kept = self.database.dereplicate(set(genomes), threshold=args.dereplicate)
            genomes = [genome for genome in genomes if genome in kept]
        self.database.add_many([(genome, True, references[genome]) for genome in genomes])
        self.database.update()

    # ------------------------------------------------------------------ sketch (metasbt.py:1145-1244)
    def _open(self, args) -> List[str]:
        db_dir = os.path.join(args.workdir, args.database)
        if not os.path.isdir(db_dir):
            raise FileNotFoundError(errno.ENOENT, os.strerror(errno.ENOENT), db_dir)
        tmp_dir = os.path.join(args.workdir, "tmp")
        if self.database is None:
            self.database = Database(args.database, db_dir, tmp_dir, flat=True, nproc=args.nproc, backend=self._backend)
        if args.genome:
            return [args.genome]
        return [line.strip() for line in open(args.genomes).readlines() if line.strip()]

    def sketch(self, argv: List[str], parse_known_args: bool = False) -> List[str]:
        parser = _genomes_parser("sketch", argv, profile=False)
        args = parser.parse_known_args(argv)[0] if parse_known_args else parser.parse_args(argv)
        genomes = self._open(args)
        entries = []
        for genome_filepath in genomes:
            genome_name = os.path.splitext(os.path.basename(genome_filepath))[0]
            entries.append(Entry(self.database, genome_name, genome_name, "genome"))
        return Entry.sketch_many(entries, genomes)  # one makebf batch instead of one spawn per genome

    # ------------------------------------------------------------------ profile (metasbt.py:1020-1143)
    def profile(self, argv: List[str], parse_known_args: bool = False):
        parser = _genomes_parser("profile", argv, profile=True)
        args = parser.parse_known_args(argv)[0] if parse_known_args else parser.parse_args(argv)
        genomes = self._open(args)
        sketches = self.sketch(argv, parse_known_args=True)
        # the reference fans `Database._profile` out over a process pool, one job per genome (metasbt.py:1124-1143);
        # here all genomes descend the trees together (Database.profile_many)
        tables = self.database.profile_many(genomes, sketches, uncertainty=args.uncertainty,
                                            pruning_threshold=args.pruning_threshold)
        return list(zip(genomes, tables))


    # ------------------------------------------------------------------ update (metasbt.py:1730-1920)
    def update(self, argv: List[str]):
        """Add new genomes (MAGs) to an existing database: profile them, assign the ones that fall inside a known species
        (is_known), cluster the rest into new taxa (characterize), then re-index every touched cluster (update)."""
        parser = _genomes_parser("update", argv, profile=True)
        parser.description = "Update a MetaSBT database with new genomes."
        for action in parser._actions:
            if action.dest == "database":
                action.required, action.default = True, None
        parser.add_argument("--dereplicate", required=False, default=0.0, type=float, dest="dereplicate",
                            help="Dereplicate genomes based of their ANI distance according the specified threshold.")
        parser.add_argument("--completeness", type=float, required=False, default=0.0,
                            help="Percentage threshold on genomes completeness.")
        parser.add_argument("--contamination", type=float, required=False, default=100.0,
                            help="Percentage threshold on genomes contamination.")
        parser.add_argument("--pack", action="store_true", default=False, help="Pack the database into a compressed tarball.")
        args = parser.parse_args(argv)
        if args.completeness > 0.0 or args.contamination < 100.0:
            raise NotImplementedError("quality control (checkm/checkv/busco) is outside the B200 hot path")
        if args.pack:
            raise NotImplementedError("--pack (tar + sha256sum) is outside the B200 hot path")
        # the reference holds the paths in a set (metasbt.py:1845): sorted here, so runs are reproducible
        genomes = sorted(set(self._open(args)))
        if args.dereplicate > 0.0:
            # metasbt.py:1858-1862: the input genomes against themselves first
            kept = self.database.dereplicate(set(genomes), threshold=args.dereplicate)
            genomes = [genome for genome in genomes if genome in kept]
        # profiles (and sketches) first: is_known / add read the memoised tables (metasbt.py:1866)
        self.profile(argv, parse_known_args=True)
        if args.dereplicate > 0.0:
            # metasbt.py:1870-1872: ... then against the genomes of the database
            kept = self.database.dereplicate(set(genomes), threshold=args.dereplicate, compare_with="database")
            genomes = [genome for genome in genomes if genome in kept]
        species_assignments = self.database.is_known_many(genomes)
        self.database.add_many([(genome, False, species_assignments[genome]) for genome in genomes])
        characterized, unassigned = dict(), set()
        try:
            characterized, unassigned = self.database.characterize()
        except Exception as exc:
            # the reference swallows every exception here (metasbt.py:1904-1910); only "nothing to characterize" is expected
            if "no unknown genomes" not in str(exc):
                raise
        self.database.update()
        return species_assignments, characterized, unassigned


def run(argv: Optional[Sequence[str]] = None) -> None:
    """Console entry point (reference: setup.py:35 `metasbt=metasbt.metasbt:run`, metasbt.py:1923-1931)."""
    t0 = time.time()
    MetaSBT(argv)
    print("Total elapsed time {}s".format(int(time.time() - t0)))


if __name__ == "__main__":
    run()
