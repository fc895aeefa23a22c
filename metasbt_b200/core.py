"""Database / Entry: the `metasbt.core` API surface that sits on the hot path, re-implemented over
in-process GPU operators (backend.CudaBackend -> libmsbt.so) instead of `howdesbt` subprocesses.

Scope: what `metasbt index` and `metasbt profile` reach -- Database.{set_configs, add,
update, dist, profile}, Entry.{sketch, index, get_density, get_boundaries, get_children,
get_full_taxonomy, is_known, add_children} -- with the same signatures, exceptions, on-disk layout
and result files (`clusters.tsv`, `genomes.tsv`, `metadata.json`, `<name>.txt`,
`tree/index.detbrief.sbt`, `tmp/profiles/<genome>.txt`) as /root/reference/metasbt/core.py.
flat=False databases get their tree topology from cluster.py (`howdesbt cluster`: greedy Hamming agglomeration over
K3 / K2) with plain union nodes; the determined/brief RRR node files of `howdesbt build --howde` are out of scope.
Out of scope and raising NotImplementedError: qc/merge/remove.  The k-mer size search (kitsune, core.py:439-468) is restated in kopt.py.
The `update` consumers of the same kernels (is_known, characterize, _estimate_boundaries) are included.

Differences that are deliberate (documented in DESIGN.md):
* process fan-outs (mp.Pool over dist/_profile, core.py:4036, metasbt.py:1124) are batch calls;
* wherever the reference iterates a Python `set` (order depends on PYTHONHASHSEED: core.py:3746,
  3979; metasbt.py:509) this code iterates in sorted order, so runs are reproducible;
* Entry.get_boundaries uses Database.tmp in the single-process branch (the reference passes the
  non-existent `self.tmp`, core.py:4072, and would raise AttributeError).
"""

from __future__ import annotations

import datetime
import errno
import json
import math
import os
import random
import re
import shutil
import statistics
import sys
from collections import OrderedDict
from pathlib import Path
from typing import Any, Dict, Iterable, List, Optional, Sequence, Set, Tuple

__version__ = "0.1.5-b200"

LEVELS = ["kingdom", "phylum", "class", "order", "family", "genus", "species"]
_MSBT_ID = re.compile(r"^MSBT[0-9]*$")

# backends shared by the static Database.dist (keyed by (kmer_size, num_bits))
_BACKENDS: Dict[Tuple[int, int], Any] = {}


def get_backend(kmer_size: int, num_bits: int):
    key = (int(kmer_size), int(num_bits))
    if key not in _BACKENDS:
        from .backend import CudaBackend  # fails loudly without libmsbt.so / a CUDA device
        from .multigpu import MultiGpuBackend, devices_from_env
        devices = devices_from_env()      # MSBT_DEVICES="0,1,..." | "all": one context, cache and host thread per GPU
        _BACKENDS[key] = MultiGpuBackend(*key, devices=devices) if devices else CudaBackend(*key)
    return _BACKENDS[key]


def register_backend(kmer_size: int, num_bits: int, backend) -> None:
    """Install a backend instance for (k, bits) -- used by tests to inject a checker backend."""
    _BACKENDS[(int(kmer_size), int(num_bits))] = backend


def ani_distance(intersect: int, union: int, kmer_size: int) -> float:
    """Jaccard (rounded to 5 decimals) -> ANI distance, exactly as core.py:1816-1829."""
    jaccard_index = round(intersect / union, 5)
    if jaccard_index == 0.0:
        return 1.0
    ani = 1 - (1 + (1 / kmer_size) * math.log((2 * jaccard_index) / (1 + jaccard_index)))
    return 1.0 if ani > 1.0 else ani


def mean_rounded(values: Sequence[float], digits: int = 5) -> float:
    """round(statistics.mean(values), digits), as core.py:2185-2188 computes it, without paying for exact rational
    arithmetic on every call: statistics.mean is the exactly rounded mean; fsum(values) / n is within one ulp of it, so
    the two can only round differently when a `digits` rounding boundary lies within a few ulps -- checked, and only
    then is statistics.mean consulted."""
    n = len(values)
    if n == 0:
        raise statistics.StatisticsError("mean requires at least one data point")
    approx = math.fsum(values) / n
    spread = 4.0 * math.ulp(approx)
    lo, hi = round(approx - spread, digits), round(approx + spread, digits)
    return lo if lo == hi else round(statistics.mean(values), digits)


class Database(object):
    """Database object (reference: core.py:44-3341)."""

    LEVELS = LEVELS

    def __init__(self, name: str, folder: str, tmp: str, flat: bool = True, nproc: int = os.cpu_count(), backend=None):
        self.version = __version__
        if not name:
            raise ValueError("Invalid Database name!")
        if not folder:
            raise ValueError("Invalid Database path!")
        self.name = name
        self.root = folder
        os.makedirs(tmp, exist_ok=True)
        self.tmp = tmp
        self.nproc = nproc
        if self.nproc is None or self.nproc < 1 or self.nproc > os.cpu_count():
            self.nproc = os.cpu_count()
        self._backend = backend
        self._touched: List[str] = list()       # taxonomies created or modified by add() (reference: __clusters)
        self._unknowns: List["Entry"] = list()
        self.clusters: Dict[str, Dict[str, "Entry"]] = {level: OrderedDict() for level in LEVELS}
        self.genomes: Dict[str, "Entry"] = OrderedDict()

        if not os.path.isdir(self.root):
            # new database (core.py:134-158)
            self.flat = flat
            os.makedirs(self.root)
            os.makedirs(os.path.join(self.root, "clusters"))
            os.makedirs(os.path.join(self.root, "sketches"))
            self.metadata: Dict[str, Any] = dict()
            self.metadata["clusters_count"] = 0
            self.metadata["flat"] = self.flat
            self.report: Dict[str, Any] = dict()
        else:
            # existing database (core.py:160-250)
            clusters_folder = os.path.join(self.root, "clusters")
            if not os.path.isdir(clusters_folder):
                raise FileNotFoundError(errno.ENOENT, os.strerror(errno.ENOENT), clusters_folder)
            sketches_folder = os.path.join(self.root, "sketches")
            if not os.path.isdir(sketches_folder):
                raise FileNotFoundError(errno.ENOENT, os.strerror(errno.ENOENT), sketches_folder)
            metadata_json_filepath = os.path.join(self.root, "metadata.json")
            if not os.path.isfile(metadata_json_filepath):
                raise FileNotFoundError(errno.ENOENT, os.strerror(errno.ENOENT), metadata_json_filepath)
            with open(metadata_json_filepath) as metadata_json_file:
                self.metadata = json.loads("".join(metadata_json_file.readlines()))
            if not self._validate_metadata(self.metadata):
                raise Exception("Database metadata did not pass the validation!")
            self.flat = self.metadata["flat"]
            report_filepath = os.path.join(self.root, "clusters.tsv")
            if not os.path.isfile(report_filepath):
                raise FileNotFoundError(errno.ENOENT, os.strerror(errno.ENOENT), report_filepath)
            self.report = self._load_report(report_filepath)
            if not self.report:
                raise Exception("Unable to retrieve the database report!")
            for cluster_id, row in self.report.items():
                taxonomy = row["taxonomy"]
                level = row["level"]
                cluster_name = taxonomy.split("|")[-1]
                cluster_folder = os.path.join(clusters_folder, self._get_cluster_batch(cluster_id), cluster_id)
                parent = None if level == "kingdom" else taxonomy.split("|")[-2]
                children = row["references"].union(row["mags"])
                cluster_obj = Entry(self, cluster_id, cluster_name, level, folder=cluster_folder, parent=parent, children=children)
                if not cluster_obj.sketch_filepath:
                    raise Exception(f"No sketch found for cluster {cluster_id}")
                self.clusters[level][cluster_name] = cluster_obj
                if level == "species":
                    for genome_id in sorted(children):
                        genome_obj = Entry(self, genome_id, genome_id, "genome", parent=cluster_name, taxonomy=taxonomy,
                                           known=genome_id in row["references"])
                        if not genome_obj.sketch_filepath:
                            raise Exception(f"No sketch found for genome {genome_id}")
                        self.genomes[genome_id] = genome_obj

    # ------------------------------------------------------------------ backend
    @property
    def backend(self):
        if not self._validate_metadata(self.metadata):
            raise Exception("No database metadata found!")
        key = (int(self.metadata["kmer_size"]), int(self.metadata["filter_size"]))
        if self._backend is None:
            self._backend = get_backend(*key)
        elif _BACKENDS.get(key) is not self._backend:
            _BACKENDS[key] = self._backend  # an injected backend also serves the static Database.dist
        return self._backend

    # ------------------------------------------------------------------ small helpers (core.py:253-368, 3104-3341)
    @staticmethod
    def _get_cluster_batch(cluster_id: str) -> str:
        if not _MSBT_ID.search(cluster_id):
            raise ValueError("The provided argument is not a valid MSBT cluster id!")
        return str(int(int(cluster_id[4:]) / 1000)).zfill(6)

    def __contains__(self, genome: str) -> bool:
        return genome in self.genomes

    def size(self) -> Tuple[int, ...]:
        return tuple(len(self.clusters[level]) for level in LEVELS)

    def taxonomy_exists(self, taxonomy: str) -> bool:
        if not self._is_defined(taxonomy):
            return False
        return all(level_id in self.clusters[level] for level, level_id in zip(LEVELS, taxonomy.split("|")))

    @classmethod
    def _is_defined(cls, taxonomy: str) -> bool:
        if not taxonomy:
            raise ValueError("No taxonomy provided!")
        if taxonomy.count("|") != 6:
            return False
        for level, level_id in zip(cls.LEVELS, taxonomy.split("|")):
            if len(level_id) <= 3 or level_id[:3] != f"{level[0]}__":
                return False
        return True

    @classmethod
    def _format_taxonomy(cls, taxonomy: str) -> str:
        # The reference validates and re-joins the seven prefixed names unchanged (core.py:3137-3199).
        if not cls._is_defined(taxonomy):
            raise ValueError(f"Malformed taxonomic label: {taxonomy}")
        names = [part[3:] for part in taxonomy.split("|")]
        return "|".join(f"{level[0]}__{name}" for level, name in zip(cls.LEVELS, names))

    @staticmethod
    def _is_supported(filepath: str) -> bool:
        if not os.path.isfile(filepath):
            raise FileNotFoundError(errno.ENOENT, os.strerror(errno.ENOENT), filepath)
        return os.path.splitext(filepath)[1] in {".fa", ".fasta", ".fna", ".bf"}

    @staticmethod
    def _validate_metadata(metadata: Dict[str, Any]) -> bool:
        required = {"clusters_count", "filter_size", "kmer_size", "min_kmer_occurrence", "flat"}
        return required.issubset(metadata.keys())

    def _dump_metadata(self) -> None:
        with open(os.path.join(self.root, "metadata.json"), "w+") as metadata_json_file:
            json.dump(self.metadata, metadata_json_file)

    @staticmethod
    def _load_report(filepath: str) -> Dict[str, Any]:
        """clusters.tsv -> {cluster_id: row} (core.py:3233-3321)."""
        if not os.path.isfile(filepath):
            raise FileNotFoundError(errno.ENOENT, os.strerror(errno.ENOENT), filepath)
        table: Dict[str, Any] = OrderedDict()
        header: List[str] = []
        with open(filepath) as handle:
            for line in handle:
                if not line.strip():
                    continue
                if line.startswith("#"):
                    header = line.strip().split("\t")
                    header[0] = header[0][1:].strip()
                    continue
                cells = line.split("\t")
                col = {name: cells[pos] for pos, name in enumerate(header) if pos < len(cells)}
                row: Dict[str, Any] = dict()
                row["level"] = col["level"]
                row["density"] = float(col["density"])
                row["references"] = set(col["references_list"].split(",")) if col["references_list"].strip() else set()
                row["mags"] = set(col["mags_list"].split(",")) if col["mags_list"].strip() else set()
                row["centroid"] = col["centroid"]
                row["known"] = col["known"]
                row["taxonomy"] = col["taxonomy"]
                row["internal"] = col["internal"]
                min_ani = float(col["min_ani"]) if col["min_ani"].strip() else None
                max_ani = float(col["max_ani"].strip()) if col["max_ani"].strip() else None
                row["boundaries"] = (min_ani, max_ani)
                table[col["cluster_id"]] = row
        return table

    # ------------------------------------------------------------------ configuration (core.py:370-533)
    def set_configs(self, filepaths: Iterable[str], min_kmer_occurrence: int = None, kmer_size: int = None,
                    kmer_max: int = None, filter_size: int = None, filter_expand_by: float = None) -> None:
        if os.path.isfile(os.path.join(self.root, "metadata.json")):
            raise Exception("A metadata.json file is already defined for this database!")
        if not filepaths:
            raise ValueError("One or more genome files must be provided in input!")
        if not min_kmer_occurrence:
            raise ValueError("The minimum number of kmer occurrences must be provided in input!")
        self.metadata["min_kmer_occurrence"] = min_kmer_occurrence
        if not kmer_size:
            if not kmer_max:
                raise ValueError("Max kmer size must be provided!")
            # core.py:439-468 runs `kitsune kopt --k-max kmer_max --canonical` here; kopt.py restates its method
            from . import kopt
            searcher = self._backend if self._backend is not None else get_backend(kmer_max, kopt.SET_BITS)
            kmer_size = kopt.optimal_kmer_size(sorted(filepaths), kmer_max, searcher)
        self.metadata["kmer_size"] = kmer_size
        if not filter_size:
            # core.py:470-529 runs ntCard here; the distinct k-mer count comes from K1 + K3' instead (estimate.py)
            from . import estimate
            counter = self._backend if self._backend is not None else get_backend(kmer_size, estimate.EST_BITS)
            f0, _ = estimate.estimate_distinct_kmers(sorted(filepaths), kmer_size, counter.count_union_bits)
            filter_size = estimate.filter_size_from_f0(f0, filter_expand_by)
        self.metadata["filter_size"] = filter_size
        self._dump_metadata()

    # ------------------------------------------------------------------ add (core.py:845-980)
    def add(self, filepath: str, reference: bool = False, taxonomy: Optional[str] = None) -> None:
        self.add_many([(filepath, reference, taxonomy)])

    def add_many(self, items: Sequence[Tuple[str, bool, Optional[str]]]) -> None:
        """add() for a batch: all sketches are built by ONE makebf batch call, then every genome is
        registered exactly as the reference's per-genome loop (metasbt.py:542-543) would."""
        if not self._validate_metadata(self.metadata):
            raise Exception("No database metadata found!")
        prepared = []
        seen = set()
        for filepath, reference, taxonomy in items:
            if not self._is_supported(filepath):
                raise ValueError("Input file format is not supported!")
            if reference and not taxonomy:
                raise ValueError("A full taxonomic label must be defined for reference genomes!")
            if not reference and not self.clusters:
                raise Exception("The database does not contain any clusters!")
            filename = os.path.splitext(os.path.basename(filepath))[0]
            if filename in self.genomes or filename in seen:
                raise Exception("A genome with the same name already exists in the database!")
            seen.add(filename)
            prepared.append((filepath, filename, Entry(self, filename, filename, "genome", taxonomy=taxonomy, known=reference), taxonomy))
        Entry.sketch_many([p[2] for p in prepared], [p[0] for p in prepared])
        unknown = [p for p in prepared if not p[3]]
        if unknown:
            # MAGs with no species close enough (core.py:972-980): their profiles in one batch
            self.profile_many([p[0] for p in unknown], [p[2].sketch_filepath for p in unknown])

        for filepath, filename, genome_obj, taxonomy in prepared:
            if not taxonomy:
                # a MAG with no species close enough (core.py:972-980): keep its profile for characterize()
                genome_obj.profile = self.profile(filepath, genome_obj.sketch_filepath)
                self._unknowns.append(genome_obj)
                continue
            if not self._is_defined(taxonomy):
                raise ValueError("The assigned taxonomy is malformed or not defined!")
            taxonomy = self._format_taxonomy(taxonomy)
            names = taxonomy.split("|")
            new_clusters = False
            # species first, kingdom last: new MSBT ids are handed out in this order (core.py:913-932)
            for position in range(len(names) - 1, -1, -1):
                level, name = LEVELS[position], names[position]
                cluster_obj = self.clusters[level].get(name)
                if cluster_obj is None:
                    if _MSBT_ID.search(name[3:]):
                        cluster_id = name[3:]
                    else:
                        self.metadata["clusters_count"] += 1
                        cluster_id = f"MSBT{self.metadata['clusters_count']}"
                    cluster_folder = os.path.join(self.root, "clusters", self._get_cluster_batch(cluster_id), cluster_id)
                    os.makedirs(cluster_folder)
                    parent = None if position == 0 else names[position - 1]
                    cluster_obj = Entry(self, cluster_id, name, level, folder=cluster_folder, parent=parent)
                    new_clusters = True
                if level == "species":
                    genome_obj.parent = name
                    self.genomes[filename] = genome_obj
                    cluster_obj.add_children(filename)
                else:
                    cluster_obj.add_children(names[position + 1])
                self.clusters[level][name] = cluster_obj
            if new_clusters:
                self._dump_metadata()
            self._touched.append(taxonomy)

    # ------------------------------------------------------------------ update (core.py:1388-1459)
    def update(self) -> None:
        if not self._touched:
            raise Exception("There is nothing to update in the database!")
        if self._unknowns:
            raise Exception("There are still unknown genomes in cache! Please run `self.characterize()` first")
        processed = set()
        for level in reversed(LEVELS):
            pos = LEVELS.index(level)
            # the clusters of one level are independent: their unions go to the device as ONE batch (Entry.index_many); the
            # reference indexes them one after the other, one `bfoperate` spawn each (core.py:1408-1427)
            batch = []
            for taxonomy in self._touched:
                partial = "|".join(taxonomy.split("|")[: pos + 1])
                if partial not in processed:
                    print(f"Update: {partial}")
                    batch.append(self.clusters[level][partial.split("|")[-1]])
                    processed.add(partial)
            Entry.index_many(batch)
        db_id = "MSBT0"
        db_folder = os.path.join(self.root, "clusters", self._get_cluster_batch(db_id), db_id)
        db_obj = Entry(self, db_id, "db", None, folder=db_folder, parent=None, children=set(self.clusters["kingdom"].keys()))
        db_obj.index()
        self._dump_report()
        self._dump_genomes()
        self.report = self._load_report(os.path.join(self.root, "clusters.tsv"))
        self._touched = list()

    def _dump_genomes(self) -> None:
        """genomes.tsv (core.py:1461-1501)."""
        count = len(list(Path(self.root).glob("genomes*.tsv")))
        path = os.path.join(self.root, "genomes.tsv")
        if os.path.isfile(path):
            shutil.move(path, os.path.join(self.root, f"genomes_{count + 1}.tsv"))
        with open(path, "w+") as handle:
            today = datetime.date.today()
            handle.write(f"# MetaSBT genomes {today.year}-{today.month}-{today.day}\n")
            handle.write("# {}\n".format("\t".join(["genome_id", "type", "assignment", "internal"])))
            for genome, obj in self.genomes.items():
                handle.write("{}\n".format("\t".join([
                    genome, "reference" if obj.is_known() else "mag", obj.get_full_taxonomy(), obj.get_full_taxonomy(internal=True)])))

    def _dump_report(self) -> None:
        """clusters.tsv (core.py:1503-1662): species first, kingdom last; density every time (one batched
        popcount call); boundaries recomputed unless the membership is unchanged."""
        count = len(list(Path(self.root).glob("clusters*.tsv")))
        path = os.path.join(self.root, "clusters.tsv")
        if os.path.isfile(path):
            shutil.move(path, os.path.join(self.root, f"clusters_{count + 1}.tsv"))
        # K3': one batched popcount for every cluster representative
        all_clusters = [self.clusters[level][name] for level in reversed(LEVELS) for name in self.clusters[level]]
        densities = Entry.get_density_many(all_clusters)
        # K3: the all-pairs matrices of every species whose boundaries will be recomputed, in one batch (the reference
        # runs a pool of `dist` jobs per cluster, core.py:4036-4081); get_boundaries picks them up by their path tuple
        todo = []
        for cluster_obj in all_clusters:
            if cluster_obj.level != "species":
                continue
            children = sorted(cluster_obj.get_children())
            refs = set(g for g in cluster_obj.children if self.genomes[g].is_known())
            previous = self.report.get(cluster_obj.identifier)
            unchanged = (previous is not None and not previous["references"].difference(refs)
                         and not previous["mags"].difference(cluster_obj.children.difference(refs)))
            if len(children) >= 3 and not unchanged:
                todo.append(tuple(self.genomes[c].sketch_filepath for c in children))
        self._allpairs_prefetch = dict(zip(todo, self.backend.dist_all_pairs_many(todo))) if todo else dict()
        header = ["cluster_id", "level", "density", "references_count", "mags_count", "references_list", "mags_list",
                  "centroid", "known", "taxonomy", "internal", "min_ani", "max_ani"]
        with open(path, "w+") as handle:
            today = datetime.date.today()
            handle.write(f"# MetaSBT report {today.year}-{today.month}-{today.day}\n")
            handle.write("# {}\n".format("\t".join(header)))
            for cluster_obj in all_clusters:
                level = cluster_obj.level
                if level == "species":
                    references = sorted(g for g in cluster_obj.children if self.genomes[g].is_known())
                else:
                    lower = LEVELS[LEVELS.index(level) + 1]
                    references = sorted(c for c in cluster_obj.children if self.clusters[lower][c].is_known())
                mags = sorted(cluster_obj.children.difference(references))
                density = densities[cluster_obj.identifier]
                previous = self.report.get(cluster_obj.identifier)
                if previous is not None and not previous["references"].difference(references) and not previous["mags"].difference(mags):
                    min_ani, max_ani = previous["boundaries"]
                    centroid, known = previous["centroid"], str(previous["known"])
                    taxonomy, internal = previous["taxonomy"], previous["internal"]
                else:
                    is_known = cluster_obj.is_known()
                    taxonomy = cluster_obj.get_full_taxonomy()
                    internal = cluster_obj.get_full_taxonomy(internal=True)
                    min_ani, max_ani, centroid = cluster_obj.get_boundaries()
                    known = str(is_known)
                    # the report is updated immediately: higher levels read the species centroids from it
                    self.report[cluster_obj.identifier] = {
                        "level": level, "density": density, "references": set(references), "mags": set(mags),
                        "centroid": centroid, "known": is_known, "taxonomy": taxonomy, "internal": internal,
                        "boundaries": (min_ani, max_ani)}
                row = [cluster_obj.identifier, level, str(density), str(len(references)), str(len(mags)),
                       ",".join(references), ",".join(mags), centroid, known, taxonomy, internal,
                       "" if min_ani is None else str(min_ani), "" if max_ani is None else str(max_ani)]
                handle.write("{}\n".format("\t".join(row)))
        self._allpairs_prefetch = dict()

    # ------------------------------------------------------------------ dist (core.py:1664-1833)
    @staticmethod
    def dist(sketch_filepath: str, sketches: List[str], kmer_size: int, tmp: str = None, resume: bool = False
             ) -> Tuple[str, Dict[str, float]]:
        """ANI distances between one sketch and a list of sketches.  One AND/OR-popcount pass replaces the
        two `howdesbt bfdistance` spawns; tmp/resume are accepted for signature compatibility (no
        intermediate text tables are written)."""
        if not sketches:
            return (sketch_filepath, dict())
        if not os.path.isfile(sketch_filepath):
            raise FileNotFoundError(errno.ENOENT, os.strerror(errno.ENOENT), sketch_filepath)
        from . import bffile
        header = bffile.read_header(sketch_filepath)
        backend = get_backend(header.kmer_size, header.num_bits)
        distances: Dict[str, float] = OrderedDict()
        for target, (intersect, union) in zip(sketches, backend.dist_one_vs_many(sketch_filepath, sketches)):
            distances[target] = ani_distance(intersect, union, kmer_size)
        return sketch_filepath, distances

    # ------------------------------------------------------------------ profile (core.py:1836-2211)
    @staticmethod
    def _profile(instance: "Database", genome_filepath: str, sketch_filepath: str, uncertainty: float = 50.0,
                 pruning_threshold: float = 0.0):
        return (genome_filepath, instance.profile(genome_filepath, sketch_filepath, uncertainty=uncertainty,
                                                  pruning_threshold=pruning_threshold))

    @staticmethod
    def _tree_leaves(tree_filepath: str) -> List[str]:
        """Leaf filter paths of an `index.detbrief.sbt`: the nodes without children.  Flat trees (core.py:3870-3875) have
        the cluster's own filter as root and one `*`-prefixed line per leaf; clustered trees (flat=False, cluster.py) nest
        deeper.  `howdesbt query` reports leaves only, and with union nodes above them the internal nodes never change a
        leaf's hit count -- so every tree is probed as the flat list of its leaves."""
        from . import cluster
        return cluster.topology_leaves(cluster.read_topology(tree_filepath))

    @staticmethod
    def _leaf_names(leaves: Sequence[str]) -> List[str]:
        """Node names as `howdesbt query` prints them: the `.bf` basename without directory and suffix (App. A6)."""
        return [os.path.splitext(os.path.basename(leaf))[0] for leaf in leaves]

    @staticmethod
    def _tree_matches(names: Sequence[str], hits: Sequence[int], total: int, threshold: float) -> "OrderedDict[str, Tuple[int, int]]":
        """What `howdesbt query --sort --distinctkmers --threshold=T` prints for one query against one flat tree:
        {leaf name: (hits, total)} in output order (hits descending, ties in tree order; App. A6)."""
        needed = math.ceil(threshold * total)
        counts = [int(h) for h in hits]
        order = sorted(range(len(names)), key=lambda i: -counts[i])
        out: "OrderedDict[str, Tuple[int, int]]" = OrderedDict()
        for i in order:
            if counts[i] >= needed:
                out[names[i]] = (counts[i], total)
        return out

    def profile(self, genome_filepath: str, sketch_filepath: str, uncertainty: float = 50.0,
                pruning_threshold: float = 0.0) -> Dict[str, Dict[str, float]]:
        """core.py:1870-2211 for one genome: a batch of one through profile_many."""
        return self.profile_many([genome_filepath], [sketch_filepath], uncertainty=uncertainty,
                                 pruning_threshold=pruning_threshold)[0]

    def _load_profile_table(self, result_filepath: str) -> Optional[Dict[str, Dict[str, float]]]:
        """Memoised table (core.py:1932-1960); a corrupted file is removed and recomputed."""
        if not os.path.isfile(result_filepath):
            return None
        profiles: Dict[str, Dict[str, float]] = {level: OrderedDict() for level in LEVELS + ["genome"]}
        try:
            with open(result_filepath) as handle:
                for line in handle:
                    line = line.strip()
                    if line and not line.startswith("#"):
                        level, label, ani = line.split("\t")[:3]
                        profiles[level][label] = float(ani)
            return profiles
        except Exception:
            os.unlink(result_filepath)
            return None

    def profile_many(self, genome_filepaths: Sequence[str], sketch_filepaths: Sequence[str], uncertainty: float = 50.0,
                     pruning_threshold: float = 0.0) -> List[Dict[str, Dict[str, float]]]:
        """`profile` for a batch of genomes -- the process pool of metasbt.py:1124-1143 (one `_profile` job per genome, each
        spawning >= 7 `howdesbt query` and as many `bfdistance` pairs, core.py:2039-2051, 2106-2198) as batch calls:

        * ONE FASTA ingest + ONE K1 launch + one batched position extraction for all genomes of a batch
          (backend.query_positions_many);
        * the descent db -> kingdom -> ... -> species is level-synchronous: at every level each flat tree is probed ONCE with
          all the genomes that reached it (backend.query_hits_many) and all hit matrices of the level come back in ONE
          device->host copy (backend.fetch_hits) -- the descent only needs the hits;
        * every ANI distance the tables report (Database.dist at core.py:2106, 2133, 2172) is deferred and computed by ONE
          pair-list launch at the end (backend.dist_pairs).

        Results, memoisation (tmp/profiles/<genome>.txt) and exceptions are those of the per-genome method."""
        if len(genome_filepaths) != len(sketch_filepaths):
            raise ValueError("One sketch file path per genome file path is required!")
        levels = ["db"] + LEVELS + ["genome"]
        profiles_dir = os.path.join(self.tmp, "profiles")
        os.makedirs(profiles_dir, exist_ok=True)
        results: List[Optional[Dict[str, Dict[str, float]]]] = [None] * len(genome_filepaths)
        result_paths: List[str] = []
        todo: List[int] = []
        for g, genome_filepath in enumerate(genome_filepaths):
            genome_filename = os.path.splitext(os.path.basename(genome_filepath))[0]
            result_paths.append(os.path.join(profiles_dir, f"{genome_filename}.txt"))
            results[g] = self._load_profile_table(result_paths[g])
            if results[g] is None:
                if not os.path.isfile(sketch_filepaths[g]):
                    raise FileNotFoundError(errno.ENOENT, os.strerror(errno.ENOENT), sketch_filepaths[g])
                todo.append(g)
        if not todo:
            return results  # type: ignore[return-value]

        backend = self.backend
        kmer_size = self.metadata["kmer_size"]
        # batches bounded by the memory their packed positions may take (a file of n bytes has fewer than n k-mers)
        budget = max(int(getattr(backend, "position_bytes", 1 << 62)) // 4, 1)
        batches: List[List[int]] = [[]]
        size = 0
        for g in todo:
            nbytes = os.path.getsize(genome_filepaths[g])
            if batches[-1] and size + nbytes > budget:
                batches.append([])
                size = 0
            batches[-1].append(g)
            size += nbytes

        leaves_of: Dict[str, List[str]] = dict()
        names_of: Dict[str, List[str]] = dict()

        def tree_leaves(cluster_id: str) -> List[str]:
            if cluster_id not in leaves_of:
                tree_filepath = os.path.join(self.root, "clusters", self._get_cluster_batch(cluster_id), cluster_id,
                                             "tree", "index.detbrief.sbt")
                if not os.path.isfile(tree_filepath):
                    raise FileNotFoundError(errno.ENOENT, os.strerror(errno.ENOENT), tree_filepath)
                leaves_of[cluster_id] = self._tree_leaves(tree_filepath)
                names_of[cluster_id] = self._leaf_names(leaves_of[cluster_id])
            return leaves_of[cluster_id]

        cluster_ids: Dict[Tuple[str, str], str] = dict()  # (level, cluster name) -> MSBT id
        # per call, every genome of the batch sees the same database: labels and centroid lists are resolved once
        label_of: Dict[Tuple[str, str], Tuple[str, str]] = dict()            # (level, name) -> (taxonomy, sketch path)
        centroids_of: Dict[Tuple[str, str], Tuple[str, List[str]]] = dict()  # (level, name) -> (taxonomy, centroid sketches)

        def cluster_label(level: str, name: str) -> Tuple[str, str]:
            key = (level, name)
            if key not in label_of:
                obj = self.clusters[level][name]
                label_of[key] = (obj.get_full_taxonomy(), obj.sketch_filepath)
            return label_of[key]

        def cluster_centroids(level: str, name: str) -> Tuple[str, List[str]]:
            key = (level, name)
            if key not in centroids_of:
                obj = self.clusters[level][name]
                species_entries = [obj.name] if level == "species" else sorted(obj.get_children(up_to="species"))
                centroids = [self.report[self.clusters["species"][s].identifier]["centroid"] for s in species_entries]
                centroids_of[key] = (obj.get_full_taxonomy(), [self.genomes[c].sketch_filepath for c in centroids])
            return centroids_of[key]

        for members in batches:
            # The reference joins multi-record FASTA with "N" (core.py:1974-1996) because howdesbt treats records as
            # separate queries; k-mers never span records or Ns, so the distinct-position set of the whole file is the
            # same thing and is computed once for the whole descent.
            qb = backend.query_positions_many([genome_filepaths[g] for g in members])
            totals = backend.query_totals(qb)
            n = len(members)
            reached: List[List[str]] = [["db"] for _ in range(n)]           # clusters each genome queries at this level
            matched: List[Dict[str, "OrderedDict[str, float]"]] = [dict() for _ in range(n)]  # level -> best matches
            for pos, level in enumerate(levels[:-1]):
                # every tree of this level once, with all the genomes that reached it
                by_tree: "OrderedDict[str, List[int]]" = OrderedDict()
                for q in range(n):
                    for cluster in reached[q]:
                        cluster_id = cluster_ids.get((level, cluster))
                        if cluster_id is None:
                            cluster_id = cluster_ids[(level, cluster)] = "MSBT0" if level == "db" else self.clusters[level][cluster].identifier
                        by_tree.setdefault(cluster_id, []).append(q)
                tree_ids = list(by_tree.keys())
                pending = [backend.query_hits_many(qb, by_tree[t], tree_leaves(t)) for t in tree_ids]
                fetched = backend.fetch_hits(pending)  # the level's only synchronisation
                hits_of: Dict[Tuple[str, int], Any] = dict()
                for t, rows in zip(tree_ids, fetched):
                    for r, q in enumerate(by_tree[t]):
                        hits_of[(t, q)] = rows[r]
                threshold = pruning_threshold if level == "kingdom" else 0.0
                for q in range(n):
                    best_level_matches: "OrderedDict[str, float]" = OrderedDict()
                    for cluster in reached[q]:
                        cluster_id = cluster_ids[(level, cluster)]
                        matches = self._tree_matches(names_of[cluster_id], hits_of[(cluster_id, q)], totals[q], threshold)
                        if not matches:
                            continue
                        scores = OrderedDict((node, common / total) for node, (common, total) in matches.items())
                        best_match = sorted(scores.keys(), key=lambda m: scores[m])[-1]
                        best_score = scores[best_match]
                        score_threshold = float((best_score * uncertainty) / 100.0)
                        for node, score in scores.items():
                            if score >= best_score - score_threshold:
                                best_level_matches[node] = score
                    matched[q][level] = best_level_matches
                    reached[q] = list(best_level_matches.keys())

            # every distance the tables report, in one pair-list call (the descent above never needed them)
            pairs: List[Tuple[str, str]] = []
            plan: List[List[Tuple[str, Any]]] = [[] for _ in range(n)]  # per genome: (kind, payload) in table order
            for q, g in enumerate(members):
                sketch_filepath = sketch_filepaths[g]
                wanted: "OrderedDict[str, None]" = OrderedDict()
                first_iter = True
                for pos, level in enumerate(levels[:-1]):
                    next_level = levels[pos + 1]
                    best_level_matches = matched[q][level]
                    if level == "db":
                        for name in best_level_matches:
                            taxonomy_label, cluster_sketch = cluster_label(next_level, name)
                            wanted[cluster_sketch] = None
                            plan[q].append(("single", (next_level, taxonomy_label, cluster_sketch)))
                    elif level == "species":
                        top = sorted(best_level_matches.keys(), key=lambda m: best_level_matches[m])[-1]
                        best_obj = self.genomes[top]
                        wanted[best_obj.sketch_filepath] = None
                        label = f"{best_obj.get_full_taxonomy()}|t__{best_obj.name}"
                        plan[q].append(("single", (next_level, label, best_obj.sketch_filepath)))
                    else:
                        for name in best_level_matches:
                            taxonomy_label, centroid_sketches = cluster_centroids(next_level, name)
                            if first_iter:
                                # the reference computes centroid distances at the first cluster level only and looks the
                                # lower levels up in the same dictionary (core.py:2160-2185)
                                for sk in centroid_sketches:
                                    wanted[sk] = None
                            plan[q].append(("mean", (next_level, taxonomy_label, centroid_sketches)))
                        first_iter = False
                plan[q].insert(0, ("targets", list(wanted.keys())))
                pairs.extend((sketch_filepath, target) for target in wanted)
            counts = backend.dist_pairs(pairs)
            cursor = 0
            for q, g in enumerate(members):
                targets = plan[q][0][1]
                dists: Dict[str, float] = dict()
                for target in targets:
                    intersect, union = counts[cursor]
                    cursor += 1
                    dists[target] = ani_distance(intersect, union, kmer_size)
                profiles: Dict[str, Dict[str, float]] = {level: OrderedDict() for level in levels[1:]}
                for kind, payload in plan[q][1:]:
                    if kind == "single":
                        level, label, target = payload
                        profiles[level][label] = round(dists[target], 5)
                    else:
                        level, label, centroid_sketches = payload
                        profiles[level][label] = mean_rounded([dists[s] for s in centroid_sketches], 5)
                with open(result_paths[g], "w+") as table:
                    table.write("# level\tclosest\tani\n")
                    for level in LEVELS + ["genome"]:
                        for match in profiles[level]:
                            table.write(f"{level}\t{match}\t{profiles[level][match]}\n")
                results[g] = profiles
        return results  # type: ignore[return-value]

    # ------------------------------------------------------------------ outside the hot path
    def _unsupported(self, *a, **k):
        raise NotImplementedError("outside the B200 hot path (SURVEY.md 8: index / profile only)")

    qc = merge = remove = cluster = _unsupported

    # ------------------------------------------------------------------ dereplicate (core.py:2535-2741)
    def dereplicate(self, genomes: Set[str], threshold: float = 0.01, compare_with: str = "self") -> Set[str]:
        """Drop input genomes that have a replica (ANI distance <= threshold) among the input genomes ("self") or in the
        database ("database").  "self": ONE all-pairs AND-popcount launch over the input sketches instead of a pool of
        `dist` jobs (core.py:2611-2640); "database": one profile_many batch with uncertainty 1.0 (core.py:2660-2700), whose
        genome level names the closest genome and its distance.

        The reference cannot get past its last loop once a replica exists (core.py:2735 indexes `replicas`, which is keyed
        by genome NAME, with a file PATH, and :2689 reads the profile with the bare genome name where the key is the full
        label): what is implemented here is what that code sets out to do -- walking the genomes in input order, a genome
        that has not been excluded excludes all of its replicas."""
        if not genomes:
            raise Exception("There are no genomes to dereplicate!")
        if threshold < 0.0 or threshold >= 1.0:
            raise ValueError("The ANI threshold must be >0.0 and <1.0!")
        if compare_with not in ["self", "database"]:
            raise ValueError("The dereplication can be performed against the input itself or the genomes in the database!")
        ordered = sorted(genomes)  # the reference iterates the input set (hash order)
        names = [os.path.splitext(os.path.basename(path))[0] for path in ordered]
        entries = [Entry(self, name, name, "genome") for name in names]
        sketches = Entry.sketch_many(entries, ordered)
        excluded: Set[str] = set()
        if compare_with == "self":
            if len(sketches) > 1:
                inter = self.backend.dist_all_pairs(sketches)
                k = self.metadata["kmer_size"]
                for i, path in enumerate(ordered):
                    if path in excluded:
                        continue
                    for j in range(i + 1, len(ordered)):
                        union = int(inter[i][i]) + int(inter[j][j]) - int(inter[i][j])
                        if ani_distance(int(inter[i][j]), union, k) <= threshold:
                            excluded.add(ordered[j])
        else:
            tables = self.profile_many(ordered, sketches, uncertainty=1.0, pruning_threshold=0.0)
            for path, table in zip(ordered, tables):
                closest = list(table["genome"].keys())[0]
                if table["genome"][closest] <= threshold:
                    excluded.add(path)
        return set(genomes).difference(excluded)

    # ------------------------------------------------------------------ N2: the `update` consumers of K1/K3/K4
    @staticmethod
    def _is_known(instance: "Database", filepath: str) -> Tuple[str, Optional[str]]:
        return (filepath, instance.is_known(filepath))

    def _closest_centroid_check(self, genome_obj: "Entry", level: str) -> Optional[str]:
        """Shared by is_known (core.py:782-841): best match of `level` in the genome's profile, distance to that species'
        centroid (one K3 call), accepted iff within the species' boundary."""
        matches = genome_obj.profile[level]
        best = sorted(matches.keys(), key=lambda match: matches[match])[0]
        if level == "genome":
            best = "|".join(best.split("|")[:-1])  # drop the trailing t__<genome>
        species_name = best.split("|")[-1]
        species_id = self.clusters["species"][species_name].identifier
        centroid = self.report[species_id]["centroid"]
        centroid_sketch = os.path.join(self.root, "sketches", f"{centroid}.bf")
        _, dists = self.dist(genome_obj.sketch_filepath, [centroid_sketch], self.metadata["kmer_size"], tmp=self.tmp, resume=False)
        _, max_boundary = self._estimate_boundaries(self.report[species_id]["taxonomy"])
        return best if dists[centroid_sketch] <= max_boundary else None

    def is_known(self, filepath: str) -> Optional[str]:
        """Can this MAG be assigned to a species of the database?  (core.py:729-843)
        Returns the species-level taxonomy or None.  Run before add()."""
        if not self._validate_metadata(self.metadata):
            raise Exception("No database metadata found!")
        if not self._is_supported(filepath):
            raise ValueError("Input file format is not supported!")
        filename = os.path.splitext(os.path.basename(filepath))[0]
        if filename in self.genomes:
            return self.genomes[filename].get_full_taxonomy()
        if not self.report:
            raise Exception("No database report defined!")
        genome_obj = Entry(self, filename, filename, "genome", taxonomy=None, known=False)
        genome_sketch_filepath = genome_obj.sketch(filepath)
        genome_obj.profile = self.profile(filepath, genome_sketch_filepath)
        # the species of the closest genome first, then the closest species (they can differ)
        for level in ["genome", "species"]:
            taxonomy = self._closest_centroid_check(genome_obj, level)
            if taxonomy is not None:
                return taxonomy
        return None

    def is_known_many(self, filepaths: Sequence[str]) -> Dict[str, Optional[str]]:
        """The pool fan-out of metasbt.py:1878-1897 as one call: every missing sketch is built by ONE makebf batch,
        then each genome is checked as is_known() does."""
        todo = []
        for filepath in filepaths:
            if not self._is_supported(filepath):
                raise ValueError("Input file format is not supported!")
            filename = os.path.splitext(os.path.basename(filepath))[0]
            if filename not in self.genomes:
                todo.append((Entry(self, filename, filename, "genome", taxonomy=None, known=False), filepath))
        if todo and self._validate_metadata(self.metadata):
            sketches = Entry.sketch_many([t[0] for t in todo], [t[1] for t in todo])
            if self.report:
                self.profile_many([t[1] for t in todo], sketches)  # one batch; is_known() below reads the memoised tables
        return OrderedDict((filepath, self.is_known(filepath)) for filepath in filepaths)

    def _estimate_boundaries(self, taxonomy: str) -> Tuple[float, float]:
        """Boundaries of a (possibly not yet existing) cluster (core.py:1295-1386): species are fixed at 5 %; otherwise the
        mean of the boundaries of the clusters of that level under the closest ancestor that has any.  As in the reference,
        the shortcut through the report only fires for labels taxonomy_exists() accepts, i.e. never for a partial label
        (core.py:348 wants seven levels), and a kingdom cannot be estimated."""
        if not taxonomy:
            raise ValueError("Taxonomy is not provided!")
        names = taxonomy.split("|")
        cluster_level = LEVELS[len(names) - 1]
        if cluster_level == "species":
            return (0.0, 0.05)
        if self.taxonomy_exists(taxonomy):
            cluster_id = self.clusters[cluster_level][names[-1]].identifier
            min_ani, max_ani = self.report[cluster_id]["boundaries"]
            if min_ani is not None and max_ani is not None:
                return (min_ani, max_ani)
        if cluster_level == "kingdom":
            raise Exception(f"Unable to retrieve boundaries for {taxonomy}")
        min_bounds: List[float] = list()
        max_bounds: List[float] = list()
        # ancestors from the direct parent up to the kingdom; stop at the first one with usable siblings
        for depth in range(len(names) - 2, -1, -1):
            parent_level, parent_name = LEVELS[depth], names[depth]
            if parent_name in self.clusters[parent_level]:
                for child in self.clusters[parent_level][parent_name].get_children(up_to=cluster_level):
                    child_obj = self.clusters[cluster_level][child]
                    if child_obj.identifier in self.report:
                        min_ani, max_ani = self.report[child_obj.identifier]["boundaries"]
                        if min_ani is not None and max_ani is not None:
                            min_bounds.append(min_ani)
                            max_bounds.append(max_ani)
            if min_bounds or max_bounds:
                break
        if not min_bounds and not max_bounds:
            raise Exception(f"Unable to retrieve boundaries for {taxonomy}")
        return (round(statistics.mean(min_bounds), 5), round(statistics.mean(max_bounds), 5))

    def characterize(self) -> Tuple[Dict[str, Set[str]], Set[str]]:
        """Cluster the unassigned genomes together and define new clusters (core.py:982-1283).

        The all-pairs ANI distances come from ONE K3 all-pairs call (the reference runs a pool of `dist` jobs); the
        dendrogram is scipy's average linkage (the reference calls fastcluster.linkage(method="average"), which
        implements the same algorithm).  One reference behaviour is kept on purpose, so that assignments are identical:
        the dendrogram cut always follows the flat cluster of the FIRST unknown genome -- the reference indexes the
        fcluster result with the inner loop's variable, which is always 0 (core.py:1107 shadows :1080, used at :1131).
        """
        if not self._validate_metadata(self.metadata):
            raise Exception("No database metadata found!")
        if not self._unknowns:
            raise Exception("There are no unknown genomes to be characterized!")
        import scipy.cluster.hierarchy as hier

        sketches = [genome.sketch_filepath for genome in self._unknowns]
        names_of = [os.path.splitext(os.path.basename(path))[0] for path in sketches]
        dendro = None
        if len(sketches) > 1:
            print(f"Computing pair-wise distances between {len(sketches)} uncharacterized genomes")
            inter = self.backend.dist_all_pairs(sketches)
            k = self.metadata["kmer_size"]
            condensed = list()
            for i in range(len(sketches)):
                for j in range(i + 1, len(sketches)):
                    union = int(inter[i][i]) + int(inter[j][j]) - int(inter[i][j])
                    condensed.append(ani_distance(int(inter[i][j]), union, k))
            dendro = hier.linkage(condensed, method="average")

        def cut(height: float):
            return hier.fcluster(dendro, height, criterion="distance") if dendro is not None else [1]

        assignments: Dict[str, str] = dict()
        unassigned: Set[str] = set()
        processed: Set[str] = set()
        print(f"Clustering {len(sketches)} genomes. This may take a while..")
        # genus up to kingdom (the species level was tried by is_known/add)
        for level in reversed(LEVELS[:-1]):
            if len(processed) == len(self._unknowns):
                break
            print(f"Clustering at the {level} level")
            for genome_obj in self._unknowns:
                if genome_obj.sketch_filepath not in processed:
                    matches = genome_obj.profile[level]
                    closest_taxonomy = sorted(matches.keys(), key=lambda match: matches[match])[0]
                    closest_id = self.clusters[level][closest_taxonomy.split("|")[-1]].identifier
                    centroid_sketch = self.genomes[self.report[closest_id]["centroid"]].sketch_filepath
                    _, dists = self.dist(genome_obj.sketch_filepath, [centroid_sketch], self.metadata["kmer_size"],
                                         tmp=self.tmp, resume=False)
                    try:
                        _, max_boundary = self._estimate_boundaries(self.report[closest_id]["taxonomy"])
                    except Exception:
                        max_boundary = None  # outside the kingdom boundaries: stays unassigned
                    if max_boundary is not None and dists[centroid_sketch] <= max_boundary:
                        clusters = cut(max_boundary)
                        cluster_id = clusters[0]  # sic: see the docstring
                        for sketch_filepath, sketch_name, assigned in zip(sketches, names_of, clusters):
                            if assigned == cluster_id and sketch_filepath not in processed:
                                assignments[sketch_name] = closest_taxonomy
                                processed.add(sketch_filepath)
                                print(f"\t[{len(processed)}/{len(self._unknowns)}] {level}={closest_taxonomy}\t{sketch_name}")
                if level == "kingdom" and genome_obj.sketch_filepath not in processed:
                    unassigned.add(genome_obj.name)
        print(f"{len(unassigned)} genomes have not been characterized")

        characterized: Dict[str, Set[str]] = dict()
        partially: "OrderedDict[str, Set[str]]" = OrderedDict()
        partial_genomes: Set[str] = set()
        for sketch_name, taxonomy in assignments.items():
            if taxonomy.count("|") + 1 == len(LEVELS):
                characterized.setdefault(taxonomy, set()).add(sketch_name)
            else:
                partially.setdefault(taxonomy, set()).add(sketch_name)
                partial_genomes.add(sketch_name)
        print(f"{len(partial_genomes)} genomes have been partially characterized")
        if partially:
            print("Processing partially characterized genomes. This will create new clusters..")

        done = 0
        new_clusters = False
        # top-down: every pass extends each partial label by one level with brand-new MSBT clusters, cutting the
        # dendrogram at the boundary estimated for a sibling-less cluster of that level
        while len(partially) > 0:
            for taxonomy in list(partially.keys()):
                print(f"Expanding {taxonomy}")
                next_prefix = LEVELS[taxonomy.count("|") + 1][0]
                _, max_boundary = self._estimate_boundaries(f"{taxonomy}|{next_prefix}__tmp")
                clusters = cut(max_boundary)
                clusters_map: Dict[int, int] = dict()
                for sketch_filepath, sketch_name, assigned in zip(sketches, names_of, clusters):
                    if sketch_name in partially[taxonomy]:
                        if assigned not in clusters_map:
                            self.metadata["clusters_count"] += 1
                            clusters_map[assigned] = self.metadata["clusters_count"]
                            new_clusters = True
                        assigned_taxonomy = f"{taxonomy}|{next_prefix}__MSBT{clusters_map[assigned]}"
                        if self._is_defined(assigned_taxonomy):
                            characterized.setdefault(assigned_taxonomy, set()).add(sketch_name)
                            # the sketch path stands in for the genome path: the sketch exists, so add() reuses it
                            self.add(sketch_filepath, reference=False, taxonomy=assigned_taxonomy)
                            done += 1
                            print(f"\t[{done}/{len(partial_genomes)}] species={assigned_taxonomy}\t{sketch_name}")
                        else:
                            partially.setdefault(assigned_taxonomy, set()).add(sketch_name)
                del partially[taxonomy]

        self._unknowns = list()
        if new_clusters:
            self._dump_metadata()
        return characterized, unassigned


class Entry(object):
    """One cluster (MSBT id) or one genome (reference: core.py:3344-4111)."""

    def __init__(self, database: Database, identifier: str, name: str, level: Optional[str], folder: Optional[str] = None,
                 parent: Optional[str] = None, children: Optional[Set[str]] = None, taxonomy: Optional[str] = None,
                 known: bool = False):
        self.database = database
        if not identifier:
            raise ValueError("Invalid Entry ID!")
        self.identifier = identifier
        if not name:
            raise ValueError("Invalid Entry name!")
        self.name = name
        if level and level not in LEVELS + ["genome"]:
            raise ValueError("Invalid taxonomic level!")
        if level == "genome" and children:
            raise ValueError("A genome cannot contain children!")
        self.level = level
        if self.level == "genome" and self.identifier != self.name:
            raise ValueError("A genome entry ID should match with the entry name!")
        if self.level in LEVELS and not _MSBT_ID.search(self.identifier):
            raise ValueError("A cluster entry ID should match with the following regular expression: ^MSBT[0-9]*$")
        self.folder = folder if folder else os.path.join(os.getcwd(), name)
        is_cluster = not self.level or self.level in LEVELS
        if is_cluster:
            os.makedirs(os.path.join(self.folder, "tree"), exist_ok=True)
        self.parent = parent
        self.children: Set[str] = set(children) if children else set()
        self._known = known
        self.taxonomy = taxonomy
        if self.level == "genome" and self._known and not self.taxonomy:
            raise ValueError("The entry is a reference but no taxonomy is provided!")
        if self.level == "genome" and self._known and not Database._is_defined(self.taxonomy):
            raise ValueError("The provided taxonomy is malformed!")
        self.profile: Dict[str, Any] = dict()
        self.sketch_filepath: Optional[str] = None
        candidate = self._sketch_path()
        if os.path.isfile(candidate):
            self.sketch_filepath = candidate

    def _sketch_path(self) -> str:
        if not self.level or self.level in LEVELS:
            return os.path.join(self.database.root, "clusters", Database._get_cluster_batch(self.identifier),
                                self.identifier, f"{self.name}.bf")
        return os.path.join(self.database.root, "sketches", f"{self.name}.bf")

    def __len__(self) -> int:
        return len(self.children)

    # ------------------------------------------------------------------ navigation (core.py:3509-3712)
    def add_children(self, child: str) -> None:
        if self.level == "genome":
            raise Exception("Genomes cannot have children!")
        self.children.add(child)

    def get_children(self, up_to: str = None) -> Set[str]:
        levels = LEVELS + ["genome"]
        if not up_to:
            up_to = "genome"
        elif up_to not in levels:
            raise Exception(f"{up_to} is not a valid taxonomic level!")
        elif levels.index(up_to) <= levels.index(self.level):
            raise Exception(f"There are no {up_to} levels under the {self.level} level!")
        found: Set[str] = set()
        stack = [self]
        while stack:
            current = stack.pop()
            if current.level == "genome":
                raise Exception("The current entry represents a genome!")
            if current.level == "species":
                found.update(current.children)
                continue
            lower = levels[levels.index(current.level) + 1]
            if lower == up_to:
                found.update(current.children)
                continue
            stack.extend(self.database.clusters[lower][child] for child in current.children)
        return found

    def get_full_taxonomy(self, current_entry: "Entry" = None, taxonomy: str = None, internal: bool = False) -> str:
        entry = current_entry if current_entry else self
        if not entry.level:
            return ""
        if entry.level == "genome":
            entry = self.database.clusters["species"][entry.parent]
        parts: List[str] = [taxonomy] if taxonomy else []
        while True:
            parts.insert(0, f"{entry.level[0]}__{entry.identifier}" if internal else entry.name)
            if entry.level == "kingdom":
                return "|".join(parts)
            entry = self.database.clusters[LEVELS[LEVELS.index(entry.level) - 1]][entry.parent]

    def is_known(self) -> bool:
        if not self.level:
            return False
        if self.level == "genome" or self._known:
            return self._known
        for genome in self.get_children():
            if genome in self.database.genomes and self.database.genomes[genome].is_known():
                self._known = True
                return True
        return False

    # ------------------------------------------------------------------ K1: sketch (core.py:3880-3943)
    def sketch(self, filepath: str) -> str:
        Entry.sketch_many([self], [filepath])
        return self.sketch_filepath

    @staticmethod
    def sketch_many(entries: Sequence["Entry"], filepaths: Sequence[str]) -> List[str]:
        """makebf for many genome entries in one batch; existing sketches are reused (core.py:3915-3918)."""
        todo_fa, todo_out = [], []
        database = None
        for entry, filepath in zip(entries, filepaths):
            database = entry.database
            if not Database._validate_metadata(database.metadata):
                raise Exception("No database metadata found!")
            if entry.level in LEVELS:
                raise Exception("This function can be used in case of genome entries only!")
            if not Database._is_supported(filepath):
                raise ValueError("Input file format is not supported!")
            out = os.path.join(database.root, "sketches", f"{entry.name}.bf")
            if not os.path.isfile(out) and out not in todo_out:
                todo_fa.append(filepath)
                todo_out.append(out)
            entry.sketch_filepath = out
        if todo_fa:
            database.backend.sketch_many(todo_fa, todo_out, database.metadata["min_kmer_occurrence"])
        return [e.sketch_filepath for e in entries]

    # ------------------------------------------------------------------ K2: index (core.py:3714-3878)
    @staticmethod
    def index_many(entries: Sequence["Entry"]) -> None:
        """index() for many independent clusters (one taxonomic level): file lists and flat trees are written per entry, the
        unions run as one batch (backend.union_many).  Clustered (flat=False) databases index entry by entry."""
        if not entries:
            return
        db = entries[0].database
        if not db.flat or not hasattr(db.backend, "union_many"):
            for entry in entries:
                entry.index()
            return
        jobs = []
        for entry in entries:
            ordered, target = entry._prepare_index()
            jobs.append((ordered, target))
        db.backend.union_many(jobs)
        for entry, (ordered, target) in zip(entries, jobs):
            entry._finish_index(ordered, target)

    def _prepare_index(self) -> Tuple[List[str], str]:
        """Checks, `<name>.txt` and the ordered child list of index() (core.py:3729-3776); -> (children, target filter)."""
        db = self.database
        if not Database._validate_metadata(db.metadata):
            raise Exception("No database metadata found!")
        if self.level == "genome":
            raise Exception("Genome entries cannot be indexed!")
        if not self.children:
            raise Exception("This entry is empty!")
        if os.path.isdir(self.folder):
            shutil.rmtree(os.path.join(self.folder, "tree"), ignore_errors=True)
            os.makedirs(os.path.join(self.folder, "tree"), exist_ok=True)
        if self.level == "species":
            sketches = {db.genomes[child].sketch_filepath for child in self.children}
        else:
            lower = "kingdom" if not self.level else LEVELS[LEVELS.index(self.level) + 1]
            sketches = set()
            for child in self.children:
                obj = db.clusters[lower][child]
                sketches.add(os.path.join(obj.folder, f"{obj.name}.bf"))
        ordered = sorted(sketches)  # the reference iterates the set (hash order)
        with open(os.path.join(self.folder, f"{self.name}.txt"), "w+") as handle:
            for path in ordered:
                handle.write(f"{path}\n")
        return ordered, os.path.join(self.folder, f"{self.name}.bf")

    def _finish_index(self, ordered: Sequence[str], target: str) -> None:
        """The flat `index.detbrief.sbt` of index() (core.py:3870-3875)."""
        with open(os.path.join(self.folder, "tree", "index.detbrief.sbt"), "w+") as flat_tree:
            flat_tree.write(f"{target}\n")
            for path in ordered:
                flat_tree.write(f"*{path}\n")
        self.sketch_filepath = target

    def index(self) -> None:
        db = self.database
        if not Database._validate_metadata(db.metadata):
            raise Exception("No database metadata found!")
        if self.level == "genome":
            raise Exception("Genome entries cannot be indexed!")
        if not self.children:
            raise Exception("This entry is empty!")
        if os.path.isdir(self.folder):
            shutil.rmtree(os.path.join(self.folder, "tree"), ignore_errors=True)
            os.makedirs(os.path.join(self.folder, "tree"), exist_ok=True)
        if self.level == "species":
            sketches = {db.genomes[child].sketch_filepath for child in self.children}
        else:
            lower = "kingdom" if not self.level else LEVELS[LEVELS.index(self.level) + 1]
            sketches = set()
            for child in self.children:
                obj = db.clusters[lower][child]
                sketches.add(os.path.join(obj.folder, f"{obj.name}.bf"))
        ordered = sorted(sketches)  # the reference iterates the set (hash order)
        with open(os.path.join(self.folder, f"{self.name}.txt"), "w+") as handle:
            for path in ordered:
                handle.write(f"{path}\n")
        target = os.path.join(self.folder, f"{self.name}.bf")
        if len(ordered) == 1:
            db.backend.copy(ordered[0], target)
        else:
            db.backend.union(ordered, target)
        index_filepath = os.path.join(self.folder, "tree", "index.detbrief.sbt")
        if db.flat:
            with open(index_filepath, "w+") as flat_tree:
                flat_tree.write(f"{target}\n")
                for path in ordered:
                    flat_tree.write(f"*{path}\n")
        else:
            # howdesbt cluster --list --bits=<filter size> --tree=union.sbt --nodename=tree/node{number} --keepallnodes
            # (core.py:3800-3808) followed by howdesbt build (core.py:3822-3828): greedy Hamming agglomeration over
            # K3 / K2 (cluster.py); the internal nodes are plain union filters, not determined/brief RRR files (N4)
            from . import cluster
            topology = cluster.cluster_tree(db.backend, ordered, os.path.join(self.folder, "tree", "node{number}"))
            cluster.write_topology(index_filepath, topology)
        self.sketch_filepath = target

    # ------------------------------------------------------------------ K3': density (core.py:3576-3607)
    def get_density(self) -> float:
        return Entry.get_density_many([self])[self.identifier]

    @staticmethod
    def get_density_many(entries: Sequence["Entry"]) -> Dict[str, float]:
        if not entries:
            return {}
        db = entries[0].database
        paths = [os.path.join(e.folder, f"{e.name}.bf") for e in entries]
        counts = db.backend.popcounts(paths)
        bits = db.metadata["filter_size"]
        return {e.identifier: c / bits for e, c in zip(entries, counts)}

    # ------------------------------------------------------------------ K3: boundaries (core.py:3945-4111)
    def get_boundaries(self, limit_number: int = 0, limit_percentage: float = 100.0):
        db = self.database
        if self.level == "genome":
            raise Exception("Cannot compute boundaries over a genome entry!")
        if self.level == "species":
            children = sorted(self.get_children())
        else:
            children = []
            for species in sorted(self.get_children(up_to="species")):
                identifier = db.clusters["species"][species].identifier
                if identifier in db.report:
                    children.append(db.report[identifier]["centroid"])
                else:
                    children.append(db.clusters["species"][species].get_boundaries(limit_number, limit_percentage)[2])
        if limit_percentage < 100.0:
            limit = math.ceil(len(children) * limit_percentage / 100.0)
        elif limit_number > 0:
            limit = min(limit_number, len(children))
        else:
            limit = len(children)
        if 0 < limit < len(children):
            random.seed(0)
            children = random.sample(children, limit)
        if len(children) < 3:
            return (None, None, sorted(children)[0])

        # all pairs in ONE call: I[i][j] = popc(A_i & A_j), diagonal = popcounts, U = I_ii + I_jj - I_ij
        paths = [db.genomes[c].sketch_filepath for c in children]
        inter = getattr(db, "_allpairs_prefetch", {}).pop(tuple(paths), None)
        if inter is None:
            inter = db.backend.dist_all_pairs(paths)
        k = db.metadata["kmer_size"]
        pairwise: Dict[str, List[float]] = {c: [] for c in children}
        for i, source in enumerate(children):
            for j in range(i + 1, len(children)):
                union = int(inter[i][i]) + int(inter[j][j]) - int(inter[i][j])
                d = ani_distance(int(inter[i][j]), union, k)
                pairwise[source].append(d)
                pairwise[children[j]].append(d)
        best = sys.float_info.max
        centroid = None
        for child in children:
            mean = statistics.mean(pairwise[child])
            if mean < best:
                best, centroid = mean, child
        min_ani = round(min(pairwise[centroid]), 5)
        max_ani = round(max(pairwise[centroid]), 5)
        if self.level != "kingdom" and max_ani == 1.0:
            return (None, None, centroid)
        return (round(min_ani, 5), round(max_ani, 5), centroid)
