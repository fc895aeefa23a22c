#!/usr/bin/env python
"""TEST INFRASTRUCTURE: a stand-in `howdesbt` executable, served by the CPU oracle, that lets the UNMODIFIED reference
(/root/reference/metasbt/core.py) run its own Database/Entry logic in this container (the real binary is not available).
It implements exactly the seven argv shapes MetaSBT spawns (core.py:1747, 1768, 2039, 3588, 3851, 3920) and prints text
the reference's parsers accept (core.py:1787-1831, 2063-2082, 3604-3605).  Used only by tests/golden/make_reference_golden.py."""
import math
import os
import sys

ROOT = os.environ["MSBT_REPO_ROOT"]
sys.path.insert(0, ROOT)
from metasbt_b200 import bffile  # noqa: E402  (file format only)
from oracle import oracle as O  # noqa: E402


def opt(argv, name, default=None):
    for a in argv:
        if a.startswith(name + "="):
            return a[len(name) + 1:]
    return default


def positional(argv):
    return [a for a in argv if not a.startswith("--")]


def load(path):
    return bffile.read_bf(path)


def main():
    cmd, argv = sys.argv[1], sys.argv[2:]
    if cmd == "makebf":
        k, mn, bits = int(opt(argv, "--k")), int(opt(argv, "--min")), int(opt(argv, "--bits"))
        assert opt(argv, "--hashes") == "1" and opt(argv, "--seed") == "0,0"
        fasta = positional(argv)[0]
        words = O.makebf(open(fasta, "rb").read(), k, bits, mn)
        bffile.write_bf(opt(argv, "--out"), bffile.BloomHeader(kmer_size=k, num_bits=bits), words)
    elif cmd == "bfoperate":
        assert "--or" in argv
        paths = [l.strip() for l in open(opt(argv, "--list")) if l.strip()]
        hdr, first = load(paths[0])
        words = O.bf_or([first] + [load(p)[1] for p in paths[1:]])
        bffile.write_bf(opt(argv, "--out"), bffile.BloomHeader(kmer_size=hdr.kmer_size, num_bits=hdr.num_bits), words)
    elif cmd == "bfdistance":
        paths = [l.strip() for l in open(opt(argv, "--list")) if l.strip()]
        _, focus = load(opt(argv, "--focus"))
        inter = "--show:intersect" in argv
        print("#filter1 filter2 " + ("intersection" if inter else "union"))
        for p in paths:
            _, t = load(p)
            c = O.bf_and_popcount(focus, t) if inter else O.bf_or_popcount(focus, t)
            print("%s:%s %d %s" % (p, opt(argv, "--focus"), c, "(intersection)" if inter else "(union)"))
    elif cmd == "dumpbf":
        path = positional(argv)[0]
        hdr, words = load(path)
        print("%s density: %r" % (path, O.bf_popcount(words) / hdr.num_bits))
    elif cmd == "query":
        tree = opt(argv, "--tree")
        threshold = float(opt(argv, "--threshold", "0.0"))
        fasta = positional(argv)[0]
        leaves = [l.strip().lstrip("*") for l in open(tree) if l.strip().startswith("*")]
        hdr = bffile.read_header(leaves[0])
        filters = [load(p)[1] for p in leaves]
        text = open(fasta, "rb").read()
        # one query per FASTA record (App. A6); MetaSBT hands over exactly one record
        records = [b">" + r for r in text.split(b">") if r.strip()]
        for rec in records:
            name = rec[1:].split(b"\n", 1)[0].split()[0].decode() if rec[1:].strip() else "query"
            pos = O.positions_distinct(rec, hdr.kmer_size, hdr.num_bits)
            hits = O.query_hits(pos, filters)
            needed = math.ceil(threshold * len(pos))
            order = sorted(range(len(leaves)), key=lambda i: -int(hits[i]))
            shown = [i for i in order if int(hits[i]) >= needed]
            print("*%s %d" % (name, len(shown)))
            for i in shown:
                node = os.path.splitext(os.path.basename(leaves[i]))[0]
                print("%s %d/%d %.6f" % (node, int(hits[i]), len(pos), int(hits[i]) / max(len(pos), 1)))
    else:
        sys.exit("fake howdesbt: unsupported command %s" % cmd)


if __name__ == "__main__":
    main()
