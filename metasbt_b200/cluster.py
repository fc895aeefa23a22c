"""`howdesbt cluster` (core.py:3800-3808, only reached with flat=False): the topology of a binary Sequence Bloom Tree by
greedy agglomerative clustering of the leaf filters, over the backend's batch operators.

Algorithm (HowDeSBT's cluster command, restated from its published description -- the binary is not
available here, so node numbering and tie-breaking are UNPINNED and isolated in this file):

  1. Hamming distance between every pair of leaves over the first `--bits` bits (MetaSBT passes the whole filter,
     core.py:3804): H(a, b) = popc(a) + popc(b) - 2 popc(a & b)  -- ONE all-pairs AND-popcount launch (K3);
  2. repeat n - 1 times: take the closest pair of active nodes (ties: smaller node ids first; leaves are 0 .. n-1 in list
     order, new nodes follow in creation order), replace it by a new node whose filter is the OR of the two (K2), and
     compute the distances from the new node to every other active node: H(x, w) = popc(x | w) - popc(x & w)
     -- one 1-vs-N launch per merge (K3);
  3. the topology is written in pre-order, one node per line, depth encoded by leading `*` (the format of the flat writer
     core.py:3870-3875 and of `howdesbt query --tree`); internal nodes are named by the `--nodename` template with
     {number} = 1, 2, ... in pre-order (the root is node1), `--keepallnodes`: no culling.

`build --howde` (core.py:3822-3828) turns the union tree into determined/brief, RRR-compressed node files; that
representation is outside this path.  Here the internal nodes are kept as plain union filters
(`node<N>.bf`, the same simple uncompressed format as every other filter), which is all a query of the tree needs: with
union nodes a leaf's hit count never depends on the internal nodes, they only prune.
"""

from __future__ import annotations

import heapq
import os
from typing import Callable, List, Sequence, Tuple


def greedy_merges(hamming: Sequence[Sequence[int]], merge: Callable[[int, int, int], None],
                  distances_to: Callable[[int, List[int]], List[int]]) -> List[Tuple[int, int]]:
    """The agglomeration proper.  hamming[i][j]: leaf distances; merge(u, v, w): create node w = u | v;
    distances_to(w, active): Hamming distances from w to each active node.  -> [(u, v)] for w = n, n + 1, ..."""
    n = len(hamming)
    heap = [(int(hamming[u][v]), u, v) for u in range(n) for v in range(u + 1, n)]
    heapq.heapify(heap)
    active = set(range(n))
    merges: List[Tuple[int, int]] = []
    for w in range(n, 2 * n - 1):
        while True:
            _d, u, v = heapq.heappop(heap)
            if u in active and v in active:
                break
        merge(u, v, w)
        active.discard(u)
        active.discard(v)
        others = sorted(active)
        for x, d in zip(others, distances_to(w, others)):
            heapq.heappush(heap, (int(d), x, w))
        active.add(w)
        merges.append((u, v))
    return merges


def preorder(n_leaves: int, merges: Sequence[Tuple[int, int]]) -> List[Tuple[int, int]]:
    """[(depth, node id)] of the binary tree in pre-order (first child = the lower-numbered node of the merged pair)."""
    if n_leaves == 0:
        return []
    children = {n_leaves + i: uv for i, uv in enumerate(merges)}
    root = n_leaves + len(merges) - 1 if merges else 0
    out: List[Tuple[int, int]] = []
    stack = [(0, root)]
    while stack:
        depth, node = stack.pop()
        out.append((depth, node))
        if node in children:
            u, v = children[node]
            stack.append((depth + 1, v))
            stack.append((depth + 1, u))
    return out


def cluster_tree(backend, leaf_paths: Sequence[str], node_template: str) -> List[Tuple[int, str]]:
    """howdesbt cluster --list=<leaves> --bits=<all> --nodename=<template> --keepallnodes, plus the union node files a
    plain (non-RRR) `build` would produce.  -> [(depth, filter path)] in pre-order.  Internal node filters are written
    to node_template.format-style paths ("{number}" replaced) as they are created."""
    if "{number}" not in node_template:
        raise ValueError("the node name template must contain {number}")
    n = len(leaf_paths)
    if n == 0:
        return []
    if n == 1:
        return [(0, leaf_paths[0])]
    inter = backend.dist_all_pairs(leaf_paths)
    hamming = [[int(inter[i][i]) + int(inter[j][j]) - 2 * int(inter[i][j]) for j in range(n)] for i in range(n)]
    paths = {i: p for i, p in enumerate(leaf_paths)}
    tmp_names = {}

    def merge(u: int, v: int, w: int) -> None:
        # final numbers are only known once the tree is complete: build under a creation-order name, rename afterwards
        tmp = node_template.replace("{number}", "tmp%d" % w) + ".bf"
        backend.union([paths[u], paths[v]], tmp)
        paths[w] = tmp
        tmp_names[w] = tmp

    def distances_to(w: int, others: List[int]) -> List[int]:
        return [union - intersect for intersect, union in backend.dist_one_vs_many(paths[w], [paths[x] for x in others])]

    merges = greedy_merges(hamming, merge, distances_to)
    order = preorder(n, merges)
    topology: List[Tuple[int, str]] = []
    number = 0
    for depth, node in order:
        if node < n:
            topology.append((depth, leaf_paths[node]))
            continue
        number += 1
        final = node_template.replace("{number}", str(number)) + ".bf"
        os.replace(tmp_names[node], final)
        if hasattr(backend, "forget"):
            backend.forget(tmp_names[node])
        topology.append((depth, final))
    return topology


def write_topology(path: str, topology: Sequence[Tuple[int, str]]) -> None:
    with open(path, "w+") as handle:
        for depth, node in topology:
            handle.write("%s%s\n" % ("*" * depth, node))


def read_topology(path: str) -> List[Tuple[int, str]]:
    out = []
    with open(path) as handle:
        for line in handle:
            line = line.strip()
            if line:
                name = line.lstrip("*")
                out.append((len(line) - len(name), name))
    return out


def topology_leaves(topology: Sequence[Tuple[int, str]]) -> List[str]:
    """Nodes without children, in file order (a node is a leaf iff the next line is not deeper)."""
    return [node for i, (depth, node) in enumerate(topology) if i + 1 == len(topology) or topology[i + 1][0] <= depth]
