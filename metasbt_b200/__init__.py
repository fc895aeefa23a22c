"""metasbt_b200 -- B200-native (sm_100a) implementation of MetaSBT's data-parallel hot path.

    ops      host operators over libmsbt.so (C-ABI in include/msbt.h), one per `howdesbt` subcommand
    backend  file-level operators with an HBM-resident filter cache
    core     Database / Entry (the metasbt.core API reached by `index` and `profile`)
    cli      `metasbt index | profile | sketch`
    shard    multi-GPU partitioning (genome blocks, bit ranges) and merges
    bffile   HowDeSBT `.bf` reader / writer

Nothing here computes on the CPU: without libmsbt.so and a CUDA device the operators raise.
"""
__version__ = "0.1.5-b200"
