# the eight strong-scaling shards on one GPU: contiguous index ranges vs interleaved (problem i on rank i mod 8)
echo contiguous; python tools/strong_shards.py 2048 1
echo interleaved; python tools/strong_shards.py 2048 1 interleaved
