# resident blocks per SM of the persistent kernel on the eight strong-scaling shards (6 = default)
for bps in 6 5 4; do echo "HSDDP_BLOCKS_PER_SM=$bps"; HSDDP_BLOCKS_PER_SM=$bps python tools/strong_shards.py 2048 1 | cut -c1-75; done
